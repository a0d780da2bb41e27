"""Minimal BAM (BGZF) writer for synthetic 10x-style inputs.  Test / bench infrastructure only."""
import struct
import zlib

_CODE = {c: i for i, c in enumerate("=ACMGRSVTWYHKDBN")}


def _bgzf_block(data, level=6):
    co = zlib.compressobj(level, zlib.DEFLATED, -15)
    comp = co.compress(data) + co.flush()
    bsize = len(comp) + 25
    hdr = struct.pack("<BBBBIBBHBBHH", 31, 139, 8, 4, 0, 0, 255, 6, 66, 67, 2, bsize)
    return hdr + comp + struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data))


def _pack_seq(seq):
    n = len(seq)
    out = bytearray((n + 1) // 2)
    for i, c in enumerate(seq):
        v = _CODE.get(c.upper(), 15)
        out[i >> 1] |= v << (4 if (i & 1) == 0 else 0)
    return bytes(out)


def encode_record(qname, flag, seq, qual, tags=(), refid=-1, pos=-1, mapq=255, next_refid=-1, next_pos=-1, tlen=0):
    """tags: sequence of (two-letter tag, BAM aux type, value)."""
    name = qname.encode() + b"\0"
    aux = b""
    for tag, ty, val in tags:
        if ty == "Z":
            aux += tag.encode() + b"Z" + str(val).encode() + b"\0"
        elif ty == "A":
            aux += tag.encode() + b"A" + str(val).encode()[:1]
        elif ty == "i":
            aux += tag.encode() + b"i" + struct.pack("<i", int(val))
        elif ty in "cCsSIf":
            aux += tag.encode() + ty.encode() + struct.pack("<" + {"c": "b", "C": "B", "s": "h", "S": "H", "I": "I", "f": "f"}[ty], val)
        elif ty == "H":
            aux += tag.encode() + b"H" + str(val).encode() + b"\0"
        elif ty == "B":      # val = (subtype in 'cCsSiIf', sequence of numbers)
            st, xs = val
            aux += tag.encode() + b"B" + st.encode() + struct.pack("<I", len(xs)) + b"".join(struct.pack("<" + {"c": "b", "C": "B", "s": "h", "S": "H", "i": "i", "I": "I", "f": "f"}[st], x) for x in xs)
        else:
            raise ValueError(ty)
    body = struct.pack("<iiBBHHHIiii", refid, pos, len(name), mapq, 4680, 0, flag, len(seq), next_refid, next_pos, tlen)
    body += name + _pack_seq(seq) + bytes(qual) + aux
    return struct.pack("<I", len(body)) + body


def write_bam(path, records, refs=(("chr1", 100000000),), block_bytes=60000, level=6):
    """records: iterable of already encoded records (encode_record)."""
    text = b"@HD\tVN:1.6\tSO:unsorted\n" + b"".join(b"@SQ\tSN:%s\tLN:%d\n" % (n.encode(), l) for n, l in refs)
    hdr = b"BAM\1" + struct.pack("<I", len(text)) + text + struct.pack("<I", len(refs))
    for n, l in refs:
        nm = n.encode() + b"\0"
        hdr += struct.pack("<I", len(nm)) + nm + struct.pack("<I", l)
    with open(path, "wb") as f:
        buf = bytearray(hdr)
        for r in records:
            buf += r
            while len(buf) >= block_bytes:
                f.write(_bgzf_block(bytes(buf[:block_bytes]), level))
                del buf[:block_bytes]
        if buf:
            f.write(_bgzf_block(bytes(buf), level))
        f.write(_bgzf_block(b"", level))   # EOF marker block
    return path
