"""CPU test of the file drivers' inflate (nimble_aligner_b200/csrc/inflate.hpp; flate2 / htslib's place in
/root/reference/src/parse/fastq.rs:21-43 and src/parse/sorted_bam_reader.rs:22-41): whatever zlib produced, at any level
and strategy, must decode to the same bytes — in one piece (a BGZF block) and through small output windows (a .fastq.gz
stream) — and a damaged stream must end in an NbError, never in a crash or in silently different bytes."""
import gzip
import random
import struct
import zlib

import pytest

import nimble_aligner_b200 as nb


def deflate(data, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, wbits=-15, memlevel=8):
    co = zlib.compressobj(level, zlib.DEFLATED, wbits, memlevel, strategy)
    return co.compress(data) + co.flush()


def corpus():
    rng = random.Random(5)
    fastq = "".join("@r%d\n%s\n+\n%s\n" % (i, "".join(rng.choice("ACGT") for _ in range(150)), "".join(rng.choice("FFFFF:,#") for _ in range(150))) for i in range(3000)).encode()
    return {
        "empty": b"",
        "one": b"x",
        "fastq": fastq,
        "zeros": bytes(300_000),                                            # distance 1, length 258 over and over
        "random": bytes(rng.randrange(256) for _ in range(200_000)),        # incompressible: stored blocks at level >= 1
        "period7": bytes(range(7)) * 30_000,                                # overlapping copies with distances below 8
        "far": (bytes(rng.randrange(256) for _ in range(32768)) * 6),       # matches at the maximum distance
        "skewed": bytes(min(255, int(rng.expovariate(0.02))) for _ in range(300_000)),   # long codes: second-level tables
        "text": (b"the quick brown fox jumps over the lazy dog. " * 5000) + fastq[:50_000],
    }


CORPUS = corpus()


@pytest.mark.parametrize("name", sorted(CORPUS))
def test_raw_deflate_in_one_piece(name):
    data = CORPUS[name]
    for level in (0, 1, 4, 6, 9):
        for strategy in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE):
            c = deflate(data, level, strategy)
            assert nb.inflate(c, raw=True, out_cap=len(data)) == data                  # exactly as large as the output (a BGZF block's isize)
            assert nb.inflate(c, raw=True, out_cap=len(data) + 1000) == data
            if len(data) > 10:
                with pytest.raises(nb.NbError):
                    nb.inflate(c, raw=True, out_cap=len(data) - 1)                     # never writes past the buffer
                with pytest.raises(nb.NbError):
                    nb.inflate(c[:len(c) // 2], raw=True, out_cap=len(data))           # cut off
    small = deflate(data, 6, memlevel=1)                                               # many small dynamic blocks
    assert nb.inflate(small, raw=True, out_cap=len(data)) == data


@pytest.mark.parametrize("name", ["fastq", "zeros", "far", "skewed", "random", "period7"])
@pytest.mark.parametrize("window", [0, 300, 777, 4096, 65536, 1 << 20])
def test_gzip_members_through_windows(name, window):
    data = CORPUS[name]
    one = gzip.compress(data, 6)
    assert nb.inflate(one, window=window, out_cap=len(data) + 64) == data
    # several members (bgzip / cat a.gz b.gz), header fields in use, trailing bytes that are no member
    parts = [data[:1000], b"", data[1000:70_000], data[70_000:]]
    hdr = b"\x1f\x8b\x08\x1c" + bytes(6) + struct.pack("<H", 5) + b"extra" + b"name.fq\0" + b"a comment\0"
    many = b"".join(hdr + deflate(p, 3) + struct.pack("<II", zlib.crc32(p), len(p) & 0xFFFFFFFF) for p in parts) + b"\0\0\0\0garbage"
    assert nb.inflate(many, window=window, out_cap=len(data) + 64) == data


def test_damaged_streams_fail_loudly():
    rng = random.Random(9)
    data = CORPUS["fastq"]
    good = gzip.compress(data, 6)
    with pytest.raises(nb.NbError):
        nb.inflate(b"", window=0)
    with pytest.raises(nb.NbError):
        nb.inflate(b"@r1\nACGT\n+\nFFFF\n", window=0)                 # not gzip
    with pytest.raises(nb.NbError):
        nb.inflate(good[:-4], window=4096)                            # trailer cut
    bad_crc = bytearray(good); bad_crc[-6] ^= 1
    with pytest.raises(nb.NbError):
        nb.inflate(bytes(bad_crc), window=4096)
    ok = err = 0
    for it in range(400):
        b = bytearray(good)
        if it % 3 == 0:
            b = b[:rng.randrange(len(b))]
        else:
            for _ in range(rng.randint(1, 4)):
                b[rng.randrange(len(b))] = rng.randrange(256)
        try:
            out = nb.inflate(bytes(b), window=rng.choice([0, 300, 5000]), out_cap=len(data) + 100_000)
            assert out == data                                        # (a flipped byte inside an ignored header field)
            ok += 1
        except nb.NbError:
            err += 1
    assert err > 350
    # raw streams carry no checksum: a damaged one may decode to other bytes, but stays inside the buffers and agrees with zlib on validity
    raw = deflate(data, 6)
    for it in range(400):
        b = bytearray(raw)
        for _ in range(rng.randint(1, 3)):
            b[rng.randrange(len(b))] = rng.randrange(256)
        d = zlib.decompressobj(-15)
        try:
            want = d.decompress(bytes(b), len(data) + 50_000)
            valid = d.eof
        except zlib.error:
            valid = False
        try:
            got = nb.inflate(bytes(b), raw=True, out_cap=len(data) + 50_000)
            assert valid and got == want
        except nb.NbError:
            assert not valid or len(want) > len(data) + 49_000


def test_code_sets_zlib_accepts_or_rejects():
    """Hand-made dynamic blocks: a distance set with one code (incomplete but legal), no distance code at all (literals only),
    an over-subscribed set and a block without an end-of-block code."""
    class Bits:
        def __init__(self):
            self.v = 0; self.n = 0
        def put(self, val, n):              # LSB first (header fields, extra bits)
            self.v |= val << self.n; self.n += n
        def code(self, val, n):             # Huffman codes go MSB first
            for i in range(n - 1, -1, -1):
                self.put((val >> i) & 1, 1)
        def bytes(self):
            return self.v.to_bytes((self.n + 7) // 8, "little")

    def block(litlen_lens, dist_lens, symbols):
        """One final dynamic block whose code-length code gives every length 0..15 a 4-bit code... kept simple: lengths are
        sent as literals through a flat 4-bit code-length code over the symbols 0..15."""
        b = Bits(); b.put(1, 1); b.put(2, 2); b.put(len(litlen_lens) - 257, 5); b.put(len(dist_lens) - 1, 5); b.put(19 - 4, 4)
        order = [16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15]
        for s in order:
            b.put(4 if s < 16 else 0, 3)
        for l in list(litlen_lens) + list(dist_lens):
            b.code(l, 4)                    # canonical flat code: symbol value == code
        def canon(lens):
            codes, code = {}, 0
            for l in range(1, 16):
                for s, sl in enumerate(lens):
                    if sl == l:
                        codes[s] = (code, l); code += 1
                code <<= 1
            return codes
        lc, dc = canon(litlen_lens), canon(dist_lens)
        for kind, sym, extra, nextra in symbols:
            if kind != "raw":
                c, l = (lc if kind == "l" else dc)[sym]
                b.code(c, l)
            if nextra:
                b.put(extra, nextra)
        return b.bytes()

    def both(stream, cap=1000):
        d = zlib.decompressobj(-15)
        try:
            want = d.decompress(stream); ok = d.eof
        except zlib.error:
            ok = False
        try:
            got = nb.inflate(stream, raw=True, out_cap=cap)
            assert ok and got == want
            return got
        except nb.NbError:
            assert not ok
            return None
    # literals 'a' (97), 'b' (98), end-of-block and length symbol 257 (length 3): four 2-bit codes
    ll = [0] * 258; ll[97] = 2; ll[98] = 2; ll[256] = 2; ll[257] = 2
    # one distance code of one bit (distance 1): incomplete, legal
    assert both(block(ll, [1], [("l", 97, 0, 0), ("l", 257, 0, 0), ("d", 0, 0, 0), ("l", 98, 0, 0), ("l", 256, 0, 0)])) == b"aaaab"
    # the unused half of that distance code is an error when it shows up
    assert both(block(ll, [1], [("l", 97, 0, 0), ("l", 257, 0, 0), ("raw", 0, 1, 1), ("l", 256, 0, 0)])) is None
    # no distance code at all: fine as long as only literals are used
    ll3 = [0] * 257; ll3[97] = 1; ll3[256] = 1
    assert both(block(ll3, [0], [("l", 97, 0, 0), ("l", 97, 0, 0), ("l", 256, 0, 0)])) == b"aa"
    # over-subscribed literal set
    bad = [0] * 257; bad[97] = 1; bad[98] = 1; bad[256] = 1
    assert both(block(bad, [1], [])) is None
    # no end-of-block code
    noeob = [0] * 257; noeob[97] = 1; noeob[98] = 1
    assert both(block(noeob, [1], [])) is None
    # incomplete literal set with more than one code
    inc = [0] * 257; inc[97] = 2; inc[98] = 2; inc[256] = 2
    assert both(block(inc, [1], [("l", 97, 0, 0), ("l", 256, 0, 0)])) is None
    # a distance before the start of the output
    assert both(block(ll, [1, 1], [("l", 97, 0, 0), ("l", 257, 0, 0), ("d", 1, 0, 0), ("l", 256, 0, 0)])) is None
    # 15-bit codes: a maximally skewed literal set (lengths 1, 2, ..., 14, 15, 15)
    sk = [0] * 257
    for i in range(14):
        sk[65 + i] = i + 1
    sk[65 + 14] = 15; sk[256] = 15
    syms = [("l", 65 + i, 0, 0) for i in range(15)] * 3 + [("l", 256, 0, 0)]
    assert both(block(sk, [0], syms)) == bytes(range(65, 80)) * 3


# ---------------------------------------------------------------------------------------------- several threads on one gzip file
@pytest.mark.parametrize("name", sorted(CORPUS))
def test_parallel_reader_equals_zlib(name):
    """pgunzip.hpp: workers enter the deflate stream at guessed block headers with the window unknown; the text must be
    zlib's whatever the chunk size (4 KiB chunks: hundreds of entry points, many of them useless) and thread count."""
    data = CORPUS[name]
    for level, strategy, memlevel in ((6, zlib.Z_DEFAULT_STRATEGY, 8), (1, zlib.Z_DEFAULT_STRATEGY, 8), (9, zlib.Z_DEFAULT_STRATEGY, 1), (0, zlib.Z_DEFAULT_STRATEGY, 8),
                                      (6, zlib.Z_FIXED, 8), (6, zlib.Z_HUFFMAN_ONLY, 8)):
        g = deflate(data, level, strategy, wbits=31, memlevel=memlevel)
        assert zlib.decompress(g, 31) == data
        for threads, chunk in ((1, 4096), (2, 4096), (5, 9000), (3, 1 << 20)):
            assert nb.gunzip_parallel(g, threads, chunk, out_cap=len(data) + 64) == data, (level, strategy, threads, chunk)


def test_parallel_reader_members_and_ratios():
    rng = random.Random(21)
    data = CORPUS["fastq"] * 3
    # many members (bgzip: one per 64 KiB), members cut anywhere, empty members, bytes behind the last member
    cuts = sorted(rng.randrange(len(data)) for _ in range(40))
    members = b"".join(gzip.compress(data[a:b], rng.choice([1, 6, 9])) for a, b in zip([0] + cuts, cuts + [len(data)]))
    for threads, chunk in ((4, 4096), (3, 50_000), (2, 1 << 20)):
        assert nb.gunzip_parallel(members, threads, chunk, out_cap=len(data) + 64) == data
        assert nb.gunzip_parallel(members + gzip.compress(b"") + b"\0" * 5000, threads, chunk, out_cap=len(data) + 64) == data
    # bgzip's layout (a BC extra field per member, an empty member at the end): workers enter at member headers
    def bgzf(d):
        out = []
        for k in range(0, len(d), 65280):
            c = d[k:k + 65280]; cd = deflate(c, 6)
            out.append(b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", len(cd) + 25) + cd + struct.pack("<II", zlib.crc32(c) & 0xFFFFFFFF, len(c)))
        return b"".join(out) + bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
    for threads, chunk in ((4, 4096), (4, 100_000)):
        assert nb.gunzip_parallel(bgzf(data), threads, chunk, out_cap=len(data) + 64) == data
    damaged = bytearray(bgzf(data)); damaged[len(damaged) // 2] ^= 0x10
    with pytest.raises(nb.NbError):
        nb.gunzip_parallel(bytes(damaged), 4, 100_000, out_cap=len(data) + 64)
    # text that expands a thousandfold: chunks are cut by their output size, the reader keeps working in bounded pieces
    big = bytes(40_000_000)
    g = gzip.compress(big, 6)
    assert len(g) < 50_000
    for threads, chunk in ((4, 4096), (2, 1 << 20)):
        out = nb.gunzip_parallel(g, threads, chunk, out_cap=len(big) + 64)
        assert len(out) == len(big) and out == big
    mixed = gzip.compress(big[:20_000_000] + data + big[:9_000_000] + data, 6)
    assert nb.gunzip_parallel(mixed, 4, 4096, out_cap=30_000_000 + 2 * len(data)) == big[:20_000_000] + data + big[:9_000_000] + data


def test_parallel_reader_fails_loudly_on_damage():
    rng = random.Random(4)
    data = CORPUS["fastq"] * 2
    good = gzip.compress(data, 6)
    with pytest.raises(nb.NbError):
        nb.gunzip_parallel(b"", 4, 4096)
    with pytest.raises(nb.NbError):
        nb.gunzip_parallel(data[:5000], 4, 4096)
    with pytest.raises(nb.NbError):
        nb.gunzip_parallel(good[:-5], 4, 4096, out_cap=len(data) + 64)
    ok = err = 0
    for it in range(150):
        b = bytearray(good)
        if it % 3 == 0:
            b = b[:rng.randrange(20, len(b))]
        else:
            for _ in range(rng.randint(1, 3)):
                b[rng.randrange(len(b))] = rng.randrange(256)
        try:
            out = nb.gunzip_parallel(bytes(b), rng.choice([2, 4]), rng.choice([4096, 30_000]), out_cap=len(data) + 200_000)
            assert out == data
            ok += 1
        except nb.NbError:
            err += 1
    assert err > 130


def test_inflate_under_sanitizers(tmp_path):
    """tests/native/inflate_harness.cpp built with -fsanitize=address,undefined: corpus files and damaged copies of them through
    the parallel and the serial reader (bounds, overflows, misaligned access; the same harness runs clean under -fsanitize=thread)."""
    import os
    import shutil
    import subprocess
    if not shutil.which("g++"):
        pytest.skip("no g++")
    here = os.path.dirname(os.path.abspath(__file__))
    exe = str(tmp_path / "harness")
    r = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-o", exe,
                        os.path.join(here, "native", "inflate_harness.cpp"), "-lz", "-lpthread"], capture_output=True, text=True)
    if r.returncode != 0 and "sanitize" in r.stderr:
        pytest.skip("sanitizer runtime not installed")
    assert r.returncode == 0, r.stderr[-2000:]
    files = []
    for i, (name, level, strategy) in enumerate([("fastq", 6, zlib.Z_DEFAULT_STRATEGY), ("fastq", 1, zlib.Z_HUFFMAN_ONLY), ("skewed", 9, zlib.Z_DEFAULT_STRATEGY),
                                                  ("far", 6, zlib.Z_DEFAULT_STRATEGY), ("random", 6, zlib.Z_DEFAULT_STRATEGY), ("period7", 6, zlib.Z_FIXED), ("zeros", 6, zlib.Z_DEFAULT_STRATEGY)]):
        p = tmp_path / ("c%d.gz" % i); p.write_bytes(deflate(CORPUS[name], level, strategy, wbits=31)); files.append(str(p))
    data = CORPUS["fastq"]
    p = tmp_path / "members.gz"; p.write_bytes(b"".join(gzip.compress(data[a:a + 70_000], 6) for a in range(0, len(data), 70_000)) + b"\0" * 50); files.append(str(p))
    r = subprocess.run([exe] + files, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "done: 0 problems" in r.stdout and "runtime error" not in r.stderr and "ERROR" not in r.stderr, (r.stdout[-1500:], r.stderr[-3000:])


# ---------------------------------------------------------------------------------------------- the BAM driver's TSV compressor
def test_fast_deflate_round_trips(tmp_path):
    """deflate_fast.hpp (the rows stage's gzip members): whatever goes in must come out of zlib and of inflate.hpp unchanged —
    corpus files, row text as the driver writes it, inputs around the block size (64 Ki tokens), tiny inputs, runs at the
    maximum match length and distance."""
    import synth
    from tests.bamcases import make_bam
    import os
    cases = dict(CORPUS)
    L = synth.SynthLibrary(seed=1234, n_fam=20, n_all=5)
    bam = make_bam(str(tmp_path / "t.bam"), L, n_groups=1500)
    os.environ["NB_BAM_DUMP_ROWFMT"] = "1"
    try:
        nb.bam_dump_groups(bam, str(tmp_path / "rows.txt"), num_cores=2)
    finally:
        del os.environ["NB_BAM_DUMP_ROWFMT"]
    rows = open(tmp_path / "rows.txt", "rb").read()
    assert len(rows) > 1_000_000
    cases["rows"] = rows
    rng = random.Random(12)
    cases["literals_over_a_block"] = bytes(rng.randrange(256) for _ in range(70_000))                 # > 64 Ki literal tokens: two blocks
    cases["exactly_a_block"] = bytes(rng.randrange(256) for _ in range(65_536 + 12))
    cases["one_symbol"] = b"a" * 100_000
    cases["two_symbols"] = bytes(rng.choice(b"ab") for _ in range(50_000))
    cases["max_distance"] = (lambda blk: blk + bytes(rng.randrange(256) for _ in range(32768 - len(blk))) + blk * 3)(bytes(rng.randrange(256) for _ in range(300)))
    for n in range(0, 40):
        cases["tiny%d" % n] = bytes(rng.choice(b"abc") for _ in range(n))
    fib = [1, 1]                                       # Fibonacci literal frequencies: the unlimited Huffman tree is 20 levels deep, the 15-bit limit has to be enforced
    while sum(fib) + fib[-1] + fib[-2] < 60_000:
        fib.append(fib[-1] + fib[-2])
    deep = [sym * 7 % 256 for sym, f in enumerate(fib) for _ in range(f)]
    rng.shuffle(deep)
    cases["deep_tree"] = bytes(deep)
    for name, data in cases.items():
        g = nb.gzip_fast(data)
        assert zlib.decompress(g, 31) == data, name
        assert nb.inflate(g, window=4096 if len(data) else 0, out_cap=len(data) + 64) == data, name
    assert len(nb.gzip_fast(rows)) < len(zlib.compress(rows, 2)) * 1.1              # no worse than the zlib level it replaces
    # the TSV.gz as the driver writes it: a zlib member (the header) followed by the rows' members, read as ONE gzip stream
    parts3 = [b"header line\n", rows[:300_000], rows[300_000:]]
    stream = gzip.compress(parts3[0], 2) + nb.gzip_fast(parts3[1]) + nb.gzip_fast(parts3[2])
    assert gzip.decompress(stream) == b"".join(parts3)
    assert nb.gunzip_parallel(stream, 3, 4096, out_cap=len(rows) + 64) == b"".join(parts3)
    for it in range(300):                                                           # structured random inputs: repeats at all distances and lengths
        parts = []
        for _ in range(rng.randint(1, 60)):
            k = rng.random()
            if k < 0.3 and parts:
                parts.append(rng.choice(parts)[:rng.randint(1, 400)])
            elif k < 0.5:
                parts.append(bytes([rng.randrange(256)]) * rng.randint(1, 600))
            else:
                parts.append(bytes(rng.randrange(256) for _ in range(rng.randint(1, 200))))
        data = b"".join(parts)
        assert zlib.decompress(nb.gzip_fast(data), 31) == data
