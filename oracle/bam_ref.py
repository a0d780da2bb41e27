"""CPU restatement of nimble-aligner's BAM-mode driver — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Independent of the product's C++ feeder: its own BAM decoder (Python gzip reads BGZF as concatenated gzip members) and
literal restatements of (paths under /root/reference)
  src/parse/sorted_bam_reader.rs:31-185   SortedBamReader
  src/parse/bam.rs:100-287                UMIReader (grouping, TSO clip, the 38 metadata fields)
  src/process/bam.rs:157-180              producer loop (the last group is dropped when >= 2 groups exist)
  src/process/bam.rs:305-405              align_umi_to_libraries (zero rows, reasons by read_key)
Parity status: UNPINNED — the reference's only BAM test (tests/bam_pipeline_run.rs) asserts nothing and its inputs are
git-LFS stubs; rust-htslib semantics (aux lookup by the first two tag bytes, read_pair_orientation) are recalled.
"""
import gzip
import struct

import numpy as np

from . import REASONS, Oracle  # noqa: F401

FIELDS = ["QNAME", "QUAL", "REVERSE", "MATE_REVERSE", "PAIRED", "PROPER_PAIRED", "PAIR_ORIENTATION", "UNMAPPED", "MATE_UNMAPPED",
          "FIRST_IN_TEMPLATE", "LAST_IN_TEMPLATE", "STRAND", "MAPQ", "POS", "MATE_POS", "SEQ", "SEQ_LEN", "INSERT_SIZE", "QUALITY_FAILED",
          "SECONDARY", "DUPLICATE", "SUPPLEMENTARY", "NH", "HI", "AS", "GN", "TX", "AN", "nM", "fx", "RE", "CR", "CY", "CB", "UR", "UY", "UB",
          "SKIP_ALIGN"]   # src/parse/bam.rs:9-49
CLIP = 13
_SEQ = "=ACMGRSVTWYHKDBN"


class Record:
    def __init__(self, buf):
        (self.refid, self.pos, l_name, self.mapq, _bin, n_cig, self.flag, l_seq, self.mrefid, self.mpos, self.tlen) = struct.unpack_from("<iiBBHHHIiii", buf, 0)
        o = 32
        self.qname = buf[o:o + l_name - 1].decode()
        o += l_name + 4 * n_cig
        packed = buf[o:o + (l_seq + 1) // 2]
        self.seq = "".join(_SEQ[(packed[i >> 1] >> (4 if (i & 1) == 0 else 0)) & 15] for i in range(l_seq))
        o += (l_seq + 1) // 2
        self.qual = bytes(buf[o:o + l_seq])
        o += l_seq
        self.aux = {}
        while o + 3 <= len(buf):
            tag = buf[o:o + 2].decode("latin1"); ty = chr(buf[o + 2]); o += 3
            if ty in "AcC":
                val = buf[o:o + 1]; o += 1
            elif ty in "sS":
                val = buf[o:o + 2]; o += 2
            elif ty in "iIf":
                val = buf[o:o + 4]; o += 4
            elif ty in "ZH":
                e = buf.index(b"\0", o); val = buf[o:e].decode("latin1"); o = e + 1
            elif ty == "B":
                st = chr(buf[o]); n = struct.unpack_from("<I", buf, o + 1)[0]; w = 1 if st in "cC" else 2 if st in "sS" else 4
                val = buf[o:o + 5 + w * n]; o += 5 + w * n
            else:
                break
            self.aux.setdefault(tag, (ty, val))   # bam_aux_get returns the first occurrence
        self.skip_align = None

    def aux_string(self, field):   # record.aux(field.as_bytes()) -> Aux::String only; htslib matches the first two bytes
        if field == "SKIP_ALIGN" and self.skip_align is not None:
            return self.skip_align
        t = self.aux.get(field[:2])
        return t[1] if t is not None and t[0] == "Z" else None

    def clone(self):
        c = Record.__new__(Record)
        c.__dict__.update(self.__dict__)
        return c

    is_paired = property(lambda s: bool(s.flag & 1))
    is_reverse = property(lambda s: bool(s.flag & 16))
    is_first = property(lambda s: bool(s.flag & 64))


def read_bam(path):
    with gzip.open(path, "rb") as f:
        d = f.read()
    assert d[:4] == b"BAM\1"
    l_text = struct.unpack_from("<I", d, 4)[0]
    p = 8 + l_text
    n_ref = struct.unpack_from("<I", d, p)[0]; p += 4
    for _ in range(n_ref):
        l = struct.unpack_from("<I", d, p)[0]; p += 4 + l + 4
    recs = []
    while p + 4 <= len(d):
        bs = struct.unpack_from("<I", d, p)[0]
        recs.append(Record(d[p + 4:p + 4 + bs])); p += 4 + bs
    return recs


def _dna(s):   # DnaString::from_acgt_bytes(...).to_string()
    return "".join(c.upper() if c in "ACGTacgt" else "A" for c in s)


def _revcomp(s):
    return s[::-1].translate(str.maketrans("ACGT", "TGCA"))


class SortedBamReader:   # src/parse/sorted_bam_reader.rs
    def __init__(self, records, force_bam_paired):
        self.it = iter(records); self.force = force_bam_paired
        self.current_umi = ""; self.next_umi = ""; self.buffer = []; self.next_records = []

    @staticmethod
    def umi_of(r):
        u = r.aux_string("UB")
        if u is None:
            u = r.aux_string("UR")
        if u is None:
            raise RuntimeError("Error -- Could not read UMI.")
        return u

    def fill_buffer(self):
        self.buffer = self.next_records; self.next_records = []
        self.current_umi = self.next_umi
        for r in self.it:
            if not r.is_paired and self.force:
                continue
            if r.aux_string("CB") is None:
                continue
            umi = self.umi_of(r)
            if umi == "AAAAAAAAAA":
                continue
            if self.current_umi == "":
                self.current_umi = umi
            if self.current_umi != umi:
                self.buffer.sort(key=lambda x: x.aux_string("CB"))   # stable, like slice::sort_by
                self.next_records.append(r); self.next_umi = umi
                return
            self.buffer.append(r)

    def add_dummy(self):
        out = []
        for r in self.buffer:
            m = r.clone(); m.skip_align = "FALSE"; out.append(m)
            if not r.is_paired:
                d = r.clone(); d.skip_align = "TRUE"; out.append(d)
        self.buffer = out

    def filter_paired(self):
        out = []; i = 0; b = self.buffer
        while i < len(b):
            if i + 1 < len(b):
                if b[i].qname == b[i + 1].qname:
                    out += [b[i], b[i + 1]] if b[i].is_first else [b[i + 1], b[i]]
                    i += 2
                else:
                    i += 1
            else:
                break
        self.buffer = out

    def next(self):
        if self.buffer:
            return self.buffer.pop()
        self.fill_buffer()
        if not self.force:
            self.add_dummy()
        self.filter_paired()
        self.buffer.reverse()
        return self.buffer.pop() if self.buffer else None


def _orientation(r):   # rust-htslib Record::read_pair_orientation
    unm, munm, mrev = bool(r.flag & 4), bool(r.flag & 8), bool(r.flag & 32)
    if r.is_paired and not unm and not munm and r.refid == r.mrefid:
        if r.pos == r.mpos:
            return "None"
        if r.is_first:
            p1, p2, f1, f2 = r.pos, r.mpos, not r.is_reverse, not mrev
        else:
            p1, p2, f1, f2 = r.mpos, r.pos, not mrev, not r.is_reverse
        if p1 < p2:
            return ("F1" if f1 else "R1") + ("F2" if f2 else "R2")
        return ("F2" if f2 else "R2") + ("F1" if f1 else "R1")
    return "None"


def record_fields(r):   # src/parse/bam.rs:186-236 -> (clipped DnaString as text, 38 fields)
    rev = r.is_reverse
    s, q = r.seq, r.qual
    if len(s) == 124:
        s, q = (s[:-CLIP], q[:-CLIP]) if rev else (s[CLIP:], q[CLIP:])
    seq = _dna(s)
    qual = q[::-1] if rev else q
    b = lambda x: "true" if x else "false"
    fl = r.flag
    named = {"QNAME": r.qname, "QUAL": qual.decode("latin1"), "REVERSE": b(rev), "MATE_REVERSE": b(fl & 32), "PAIRED": b(fl & 1), "PROPER_PAIRED": b(fl & 2),
             "PAIR_ORIENTATION": _orientation(r), "UNMAPPED": b(fl & 4), "MATE_UNMAPPED": b(fl & 8), "FIRST_IN_TEMPLATE": b(fl & 64),
             "LAST_IN_TEMPLATE": b(fl & 128), "STRAND": "-" if rev else "+", "MAPQ": str(r.mapq), "POS": str(r.pos), "MATE_POS": str(r.mpos), "SEQ": seq,
             "SEQ_LEN": str(len(r.seq)), "INSERT_SIZE": str(r.tlen), "QUALITY_FAILED": b(fl & 512), "SECONDARY": b(fl & 256), "DUPLICATE": b(fl & 1024),
             "SUPPLEMENTARY": b(fl & 2048)}
    out = []
    for f in FIELDS:
        z = r.aux_string(f)
        out.append(z if z is not None else named.get(f, ""))
    return seq, qual, out


class UMIReader:   # src/parse/bam.rs:70-253
    def __init__(self, records, force_bam_paired):
        self.rd = SortedBamReader(records, force_bam_paired)
        self.cur = []; self.nxt = []; self.cur_key = ""; self.nxt_key = ""

    def next(self):   # -> final_umi
        self.cur, self.nxt = self.nxt, []
        self.cur_key, self.nxt_key = self.nxt_key, ""
        while True:
            r = self.rd.next()
            if r is None:
                return True
            umi = SortedBamReader.umi_of(r)
            cb = r.aux_string("CB")
            if cb is None:
                raise RuntimeError("Error Read without cell barcode, cannot excise read-mate.")
            key = umi + cb[:-2]
            if self.cur_key == "":
                self.cur_key = key
            seq, qual, fields = record_fields(r)
            item = dict(seq=seq, qual=qual, f=fields, umi=umi, cb=cb[:-2])
            if self.cur_key == key:
                self.cur.append(item)
            else:
                self.nxt.append(item); self.nxt_key = key
                return False


def groups_of(records, force_bam_paired=False):
    """The groups process::bam::process actually sends to the aligner (src/process/bam.rs:157-180)."""
    rd = UMIReader(records, force_bam_paired)
    out = []; has_aligned = False
    while True:
        final = rd.next()
        if final and has_aligned:
            break
        out.append(list(rd.cur))
        has_aligned = True
        if final:
            break
    return out


def data_values(f):
    return "\t".join(v for i, v in enumerate(f) if i not in (1, 15))


def header_line():
    dh = lambda p: "\t".join("%s_%s" % (p, n) for i, n in enumerate(FIELDS) if i not in (1, 15))
    return ("nimble_features\tnimble_score\t%s\t%s\t" % (dh("r1"), dh("r2")) +
            "r1_filter_forward\tr1_forward_score\tr1_filter_reverse\tr1_reverse_score\tr2_filter_forward\tr2_forward_score\tr2_filter_reverse\tr2_reverse_score\ttriage_reason\taligndirection")


def align_groups(oracle, groups, threads=4):
    """align_umi_to_libraries for one library over all groups.  Returns per group: dict(rows=[(features, count)],
    n_zero_rows, pairs=[dict(qname, fr1, fr2, triage, score1, score2, callset)])."""
    r1, r2, q1, q2, s1, s2 = [], [], [], [], [], []
    scope_off = [0]
    for g in groups:
        for j in range(0, len(g) - 1, 2):
            a, b = g[j], g[j + 1]
            rv = lambda x: _revcomp(x["seq"]) if x["f"][2] == "true" else x["seq"]   # reverse_comp_if_needed (src/process/bam.rs:407-415)
            r1.append(rv(a)); r2.append(rv(b)); q1.append(a["qual"]); q2.append(b["qual"])
            s1.append(1 if a["f"][37] == "TRUE" else 0); s2.append(1 if b["f"][37] == "TRUE" else 0)
        scope_off.append(len(r1))
    from . import pack_reads
    d1, o1 = pack_reads(r1); d2, o2 = pack_reads(r2)
    qa = np.frombuffer(b"".join(q1), dtype=np.uint8).copy() if q1 else np.zeros(1, np.uint8)
    qb = np.frombuffer(b"".join(q2), dtype=np.uint8).copy() if q2 else np.zeros(1, np.uint8)
    res = oracle.run(d1, o1, d2, o2, q1=qa, q2=qb, skip1=np.array(s1, dtype=np.uint8), skip2=np.array(s2, dtype=np.uint8),
                     scope_off=np.array(scope_off, dtype=np.uint64), threads=threads)
    out = []
    for gi, g in enumerate(groups):
        rows = res["scopes"][gi]
        p0, p1 = scope_off[gi], scope_off[gi + 1]
        pairs = []
        for p in range(p0, p1):
            cs = rows[res["pair_callset"][p]][0] if res["pair_counted"][p] else None
            pairs.append(dict(qname=g[2 * (p - p0)]["f"][0], fr1=int(res["pair_fr1"][p]), fr2=int(res["pair_fr2"][p]), triage=int(res["pair_triage"][p]),
                              score1=int(res["read_score"][2 * p]) if res["read_pass"][2 * p] else 0,
                              score2=int(res["read_score"][2 * p + 1]) if res["read_pass"][2 * p + 1] else 0, callset=cs))
        # zero rows: one per pair whose mate-slot QNAME is not a representative's QNAME (src/process/bam.rs:332-353); every
        # callset row has its own representative pair, so with unique qnames the count is n_pairs - n_callsets
        n_zero = (p1 - p0 - len(rows)) if rows else 0
        out.append(dict(rows=rows, n_zero_rows=n_zero, pairs=pairs, key=(g[0]["umi"], g[0]["cb"]) if g else None))
    return out
