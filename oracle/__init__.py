"""CPU oracle for the nimble-aligner hot path — TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this package; the product (``nimble_aligner_b200``) never does.  The arithmetic lives in ``oracle.cpp``
(see its header for the reference file:line map and the pinned / unpinned statement); this module restates the
JSON library loader of ``/root/reference/src/reference_library.rs:20-174`` in Python and marshals data.
"""
import ctypes as C
import json
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")

REV_SEP = "§"  # src/reference_library.rs:8

REASONS = [  # src/align.rs:32-51 (enum order) with Display strings 53-77
    "Score Below Threshold", "Discarded Multiple Match", "Discarded Nonzero Mismatch", "No Match",
    "No Match and Score Below Threshold", "Different Filter Reasons", "Required Valid Pair Not Matching",
    "Force Intersect Failure", "Short Read", "Max Hits Exceeded", "Low Entropy", "Successful Match",
    "Strandedness Filtered", "Equivalence Class Empty After Filters", "Above Mismatch Threshold",
    "SKipped Align Due To Unpaired Dummy Read", "None",
]
R = {n: i for i, n in enumerate([
    "ScoreBelowThreshold", "DiscardedMultipleMatch", "DiscardedNonzeroMismatch", "NoMatch", "NoMatchAndScoreBelowThreshold",
    "DifferentFilterReasons", "NotMatchingPair", "ForceIntersectFailure", "ShortRead", "MaxHitsExceeded", "HighEntropy",
    "SuccessfulMatch", "StrandWasWrong", "TriageEmptyEquivalenceClass", "AboveMismatchThreshold",
    "SkippedAlignDueToUnpairedDummy", "None"])}
CHEM = {"unstranded": 0, "fiveprime": 1, "threeprime": 2, "none": 3}  # src/bin/main.rs:40-47


def build(force=False):
    """Compile oracle.cpp with the committed Makefile (building the checker is not using it)."""
    src = os.path.join(_HERE, "oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


class OrcCfg(C.Structure):
    _fields_ = [("score_percent", C.c_double), ("score_threshold", C.c_uint64), ("num_mismatches", C.c_uint64),
                ("discard_multiple_matches", C.c_int32), ("require_valid_pair", C.c_int32),
                ("discard_multi_hits", C.c_uint64), ("max_hits_to_report", C.c_uint64),
                ("intersect_level", C.c_int32), ("strand_filter", C.c_int32),
                ("trim_target_length", C.c_uint64), ("trim_strictness", C.c_double),
                ("group_header_is_nt_sequence", C.c_int32), ("faithful_cost", C.c_int32)]


class OrcInput(C.Structure):
    _fields_ = [("r1", C.c_void_p), ("r1_off", C.c_void_p), ("r2", C.c_void_p), ("r2_off", C.c_void_p),
                ("q1", C.c_void_p), ("q2", C.c_void_p), ("skip1", C.c_void_p), ("skip2", C.c_void_p)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_uint32, C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(OrcCfg)]
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_set_cfg.argtypes = [C.c_void_p, C.POINTER(OrcCfg)]
        L.orc_index_stats.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_index_dump.restype = C.c_uint64
        L.orc_index_dump.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64]
        L.orc_run.restype = C.c_int
        L.orc_run.argtypes = [C.c_void_p, C.POINTER(OrcInput), C.c_uint64, C.c_void_p, C.c_uint64, C.c_int, C.c_int]
        L.orc_work.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_read_count.restype = C.c_uint64
        L.orc_read_count.argtypes = [C.c_void_p]
        L.orc_read_ec_total.restype = C.c_uint64
        L.orc_read_ec_total.argtypes = [C.c_void_p]
        L.orc_read_records.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_pair_records.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_results.restype = C.c_uint64
        L.orc_results.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64]
        L.orc_shannon_entropy.restype = C.c_double
        L.orc_shannon_entropy.argtypes = [C.c_char_p]
        L.orc_maxinfo.restype = C.c_uint64
        L.orc_maxinfo.argtypes = [C.c_char_p, C.c_uint64, C.c_uint64, C.c_double]
        L.orc_natural_lexical_cmp.restype = C.c_int
        L.orc_natural_lexical_cmp.argtypes = [C.c_char_p, C.c_char_p]
        L.orc_filter_pair.restype = C.c_int
        L.orc_filter_pair.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
        L.orc_pseudoalign.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_void_p, C.POINTER(C.c_double)]
        for f in ("orc_feature_list",):
            getattr(L, f).restype = C.c_uint64
        L.orc_feature_list.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_char_p, C.c_uint64]
        L.orc_filter_read_calls.restype = C.c_uint64
        L.orc_filter_read_calls.argtypes = [C.c_char_p, C.c_char_p, C.c_uint64]
        L.orc_filter_chemistry.restype = C.c_uint64
        L.orc_filter_chemistry.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_char_p, C.c_uint64]
        L.orc_parse_calls.restype = C.c_uint64
        L.orc_parse_calls.argtypes = [C.c_char_p, C.c_char_p, C.c_uint64]
        L.orc_intersect.restype = C.c_uint64
        L.orc_intersect.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_uint64]
        L.orc_unmap.restype = C.c_int
        L.orc_unmap.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
        _lib = L
    return _lib


# ---------------------------------------------------------------- reference_library.rs restated
def _revcomp(seq):  # src/utils.rs:61-94
    m = {"a": "t", "c": "g", "t": "a", "g": "c", "u": "a", "A": "T", "C": "G", "T": "A", "G": "C", "U": "A"}
    out = []
    for bp in reversed(seq):
        if bp not in "AaCcGgTtUuNn":
            raise ValueError("Input sequence base is not DNA: %s" % bp)
        out.append(m.get(bp, "N"))
    return "".join(out)


class Reference:  # src/reference_library.rs:11-17
    def __init__(self, group_on, headers, columns, sequence_name_idx, sequence_idx):
        self.group_on, self.headers, self.columns = group_on, headers, columns
        self.sequence_name_idx, self.sequence_idx = sequence_name_idx, sequence_idx


def parse_reference_library(v, strand_filter="unstranded"):
    """v = parsed JSON (list of 2 objects). Returns (config dict, Reference). src/reference_library.rs:20-174."""
    c = v[0]

    def as_int(k):
        if not isinstance(c[k], int) or isinstance(c[k], bool):
            raise ValueError("could not parse %s as int64" % k)
        return c[k]

    def as_bool(k):
        if not isinstance(c[k], bool):
            raise ValueError("could not parse %s as boolean" % k)
        return c[k]

    def as_f64(k):
        if isinstance(c[k], bool) or not isinstance(c[k], (int, float)):
            raise ValueError("could not parse %s as f64" % k)
        return float(c[k])

    cfg = dict(score_percent=as_f64("score_percent"), score_filter=as_int("score_filter"),
               score_threshold=as_int("score_threshold"), num_mismatches=as_int("num_mismatches"),
               discard_multiple_matches=as_bool("discard_multiple_matches"), require_valid_pair=as_bool("require_valid_pair"),
               discard_multi_hits=as_int("discard_multi_hits"), intersect_level=as_int("intersect_level"),
               max_hits_to_report=as_int("max_hits_to_report"), trim_target_length=as_int("trim_target_length"),
               trim_strictness=as_f64("trim_strictness"), strand_filter=strand_filter, discard_nonzero_mismatch=False)
    if cfg["intersect_level"] not in (0, 1, 2):
        raise ValueError("invalid intersect level")
    group_on = c["group_on"]
    if not isinstance(group_on, str):
        raise ValueError("could not parse group_on as string")
    headers = list(v[1]["headers"])
    cols = v[1]["columns"]
    for h in headers:
        if not isinstance(h, str):
            raise ValueError("headers element not a string")
    name_idx = headers.index("sequence_name")
    group_idx = name_idx if group_on == "" else headers.index(group_on)
    seq_idx = headers.index("sequence")
    for col in cols:
        for x in col:
            if not isinstance(x, str):
                raise ValueError("column element not a string")
    n_rows = len(cols[0])
    new_cols = [[] for _ in cols]
    for r in range(n_rows):
        row = [col[r] for col in cols]
        row[seq_idx] = row[seq_idx].replace("U", "T").replace("u", "t")
        rev = list(row)
        rev[name_idx] = rev[name_idx] + REV_SEP + "rev"
        rev[seq_idx] = _revcomp(rev[seq_idx])
        for i in range(len(cols)):
            new_cols[i].append(row[i])
            new_cols[i].append(rev[i])
    cfg["reference_genome_size"] = len(cols[name_idx])
    if not (0.0 <= cfg["score_percent"] <= 1.0):
        raise ValueError("score_percent must be between 0 and 1")
    if cfg["score_filter"] < 0:
        raise ValueError("score_filter must be positive")
    if not (0.0 <= cfg["trim_strictness"] <= 1.0):
        raise ValueError("trim_strictness must be between 0 and 1")
    return cfg, Reference(group_idx, headers, new_cols, name_idx, seq_idx)


def get_reference_library(path, strand_filter="unstranded"):
    with open(path) as f:
        return parse_reference_library(json.load(f), strand_filter)


def _blob(strs):
    return b"".join(s.encode("utf-8") + b"\0" for s in strs)


def _mkcfg(cfg, reference, faithful_cost=False):
    return OrcCfg(cfg["score_percent"], cfg["score_threshold"], cfg["num_mismatches"], int(cfg["discard_multiple_matches"]),
                  int(cfg["require_valid_pair"]), cfg["discard_multi_hits"], cfg["max_hits_to_report"], cfg["intersect_level"],
                  CHEM[cfg["strand_filter"]], cfg["trim_target_length"], cfg["trim_strictness"],
                  int(reference.headers[reference.group_on] == "nt_sequence"), int(faithful_cost))


def pack_reads(reads):
    """list of str/bytes -> (uint8 array, uint64 offsets[n+1])."""
    bs = [r.encode() if isinstance(r, str) else bytes(r) for r in reads]
    off = np.zeros(len(bs) + 1, dtype=np.uint64)
    if bs:
        off[1:] = np.cumsum([len(b) for b in bs], dtype=np.uint64)
    data = np.frombuffer(b"".join(bs), dtype=np.uint8).copy() if bs else np.zeros(0, dtype=np.uint8)
    return data, off


class Oracle:
    """Index + config for one reference library (the PseudoAligner/Reference/AlignFilterConfig triple)."""

    def __init__(self, cfg, reference, faithful_cost=False):
        self.cfg, self.reference, self.faithful_cost = dict(cfg), reference, faithful_cost
        names = reference.columns[reference.sequence_name_idx]
        groups = reference.columns[reference.group_on]
        seqs = reference.columns[reference.sequence_idx]
        self._ccfg = _mkcfg(self.cfg, reference, faithful_cost)
        self.h = lib().orc_create(len(names), _blob(names), _blob(groups), _blob(seqs), C.byref(self._ccfg))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_free(self.h)
            self.h = None

    def set_config(self, **kw):
        self.cfg.update(kw)
        self._ccfg = _mkcfg(self.cfg, self.reference, self.faithful_cost)
        lib().orc_set_cfg(self.h, C.byref(self._ccfg))

    def index_stats(self):
        out = np.zeros(5, dtype=np.uint64)
        lib().orc_index_stats(self.h, out.ctypes.data)
        return dict(zip(["n_kmers", "n_nodes", "n_colours", "colour_elems", "unitig_bases"], out.tolist()))

    def index_dump(self):
        n = lib().orc_index_dump(self.h, None, 0)
        buf = C.create_string_buffer(n)
        lib().orc_index_dump(self.h, buf, n)
        return buf.raw[:n].decode()

    def run(self, r1, r1_off, r2=None, r2_off=None, q1=None, q2=None, skip1=None, skip2=None, scope_off=None,
            threads=1, shard_single=False, want_records=True):
        """get_calls over the given pairs. Arrays are numpy (uint8 data, uint64 offsets). Returns dict."""
        keep = [np.ascontiguousarray(a) if a is not None else None for a in (r1, r1_off, r2, r2_off, q1, q2, skip1, skip2)]
        ptr = [a.ctypes.data if a is not None else None for a in keep]
        inp = OrcInput(*ptr)
        n = len(r1_off) - 1
        so = np.ascontiguousarray(scope_off, dtype=np.uint64) if scope_off is not None else None
        rc = lib().orc_run(self.h, C.byref(inp), n, so.ctypes.data if so is not None else None,
                           (len(so) - 1) if so is not None else 0, threads, int(shard_single))
        if rc != 0:
            raise RuntimeError("oracle: Feature not found in reference columns (reference would panic)")
        out = {}
        w = np.zeros(6, dtype=np.uint64)
        lib().orc_work(self.h, w.ctypes.data)
        out["work"] = dict(zip(["probes", "nodes", "bases", "colour_elems", "reads", "in_bases"], w.tolist()))
        sz = lib().orc_results(self.h, None, 0)
        buf = C.create_string_buffer(max(sz, 1))
        lib().orc_results(self.h, buf, sz)
        scopes = []
        for line in buf.raw[:sz].decode().split("\n"):
            if not line:
                continue
            if line.startswith("#scope "):
                scopes.append([])
                continue
            parts = line.split("\t")
            scopes[-1].append((parts[:-1], int(parts[-1])))
        out["scopes"] = scopes
        if want_records:
            nr = lib().orc_read_count(self.h)
            rec = np.zeros((nr, 5), dtype=np.uint32)
            ec = np.zeros(lib().orc_read_ec_total(self.h), dtype=np.uint32)
            lib().orc_read_records(self.h, rec.ctypes.data, ec.ctypes.data)
            pr = np.zeros((n, 2), dtype=np.uint32)
            lib().orc_pair_records(self.h, pr.ctypes.data)
            out["read_reason"] = (rec[:, 0] & 0xFF).astype(np.uint8)
            out["read_pass"] = ((rec[:, 0] >> 8) & 1).astype(np.uint8)
            out["read_score"] = rec[:, 1].copy()
            out["read_mm"] = rec[:, 2].copy()
            out["read_trimmed_len"] = rec[:, 3].copy()
            out["read_ec_len"] = rec[:, 4].copy()
            out["read_ec"] = ec
            out["pair_triage"] = (pr[:, 0] & 0xFF).astype(np.uint8)
            out["pair_fr1"] = ((pr[:, 0] >> 8) & 0xFF).astype(np.uint8)
            out["pair_fr2"] = ((pr[:, 0] >> 16) & 0xFF).astype(np.uint8)
            out["pair_counted"] = ((pr[:, 0] >> 24) & 0xFF).astype(np.uint8)
            out["pair_callset"] = pr[:, 1].astype(np.int32)
        return out

    def get_calls(self, reads, mates=None, **kw):
        """Convenience for small cases: lists of strings -> sorted [(callset, count)] of the single scope."""
        r1, o1 = pack_reads(reads)
        r2 = o2 = None
        if mates is not None:
            r2, o2 = pack_reads(mates)
        return self.run(r1, o1, r2, o2, **kw)

    def pseudoalign(self, seq, min_len=40):
        out = np.zeros(64, dtype=np.uint32)
        norm = C.c_double(0)
        lib().orc_pseudoalign(self.h, seq.encode(), min_len, out.ctypes.data, C.byref(norm))
        return dict(reason=int(out[0]), passed=bool(out[1]), score=int(out[2]), mm=int(out[3]),
                    ec=out[5:5 + int(out[4])].tolist(), norm=norm.value)

    def feature_list(self, ec, ignore_rollup):
        a = np.asarray(ec, dtype=np.uint32)
        buf = C.create_string_buffer(1 << 16)
        n = lib().orc_feature_list(self.h, a.ctypes.data, len(a), int(ignore_rollup), buf, len(buf))
        return [x for x in buf.raw[:n].decode().split("\n") if x]

    def unmap(self, feats):
        out = np.zeros(max(len(feats), 1), dtype=np.uint32)
        n = lib().orc_unmap(self.h, "\n".join(feats).encode(), out.ctypes.data)
        if n < 0:
            raise KeyError("Feature not found in reference columns")
        return out[:n].tolist()


def _lines(fn, *args):
    buf = C.create_string_buffer(1 << 16)
    n = fn(*args, buf, len(buf))
    return [x for x in buf.raw[:n].decode().split("\n") if x]


def shannon_entropy(s):
    return lib().orc_shannon_entropy(s.encode())


def maxinfo(qual_bytes, target, strictness):
    q = bytes(qual_bytes)
    return lib().orc_maxinfo(q, len(q), target, strictness)


def natural_lexical_cmp(a, b):
    return lib().orc_natural_lexical_cmp(a.encode(), b.encode())


def filter_pair(a, b):
    a = np.asarray(a, dtype=np.uint32)
    b = np.asarray(b, dtype=np.uint32)
    return bool(lib().orc_filter_pair(a.ctypes.data, len(a), b.ctypes.data, len(b)))


def filter_read_calls_with_orientation(calls):
    return _lines(lib().orc_filter_read_calls, "\n".join(calls).encode())


def parse_calls(calls):
    return [(x.split("\t")[0], x.split("\t")[1] == "1") for x in _lines(lib().orc_parse_calls, "\n".join(calls).encode())]


def filter_chemistry(a, b, chem):
    out = _lines(lib().orc_filter_chemistry, "\n".join(a).encode(), "\n".join(b).encode(), CHEM[chem])
    i = out.index("--")
    return out[:i], out[i + 1:]


def intersect(a, b):
    return _lines(lib().orc_intersect, "\n".join(a).encode(), "\n".join(b).encode())


def read_fastq(path):
    """Minimal FASTQ reader for fixtures (4-line records)."""
    seqs, quals = [], []
    with open(path) as f:
        lines = [l.rstrip("\n") for l in f]
    i = 0
    while i + 3 < len(lines) + 1 and i < len(lines):
        if not lines[i].startswith("@"):
            break
        seqs.append(lines[i + 1])
        quals.append(lines[i + 3] if i + 3 < len(lines) else "")
        i += 4
    return seqs, quals
