"""Seeded differential fuzzing of the whole path against the oracle: random libraries built from a pool of segments
(shared, reverse-complemented, tandem-repeated, low-complexity), random grouping, random AlignFilterConfig, random reads
(fragments with substitutions, N, lowercase, junk tails, lengths 0..260) — every per-read, per-pair and per-scope output
must be bit-identical (tests.test_gpu_parity.compare)."""
import json

import numpy as np
import pytest

import nimble_aligner_b200 as nb
import oracle as orc
import synth
from tests.test_gpu_parity import compare

pytestmark = pytest.mark.gpu
RC = str.maketrans("ACGT", "TGCA")


def _case(seed):
    rng = np.random.default_rng(seed)
    rs = lambda n: "".join("ACGT"[i] for i in rng.integers(0, 4, n))
    rc = lambda s: s[::-1].translate(RC)
    pool = [rs(int(rng.integers(35, 260))) for _ in range(int(rng.integers(4, 14)))]
    pool += [rc(p) for p in pool[:3]] + [("ACG" * 40)[: int(rng.integers(31, 100))], "A" * int(rng.integers(31, 60)), (rs(7) * 30)[: int(rng.integers(40, 150))]]
    n_seq = int(rng.integers(3, 40))
    seqs = []
    for _ in range(n_seq):
        s = "".join(pool[int(rng.integers(len(pool)))] for _ in range(int(rng.integers(1, 6))))
        if rng.random() < 0.3:   # a near-copy of an earlier sequence (allele)
            if seqs:
                s = list(seqs[int(rng.integers(len(seqs)))])
                for pos in rng.integers(0, len(s), size=int(rng.integers(1, 5))):
                    s[pos] = "ACGT"[int(rng.integers(4))]
                s = "".join(s)
        seqs.append(s[: int(rng.integers(40, 900))] if len(s) > 40 else s + rs(40))
    names = ["S%03d" % i for i in range(n_seq)]
    groups = [("g%d" % int(rng.integers(max(1, n_seq // 3)))) if rng.random() < 0.8 else "" for _ in range(n_seq)]
    cfg = dict(synth.BASE_CONFIG)
    cfg.update(group_on="grp" if rng.random() < 0.6 else "", num_mismatches=int(rng.integers(0, 4)), score_percent=float(rng.choice([0.0, 0.2, 0.33, 0.7, 1.0])),
               score_threshold=int(rng.choice([0, 30, 50, 90])), discard_multiple_matches=bool(rng.random() < 0.2), require_valid_pair=bool(rng.random() < 0.3),
               discard_multi_hits=int(rng.choice([0, 0, 1, 2, 5])), max_hits_to_report=int(rng.choice([1, 2, 5, 10, 20])), intersect_level=int(rng.integers(0, 3)))
    obj = [cfg, {"headers": ["sequence_name", "grp", "sequence"], "columns": [names, groups, seqs]}]
    chem = str(rng.choice(["unstranded", "fiveprime", "threeprime", "none"]))
    reads, mates = [], []
    both = seqs + [rc(s) for s in seqs]
    def frag():
        k = rng.random()
        if k < 0.08:
            return rs(int(rng.integers(0, 200)))
        a = both[int(rng.integers(len(both)))]
        n = int(rng.integers(20, 260)); st = int(rng.integers(0, max(1, len(a) - 30)))
        r = list(a[st:st + n])
        for _ in range(int(rng.choice([0, 0, 0, 1, 1, 2, 4]))):
            if r:
                r[int(rng.integers(len(r)))] = "ACGTN"[int(rng.integers(5))]
        r = "".join(r)
        if rng.random() < 0.1:
            r = r + rs(int(rng.integers(1, 40)))
        if rng.random() < 0.1:
            r = rs(int(rng.integers(1, 40))) + r
        if rng.random() < 0.1:
            r = r.lower()
        return r[:300]
    for _ in range(1500):
        r = frag()
        reads.append(r)
        mates.append(reads[int(rng.integers(len(reads)))] if rng.random() < 0.15 else frag())   # some exact duplicates of earlier reads
    paired = bool(rng.random() < 0.75)
    return obj, chem, reads, (mates if paired else None)


@pytest.mark.parametrize("seed", list(range(100, 180)))
def test_random_library_config_and_reads_match_the_oracle(seed):
    obj, chem, reads, mates = _case(seed)
    ocfg, oref = orc.parse_reference_library(obj, chem)
    lib = nb.Library.from_text(json.dumps(obj), chem)
    ix = nb.build_index(lib, 4, device=0 if seed % 2 else None)
    o = orc.Oracle(ocfg, oref)
    assert o.index_dump() == ix.dump()
    ctx = nb.Context(ix, lib)
    r1, o1 = orc.pack_reads(reads)
    r2, o2 = orc.pack_reads(mates) if mates is not None else (None, None)
    compare(ctx, o, ocfg, r1, o1, r2, o2)


@pytest.mark.parametrize("seed", list(range(500, 540)))
def test_random_scoped_batches_with_quals_trim_and_dummy_mates_match_the_oracle(seed):
    """BAM-shaped use of the same path: raw-Phred qualities with random low tails (MAXINFO trim with a random target /
    strictness), SKIP_ALIGN dummies in either slot, consecutive pairs grouped into scopes of random size."""
    obj, chem, reads, mates = _case(seed)
    rng = np.random.default_rng(seed + 7)
    obj[0]["trim_target_length"] = int(rng.choice([0, 15, 40, 90, 200]))
    obj[0]["trim_strictness"] = float(rng.choice([0.0, 0.1, 0.5, 0.9, 1.0]))
    if mates is None:
        mates = [reads[int(rng.integers(len(reads)))] for _ in reads]
    ocfg, oref = orc.parse_reference_library(obj, chem)
    lib = nb.Library.from_text(json.dumps(obj), chem)
    ix = nb.build_index(lib, 4, device=0 if seed % 2 else None)
    o = orc.Oracle(ocfg, oref)
    ctx = nb.Context(ix, lib)
    r1, o1 = orc.pack_reads(reads)
    r2, o2 = orc.pack_reads(mates)
    def quals(off):
        q = rng.integers(20, 42, size=int(off[-1])).astype(np.uint8)
        for i in np.flatnonzero(rng.random(len(off) - 1) < 0.4):
            a, b = int(off[i]), int(off[i + 1])
            if b - a > 5:
                cut = a + int(rng.integers(1, b - a))
                q[cut:b] = rng.integers(0, 8, size=b - cut)
        return q
    q1, q2 = quals(o1), quals(o2)
    n = len(reads)
    k = rng.random(n)
    flags1 = (k < 0.25).astype(np.uint8) * nb.FLAG_SKIP_ALIGN
    flags2 = ((k >= 0.25) & (k < 0.35)).astype(np.uint8) * nb.FLAG_SKIP_ALIGN
    sizes = []
    while sum(sizes) < n:
        sizes.append(int(rng.integers(1, 9)))
    sizes[-1] -= sum(sizes) - n
    scope = np.repeat(np.arange(len(sizes), dtype=np.uint32), sizes)
    compare(ctx, o, ocfg, r1, o1, r2, o2, q1=q1, q2=q2, flags1=flags1, flags2=flags2, scope=scope)
