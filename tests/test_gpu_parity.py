"""GPU parity tests (run with -m gpu on the B200 box).  Every test drives the CUDA path through the C ABI
(libnimble_b200.so) and compares it with the CPU oracle and/or the reference's literal golden expectations:
bit-exact per-read (reason, score, mismatches, trimmed length, equivalence class), per-pair (filter reasons, triage,
callset) and per-scope callset counts."""
import json
import os

import numpy as np
import pytest

import nimble_aligner_b200 as nb
import oracle as orc
import synth

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")
EXP = json.load(open(os.path.join(G, "expected.json")))
CHEMS = ["unstranded", "fiveprime", "threeprime", "none"]


def gpu_cfg(lib, ocfg):
    return lib.config.copy(score_percent=ocfg["score_percent"], score_threshold=ocfg["score_threshold"], num_mismatches=ocfg["num_mismatches"],
                           discard_multiple_matches=int(ocfg["discard_multiple_matches"]), require_valid_pair=int(ocfg["require_valid_pair"]),
                           discard_multi_hits=ocfg["discard_multi_hits"], max_hits_to_report=ocfg["max_hits_to_report"],
                           intersect_level=ocfg["intersect_level"], strand_filter=ocfg["strand_filter"],
                           trim_target_length=ocfg["trim_target_length"], trim_strictness=ocfg["trim_strictness"])


def compare(ctx, o, ocfg, r1, o1, r2=None, o2=None, q1=None, q2=None, flags1=None, flags2=None, scope=None, check_ecs=True, **ctx_kw):
    """Runs one batch on the GPU and on the oracle and asserts bit-exact agreement. Returns the GPU counts."""
    n = len(o1) - 1
    skip1 = (flags1 & 1).astype(np.uint8) if flags1 is not None else None
    skip2 = (flags2 & 1).astype(np.uint8) if flags2 is not None else None
    scope_off = None
    if scope is not None:
        b = np.flatnonzero(np.diff(scope)) + 1
        scope_off = np.concatenate([[0], b, [n]]).astype(np.uint64)
    ref = o.run(r1, o1, r2, o2, q1=q1, q2=q2, skip1=skip1, skip2=skip2, scope_off=scope_off, threads=4)
    ctx.reset()
    ctx.set_config(gpu_cfg(ctx.library, ocfg))
    reads, pairs = ctx.align_batch(r1, o1, r2, o2, q1=q1, q2=q2, flags1=flags1, flags2=flags2, scope_id=scope, want_reads=True, want_pairs=True)
    res = ctx.counts()
    # ---- per read
    assert np.array_equal(reads["reason"], ref["read_reason"]), np.flatnonzero(reads["reason"] != ref["read_reason"])[:10]
    assert np.array_equal(reads["pass"], ref["read_pass"])
    assert np.array_equal(reads["score"], ref["read_score"]), np.flatnonzero(reads["score"] != ref["read_score"])[:10]
    assert np.array_equal(reads["mismatches"], ref["read_mm"]), np.flatnonzero(reads["mismatches"] != ref["read_mm"])[:10]
    assert np.array_equal(reads["ec_len"], ref["read_ec_len"]), np.flatnonzero(reads["ec_len"] != ref["read_ec_len"])[:10]
    notskip = reads["reason"] != nb.R["SkippedAlignDueToUnpairedDummy"]
    assert np.array_equal(reads["trimmed_len"][notskip], ref["read_trimmed_len"][notskip])
    if check_ecs and n <= ctx_kw.get("max_batch_pairs", 1 << 20):
        off, ids = ctx.last_batch_ecs(len(reads))
        assert np.array_equal(ids, ref["read_ec"])
    # ---- per pair
    assert np.array_equal(pairs["fr1"], ref["pair_fr1"]), np.flatnonzero(pairs["fr1"] != ref["pair_fr1"])[:10]
    assert np.array_equal(pairs["fr2"], ref["pair_fr2"])
    assert np.array_equal(pairs["triage"], ref["pair_triage"]), np.flatnonzero(pairs["triage"] != ref["pair_triage"])[:10]
    has = pairs["callset"] != 0xFFFFFFFF
    assert np.array_equal(has, ref["pair_counted"].astype(bool))
    s2c = res["slot_to_callset"]
    scope_of = np.zeros(n, dtype=np.int64) if scope is None else (np.searchsorted(scope_off, np.arange(n), side="right") - 1)
    for p in np.flatnonzero(has)[:: max(1, int(has.sum()) // 5000)]:
        assert res["callsets"][s2c[pairs["callset"][p]]] == ref["scopes"][scope_of[p]][ref["pair_callset"][p]][0], p
    # ---- per scope counts (rows sorted by Vec<String> Ord, like utils::sort_score_vector)
    got = {}
    for sc, cs, cnt in res["rows"]:
        got.setdefault(sc, []).append((cs, cnt))
    scope_ids = [0] if scope is None else [int(scope[int(a)]) for a in scope_off[:-1]]
    for si, sid in enumerate(scope_ids):
        assert got.get(sid, []) == ref["scopes"][si], (sid, got.get(sid), ref["scopes"][si])
    assert sum(len(v) for v in got.values()) == sum(len(s) for s in ref["scopes"])
    return res, ref


def load_fixture(lib_name, chem="none", group=False):
    path = os.path.join(G, "ref", "libraries", lib_name)
    ocfg, oref = orc.get_reference_library(path, chem)
    _, lib = nb.get_reference_library(path, chem)
    if group:
        oref.group_on = 4
        oref.headers.append("test_group_on")
        oref.columns.append(list(EXP["group_column"]))
        lib.push_column("test_group_on", EXP["group_column"], set_group_on=True)
    return ocfg, oref, lib


# ------------------------------------------------------------------ C1: the reference's golden vectors
@pytest.mark.parametrize("case", EXP["get_calls"], ids=[c["src"] for c in EXP["get_calls"]])
def test_reference_goldens(case):
    ocfg, oref, lib = load_fixture(case["lib"], "none", case["group"])
    ocfg["num_mismatches"] = case["mm"]
    reads, _ = orc.read_fastq(os.path.join(G, "ref", "reads", case["reads"]))
    ix = nb.build_index(lib, 2)
    rows, _, _ = nb.get_calls(reads, None, None, ix, lib, gpu_cfg(lib, ocfg))
    assert [[cs, n] for cs, n in rows] == case["expect"]
    r1, o1 = orc.pack_reads(reads)
    compare(nb.Context(ix, lib), orc.Oracle(ocfg, oref), ocfg, r1, o1)


def test_pseudoalign_known_answers():   # src/align.rs:1061-1107 (min_read_length 12 there)
    cfg = nb.Config(score_percent=0.1, score_threshold=50, num_mismatches=3, max_hits_to_report=5, intersect_level=1, strand_filter=1,
                    trim_target_length=15, trim_strictness=0.5)
    lib = nb.Library.from_columns(["sequence_name", "sequence"], [["Gene1", "Gene2"], ["ACGT" * 8, "TGCA" * 8]], 0, 0, 1, cfg)
    ix = nb.build_index(lib, 1)
    ctx = nb.Context(ix, lib, min_read_length=12)
    seqs = ["ACG", "A" * 30, "CCTGAGATTTCGAGCTCGTAACGTGACCTACGGACAC", "TGCA" * 8]
    r1, o1 = nb.pack_reads(seqs)

    def run(**kw):
        ctx.reset()
        ctx.set_config(cfg.copy(**kw))
        reads, _ = ctx.align_batch(r1, o1, want_reads=True)
        return reads

    r = run(score_threshold=32)
    assert [int(x) for x in r["reason"][:3]] == [nb.R["ShortRead"], nb.R["HighEntropy"], nb.R["NoMatch"]]
    assert r["pass"][3] == 1 and r["score"][3] == 32 and r["ec_len"][3] == 1   # Some(([1], 1.0, 32))
    off, ids = ctx.last_batch_ecs(4)
    assert ids.tolist() == [1]
    r = run(score_threshold=1000)
    assert r["reason"][3] == nb.R["ScoreBelowThreshold"] and r["score"][3] == 32 and r["pass"][3] == 0


# ------------------------------------------------------------------ C2-shaped synthetic parity (1k-transcript family library)
@pytest.fixture(scope="module")
def c2():
    L = synth.SynthLibrary(seed=1234, n_fam=200, n_all=5, group_on="")
    obj = L.to_json_obj()
    return L, obj


def make(obj, chem, group_on=""):
    import copy
    obj = [dict(obj[0]), obj[1]]
    obj[0]["group_on"] = group_on
    ocfg, oref = orc.parse_reference_library(obj, chem)
    lib = nb.Library.from_text(json.dumps(obj), chem)
    return ocfg, oref, lib


@pytest.fixture(scope="module")
def c2_built(c2):
    L, obj = c2
    built = {}
    for group_on in ("", "family"):
        ocfg, oref, lib = make(obj, "unstranded", group_on)
        ix = nb.build_index(lib, 8)
        built[group_on] = (ocfg, oref, lib, ix, orc.Oracle(ocfg, oref), nb.Context(ix, lib))
    return L, built


def test_c2_index_matches_oracle(c2_built):
    L, built = c2_built
    ocfg, oref, lib, ix, o, ctx = built[""]
    so, sp = o.index_stats(), ix.stats()
    for k in so:
        assert so[k] == sp[k], k
    assert o.index_dump() == ix.dump()


@pytest.mark.parametrize("mm", [0, 1, 2])
@pytest.mark.parametrize("paired", [True, False])
def test_c2_parity_mismatch_sweep(c2_built, mm, paired):   # also config C5's sweep
    L, built = c2_built
    ocfg, oref, lib, ix, o, ctx = built[""]
    cfg = dict(ocfg, num_mismatches=mm)
    o.set_config(**cfg)
    r1, o1, r2, o2 = synth.pairs(L, 0, 60000, paired=paired)
    res, ref = compare(ctx, o, cfg, r1, o1, r2, o2)
    assert res["n_unique_keys"] < 60000 and len(res["rows"]) > 100   # duplicates exist and many callsets are hit


@pytest.mark.parametrize("chem", CHEMS)
@pytest.mark.parametrize("level", [0, 1, 2])
def test_c2_parity_chemistry_and_intersect(c2_built, chem, level):
    L, built = c2_built
    for group_on in ("", "family"):
        ocfg, oref, lib, ix, o, ctx = built[group_on]
        cfg = dict(ocfg, strand_filter=chem, intersect_level=level, num_mismatches=1)
        o.set_config(**cfg)
        r1, o1, r2, o2 = synth.pairs(L, 100000, 20000, paired=True)
        compare(ctx, o, cfg, r1, o1, r2, o2, check_ecs=False)


@pytest.mark.parametrize("kw", [dict(require_valid_pair=True), dict(discard_multiple_matches=True), dict(discard_multi_hits=1),
                                dict(max_hits_to_report=1), dict(discard_multi_hits=2, max_hits_to_report=1), dict(score_threshold=140, score_percent=0.95),
                                dict(require_valid_pair=True, strand_filter="none", intersect_level=2)])
def test_c2_parity_filters(c2_built, kw):
    L, built = c2_built
    ocfg, oref, lib, ix, o, ctx = built["family"]
    cfg = dict(ocfg, num_mismatches=2)
    cfg.update(kw)
    o.set_config(**cfg)
    r1, o1, r2, o2 = synth.pairs(L, 200000, 20000, paired=True)
    compare(ctx, o, cfg, r1, o1, r2, o2, check_ecs=False)


def test_multi_batch_equals_single_batch_and_is_idempotent(c2_built):
    """Whole-run de-duplication across batches: chunked == single; feeding the same pairs twice changes nothing."""
    L, built = c2_built
    ocfg, oref, lib, ix, o, _ = built[""]
    o.set_config(**ocfg)
    r1, o1, r2, o2 = synth.pairs(L, 0, 50000, paired=True)
    ref = o.run(r1, o1, r2, o2, threads=4, want_records=False)["scopes"][0]
    ctx = nb.Context(ix, lib, max_batch_pairs=7001, key_slots=1024)   # forces 8 chunks and several key-table growths
    ctx.set_config(gpu_cfg(lib, ocfg))
    ctx.align_batch(r1, o1, r2, o2)
    ctx.align_batch(r1, o1, r2, o2)   # every pair again: duplicates of existing keys
    res = ctx.counts()
    assert [(cs, n) for _, cs, n in res["rows"]] == ref
    assert res["n_pairs_seen"] == 100000


# ------------------------------------------------------------------ C3-shaped: scoped (UMI, CB) groups, quality trim, dummy mates
def test_c3_scoped_bam_like_parity(c2_built):
    L, built = c2_built
    ocfg, oref, lib, ix, o, ctx = built[""]
    u = synth.umi_reads(L, 0, 6000)
    n = u["n_reads"]
    # 10x single-end records become (dummy, real) pairs (sorted_bam_reader.rs:109-145): sequence slot = SKIP_ALIGN dummy
    rng = np.random.default_rng(3)
    flags2 = (rng.random(n) < 0.3).astype(np.uint8) * nb.FLAG_REVCOMP     # REVERSE records are reverse-complemented
    flags1 = (flags2 | nb.FLAG_SKIP_ALIGN).astype(np.uint8)               # the dummy is a clone of the record (same REVERSE flag)
    for chem in ("unstranded", "fiveprime", "threeprime"):
        cfg = dict(ocfg, strand_filter=chem, trim_target_length=40, trim_strictness=0.9, num_mismatches=1)
        o.set_config(**cfg)
        # the oracle gets the reads already reverse-complemented / quals reversed, as process::bam::get_calls hands them over
        bases = u["bases"][: n * 91].reshape(n, 91).copy()
        qual = u["qual"][: n * 91].reshape(n, 91).copy()
        rc = flags2.astype(bool)
        comp = np.zeros(256, dtype=np.uint8)
        for a, b in zip(b"ACGT", b"TGCA"):
            comp[a] = b
        ob, oq = bases.copy(), qual.copy()
        ob[rc] = comp[bases[rc][:, ::-1]]
        oq[rc] = qual[rc][:, ::-1]
        skip1 = np.ones(n, dtype=np.uint8)
        scope_off = np.concatenate([[0], np.cumsum(u["sizes"])]).astype(np.uint64)
        ref = o.run(ob.reshape(-1), u["off"], ob.reshape(-1), u["off"], q1=oq.reshape(-1), q2=oq.reshape(-1), skip1=skip1, skip2=None, scope_off=scope_off, threads=4)
        ctx.reset()
        ctx.set_config(gpu_cfg(lib, cfg))
        reads, pairs = ctx.align_batch(u["bases"], u["off"], u["bases"], u["off"], q1=u["qual"], q2=u["qual"], flags1=flags1, flags2=flags2,
                                       scope_id=u["scope"], want_reads=True, want_pairs=True)
        res = ctx.counts()
        real = np.arange(n) * 2 + 1
        assert np.array_equal(reads["reason"][real], ref["read_reason"][real])
        assert np.all(reads["reason"][real - 1] == nb.R["SkippedAlignDueToUnpairedDummy"])
        assert np.array_equal(reads["trimmed_len"][real], ref["read_trimmed_len"][real])
        assert np.array_equal(reads["score"][real], ref["read_score"][real]) and np.array_equal(reads["mismatches"][real], ref["read_mm"][real])
        assert len(np.unique(reads["trimmed_len"][real])) > 5          # the Q2 tails really trim
        assert np.array_equal(pairs["triage"], ref["pair_triage"]) and np.array_equal(pairs["fr2"], ref["pair_fr2"])
        got = {}
        for sc, cs, cnt in res["rows"]:
            got.setdefault(sc, []).append((cs, cnt))
        for s in range(len(u["sizes"])):
            assert got.get(s, []) == ref["scopes"][s], s
        # per-cell table (cell_id given): sum over the cell's scopes
        ctx.reset()
        ctx.align_batch(u["bases"], u["off"], u["bases"], u["off"], q1=u["qual"], q2=u["qual"], flags1=flags1, flags2=flags2, scope_id=u["scope"], cell_id=u["cell"])
        cells = {}
        for sc, cs, cnt in ctx.counts()["rows"]:
            cells[(sc, tuple(cs))] = cnt
        want = {}
        first = scope_off[:-1].astype(np.int64)
        for s in range(len(u["sizes"])):
            for cs, cnt in ref["scopes"][s]:
                k = (int(u["cell"][first[s]]), tuple(cs))
                want[k] = want.get(k, 0) + cnt
        assert cells == want


# ------------------------------------------------------------------ edge cases
def test_edge_cases_ragged_empty_nonacgt(c2_built):
    L, built = c2_built
    ocfg, oref, lib, ix, o, ctx = built[""]
    o.set_config(**ocfg)
    seqs = L.sequences()
    t = seqs[17]
    reads = ["", "A", t[:39], t[:40], t[100:250], t[100:250].lower(), t[100:180] + "N" + t[181:250], "N" * 150, "ACGT" * 40, t[:1024],
             t[-150:], t[-100:] + "ACGTTGCA" * 6, "G" * 29 + t[200:321], t[5:155][:75] + "T" + t[5:155][76:]]
    mates = [r[::-1] for r in reads]
    r1, o1 = orc.pack_reads(reads)
    r2, o2 = orc.pack_reads(mates)
    for mm in (0, 2):
        cfg = dict(ocfg, num_mismatches=mm)
        o.set_config(**cfg)
        compare(ctx, o, cfg, r1, o1)
        compare(ctx, o, cfg, r1, o1, r2, o2)
    ctx.reset()
    e = np.zeros(1, dtype=np.uint64)
    ctx.align_batch(np.zeros(1, dtype=np.uint8), e)   # empty batch
    assert ctx.counts()["rows"] == []
    with pytest.raises(nb.NbError):   # longer than the device path supports -> loud error, not a silent wrong answer
        big, ob = orc.pack_reads(["ACGT" * 300])
        ctx.align_batch(big, ob)


def test_left_extension_and_reseed_paths_are_exercised(c2_built):
    """Reads with an error at positions 27..29 make every seed before position 30 miss, so the first hit is at
    >= floor(0.2*len) and the left extension runs; the oracle's work counters must match the device's."""
    L, built = c2_built
    ocfg, oref, lib, ix, o, _ = built[""]
    seqs = L.sequences()
    rng = np.random.default_rng(11)
    reads = []
    for i in range(4000):
        t = seqs[int(rng.integers(len(seqs)))]
        s = int(rng.integers(0, len(t) - 150))
        r = list(t[s:s + 150])
        for pos in rng.choice([27, 28, 29, 57, 58, 59, 100], size=int(rng.integers(1, 4)), replace=False):
            r[pos] = "ACGT"[("ACGT".index(r[pos]) + 1 + int(rng.integers(3))) % 4]
        reads.append("".join(r))
    r1, o1 = orc.pack_reads(reads)
    ctx = nb.Context(ix, lib, count_work=1)
    for mm in (0, 1, 3):
        cfg = dict(ocfg, num_mismatches=mm)
        o.set_config(**cfg)
        res, ref = compare(ctx, o, cfg, r1, o1)
        w = ctx.work_counters()   # reset with the tables by compare()'s ctx.reset()
        for k in ("probes", "nodes", "bases"):
            assert w[k] == ref["work"][k], (mm, k, w[k], ref["work"][k])
    assert w["probes"] > 4000 and w["nodes"] > 4000


def test_long_reads_multi_round_seed_search(c2_built):
    """Reads of 300-1000 bases: the cooperative seed search needs several 64-seed rounds (first hit in round 2, 3, ...;
    off-target reads exhaust up to 6 rounds), the walk crosses a whole transcript, long junk prefixes trigger the left
    extension late.  Results and probe counts (= the sequential search's) must match the oracle."""
    L, built = c2_built
    ocfg, oref, lib, ix, o, _ = built[""]
    seqs = L.sequences()
    rng = np.random.default_rng(23)
    rnd = lambda n: "".join("ACGT"[int(x)] for x in rng.integers(0, 4, size=n))
    reads = []
    for i in range(600):
        t = seqs[int(rng.integers(len(seqs)))]
        kind = i % 6
        if kind == 0:   # off-target, every seed of every round misses
            reads.append(rnd(int(rng.integers(300, 1001))))
        elif kind == 1:   # junk prefix: the first hit is at seed index prefix/3 (rounds 2..5), then the left extension
            pre = int(rng.integers(200, 700)); seg = int(rng.integers(60, 300)); s0 = int(rng.integers(0, len(t) - seg))
            reads.append((rnd(pre) + t[s0:s0 + seg])[:1024])
        elif kind == 2:   # on-target, one error every 25 bases over the first 250-600 bases: no clean 30-mer until then
            n = int(rng.integers(700, min(1000, len(t)))); r = list(t[:n]); stop = int(rng.integers(250, 600))
            for pos in range(int(rng.integers(0, 25)), stop, 25):
                r[pos] = "ACGT"[("ACGT".index(r[pos]) + 1 + int(rng.integers(3))) % 4]
            reads.append("".join(r))
        elif kind == 3:   # clean long read (walks hundreds of bases, many unitigs)
            n = int(rng.integers(300, min(1024, len(t)))); s0 = int(rng.integers(0, len(t) - n + 1)); reads.append(t[s0:s0 + n])
        elif kind == 4:   # chimera of two transcripts with junk in between (re-seed far into the read)
            u = seqs[int(rng.integers(len(seqs)))]; reads.append((t[50:250] + rnd(int(rng.integers(100, 300))) + u[100:400])[:1024])
        else:             # ordinary 150 bp read next to the long ones (ragged batch)
            s0 = int(rng.integers(0, len(t) - 150)); reads.append(t[s0:s0 + 150])
    assert max(len(r) for r in reads) > 900
    r1, o1 = orc.pack_reads(reads)
    r2, o2 = orc.pack_reads(reads[::-1])
    ctx = nb.Context(ix, lib, count_work=1)
    for mm in (0, 2):
        cfg = dict(ocfg, num_mismatches=mm)
        o.set_config(**cfg)
        res, ref = compare(ctx, o, cfg, r1, o1)
        w = ctx.work_counters()
        for k in ("probes", "nodes", "bases"):
            assert w[k] == ref["work"][k], (mm, k, w[k], ref["work"][k])
        compare(ctx, o, cfg, r1, o1, r2, o2)
    assert w["probes"] > 100 * 64     # the multi-round searches really ran


def test_work_counters_match_oracle(c2_built):
    L, built = c2_built
    ocfg, oref, lib, ix, o, _ = built[""]
    o.set_config(**ocfg)
    r1, o1, r2, o2 = synth.pairs(L, 0, 30000, paired=True)
    ref = o.run(r1, o1, r2, o2, threads=4, want_records=False)
    ctx = nb.Context(ix, lib, count_work=1)
    ctx.set_config(gpu_cfg(lib, ocfg))
    ctx.align_batch(r1, o1, r2, o2)
    w = ctx.work_counters()
    for k in ("probes", "nodes", "bases"):
        assert w[k] == ref["work"][k], (k, w[k], ref["work"][k])


def test_fastq_driver_and_tsv_format(tmp_path, c2):
    """process::fastq::process end to end: FASTQ(.gz) in, TSV out (append mode, header once) — src/utils.rs:27-51."""
    import gzip
    L, obj = c2
    lib_json = tmp_path / "lib.json"
    lib_json.write_text(json.dumps(obj))
    r1, o1, r2, o2 = synth.pairs(L, 0, 3000, paired=True)

    def fq(path, data, off, gz):
        op = gzip.open if gz else open
        with op(path, "wt") as f:
            for i in range(len(off) - 1):
                s = bytes(data[int(off[i]):int(off[i + 1])]).decode()
                f.write("@r%d\n%s\n+\n%s\n" % (i, s, "I" * len(s)))
    fq(tmp_path / "a_R1.fastq.gz", r1, o1, True)
    fq(tmp_path / "a_R2.fastq", r2, o2, False)
    out = tmp_path / "out.tsv"
    nb.process_fastq([str(tmp_path / "a_R1.fastq.gz"), str(tmp_path / "a_R2.fastq")], [str(lib_json)], [str(out)], "unstranded", 4)
    ocfg, oref = orc.parse_reference_library(obj, "unstranded")
    ref = orc.Oracle(ocfg, oref).run(r1, o1, r2, o2, want_records=False)["scopes"][0]
    want = "feature\tscore\n" + "".join("\t".join(cs) + "\t%d\n" % n for cs, n in ref)
    assert out.read_text() == want
    nb.process_fastq([str(tmp_path / "a_R1.fastq.gz"), str(tmp_path / "a_R2.fastq")], [str(lib_json)], [str(out)], "unstranded", 4)
    assert out.read_text() == want + want[len("feature\tscore\n"):]   # append, no second header


def test_big_component_list_and_arena_paths():
    """Components of > 64 sequences have no bitmap colours: exercises the list-mask path (colour <= 64 ids), the arena
    path (colour > 64 ids) and their transitions, plus mixing with small components."""
    Lbig = synth.SynthLibrary(seed=99, n_fam=2, n_all=90)
    # make the 90 alleles of each family near-identical (2 SNPs each) so that most k-mers are shared by > 64 sequences
    rng = np.random.default_rng(17)
    for f in range(2):
        base = Lbig.seqs[int(Lbig.off[f * 90]):int(Lbig.off[f * 90 + 1])].copy()
        for a in range(90):
            s = base.copy()
            for pos in rng.integers(0, len(s), size=2):
                s[pos] = ord("ACGT"[(("ACGT".index(chr(s[pos]))) + 1 + int(rng.integers(3))) % 4])
            Lbig.seqs[int(Lbig.off[f * 90 + a]):int(Lbig.off[f * 90 + a + 1])] = s
    Lsmall = synth.SynthLibrary(seed=5, n_fam=6, n_all=4)
    obj = Lbig.to_json_obj()
    so = Lsmall.to_json_obj()
    for ci in range(len(obj[1]["columns"])):
        obj[1]["columns"][ci] = obj[1]["columns"][ci] + [("S" + x if ci in (1, 2) else x) for x in so[1]["columns"][ci]]
    for group_on, kw in (("family", {}), ("", dict(max_hits_to_report=60, discard_multi_hits=0)), ("", dict(max_hits_to_report=3))):
        ocfg, oref, lib = make(obj, "unstranded", group_on)
        ocfg.update(kw)
        ix = nb.build_index(lib, 4)
        st = ix.stats()
        o = orc.Oracle(ocfg, oref)
        assert o.index_dump() == ix.dump()
        ctx = nb.Context(ix, lib)
        for mm in (0, 2):
            cfg = dict(ocfg, num_mismatches=mm)
            o.set_config(**cfg)
            r1, o1, r2, o2 = synth.pairs(Lbig, 0, 6000, seed=77, paired=True)
            s1, so1, s2, so2 = synth.pairs(Lsmall, 0, 2000, seed=78, paired=True)
            a1 = np.concatenate([r1[: int(o1[-1])], s1[: int(so1[-1])]]); a2 = np.concatenate([r2[: int(o2[-1])], s2[: int(so2[-1])]])
            b1 = np.concatenate([o1, so1[1:] + o1[-1]]); b2 = np.concatenate([o2, so2[1:] + o2[-1]])
            res, ref = compare(ctx, o, cfg, a1, b1, a2, b2)
            assert ref["read_ec_len"].max() > 64      # the arena path really ran


def test_tangled_orientations_all_configs():
    """Adversarial library for the pair stage: features that contain each other's reverse complements, a feature that
    carries one segment in both orientations, a reverse-palindromic feature (its fwd and §rev rows are identical), shared
    and empty group strings.  Reads and mates in every orientation combination, then every chemistry x intersect level
    x pair/hit filter combination is compared with the oracle's literal string pipeline."""
    rng = np.random.default_rng(42)
    rs = lambda n: "".join("ACGT"[i] for i in rng.integers(0, 4, n))
    rc = lambda s: s[::-1].translate(str.maketrans("ACGT", "TGCA"))
    f0 = rs(600)
    f1 = rc(f0[100:400]) + rs(300)                 # F1 forward shares k-mers with F0 reverse
    f2 = f0[0:300] + rs(50) + rc(f0[0:300])        # one segment in both orientations inside one feature
    x = rs(200)
    f3 = x + rc(x)                                 # reverse palindrome: F3 == revcomp(F3)
    f4 = f0[:250] + rs(350)                        # plain sharing with F0 (same orientation)
    f5 = rs(500)
    f6 = f5[:200] + rc(f1[350:550]) + rs(100)
    seqs = [f0, f1, f2, f3, f4, f5, f6]
    names = ["T%d" % i for i in range(len(seqs))]
    groups = ["gA", "gA", "", "gB", "gB", "", "gC"]
    cfg0 = dict(synth.BASE_CONFIG, group_on="grp", num_mismatches=1)
    obj = [cfg0, {"headers": ["sequence_name", "grp", "sequence"], "columns": [names, groups, seqs]}]
    reads, mates = [], []
    pool = seqs + [rc(s) for s in seqs]
    for _ in range(2500):
        a = pool[int(rng.integers(len(pool)))]
        s = int(rng.integers(0, len(a) - 120)); r = a[s:s + 120]
        k = rng.random()
        if k < 0.4:      # proper FR mate from the same molecule
            t = int(rng.integers(s, min(len(a) - 120, s + 200) + 1)); m = rc(a[t:t + 120])
        elif k < 0.6:    # same orientation (FF / RR)
            t = int(rng.integers(0, len(a) - 120)); m = a[t:t + 120]
        elif k < 0.8:    # unrelated molecule
            b = pool[int(rng.integers(len(pool)))]; t = int(rng.integers(0, len(b) - 120)); m = b[t:t + 120]
        else:            # junk mate
            m = rs(120)
        if rng.random() < 0.1:
            r = r[:60] + "ACGT"[("ACGT".index(r[60]) + 1) % 4] + r[61:]
        reads.append(r); mates.append(m)
    r1, o1 = orc.pack_reads(reads); r2, o2 = orc.pack_reads(mates)
    n_callsets = 0
    for group_on in ("grp", ""):
        obj[0]["group_on"] = group_on
        ocfg, oref, lib = make(obj, "unstranded", group_on)
        ix = nb.build_index(lib, 2)
        o = orc.Oracle(ocfg, oref)
        assert o.index_dump() == ix.dump()
        ctx = nb.Context(ix, lib)
        for chem in CHEMS:
            for level in (0, 1, 2):
                for kw in (dict(), dict(require_valid_pair=True), dict(discard_multi_hits=1), dict(max_hits_to_report=1, discard_multiple_matches=True)):
                    cfg = dict(ocfg, strand_filter=chem, intersect_level=level)
                    cfg.update(kw)
                    o.set_config(**cfg)
                    res, ref = compare(ctx, o, cfg, r1, o1, r2, o2, check_ecs=(chem == "none" and level == 0 and not kw))
                    n_callsets += len(res["rows"])
        o.set_config(**dict(ocfg, strand_filter="fiveprime"))
        compare(ctx, o, dict(ocfg, strand_filter="fiveprime"), r1, o1)     # single-end through the same logic
    assert n_callsets > 200


def test_scope_larger_than_max_batch_pairs_is_still_deduplicated_as_a_whole(c2_built):
    """A (UMI, CB) scope with more pairs than max_batch_pairs: the chunk must grow to the end of the scope — cutting it would
    clear the key table in the middle and count duplicate read_keys twice (src/align.rs:576-579, 685)."""
    L, built = c2_built
    ocfg, oref, lib, ix, o, _ = built[""]
    o.set_config(**ocfg)
    r1, o1, r2, o2 = synth.pairs(L, 0, 700, dup_rate=0.5, paired=True)
    n = len(o1) - 1
    # scopes: 0 = pairs [0, 40), 1 = one big scope [40, 600) with many duplicates, 2.. = small ones
    scope = np.zeros(n, dtype=np.uint32)
    scope[40:600] = 1
    scope[600:] = 2 + (np.arange(n - 600) // 7)
    ctx = nb.Context(ix, lib, max_batch_pairs=64)
    res, ref = compare(ctx, o, ocfg, r1, o1, r2, o2, scope=scope, check_ecs=False)
    big = nb.Context(ix, lib, max_batch_pairs=1 << 20)
    big.set_config(gpu_cfg(lib, ocfg))
    big.align_batch(r1, o1, r2, o2, scope_id=scope)
    assert big.counts()["rows"] == res["rows"]
    assert sum(c for sc, _cs, c in res["rows"] if sc == 1) < 560     # duplicates inside the big scope really were merged


def test_hbm_resident_kernel_variant_on_a_small_library(c2_built, monkeypatch):
    """The kernels compiled for an HBM-resident index (L2 eviction hints on the prefilter and the table buckets) are chosen
    by table size; NB_L2_HINTS=1 forces them for a small library so that both variants see the same parity cases."""
    L, built = c2_built
    ocfg, oref, lib, ix, o, _ = built[""]
    o.set_config(**ocfg)
    monkeypatch.setenv("NB_L2_HINTS", "1")
    ctx = nb.Context(ix, lib, count_work=1)
    monkeypatch.delenv("NB_L2_HINTS")
    r1, o1, r2, o2 = synth.pairs(L, 300_000, 20_000, paired=True)
    res, ref = compare(ctx, o, ocfg, r1, o1, r2, o2)
    w = ctx.work_counters()
    for k in ("probes", "nodes", "bases"):
        assert w[k] == ref["work"][k], (k, w[k], ref["work"][k])


def test_bulk_copy_pack_variant(c2_built, monkeypatch):
    """NB_PACK_BULK=1 routes ASCII batches through the staged K0 (cp.async.bulk into shared memory, 2-bit image, word cut):
    same per-read / per-pair records and counts as the direct kernel, on ragged reads (empty, 1 base, 1024 bases), reverse
    complement flags, lower case and non-ACGT letters, single and paired, chunk boundaries that are not block multiples."""
    L, built = c2_built
    ocfg, oref, lib, ix, o, _ = built[""]
    o.set_config(**ocfg)
    seqs = L.sequences()
    t = seqs[23]
    reads = ["", "A", t[:39], t[:40], t[100:250], t[100:250].lower(), t[100:180] + "N" + t[181:250], "NRYKM" * 30, "ACGT" * 40, t[:1024], t[-150:], "G" * 29 + t[200:321]] * 7
    r1, o1 = orc.pack_reads(reads)
    r2, o2 = orc.pack_reads([r[::-1] for r in reads])
    rng = np.random.default_rng(11)
    f1 = (rng.random(len(reads)) < 0.4).astype(np.uint8) * nb.FLAG_REVCOMP
    f2 = (rng.random(len(reads)) < 0.4).astype(np.uint8) * nb.FLAG_REVCOMP
    a1, b1, a2, b2 = synth.pairs(L, 900_000, 20_011, paired=True)
    for args, kw in (((r1, o1), dict(flags1=f1)), ((r1, o1, r2, o2), dict(flags1=f1, flags2=f2)), ((a1, b1, a2, b2), {}), ((a1, b1), {})):
        out = []
        for bulk in ("0", "1"):
            monkeypatch.setenv("NB_PACK_BULK", bulk)
            ctx = nb.Context(ix, lib, max_batch_pairs=6_007)
            rr, _ = ctx.align_batch(*args, want_reads=True, **kw)
            out.append((rr.tobytes(), ctx.counts()["rows"]))
            ctx.close()
        assert out[0] == out[1]
    monkeypatch.setenv("NB_PACK_BULK", "1")
    ctx = nb.Context(ix, lib)
    compare(ctx, o, ocfg, a1, b1, a2, b2)        # and the staged kernel against the oracle directly
    ctx.close()


# ------------------------------------------------------------------ packed input encodings (nb_batch.encoding)
@pytest.mark.parametrize("enc", ["2bit", "bam4"])
@pytest.mark.parametrize("read_len", [150, 151, 45])
def test_packed_encodings_equal_the_ascii_path(c2_built, enc, read_len):
    """NB_SEQ_2BIT / NB_SEQ_BAM4 batches (offsets in bases, reads starting at any base position — odd lengths put reads at
    odd nibble / quarter-byte positions) must give the ASCII path's per-read and per-pair records and counts, which the
    oracle pins; reverse-complement flags and non-ACGT letters included."""
    L, built = c2_built
    ocfg, oref, lib, ix, o, _ = built[""]
    o.set_config(**ocfg)
    n = 30_000
    r1, o1, r2, o2 = synth.pairs(L, 700_000, n, read_len=read_len, paired=True)
    rng = np.random.default_rng(5)
    for r, off in ((r1, o1), (r2, o2)):      # sprinkle N / IUPAC letters: both packed paths must turn them into A like from_acgt_bytes
        pos = rng.integers(0, int(off[-1]), size=300)
        r[pos] = np.frombuffer(b"NRYKMn", dtype=np.uint8)[rng.integers(0, 6, size=300)]
    flags1 = (rng.random(n) < 0.3).astype(np.uint8) * nb.FLAG_REVCOMP
    flags2 = (rng.random(n) < 0.3).astype(np.uint8) * nb.FLAG_REVCOMP
    ctx = nb.Context(ix, lib, max_batch_pairs=7_001)       # chunk boundaries at arbitrary base offsets too
    def dense(pairs):   # nb_pair_result.callset is a dictionary SLOT (which of two colliding callsets gets which slot is a race): compare callset ids
        p = pairs.copy()
        s2c = ctx.counts_raw()["slot_to_callset"]
        has = p["callset"] != 0xFFFFFFFF
        p["callset"][has] = s2c[p["callset"][has]]
        return p
    ra, pa = ctx.align_batch(r1, o1, r2, o2, flags1=flags1, flags2=flags2, want_reads=True, want_pairs=True)
    ca = ctx.counts()["rows"]
    pa = dense(pa)
    f = nb.encode_2bit if enc == "2bit" else nb.encode_bam4
    code = nb.NB_SEQ_2BIT if enc == "2bit" else nb.NB_SEQ_BAM4
    p1, p2 = f(r1, o1), f(r2, o2)
    ctx.reset()
    rb, pb = ctx.align_batch(p1, o1, p2, o2, flags1=flags1, flags2=flags2, want_reads=True, want_pairs=True, encoding=code)
    cb = ctx.counts()["rows"]
    pb = dense(pb)
    assert ra.tobytes() == rb.tobytes() and pa.tobytes() == pb.tobytes() and ca == cb and (len(ca) > 100 or read_len < 50)   # (45-base reads never reach score_threshold 50)
    if read_len == 150 and enc == "2bit":    # and the ASCII path itself is the oracle's (unflagged reads)
        ctx.reset()
        ctx.align_batch(p1, o1, p2, o2, encoding=code)
        ref = o.run(r1, o1, r2, o2, threads=4, want_records=False)
        assert [(cs, c) for _, cs, c in ctx.counts()["rows"]] == ref["scopes"][0]


def test_packed_encoding_with_explicit_lengths_and_gaps(c2_built):
    """r1_len / r2_len: reads that are not densely packed (every read starts on a 32-base boundary of the 2-bit stream, the
    natural layout of a host that keeps one word-aligned DnaString per read)."""
    L, built = c2_built
    ocfg, oref, lib, ix, o, _ = built[""]
    n = 5_000
    r1, o1, _, _ = synth.pairs(L, 50_000, n, paired=False)
    lens = np.diff(o1).astype(np.uint32)
    starts = np.concatenate([[0], np.cumsum((lens + 31) // 32 * 32)]).astype(np.uint64)
    wide = np.zeros(int(starts[-1]) + 64, dtype=np.uint8) + ord("A")
    for i in range(n):
        wide[int(starts[i]):int(starts[i]) + int(lens[i])] = r1[int(o1[i]):int(o1[i + 1])]
    packed = nb.encode_2bit(wide, starts)
    ctx = nb.Context(ix, lib)
    ra, _ = ctx.align_batch(r1, o1, want_reads=True)
    ca = ctx.counts()["rows"]
    ctx.reset()
    rb, _ = ctx.align_batch(packed, starts, want_reads=True, encoding=nb.NB_SEQ_2BIT, r1_len=lens)
    assert ra.tobytes() == rb.tobytes() and ctx.counts()["rows"] == ca
