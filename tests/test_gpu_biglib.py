"""Parity with an index that does NOT fit the 126 MB L2 (BASELINE.json config 4 shape): 8 000 families x 5 alleles =
40 000 transcripts, 80 000 index sequences, 44 M k-mers, ~2 GB on the device.  Everything the toy libraries cannot
exercise — multiply-high bucket reduction over a 10^8-slot table, `start_hi` bits of the unitig store, 64-byte walk
records in HBM, the L2 prefilter in front of the table, 10^5 colours — is compared with the oracle's independent index
and walk: the index dump itself, per-read reason / score / mismatches / equivalence classes, per-pair callsets, counts
and the probe / unitig / base work counters.  The oracle's own index build takes most of the time (under a minute on the GPU box)."""
import json

import numpy as np
import pytest

import nimble_aligner_b200 as nb
import oracle as orc
import synth
from tests.test_gpu_parity import compare

pytestmark = pytest.mark.gpu
N_FAM = 8000


@pytest.fixture(scope="module")
def big():
    L = synth.SynthLibrary(seed=3456, n_fam=N_FAM, n_all=5, group_on="")
    obj = L.to_json_obj()
    lib = nb.Library.from_text(json.dumps(obj), "unstranded")
    ix = nb.build_index(lib, 16, device=0)
    ocfg, oref = orc.parse_reference_library(obj, "unstranded")
    o = orc.Oracle(ocfg, oref)
    return L, lib, ix, ocfg, o


def test_big_index_has_the_oracles_graph(big):
    L, lib, ix, ocfg, o = big
    st, ost = ix.stats(), o.index_stats()
    assert st["device_bytes"] > 8 * (126 << 20), "the index must be far larger than L2 for this test to mean anything"
    for k in ("n_kmers", "n_nodes", "n_colours", "unitig_bases"):
        assert st[k] == ost[k], (k, st[k], ost[k])
    assert st["n_sequences"] == 2 * 5 * N_FAM


@pytest.mark.parametrize("mm", [0, 2])
def test_big_index_single_end_reads_match_the_oracle(big, mm):
    L, lib, ix, ocfg, o = big
    n = 150_000
    r1, o1, _, _ = synth.pairs(L, 1_000_000 * mm, n, seed=3456, paired=False, threads=16)
    oc = dict(ocfg, num_mismatches=mm)
    o.set_config(num_mismatches=mm)
    ctx = nb.Context(ix, lib, count_work=1, callset_slots=1 << 20)
    res, ref = compare(ctx, o, oc, r1, o1)
    w = ctx.work_counters()
    for k in ("probes", "nodes", "bases"):   # (colour ids: the device skips a colour equal to the previous unitig's)
        assert w[k] == ref["work"][k], (k, w[k], ref["work"][k])
    assert len(res["rows"]) > 5000          # thousands of families are hit: the colour / callset space is really exercised
    o.set_config(num_mismatches=0)


def test_big_index_pairs_match_the_oracle(big):
    L, lib, ix, ocfg, o = big
    n = 60_000
    r1, o1, r2, o2 = synth.pairs(L, 5_000_000, n, seed=3456, threads=16)
    ctx = nb.Context(ix, lib, count_work=1, callset_slots=1 << 20)
    res, ref = compare(ctx, o, ocfg, r1, o1, r2, o2)
    w = ctx.work_counters()
    for k in ("probes", "nodes", "bases"):   # (colour ids: the device skips a colour equal to the previous unitig's)
        assert w[k] == ref["work"][k], (k, w[k], ref["work"][k])


def test_big_index_chunking_and_device_resident_input_do_not_change_counts(big):
    L, lib, ix, ocfg, o = big
    import torch
    n = 1_500_000
    r1, o1, _, _ = synth.pairs(L, 0, n, seed=3456, paired=False, threads=16)
    a = nb.Context(ix, lib, max_batch_pairs=1 << 20, callset_slots=1 << 20)
    a.align_batch(r1, o1, max_read_len=150)
    ca = {tuple(cs): int(c) for _, cs, c in a.counts()["rows"]}
    b = nb.Context(ix, lib, max_batch_pairs=333_333, callset_slots=1 << 20)
    d1, do1 = torch.from_numpy(r1).cuda(), torch.from_numpy(o1.astype(np.int64)).cuda()
    b.align_batch(d1, do1, n_pairs=n, max_read_len=150, location=nb.NB_MEM_DEVICE)
    b.sync()
    cb = {tuple(cs): int(c) for _, cs, c in b.counts()["rows"]}
    assert ca == cb and len(ca) > 20000
    # the oracle pins a prefix of the same stream
    m = 100_000
    ref = o.run(r1, o1[: m + 1], threads=16, want_records=False)
    c = nb.Context(ix, lib, callset_slots=1 << 20)
    c.align_batch(r1, o1[: m + 1], max_read_len=150)
    got = sorted((tuple(cs), int(k)) for _, cs, k in c.counts()["rows"])
    assert got == sorted((tuple(cs), int(k)) for cs, k in ref["scopes"][0])
