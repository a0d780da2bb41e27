"""Multi-GPU merge inside the library (nb_comm_*, nb_merge_whole_run, nb_merge_scoped): one process drives one context
per GPU (ncclCommInitAll), reads / scopes shard over the contexts, and the merged result must equal the oracle's and one
GPU's over the union.  Needs >= 2 GPUs (gpurun --gpus 2); on one GPU the tests skip — the protocol's host logic is then
covered by tests/test_multirank_cpu.py (gloo, world_size 2) and the routed kernels by tests/test_gpu_route.py."""
import json
import threading

import numpy as np
import pytest

import nimble_aligner_b200 as nb
import oracle as orc
import synth

pytestmark = pytest.mark.gpu


def _n_gpus():
    try:
        return nb.lib().nb_device_count()
    except Exception:
        return 0


needs2 = pytest.mark.skipif(_n_gpus() < 2, reason="needs >= 2 GPUs")


def _parallel(fns):
    out, err = [None] * len(fns), [None] * len(fns)

    def run(i):
        try:
            out[i] = fns[i]()
        except Exception as e:   # noqa: BLE001 - re-raised below
            err[i] = e
    th = [threading.Thread(target=run, args=(i,)) for i in range(len(fns))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for e in err:
        if e is not None:
            raise e
    return out


@pytest.fixture(scope="module")
def lib2():
    L = synth.SynthLibrary(seed=1234, n_fam=200, n_all=5, group_on="")
    obj = L.to_json_obj()
    lib = nb.Library.from_text(json.dumps(obj), "unstranded")
    ix = nb.build_index(lib, 8)
    ocfg, oref = orc.parse_reference_library(obj, "unstranded")
    return L, obj, lib, ix, orc.Oracle(ocfg, oref)


@needs2
@pytest.mark.parametrize("world", [2, 4, 8])
def test_whole_run_merge_in_library_matches_oracle_and_one_gpu(lib2, world):
    if _n_gpus() < world:
        pytest.skip("needs %d GPUs" % world)
    L, obj, lib, ix, o = lib2
    per = 60_000
    r1, o1, r2, o2 = synth.pairs(L, 0, per * world, dup_rate=0.3, paired=True, threads=8)
    ref = o.run(r1, o1, r2, o2, threads=8, want_records=False)
    want = {tuple(cs): int(c) for cs, c in ref["scopes"][0]}
    one = nb.Context(ix, lib, device=0)
    one.align_batch(r1, o1, r2, o2, max_read_len=150)
    single = one.counts_raw()
    assert {tuple(cs): int(c) for _, cs, c in one.decode_counts(single)["rows"]} == want
    ctxs = [nb.Context(ix, lib, device=d, max_batch_pairs=25_000) for d in range(world)]
    nb.comm_init_all(ctxs)
    for d, c in enumerate(ctxs):
        c.route_create(world, per * 2)
    for d, c in enumerate(ctxs):
        c.route_attach_ctx(world, d, ctxs, d * per)
    for rep in range(2):   # twice: the second job must start from clean cursors / inboxes
        def shard(d):
            c = ctxs[d]
            c.reset()
            lo, hi = d * per, (d + 1) * per
            a1 = np.concatenate([r1[int(o1[lo]):int(o1[hi])], np.zeros(64, np.uint8)]); b1 = (o1[lo:hi + 1] - o1[lo]).astype(np.uint64)
            a2 = np.concatenate([r2[int(o2[lo]):int(o2[hi])], np.zeros(64, np.uint8)]); b2 = (o2[lo:hi + 1] - o2[lo]).astype(np.uint64)
            c.align_batch(a1, b1, a2, b2, max_read_len=150)
            return c.merge_whole_run()
        raws = _parallel([lambda d=d: shard(d) for d in range(world)])
        for d, raw in enumerate(raws):
            got = {tuple(cs): int(c) for _, cs, c in ctxs[d].decode_counts(raw)["rows"]}
            assert got == want, (rep, d)
            assert raw["n_unique_keys"] == single["n_unique_keys"]
    for c in ctxs:
        c.close()


@needs2
def test_scoped_merge_in_library_sums_the_per_cell_tables(lib2):
    L, obj, lib, ix, o = lib2
    world = 2
    cfgs = dict(trim_target_length=40, trim_strictness=0.9)
    L3 = synth.SynthLibrary(seed=1234, n_fam=200, n_all=5, group_on="", **cfgs)
    lib3 = nb.Library.from_text(json.dumps(L3.to_json_obj()), "unstranded")
    ix3 = nb.build_index(lib3, 8)
    groups = 20_000
    us = [synth.umi_reads(L3, d * groups, groups, seed=2345, threads=8) for d in range(world)]
    n_cells = 8000

    def run(c, u):
        n = u["n_reads"]
        f1 = np.full(n, nb.FLAG_SKIP_ALIGN, dtype=np.uint8); f2 = np.zeros(n, dtype=np.uint8)
        c.align_batch(u["bases"], u["off"], u["bases"], u["off"], q1=u["qual"], q2=u["qual"], flags1=f1, flags2=f2, scope_id=u["scope"], cell_id=u["cell"], max_read_len=91)
    # one GPU over both shards (scope ids of the shards are disjoint runs of the same stream)
    one = nb.Context(ix3, lib3, device=0, max_batch_pairs=30_000)
    for u in us:
        run(one, u)
    want = {(int(c), tuple(cs)): int(k) for c, cs, k in one.counts()["rows"]}
    ctxs = [nb.Context(ix3, lib3, device=d, max_batch_pairs=30_000) for d in range(world)]
    nb.comm_init_all(ctxs)

    def shard(d):
        run(ctxs[d], us[d])
        return ctxs[d].merge_scoped(n_cells)
    raws = _parallel([lambda d=d: shard(d) for d in range(world)])
    for d, raw in enumerate(raws):
        got = {(int(c), tuple(cs)): int(k) for c, cs, k in ctxs[d].decode_counts(raw)["rows"]}
        assert got == want, d
    assert len(want) > 10_000
    # nb_merge_scoped_sharded: every rank keeps its own range of cells; the shards are disjoint and their union is the table
    for c in ctxs:
        c.reset()

    def shard2(d):
        run(ctxs[d], us[d])
        return ctxs[d].merge_scoped(n_cells, sharded=True)
    raws = _parallel([lambda d=d: shard2(d) for d in range(world)])
    union, per = {}, (n_cells + world - 1) // world
    for d, raw in enumerate(raws):
        rows = ctxs[d].decode_counts(raw)["rows"]
        assert all(d * per <= int(c) < (d + 1) * per for c, _, _ in rows) and len(rows) > 1000
        for c, cs, k in rows:
            assert (int(c), tuple(cs)) not in union
            union[(int(c), tuple(cs))] = int(k)
    assert union == want


@needs2
@pytest.mark.parametrize("world", [2, 4])
def test_fastq_driver_on_several_gpus(tmp_path, world):
    """nb_process_fastq_devices: one feeder, one context per GPU, key records routed inside k_pair, nb_merge_whole_run called
    from the C++ driver.  The TSV must be byte-identical to the one-GPU driver's and to the oracle's counts (duplicated pairs
    land on different GPUs: batches are dealt to the contexts in turn)."""
    if _n_gpus() < world:
        pytest.skip("needs %d GPUs" % world)
    L = synth.SynthLibrary(seed=1234, n_fam=40, n_all=5, group_on="")
    obj = L.to_json_obj()
    (tmp_path / "lib.json").write_text(json.dumps(obj))
    n = 60_000
    r1, o1, r2, o2 = synth.pairs(L, 0, n, seed=77, paired=True)
    synth.write_fastq(str(tmp_path / "r1.fastq"), r1, o1, 1)
    synth.write_fastq(str(tmp_path / "r2.fastq"), r2, o2, 2)
    import os
    os.environ["NB_FASTQ_CHUNK"] = "262144"          # many small batches: every context gets dozens of them
    try:
        nb.process_fastq([tmp_path / "r1.fastq", tmp_path / "r2.fastq"], [tmp_path / "lib.json"], [tmp_path / "one.tsv"], num_cores=4)
        nb.process_fastq([tmp_path / "r1.fastq", tmp_path / "r2.fastq"], [tmp_path / "lib.json"], [tmp_path / "many.tsv"], num_cores=4, devices=list(range(world)))
    finally:
        del os.environ["NB_FASTQ_CHUNK"]
    one, many = (tmp_path / "one.tsv").read_text(), (tmp_path / "many.tsv").read_text()
    assert one == many and one.count("\n") > 50
    ocfg, oref = orc.parse_reference_library(obj, "unstranded")
    ref = orc.Oracle(ocfg, oref).run(r1, o1, r2, o2, threads=4, want_records=False)["scopes"][0]
    assert one == "feature\tscore\n" + "".join("\t".join(cs) + "\t%d\n" % k for cs, k in ref)
