"""Peer routing of the whole-run scope (nb_route_*, include/nimble_b200.h): k_pair appends every record whose read_key
another rank owns to that rank's inbox while the alignment runs (NVLink peer stores between GPUs; here the "ranks" are
contexts of one process on one GPU, attached by pointer with nb_route_attach_ctx, so the same kernel path runs on the
driver's single-GPU box).  The merged counts must equal one context over the union of the shards and the CPU oracle:
read_keys duplicated across ranks count once (src/align.rs:576-579, 685), callsets only one rank saw survive the merge."""
import ctypes as C
import json

import numpy as np
import pytest

import nimble_aligner_b200 as nb
import oracle as orc
import synth

pytestmark = pytest.mark.gpu


def _part(r, o, lo, hi):
    oo = (o[lo:hi + 1] - o[lo]).astype(np.uint64)
    return np.concatenate([r[int(o[lo]):int(o[hi])], np.zeros(64, dtype=np.uint8)]), oo


def _callset_rows(ctx):
    nout, gcap = C.c_uint64(0), C.c_uint32(0)
    nb._ck(nb.lib().nb_callsets_export(ctx.h, None, 0, C.byref(nout), C.byref(gcap)))
    rows = np.zeros((max(1, nout.value), 4 + gcap.value), dtype=np.uint32)
    nb._ck(nb.lib().nb_callsets_export(ctx.h, rows.ctypes.data, rows.shape[0], C.byref(nout), C.byref(gcap)))
    return rows[: nout.value]


def _routed_counts(ix, lib, data, world, n, chunk, cfg=None):
    r1, o1, r2, o2 = data
    ctxs = [nb.Context(ix, lib, max_batch_pairs=chunk) for _ in range(world)]
    if cfg is not None:
        for c in ctxs:
            c.set_config(cfg)
    per = n // world
    bounds = [(r * per, (r + 1) * per if r + 1 < world else n) for r in range(world)]
    for c in ctxs:
        c.route_create(world, n // world + 64)
    for r, c in enumerate(ctxs):
        c.route_attach_ctx(world, r, ctxs, bounds[r][0])
    out = []
    for job in range(2):          # second job: the inboxes were emptied by nb_route_import, the routes stay attached
        for r, c in enumerate(ctxs):
            c.reset()
            lo, hi = bounds[r]
            a1, b1 = _part(r1, o1, lo, hi); a2, b2 = _part(r2, o2, lo, hi)
            c.align_batch(a1, b1, a2, b2, max_read_len=150)
        sent = np.stack([c.route_sent() for c in ctxs])     # waits for each "peer's" batches: the barrier / collective of a real job
        assert all(sent[r, r] == 0 for r in range(world))
        rows = [_callset_rows(c) for c in ctxs]
        for r, c in enumerate(ctxs):
            others = np.ascontiguousarray(np.concatenate([rows[q] for q in range(world) if q != r]))
            nb._ck(nb.lib().nb_callsets_import(c.h, others.ctypes.data, others.shape[0]))
        imported = [c.route_import(sent[:, r]) for r, c in enumerate(ctxs)]
        merged, uniq = {}, 0
        for c in ctxs:
            d = c.counts()
            uniq += d["n_unique_keys"]
            for _, cs, k in d["rows"]:
                merged[tuple(cs)] = merged.get(tuple(cs), 0) + int(k)
        out.append((merged, uniq, imported))
    for c in ctxs:
        c.route_detach()
        c.close()
    return out


@pytest.mark.parametrize("world,mm", [(2, 0), (3, 1), (4, 2)])
def test_routed_shards_equal_one_context_and_the_oracle(world, mm):
    n = 60_000
    L = synth.SynthLibrary(seed=4321, n_fam=60, n_all=5, group_on="", num_mismatches=mm)
    obj = L.to_json_obj()
    lib = nb.Library.from_text(json.dumps(obj), "unstranded")
    ix = nb.build_index(lib, 8)
    data = synth.pairs(L, 0, n, seed=99, dup_rate=0.3)
    one = nb.Context(ix, lib, max_batch_pairs=1 << 14)
    one.align_batch(*[data[i] for i in (0, 1, 2, 3)], max_read_len=150)
    d = one.counts()
    want = {tuple(cs): int(k) for _, cs, k in d["rows"]}
    jobs = _routed_counts(ix, lib, data, world, n, 1 << 13)
    for merged, uniq, imported in jobs:
        assert merged == want
        assert uniq == d["n_unique_keys"]
        assert sum(imported) > n // 4                  # most insertable pairs of a rank belong to another rank
    # the oracle over the union (sizes it finishes in seconds)
    ocfg, oref = orc.parse_reference_library(obj, "unstranded")
    ref = orc.Oracle(ocfg, oref).run(*data, threads=8, want_records=False)
    assert want == {tuple(cs): int(c) for cs, c in ref["scopes"][0]}


def test_inbox_overflow_and_misuse_fail_loudly():
    L = synth.SynthLibrary(seed=4321, n_fam=20, n_all=5, group_on="")
    lib = nb.Library.from_text(json.dumps(L.to_json_obj()), "unstranded")
    ix = nb.build_index(lib, 4)
    r1, o1, r2, o2 = synth.pairs(L, 0, 20_000, seed=5)
    a, b = nb.Context(ix, lib), nb.Context(ix, lib)
    with pytest.raises(nb.NbError):
        a.route_import(np.zeros(2, dtype=np.uint64))    # no routes yet
    with pytest.raises(nb.NbError):
        a.route_create(1, 64)                           # world must be >= 2
    a.route_create(2, 64); b.route_create(2, 64)        # far too small for ~10k routed records
    with pytest.raises(nb.NbError):
        a.route_attach_ctx(2, 1, [a, b], 0)             # peers[rank] must be the context itself
    a.route_attach_ctx(2, 0, [a, b], 0); b.route_attach_ctx(2, 1, [a, b], 20_000)
    with pytest.raises(nb.NbError):
        a.route_create(2, 128)                          # attached: detach first
    a.align_batch(r1, o1, r2, o2, max_read_len=150)
    with pytest.raises(nb.NbError) as e:
        a.route_sent()
    assert "inbox" in str(e.value)
    with pytest.raises(nb.NbError) as e:
        a.counts()
    assert "inbox" in str(e.value)
    with pytest.raises(nb.NbError) as e:
        b.route_import(np.array([1000, 0], dtype=np.uint64))   # more than a region holds
    assert "inbox" in str(e.value)
    a.route_detach(); b.route_detach()
