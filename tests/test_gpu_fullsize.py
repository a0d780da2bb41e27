"""Size-independent properties at BASELINE.json's full C2 size (10 M pairs, through the C ABI), where the oracle would take
minutes: submission order, batch splitting and resubmission must not change a single count, and the count table must be
consistent with the number of unique read_keys.  The oracle pins a prefix of the same stream."""
import json

import numpy as np
import pytest

import nimble_aligner_b200 as nb
import oracle as orc
import synth

pytestmark = pytest.mark.gpu
N = 10_000_000


@pytest.fixture(scope="module")
def world():
    L = synth.SynthLibrary(seed=1234, n_fam=200, n_all=5, group_on="")
    obj = L.to_json_obj()
    lib = nb.Library.from_text(json.dumps(obj), "unstranded")
    ix = nb.build_index(lib, 16, device=0)
    r1, o1, r2, o2 = synth.pairs(L, 0, N, seed=1234, threads=16)
    return L, obj, lib, ix, (r1, o1, r2, o2)


def _counts(ctx):
    return {tuple(cs): int(c) for _, cs, c in ctx.counts()["rows"]}


def test_full_size_counts_are_independent_of_chunking_and_resubmission(world):
    L, obj, lib, ix, (r1, o1, r2, o2) = world
    ctx = nb.Context(ix, lib, max_batch_pairs=1 << 20)
    ctx.align_batch(r1, o1, r2, o2, max_read_len=150)
    raw = ctx.counts_raw()
    a = _counts(ctx)
    assert sum(a.values()) <= raw["n_unique_keys"] <= N and len(a) > 2000 and sum(a.values()) > N // 3
    # other chunk size
    ctx2 = nb.Context(ix, lib, max_batch_pairs=700_001)
    ctx2.align_batch(r1, o1, r2, o2, max_read_len=150)
    assert _counts(ctx2) == a and ctx2.counts_raw()["n_unique_keys"] == raw["n_unique_keys"]
    # every pair submitted twice (two calls): duplicates of a read_key vote once (src/align.rs:576-579, 685)
    ctx2.reset()
    ctx2.align_batch(r1, o1, r2, o2, max_read_len=150)
    ctx2.align_batch(r1, o1, r2, o2, max_read_len=150)
    assert _counts(ctx2) == a and ctx2.counts_raw()["n_unique_keys"] == raw["n_unique_keys"]


def test_full_size_counts_are_independent_of_submission_order(world):
    L, obj, lib, ix, (r1, o1, r2, o2) = world
    ctx = nb.Context(ix, lib, max_batch_pairs=1 << 20)
    ctx.align_batch(r1, o1, r2, o2, max_read_len=150)
    a = _counts(ctx)
    # second half first, as two calls into the same whole-run scope
    h = N // 2
    def part(r, o, lo, hi):
        oo = (o[lo:hi + 1] - o[lo]).astype(np.uint64)
        return np.concatenate([r[int(o[lo]):int(o[hi])], np.zeros(64, dtype=np.uint8)]), oo
    ctx.reset()
    for lo, hi in ((h, N), (0, h)):
        a1, b1 = part(r1, o1, lo, hi); a2, b2 = part(r2, o2, lo, hi)
        ctx.align_batch(a1, b1, a2, b2, max_read_len=150)
    assert _counts(ctx) == a


def test_full_size_prefix_matches_the_oracle(world):
    L, obj, lib, ix, (r1, o1, r2, o2) = world
    m = 200_000
    ocfg, oref = orc.parse_reference_library(obj, "unstranded")
    o = orc.Oracle(ocfg, oref)
    ref = o.run(r1, o1[: m + 1], r2, o2[: m + 1], threads=16, want_records=False)
    ctx = nb.Context(ix, lib)
    ctx.align_batch(r1, o1[: m + 1], r2, o2[: m + 1], max_read_len=150)
    got = sorted((cs, c) for cs, c in _counts(ctx).items())
    want = sorted((tuple(cs), int(c)) for cs, c in ref["scopes"][0])
    assert got == want
