"""One process, several k_map configurations back to back (for a single ncu invocation): NB_KMAP / NB_WALK_SMEM are re-read per launch."""
import json, numpy as np, sys, os
sys.path.insert(0, '.')
import nimble_aligner_b200 as nb, synth
L = synth.SynthLibrary(seed=1234, n_fam=200, n_all=5, group_on=""); obj = L.to_json_obj()
lib = nb.Library.from_text(json.dumps(obj), "unstranded"); ix = nb.build_index(lib, 8)
ctx = nb.Context(ix, lib)
r1, o1, r2, o2 = synth.pairs(L, 0, 1000000, seed=1234)
for env in ({"NB_KMAP": "fused"}, {"NB_KMAP": "split", "NB_WALK_SMEM": "0"}, {"NB_KMAP": "split", "NB_WALK_SMEM": "1"}):
    os.environ.update(env)
    for _ in range(2):
        ctx.reset(); ctx.align_batch(r1, o1, r2, o2, max_read_len=150)
    print(env, ctx.kernel_stats(reset=True))
