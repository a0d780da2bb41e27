import json, numpy as np, sys
sys.path.insert(0, '.')
import nimble_aligner_b200 as nb, oracle as orc, synth
L = synth.SynthLibrary()
obj = L.to_json_obj()
lib = nb.Library.from_text(json.dumps(obj), "unstranded")
ix = nb.build_index(lib, 8)
ctx = nb.Context(ix, lib)
u = synth.umi_reads(L, 0, 600)
n = u["n_reads"]
print("n", n, "scope head", u["scope"][:20], "sizes", u["sizes"][:5])
flags1 = np.full(n, 1, dtype=np.uint8)
reads, pairs = ctx.align_batch(u["bases"], u["off"], u["bases"], u["off"], q1=u["qual"], q2=u["qual"], flags1=flags1, flags2=None, scope_id=u["scope"], want_reads=True, want_pairs=True)
res = ctx.counts()
print("rows", len(res["rows"]), res["rows"][:5], "unique", res["n_unique_keys"], "pairs_seen", res["n_pairs_seen"])
print("row_scope", res["row_scope"][:20], res["row_count"][:20])
print("pairs callset head", pairs["callset"][:20], pairs["triage"][:20], pairs["insertable"][:20])
