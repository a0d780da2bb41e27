import json, numpy as np, sys, os
sys.path.insert(0, '.')
os.environ["NB_DEBUG_KMAP"] = "1"
import nimble_aligner_b200 as nb, synth
L = synth.SynthLibrary(); obj = L.to_json_obj()
lib = nb.Library.from_text(json.dumps(obj), "unstranded"); ix = nb.build_index(lib, 8)
ctx = nb.Context(ix, lib, count_work=1)
r1,o1,r2,o2 = synth.pairs(L, 0, 1000000)
ctx.align_batch(r1,o1,r2,o2, max_read_len=150)
print(ctx.work_counters())
