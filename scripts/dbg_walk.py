#!/usr/bin/env python
"""k_walk lane statistics (COUNT_WORK build of the kernel): iterations, lanes walking / re-seeding per iteration."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["NB_DEBUG_KMAP"] = "1"
import numpy as np
import nimble_aligner_b200 as nb, synth
L = synth.SynthLibrary(seed=1234, n_fam=200, n_all=5, group_on="")
lib = nb.Library.from_text(json.dumps(L.to_json_obj()), "unstranded")
ix = nb.build_index(lib, 8)
n = 1_000_000
r1, o1, r2, o2 = synth.pairs(L, 0, n, seed=1234, threads=8)
ctx = nb.Context(ix, lib, max_batch_pairs=1 << 20, count_work=1)
ctx.align_batch(r1, o1, r2, o2, max_read_len=150)
print(ctx.work_counters())
