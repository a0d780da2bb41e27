"""Where does a scoped (BAM-shaped) nb_align_batch call spend its time?  Same batch with / without per-read and per-pair
outputs, pinned vs pageable inputs, shared vs separate mate buffers."""
import json, os, sys, time, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import nimble_aligner_b200 as nb
import synth
cores = os.cpu_count() or 1
L = synth.SynthLibrary(seed=1234, n_fam=200, n_all=5, group_on="", trim_target_length=40, trim_strictness=0.9)
lib = nb.Library.from_text(json.dumps(L.to_json_obj()), "unstranded")
ix = nb.build_index(lib, cores)
u = synth.umi_reads(L, 0, 262144, seed=2345, threads=cores)
n = u["n_reads"]
ctx = nb.Context(ix, lib, max_batch_pairs=1 << 21)
def run(pinned, outputs, shared, reps=3):
    mk = (lambda x: torch.from_numpy(x).pin_memory()) if pinned else (lambda x: torch.from_numpy(x.copy()))
    bases, qual, off = mk(u["bases"]), mk(u["qual"]), mk(u["off"].astype(np.int64))
    b2, q2, o2 = (bases, qual, off) if shared else (mk(u["bases"]), mk(u["qual"]), mk(u["off"].astype(np.int64)))
    scope = mk(u["scope"].astype(np.int32))
    f1 = mk(np.full(n, nb.FLAG_SKIP_ALIGN, dtype=np.uint8)); f2 = mk(np.zeros(n, dtype=np.uint8))
    rr = np.zeros(2 * n * 16, dtype=np.uint8); pr = np.zeros(n * 24, dtype=np.uint8)
    best = 1e9
    for _ in range(reps):
        ctx.reset()
        b = nb.Batch(n, nb.NB_MEM_HOST, 91, bases.data_ptr(), off.data_ptr(), b2.data_ptr(), o2.data_ptr(), qual.data_ptr(), q2.data_ptr(),
                     f1.data_ptr(), f2.data_ptr(), scope.data_ptr(), None)
        t = time.time()
        nb._ck(nb.lib().nb_align_batch(ctx.h, C.byref(b), rr.ctypes.data if outputs else None, pr.ctypes.data if outputs else None))
        ctx.sync(); t1 = time.time()
        ctx.counts_raw(); t2 = time.time()
        best = min(best, t1 - t)
    print("pinned=%d outputs=%d shared_mate=%d: align %.1f ms (%d pairs), finalize %.1f ms" % (pinned, outputs, shared, best * 1e3, n, (t2 - t1) * 1e3), flush=True)
for pinned in (1, 0):
    for outputs in (0, 1):
        for shared in (1, 0):
            run(pinned, outputs, shared)
