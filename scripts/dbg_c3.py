import json, os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nimble_aligner_b200 as nb, synth
import bench
L = bench.c3_library()
lib = nb.Library.from_text(json.dumps(L.to_json_obj()), "unstranded")
ix = nb.build_index(lib, 8)
u = synth.umi_reads(L, 0, 600000, seed=2345, threads=8)
n = u["n_reads"]; chunk = 1 << 20
pin = lambda x: torch.from_numpy(x).pin_memory()
bases, qual, off = pin(u["bases"]), pin(u["qual"]), pin(u["off"].astype(np.int64))
scope, cell = pin(u["scope"].astype(np.int32)), pin(u["cell"].astype(np.int32))
f1, f2 = pin(np.full(n, 1, dtype=np.uint8)), pin(np.zeros(n, dtype=np.uint8))
dv = [t.cuda() for t in (bases, qual, off, scope, cell, f1, f2)]
ctx = nb.Context(ix, lib, max_batch_pairs=chunk, agg_slots=1 << 24)
cuts, sc = [0], u["scope"]
while cuts[-1] < n:
    p1 = min(n, cuts[-1] + chunk)
    while p1 < n and p1 > cuts[-1] + 1 and sc[p1] == sc[p1 - 1]:
        p1 -= 1
    cuts.append(p1)
print("n", n, "cuts", cuts, "scope monotone", bool(np.all(np.diff(sc.astype(np.int64)) >= 0)))
def dev():
    ctx.reset()
    db, dq, do, dsc, dce, df1, df2 = dv
    for a, b_ in zip(cuts[:-1], cuts[1:]):
        bt = nb.Batch(b_ - a, nb.NB_MEM_DEVICE, 91, db.data_ptr(), do.data_ptr() + 8 * a, db.data_ptr(), do.data_ptr() + 8 * a, dq.data_ptr(), dq.data_ptr(),
                      df1.data_ptr() + a, df2.data_ptr() + a, dsc.data_ptr() + 4 * a, dce.data_ptr() + 4 * a)
        nb._ck(nb.lib().nb_align_batch(ctx.h, C.byref(bt), None, None))
    return ctx.counts_raw()
def host():
    ctx.reset()
    bt = nb.Batch(n, nb.NB_MEM_HOST, 91, bases.data_ptr(), off.data_ptr(), bases.data_ptr(), off.data_ptr(), qual.data_ptr(), qual.data_ptr(),
                  f1.data_ptr(), f2.data_ptr(), scope.data_ptr(), cell.data_ptr())
    nb._ck(nb.lib().nb_align_batch(ctx.h, C.byref(bt), None, None))
    return ctx.counts_raw()
for name, fn in (("dev", dev), ("host", host), ("dev", dev), ("host", host)):
    r = fn()
    print(name, "rows", len(r["row_count"]), "sum", int(r["row_count"].sum()), "callsets", len(r["callset_off"]) - 1, "uniq", r["n_unique_keys"], "pairs", r["n_pairs_seen"],
          "hash", hash((r["row_scope"].tobytes(), r["row_callset"].tobytes(), r["row_count"].tobytes())))
