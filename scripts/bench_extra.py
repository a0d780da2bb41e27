#!/usr/bin/env python
"""Secondary measurements (not the driver's bench line): the other BASELINE.json config shapes on one GPU.
  --workload c3   10x-style single-end records (L=91, quals, (UMI,CB) scopes, dummy mates, MAXINFO trim), per-cell counts
  --workload c4s  scaled-down C4: a library whose index is far larger than the 126 MB L2 (HBM-bound probes), single-end 150 bp
Prints one JSON line per run with reads/s device-timed from pinned host buffers (e2e) and k_map's launch time."""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import nimble_aligner_b200 as nb
import synth


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--reads", type=int, default=20_000_000)
    ap.add_argument("--families", type=int, default=8000)
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    rank, world, local_rank = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    cores = max(1, (os.cpu_count() or 1) // world)
    torch.cuda.set_device(local_rank)
    if world > 1:   # C3 only: (UMI, CB) scopes shard over ranks, the per-cell tables are combined at the end of the job
        import torch.distributed as dist
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
            os.environ.pop("NCCL_DEBUG")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        assert a.workload == "c3"
    T = {"align": 0.0}
    s = torch.cuda.Stream(); torch.cuda.set_stream(s)
    if a.workload == "c3":
        L = synth.SynthLibrary(seed=1234, n_fam=200, n_all=5, group_on="", trim_target_length=40, trim_strictness=0.9)
        lib = nb.Library.from_text(json.dumps(L.to_json_obj()), "unstranded")
        ix = nb.build_index(lib, cores)
        groups = a.reads // 4
        u = synth.umi_reads(L, rank * groups, groups, seed=2345, threads=cores)   # weak scaling: every rank takes its own run of (UMI, CB) groups
        n = u["n_reads"]
        pin = lambda x: torch.from_numpy(x).pin_memory()
        bases, qual, off = pin(u["bases"]), pin(u["qual"]), pin(u["off"].astype(np.int64))
        scope, cell = pin(u["scope"].astype(np.int32)), pin(u["cell"].astype(np.int32))
        f1 = pin(np.full(n, nb.FLAG_SKIP_ALIGN, dtype=np.uint8)); f2 = pin(np.zeros(n, dtype=np.uint8))
        ctx = nb.Context(ix, lib, device=local_rank, stream=s.cuda_stream, max_batch_pairs=1 << 20, agg_slots=1 << 24)
        import ctypes as C
        if world > 1:
            from nimble_aligner_b200.multigpu import merge_scoped_across_ranks, DeviceShard
            shard = DeviceShard(ctx, nb, torch, 0, 0)
        def step():
            ctx.reset()
            b = nb.Batch(n, nb.NB_MEM_HOST, 91, bases.data_ptr(), off.data_ptr(), bases.data_ptr(), off.data_ptr(), qual.data_ptr(), qual.data_ptr(),
                         f1.data_ptr(), f2.data_ptr(), scope.data_ptr(), cell.data_ptr())
            nb._ck(nb.lib().nb_align_batch(ctx.h, C.byref(b), None, None))
            ctx.sync(); T["align"] += time.time()
            if world > 1:
                raw, cells, css, vals = merge_scoped_across_ranks(shard, torch, dist, rank, world, "cuda", 8000)
                raw = dict(raw); raw["row_count"] = vals
                return raw
            return ctx.counts_raw()
        desc = "C3-shaped: %d single-end 91 bp records with quals in %d (UMI,CB) scopes, 8000 cells, 1k-transcript library" % (n, len(u["sizes"]))
        h2d = int(u["off"][-1]) * 2 + n * (8 + 4 + 4 + 2)
    else:
        L = synth.SynthLibrary(seed=3456, n_fam=a.families, n_all=5, group_on="")
        t0 = time.time(); lib = nb.Library.from_text(json.dumps(L.to_json_obj()), "unstranded"); t1 = time.time()
        ix = nb.build_index(lib, cores, device=0); t2 = time.time()
        n = a.reads
        r1, o1, _, _ = synth.pairs(L, 0, n, seed=3456, paired=False, threads=cores)
        hb, ho = torch.from_numpy(r1).pin_memory(), torch.from_numpy(o1.astype(np.int64)).pin_memory()
        ctx = nb.Context(ix, lib, stream=s.cuda_stream, max_batch_pairs=1 << 20, callset_slots=1 << 22, agg_slots=1 << 23)
        import ctypes as C
        def step():
            ctx.reset()
            b = nb.Batch(n, nb.NB_MEM_HOST, 150, hb.data_ptr(), ho.data_ptr(), None, None, None, None, None, None, None, None)
            nb._ck(nb.lib().nb_align_batch(ctx.h, C.byref(b), None, None))
            ctx.sync(); T["align"] += time.time()
            return ctx.counts_raw()
        st = ix.stats()
        desc = "C4-scaled: %d transcripts, index %.0f MB on the device (%d k-mers; library parse %.1fs, GPU index build %.1fs), %d single-end 150 bp reads" % (5 * a.families, st["device_bytes"] / 1e6, st["n_kmers"], t1 - t0, t2 - t1, n)
        h2d = int(o1[-1]) + 8 * n
    for _ in range(2):
        raw = step()
    ctx.kernel_stats(reset=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(); t0 = time.time(); T["align"] = 0.0; starts = 0.0
    for _ in range(a.steps):
        starts += time.time()
        raw = step()
    torch.cuda.synchronize(); dt = (time.time() - t0) / a.steps
    if world > 1:   # max over ranks; rank 0 reports the whole job
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda"); dist.all_reduce(tt, op=dist.ReduceOp.MAX); dt = float(tt.item())
        n = n * world
        if rank != 0:
            dist.destroy_process_group(); return
        desc += " per rank x %d ranks (scopes sharded, per-cell tables merged)" % world
    ks = ctx.kernel_stats()
    align_s = (T["align"] - starts) / a.steps
    roof = None
    if a.workload != "c3":
        # algorithmic bytes per read from the device's own work counters (equal to the oracle's on every parity test):
        # packed read + 16 B per probe + 32 B per visited unitig + compared bases / 4 + 4 B per colour id + 32 B result
        m = min(n, 1 << 20)
        wctx = nb.Context(ix, lib, stream=s.cuda_stream, max_batch_pairs=1 << 20, callset_slots=1 << 22, agg_slots=1 << 23, count_work=1)
        wb = nb.Batch(m, nb.NB_MEM_HOST, 150, hb.data_ptr(), ho.data_ptr(), None, None, None, None, None, None, None, None)
        nb._ck(nb.lib().nb_align_batch(wctx.h, C.byref(wb), None, None)); wctx.sync()
        w = wctx.work_counters()
        bpr = (8 * 5 * m + 16 * w["probes"] + 32 * w["nodes"] + w["bases"] / 4.0 + 4 * w["colour_elems"] + 32 * m) / m
        ms = ks["map_ms"] / max(1, ks["map_launches"]); rpl = ks["map_reads"] / max(1, ks["map_launches"])
        peak = 6458.4
        try:
            peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
        except Exception:
            pass
        roof = {"bound": "hbm", "algorithmic_bytes_per_read": bpr, "achieved_gb_s": bpr * rpl / (ms / 1e3) / 1e9, "peak_gb_s": peak, "frac": bpr * rpl / (ms / 1e3) / 1e9 / peak,
                "work_per_read": {k: w[k] / m for k in w}}
    print(json.dumps({"workload": desc, "map_stage_roofline": roof, "e2e_reads_per_s": n / dt, "align_only_reads_per_s": n / align_s, "align_ms": align_s * 1e3, "finalize_ms": (dt - align_s) * 1e3, "ms_per_step": dt * 1e3, "h2d_gb_per_s": h2d / dt / 1e9, "count_rows": int(len(raw["row_count"])),
                      "unique_keys": int(raw["n_unique_keys"]), "k_map_ms_per_launch": ks["map_ms"] / max(1, ks["map_launches"]), "k_map_reads_per_launch": ks["map_reads"] / max(1, ks["map_launches"]),
                      "k_map_share": ks["map_ms"] / (dt * 1e3 * a.steps), "launches_per_step": ks["launches"] / a.steps}))


if __name__ == "__main__":
    main()
