#!/usr/bin/env python
"""bench.py — aligned reads/sec of the nimble-aligner hot path on B200 (BASELINE.json metric).

Workload (config.workload "C2"): synthetic 1k-transcript family library (200 families x 5 alleles, seed 1234) and
10 M 2x150 bp read pairs per GPU, FASTQ-mode semantics: one whole-run aggregation scope, counts over
unique read pairs.  A "step" is one complete pass of the hot path (pack -> seed-and-walk map -> pair/orientation ->
de-duplicate -> callset histogram -> counts on the host) over that input.

  value   reads/s with the ASCII reads already resident in HBM (device-timed, CUDA events on the launching stream)
  e2e     reads/s through the C ABI with pinned HOST buffers: H2D copies and the D2H of the counts inside the timed region
  roofline  k_map (dominant kernel): algorithmic bytes per launch / mean launch time (CUDA events inside the library)
  cpu_baseline  the CPU oracle ("port" of the reference; the Rust reference cannot be built here) on a bounded sample

`--impl reference` times the oracle (reference cost structure: string-keyed maps, linear unmap) with all host threads.
N > 1 (torchrun): reads shard over ranks, index replicated; the whole-run de-duplication exchanges 32-byte key
records by key range (all_to_all over NCCL) and the per-callset counts are all-reduced.  Weak scaling: 10 M pairs per GPU.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "aligned reads/sec per box, bit-exact counts vs CPU ref"
UNIT = "reads/s"
SEED = 1234
READ_LEN = 150


def algorithmic_bytes_per_read(work, n_reads_total):
    """DESIGN.md 'Roofline': bytes k_map must touch per read = packed read words + 16 B per hash probe + 32 B node
    record and c_v/4 unitig bytes per visited unitig + 4 B per colour id touched + 32 B result record."""
    packed = 8 * ((READ_LEN + 31) // 32)
    tot = packed * n_reads_total + 16 * work["probes"] + 32 * work["nodes"] + work["bases"] / 4.0 + 4 * work["colour_elems"] + 32 * n_reads_total
    return tot / n_reads_total


class ClockSampler:
    """Polls NVML (same counters as the nvidia-smi clocks line of B200_PROFILING.md) every ~2 ms from a thread, so that
    even a sub-second timed region gets samples DURING it."""
    REASONS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "hw_power_brake": 0x80}

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.stop_flag, self.thread, self.h = gpu_index, [], False, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        except Exception:
            self.h = None

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((time.time(), sm, rs))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.h is not None:
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()

    def stop(self, t0=None, t1=None):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=1.0)
        rows = [r for r in self.rows if (t0 is None or r[0] >= t0) and (t1 is None or r[0] <= t1)]
        mx = None
        try:
            mx = float(self.nv.nvmlDeviceGetMaxClockInfo(self.h, self.nv.NVML_CLOCK_SM))
        except Exception:
            pass
        reasons = set()
        for _, _, rs in rows:
            for name, bit in self.REASONS.items():
                if rs & bit:
                    reasons.add(name)
        return {"sm_mhz": float(np.median([r[1] for r in rows])) if rows else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(rows)}


def build_library(mismatches=0):
    import synth
    import nimble_aligner_b200 as nb
    L = synth.SynthLibrary(seed=SEED, n_fam=200, n_all=5, group_on="", num_mismatches=mismatches)
    obj = L.to_json_obj()
    lib = nb.Library.from_text(json.dumps(obj), "unstranded")
    return L, obj, lib


def run_reference(args):
    """CPU arm: the oracle with the reference's cost structure on a bounded sample, all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle as orc
    import synth
    L = synth.SynthLibrary(seed=SEED, n_fam=200, n_all=5, group_on="", num_mismatches=args.mismatches)
    ocfg, oref = orc.parse_reference_library(L.to_json_obj(), "unstranded")
    o = orc.Oracle(ocfg, oref, faithful_cost=True)
    cores = os.cpu_count() or 1
    n = args.ref_pairs
    r1, o1, r2, o2 = synth.pairs(L, 0, n, seed=SEED, threads=cores)
    for _ in range(args.warmup):
        o.run(r1, o1[:n // 8 + 1], r2, o2[:n // 8 + 1], threads=cores, shard_single=True, want_records=False)
    t0 = time.time()
    for _ in range(args.steps):
        o.run(r1, o1, r2, o2, threads=cores, shard_single=True, want_records=False)
    dt = (time.time() - t0) / args.steps
    v = 2 * n / dt
    sample = "%d of the 10M C2 pairs per step (pairs 0..%d of the same seeded stream), %d threads over contiguous shards" % (n, n, cores)
    emit_json({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                      "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/u64 integer (f64 thresholds)",
                      "data": "synthetic", "config": {"workload": "C2: 1k-transcript family library x 2x150 bp pairs, FASTQ-mode scope", "sample_pairs": n, "num_mismatches": args.mismatches},
                      "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
                      "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})


_JSON_OUT = None


def guard_stdout():
    """stdout carries exactly one JSON line: everything else that any library prints there (NCCL's version banner,
    torchrun notices of child ranks) is sent to stderr; emit_json writes to the saved descriptor."""
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit_json(obj):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(obj) + "\n")
    out.flush()


def main():
    guard_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs", type=int, default=10_000_000, help="read pairs per GPU per step (C2: 10M)")
    ap.add_argument("--ref-pairs", type=int, default=2_000_000, help="pairs per step of the CPU reference arm / cpu_baseline sample")
    ap.add_argument("--chunk", type=int, default=1 << 20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--verify", action="store_true", help="N>1: check the merged counts against one GPU over the union of shards")
    ap.add_argument("--merge", default="p2p", choices=["p2p", "nccl"], help="N>1: key records routed inside k_pair over NVLink peer stores (default), or exchanged with an NCCL all-to-all when the job ends")
    ap.add_argument("--mismatches", type=int, default=0, help="num_mismatches of the library config (C5 sweeps 0/1/2)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = max(args.warmup, 1)
    if args.impl == "reference":
        return run_reference(args)

    rank, world, local_rank = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    numa_cpus = None
    if world > 1:   # one rank per GPU: keep each rank's pinned batches on its GPU's NUMA node (N=1 stays unbound: its cpu_baseline leg uses every core)
        from nimble_aligner_b200.multigpu import bind_to_gpu_numa_node
        numa_cpus = bind_to_gpu_numa_node(local_rank)
    import torch
    import torch.distributed as dist
    import nimble_aligner_b200 as nb
    import synth
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
            os.environ.pop("NCCL_DEBUG")             # at these two levels NCCL prints its version banner on stdout
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n = args.pairs
    cores = os.cpu_count() or 1
    L, obj, lib = build_library(args.mismatches)
    t0 = time.time()
    ix = nb.build_index(lib, max(1, cores // max(1, world)))
    index_build_s = time.time() - t0
    # one explicit CUDA stream shared by torch (events, NCCL ordering) and the library (kernels, D2H); the legacy default
    # stream has handle 0, which nb_ctx_create reads as "create your own"
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0
    ctx = nb.Context(ix, lib, device=local_rank, stream=stream, max_batch_pairs=args.chunk)
    # ---- inputs: this rank's shard of the seeded stream, in pinned host memory and resident in HBM
    pair_base = rank * n
    o1 = np.zeros(n + 1, dtype=np.uint64); o2 = np.zeros(n + 1, dtype=np.uint64)
    synth.lib().synth_pair_offsets(SEED, pair_base, n, READ_LEN, 0.1, o1.ctypes.data, o2.ctypes.data, cores)
    h1 = torch.empty(int(o1[-1]) + 64, dtype=torch.uint8).pin_memory(); h2 = torch.empty(int(o2[-1]) + 64, dtype=torch.uint8).pin_memory()
    synth.pairs(L, pair_base, n, seed=SEED, threads=max(1, cores // max(1, world)), out=(h1.numpy(), h2.numpy()))
    ho1, ho2 = torch.from_numpy(o1.astype(np.int64)).pin_memory(), torch.from_numpy(o2.astype(np.int64)).pin_memory()
    d1, d2, do1, do2 = h1.cuda(), h2.cuda(), ho1.cuda(), ho2.cuda()
    n_reads = 2 * n

    from nimble_aligner_b200.multigpu import merge_across_ranks, DeviceShard, setup_routes
    routed = world > 1 and args.merge == "p2p" and setup_routes(ctx, torch, dist, rank, world, "cuda", pair_base, (n + n // 2) // world + 4096)
    shard = DeviceShard(ctx, nb, torch, pair_base, n, routed=routed) if world > 1 else None   # merge buffers are allocated once, outside the job

    def step_device():
        ctx.reset()
        ctx.align_batch(d1, do1, d2, do2, n_pairs=n, max_read_len=READ_LEN, location=nb.NB_MEM_DEVICE)
        if world > 1:
            return merge_across_ranks(shard, torch, dist, rank, world, "cuda", routed=routed)
        raw = ctx.counts_raw()          # nb_counts_finalize: the job's result (group indices + counts) on the host
        return raw, raw["n_unique_keys"]

    def step_host():
        ctx.reset()
        # chunked submissions from pinned host memory; copies and kernels are stream-ordered inside the library
        b = nb.Batch(n, nb.NB_MEM_HOST, READ_LEN, h1.data_ptr(), ho1.data_ptr(), h2.data_ptr(), ho2.data_ptr(), None, None, None, None, None, None)
        import ctypes as C
        nb._ck(nb.lib().nb_align_batch(ctx.h, C.byref(b), None, None))
        if world > 1:
            return merge_across_ranks(shard, torch, dist, rank, world, "cuda", routed=routed)
        raw = ctx.counts_raw()          # nb_counts_finalize: the job's result (group indices + counts) on the host
        return raw, raw["n_unique_keys"]

    def timed(fn, steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.time()
        e0.record()
        out = None
        for _ in range(steps):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = max(e0.elapsed_time(e1), 0.0)
        wall = (time.time() - w0) * 1e3
        # the step ends with host-side work (count read-back + sort) after the last kernel: take the larger of the two clocks
        ms = max(ms, wall)
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps, out

    for _ in range(args.warmup):
        step_device()
    ctx.kernel_stats(reset=True)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    tw0 = time.time()
    ms_dev, (counts_dev, uniq_dev) = timed(step_device, args.steps)
    tw1 = time.time()
    ks = ctx.kernel_stats(reset=True)
    clocks = sampler.stop(tw0, tw1) if rank == 0 else None
    step_host()
    ms_host, (counts_host, uniq_host) = timed(step_host, args.steps)
    def decode(raw):   # names are decoded outside the timed region (a C host would print them straight into the TSV)
        d = ctx.decode_counts(raw)
        if "dense_counts" in raw:
            return {tuple(cs): int(c) for cs, c in zip(d["callsets"], raw["dense_counts"].tolist()) if c}
        return {tuple(cs): int(c) for _, cs, c in d["rows"]}
    counts_dev, counts_host = decode(counts_dev), decode(counts_host)
    if counts_host != counts_dev:
        diff = [(k, counts_dev.get(k), counts_host.get(k)) for k in set(counts_dev) | set(counts_host) if counts_dev.get(k) != counts_host.get(k)]
        print("rank %d: %d callsets differ (dev total %d, host total %d, %d vs %d callsets); e.g. %s" % (rank, len(diff), sum(counts_dev.values()), sum(counts_host.values()), len(counts_dev), len(counts_host), diff[:3]), file=sys.stderr)
    assert counts_host == counts_dev, "host-fed and device-resident runs disagree"
    if args.verify and world > 1 and rank == 0:
        # the merged multi-GPU counts must equal one GPU processing the union of all ranks' shards
        vo1 = np.zeros(n * world + 1, dtype=np.uint64); vo2 = np.zeros(n * world + 1, dtype=np.uint64)
        synth.lib().synth_pair_offsets(SEED, 0, n * world, READ_LEN, 0.1, vo1.ctypes.data, vo2.ctypes.data, cores)
        v1 = np.empty(int(vo1[-1]) + 64, dtype=np.uint8); v2 = np.empty(int(vo2[-1]) + 64, dtype=np.uint8)
        synth.pairs(L, 0, n * world, seed=SEED, threads=cores, out=(v1, v2))
        vctx = nb.Context(ix, lib, device=local_rank, max_batch_pairs=args.chunk)
        vctx.align_batch(v1, vo1, v2, vo2, max_read_len=READ_LEN)
        union = {tuple(cs): int(c) for _, cs, c in vctx.counts()["rows"]}
        assert union == counts_dev, "multi-GPU merge differs from the single-GPU result over the union of shards"
        print("verify: %d-GPU merged counts == single-GPU counts over the union (%d callsets)" % (world, len(union)), file=sys.stderr)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    value = n_reads * world / (ms_dev / 1e3)
    e2e = n_reads * world / (ms_host / 1e3)
    h2d = int(o1[-1]) + int(o2[-1]) + 2 * 8 * (n + 1)
    d2h = 16 * len(counts_dev) + (2 + 16) * 4 * len(counts_dev) + 2 * 96   # compacted count rows + callset rows + counters read by nb_counts_finalize
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/u64 integer (f64 thresholds)", "data": "synthetic",
           "config": {"workload": "C2: synthetic 1k-transcript family library (200x5, seed 1234) x %d 2x150 bp pairs per GPU, FASTQ-mode whole-run scope" % n,
                      "num_mismatches": args.mismatches, "merge": ("p2p-routed (k_pair stores key records into their owners' inboxes over NVLink)" if routed else "nccl all-to-all at job end") if world > 1 else "none (one GPU)",
                      "numa_bind": ("each rank bound to the %d CPUs local to its GPU" % numa_cpus) if numa_cpus else "none",
                      "pairs_per_gpu": n, "reads_per_step": n_reads * world, "chunk_pairs": args.chunk, "l2": "inputs (%.1f GB ASCII per step) exceed the 126 MB L2" % ((int(o1[-1]) + int(o2[-1])) / 1e9),
                      "index_device_mb": ix.stats()["device_bytes"] / 1e6, "index_build_s": index_build_s, "unique_pair_keys": int(uniq_dev), "callsets_counted": len(counts_dev)},
           "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_host},
           "gpu_launches": ks["launches"], "clocks": clocks,
           "k_map_ms_per_launch": ks["map_ms"] / max(1, ks["map_launches"]), "k_map_reads_per_launch": ks["map_reads"] / max(1, ks["map_launches"])}
    # ---- cpu baseline + roofline (rank 0, N=1 only)
    if world == 1 and not args.no_cpu_baseline:
        import oracle as orc
        ocfg, oref = orc.parse_reference_library(obj, "unstranded")
        o = orc.Oracle(ocfg, oref, faithful_cost=True)
        m = min(args.ref_pairs, n)
        r1, oo1, r2, oo2 = synth.pairs(L, 0, m, seed=SEED, threads=cores)
        t0 = time.time()
        ref = o.run(r1, oo1, r2, oo2, threads=cores, shard_single=True, want_records=False)
        dt = time.time() - t0
        out["cpu_baseline"] = {"value": 2 * m / dt, "unit": UNIT, "cores": cores, "kind": "port",
                               "sample": "first %d of the step's %d pairs, oracle with the reference's cost structure, %d threads over contiguous shards" % (m, n, cores)}
        bpr = algorithmic_bytes_per_read(ref["work"], 2 * m)
        launch_ms = ks["map_ms"] / max(1, ks["map_launches"])
        reads_per_launch = ks["map_reads"] / max(1, ks["map_launches"])
        achieved = bpr * reads_per_launch / (launch_ms / 1e3) / 1e9
        peak, peak_src = 6650.0, "fallback"
        try:
            peak, peak_src = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
        except Exception:
            pass
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "k_map_traffic.json")))["dram_bytes_per_launch"]
        except Exception:
            pass
        out["roofline"] = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                           "kernel": "map stage = k_seed + k_walk (one launch pair per chunk; events bracket the pair)", "peak_source": peak_src, "algorithmic_bytes_per_read": bpr, "reads_per_launch": reads_per_launch,
                           "launch_ms": launch_ms, "kernel_share_of_step": ks["map_ms"] / (ms_dev * args.steps),
                           "work_per_read": {k: ref["work"][k] / (2.0 * m) for k in ("probes", "nodes", "bases", "colour_elems")}}
    emit_json(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
