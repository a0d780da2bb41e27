#!/usr/bin/env python
"""C3-shaped end-to-end BAM run (not the driver's bench line): synthetic 10x-style unaligned BAM (single-end 91 bp records,
CB/UB tags, quals with Q2 tails) -> nb_process_bam (BGZF inflate, UMI/CB grouping, scoped batches on the GPU, TSV.gz rows).
Prints one JSON line: records/s through the whole driver (file in, file out), and the CPU oracle with the reference's
cost structure on the same records (align stage only: it gets the reads already decoded and grouped)."""
import argparse, gzip, json, os, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nimble_aligner_b200 as nb
import synth


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--groups", type=int, default=1_250_000, help="UMI groups (about 4 records each)")
    ap.add_argument("--cpu-groups", type=int, default=100_000)
    ap.add_argument("--repeat", type=int, default=2)
    ap.add_argument("--keep", action="store_true")
    ap.add_argument("--cli", action="store_true", help="run the nimble binary in a subprocess (reports its peak RSS)")
    a = ap.parse_args()
    cores = os.cpu_count() or 1
    L = synth.SynthLibrary(seed=1234, n_fam=200, n_all=5, group_on="", trim_target_length=40, trim_strictness=0.9)
    tmp = tempfile.mkdtemp(prefix="nb_bam_")
    lib_path = os.path.join(tmp, "lib.json"); json.dump(L.to_json_obj(), open(lib_path, "w"))
    t0 = time.time(); u = synth.umi_reads(L, 0, a.groups, seed=2345, threads=cores); t1 = time.time()
    bam = os.path.join(tmp, "c3.bam"); size = synth.write_umi_bam(bam, u, threads=cores); t2 = time.time()
    n = u["n_reads"]
    out = os.path.join(tmp, "out.tsv.gz")
    best = None
    peak_rss_mb = None
    for _ in range(a.repeat):
        if a.cli:   # through the `nimble` binary in its own process: the whole call as a user makes it, and the driver's own peak RSS
            import resource, subprocess
            del u; u = None
            exe = os.path.join(os.path.dirname(os.path.abspath(nb.__file__)), "nimble")
            if os.path.exists(out):
                os.remove(out)
            env = dict(os.environ, NB_BAM_STATS="1")
            t = time.time(); pr = subprocess.run([exe, "-r", lib_path, "-o", out, "-i", bam, "-c", str(cores)], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, env=env); dt = time.time() - t
            sys.stderr.write(pr.stderr)
            import re   # the driver reports its own VmHWM (ru_maxrss of a forked child starts at the parent's size)
            m = re.search(r"peak RSS (\d+) MB", pr.stderr)
            peak_rss_mb = float(m.group(1)) if m else None
        else:
            t = time.time(); nb.process_bam(bam, [lib_path], [out], strand_filter="unstranded", num_cores=cores); dt = time.time() - t
        best = dt if best is None else min(best, dt)
    osz = os.path.getsize(out)
    rows = 0
    with gzip.open(out, "rb") as f:
        for _ in f:
            rows += 1
    res = {"workload": "C3-shaped BAM: %d single-end 91 bp records in %d (UMI,CB) groups, 8000 cells, 1k-transcript library; %.0f MB BAM" % (n, a.groups, size / 1e6),
           "records_per_s": n / best, "seconds": best, "host_threads": cores, "tsv_rows": rows - 1, "tsv_gz_mb": osz / 1e6,
           "synth_s": round(t1 - t0, 2), "bam_write_s": round(t2 - t1, 2), "driver_peak_rss_mb": peak_rss_mb, "bam_mb": size / 1e6}
    if a.cpu_groups:
        import oracle as orc
        ocfg, oref = orc.parse_reference_library(L.to_json_obj(), "unstranded")
        o = orc.Oracle(ocfg, oref, faithful_cost=True)
        v = synth.umi_reads(L, 0, a.cpu_groups, seed=2345, threads=cores); m = v["n_reads"]
        scope_off = np.concatenate([[0], np.cumsum(v["sizes"], dtype=np.uint64)]).astype(np.uint64)
        skip1 = np.ones(m, dtype=np.uint8); skip2 = np.zeros(m, dtype=np.uint8)   # (dummy, real) pairs like add_dummy_paired_reads
        t = time.time(); o.run(v["bases"], v["off"], v["bases"], v["off"], q1=v["qual"], q2=v["qual"], skip1=skip1, skip2=skip2, scope_off=scope_off, threads=cores, want_records=False); dt = time.time() - t
        res["cpu_oracle"] = {"records_per_s": m / dt, "records": m, "threads": cores, "note": "align stage only (get_calls per scope with the reference's cost structure); no BAM decode, no TSV"}
    print(json.dumps(res))
    if not a.keep:
        import shutil; shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
