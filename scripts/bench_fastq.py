#!/usr/bin/env python
"""C2 end to end through the FASTQ driver (not the driver's bench line): two FASTQ files (plain and .gz) -> nb_process_fastq
-> TSV.  Prints one JSON line per input kind with reads/s through the whole call (file in, file out)."""
import argparse, json, os, sys, tempfile, time, shutil
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nimble_aligner_b200 as nb
import synth


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=4_000_000)
    a = ap.parse_args()
    cores = os.cpu_count() or 1
    L = synth.SynthLibrary(seed=1234, n_fam=200, n_all=5, group_on="")
    tmp = tempfile.mkdtemp(prefix="nb_fq_")
    lib_path = os.path.join(tmp, "lib.json"); json.dump(L.to_json_obj(), open(lib_path, "w"))
    r1, o1, r2, o2 = synth.pairs(L, 0, a.pairs, seed=1234, threads=cores)
    for ext in (".fastq", ".fastq.gz"):
        f1, f2 = os.path.join(tmp, "r1" + ext), os.path.join(tmp, "r2" + ext)
        t = time.time(); synth.write_fastq(f1, r1, o1, 1); synth.write_fastq(f2, r2, o2, 2); tw = time.time() - t
        best = None
        for rep in range(2):
            out = os.path.join(tmp, "out%s%d.tsv" % (ext.replace(".", "_"), rep))
            t = time.time(); nb.process_fastq([f1, f2], [lib_path], [out], strand_filter="unstranded", num_cores=cores); dt = time.time() - t
            best = dt if best is None else min(best, dt)
        rows = sum(1 for _ in open(out)) - 1
        print(json.dumps({"input": "2 x %s, %d pairs 2x150 (%.0f MB each on disk)" % (ext, a.pairs, os.path.getsize(f1) / 1e6), "reads_per_s": 2 * a.pairs / best, "seconds": best,
                          "host_threads": cores, "tsv_rows": rows, "write_inputs_s": round(tw, 1)}), flush=True)
    shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
