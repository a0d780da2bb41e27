#!/usr/bin/env python
"""Times the index build (K5) on the host builder and on the CUDA builder for the C2 library and a C4-scaled library,
and checks the two artefacts are the same index.  One JSON line per library."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nimble_aligner_b200 as nb
import synth


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--families", type=int, nargs="+", default=[200, 8000])
    a = ap.parse_args()
    cores = os.cpu_count() or 1
    for nf in a.families:
        L = synth.SynthLibrary(seed=3456, n_fam=nf, n_all=5, group_on="")
        lib = nb.Library.from_text(json.dumps(L.to_json_obj()), "unstranded")
        nb.build_index(lib, cores, device=0)      # warm-up: CUDA context + module load
        t0 = time.time(); dev = nb.build_index(lib, cores, device=0); t1 = time.time()
        host = nb.build_index(lib, cores); t2 = time.time()
        st = dev.stats()
        print(json.dumps({"library": "%d transcripts" % (5 * nf), "n_kmers": st["n_kmers"], "n_nodes": st["n_nodes"], "n_colours": st["n_colours"],
                          "device_bytes": st["device_bytes"], "gpu_build_s": round(t1 - t0, 3), "host_build_s": round(t2 - t1, 3), "host_threads": cores,
                          "same_index": dev.compare(host) == 0}), flush=True)


if __name__ == "__main__":
    main()
