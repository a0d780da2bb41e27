#!/bin/bash
set -u
mkdir -p gpurun_out
python scripts/dbg_c3.py > gpurun_out/r02_dbg_c3.txt 2>&1; tail -8 gpurun_out/r02_dbg_c3.txt
(time python -m pytest tests -m gpu -x -q) > gpurun_out/r02_gpu_tests_b.log 2>&1; tail -15 gpurun_out/r02_gpu_tests_b.log
python bench.py --steps 3 --warmup 3 --blocks c4 > gpurun_out/r02_bench_b.json 2> gpurun_out/r02_bench_b.err; echo "bench rc $?"; tail -5 gpurun_out/r02_bench_b.err; cat gpurun_out/r02_bench_b.json
