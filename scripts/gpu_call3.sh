#!/bin/bash
set -u
mkdir -p gpurun_out
(time python bench.py --steps 3 --warmup 3) > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err; echo "bench rc $?"; tail -5 gpurun_out/r02_bench_a.err; cat gpurun_out/r02_bench_a.json
(time python -m pytest tests/test_gpu_biglib.py -x -q) > gpurun_out/r02_biglib.log 2>&1; tail -15 gpurun_out/r02_biglib.log
