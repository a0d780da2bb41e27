#!/bin/bash
# session-2 baseline: GPU tests, default bench, ncu launch list of the same bench
set -u
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/s2_tests.log 2>&1; tail -5 gpurun_out/s2_tests.log
(time python bench.py) > gpurun_out/s2_bench.json 2> gpurun_out/s2_bench.err; echo "bench rc $?"; tail -5 gpurun_out/s2_bench.err | cut -c1-400
cat gpurun_out/s2_bench.json | cut -c1-6000
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/s2_launches.csv python bench.py --steps 2 --warmup 1 --blocks none > gpurun_out/s2_ncu.log 2>&1; echo "ncu rc $?"
python scripts/launch_list_summary.py gpurun_out/s2_launches.csv > gpurun_out/s2_launch_summary.csv 2>&1; head -30 gpurun_out/s2_launch_summary.csv
