#!/bin/bash
# k_seed with two probes in flight per thread; k_pair register caps: parity, then C2 step per variant + launch list of the default
set -u
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/s14_tests.log 2>&1; grep -n "passed\|failed" gpurun_out/s14_tests.log | tail -1
VARIANTS="- _pm10 _pm12 _pm16" PAIRS=10000000 bash scripts/run_variants.sh 2>&1 | tee gpurun_out/s14_c2.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/s14_launches.csv python bench.py --steps 2 --warmup 1 --blocks none --no-cpu-baseline > gpurun_out/s14_ncu.log 2>&1
python scripts/launch_list_summary.py gpurun_out/s14_launches.csv 2>&1 | head -8
