#!/bin/bash
# bulk-copy k_pack + block-aggregated key counting: parity, C2 step, launch list
set -u
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/s6_tests.log 2>&1; tail -4 gpurun_out/s6_tests.log
PAIRS=10000000 bash scripts/run_variants.sh 2>&1 | tee gpurun_out/s6_c2.txt
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/s6_launches.csv python bench.py --steps 2 --warmup 1 --blocks none --no-cpu-baseline > gpurun_out/s6_ncu.log 2>&1; echo "ncu rc $?"
python scripts/launch_list_summary.py gpurun_out/s6_launches.csv > gpurun_out/s6_launch_summary.csv 2>&1; head -14 gpurun_out/s6_launch_summary.csv
