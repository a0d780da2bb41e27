#!/bin/bash
# k_seed micro-optimisations (32-bit Bloom test, 3-popcount entropy): parity, C2 step, ncu of k_seed / k_pair / k_fold at C2
set -u
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/s13_tests.log 2>&1; grep -n "passed\|failed" gpurun_out/s13_tests.log | tail -1
PAIRS=10000000 bash scripts/run_variants.sh 2>&1 | tee gpurun_out/s13_c2.txt
C2="python bench.py --pairs 2000000 --steps 1 --warmup 1 --no-cpu-baseline --blocks none"
ncu --set full --clock-control none --import-source on -k regex:"^k_seed$|^k_pair$|^k_fold$" -s 5 -c 3 -f -o gpurun_out/prof_s13_c2 $C2 > gpurun_out/s13_ncu.log 2>&1; echo "ncu rc $?"
