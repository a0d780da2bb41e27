#!/bin/bash
# ncu --set full of the final map-stage kernels (after the last k_walk change): C2 and C4
set -u
mkdir -p gpurun_out
C2="python bench.py --pairs 2000000 --steps 1 --warmup 1 --no-cpu-baseline --blocks none"
ncu --set full --clock-control none --import-source on -k regex:"^k_seed$|^k_walk$" -s 4 -c 2 -f -o gpurun_out/prof_f3_c2 $C2 > gpurun_out/f3_ncu_c2.log 2>&1; echo "ncu c2 rc $?"
C4="python bench.py --pairs 1000000 --steps 1 --warmup 1 --no-cpu-baseline --blocks c4 --c4-reads 2000000"
ncu --set full --clock-control none --import-source on -k regex:"^k_seed$|^k_walk$" -s 16 -c 2 -f -o gpurun_out/prof_f3_c4 $C4 > gpurun_out/f3_ncu_c4.log 2>&1; echo "ncu c4 rc $?"
