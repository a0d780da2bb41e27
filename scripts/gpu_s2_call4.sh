#!/bin/bash
set -u
mkdir -p gpurun_out
C2="python bench.py --pairs 2000000 --steps 1 --warmup 1 --no-cpu-baseline --blocks none"
for v in "" _mb8; do
NIMBLE_B200_SO=$PWD/nimble_aligner_b200/libnimble_b200$v.so ncu --set full --clock-control none --import-source on -k regex:"k_walk" -s 2 -c 1 -f -o gpurun_out/prof_s4_walk$v $C2 > gpurun_out/s4_ncu$v.log 2>&1
echo "ncu $v rc $?"
done
