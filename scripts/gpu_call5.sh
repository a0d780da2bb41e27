#!/bin/bash
set -u
mkdir -p gpurun_out
CMD="python bench.py --pairs 2000000 --steps 1 --warmup 1 --no-cpu-baseline --blocks none"
$CMD > gpurun_out/r02_plain5.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_seed|k_walk" -s 2 -c 2 -f -o gpurun_out/prof_r02b_c2 $CMD > gpurun_out/r02_ncu5.log 2>&1
echo "ncu rc $?"
python scripts/dbg_walk.py > gpurun_out/r02_dbg_walk2.txt 2>&1; grep -a "k_map dbg\|probes" gpurun_out/r02_dbg_walk2.txt
python scripts/dbg_c3.py 2>/dev/null | head -5
