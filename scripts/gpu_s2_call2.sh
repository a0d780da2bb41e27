#!/bin/bash
# full ncu captures of the current map-stage kernels at C2 and C4 (+ k_pack / k_pair), launch list of the c3 / c4 blocks
set -u
mkdir -p gpurun_out
C2="python bench.py --pairs 2000000 --steps 1 --warmup 1 --no-cpu-baseline --blocks none"
$C2 > gpurun_out/s2_plain_c2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_seed|k_walk|k_pair|k_pack$" -s 4 -c 4 -f -o gpurun_out/prof_s2_c2 $C2 > gpurun_out/s2_ncu_c2.log 2>&1
echo "ncu c2 rc $?"
C4="python bench.py --pairs 1000000 --steps 1 --warmup 1 --no-cpu-baseline --blocks c4 --c4-reads 2000000"
$C4 > gpurun_out/s2_plain_c4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_seed<0, 1>|k_walk<0, 1>" -s 2 -c 2 -f -o gpurun_out/prof_s2_c4 $C4 > gpurun_out/s2_ncu_c4.log 2>&1
echo "ncu c4 rc $?"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 --csv --log-file gpurun_out/s2_launches_blocks.csv python bench.py --pairs 1000000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/s2_ncu_blocks.log 2>&1; echo "ncu blocks rc $?"
python scripts/launch_list_summary.py gpurun_out/s2_launches_blocks.csv > gpurun_out/s2_launch_summary_blocks.csv 2>&1; head -40 gpurun_out/s2_launch_summary_blocks.csv
ls -la gpurun_out/
