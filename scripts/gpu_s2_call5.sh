#!/bin/bash
# live key counter (no k_count_keys pass), FASTQ driver timing, C4 map-stage ncu capture
set -u
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/s5_tests.log 2>&1; tail -4 gpurun_out/s5_tests.log
PAIRS=10000000 bash scripts/run_variants.sh 2>&1 | tee gpurun_out/s5_c2.txt
NB_FASTQ_STATS=1 python scripts/bench_fastq.py --pairs 10000000 > gpurun_out/s5_fastq.json 2> gpurun_out/s5_fastq.err; cat gpurun_out/s5_fastq.json; grep nb_process_fastq gpurun_out/s5_fastq.err
C4="python bench.py --pairs 1000000 --steps 1 --warmup 1 --no-cpu-baseline --blocks c4 --c4-reads 2000000"
ncu --set full --clock-control none --import-source on -k regex:"k_seed|k_walk" -s 10 -c 4 -f -o gpurun_out/prof_s5_c4 $C4 > gpurun_out/s5_ncu_c4.log 2>&1
echo "ncu c4 rc $?"
