#!/bin/bash
# BASELINE config 3 at its real size through the BAM driver: 100 M records (25 M UMI groups) -> nimble CLI -> TSV.gz
set -u
mkdir -p gpurun_out
df -h /tmp | tail -1; free -g | head -2
G=${G:-25000000}
NB_BAM_STATS=1 timeout 1500 python scripts/bench_bam.py --groups $G --repeat 1 --cpu-groups 100000 --cli > gpurun_out/s18_bam.json 2> gpurun_out/s18_bam.err; echo "rc $?"; cat gpurun_out/s18_bam.json | cut -c1-900; grep -a "nb_process_bam" gpurun_out/s18_bam.err | tail -2 | cut -c1-600; tail -3 gpurun_out/s18_bam.err | cut -c1-300
