#!/bin/bash
# .fastq.gz through the parallel reader on the box (2 M pairs; the round's last GPU seconds)
set -u
mkdir -p gpurun_out
NB_GZ_STATS=1 NB_FASTQ_STATS=1 timeout 42 python scripts/bench_fastq.py --pairs 2000000 > gpurun_out/s37_fastq.json 2> gpurun_out/s37_fastq.err
cat gpurun_out/s37_fastq.json; grep -a "parallel gunzip\|thread time\|nb_process_fastq" gpurun_out/s37_fastq.err | tail -8
