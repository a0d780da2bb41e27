#!/bin/bash
# file drivers after the host-side work (own inflate, row formatter): BAM driver on the 20 M-record file of call 34, then .fastq.gz
set -u
mkdir -p gpurun_out
timeout 100 python scripts/bench_bam.py --groups 5000000 --repeat 1 --cpu-groups 0 --cli > gpurun_out/s35_bam.json 2> gpurun_out/s35_bam.err
grep -a "nb_process_bam:" gpurun_out/s35_bam.err | tail -1 | cut -c1-500
timeout 100 python scripts/bench_fastq.py --pairs 2000000 > gpurun_out/s35_fastq.json 2> gpurun_out/s35_fastq.err
tail -c 1500 gpurun_out/s35_fastq.json
