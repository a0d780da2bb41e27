#!/bin/bash
# BAM driver feeding 4-bit nibbles: BAM parity tests, then a 20 M-record run through the CLI (throughput, phases, peak RSS)
set -u
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests/test_gpu_bam.py tests/test_gpu_errors.py -x -q) > gpurun_out/s20_tests.log 2>&1; grep -n "passed\|failed" gpurun_out/s20_tests.log | tail -1; grep -n "^E " gpurun_out/s20_tests.log | head -5
timeout 900 python scripts/bench_bam.py --groups ${G:-5000000} --repeat 1 --cpu-groups 0 --cli > gpurun_out/s20_bam.json 2> gpurun_out/s20_bam.err; echo "rc $?"; cut -c1-700 gpurun_out/s20_bam.json; grep -a "nb_process_bam:" gpurun_out/s20_bam.err | tail -1 | cut -c1-600
