#!/bin/bash
# the file drivers' GPU tests after the host-side work (parallel .gz reader with pinned segments, row formatter)
set -u
mkdir -p gpurun_out
timeout 68 python -m pytest tests/test_gpu_parity.py tests/test_gpu_errors.py tests/test_gpu_bam.py -k "fastq_driver or process_bam" -x -q > gpurun_out/s36_tests.log 2>&1
tail -5 gpurun_out/s36_tests.log
