#!/bin/bash
# BAM driver GPU tests after the rows-stage touch-ups (the very last GPU seconds)
mkdir -p gpurun_out
timeout 20 python -m pytest tests/test_gpu_bam.py -x -q > gpurun_out/s38_tests.log 2>&1
tail -3 gpurun_out/s38_tests.log
