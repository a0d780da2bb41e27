#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 scripts/micro/gather_bench > gpurun_out/r02_gather_bench.txt 2>&1; cat gpurun_out/r02_gather_bench.txt
python scripts/dbg_walk.py > gpurun_out/r02_dbg_walk.txt 2>&1; tail -3 gpurun_out/r02_dbg_walk.txt
VARIANTS="- _pmin1 _pmin3 _pmin10 _smin1 _smin2" bash scripts/run_variants.sh > gpurun_out/r02_variants2.txt 2>&1; cat gpurun_out/r02_variants2.txt
