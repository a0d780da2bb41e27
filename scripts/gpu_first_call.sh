#!/bin/bash
# The measurements to take first when GPU time is available again (run on the GPU box from the repo root, e.g.
#   gpurun --timeout 900 -- 'bash scripts/gpu_first_call.sh'
# after building the variants HERE with scripts/build_variant.sh, since the box has no compiler budget to waste):
#   1. the GPU suite, 2. the bench line, 3. launch list + source-level ncu capture of the map stage (k_seed, k_walk),
#   4. same-run A/B of the prepared kernel variants (scripts/run_variants.sh).
# Everything lands in gpurun_out/; copy what is to be judged into profiles/.
set -u
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/gpu_tests.log 2>&1; tail -3 gpurun_out/gpu_tests.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc $?"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv \
    --log-file gpurun_out/launches.csv python bench.py --pairs 2000000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_seed|k_walk" -c 2 -f -o gpurun_out/prof_map \
    python bench.py --pairs 2000000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
python scripts/launch_list_summary.py gpurun_out/launches.csv "launch list" | head -8
VARIANTS="${VARIANTS:--}" bash scripts/run_variants.sh
