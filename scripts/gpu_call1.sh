#!/bin/bash
# round 2, first measurement pass: roofs of the box, the two prepared variants, C4-scaled plain run + ncu of the map stage
set -u
mkdir -p gpurun_out
python scripts/measure_roofs.py > gpurun_out/r02_roofs.json 2> gpurun_out/r02_roofs.err; echo "roofs rc $?"; cat gpurun_out/r02_roofs.json
VARIANTS="- _lo32 _nopad" bash scripts/run_variants.sh > gpurun_out/r02_variants.txt 2>&1; cat gpurun_out/r02_variants.txt
python scripts/bench_extra.py --workload c4s --families 8000 --reads 8000000 --steps 3 > gpurun_out/r02_c4s_8m.json 2> gpurun_out/r02_c4s_8m.err; cat gpurun_out/r02_c4s_8m.json
python scripts/bench_extra.py --workload c4s --families 8000 --reads 2000000 --steps 1 > gpurun_out/r02_c4s_plain.json 2> gpurun_out/r02_c4s_plain.err &&
ncu --set full --clock-control none --import-source on -k regex:"k_seed|k_walk" -s 4 -c 2 -f -o gpurun_out/prof_r02_c4_base \
    python scripts/bench_extra.py --workload c4s --families 8000 --reads 2000000 --steps 1 > gpurun_out/r02_c4s_ncu.log 2>&1
echo "ncu rc $?"
