#!/bin/bash
set -u
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/r02_gpu_tests_c.log 2>&1; tail -5 gpurun_out/r02_gpu_tests_c.log
python bench.py --steps 3 --warmup 3 --blocks c4 --no-cpu-baseline > gpurun_out/r02_bench_c.json 2> gpurun_out/r02_bench_c.err; echo "bench rc $?"; tail -3 gpurun_out/r02_bench_c.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_c.json'))
print("C2 value %.3f G  ms %.3f  map %.4f ms/2M  e2e %.3f G" % (d['value']/1e9, d['ms_per_step'], d['k_map_ms_per_launch'], d['e2e']['value']/1e9))
c=d['c4']; print("C4 value %.3f G ms %.3f e2e %.3f" % (c['value']/1e9, c['ms_per_step'], c['e2e']['value']/1e9))
PY
CMD="python bench.py --pairs 2000000 --steps 1 --warmup 1 --no-cpu-baseline --blocks none"
$CMD > gpurun_out/r02_plain6.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_seed|k_walk" -s 2 -c 2 -f -o gpurun_out/prof_r02c_c2 $CMD > gpurun_out/r02_ncu6.log 2>&1
echo "ncu rc $?"
