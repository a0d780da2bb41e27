#!/bin/bash
# branch-free k_trim + zero-copy result views: parity, C3 block, ncu capture of k_trim / k_pair on C3-shaped batches
set -u
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/s8_tests.log 2>&1; grep -n "passed\|failed" gpurun_out/s8_tests.log | tail -1
NB_FINALIZE_STATS=1 python bench.py --pairs 2000000 --steps 3 --warmup 3 --no-cpu-baseline --blocks c3 > gpurun_out/s8_c3.json 2> gpurun_out/s8_c3.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/s8_c3.json')); c=d['c3']
print("C3 value %.1f M rec/s ms %.2f | e2e %.1f M (%.1f ms) | map %.4f ms" % (c['value']/1e6, c['ms_per_step'], c['e2e']['value']/1e6, c['e2e']['ms_per_step'], c['k_map_ms_per_launch']))
PY
grep "finalize:" gpurun_out/s8_c3.err | tail -3
CMD="python bench.py --pairs 1000000 --steps 1 --warmup 1 --no-cpu-baseline --blocks c3 --c3-records 4000000"
ncu --set full --clock-control none --import-source on -k regex:"^k_trim$|^k_pair$" -s 14 -c 2 -f -o gpurun_out/prof_s8_c3 $CMD > gpurun_out/s8_ncu.log 2>&1; echo "ncu rc $?"
