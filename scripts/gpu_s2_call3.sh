#!/bin/bash
# k_walk generation 3: parity tests, then C2 / C4 map-stage times
set -u
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/s3_tests.log 2>&1; tail -5 gpurun_out/s3_tests.log
VARIANTS="${VARIANTS:--}" PAIRS=4000000 bash scripts/run_variants.sh 2>&1 | tee gpurun_out/s3_variants.txt
python bench.py --pairs 2000000 --steps 3 --warmup 3 --no-cpu-baseline --blocks c4 --c4-reads 4000000 > gpurun_out/s3_c4.json 2> gpurun_out/s3_c4.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/s3_c4.json')); c=d['c4']
print("C2 map %.4f ms/2M | C4 value %.1f M ms %.2f launches %d" % (d['k_map_ms_per_launch'], c['value']/1e6, c['ms_per_step'], c['gpu_launches']))
PY
NB_DEBUG_KMAP=1 python scripts/dbg_walk.py 2>&1 | grep -a "k_map dbg\|probes" | tail -3
