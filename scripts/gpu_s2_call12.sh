#!/bin/bash
# parallel callset sort: parity, C4 block, finalize phases of C2 and C4
set -u
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/s12_tests.log 2>&1; grep -n "passed\|failed" gpurun_out/s12_tests.log | tail -1
NB_FINALIZE_STATS=1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --blocks c4 > gpurun_out/s12.json 2> gpurun_out/s12.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/s12.json')); c=d['c4']
print("C2 value %.3f G ms %.3f map %.4f" % (d['value']/1e9, d['ms_per_step'], d['k_map_ms_per_launch']))
print("C4 value %.1f M ms %.2f | e2e %.1f M (%.1f ms)" % (c['value']/1e6, c['ms_per_step'], c['e2e']['value']/1e6, c['e2e']['ms_per_step']))
PY
grep "finalize:" gpurun_out/s12.err | sed -n '4p;5p;$p' | cut -c1-300
