#!/bin/bash
# pair-stage row tables (O(1) membership), k_trim side-homogeneous + skip: parity, C2 step, C3 block, launch list
set -u
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/s9_tests.log 2>&1; grep -n "passed\|failed" gpurun_out/s9_tests.log | tail -1
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --blocks c3 > gpurun_out/s9.json 2> gpurun_out/s9.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/s9.json')); c=d['c3']
print("C2 value %.3f G ms %.3f map %.4f | e2e %.3f G" % (d['value']/1e9, d['ms_per_step'], d['k_map_ms_per_launch'], d['e2e']['value']/1e9))
print("C3 value %.1f M rec/s ms %.2f | e2e %.1f M (%.1f ms) | map %.4f ms" % (c['value']/1e6, c['ms_per_step'], c['e2e']['value']/1e6, c['e2e']['ms_per_step'], c['k_map_ms_per_launch']))
PY
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1200 --csv --log-file gpurun_out/s9_launches.csv python bench.py --pairs 1000000 --steps 1 --warmup 1 --no-cpu-baseline --blocks c3 > gpurun_out/s9_ncu.log 2>&1; echo "ncu rc $?"
python scripts/launch_list_summary.py gpurun_out/s9_launches.csv > gpurun_out/s9_launch_summary.csv 2>&1; head -9 gpurun_out/s9_launch_summary.csv
