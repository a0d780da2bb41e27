#!/bin/bash
# k_fold row counting per block, vectorised k_trim, bulk-pack variant test: parity, then the C3 block and its launch list
set -u
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/s7_tests.log 2>&1; tail -4 gpurun_out/s7_tests.log | head -2
python bench.py --pairs 2000000 --steps 3 --warmup 3 --no-cpu-baseline --blocks c3 > gpurun_out/s7_c3.json 2> gpurun_out/s7_c3.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/s7_c3.json')); c=d['c3']
print("C3 value %.1f M rec/s ms %.2f | e2e %.1f M (%.1f ms) | map %.4f ms" % (c['value']/1e6, c['ms_per_step'], c['e2e']['value']/1e6, c['e2e']['ms_per_step'], c['k_map_ms_per_launch']))
PY
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1200 --csv --log-file gpurun_out/s7_launches.csv python bench.py --pairs 1000000 --steps 1 --warmup 1 --no-cpu-baseline --blocks c3 > gpurun_out/s7_ncu.log 2>&1; echo "ncu rc $?"
python scripts/launch_list_summary.py gpurun_out/s7_launches.csv > gpurun_out/s7_launch_summary.csv 2>&1; head -14 gpurun_out/s7_launch_summary.csv
