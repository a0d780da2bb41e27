#!/bin/bash
# bench line with the c5 block (mismatch sweep with parity), N = 1
set -u
mkdir -p gpurun_out
(time python bench.py) > gpurun_out/s22_bench.json 2> gpurun_out/s22_bench.err; echo "bench rc $?"; tail -3 gpurun_out/s22_bench.err | cut -c1-300
python - <<'PY'
import json
d=json.load(open('gpurun_out/s22_bench.json'))
print("C2 value %.3f G ms %.3f | e2e %.3f G | map %.4f frac %.3f parity %s" % (d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['k_map_ms_per_launch'], d['roofline']['frac'], d.get('parity_checked')))
for k,e in sorted(d['c5']['sweep'].items()): print("c5", k, "value %.3f G ms %.3f map %.4f parity %s" % (e['value']/1e9, e['ms_per_step'], e['k_map_ms_per_launch'], e.get('parity_checked')))
print("c5 wall", d['c5']['block_wall_s'], "| c4 %.1f M | c3 %.1f M" % (d['c4']['value']/1e6, d['c3']['value']/1e6))
PY
