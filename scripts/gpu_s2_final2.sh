#!/bin/bash
# last pass on the final tree: full GPU suite, the driver's bench line (with c5 / c4 / c3 blocks), the reference arm
set -u
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/f2_tests.log 2>&1; grep -n "passed\|failed" gpurun_out/f2_tests.log | tail -1
(time python bench.py) > gpurun_out/f2_bench.json 2> gpurun_out/f2_bench.err; echo "bench rc $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/f2_bench.json'))
print("C2 value %.3f G ms %.3f | e2e %.3f G | packed %.3f G | map %.4f ms frac %.3f | cpu %.2f M | parity %s | launches %d" % (d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['e2e_packed']['value']/1e9, d['k_map_ms_per_launch'], d['roofline']['frac'], d['cpu_baseline']['value']/1e6, d.get('parity_checked'), d['gpu_launches']))
for k,e in sorted(d['c5']['sweep'].items()): print("c5", k, "value %.3f G parity %s" % (e['value']/1e9, e.get('parity_checked')))
for k in ('c4','c3'):
    c=d[k]; print(k, "value %.1f M ms %.2f | e2e %.1f M (%.1f ms) | frac %.3f launch %.4f ms | cpu %.3f M | parity %s" % (c['value']/1e6, c['ms_per_step'], c['e2e']['value']/1e6, c['e2e']['ms_per_step'], c['roofline']['frac'], c['roofline']['launch_ms'], c['cpu_baseline']['value']/1e6, c.get('parity_checked')))
PY
(time python bench.py --impl reference --steps 2 --warmup 1) > gpurun_out/f2_ref.json 2> gpurun_out/f2_ref.err; python -c "
import json; d=json.load(open('gpurun_out/f2_ref.json')); print('reference arm: C2 %.2f M reads/s | c3 %.3f M | c4 %.3f M (cores %d)' % (d['value']/1e6, d['c3']['value']/1e6, d['c4']['value']/1e6, d['cpu_baseline']['cores']))"
python __graft_entry__.py --smoke 2>&1 | tail -1
