#!/bin/bash
set -u
mkdir -p gpurun_out
(time python -m pytest tests/test_gpu_parity.py -x -q -k "packed or scope_larger or hbm_resident") > gpurun_out/r02_enc_tests.log 2>&1; tail -8 gpurun_out/r02_enc_tests.log
(time python bench.py --steps 3 --warmup 3) > gpurun_out/r02_bench_d.json 2> gpurun_out/r02_bench_d.err; echo "bench rc $?"; tail -4 gpurun_out/r02_bench_d.err | cut -c1-400
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_d.json'))
print("C2 value %.3f G ms %.3f | e2e %.3f G (%.1f ms, %.1f GB/s) | packed e2e %.3f G (%.1f ms) | map %.4f ms frac %.3f | cpu %.2f M" % (d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['e2e']['ms_per_step'], d['e2e']['h2d_gbs_per_gpu'], d['e2e_packed']['value']/1e9, d['e2e_packed']['ms_per_step'], d['k_map_ms_per_launch'], d['roofline']['frac'], d['cpu_baseline']['value']/1e6))
for k in ('c4','c3'):
    c=d[k]; print(k, "value %.1f M ms %.2f | e2e %.1f M (%.1f ms) | frac %.3f launch %.4f ms | cpu %.3f M | parity %s | wall %.0fs" % (c['value']/1e6, c['ms_per_step'], c['e2e']['value']/1e6, c['e2e']['ms_per_step'], c['roofline']['frac'], c['roofline']['launch_ms'], c['cpu_baseline']['value']/1e6, c.get('parity_checked'), c['block_wall_s']))
PY
