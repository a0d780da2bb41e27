#!/bin/bash
# round-2 closing pass: full GPU test suite, the driver's bench line, launch lists and ncu --set full captures for profiles/
set -u
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/f_tests.log 2>&1; grep -n "passed\|failed" gpurun_out/f_tests.log | tail -1
(time python bench.py) > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/f_bench.json'))
print("C2 value %.3f G ms %.3f | e2e %.3f G | packed %.3f G | map %.4f ms frac %.3f | cpu %.2f M | parity %s" % (d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['e2e_packed']['value']/1e9, d['k_map_ms_per_launch'], d['roofline']['frac'], d['cpu_baseline']['value']/1e6, d.get('parity_checked')))
for k in ('c4','c3'):
    c=d[k]; print(k, "value %.1f M ms %.2f | e2e %.1f M (%.1f ms) | frac %.3f launch %.4f ms | cpu %.3f M | parity %s" % (c['value']/1e6, c['ms_per_step'], c['e2e']['value']/1e6, c['e2e']['ms_per_step'], c['roofline']['frac'], c['roofline']['launch_ms'], c['cpu_baseline']['value']/1e6, c.get('parity_checked')))
PY
python bench.py --impl reference --steps 1 --warmup 1 --blocks none > gpurun_out/f_ref.json 2> gpurun_out/f_ref.err; cut -c1-300 gpurun_out/f_ref.json
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2000 --csv --log-file gpurun_out/f_launches.csv python bench.py --pairs 2000000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/f_ncu_list.log 2>&1; echo "ncu list rc $?"
python scripts/launch_list_summary.py gpurun_out/f_launches.csv > gpurun_out/f_launch_summary.csv 2>&1; head -16 gpurun_out/f_launch_summary.csv | cut -c1-160
C2="python bench.py --pairs 2000000 --steps 1 --warmup 1 --no-cpu-baseline --blocks none"
ncu --set full --clock-control none --import-source on -k regex:"^k_pack$|^k_seed$|^k_walk$|^k_pair$|^k_fold$" -s 8 -c 9 -f -o gpurun_out/prof_f_c2 $C2 > gpurun_out/f_ncu_c2.log 2>&1; echo "ncu c2 rc $?"
C4="python bench.py --pairs 1000000 --steps 1 --warmup 1 --no-cpu-baseline --blocks c4 --c4-reads 2000000"
ncu --set full --clock-control none --import-source on -k regex:"^k_seed$|^k_walk$" -s 16 -c 2 -f -o gpurun_out/prof_f_c4 $C4 > gpurun_out/f_ncu_c4.log 2>&1; echo "ncu c4 rc $?"
C3="python bench.py --pairs 1000000 --steps 1 --warmup 1 --no-cpu-baseline --blocks c3 --c3-records 4000000"
ncu --set full --clock-control none --import-source on -k regex:"^k_trim$|^k_fold$|^k_pair$" -s 12 -c 3 -f -o gpurun_out/prof_f_c3 $C3 > gpurun_out/f_ncu_c3.log 2>&1; echo "ncu c3 rc $?"
ls -la gpurun_out/*.ncu-rep
