#!/bin/bash
# where the C4 step goes: finalize phases + launch list of the C4 block
set -u
mkdir -p gpurun_out
NB_FINALIZE_STATS=1 python bench.py --pairs 1000000 --steps 3 --warmup 3 --no-cpu-baseline --blocks c4 > gpurun_out/s10.json 2> gpurun_out/s10.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/s10.json')); c=d['c4']
print("C4 value %.1f M ms %.2f | e2e %.1f M (%.1f ms) launches %d uniq %d callsets %d" % (c['value']/1e6, c['ms_per_step'], c['e2e']['value']/1e6, c['e2e']['ms_per_step'], c['gpu_launches'], c['unique_read_keys'], c['callsets_counted']))
PY
grep "finalize:" gpurun_out/s10.err | tail -2
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 --csv --log-file gpurun_out/s10_launches.csv python bench.py --pairs 1000000 --steps 1 --warmup 1 --no-cpu-baseline --blocks c4 > gpurun_out/s10_ncu.log 2>&1; echo "ncu rc $?"
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/s10_launches.csv')) if len(r)>10]
hdr=rows[0]; ik=hdr.index("Kernel Name"); im=hdr.index("Metric Name"); iv=hdr.index("Metric Value"); iid=hdr.index("ID")
# keep launches after the C4 index build (k_probe_build of the big table is the marker): print per-kernel totals of the last 40% of launches
recs=[(int(r[iid]), r[ik], float(r[iv].replace(',',''))) for r in rows[1:] if r[im]=="gpu__time_duration.sum"]
first=[i for i,(id_,k,v) in enumerate(recs) if 'k_walk<0, 1>' in k or 'k_walk<(int)0, (int)1>' in k]
start=first[0] if first else 0
agg=collections.defaultdict(lambda:[0,0.0])
for id_,k,v in recs[start:]:
    a=agg[k.split('(')[0][:70]]; a[0]+=1; a[1]+=v
tot=sum(a[1] for a in agg.values())
for k,a in sorted(agg.items(), key=lambda x:-x[1][1])[:14]: print("%-72s n=%4d total %9.1f us mean %8.1f" % (k, a[0], a[1]/1e3, a[1]/1e3/a[0]))
print("total us", tot/1e3)
PY
