#!/bin/bash
# the driver's N-GPU command on the final tree (c5 / c3 blocks, verification passes)
set -u
N=${N:-8}
mkdir -p gpurun_out
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3) > gpurun_out/n${N}b_bench.json 2> gpurun_out/n${N}b_bench.err; echo "bench N=$N rc $?"; grep -a "verify\|rror\|real" gpurun_out/n${N}b_bench.err | tail -6 | cut -c1-300
python - <<PY
import json
d=json.load(open('gpurun_out/n${N}b_bench.json'))
print("N=%d value %.3f G ms %.3f | e2e %.3f G | packed e2e %.3f G | parity %s" % (d['n_gpus'], d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['e2e_packed']['value']/1e9, d.get('parity_checked')))
for k,e in sorted(d['c5']['sweep'].items()): print("c5", k, "value %.3f G parity %s" % (e['value']/1e9, e.get('parity_checked')))
c=d['c3']; print("c3 value %.1f M/s ms %.2f e2e %.1f M/s rows %d parity %s" % (c['value']/1e6, c['ms_per_step'], c['e2e']['value']/1e6, c['count_rows'], c.get('parity_checked')))
PY
