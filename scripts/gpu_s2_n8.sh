#!/bin/bash
# N-GPU bench line with the sharded scoped merge and the per-warp import counters (+ merge phase times)
set -u
N=${N:-8}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/n${N}_bench.json 2> gpurun_out/n${N}_bench.err; echo "bench N=$N rc $?"; grep -a "verify\|rror" gpurun_out/n${N}_bench.err | tail -5 | cut -c1-300
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/n${N}_bench.json'))
    print("N=%d value %.3f G ms %.3f | e2e %.3f G (%.2f ms, %.1f GB/s per GPU) | packed e2e %.3f G | parity %s" % (d['n_gpus'], d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['e2e']['ms_per_step'], d['e2e']['h2d_gbs_per_gpu'], d['e2e_packed']['value']/1e9, d.get('parity_checked')))
    c=d.get('c3')
    if c: print("c3 value %.1f M/s ms %.2f e2e %.1f M/s (%.1f ms) rows %d parity %s" % (c['value']/1e6, c['ms_per_step'], c['e2e']['value']/1e6, c['e2e']['ms_per_step'], c['count_rows'], c.get('parity_checked')))
except Exception as e: print("no json", e)
PY
NB_MERGE_STATS=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 2 --warmup 3 --blocks none --no-verify > gpurun_out/n${N}_stats.json 2> gpurun_out/n${N}_stats.err; grep -a "nb_merge_whole_run" gpurun_out/n${N}_stats.err | tail -1 | cut -c1-400
