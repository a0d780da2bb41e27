#!/bin/bash
set -u
mkdir -p gpurun_out
(time python -m pytest tests/test_gpu_merge.py -x -q) > gpurun_out/r02_merge_tests.log 2>&1; tail -15 gpurun_out/r02_merge_tests.log
N=${N:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "bench N=$N rc $?"; tail -8 gpurun_out/r02_bench_n$N.err | cut -c1-300
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02_bench_n$N.json'))
    print("N=%d value %.3f G ms %.3f e2e %.3f G (%.3f ms) parity %s" % (d['n_gpus'], d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['e2e']['ms_per_step'], d.get('parity')))
    c=d.get('c3'); 
    if c: print("c3 value %.1f M/s ms %.2f e2e %.1f M/s rows %d" % (c['value']/1e6, c['ms_per_step'], c['e2e']['value']/1e6, c['count_rows']))
except Exception as e: print("no json", e)
PY
NB_MERGE_STATS=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 --merge py-p2p --blocks none --no-verify > gpurun_out/r02_bench_n${N}_py.json 2> gpurun_out/r02_bench_n${N}_py.err; echo "py-p2p rc $?"
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n${N}_py.json')); print('py-p2p N=%d value %.3f G ms %.3f' % (d['n_gpus'], d['value']/1e9, d['ms_per_step']))"
