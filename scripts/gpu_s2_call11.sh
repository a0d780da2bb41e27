#!/bin/bash
# multi-GPU: merge tests (need >= 2 GPUs), bench at N = $N with the in-library merge (+ verification pass), merge phase times
set -u
N=${N:-2}
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests/test_gpu_merge.py tests/test_gpu_route.py -x -q) > gpurun_out/s11_merge_tests_n$N.log 2>&1; grep -n "passed\|failed" gpurun_out/s11_merge_tests_n$N.log | tail -1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/s11_bench_n$N.json 2> gpurun_out/s11_bench_n$N.err; echo "bench N=$N rc $?"; grep -a "verify\|rror" gpurun_out/s11_bench_n$N.err | tail -5 | cut -c1-300
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/s11_bench_n$N.json'))
    print("N=%d value %.3f G ms %.3f | e2e %.3f G (%.2f ms, %.1f GB/s per GPU) | packed e2e %.3f G | parity %s" % (d['n_gpus'], d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['e2e']['ms_per_step'], d['e2e']['h2d_gbs_per_gpu'], d['e2e_packed']['value']/1e9, d.get('parity_checked')))
    c=d.get('c3')
    if c: print("c3 value %.1f M/s ms %.2f e2e %.1f M/s rows %d parity %s" % (c['value']/1e6, c['ms_per_step'], c['e2e']['value']/1e6, c['count_rows'], c.get('parity_checked')))
except Exception as e: print("no json", e)
PY
NB_MERGE_STATS=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 --blocks none --no-verify > gpurun_out/s11_bench_n${N}_stats.json 2> gpurun_out/s11_bench_n${N}_stats.err; grep -a "nb_merge_whole_run" gpurun_out/s11_bench_n${N}_stats.err | tail -2 | cut -c1-400
