#!/bin/bash
# BAM driver: gzip level of the TSV members vs wall time (20 M records)
set -u
mkdir -p gpurun_out
for lv in 4 2 1; do
NB_BAM_GZ_LEVEL=$lv timeout 600 python scripts/bench_bam.py --groups 5000000 --repeat 1 --cpu-groups 0 --cli > gpurun_out/s31_bam_$lv.json 2> gpurun_out/s31_bam_$lv.err
python -c "
import json; d=json.load(open('gpurun_out/s31_bam_$lv.json')); print('level $lv: %.2f s whole call, tsv.gz %.0f MB' % (d['seconds'], d['tsv_gz_mb']))"; grep -a "nb_process_bam:" gpurun_out/s31_bam_$lv.err | tail -1 | cut -c60-400
done
