# usage: VARIANTS="_w9 _w10" [PAIRS=4000000] [ARGS="..."] bash scripts/run_variants.sh   (kernel tuning builds from scripts/build_variant.sh; "-" = the default build)
for v in ${VARIANTS:--}; do
  [ "$v" = "-" ] && v=""
  echo "== variant ${v:-default} ${ARGS:-}"
  NIMBLE_B200_SO=$PWD/nimble_aligner_b200/libnimble_b200$v.so timeout 300 python bench.py --pairs ${PAIRS:-4000000} --steps 3 --warmup 3 --no-cpu-baseline --blocks none ${ARGS:-} 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('value %.3f G/s  ms/step %.3f  e2e %.3f  map stage %.4f ms / %d reads' % (d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['k_map_ms_per_launch'], d['k_map_reads_per_launch']))
    elif 'rror' in l or 'assert' in l: print(l.rstrip())
"
done
