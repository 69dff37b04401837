for cfg in "160 4096" "200 4096" "216 4096" "200 3072"; do
  set -- $cfg
  echo "== WBUDGET=${1}K STAGE_MAX=$2"
  GGB200_GEMV_WBUDGET=$(($1*1024)) GGB200_GEMV_STAGE_MAX=$2 python benchmarks/bench_configs.py --only cfg0,cfg2,cfg1 --iters 20 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l)
    if 'GB/s' in d and 'isolated' not in d['config'] and 'quantize' not in d['config']: print('   %-60s %7.1f us %6.0f GB/s' % (d['config'][:60], d['ms']*1e3, d['GB/s']))
"
done
