for c in 2 3 4 6 8; do
  echo -n "chunks=$c: "
  GGB200_LANE_CHUNKS=$c timeout 200 python bench.py --steps 100 --warmup 10 --no-configs --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['e2e']['value']), round(d['e2e']['ms_per_step']*1e3,1), round(d['e2e']['device_ms_per_step']*1e3,1))"
done
