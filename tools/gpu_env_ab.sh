#!/bin/bash
# A/B of an environment knob on the in-tree library: tools/gpu_env_ab.sh <tag> VAR v1 v2 ...
TAG=$1; VAR=$2; shift 2
mkdir -p gpurun_out
show() { python -c "
import json,sys
try:
    d=json.load(open(sys.argv[1])); r=d['roofline']
    print(sys.argv[2], round(d['value']), 'img/s', round(d['ms_per_step'],4), 'ms/step lat', round(d.get('latency_ms',0),4), 'frac', round(r['frac'],3), {k: (round(v*1e3,1) if v is not None else None) for k,v in r['stage_ms'].items()}, d.get('verified'))
except Exception as e: print(sys.argv[2], 'FAILED', e)
" $1 $2; }
for rep in 1 2; do
  for v in "$@"; do
    env $VAR=$v timeout 120 python bench.py --steps 300 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_${VAR}_${v}_$TAG.json 2>/dev/null; show gpurun_out/bench_${VAR}_${v}_$TAG.json $VAR=$v
  done
done
