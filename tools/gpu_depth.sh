#!/bin/bash
TAG=${1:-dp}
mkdir -p gpurun_out
for d in 2 3 4 6 8 12; do
  timeout 120 python bench.py --steps 300 --warmup 5 --pipeline-depth $d --no-cpu-baseline --no-e2e --no-verify > gpurun_out/bench_depth${d}_$TAG.json 2>/dev/null
  python -c "
import json,sys
d=json.load(open(sys.argv[1])); print('depth', sys.argv[2], round(d['value']), 'img/s', round(d['ms_per_step']*1e3,1), 'us/step')" gpurun_out/bench_depth${d}_$TAG.json $d
done
