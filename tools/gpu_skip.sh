#!/bin/bash
# which kernel costs the pipelined period what: the same bench with the top-k and / or the NMS launch left out
# (experiment build tools/var/lib_knobs.so, stale results: --no-verify)
TAG=${1:-sk}
mkdir -p gpurun_out
for v in 0 1 2 3; do
  for depth in 6 1; do
  YPP_SKIP=$v YOLOPP_LIB=$PWD/tools/var/lib_knobs.so timeout 120 python bench.py --steps 300 --warmup 5 --pipeline-depth $depth --no-cpu-baseline --no-e2e --no-verify > gpurun_out/bench_skip${v}_d${depth}_$TAG.json 2>/dev/null
  python -c "
import json,sys
d=json.load(open(sys.argv[1])); print('skip', sys.argv[2], 'depth', sys.argv[3], round(d['value']), 'img/s', round(d['ms_per_step']*1e3,1), 'us/step')" gpurun_out/bench_skip${v}_d${depth}_$TAG.json $v $depth
  done
done
