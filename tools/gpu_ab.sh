#!/bin/bash
# A/B: every tools/var/lib_*.so (earlier / alternative builds) vs the in-tree build, headline workload.
# usage: tools/gpu_ab.sh <tag>
TAG=${1:-ab}
mkdir -p gpurun_out
show() { python -c "
import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[2], round(d['value']), 'img/s', round(d['ms_per_step'],4), 'ms/step lat', round(d.get('latency_ms',0),4), 'frac', round(d['roofline']['frac'],3), {k: round(v*1e3,1) for k,v in d['roofline']['stage_ms'].items()})" $1 $2; }
for rep in 1 2; do
  for lib in tools/var/lib_*.so; do
    n=$(basename $lib .so)
    YOLOPP_LIB=$PWD/$lib timeout 120 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_${n}_$TAG.json 2>/dev/null; show gpurun_out/bench_${n}_$TAG.json $n
  done
  timeout 120 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_tree_$TAG.json 2>/dev/null; show gpurun_out/bench_tree_$TAG.json in-tree
done
