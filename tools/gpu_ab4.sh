#!/bin/bash
# A/B of library variants incl. a parity check of each variant (bench --verify compares the last batch with the oracle)
TAG=${1:-ab}
mkdir -p gpurun_out
show() { python -c "
import json,sys
try:
    d=json.load(open(sys.argv[1])); r=d['roofline']
    print(sys.argv[2], round(d['value']), 'img/s', round(d['ms_per_step'],4), 'ms/step lat', round(d.get('latency_ms',0),4), 'frac', round(r['frac'],3), {k: (round(v*1e3,1) if v is not None else None) for k,v in r['stage_ms'].items()}, d.get('verified'))
except Exception as e: print(sys.argv[2], 'FAILED', e)
" $1 $2; }
for rep in 1 2; do
  for lib in tools/var/lib_*.so intree; do
    n=$(basename $lib .so); if [ $lib = intree ]; then L=""; else L=$PWD/$lib; fi
    YOLOPP_LIB=$L timeout 120 python bench.py --steps 300 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_${n}_$TAG.json 2>/dev/null; show gpurun_out/bench_${n}_$TAG.json $n
    YOLOPP_LIB=$L timeout 120 python bench.py --workload yolov5_640_b128_sparse --steps 100 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench640_${n}_$TAG.json 2>/dev/null; show gpurun_out/bench640_${n}_$TAG.json 640/$n
  done
done
