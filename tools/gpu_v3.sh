#!/bin/bash
TAG=${1:-v3}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x --timeout 240 --timeout-method thread -k "v3 or V3 or pkl or full_size" > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest exit=$?" >> gpurun_out/pytest_gpu_$TAG.log; tail -8 gpurun_out/pytest_gpu_$TAG.log
show() { python -c "
import json,sys
try:
    d=json.load(open(sys.argv[1])); r=d['roofline']
    print(sys.argv[2], round(d['value']), 'img/s', round(d['ms_per_step'],4), 'ms/step lat', round(d.get('latency_ms',0),4), 'frac', round(r['frac'],3), {k: (round(v*1e3,1) if v is not None else None) for k,v in r['stage_ms'].items()}, d.get('verified'))
except Exception as e: print(sys.argv[2], 'FAILED', e)
" $1 $2; }
for w in yolov3_640_b128_sparse yolov5_640_b128_sparse; do
  timeout 200 python bench.py --workload $w --steps 50 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_${w}_$TAG.json 2>/dev/null; show gpurun_out/bench_${w}_$TAG.json $w
done
