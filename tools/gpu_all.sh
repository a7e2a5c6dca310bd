#!/bin/bash
# parity tests + every workload, in-tree build against tools/var/lib_*.so
TAG=${1:-all}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x --timeout 240 --timeout-method thread > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest exit=$?" >> gpurun_out/pytest_gpu_$TAG.log; tail -6 gpurun_out/pytest_gpu_$TAG.log
show() { python -c "
import json,sys
try:
    d=json.load(open(sys.argv[1])); r=d['roofline']
    print(sys.argv[2], round(d['value']), 'img/s', round(d['ms_per_step'],4), 'ms/step lat', round(d.get('latency_ms',0),4), 'frac', round(r['frac'],3), {k: (round(v*1e3,1) if v is not None else None) for k,v in r['stage_ms'].items()}, d.get('verified'))
except Exception as e: print(sys.argv[2], 'FAILED', e)
" $1 $2; }
for lib in tools/var/lib_*.so intree; do
  n=$(basename $lib .so); if [ $lib = intree ]; then L=""; else L=$PWD/$lib; fi
  for w in yolov4_608_b64_dense yolov3_640_b128_sparse yolov4_1280_b128_sparse; do
    YOLOPP_LIB=$L timeout 200 python bench.py --workload $w --steps 50 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_${w}_${n}_$TAG.json 2>/dev/null; show gpurun_out/bench_${w}_${n}_$TAG.json $w/$n
  done
done
timeout 100 python tools/prof_timeline.py csp608_sparse 64 > gpurun_out/timeline_608_$TAG.txt 2>&1; sed -n 2,8p gpurun_out/timeline_608_$TAG.txt
