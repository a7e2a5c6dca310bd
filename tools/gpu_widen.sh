#!/bin/bash
# full GPU suite + one bench line per non-headline workload.  usage: tools/gpu_widen.sh <tag>
TAG=${1:-w}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "exit=$?" >> gpurun_out/pytest_gpu_$TAG.log; tail -5 gpurun_out/pytest_gpu_$TAG.log
for w in yolov4_608_b64_dense yolov5_640_b128_sparse yolov3_640_b128_sparse yolov4_1280_b128_sparse; do
  timeout 200 python bench.py --workload $w --steps 30 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_${w}_$TAG.json 2> gpurun_out/bench_${w}_$TAG.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_${w}_$TAG.json')); print('$w', round(d['value']), 'img/s', round(d['ms_per_step'],4), 'ms/step lat', round(d.get('latency_ms',0),4), 'frac', round(d['roofline']['frac'],3), {k: round(v*1e3,1) for k,v in d['roofline']['stage_ms'].items()})" || tail -3 gpurun_out/bench_${w}_$TAG.err
done
