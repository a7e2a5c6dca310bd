#!/bin/bash
# quick GPU check: parity tests + bench summary.  usage: tools/gpu_quick.sh <tag>
TAG=${1:-q}
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "exit=$?" >> gpurun_out/pytest_gpu_$TAG.log; tail -4 gpurun_out/pytest_gpu_$TAG.log
timeout 120 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python -c "
import json; d=json.load(open('gpurun_out/bench_$TAG.json')); print(round(d['value']), round(d['ms_per_step'],4), 'lat', round(d.get('latency_ms',0),4), round(d['roofline']['frac'],3), {k: round(v*1e3,1) for k,v in d['roofline']['stage_ms'].items()}, round(d['e2e']['value']))"; tail -3 gpurun_out/bench_$TAG.err
