#!/bin/bash
TAG=${1:-i1}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x --timeout 240 --timeout-method thread > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest exit=$?" >> gpurun_out/pytest_gpu_$TAG.log; tail -12 gpurun_out/pytest_gpu_$TAG.log
timeout 120 python bench.py --steps 300 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_$TAG.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/bench_$TAG.json')); r=d['roofline']; print(round(d['value']), d['ms_per_step'], d['latency_ms'], r['frac'], r['stage_ms'], d['verified'])"
