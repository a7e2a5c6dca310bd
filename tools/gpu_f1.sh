#!/bin/bash
TAG=${1:-f1}
mkdir -p gpurun_out
show() { python -c "
import json,sys
try:
    d=json.load(open(sys.argv[1])); r=d['roofline']
    print(sys.argv[2], round(d['value']), 'img/s', round(d['ms_per_step'],4), 'ms/step lat', round(d.get('latency_ms',0),4), 'submit', round(d.get('host_submit_ms_per_step',0),4), 'frac', round(r['frac'],3), {k: (round(v*1e3,1) if v is not None else None) for k,v in r['stage_ms'].items()}, 'verified', d.get('verified'))
except Exception as e: print(sys.argv[2], 'FAILED', e)
" $1 $2; }
timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x --timeout 200 --timeout-method thread -k "matches_oracle or golden or full_size or run_to_run or pipelined or select_exact" > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest exit=$?" >> gpurun_out/pytest_gpu_$TAG.log; tail -6 gpurun_out/pytest_gpu_$TAG.log
for rep in 1 2; do
timeout 200 python bench.py --steps 200 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_$TAG.json 2>gpurun_out/bench_$TAG.err; show gpurun_out/bench_$TAG.json default
timeout 200 python bench.py --steps 200 --warmup 3 --pipeline-depth 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_d1_$TAG.json 2>gpurun_out/bench_$TAG.err; show gpurun_out/bench_d1_$TAG.json depth1
done
timeout 100 python tools/prof_timeline.py csp608_sparse 64 > gpurun_out/timeline_608_$TAG.txt 2>&1; head -12 gpurun_out/timeline_608_$TAG.txt
tail -3 gpurun_out/bench_$TAG.err
