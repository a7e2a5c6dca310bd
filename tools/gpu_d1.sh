#!/bin/bash
TAG=d1
mkdir -p gpurun_out
show() { python -c "
import json,sys
try:
    d=json.load(open(sys.argv[1])); r=d['roofline']
    print(sys.argv[2], round(d['value']), 'img/s', round(d['ms_per_step'],4), 'ms/step lat', round(d.get('latency_ms',0),4), 'submit', round(d.get('host_submit_ms_per_step',0),4), 'frac', round(r['frac'],3), {k: (round(v*1e3,1) if v is not None else None) for k,v in r['stage_ms'].items()}, 'verified', d.get('verified'))
except Exception as e: print(sys.argv[2], 'FAILED', e)
" $1 $2; }
timeout 300 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x --timeout 200 --timeout-method thread -k "csp1280 or v3_ or nopre or full_size_other" > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest exit=$?" >> gpurun_out/pytest_gpu_$TAG.log; tail -6 gpurun_out/pytest_gpu_$TAG.log
for w in yolov4_1280_b128_sparse yolov3_640_b128_sparse yolov4_608_b64_coco_sparse; do
  timeout 200 python bench.py --workload $w --steps 50 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_${w}_$TAG.json 2>gpurun_out/bench_$TAG.err; show gpurun_out/bench_${w}_$TAG.json $w
done
timeout 100 python tools/prof_phases.py csp1280_sparse 64 > gpurun_out/phases_1280_$TAG.txt 2>&1; head -11 gpurun_out/phases_1280_$TAG.txt
timeout 200 python tools/mish_bench.py gpurun_out/mish_bench_$TAG.json 2>&1 | tail -5
timeout 100 python tools/host_submit_probe.py 2>&1 | tail -7
timeout 300 ncu --set full --clock-control none --import-source on -k regex:decode_dense -s 2 -c 1 -f -o gpurun_out/prof_dense_$TAG python bench.py --workload yolov3_640_b128_sparse --pipeline-depth 1 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-verify > /dev/null 2>>gpurun_out/bench_$TAG.err
ls -la gpurun_out/prof_dense_$TAG.ncu-rep
tail -3 gpurun_out/bench_$TAG.err
