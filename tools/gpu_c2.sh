#!/bin/bash
TAG=c2
mkdir -p gpurun_out
show() { python -c "
import json,sys
try:
    d=json.load(open(sys.argv[1])); r=d['roofline']
    print(sys.argv[2], round(d['value']), 'img/s', round(d['ms_per_step'],4), 'ms/step lat', round(d.get('latency_ms',0),4), 'submit', round(d.get('host_submit_ms_per_step',0),4), 'frac', round(r['frac'],3), {k: (round(v*1e3,1) if v is not None else None) for k,v in r['stage_ms'].items()})
except Exception as e: print(sys.argv[2], 'FAILED', e)
" $1 $2; }
for rep in 1 2; do
for lib in intree noquad; do
    if [ $lib = intree ]; then L=""; else L="$PWD/tools/var/lib_noquad.so"; fi
    YOLOPP_LIB=$L timeout 100 python bench.py --steps 300 --warmup 5 --no-cpu-baseline --no-e2e --no-verify > gpurun_out/bench_${lib}_$TAG.json 2>/dev/null; show gpurun_out/bench_${lib}_$TAG.json ${lib}
done
done
timeout 100 python tools/host_submit_probe.py 2>&1 | tail -8
timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x --timeout 240 --timeout-method thread -k "plan_handle or nms_pre or pkl or mish or gather or score_threshold or host_pipeline or pipelined" > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest exit=$?" >> gpurun_out/pytest_gpu_$TAG.log; tail -25 gpurun_out/pytest_gpu_$TAG.log
