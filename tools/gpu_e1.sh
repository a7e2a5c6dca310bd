#!/bin/bash
TAG=e1
mkdir -p gpurun_out
show() { python -c "
import json,sys
try:
    d=json.load(open(sys.argv[1])); r=d['roofline']
    print(sys.argv[2], round(d['value']), 'img/s', round(d['ms_per_step'],4), 'ms/step lat', round(d.get('latency_ms',0),4), 'submit', round(d.get('host_submit_ms_per_step',0),4), 'frac', round(r['frac'],3), {k: (round(v*1e3,1) if v is not None else None) for k,v in r['stage_ms'].items()}, 'verified', d.get('verified'))
except Exception as e: print(sys.argv[2], 'FAILED', e)
" $1 $2; }
timeout 1000 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x --timeout 240 --timeout-method thread > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest exit=$?" >> gpurun_out/pytest_gpu_$TAG.log; tail -8 gpurun_out/pytest_gpu_$TAG.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for w in yolov4_608_b64_coco_sparse yolov3_640_b128_sparse; do
  timeout 200 python bench.py --workload $w --steps 200 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_${w}_$TAG.json 2>gpurun_out/bench_$TAG.err; show gpurun_out/bench_${w}_$TAG.json $w
done
timeout 100 python tools/host_submit_probe.py 2>&1 | tail -7
echo "mish in-tree"; timeout 100 python tools/mish_bench.py gpurun_out/mish_bench_$TAG.json 2>&1 | python -c "
import sys,ast
for l in sys.stdin:
    try: d=ast.literal_eval(l); print(d['case'], 'fwd', d['fwd_frac_of_copy_peak'], 'bwd', d['bwd_frac_of_copy_peak'], 'copy', round(d['copy_gbs_here']))
    except Exception: pass"
for lib in tools/var/lib_DMISH*.so; do echo $lib; YOLOPP_LIB=$PWD/$lib timeout 100 python tools/mish_bench.py 2>&1 | python -c "
import sys,ast
for l in sys.stdin:
    try: d=ast.literal_eval(l); print(d['case'], 'fwd', d['fwd_frac_of_copy_peak'], 'bwd', d['bwd_frac_of_copy_peak'])
    except Exception: pass"; done
tail -3 gpurun_out/bench_$TAG.err
