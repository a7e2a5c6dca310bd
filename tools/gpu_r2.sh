#!/bin/bash
# Round-2 GPU pass (run through gpurun from the repo root): parity tests, smoke, bench lines of every workload,
# decode timeline. usage: tools/gpu_r2.sh <tag> [quick]
TAG=${1:-r2}
MODE=${2:-full}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu_$TAG.txt
nvidia-smi topo -m >> gpurun_out/gpu_$TAG.txt 2>&1
nproc >> gpurun_out/gpu_$TAG.txt; lscpu | grep -E "NUMA|Model name|Socket" >> gpurun_out/gpu_$TAG.txt
timeout 1000 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x --timeout 240 --timeout-method thread > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest exit=$?" >> gpurun_out/pytest_gpu_$TAG.log
tail -25 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; tail -2 gpurun_out/smoke_$TAG.log
show() { python -c "
import json,sys
try:
    d=json.load(open(sys.argv[1])); r=d['roofline']
    print(sys.argv[2], round(d['value']), 'img/s', round(d['ms_per_step'],4), 'ms/step lat', round(d.get('latency_ms',0),4), 'submit', round(d.get('host_submit_ms_per_step',0),4), 'frac', round(r['frac'],3), {k: (round(v*1e3,1) if v is not None else None) for k,v in r['stage_ms'].items()}, 'verified', d.get('verified'), 'e2e', d['e2e'] and round(d['e2e']['value']))
except Exception as e: print(sys.argv[2], 'FAILED', e)
" $1 $2; }
timeout 300 python bench.py --steps 200 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; show gpurun_out/bench_$TAG.json default200
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench20_$TAG.json 2>> gpurun_out/bench_$TAG.err; show gpurun_out/bench20_$TAG.json default20
if [ "$MODE" = "full" ]; then
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err
timeout 200 python bench.py --layout nhwc --steps 200 --no-cpu-baseline --no-e2e > gpurun_out/bench_nhwc_$TAG.json 2>> gpurun_out/bench_$TAG.err; show gpurun_out/bench_nhwc_$TAG.json nhwc
timeout 200 python bench.py --pipeline-depth 1 --steps 100 --no-cpu-baseline --no-e2e > gpurun_out/bench_depth1_$TAG.json 2>> gpurun_out/bench_$TAG.err; show gpurun_out/bench_depth1_$TAG.json depth1
timeout 200 python bench.py --workload yolov4_608_b64_dense --steps 50 --no-cpu-baseline --no-e2e > gpurun_out/bench_dense_$TAG.json 2>> gpurun_out/bench_$TAG.err; show gpurun_out/bench_dense_$TAG.json dense
for w in yolov5_640_b128_sparse yolov3_640_b128_sparse yolov4_1280_b128_sparse; do
  timeout 300 python bench.py --workload $w --steps 30 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_${w}_$TAG.json 2>> gpurun_out/bench_$TAG.err; show gpurun_out/bench_${w}_$TAG.json $w
done
timeout 400 python bench.py --workload yolov4_1280_b1024_sparse --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_1280_b1024_n1_$TAG.json 2>> gpurun_out/bench_$TAG.err; show gpurun_out/bench_1280_b1024_n1_$TAG.json 1280_b1024_n1
timeout 100 python tools/prof_timeline.py csp608_sparse 64 > gpurun_out/timeline_608_$TAG.txt 2>> gpurun_out/bench_$TAG.err; head -12 gpurun_out/timeline_608_$TAG.txt
fi
tail -5 gpurun_out/bench_$TAG.err
if [ "$MODE" = "full" ]; then
timeout 100 python tools/pipe_timeline.py 3 30 > gpurun_out/pipe_timeline_$TAG.txt 2>&1; tail -6 gpurun_out/pipe_timeline_$TAG.txt
timeout 100 python tools/prof_phases.py > gpurun_out/phases_$TAG.txt 2>&1; head -30 gpurun_out/phases_$TAG.txt
for lib in tools/var/lib_*.so; do n=$(basename $lib .so)
  YOLOPP_LIB=$PWD/$lib timeout 120 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_${n}_$TAG.json 2>/dev/null; show gpurun_out/bench_${n}_$TAG.json $n
done
fi
