#!/bin/bash
# usage: tools_gpu_round.sh <tag>   — runs on the GPU box (via gpurun): tests, bench, ncu launch list + captures
TAG=${1:-x}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu_$TAG.txt
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest exit=$?" >> gpurun_out/pytest_gpu_$TAG.log
tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 3000 gpurun_out/bench_$TAG.json
timeout 300 python bench.py --workload yolov4_608_b64_dense --steps 20 --no-cpu-baseline --no-e2e > gpurun_out/bench_dense_$TAG.json 2>> gpurun_out/bench_$TAG.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>> gpurun_out/bench_$TAG.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_tma -s 3 -c 1 -f -o gpurun_out/prof_decode_$TAG python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>> gpurun_out/bench_$TAG.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nms_image -s 3 -c 1 -f -o gpurun_out/prof_nms_$TAG python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>> gpurun_out/bench_$TAG.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:select_kernel -s 3 -c 1 -f -o gpurun_out/prof_select_$TAG python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>> gpurun_out/bench_$TAG.err
ls -la gpurun_out | tail -20
