#!/bin/bash
TAG=${1:-g1}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x --timeout 240 --timeout-method thread > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest exit=$?" >> gpurun_out/pytest_gpu_$TAG.log; tail -6 gpurun_out/pytest_gpu_$TAG.log
bash tools/gpu_ab2.sh $TAG
timeout 100 python tools/prof_phases.py csp608_sparse 64 > gpurun_out/phases_608_$TAG.txt 2>&1; sed -n 11,32p gpurun_out/phases_608_$TAG.txt
