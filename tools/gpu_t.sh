#!/bin/bash
# parity tests only: tools/gpu_t.sh <tag> [pytest -k expression]
TAG=${1:-t}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 240 --timeout-method thread ${2:+-k "$2"} > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest exit=$?" >> gpurun_out/pytest_gpu_$TAG.log; tail -30 gpurun_out/pytest_gpu_$TAG.log
