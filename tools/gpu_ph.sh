#!/bin/bash
# phase clocks of the per-image kernels (profiling build): tools/gpu_ph.sh <tag>
TAG=${1:-ph}
mkdir -p gpurun_out
timeout 100 python tools/prof_phases.py csp608_sparse 64 > gpurun_out/phases_608_$TAG.txt 2>&1; cat gpurun_out/phases_608_$TAG.txt
