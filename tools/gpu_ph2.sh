#!/bin/bash
TAG=${1:-ph}; shift
mkdir -p gpurun_out
for c in "$@"; do
  timeout 100 python tools/prof_phases.py $c 64 > gpurun_out/phases_${c}_$TAG.txt 2>&1; echo "== $c"; grep "first chunk:\|  total\|chunks " gpurun_out/phases_${c}_$TAG.txt
done
