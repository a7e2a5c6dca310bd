#!/bin/bash
TAG=${1:-ph}; shift
mkdir -p gpurun_out
for c in "$@"; do
  timeout 100 python tools/prof_phases.py $c 2 > gpurun_out/phases_${c}_$TAG.txt 2>&1; echo "== $c"; grep "first chunk:\|chunks \|groups  \|Error\|error" gpurun_out/phases_${c}_$TAG.txt | head -8
done
