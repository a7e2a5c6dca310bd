#!/bin/bash
TAG=${1:-tl}
mkdir -p gpurun_out
for c in v3_640_sparse csp640_sparse; do
  timeout 100 python tools/prof_timeline.py $c 128 > gpurun_out/timeline_${c}_$TAG.txt 2>&1; echo "== $c"; cat gpurun_out/timeline_${c}_$TAG.txt
  timeout 100 python tools/prof_decode.py $c 128 > gpurun_out/decode_${c}_$TAG.txt 2>&1; cat gpurun_out/decode_${c}_$TAG.txt | head -30
done
