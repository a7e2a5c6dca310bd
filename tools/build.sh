#!/bin/bash
# builds mmdet-yolov4_b200/csrc/libyolopp.so (same flags as __graft_entry__.build())
set -e
cd "$(dirname "$0")/../mmdet-yolov4_b200/csrc"
nvcc -shared -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 "$@" -o libyolopp.so yolopp_capi.cu 2>&1 | grep -E "error|warning|spill|Used" || true
ls -la libyolopp.so | awk '{print $5, $6, $7, $8}'
