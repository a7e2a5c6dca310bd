#!/bin/bash
mkdir -p gpurun_out
for args in "csp416_rescale 4" "csp416_rescale 2" "csp608_dense 4" "csp608_sparse 64"; do
  echo "=== in-tree $args"; timeout 60 python tools/dbg_quad.py $args 2>&1 | tail -6
done
