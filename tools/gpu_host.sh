#!/bin/bash
# is the pipelined loop host-bound? host submit time vs step period at several depths; 1-GPU torchrun smoke of bench
for d in 2 3 4 6; do
  timeout 120 python bench.py --steps 300 --warmup 5 --pipeline-depth $d --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('depth', d['config']['pipeline_depth'], round(d['value']), 'img/s  step', round(d['ms_per_step'],4), 'ms  host submit', round(d['host_submit_ms_per_step'],4), 'ms')"
done
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 1 --steps 50 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-300
