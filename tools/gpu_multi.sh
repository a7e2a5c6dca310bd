#!/bin/bash
# Multi-GPU pass on ONE 8-GPU box: both workloads at N = 2, 4, 8 (torchrun, one rank per GPU), H2D topology.
TAG=${1:-m8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$TAG.txt 2>&1
show() { python -c "
import json,sys
try:
    d=json.load(open(sys.argv[1])); e=d.get('e2e') or {}
    print(sys.argv[2], 'N', d['n_gpus'], round(d['value']), 'img/s', round(d['ms_per_step'],4), 'ms/step', d['scaling'], 'verified', d.get('verified'), 'e2e', e.get('value') and round(e['value']), e.get('ms_per_step') and round(e['ms_per_step'],2), e.get('h2d_gbs_per_rank'))
except Exception as ex: print(sys.argv[2], 'FAILED', ex)
" $1 $2; }
P=29500
for N in 8 4 2; do
  P=$((P+1))
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/bench_n${N}_$TAG.json 2> gpurun_out/bench_n${N}_$TAG.err; show gpurun_out/bench_n${N}_$TAG.json default
  P=$((P+1))
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --workload yolov4_1280_b1024_sparse --steps 5 --warmup 3 --no-e2e > gpurun_out/bench_1280_b1024_n${N}_$TAG.json 2> gpurun_out/bench_1280_n${N}_$TAG.err; show gpurun_out/bench_1280_b1024_n${N}_$TAG.json 1280_b1024
done
timeout 120 python tools/h2d_topology.py gpurun_out/h2d_topology_$TAG.json 2>&1 | tail -14
tail -2 gpurun_out/bench_n8_$TAG.err
