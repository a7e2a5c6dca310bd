#!/bin/bash
# decode kernel duration against its grid size (experiment build): does the stream need every SM?
TAG=${1:-gr}
mkdir -p gpurun_out
for g in 296 256 222 200 168 148 120; do
  YPP_DEC_GRID=$g YOLOPP_LIB=$PWD/tools/var/lib_knobs.so timeout 120 python bench.py --steps 100 --warmup 5 --pipeline-depth 1 --no-cpu-baseline --no-e2e --no-verify > gpurun_out/bench_grid${g}_$TAG.json 2>/dev/null
  python -c "
import json,sys
d=json.load(open(sys.argv[1])); r=d['roofline']; print('grid', sys.argv[2], round(d['value']), 'img/s', {k: (round(v*1e3,1) if v is not None else None) for k,v in r['stage_ms'].items()})" gpurun_out/bench_grid${g}_$TAG.json $g
done
