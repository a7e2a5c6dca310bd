#!/bin/bash
# bench every tools/var_*.so variant (decode experiments); prints value / stage times
for f in tools/var_*.so; do
  YOLOPP_LIB=$PWD/$f timeout 120 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-e2e > /tmp/vb.json 2>/tmp/vb.err
  python -c "
import json; d=json.load(open('/tmp/vb.json')); print('$f', round(d['value']), d['ms_per_step'], round(d['roofline']['frac'],3), {k: round(v*1e3,1) for k,v in d['roofline']['stage_ms'].items()})" || tail -3 /tmp/vb.err
done
