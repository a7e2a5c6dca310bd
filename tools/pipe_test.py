"""Feasibility: steps issued round-robin on D streams (D sessions / workspaces) vs one stream."""
import sys, time
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'mmdet-yolov4_b200'), os.path.join(ROOT, 'tests')]
import torch, cases, yolopp
from yolopp.ops import Session
case = dict(cases.CASES['csp608_sparse'], batch=64)
p = cases.build_params(case)
inputs = [yolopp.synth.synth_levels(p, 11 + i, 'sparse') for i in range(2)]
K = 200
for depth in (1, 2, 3):
    sess = [Session(p) for _ in range(depth)]
    streams = [torch.cuda.Stream() for _ in range(depth)]
    def loop(n):
        for i in range(n):
            with torch.cuda.stream(streams[i % depth]):
                sess[i % depth].run(inputs[i % 2])
    loop(10)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for s in streams: s.wait_event(e0)
    loop(K)
    cur = torch.cuda.current_stream()
    for s in streams: cur.wait_stream(s)
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print('depth %d: %.1f us/step  %.0f img/s   (host issue %.1f us/step)' % (depth, ms * 1e3 / K, 64 * K / ms * 1e3, (t1 - t0) * 1e6 / K))
