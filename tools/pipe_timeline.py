"""Kernel timeline of the pipelined region (what bench.py times): the stage events of yolopp_plan_run_profiled on
every pipeline stream, placed on one time axis. Shows when select / decode / NMS of neighbouring batches start and
finish relative to each other — i.e. which kernel the step period is waiting for.
    python tools/pipe_timeline.py [depth] [steps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'mmdet-yolov4_b200')]
import numpy as np, torch, workloads, yolopp
from yolopp import _capi
from yolopp.ops import Pipeline

depth = int(sys.argv[1]) if len(sys.argv) > 1 else 3
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
case = workloads.WORKLOADS['yolov4_608_b64_coco_sparse']
p = workloads.build_params(case)
inputs = [yolopp.synth.synth_levels(p, 11 + 7 * j, case['dist']) for j in range(2)]
pipe = Pipeline(p, depth)
for i in range(4 * depth):
    pipe.submit(inputs[i % 2])
torch.cuda.synchronize()
# per-step event sets (5 stage events each)
evs = [[torch.cuda.Event(enable_timing=True) for _ in range(_capi.NUM_STAGE_EVENTS)] for _ in range(steps)]
for es in evs:
    for e in es:
        e.record()
torch.cuda.synchronize()
ev0 = torch.cuda.Event(enable_timing=True)
cur = torch.cuda.current_stream()
ev0.record()
import ctypes
for i in range(steps):
    slot = i % depth
    st = pipe.streams[slot]
    st.wait_stream(cur)
    sess = pipe.sessions[slot]
    ent = sess._plan(inputs[i % 2], None)
    evp = (ctypes.c_void_p * 5)(*[e.cuda_event for e in evs[i]])
    rc = sess.lib.yolopp_plan_run_profiled(ent[0], ctypes.c_void_p(st.cuda_stream), evp, 5)
    assert rc == 0
torch.cuda.synchronize()
T = np.array([[ev0.elapsed_time(e) * 1e3 for e in es] for es in evs])  # us
print('step  slot   select: start..end     decode: start..end     nms: start..end      (us after the first submit)')
for i in range(steps):
    t = T[i]
    print(f'{i:4d}  {i % depth:4d}   {t[0]:8.1f}..{t[1]:8.1f}   {t[1]:8.1f}..{t[3]:8.1f}   {t[3]:8.1f}..{t[4]:8.1f}')
half = steps // 2
print('steady state (second half): period %.1f us; select %.1f  decode %.1f  nms %.1f us (stream-time incl. waiting for SMs)' % (
    (T[-1, 4] - T[half, 4]) / (steps - 1 - half), np.mean(T[half:, 1] - T[half:, 0]), np.mean(T[half:, 3] - T[half:, 1]),
    np.mean(T[half:, 4] - T[half:, 3])))
print('decode(i) start - decode(i-1) end: %.1f us;  select(i) end - decode(i-1) end: %.1f us' % (
    np.mean(T[half:, 1] - T[half - 1:-1, 3]), np.mean(T[half:, 1] - T[half - 1:-1, 3])))
