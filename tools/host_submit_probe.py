"""Where the host time of one pipelined submit goes (microseconds per call, 2000 calls each)."""
import os, sys, time, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'mmdet-yolov4_b200')]
import torch, workloads, yolopp
from yolopp.ops import Pipeline
case = workloads.WORKLOADS['yolov4_608_b64_coco_sparse']
p = workloads.build_params(case)
inputs = [yolopp.synth.synth_levels(p, 11 + 7 * j, case['dist']) for j in range(2)]
pipe = Pipeline(p, 6)
for i in range(24): pipe.submit(inputs[i % 2])
torch.cuda.synchronize()
N = 40     # few calls per burst: the launch queue never fills, so this is host time, not GPU back-pressure
REPS = 25
def t(f):
    tot = 0.0
    for r in range(REPS):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for i in range(N):
            f(i)
        tot += time.perf_counter() - t0
    torch.cuda.synchronize(); return tot * 1e6 / (N * REPS)
cur = torch.cuda.current_stream()
st = pipe.streams[0]; sess = pipe.sessions[0]; ent = sess._plan(inputs[0], None); sp = ctypes.c_void_p(st.cuda_stream)
print('submit (whole)            %.1f us' % t(lambda i: pipe.submit(inputs[i % 2])))
print('  wait_stream             %.1f us' % t(lambda i: st.wait_stream(cur)))
print('  plan lookup (_plan)     %.1f us' % t(lambda i: sess._plan(inputs[i % 2], None)))
print('  yolopp_plan_run (ctypes)%.1f us' % t(lambda i: sess.lib.yolopp_plan_run(ent[0], sp)))
print('  event record            %.1f us' % t(lambda i: pipe.done[0].record(st)))
# (yolopp_plan_run itself is a cudaGraphLaunch of the call's three kernels)
