"""Whole-kernel timeline of the persistent decode kernel on ONE time axis (profiling build, -DYPP_PROFILE):
where do the microseconds beyond bytes / bandwidth go — ramp-up, steady state, tail?
    python tools/prof_timeline.py [case] [batch]"""
import sys, ctypes
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'mmdet-yolov4_b200'), os.path.join(ROOT, 'tests')]
import numpy as np, torch, cases
from yolopp import _capi
_capi.LIB_PATH = os.environ.get('YPP_PROF_LIB') or os.path.join(ROOT, 'tools', 'libyolopp_prof.so')
import yolopp
from yolopp.ops import Session
lib = _capi.load_library()
name = sys.argv[1] if len(sys.argv) > 1 else 'csp608_sparse'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
case = dict(cases.CASES[name], batch=B)
p = cases.build_params(case)
levels = yolopp.synth.synth_levels(p, 11, case['dist'])
s = Session(p)
for _ in range(4): s.run(levels)
torch.cuda.synchronize()
nt = min(s.info.tma_tiles, 1 << 16)
buf = np.zeros((nt, 8), np.int64)
lib.yolopp_prof_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
lib.yolopp_prof_read(buf.ctypes.data_as(ctypes.c_void_p), nt * 8)
cta = np.zeros((1024, 8), np.int64)
lib.yolopp_prof_cta_read.argtypes = [ctypes.c_void_p]
lib.yolopp_prof_cta_read(cta.ctypes.data_as(ctypes.c_void_p))
G = 296
cta = cta[:G]
# per-CTA clock: ns = gt0 + (clk - clk0) * rate, rate from the CTA's own two (clock64, globaltimer) pairs
rate = (cta[:, 3] - cta[:, 1]) / np.maximum(cta[:, 2] - cta[:, 0], 1)
print('SM clock (cycles/ns) median %.3f' % np.median(1 / rate))
T0 = cta[:, 1].min()
print('CTA start spread: %.2f us; CTA end: first %.2f last %.2f us after kernel start' % ((cta[:, 1].max() - T0) / 1e3, (cta[:, 3].min() - T0) / 1e3, (cta[:, 3].max() - T0) / 1e3))
valid = (buf[:, 1] > 0) & (buf[:, 2] > 0) & (buf[:, 4] > 0)  # gather tiles leave before the last stamps
print('tiles', nt, 'with complete stamps', int(valid.sum()))
buf = buf[valid]
tile_cta = np.minimum((buf[:, 5] >> 16).astype(np.int64), G - 1)  # tiles are claimed dynamically: the consumer records its CTA
def ns(col):
    return cta[tile_cta, 1] + (buf[:, col] - cta[tile_cta, 0]) * rate[tile_cta] - T0
iss, land, rel, done = ns(1), ns(2), ns(3), ns(4)
print('first issue %.2f us, first landed %.2f us, last issue %.2f, last landed %.2f us, last done %.2f us' % (iss.min() / 1e3, land.min() / 1e3, iss.max() / 1e3, land.max() / 1e3, done.max() / 1e3))
NA = p.num_attrib
tile_bytes = NA * (s.info.decode_tile_positions or 64) * 4
end = done.max()
edges = np.arange(0, end + 4000, 4000)
h, _ = np.histogram(land, edges)
print('landed GB/s per 4 us bin:', ' '.join('%.0f' % (x * tile_bytes / 4e-6 / 1e9) for x in h))
h2, _ = np.histogram(iss, edges)
print('issued GB/s per 4 us bin:', ' '.join('%.0f' % (x * tile_bytes / 4e-6 / 1e9) for x in h2))
# per-CTA finishing times (last done of each CTA)
last = np.array([done[tile_cta == c].max() if (tile_cta == c).any() else 0 for c in range(G)])
print('tiles per CTA: min %d max %d' % (np.bincount(tile_cta, minlength=G).min(), np.bincount(tile_cta, minlength=G).max()))
print('per-CTA last tile done: min %.1f p50 %.1f p90 %.1f max %.1f us' % (last.min() / 1e3, np.median(last) / 1e3, np.percentile(last, 90) / 1e3, last.max() / 1e3))
smid = cta[:, 4]
per_sm = {}
for c in range(G): per_sm.setdefault(int(smid[c]), []).append(last[c] / 1e3)
v = sorted((max(x), k) for k, x in per_sm.items())
print('slowest SMs (us, smid):', v[-6:], ' fastest:', v[:4])
