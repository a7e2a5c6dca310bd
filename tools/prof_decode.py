"""Per-tile timeline of the persistent decode kernel (profiling build tools/libyolopp_prof.so, -DYPP_PROFILE)."""
import sys, ctypes
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'mmdet-yolov4_b200'), os.path.join(ROOT, 'tests')]
import numpy as np, torch, cases
from yolopp import _capi
_capi.LIB_PATH = os.path.join(ROOT, 'tools', 'libyolopp_prof.so')
import yolopp
from yolopp.ops import Session
lib = _capi.load_library()
case = dict(cases.CASES['csp608_sparse'], batch=64)
p = cases.build_params(case)
levels = yolopp.synth.synth_levels(p, 11, 'sparse')
s = Session(p)
for _ in range(4): s.run(levels)
torch.cuda.synchronize()
info = s.info
nt = info.tma_tiles
buf = np.zeros((nt, 8), np.int64)
lib.yolopp_prof_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
lib.yolopp_prof_read(buf.ctypes.data_as(ctypes.c_void_p), nt * 8)
t0 = buf[:, 0].min()
iss0, iss1, land, rel, done, smid = [buf[:, i] for i in range(6)]
print('tiles', nt, 'kernel span cycles', done.max() - t0)
gather = np.arange(nt) < 64 * 3 * 6
for name, sel in (('gather', gather), ('tma', ~gather)):
    print(name, 'n=%d' % sel.sum(), ' producer wait-empty %.0f | issue->landed %.0f | landed->released %.0f | released->done %.0f  (medians)' % (
        np.median((iss1 - iss0)[sel]), np.median((land - iss1)[sel]), np.median((rel - land)[sel]), np.median((done - rel)[sel])))
    print('      means: wait-empty %.0f | issue->landed %.0f | landed->released %.0f | released->done %.0f' % ((iss1 - iss0)[sel].mean(), (land - iss1)[sel].mean(), (rel - land)[sel].mean(), (done - rel)[sel].mean()))
    print('      p90: wait-empty %.0f | issue->landed %.0f | landed->released %.0f | released->done %.0f' % (
        np.percentile((iss1 - iss0)[sel], 90), np.percentile((land - iss1)[sel], 90), np.percentile((rel - land)[sel], 90), np.percentile((done - rel)[sel], 90)))
m6, m7 = buf[:, 6], buf[:, 7]; ok = (~gather) & (m7 > 0)
print('   landed->mask %.0f | mask->loads issued %.0f | loads->released %.0f (medians, tiles with admitted anchors)' % (np.median((m6-land)[ok]), np.median((m7-m6)[ok]), np.median((rel-m7)[ok])))
# per CTA: time between consecutive issues
cta = np.arange(nt) % 296
for c in (0, 100, 295):
    idx = np.where(cta == c)[0]
    d = np.diff(iss1[idx])
    print('cta', c, 'tiles', len(idx), 'median issue interval', np.median(d), 'sum', d.sum())
