"""Phase timeline of the per-image kernels (profiling build tools/libyolopp_prof.so, -DYPP_PROFILE)."""
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
case = dict(cases.CASES[sys.argv[1] if len(sys.argv) > 1 else 'csp608_sparse'], batch=int(sys.argv[2]) if len(sys.argv) > 2 else 64)
p = cases.build_params(case)
levels = cases.device_levels(case, p) if case['dist'] in cases.HOST_DISTS else yolopp.synth.synth_levels(p, 11, case['dist'] if isinstance(case['dist'], (str, tuple)) else 'sparse')
s = Session(p)
for _ in range(4): s.run(levels)
torch.cuda.synchronize()
buf = np.zeros((2, 256, 16), np.int64)
lib.yolopp_phase_read.argtypes = [ctypes.c_void_p]
lib.yolopp_phase_read(buf.ctypes.data_as(ctypes.c_void_p))
B = case['batch']
sel = buf[0, :B]; nms = buf[1, :B]
def stats(name, d):
    print('  %-34s median %7.0f  p10 %7.0f  p90 %7.0f  max %7.0f' % (name, np.median(d), np.percentile(d, 10), np.percentile(d, 90), d.max()))
print('select_kernel (cycles)')
stats('stage logits + ord + minmax', sel[:, 1] - sel[:, 0])
stats('   issue async copies', sel[:, 5] - sel[:, 0]); stats('   wait copies', sel[:, 6] - sel[:, 5]); stats('   ord + rank clear', sel[:, 7] - sel[:, 6]); stats('   keys of the stash', sel[:, 8] - sel[:, 2])
stats('sampled pivot + stash', sel[:, 2] - sel[:, 1])
stats('keys + sort', sel[:, 3] - sel[:, 2])
stats('guard + rank scatter', sel[:, 4] - sel[:, 3])
stats('total', sel[:, 4] - sel[:, 0])
print('  kernel span', sel[:, 4].max() - sel[:, 0].min())
print('nms_image_kernel (cycles)')
stats('row_stat reduction', nms[:, 1] - nms[:, 0])
stats('row-best select', nms[:, 2] - nms[:, 1])
stats('chunk 1 select', nms[:, 3] - nms[:, 2])
stats('   staged scan: issue copies', nms[:, 12]); stats('   staged scan: wait', nms[:, 13]); stats('   staged scan: test + stash', nms[:, 14]); stats('   staged scan: passes', nms[:, 15])
stats('chunk 1 staging', nms[:, 4] - nms[:, 3])
stats('groups (+ later chunks)', nms[:, 5] - nms[:, 4])
stats('outputs', nms[:, 6] - nms[:, 5])
stats('total', nms[:, 6] - nms[:, 0])
print('  kernel span', nms[:, 6].max() - nms[:, 0].min())
stats('chunks', nms[:, 7]); stats('groups', nms[:, 8]); stats('candidates', nms[:, 9])
per_group = (nms[:, 5] - nms[:, 4]) / np.maximum(nms[:, 8], 1)
stats('cycles per group', per_group)
stats('  A + B1 per group', nms[:, 10] / np.maximum(nms[:, 8], 1)); stats('  B2 + append per group', nms[:, 11] / np.maximum(nms[:, 8], 1))
sub = np.zeros((256, 32), np.int64)
if hasattr(lib, 'yolopp_sub_read'):
    lib.yolopp_sub_read.argtypes = [ctypes.c_void_p]
    lib.yolopp_sub_read(sub.ctypes.data_as(ctypes.c_void_p))
    sub = sub[:B]
    print('nms first chunk, finer stamps (cycles, median):')
    for a, z, n in ((0, 1, 'pick rows'), (1, 2, 'bulk scan'), (2, 3, 'stage boxes'), (8, 9, 'classes: sizes + offsets'),
                    (9, 10, 'classes: member lists'), (10, 11, 'classes: rank inside class'), (11, 12, 'classes: pair masks'),
                    (12, 13, 'classes: decide'), (13, 14, 'kept: compact'), (14, 15, 'kept: sort'), (15, 16, 'kept: append')):
        if sub[:, z].max() > 0 and sub[:, a].max() > 0: print('    %-36s %7.0f' % (n, np.median(sub[:, z] - sub[:, a])))
if sub[:, 20].max() > 0:
    print('select start-up (cycles from kernel stamp 0, median): ' + ', '.join('%s %d' % (n, np.median(sub[:, i] - sel[:, 0])) for i, n in ((20, 'hist cleared'), (21, 'barrier armed'), (22, 'block barrier'), (23, 'level 0 issued'), (24, 'level 1 issued'), (25, 'level 2 issued'))))
if sub[:, 27].max() > 0:
    for i, n in ((26, 'rows scanned'), (27, 'stash'), (28, 'kept in the chunk'), (29, 'kept after the chunk')):
        print('    first chunk: %-24s median %5d  min %5d  max %5d' % (n, np.median(sub[:, i]), sub[:, i].min(), sub[:, i].max()))
ssp = np.zeros((2, 64, 4, 10), np.int64)
lib.yolopp_ssp_read.argtypes = [ctypes.c_void_p]
lib.yolopp_ssp_read(ssp.ctypes.data_as(ctypes.c_void_p))
names = ['clear', 'pass1 (source scan -> stash)', 'histogram', 'scan + pivot', 'scatter', 'rank']
for k, kn in ((0, 'select_kernel'), (1, 'nms_image_kernel')):
    for call in range(2 if k == 0 else 2):
        d = ssp[k, :, call]
        if d[:, 6].max() == 0: continue
        print('%s select_sorted_prefix call %d: groups %d stash %.0f cnt %.0f' % (kn, call, np.median(d[:, 9]), np.median(d[:, 7]), np.median(d[:, 8])))
        for i, n in enumerate(names[:6]):
            print('    %-30s %7.0f' % (n, np.median(d[:, i + 1] - d[:, i])))
ssp2 = np.zeros((2, 64, 4, 4), np.int64)
lib.yolopp_ssp2_read.argtypes = [ctypes.c_void_p]
lib.yolopp_ssp2_read(ssp2.ctypes.data_as(ctypes.c_void_p))
for k in (0, 1):
    for call in (0, 1):
        d = ssp2[k, :, call]
        print('kernel %d call %d pass1 (thread 0 view): load+test %.0f  vote+slots %.0f  stash writes %.0f  | wait at barrier %.0f' % (
            k, call, np.median(d[:, 0]), np.median(d[:, 1]), np.median(d[:, 2]), np.median(ssp[k, :, call, 2] - d[:, 3])))
