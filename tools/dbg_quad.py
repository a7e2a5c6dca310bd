"""Debug: decode tap vs oracle tap per level for a case (locates wrong rows)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'mmdet-yolov4_b200'), os.path.join(ROOT, 'tests')]
import numpy as np, torch, cases, yolopp
from oracle import oracle
from yolopp import _capi as capi
name = sys.argv[1]; B = int(sys.argv[2])
case = dict(cases.CASES[name], batch=B)
p = cases.build_params(case)
levels = yolopp.synth.synth_levels(p, case['seed'], case['dist'])
info = capi.describe(p)
print(name, 'B', B, 'tma mask', bin(info.tma_level_mask), 'tiles', info.tma_tiles, 'smem', info.decode_smem_bytes, flush=True)
R = info.rows_per_image
sf = cases.scale_factors(case)
if sf is not None: sf = np.tile(sf, (B // len(sf) + 1, 1))[:B]
sf_t = torch.from_numpy(sf).cuda() if sf is not None else None
boxes, scores, inds = yolopp.decode(p, levels, sf_t)
torch.cuda.synchronize()
print('decode done', flush=True)
o_topk, o_boxes, o_scores = oracle.get_taps(p, [x.cpu().numpy() for x in levels], R, sf)
inds = inds.cpu().numpy(); s = scores.cpu().numpy(); bx = boxes.cpu().numpy()
print('topk equal', np.array_equal(inds, o_topk))
nan_o, nan_s = np.isnan(o_scores), np.isnan(s)
bad_rows = np.argwhere((nan_o != nan_s).any(axis=2) | (np.where(nan_o, 0, o_scores).view(np.uint32) != np.where(nan_s, 0, s).view(np.uint32)).any(axis=2))
print('rows with wrong scores:', len(bad_rows), 'of', B * R)
# level of each bad row
offs = np.cumsum([0] + [p.height[l] * p.width[l] * p.num_anchors for l in range(p.num_levels)])
for b, r in bad_rows[:12]:
    n = o_topk[b, r]; l = int(np.searchsorted(offs, n, side='right') - 1); loc = n - offs[l]
    hw, a = divmod(loc, p.num_anchors)
    print(f'  img {b} row {r} anchor {n} level {l} hw {hw} a {a}: got {s[b, r, :3]} want {o_scores[b, r, :3]} box got {bx[b, r]} want {o_boxes[b, r]}')
lv_bad = [int(np.searchsorted(offs, o_topk[b, r], side='right') - 1) for b, r in bad_rows]
print('bad rows per level', np.bincount(lv_bad, minlength=p.num_levels) if lv_bad else 0)
