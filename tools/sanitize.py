"""Small end-to-end runs for compute-sanitizer (memcheck / racecheck / synccheck): every kernel, both decode paths."""
import sys
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'mmdet-yolov4_b200'), os.path.join(ROOT, 'tests')]
import numpy as np, torch, cases, yolopp
from oracle import oracle
names = sys.argv[1:] or ['csp_tiny', 'csp_odd', 'csp608_sparse', 'v3_tiny_nopre', 'tencent_agnostic', 'csp_force_global']
for name in names:
    case = cases.CASES[name]
    p = cases.build_params(case)
    levels = cases.device_levels(case, p)
    sf = cases.scale_factors(case)
    out = yolopp.get_bboxes_raw(p, levels, torch.from_numpy(sf).cuda() if sf is not None else None)
    torch.cuda.synchronize()
    orc = oracle.get_bboxes(p, [x.cpu().numpy() for x in levels], sf)
    ok = np.array_equal(out['count'].cpu().numpy(), orc['count'])
    print(name, 'count ok' if ok else 'COUNT MISMATCH', out['count'].cpu().tolist())
rng = np.random.RandomState(0)
b = rng.rand(500, 4).astype(np.float32) * 100; b[:, 2:] += b[:, :2]
d, k = yolopp.batched_nms(torch.from_numpy(b).cuda(), torch.from_numpy(rng.rand(500).astype(np.float32)).cuda(),
                          torch.from_numpy(rng.randint(0, 5, 500)).cuda(), dict(type='nms', iou_threshold=0.5))
print('batched_nms kept', k.numel())
