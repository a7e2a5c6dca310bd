import sys
sys.path[:0] = ['/root/repo', '/root/repo/mmdet-yolov4_b200', '/root/repo/tests']
import numpy as np, torch, cases, yolopp
case = dict(cases.CASES['csp608_sparse'], batch=64)
p = cases.build_params(case)
levels = yolopp.synth.synth_levels(p, case['seed'], case['dist'])
def run(pp, lv):
    out = yolopp.get_bboxes_raw(pp, lv); torch.cuda.synchronize()
    return {k: v.cpu().numpy().copy() for k, v in out.items()}
a = run(p, levels); b = run(p, levels)
print('b64 run-to-run equal:', all(np.array_equal(a[k], b[k]) for k in a))
p1 = cases.build_params(case, batch=1)
bad = []
for i in range(64):
    o = run(p1, [x[i:i + 1].contiguous() for x in levels])
    n = int(o['count'][0])
    ok = n == a['count'][i] and np.array_equal(o['dets'][0, :n].view(np.uint32), a['dets'][i, :n].view(np.uint32)) and np.array_equal(o['num_candidates'][0], a['num_candidates'][i])
    if not ok: bad.append((i, n, int(a['count'][i]), int(o['num_candidates'][0]), int(a['num_candidates'][i])))
print('images differing alone vs batch:', bad)
o1 = run(p1, [x[0:1].contiguous() for x in levels]); o2 = run(p1, [x[0:1].contiguous() for x in levels])
print('b1 run-to-run equal:', all(np.array_equal(o1[k], o2[k]) for k in o1))
