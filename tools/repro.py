"""Run-to-run reproducibility of the whole path at full size: python tools/repro.py [batch] [runs]"""
import sys
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'mmdet-yolov4_b200'), os.path.join(ROOT, 'tests')]
import numpy as np, torch, cases, yolopp
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
R = int(sys.argv[2]) if len(sys.argv) > 2 else 10
case = dict(cases.CASES['csp608_sparse'], batch=B)
p = cases.build_params(case)
nbad = 0
for seed in (11, 12, 13):
    levels = yolopp.synth.synth_levels(p, seed, case['dist'])
    def run():
        out = yolopp.get_bboxes_raw(p, levels); torch.cuda.synchronize()
        return {k: v.cpu().numpy().copy() for k, v in out.items()}
    a = run()
    for r in range(R):
        b = run()
        if not all(np.array_equal(a[k], b[k]) for k in a):
            nbad += 1
            print('seed', seed, 'run', r, 'differs in images', [i for i in range(B) if not np.array_equal(a['dets'][i], b['dets'][i]) or a['num_candidates'][i] != b['num_candidates'][i]])
print('batch', B, 'runs', 3 * R, 'differing:', nbad)
