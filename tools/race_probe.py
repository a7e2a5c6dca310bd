"""Run-to-run reproducibility probe: python tools/race_probe.py [batch] [runs] [case]
Prints one fingerprint per run (so a cold first run, a rare glitch and a steady disagreement can be told apart)
and, for every run that is not in the majority, the images that differ from the majority result."""
import sys, hashlib, collections
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'mmdet-yolov4_b200'), os.path.join(ROOT, 'tests')]
import numpy as np, torch, cases, yolopp
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
R = int(sys.argv[2]) if len(sys.argv) > 2 else 30
name = sys.argv[3] if len(sys.argv) > 3 else 'csp608_sparse'
case = dict(cases.CASES[name], batch=B)
p = cases.build_params(case)
levels = yolopp.synth.synth_levels(p, 12, case['dist'])
torch.cuda.synchronize()
outs, fps = [], []
for r in range(R):
    out = yolopp.get_bboxes_raw(p, levels); torch.cuda.synchronize()
    cur = {k: v.cpu().numpy().copy() for k, v in out.items()}
    h = hashlib.md5()
    for k in sorted(cur): h.update(cur[k].tobytes())
    outs.append(cur); fps.append(h.hexdigest()[:8])
cnt = collections.Counter(fps)
maj = cnt.most_common(1)[0][0]
ref = outs[fps.index(maj)]
print('lib', yolopp._capi.LIB_PATH.split('/')[-1], 'batch', B, 'runs', R, 'distinct', len(cnt), 'majority', cnt[maj])
for r, f in enumerate(fps):
    if f != maj:
        imgs = [i for i in range(B) if not np.array_equal(ref['dets'][i], outs[r]['dets'][i]) or ref['num_candidates'][i] != outs[r]['num_candidates'][i]]
        print('  run', r, 'differs in images', imgs[:16])
