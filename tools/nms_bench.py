"""Standalone NMS entries at RPN-like sizes: yolopp.batched_nms / yolopp.nms against
torchvision.ops.batched_nms / nms on the same CUDA tensors (event-timed, results compared).
    python tools/nms_bench.py [out.txt]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'mmdet-yolov4_b200')]
import numpy as np, torch, torchvision, yolopp

lines = []


def timed(f, iters=10):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us


rng = np.random.RandomState(0)
for n, ncls, max_num, thr in ((1000, 80, 100, 0.5), (2000, 5, 1000, 0.7), (6000, 5, 1000, 0.7), (12000, 5, 2000, 0.7), (12000, 1, 2000, 0.7)):
    xy = rng.rand(n, 2).astype(np.float32) * 1000
    wh = rng.rand(n, 2).astype(np.float32) * 120 + 4
    b = torch.from_numpy(np.concatenate([xy, xy + wh], 1)).cuda()
    s = torch.from_numpy((rng.permutation(n).astype(np.float32) + 1) / n).cuda()
    i = torch.from_numpy(rng.randint(0, ncls, n)).cuda()
    cfg = dict(type='nms', iou_threshold=thr, max_num=max_num)
    d, k = yolopp.batched_nms(b, s, i, dict(cfg))
    ref = torchvision.ops.batched_nms(b, s, i, thr)[:max_num]
    same = bool(torch.equal(k, ref))
    t_own = timed(lambda: yolopp.batched_nms(b, s, i, dict(cfg)))
    t_tv = timed(lambda: torchvision.ops.batched_nms(b, s, i, thr)[:max_num])
    lines.append(f'batched_nms n={n:6d} classes={ncls:3d} max_num={max_num:5d} iou={thr}: kept {k.numel():5d}  yolopp {t_own:8.1f} us   torchvision {t_tv:8.1f} us   same result: {same}')
    print(lines[-1], flush=True)
if len(sys.argv) > 1:
    open(sys.argv[1], 'w').write('\n'.join(lines) + '\n')
