"""'Stock GPU' comparison point: the reference's algorithm restated with stock torch CUDA ops and
torchvision.ops.batched_nms, batched decode + per-image loop exactly like YOLOCSPHead.get_bboxes /
_get_bboxes_single / multiclass_nms (yolocsp_head.py:225-382, bbox_nms.py:7-93). Not a parity artefact (library
sigmoid, library tie order) and not the product: a number to put next to bench.py's.
    python tools/stock_gpu.py [batch] [steps]"""
import sys, time
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'mmdet-yolov4_b200'), os.path.join(ROOT, 'tests')]
import torch, torchvision
import cases, yolopp

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
K = int(sys.argv[2]) if len(sys.argv) > 2 else 5
case = dict(cases.CASES['csp608_sparse'], batch=B)
p = cases.build_params(case)
levels = yolopp.synth.synth_levels(p, case['seed'], case['dist'])
dev = levels[0].device
strides = case['strides']
anchors = []
for (h, w), s, sizes in zip(case['sizes'], strides, case['base_sizes']):
    ys, xs = torch.meshgrid(torch.arange(h, device=dev), torch.arange(w, device=dev), indexing='ij')
    cx = (xs.reshape(-1, 1).float() + 0.5) * s
    cy = (ys.reshape(-1, 1).float() + 0.5) * s
    wh = torch.tensor(sizes, dtype=torch.float32, device=dev)
    a = torch.stack([cx - wh[:, 0] / 2, cy - wh[:, 1] / 2, cx + wh[:, 0] / 2, cy + wh[:, 1] / 2], -1).reshape(-1, 4)
    anchors.append(a)


def get_bboxes(levels):
    cls_l, conf_l, box_l = [], [], []
    for x, a, s in zip(levels, anchors, strides):
        pred = x.permute(0, 2, 3, 1).reshape(B, -1, 85).sigmoid()
        xy = pred[..., :2] * 2 - 1
        wh = (pred[..., 2:4] * 2) ** 2
        ac = (a[:, :2] + a[:, 2:]) * 0.5
        aw = a[:, 2:] - a[:, :2]
        c = xy * s + ac
        half = wh * aw * 0.5
        box_l.append(torch.cat([c - half, c + half], -1))
        conf_l.append(pred[..., 4])
        cls_l.append(pred[..., 5:])
    cls, conf, box = torch.cat(cls_l, 1), torch.cat(conf_l, 1), torch.cat(box_l, 1)
    out = []
    for b in range(B):
        _, idx = conf[b].topk(case['nms_pre'])
        bx, sc = box[b, idx], cls[b, idx] * conf[b, idx, None]
        valid = sc > case['score_thr']
        r, c = valid.nonzero(as_tuple=True)
        keep = torchvision.ops.batched_nms(bx[r], sc[r, c], c, case['nms']['iou_threshold'])[:case['max_per_img']]
        out.append((torch.cat([bx[r[keep]], sc[r[keep], c[keep], None]], -1), c[keep]))
    return out


for _ in range(2):
    res = get_bboxes(levels)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(K):
    res = get_bboxes(levels)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / K
print('stock torch CUDA + torchvision.ops.batched_nms: %.1f ms per batch of %d -> %.0f img/s (dets of image 0: %d)' % (dt * 1e3, B, B / dt, len(res[0][1])))
