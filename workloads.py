"""Path configurations shared by bench.py and the tests: the BASELINE.json workloads and the helpers that turn a
configuration dict into yolopp_params / the reference's test_cfg. (tests/cases.py adds the parity-only cases.)"""
import numpy as np

from yolopp import _capi as capi
from yolopp import synth as ysynth

V4_SIZES = [[(12, 16), (19, 36), (40, 28)], [(36, 75), (76, 55), (72, 146)], [(142, 110), (192, 243), (459, 401)]]
V3_SIZES = [[(116, 90), (156, 198), (373, 326)], [(30, 61), (62, 45), (59, 119)], [(10, 13), (16, 30), (33, 23)]]

COCO_NMS = dict(type='nms', iou_threshold=0.65)


def csp_case(img, batch, dist, seed, C=80, nms_pre=1000, score_thr=0.001, nms=None, max_per_img=300, **kw):
    """YOLOCSPHead + YOLOV4BBoxCoder (configs/yolov4, configs/yolov5), COCO eval setting by default."""
    strides = [8, 16, 32]
    d = dict(mode=capi.MODE_CSP, batch=batch, sizes=[(img // s, img // s) for s in strides], strides=strides,
             base_sizes=V4_SIZES, num_classes=C, nms_pre=nms_pre, score_thr=score_thr, nms=dict(nms or COCO_NMS),
             max_per_img=max_per_img, dist=dist, seed=seed)
    d.update(kw)
    return d


def v3_case(img, batch, dist, seed, C=80, nms_pre=1000, score_thr=0.05, conf_thr=0.005, nms=None, max_per_img=100, **kw):
    """YOLOV3Head + YOLOBBoxCoder (configs/yolo/yolov3_d53_mstrain-608_273e_coco.py:48-54)."""
    strides = [32, 16, 8]
    d = dict(mode=capi.MODE_V3, batch=batch, sizes=[(img // s, img // s) for s in strides], strides=strides,
             base_sizes=V3_SIZES, num_classes=C, nms_pre=nms_pre, score_thr=score_thr, conf_thr=conf_thr,
             nms=dict(nms or dict(type='nms', iou_threshold=0.45)), max_per_img=max_per_img, dist=dist, seed=seed)
    d.update(kw)
    return d


# bench.py workloads (BASELINE.json configs[1..4]); `batch` = images per call on one GPU, `global_batch` (strong
# scaling only) = images of the whole job, split over the GPUs and processed `batch` at a time.
WORKLOADS = {
    'yolov4_608_b64_coco_sparse': csp_case(608, 64, 'sparse', 11),      # configs[1]: the headline
    'yolov4_608_b64_dense': csp_case(608, 64, 'dense', 12),             # configs[2]
    'yolov5_640_b128_sparse': csp_case(640, 128, 'sparse', 46),         # configs[3] (i): configs/yolov5 = the same head
    'yolov3_640_b128_sparse': v3_case(640, 128, 'sparse', 47),          # configs[3] (ii): the other decode convention
    'yolov4_1280_b128_sparse': csp_case(1280, 128, 'sparse', 48),       # one GPU's shard of configs[4]
    'yolov4_1280_b1024_sparse': csp_case(1280, 128, 'sparse', 48, global_batch=1024),  # configs[4]: strong scaling
}


def build_params(case, batch=None):
    from yolopp.heads import parse_nms_cfg
    mode = case['mode']
    return capi.make_params(
        mode, batch or case['batch'], case['sizes'], case['strides'], case['strides'], case['base_sizes'],
        case['num_classes'], class_agnostic=case.get('class_agnostic', False), nms_pre=case['nms_pre'],
        score_thr=case['score_thr'], conf_thr=case.get('conf_thr', -1.0) if mode == capi.MODE_V3 else -1.0,
        max_per_img=case['max_per_img'], rescale=case.get('rescale', False), out_capacity=case.get('out_capacity', 0),
        **parse_nms_cfg(case['nms']))


def ref_cfg(case):
    """The reference's test_cfg for this case."""
    cfg = dict(nms_pre=case['nms_pre'], score_thr=case['score_thr'], nms=dict(case['nms']),
               max_per_img=case['max_per_img'], min_bbox_size=0)
    if case['mode'] == capi.MODE_V3:
        cfg['conf_thr'] = case.get('conf_thr', -1)
    return cfg


def scale_factors(case):
    if not case.get('rescale', False):
        return None
    return np.asarray(case['scale_factors'], np.float32)


def host_levels(case, params=None):
    """The case's synthetic head tensors generated on the HOST by the checker's generator (oracle.synth_level; same
    bits as the device generator). Test / CPU-baseline infrastructure: imports oracle/."""
    from oracle import oracle
    p = params or build_params(case)
    if case['dist'] == 'gauss':
        return gauss_levels(case, p)
    if case['dist'] == 'blobs':
        return blob_levels(case, p)
    mean, std = ysynth.dist_stats(case['dist'])
    na = p.num_attrib
    m = np.array([mean[0]] * 4 + [mean[1]] + [mean[2]] * (na - 5), np.float32)
    s = np.array([std[0]] * 4 + [std[1]] + [std[2]] * (na - 5), np.float32)
    out = []
    for l in range(p.num_levels):
        hw = p.height[l] * p.width[l]
        x = oracle.synth_level(p.batch, p.num_anchors, na, hw, m, s, ysynth.level_seed(case['seed'], l))
        out.append(x.reshape(p.level_shape(l)))
    return out


HOST_DISTS = ('gauss', 'blobs')  # generated on the host and uploaded (everything else: the device generator)


def blob_levels(case, p):
    """Detector-like head tensors: a few "objects" per image (centre, extent, class); cells near an object get a high
    objectness logit and a high logit for the object's class on every level and anchor, box logits stay near 0 — so
    the best-scored candidates come in clusters of heavily overlapping boxes of the SAME class, which is what NMS
    exists for (the i.i.d. distributions suppress next to nothing). Deterministic (numpy MT19937)."""
    rng = np.random.RandomState(case['seed'])
    na = p.num_attrib
    C = na - 5
    n_obj = case.get('objects', 12)
    out = []
    B = p.batch
    img_w = p.width[0] * p.stride_w[0]
    img_h = p.height[0] * p.stride_h[0]
    objs = [(rng.uniform(0.1, 0.9, n_obj) * img_w, rng.uniform(0.1, 0.9, n_obj) * img_h,
             rng.uniform(*case.get('extent', (0.04, 0.25)), n_obj) * img_w, rng.randint(0, max(C, 1), n_obj),
             rng.uniform(*case.get('amp', (4.0, 9.0)), n_obj))
            for _ in range(B)]
    for l in range(p.num_levels):
        _, _, H, W = p.level_shape(l)
        ys, xs = np.meshgrid((np.arange(H) + 0.5) * p.stride_h[l], (np.arange(W) + 0.5) * p.stride_w[l], indexing='ij')
        x = np.empty((B, p.num_anchors, na, H, W), np.float32)
        x[:, :, :4] = rng.standard_normal((B, p.num_anchors, 4, H, W)).astype(np.float32) * 0.2
        x[:, :, 4] = rng.standard_normal((B, p.num_anchors, H, W)).astype(np.float32) * 1.0 - 7.0
        x[:, :, 5:] = rng.standard_normal((B, p.num_anchors, C, H, W)).astype(np.float32) * 1.0 - 5.0
        for b in range(B):
            cx, cy, ext, cls, amp = objs[b]
            for o in range(n_obj):
                g = np.exp(-((xs - cx[o]) ** 2 + (ys - cy[o]) ** 2) / (2.0 * ext[o] ** 2)).astype(np.float32)
                x[b, :, 4] += (amp[o] * g)[None]
                if C > 0:
                    x[b, :, 5 + cls[o]] += (1.2 * amp[o] * g)[None]
        out.append(np.ascontiguousarray(x.reshape(p.level_shape(l)), np.float32))
    return out


def gauss_levels(case, p):
    """SURVEY.md §8(d)'s "COCO-like sparse" distribution with TRUE Gaussian tails (the bit-reproducible device
    generator is Irwin-Hall(4), support +-3.46 sigma): box logits N(0,1), objectness N(-5, 2^2), class logits
    N(-4.9, 1.5^2), from numpy's MT19937 (deterministic across platforms) — generated on the host and uploaded."""
    rng = np.random.RandomState(case['seed'])
    na = p.num_attrib
    mean = np.array([0.0] * 4 + [-5.0] + [-4.9] * (na - 5), np.float32)
    std = np.array([1.0] * 4 + [2.0] + [1.5] * (na - 5), np.float32)
    out = []
    for l in range(p.num_levels):
        B, _, H, W = p.level_shape(l)
        z = rng.standard_normal((B, p.num_anchors, na, H * W)).astype(np.float32)
        x = z * std[None, None, :, None] + mean[None, None, :, None]
        out.append(np.ascontiguousarray(x.reshape(p.level_shape(l)), np.float32))
    return out
