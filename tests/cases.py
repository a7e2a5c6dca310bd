"""Shared table of path configurations used by the parity tests, the golden generator and the benchmark.

Every case is a dict understood by `build_params` (-> yolopp_params) and `ref_cfg` (-> the reference's
test_cfg dict). Inputs are synthetic, from the bit-reproducible generator (yolopp.synth / oracle.synth_level).
"""
import numpy as np

from yolopp import _capi as capi
from yolopp import synth as ysynth

V4_SIZES = [[(12, 16), (19, 36), (40, 28)], [(36, 75), (76, 55), (72, 146)], [(142, 110), (192, 243), (459, 401)]]
V3_SIZES = [[(116, 90), (156, 198), (373, 326)], [(30, 61), (62, 45), (59, 119)], [(10, 13), (16, 30), (33, 23)]]
TENCENT_SIZES = [[(8, 8)], [(16, 16)], [(32, 32)], [(64, 64)]]

COCO_NMS = dict(type='nms', iou_threshold=0.65)


def _csp(img, batch, dist, seed, C=80, nms_pre=1000, score_thr=0.001, nms=None, max_per_img=300, **kw):
    strides = [8, 16, 32]
    d = dict(mode=capi.MODE_CSP, batch=batch, sizes=[(img // s, img // s) for s in strides], strides=strides,
             base_sizes=V4_SIZES, num_classes=C, nms_pre=nms_pre, score_thr=score_thr, nms=dict(nms or COCO_NMS),
             max_per_img=max_per_img, dist=dist, seed=seed)
    d.update(kw)
    return d


def _v3(img, batch, dist, seed, C=80, nms_pre=1000, score_thr=0.05, conf_thr=0.005, nms=None, max_per_img=100, **kw):
    strides = [32, 16, 8]
    d = dict(mode=capi.MODE_V3, batch=batch, sizes=[(img // s, img // s) for s in strides], strides=strides,
             base_sizes=V3_SIZES, num_classes=C, nms_pre=nms_pre, score_thr=score_thr, conf_thr=conf_thr,
             nms=dict(nms or dict(type='nms', iou_threshold=0.45)), max_per_img=max_per_img, dist=dist, seed=seed)
    d.update(kw)
    return d


SAT = ((0.0, 0.0, 0.0), (30.0, 30.0, 30.0))        # saturated logits: sigmoid hits 0/1 exactly -> many exact ties
MIDSPARSE = ((0.0, -3.0, -3.0), (1.0, 2.0, 2.0))

# name -> case. Sizes are chosen so that the CPU oracle finishes each in seconds.
CASES = {
    # BASELINE.json configs 1-3 at reduced batch
    'csp608_sparse': _csp(608, 2, 'sparse', 11),
    'csp608_dense': _csp(608, 2, 'dense', 12),
    # single-problem regime of batched_nms (n < split_thr), mostly separable by class
    'csp608_dense_thr07': _csp(608, 2, 'dense', 13, score_thr=0.7),
    'csp608_sparse_thr002': _csp(608, 2, 'sparse', 13, score_thr=0.02),
    'csp608_sparse_thr01': _csp(608, 3, 'sparse', 14, score_thr=0.1),
    # the fork's own config: no objectness top-k (nms_pre=-1) -> >1024 boxes per class, chunked NMS stream
    'csp320_nopre_dense': _csp(320, 2, 'dense', 15, nms_pre=-1),
    'csp320_nopre_sparse': _csp(320, 2, 'sparse', 16, nms_pre=-1),
    # tiny maps, odd sizes (no 16-byte aligned plane on any level)
    'csp_tiny': _csp(64, 3, 'dense', 17, C=4, nms_pre=20),
    'csp_odd': dict(mode=capi.MODE_CSP, batch=2, sizes=[(9, 13), (5, 7), (3, 3)], strides=[8, 16, 32],
                    base_sizes=V4_SIZES, num_classes=7, nms_pre=50, score_thr=0.001, nms=dict(COCO_NMS),
                    max_per_img=30, dist='dense', seed=18),
    # nms_pre == N, nms_pre > N (no top-k), nms_pre = N-1
    'csp_pre_eq_n': _csp(64, 2, 'dense', 19, C=3, nms_pre=252),
    'csp_pre_nm1': _csp(64, 2, 'dense', 20, C=3, nms_pre=251),
    # rescale=True with per-image scale factors
    'csp416_rescale': _csp(416, 2, 'sparse', 21, rescale=True,
                           scale_factors=[[0.65, 0.65, 0.65, 0.65], [1.3, 0.8671875, 1.3, 0.8671875]]),
    # empty output: nothing passes the threshold
    'csp_empty': _csp(128, 2, 'sparse', 22, score_thr=0.99999),
    # exact ties: saturated logits
    'csp_saturated': _csp(128, 2, SAT, 23, C=8, nms_pre=100),
    'csp_saturated_nopre': _csp(64, 2, SAT, 24, C=8, nms_pre=-1),
    # single class
    'csp_single_class': _csp(128, 2, 'dense', 25, C=1, nms_pre=200),
    # class agnostic head (configs/tencent/tencent_traffic_sign_yolov4l.py:16-38)
    'tencent_agnostic': dict(mode=capi.MODE_CSP, batch=2, sizes=[(40, 40), (20, 20), (10, 10), (5, 5)],
                             strides=[4, 8, 16, 32], base_sizes=TENCENT_SIZES, num_classes=1, class_agnostic=True,
                             nms_pre=-1, score_thr=0.3, nms=dict(type='nms', iou_threshold=0.1), max_per_img=300,
                             dist=MIDSPARSE, seed=26),
    # nms_cfg variants
    'csp_nms_agnostic': _csp(128, 2, 'dense', 27, C=6, nms_pre=300,
                             nms=dict(type='nms', iou_threshold=0.5, class_agnostic=True)),
    'csp_nms_agnostic_split': _csp(128, 2, 'dense', 28, C=6, nms_pre=300,
                                   nms=dict(type='nms', iou_threshold=0.5, class_agnostic=True, split_thr=500)),
    'csp_nms_offset1': _csp(128, 2, 'dense', 29, C=6, nms_pre=300, nms=dict(type='nms', iou_threshold=0.5, offset=1)),
    'csp_nms_maxnum': _csp(128, 2, 'dense', 30, C=6, nms_pre=300, nms=dict(type='nms', iou_threshold=0.5, max_num=17)),
    'csp_force_split': _csp(128, 2, 'dense', 31, C=6, nms_pre=300, nms=dict(type='nms', iou_threshold=0.5, split_thr=10)),
    'csp_force_global': _csp(160, 2, 'dense', 32, C=20, nms_pre=500,
                             nms=dict(type='nms', iou_threshold=0.5, split_thr=1000000)),
    'csp_keep_all': _csp(64, 2, 'dense', 33, C=3, nms_pre=100, max_per_img=-1, out_capacity=400),
    # YOLOv3 convention (BASELINE.json config 4 (ii))
    'v3_416_sparse': _v3(416, 2, 'sparse', 41),
    'v3_416_dense': _v3(416, 2, 'dense', 42),
    'v3_320_mid': _v3(320, 3, MIDSPARSE, 43, score_thr=0.3),
    'v3_tiny_nopre': _v3(96, 2, 'dense', 44, C=4, nms_pre=-1, conf_thr=-1),
    'v3_rescale': _v3(256, 2, MIDSPARSE, 45, C=12, rescale=True,
                      scale_factors=[[1.5, 1.5, 1.5, 1.5], [0.4, 0.6, 0.4, 0.6]]),
    # BASELINE.json config 4 (i): the yolov5 configs build the same YOLOCSPHead at 640^2
    'csp640_sparse': _csp(640, 2, 'sparse', 46),
    # BASELINE.json config 4 (ii): YOLOv3 convention at 640^2 (per-level top-k, conf_thr, score_factors)
    'v3_640_sparse': _v3(640, 2, 'sparse', 47),
    # BASELINE.json config 5: 1280^2 (100 800 anchors per image: the objectness top-k runs its exact path)
    'csp1280_sparse': _csp(1280, 2, 'sparse', 48),
}

# subset that is also frozen as golden vectors produced by the reference's own source (tests/golden)
GOLDEN_CASES = ['csp608_sparse', 'csp608_dense', 'csp608_dense_thr07', 'csp608_sparse_thr002', 'csp320_nopre_sparse', 'csp_odd', 'csp416_rescale',
                'csp_saturated', 'tencent_agnostic', 'csp_nms_agnostic', 'csp_nms_offset1', 'csp_nms_maxnum',
                'csp_force_global', 'v3_416_sparse', 'v3_416_dense', 'v3_320_mid', 'v3_rescale', 'csp640_sparse',
                'csp_empty', 'v3_640_sparse', 'csp1280_sparse']


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
    """The case's synthetic head tensors generated on the HOST (oracle.synth_level; same bits as the device)."""
    from oracle import oracle
    p = params or build_params(case)
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


def asis_rel_err(ref_dets, got_dets):
    """Largest relative deviation between reference-as-it-runs detections and ours, (n,5) each. north_star's bar
    is 1e-5 relative (fp32). A box corner is `centre -/+ half-size` (yolov4_bbox_coder.py:62-65), so a corner near 0
    is a difference of two large numbers: a 1-ulp sigmoid difference (torch's SIMD sigmoid vs the canonical
    polynomial) is an ABSOLUTE error of ~1 ulp of the operands, not of the corner. Coordinates are therefore
    measured against the magnitude of the box they belong to (max |coordinate| of that detection); the score column
    is measured against itself."""
    import numpy as np
    a = np.asarray(ref_dets, np.float64).reshape(-1, 5)
    o = np.asarray(got_dets, np.float64).reshape(-1, 5)
    if a.shape[0] == 0:
        return 0.0
    scale = np.maximum(np.abs(a[:, :4]).max(axis=1, keepdims=True), 1e-3)
    box = (np.abs(a[:, :4] - o[:, :4]) / np.maximum(np.abs(a[:, :4]), scale)).max()
    score = (np.abs(a[:, 4] - o[:, 4]) / np.maximum(np.abs(a[:, 4]), 1e-3)).max()
    return float(max(box, score))
