"""Host-side mirror of the reference's plugin interface for the path: same class / method names, argument
meaning and error behaviour as

    YOLOCSPHead.get_bboxes           mmdet/models/dense_heads/yolocsp_head.py:225-382
    YOLOV3Head.get_bboxes            mmdet/models/dense_heads/yolo_head.py:171-393
    YOLOV4BBoxCoder.decode           mmdet/core/bbox/coder/yolov4_bbox_coder.py:39-67
    YOLOBBoxCoder.decode             mmdet/core/bbox/coder/yolo_bbox_coder.py:60-89
    YOLOAnchorGenerator / YOLOV4AnchorGenerator (inference part)   mmdet/core/anchor/anchor_generator.py:595-665

The classes hold only what inference post-processing reads (no conv layers, no losses — training is out of
scope). `patch_head(head)` rebinds `get_bboxes` of a live mmdet head instance to this
implementation. All compute runs in the CUDA library (ops.py); nothing here touches pixels.
"""
import types

import numpy as np
import torch

from . import _capi, ops


def _cfg_get(cfg, key, default=None):
    if cfg is None:
        return default
    if isinstance(cfg, dict):
        return cfg.get(key, default)
    return getattr(cfg, key, default) if not hasattr(cfg, 'get') else cfg.get(key, default)


def _cfg_attr(cfg, key):
    """cfg.<key> as the reference accesses it (AttributeError/KeyError when absent)."""
    if isinstance(cfg, dict):
        if key not in cfg:
            raise AttributeError(key)
        return cfg[key]
    return getattr(cfg, key)


def parse_nms_cfg(nms_cfg):
    """nms_cfg dict as consumed by mmcv.ops.nms.batched_nms / nms (mmcv-full 1.3.x)."""
    c = dict(nms_cfg)
    nms_type = c.pop('type', 'nms')
    if nms_type != 'nms':
        raise NotImplementedError(f"nms type '{nms_type}' is not on the accelerated path (only 'nms')")
    out = dict(iou_thr=float(c.pop('iou_threshold')), nms_offset=int(c.pop('offset', 0)),
               split_thr=int(c.pop('split_thr', 10000)), nms_class_agnostic=bool(c.pop('class_agnostic', False)),
               nms_max_num=int(c.pop('max_num', -1)))
    out['nms_score_thr'] = float(c.pop('score_threshold', 0))  # mmcv NMSop prefilter (0 = off)
    if c:
        raise TypeError(f'unexpected nms_cfg keys: {sorted(c)}')
    return out


# ----------------------------------------------------------------------------------------------------
# anchor generators / coders
# ----------------------------------------------------------------------------------------------------
class YOLOAnchorGenerator:
    """Inference part of mmdet's YOLOAnchorGenerator (anchor_generator.py:595-665)."""

    def __init__(self, strides, base_sizes):
        self.strides = [_capi._pair(s) for s in strides]
        self.centers = [(s[0] / 2., s[1] / 2.) for s in self.strides]
        n = len(base_sizes[0])
        self.base_sizes = []
        for per_level in base_sizes:
            assert n == len(per_level)
            self.base_sizes.append([_capi._pair(b) if isinstance(b, (tuple, list)) else (b, b) for b in per_level])
        self._base = _capi.yolo_base_anchors(self.base_sizes, self.strides)
        self.base_anchors = [torch.from_numpy(b.copy()) for b in self._base]

    @property
    def num_levels(self):
        return len(self.base_sizes)

    @property
    def num_base_anchors(self):
        return [b.shape[0] for b in self._base]

    def grid_anchors(self, featmap_sizes, device='cuda'):
        """list[Tensor(H*W*A, 4)], row (y*W+x)*A+a (anchor_generator.py:207-270). API helper only: the decode
        kernel forms anchors in closed form and never materialises them."""
        assert self.num_levels == len(featmap_sizes)
        out = []
        for l, (h, w) in enumerate(featmap_sizes):
            sx = torch.arange(int(w), device=device) * self.strides[l][0]
            sy = torch.arange(int(h), device=device) * self.strides[l][1]
            gx = sx.repeat(int(h))
            gy = sy.repeat_interleave(int(w))
            shifts = torch.stack([gx, gy, gx, gy], dim=-1).to(torch.float32)
            base = self.base_anchors[l].to(device)
            out.append((shifts[:, None, :] + base[None, :, :]).reshape(-1, 4))
        return out


class YOLOV4AnchorGenerator(YOLOAnchorGenerator):
    """mmdet/core/anchor/yolov4_anchor_generator.py — inference uses the inherited grid only;
    responsible_indices (:12-134) is training-only and out of scope."""


class YOLOV4BBoxCoder:
    """mmdet/core/bbox/coder/yolov4_bbox_coder.py"""

    def __init__(self, eps=1e-6):
        self.eps = eps

    def encode(self, bboxes, gt_bboxes, stride):
        raise NotImplementedError

    def decode(self, bboxes, pred_bboxes, stride):
        return ops.coder_decode(bboxes, pred_bboxes, stride, _capi.MODE_CSP)


class YOLOBBoxCoder:
    """mmdet/core/bbox/coder/yolo_bbox_coder.py (decode :60-89; encode is training-only, out of scope)"""

    def __init__(self, eps=1e-6):
        self.eps = eps

    def encode(self, bboxes, gt_bboxes, stride):
        raise NotImplementedError('YOLOBBoxCoder.encode is training-only (out of scope of the accelerated path)')

    def decode(self, bboxes, pred_bboxes, stride):
        return ops.coder_decode(bboxes, pred_bboxes, stride, _capi.MODE_V3)


_ANCHOR_GENERATORS = dict(YOLOAnchorGenerator=YOLOAnchorGenerator, YOLOV4AnchorGenerator=YOLOV4AnchorGenerator)
_CODERS = dict(YOLOV4BBoxCoder=YOLOV4BBoxCoder, YOLOBBoxCoder=YOLOBBoxCoder)


def _build(cfg, table):
    if not isinstance(cfg, dict):
        return cfg
    args = dict(cfg)
    return table[args.pop('type')](**args)


# ----------------------------------------------------------------------------------------------------
# get_bboxes
# ----------------------------------------------------------------------------------------------------
def _scale_factors(img_metas, batch):
    """(B,4) fp32, each row what `bbox_pred.new_tensor(scale_factor)` broadcasts to (yolocsp_head.py:365-366)."""
    sf = np.empty((batch, 4), np.float32)
    for b in range(batch):
        v = np.asarray(img_metas[b]['scale_factor'], dtype=np.float32).reshape(-1)
        if v.size == 1:
            sf[b, :] = v[0]
        elif v.size == 4:
            sf[b, :] = v
        else:
            raise AssertionError('scale_factor must be a scalar or 4 values')
    return torch.from_numpy(sf)


def _run_raw(mode, head, pred_maps, img_metas, cfg, rescale, with_nms):
    cfg = head.test_cfg if cfg is None else cfg
    num_levels = len(pred_maps)
    assert num_levels == head.num_levels
    if not with_nms:
        raise NotImplementedError('with_nms=False (TTA merge path) is not on the accelerated path')
    nms_cfg = _cfg_get(cfg, 'nms', None)
    if nms_cfg is None:
        raise NotImplementedError('test_cfg without `nms` (raw export path) is not on the accelerated path')
    batch = pred_maps[0].shape[0]
    if mode == _capi.MODE_CSP:
        assert len(img_metas) == batch
    num_anchors = head.num_anchors[0] if isinstance(head.num_anchors, (list, tuple)) else head.num_anchors
    for l in range(num_levels):
        assert pred_maps[l].shape[1] == num_anchors * head.num_attrib
    params = head_params(mode, head, [tuple(m.shape[-2:]) for m in pred_maps], batch, cfg, rescale)
    sf = _scale_factors(img_metas, batch) if rescale else None
    if sf is not None:
        sf = sf.to(pred_maps[0].device, non_blocking=True)
    return params, ops.get_bboxes_raw(params, pred_maps, sf)


def _get_bboxes_impl(mode, head, pred_maps, img_metas, cfg, rescale, with_nms):
    params, out = _run_raw(mode, head, pred_maps, img_metas, cfg, rescale, with_nms)
    batch = params.batch
    # one small device->host read: counts + status (the reference syncs dozens of times per image)
    host = _pull(head, params, out, ())['meta'].clone()
    count, ncand = host[:batch], host[batch:2 * batch]
    result = []
    for b in range(batch):
        n = int(count[b])
        if int(ncand[b]) == 0:
            # multiclass_nms returns the (0, 4) boxes tensor when nothing passes score_thr (bbox_nms.py:75-82)
            result.append((out['dets'].new_zeros((0, 4)), out['labels'].new_zeros((0, ))))
        else:
            result.append((out['dets'][b, :n], out['labels'][b, :n]))
    return result


def _host_block(head, out, B, cap, C):
    """Pinned host staging block of a head (one per (B, cap, device)); overwritten by every call — callers get
    copies, never views of it."""
    key = (B, cap, C, out['dets'].device)
    buf = getattr(head, '_yolopp_host', None)
    if buf is None or buf[0] != key:
        buf = (key, dict(dets=torch.empty((B, cap, 5), dtype=torch.float32).pin_memory(),
                         labels=torch.empty((B, cap), dtype=torch.int64).pin_memory(),
                         cls_dets=torch.empty((B, cap, 5), dtype=torch.float32).pin_memory(),
                         cls_offsets=torch.empty((B, C + 1), dtype=torch.int32).pin_memory(),
                         meta=torch.empty((2 * B + 1, ), dtype=torch.int32).pin_memory()))
        head._yolopp_host = buf
    return buf[1]


def _pull(head, params, out, names):
    """ONE stream-ordered batch of device->host copies of the named output blocks (+ counts / status) into pinned
    memory, one synchronize, status check. Returns the pinned block dict."""
    B, cap, C = params.batch, params.capacity, params.eff_classes
    h = _host_block(head, out, B, cap, C)
    for n in names:
        h[n].copy_(out[n], non_blocking=True)
    h['meta'][:B].copy_(out['count'], non_blocking=True)
    h['meta'][B:2 * B].copy_(out['num_candidates'], non_blocking=True)
    h['meta'][2 * B:].copy_(out['status'], non_blocking=True)
    torch.cuda.current_stream().synchronize()
    status = int(h['meta'][-1])
    if status != 0:
        raise RuntimeError(f'yolopp_get_bboxes: {_capi.load_library().yolopp_strerror(status).decode()}')
    return h


def _get_results_host(mode, head, pred_maps, img_metas, cfg, rescale):
    """get_bboxes + bbox2result's device->host step fused: the fixed-capacity block goes to pinned memory in one
    batch of copies. Returns list[(ndarray(n,5) f32, ndarray(n,) i64)] — independent arrays (the reference returns
    fresh tensors per call, so results collected over a dataset loop must not alias the staging block)."""
    params, out = _run_raw(mode, head, pred_maps, img_metas, cfg, rescale, True)
    B = params.batch
    h = _pull(head, params, out, ('dets', 'labels'))
    d, l, cnt = h['dets'].numpy(), h['labels'].numpy(), h['meta'].numpy()[:B]
    return [(d[b, :cnt[b]].copy(), l[b, :cnt[b]].copy()) for b in range(B)]


def _get_bbox_results(mode, head, pred_maps, img_metas, cfg, rescale):
    """The tail of SingleStageDetector.simple_test (mmdet/models/detectors/single_stage.py:102-111):
    `[bbox2result(det_bboxes, det_labels, num_classes) for ...]`. The NMS kernel already wrote the detections grouped
    by label with the group offsets (yolopp_outputs.cls_dets / cls_offsets), so the per-class split is
    `num_classes` zero-copy VIEWS of one small per-image array instead of `num_classes` boolean-mask gathers."""
    params, out = _run_raw(mode, head, pred_maps, img_metas, cfg, rescale, True)
    B, C = params.batch, params.eff_classes
    h = _pull(head, params, out, ('cls_dets', 'cls_offsets'))
    d, off, cnt = h['cls_dets'].numpy(), h['cls_offsets'].numpy(), h['meta'].numpy()[:B]
    results = []
    for b in range(B):
        block = d[b, :cnt[b]].copy()  # one small copy per image; the class arrays below are views of it
        o = off[b]
        results.append([block[o[c]:o[c + 1]] for c in range(C)])
    return results


class _HeadBase:
    _mode = None
    _default_coder = None

    def __init__(self, num_classes, in_channels=None, anchor_generator=None, bbox_coder=None, featmap_strides=None,
                 class_agnostic=False, train_cfg=None, test_cfg=None, **unused):
        self.num_classes = num_classes
        self.in_channels = in_channels
        self.featmap_strides = list(featmap_strides)
        self.train_cfg = train_cfg
        self.test_cfg = test_cfg
        self.class_agnostic = class_agnostic
        self.bbox_coder = _build(bbox_coder or dict(type=self._default_coder), _CODERS)
        self.anchor_generator = _build(anchor_generator, _ANCHOR_GENERATORS)
        assert len(self.anchor_generator.num_base_anchors) == len(self.featmap_strides)
        self.num_anchors = self.anchor_generator.num_base_anchors

    @property
    def num_levels(self):
        return len(self.featmap_strides)

    @property
    def num_attrib(self):
        return 5 if self.class_agnostic else 5 + self.num_classes

    def get_bboxes(self, pred_maps, img_metas, cfg=None, rescale=False, with_nms=True):
        """list[(Tensor(n,5), Tensor(n,))] — same contract as the reference's get_bboxes."""
        return _get_bboxes_impl(self._mode, self, pred_maps, img_metas, cfg, rescale, with_nms)

    def get_results_host(self, pred_maps, img_metas, cfg=None, rescale=False):
        """Host-side results (numpy) with a single device->host copy per batch — what `simple_test` needs
        before `bbox2result` (mmdet/models/detectors/single_stage.py:102-111)."""
        return _get_results_host(self._mode, self, pred_maps, img_metas, cfg, rescale)

    def get_bbox_results(self, pred_maps, img_metas, cfg=None, rescale=False):
        """`bbox_results` of SingleStageDetector.simple_test (single_stage.py:102-111): per image a list of
        `num_classes` arrays (k_c, 5), == [bbox2result(*r, num_classes) for r in get_bboxes(...)], with the
        per-class grouping done on the device."""
        return _get_bbox_results(self._mode, self, pred_maps, img_metas, cfg, rescale)


class YOLOCSPHead(_HeadBase):
    """Post-processing mirror of mmdet's YOLOCSPHead (yolocsp_head.py:54; defaults :83-92)."""
    _mode = _capi.MODE_CSP
    _default_coder = 'YOLOV4BBoxCoder'

    def __init__(self, num_classes, in_channels=None, anchor_generator=None, bbox_coder=None,
                 featmap_strides=(8, 16, 32), class_agnostic=False, train_cfg=None, test_cfg=None, **unused):
        anchor_generator = anchor_generator or dict(
            type='YOLOV4AnchorGenerator',
            base_sizes=[[(12, 16), (19, 36), (40, 28)], [(36, 75), (76, 55), (72, 146)],
                        [(142, 110), (192, 243), (459, 401)]], strides=[8, 16, 32])
        super().__init__(num_classes, in_channels, anchor_generator, bbox_coder, featmap_strides, class_agnostic,
                         train_cfg, test_cfg)


class YOLOV3Head(_HeadBase):
    """Post-processing mirror of mmdet's YOLOV3Head (yolo_head.py:20; defaults :49-59)."""
    _mode = _capi.MODE_V3
    _default_coder = 'YOLOBBoxCoder'

    def __init__(self, num_classes, in_channels=None, out_channels=None, anchor_generator=None, bbox_coder=None,
                 featmap_strides=(32, 16, 8), train_cfg=None, test_cfg=None, **unused):
        anchor_generator = anchor_generator or dict(
            type='YOLOAnchorGenerator',
            base_sizes=[[(116, 90), (156, 198), (373, 326)], [(30, 61), (62, 45), (59, 119)],
                        [(10, 13), (16, 30), (33, 23)]], strides=[32, 16, 8])
        super().__init__(num_classes, in_channels, anchor_generator, bbox_coder, featmap_strides, False, train_cfg,
                         test_cfg)
        self.out_channels = out_channels


def patch_head(head):
    """Rebinds `get_bboxes` of a live mmdet YOLOCSPHead / YOLOV3Head instance to the CUDA path. The instance
    keeps its own convs / loss; only inference post-processing changes. Returns the head."""
    name = type(head).__name__
    if name == 'YOLOCSPHead':
        mode = _capi.MODE_CSP
    elif name == 'YOLOV3Head':
        mode = _capi.MODE_V3
    else:
        raise TypeError(f'patch_head: unsupported head type {name}')

    def get_bboxes(self, pred_maps, img_metas, cfg=None, rescale=False, with_nms=True):
        return _get_bboxes_impl(mode, self, [m.detach() for m in pred_maps], img_metas, cfg, rescale, with_nms)

    def get_results_host(self, pred_maps, img_metas, cfg=None, rescale=False):
        return _get_results_host(mode, self, [m.detach() for m in pred_maps], img_metas, cfg, rescale)

    def get_bbox_results(self, pred_maps, img_metas, cfg=None, rescale=False):
        return _get_bbox_results(mode, self, [m.detach() for m in pred_maps], img_metas, cfg, rescale)

    head.get_bboxes = types.MethodType(get_bboxes, head)
    head.get_results_host = types.MethodType(get_results_host, head)
    head.get_bbox_results = types.MethodType(get_bbox_results, head)
    return head


def head_params(mode_or_head, head=None, pred_shapes=None, batch=None, cfg=None, rescale=False):
    """yolopp_params for a head instance (ours or a live mmdet head) and a list of level shapes (H, W) — what
    `get_bboxes` builds internally; exposed for Session / Pipeline / HostPipeline users."""
    if head is None:
        head = mode_or_head
        mode = _capi.MODE_CSP if type(head).__name__ == 'YOLOCSPHead' else _capi.MODE_V3
    else:
        mode = mode_or_head
    cfg = head.test_cfg if cfg is None else cfg
    ag = head.anchor_generator
    return _capi.make_params(
        mode, batch, [tuple(s) for s in pred_shapes], ag.strides, head.featmap_strides, ag.base_sizes,
        head.num_classes, class_agnostic=bool(getattr(head, 'class_agnostic', False)),
        nms_pre=int(_cfg_get(cfg, 'nms_pre', -1)), score_thr=float(_cfg_attr(cfg, 'score_thr')),
        conf_thr=float(_cfg_get(cfg, 'conf_thr', -1)) if mode == _capi.MODE_V3 else -1.0,
        max_per_img=int(_cfg_attr(cfg, 'max_per_img')), rescale=bool(rescale), **parse_nms_cfg(_cfg_get(cfg, 'nms')))


def bbox2result(bboxes, labels, num_classes):
    """mmdet.core.bbox2result (mmdet/core/bbox/transforms.py:99-116): list of `num_classes` arrays (k_c, 5).
    Accepts the host arrays of `get_results_host` (no further device traffic) or device tensors (one D2H each,
    like the reference)."""
    if bboxes.shape[0] == 0:
        return [np.zeros((0, 5), dtype=np.float32) for _ in range(num_classes)]
    if isinstance(bboxes, torch.Tensor):
        bboxes = bboxes.detach().cpu().numpy()
        labels = labels.detach().cpu().numpy()
    return [bboxes[labels == i, :] for i in range(num_classes)]
