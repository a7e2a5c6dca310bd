"""Execute the reference's OWN source files for the hot path.            *** TEST INFRASTRUCTURE ***

Only usable where /root/reference exists (this container); never on the GPU box. It is used by
tests/golden/make_golden.py to produce the committed golden vectors and by tests that cross-check the C
oracle against the real reference when the tree is present.

The reference cannot be imported as a package (``import mmdet`` needs mmcv-full, pycocotools, the compiled
Cython/CUDA extensions — SURVEY.md §8c), so the files on the path are loaded one by one with importlib under
their real dotted names, on top of a minimal stub of the third-party ``mmcv`` package:

  mmdet/core/anchor/{builder,anchor_generator,yolov4_anchor_generator}.py
  mmdet/core/bbox/builder.py, mmdet/core/bbox/coder/{base_bbox_coder,yolov4_bbox_coder,yolo_bbox_coder}.py
  mmdet/core/post_processing/bbox_nms.py, mmdet/core/export/onnx_helper.py
  mmdet/models/dense_heads/{base_dense_head,dense_test_mixins,yolocsp_head,yolo_head}.py

``mmcv.ops.nms`` (third party, mmcv-full 1.3.2..1.4.0, NOT in the tree) is restated here in torch following
upstream mmcv/ops/nms.py (batched_nms / nms / NMSop.forward); the inner greedy kernel is
``torchvision.ops.nms`` — an independent implementation of the same algorithm (strict ``>``, offset 0,
division form), fed boxes already in canonical (score desc, index asc) order.

Nothing here is copied from the reference: the files are read from /root/reference at run time.
"""
import contextlib
import importlib.util
import os
import sys
import types

import torch

REF_ROOT = os.environ.get('YOLOPP_REFERENCE_ROOT', '/root/reference')


def available():
    return os.path.isfile(os.path.join(REF_ROOT, 'mmdet/models/dense_heads/yolocsp_head.py'))


class Cfg(dict):
    """mmcv.Config stand-in: dict with attribute access (the heads use both cfg.get() and cfg.attr)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


# ----------------------------------------------------------------------------------------------------
# mmcv.ops.nms restated (upstream mmcv 1.3.x python/nms.py), canonical stable ordering
# ----------------------------------------------------------------------------------------------------
def _stable_desc_order(scores):
    return torch.sort(scores, descending=True, stable=True)[1]


def _nms_kernel(boxes, scores, iou_threshold, offset):
    """ext_module.nms (nms_cpu): indices kept, descending score order."""
    import torchvision
    if boxes.numel() == 0:
        return torch.zeros((0, ), dtype=torch.long)
    order = _stable_desc_order(scores)
    if offset == 0:
        ramp = torch.arange(order.numel(), 0, -1, dtype=torch.float32)  # strictly decreasing: no ties inside tv
        kept = torchvision.ops.nms(boxes[order].float(), ramp, float(iou_threshold))
        return order[kept]
    # offset == 1: plain python greedy (rare; small inputs only)
    b = boxes[order]
    areas = (b[:, 2] - b[:, 0] + offset) * (b[:, 3] - b[:, 1] + offset)
    n = b.shape[0]
    sel = [True] * n
    for i in range(n):
        if not sel[i]:
            continue
        for j in range(i + 1, n):
            if not sel[j]:
                continue
            w = torch.clamp(torch.min(b[i, 2], b[j, 2]) - torch.max(b[i, 0], b[j, 0]) + offset, min=0)
            h = torch.clamp(torch.min(b[i, 3], b[j, 3]) - torch.max(b[i, 1], b[j, 1]) + offset, min=0)
            inter = w * h
            if inter / (areas[i] + areas[j] - inter) > iou_threshold:
                sel[j] = False
    return order[torch.tensor(sel, dtype=torch.bool)]


def nms(boxes, scores, iou_threshold, offset=0, score_threshold=0, max_num=-1):
    assert boxes.size(1) == 4
    assert boxes.size(0) == scores.size(0)
    assert offset in (0, 1)
    if score_threshold > 0:
        valid_mask = scores > score_threshold
        b, s = boxes[valid_mask], scores[valid_mask]
        valid_inds = torch.nonzero(valid_mask, as_tuple=False).squeeze(dim=1)
    else:
        b, s = boxes, scores
    inds = _nms_kernel(b, s, float(iou_threshold), offset)
    if max_num > 0:
        inds = inds[:max_num]
    if score_threshold > 0:
        inds = valid_inds[inds]
    dets = torch.cat((boxes[inds], scores[inds].reshape(-1, 1)), dim=1)
    return dets, inds


def batched_nms(boxes, scores, idxs, nms_cfg, class_agnostic=False):
    nms_cfg_ = nms_cfg.copy()
    class_agnostic = nms_cfg_.pop('class_agnostic', class_agnostic)
    if class_agnostic:
        boxes_for_nms = boxes
    else:
        max_coordinate = boxes.max()
        offsets = idxs.to(boxes) * (max_coordinate + torch.tensor(1).to(boxes))
        boxes_for_nms = boxes + offsets[:, None]
    nms_type = nms_cfg_.pop('type', 'nms')
    assert nms_type == 'nms'
    split_thr = nms_cfg_.pop('split_thr', 10000)
    if boxes_for_nms.shape[0] < split_thr:
        dets, keep = nms(boxes_for_nms, scores, **nms_cfg_)
        boxes = boxes[keep]
        scores = dets[:, -1]
    else:
        max_num = nms_cfg_.pop('max_num', -1)
        total_mask = scores.new_zeros(scores.size(), dtype=torch.bool)
        scores_after_nms = scores.new_zeros(scores.size())
        for id in torch.unique(idxs):
            mask = (idxs == id).nonzero(as_tuple=False).view(-1)
            dets, keep = nms(boxes_for_nms[mask], scores[mask], **nms_cfg_)
            total_mask[mask[keep]] = True
            scores_after_nms[mask[keep]] = dets[:, -1]
        keep = total_mask.nonzero(as_tuple=False).view(-1)
        scores, inds = scores_after_nms[keep].sort(descending=True, stable=True)
        keep = keep[inds]
        boxes = boxes[keep]
        if max_num > 0:
            keep = keep[:max_num]
            boxes = boxes[:max_num]
            scores = scores[:max_num]
    return torch.cat([boxes, scores[:, None]], -1), keep


# ----------------------------------------------------------------------------------------------------
# mmcv stub + loader
# ----------------------------------------------------------------------------------------------------
class _Registry:

    def __init__(self, name):
        self.name = name
        self._d = {}

    def register_module(self, name=None, force=False, module=None):

        def deco(cls):
            self._d[name or cls.__name__] = cls
            return cls

        return deco(module) if module is not None else deco

    def get(self, k):
        return self._d.get(k)


def _build_from_cfg(cfg, registry, default_args=None):
    args = dict(cfg)
    if default_args:
        for k, v in default_args.items():
            args.setdefault(k, v)
    t = args.pop('type')
    cls = registry.get(t) if isinstance(t, str) else t
    if cls is None:
        raise KeyError(f'{t} is not in the {registry.name} registry')
    return cls(**args)


def _identity_decorator_factory(*a, **k):

    def deco(f):
        return f

    return deco


class _DummyLoss(torch.nn.Module):

    def __init__(self, loss_weight=1.0, **kw):
        super().__init__()
        self.loss_weight = loss_weight


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__path__ = []  # behave like a package so sub-modules can hang off it
    sys.modules[name] = m
    return m


def _load(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF_ROOT, rel))
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


_LOADED = None


def load_reference():
    """Returns a namespace with the reference's own YOLOCSPHead, YOLOV3Head, coders, anchor generators,
    multiclass_nms (loaded from /root/reference) — cached."""
    global _LOADED
    if _LOADED is not None:
        return _LOADED
    if not available():
        raise RuntimeError(f'reference tree not found under {REF_ROOT}')
    import torch.nn as nn

    class BaseModule(nn.Module):

        def __init__(self, init_cfg=None):
            super().__init__()
            self.init_cfg = init_cfg

    class ConvModule(nn.Module):  # YOLOV3Head._init_layers only constructs it; never run here

        def __init__(self, cin, cout, k, padding=0, **kw):
            super().__init__()
            self.conv = nn.Conv2d(cin, cout, k, padding=padding)

    _mod('mmcv', __version__='1.3.8', jit=_identity_decorator_factory,
         is_tuple_of=lambda seq, t: isinstance(seq, tuple) and all(isinstance(x, t) for x in seq))
    _mod('mmcv.utils', Registry=_Registry, build_from_cfg=_build_from_cfg)
    _mod('mmcv.runner', BaseModule=BaseModule, force_fp32=_identity_decorator_factory,
         auto_fp16=_identity_decorator_factory)
    _mod('mmcv.runner.fp16_utils', auto_fp16=_identity_decorator_factory)
    _mod('mmcv.cnn', normal_init=lambda *a, **k: None, ConvModule=ConvModule)
    _mod('mmcv.ops')
    _mod('mmcv.ops.nms', batched_nms=batched_nms, nms=nms)

    mmdet = _mod('mmdet')
    core = _mod('mmdet.core')
    _mod('mmdet.core.anchor')
    _mod('mmdet.core.bbox')
    _mod('mmdet.core.bbox.coder')
    _mod('mmdet.core.bbox.iou_calculators', bbox_overlaps=None)
    _mod('mmdet.core.post_processing')
    _mod('mmdet.core.export')
    _mod('mmdet.models')
    _mod('mmdet.models.dense_heads')

    ab = _load('mmdet.core.anchor.builder', 'mmdet/core/anchor/builder.py')
    ag = _load('mmdet.core.anchor.anchor_generator', 'mmdet/core/anchor/anchor_generator.py')
    ag4 = _load('mmdet.core.anchor.yolov4_anchor_generator', 'mmdet/core/anchor/yolov4_anchor_generator.py')
    bb = _load('mmdet.core.bbox.builder', 'mmdet/core/bbox/builder.py')
    _load('mmdet.core.bbox.coder.base_bbox_coder', 'mmdet/core/bbox/coder/base_bbox_coder.py')
    c4 = _load('mmdet.core.bbox.coder.yolov4_bbox_coder', 'mmdet/core/bbox/coder/yolov4_bbox_coder.py')
    c3 = _load('mmdet.core.bbox.coder.yolo_bbox_coder', 'mmdet/core/bbox/coder/yolo_bbox_coder.py')
    pp = _load('mmdet.core.post_processing.bbox_nms', 'mmdet/core/post_processing/bbox_nms.py')
    ex = _load('mmdet.core.export.onnx_helper', 'mmdet/core/export/onnx_helper.py')
    sys.modules['mmdet.core.export'].get_k_for_topk = ex.get_k_for_topk
    sys.modules['mmdet.core.export'].add_dummy_nms_for_onnx = ex.add_dummy_nms_for_onnx

    def multi_apply(func, *args, **kwargs):
        from functools import partial
        pfunc = partial(func, **kwargs) if kwargs else func
        return tuple(map(list, zip(*map(pfunc, *args))))

    core.__dict__.update(
        build_anchor_generator=ab.build_anchor_generator, build_assigner=bb.build_assigner,
        build_bbox_coder=bb.build_bbox_coder, build_sampler=bb.build_sampler, multi_apply=multi_apply,
        multiclass_nms=pp.multiclass_nms, bbox2result=None, bbox_mapping_back=None, images_to_levels=None)

    HEADS, LOSSES = _Registry('head'), _Registry('loss')
    for n in ('CrossEntropyLoss', 'GIoULoss', 'MSELoss'):
        LOSSES.register_module(name=n, module=type(n, (_DummyLoss, ), {}))
    _mod('mmdet.models.builder', HEADS=HEADS, LOSSES=LOSSES, build_loss=lambda cfg: _build_from_cfg(cfg, LOSSES))
    _mod('mmdet.models.losses', reduce_loss=None)
    _load('mmdet.models.dense_heads.base_dense_head', 'mmdet/models/dense_heads/base_dense_head.py')
    _load('mmdet.models.dense_heads.dense_test_mixins', 'mmdet/models/dense_heads/dense_test_mixins.py')
    hc = _load('mmdet.models.dense_heads.yolocsp_head', 'mmdet/models/dense_heads/yolocsp_head.py')
    h3 = _load('mmdet.models.dense_heads.yolo_head', 'mmdet/models/dense_heads/yolo_head.py')

    _LOADED = types.SimpleNamespace(
        YOLOCSPHead=hc.YOLOCSPHead, YOLOV3Head=h3.YOLOV3Head, YOLOV4BBoxCoder=c4.YOLOV4BBoxCoder,
        YOLOBBoxCoder=c3.YOLOBBoxCoder, YOLOAnchorGenerator=ag.YOLOAnchorGenerator,
        YOLOV4AnchorGenerator=ag4.YOLOV4AnchorGenerator, multiclass_nms=pp.multiclass_nms,
        bbox_nms_module=pp, yolocsp_module=hc, yolo_module=h3, batched_nms=batched_nms, nms=nms, Cfg=Cfg,
        build_anchor_generator=ab.build_anchor_generator)
    return _LOADED


# ----------------------------------------------------------------------------------------------------
# canonicalisation patches (applied from OUTSIDE; the reference files stay unmodified)
# ----------------------------------------------------------------------------------------------------
@contextlib.contextmanager
def canonical_ties(record=None):
    """torch.topk tie order is implementation defined -> stable (value desc, index asc). `record` (a list)
    receives every index tensor returned."""
    orig = torch.Tensor.topk

    def topk(self, k, dim=-1, largest=True, sorted=True):
        assert largest
        v, i = torch.sort(self, dim=dim, descending=True, stable=True)
        v, i = v.narrow(dim, 0, k), i.narrow(dim, 0, k)
        if record is not None:
            record.append(i.clone())
        return v, i

    torch.Tensor.topk = topk
    try:
        yield
    finally:
        torch.Tensor.topk = orig


@contextlib.contextmanager
def canonical_transcendentals(oracle):
    """Swap torch's sigmoid/exp (host-ISA dependent SIMD approximations) for the canonical polynomial of
    oracle.c while the reference code runs, so every OTHER operation of the path can be pinned bit-exactly."""
    o_sig_t, o_sig_f, o_exp = torch.Tensor.sigmoid, torch.sigmoid, torch.exp

    def sig(x):
        return torch.from_numpy(oracle.sigmoid(x.detach().contiguous().numpy())).reshape(x.shape)

    def exp(x):
        return torch.from_numpy(oracle.expf(x.detach().contiguous().numpy())).reshape(x.shape)

    torch.Tensor.sigmoid, torch.sigmoid, torch.exp = sig, sig, exp
    try:
        yield
    finally:
        torch.Tensor.sigmoid, torch.sigmoid, torch.exp = o_sig_t, o_sig_f, o_exp
