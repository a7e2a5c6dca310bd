"""ctypes wrapper of oracle/liboracle.so (oracle.c).        *** TEST INFRASTRUCTURE — see oracle.c ***

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes
import os
import subprocess
from ctypes import c_float, c_int, c_int32, c_int64, c_uint64, c_void_p

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, 'liboracle.so')


def build(force=False):
    src = os.path.join(HERE, 'oracle.c')
    if force or not os.path.isfile(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(['make', '-C', HERE, '-B', 'liboracle.so'], check=True, capture_output=True)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB):
            build()
        L = ctypes.CDLL(LIB)
        L.oracle_expf.restype = c_float
        L.oracle_expf.argtypes = [c_float]
        L.oracle_sigmoid.restype = c_float
        L.oracle_sigmoid.argtypes = [c_float]
        L.oracle_expf_array.argtypes = [c_void_p, c_void_p, c_int64]
        L.oracle_sigmoid_array.argtypes = [c_void_p, c_void_p, c_int64]
        L.oracle_coder_decode.argtypes = [c_int, c_void_p, c_void_p, c_float, c_int64, c_void_p]
        L.oracle_grid_anchors.argtypes = [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]
        L.oracle_nms.restype = c_int64
        L.oracle_nms.argtypes = [c_void_p, c_void_p, c_int64, c_float, c_int, c_void_p]
        L.oracle_batched_nms2.restype = c_int64
        L.oracle_batched_nms2.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_float, c_int, c_int, c_int, c_int,
                                          c_float, c_void_p, c_void_p]
        L.oracle_multiclass_nms.restype = c_int64
        L.oracle_multiclass_nms.argtypes = [c_void_p, c_int, c_void_p, c_int64, c_int, c_float, c_float, c_int, c_int,
                                            c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                            c_void_p]
        L.oracle_get_bboxes.restype = c_int
        L.oracle_get_bboxes.argtypes = [c_void_p, ctypes.POINTER(c_void_p), c_void_p, c_int64, c_void_p, c_void_p,
                                        c_void_p, c_void_p, c_void_p, c_void_p, c_int]
        L.oracle_get_taps.restype = c_int
        L.oracle_get_taps.argtypes = [c_void_p, ctypes.POINTER(c_void_p), c_void_p, c_int64, c_void_p, c_void_p, c_void_p]
        L.oracle_mish_fwd_array.argtypes = [c_void_p, c_void_p, c_int64]
        L.oracle_mish_bwd_array.argtypes = [c_void_p, c_void_p, c_void_p, c_int64]
        L.oracle_synth_level.argtypes = [c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_uint64]
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a):
    return a.ctypes.data_as(c_void_p) if a is not None else None


def expf(x):
    x = _f32(x)
    out = np.empty_like(x)
    lib().oracle_expf_array(_ptr(x), _ptr(out), x.size)
    return out


def sigmoid(x):
    x = _f32(x)
    out = np.empty_like(x)
    lib().oracle_sigmoid_array(_ptr(x), _ptr(out), x.size)
    return out


def coder_decode(mode, bboxes, pred, stride):
    b, p = _f32(bboxes).reshape(-1, 4), _f32(pred).reshape(-1, 4)
    assert b.shape == p.shape
    out = np.empty_like(b)
    lib().oracle_coder_decode(int(mode), _ptr(b), _ptr(p), float(stride), b.shape[0], _ptr(out))
    return out.reshape(np.shape(pred))


def grid_anchors(base, H, W, stride_w, stride_h):
    base = _f32(base).reshape(-1, 4)
    out = np.empty((H * W * base.shape[0], 4), np.float32)
    lib().oracle_grid_anchors(_ptr(base), base.shape[0], H, W, stride_w, stride_h, _ptr(out))
    return out


def nms(boxes, scores, iou_thr, offset=0):
    b, s = _f32(boxes).reshape(-1, 4), _f32(scores).reshape(-1)
    keep = np.empty(max(s.size, 1), np.int64)
    k = lib().oracle_nms(_ptr(b), _ptr(s), s.size, float(iou_thr), int(offset), _ptr(keep))
    return keep[:k].copy()


def batched_nms(boxes, scores, idxs, iou_thr, offset=0, split_thr=10000, class_agnostic=False, max_num=-1,
                score_threshold=0.0):
    b, s = _f32(boxes).reshape(-1, 4), _f32(scores).reshape(-1)
    i = None if idxs is None else np.ascontiguousarray(idxs, dtype=np.int64)
    n = s.size
    dets = np.empty((max(n, 1), 5), np.float32)
    keep = np.empty(max(n, 1), np.int64)
    k = lib().oracle_batched_nms2(_ptr(b), _ptr(s), _ptr(i), n, float(iou_thr), int(offset), int(split_thr),
                                  int(bool(class_agnostic)), int(max_num), float(score_threshold), _ptr(dets), _ptr(keep))
    return dets[:k].copy(), keep[:k].copy()


def multiclass_nms(multi_bboxes, multi_scores, score_thr, iou_thr, max_num=-1, score_factors=None, offset=0,
                   split_thr=10000, class_agnostic=False, nms_max_num=-1):
    ms = _f32(multi_scores)
    n, C = ms.shape[0], ms.shape[1] - 1
    mb = _f32(multi_bboxes).reshape(n, -1)
    per_class = int(mb.shape[1] > 4)
    sf = None if score_factors is None else _f32(score_factors).reshape(-1)
    cap = max(n * C, 1)
    dets = np.empty((cap, 5), np.float32)
    labels = np.empty(cap, np.int64)
    inds = np.empty(cap, np.int64)
    flat = np.empty(cap, np.int64)
    ncand = np.zeros(1, np.int64)
    k = lib().oracle_multiclass_nms(_ptr(mb), per_class, _ptr(ms), n, C, float(score_thr), float(iou_thr), int(offset),
                                    int(split_thr), int(bool(class_agnostic)), int(nms_max_num), int(max_num),
                                    _ptr(sf), _ptr(dets), _ptr(labels), _ptr(inds), _ptr(flat), _ptr(ncand))
    return dets[:k].copy(), labels[:k].copy(), inds[:k].copy(), flat[:k].copy(), int(ncand[0])


def get_bboxes(params, levels, scale_factors=None, num_threads=0):
    """Host mirror of yolopp_get_bboxes. `params` is a yolopp_params ctypes struct, `levels` a list of numpy
    arrays (B, A*(5+C), H, W). Returns dict of per-image lists."""
    B = params.batch
    lv = [_f32(x) for x in levels]
    cap = params.out_capacity if params.out_capacity > 0 else params.max_per_img
    if cap <= 0:
        cap = sum(params.height[l] * params.width[l] for l in range(params.num_levels)) * params.num_anchors * \
            (1 if params.class_agnostic else params.num_classes)
    ptrs = (c_void_p * len(lv))(*[x.ctypes.data for x in lv])
    sf = None if scale_factors is None else _f32(scale_factors).reshape(B, 4)
    dets = np.zeros((B, cap, 5), np.float32)
    labels = np.zeros((B, cap), np.int64)
    anchors = np.zeros((B, cap), np.int32)
    rows = np.zeros((B, cap), np.int32)
    count = np.zeros(B, np.int32)
    ncand = np.zeros(B, np.int32)
    rc = lib().oracle_get_bboxes(ctypes.byref(params), ptrs, _ptr(sf), cap, _ptr(dets), _ptr(labels), _ptr(anchors),
                                 _ptr(rows), _ptr(count), _ptr(ncand), int(num_threads))
    if rc != 0:
        raise RuntimeError(f'oracle_get_bboxes failed: {rc}')
    return dict(dets=[dets[b, :count[b]].copy() for b in range(B)], labels=[labels[b, :count[b]].copy() for b in range(B)],
                anchors=[anchors[b, :count[b]].copy() for b in range(B)], rows=[rows[b, :count[b]].copy() for b in range(B)],
                count=count, num_candidates=ncand)


def synth_level(batch, num_anchors, num_attrib, hw, mean, std, seed):
    out = np.empty((batch, num_anchors * num_attrib, hw), np.float32)
    m, s = _f32(mean), _f32(std)
    assert m.size == num_attrib and s.size == num_attrib
    lib().oracle_synth_level(_ptr(out), batch, num_anchors, num_attrib, hw, _ptr(m), _ptr(s), c_uint64(seed))
    return out


def get_taps(params, levels, rows, scale_factors=None):
    """Host mirror of yolopp_topk_conf / yolopp_decode (SURVEY.md A.3 taps): topk_inds (B,R) int32, boxes (B,R,4),
    scores (B,R,C) with NaN where (row, class) is not a candidate."""
    B = params.batch
    C = 1 if params.class_agnostic else params.num_classes
    lv = [_f32(x) for x in levels]
    ptrs = (c_void_p * len(lv))(*[x.ctypes.data for x in lv])
    sf = None if scale_factors is None else _f32(scale_factors).reshape(B, 4)
    topk = np.zeros((B, rows), np.int32)
    boxes = np.zeros((B, rows, 4), np.float32)
    scores = np.zeros((B, rows, C), np.float32)
    rc = lib().oracle_get_taps(ctypes.byref(params), ptrs, _ptr(sf), rows, _ptr(topk), _ptr(boxes), _ptr(scores))
    assert rc == 0
    return topk, boxes, scores


def mish_forward(x):
    x = _f32(x)
    out = np.empty_like(x)
    lib().oracle_mish_fwd_array(_ptr(x), _ptr(out), x.size)
    return out


def mish_backward(grad_out, x):
    g, x = _f32(grad_out), _f32(x)
    out = np.empty_like(x)
    lib().oracle_mish_bwd_array(_ptr(g), _ptr(x), _ptr(out), x.size)
    return out
