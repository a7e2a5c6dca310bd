"""Drop-in mirrors of the NMS entry points that the reference imports by name, on the CUDA library:

    multiclass_nms   mmdet/core/post_processing/bbox_nms.py:7-93
    batched_nms      mmcv.ops.nms.batched_nms   (third party, mmcv-full 1.3.2..1.4.0; call site bbox_nms.py:2,84)
    nms              mmcv.ops.nms.nms

Same argument meaning and return types; results follow the canonical tie order (score desc, index asc).
Not supported (raise): nms types other than 'nms', numpy inputs.
CUDA tensors only — there is no CPU fallback.
"""
import ctypes

import torch

from . import _capi
from .heads import parse_nms_cfg

INT_MAX = 2**31 - 1


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise NotImplementedError('yolopp NMS ops run on CUDA tensors only (no CPU fallback)')


def _run_batched(boxes, scores, idxs, iou_thr, offset, split_thr, class_agnostic, max_num, score_threshold=0.0):
    lib = _capi.load_library()
    assert boxes.size(1) == 4
    assert boxes.size(0) == scores.size(0)
    assert offset in (0, 1)
    _require_cuda(boxes, scores, idxs)
    n = boxes.shape[0]
    b = boxes.to(torch.float32).contiguous()
    s = scores.to(torch.float32).contiguous()
    lab, num_labels = None, 0
    if idxs is not None and n > 0:
        lab = idxs.to(torch.int64).contiguous()
        num_labels = int(lab.max().item()) + 1  # mmcv's batched_nms synchronises too (boxes.max(), torch.unique)
        if int(lab.min().item()) < 0:
            raise ValueError('negative class index')
    cap = max_num if 0 < max_num < n else n
    dets = torch.empty((max(cap, 1), 5), dtype=torch.float32, device=boxes.device)
    keep = torch.empty((max(cap, 1), ), dtype=torch.int64, device=boxes.device)
    cnt = torch.zeros((2, ), dtype=torch.int32, device=boxes.device)
    with torch.cuda.device(boxes.device):
        ws_bytes = lib.yolopp_nms_workspace_bytes(n, 0) if cap > 4096 else 0  # kept list beyond the smem capacity
        ws = torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device=boxes.device) if ws_bytes else None
        rc = lib.yolopp_batched_nms(_p(b), _p(s), _p(lab), n, num_labels, float(iou_thr), float(score_threshold),
                                    int(offset), int(split_thr),
                                    int(bool(class_agnostic)), int(max_num), _p(dets), _p(keep), _p(cnt), _p(ws),
                                    ws.numel() if ws is not None else 0, _stream())
        if ws is not None:
            ws.record_stream(torch.cuda.current_stream())
    _capi.check(rc, 'yolopp_batched_nms')
    k, status = (int(v) for v in cnt.tolist())
    if status:
        raise RuntimeError(f'yolopp_batched_nms: status {status}')
    return dets[:k].to(boxes.dtype), keep[:k]


def nms(boxes, scores, iou_threshold, offset=0, score_threshold=0, max_num=-1):
    """mmcv.ops.nms.nms -> (dets (k,5), inds (k,)) in descending score order."""
    if not isinstance(boxes, torch.Tensor):
        raise NotImplementedError('numpy inputs are not supported')
    return _run_batched(boxes, scores, None, iou_threshold, offset, INT_MAX, True, max_num, score_threshold)


def batched_nms(boxes, scores, idxs, nms_cfg, class_agnostic=False):
    """mmcv.ops.nms.batched_nms -> (dets (k,5), keep (k,))."""
    cfg = parse_nms_cfg(nms_cfg)
    agnostic = bool(cfg['nms_class_agnostic'] or class_agnostic)
    return _run_batched(boxes, scores, idxs, cfg['iou_thr'], cfg['nms_offset'], cfg['split_thr'], agnostic,
                        cfg['nms_max_num'], cfg['nms_score_thr'])


def multiclass_nms(multi_bboxes, multi_scores, score_thr, nms_cfg, max_num=-1, score_factors=None, return_inds=False):
    """mmdet multiclass_nms -> (dets (k,5), labels (k,)[, inds (k,)])."""
    lib = _capi.load_library()
    _require_cuda(multi_bboxes, multi_scores, score_factors)
    cfg = parse_nms_cfg(nms_cfg)
    n = multi_scores.size(0)
    C = multi_scores.size(1) - 1
    per_class = multi_bboxes.shape[1] > 4
    if per_class:
        assert multi_bboxes.shape[1] == 4 * C
    dev = multi_scores.device
    mb = multi_bboxes.to(torch.float32).contiguous()
    ms = multi_scores.to(torch.float32).contiguous()
    sf = score_factors.to(torch.float32).contiguous().view(-1) if score_factors is not None else None
    m_eff = -1
    if max_num > 0:
        m_eff = max_num
    if cfg['nms_max_num'] > 0:
        m_eff = min(m_eff, cfg['nms_max_num']) if m_eff > 0 else cfg['nms_max_num']
    total = n * C
    cap = m_eff if 0 < m_eff < total else total
    dets = torch.empty((max(cap, 1), 5), dtype=torch.float32, device=dev)
    labels = torch.empty((max(cap, 1), ), dtype=torch.int64, device=dev)
    flat = torch.empty((max(cap, 1), ), dtype=torch.int64, device=dev)
    cnt = torch.zeros((3, ), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        ws_bytes = lib.yolopp_nms_workspace_bytes(n, C)
        ws = torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device=dev)
        rc = lib.yolopp_multiclass_nms(_p(mb), int(per_class), _p(ms), n, C, float(score_thr), _p(sf), cfg['iou_thr'],
                                       cfg['nms_score_thr'], cfg['nms_offset'], cfg['split_thr'], int(cfg['nms_class_agnostic']),
                                       cfg['nms_max_num'], int(max_num), _p(dets), _p(labels), _p(flat), _p(cnt[0:2]),
                                       _p(cnt[2:3]), _p(ws), ws.numel(), _stream())
        ws.record_stream(torch.cuda.current_stream())
    _capi.check(rc, 'yolopp_multiclass_nms')
    k, status, ncand = (int(v) for v in cnt.tolist())
    if status:
        raise RuntimeError(f'yolopp_multiclass_nms: status {status}')
    if ncand == 0:
        # the reference returns the (0, 4) boxes tensor here (bbox_nms.py:75-82)
        out = (mb.new_zeros((0, 4)).to(multi_bboxes.dtype), labels[:0])
        return out + (labels[:0], ) if return_inds else out
    dets, labels, flat = dets[:k].to(multi_bboxes.dtype), labels[:k], flat[:k]
    if not return_inds:
        return dets, labels
    # the reference's `keep` indexes the thresholded candidate list: rank of each flat index among the valid ones
    valid = (ms[:, :-1].reshape(-1) > score_thr)
    rank = torch.cumsum(valid.to(torch.int64), 0) - 1
    return dets, labels, rank[flat]
