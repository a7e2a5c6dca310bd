#!/usr/bin/env python
"""Generates tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN SOURCE (oracle/refexec.py) — run in the build
container only (needs /root/reference); the GPU box and the CPU test-suite only read the committed vectors.

    python tests/golden/make_golden.py [case ...]

For every case of tests/cases.py::GOLDEN_CASES the reference head (YOLOCSPHead / YOLOV3Head, unmodified files
from /root/reference) runs on the case's synthetic head tensors (bit-reproducible generator, so the inputs are
NOT stored) in two flavours:

  canon  torch.sigmoid / torch.exp swapped (from outside) for the canonical polynomial of DESIGN.md, tie order
         canonicalised (stable top-k).  Everything else is the reference's own arithmetic.  The C oracle and
         the CUDA kernels must reproduce these vectors BIT-EXACTLY (boxes, scores, labels, kept anchors, order).
  asis   the reference exactly as it runs here (torch's own CPU sigmoid/exp; only the tie order canonicalised).
         Boxes / scores must agree within 1e-5 relative (north_star); kept (anchor, class) sets are compared
         exactly and the outcome of that comparison at generation time is stored (`asis_index_equal`).

mmcv.ops.nms is NOT in the reference tree (third party, mmcv-full 1.3.2..1.4.0): refexec restates batched_nms/nms
from upstream and uses torchvision.ops.nms as the inner greedy kernel. That boundary is "parity unpinned".
"""
import contextlib
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, 'mmdet-yolov4_b200'), os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)

import cases  # noqa: E402
from oracle import oracle, refexec  # noqa: E402
from yolopp import _capi as capi  # noqa: E402


def build_ref_head(ref, case):
    cfg = ref.Cfg(cases.ref_cfg(case))
    n_lvl = len(case['sizes'])
    if case['mode'] == capi.MODE_CSP:
        ag = dict(type='YOLOV4AnchorGenerator', base_sizes=case['base_sizes'], strides=case['strides'])
        head = ref.YOLOCSPHead(num_classes=case['num_classes'], in_channels=[8] * n_lvl, anchor_generator=ag,
                               featmap_strides=case['strides'], class_agnostic=case.get('class_agnostic', False),
                               test_cfg=cfg)
    else:
        ag = dict(type='YOLOAnchorGenerator', base_sizes=case['base_sizes'], strides=case['strides'])
        head = ref.YOLOV3Head(num_classes=case['num_classes'], in_channels=[8] * n_lvl, out_channels=[8] * n_lvl,
                              anchor_generator=ag, featmap_strides=case['strides'], test_cfg=cfg)
    return head


def run_reference(case, levels, canonical):
    """-> list per image of dict(dets, labels, rows, anchors?)"""
    ref = refexec.load_reference()
    head = build_ref_head(ref, case)
    B = case['batch']
    sf = cases.scale_factors(case)
    metas = [dict(scale_factor=(sf[b] if sf is not None else 1.0)) for b in range(B)]
    C = 1 if case.get('class_agnostic', False) else case['num_classes']
    topk_rec, nms_rec = [], []
    mod = ref.yolocsp_module if case['mode'] == capi.MODE_CSP else ref.yolo_module
    orig_mnms = mod.multiclass_nms

    def mnms(multi_bboxes, multi_scores, score_thr, nms_cfg, max_num=-1, score_factors=None, return_inds=False):
        assert not return_inds
        valid = (multi_scores[:, :-1].reshape(-1) > score_thr).nonzero(as_tuple=False).squeeze(1)
        out = orig_mnms(multi_bboxes, multi_scores, score_thr, nms_cfg, max_num, score_factors, return_inds=True)
        dets, labels, keep = out
        flat = valid[keep] if keep.numel() else keep
        nms_rec.append(dict(n_rows=multi_scores.shape[0], flat=flat.clone(), ncand=int(valid.numel())))
        return dets, labels

    mod.multiclass_nms = mnms
    ctx = refexec.canonical_transcendentals(oracle) if canonical else contextlib.nullcontext()
    try:
        with refexec.canonical_ties(topk_rec), ctx, torch.no_grad():
            out = head.get_bboxes([torch.from_numpy(x.copy()) for x in levels], metas, rescale=case.get('rescale', False))
    finally:
        mod.multiclass_nms = orig_mnms
    res = []
    for b in range(B):
        dets, labels = out[b]
        d = dets.numpy().astype(np.float32).reshape(-1, dets.shape[1] if dets.ndim == 2 else 5)
        item = dict(dets=d, labels=labels.numpy().astype(np.int64), ncand=nms_rec[b]['ncand'] if b < len(nms_rec) else 0)
        if case['mode'] == capi.MODE_CSP and b < len(nms_rec):
            rows = (nms_rec[b]['flat'] // C).numpy()
            if len(topk_rec) == B:  # a top-k ran for every image: row -> anchor through its indices
                anchors = topk_rec[b].numpy()[rows]
            else:
                anchors = rows
            item['rows'] = rows.astype(np.int32)
            item['anchors'] = anchors.astype(np.int32)
        res.append(item)
    return res


def run_reference_taps(case, levels):
    """Intermediate tensors of the reference (canonical flavour, CSP convention), per image: what enters
    multiclass_nms (yolocsp_head.py:374-376) — `bbox_pred` (R,4), `cls_pred` (R,C) — the candidate mask / scores the
    reference derives from them (bbox_nms.py:42-62) and `topk_inds` (yolocsp_head.py:350-355). SURVEY.md A.3."""
    assert case['mode'] == capi.MODE_CSP
    ref = refexec.load_reference()
    head = build_ref_head(ref, case)
    B = case['batch']
    sf = cases.scale_factors(case)
    metas = [dict(scale_factor=(sf[b] if sf is not None else 1.0)) for b in range(B)]
    topk_rec, rec = [], []
    mod = ref.yolocsp_module
    orig_mnms = mod.multiclass_nms

    def mnms(multi_bboxes, multi_scores, score_thr, nms_cfg, max_num=-1, score_factors=None, return_inds=False):
        scores = multi_scores[:, :-1]
        valid = scores > score_thr                      # bbox_nms.py:54
        if score_factors is not None:
            scores = scores * score_factors[:, None]    # bbox_nms.py:57-62
        tap = torch.where(valid, scores, torch.full_like(scores, float('nan')))
        rec.append(dict(boxes=multi_bboxes.clone(), scores=tap.clone()))
        return orig_mnms(multi_bboxes, multi_scores, score_thr, nms_cfg, max_num, score_factors, return_inds)

    mod.multiclass_nms = mnms
    try:
        with refexec.canonical_ties(topk_rec), refexec.canonical_transcendentals(oracle), torch.no_grad():
            head.get_bboxes([torch.from_numpy(x.copy()) for x in levels], metas, rescale=case.get('rescale', False))
    finally:
        mod.multiclass_nms = orig_mnms
    assert len(rec) == B and len(topk_rec) == B
    return (np.stack([t.numpy().astype(np.int32) for t in topk_rec]),
            np.stack([r['boxes'].numpy() for r in rec]).astype(np.float32),
            np.stack([r['scores'].numpy() for r in rec]).astype(np.float32))


TAP_CASES = ['csp608_sparse', 'csp416_rescale', 'csp_odd', 'csp_saturated']


def main_taps(names):
    """tests/golden/taps_<case>.npz: topk_inds (B,R), boxes bits (B,R,4), scores bits (B,R,C; NaN = no candidate)."""
    for name in names or TAP_CASES:
        case = cases.CASES[name]
        p = cases.build_params(case)
        levels = cases.host_levels(case, p)
        topk, boxes, scores = run_reference_taps(case, levels)
        R = topk.shape[1]
        o_topk, o_boxes, o_scores = oracle.get_taps(p, levels, R, cases.scale_factors(case))
        nan_ref, nan_orc = np.isnan(scores), np.isnan(o_scores)
        ok = (np.array_equal(topk, o_topk) and np.array_equal(boxes.view(np.uint32), o_boxes.view(np.uint32))
              and np.array_equal(nan_ref, nan_orc)
              and np.array_equal(scores[~nan_ref].view(np.uint32), o_scores[~nan_orc].view(np.uint32)))
        np.savez_compressed(os.path.join(HERE, f'taps_{name}.npz'), topk_inds=topk, boxes_bits=boxes.view(np.uint32),
                            scores_bits=np.where(nan_ref, np.uint32(0xFFFFFFFF), scores.view(np.uint32)))
        print(f'taps {name:22s} rows={R} candidates={int((~nan_ref).sum())} oracle_bit_exact={ok}')
        if not ok:
            raise SystemExit(f'{name}: the C oracle taps do not reproduce the reference bit-exactly')


def main_pkl():
    """tests/golden/v3_onnx_pkl.npz: the reference's own fixed input tensors (tests/test_onnx/data/
    yolov3_head_get_bboxes.pkl, head config of tests/test_onnx/test_head.py:103-129) and what its YOLOV3Head.get_bboxes
    returns for them (with_nms=True), both flavours."""
    import pickle
    with open(os.path.join(refexec.REF_ROOT, 'tests/test_onnx/data/yolov3_head_get_bboxes.pkl'), 'rb') as f:
        maps = pickle.load(f)
    levels = [np.ascontiguousarray(t.numpy(), np.float32) for t in maps]
    case = cases.PKL_CASE
    p = cases.build_params(case)
    assert [tuple(x.shape) for x in levels] == [p.level_shape(l) for l in range(p.num_levels)]
    canon = pack(run_reference(case, levels, True), p.capacity)
    asis = pack(run_reference(case, levels, False), p.capacity)
    orc = oracle.get_bboxes(p, levels)
    n = canon['count'][0]
    ok = (np.array_equal(orc['count'], canon['count']) and np.array_equal(orc['labels'][0], canon['labels'][0, :n])
          and np.array_equal(orc['dets'][0].view(np.uint32), canon['dets'][0, :n].view(np.uint32)))
    np.savez_compressed(os.path.join(HERE, 'v3_onnx_pkl.npz'), level0=levels[0], level1=levels[1], level2=levels[2],
                        canon_count=canon['count'], canon_ncand=canon['ncand'], canon_dets_bits=canon['dets'].view(np.uint32),
                        canon_labels=canon['labels'], asis_count=asis['count'], asis_dets=asis['dets'],
                        asis_labels=asis['labels'])
    print(f'v3_onnx_pkl count={canon["count"]} ncand={canon["ncand"]} oracle_bit_exact={ok} '
          f'asis_labels_equal={np.array_equal(canon["labels"], asis["labels"])}')
    if not ok:
        raise SystemExit('v3_onnx_pkl: the C oracle does not reproduce the reference bit-exactly')


def pack(res, cap):
    B = len(res)
    out = dict(count=np.zeros(B, np.int32), ncand=np.zeros(B, np.int32), dets=np.zeros((B, cap, 5), np.float32),
               labels=np.zeros((B, cap), np.int64), anchors=np.full((B, cap), -1, np.int32))
    for b, r in enumerate(res):
        n = r['labels'].shape[0]
        out['count'][b] = n
        out['ncand'][b] = r['ncand']
        if n:
            out['dets'][b, :n] = r['dets'][:, :5]
            out['labels'][b, :n] = r['labels']
            if 'anchors' in r:
                out['anchors'][b, :n] = r['anchors']
    return out


def main():
    if len(sys.argv) > 1 and sys.argv[1] == '--taps':
        assert refexec.available(), 'reference tree not found'
        torch.set_num_threads(8)
        return main_taps(sys.argv[2:])
    if len(sys.argv) > 1 and sys.argv[1] == '--pkl':
        assert refexec.available(), 'reference tree not found'
        torch.set_num_threads(8)
        return main_pkl()
    names = sys.argv[1:] or cases.GOLDEN_CASES
    assert refexec.available(), 'reference tree not found'
    torch.set_num_threads(8)
    for name in names:
        case = cases.CASES[name]
        p = cases.build_params(case)
        levels = cases.host_levels(case, p)
        cap = p.capacity
        t0 = time.time()
        canon = pack(run_reference(case, levels, True), cap)
        asis = pack(run_reference(case, levels, False), cap)
        # how the two flavours relate on this input (stored, asserted by the tests)
        idx_equal = bool(np.array_equal(canon['count'], asis['count']) and np.array_equal(canon['labels'], asis['labels'])
                         and np.array_equal(canon['anchors'], asis['anchors']))
        rel = 0.0
        if idx_equal:
            for b in range(p.batch):
                n = canon['count'][b]
                if n:
                    a, c = asis['dets'][b, :n].astype(np.float64), canon['dets'][b, :n].astype(np.float64)
                    rel = max(rel, float(np.max(np.abs(a - c) / np.maximum(np.abs(c), 1e-3))))
        # cross-check with the C oracle right away
        orc = oracle.get_bboxes(p, levels, cases.scale_factors(case))
        ok = np.array_equal(orc['count'], canon['count'])
        for b in range(p.batch):
            n = canon['count'][b]
            ok = ok and np.array_equal(orc['dets'][b].view(np.uint32), canon['dets'][b, :n].view(np.uint32))
            ok = ok and np.array_equal(orc['labels'][b], canon['labels'][b, :n])
            if case['mode'] == capi.MODE_CSP:
                ok = ok and np.array_equal(orc['anchors'][b], canon['anchors'][b, :n])
        np.savez_compressed(os.path.join(HERE, f'{name}.npz'), canon_count=canon['count'], canon_ncand=canon['ncand'],
                            canon_dets_bits=canon['dets'].view(np.uint32), canon_labels=canon['labels'],
                            canon_anchors=canon['anchors'], asis_count=asis['count'], asis_dets=asis['dets'],
                            asis_labels=asis['labels'], asis_anchors=asis['anchors'],
                            asis_index_equal=np.array(idx_equal), asis_max_rel=np.array(rel),
                            has_anchors=np.array(case['mode'] == capi.MODE_CSP))
        print(f'{name:26s} count={canon["count"]} ncand={canon["ncand"]} oracle_bit_exact={ok} '
              f'asis_index_equal={idx_equal} asis_max_rel={rel:.2e}  ({time.time() - t0:.1f}s)')
        if not ok:
            raise SystemExit(f'{name}: the C oracle does not reproduce the reference bit-exactly')


if __name__ == '__main__':
    main()
