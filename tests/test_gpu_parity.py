"""GPU parity tests (run on the B200 box: pytest -m gpu). Everything goes through the C ABI of
mmdet-yolov4_b200/csrc/libyolopp.so via the registered torch custom ops, and is compared with the CPU oracle
(oracle/oracle.c) on the same inputs. Bars: indices / labels / counts bit-exact; boxes and scores bit-exact too
(the oracle and the kernels share the canonical fp32 arithmetic), which is stricter than the 1e-5 relative
tolerance north_star states for floating point.
"""
import ctypes

import numpy as np
import pytest
import torch

import cases
from oracle import oracle
from yolopp import _capi as capi

pytestmark = pytest.mark.gpu


def _u32(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def run_cuda(case, params=None, levels=None):
    import yolopp
    p = params or cases.build_params(case)
    if levels is None:
        levels = yolopp.synth.synth_levels(p, case['seed'], case['dist'])
    sf = cases.scale_factors(case)
    sf_t = torch.from_numpy(sf).cuda() if sf is not None else None
    out = yolopp.get_bboxes_raw(p, levels, sf_t)
    torch.cuda.synchronize()
    res = {k: v.cpu().numpy() for k, v in out.items()}
    return p, levels, res


def compare(p, res, orc, name):
    B = p.batch
    assert int(res['status'][0]) == 0, f'{name}: status {res["status"]}'
    np.testing.assert_array_equal(res['num_candidates'], orc['num_candidates'], err_msg=f'{name}: num_candidates')
    np.testing.assert_array_equal(res['count'], orc['count'], err_msg=f'{name}: count')
    for b in range(B):
        n = int(orc['count'][b])
        np.testing.assert_array_equal(res['anchors'][b, :n], orc['anchors'][b], err_msg=f'{name}: anchors img {b}')
        np.testing.assert_array_equal(res['labels'][b, :n], orc['labels'][b], err_msg=f'{name}: labels img {b}')
        np.testing.assert_array_equal(res['rows'][b, :n], orc['rows'][b], err_msg=f'{name}: rows img {b}')
        np.testing.assert_array_equal(_u32(res['dets'][b, :n]), _u32(orc['dets'][b]), err_msg=f'{name}: dets img {b}')


def test_transcendentals_bit_exact():
    """canonical exp / sigmoid: GPU bits == CPU oracle bits on a dense sweep + random bit patterns."""
    import yolopp
    rng = np.random.RandomState(0)
    xs = np.concatenate([
        np.linspace(-110, 95, 1 << 20).astype(np.float32),
        np.linspace(-30, 30, 1 << 21).astype(np.float32),
        rng.randint(0, 2**32, size=1 << 20, dtype=np.uint64).astype(np.uint32).view(np.float32),
        np.array([0.0, -0.0, np.inf, -np.inf, 88.72283, 88.72284, -87.33654, -103.97, 1e-45, -1e-45], np.float32),
    ])
    xs = xs[~np.isnan(xs)]
    x = torch.from_numpy(xs).cuda()
    e = yolopp.exp(x).cpu().numpy()
    s = yolopp.sigmoid(x).cpu().numpy()
    np.testing.assert_array_equal(_u32(e), _u32(oracle.expf(xs)))
    np.testing.assert_array_equal(_u32(s), _u32(oracle.sigmoid(xs)))


def test_synth_bit_exact():
    """device generator == host generator (so full-size device runs are reproducible on the host)."""
    import yolopp
    case = cases.CASES['csp_odd']
    p = cases.build_params(case)
    dev = yolopp.synth.synth_levels(p, case['seed'], case['dist'])
    host = cases.host_levels(case, p)
    for d, h in zip(dev, host):
        np.testing.assert_array_equal(_u32(d.cpu().numpy()), _u32(h))


@pytest.mark.parametrize('name', sorted(cases.CASES))
def test_get_bboxes_matches_oracle(name):
    case = cases.CASES[name]
    p, levels, res = run_cuda(case)
    host = [x.cpu().numpy() for x in levels]
    orc = oracle.get_bboxes(p, host, cases.scale_factors(case))
    compare(p, res, orc, name)


@pytest.mark.parametrize('name', cases.GOLDEN_CASES)
def test_get_bboxes_matches_reference_golden(name):
    """CUDA path == the committed golden vectors that the reference's OWN source produced
    (tests/golden/make_golden.py, canonical transcendental): bit-exact boxes, scores, labels, anchors, order."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', f'{name}.npz'))
    case = cases.CASES[name]
    p, levels, res = run_cuda(case)
    assert int(res['status'][0]) == 0
    np.testing.assert_array_equal(res['count'], g['canon_count'])
    np.testing.assert_array_equal(res['num_candidates'], g['canon_ncand'])
    for b in range(p.batch):
        n = int(g['canon_count'][b])
        np.testing.assert_array_equal(_u32(res['dets'][b, :n]), g['canon_dets_bits'][b, :n])
        np.testing.assert_array_equal(res['labels'][b, :n], g['canon_labels'][b, :n])
        if bool(g['has_anchors']):
            np.testing.assert_array_equal(res['anchors'][b, :n], g['canon_anchors'][b, :n])
        # and within north_star's 1e-5 relative of the reference exactly as it runs (torch CPU sigmoid)
        if n:
            assert cases.asis_rel_err(g['asis_dets'][b, :n], res['dets'][b, :n]) <= 1e-5


def test_coder_decode_matches_oracle():
    import yolopp
    rng = np.random.RandomState(1)
    anchors = (rng.rand(5000, 4).astype(np.float32) * 600)
    anchors[:, 2:] += anchors[:, :2]
    pred = rng.randn(5000, 4).astype(np.float32)
    for mode, coder in ((capi.MODE_CSP, yolopp.YOLOV4BBoxCoder()), (capi.MODE_V3, yolopp.YOLOBBoxCoder())):
        for stride in (8, 16, 32):
            got = coder.decode(torch.from_numpy(anchors).cuda(), torch.from_numpy(pred).cuda(), stride).cpu().numpy()
            np.testing.assert_array_equal(_u32(got), _u32(oracle.coder_decode(mode, anchors, pred, stride)))


def test_reference_coder_kat():
    """The reference's own known-answer test for YOLOBBoxCoder.decode (tests/test_utils/test_coder.py:8-23)."""
    import yolopp
    coder = yolopp.YOLOBBoxCoder()
    bboxes = torch.Tensor([[-42., -29., 74., 61.], [-10., -29., 106., 61.], [22., -29., 138., 61.],
                           [54., -29., 170., 61.]]).cuda()
    pred_bboxes = torch.Tensor([[0.4709, 0.6152, 0.1690, -0.4056], [0.5399, 0.6653, 0.1162, -0.4162],
                                [0.4654, 0.6618, 0.1548, -0.4301], [0.4786, 0.6197, 0.1896, -0.4479]]).cuda()
    expected = torch.Tensor([[-53.6102, -10.3096, 83.7478, 49.6824], [-15.8700, -8.3901, 114.4236, 50.9693],
                             [11.1822, -8.0924, 146.6034, 50.4476], [41.2068, -8.9232, 181.4236, 48.5840]])
    assert expected.allclose(coder.decode(bboxes, pred_bboxes, 32).cpu())


def test_head_api_list_of_tuples():
    """get_bboxes through the reference-shaped head API returns list[(n,5),(n,)] equal to the raw op."""
    import yolopp
    case = cases.CASES['csp608_sparse']
    p = cases.build_params(case)
    levels = yolopp.synth.synth_levels(p, case['seed'], case['dist'])
    head = yolopp.YOLOCSPHead(num_classes=80, in_channels=[256, 512, 1024], test_cfg=cases.ref_cfg(case))
    metas = [dict(scale_factor=1.0) for _ in range(p.batch)]
    out = head.get_bboxes(levels, metas)
    orc = oracle.get_bboxes(p, [x.cpu().numpy() for x in levels])
    assert len(out) == p.batch
    for b, (dets, labels) in enumerate(out):
        assert dets.dtype == torch.float32 and labels.dtype == torch.int64 and dets.is_cuda
        np.testing.assert_array_equal(_u32(dets.cpu().numpy()), _u32(orc['dets'][b]))
        np.testing.assert_array_equal(labels.cpu().numpy(), orc['labels'][b])


def test_cpu_tensors_are_rejected():
    """No CPU fallback: the custom op only has a CUDA kernel."""
    import yolopp
    case = cases.CASES['csp_tiny']
    p = cases.build_params(case)
    host = [torch.from_numpy(x) for x in cases.host_levels(case, p)]
    with pytest.raises((NotImplementedError, RuntimeError)):
        yolopp.get_bboxes_raw(p, host)


def test_full_size_batch64_properties():
    """BASELINE config 2 at full size (608^2, batch 64): size-independent properties — every image of the
    batch equals the same image processed alone (images are independent), counts within bounds, scores
    sorted, and a prefix of the batch matches the oracle."""
    import yolopp
    case = dict(cases.CASES['csp608_sparse'], batch=64)
    p, levels, res = run_cuda(case)
    assert int(res['status'][0]) == 0
    assert (res['count'] <= 300).all() and (res['count'] > 0).all()
    # run-to-run bit-identical (the persistent decode kernel hands tiles to whichever warp is free)
    for _ in range(3):
        _, _, again = run_cuda(case, p, levels)
        for k in res:
            np.testing.assert_array_equal(again[k], res[k], err_msg=f'run-to-run difference in {k}')
    for b in range(64):
        n = res['count'][b]
        s = res['dets'][b, :n, 4]
        assert (np.diff(s) <= 0).all()
    # image b alone == image b inside the batch
    for b in (0, 1, 2, 3, 17, 40, 63):
        p1 = cases.build_params(case, batch=1)
        lv1 = [x[b:b + 1].contiguous() for x in levels]
        out1 = yolopp.get_bboxes_raw(p1, lv1)
        n = int(out1['count'][0])
        assert n == res['count'][b]
        np.testing.assert_array_equal(_u32(out1['dets'][0, :n].cpu().numpy()), _u32(res['dets'][b, :n]))
        np.testing.assert_array_equal(out1['labels'][0, :n].cpu().numpy(), res['labels'][b, :n])
    # first images against the oracle
    host = [x[:2].cpu().numpy() for x in levels]
    p2 = cases.build_params(case, batch=2)
    orc = oracle.get_bboxes(p2, host)
    for b in range(2):
        n = int(orc['count'][b])
        assert n == res['count'][b]
        np.testing.assert_array_equal(_u32(res['dets'][b, :n]), _u32(orc['dets'][b]))
        np.testing.assert_array_equal(res['anchors'][b, :n], orc['anchors'][b])


@pytest.mark.parametrize('name,batch', [('csp608_dense', 64), ('csp640_sparse', 128), ('v3_640_sparse', 128),
                                        ('csp1280_sparse', 128)])
def test_full_size_other_configs(name, batch):
    """BASELINE configs 3, 4 (both decode conventions) and 5 (one GPU's shard of the 1024-image batch) at full
    size: repeatable, sorted, within bounds, sampled images equal the same image processed alone, and the first
    image equals the oracle."""
    import yolopp
    case = dict(cases.CASES[name], batch=batch)
    p, levels, res = run_cuda(case)
    assert int(res['status'][0]) == 0
    cap = case['max_per_img']
    assert (res['count'] <= cap).all() and (res['count'] > 0).all()
    for _ in range(5):
        _, _, again = run_cuda(case, p, levels)
        for k in res:
            np.testing.assert_array_equal(again[k], res[k], err_msg=f'run-to-run difference in {k}')
    for b in range(batch):
        n = res['count'][b]
        assert (np.diff(res['dets'][b, :n, 4]) <= 0).all()
        assert (res['labels'][b, :n] >= 0).all() and (res['labels'][b, :n] < case['num_classes']).all()
    p1 = cases.build_params(case, batch=1)
    for b in (0, 1, batch // 2 + 3, batch - 1):
        out1 = yolopp.get_bboxes_raw(p1, [x[b:b + 1].contiguous() for x in levels])
        n = int(out1['count'][0])
        assert n == res['count'][b]
        np.testing.assert_array_equal(_u32(out1['dets'][0, :n].cpu().numpy()), _u32(res['dets'][b, :n]))
        np.testing.assert_array_equal(out1['labels'][0, :n].cpu().numpy(), res['labels'][b, :n])
    orc = oracle.get_bboxes(p1, [x[:1].cpu().numpy() for x in levels])
    n = int(orc['count'][0])
    assert n == res['count'][0]
    np.testing.assert_array_equal(_u32(res['dets'][0, :n]), _u32(orc['dets'][0]))
    np.testing.assert_array_equal(res['labels'][0, :n], orc['labels'][0])


@pytest.mark.parametrize('batch,runs', [(64, 40), (128, 40)])
def test_run_to_run_reproducible(batch, runs):
    """The decode ring hands a stage from generic-proxy readers (ld.shared) to async-proxy writers (TMA); without
    the cross-proxy fences the first tiles of a launch differed in 3 % (batch 64) / 35 % (batch 128) of the runs.
    Every run must be bit-identical, and the first images (the ones the race hit) must equal the oracle."""
    case = dict(cases.CASES['csp608_sparse'], batch=batch)
    p, levels, res = run_cuda(case)
    for r in range(runs):
        _, _, again = run_cuda(case, p, levels)
        for k in res:
            np.testing.assert_array_equal(again[k], res[k], err_msg=f'run {r}: run-to-run difference in {k}')
    # the schedule for overlapped batches (everything dealt statically) gives the same bits as the hybrid one
    p_static = type(p).from_buffer_copy(p)
    p_static.batches_in_flight = 3
    for r in range(5):
        _, _, other = run_cuda(case, p_static, levels)
        for k in res:
            np.testing.assert_array_equal(other[k], res[k], err_msg=f'static schedule, run {r}: difference in {k}')
    p2 = cases.build_params(case, batch=2)
    orc = oracle.get_bboxes(p2, [x[:2].cpu().numpy() for x in levels])
    for b in range(2):
        n = int(orc['count'][b])
        assert n == res['count'][b]
        np.testing.assert_array_equal(_u32(res['dets'][b, :n]), _u32(orc['dets'][b]))
        np.testing.assert_array_equal(res['anchors'][b, :n], orc['anchors'][b])


# ----------------------------------------------------------------------------------------------------
# standalone NMS entry points (multiclass_nms / batched_nms / nms shims) vs the oracle
# ----------------------------------------------------------------------------------------------------
def _rand_boxes(rng, n, span=300., wh=90.):
    xy = rng.rand(n, 2).astype(np.float32) * span - 20
    sz = rng.rand(n, 2).astype(np.float32) * wh + 1
    return np.concatenate([xy, xy + sz], 1)


@pytest.mark.parametrize('n,ncls,split_thr,max_num,agnostic,offset', [
    (0, 3, 10000, -1, False, 0), (1, 1, 10000, -1, False, 0), (700, 5, 10000, -1, False, 0), (700, 5, 100, -1, False, 0),
    (3000, 20, 10000, 300, False, 0), (3000, 20, 1000, 300, False, 0), (2500, 4, 10000, 100, True, 0),
    (900, 3, 10000, -1, False, 1), (12000, 80, 10000, 500, False, 0),
])
def test_batched_nms_matches_oracle(n, ncls, split_thr, max_num, agnostic, offset):
    import yolopp
    rng = np.random.RandomState(n + ncls)
    boxes = _rand_boxes(rng, n)
    scores = rng.rand(n).astype(np.float32)
    if n > 10:
        scores[rng.randint(0, n, 20)] = scores[0]  # exact ties
    idxs = rng.randint(0, ncls, n).astype(np.int64)
    cfg = dict(type='nms', iou_threshold=0.5, split_thr=split_thr)
    if max_num > 0:
        cfg['max_num'] = max_num
    if agnostic:
        cfg['class_agnostic'] = True
    if offset:
        cfg['offset'] = offset
    dets, keep = yolopp.batched_nms(torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda(),
                                    torch.from_numpy(idxs).cuda(), cfg)
    od, ok = oracle.batched_nms(boxes, scores, idxs, 0.5, offset=offset, split_thr=split_thr, class_agnostic=agnostic,
                                max_num=max_num)
    np.testing.assert_array_equal(keep.cpu().numpy(), ok)
    np.testing.assert_array_equal(_u32(dets.cpu().numpy()), _u32(od))


def test_nms_matches_oracle_and_torchvision():
    import yolopp
    rng = np.random.RandomState(5)
    boxes = _rand_boxes(rng, 2000)
    scores = (rng.rand(2000).astype(np.float32) - 0.3)  # negative scores too
    dets, inds = yolopp.nms(torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda(), 0.45)
    keep = oracle.nms(boxes, scores, 0.45)
    np.testing.assert_array_equal(inds.cpu().numpy(), keep)
    np.testing.assert_array_equal(_u32(dets.cpu().numpy()[:, 4]), _u32(scores[keep]))
    dets2, inds2 = yolopp.nms(torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda(), 0.45, max_num=50)
    np.testing.assert_array_equal(inds2.cpu().numpy(), keep[:50])


@pytest.mark.parametrize('n,C,per_class,factors,thr,max_num', [
    (500, 8, False, False, 0.05, 100), (500, 8, True, False, 0.05, 100), (1500, 80, False, True, 0.3, 200),
    (300, 4, False, False, 0.999, 100), (300, 20, False, False, 0.5, -1),
])
def test_multiclass_nms_matches_oracle(n, C, per_class, factors, thr, max_num):
    import yolopp
    rng = np.random.RandomState(n + C)
    boxes = _rand_boxes(rng, n * (C if per_class else 1)).reshape(n, -1)
    scores = rng.rand(n, C + 1).astype(np.float32)
    sf = rng.rand(n).astype(np.float32) if factors else None
    cfg = dict(type='nms', iou_threshold=0.5)
    out = yolopp.multiclass_nms(torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda(), thr, cfg, max_num,
                                score_factors=None if sf is None else torch.from_numpy(sf).cuda(), return_inds=True)
    od, ol, oi, oflat, ncand = oracle.multiclass_nms(boxes, scores, thr, 0.5, max_num=max_num, score_factors=sf)
    if ncand == 0:
        assert out[0].shape == (0, 4) and out[1].numel() == 0
        return
    dets, labels, inds = out
    np.testing.assert_array_equal(labels.cpu().numpy(), ol)
    np.testing.assert_array_equal(inds.cpu().numpy(), oi)
    np.testing.assert_array_equal(_u32(dets.cpu().numpy()), _u32(od))


def test_keep_all_overflow_is_reported_not_truncated():
    """No max_num and more than 4096 survivors: the ops raise instead of silently truncating."""
    import yolopp
    rng = np.random.RandomState(9)
    n = 6000
    boxes = np.concatenate([np.arange(n, dtype=np.float32)[:, None] * 10 + np.zeros((n, 2), np.float32),
                            np.arange(n, dtype=np.float32)[:, None] * 10 + 5 + np.zeros((n, 2), np.float32)], 1)
    scores = rng.rand(n).astype(np.float32)
    with pytest.raises(RuntimeError):
        yolopp.nms(torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda(), 0.5)
    dets, inds = yolopp.nms(torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda(), 0.5, max_num=4096)
    assert inds.numel() == 4096
    np.testing.assert_array_equal(inds.cpu().numpy(), oracle.nms(boxes, scores, 0.5)[:4096])


def test_get_results_host_and_bbox2result():
    """simple_test's tail (single_stage.py:102-111): one pinned D2H of the detections, then the per-class split."""
    import yolopp
    case = cases.CASES['csp608_sparse']
    p = cases.build_params(case)
    levels = yolopp.synth.synth_levels(p, case['seed'], case['dist'])
    head = yolopp.YOLOCSPHead(num_classes=80, test_cfg=cases.ref_cfg(case))
    metas = [dict(scale_factor=1.0) for _ in range(p.batch)]
    res = head.get_results_host(levels, metas)
    orc = oracle.get_bboxes(p, [x.cpu().numpy() for x in levels])
    for b, (dets, labels) in enumerate(res):
        np.testing.assert_array_equal(_u32(dets), _u32(orc['dets'][b]))
        np.testing.assert_array_equal(labels, orc['labels'][b])
        per_class = yolopp.bbox2result(dets, labels, 80)
        assert len(per_class) == 80 and sum(len(x) for x in per_class) == len(labels)
        for c in range(80):
            np.testing.assert_array_equal(per_class[c], orc['dets'][b][orc['labels'][b] == c])


def test_pipeline_matches_single_stream():
    """Batches in flight on several streams (yolopp.ops.Pipeline) give the results of one batch at a time."""
    import yolopp
    from yolopp.ops import Pipeline
    case = dict(cases.CASES['csp608_sparse'], batch=8)
    p = cases.build_params(case)
    inputs = [yolopp.synth.synth_levels(p, 100 + j, case['dist']) for j in range(5)]
    want = []
    for lv in inputs:
        out = yolopp.get_bboxes_raw(p, lv)
        torch.cuda.synchronize()
        want.append({k: v.cpu().numpy().copy() for k, v in out.items()})
    pipe = Pipeline(p, depth=3)
    for rnd in range(2):  # second round reuses the slots
        tickets = []
        for j in (0, 1, 2):
            tickets.append((j, pipe.submit(inputs[j])))
        for j, t in tickets:
            out = pipe.result(t)
            for k in ('count', 'num_candidates', 'labels', 'anchors', 'rows'):
                np.testing.assert_array_equal(out[k].cpu().numpy(), want[j][k], err_msg=f'{k} batch {j}')
            np.testing.assert_array_equal(_u32(out['dets'].cpu().numpy()), _u32(want[j]['dets']))
        # keep all three slots busy with other batches before the next round
        for j in (3, 4, 3):
            pipe.submit(inputs[j])
    torch.cuda.synchronize()


def test_select_exact_path_on_tied_objectness():
    """Objectness logits quantised to a handful of values (mass ties at the top-k cut): the raw-logit fast path
    of the select kernel must decline and the exact path must reproduce the canonical (conf desc, index asc) order."""
    import yolopp
    for name, q in (('csp_tiny', 2.0), ('csp608_sparse', 2.0), ('csp608_sparse', 64.0)):
        case = dict(cases.CASES[name])
        p = cases.build_params(case)
        levels = yolopp.synth.synth_levels(p, 5, case['dist'])
        A, NA = p.num_anchors, 5 + p.num_classes
        for x in levels:
            B, _, H, W = x.shape
            v = x.view(B, A, NA, H, W)
            v[:, :, 4] = torch.round(v[:, :, 4] * q) / q  # multiples of 1/q
        host = [x.cpu().numpy() for x in levels]
        orc = oracle.get_bboxes(p, host)
        _, _, res = run_cuda(case, p, levels)
        compare(p, res, orc, f'tied objectness {name} 1/{q}')
