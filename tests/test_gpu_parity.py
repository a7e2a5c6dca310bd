"""GPU parity tests (run on the B200 box: pytest -m gpu). Everything goes through the C ABI of
mmdet-yolov4_b200/csrc/libyolopp.so via the registered torch custom ops, and is compared with the CPU oracle
(oracle/oracle.c) on the same inputs. Bars: indices / labels / counts bit-exact; boxes and scores bit-exact too
(the oracle and the kernels share the canonical fp32 arithmetic), which is stricter than the 1e-5 relative
tolerance north_star states for floating point.
"""
import ctypes
import os

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
        levels = cases.device_levels(case, p)
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


def _oracle_full(case, p, levels):
    """The C oracle on the WHOLE batch (OpenMP over images: 64 images of 608^2 take ~0.4 s on the box)."""
    return oracle.get_bboxes(p, [x.cpu().numpy() for x in levels], cases.scale_factors(case))


def test_full_size_batch64_properties():
    """BASELINE config 2 at full size (608^2, batch 64): EVERY image equals the oracle bit for bit; plus the
    size-independent properties — images are independent (image b alone == image b inside the batch), counts within
    bounds, scores sorted, run-to-run identical."""
    import yolopp
    case = dict(cases.CASES['csp608_sparse'], batch=64)
    p, levels, res = run_cuda(case)
    assert int(res['status'][0]) == 0
    assert (res['count'] <= 300).all() and (res['count'] > 0).all()
    compare(p, res, _oracle_full(case, p, levels), 'csp608_sparse b64')
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


@pytest.mark.parametrize('name,batch', [('csp608_dense', 64), ('csp640_sparse', 128), ('v3_640_sparse', 128),
                                        ('csp1280_sparse', 128)])
def test_full_size_other_configs(name, batch):
    """BASELINE configs 3, 4 (both decode conventions) and 5 (one GPU's shard of the 1024-image batch) at full
    size: EVERY image equals the oracle; repeatable, sorted, within bounds, sampled images equal the same image
    processed alone."""
    import yolopp
    case = dict(cases.CASES[name], batch=batch)
    p, levels, res = run_cuda(case)
    assert int(res['status'][0]) == 0
    cap = case['max_per_img']
    assert (res['count'] <= cap).all() and (res['count'] > 0).all()
    compare(p, res, _oracle_full(case, p, levels), f'{name} b{batch}')
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


@pytest.mark.parametrize('batch,runs', [(64, 40), (128, 40)])
def test_run_to_run_reproducible(batch, runs):
    """The decode ring hands a stage from generic-proxy readers (ld.shared) to async-proxy writers (TMA); without
    the cross-proxy fences the first tiles of a launch differed in 3 % (batch 64) / 35 % (batch 128) of the runs.
    Every run must be bit-identical, and ALL images must equal the oracle — for the hybrid schedule (a batch alone)
    and for the static schedule that overlapped batches use (`batches_in_flight = 3`, what bench.py times)."""
    case = dict(cases.CASES['csp608_sparse'], batch=batch)
    p, levels, res = run_cuda(case)
    orc = _oracle_full(case, p, levels)
    compare(p, res, orc, f'hybrid schedule b{batch}')
    for r in range(runs):
        _, _, again = run_cuda(case, p, levels)
        for k in res:
            np.testing.assert_array_equal(again[k], res[k], err_msg=f'run {r}: run-to-run difference in {k}')
    p_static = type(p).from_buffer_copy(p)
    p_static.batches_in_flight = 3
    for r in range(5):
        _, _, other = run_cuda(case, p_static, levels)
        compare(p, other, orc, f'static schedule b{batch} run {r}')


def test_pipelined_batches_all_images_match_oracle():
    """What bench.py times: batches in flight on 3 streams (Pipeline, plan handles, static schedule), full size.
    Every image of every in-flight batch equals the oracle."""
    import yolopp
    from yolopp.ops import Pipeline
    case = dict(cases.CASES['csp608_sparse'], batch=64)
    p = cases.build_params(case)
    inputs = [yolopp.synth.synth_levels(p, case['seed'] + 7 * j, case['dist']) for j in range(2)]
    want = [_oracle_full(case, p, lv) for lv in inputs]
    pipe = Pipeline(p, depth=3)
    for rnd in range(3):
        tickets = [(i % 2, pipe.submit(inputs[i % 2])) for i in range(3)]
        for j, t in tickets:
            out = pipe.result(t)
            compare(p, {k: v.cpu().numpy() for k, v in out.items()}, want[j], f'pipelined round {rnd} input {j}')
        for i in range(3):  # keep the slots busy before the next round
            pipe.submit(inputs[(i + 1) % 2])
    torch.cuda.synchronize()


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


@pytest.mark.parametrize('n,ncls,split_thr,thr', [(800, 5, 10000, 0.4), (800, 5, 100, 0.4), (500, 1, 10000, 0.999), (3000, 20, 10000, 0.7)])
def test_batched_nms_score_threshold(n, ncls, split_thr, thr):
    """nms_cfg.score_threshold (mmcv NMSop.forward prefilter): only boxes with score > threshold enter the greedy
    pass; both regimes."""
    import yolopp
    rng = np.random.RandomState(n + ncls + split_thr)
    boxes = _rand_boxes(rng, n)
    scores = rng.rand(n).astype(np.float32)
    idxs = rng.randint(0, ncls, n).astype(np.int64)
    cfg = dict(type='nms', iou_threshold=0.5, split_thr=split_thr, score_threshold=thr)
    dets, keep = yolopp.batched_nms(torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda(),
                                    torch.from_numpy(idxs).cuda(), cfg)
    od, ok = oracle.batched_nms(boxes, scores, idxs, 0.5, split_thr=split_thr, score_threshold=thr)
    np.testing.assert_array_equal(keep.cpu().numpy(), ok)
    np.testing.assert_array_equal(_u32(dets.cpu().numpy()), _u32(od))
    d2, i2 = yolopp.nms(torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda(), 0.5, score_threshold=thr)
    _, ok2 = oracle.batched_nms(boxes, scores, None, 0.5, class_agnostic=True, score_threshold=thr)
    np.testing.assert_array_equal(i2.cpu().numpy(), ok2)


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


def test_keep_all_beyond_the_smem_kept_list():
    """No max_num and more than 4096 survivors (RPN-sized standalone NMS): the kept list lives in the workspace;
    nothing is truncated."""
    import yolopp
    rng = np.random.RandomState(9)
    n = 6000
    boxes = np.concatenate([np.arange(n, dtype=np.float32)[:, None] * 10 + np.zeros((n, 2), np.float32),
                            np.arange(n, dtype=np.float32)[:, None] * 10 + 5 + np.zeros((n, 2), np.float32)], 1)
    scores = rng.rand(n).astype(np.float32)
    dets, inds = yolopp.nms(torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda(), 0.5)
    want = oracle.nms(boxes, scores, 0.5)
    assert inds.numel() == 6000 == len(want)
    np.testing.assert_array_equal(inds.cpu().numpy(), want)
    dets, inds = yolopp.nms(torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda(), 0.5, max_num=4096)
    np.testing.assert_array_equal(inds.cpu().numpy(), want[:4096])
    # RPN-like: 12000 overlapping boxes, 5 levels as classes, keep all (rpn_head.py:247 then dets[:max_per_img])
    boxes = _rand_boxes(rng, 12000, span=800., wh=200.)
    scores = rng.rand(12000).astype(np.float32)
    idxs = rng.randint(0, 5, 12000).astype(np.int64)
    dets, keep = yolopp.batched_nms(torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda(),
                                    torch.from_numpy(idxs).cuda(), dict(type='nms', iou_threshold=0.7))
    od, ok = oracle.batched_nms(boxes, scores, idxs, 0.7)
    assert len(ok) > 4096
    np.testing.assert_array_equal(keep.cpu().numpy(), ok)
    np.testing.assert_array_equal(_u32(dets.cpu().numpy()), _u32(od))


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


# ----------------------------------------------------------------------------------------------------
# round 2: NHWC layout, stage taps, class-grouped output, plan handles, large nms_pre, the reference's pkl input, Mish
# ----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('name', ['csp608_sparse', 'csp608_dense', 'csp_odd', 'csp_tiny', 'csp320_nopre_sparse',
                                  'tencent_agnostic', 'v3_416_sparse', 'v3_tiny_nopre', 'csp416_rescale',
                                  'csp_saturated'])
def test_channels_last_inputs_match_oracle(name):
    """Channels-last head tensors (what a cuDNN NHWC convolution writes) are consumed IN PLACE by the row-driven
    decode; same bits as the NCHW path and the oracle."""
    import yolopp
    case = cases.CASES[name]
    p = cases.build_params(case)
    levels = yolopp.synth.synth_levels(p, case['seed'], case['dist'])
    cl = [x.contiguous(memory_format=torch.channels_last) for x in levels]
    if all(x.shape[2] * x.shape[3] > 1 for x in cl):
        assert not any(x.is_contiguous() for x in cl)
    sf = cases.scale_factors(case)
    out = yolopp.get_bboxes_raw(p, cl, torch.from_numpy(sf).cuda() if sf is not None else None)
    torch.cuda.synchronize()
    res = {k: v.cpu().numpy() for k, v in out.items()}
    orc = oracle.get_bboxes(p, [x.cpu().numpy() for x in levels], sf)
    compare(p, res, orc, f'{name} (NHWC)')


def test_channels_last_full_size():
    import yolopp
    case = dict(cases.CASES['csp608_sparse'], batch=64)
    p = cases.build_params(case)
    levels = yolopp.synth.synth_levels(p, case['seed'], case['dist'])
    cl = [x.contiguous(memory_format=torch.channels_last) for x in levels]
    out = yolopp.get_bboxes_raw(p, cl)
    torch.cuda.synchronize()
    compare(p, {k: v.cpu().numpy() for k, v in out.items()}, _oracle_full(case, p, levels), 'csp608_sparse b64 NHWC')


@pytest.mark.parametrize('name', ['csp608_sparse', 'csp416_rescale', 'csp_odd', 'csp_saturated'])
def test_stage_taps_match_reference_golden(name):
    """yolopp_topk_conf / yolopp_decode (SURVEY.md A.3 taps) == the reference's own intermediate tensors
    (tests/golden/taps_*.npz from make_golden.py --taps): topk_inds, the boxes entering multiclass_nms, and the
    candidate scores with the non-candidates marked."""
    import yolopp
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', f'taps_{name}.npz'))
    case = cases.CASES[name]
    p = cases.build_params(case)
    levels = yolopp.synth.synth_levels(p, case['seed'], case['dist'])
    sf = cases.scale_factors(case)
    sf_t = torch.from_numpy(sf).cuda() if sf is not None else None
    topk = yolopp.topk_conf(p, levels).cpu().numpy()
    np.testing.assert_array_equal(topk, g['topk_inds'])
    boxes, scores, inds = yolopp.decode(p, levels, sf_t)
    np.testing.assert_array_equal(inds.cpu().numpy(), g['topk_inds'])
    sb = scores.cpu().numpy().view(np.uint32)
    np.testing.assert_array_equal(sb, g['scores_bits'])
    # boxes of rows that own a candidate (rows without one are not needed downstream; CSP writes them all anyway)
    np.testing.assert_array_equal(_u32(boxes.cpu().numpy()), g['boxes_bits'])


@pytest.mark.parametrize('name', ['v3_416_sparse', 'v3_320_mid', 'csp320_nopre_sparse', 'tencent_agnostic', 'csp608_dense'])
def test_stage_taps_match_oracle(name):
    """The taps against the oracle's, for the configurations without a reference-generated tap fixture (YOLOv3
    convention with its conf_thr row filter, no top-k, class agnostic, dense)."""
    import yolopp
    case = cases.CASES[name]
    p = cases.build_params(case)
    levels = yolopp.synth.synth_levels(p, case['seed'], case['dist'])
    R = capi.describe(p).rows_per_image
    o_topk, o_boxes, o_scores = oracle.get_taps(p, [x.cpu().numpy() for x in levels], R, cases.scale_factors(case))
    np.testing.assert_array_equal(yolopp.topk_conf(p, levels).cpu().numpy(), o_topk)
    boxes, scores, inds = yolopp.decode(p, levels)
    np.testing.assert_array_equal(inds.cpu().numpy(), o_topk)
    s = scores.cpu().numpy()
    nan = np.isnan(o_scores)
    np.testing.assert_array_equal(np.isnan(s), nan)
    np.testing.assert_array_equal(_u32(s[~nan]), _u32(o_scores[~nan]))
    has = ~nan.all(axis=2)  # rows that own a candidate
    np.testing.assert_array_equal(_u32(boxes.cpu().numpy()[has]), _u32(o_boxes[has]))


@pytest.mark.parametrize('name', ['csp608_sparse', 'csp608_dense', 'v3_416_sparse', 'tencent_agnostic', 'csp_empty'])
def test_class_grouped_output_is_bbox2result(name):
    """yolopp_outputs.cls_dets / cls_offsets == bbox2result (mmdet/core/bbox/transforms.py:99-116) of the same
    detections: group c holds `bboxes[labels == c]` in score order."""
    import yolopp
    case = cases.CASES[name]
    p, levels, res = run_cuda(case)
    C = p.eff_classes
    for b in range(p.batch):
        n = int(res['count'][b])
        off = res['cls_offsets'][b]
        assert off[0] == 0 and off[C] == n and (np.diff(off) >= 0).all()
        want = yolopp.bbox2result(res['dets'][b, :n], res['labels'][b, :n], C)
        for c in range(C):
            np.testing.assert_array_equal(_u32(res['cls_dets'][b, off[c]:off[c + 1]]), _u32(want[c]))
    # through the head API: bbox_results of simple_test (single_stage.py:108-111)
    if case['mode'] == capi.MODE_V3:
        head = yolopp.YOLOV3Head(num_classes=case['num_classes'], test_cfg=cases.ref_cfg(case))
    else:
        head = yolopp.YOLOCSPHead(num_classes=case['num_classes'], test_cfg=cases.ref_cfg(case),
                                  class_agnostic=case.get('class_agnostic', False), featmap_strides=case['strides'],
                                  anchor_generator=dict(type='YOLOV4AnchorGenerator', base_sizes=case['base_sizes'],
                                                        strides=case['strides']))
    metas = [dict(scale_factor=1.0) for _ in range(p.batch)]
    per_img = head.get_bbox_results(levels, metas)
    flat = head.get_results_host(levels, metas)
    for b in range(p.batch):
        want = yolopp.bbox2result(flat[b][0], flat[b][1], C)
        assert len(per_img[b]) == C
        for c in range(C):
            np.testing.assert_array_equal(_u32(per_img[b][c]), _u32(want[c]))


def test_results_host_are_independent_copies():
    """Results of an earlier batch must not change when the next batch reuses the pinned staging block."""
    import yolopp
    case = cases.CASES['csp608_sparse']
    p = cases.build_params(case)
    head = yolopp.YOLOCSPHead(num_classes=80, test_cfg=cases.ref_cfg(case))
    metas = [dict(scale_factor=1.0) for _ in range(p.batch)]
    a = head.get_results_host(yolopp.synth.synth_levels(p, 1, case['dist']), metas)
    keep = [(d.copy(), l.copy()) for d, l in a]
    head.get_results_host(yolopp.synth.synth_levels(p, 2, case['dist']), metas)
    for (d, l), (d0, l0) in zip(a, keep):
        np.testing.assert_array_equal(d, d0)
        np.testing.assert_array_equal(l, l0)


def test_plan_handle_reuse_and_validation():
    """Session: one yolopp_plan per set of buffers, reused across calls; alternating input sets; bad inputs are
    rejected before any pointer reaches the C ABI."""
    import yolopp
    from yolopp.ops import Session
    case = dict(cases.CASES['csp608_sparse'], batch=4)
    p = cases.build_params(case)
    sess = Session(p)
    ins = [yolopp.synth.synth_levels(p, 50 + j, case['dist']) for j in range(3)]
    want = [oracle.get_bboxes(p, [x.cpu().numpy() for x in lv]) for lv in ins]
    for rnd in range(3):
        for j, lv in enumerate(ins):
            out = sess.run(lv)
            torch.cuda.synchronize()
            compare(p, {k: v.cpu().numpy() for k, v in out.items()}, want[j], f'plan reuse round {rnd} input {j}')
    assert len(sess._plans) == 3
    with pytest.raises(TypeError):
        sess.run([x.half() for x in ins[0]])
    with pytest.raises(AssertionError):
        sess.run([x[:, :, :-1] for x in ins[0]])
    with pytest.raises(ValueError):  # neither NCHW- nor channels-last-contiguous: a Session never copies silently
        sess.run([x.permute(0, 1, 3, 2).contiguous().permute(0, 1, 3, 2) for x in ins[0]])
    # channels-last inputs get their own plan (row-driven decode), same results
    out = sess.run([x.contiguous(memory_format=torch.channels_last) for x in ins[1]])
    torch.cuda.synchronize()
    compare(p, {k: v.cpu().numpy() for k, v in out.items()}, want[1], 'plan, channels-last')


def test_host_pipeline_matches_oracle():
    """HostPipeline (bench.py's e2e path): pinned host tensors in, host results out, two slots in flight."""
    import yolopp
    from yolopp.ops import HostPipeline
    case = dict(cases.CASES['csp608_sparse'], batch=8)
    p = cases.build_params(case)
    hosts, want = [], []
    for j in range(3):
        lv = yolopp.synth.synth_levels(p, 70 + j, case['dist'])
        hosts.append([x.cpu().pin_memory() for x in lv])
        want.append(oracle.get_bboxes(p, [x.numpy() for x in hosts[-1]]))
    hp = HostPipeline(p, depth=2)
    prev = None
    got = {}
    for i in range(6):
        t = hp.submit(hosts[i % 3])
        if prev is not None:
            got[prev[0]] = hp.result(prev[1])
        prev = (i, t)
    got[prev[0]] = hp.result(prev[1])
    for i in range(6):
        w = want[i % 3]
        for b in range(p.batch):
            np.testing.assert_array_equal(_u32(got[i][b][0]), _u32(w['dets'][b]))
            np.testing.assert_array_equal(got[i][b][1], w['labels'][b])


@pytest.mark.parametrize('name,nms_pre', [('csp608_sparse', 5000), ('csp320_nopre_dense', 5000), ('v3_416_dense', 4500)])
def test_nms_pre_beyond_the_smem_sort_capacity(name, nms_pre):
    """nms_pre in (4096, N): the reference takes any k (yolocsp_head.py:348-355); the select kernel produces the
    sorted top-k in chunks."""
    case = dict(cases.CASES[name], nms_pre=nms_pre, batch=1, score_thr=max(cases.CASES[name]['score_thr'], 0.3))
    p, levels, res = run_cuda(case)
    orc = oracle.get_bboxes(p, [x.cpu().numpy() for x in levels], cases.scale_factors(case))
    compare(p, res, orc, f'{name} nms_pre={nms_pre}')


def test_reference_pkl_input():
    """The reference's own deterministic input (tests/test_onnx/data/yolov3_head_get_bboxes.pkl; head of
    tests/test_onnx/test_head.py:103-129), stored with the reference's outputs in tests/golden/v3_onnx_pkl.npz."""
    import yolopp
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'v3_onnx_pkl.npz'))
    p = cases.build_params(cases.PKL_CASE)
    levels = [torch.from_numpy(g[f'level{l}']).cuda() for l in range(3)]
    out = yolopp.get_bboxes_raw(p, levels)
    torch.cuda.synchronize()
    n = int(g['canon_count'][0])
    assert int(out['count'][0]) == n and int(out['num_candidates'][0]) == int(g['canon_ncand'][0])
    np.testing.assert_array_equal(_u32(out['dets'][0, :n].cpu().numpy()), g['canon_dets_bits'][0, :n])
    np.testing.assert_array_equal(out['labels'][0, :n].cpu().numpy(), g['canon_labels'][0, :n])
    assert cases.asis_strict_rel_err(g['asis_dets'][0, :n], out['dets'][0, :n].cpu().numpy()) <= 1e-5
    # through the head mirror, exactly as the reference's test builds it
    head = yolopp.YOLOV3Head(num_classes=4, in_channels=[1, 1, 1], out_channels=[16, 8, 4],
                             test_cfg=dict(deploy_nms_pre=0, min_bbox_size=0, score_thr=0.05, conf_thr=0.005,
                                           nms=dict(type='nms', iou_threshold=0.45), max_per_img=100))
    dets, labels = head.get_bboxes(levels, [dict(scale_factor=1)])[0]
    np.testing.assert_array_equal(_u32(dets.cpu().numpy()), g['canon_dets_bits'][0, :n])
    np.testing.assert_array_equal(labels.cpu().numpy(), g['canon_labels'][0, :n])


def _mish_inputs(n=1 << 20):
    rng = np.random.RandomState(11)
    x = np.concatenate([np.linspace(-30, 30, n // 2), rng.randn(n // 2 - 7) * 3,
                        [0.0, -0.0, 19.999, 20.0, 20.001, -87.0, 60.0]]).astype(np.float32)
    return x, rng.randn(x.size).astype(np.float32)


def test_mish_forward_backward_fp32():
    """Mish vs the oracle (mish.h math, pinned bit-exactly to the reference's own header). Tolerance (fp32): the
    kernel uses exp + one division instead of log1p / tanh — |err| <= 2e-6 * max(|y|, 1) forward,
    <= 4e-6 * max(|dy|, 1) backward (a few ulp; the reference's own float chain on CUDA differs from its CPU double
    chain by as much)."""
    import yolopp
    x, g = _mish_inputs()
    xt, gt = torch.from_numpy(x).cuda(), torch.from_numpy(g).cuda()
    y = yolopp.mish_forward(xt).cpu().numpy()
    ref = oracle.mish_forward(x)
    assert (np.abs(y - ref) <= 2e-6 * np.maximum(np.abs(ref), 1.0)).all(), np.abs(y - ref).max()
    dx = yolopp.mish_backward(gt, xt).cpu().numpy()
    refb = oracle.mish_backward(g, x)
    assert (np.abs(dx - refb) <= 4e-6 * np.maximum(np.abs(g), 1.0)).all(), np.abs(dx - refb).max()
    # odd sizes: scalar tail, tiny tensors, empty
    for n in (0, 1, 3, 5, 1027):
        yy = yolopp.mish_forward(xt[:n].clone()).cpu().numpy()
        assert (np.abs(yy - ref[:n]) <= 2e-6 * np.maximum(np.abs(ref[:n]), 1.0)).all()
        dd = yolopp.mish_backward(gt[:n].clone(), xt[:n].clone()).cpu().numpy()
        assert (np.abs(dd - refb[:n]) <= 4e-6 * np.maximum(np.abs(g[:n]), 1.0)).all()


@pytest.mark.parametrize('dtype,tol', [(torch.float16, 1e-3), (torch.bfloat16, 8e-3)])
def test_mish_half_precisions(dtype, tol):
    """Half / bfloat16 compute in fp32 and round once (mish.h:33-50): result == round(oracle(fp32(x))) within one
    rounding of the output type."""
    import yolopp
    x, g = _mish_inputs(1 << 16)
    xt, gt = torch.from_numpy(x).cuda().to(dtype), torch.from_numpy(g).cuda().to(dtype)
    xf, gf = xt.float().cpu().numpy(), gt.float().cpu().numpy()
    y = yolopp.mish_forward(xt)
    assert y.dtype == dtype
    ref = oracle.mish_forward(xf)
    assert (np.abs(y.float().cpu().numpy() - ref) <= tol * np.maximum(np.abs(ref), 1.0)).all()
    dx = yolopp.mish_backward(gt, xt)
    refb = oracle.mish_backward(gf, xf)
    assert (np.abs(dx.float().cpu().numpy() - refb) <= tol * np.maximum(np.abs(gf), 1.0)).all()


def test_mish_module_autograd():
    """yolopp.Mish mirrors the reference's Mish / MishCudaFunction (mish.py:18-48): non-contiguous inputs, autograd."""
    import yolopp
    m = yolopp.Mish()
    x = torch.randn(4, 8, 16, 16, device='cuda').permute(0, 2, 3, 1).requires_grad_(True)
    y = m(x)
    (y * 2).sum().backward()
    xr = x.detach().cpu().numpy()
    np.testing.assert_allclose(y.detach().cpu().numpy(), oracle.mish_forward(xr).reshape(xr.shape), rtol=0, atol=4e-6)
    np.testing.assert_allclose(x.grad.cpu().numpy(), oracle.mish_backward(np.full(xr.shape, 2, np.float32), xr).reshape(xr.shape),
                               rtol=0, atol=1e-5)
    with pytest.raises((NotImplementedError, RuntimeError)):
        m(torch.randn(4))  # CPU tensor: no fallback


def test_gather_detections_nccl_single_rank():
    """The device-side replacement of collect_results_gpu (mmdet/apis/test.py:160-190): fixed-size all_gather of the
    detections over NCCL instead of pickles. One rank here (the 2-rank exchange runs under gloo in the CPU suite and
    under NCCL in bench.py --gpus 2)."""
    import torch.distributed as dist
    import yolopp
    from yolopp.shard import gather_detections
    if not dist.is_initialized():
        import socket
        with socket.socket() as s:
            s.bind(('127.0.0.1', 0))
            port = s.getsockname()[1]
        dist.init_process_group('nccl', init_method=f'tcp://127.0.0.1:{port}', rank=0, world_size=1,
                                device_id=torch.device('cuda', 0))
    try:
        case = cases.CASES['csp608_sparse']
        p = cases.build_params(case)
        levels = yolopp.synth.synth_levels(p, case['seed'], case['dist'])
        head = yolopp.YOLOCSPHead(num_classes=80, test_cfg=cases.ref_cfg(case))
        local = head.get_results_host(levels, [dict(scale_factor=1.0)] * p.batch)
        res = gather_detections(local, p.batch)
        for b in range(p.batch):
            np.testing.assert_array_equal(_u32(res[b][0].numpy()), _u32(local[b][0]))
            np.testing.assert_array_equal(res[b][1].numpy(), local[b][1])
    finally:
        dist.destroy_process_group()
