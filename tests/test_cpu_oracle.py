"""CPU tests (no GPU): the oracle against the golden vectors produced by the reference's own source, against the
reference's own known-answer tests, and (when /root/reference is present) against the live reference."""
import os

import numpy as np
import pytest

import cases
from oracle import oracle, refexec
from yolopp import _capi as capi

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _u32(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.mark.parametrize('name', cases.GOLDEN_CASES)
def test_oracle_matches_reference_golden(name):
    """Bit-exact: boxes, scores, labels, kept anchors and their order == the reference's own code run with the
    canonical transcendental (tests/golden/make_golden.py)."""
    g = np.load(os.path.join(GOLDEN_DIR, f'{name}.npz'))
    case = cases.CASES[name]
    p = cases.build_params(case)
    out = oracle.get_bboxes(p, cases.host_levels(case, p), cases.scale_factors(case))
    np.testing.assert_array_equal(out['count'], g['canon_count'])
    np.testing.assert_array_equal(out['num_candidates'], g['canon_ncand'])
    for b in range(p.batch):
        n = int(g['canon_count'][b])
        np.testing.assert_array_equal(_u32(out['dets'][b]), g['canon_dets_bits'][b, :n])
        np.testing.assert_array_equal(out['labels'][b], g['canon_labels'][b, :n])
        if bool(g['has_anchors']):
            np.testing.assert_array_equal(out['anchors'][b], g['canon_anchors'][b, :n])


@pytest.mark.parametrize('name', cases.GOLDEN_CASES)
def test_oracle_vs_reference_as_is(name):
    """Against the reference exactly as it runs (torch's own CPU sigmoid/exp): kept (anchor, class) sets and
    order identical, boxes and scores within 1e-5 relative — north_star's tolerance, written here."""
    g = np.load(os.path.join(GOLDEN_DIR, f'{name}.npz'))
    assert bool(g['asis_index_equal']), 'recorded at generation time: index sets were equal for every golden case'
    case = cases.CASES[name]
    p = cases.build_params(case)
    out = oracle.get_bboxes(p, cases.host_levels(case, p), cases.scale_factors(case))
    np.testing.assert_array_equal(out['count'], g['asis_count'])
    for b in range(p.batch):
        n = int(g['asis_count'][b])
        np.testing.assert_array_equal(out['labels'][b], g['asis_labels'][b, :n])
        if bool(g['has_anchors']):
            np.testing.assert_array_equal(out['anchors'][b], g['asis_anchors'][b, :n])
        if n:
            # two readings of "1e-5 relative", both reported: per coordinate (north_star's literal text) and per
            # coordinate measured against the magnitude of its box (cancellation-aware, see cases.asis_rel_err)
            strict = cases.asis_strict_rel_err(g['asis_dets'][b, :n], out['dets'][b])
            rel = cases.asis_rel_err(g['asis_dets'][b, :n], out['dets'][b])
            print(f'{name} img {b}: as-is deviation strict per-coordinate {strict:.3e}, box-relative {rel:.3e}')
            assert rel <= 1e-5, rel
            assert strict <= cases.STRICT_OUTLIERS.get(name, 1e-5), (name, strict)


def test_reference_kat_yolo_bbox_coder():
    """tests/test_utils/test_coder.py:8-23 of the reference (YOLOBBoxCoder.decode known answer)."""
    bboxes = np.array([[-42., -29., 74., 61.], [-10., -29., 106., 61.], [22., -29., 138., 61.], [54., -29., 170., 61.]],
                      np.float32)
    pred = np.array([[0.4709, 0.6152, 0.1690, -0.4056], [0.5399, 0.6653, 0.1162, -0.4162],
                     [0.4654, 0.6618, 0.1548, -0.4301], [0.4786, 0.6197, 0.1896, -0.4479]], np.float32)
    expected = np.array([[-53.6102, -10.3096, 83.7478, 49.6824], [-15.8700, -8.3901, 114.4236, 50.9693],
                         [11.1822, -8.0924, 146.6034, 50.4476], [41.2068, -8.9232, 181.4236, 48.5840]], np.float32)
    got = oracle.coder_decode(capi.MODE_V3, bboxes, pred, 32)
    assert np.allclose(got, expected, rtol=1e-5, atol=1e-8)  # torch.allclose defaults


def test_reference_kat_yolo_anchor_generator():
    """tests/test_utils/test_anchor.py:148-188 of the reference (YOLO base anchors known answer)."""
    base = capi.yolo_base_anchors([[(116, 90), (156, 198), (373, 326)], [(30, 61), (62, 45), (59, 119)],
                                   [(10, 13), (16, 30), (33, 23)]], [32, 16, 8])
    expected = [
        np.array([[-42.0, -29.0, 74.0, 61.0], [-62.0, -83.0, 94.0, 115.0], [-170.5, -147.0, 202.5, 179.0]], np.float32),
        np.array([[-7.0, -22.5, 23.0, 38.5], [-23.0, -14.5, 39.0, 30.5], [-21.5, -51.5, 37.5, 67.5]], np.float32),
        np.array([[-1.0, -2.5, 9.0, 10.5], [-4.0, -11.0, 12.0, 19.0], [-12.5, -7.5, 20.5, 15.5]], np.float32),
    ]
    assert [b.shape[0] for b in base] == [3, 3, 3]
    for b, e in zip(base, expected):
        np.testing.assert_array_equal(b, e)
    for (h, w), b, s in zip([(14, 18), (28, 36), (56, 72)], base, [32, 16, 8]):
        a = oracle.grid_anchors(b, h, w, s, s)
        assert a.shape == (h * w * 3, 4)
        np.testing.assert_array_equal(a[:3], b)                      # cell (0,0)
        np.testing.assert_array_equal(a[3:6], b + np.float32(s) * np.array([1, 0, 1, 0], np.float32))  # cell x=1


def test_canonical_exp_accuracy():
    """<= 1.02 ulp against double precision, monotone (DESIGN.md "Canonical arithmetic")."""
    x = np.linspace(-30, 30, 2_000_001).astype(np.float32)
    y = oracle.expf(x).astype(np.float64)
    ref = np.exp(x.astype(np.float64))
    ulp = np.spacing(ref.astype(np.float32)).astype(np.float64)
    assert (np.abs(y - ref) / ulp).max() <= 1.02
    assert (np.diff(y) >= 0).all()
    s = oracle.sigmoid(x)
    assert (np.diff(s) >= 0).all() and s.min() >= 0 and s.max() <= 1
    # special values
    sp = oracle.expf(np.array([np.inf, -np.inf, 0.0, 89.0, -104.0], np.float32))
    assert sp[0] == np.inf and sp[1] == 0 and sp[2] == 1 and sp[3] == np.inf and sp[4] == 0
    assert np.isnan(oracle.expf(np.array([np.nan], np.float32))[0])


def test_oracle_nms_matches_torchvision():
    """The restated mmcv nms_cpu against an independent implementation of the same greedy algorithm."""
    torch = pytest.importorskip('torch')
    tv = pytest.importorskip('torchvision')
    rng = np.random.RandomState(3)
    for n, thr in ((1, 0.5), (50, 0.3), (700, 0.5), (2000, 0.65)):
        xy = rng.rand(n, 2).astype(np.float32) * 200
        wh = rng.rand(n, 2).astype(np.float32) * 80 + 1
        boxes = np.concatenate([xy, xy + wh], 1)
        scores = rng.permutation(n).astype(np.float32)  # no ties
        keep = oracle.nms(boxes, scores, thr)
        ref = tv.ops.nms(torch.from_numpy(boxes), torch.from_numpy(scores), thr).numpy()
        np.testing.assert_array_equal(keep, ref)


def test_oracle_batched_nms_regimes():
    """split_thr: below it one problem over offset boxes, at/above it per class — same result when the classes
    are far apart, and the empty / single-box edge cases."""
    rng = np.random.RandomState(4)
    n = 600
    xy = rng.rand(n, 2).astype(np.float32) * 100
    wh = rng.rand(n, 2).astype(np.float32) * 40 + 1
    boxes = np.concatenate([xy, xy + wh], 1)
    scores = rng.rand(n).astype(np.float32)
    idxs = rng.randint(0, 7, n)
    d1, k1 = oracle.batched_nms(boxes, scores, idxs, 0.5, split_thr=10000)
    d2, k2 = oracle.batched_nms(boxes, scores, idxs, 0.5, split_thr=10)
    np.testing.assert_array_equal(k1, k2)  # boxes are >= 0 so class offset ranges never overlap
    np.testing.assert_array_equal(d1, d2)
    # independent implementation of the classes-independent regime (the >= split_thr branch, written by another hand):
    # torchvision.ops.batched_nms, both of its strategies (coordinate trick below 4000 boxes, per-class loop above)
    torch = pytest.importorskip('torch')
    tv = pytest.importorskip('torchvision')
    for nn, ncls, thr in ((600, 7, 0.5), (600, 7, 0.2), (5000, 40, 0.45)):
        xy2 = rng.rand(nn, 2).astype(np.float32) * 100
        wh2 = rng.rand(nn, 2).astype(np.float32) * 40 + 1
        b2 = np.concatenate([xy2, xy2 + wh2], 1)
        s2 = rng.permutation(nn).astype(np.float32) / nn  # no ties
        i2 = rng.randint(0, ncls, nn)
        _, kk = oracle.batched_nms(b2, s2, i2, thr, split_thr=10)
        ref = tv.ops.batched_nms(torch.from_numpy(b2), torch.from_numpy(s2), torch.from_numpy(i2), thr).numpy()
        np.testing.assert_array_equal(kk, ref)
    d0, k0 = oracle.batched_nms(np.zeros((0, 4), np.float32), np.zeros(0, np.float32), np.zeros(0, np.int64), 0.5)
    assert d0.shape == (0, 5) and k0.shape == (0, )
    d3, k3 = oracle.batched_nms(boxes[:1], scores[:1], idxs[:1], 0.5)
    assert k3.tolist() == [0]
    # class agnostic = plain nms
    da, ka = oracle.batched_nms(boxes, scores, idxs, 0.5, class_agnostic=True)
    np.testing.assert_array_equal(ka, oracle.nms(boxes, scores, 0.5))


def test_synth_generator_is_stable():
    """The synthetic generator is part of the parity contract (inputs are regenerated from seeds)."""
    x = oracle.synth_level(1, 1, 6, 4, np.array([0, 0, 0, 0, -5, -4.9], np.float32),
                           np.array([1, 1, 1, 1, 2, 1.5], np.float32), 12345)
    assert x.shape == (1, 6, 4)
    np.testing.assert_array_equal(
        x.reshape(-1)[:6].view(np.uint32),
        oracle.synth_level(1, 1, 6, 4, np.array([0, 0, 0, 0, -5, -4.9], np.float32),
                           np.array([1, 1, 1, 1, 2, 1.5], np.float32), 12345).reshape(-1)[:6].view(np.uint32))
    big = oracle.synth_level(2, 3, 85, 361, np.array([0] * 4 + [-5] + [-4.9] * 80, np.float32),
                             np.array([1] * 4 + [2] + [1.5] * 80, np.float32), 99).reshape(2, 3, 85, 361)
    assert abs(big[:, :, :4].mean()) < 0.05 and abs(big[:, :, :4].std() - 1) < 0.05
    assert abs(big[:, :, 4].mean() + 5) < 0.1 and abs(big[:, :, 5:].std() - 1.5) < 0.05


@pytest.mark.skipif(not refexec.available(), reason='reference tree not present (GPU box)')
@pytest.mark.parametrize('name', ['csp_tiny', 'csp_pre_nm1', 'v3_tiny_nopre', 'csp_keep_all', 'csp_force_split'])
def test_oracle_against_live_reference(name):
    """Where /root/reference exists, cases WITHOUT a committed golden are checked against the reference's own
    source on the spot (same harness as tests/golden/make_golden.py)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location('make_golden', os.path.join(GOLDEN_DIR, 'make_golden.py'))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    case = cases.CASES[name]
    p = cases.build_params(case)
    levels = cases.host_levels(case, p)
    ref = mg.run_reference(case, levels, True)
    out = oracle.get_bboxes(p, levels, cases.scale_factors(case))
    for b in range(p.batch):
        d = ref[b]['dets']
        assert out['count'][b] == ref[b]['labels'].shape[0]
        if out['count'][b]:
            np.testing.assert_array_equal(_u32(out['dets'][b]), _u32(d[:, :5]))
            np.testing.assert_array_equal(out['labels'][b], ref[b]['labels'])


# ----------------------------------------------------------------------------------------------------
# intermediate taps, the reference's own pkl input, Mish
# ----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('name', ['csp608_sparse', 'csp416_rescale', 'csp_odd', 'csp_saturated'])
def test_oracle_taps_match_reference_golden(name):
    """topk_inds (yolocsp_head.py:350-355), the boxes entering multiclass_nms and the candidate scores
    (bbox_nms.py:42-67) == what the reference's own source computed (tests/golden/make_golden.py --taps)."""
    g = np.load(os.path.join(GOLDEN_DIR, f'taps_{name}.npz'))
    case = cases.CASES[name]
    p = cases.build_params(case)
    R = g['topk_inds'].shape[1]
    assert R == capi.describe(p).rows_per_image
    topk, boxes, scores = oracle.get_taps(p, cases.host_levels(case, p), R, cases.scale_factors(case))
    np.testing.assert_array_equal(topk, g['topk_inds'])
    np.testing.assert_array_equal(_u32(boxes), g['boxes_bits'])
    bits = np.where(np.isnan(scores), np.uint32(0xFFFFFFFF), _u32(scores))
    np.testing.assert_array_equal(bits, g['scores_bits'])


def test_oracle_on_reference_pkl_input():
    """The reference's own deterministic input (tests/test_onnx/data/yolov3_head_get_bboxes.pkl, head of
    tests/test_onnx/test_head.py:103-129): oracle == the reference's YOLOV3Head.get_bboxes on it."""
    g = np.load(os.path.join(GOLDEN_DIR, 'v3_onnx_pkl.npz'))
    p = cases.build_params(cases.PKL_CASE)
    out = oracle.get_bboxes(p, [g['level0'], g['level1'], g['level2']])
    n = int(g['canon_count'][0])
    assert out['count'][0] == n and out['num_candidates'][0] == g['canon_ncand'][0]
    np.testing.assert_array_equal(_u32(out['dets'][0]), g['canon_dets_bits'][0, :n])
    np.testing.assert_array_equal(out['labels'][0], g['canon_labels'][0, :n])
    np.testing.assert_array_equal(out['labels'][0], g['asis_labels'][0, :n])
    assert cases.asis_strict_rel_err(g['asis_dets'][0, :n], out['dets'][0]) <= 1e-5


def _mish_inputs():
    rng = np.random.RandomState(7)
    x = np.concatenate([np.linspace(-30, 30, 100001), np.linspace(-100, 100, 20001), rng.randn(50000) * 3,
                        [0.0, -0.0, 19.999, 20.0, 20.001, -87.0, 88.0]]).astype(np.float32)
    return x, rng.randn(x.size).astype(np.float32)


def test_oracle_mish_matches_reference_build():
    """oracle_mish_fwd/bwd == the reference's own mish.h compiled from /root/reference (oracle/_ref/libmish_ref.so),
    bit for bit. Skipped where the reference build is absent."""
    import ctypes
    path = os.path.join(os.path.dirname(GOLDEN_DIR), '..', 'oracle', '_ref', 'libmish_ref.so')
    if not os.path.isfile(path):
        pytest.skip('oracle/_ref/libmish_ref.so not built (needs /root/reference)')
    L = ctypes.CDLL(os.path.abspath(path))
    x, g = _mish_inputs()
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    o = np.empty_like(x)
    L.ref_mish_fwd(vp(x), vp(o), ctypes.c_longlong(x.size))
    np.testing.assert_array_equal(_u32(o), _u32(oracle.mish_forward(x)))
    L.ref_mish_bwd(vp(g), vp(x), vp(o), ctypes.c_longlong(x.size))
    np.testing.assert_array_equal(_u32(o), _u32(oracle.mish_backward(g, x)))


def test_oracle_mish_math():
    """Closed form used by the CUDA kernel (one exp + one division) against the oracle: x*n/(n+2), n = e(e+2), and
    the backward 4xe(1+e)/(n+2)^2 + n/(n+2); gradient also against a central difference of the forward."""
    x, g = _mish_inputs()
    xd = x.astype(np.float64)
    e = np.exp(np.minimum(xd, 20.0))
    n = e * (e + 2)
    fwd = np.where(xd >= 20, xd * np.tanh(xd), xd * n / (n + 2))
    ref = oracle.mish_forward(x).astype(np.float64)
    assert np.abs(fwd - ref).max() <= 1e-6 * np.maximum(np.abs(ref), 1.0).max()
    assert (np.abs(fwd - ref) / np.maximum(np.abs(ref), 1e-3)).max() <= 2e-6
    grad = np.where(xd >= 20, 1.0, 4 * xd * e * (1 + e) / (n + 2) ** 2 + n / (n + 2))
    refb = oracle.mish_backward(np.ones_like(x), x).astype(np.float64)
    assert np.abs(grad - refb).max() <= 2e-6
    h = 1e-3
    num = (oracle.mish_forward((xd + h).astype(np.float32)).astype(np.float64) -
           oracle.mish_forward((xd - h).astype(np.float32)).astype(np.float64)) / (2 * h)
    sel = np.abs(xd) < 15
    assert np.abs(num[sel] - refb[sel]).max() < 5e-3
