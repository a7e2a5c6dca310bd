"""CPU tests (no GPU): the C-ABI library loads and exports every symbol include/yolopp.h declares, the host-side
planning calls work without a device, the reference-shaped host API validates like the reference, and the
multi-GPU sharding helpers work under gloo with world_size 2."""
import ctypes
import os
import re
import socket
import subprocess
import sys

import numpy as np
import pytest

import cases
from yolopp import _capi as capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, 'include', 'yolopp.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    declared = set(re.findall(r'\b(yolopp_[a-z_0-9]+)\s*\(', hdr))
    assert declared == set(capi.EXPORTED_SYMBOLS), declared ^ set(capi.EXPORTED_SYMBOLS)
    lib = capi.load_library()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.yolopp_abi_version() == capi.ABI_VERSION
    assert b'ok' == lib.yolopp_strerror(0)
    assert ctypes.sizeof(capi.YoloppParams) == 4 * (7 + 5 * 8 + 8 * 8 * 4 + 11 + 7)


def test_header_struct_layouts_match_the_ctypes_mirrors(tmp_path):
    """include/yolopp.h compiled as plain C (the header is the boundary a cgo / JNI / ctypes binding reads): sizes and
    a few member offsets of the three structs must equal the Python mirrors the tests and the shim go through."""
    import subprocess
    src = tmp_path / 'layout.c'
    src.write_text('#include <stddef.h>\n#include <stdio.h>\n#include "yolopp.h"\nint main(void) {\n'
                   '  printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(yolopp_params), sizeof(yolopp_outputs), sizeof(yolopp_plan_info),\n'
                   '         offsetof(yolopp_params, base_anchors), offsetof(yolopp_outputs, cls_offsets),\n'
                   '         offsetof(yolopp_plan_info, workspace_bytes), offsetof(yolopp_plan_info, decode_tile_positions));\n'
                   '  return YOLOPP_ABI_VERSION == ' + str(capi.ABI_VERSION) + ' ? 0 : 1;\n}\n')
    exe = tmp_path / 'layout'
    subprocess.run(['gcc', '-std=c99', '-Wall', '-Werror', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    got = [int(x) for x in out]
    want = [ctypes.sizeof(capi.YoloppParams), ctypes.sizeof(capi.YoloppOutputs), ctypes.sizeof(capi.YoloppPlanInfo),
            capi.YoloppParams.base_anchors.offset, capi.YoloppOutputs.cls_offsets.offset,
            capi.YoloppPlanInfo.workspace_bytes.offset, capi.YoloppPlanInfo.decode_tile_positions.offset]
    assert got == want, (got, want)


def test_planning_calls_need_no_device():
    lib = capi.load_library()
    p = cases.build_params(dict(cases.CASES['csp608_sparse'], batch=64))
    ws = lib.yolopp_workspace_bytes(ctypes.byref(p))
    assert 20e6 < ws < 200e6
    info = capi.describe(p)
    assert info.anchors_per_image == 22743 and info.rows_per_image == 1000 and info.num_attrib == 85
    assert info.tma_level_mask == 0b111          # 76^2 / 38^2: TMA tiles; 19^2 (plane stride not 16-byte aligned):
                                                 # gather tiles of the same persistent kernel
    assert info.tma_bytes_per_image + info.ldg_bytes_per_image == 7732620  # 4*3*85*7581
    assert info.kernel_launches == 3             # select, persistent decode (TMA + gather tiles), per-image NMS
    assert info.ldg_blocks == 0 and info.ldg_bytes_per_image == 0
    # no top-k -> no select kernel, everything through the generic decode kernel
    p2 = cases.build_params(cases.CASES['csp320_nopre_dense'])
    i2 = capi.describe(p2)
    assert i2.tma_level_mask == 0 and i2.kernel_launches == 2 and i2.rows_per_image == i2.anchors_per_image
    assert i2.dense_tiles > 0 and i2.ldg_blocks == 0   # dense admission: thread-per-position kernel
    # V3: one top-k segment per level
    p3 = cases.build_params(cases.CASES['v3_416_sparse'])
    assert capi.describe(p3).rows_per_image == 507 + 1000 + 1000   # 13^2*3 < nms_pre: that level keeps all rows


@pytest.mark.parametrize('mutate', [
    lambda p: setattr(p, 'abi_version', 99), lambda p: setattr(p, 'mode', 5), lambda p: setattr(p, 'batch', 0),
    lambda p: setattr(p, 'num_levels', 9), lambda p: setattr(p, 'num_anchors', 0), lambda p: setattr(p, 'num_classes', 5000),
    lambda p: setattr(p, 'nms_offset', 2), lambda p: setattr(p, 'layout', 2), lambda p: setattr(p, 'max_per_img', 1 << 21),
    lambda p: (setattr(p, 'max_per_img', -1), setattr(p, 'out_capacity', 0)),
])
def test_invalid_params_are_rejected(mutate):
    lib = capi.load_library()
    p = cases.build_params(cases.CASES['csp608_sparse'])
    mutate(p)
    assert lib.yolopp_workspace_bytes(ctypes.byref(p)) == 0
    assert lib.yolopp_describe(ctypes.byref(p), ctypes.byref(capi.YoloppPlanInfo())) == capi.E_INVALID


def test_get_bboxes_rejects_bad_calls_without_touching_the_device():
    lib = capi.load_library()
    p = cases.build_params(cases.CASES['csp_tiny'])
    out = capi.YoloppOutputs()
    ptrs = (ctypes.c_void_p * 3)()
    assert lib.yolopp_get_bboxes(ctypes.byref(p), ptrs, None, ctypes.byref(out), None, 0, None) == capi.E_INVALID
    out = capi.YoloppOutputs(1, 1, 1, 1, 1, 1, 1, None, None)
    assert lib.yolopp_get_bboxes(ctypes.byref(p), ptrs, None, ctypes.byref(out), None, 0, None) == capi.E_WORKSPACE


def test_host_api_mirrors_reference_validation():
    import yolopp
    from yolopp.heads import parse_nms_cfg
    assert parse_nms_cfg(dict(type='nms', iou_threshold=0.65)) == dict(
        iou_thr=0.65, nms_offset=0, split_thr=10000, nms_class_agnostic=False, nms_max_num=-1, nms_score_thr=0.0)
    assert parse_nms_cfg(dict(type='nms', iou_threshold=0.5, score_threshold=0.25))['nms_score_thr'] == 0.25
    with pytest.raises(NotImplementedError):
        parse_nms_cfg(dict(type='soft_nms', iou_threshold=0.5))
    with pytest.raises(TypeError):
        parse_nms_cfg(dict(type='nms', iou_threshold=0.5, bogus=1))
    head = yolopp.YOLOCSPHead(num_classes=80, in_channels=[256, 512, 1024],
                              test_cfg=dict(nms_pre=1000, score_thr=0.001, nms=dict(type='nms', iou_threshold=0.65),
                                            max_per_img=300))
    assert head.num_levels == 3 and head.num_attrib == 85 and head.num_anchors == [3, 3, 3]
    assert head.featmap_strides == [8, 16, 32]
    v3 = yolopp.YOLOV3Head(num_classes=80, in_channels=[1024, 512, 256], out_channels=[1024, 512, 256])
    assert v3.featmap_strides == [32, 16, 8] and v3.anchor_generator.base_sizes[0][0] == (116, 90)
    with pytest.raises(NotImplementedError):
        yolopp.YOLOV4BBoxCoder().encode(None, None, 8)
    ag = yolopp.YOLOAnchorGenerator(strides=[32, 16, 8], base_sizes=[[(116, 90)], [(30, 61)], [(10, 13)]])
    anchors = ag.grid_anchors([(2, 3), (4, 6), (8, 12)], device='cpu')
    assert [tuple(a.shape) for a in anchors] == [(6, 4), (24, 4), (96, 4)]
    from oracle import oracle
    np.testing.assert_array_equal(anchors[0].numpy(), oracle.grid_anchors(ag._base[0], 2, 3, 32, 32))


def test_scale_factor_broadcast():
    from yolopp.heads import _scale_factors
    sf = _scale_factors([dict(scale_factor=2.0), dict(scale_factor=np.array([1.5, 2.5, 1.5, 2.5], np.float32))], 2)
    np.testing.assert_array_equal(sf.numpy(), np.array([[2, 2, 2, 2], [1.5, 2.5, 1.5, 2.5]], np.float32))


def test_shard_ranges_partition_the_batch():
    from yolopp.shard import shard_range
    for n in (1, 7, 64, 1024):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


_WORKER = r'''
import os, sys
sys.path[:0] = [{root!r}, os.path.join({root!r}, 'mmdet-yolov4_b200'), os.path.join({root!r}, 'tests')]
import numpy as np, torch, torch.distributed as dist
import cases
from oracle import oracle
from yolopp.shard import shard_range, gather_detections
dist.init_process_group('gloo', init_method='tcp://127.0.0.1:{port}', rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
case = dict(cases.CASES['csp_odd'], batch=5)
p = cases.build_params(case)
levels = cases.host_levels(case, p)
lo, hi = shard_range(5, 2, rank)
ps = cases.build_params(case, batch=hi - lo)
# the CPU checker stands in for the per-GPU kernels here: this test is about the sharding / gather plumbing
loc = oracle.get_bboxes(ps, [x[lo:hi] for x in levels])
res = gather_detections([(loc['dets'][i], loc['labels'][i]) for i in range(hi - lo)], 5)
if rank == 0:
    full = oracle.get_bboxes(p, levels)
    assert len(res) == 5
    for b in range(5):
        assert np.array_equal(res[b][0].numpy().view(np.uint32), full['dets'][b].view(np.uint32)), b
        assert np.array_equal(res[b][1].numpy(), full['labels'][b]), b
    print('SHARD_OK')
dist.destroy_process_group()
'''


def test_two_rank_sharding_with_gloo(tmp_path):
    """world_size-2 gloo run on CPU: each rank post-processes its contiguous shard, rank 0 gathers the detections
    in image order; equals the unsharded result (images are independent: no data-path collective)."""
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    script = tmp_path / 'worker.py'
    script.write_text(_WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                              text=True) for r in range(2)]
    outs = [pr.communicate(timeout=240)[0] for pr in procs]
    assert all(pr.returncode == 0 for pr in procs), outs
    assert 'SHARD_OK' in outs[0]


def test_bbox2result_split():
    """Same contract as mmdet.core.bbox2result (mmdet/core/bbox/transforms.py:99-116)."""
    import yolopp
    dets = np.arange(20, dtype=np.float32).reshape(4, 5)
    labels = np.array([2, 0, 2, 1], np.int64)
    out = yolopp.bbox2result(dets, labels, 3)
    assert [o.shape for o in out] == [(1, 5), (1, 5), (2, 5)]
    np.testing.assert_array_equal(out[2], dets[[0, 2]])
    empty = yolopp.bbox2result(np.zeros((0, 5), np.float32), np.zeros((0, ), np.int64), 3)
    assert [o.shape for o in empty] == [(0, 5)] * 3 and empty[0].dtype == np.float32


def test_plans_for_layouts_and_large_nms_pre():
    """NHWC: row-driven decode (select + rows + NMS), bytes = the admitted rows only; nms_pre beyond the select
    kernel's shared-memory sort capacity is planned (chunked sort), not rejected; 1280^2 keeps the fast top-k by
    staging a sample."""
    nh = cases.build_params(dict(cases.CASES['csp608_sparse'], batch=64))
    nh.layout = capi.LAYOUT_NHWC
    i = capi.describe(nh)
    assert i.kernel_launches == 3 and i.tma_tiles == 0 and i.dense_tiles == 0
    assert i.ldg_bytes_per_image == 4 * 85 * 1000 and i.tma_bytes_per_image == 0
    big = cases.build_params(dict(cases.CASES['csp608_sparse'], nms_pre=5000))
    assert capi.describe(big).rows_per_image == 5000
    assert capi.load_library().yolopp_workspace_bytes(ctypes.byref(big)) > 0
    p1280 = cases.build_params(dict(cases.CASES['csp1280_sparse'], batch=8))
    assert capi.describe(p1280).tma_level_mask == 0b111


def test_decode_tile_size_follows_the_admitted_density():
    """The persistent decode kernel streams 32-position tiles in 8 stages where a 64-position tile would hold two or more
    admitted anchors on average (608^2: 2.8, YOLOv3 640^2: 5.3), 64-position tiles in 4 stages where tiles are mostly empty
    (1280^2: 0.6); either way two CTAs fit an SM, and the tile count follows (the gathered 19^2 level keeps 64)."""
    i608 = capi.describe(cases.build_params(dict(cases.CASES['csp608_sparse'], batch=64)))
    assert i608.decode_tile_positions == 32 and i608.decode_ctas_per_sm == 2
    assert i608.tma_tiles == 64 * 3 * (-(-76 * 76 // 32) + -(-38 * 38 // 32) + -(-19 * 19 // 64))
    i1280 = capi.describe(cases.build_params(dict(cases.CASES['csp1280_sparse'], batch=8)))
    assert i1280.decode_tile_positions == 64 and i1280.decode_ctas_per_sm == 2
    iv3 = capi.describe(cases.build_params(dict(cases.CASES['v3_640_sparse'], batch=8)))
    assert iv3.decode_tile_positions == 32 and iv3.dense_tiles > 0
    nh = cases.build_params(dict(cases.CASES['csp608_sparse'], batch=2))
    nh.layout = capi.LAYOUT_NHWC
    assert capi.describe(nh).decode_tile_positions == 0  # kernel not launched


def test_plan_and_stage_entries_reject_bad_calls_without_a_device():
    lib = capi.load_library()
    p = cases.build_params(cases.CASES['csp_tiny'])
    ptrs = (ctypes.c_void_p * 3)()
    h = ctypes.c_void_p()
    out = capi.YoloppOutputs(1, 1, 1, 1, 1, 1, 1, None, None)
    assert lib.yolopp_plan_create(ctypes.byref(p), ptrs, None, ctypes.byref(out), None, 0, ctypes.byref(h)) == capi.E_WORKSPACE
    assert not h.value
    assert lib.yolopp_plan_run(None, None) == capi.E_INVALID
    lib.yolopp_plan_destroy(None)
    assert lib.yolopp_topk_conf(ctypes.byref(p), ptrs, None, None, 0, None) == capi.E_INVALID
    assert lib.yolopp_decode(ctypes.byref(p), ptrs, None, None, None, None, None, 0, None) == capi.E_INVALID
    assert lib.yolopp_mish_forward(None, None, -1, 0, None) == capi.E_INVALID
    assert lib.yolopp_mish_forward(ctypes.c_void_p(16), ctypes.c_void_p(16), 4, 7, None) == capi.E_INVALID
    assert lib.yolopp_mish_backward(ctypes.c_void_p(16), ctypes.c_void_p(16), ctypes.c_void_p(8), 4, 0, None) == capi.E_INVALID


@pytest.mark.skipif(not os.path.isfile('/root/reference/mmdet/models/dense_heads/yolocsp_head.py'),
                    reason='reference tree not present (GPU box)')
@pytest.mark.parametrize('name', ['csp608_sparse', 'tencent_agnostic', 'csp416_rescale', 'v3_416_sparse', 'csp_nms_offset1'])
def test_patch_head_on_the_live_reference_head(name):
    """patch_head() on an instance of the REFERENCE's own head class (mmdet/models/dense_heads/yolocsp_head.py:54-178,
    yolo_head.py:20-110, executed from /root/reference): the yolopp_params it extracts from that instance
    (anchor_generator, featmap_strides, num_classes, class_agnostic, test_cfg) are byte-identical to the params the
    parity tests use for the same configuration — i.e. the drop-in binding reads a real mmdet head correctly."""
    import importlib.util
    import torch
    import yolopp
    from yolopp import heads
    from oracle import refexec
    spec = importlib.util.spec_from_file_location('make_golden', os.path.join(ROOT, 'tests', 'golden', 'make_golden.py'))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    case = cases.CASES[name]
    ref = refexec.load_reference()
    head = mg.build_ref_head(ref, case)
    assert type(head).__module__ != yolopp.heads.__name__  # really the reference's class
    yolopp.patch_head(head)
    p_ref = heads.head_params(head, pred_shapes=case['sizes'], batch=case['batch'], rescale=case.get('rescale', False))
    p_own = cases.build_params(case)
    assert bytes(p_ref) == bytes(p_own)
    # the patched method validates like the reference (level count) before touching the device
    with pytest.raises(AssertionError):
        head.get_bboxes([torch.zeros(1, 1, 1, 1)], [dict(scale_factor=1.0)])
