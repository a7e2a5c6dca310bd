"""ctypes view of include/yolopp.h: the params/outputs structs, the library loader and the params builder.

The product path has NO CPU fallback: `load_library()` raises if the sm_100a shared library has not been
built (python -c "import __graft_entry__ as g; g.build()"), and every op in this package goes through it.
"""
import ctypes
import os
from ctypes import c_float, c_int32, c_int64, c_size_t, c_uint64, c_void_p

import numpy as np

ABI_VERSION = 3
MAX_LEVELS = 8
MAX_ANCHORS = 8
MAX_CLASSES = 4096
MAX_ROWS = 1 << 20

MODE_CSP = 0
MODE_V3 = 1
LAYOUT_NCHW = 0
LAYOUT_NHWC = 1
DTYPE_F32, DTYPE_F16, DTYPE_BF16 = 0, 1, 2

OK, E_INVALID, E_WORKSPACE, E_OVERFLOW, E_NO_DEVICE, E_CUDA = 0, 1, 2, 3, 4, 1000

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('YOLOPP_LIB') or os.path.join(os.path.dirname(PKG_DIR), 'csrc', 'libyolopp.so')


class YoloppParams(ctypes.Structure):
    """struct yolopp_params (include/yolopp.h)."""
    _fields_ = [
        ('abi_version', c_int32),
        ('mode', c_int32),
        ('batch', c_int32),
        ('num_levels', c_int32),
        ('num_anchors', c_int32),
        ('num_classes', c_int32),
        ('class_agnostic', c_int32),
        ('height', c_int32 * MAX_LEVELS),
        ('width', c_int32 * MAX_LEVELS),
        ('stride_w', c_int32 * MAX_LEVELS),
        ('stride_h', c_int32 * MAX_LEVELS),
        ('coder_stride', c_int32 * MAX_LEVELS),
        ('base_anchors', ((c_float * 4) * MAX_ANCHORS) * MAX_LEVELS),
        ('nms_pre', c_int32),
        ('score_thr', c_float),
        ('conf_thr', c_float),
        ('iou_thr', c_float),
        ('nms_offset', c_int32),
        ('split_thr', c_int32),
        ('nms_class_agnostic', c_int32),
        ('nms_max_num', c_int32),
        ('max_per_img', c_int32),
        ('rescale', c_int32),
        ('out_capacity', c_int32),
        ('batches_in_flight', c_int32),
        ('layout', c_int32),
        ('nms_score_thr', c_float),
        ('reserved', c_int32 * 4),
    ]

    # convenience -----------------------------------------------------------------------------
    @property
    def num_attrib(self):
        return 5 if self.class_agnostic else 5 + self.num_classes

    @property
    def eff_classes(self):
        return 1 if self.class_agnostic else self.num_classes

    def level_shape(self, l):
        """LOGICAL shape of level l (both layouts: NHWC is the channels-last memory of this shape)."""
        return (self.batch, self.num_anchors * self.num_attrib, self.height[l], self.width[l])

    @property
    def rows_per_image(self):
        return describe(self).rows_per_image

    @property
    def anchors_per_image(self):
        return sum(self.height[l] * self.width[l] * self.num_anchors for l in range(self.num_levels))

    @property
    def capacity(self):
        return self.out_capacity if self.out_capacity > 0 else self.max_per_img


class YoloppOutputs(ctypes.Structure):
    """struct yolopp_outputs (include/yolopp.h) — device pointers."""
    _fields_ = [
        ('dets', c_void_p),
        ('labels', c_void_p),
        ('anchors', c_void_p),
        ('rows', c_void_p),
        ('count', c_void_p),
        ('num_candidates', c_void_p),
        ('status', c_void_p),
        ('cls_dets', c_void_p),
        ('cls_offsets', c_void_p),
    ]


class YoloppPlanInfo(ctypes.Structure):
    """struct yolopp_plan_info (include/yolopp.h)."""
    _fields_ = [
        ('anchors_per_image', c_int32),
        ('rows_per_image', c_int32),
        ('num_attrib', c_int32),
        ('tma_level_mask', c_int32),
        ('tma_tiles', c_int32),
        ('ldg_blocks', c_int32),
        ('decode_smem_bytes', c_int32),
        ('decode_ctas_per_sm', c_int32),
        ('kernel_launches', c_int32),
        ('dense_tiles', c_int32),
        ('tma_bytes_per_image', c_int64),
        ('ldg_bytes_per_image', c_int64),
        ('workspace_bytes', c_int64),
        ('decode_tile_positions', c_int32),
        ('reserved_', c_int32),
    ]


NUM_STAGE_EVENTS = 5
STAGE_NAMES = ('select', 'decode_tma', 'decode_ldg', 'nms_image')


def _pair(v):
    return (int(v[0]), int(v[1])) if isinstance(v, (tuple, list, np.ndarray)) else (int(v), int(v))


def yolo_base_anchors(base_sizes, strides):
    """YOLOAnchorGenerator.gen_base_anchors (mmdet/core/anchor/anchor_generator.py:605-616,639-665):
    centre = stride/2, corners computed in double and rounded ONCE to fp32 (torch.Tensor([...]))."""
    out = []
    for sizes, stride in zip(base_sizes, strides):
        sw, sh = _pair(stride)
        cx, cy = sw / 2., sh / 2.
        lvl = []
        for size in sizes:
            w, h = (size if isinstance(size, (tuple, list)) else (size, size))
            lvl.append([cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h])
        out.append(np.asarray(lvl, dtype=np.float64).astype(np.float32))
    return out


def make_params(mode, batch, featmap_sizes, anchor_strides, coder_strides, base_sizes, num_classes,
                class_agnostic=False, nms_pre=-1, score_thr=0.0, conf_thr=-1.0, iou_thr=0.5, nms_offset=0,
                split_thr=10000, nms_class_agnostic=False, nms_max_num=-1, max_per_img=-1, rescale=False,
                out_capacity=0, base_anchors=None, layout=LAYOUT_NCHW, nms_score_thr=0.0):
    """Builds yolopp_params from what the reference reads off the head instance and test_cfg
    (yolocsp_head.py:112-114,151,162,170-178,345-348,374-376; yolo_head.py:52-59,281,365,378-384)."""
    L = len(featmap_sizes)
    if not (1 <= L <= MAX_LEVELS):
        raise ValueError(f'num_levels must be in [1, {MAX_LEVELS}]')
    if not (len(anchor_strides) == len(coder_strides) == L):
        raise ValueError('strides / featmap_sizes length mismatch')
    if base_anchors is None:
        if len(base_sizes) != L:
            raise ValueError('base_sizes / featmap_sizes length mismatch')
        base_anchors = yolo_base_anchors(base_sizes, anchor_strides)
    A = int(base_anchors[0].shape[0])
    if not (1 <= A <= MAX_ANCHORS) or any(b.shape != (A, 4) for b in base_anchors):
        raise ValueError(f'every level needs the same number (1..{MAX_ANCHORS}) of base anchors')
    p = YoloppParams()
    p.abi_version = ABI_VERSION
    p.mode = int(mode)
    p.batch = int(batch)
    p.num_levels = L
    p.num_anchors = A
    p.num_classes = int(num_classes)
    p.class_agnostic = int(bool(class_agnostic))
    for l in range(L):
        p.height[l], p.width[l] = int(featmap_sizes[l][0]), int(featmap_sizes[l][1])
        p.stride_w[l], p.stride_h[l] = _pair(anchor_strides[l])
        p.coder_stride[l] = int(coder_strides[l])
        for a in range(A):
            for k in range(4):
                p.base_anchors[l][a][k] = float(base_anchors[l][a, k])
    p.nms_pre = int(nms_pre)
    p.score_thr = float(score_thr)
    p.conf_thr = float(conf_thr)
    p.iou_thr = float(iou_thr)
    p.nms_offset = int(nms_offset)
    p.split_thr = int(split_thr)
    p.nms_class_agnostic = int(bool(nms_class_agnostic))
    p.nms_max_num = int(nms_max_num)
    p.max_per_img = int(max_per_img)
    p.rescale = int(bool(rescale))
    p.out_capacity = int(out_capacity)
    p.layout = int(layout)
    p.nms_score_thr = float(nms_score_thr)
    return p


_LIB = None


def load_library(path=None):
    """dlopen the sm_100a shared library and declare every entry point of include/yolopp.h.
    Raises RuntimeError when it has not been built — there is deliberately no fallback."""
    global _LIB
    if _LIB is not None and path is None:
        return _LIB
    path = path or LIB_PATH
    if not os.path.isfile(path):
        raise RuntimeError(
            f'{path} not found: build the CUDA library first (python -c "import __graft_entry__ as g; g.build()"). '
            'yolopp has no CPU / PyTorch fallback.')
    lib = ctypes.CDLL(path)
    pp = ctypes.POINTER(YoloppParams)
    po = ctypes.POINTER(YoloppOutputs)
    lib.yolopp_abi_version.restype = ctypes.c_int
    lib.yolopp_abi_version.argtypes = []
    lib.yolopp_strerror.restype = ctypes.c_char_p
    lib.yolopp_strerror.argtypes = [ctypes.c_int]
    lib.yolopp_workspace_bytes.restype = c_size_t
    lib.yolopp_workspace_bytes.argtypes = [pp]
    lib.yolopp_get_bboxes.restype = ctypes.c_int
    lib.yolopp_get_bboxes.argtypes = [pp, ctypes.POINTER(c_void_p), c_void_p, po, c_void_p, c_size_t, c_void_p]
    lib.yolopp_get_bboxes_profiled.restype = ctypes.c_int
    lib.yolopp_get_bboxes_profiled.argtypes = [pp, ctypes.POINTER(c_void_p), c_void_p, po, c_void_p, c_size_t,
                                               c_void_p, ctypes.POINTER(c_void_p), ctypes.c_int]
    lib.yolopp_plan_create.restype = ctypes.c_int
    lib.yolopp_plan_create.argtypes = [pp, ctypes.POINTER(c_void_p), c_void_p, po, c_void_p, c_size_t,
                                       ctypes.POINTER(c_void_p)]
    lib.yolopp_plan_run.restype = ctypes.c_int
    lib.yolopp_plan_run.argtypes = [c_void_p, c_void_p]
    lib.yolopp_plan_run_profiled.restype = ctypes.c_int
    lib.yolopp_plan_run_profiled.argtypes = [c_void_p, c_void_p, ctypes.POINTER(c_void_p), ctypes.c_int]
    lib.yolopp_plan_destroy.restype = None
    lib.yolopp_plan_destroy.argtypes = [c_void_p]
    lib.yolopp_topk_conf.restype = ctypes.c_int
    lib.yolopp_topk_conf.argtypes = [pp, ctypes.POINTER(c_void_p), c_void_p, c_void_p, c_size_t, c_void_p]
    lib.yolopp_decode.restype = ctypes.c_int
    lib.yolopp_decode.argtypes = [pp, ctypes.POINTER(c_void_p), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_size_t, c_void_p]
    lib.yolopp_mish_forward.restype = ctypes.c_int
    lib.yolopp_mish_forward.argtypes = [c_void_p, c_void_p, c_int64, ctypes.c_int, c_void_p]
    lib.yolopp_mish_backward.restype = ctypes.c_int
    lib.yolopp_mish_backward.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, ctypes.c_int, c_void_p]
    lib.yolopp_describe.restype = ctypes.c_int
    lib.yolopp_describe.argtypes = [pp, ctypes.POINTER(YoloppPlanInfo)]
    lib.yolopp_coder_decode.restype = ctypes.c_int
    lib.yolopp_coder_decode.argtypes = [ctypes.c_int, c_void_p, c_void_p, c_float, c_int64, c_void_p, c_void_p]
    lib.yolopp_nms_workspace_bytes.restype = c_size_t
    lib.yolopp_nms_workspace_bytes.argtypes = [c_int64, c_int32]
    lib.yolopp_batched_nms.restype = ctypes.c_int
    lib.yolopp_batched_nms.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_float, c_float, ctypes.c_int,
                                       ctypes.c_int, ctypes.c_int, ctypes.c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                       c_size_t, c_void_p]
    lib.yolopp_multiclass_nms.restype = ctypes.c_int
    lib.yolopp_multiclass_nms.argtypes = [c_void_p, ctypes.c_int, c_void_p, c_int64, c_int32, c_float, c_void_p, c_float,
                                          c_float, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_void_p,
                                          c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]
    lib.yolopp_synth_level.restype = ctypes.c_int
    lib.yolopp_synth_level.argtypes = [c_void_p, c_int32, c_int32, c_int32, c_int32, ctypes.POINTER(c_float),
                                       ctypes.POINTER(c_float), c_uint64, c_void_p]
    lib.yolopp_sigmoid.restype = ctypes.c_int
    lib.yolopp_sigmoid.argtypes = [c_void_p, c_void_p, c_int64, c_void_p]
    lib.yolopp_exp.restype = ctypes.c_int
    lib.yolopp_exp.argtypes = [c_void_p, c_void_p, c_int64, c_void_p]
    if lib.yolopp_abi_version() != ABI_VERSION:
        raise RuntimeError(f'{path}: ABI version {lib.yolopp_abi_version()} != {ABI_VERSION}')
    if path == LIB_PATH:
        _LIB = lib
    return lib


EXPORTED_SYMBOLS = ('yolopp_abi_version', 'yolopp_strerror', 'yolopp_workspace_bytes', 'yolopp_get_bboxes',
                    'yolopp_get_bboxes_profiled', 'yolopp_describe', 'yolopp_plan_create', 'yolopp_plan_run',
                    'yolopp_plan_run_profiled', 'yolopp_plan_destroy', 'yolopp_topk_conf', 'yolopp_decode',
                    'yolopp_mish_forward', 'yolopp_mish_backward',
                    'yolopp_coder_decode', 'yolopp_nms_workspace_bytes', 'yolopp_batched_nms', 'yolopp_multiclass_nms',
                    'yolopp_synth_level',
                    'yolopp_sigmoid', 'yolopp_exp')


def describe(params):
    info = YoloppPlanInfo()
    check(load_library().yolopp_describe(ctypes.byref(params), ctypes.byref(info)), 'yolopp_describe')
    return info


def check(code, what='yolopp'):
    if code != OK:
        lib = load_library()
        raise RuntimeError(f'{what} failed: {lib.yolopp_strerror(code).decode()} (code {code})')
