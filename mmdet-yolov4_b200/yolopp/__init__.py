"""yolopp — B200-native (sm_100a) YOLOv4/v5/v3 detection post-processing.

Drop-in for ONE hot path of zhanggefan/mmdet-yolov4 (`get_bboxes -> bbox_coder.decode -> multiclass_nms ->
batched_nms`), behind the reference's own call signatures. See DESIGN.md / INTEGRATION.md.
There is no CPU or PyTorch fallback: every op runs the hand-written CUDA library through its C ABI.
"""
from . import _capi
from ._capi import MODE_CSP, MODE_V3, make_params, load_library
from .ops import get_bboxes_raw, coder_decode, sigmoid, exp, topk_conf, decode
from .heads import (YOLOCSPHead, YOLOV3Head, YOLOV4BBoxCoder, YOLOBBoxCoder, YOLOAnchorGenerator,
                    YOLOV4AnchorGenerator, patch_head, bbox2result, head_params)
from .nms import multiclass_nms, batched_nms, nms
from .mish import Mish, MishCudaFunction, mish_forward, mish_backward
from . import synth
from . import shard

__all__ = [
    'MODE_CSP', 'MODE_V3', 'make_params', 'load_library', 'get_bboxes_raw', 'coder_decode', 'sigmoid', 'exp',
    'YOLOCSPHead', 'YOLOV3Head', 'YOLOV4BBoxCoder', 'YOLOBBoxCoder', 'YOLOAnchorGenerator', 'YOLOV4AnchorGenerator',
    'patch_head', 'bbox2result', 'head_params', 'multiclass_nms', 'batched_nms', 'nms', 'synth', 'shard', 'topk_conf',
    'decode', 'Mish', 'MishCudaFunction', 'mish_forward', 'mish_backward'
]
