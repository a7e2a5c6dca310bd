"""Bit-reproducible synthetic head tensors, generated on the device (yolopp_synth_level).

Distributions:
  'sparse'  COCO-like: box logits N(0,1), objectness N(-5, 2^2), class logits N(-4.9, 1.5^2)
  'dense'   every logit N(0,1): (almost) every anchor above score_thr — stresses compaction / top-k / NMS
The variates are Irwin-Hall(4) sums from a counter-based hash (support +-3.46 sigma).
"""
import ctypes

import torch

from . import _capi

DISTS = {
    'sparse': ((0.0, -5.0, -4.9), (1.0, 2.0, 1.5)),
    'dense': ((0.0, 0.0, 0.0), (1.0, 1.0, 1.0)),
}


def level_seed(seed, level):
    return (int(seed) * 1000003 + int(level) * 7919 + 12345) % (1 << 64)


def dist_stats(dist):
    return DISTS[dist] if isinstance(dist, str) else dist


def synth_levels(params, seed, dist='sparse', device='cuda'):
    """Returns the list of raw head tensors (B, A*(5+C), H, W) for `params`."""
    lib = _capi.load_library()
    mean, std = dist_stats(dist)
    m3 = (ctypes.c_float * 3)(*mean)
    s3 = (ctypes.c_float * 3)(*std)
    out = []
    dev = torch.device(device)
    with torch.cuda.device(dev):
        for l in range(params.num_levels):
            t = torch.empty(params.level_shape(l), dtype=torch.float32, device=dev)
            rc = lib.yolopp_synth_level(ctypes.c_void_p(t.data_ptr()), params.batch, params.num_anchors,
                                        params.num_attrib, params.height[l] * params.width[l], m3, s3,
                                        ctypes.c_uint64(level_seed(seed, l)),
                                        ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
            _capi.check(rc, 'yolopp_synth_level')
            out.append(t)
    return out
