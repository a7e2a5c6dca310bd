"""Mish activation on the CUDA library — mirror of mmdet/ops/mish_cuda/mish.py:18-48 (MishCudaFunction, Mish) over
`torch.ops.yolopp.mish_forward / mish_backward` (sm_100a kernels in csrc/yolopp_mish.cuh; current stream, 128-bit
accesses, float32 / float16 / bfloat16 with fp32 math like mish.h:33-50). CUDA tensors only: the reference's CPU branch
(mish_cpu.cc) is NOT mirrored — a CPU tensor raises from the dispatcher."""
import torch

from . import ops  # noqa: F401  (registers torch.ops.yolopp.*)

try:
    from torch.amp import custom_bwd, custom_fwd

    def _fwd(f):
        return custom_fwd(f, device_type='cuda')

    def _bwd(f):
        return custom_bwd(f, device_type='cuda')
except ImportError:  # older torch
    from torch.cuda.amp import custom_bwd as _bwd, custom_fwd as _fwd


class MishCudaFunction(torch.autograd.Function):

    @staticmethod
    @_fwd
    def forward(ctx, inp):
        if not inp.is_contiguous():
            inp = inp.contiguous()
        ctx.save_for_backward(inp)
        return torch.ops.yolopp.mish_forward(inp)

    @staticmethod
    @_bwd
    def backward(ctx, grad_out):
        inp, = ctx.saved_tensors
        if not grad_out.is_contiguous():
            grad_out = grad_out.contiguous()
        if not ctx.needs_input_grad[0]:
            return (None, )
        return torch.ops.yolopp.mish_backward(grad_out, inp)


class Mish(torch.nn.Module):
    """Drop-in for the `Mish` the reference registers in mmcv's ACTIVATION_LAYERS (mish.py:41-48)."""

    def __init__(self, **kwargs):
        super().__init__()

    def forward(self, inp):
        return MishCudaFunction.apply(inp)


def mish_forward(inp):
    return torch.ops.yolopp.mish_forward(inp)


def mish_backward(grad_out, inp):
    return torch.ops.yolopp.mish_backward(grad_out, inp)
