"""fused_leaky_relu / FusedLeakyReLU with the reference's signatures and autograd behaviour
(models/networks/stylegan2_op/fused_act.py:23-96), on ppst_fused_bias_act (include/ppst_b200.h).

Differences from the reference, all deliberate:
  * no JIT build at import, no pybind module: one C-ABI call per op on the current CUDA stream;
  * bfloat16 is supported (the reference dispatches FLOATING_TYPES_AND_HALF only);
  * non-CUDA input raises instead of silently taking a PyTorch fallback.
"""
import torch
from torch import nn
from torch.autograd import Function

from .. import _lib

ACT_LRELU = 3


def _bias_act(x, b, ref, act, grad, alpha, scale):
    """The reference's `fused.fused_bias_act(input, bias, refer, act, grad, alpha, scale)`
    (fused_bias_act.cpp:11-20): empty tensors mean "absent"."""
    _lib.require_cuda(x, b, ref)
    lib = _lib.load()
    x = x.contiguous()
    y = torch.empty_like(x)
    has_b = b is not None and b.numel() > 0
    has_ref = ref is not None and ref.numel() > 0
    if has_b:
        b = b.contiguous().to(x.dtype)
    if has_ref:
        ref = ref.contiguous().to(x.dtype)
    step_b = 1
    for s in x.shape[2:]:
        step_b *= s
    if x.numel():
        _lib.check(lib.ppst_fused_bias_act(y.data_ptr(), x.data_ptr(), b.data_ptr() if has_b else None,
                                           ref.data_ptr() if has_ref else None, x.numel(), step_b,
                                           b.numel() if has_b else 0, act, grad, float(alpha), float(scale),
                                           _lib.dt(x), _lib.stream()))
    return y


class FusedLeakyReLUFunctionBackward(Function):
    """fused_act.py:23-52."""

    @staticmethod
    def forward(ctx, grad_output, out, negative_slope, scale):
        ctx.save_for_backward(out)
        ctx.negative_slope = negative_slope
        ctx.scale = scale
        grad_input = _bias_act(grad_output, None, out, ACT_LRELU, 1, negative_slope, scale)
        lib = _lib.load()
        c = grad_input.shape[1] if grad_input.dim() > 1 else 1
        outer = grad_input.shape[0]
        step = grad_input.numel() // max(outer * c, 1)
        grad_bias = torch.empty(c, device=grad_input.device, dtype=torch.float32)
        _lib.check(lib.ppst_bias_grad_reduce(grad_bias.data_ptr(), grad_input.data_ptr(), outer, c, step,
                                             _lib.dt(grad_input), _lib.stream()))
        return grad_input, grad_bias.to(grad_input.dtype)

    @staticmethod
    def backward(ctx, gradgrad_input, gradgrad_bias):
        out, = ctx.saved_tensors
        gradgrad_out = _bias_act(gradgrad_input, gradgrad_bias, out, ACT_LRELU, 1, ctx.negative_slope, ctx.scale)
        return gradgrad_out, None, None, None


class FusedLeakyReLUFunction(Function):
    """fused_act.py:55-74."""

    @staticmethod
    def forward(ctx, input, bias, negative_slope, scale):
        out = _bias_act(input, bias, None, ACT_LRELU, 0, negative_slope, scale)
        ctx.save_for_backward(out)
        ctx.negative_slope = negative_slope
        ctx.scale = scale
        return out

    @staticmethod
    def backward(ctx, grad_output):
        out, = ctx.saved_tensors
        grad_input, grad_bias = FusedLeakyReLUFunctionBackward.apply(grad_output, out, ctx.negative_slope, ctx.scale)
        return grad_input, grad_bias, None, None


class FusedLeakyReLU(nn.Module):
    """fused_act.py:77-86: `.bias` Parameter [channel], zero-initialised (state-dict key `...bias`)."""

    def __init__(self, channel, negative_slope=0.2, scale=2 ** 0.5):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(channel))
        self.negative_slope = negative_slope
        self.scale = scale

    def forward(self, input):
        return fused_leaky_relu(input, self.bias, self.negative_slope, self.scale)


def fused_leaky_relu(input, bias, negative_slope=0.2, scale=2 ** 0.5):
    """fused_act.py:89-96 (the custom-kernel branch; there is no other branch here)."""
    return FusedLeakyReLUFunction.apply(input, bias, negative_slope, scale)
