"""upfirdn2d with the reference's signature and autograd behaviour
(models/networks/stylegan2_op/upfirdn2d.py:24-159) on ppst_upfirdn2d (include/ppst_b200.h).

Unlike the reference's native op, every (up, down, kernel) combination up to 32x32 taps is computed
(the reference silently returns uninitialised memory outside its six template modes,
upfirdn2d_kernel.cu:172-223), bfloat16 is accepted, and non-CUDA input raises.
"""
import torch
from torch.autograd import Function

from .. import _lib


def _upfirdn2d_raw(x4, kernel, up_x, up_y, down_x, down_y, px0, px1, py0, py1):
    """The reference's `upfirdn2d_op.upfirdn2d(input[major,H,W,minor], kernel, ...)` (upfirdn2d.cpp:12-23)."""
    _lib.require_cuda(x4, kernel)
    lib = _lib.load()
    x4 = x4.contiguous()
    k = kernel.contiguous().to(torch.float32)
    major, in_h, in_w, minor = x4.shape
    kh, kw = k.shape
    out_h = (in_h * up_y + py0 + py1 - kh + down_y) // down_y
    out_w = (in_w * up_x + px0 + px1 - kw + down_x) // down_x
    if out_h <= 0 or out_w <= 0:
        raise RuntimeError("ppst_b200.upfirdn2d: empty output (%d x %d)" % (out_h, out_w))
    y = torch.empty((major, out_h, out_w, minor), device=x4.device, dtype=x4.dtype)
    _lib.check(lib.ppst_upfirdn2d(y.data_ptr(), x4.data_ptr(), k.data_ptr(), major, in_h, in_w, minor, kh, kw,
                                  up_x, up_y, down_x, down_y, px0, px1, py0, py1, _lib.dt(x4), _lib.stream()))
    return y


class UpFirDn2dBackward(Function):
    """upfirdn2d.py:24-90."""

    @staticmethod
    def forward(ctx, grad_output, kernel, grad_kernel, up, down, pad, g_pad, in_size, out_size):
        up_x, up_y = up
        down_x, down_y = down
        g_pad_x0, g_pad_x1, g_pad_y0, g_pad_y1 = g_pad
        grad_output = grad_output.reshape(-1, out_size[0], out_size[1], 1)
        grad_input = _upfirdn2d_raw(grad_output, grad_kernel, down_x, down_y, up_x, up_y,
                                    g_pad_x0, g_pad_x1, g_pad_y0, g_pad_y1)
        grad_input = grad_input.view(in_size[0], in_size[1], in_size[2], in_size[3])
        ctx.save_for_backward(kernel)
        ctx.up, ctx.down, ctx.pad = up, down, pad
        ctx.in_size, ctx.out_size = in_size, out_size
        return grad_input

    @staticmethod
    def backward(ctx, gradgrad_input):
        kernel, = ctx.saved_tensors
        gradgrad_input = gradgrad_input.reshape(-1, ctx.in_size[2], ctx.in_size[3], 1)
        gradgrad_out = _upfirdn2d_raw(gradgrad_input, kernel, ctx.up[0], ctx.up[1], ctx.down[0], ctx.down[1], *ctx.pad)
        gradgrad_out = gradgrad_out.view(ctx.in_size[0], ctx.in_size[1], ctx.out_size[0], ctx.out_size[1])
        return gradgrad_out, None, None, None, None, None, None, None, None


class UpFirDn2d(Function):
    """upfirdn2d.py:93-147."""

    @staticmethod
    def forward(ctx, input, kernel, up, down, pad):
        up_x, up_y = up
        down_x, down_y = down
        pad_x0, pad_x1, pad_y0, pad_y1 = pad
        kernel_h, kernel_w = kernel.shape
        batch, channel, in_h, in_w = input.shape
        ctx.in_size = input.shape
        input = input.reshape(-1, in_h, in_w, 1)
        ctx.save_for_backward(kernel, torch.flip(kernel, [0, 1]))
        out_h = (in_h * up_y + pad_y0 + pad_y1 - kernel_h) // down_y + 1
        out_w = (in_w * up_x + pad_x0 + pad_x1 - kernel_w) // down_x + 1
        ctx.out_size = (out_h, out_w)
        ctx.up, ctx.down, ctx.pad = (up_x, up_y), (down_x, down_y), (pad_x0, pad_x1, pad_y0, pad_y1)
        g_pad_x0 = kernel_w - pad_x0 - 1
        g_pad_y0 = kernel_h - pad_y0 - 1
        g_pad_x1 = in_w * up_x - out_w * down_x + pad_x0 - up_x + 1
        g_pad_y1 = in_h * up_y - out_h * down_y + pad_y0 - up_y + 1
        ctx.g_pad = (g_pad_x0, g_pad_x1, g_pad_y0, g_pad_y1)
        out = _upfirdn2d_raw(input, kernel, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1)
        return out.view(-1, channel, out_h, out_w)

    @staticmethod
    def backward(ctx, grad_output):
        kernel, grad_kernel = ctx.saved_tensors
        grad_input = UpFirDn2dBackward.apply(grad_output, kernel, grad_kernel, ctx.up, ctx.down, ctx.pad, ctx.g_pad,
                                             ctx.in_size, ctx.out_size)
        return grad_input, None, None, None, None


def upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0)):
    """upfirdn2d.py:150-159: same pad on x and y; negative pads crop."""
    return UpFirDn2d.apply(input, kernel, (up, up), (down, down), (pad[0], pad[1], pad[0], pad[1]))
