"""models/networks/function.py:3-35 (AdaIN helpers; dead code in the reference, named by north_star) on the
segmented-reduction kernel.  NCHW CUDA tensors in, same shapes out as the reference."""
import torch

from . import ops


def _stats(feat):
    assert feat.dim() == 4
    n, c, h, w = feat.shape
    x = feat.contiguous()
    r1 = ops.plane_reduce(x, False, want=("sum",))
    mean = (r1["sum"].view(n, c) / float(h * w)).contiguous()
    r2 = ops.plane_reduce(x, False, shift=mean, want=("sum", "sumsq"))
    hw = float(h * w)
    m2 = r2["sumsq"].view(n, c) - r2["sum"].view(n, c) ** 2 / hw          # sum (x-mean)^2, exact two-pass
    return mean, m2, hw


def calc_mean_std(feat, eps=1e-5):
    """function.py:3-11: UNBIASED variance + eps."""
    n, c = feat.shape[:2]
    mean, m2, hw = _stats(feat)
    std = (m2 / (hw - 1.0) + eps).sqrt()
    return mean.view(n, c, 1, 1).to(feat.dtype), std.view(n, c, 1, 1).to(feat.dtype)


def calc_mean(feat):
    """function.py:13-18."""
    n, c = feat.shape[:2]
    return _stats(feat)[0].view(n, c, 1, 1).to(feat.dtype)


def nor_mean_std(feat):
    """function.py:20-24."""
    mean, std = calc_mean_std(feat)
    n, c, h, w = feat.shape
    a = (1.0 / std.float()).view(n, c).contiguous()
    b = (-mean.float().view(n, c) * a).contiguous()
    x = ops.nchw_to_nhwc(feat)
    return ops.nhwc_to_nchw(ops.affine_apply(x, c, a, b))


def nor_mean(feat):
    """function.py:26-30."""
    mean = calc_mean(feat)
    n, c, h, w = feat.shape
    x = ops.nchw_to_nhwc(feat)
    b = (-mean.float().view(n, c)).contiguous()
    return ops.nhwc_to_nchw(ops.affine_apply(x, c, None, b)), mean


def calc_cov(feat):
    """function.py:32-35: [N,C,H,W] -> [N,C,C] = X X^T / (H*W)."""
    n, c, h, w = feat.shape
    x = feat.float().reshape(n, c, h * w).contiguous()
    return ops.gemm_f32(x, x, True, alpha=1.0 / (h * w)).to(feat.dtype)
