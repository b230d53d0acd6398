"""One-time weight packing: reference state_dict tensors (SURVEY.md Appendix A key layout) -> the layouts the
kernels consume.  Pure tensor reshuffling at load time; no per-step work happens here."""
import math
from dataclasses import dataclass, field
from typing import Callable, Optional, Tuple

import torch
import torch.nn.functional as F

SQRT2 = math.sqrt(2.0)


def pack_taps(w):
    """[co, ci, kh, kw] -> [co, kh*kw*ci]  (tap-major, channel-minor: matches NHWC implicit GEMM)."""
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous()


def fold_fir(w, k):
    """conv(w) o upfirdn2d(k)  ==  conv(w (*) flip(k)): fold the FIR of Blur (stylegan2_layers.py:142-164,
    upfirdn2d with up=down=1) into the following convolution's taps.  w [co,ci,kh,kw], k [fh,fw] ->
    [co, ci, kh+fh-1, kw+fw-1].  Exact in real arithmetic."""
    co, ci, kh, kw = w.shape
    kf = torch.flip(k.to(w.dtype), [0, 1]).view(1, 1, *k.shape)
    out = F.conv_transpose2d(w.reshape(co * ci, 1, kh, kw), kf)      # polynomial product of the two stencils
    return out.view(co, ci, kh + k.shape[0] - 1, kw + k.shape[1] - 1)


def pack_upscale_phases(w):
    """EqualizedConv2d fused upscale (stylegan2_layers.py:312-321): the 3x3 weight is box-summed to a 4x4
    transposed-conv kernel W'[ci,co,4,4] (stride 2, padding 1).  Output pixel (2i+a, 2j+b) only touches the 2x2
    input window starting at (i-1+a, j-1+b) with taps W'[.., 3-a-2t, 3-b-2u]  ->  four plain 2x2 convolutions.
    Returns [4 (phase a*2+b), co, 2*2*ci]."""
    wt = F.pad(w.permute(1, 0, 2, 3), [1, 1, 1, 1])
    wt = wt[:, :, 1:, 1:] + wt[:, :, :-1, 1:] + wt[:, :, 1:, :-1] + wt[:, :, :-1, :-1]      # [ci, co, 4, 4]
    ci, co = wt.shape[:2]
    out = w.new_empty(2, 2, co, 2, 2, ci)
    for a in range(2):
        for b in range(2):
            for t in range(2):
                for u in range(2):
                    out[a, b, :, t, u, :] = wt[:, :, 3 - a - 2 * t, 3 - b - 2 * u].t()
    return out.reshape(4, co, 4 * ci).contiguous()


@dataclass
class ConvSpec:
    """A convolution ready to launch.  `w32` [phases][co][K] f32 (CUDA-core path / fp32 mode) and `w16` bf16
    (tcgen05 path) hold the same numbers."""
    cin: int
    cout: int
    kh: int
    kw: int
    stride: int = 1
    off: Tuple[int, int] = (0, 0)                  # input coordinate of tap (0,0) for output (0,0)
    phases: Tuple[int, int] = (1, 1)
    phase_off: Tuple[int, int] = (0, 0)
    out_hw: Callable = None                        # (in_h, in_w) -> (out_h, out_w) of one phase
    w32: torch.Tensor = None
    w16: Optional[torch.Tensor] = None
    bias: Optional[torch.Tensor] = None            # f32 [co] (all additive per-channel terms pre-summed)
    act: bool = False                              # fused leaky-ReLU(0.2) * sqrt(2) in the conv epilogue
    noise_w: float = 0.0
    extra: dict = field(default_factory=dict)


def _dev(t, device, dtype=torch.float32):
    return None if t is None else t.detach().to(device=device, dtype=dtype).contiguous()


def make_spec(w, device, scale=1.0, bias=None, act=False, stride=1, pad=0, off=None, out_hw=None, tc=True,
              noise_w=0.0):
    """w [co,ci,kh,kw] (already folded if needed)."""
    co, ci, kh, kw = w.shape
    w = w.detach().to(torch.float32) * scale
    packed = pack_taps(w).unsqueeze(0)
    if off is None:
        off = (-pad, -pad)
    if out_hw is None:
        def out_hw(h, wd, _k=(kh, kw), _s=stride, _o=off):
            # all taps of the last output stay inside [off, in + |off|): symmetric-pad formula
            return ((h + 2 * (-_o[0]) - _k[0]) // _s + 1, (wd + 2 * (-_o[1]) - _k[1]) // _s + 1)
    s = ConvSpec(cin=ci, cout=co, kh=kh, kw=kw, stride=stride, off=off, out_hw=out_hw, act=act, noise_w=noise_w)
    s.w32 = _dev(packed, device)
    s.w16 = _dev(packed, device, torch.bfloat16) if (tc and ci % 32 == 0 and co % 32 == 0) else None
    s.bias = _dev(bias, device)
    return s


def make_upscale_spec(w, device, bias, act, noise_w=0.0):
    co, ci = w.shape[:2]
    packed = pack_upscale_phases(w.detach().to(torch.float32))
    s = ConvSpec(cin=ci, cout=co, kh=2, kw=2, stride=1, off=(-1, -1), phases=(2, 2), phase_off=(1, 1),
                 out_hw=lambda h, wd: (h, wd), act=act, noise_w=noise_w)
    s.w32 = _dev(packed, device)
    s.w16 = _dev(packed, device, torch.bfloat16) if (ci % 32 == 0 and co % 32 == 0) else None
    s.bias = _dev(bias, device)
    return s
