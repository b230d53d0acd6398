"""Host-side weight packing of the inference engine on CPU (no kernels run): the algebra that lets the engine skip whole
passes must hold exactly in real arithmetic -- Blur folded into the stride-2 convolutions, the fused-upscale transposed
convolution as four 2x2 phase convolutions, the tap-major packing.  Checked against torch CPU ops / the oracle."""
import itertools

import pytest
import torch
import torch.nn.functional as F

from oracle import ppst_oracle as O
from ppst_b200 import packing


def test_fold_fir_equals_blur_then_strided_convolution():
    """ConvLayer(downsample=True) of the encoders (stylegan2_layers.py:497-555): reflect pad (2,1) -> Blur([1,2,1]) -> 3x3 stride 2
    == ONE 5x5 stride-2 convolution on the reflect-padded input; skip: Blur(zero pad (1,0)) -> 1x1 stride 2 == ONE 3x3 stride-2
    convolution with zero padding 1."""
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 6, 16, 20, generator=g, dtype=torch.float64)
    k = O.make_kernel([1, 2, 1]).double()
    w3 = torch.randn(8, 6, 3, 3, generator=g, dtype=torch.float64)
    xp = F.pad(x, (2, 1, 2, 1), mode="reflect")
    want = F.conv2d(O.upfirdn2d(xp.float(), k.float(), pad=(0, 0)).double(), w3, stride=2)
    got = F.conv2d(xp, packing.fold_fir(w3, k), stride=2)
    assert got.shape == want.shape and (got - want).abs().max() <= 1e-5      # the oracle blur runs in fp32
    w1 = torch.randn(8, 6, 1, 1, generator=g, dtype=torch.float64)
    want = F.conv2d(O.upfirdn2d(x.float(), k.float(), pad=(1, 0)).double(), w1, stride=2)
    got = F.conv2d(x, packing.fold_fir(w1, k), stride=2, padding=1)
    assert got.shape == want.shape and (got - want).abs().max() <= 1e-5
    # discriminator flavour: Blur([1,3,3,1], pad (2,2)) -> 3x3 stride 2 == 6x6 stride 2, padding 2
    k4 = O.make_kernel([1, 3, 3, 1]).double()
    want = F.conv2d(O.upfirdn2d(x.float(), k4.float(), pad=(2, 2)).double(), w3, stride=2)
    got = F.conv2d(x, packing.fold_fir(w3, k4), stride=2, padding=2)
    assert got.shape == want.shape and (got - want).abs().max() <= 1e-5


def test_upscale_phases_equal_the_fused_transposed_convolution():
    """EqualizedConv2d fused upscale (stylegan2_layers.py:312-321): box-summed 4x4 kernel, conv_transpose2d(stride 2, padding
    1) == four plain 2x2 convolutions over the input, output pixel (2i+a, 2j+b) from phase (a, b) reading the window that
    starts at (i-1+a, j-1+b)."""
    g = torch.Generator().manual_seed(1)
    ci, co, h, w = 5, 7, 6, 9
    x = torch.randn(2, ci, h, w, generator=g, dtype=torch.float64)
    wt = torch.randn(co, ci, 3, 3, generator=g, dtype=torch.float64)
    wp = F.pad(wt.permute(1, 0, 2, 3), [1, 1, 1, 1])
    w4 = wp[:, :, 1:, 1:] + wp[:, :, :-1, 1:] + wp[:, :, 1:, :-1] + wp[:, :, :-1, :-1]
    ref = F.conv_transpose2d(x, w4, stride=2, padding=1)
    packed = packing.pack_upscale_phases(wt)                                  # [4][co][2*2*ci], tap-major, channel-minor
    out = torch.zeros_like(ref)
    xpad = F.pad(x, (1, 1, 1, 1))
    for a, b in itertools.product(range(2), range(2)):
        wk = packed[a * 2 + b].view(co, 2, 2, ci).permute(0, 3, 1, 2)          # [co, ci, 2, 2]
        # window starts at (i - 1 + a, j - 1 + b) -> in the 1-padded tensor at (i + a, j + b)
        full = F.conv2d(xpad, wk)                                              # [.., h + 1, w + 1]
        out[:, :, a::2, b::2] = full[:, :, a:a + h, b:b + w]
    assert (out - ref).abs().max() <= 1e-10


def test_pack_taps_is_tap_major_channel_minor():
    w = torch.arange(2 * 3 * 2 * 2, dtype=torch.float32).view(2, 3, 2, 2)
    p = packing.pack_taps(w)
    assert p.shape == (2, 12)
    for co, ty, tx, c in itertools.product(range(2), range(2), range(2), range(3)):
        assert p[co, (ty * 2 + tx) * 3 + c] == w[co, c, ty, tx]
