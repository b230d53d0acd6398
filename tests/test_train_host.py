"""Host-side logic of the training path on CPU (no kernels run): the sub-pixel phase decomposition of a stride-2 input
gradient, the weight-packing slot tables, output-size algebra, and the state_dict contract of the trainable modules."""
import itertools
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import ppst_oracle as O
from ppst_b200 import train, train_nets


def _dgrad_reference(gy, w, stride, pad, in_hw):
    """d(conv2d)/dx by autograd (NCHW, fp64)."""
    x = torch.zeros(gy.shape[0], w.shape[1], *in_hw, dtype=torch.float64, requires_grad=True)
    F.conv2d(x, w, stride=stride, padding=pad).backward(gy)
    return x.grad


@pytest.mark.parametrize("k,pad", [(6, 2), (4, 1), (3, 0), (1, 0), (5, 0), (3, 1)])
def test_stride2_input_gradient_is_four_phase_convolutions(k, pad):
    """The index algebra behind train._dgrad_raw (what ppst_conv_tc is asked to compute): input pixel 2a + r receives
    dy[a - u] * w[2u + r + pad]; per phase that is a plain J-tap correlation over dy with offset -u_hi(r)."""
    g = torch.Generator().manual_seed(k * 10 + pad)
    ci, co, h = 3, 2, 12
    oh = (h + 2 * pad - k) // 2 + 1
    w = torch.randn(co, ci, k, k, generator=g, dtype=torch.float64)
    gy = torch.randn(1, co, oh, oh, generator=g, dtype=torch.float64)
    want = _dgrad_reference(gy, w, 2, pad, (h, h))
    (hi0, lo0), (hi1, lo1) = train._phase_taps(k, pad, 0), train._phase_taps(k, pad, 1)
    J = max(hi0 - lo0, hi1 - lo1) + 1
    his = (hi0, hi1)
    got = torch.zeros_like(want)
    for ry, rx in itertools.product(range(2), range(2)):
        for a, b in itertools.product(range(h // 2), range(h // 2)):
            acc = torch.zeros(ci, dtype=torch.float64)
            for jy, jx in itertools.product(range(J), range(J)):
                ty, tx = 2 * (his[ry] - jy) + ry + pad, 2 * (his[rx] - jx) + rx + pad
                if not (0 <= ty < k and 0 <= tx < k):
                    continue                                        # slot packed as zero
                oy, ox = a + jy - his[ry], b + jx - his[rx]         # off = -u_hi(r), + one per tap
                if 0 <= oy < oh and 0 <= ox < oh:                   # outside: TMA zero fill
                    acc += gy[0, :, oy, ox] @ w[:, :, ty, tx]
            got[0, :, 2 * a + ry, 2 * b + rx] = acc
    assert torch.allclose(got, want, atol=1e-12)


def test_geom_output_sizes_match_torch():
    for k, s, p, h, w in [(3, 1, 1, 17, 9), (1, 1, 0, 5, 5), (6, 2, 2, 32, 48), (4, 2, 1, 16, 32), (3, 2, 0, 34, 50), (5, 2, 0, 260, 260)]:
        y = F.conv2d(torch.zeros(1, 1, h, w), torch.zeros(1, 1, k, k), stride=s, padding=p)
        assert train.Geom(k, s, p).out_hw(h, w) == tuple(y.shape[2:])


def test_discriminator_parameters_follow_the_reference_state_dict():
    for size in (256, 512):
        sd = O.make_discriminator(size, seed=3)
        D = train.Discriminator(size, 2.0)
        D.load_reference_state_dict(sd)
        out = D.reference_state_dict()
        assert set(out) == set(sd)                   # key-complete for the reference's loader, Blur.kernel buffers included
        for k in sd:
            if "Blur" in k:
                assert torch.allclose(out[k], sd[k], atol=1e-7), k
            else:
                assert torch.equal(out[k], sd[k]), k
    with pytest.raises(AssertionError):
        train.Discriminator(256, 2.0).load_reference_state_dict(O.make_discriminator(512, seed=3))


def test_paramset_round_trips_reference_keys_and_excludes_fir_buffers():
    sdE1, sdE2, sdG = O.make_networks(0)
    for sd in (sdE1, sdE2, sdG):
        P = train_nets.ParamSet(sd)
        back = P.reference_state_dict()
        assert set(back) == set(sd)
        for k, v in sd.items():
            assert torch.equal(back[k], v), k
            assert (k.replace(".", "/") in P.P) == (not k.endswith(".kernel"))
        assert all(p.requires_grad for p in P.parameters())
    assert sum(p.numel() for p in train_nets.ParamSet(sdG).parameters()) > 40e6      # G 42.5 M


def test_training_ops_refuse_cpu_tensors():
    x = torch.zeros(1, 8, 8, 32, dtype=torch.bfloat16)
    w = torch.zeros(32, 32, 3, 3)
    with pytest.raises(RuntimeError):
        train.conv2d(x, w, 1, 1)
    with pytest.raises(RuntimeError):
        train.instance_norm_mod(x)
