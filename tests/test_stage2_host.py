"""Host-side pieces of training stage 2 (ppst_b200/train_stage2.py) that are plain tensor code: checked on CPU against torch
and against the oracle's restatement of the reference (models/ppst_model.py:59-66,330-339; generator.py:12-26)."""
import torch
import torch.nn.functional as F

from oracle import ppst_oracle as O
from ppst_b200 import train_stage2 as T2


def test_swap_pairs_is_the_reference_minibatch_swap():
    x = torch.arange(6 * 3).reshape(6, 3).float()
    assert torch.equal(T2.swap_pairs(x), O._swap_pairs(x))
    assert torch.equal(T2.swap_pairs(x)[0], x[1]) and torch.equal(T2.swap_pairs(x)[5], x[4])


def test_pad_to_granularity():
    t = torch.randn(2, 5, 7)
    p = T2._pad_to(T2._pad_to(t, 2, 64), 1, 32)
    assert p.shape == (2, 32, 64) and torch.equal(p[:, :5, :7], t) and float(p[:, 5:].abs().max()) == 0.0
    assert T2._pad_to(t, 1, 5) is t


def test_replicate_pad_matches_replication_pad2d_with_gradient():
    x = torch.randn(2, 5, 6, 3, requires_grad=True)
    y = T2.replicate_pad(x, 1)
    ref = F.pad(x.permute(0, 3, 1, 2), (1, 1, 1, 1), mode="replicate").permute(0, 2, 3, 1)
    assert torch.equal(y, ref)
    g = torch.randn_like(y)
    (gx,) = torch.autograd.grad(y, x, g)
    (gr,) = torch.autograd.grad(ref, x, g)
    assert torch.allclose(gx, gr, atol=1e-6)


def test_prelu_function_matches_torch_prelu_and_its_gradients():
    x = torch.randn(3, 4, 5, 8, requires_grad=True)
    a = torch.tensor([0.25], requires_grad=True)
    y = T2.prelu(x, a)
    ref = F.prelu(x, a)
    assert torch.equal(y, ref)
    g = torch.randn_like(y)
    gx, ga = torch.autograd.grad(y, (x, a), g)
    rx, ra = torch.autograd.grad(ref, (x, a), g)
    assert torch.allclose(gx, rx) and torch.allclose(ga, ra, rtol=1e-5, atol=1e-6)


def test_rselfcorr_nhwc_matches_the_oracle():
    f = torch.randn(2, 64, 32, 32)                       # NCHW for the oracle (any H, W divisible by 4)
    ref = O.rselfcorr(F.interpolate(f, (256, 256)))      # the reference reshapes to 64 x 64 patches: needs 256 x 256
    got = T2.rselfcorr(F.interpolate(f, (256, 256)).permute(0, 2, 3, 1).contiguous())
    assert got.shape == (2, 64, 64, 256)
    assert torch.allclose(got.permute(0, 3, 1, 2), ref, atol=2e-6)


def test_corr_prep_matches_the_oracle_normalisation():
    f = torch.randn(2, 512, 8, 8)
    rows = f.reshape(2, 512, -1).permute(0, 2, 1)        # [b, L, C]
    got = T2._corr_prep(rows)
    h1 = f.reshape(2, 512, -1)
    h1 = torch.cat((h1[:, :256] - h1[:, :256].mean(dim=1, keepdim=True), h1[:, 256:]), dim=1)
    ref = (h1 / (h1.pow(2).sum(dim=1, keepdim=True).sqrt() + O._EPS)).permute(0, 2, 1)
    assert torch.allclose(got, ref, atol=1e-6)
