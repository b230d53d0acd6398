"""Training-side kernels through the C ABI on a B200: tcgen05 convolution forward / input gradient /
weight gradient, their second derivatives, the thin input layer, and the discriminator with its stage-1 losses and the
R1 penalty -- against torch fp32 math of the same op, the CPU oracle and the reference-run fixtures
(tests/golden/discriminator.npz).  Tolerances: bf16 operands with fp32 accumulation, rel-L2 <= 1e-2 per op (north_star's
bf16 per-layer bound), looser only where a whole network's drift accumulates (stated at the assertion)."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from golden_util import sample_idx

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _bf(t):
    return t.to(torch.bfloat16).float()


def _nchw(x_nhwc):
    return x_nhwc.float().permute(0, 3, 1, 2).contiguous()


GEOMS = [  # (k, stride, pad, cin, cout, n, h, w)
    (3, 1, 1, 32, 64, 2, 24, 40),
    (3, 1, 1, 64, 128, 3, 16, 16),
    (1, 1, 0, 128, 32, 2, 20, 12),
    (3, 1, 0, 64, 64, 2, 18, 34),          # valid conv (the encoders' reflect-padded inputs)
    (6, 2, 2, 64, 128, 2, 32, 48),         # Blur[1,3,3,1] folded into conv2 of a discriminator ResBlock
    (4, 2, 1, 64, 256, 2, 16, 32),         # Blur folded into the 1x1 stride-2 skip
    (3, 2, 0, 64, 128, 2, 34, 50),         # the unfolded stride-2 convolutions behind ppst_blur_nhwc
    (1, 2, 0, 64, 128, 2, 32, 48),
    (3, 1, 1, 256, 256, 1, 64, 64),
    (3, 1, 1, 512, 512, 2, 8, 8),          # tiny maps (the 4x4 / 8x8 end of the discriminator)
    (3, 1, 1, 512, 512, 2, 4, 4),
]


@pytest.mark.parametrize("k,s,p,ci,co,n,h,w", GEOMS)
def test_conv_forward_dgrad_wgrad_match_fp32(k, s, p, ci, co, n, h, w):
    from ppst_b200 import train
    g = torch.Generator().manual_seed(k * 1000 + ci + h)
    x = _bf(torch.randn(n, h, w, ci, generator=g)).cuda()
    wt = _bf(torch.randn(co, ci, k, k, generator=g) / math.sqrt(ci * k * k)).cuda()
    geom = train.Geom(k, s, p)
    oh, ow = geom.out_hw(h, w)
    gy = _bf(torch.randn(n, oh, ow, co, generator=g)).cuda()
    xr = _nchw(x).requires_grad_(True)
    wr = wt.clone().requires_grad_(True)
    with torch.backends.cudnn.flags(allow_tf32=False):
        yr = F.conv2d(xr, wr, stride=s, padding=p)
        yr.backward(_nchw(gy))
    xb = x.to(torch.bfloat16).requires_grad_(True)
    wb = wt.clone().requires_grad_(True)
    y = train.conv2d(xb, wb, s, p)
    assert y.shape == (n, oh, ow, co) and y.dtype == torch.bfloat16
    y.backward(gy.to(torch.bfloat16))
    assert rel_l2(_nchw(y), yr) <= 5e-3                       # bf16 rounding of the stored output
    assert rel_l2(_nchw(xb.grad), xr.grad) <= 5e-3
    assert rel_l2(wb.grad, wr.grad) <= 2e-3                   # fp32 output: only the accumulation order differs


def test_conv_bias_act_fused_and_bias_gradient():
    from ppst_b200 import train
    g = torch.Generator().manual_seed(5)
    n, h, w, ci, co = 2, 32, 32, 64, 128
    x = _bf(torch.randn(n, h, w, ci, generator=g)).cuda()
    wt = _bf(torch.randn(co, ci, 3, 3, generator=g) / math.sqrt(ci * 9)).cuda()
    b = (0.3 * torch.randn(co, generator=g)).cuda()
    gy = _bf(torch.randn(n, h, w, co, generator=g)).cuda()
    xr, wr, br = _nchw(x).requires_grad_(True), wt.clone().requires_grad_(True), b.clone().requires_grad_(True)
    with torch.backends.cudnn.flags(allow_tf32=False):
        yr = F.leaky_relu(F.conv2d(xr, wr, padding=1) + br.view(1, -1, 1, 1), 0.2) * math.sqrt(2)
        yr.backward(_nchw(gy))
    xb, wb, bb = x.to(torch.bfloat16).requires_grad_(True), wt.clone().requires_grad_(True), b.clone().requires_grad_(True)
    y = train.conv2d(xb, wb, 1, 1, bias=bb, act=True)
    y.backward(gy.to(torch.bfloat16))
    assert rel_l2(_nchw(y), yr) <= 5e-3
    assert rel_l2(_nchw(xb.grad), xr.grad) <= 1e-2
    assert rel_l2(wb.grad, wr.grad) <= 1e-2                   # the bf16-rounded masked gradient feeds the GEMM
    assert rel_l2(bb.grad, br.grad) <= 1e-2


def test_wgrad_adjoint_identity_at_full_size():
    """<dw, R> == <dy, conv(x, R)> for a random probe R, at the 512^2 x 128-channel size of the generator's last block
    (a size the CPU oracle cannot reach in seconds); both sides on the tensor cores, so the check is of the index
    algebra (taps, padding, K split), with the tolerance of one bf16 rounding of conv(x, R)."""
    from ppst_b200 import train
    g = torch.Generator().manual_seed(6)
    n, h, w, c = 2, 512, 512, 128
    x = torch.randn(n, h, w, c, generator=g).to(torch.bfloat16).cuda()
    gy = torch.randn(n, h, w, c, generator=g).to(torch.bfloat16).cuda()
    R = _bf(torch.randn(c, c, 3, 3, generator=g) / math.sqrt(c * 9)).cuda()
    geom = train.Geom(3, 1, 1)
    dw = train._wgrad_raw(x, gy, geom, R.shape)
    lhs = float((dw.double() * R.double()).sum())
    y = train._conv_raw(x, R, geom)
    rhs = float((y.double() * gy.double()).sum())
    # both sides are sums of N products of magnitude sy * sg: |sum| ~ sqrt(N) * sy * sg =: scale; rounding y to bf16
    # perturbs the right side by ~2^-9 * scale
    scale = math.sqrt(float((y.double() ** 2).sum()) * float((gy.double() ** 2).sum()) / y.numel())
    assert abs(lhs - rhs) <= 2e-2 * scale, (lhs, rhs, scale)
    # determinism: bit-identical on a second run
    assert torch.equal(dw, train._wgrad_raw(x, gy, geom, R.shape))


def test_second_derivative_gradient_penalty():
    """R1-style penalty |d out / d x|^2 through conv -> lrelu -> stride-2 conv: parameter gradients via double backward."""
    from ppst_b200 import train
    g = torch.Generator().manual_seed(7)
    n, h, w, c0, c1, c2 = 2, 16, 16, 32, 64, 64
    x = _bf(torch.randn(n, h, w, c0, generator=g)).cuda()
    w1 = _bf(torch.randn(c1, c0, 3, 3, generator=g) / math.sqrt(c0 * 9)).cuda()
    b1 = (0.2 * torch.randn(c1, generator=g)).cuda()
    w2 = _bf(torch.randn(c2, c1, 4, 4, generator=g) / math.sqrt(c1 * 16)).cuda()
    # reference
    xr = _nchw(x).requires_grad_(True)
    w1r, w2r = w1.clone().requires_grad_(True), w2.clone().requires_grad_(True)
    with torch.backends.cudnn.flags(allow_tf32=False):
        out = F.conv2d(F.leaky_relu(F.conv2d(xr, w1r, padding=1) + b1.view(1, -1, 1, 1), 0.2) * math.sqrt(2), w2r, stride=2, padding=1)
        gx, = torch.autograd.grad(out.sum(), [xr], create_graph=True)
        gx.pow(2).sum().backward()
    xb = x.to(torch.bfloat16).requires_grad_(True)
    w1b, w2b = w1.clone().requires_grad_(True), w2.clone().requires_grad_(True)
    o = train.conv2d(train.conv2d(xb, w1b, 1, 1, bias=b1, act=True), w2b, 2, 1)
    gxb, = torch.autograd.grad(o.float().sum(), [xb], create_graph=True)
    assert rel_l2(_nchw(gxb), gx) <= 1e-2
    gxb.float().pow(2).sum().backward()
    assert rel_l2(w1b.grad, w1r.grad) <= 2e-2
    assert rel_l2(w2b.grad, w2r.grad) <= 2e-2


def test_thin_input_layer_all_gradients():
    from ppst_b200 import train
    g = torch.Generator().manual_seed(8)
    n, h, w, co = 2, 40, 24, 64
    img = (torch.rand(n, 3, h, w, generator=g) * 2 - 1).cuda()
    wt = torch.randn(co, 3, 1, 1, generator=g).cuda()
    b = (0.2 * torch.randn(co, generator=g)).cuda()
    gy = _bf(torch.randn(n, h, w, co, generator=g)).cuda()
    ir, wr, br = img.clone().requires_grad_(True), wt.clone().requires_grad_(True), b.clone().requires_grad_(True)
    with torch.backends.cudnn.flags(allow_tf32=False):       # TF32 flips the leaky-ReLU branch of near-zero activations
        yr = F.leaky_relu(F.conv2d(ir, wr) + br.view(1, -1, 1, 1), 0.2) * math.sqrt(2)
        yr.backward(_nchw(gy))
    ib, wb, bb = img.clone().requires_grad_(True), wt.clone().requires_grad_(True), b.clone().requires_grad_(True)
    y = train.conv1x1_from_image(ib, wb, bb, act=True)
    y.backward(gy.to(torch.bfloat16))
    assert rel_l2(_nchw(y), yr) <= 5e-3
    assert rel_l2(ib.grad, ir.grad) <= 1e-2
    assert rel_l2(wb.grad, wr.grad) <= 1e-2
    assert rel_l2(bb.grad, br.grad) <= 1e-2


# ---------------------------------------------------------------------------------------------- discriminator
D_KEYS = ["convs.0.Conv.weight", "convs.0.Act.bias", "convs.1.conv1.Conv.weight", "convs.1.conv2.Conv.weight",
          "convs.1.conv2.Act.bias", "convs.1.skip.Conv.weight", "convs.3.conv1.Conv.weight", "convs.6.conv2.Conv.weight",
          "final_conv.Conv.weight",
          "final_linear.0.weight", "final_linear.0.bias", "final_linear.1.weight"]


def _sample_rel(store, key, t):
    """rel-L2 between the stored 4096-point subsample of the reference's tensor and the same points of t."""
    flat = t.detach().float().cpu().reshape(-1)
    got = flat[sample_idx(flat.numel(), key)].numpy()
    want = store[key + "/sample"]
    return float(np.linalg.norm(got - want) / max(np.linalg.norm(want), 1e-30))


# Whole-network gradients in bf16 differ from the fp32 reference mostly through leaky-ReLU BRANCH FLIPS: a forward
# activation that drifted by ~0.5 % (bf16 storage, 1e-2 per layer allowed) changes sign where |v| is below that, and each
# flipped element changes its gradient multiplier by 0.8.  A fraction f of flipped elements costs ~sqrt(f) in rel-L2 per
# layer, accumulating towards the input; the least-squares loss adds the relative error of (pred - 1) and pred as a common
# factor.  Against the fp32 reference fixture the bound is therefore 0.15 (cosine >= 0.99); the tight check is
# test_discriminator_matches_bf16_emulation below, where the rounding points coincide; the per-operator bounds (no drift
# in the inputs) are the 1e-2 of the tests above.
GRAD_TOL_DEEP = 0.15


def _check_depth_bounds(errs, second_order=False):
    for k, r in errs.items():
        assert r <= GRAD_TOL_DEEP, (k, r, errs)


def _make_d(size=256):
    from oracle import ppst_oracle as O
    from ppst_b200 import train
    sd = O.make_discriminator(size, seed=0)
    D = train.Discriminator(size, 2.0).cuda()
    D.load_reference_state_dict(sd)
    return D, sd


def test_discriminator_matches_reference_fixture(golden):
    """Forward, stage-1 D losses and their parameter gradients against the reference run (gen_golden.py)."""
    from ppst_b200 import train
    g = golden("discriminator")
    D, _ = _make_d(256)
    real, fake = torch.from_numpy(g["real"]).cuda(), torch.from_numpy(g["fake"]).cuda()
    pred = D(real)
    assert pred.shape == (2, 1)
    assert np.abs(pred.detach().cpu().numpy() - g["pred_real"]).max() <= 2e-2 * max(1.0, np.abs(g["pred_real"]).max())
    losses = train.d_losses(D, real, fake)
    assert abs(losses["D_real"].item() - float(g["D_real"])) <= 2e-2 * float(g["D_real"])
    (losses["D_real"] + losses["D_rec"]).backward()
    errs = {k: _sample_rel(g, "dstep_grad/" + k, D.p(k).grad) for k in D_KEYS}
    print("D-step gradient rel-L2 vs the reference:", {k: round(v, 4) for k, v in errs.items()})
    _check_depth_bounds(errs)


def test_discriminator_r1_penalty_matches_reference_fixture(golden):
    """R1 (ppst_model.py:139-156): input gradient, penalty and its double-backward parameter gradients."""
    from ppst_b200 import train
    g = golden("discriminator")
    D, _ = _make_d(256)
    real = torch.from_numpy(g["real"]).cuda().requires_grad_(True)
    pred = D(real).sum()
    grad_real, = torch.autograd.grad(outputs=pred, inputs=[real], create_graph=True, retain_graph=True)
    r_in = _sample_rel(g, "r1/grad_real", grad_real)
    assert r_in <= GRAD_TOL_DEEP, r_in
    pen = grad_real.pow(2).sum(dim=(1, 2, 3)) * 5.0
    assert np.abs(pen.detach().cpu().numpy() - g["r1/penalty"]).max() <= 5e-2 * np.abs(g["r1/penalty"]).max()
    pen.mean().backward()
    errs = {k: _sample_rel(g, "r1_grad/" + k, D.p(k).grad) for k in D_KEYS if not k.endswith("bias")}
    print("R1 gradient rel-L2 vs the reference: input grad %.4f" % r_in, {k: round(v, 4) for k, v in errs.items()})
    _check_depth_bounds(errs, second_order=True)


def test_discriminator_state_dict_round_trip_and_determinism():
    D, sd = _make_d(256)
    out = D.reference_state_dict()
    assert set(out) == set(sd)
    for k, v in out.items():
        assert torch.allclose(v.cpu(), sd[k], atol=1e-7) if "Blur" in k else torch.equal(v.cpu(), sd[k]), k
    x = (torch.rand(2, 3, 256, 256, generator=torch.Generator().manual_seed(1)) * 2 - 1).cuda()
    with torch.no_grad():
        a, b = D(x), D(x)
    assert torch.equal(a, b)


def test_discriminator_matches_bf16_emulation():
    """Same network, same rounding points (oracle.discriminator_bf16_emulation, CPU fp32 arithmetic with bf16 storage):
    the leaky-ReLU branches coincide, so forward, D-step gradients, the R1 input gradient and the R1 double-backward
    two bf16 executions with different accumulation orders still round some values to different neighbours, drift apart
    by ~1e-3 per layer and flip some leaky-ReLU branches; with a batch of 2 and 512 hidden units at the head ONE flipped
    unit is already 3 % of that layer's gradient.  Measured 3e-2 .. 8e-2; bound = the whole-network bound above (cosine
    >= 0.99).  The forward prediction agrees to 1e-2."""
    from oracle import ppst_oracle as O
    from ppst_b200 import train
    size = 256
    D, sd = _make_d(size)
    g = torch.Generator().manual_seed(123)
    real = torch.rand(4, 3, size, size, generator=g) * 2 - 1          # 4 + 4 samples: less head-flip noise than the fixture's 2 + 2
    fake = torch.rand(4, 3, size, size, generator=g) * 2 - 1
    keys = ["stylegan2_D." + k for k in D_KEYS]

    def leaf_sd():
        return {k: (v.clone().requires_grad_(True) if "Blur" not in k else v) for k, v in sd.items()}

    # D step
    e = leaf_sd()
    pr, pf = O.discriminator_bf16_emulation(e, real), O.discriminator_bf16_emulation(e, fake)
    (O.gan_loss(pr, True) + 0.5 * O.gan_loss(pf, False)).backward()
    pred = D(real.cuda())
    assert (pred.detach().cpu() - pr.detach()).abs().max() <= 1e-2 * max(1.0, float(pr.abs().max()))
    losses = train.d_losses(D, real.cuda(), fake.cuda())
    (losses["D_real"] + losses["D_rec"]).backward()
    errs = {k: rel_l2(D.p(k[len("stylegan2_D."):]).grad, e[k].grad) for k in keys}
    print("D-step gradient rel-L2 vs the bf16 emulation:", {k[12:]: round(v, 4) for k, v in errs.items()})
    assert max(errs.values()) <= GRAD_TOL_DEEP, errs
    # R1
    e = leaf_sd()
    r = real.clone().requires_grad_(True)
    gr, = torch.autograd.grad(O.discriminator_bf16_emulation(e, r).sum(), [r], create_graph=True)
    (gr.pow(2).sum(dim=(1, 2, 3)) * 5.0).mean().backward()
    D.zero_grad()
    rc = real.cuda().requires_grad_(True)
    gc, = torch.autograd.grad(D(rc).sum(), [rc], create_graph=True)
    assert rel_l2(gc, gr) <= GRAD_TOL_DEEP
    (gc.pow(2).sum(dim=(1, 2, 3)) * 5.0).mean().backward()
    errs = {k: rel_l2(D.p(k[len("stylegan2_D."):]).grad, e[k].grad) for k in keys if not k.endswith("bias")}
    print("R1 gradient rel-L2 vs the bf16 emulation:", {k[12:]: round(v, 4) for k, v in errs.items()})
    assert max(errs.values()) <= GRAD_TOL_DEEP, errs


# ---------------------------------------------------------------------------------------------- generator side (stage 1)
def test_norm_mod_affine_reflect_pad_gradients():
    """InstanceNorm + StyleMod (+ leaky-ReLU), the per-(n,c) affine and the reflection pad against torch autograd."""
    from ppst_b200 import train
    g = torch.Generator().manual_seed(11)
    n, h, w, c = 2, 24, 40, 64
    x = _bf(torch.randn(n, h, w, c, generator=g) * 1.5 + 0.3).cuda()
    sc, sh = (0.3 * torch.randn(n, c, generator=g)).cuda(), (0.3 * torch.randn(n, c, generator=g)).cuda()
    gy = _bf(torch.randn(n, h, w, c, generator=g)).cuda()
    for act in (False, True):
        xr, scr, shr = x.clone().requires_grad_(True), sc.clone().requires_grad_(True), sh.clone().requires_grad_(True)
        mu = xr.mean(dim=(1, 2), keepdim=True)
        var = ((xr - mu) ** 2).mean(dim=(1, 2), keepdim=True)
        yr = (xr - mu) * torch.rsqrt(var + 1e-5) * (1 + scr[:, None, None, :]) + shr[:, None, None, :]
        if act:
            yr = F.leaky_relu(yr, 0.2) * math.sqrt(2)
        yr.backward(gy)
        xb, scb, shb = x.to(torch.bfloat16).requires_grad_(True), sc.clone().requires_grad_(True), sh.clone().requires_grad_(True)
        y = train.instance_norm_mod(xb, scb, shb, act=act)
        y.backward(gy.to(torch.bfloat16))
        assert rel_l2(y, yr) <= 5e-3
        assert rel_l2(xb.grad, xr.grad) <= 1e-2, (act, rel_l2(xb.grad, xr.grad))
        assert rel_l2(scb.grad, scr.grad) <= 1e-2 and rel_l2(shb.grad, shr.grad) <= 1e-2
    # affine
    xr, ar, br = x.clone().requires_grad_(True), sc.clone().requires_grad_(True), sh.clone().requires_grad_(True)
    (xr * ar[:, None, None, :] + br[:, None, None, :]).backward(gy)
    xb, ab, bb = x.to(torch.bfloat16).requires_grad_(True), sc.clone().requires_grad_(True), sh.clone().requires_grad_(True)
    train.channel_affine(xb, ab, bb).backward(gy.to(torch.bfloat16))
    assert rel_l2(xb.grad, xr.grad) <= 5e-3 and rel_l2(ab.grad, ar.grad) <= 5e-3 and rel_l2(bb.grad, br.grad) <= 5e-3
    # reflection pad, asymmetric; also an image smaller than twice the pad (a pixel mirrored from both sides)
    for pad, xs in (((2, 1, 2, 2), x), ((2, 2, 2, 2), x[:, :3, :4].contiguous())):
        xr = xs.clone().requires_grad_(True)
        yr = F.pad(xr.permute(0, 3, 1, 2), (pad[2], pad[3], pad[0], pad[1]), mode="reflect").permute(0, 2, 3, 1)
        gp = _bf(torch.randn(yr.shape, generator=g)).cuda()
        yr.backward(gp)
        xb = xs.to(torch.bfloat16).requires_grad_(True)
        y = train.reflect_pad(xb, pad)
        assert torch.equal(y.float(), yr.detach())
        y.backward(gp.to(torch.bfloat16))
        assert rel_l2(xb.grad, xr.grad) <= 5e-3, pad
    pad = (2, 1, 2, 2)
    xr = x.clone().requires_grad_(True)
    yr = F.pad(xr.permute(0, 3, 1, 2), (pad[2], pad[3], pad[0], pad[1]), mode="reflect").permute(0, 2, 3, 1)
    gp = _bf(torch.randn(yr.shape, generator=g)).cuda()
    yr.backward(gp)
    xb = x.to(torch.bfloat16).requires_grad_(True)
    y = train.reflect_pad(xb, pad)
    assert torch.equal(y.float(), yr.detach())
    y.backward(gp.to(torch.bfloat16))
    assert rel_l2(xb.grad, xr.grad) <= 5e-3


def test_conv_transpose_fused_upscale_matches_torch():
    from ppst_b200 import train
    g = torch.Generator().manual_seed(12)
    n, h, w, ci, co = 2, 64, 64, 64, 128
    x = _bf(torch.randn(n, h, w, ci, generator=g)).cuda()
    wt = _bf(torch.randn(ci, co, 4, 4, generator=g) / math.sqrt(ci * 4)).cuda()
    b = (0.2 * torch.randn(co, generator=g)).cuda()
    nz = (0.3 * torch.randn(n, 2 * h, 2 * w, generator=g)).cuda()
    gy = _bf(torch.randn(n, 2 * h, 2 * w, co, generator=g)).cuda()
    xr, wr, br, nr = _nchw(x).requires_grad_(True), wt.clone().requires_grad_(True), b.clone().requires_grad_(True), nz.clone().requires_grad_(True)
    with torch.backends.cudnn.flags(allow_tf32=False):
        yr = F.conv_transpose2d(xr, wr, stride=2, padding=1) + br.view(1, -1, 1, 1) + nr[:, None]
        yr = F.leaky_relu(yr, 0.2) * math.sqrt(2)
        yr.backward(_nchw(gy))
    xb, wb, bb, nb = x.to(torch.bfloat16).requires_grad_(True), wt.clone().requires_grad_(True), b.clone().requires_grad_(True), nz.clone().requires_grad_(True)
    y = train.conv_transpose2d(xb, wb, 2, 1, (2 * h, 2 * w), bias=bb, act=True, noise=nb)
    y.backward(gy.to(torch.bfloat16))
    assert rel_l2(_nchw(y), yr) <= 5e-3
    assert rel_l2(_nchw(xb.grad), xr.grad) <= 1e-2
    assert rel_l2(wb.grad, wr.grad) <= 1e-2 and rel_l2(bb.grad, br.grad) <= 1e-2 and rel_l2(nb.grad, nr.grad) <= 1e-2


def test_stage1_networks_forward_and_l1_gradients_match_oracle():
    """E1 / E2 / G of training stage 1 at 128x128 against the CPU oracle: forward (bf16 drift bounds of the inference
    tests) and the gradients of the L1 reconstruction loss w.r.t. parameters of all three networks (torch.autograd through
    the oracle's fp32 expression; whole-network bound: cosine >= 0.99, see GRAD_TOL_DEEP)."""
    from oracle import ppst_oracle as O
    from ppst_b200 import train_nets as TN
    sdE1, sdE2, sdG = O.make_networks(0)
    size, n = 128, 2
    real = O.make_portraits(n, size, size, seed=21, smooth=True)
    noises = O.make_noises(n, size, size, seed=5)
    keys = {"E1": ["FromRGB.Conv.weight", "DownToSpatialCode.ResBlockDownBy2.conv2.Conv.weight", "ToSpatialCode.1.Conv.weight"],
            "E2": ["DownToGlobalCode1.ResBlockDownBy1.conv1.Conv.weight", "projector1.3.weight", "conv1x1_2.weight"],
            "G": ["HeadResnetBlock0.conv1.conv.weight", "HeadResnetBlock2.skip.Conv.weight",
                  "UpsamplingResBlock32.conv1.conv.weight", "UpsamplingResBlock64.conv2.epi1.style_mod.lin.weight",
                  "UpsamplingResBlock64.conv2.noise.weight", "ToRGB.conv.weight", "SpatialCodeModulation.scale.weight"]}
    ref = {}
    for name, sd in (("E1", sdE1), ("E2", sdE2), ("G", sdG)):
        ref[name] = {k: (v.clone().requires_grad_(True) if (v.is_floating_point() and not k.endswith("kernel")) else v)
                     for k, v in sd.items()}
    sp_r = O.encoder_con(ref["E1"], real)
    gl_r = O.encoder_col(ref["E2"], real)[0]
    rec_r = O.generator(ref["G"], sp_r, gl_r, noises=noises)
    (rec_r - real).abs().mean().backward()
    E1, E2, G = TN.ParamSet(sdE1).cuda(), TN.ParamSet(sdE2).cuda(), TN.ParamSet(sdG).cuda()
    rc = real.cuda()
    sp = TN.encoder_con(E1, rc)
    gl = TN.encoder_col(E2, rc)
    rec = TN.generator(G, sp, gl, noises=noises)
    assert rel_l2(sp.float().permute(0, 3, 1, 2), sp_r) <= 2e-2
    for a, b in zip(gl, gl_r):
        assert rel_l2(a, b) <= 2e-2
    peak = float(rec_r.max() - rec_r.min())
    psnr = 10 * math.log10(peak * peak / float(((rec.detach().cpu() - rec_r.detach()) ** 2).mean()))
    assert psnr >= 40.0, psnr
    (rec - rc).abs().mean().backward()
    errs = {}
    for name, P in (("E1", E1), ("E2", E2), ("G", G)):
        for k in keys[name]:
            errs[name + "." + k] = rel_l2(P[k].grad, ref[name][k].grad)
    print("stage-1 L1-loss gradient rel-L2 vs the oracle:", {k: round(v, 4) for k, v in errs.items()}, "PSNR %.1f" % psnr)
    # ~40 bf16 layers with leaky-ReLU between the loss and E1's first layer, 16x16 .. 128x128 maps, batch 2: a forward
    # drift of 1 % flips ~0.8 % of the branches of a layer, i.e. ~0.1 of that layer's gradient in rel-L2 (tools/train_diag.py
    # decomposes it per network: 0.05-0.11 for E1 / E2 alone, 0.02 -> 0.2 from ToRGB down to the spatial code for G alone).
    # Bound: cosine >= 0.95 for every probed parameter, >= 0.99 next to the loss.
    assert max(errs.values()) <= 0.30, errs
    assert errs["G.ToRGB.conv.weight"] <= GRAD_TOL_DEEP and errs["G.UpsamplingResBlock64.conv2.epi1.style_mod.lin.weight"] <= GRAD_TOL_DEEP


def test_linear_forward_backward_and_second_order():
    """ppst_linear / _dgrad / _wgrad closure against torch (f32): first derivatives and a gradient-penalty second derivative."""
    from ppst_b200 import train
    g = torch.Generator().manual_seed(13)
    for (m, k, n) in [(8, 2048, 512), (3, 512, 1), (16, 260, 2048)]:
        x = torch.randn(m, k, generator=g).cuda()
        w = (torch.randn(n, k, generator=g) / math.sqrt(k)).cuda()
        gy = torch.randn(m, n, generator=g).cuda()
        xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
        yr = xr @ wr.t()
        yr.backward(gy)
        xb, wb = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
        y = train.linear(xb, wb)
        y.backward(gy)
        assert rel_l2(y, yr) <= 1e-5 and rel_l2(xb.grad, xr.grad) <= 1e-5 and rel_l2(wb.grad, wr.grad) <= 1e-5, (m, k, n)
    m, k, n = 4, 64, 32
    x = torch.randn(m, k, generator=g).cuda()
    w = (torch.randn(n, k, generator=g) / math.sqrt(k)).cuda()
    xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    gx, = torch.autograd.grad((torch.tanh(xr @ wr.t())).sum(), [xr], create_graph=True)
    gx.pow(2).sum().backward()
    xb, wb = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    gxb, = torch.autograd.grad((torch.tanh(train.linear(xb, wb))).sum(), [xb], create_graph=True)
    gxb.pow(2).sum().backward()
    assert rel_l2(gxb, gx) <= 1e-5 and rel_l2(wb.grad, wr.grad) <= 1e-4 and rel_l2(xb.grad, xr.grad) <= 1e-4


def test_stage1_graph_replay_trains_like_eager_steps():
    """train.capture_step: a generator step (E1, E2, G, D, L1 + GAN losses, Adam) replayed from a CUDA graph moves the
    parameters like the same number of eager steps, and the loss goes down."""
    from oracle import ppst_oracle as O
    from ppst_b200 import train, train_nets as TN
    sdE1, sdE2, sdG = O.make_networks(0)
    size, n = 256, 2
    real = O.make_portraits(n, size, size, seed=31, smooth=True).cuda()

    def build():
        D = train.Discriminator(size, 2.0)
        D.load_reference_state_dict(O.make_discriminator(size, seed=0))
        m = TN.Stage1(sdE1, sdE2, sdG, D, noise_seed=None).cuda()
        for p in D.parameters():
            p.requires_grad_(False)
        opt = torch.optim.Adam(m.g_params(), lr=1e-3, betas=(0.0, 0.99), fused=True, capturable=True)
        return m, opt

    losses_seen = []

    def make_body(m, opt, log=None):
        def body():
            losses, _ = m.generator_losses(real)
            total = sum(losses.values())
            total.backward()
            opt.step()
            if log is not None:
                log.append(total.detach())
        return body

    ma, oa = build()
    body_a = make_body(ma, oa, losses_seen)
    for _ in range(5):
        oa.zero_grad(set_to_none=True)
        body_a()
    vals = [float(v) for v in losses_seen]
    assert vals[-1] < vals[0], vals
    # 2 steps eagerly vs 1 warm-up step + 1 replay: same kernels in the same order, no atomics anywhere in the step
    # (the bilinear skip's adjoint is a gather), so the parameters must agree bit for bit
    ma, oa = build()
    body_a = make_body(ma, oa)
    for _ in range(2):
        oa.zero_grad(set_to_none=True)
        body_a()
    mb, ob = build()
    graph = train.capture_step(make_body(mb, ob), lambda: ob.zero_grad(set_to_none=True), warmup=1)
    graph.replay()
    torch.cuda.synchronize()
    for name, pa, pb in (("G", ma.G, mb.G), ("E1", ma.E1, mb.E1), ("E2", ma.E2, mb.E2)):
        for k in pa.P:
            assert torch.equal(pa.P[k], pb.P[k]), (name, k)
    assert not torch.equal(ma.G["ToRGB.conv.weight"].detach().cpu(), sdG["ToRGB.conv.weight"])     # and they did move


def test_stage1_iterations_are_bit_deterministic():
    """Two identical runs of G step + D step + R1 step (fresh NoiseInjection draws from the same seed) end with identical
    parameters in all four networks."""
    from oracle import ppst_oracle as O
    from ppst_b200 import train, train_nets as TN
    sdE1, sdE2, sdG = O.make_networks(0)
    size = 256
    real = O.make_portraits(2, size, size, seed=33, smooth=True).cuda()

    def run():
        D = train.Discriminator(size, 2.0)
        D.load_reference_state_dict(O.make_discriminator(size, seed=0))
        m = TN.Stage1(sdE1, sdE2, sdG, D, noise_seed=7).cuda()
        og = torch.optim.Adam(m.g_params(), lr=1e-3, betas=(0.0, 0.99), fused=True)
        od = torch.optim.Adam(D.parameters(), lr=1e-3, betas=(0.0, 0.99), fused=True)
        for _ in range(2):
            og.zero_grad(set_to_none=True)
            for p in D.parameters():
                p.requires_grad_(False)
            sum(m.generator_losses(real)[0].values()).backward()
            og.step()
            for p in D.parameters():
                p.requires_grad_(True)
            od.zero_grad(set_to_none=True)
            sum(m.discriminator_losses(real).values()).backward()
            od.step()
            od.zero_grad(set_to_none=True)
            train.r1_loss(D, real).mean().backward()
            od.step()
        return m, D
    (a, da), (b, db) = run(), run()
    for pa, pb in ((a.G, b.G), (a.E1, b.E1), (a.E2, b.E2), (da, db)):
        for k in pa.P:
            assert torch.equal(pa.P[k], pb.P[k]), k


def test_bilinear2x_forward_and_deterministic_adjoint():
    from ppst_b200 import train
    g = torch.Generator().manual_seed(14)
    for (n, h, w, c) in [(2, 16, 24, 64), (1, 1, 5, 8), (2, 7, 3, 32)]:
        x = _bf(torch.randn(n, h, w, c, generator=g)).cuda()
        gy = _bf(torch.randn(n, 2 * h, 2 * w, c, generator=g)).cuda()
        xr = _nchw(x).requires_grad_(True)
        yr = F.interpolate(xr, scale_factor=2, mode="bilinear", align_corners=False)
        yr.backward(_nchw(gy))
        xb = x.to(torch.bfloat16).requires_grad_(True)
        y = train.bilinear2x(xb)
        y.backward(gy.to(torch.bfloat16))
        assert rel_l2(_nchw(y), yr) <= 5e-3, (n, h, w, c)
        assert rel_l2(_nchw(xb.grad), xr.grad) <= 5e-3, (n, h, w, c, rel_l2(_nchw(xb.grad), xr.grad))
        g2 = xb.grad.clone()
        xb.grad = None
        train.bilinear2x(xb).backward(gy.to(torch.bfloat16))
        assert torch.equal(g2, xb.grad)
        # nearest x2 (Upscale2d) and its 2x2 box-sum adjoint
        xr2 = _nchw(x).requires_grad_(True)
        yr2 = F.interpolate(xr2, scale_factor=2, mode="nearest")
        yr2.backward(_nchw(gy))
        xb2 = x.to(torch.bfloat16).requires_grad_(True)
        y2 = train.nearest2x(xb2)
        y2.backward(gy.to(torch.bfloat16))
        assert torch.equal(_nchw(y2), yr2.detach()) and rel_l2(_nchw(xb2.grad), xr2.grad) <= 5e-3


def test_discriminator_512_and_stage1_256_forward_match_oracle():
    """The sizes the reference trains at: D at crop 512 (first layer 3 -> 64) and the stage-1 reconstruction at 256^2."""
    from oracle import ppst_oracle as O
    from ppst_b200 import train, train_nets as TN
    x = O.make_portraits(1, 512, 512, seed=41, smooth=True)
    sdD = O.make_discriminator(512, seed=1)
    D = train.Discriminator(512, 2.0).cuda()
    D.load_reference_state_dict(sdD)
    with torch.no_grad():
        want = O.discriminator(sdD, x)
        got = D(x.cuda()).cpu()
    assert (got - want).abs().max() <= 2e-2 * max(1.0, float(want.abs().max())), (got, want)
    sdE1, sdE2, sdG = O.make_networks(0)
    real = O.make_portraits(1, 256, 256, seed=42, smooth=True)
    noises = O.make_noises(1, 256, 256, seed=6)
    with torch.no_grad():
        rec_r = O.generator(sdG, O.encoder_con(sdE1, real), O.encoder_col(sdE2, real)[0], noises=noises)
        E1, E2, G = TN.ParamSet(sdE1).cuda(), TN.ParamSet(sdE2).cuda(), TN.ParamSet(sdG).cuda()
        rec = TN.generator(G, TN.encoder_con(E1, real.cuda()), TN.encoder_col(E2, real.cuda()), noises=noises).cpu()
    peak = float(rec_r.max() - rec_r.min())
    psnr = 10 * math.log10(peak * peak / float(((rec - rec_r) ** 2).mean()))
    assert psnr >= 45.0, psnr                      # north_star's bf16 output bound


def test_blur_nhwc_matches_upfirdn2d_and_is_self_adjoint_to_second_order():
    from oracle import ppst_oracle as O
    from ppst_b200 import train
    g = torch.Generator().manual_seed(15)
    for taps, pad in (([1, 3, 3, 1], (2, 3)), ([1, 3, 3, 1], (1, 2)), ([1, 2, 1], (0, 0)), ([1, 2, 1], (1, 1))):
        k = O.make_kernel(taps)
        x = _bf(torch.randn(2, 18, 26, 32, generator=g))
        want = O.upfirdn2d(x.permute(0, 3, 1, 2), k, pad=pad)                  # NCHW oracle
        xb = x.cuda().to(torch.bfloat16).requires_grad_(True)
        y = train.blur(xb, k.cuda(), pad)
        assert rel_l2(_nchw(y), want) <= 5e-3, (taps, pad)
        gy = _bf(torch.randn(y.shape, generator=g)).cuda()
        xr = x.clone().requires_grad_(True)
        O.upfirdn2d(xr.permute(0, 3, 1, 2), k, pad=pad).backward(_nchw(gy).cpu())
        gx, = torch.autograd.grad(y, [xb], gy.to(torch.bfloat16), create_graph=True)
        assert rel_l2(gx, xr.grad) <= 5e-3, (taps, pad)
        # second order: d/d(gy) of <gx, R> is the blur of R
        R = _bf(torch.randn(x.shape, generator=g)).cuda()
        gyb = gy.to(torch.bfloat16).requires_grad_(True)
        gx2, = torch.autograd.grad(train.blur(xb, k.cuda(), pad), [xb], gyb, create_graph=True)
        d, = torch.autograd.grad((gx2.float() * R).sum(), [gyb])
        assert rel_l2(_nchw(d), O.upfirdn2d(R.cpu().permute(0, 3, 1, 2), k, pad=pad)) <= 5e-3


def test_gap_gmp_heads_pooling_with_ties():
    from ppst_b200 import train
    g = torch.Generator().manual_seed(16)
    x = _bf(torch.randn(2, 20, 12, 64, generator=g))
    x[0, 3, 4, :8] = 9.0                         # a tie for the maximum: two pixels attain it in 8 channels
    x[0, 7, 1, :8] = 9.0
    x = x.cuda()
    xr = x.clone().requires_grad_(True)
    gap_r, gmp_r = xr.mean(dim=(1, 2)), xr.amax(dim=(1, 2))
    ga, gm = torch.randn(2, 64, generator=g).cuda(), torch.randn(2, 64, generator=g).cuda()
    ((gap_r * ga).sum() + (gmp_r * gm).sum()).backward()
    xb = x.to(torch.bfloat16).requires_grad_(True)
    gap, gmp = train.gap_gmp(xb)
    assert rel_l2(gap, gap_r) <= 1e-5 and torch.equal(gmp, gmp_r.detach())
    ((gap * ga).sum() + (gmp * gm).sum()).backward()
    assert rel_l2(xb.grad, xr.grad) <= 5e-3
