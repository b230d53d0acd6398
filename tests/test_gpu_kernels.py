"""Kernel-level parity on the GPU: every C-ABI entry point against the oracle / plain fp32 PyTorch math,
called through the ctypes binding (no fallbacks anywhere)."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import ppst_oracle as O  # noqa: E402

T = torch.from_numpy
DEV = "cuda"


def gen(seed):
    return torch.Generator().manual_seed(seed)


# ------------------------------------------------------------------------------------------------ plugin ops
def test_fused_leaky_relu_golden_bit_exact(golden):
    from ppst_b200.stylegan2_op import fused_leaky_relu
    g = golden("ops")
    for name in g["flrelu/names"]:
        slope, scale = (float(v) for v in g["flrelu/%s/cfg" % name])
        x = T(g["flrelu/%s/x" % name]).to(DEV).requires_grad_(True)
        b = T(g["flrelu/%s/b" % name]).to(DEV).requires_grad_(True)
        y = fused_leaky_relu(x, b, slope, scale)
        assert np.array_equal(y.detach().cpu().numpy(), g["flrelu/%s/y" % name]), name   # bit-exact in fp32
        go = T(g["flrelu/%s/go" % name]).to(DEV)
        y.backward(go)
        assert np.abs(x.grad.cpu().numpy() - g["flrelu/%s/gx" % name]).max() <= 1e-6
        assert np.abs(b.grad.cpu().numpy() - g["flrelu/%s/gb" % name]).max() <= 2e-4


def test_fused_leaky_relu_large_dtypes_and_module():
    from ppst_b200.stylegan2_op import FusedLeakyReLU, fused_leaky_relu
    x = torch.randn(2, 64, 128, 128, generator=gen(1))
    b = torch.randn(64, generator=gen(2))
    want = O.fused_leaky_relu(x, b)
    got = fused_leaky_relu(x.to(DEV), b.to(DEV))
    assert torch.equal(got.cpu(), want)
    for dtp, tol in ((torch.bfloat16, 2e-2), (torch.float16, 2e-3)):
        got = fused_leaky_relu(x.to(DEV, dtp), b.to(DEV, dtp)).float().cpu()
        ref = O.fused_leaky_relu(x.to(dtp).float(), b.to(dtp).float())
        assert (got - ref).abs().max() <= tol * ref.abs().max()
    # odd plane length (scalar path) and the [B, C] form
    x = torch.randn(3, 5, 33, 31, generator=gen(3))
    b = torch.randn(5, generator=gen(4))
    assert torch.equal(fused_leaky_relu(x.to(DEV), b.to(DEV)).cpu(), O.fused_leaky_relu(x, b))
    x = torch.randn(7, 2048, generator=gen(5))
    b = torch.randn(2048, generator=gen(6))
    assert torch.equal(fused_leaky_relu(x.to(DEV), b.to(DEV), 0.1, 1.5).cpu(), O.fused_leaky_relu(x, b, 0.1, 1.5))
    m = FusedLeakyReLU(5).to(DEV)
    assert list(m.state_dict()) == ["bias"] and float(m.bias.abs().sum()) == 0.0
    with pytest.raises(RuntimeError):
        fused_leaky_relu(torch.zeros(1, 2, 3, 3), torch.zeros(2))        # CPU tensors: no fallback


def test_fused_leaky_relu_double_backward():
    from ppst_b200.stylegan2_op import fused_leaky_relu
    x = torch.randn(2, 3, 9, 9, generator=gen(7)).to(DEV).requires_grad_(True)
    b = torch.randn(3, generator=gen(8)).to(DEV).requires_grad_(True)
    y = fused_leaky_relu(x, b)
    gx, = torch.autograd.grad(y.sum(), x, create_graph=True)
    # R1-style penalty: second order through the op (ppst_model.py:145-150 uses this on D)
    (gx.pow(2).sum()).backward()
    xr = x.detach().clone().requires_grad_(True)
    br = b.detach().clone().requires_grad_(True)
    yr = F.leaky_relu(xr + br.view(1, -1, 1, 1), 0.2) * math.sqrt(2)
    gxr, = torch.autograd.grad(yr.sum(), xr, create_graph=True)
    gxr.pow(2).sum().backward()
    assert torch.allclose(gx, gxr, atol=1e-6)
    # gate is piecewise constant: second derivative w.r.t. x is zero in both implementations
    assert x.grad is None or float(x.grad.abs().max()) == 0.0 or torch.allclose(x.grad, xr.grad, atol=1e-6)


def test_upfirdn2d_golden_all_modes(golden):
    from ppst_b200.stylegan2_op import upfirdn2d
    g = golden("ops")
    for name in g["upfirdn2d/names"]:
        up, down, p0, p1 = (int(v) for v in g["upfirdn2d/%s/cfg" % name])
        y = upfirdn2d(T(g["upfirdn2d/%s/x" % name]).to(DEV), T(g["upfirdn2d/%s/k" % name]).to(DEV), up, down, (p0, p1))
        want = g["upfirdn2d/%s/y" % name]
        assert tuple(y.shape) == want.shape, name
        assert np.abs(y.cpu().numpy() - want).max() <= 2e-6, name


def test_upfirdn2d_hot_path_sizes_and_dtypes():
    from ppst_b200.stylegan2_op import upfirdn2d
    k = O.make_kernel([1, 2, 1])
    for shape, pad in (((1, 32, 259, 259), (0, 0)), ((2, 16, 256, 256), (1, 0)), ((1, 8, 131, 67), (0, 0))):
        x = torch.randn(*shape, generator=gen(shape[2]))
        want = O.upfirdn2d(x, k, pad=pad)
        got = upfirdn2d(x.to(DEV), k.to(DEV), pad=pad).cpu()
        assert got.shape == want.shape and (got - want).abs().max() <= 1e-6
        got16 = upfirdn2d(x.to(DEV, torch.bfloat16), k.to(DEV), pad=pad).float().cpu()
        assert (got16 - O.upfirdn2d(x.bfloat16().float(), k, pad=pad)).abs().max() <= 2e-2
    k4 = O.make_kernel([1, 3, 3, 1])
    x = torch.randn(2, 8, 64, 64, generator=gen(11))
    for up, down, pad in ((1, 1, (2, 1)), (2, 1, (2, 1)), (1, 2, (1, 1))):
        kk = k4 * (up ** 2)
        want = O.upfirdn2d(x, kk, up, down, pad)
        got = upfirdn2d(x.to(DEV), kk.to(DEV), up, down, pad).cpu()
        assert got.shape == want.shape and (got - want).abs().max() <= 2e-6


def test_upfirdn2d_strip_path_edge_cases():
    """up = down = 1 fast path: cropping (negative) and oversized pads, non-square / short taps, rows long enough to be
    split into column chunks, one-row strips, and a base pointer that is not 16-byte aligned."""
    from ppst_b200.stylegan2_op import upfirdn2d
    cases = [
        ((1, 2, 37, 41), torch.randn(3, 3, generator=gen(1)), (-1, -2)),
        ((1, 2, 37, 41), torch.randn(2, 2, generator=gen(2)), (5, 6)),
        ((2, 1, 9, 1400), torch.randn(4, 3, generator=gen(3)), (2, 1)),     # out_w > 1280 -> two column chunks
        ((1, 3, 33, 2600), torch.randn(3, 3, generator=gen(4)), (0, 0)),
        ((3, 1, 3, 3), torch.randn(3, 3, generator=gen(5)), (0, 0)),       # single output element per plane
        ((1, 1, 70, 1027), torch.randn(1, 4, generator=gen(6)), (1, 2)),
    ]
    for shape, k, pad in cases:
        x = torch.randn(*shape, generator=gen(shape[3]))
        want = O.upfirdn2d(x, k, pad=pad)
        got = upfirdn2d(x.to(DEV), k.to(DEV), pad=pad).cpu()
        assert got.shape == want.shape, (shape, pad)
        assert (got - want).abs().max() <= 2e-5, (shape, pad)
        for dt in (torch.bfloat16, torch.float16):
            g16 = upfirdn2d(x.to(DEV, dt), k.to(DEV), pad=pad).float().cpu()
            w16 = O.upfirdn2d(x.to(dt).float(), k, pad=pad)
            assert (g16 - w16).abs().max() <= 2e-2 * max(1.0, float(w16.abs().max())), (shape, pad, dt)
    # contiguous view at an odd element offset into its storage
    k = O.make_kernel([1, 2, 1])
    for dt, tol in ((torch.float32, 1e-6), (torch.bfloat16, 2e-2)):
        flat = torch.randn(2 * 5 * 131 * 67 + 3, generator=gen(9)).to(DEV, dt)
        x = flat[3:].view(2, 5, 131, 67)
        assert x.is_contiguous() and x.data_ptr() % 16 != 0
        got = upfirdn2d(x, k.to(DEV)).float().cpu()
        assert (got - O.upfirdn2d(x.float().cpu(), k)).abs().max() <= tol
    # Inf in the input reaches exactly the outputs whose kh x kw window covers it (no 0 * Inf through padded taps)
    for kk, pad in ((torch.ones(2, 2), (0, -1)), (torch.ones(1, 3), (1, 1)), (torch.ones(3, 3), (0, 0)),
                    (torch.ones(4, 2), (2, 2))):
        x = torch.randn(1, 2, 40, 40, generator=gen(10))
        x[0, 0, -1, :] = float("inf")
        x[0, 1, 17, 23] = float("inf")
        want = O.upfirdn2d(x, kk, pad=pad)
        got = upfirdn2d(x.to(DEV), kk.to(DEV), pad=pad).cpu()
        fin = torch.isfinite(want)
        assert (torch.isfinite(got) == fin).all(), (tuple(kk.shape), pad)
        assert (got[fin] - want[fin]).abs().max() <= 1e-5


def test_upfirdn2d_16bit_pair_kernel_is_bit_identical_to_the_fp32_kernel():
    """bf16 / fp16 strips run two columns per thread on packed FFMA2 (upfirdn2d_bulk16_kernel); per column the fp32
    arithmetic and its order are those of the fp32 kernel, so y16 == round16(fp32 kernel on the same 16-bit values),
    exactly -- over odd / even widths, every tap shape up to 4x4, zero / cropping pads, ragged strips, tiny planes, Inf."""
    from ppst_b200.stylegan2_op import upfirdn2d
    cases = [((8, 4, 515, 515), (3, 3), (0, 0)), ((2, 8, 512, 512), (4, 4), (2, 2)), ((2, 3, 256, 256), (3, 3), (1, 0)),
             ((2, 3, 259, 259), (4, 4), (2, 1)), ((1, 2, 131, 130), (4, 4), (1, 1)), ((1, 2, 37, 41), (3, 3), (-1, -2)),
             ((1, 2, 37, 41), (2, 2), (5, 6)), ((1, 1, 70, 1027), (1, 4), (1, 2)), ((3, 1, 9, 9), (4, 3), (3, 3)),
             ((1, 2, 8, 8), (3, 3), (0, 0)), ((1, 2, 5, 12), (1, 1), (0, 0)), ((1, 1, 40, 64), (2, 4), (0, 3)),
             ((1, 2, 300, 24), (4, 4), (2, 2)), ((2, 2, 16, 2050), (3, 3), (1, 1))]
    for shape, ks, pad in cases:
        k = torch.randn(*ks, generator=gen(ks[0] * 7 + ks[1])).to(DEV)
        x = torch.randn(*shape, generator=gen(shape[2] + shape[3])).to(DEV)
        for dt in (torch.bfloat16, torch.float16):
            x16 = x.to(dt)
            got = upfirdn2d(x16, k, pad=pad)
            want = upfirdn2d(x16.float(), k, pad=pad).to(dt)
            assert got.dtype == dt and got.shape == want.shape
            assert torch.equal(got, want), (shape, ks, pad, dt, float((got.float() - want.float()).abs().max()))
    # a view at an odd element offset (2-byte aligned planes: 16-bit stores) and Inf next to the border
    k = torch.randn(4, 4, generator=gen(3)).to(DEV)
    flat = torch.randn(2 * 3 * 66 * 130 + 1, generator=gen(4)).to(DEV, torch.bfloat16)
    x = flat[1:].view(2, 3, 66, 130)
    x[0, 0, 0, 0] = float("inf")
    x[1, 2, -1, -2] = float("-inf")
    got = upfirdn2d(x, k.abs(), pad=(2, 2))
    want = upfirdn2d(x.float(), k.abs(), pad=(2, 2)).to(torch.bfloat16)
    assert torch.equal(torch.isnan(got), torch.isnan(want))
    ok = ~torch.isnan(want)
    assert torch.equal(got[ok], want[ok])


def test_upfirdn2d_backward_and_double_backward():
    from ppst_b200.stylegan2_op import upfirdn2d
    k = O.make_kernel([1, 3, 3, 1])
    for up, down, pad in ((1, 1, (2, 1)), (2, 1, (2, 1)), (1, 2, (1, 1))):
        x = torch.randn(1, 2, 12, 12, generator=gen(up * 10 + down)).to(DEV).requires_grad_(True)
        y = upfirdn2d(x, k.to(DEV), up, down, pad)
        go = torch.randn(y.shape, generator=gen(5)).to(DEV).requires_grad_(True)
        gx, = torch.autograd.grad(y, x, go, create_graph=True)
        xr = x.detach().cpu().requires_grad_(True)
        yr = O.upfirdn2d(xr, k, up, down, pad)
        gxr, = torch.autograd.grad(yr, xr, go.detach().cpu())
        assert (gx.detach().cpu() - gxr).abs().max() <= 2e-6
        # double backward: d(gx . v)/d(go) = upfirdn2d(v) (linear op)
        v = torch.randn(gx.shape, generator=gen(6)).to(DEV)
        ggo, = torch.autograd.grad(gx, go, v)
        assert (ggo.cpu() - O.upfirdn2d(v.cpu(), k, up, down, pad)).abs().max() <= 2e-6


# ------------------------------------------------------------------------------------------------ reductions
def test_plane_reduce_layouts_and_regions():
    from ppst_b200 import ops
    x = torch.randn(2, 24, 37, 29, generator=gen(21)) * 2 + 0.3
    ref_sum = x.sum((2, 3))
    ref_sq = (x * x).sum((2, 3))
    ref_max = x.amax((2, 3))
    for dtype, tol in ((torch.float32, 1e-4), (torch.bfloat16, 0.0)):
        xd = x.to(dtype)
        xf = xd.float()
        for cl in (False, True):
            xin = xd.to(DEV)
            if cl:
                xin = xin.permute(0, 2, 3, 1).contiguous()
            r = ops.plane_reduce(xin, cl)
            assert torch.allclose(r["sum"][:, 0].cpu(), xf.sum((2, 3)), rtol=1e-4, atol=1e-2)
            assert torch.allclose(r["sumsq"][:, 0].cpu(), (xf * xf).sum((2, 3)), rtol=1e-4, atol=1e-2)
            assert torch.equal(r["max"][:, 0].cpu(), xf.amax((2, 3)))
    # exact two-pass variance
    xin = x.to(DEV).permute(0, 2, 3, 1).contiguous()
    r1 = ops.plane_reduce(xin, True, want=("sum",))
    mean = (r1["sum"][:, 0] / (37 * 29)).contiguous()
    r2 = ops.plane_reduce(xin, True, shift=mean, want=("sum", "sumsq"))
    var = r2["sumsq"][:, 0] / (37 * 29) - (r2["sum"][:, 0] / (37 * 29)) ** 2
    assert torch.allclose(var.cpu(), x.var((2, 3), unbiased=False), rtol=1e-5, atol=1e-6)
    # regions keyed by the integer label map: bit-exact membership
    lab = O.make_label_map(2, 37, 29, blocky=False)
    want = O.masked_gap_gmp(x, lab)
    for cl in (False, True):
        xin = x.to(DEV)
        if cl:
            xin = xin.permute(0, 2, 3, 1).contiguous()
        r = ops.plane_reduce(xin, cl, labels=lab.to(DEV), n_regions=3, want=("sum", "max"))
        for i in range(3):
            assert torch.allclose(r["sum"][:, i].cpu() / (37 * 29), want[i][0], rtol=1e-4, atol=1e-5)
            assert torch.equal(r["max"][:, i].cpu(), want[i][1])
    # one-hot and pooled membership bitmasks == MaxPool2d of the one-hot mask
    oh = ops.labels_to_onehot(lab.to(DEV)).cpu()
    assert np.array_equal(oh[0].numpy(), O.mask_labels(lab[0].numpy()).astype(np.float32))
    lab2 = O.make_label_map(2, 32, 48)
    oh2 = torch.stack([torch.tensor(O.mask_labels(lab2[b].numpy()), dtype=torch.float) for b in range(2)])
    pooled = F.max_pool2d(oh2, 2, 2)
    bits = ops.labels_pool_or(lab2.to(DEV), 2, 0).cpu()
    for i in range(3):
        assert torch.equal(((bits >> i) & 1).float(), pooled[:, i])
    x2 = torch.randn(2, 8, 16, 24, generator=gen(23))
    r = ops.plane_reduce(x2.to(DEV).permute(0, 2, 3, 1).contiguous(), True, labels=bits.to(DEV), n_regions=3,
                         label_mode=1, want=("sum", "max"))
    for i in range(3):
        xm = x2 * pooled[:, i:i + 1]
        assert torch.allclose(r["sum"][:, i].cpu(), xm.sum((2, 3)), rtol=1e-4, atol=1e-4)
        assert torch.equal(r["max"][:, i].cpu(), xm.amax((2, 3)))


def test_upsample_reduce_matches_interpolate_then_pool():
    """encoder_col.py:119-130 + :160-166: F.interpolate(bilinear, align_corners=False) followed by GAP / GMP, fused."""
    from ppst_b200 import ops
    for (n, lh, lw, cs, c_off, c, scale) in [(2, 64, 64, 512, 0, 32, 8), (2, 64, 64, 512, 32, 64, 4),
                                             (3, 64, 64, 512, 96, 128, 2), (2, 64, 64, 512, 224, 256, 1),
                                             (1, 5, 7, 16, 4, 8, 4), (2, 3, 3, 8, 0, 8, 2), (1, 1, 1, 4, 0, 4, 8)]:
        x = torch.randn(n, lh, lw, cs, generator=gen(lh * 31 + scale))
        r = ops.upsample_reduce(x.to(DEV), c_off, c, scale)
        xs = x[..., c_off:c_off + c].permute(0, 3, 1, 2).contiguous()
        up = F.interpolate(xs, scale_factor=scale, mode="bilinear", align_corners=False) if scale > 1 else xs
        ref_sum, ref_max = up.sum(dim=(2, 3)), up.amax(dim=(2, 3))
        tol = 1e-5 * max(1.0, float(ref_sum.abs().max()))
        assert (r["sum"].cpu() - ref_sum).abs().max() <= max(tol, 2e-3), (lh, lw, c, scale)
        assert (r["max"].cpu() - ref_max).abs().max() <= 1e-5, (lh, lw, c, scale)
    # a sample's result does not depend on the batch it is reduced in
    x = torch.randn(4, 64, 64, 512, generator=gen(5)).to(DEV)
    a = ops.upsample_reduce(x, 32, 64, 4)
    b = ops.upsample_reduce(x[1:2].contiguous(), 32, 64, 4)
    assert torch.equal(a["sum"][1:2], b["sum"]) and torch.equal(a["max"][1:2], b["max"])


def test_affine_apply_torgb_and_affine_layout_conversion():
    """UpsamplingResBlock64 tail + ToRGB conv in one pass (generator.py:63-77, stylegan2_layers.py:477-495), and the
    ToRGB IN + StyleMod affine folded into the NHWC -> NCHW conversion."""
    from ppst_b200 import ops
    g = gen(77)
    n, h, w, c = 2, 8, 128, 128
    y = torch.randn(n, c, h, w, generator=g).bfloat16()
    lo = torch.randn(n, c, h // 2, w // 2, generator=g).bfloat16()
    a, b = torch.rand(n, c, generator=g) + 0.5, torch.randn(n, c, generator=g)
    wr, br = torch.randn(3, c, generator=g) / math.sqrt(c), torch.randn(3, generator=g)
    want_x = ((y.float() * a[:, :, None, None] + b[:, :, None, None] +
               F.interpolate(lo.float(), scale_factor=2, mode="bilinear", align_corners=False)) / math.sqrt(2)).bfloat16()
    want_rgb = F.conv2d(want_x.float(), wr.view(3, c, 1, 1), br)
    yn, ln = ops.nchw_to_nhwc(y.to(DEV)), ops.nchw_to_nhwc(lo.to(DEV))
    for write_out in (True, False):
        x, rgb = ops.affine_apply_torgb(yn, a.to(DEV), b.to(DEV), ln, wr.to(DEV), br.to(DEV), post_gain=1 / math.sqrt(2),
                                        write_out=write_out)
        assert (x is None) == (not write_out)
        if write_out:
            # same values as the unfused kernel, bit for bit
            ref = ops.affine_apply(yn, c, a.to(DEV), b.to(DEV), res=ln, res_mode=2, post_gain=1 / math.sqrt(2))
            assert torch.equal(x, ref)
            assert (x.float().permute(0, 3, 1, 2).cpu() - want_x.float()).abs().max() <= 4e-2
            thin = ops.conv1x1_thin_out(ref, wr.to(DEV), br.to(DEV))
            assert (rgb - thin).abs().max() <= 1e-4
        assert (rgb.permute(0, 3, 1, 2).cpu() - want_rgb).abs().max() <= 5e-2
    ra, rb = torch.rand(n, 3, generator=g) + 0.5, torch.randn(n, 3, generator=g)
    img = ops.nhwc_to_nchw(rgb, a=ra.to(DEV), b=rb.to(DEV)).cpu()
    assert (img - (rgb.permute(0, 3, 1, 2).cpu() * ra[:, :, None, None] + rb[:, :, None, None])).abs().max() <= 1e-5


def test_layout_roundtrip():
    from ppst_b200 import ops
    x = torch.randn(2, 19, 13, 21, generator=gen(31))
    nhwc = ops.nchw_to_nhwc(x.to(DEV))
    assert torch.equal(nhwc.cpu(), x.permute(0, 2, 3, 1))
    assert torch.equal(ops.nhwc_to_nchw(nhwc).cpu(), x)
    b16 = ops.nchw_to_nhwc(x.to(DEV), dtype=torch.bfloat16)
    assert torch.equal(b16.float().cpu(), x.bfloat16().float().permute(0, 2, 3, 1))
    buf = torch.zeros(2, 13, 21, 40, device=DEV)
    ops.nchw_to_nhwc(x.to(DEV), out=buf, c_off=8)
    assert torch.equal(buf[..., 8:27].cpu(), x.permute(0, 2, 3, 1)) and float(buf[..., :8].abs().sum()) == 0
    assert torch.equal(ops.nhwc_to_nchw(buf, c=19, c_off=8).cpu(), x)


# ------------------------------------------------------------------------------------------------ convolution
def pack_w(w):
    """[co, ci, kh, kw] -> [co, kh*kw*ci] (tap-major, channel-minor)."""
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous()


def run_conv(kind, x_nchw, w, bias, stride, off, out_hw, act=False, noise=None, noise_w=0.0, dtype=torch.float32,
             stats=False):
    from ppst_b200 import ops
    from ppst_b200._lib import F32, BF16
    n, ci, h, wd = x_nchw.shape
    co = w.shape[0]
    x = ops.nchw_to_nhwc(x_nchw.to(DEV), dtype=dtype)
    code = F32 if dtype == torch.float32 else BF16
    d = ops.conv_desc((n, h, wd, ci), co, w.shape[2], w.shape[3], stride=stride, off=off, out_hw=out_hw, act=act,
                      slope=0.2, gain=math.sqrt(2) if act else 1.0, noise_w=noise_w, in_dtype=code, out_dtype=code)
    out = torch.empty((n, out_hw[0], out_hw[1], co), device=DEV, dtype=dtype)
    b = bias.to(DEV) if bias is not None else None
    nz = noise.to(DEV).contiguous() if noise is not None else None
    if kind == "simt":
        ops.conv_simt(out, x, pack_w(w).to(DEV), b, nz, d)
        return out
    st = torch.zeros(ops.conv_tc_stats_elems(d), device=DEV) if stats else None
    ops.conv_tc(out, x, pack_w(w).to(DEV, torch.bfloat16), b, nz, st, d)
    return (out, st) if stats else out


CONV_CASES = [  # ci, co, k, stride, pad (symmetric zero pad), h, w
    (16, 24, 3, 1, 1, 19, 23),
    (3, 32, 1, 1, 0, 17, 17),
    (32, 8, 3, 2, 0, 21, 21),
    (16, 16, 5, 2, 0, 19, 19),
    (7, 5, 3, 1, 1, 9, 11),
]


def test_conv_simt_matches_conv2d():
    for ci, co, k, s, p, h, w in CONV_CASES:
        g = gen(ci * 100 + co)
        x = torch.randn(2, ci, h, w, generator=g)
        wt = torch.randn(co, ci, k, k, generator=g) / math.sqrt(ci * k * k)
        b = torch.randn(co, generator=g)
        want = F.conv2d(x, wt, b, stride=s, padding=p)
        oh, ow = want.shape[2:]
        got = run_conv("simt", x, wt, b, s, (-p, -p), (oh, ow))
        assert (got.permute(0, 3, 1, 2).cpu() - want).abs().max() <= 2e-5, (ci, co, k, s)
    # fused epilogue: + bias + noise -> lrelu * sqrt2
    g = gen(77)
    x = torch.randn(2, 16, 12, 12, generator=g)
    wt = torch.randn(8, 16, 3, 3, generator=g) / 12
    b = torch.randn(8, generator=g)
    nz = torch.randn(2, 12, 12, generator=g)
    want = O.fused_leaky_relu(F.conv2d(x, wt, None, padding=1) + 0.3 * nz[:, None], b)
    got = run_conv("simt", x, wt, b, 1, (-1, -1), (12, 12), act=True, noise=nz, noise_w=0.3)
    assert (got.permute(0, 3, 1, 2).cpu() - want).abs().max() <= 2e-5


def pack_convT(w):
    """EqualizedConv2d fused upscale (stylegan2_layers.py:312-321) as 4 sub-pixel 2x2 phase convolutions.
    w: [co, ci, 3, 3] -> packed [4][co][2*2*ci] with tap (t,u) of phase (a,b) = W'[ci,co,3-a-2t,3-b-2u]."""
    wt = F.pad(w.permute(1, 0, 2, 3), [1, 1, 1, 1])
    wt = wt[:, :, 1:, 1:] + wt[:, :, :-1, 1:] + wt[:, :, 1:, :-1] + wt[:, :, :-1, :-1]      # [ci, co, 4, 4]
    ci, co = wt.shape[:2]
    out = torch.empty(2, 2, co, 2, 2, ci)
    for a in range(2):
        for b in range(2):
            for t in range(2):
                for u in range(2):
                    out[a, b, :, t, u, :] = wt[:, :, 3 - a - 2 * t, 3 - b - 2 * u].t()
    return out.reshape(4, co, 4 * ci).contiguous(), wt


def test_conv_transpose_as_phases():
    from ppst_b200 import ops
    from ppst_b200._lib import F32, BF16
    g = gen(99)
    ci, co, h, w = 64, 32, 18, 20
    x = torch.randn(2, ci, h, w, generator=g)
    wt = torch.randn(co, ci, 3, 3, generator=g) / math.sqrt(ci * 9)
    b = torch.randn(co, generator=g)
    packed, w4 = pack_convT(wt)
    want = F.conv_transpose2d(x, w4, stride=2, padding=1) + b.view(1, -1, 1, 1)
    for kind, dtype, tol in (("simt", torch.float32, 3e-5), ("tc", torch.bfloat16, 4e-2)):
        if kind == "tc" and not ops.conv_tc_available():
            pytest.skip("tcgen05 path unavailable")
        code = F32 if dtype == torch.float32 else BF16
        xin = ops.nchw_to_nhwc(x.to(DEV), dtype=dtype)
        d = ops.conv_desc((2, h, w, ci), co, 2, 2, off=(-1, -1), out_hw=(h, w), phases=(2, 2), phase_off=(1, 1),
                          in_dtype=code, out_dtype=code)
        out = torch.empty((2, 2 * h, 2 * w, co), device=DEV, dtype=dtype)
        if kind == "simt":
            ops.conv_simt(out, xin, packed.to(DEV), b.to(DEV), None, d)
            ref = want
        else:
            ops.conv_tc(out, xin, packed.to(DEV, torch.bfloat16), b.to(DEV), None, None, d)
            ref = F.conv_transpose2d(x.bfloat16().float(), w4.bfloat16().float(), stride=2, padding=1) + b.view(1, -1, 1, 1)
        err = (out.float().permute(0, 3, 1, 2).cpu() - ref).abs().max()
        assert err <= tol, (kind, float(err))


TC_CASES = [  # ci, co, k, stride, pad, h, w, batch
    (64, 128, 3, 1, 1, 32, 32, 2),      # BN=128, 128B swizzle
    (128, 256, 3, 1, 1, 16, 24, 1),     # BN=256, ragged tile edge
    (64, 64, 1, 1, 0, 20, 20, 2),       # 1x1, BN=64
    (32, 32, 3, 1, 1, 24, 24, 1),       # cin 32: 64B swizzle, BN=32
    (32, 64, 5, 2, 0, 35, 35, 1),       # stride 2 via tensor-map element strides
    (64, 96, 3, 1, 1, 9, 70, 1),        # cout = 3*32, odd sizes
    (256, 384, 3, 1, 1, 16, 16, 1),     # BN=128 x 3
    (64, 128, 3, 1, 1, 200, 190, 2),    # >= 4 x 148 tiles: halo kernel, 2 stacked sub-tiles; odd tile-row count
    (32, 32, 3, 1, 1, 200, 190, 2),     # same with 64B-swizzled 32-channel blocks
    (64, 64, 3, 1, 1, 203, 97, 3),      # BN=64, ragged bottom / right edges
]


def test_conv_tc_matches_reference():
    from ppst_b200 import ops
    if not ops.conv_tc_available():
        pytest.skip("tcgen05 path unavailable")
    for ci, co, k, s, p, h, w, n in TC_CASES:
        g = gen(ci + co + k)
        x = torch.randn(n, ci, h, w, generator=g)
        wt = torch.randn(co, ci, k, k, generator=g) / math.sqrt(ci * k * k)
        b = torch.randn(co, generator=g)
        ref = F.conv2d(x.bfloat16().float(), wt.bfloat16().float(), b, stride=s, padding=p)
        oh, ow = ref.shape[2:]
        got, st = run_conv("tc", x, wt, b, s, (-p, -p), (oh, ow), dtype=torch.bfloat16, stats=True)
        gotf = got.float().permute(0, 3, 1, 2).cpu()
        err = (gotf - ref).abs().max()
        assert err <= 2e-2 * max(1.0, float(ref.abs().max())), ((ci, co, k, s), float(err))
        # per-tile statistics of the stored values reduce to the plane sums
        half = st.numel() // 2
        parts = half // (n * co)
        ssum = st[:half].view(n, parts, co).sum(1).cpu()
        ssq = st[half:].view(n, parts, co).sum(1).cpu()
        assert torch.allclose(ssum, gotf.sum((2, 3)), rtol=2e-3, atol=2e-2), (ci, co, k, s)
        assert torch.allclose(ssq, (gotf * gotf).sum((2, 3)), rtol=2e-3, atol=2e-2), (ci, co, k, s)
    # fused epilogue on the tensor-core path
    g = gen(5)
    x = torch.randn(2, 64, 16, 16, generator=g)
    wt = torch.randn(64, 64, 3, 3, generator=g) / 24
    b = torch.randn(64, generator=g)
    nz = torch.randn(2, 16, 16, generator=g)
    ref = O.fused_leaky_relu(F.conv2d(x.bfloat16().float(), wt.bfloat16().float(), None, padding=1) + 0.3 * nz[:, None], b)
    got = run_conv("tc", x, wt, b, 1, (-1, -1), (16, 16), act=True, noise=nz, noise_w=0.3, dtype=torch.bfloat16)
    assert (got.float().permute(0, 3, 1, 2).cpu() - ref).abs().max() <= 3e-2


def test_conv_tc_per_sample_scaled_weights():
    """ops.scale_weights + w_img_rows: conv(W * a_n, x_n) per sample (the feature heads' folded InstanceNorm), on the
    3x3 halo kernel and on the 1x1 plain kernel."""
    from ppst_b200 import ops
    from ppst_b200._lib import BF16
    if not ops.conv_tc_available():
        pytest.skip("tcgen05 path unavailable")
    for ci, co, k, h, w, n in ((64, 128, 3, 40, 48, 3), (128, 64, 1, 32, 32, 2)):
        g = gen(ci + k)
        x = torch.randn(n, ci, h, w, generator=g)
        wt = torch.randn(co, ci, k, k, generator=g) / math.sqrt(ci * k * k)
        a = torch.rand(n, ci, generator=g) + 0.5
        w32 = pack_w(wt).to(DEV).contiguous()
        wn = ops.scale_weights(w32, a.to(DEV), k * k, ci)
        assert wn.shape == (n, co, k * k * ci) and wn.dtype == torch.bfloat16
        want_w = (pack_w(wt).view(co, k * k, ci) * a[:, None, None, :]).view(n, co, -1)
        assert (wn.float().cpu() - want_w.bfloat16().float()).abs().max() == 0.0
        d = ops.conv_desc((n, h, w, ci), co, k, k, off=(-(k // 2), -(k // 2)), out_hw=(h, w), in_dtype=BF16,
                          out_dtype=BF16, w_img_rows=co)
        out = torch.empty((n, h, w, co), device=DEV, dtype=torch.bfloat16)
        ops.conv_tc(out, ops.nchw_to_nhwc(x.to(DEV), dtype=torch.bfloat16), wn, None, None, None, d)
        for i in range(n):
            wi = (wt * a[i].view(1, ci, 1, 1)).bfloat16().float()
            ref = F.conv2d(x[i:i + 1].bfloat16().float(), wi, None, padding=k // 2)
            err = (out[i:i + 1].float().permute(0, 3, 1, 2).cpu() - ref).abs().max()
            assert err <= 2e-2 * max(1.0, float(ref.abs().max())), (ci, co, k, i, float(err))


def test_conv_tc_lazy_affine_input_padded_output():
    """StyledConv conv1 -> [IN + StyleMod never materialised] -> conv2 (stylegan2_layers.py:439-475):
    conv1 stores into the interior of a padded buffer (out_pad), ppst_fill_border sets the border to -b/a,
    ppst_modulate_weights folds the affine into per-sample weights and a per-sample bias, conv2 reads the padded
    buffer as a valid convolution.  Must equal conv2(zero_pad(a*y + b))."""
    from ppst_b200 import ops
    from ppst_b200._lib import BF16
    if not ops.conv_tc_available():
        pytest.skip("tcgen05 path unavailable")
    for ci, cm, co, h, w, n in ((64, 128, 128, 40, 48, 3), (128, 256, 256, 24, 32, 2), (32, 64, 32, 200, 190, 2)):
        g = gen(ci + co + h)
        x = torch.randn(n, ci, h, w, generator=g)
        w1 = torch.randn(cm, ci, 3, 3, generator=g) / math.sqrt(ci * 9)
        b1 = torch.randn(cm, generator=g) * 0.3
        w2 = torch.randn(co, cm, 3, 3, generator=g) / math.sqrt(cm * 9)
        b2 = torch.randn(co, generator=g)
        a = (torch.rand(n, cm, generator=g) + 0.5) * torch.where(torch.rand(n, cm, generator=g) < 0.3, -1.0, 1.0)
        b = torch.randn(n, cm, generator=g)
        xin = ops.nchw_to_nhwc(x.to(DEV), dtype=torch.bfloat16)
        # conv1 -> interior of a padded buffer, with statistics of the stored interior
        d1 = ops.conv_desc((n, h, w, ci), cm, 3, 3, off=(-1, -1), out_hw=(h, w), in_dtype=BF16, out_dtype=BF16, out_pad=1)
        ypad = torch.full((n, h + 2, w + 2, cm), float("nan"), device=DEV, dtype=torch.bfloat16)
        st = torch.empty(ops.conv_tc_stats_elems(d1), device=DEV)
        ops.conv_tc(ypad, xin, pack_w(w1).to(DEV).bfloat16().contiguous(), b1.to(DEV), None, st, d1)
        d0 = ops.conv_desc((n, h, w, ci), cm, 3, 3, off=(-1, -1), out_hw=(h, w), in_dtype=BF16, out_dtype=BF16)
        y = torch.empty((n, h, w, cm), device=DEV, dtype=torch.bfloat16)
        st0 = torch.empty(ops.conv_tc_stats_elems(d0), device=DEV)
        ops.conv_tc(y, xin, pack_w(w1).to(DEV).bfloat16().contiguous(), b1.to(DEV), None, st0, d0)
        assert torch.equal(ypad[:, 1:-1, 1:-1, :], y) and torch.equal(st, st0)
        assert bool(torch.isnan(ypad[:, 0].float()).all()) and bool(torch.isnan(ypad[:, :, -1].float()).all())
        ops.fill_border(ypad, a.to(DEV), b.to(DEV), 1)
        want_border = (-b / a).bfloat16().float()
        for sl in (ypad[:, 0], ypad[:, -1], ypad[:, :, 0], ypad[:, :, -1]):
            assert (sl.float().cpu() - want_border[:, None, :]).abs().max() == 0.0
        assert torch.equal(ypad[:, 1:-1, 1:-1, :], y)
        wn, bn = ops.modulate_weights(pack_w(w2).to(DEV).contiguous(), a.to(DEV), b.to(DEV), 9, cm, base_bias=b2.to(DEV))
        want_bias = b2[None] + torch.einsum("ockl,nc->no", w2, b)
        assert (bn.cpu() - want_bias).abs().max() <= 1e-4 * max(1.0, float(want_bias.abs().max()))
        d2 = ops.conv_desc((n, h + 2, w + 2, cm), co, 3, 3, off=(0, 0), out_hw=(h, w), in_dtype=BF16, out_dtype=BF16,
                           w_img_rows=co, bias_img_stride=co, act=True, gain=math.sqrt(2))
        out = torch.empty((n, h, w, co), device=DEV, dtype=torch.bfloat16)
        ops.conv_tc(out, ypad, wn, bn, None, None, d2)
        yf = y.float().permute(0, 3, 1, 2).cpu()
        xn = yf * a[:, :, None, None] + b[:, :, None, None]
        ref = F.leaky_relu(F.conv2d(xn, w2, b2, padding=1), 0.2) * math.sqrt(2)
        err = (out.float().permute(0, 3, 1, 2).cpu() - ref).abs()
        scale = max(1.0, float(ref.abs().max()))
        assert err.max() <= 4e-2 * scale and err.mean() <= 4e-3 * scale, ((ci, cm, co), float(err.max()), float(err.mean()))
        # the border rows / columns (where the padding matters) are as accurate as the interior
        edge = torch.cat([err[:, :, 0].flatten(), err[:, :, -1].flatten(), err[:, :, :, 0].flatten(), err[:, :, :, -1].flatten()])
        assert edge.max() <= 4e-2 * scale and edge.mean() <= 5e-3 * scale


def test_conv_reads_interior_of_padded_buffer():
    from ppst_b200 import ops
    from ppst_b200._lib import F32
    g = gen(41)
    x = torch.randn(1, 16, 10, 12, generator=g)
    wt = torch.randn(8, 16, 3, 3, generator=g) / 12
    want = F.conv2d(x, wt, None, stride=2, padding=1)
    xp = F.pad(x, (1, 1, 1, 1), mode="reflect")                  # the conv must NOT see this border
    buf = ops.nchw_to_nhwc(xp.to(DEV))
    view = buf[:, 1:, 1:, :]
    d = ops.conv_desc((1, 10, 12, 16), 8, 3, 3, stride=2, off=(-1, -1), out_hw=tuple(want.shape[2:]), in_dtype=F32,
                      out_dtype=F32, in_row_pitch=14 * 16, in_img_pitch=12 * 14 * 16)
    out = torch.empty((1,) + tuple(want.shape[2:]) + (8,), device=DEV)
    ops.conv_simt(out, view, pack_w(wt).to(DEV), None, None, d)
    assert (out.permute(0, 3, 1, 2).cpu() - want).abs().max() <= 2e-5


# ------------------------------------------------------------------------------------------------ epilogues
def test_in_coeffs_and_affine_apply():
    from ppst_b200 import ops
    g = gen(51)
    x = torch.randn(2, 16, 10, 14, generator=g) * 1.5 + 0.4
    style = torch.randn(2, 32, generator=g)
    bias = torch.randn(16, generator=g)
    xin = ops.nchw_to_nhwc(x.to(DEV))
    r = ops.plane_reduce(xin, True, want=("sum", "sumsq"))
    a, b = ops.in_coeffs(2, 16, r["sum"][:, 0].contiguous(), r["sumsq"][:, 0].contiguous(), count=140,
                         style=style.to(DEV), bias=bias.to(DEV))
    y = ops.affine_apply(xin, 16, a, b)
    s = style.view(2, 2, 16, 1, 1)
    want = O.instance_norm(x) * (s[:, 0] + 1) + s[:, 1] + bias.view(1, -1, 1, 1)
    assert (y.permute(0, 3, 1, 2).cpu() - want).abs().max() <= 2e-5
    # pre-activation, residual with its own affine, post gain  (encoder ResBlock tail)
    res = torch.randn(2, 16, 10, 14, generator=g)
    ra, rb = torch.rand(2, 16, generator=g) + 0.5, torch.randn(2, 16, generator=g)
    y = ops.affine_apply(xin, 16, a, b, res=ops.nchw_to_nhwc(res.to(DEV)), ra=ra.to(DEV), rb=rb.to(DEV),
                         pre_slope=0.2, pre_gain=math.sqrt(2), res_mode=1, post_gain=1 / math.sqrt(2))
    w2 = (F.leaky_relu(want, 0.2) * math.sqrt(2) + res * ra[:, :, None, None] + rb[:, :, None, None]) / math.sqrt(2)
    assert (y.permute(0, 3, 1, 2).cpu() - w2).abs().max() <= 2e-5
    # post activation after the residual (ResidualBlock tail, generator.py:29-31)
    y = ops.affine_apply(xin, 16, a, b, res=ops.nchw_to_nhwc(res.to(DEV)), res_mode=1, post_slope=0.3)
    w3 = want + res
    w3 = torch.where(w3 >= 0, w3, 0.3 * w3)
    assert (y.permute(0, 3, 1, 2).cpu() - w3).abs().max() <= 2e-5
    # paddings, nearest x2, bilinear x2 skip
    for mode, name in ((1, "reflect"), (2, "replicate"), (0, "constant")):
        y = ops.affine_apply(xin, 16, a, b, pad=(2, 1, 2, 1), pad_mode=mode)
        assert (y.permute(0, 3, 1, 2).cpu() - F.pad(want, (2, 1, 2, 1), mode=name)).abs().max() <= 2e-5
    y = ops.affine_apply(xin, 16, a, b, up=2, pad=(1, 1, 1, 1), pad_mode=0)
    wu = F.pad(want.repeat_interleave(2, 2).repeat_interleave(2, 3), (1, 1, 1, 1))
    assert (y.permute(0, 3, 1, 2).cpu() - wu).abs().max() <= 2e-5
    lo = torch.randn(2, 16, 5, 7, generator=g)
    y = ops.affine_apply(xin, 16, a, b, res=ops.nchw_to_nhwc(lo.to(DEV)), res_mode=2, post_gain=1 / math.sqrt(2))
    w4 = (want + F.interpolate(lo, scale_factor=2, mode="bilinear", align_corners=False)) / math.sqrt(2)
    assert (y.permute(0, 3, 1, 2).cpu() - w4).abs().max() <= 2e-5
    y = ops.affine_apply(xin, 16, a, b, res=ops.nchw_to_nhwc(lo.to(DEV)), ra=ra.to(DEV), rb=rb.to(DEV), res_mode=2)
    w5 = want + F.interpolate(lo, scale_factor=2, mode="bilinear", align_corners=False) * ra[:, :, None, None] + rb[:, :, None, None]
    assert (y.permute(0, 3, 1, 2).cpu() - w5).abs().max() <= 2e-5
    xb = torch.randn(2, 128, 32, 64, generator=g).bfloat16()                     # an UpsamplingResBlock tail shape
    lob = torch.randn(2, 128, 16, 32, generator=g).bfloat16()
    ab, bb = torch.rand(2, 128, generator=g) + 0.5, torch.randn(2, 128, generator=g)
    y = ops.affine_apply(ops.nchw_to_nhwc(xb.to(DEV)), 128, ab.to(DEV), bb.to(DEV), res=ops.nchw_to_nhwc(lob.to(DEV)),
                         res_mode=2, post_gain=1 / math.sqrt(2), out_dtype=torch.float32)
    w6 = (xb.float() * ab[:, :, None, None] + bb[:, :, None, None] +
          F.interpolate(lob.float(), scale_factor=2, mode="bilinear", align_corners=False)) / math.sqrt(2)
    assert (y.permute(0, 3, 1, 2).cpu() - w6).abs().max() <= 2e-5
    # bf16 storage, 3-channel scalar path, channel-slice destination
    x3 = torch.randn(1, 3, 6, 6, generator=g)
    y = ops.affine_apply(ops.nchw_to_nhwc(x3.to(DEV)), 3, None, None, post_gain=2.0)
    assert torch.equal(y.permute(0, 3, 1, 2).cpu(), x3 * 2)
    dst = torch.zeros(2, 10, 14, 48, device=DEV, dtype=torch.bfloat16)
    ops.affine_apply(ops.nchw_to_nhwc(x.to(DEV), dtype=torch.bfloat16), 16, a, b, out=dst, out_c_off=16)
    assert (dst[..., 16:32].float().permute(0, 3, 1, 2).cpu() - want).abs().max() <= 1e-1
    assert float(dst[..., :16].abs().sum()) == 0 and float(dst[..., 32:].abs().sum()) == 0


def test_resample():
    from ppst_b200 import ops
    x = torch.randn(2, 16, 32, 32, generator=gen(61))
    xin = ops.nchw_to_nhwc(x.to(DEV))
    for size in ((64, 64), (16, 16), (32, 32), (256, 256)):
        y = ops.resample(xin, size, 0)
        assert (y.permute(0, 3, 1, 2).cpu() - F.interpolate(x, size, mode="bilinear")).abs().max() <= 1e-5
    for size in ((8, 8), (16, 16), (32, 32)):
        y = ops.resample(xin, size, 1)
        assert (y.permute(0, 3, 1, 2).cpu() - F.adaptive_avg_pool2d(x, size)).abs().max() <= 1e-5
    dst = torch.zeros(2, 8, 8, 64, device=DEV)
    ops.resample(xin, (8, 8), 1, out=dst, out_c_off=32)
    assert (dst[..., 32:48].permute(0, 3, 1, 2).cpu() - F.adaptive_avg_pool2d(x, 8)).abs().max() <= 1e-5
    # normalisation + PReLU of the source applied on load (the feature-head tails, generator.py:272-278)
    a, b = torch.randn(2, 16, generator=gen(62)), torch.randn(2, 16, generator=gen(63))
    xt = F.leaky_relu(x * a[:, :, None, None] + b[:, :, None, None], 0.25)
    for xi, xr in ((xin, xt), (xin.bfloat16(), F.leaky_relu(x.bfloat16().float() * a[:, :, None, None] + b[:, :, None, None], 0.25))):
        y = ops.resample(xi, (16, 16), 0, a=a.to(DEV), b=b.to(DEV), slope=0.25, out_dtype=torch.float32)
        assert (y.permute(0, 3, 1, 2).cpu() - F.interpolate(xr, (16, 16), mode="bilinear")).abs().max() <= 1e-5
        y = ops.resample(xi, (4, 4), 1, a=a.to(DEV), b=b.to(DEV), slope=0.25, out_dtype=torch.float32)
        assert (y.permute(0, 3, 1, 2).cpu() - F.adaptive_avg_pool2d(xr, 4)).abs().max() <= 1e-5


def test_resample_integer_upsampling_kernel_matches_the_generic_one():
    """resample_up_kernel (x1 / x2 / x4 / x8 bilinear, the four neighbours of a column pair loaded and transformed once for
    the F outputs between them) against the generic per-output kernel (PPST_RESAMPLE_UP=0) and F.interpolate: every dtype
    pair, affine + PReLU on load, bordered input and output buffers, a channel slice of a wider output, 1-pixel-wide maps."""
    import os
    from ppst_b200 import ops
    g = gen(64)
    cases = [(2, 16, 16, 64, 4, 0, 0), (1, 8, 24, 16, 8, 0, 0), (2, 32, 20, 64, 2, 1, 0), (1, 64, 64, 64, 1, 0, 1),
             (2, 9, 1, 8, 4, 0, 0), (1, 1, 7, 8, 2, 0, 0), (3, 5, 6, 24, 8, 1, 1), (1, 3, 3, 8, 1, 1, 0)]
    old = os.environ.get("PPST_RESAMPLE_UP")
    try:
        for n, h, w, c, f, ib, ob in cases:
            x = torch.randn(n, c, h, w, generator=g)
            a, b = torch.randn(n, c, generator=g).to(DEV), torch.randn(n, c, generator=g).to(DEV)
            for dt_in, dt_out in ((torch.float32, torch.float32), (torch.bfloat16, torch.bfloat16), (torch.bfloat16, torch.float32),
                                  (torch.float32, torch.bfloat16)):
                xin = ops.nchw_to_nhwc(x.to(DEV), dtype=dt_in)
                if ib:
                    xin = torch.nn.functional.pad(xin, (0, 0, ib, ib, ib, ib), value=float("nan")).contiguous()
                for tf in (False, True):
                    outs = {}
                    for flag in ("0", "1"):
                        os.environ["PPST_RESAMPLE_UP"] = flag
                        dst = torch.full((n, h * f + 2 * ob, w * f + 2 * ob, c + 16), 7.0, device=DEV, dtype=dt_out)
                        ops.resample(xin, (h * f, w * f), 0, out=dst, out_c_off=8, in_border=ib, out_border=ob,
                                     **(dict(a=a, b=b, slope=0.25) if tf else {}))
                        outs[flag] = dst
                    tol = 0.0 if dt_out == torch.float32 else 0.0
                    d = (outs["0"].float() - outs["1"].float()).abs().max()
                    assert float(d) <= 1e-6 * (1 if dt_out == torch.float32 else 0) + (4e-2 if dt_out == torch.bfloat16 else 0), \
                        (n, h, w, c, f, ib, ob, dt_in, dt_out, tf, float(d))
                    inner = outs["1"][:, ob:ob + h * f, ob:ob + w * f]
                    assert float((inner[..., :8].float() - 7.0).abs().max()) == 0 and float((inner[..., 8 + c:].float() - 7.0).abs().max()) == 0
                    if ob:
                        assert float((outs["1"][:, 0].float() - 7.0).abs().max()) == 0        # the border is the caller's
                    xr = x.to(dt_in).float()
                    if tf:
                        xr = F.leaky_relu(xr * a.cpu()[:, :, None, None] + b.cpu()[:, :, None, None], 0.25)
                    want = F.interpolate(xr, (h * f, w * f), mode="bilinear")
                    got = inner[..., 8:8 + c].float().permute(0, 3, 1, 2).cpu()
                    assert (got - want).abs().max() <= (1e-5 if dt_out == torch.float32 else 4e-2) * max(1.0, float(want.abs().max()))
    finally:
        if old is None:
            os.environ.pop("PPST_RESAMPLE_UP", None)
        else:
            os.environ["PPST_RESAMPLE_UP"] = old


# ------------------------------------------------------------------------------------------------ dense / corr
def test_linear_l2norm_lerp():
    from ppst_b200 import ops
    g = gen(71)
    for m, k, n in ((1, 64, 1024), (3, 2048, 2048), (9, 30, 17)):
        x = torch.randn(m, k, generator=g)
        w = torch.randn(n, k, generator=g) / math.sqrt(k)
        b = torch.randn(n, generator=g)
        y = ops.linear(x.to(DEV), w.to(DEV), b.to(DEV), w_scale=0.5, b_scale=2.0, relu_in=True)
        want = F.linear(F.relu(x), w * 0.5, b * 2.0)
        assert (y.cpu() - want).abs().max() <= 1e-4
    v = torch.randn(5, 2048, generator=g)
    assert torch.allclose(ops.l2_normalize(v.to(DEV), 1e-8, 0).cpu(), O.l2_normalize_rows(v), atol=1e-7)
    assert torch.allclose(ops.l2_normalize(v.to(DEV), 1e-12, 1).cpu(), F.normalize(v), atol=1e-7)
    a, b = torch.randn(4, 2048, generator=g), torch.randn(4, 2048, generator=g)
    assert torch.allclose(ops.lerp(a.to(DEV), b.to(DEV), 0.3).cpu(), O.lerp(a, b, 0.3), atol=1e-7)


def test_gemm_and_softmax():
    from ppst_b200 import ops
    g = gen(81)
    a = torch.randn(2, 300, 70, generator=g)
    bt = torch.randn(2, 130, 70, generator=g)
    bn = torch.randn(2, 70, 130, generator=g)
    assert (ops.gemm_f32(a.to(DEV), bt.to(DEV), True, alpha=0.5).cpu() - 0.5 * a @ bt.transpose(1, 2)).abs().max() <= 1e-4
    assert (ops.gemm_f32(a.to(DEV), bn.to(DEV), False).cpu() - a @ bn).abs().max() <= 1e-4
    x = torch.randn(33, 4096, generator=g) * 0.3
    got = ops.softmax_rows_(x.to(DEV).clone(), 0.01).cpu()
    want = F.softmax(x / 0.01, dim=-1)
    assert (got - want).abs().max() <= 1e-5 and (got.sum(-1) - 1).abs().max() <= 1e-5


def test_rselfcorr_and_corr_prepare():
    from ppst_b200 import ops
    g = gen(91)
    fea = torch.randn(2, 64, 32, 48, generator=g)
    # oracle reshape is hard-wired to 64x64 patches -> use the general restatement on 8x12 patches
    b, c, h, w = fea.shape
    p = fea.reshape(b, c, h // 4, 4, w // 4, 4).permute(0, 1, 2, 4, 3, 5).reshape(b, c, -1, 16)
    p = p - p.mean(1, keepdim=True)
    p = p / (p.pow(2).sum(1, keepdim=True).sqrt() + 2.220446049250313e-16)
    want = torch.einsum("bcli,bclj->blij", p, p).reshape(b, h // 4, w // 4, 256)
    got = ops.rselfcorr(ops.nchw_to_nhwc(fea.to(DEV)))
    assert (got.cpu() - want).abs().max() <= 1e-5
    # bf16 input written into an f32 channel slice (the swap path), and a width that takes the generic kernel
    fb = fea.bfloat16()
    pb = fb.float().reshape(b, c, h // 4, 4, w // 4, 4).permute(0, 1, 2, 4, 3, 5).reshape(b, c, -1, 16)
    pb = pb - pb.mean(1, keepdim=True)
    pb = pb / (pb.pow(2).sum(1, keepdim=True).sqrt() + 2.220446049250313e-16)
    wantb = torch.einsum("bcli,bclj->blij", pb, pb).reshape(b, h // 4, w // 4, 256)
    slab = torch.full((b, h // 4, w // 4, 512), 7.0, device=DEV)
    ops.rselfcorr(ops.nchw_to_nhwc(fb.to(DEV)), out=slab, out_c_off=256)
    assert (slab[..., 256:].cpu() - wantb).abs().max() <= 1e-5 and bool((slab[..., :256] == 7.0).all())
    f32c = torch.randn(1, 32, 8, 8, generator=g)
    p32 = f32c.reshape(1, 32, 2, 4, 2, 4).permute(0, 1, 2, 4, 3, 5).reshape(1, 32, -1, 16)
    p32 = p32 - p32.mean(1, keepdim=True)
    p32 = p32 / (p32.pow(2).sum(1, keepdim=True).sqrt() + 2.220446049250313e-16)
    want32 = torch.einsum("bcli,bclj->blij", p32, p32).reshape(1, 2, 2, 256)
    assert (ops.rselfcorr(ops.nchw_to_nhwc(f32c.to(DEV))).cpu() - want32).abs().max() <= 1e-5
    f = torch.randn(1, 512, 8, 8, generator=g)
    rows = ops.nchw_to_nhwc(f.to(DEV))
    prep = ops.corr_prepare(rows, 256).cpu().reshape(64, 512)
    ff = f.reshape(1, 512, 64)
    h1 = ff[:, :256] - ff[:, :256].mean(1, keepdim=True)
    ff = torch.cat((h1, ff[:, 256:]), 1)
    ff = ff / (ff.pow(2).sum(1, keepdim=True).sqrt() + 2.220446049250313e-16)
    assert (prep - ff[0].t()).abs().max() <= 1e-6


def test_split_gemm_on_tensor_cores_is_fp32_accurate():
    """corrm logits and corr @ feat in bf16 mode: hi*hi + hi*lo + lo*hi on tcgen05 (ppst_conv_tc, split_terms=3)."""
    from ppst_b200 import ops
    if not ops.conv_tc_available():
        pytest.skip("tcgen05 path unavailable")
    g = gen(101)
    a = torch.randn(2, 256, 512, generator=g) / math.sqrt(512)
    b = torch.randn(2, 320, 512, generator=g) / math.sqrt(512)
    want = (a.double() @ b.double().transpose(1, 2)).float()
    got = ops.gemm_split_tc(ops.split_bf16(a.to(DEV)), ops.split_bf16(b.to(DEV)), (4, 64)).cpu()
    assert (got - want).abs().max() <= 2e-5 * float(want.abs().max()) + 1e-6
    # transposed split (NHWC feature map -> K-major B operand) and the split softmax
    f = torch.randn(2, 512, 96, generator=g)
    corr = torch.softmax(torch.randn(2, 128, 512, generator=g) * 3, -1)
    want = (corr.double() @ f.double()).float()
    got = ops.gemm_split_tc(ops.split_bf16(corr.to(DEV)), ops.split_bf16(f.to(DEV), transpose=True), (2, 64)).cpu()
    assert (got - want).abs().max() <= 1e-4
    logits = torch.randn(64, 4096, generator=g) * 0.3
    sp = ops.softmax_rows_split(logits.to(DEV), 0.01).float().cpu()
    rec = sp[:, :4096] + sp[:, 4096:]
    assert (rec - F.softmax(logits / 0.01, -1)).abs().max() <= 1e-5
    for cols in (1000, 5000, 4098):          # short rows (register path), long rows and odd lengths (three-pass path)
        logits = torch.randn(5, cols, generator=g) * 0.3
        sp = ops.softmax_rows_split(logits.to(DEV), 0.01).float().cpu()
        assert (sp[:, :cols] + sp[:, cols:] - F.softmax(logits / 0.01, -1)).abs().max() <= 1e-5
    x = torch.randn(1, 8, 8, 512, generator=g)
    sp = ops.corr_prepare_split(x.to(DEV), 256).float().cpu().view(64, 1024)
    ref = ops.corr_prepare(x.to(DEV), 256).cpu().view(64, 512)
    assert (sp[:, :512] + sp[:, 512:] - ref).abs().max() <= 1e-6


def test_guided_smooth_matches_oracle():
    """decode(target=...) post-process (ppst_model.py:288-306, photo_gif.py:13-46) on the device vs the numpy restatement:
    the integer window sums are exact and every fp32 step is order-matched, so at most a handful of pixels (double-
    precision sums landing on a rounding boundary) may differ, and then by one uint8 level."""
    from oracle import guided_oracle as G
    from ppst_b200 import ops
    for (n, h, w, r) in ((2, 96, 130, 30), (1, 512, 512, 30), (1, 70, 64, 5)):
        out = O.make_portraits(n, h, w, seed=5, smooth=True) * 1.2        # values outside [-1, 1] exercise the clip
        tgt = O.make_portraits(n, h, w, seed=6, smooth=True) * 0.7 + 0.2 * O.make_portraits(n, h, w, seed=7)
        want = G.guided_smooth(out.numpy(), tgt.numpy(), r=r)
        got = ops.guided_smooth(out.to(DEV), tgt.to(DEV), radius=r).cpu().numpy()
        lv = np.abs(np.rint((got - want) * 127.5))
        assert lv.max() <= 1, (n, h, w, lv.max())
        assert (lv > 0).mean() <= 1e-4, (n, h, w, float((lv > 0).mean()))
    with pytest.raises(RuntimeError):
        ops.guided_smooth(torch.zeros(1, 3, 20, 20, device=DEV), torch.zeros(1, 3, 20, 20, device=DEV), radius=30)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [
    # (n, h, w, cin, cout, mode)  mode: "3x3" zero-padded, "up" = fused 2x2-phase transposed conv, "lazy" = per-sample weights
    (4, 128, 128, 128, 128, "3x3"), (2, 64, 96, 256, 128, "3x3"), (3, 64, 64, 256, 384, "3x3"),
    (2, 128, 128, 256, 128, "up"), (2, 96, 64, 128, 128, "lazy"), (2, 128, 128, 256, 256, "3x3"), (2, 64, 64, 512, 512, "up"),
])
def test_conv_tc_cta_pair_kernel_is_bit_identical_to_single_cta(shape):
    """conv_tc_halo2_kernel (tcgen05.mma.cta_group::2, B tiles split over the two CTAs of a pair) against the 1-CTA halo
    kernel: same K order per output element, so outputs AND instance-norm partials must agree bit for bit."""
    import math
    import os
    from ppst_b200 import ops, packing
    from ppst_b200._lib import BF16
    from ppst_b200.engine import Engine
    if not ops.conv_tc_available():
        pytest.skip("tcgen05 path unavailable")
    n, h, w, ci, co, mode = shape
    eng = Engine.__new__(Engine)
    eng.use_tc, eng.adt, eng.device, eng.mode, eng.acode = True, torch.bfloat16, torch.device("cuda"), "bf16", BF16
    g = torch.Generator().manual_seed(5)
    x = torch.randn(n, h, w, ci, generator=g).to(torch.bfloat16).cuda()
    wt = torch.randn(co, ci, 3, 3, generator=g) / math.sqrt(ci * 9)
    bias = torch.randn(co, generator=g)
    kw = {}
    if mode == "up":
        spec = packing.make_upscale_spec(wt, "cuda", bias=bias, act=True, noise_w=0.3)
        noise = torch.randn(n, 1, 2 * h, 2 * w, generator=g).cuda()
    else:
        spec = packing.make_spec(wt, "cuda", bias=bias, act=True, pad=1, noise_w=0.3)
        noise = torch.randn(n, 1, h, w, generator=g).cuda()
        if mode == "lazy":
            a = (torch.rand(n, ci, generator=g) + 0.5).cuda()
            b = torch.randn(n, ci, generator=g).cuda()
            wn, bn = ops.modulate_weights(spec.w32, a, b, 9, ci, base_bias=spec.bias)
            kw = dict(w_per_image=wn, bias_per_image=bn)
    res = {}
    old = os.environ.get("PPST_CONV_TC_2CTA")
    try:
        for flag in ("0", "2"):
            os.environ["PPST_CONV_TC_2CTA"] = flag
            y, st = eng.conv(x, spec, noise=noise, want_stats=True, out_pad=1 if mode == "up" else 0, **kw)
            torch.cuda.synchronize()
            res[flag] = (y.clone(), st["sum"].clone(), st["sumsq"].clone())
    finally:
        if old is None:
            os.environ.pop("PPST_CONV_TC_2CTA", None)
        else:
            os.environ["PPST_CONV_TC_2CTA"] = old
    if mode == "up":
        i0, i2 = res["0"][0][:, 1:-1, 1:-1], res["2"][0][:, 1:-1, 1:-1]        # the border belongs to the caller
    else:
        i0, i2 = res["0"][0], res["2"][0]
    assert torch.equal(i0, i2)
    assert torch.equal(res["0"][1], res["2"][1]) and torch.equal(res["0"][2], res["2"][2])


@pytest.mark.gpu
@pytest.mark.parametrize("case", [(2, 64, 64, 32, 64, 5, 0), (1, 128, 96, 32, 64, 3, -1), (2, 48, 80, 64, 128, 5, 0),
                                  (1, 64, 64, 128, 256, 3, -1), (3, 32, 48, 32, 32, 6, -2), (16, 256, 256, 32, 64, 5, 0),
                                  (8, 256, 192, 64, 128, 3, -1), (1, 70, 50, 32, 64, 4, -1), (16, 256, 256, 64, 128, 5, 0),
                                  (8, 128, 256, 128, 128, 6, -2), (16, 128, 128, 128, 256, 5, 0)])
def test_conv_tc_stride2_halo_boxes(case):
    """The blurred stride-2 convolutions of ResBlock: plain kernel (one strided box per tap) and the halo kernel with row
    stride 2 (one even-row + one odd-row box per kernel column, taps = row offsets inside the box of their parity, 1 / 2 /
    4 sub-tiles per work item) against F.conv2d on the same bf16 operands; statistics partials must add up to the sums of
    the stored outputs, and the super-tile heights must agree bit for bit (same K order)."""
    import math
    import os
    import torch.nn.functional as F
    from ppst_b200 import ops, packing
    from ppst_b200._lib import BF16
    if not ops.conv_tc_available():
        pytest.skip("tcgen05 path unavailable")
    n, h, w, ci, co, k, off = case
    g = torch.Generator().manual_seed(k * 100 + ci)
    pad = -off
    ih, iw = (h, w) if off else (h + k - 1, w + k - 1)       # off = 0: the caller pre-padded
    oh, ow = ((h - 2) // 2 + 1, (w - 2) // 2 + 1) if off else (h // 2, w // 2)
    x = torch.randn(n, ih, iw, ci, generator=g).to(torch.bfloat16).cuda()
    wt = (torch.randn(co, ci, k, k, generator=g) / math.sqrt(ci * k * k)).to(torch.bfloat16)
    bias = torch.randn(co, generator=g).cuda()
    xr = x.float().permute(0, 3, 1, 2)
    ref = F.conv2d(F.pad(xr, [pad, pad + k, pad, pad + k]), wt.float().cuda(), stride=2)[:, :, :oh, :ow] + bias.view(1, -1, 1, 1)
    ref = ref.permute(0, 2, 3, 1)
    wq = packing.pack_taps(wt.float()).unsqueeze(0).to(torch.bfloat16).cuda()
    d = ops.conv_desc((n, ih, iw, ci), co, k, k, stride=2, off=(off, off), out_hw=(oh, ow), in_dtype=BF16, out_dtype=BF16)
    old = os.environ.get("PPST_CONV_TC_HALO_S2")
    outs = {}
    try:
        for flag in ("0", "1", "2", "3"):
            os.environ["PPST_CONV_TC_HALO_S2"] = flag
            y = torch.full((n, oh, ow, co), float("nan"), device="cuda", dtype=torch.bfloat16)
            st = torch.empty(ops.conv_tc_stats_elems(d), device="cuda")
            ops.conv_tc(y, x, wq, bias, None, st, d)
            torch.cuda.synchronize()
            err = float((y.float() - ref).norm() / ref.norm())
            assert err <= 4e-3, (case, flag, err)
            half = st.numel() // 2
            s1 = st[:half].view(n, -1, co).sum(1)
            assert torch.allclose(s1, y.float().sum((1, 2)), rtol=1e-3, atol=2e-2), (case, flag)
            outs[flag] = (y, st.clone())
    finally:
        if old is None:
            os.environ.pop("PPST_CONV_TC_HALO_S2", None)
        else:
            os.environ["PPST_CONV_TC_HALO_S2"] = old
    for flag in ("2", "3"):
        assert torch.equal(outs["1"][0], outs[flag][0]), (case, flag)
