"""Tensors produced INSIDE their consumer's padding border (no padding passes): the border-aware variants of the kernels are
bit-identical to the plain ones, ppst_pad_border equals F.pad, and the whole swap is bit-identical with the feature on or
off (engine.padded)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ops():
    from ppst_b200 import ops
    return ops


def _rand(shape, dtype, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g).to(dtype).cuda()


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("mode,pad", [(1, (1, 1, 1, 1)), (2, (1, 1, 1, 1)), (1, (2, 1, 2, 1)), (0, (1, 2, 3, 0))])
def test_pad_border_matches_f_pad(dtype, mode, pad):
    ops = _ops()
    n, h, w, c = 3, 9, 13, 16
    x = _rand((n, h, w, c), dtype, 1)
    pt, pb, pl, pr = pad
    buf = torch.full((n, h + pt + pb, w + pl + pr, c), 7.0, dtype=dtype, device="cuda")
    buf[:, pt:pt + h, pl:pl + w] = x
    ops.pad_border(buf, pad, mode)
    ref = F.pad(x.permute(0, 3, 1, 2).float(), (pl, pr, pt, pb), mode=["constant", "reflect", "replicate"][mode])
    assert torch.equal(buf.float(), ref.permute(0, 2, 3, 1))


@pytest.mark.parametrize("res_mode", [0, 1, 2])
def test_affine_apply_borders_are_bit_identical(res_mode):
    ops = _ops()
    n, h, w, c = 2, 16, 64, 128
    x = _rand((n, h, w, c), torch.bfloat16, 2)
    a, b = _rand((n, c), torch.float32, 3), _rand((n, c), torch.float32, 4)
    res = None
    if res_mode == 1:
        res = _rand((n, h, w, c), torch.bfloat16, 5)
    elif res_mode == 2:
        res = _rand((n, h // 2, w // 2, c), torch.bfloat16, 5)
    kw = dict(a=a, b=b, res=res, res_mode=res_mode, pre_slope=0.2, post_gain=0.7)
    y0 = ops.affine_apply(x, c, **kw)
    y1 = ops.affine_apply(x, c, out_border=1, **kw)
    assert torch.equal(y1[:, 1:-1, 1:-1], y0)
    if res is not None:
        resp = F.pad(res.permute(0, 3, 1, 2).float(), (1, 1, 1, 1), value=99.0).permute(0, 2, 3, 1).to(torch.bfloat16).contiguous()
        y2 = ops.affine_apply(x, c, res_border=1, out_border=1, **dict(kw, res=resp))
        assert torch.equal(y2[:, 1:-1, 1:-1], y0)
    # in-kernel padding together with an output border (generic residual kernel)
    y3 = ops.affine_apply(x, c, pad=(1, 1, 1, 1), pad_mode=2, **kw)
    y4 = ops.affine_apply(x, c, pad=(1, 1, 1, 1), pad_mode=2, out_border=2, **kw)
    assert torch.equal(y4[:, 2:-2, 2:-2], y3)


@pytest.mark.parametrize("mode,out_hw", [(0, (24, 40)), (1, (4, 8))])
def test_resample_borders_are_bit_identical(mode, out_hw):
    ops = _ops()
    n, h, w, c = 2, 16, 32, 64
    x = _rand((n, h, w, c), torch.bfloat16, 6)
    a, b = _rand((n, c), torch.float32, 7), _rand((n, c), torch.float32, 8)
    y0 = ops.resample(x, out_hw, mode, a=a, b=b, slope=0.3)
    xp = torch.full((n, h + 2, w + 2, c), 5.0, dtype=torch.bfloat16, device="cuda")
    xp[:, 1:-1, 1:-1] = x
    out = torch.zeros((n, out_hw[0] + 2, out_hw[1] + 2, 2 * c), dtype=torch.bfloat16, device="cuda")
    ops.resample(xp, out_hw, mode, out=out, out_c_off=c, a=a, b=b, slope=0.3, in_border=1, out_border=1)
    assert torch.equal(out[:, 1:-1, 1:-1, c:], y0)
    assert float(out[:, 0].abs().max()) == 0.0 and float(out[..., :c].abs().max()) == 0.0


def test_plane_reduce_and_from_rgb_borders():
    ops = _ops()
    n, h, w, c = 3, 20, 24, 64
    x = _rand((n, h, w, c), torch.bfloat16, 9)
    xp = torch.full((n, h + 2, w + 2, c), 50.0, dtype=torch.bfloat16, device="cuda")
    xp[:, 1:-1, 1:-1] = x
    r0 = ops.plane_reduce(x, True, want=("sum", "sumsq", "max"))
    r1 = ops.plane_reduce(xp, True, want=("sum", "sumsq", "max"), border=1)
    for k in r0:
        assert torch.equal(r0[k], r1[k]), k
    labels = torch.randint(0, 3, (n, h, w), generator=torch.Generator().manual_seed(1), dtype=torch.uint8).cuda()
    r0 = ops.plane_reduce(x, True, labels=labels, n_regions=3, want=("sum", "max"))
    r1 = ops.plane_reduce(xp, True, labels=labels, n_regions=3, want=("sum", "max"), border=1)
    for k in r0:
        assert torch.equal(r0[k], r1[k]), k
    img = _rand((2, 3, 40, 56), torch.float32, 10)
    wt, bias = _rand((32, 3), torch.float32, 11), _rand((32,), torch.float32, 12)
    y0 = ops.conv1x1_in_nchw(img, wt, bias, True, torch.bfloat16)
    y1 = ops.conv1x1_in_nchw(img, wt, bias, True, torch.bfloat16, out_border=1)
    assert torch.equal(y1[:, 1:-1, 1:-1], y0)


def test_conv_tc_interior_view_and_trimmed_output_border():
    import math
    from ppst_b200 import packing
    from ppst_b200.engine import Engine
    ops = _ops()
    if not ops.conv_tc_available():
        pytest.skip("tcgen05 path unavailable")
    eng = Engine.__new__(Engine)
    eng.use_tc, eng.adt, eng.device, eng.mode = True, torch.bfloat16, torch.device("cuda"), "bf16"
    from ppst_b200._lib import BF16
    eng.acode = BF16
    n, h, w, ci, co = 2, 32, 48, 64, 64
    x = _rand((n, h, w, ci), torch.bfloat16, 13)
    wt = torch.randn(co, ci, 3, 3, generator=torch.Generator().manual_seed(14)) / math.sqrt(ci * 9)
    bias = torch.randn(co, generator=torch.Generator().manual_seed(15))
    spec = packing.make_spec(wt, "cuda", bias=bias, act=True, pad=1)
    y0, st0 = eng.conv(x, spec, want_stats=True)
    xp = torch.full((n, h + 2, w + 2, ci), 3.0, dtype=torch.bfloat16, device="cuda")
    xp[:, 1:-1, 1:-1] = x
    y1, st1 = eng.conv(xp, spec, want_stats=True, in_border=1)          # zero padding must NOT see the border of 3s
    assert torch.equal(y1, y0) and torch.equal(st1["sum"], st0["sum"])
    y2, _ = eng.conv(x, spec, out_pad=2, out_pad_trim=1)
    assert y2.shape == (n, h + 3, w + 3, co) and torch.equal(y2[:, 2:-1, 2:-1], y0)
    # stride-2 folded skip (3x3, offset -1) through the interior view
    ws = torch.randn(co, ci, 3, 3, generator=torch.Generator().manual_seed(16)) / math.sqrt(ci * 9)
    sk = packing.make_spec(ws, "cuda", stride=2, off=(-1, -1), out_hw=lambda hh, ww: ((hh - 2) // 2 + 1, (ww - 2) // 2 + 1))
    z0, _ = eng.conv(x, sk)
    z1, _ = eng.conv(xp, sk, in_border=1)
    assert torch.equal(z1, z0)


@pytest.mark.parametrize("res_mode,c,hw", [(2, 128, (32, 64)), (2, 512, (16, 16)), (1, 512, (16, 24)), (0, 64, (8, 40))])
@pytest.mark.parametrize("P", [0, 1])
def test_affine_apply_stats_matches_apply_plus_pad_plus_reduce(res_mode, c, hw, P):
    ops = _ops()
    n, (h, w) = 3, hw
    x = _rand((n, h, w, c), torch.bfloat16, 20)
    a, b = _rand((n, c), torch.float32, 21), _rand((n, c), torch.float32, 22)
    res = None
    if res_mode == 1:
        res = _rand((n, h, w, c), torch.bfloat16, 23)
    elif res_mode == 2:
        res = _rand((n, h // 2, w // 2, c), torch.bfloat16, 23)
    kw = dict(a=a, b=b, res=res, res_mode=res_mode, post_gain=0.7071)
    y0 = ops.affine_apply(x, c, **kw)
    ref = F.pad(y0.permute(0, 3, 1, 2).float(), (P, P, P, P), mode="replicate").permute(0, 2, 3, 1) if P else y0.float()
    y1, st = ops.affine_apply_stats(x, c, pad=(P,) * 4, pad_mode=2, **kw)
    assert torch.equal(y1.float(), ref)
    assert st["n_partials"] == h + 2 * P and st["count"] == (h + 2 * P) * (w + 2 * P)
    s1 = st["sum"].double().sum(1)
    s2 = st["sumsq"].double().sum(1)
    r1, r2 = ref.double().sum((1, 2)), (ref.double() ** 2).sum((1, 2))
    assert float((s1.cpu() - r1.cpu()).abs().max()) <= 1e-4 * float(r2.max()) ** 0.5 + 1e-3
    assert float(((s2.cpu() - r2.cpu()).abs() / r2.cpu()).max()) <= 1e-5
    if res_mode == 2:                                   # padded residual input (identity skip of UpsamplingResBlock16)
        resp = F.pad(res.permute(0, 3, 1, 2).float(), (1, 1, 1, 1), value=9.0).permute(0, 2, 3, 1).to(torch.bfloat16).contiguous()
        y2, st2 = ops.affine_apply_stats(x, c, pad=(P,) * 4, pad_mode=2, res_border=1, **dict(kw, res=resp))
        assert torch.equal(y2, y1) and torch.equal(st2["sum"], st["sum"])


def test_swap_with_and_without_padded_producers():
    import ppst_b200
    from ppst_b200 import synthetic as S
    m = ppst_b200.SwapModel(*S.make_networks(0), mode="bf16")
    if not m.engine.use_tc:
        pytest.skip("tcgen05 path unavailable")
    m.noises = S.make_noises(2, 512, 512, seed=5)
    c = S.make_portraits(2, 512, 512, seed=1).cuda()
    s = S.make_portraits(2, 512, 512, seed=2, smooth=True).cuda()
    import math
    e = m.engine
    res = {}
    for flag in (True, False):
        e.padded = flag
        sp = e.encoder_con(c)
        codes, _ = e.encoder_col(s)
        res[flag] = (sp.clone(), [v.clone() for v in codes], m.swap(c, s).clone(),
                     [t.clone() for t in m.extract_feat_from_image(c)])
    # the encoders run the same arithmetic either way: bit-identical
    assert torch.equal(res[True][0], res[False][0])
    assert all(torch.equal(a, b) for a, b in zip(res[True][1], res[False][1]))
    # the feature heads' first InstanceNorm takes its statistics from the producing pass (another summation order than
    # the stand-alone reduction): the coefficients agree to fp32 rounding, which flips the bf16 rounding of a few of the
    # per-sample weights they scale; the correlation softmax (T = 0.01) amplifies that like any other bf16-level noise
    # (the path's bound against the fp32 reference is 45 dB)
    def psnr(a, b):
        peak = float(b.max() - b.min())
        return 10 * math.log10(peak * peak / max(float(((a - b) ** 2).mean()), 1e-30))
    p_img = psnr(res[True][2], res[False][2])
    p_fea = [psnr(a, b) for a, b in zip(res[True][3], res[False][3])]
    print("padded vs plain: image %.1f dB, features %s dB" % (p_img, ["%.1f" % v for v in p_fea]))
    assert p_img >= 50.0 and min(p_fea) >= 50.0, (p_img, p_fea)
