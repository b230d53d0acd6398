"""Seeded synthetic weights and inputs for benchmarks, tools and tests: random-init networks keyed like
the reference's state_dicts (Appendix A), uniform [-1, 1] portraits, {0,1,2} label maps and explicit NoiseInjection maps.
Product-side generators for bench.py's GPU arm and tools/ (tests/test_synthetic.py keeps them bit-identical to the test
infrastructure's own copies, which the product never imports)."""
import math

import torch
import torch.nn.functional as F

SQRT2 = math.sqrt(2.0)
HEAD_CH = [(256, 256), (256, 256), (256, 384), (384, 512)]     # generator.py:156-165
UP_CH = [(16, 512, 512), (32, 512, 256), (64, 256, 128)]       # generator.py:167-175
D_CHANNELS = {4: 512, 8: 512, 16: 512, 32: 512, 64: 512, 128: 256, 256: 128, 512: 64, 1024: 32}


def make_kernel(k):
    """stylegan2_layers.py:28-36."""
    k = torch.tensor(k, dtype=torch.float32)
    if k.dim() == 1:
        k = k[None, :] * k[:, None]
    return k / k.sum()


def d_block_names(size):
    """stylegan2_layers.py:603-611: one ResBlock per octave from `size` down to 4x4."""
    log_size = int(math.log(2 ** int(round(math.log(size, 2))), 2))
    out = []
    for i in range(log_size, 2, -1):
        out.append(("convs." + (str(9 - i) if i <= 8 else "%dx%d" % (2 ** i, 2 ** i)), D_CHANNELS[2 ** i], D_CHANNELS[2 ** (i - 1)]))
    return out


def shapes_D(size):
    sh = {}
    c0 = D_CHANNELS[2 ** int(round(math.log(size, 2)))]
    sh["convs.0.Conv.weight"] = (c0, 3, 1, 1)
    sh["convs.0.Act.bias"] = (c0,)
    c = c0
    for name, ci, co in d_block_names(size):
        sh[name + ".conv1.Conv.weight"] = (ci, ci, 3, 3)
        sh[name + ".conv1.Act.bias"] = (ci,)
        sh[name + ".conv2.Conv.weight"] = (co, ci, 3, 3)
        sh[name + ".conv2.Act.bias"] = (co,)
        sh[name + ".skip.Conv.weight"] = (co, ci, 1, 1)
        c = co
    sh["final_conv.Conv.weight"] = (512, c, 3, 3)
    sh["final_conv.Act.bias"] = (512,)
    sh["final_linear.0.weight"] = (512, 512 * 16)
    sh["final_linear.0.bias"] = (512,)
    sh["final_linear.1.weight"] = (1, 512)
    sh["final_linear.1.bias"] = (1,)
    return sh


def make_discriminator(size, seed=0, prefix="stylegan2_D."):
    """Seeded synthetic discriminator weights under the reference's state_dict keys (N(0,1) weights with the runtime
    equalised-lr scaling, small non-zero biases)."""
    g = torch.Generator().manual_seed(seed * 3 + 14)
    sd = {}
    shapes = shapes_D(size)
    for k in sorted(shapes):
        sd[prefix + k] = (0.1 * torch.randn(shapes[k], generator=g)) if k.endswith("bias") else torch.randn(shapes[k], generator=g)
    for name, _, _ in d_block_names(size):
        for l in ("conv2", "skip"):
            sd[prefix + name + "." + l + ".Blur.kernel"] = make_kernel([1, 3, 3, 1])
    return sd


def _shapes_encoder(prefix_down, with_in_tail):
    sh = {"FromRGB.Conv.weight": (32, 3, 1, 1), "FromRGB.Act.bias": (32,)}
    for i, (ci, co) in zip((1, 2, 4), ((32, 64), (64, 128), (128, 256))):
        p = "%s.ResBlockDownBy%d." % (prefix_down, i)
        sh[p + "conv1.Conv.weight"] = (ci, ci, 3, 3)
        sh[p + "conv1.Act.bias"] = (ci,)
        sh[p + "conv2.Blur.kernel"] = (3, 3)
        sh[p + "conv2.Conv.weight"] = (co, ci, 3, 3)
        sh[p + "conv2.Act.bias"] = (co,)
        sh[p + "skip.Blur.kernel"] = (3, 3)
        sh[p + "skip.Conv.weight"] = (co, ci, 1, 1)
    return sh


def shapes_E1():
    sh = _shapes_encoder("DownToSpatialCode", True)
    sh.update({"mlp_01.0.weight": (256, 32), "mlp_01.0.bias": (256,), "mlp_01.2.weight": (256, 256),
               "mlp_01.2.bias": (256,),
               "ToSpatialCode.0.Conv.weight": (256, 256, 1, 1), "ToSpatialCode.0.Act.bias": (256,),
               "ToSpatialCode.1.Conv.weight": (256, 256, 1, 1), "ToSpatialCode.1.Conv.bias": (256,)})
    return sh


def shapes_E2():
    sh = _shapes_encoder("DownToGlobalCode1", False)
    sh.update({"ToGlobalCode.0.weight": (2048, 256), "ToGlobalCode.0.bias": (2048,)})
    for lvl, c in (("9", 32), ("0", 64), ("1", 128), ("2", 256)):
        sh["conv1x1_%s.weight" % lvl] = (c, 2 * c, 1, 1)
        sh["conv1x1_%s.bias" % lvl] = (c,)
        for j, (o, i) in zip((1, 3, 5), ((1024, c), (2048, 1024), (2048, 2048))):
            sh["projector%s.%d.weight" % (lvl, j)] = (o, i)
            sh["projector%s.%d.bias" % (lvl, j)] = (o,)
    return sh


def shapes_G():
    sh = {}
    for n in ("scale", "bias"):
        sh["SpatialCodeModulation.%s.weight" % n] = (256, 2048)
        sh["SpatialCodeModulation.%s.bias" % n] = (256,)

    def styled(p, ci, co):
        sh[p + "bias"] = (1, co, 1, 1)
        sh[p + "conv.weight"] = (co, ci, 3, 3)
        sh[p + "conv.bias"] = (co,)
        sh[p + "epi1.style_mod.lin.weight"] = (2 * co, 2048)
        sh[p + "epi1.style_mod.lin.bias"] = (2 * co,)
        sh[p + "noise.weight"] = (1,)
        sh[p + "activate.bias"] = (co,)
    for i, (ci, co) in enumerate(HEAD_CH):
        p = "HeadResnetBlock%d." % i
        styled(p + "conv1.", ci, co)
        styled(p + "conv2.", co, co)
        if ci != co:
            sh[p + "skip.Conv.weight"] = (co, ci, 1, 1)
    for res, ci, co in UP_CH:
        p = "UpsamplingResBlock%d." % res
        styled(p + "conv1.", ci, co)
        styled(p + "conv2.", co, co)
        if ci != co:
            sh[p + "skip.Conv.weight"] = (co, ci, 1, 1)
            sh[p + "skip.Act.bias"] = (co,)
    sh.update({"ToRGB.bias": (1, 3, 1, 1), "ToRGB.upsample.kernel": (4, 4), "ToRGB.conv.weight": (3, 128, 1, 1),
               "ToRGB.conv.bias": (3,), "ToRGB.epi1.style_mod.lin.weight": (6, 2048),
               "ToRGB.epi1.style_mod.lin.bias": (6,)})
    for name, ci, k in (("layer32", 512, 3), ("layer64", 512, 3), ("layer128", 256, 3), ("layer256", 128, 1)):
        mid = 128 if k == 3 else 64
        sh[name + ".2.weight"] = (mid, ci, k, k)
        sh[name + ".2.bias"] = (mid,)
        sh[name + ".4.weight"] = (1,)
        sh[name + ".6.weight"] = (64, mid, k, k)
        sh[name + ".6.bias"] = (64,)
        sh[name + ".8.weight"] = (1,)
    for p in ("layert.0.", "layert.1.", "layert.2.", "layert1.0."):
        sh[p + "conv1.weight"] = (256, 256, 3, 3)
        sh[p + "conv1.bias"] = (256,)
        sh[p + "prelu.weight"] = (1,)
        sh[p + "conv2.weight"] = (256, 256, 3, 3)
        sh[p + "conv2.bias"] = (256,)
    sh["layert1.1.weight"] = (64, 256, 1, 1)
    sh["layert1.1.bias"] = (64,)
    return sh


def _fan_in(shape):
    n = 1
    for s in shape[1:]:
        n *= s
    return max(n, 1)


def make_state_dict(shapes, seed):
    """Deterministic synthetic weights keyed like the reference state_dict.  Statistics follow the
    reference initialisers (EqualConv2d/EqualLinear: N(0,1) with runtime scaling; EqualizedConv2d /
    EqualizedLinear / nn.Conv2d / nn.Linear: ~He-scaled), but biases, noise weights and PReLU slopes
    are made NON-zero so that every term of the epilogue is exercised by the parity tests."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k in sorted(shapes):
        shp = shapes[k]
        if k.endswith("Blur.kernel"):
            sd[k] = make_kernel([1, 2, 1])
        elif k.endswith("upsample.kernel"):
            sd[k] = make_kernel([1, 3, 3, 1]) * 4
        elif k.endswith("noise.weight"):
            sd[k] = torch.full(shp, 0.1) * (1 + torch.rand(shp, generator=g))
        elif k.endswith("prelu.weight") or (k.startswith("layer") and k.endswith((".4.weight", ".8.weight"))):
            sd[k] = torch.full(shp, 0.25) + 0.1 * torch.rand(shp, generator=g)
        elif k.endswith("bias"):
            sd[k] = 0.1 * torch.randn(shp, generator=g)
        elif k.endswith("weight"):
            w = torch.randn(shp, generator=g)
            equalized_runtime = (k.endswith("Conv.weight") or k.startswith("SpatialCodeModulation")
                                 or k.startswith("ToGlobalCode") or k == "ToRGB.conv.weight"
                                 or k.endswith("style_mod.lin.weight"))
            if not equalized_runtime:
                gain = SQRT2 if (k.endswith("conv.weight") and not k.startswith("ToRGB")) else 1.0
                w = w * (gain / math.sqrt(_fan_in(shp)))
            sd[k] = w
        else:
            raise KeyError(k)
    return sd


def make_networks(seed=0):
    return (make_state_dict(shapes_E1(), seed * 3 + 11), make_state_dict(shapes_E2(), seed * 3 + 12),
            make_state_dict(shapes_G(), seed * 3 + 13))


def make_portraits(batch, h, w, seed=1234, smooth=False):
    """Uniform [-1,1] synthetic portraits (matches Normalize(.5,.5) range); `smooth` = bicubic-upsampled
    32x32 noise so instance-norm statistics are portrait-like."""
    g = torch.Generator().manual_seed(seed)
    if smooth:
        base = torch.rand(batch, 3, 32, 32, generator=g) * 2 - 1
        return F.interpolate(base, (h, w), mode="bicubic", align_corners=False).clamp(-1, 1)
    return torch.rand(batch, 3, h, w, generator=g) * 2 - 1


def make_label_map(batch, h, w, seed=77, blocky=True):
    """uint8 label map in {0,1,2}: nested ellipses (blocky) or iid."""
    g = torch.Generator().manual_seed(seed)
    if not blocky:
        return torch.randint(0, 3, (batch, h, w), generator=g, dtype=torch.uint8)
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, h), torch.linspace(-1, 1, w), indexing="ij")
    out = torch.zeros(batch, h, w, dtype=torch.uint8)
    for b in range(batch):
        cx, cy = (torch.rand(2, generator=g) - 0.5) * 0.3
        r = ((xx - cx) / 0.55) ** 2 + ((yy - cy) / 0.75) ** 2
        out[b][r < 1.0] = 2
        out[b][r < 0.45] = 1
    return out


def make_noises(batch, h, w, seed=99):
    """Explicit NoiseInjection inputs for every StyledConv of G at input size (h, w): key '<block>.conv{1,2}'."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    sh, sw = h // 8, w // 8
    for i in range(4):
        for c in ("conv1", "conv2"):
            out["HeadResnetBlock%d.%s" % (i, c)] = torch.randn(batch, 1, sh, sw, generator=g)
    for j, res in enumerate((16, 32, 64)):
        f = 2 ** (j + 1)
        for c in ("conv1", "conv2"):
            out["UpsamplingResBlock%d.%s" % (res, c)] = torch.randn(batch, 1, sh * f, sw * f, generator=g)
    return out
