"""Differentiable content encoder, colour encoder and generator for the training step (BASELINE config 5,
training stage 1: `PPSTModel.compute_generator_losses` / `compute_discriminator_losses`, models/ppst_model.py:113-137,
158-236 with training_stage == 1).

Same networks as the inference engine, but laid out for autograd: parameters are torch Parameters under the REFERENCE's
state_dict keys (E1.* / E2.* / G.*), activations are NHWC bf16, and every activation-sized operator
is a differentiable wrapper around the C-ABI kernels (ppst_b200/train.py): tcgen05 convolution forward / input gradient /
weight gradient (plain, strided with the blur folded in, transposed for the fused upscale), InstanceNorm + StyleMod as
statistics + one apply pass each way, the thin RGB layers, the residual joins, the x2 upsamplings and the reflection pads
with their adjoints, the global average / max pooling of the colour encoder's heads.  What still runs as ATen kernels:
the ToRGB instance norm on the 3-channel image and everything that is [n, c]- or parameter-sized
(style linears' bias adds, weight scaling, normalisations, losses, Adam).

Covers what training stage 1 executes: E1, E2 without the correlation / mask branches, G without feature extraction.
"""
import math

import torch
import torch.nn.functional as F

from . import train
from .train import SQRT2

HEAD_CH = [(256, 256), (256, 256), (256, 384), (384, 512)]     # generator.py:156-165
UP_CH = [(16, 512, 512), (32, 512, 256), (64, 256, 128)]       # generator.py:167-175


class ParamSet(torch.nn.Module):
    """Parameters of one reference network under its own state_dict keys ('.' <-> '/' for ParameterDict)."""

    def __init__(self, sd, trainable=lambda k: not k.endswith(".kernel")):
        super().__init__()
        self.P = torch.nn.ParameterDict()
        self.B = {}
        for k, v in sd.items():
            if trainable(k) and v.is_floating_point():
                self.P[k.replace(".", "/")] = torch.nn.Parameter(v.detach().clone().to(torch.float32))
            else:
                self.B[k] = v.detach().clone()

    def __getitem__(self, k):
        kk = k.replace(".", "/")
        return self.P[kk] if kk in self.P else self.B[k]

    def __contains__(self, k):
        return k.replace(".", "/") in self.P or k in self.B

    def _apply(self, fn, *a, **kw):
        self.B = {k: fn(v) for k, v in self.B.items()}
        return super()._apply(fn, *a, **kw)

    def reference_state_dict(self):
        out = {k.replace("/", "."): v.detach().clone() for k, v in self.P.items()}
        out.update({k: v.clone() for k, v in self.B.items()})
        return out


_consts = {}


def _const(name, device, make):
    """Small constant tensors (FIR kernels), created once per device -- not inside a CUDA-graph capture."""
    key = (name, str(device))
    t = _consts.get(key)
    if t is None:
        t = _consts[key] = make().to(device).contiguous()
    return t


def _k121(device):
    def make():
        k = torch.tensor([1.0, 2.0, 1.0])
        k = k[None, :] * k[:, None]
        return k / k.sum()
    return _const("k121", device, make)


def _eq(w):
    """EqualConv2d's run-time scale (stylegan2_layers.py:177)."""
    return w * (1.0 / math.sqrt(w.shape[1] * w.shape[2] * w.shape[3]))


# ------------------------------------------------------------------------------------------------ encoders
def _res_block(P, p, x, norm):
    """ResBlock with reflection padding and Blur([1,2,1]) (stylegan2_layers.py:559-579 as encoder_con.py / encoder_col.py
    build it).  conv2's Blur + stride-2 conv = ONE 5x5 stride-2 convolution on the reflect-padded input, the skip's
    Blur(zero pad (1,0)) + 1x1 stride-2 conv = ONE 3x3 stride-2 convolution with zero padding."""
    blur = _k121(x.device)

    def tail(ys, name, act):
        y, stats = ys
        b = P[p + name + ".Act.bias"] if act else None
        n, c = y.shape[0], y.shape[3]
        return train.instance_norm_mod(y, None, None if b is None else b.view(1, c).expand(n, c), act=act, stats=stats)

    # conv1: RefPad(1) -> 3x3 -> [IN] -> FusedLeakyReLU
    w1 = _eq(P[p + "conv1.Conv.weight"])
    xp = train.reflect_pad(x, (1, 1, 1, 1))
    if norm:
        y = tail(train.conv2d(xp, w1, 1, 0, want_stats=True), "conv1", True)
    else:
        y = train.conv2d(xp, w1, 1, 0, bias=P[p + "conv1.Act.bias"], act=True)
    # conv2: RefPad(2,1) -> Blur -> 3x3 stride 2; one more reflected row / column keeps the size even, no tap reads it
    w2 = train._fold(_eq(P[p + "conv2.Conv.weight"]), blur)                      # [co, ci, 5, 5]
    yp = train.reflect_pad(y, (2, 2, 2, 2))
    if norm:
        y = tail(train.conv2d(yp, w2, 2, 0, want_stats=True), "conv2", True)
    else:
        y = train.conv2d(yp, w2, 2, 0, bias=P[p + "conv2.Act.bias"], act=True)
    # skip: Blur(pad (1, 0)) -> 1x1 stride 2, no bias, no activation
    ws = train._fold(_eq(P[p + "skip.Conv.weight"]), blur)                       # [co, ci, 3, 3]
    if norm:
        sk, st = train.conv2d(x, ws, 2, 1, want_stats=True)
        sk = train.instance_norm_mod(sk, stats=st)
    else:
        sk = train.conv2d(x, ws, 2, 1)
    return train.add_scale(y, sk, 1.0 / SQRT2)


def encoder_con(P, img):
    """StyleGAN2ResnetEncodercon.forward (encoder_con.py:82-92): NCHW f32 image -> sp NHWC bf16 [n, h/8, w/8, 256]."""
    x = train.conv1x1_from_image(img, _eq(P["FromRGB.Conv.weight"]), P["FromRGB.Act.bias"], act=True)
    for i in (1, 2, 4):
        x = _res_block(P, "DownToSpatialCode.ResBlockDownBy%d." % i, x, norm=True)
    n, c = x.shape[0], x.shape[3]
    y, st = train.conv2d(x, _eq(P["ToSpatialCode.0.Conv.weight"]), 1, 0, want_stats=True)
    y = train.instance_norm_mod(y, None, P["ToSpatialCode.0.Act.bias"].view(1, -1).expand(n, -1), act=True, stats=st)
    y, st = train.conv2d(y, _eq(P["ToSpatialCode.1.Conv.weight"]), 1, 0, bias=P["ToSpatialCode.1.Conv.bias"], want_stats=True)
    return train.instance_norm_mod(y, stats=st)


def _head(P, level, x):
    """GAP || GMP -> conv1x1 -> projector MLP -> F.normalize (encoder_col.py:39-40,50-86,160-166)."""
    gap, gmp = train.gap_gmp(x)
    v = torch.cat([gap, gmp], 1)
    w = P["conv1x1_%s.weight" % level]
    v = train.linear(v, w.view(w.shape[0], -1)) + P["conv1x1_%s.bias" % level]
    for j in (1, 3, 5):
        v = train.linear(F.relu(v), P["projector%s.%d.weight" % (level, j)]) + P["projector%s.%d.bias" % (level, j)]
    return v / v.norm(dim=1, keepdim=True).clamp_min(1e-12)


def encoder_col(P, img):
    """StyleGAN2ResnetEncodercol.forward without correlation matrix / mask (encoder_col.py:150-170): four [n, 2048] codes."""
    x = train.conv1x1_from_image(img, _eq(P["FromRGB.Conv.weight"]), P["FromRGB.Act.bias"], act=True)
    out = []
    for li, level in enumerate(["9", "0", "1", "2"]):
        if li > 0:
            x = _res_block(P, "DownToGlobalCode1.ResBlockDownBy%d." % (2 ** (li - 1)), x, norm=False)
        out.append(_head(P, level, x))
    return out


# ------------------------------------------------------------------------------------------------ generator
def _styled(P, p, x, code, upsample, noise_gen, noises=None):
    """StyledConv (stylegan2_layers.py:439-475): conv (+ fused upscale) + noise + biases + leaky-ReLU in one launch, then
    InstanceNorm + StyleMod."""
    w = P[p + "conv.weight"]
    co = w.shape[0]
    bias = P[p + "conv.bias"] + P[p + "bias"].view(-1) + P[p + "activate.bias"]
    n, h, wd, _ = x.shape
    oh, ow = (2 * h, 2 * wd) if upsample else (h, wd)
    noise = None
    if noises is not None and p[:-1] in noises:                   # explicit NoiseInjection input [n, 1, oh, ow]
        noise = noises[p[:-1]].to(x.device, torch.float32).reshape(n, oh, ow) * P[p + "noise.weight"].view(())
    elif noise_gen is not None:
        noise = torch.randn((n, oh, ow), device=x.device, generator=noise_gen) * P[p + "noise.weight"].view(())
    if upsample and min(h, wd) * 2 >= 128:
        # W' = 2x2 box sum of the zero-padded transposed weight -> conv_transpose2d(stride 2, padding 1) (:312-321)
        wt = train._fold(w.permute(1, 0, 2, 3), _const("box2", w.device, lambda: torch.ones(2, 2)))   # [ci, co, 4, 4]
        y, st = train.conv_transpose2d(x, wt, 2, 1, (oh, ow), bias=bias, act=True, noise=noise, want_stats=True)
    else:
        if upsample:
            x = train.nearest2x(x)              # Upscale2d below the fused-upscale threshold (stylegan2_layers.py:322-323)
        y, st = train.conv2d(x, w, 1, w.shape[2] // 2, bias=bias, act=True, noise=noise, want_stats=True)
    lw = P[p + "epi1.style_mod.lin.weight"]
    s = train.linear(code, lw * (lw.shape[1] ** -0.5)) + P[p + "epi1.style_mod.lin.bias"]
    return train.instance_norm_mod(y, s[:, :co], s[:, co:], act=False, stats=st)


def generator(P, sp, codes, noise_gen=None, noises=None):
    """StyleGAN2ResnetGenerator.forward (generator.py:244-271, extract_features=False): sp NHWC bf16 [n, h, w, 256], four
    [n, 2048] codes -> NCHW f32 image [n, 3, 8h, 8w].  noise_gen: torch.Generator for NoiseInjection's fresh N(0,1)
    noise (:376-399), or None to skip the term (its weight is 0 at initialisation); noises: explicit maps
    {'<block>.conv{1,2}': [n, 1, H, W]} instead (parity tests)."""
    codes = [c * torch.rsqrt(torch.sum(c * c, dim=1, keepdim=True) + 1e-8) for c in codes]       # util.normalize
    g = codes[-1]

    def eql(name):
        w = P["SpatialCodeModulation.%s.weight" % name]
        return train.linear(g, w * (1.0 / math.sqrt(w.shape[1]))) + P["SpatialCodeModulation.%s.bias" % name]
    x = train.channel_affine(sp, eql("scale"), eql("bias"))
    for i, (cin, cout) in enumerate(HEAD_CH):
        p = "HeadResnetBlock%d." % i
        skip = train.conv2d(x, _eq(P[p + "skip.Conv.weight"]), 1, 0) if cin != cout else x
        r = _styled(P, p + "conv1.", x, g, False, noise_gen, noises)
        r = _styled(P, p + "conv2.", r, g, False, noise_gen, noises)
        x = train.add_scale(skip, r, 1.0 / SQRT2)
    for j, (res, cin, cout) in enumerate(UP_CH):
        p = "UpsamplingResBlock%d." % res
        code = codes[-2 - j]
        skip = train.conv2d(x, _eq(P[p + "skip.Conv.weight"]), 1, 0, bias=P[p + "skip.Act.bias"], act=True) if cin != cout else x
        skip = train.bilinear2x(skip)
        r = _styled(P, p + "conv1.", x, code, True, noise_gen, noises)
        r = _styled(P, p + "conv2.", r, code, False, noise_gen, noises)
        x = train.add_scale(skip, r, 1.0 / SQRT2)
    # ToRGB (stylegan2_layers.py:477-495): 1x1 conv 128 -> 3 (+ biases) -> IN -> StyleMod, on the 3-channel image
    w = _eq(P["ToRGB.conv.weight"])
    rgb = train._ThinOut.apply(x, w.reshape(w.shape[0], w.shape[1]).t())                 # NCHW f32 [n, 3, H, W]
    rgb = rgb + (P["ToRGB.conv.bias"] + P["ToRGB.bias"].view(-1)).view(1, -1, 1, 1)
    mu = rgb.mean(dim=(2, 3), keepdim=True)
    var = ((rgb - mu) ** 2).mean(dim=(2, 3), keepdim=True)
    rgb = (rgb - mu) * torch.rsqrt(var + 1e-5)
    lw = P["ToRGB.epi1.style_mod.lin.weight"]
    s = train.linear(codes[0], lw * (lw.shape[1] ** -0.5)) + P["ToRGB.epi1.style_mod.lin.bias"]
    return rgb * (s[:, :3, None, None] + 1.0) + s[:, 3:, None, None]


# ------------------------------------------------------------------------------------------------ training stage 1
class Stage1(torch.nn.Module):
    """E1, E2, G, D and the two loss functions of training stage 1 (ppst_model.py:113-137,158-236; lambda_StyleCon = 0,
    the only stage-1 setting that runs upstream)."""

    def __init__(self, sdE1, sdE2, sdG, D, lambda_l1=1.0, lambda_gan=1.0, noise_seed=None):
        super().__init__()
        self.E1, self.E2, self.G, self.D = ParamSet(sdE1), ParamSet(sdE2), ParamSet(sdG), D
        self.lambda_l1, self.lambda_gan = lambda_l1, lambda_gan
        self.noise_seed = noise_seed
        self._gen = None

    def _noise(self, device):
        if self.noise_seed is None:
            return None
        if self._gen is None:
            self._gen = torch.Generator(device=device)
            self._gen.manual_seed(self.noise_seed)
        return self._gen

    def reconstruct(self, real):
        sp = encoder_con(self.E1, real)
        gl = encoder_col(self.E2, real)
        return generator(self.G, sp, gl, self._noise(real.device))

    def generator_losses(self, real):
        """G_L1 + G_GAN_rec (ppst_model.py:184-186,226-230)."""
        rec = self.reconstruct(real)
        losses = {"G_L1": (rec - real).abs().mean() * self.lambda_l1}
        if self.lambda_gan > 0:
            losses["G_GAN_rec"] = train.gan_loss(self.D(rec), True) * (0.5 * self.lambda_gan)
        return losses, rec

    def discriminator_losses(self, real):
        """D_real + D_rec on half of the reconstructions (ppst_model.py:113-137)."""
        with torch.no_grad():
            half = real[: max(real.shape[0] // 2, 1)]
            rec = self.reconstruct(half)
        return train.d_losses(self.D, real, rec, self.lambda_gan)

    def g_params(self):
        return list(self.E1.parameters()) + list(self.E2.parameters()) + list(self.G.parameters())
