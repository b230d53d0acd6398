"""Training stage 2 of the reference (its default, models/ppst_model.py:24): the generator and discriminator losses of
`PPSTModel.compute_generator_losses` / `compute_discriminator_losses` (ppst_model.py:105-137,158-236) -- correlation-warped
mix, cycle warp, mask warp, region style contrast -- on the differentiable operators of ppst_b200/train.py.

What stage 2 adds to stage 1 (train_nets.py), all differentiable here:
  * G(extract_features=True): the four feature heads (generator.py:180-231, fed x.detach()), their pooled / bilinear
    concatenations, layert / layert1 (ResidualBlock, generator.py:10-32);
  * PPSTModel.Rselfcorr / corrm / warp (ppst_model.py:330-387) and StyleGAN2ResnetEncodercol.warp with the mask branch
    (encoder_col.py:100-138,150-251, incl. its detach points: only the FIRST level's warp sees the live matrix);
  * the correlation GEMMs (4096 x 4096 x 512 logits at softmax temperature 0.01, corr @ features) forward AND backward as
    3-term bf16-split tensor-core GEMMs (`_SplitGemm`: fp32-accurate, ppst_conv_tc in split mode), no library GEMM;
  * the region style contrastive loss with its queues (ppst_b200/rscl.py).
Convolutions, InstanceNorm (+ StyleMod), the global poolings and the thin RGB layers are the tcgen05 / fused kernels of
train.py; replicate padding, PReLU, the bilinear / average resamplings of the 64-channel head outputs, softmax, unfold / fold
of the 3-channel image warp and the [n, c]-sized arithmetic are ATen glue (activation-sized ones on channels-last views).

LPIPS (lpips.LPIPS(net='alex'), ppst_model.py:48,178) needs weights that are not available offline: `lpips_fn` defaults to the
same mean |a - b| stand-in the reference-run fixture was made with, and can be replaced by the real module.
"""
import sys

import torch
import torch.nn.functional as F
from torch.autograd import Function

from . import ops, train
from . import train_nets as TN
from .rscl import RsclLoss
from .train_nets import HEAD_CH, UP_CH, ParamSet, _eq

SQRT2 = train.SQRT2
_EPS = sys.float_info.epsilon
LEVELS = ["9", "0", "1", "2"]


def swap_pairs(x):
    """PPSTModel.swap (ppst_model.py:59-66): flip adjacent pairs of the minibatch."""
    s = x.shape
    assert s[0] % 2 == 0, "Minibatch size must be a multiple of 2"
    return torch.flip(x.reshape(s[0] // 2, 2, *s[1:]), [1]).reshape(*s)


# ------------------------------------------------------------------------------------------------ split tensor-core GEMM
def _pad_to(t, dim, mult):
    r = (-t.shape[dim]) % mult
    if r == 0:
        return t
    pad = [0, 0] * t.dim()
    pad[2 * (t.dim() - 1 - dim) + 1] = r
    return F.pad(t, pad)


def _split_gemm_raw(a, b):
    """C[b] = A[b] . B[b]^T at fp32 accuracy on the tensor cores: a f32 [B, M, K], b f32 [B, N, K] -> f32 [B, M, N].
    Operands are zero-padded to the kernel's granularity (K: 64, N: 32, M: 64)."""
    bt, M, K = a.shape
    N = b.shape[1]
    a = _pad_to(_pad_to(a.to(torch.float32), 2, 64), 1, 64).contiguous()
    b = _pad_to(_pad_to(b.to(torch.float32), 2, 64), 1, 32).contiguous()
    asp, bsp = ops.split_bf16(a), ops.split_bf16(b)
    Mp = a.shape[1]
    out = ops.gemm_split_tc(asp, bsp, (Mp // 64, 64))
    return out[:, :M, :N]


class _SplitGemm(Function):
    """C = A . B^T (batched), forward and both gradients as 3-term bf16-split tcgen05 GEMMs:
    dA = dC . B, dB = dC^T . A."""

    @staticmethod
    def forward(ctx, a, b):
        ctx.save_for_backward(a, b)
        return _split_gemm_raw(a, b).contiguous()

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gc):
        a, b = ctx.saved_tensors
        gc = gc.to(torch.float32)
        da = db = None
        if ctx.needs_input_grad[0]:
            da = _split_gemm_raw(gc, b.transpose(1, 2)).contiguous()            # [M, N] . ([K, N])^T
        if ctx.needs_input_grad[1]:
            db = _split_gemm_raw(gc.transpose(1, 2), a.transpose(1, 2)).contiguous()   # [N, M] . ([K, M])^T
        return da, db


def split_matmul_nt(a, b):
    return _SplitGemm.apply(a, b)


# ------------------------------------------------------------------------------------------------ small NHWC helpers
def _nchw(x_nhwc):
    return x_nhwc.permute(0, 3, 1, 2)                 # channels-last view for ATen glue


def _nhwc(x_nchw):
    return x_nchw.permute(0, 2, 3, 1).contiguous()


def replicate_pad(x, p=1):
    """nn.ReplicationPad2d(p) on NHWC (generator.py:12-26,180-231)."""
    n, h, w, c = x.shape
    iy = torch.arange(-p, h + p, device=x.device).clamp_(0, h - 1)
    ix = torch.arange(-p, w + p, device=x.device).clamp_(0, w - 1)
    return x.index_select(1, iy).index_select(2, ix)


class _PReLU(Function):
    """nn.PReLU() with its single learnable slope.  The slope's gradient sum_{x<0} g*x is a sum of millions of signed terms
    with heavy cancellation: it is accumulated in fp32 from fp32 products (autograd's bf16 elementwise product would round
    every term first)."""

    @staticmethod
    def forward(ctx, x, a):
        ctx.save_for_backward(x, a)
        return torch.where(x >= 0, x, x * a.to(x.dtype))

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        x, a = ctx.saved_tensors
        neg = x < 0
        gx = torch.where(neg, g * a.to(g.dtype), g) if ctx.needs_input_grad[0] else None
        ga = None
        if ctx.needs_input_grad[1]:
            ga = torch.where(neg, g.to(torch.float32) * x.to(torch.float32), torch.zeros((), device=x.device)).sum().reshape(a.shape)
        return gx, ga


def prelu(x, a):
    return _PReLU.apply(x, a)


# ------------------------------------------------------------------------------------------------ generator with features
def _feature_head(P, name, x, k):
    """layer32 / layer64 / layer128 (3x3, ReplicationPad2d(1)) and layer256 (1x1), generator.py:180-231; the first
    InstanceNorm runs on the PADDED tensor; fed x.detach() (generator.py:257,268)."""
    x = x.detach()
    if k == 3:
        x = replicate_pad(x)
    x = train.instance_norm_mod(x)
    y, st = train.conv2d(x, P[name + ".2.weight"], 1, 0, bias=P[name + ".2.bias"], want_stats=True)
    y = prelu(train.instance_norm_mod(y, stats=st), P[name + ".4.weight"])
    if k == 3:
        y = replicate_pad(y)
    y, st = train.conv2d(y, P[name + ".6.weight"], 1, 0, bias=P[name + ".6.bias"], want_stats=True)
    return prelu(train.instance_norm_mod(y, stats=st), P[name + ".8.weight"])


def _residual_block(P, p, x):
    """ResidualBlock (generator.py:10-32): one PReLU shared by both activations."""
    a = P[p + "prelu.weight"]
    y, st = train.conv2d(replicate_pad(x), P[p + "conv1.weight"], 1, 0, bias=P[p + "conv1.bias"], want_stats=True)
    y = prelu(train.instance_norm_mod(y, stats=st), a)
    y, st = train.conv2d(replicate_pad(y), P[p + "conv2.weight"], 1, 0, bias=P[p + "conv2.bias"], want_stats=True)
    return prelu(train.instance_norm_mod(y, stats=st) + x, a)


def generator(P, sp, codes, noise_gen=None, noises=None, extract_features=False):
    """StyleGAN2ResnetGenerator.forward (generator.py:244-281).  extract_features: -> (rgb NCHW f32, feat NHWC bf16
    [n, h, w, 256], feat1 NHWC bf16 [n, 256, 256, 64])."""
    import math
    codes = [c * torch.rsqrt(torch.sum(c * c, dim=1, keepdim=True) + 1e-8) for c in codes]       # util.normalize
    g = codes[-1]

    def eql(name):
        w = P["SpatialCodeModulation.%s.weight" % name]
        return train.linear(g, w * (1.0 / math.sqrt(w.shape[1]))) + P["SpatialCodeModulation.%s.bias" % name]
    x = train.channel_affine(sp, eql("scale"), eql("bias"))
    for i, (cin, cout) in enumerate(HEAD_CH):
        p = "HeadResnetBlock%d." % i
        skip = train.conv2d(x, _eq(P[p + "skip.Conv.weight"]), 1, 0) if cin != cout else x
        r = TN._styled(P, p + "conv1.", x, g, False, noise_gen, noises)
        r = TN._styled(P, p + "conv2.", r, g, False, noise_gen, noises)
        x = train.add_scale(skip, r, 1.0 / SQRT2)
    feas = []
    if extract_features:
        pool_hw = (x.shape[1], x.shape[2])
        feas.append(_feature_head(P, "layer32", x, 3))
    for j, (res, cin, cout) in enumerate(UP_CH):
        p = "UpsamplingResBlock%d." % res
        code = codes[-2 - j]
        skip = train.conv2d(x, _eq(P[p + "skip.Conv.weight"]), 1, 0, bias=P[p + "skip.Act.bias"], act=True) if cin != cout else x
        skip = train.bilinear2x(skip)
        r = TN._styled(P, p + "conv1.", x, code, True, noise_gen, noises)
        r = TN._styled(P, p + "conv2.", r, code, False, noise_gen, noises)
        x = train.add_scale(skip, r, 1.0 / SQRT2)
        if extract_features:
            feas.append(_feature_head(P, "layer%d" % (2 ** (j + 6)), x, 3 if j < 2 else 1))
    w = _eq(P["ToRGB.conv.weight"])
    rgb = train._ThinOut.apply(x, w.reshape(w.shape[0], w.shape[1]).t())                 # NCHW f32 [n, 3, H, W]
    rgb = rgb + (P["ToRGB.conv.bias"] + P["ToRGB.bias"].view(-1)).view(1, -1, 1, 1)
    mu = rgb.mean(dim=(2, 3), keepdim=True)
    var = ((rgb - mu) ** 2).mean(dim=(2, 3), keepdim=True)
    rgb = (rgb - mu) * torch.rsqrt(var + 1e-5)
    lw = P["ToRGB.epi1.style_mod.lin.weight"]
    s = train.linear(codes[0], lw * (lw.shape[1] ** -0.5)) + P["ToRGB.epi1.style_mod.lin.bias"]
    rgb = rgb * (s[:, :3, None, None] + 1.0) + s[:, 3:, None, None]
    if not extract_features:
        return rgb
    # generator.py:272-280: pooled / bilinear concatenations (ATen on channels-last views of the 64-channel head outputs)
    pooled = [feas[0]] + [_nhwc(F.adaptive_avg_pool2d(_nchw(f), pool_hw)) for f in feas[1:]]
    up = [_nhwc(F.interpolate(_nchw(f), (256, 256), mode="bilinear")) for f in feas]
    feat = torch.cat(pooled, dim=3)
    feat1 = torch.cat(up, dim=3)
    for i in range(3):
        feat = _residual_block(P, "layert.%d." % i, feat)
    feat1 = _residual_block(P, "layert1.0.", feat1)
    feat1 = train.conv2d(feat1, P["layert1.1.weight"], 1, 0, bias=P["layert1.1.bias"])
    return rgb, feat, feat1


# ------------------------------------------------------------------------------------------------ correlation
def rselfcorr(f):
    """PPSTModel.Rselfcorr (ppst_model.py:330-339) on NHWC: [b, 256, 256, 64] -> f32 NHWC [b, 64, 64, 256]."""
    b, h, w, c = f.shape
    ph, pw = h // 4, w // 4
    p = f.to(torch.float32).reshape(b, ph, 4, pw, 4, c).permute(0, 1, 3, 2, 4, 5).reshape(b, ph * pw, 16, c)
    p = p - p.mean(dim=3, keepdim=True)
    p = p / (p.pow(2).sum(dim=3, keepdim=True).sqrt() + _EPS)
    gram = torch.matmul(p, p.transpose(2, 3))                     # [b, L, 16, 16]: 16 x 16 x 64 per patch
    return gram.reshape(b, ph, pw, 256)


def _corr_prep(f):
    """rows [b, L, 512]: first 256 channels centred over the channel dim, all 512 L2-normalised (ppst_model.py:349-360)."""
    h1 = f[..., :256] - f[..., :256].mean(dim=-1, keepdim=True)
    f = torch.cat((h1, f[..., 256:]), dim=-1)
    return f / (f.pow(2).sum(dim=-1, keepdim=True).sqrt() + _EPS)


def corrm(fea, fea0):
    """PPSTModel.corrm (match_kernel == 1, ppst_model.py:341-364) on NHWC f32 [b, 64, 64, 512] -> f32 [b, L0, L]."""
    b = fea.shape[0]
    a = _corr_prep(fea.to(torch.float32).reshape(b, -1, fea.shape[3]))
    q = _corr_prep(fea0.to(torch.float32).reshape(b, -1, fea0.shape[3]))
    return torch.softmax(split_matmul_nt(q, a) / 0.01, dim=-1)


def model_warp(img, corr):
    """PPSTModel.warp (ppst_model.py:366-387) on an NCHW f32 image / mask whose pixel count is s^2 x the matrix size."""
    b, c, h, w = img.shape
    H = corr.shape[1]
    if H != h * w:
        s = int(((h * w) / H) ** 0.5)
        feas = F.unfold(img, s, stride=s)                         # [b, c*s*s, H]
        wf = split_matmul_nt(corr, feas)                          # [b, H, c*s*s]
        return F.fold(wf.permute(0, 2, 1), (h, w), s, stride=s)
    wf = split_matmul_nt(corr, img.reshape(b, c, -1))
    return wf.permute(0, 2, 1).reshape(b, c, h, w)


def e2_warp(x, corr, resize, scale):
    """StyleGAN2ResnetEncodercol.warp (encoder_col.py:100-138, square inputs): NHWC bf16 level features -> NCHW f32."""
    n, h, w, c = x.shape
    xf = _nchw(x).to(torch.float32)
    if resize:
        xf = F.adaptive_avg_pool2d(xf, (64, 64))
    wf = split_matmul_nt(corr, xf.reshape(n, c, -1))              # [n, 4096, c]
    wf = wf.permute(0, 2, 1).reshape(n, c, 64, 64)
    if resize:
        wf = F.interpolate(wf, scale_factor=scale, mode="bilinear")
    return wf


# ------------------------------------------------------------------------------------------------ colour encoder, stage 2
def _code_head(P, level, gap, gmp):
    v = torch.cat([gap, gmp], 1)
    w = P["conv1x1_%s.weight" % level]
    v = train.linear(v, w.view(w.shape[0], -1)) + P["conv1x1_%s.bias" % level]
    for j in (1, 3, 5):
        v = train.linear(F.relu(v), P["projector%s.%d.weight" % (level, j)]) + P["projector%s.%d.bias" % (level, j)]
    return v / v.norm(dim=1, keepdim=True).clamp_min(1e-12)


def encoder_col(P, img, corr=None, mask=None):
    """StyleGAN2ResnetEncodercol.forward (encoder_col.py:150-251) with the correlation and mask branches.
    mask: one-hot f32 NCHW [n, 3, H, W].  -> (vectors, vectors_w) or (vectors, projections_m, vectors_w, projections_mw)."""
    x = train.conv1x1_from_image(img, _eq(P["FromRGB.Conv.weight"]), P["FromRGB.Act.bias"], act=True)
    vectors, vectors_w, pm, pmw = [], [], [], []
    for li, level in enumerate(LEVELS):
        if li > 0:
            x = TN._res_block(P, "DownToGlobalCode1.ResBlockDownBy%d." % (2 ** (li - 1)), x, norm=False)
        vectors.append(_code_head(P, level, *train.gap_gmp(x)))
        xx = None
        if corr is not None:
            # encoder_col.py:168 warps the first level with the live matrix, :196 the others with corrmatrix.detach()
            xx = e2_warp(x, corr if li == 0 else corr.detach(), resize=(li < 3), scale=2 ** (3 - li))
            vectors_w.append(_code_head(P, level, xx.mean(dim=(2, 3)), xx.amax(dim=(2, 3))))
        if mask is not None:
            if li > 0:
                mask = F.max_pool2d(mask, 2, stride=2)
            mh = mask.permute(0, 2, 3, 1).to(x.dtype)             # NHWC, exact in bf16 (0 / 1)
            msw = swap_pairs(mask) if corr is not None else None
            for i in range(3):
                pm.append(_code_head(P, level, *train.gap_gmp(x * mh[..., i:i + 1])))
                if corr is not None:
                    xm = xx * msw[:, i:i + 1]
                    pmw.append(_code_head(P, level, xm.mean(dim=(2, 3)), xm.amax(dim=(2, 3))))
    if mask is not None:
        return vectors, pm, vectors_w, pmw
    return vectors, vectors_w


# ------------------------------------------------------------------------------------------------ the two steps
def _lpips_stub(a, b):
    return (a - b).abs().mean()


class Stage2(torch.nn.Module):
    """E1, E2, G, D, the region-contrast criterion and the two loss functions of training stage 2."""

    def __init__(self, sdE1, sdE2, sdG, D, lambda_l1=3.0, lambda_gan=1.0, lambda_stylecon=1.0, lambda_maskwarp=10.0,
                 lambda_cycwarp=5.0, nce_T=0.07, noise_seed=None, lpips_fn=None):
        super().__init__()
        self.E1, self.E2, self.G, self.D = ParamSet(sdE1), ParamSet(sdE2), ParamSet(sdG), D
        self.criterionNCE = RsclLoss(nce_T=nce_T)
        self.lam = {"L1": lambda_l1, "GAN": lambda_gan, "StyleCon": lambda_stylecon, "Maskwarp": lambda_maskwarp,
                    "Cycwarp": lambda_cycwarp}
        self.noise_seed = noise_seed
        self._gen = None
        self.lpips_fn = lpips_fn or _lpips_stub

    def _noise(self, device):
        if self.noise_seed is None:
            return None
        if self._gen is None:
            self._gen = torch.Generator(device=device)
            self._gen.manual_seed(self.noise_seed)
        return self._gen

    def _features(self, sp, gl, ng):
        """ppst_model.py:164-166: sps = cat(feas, Rselfcorr(feas1)) as NHWC f32 [b, 64, 64, 512]."""
        _, feas, feas1 = generator(self.G, sp, gl, ng, extract_features=True)
        return torch.cat((feas.to(torch.float32), rselfcorr(feas1)), dim=3)

    def generator_losses(self, real, mask, enqueue=True):
        """compute_generator_losses, training_stage == 2 (ppst_model.py:158-236).  real NCHW f32 [B,3,512,512], mask
        one-hot f32 [B,3,512,512].  Returns the reference's loss dict; the queues are updated like the reference does (one
        fused all_gather for the 24 keys)."""
        lam, ng = self.lam, self._noise(real.device)
        losses = {}
        B = real.shape[0]
        sp = TN.encoder_con(self.E1, real)
        gl, _ = encoder_col(self.E2, real)
        sps = self._features(sp, gl, ng)
        corr = corrm(sps, swap_pairs(sps))
        corr_self = corrm(sps, sps)
        _, gl = encoder_col(self.E2, real, corr=corr_self)
        _, pro_ms, gl_w, pro_mw = encoder_col(self.E2, real, corr=corr, mask=mask)
        image_warp = model_warp(real, corr)
        image_rec = model_warp(image_warp, swap_pairs(corr))
        losses["image_warp_reg"] = self.lpips_fn(image_rec, real) * lam["Cycwarp"]
        losses["Mask_warp"] = (model_warp(mask, corr) - swap_pairs(mask)).abs().mean() * lam["Maskwarp"]
        rec = generator(self.G, sp, gl, ng)
        losses["G_L1"] = (rec - real).abs().mean() * lam["L1"]
        mix = generator(self.G, swap_pairs(sp), gl_w, ng)
        _, pro_3m, _, _ = encoder_col(self.E2, mix, mask=swap_pairs(mask))
        _, pro_2m, _, _ = encoder_col(self.E2, rec, mask=mask)
        sp_3 = TN.encoder_con(self.E1, mix)
        cyc = generator(self.G, swap_pairs(sp_3)[:B // 2], [g[:B // 2] for g in gl], ng)
        losses["G_L1_cyc"] = (cyc - real[:B // 2]).abs().mean() * 3
        style1 = style2 = 0.0
        pushes = []
        for lvl in range(4):
            key0 = torch.cat(pro_ms[3 * lvl:3 * lvl + 3], dim=0)
            keyw = torch.cat(pro_mw[3 * lvl:3 * lvl + 3], dim=0)
            query = torch.cat(pro_3m[3 * lvl:3 * lvl + 3], dim=0)
            query_r = torch.cat(pro_2m[3 * lvl:3 * lvl + 3], dim=0)
            style1 = style1 + self.criterionNCE(query, keyw.detach(), key0.detach(), lvl)
            style2 = style2 + self.criterionNCE(query_r, key0.detach(), keyw.detach(), lvl)
            pushes += [(key0[i:i + 1].detach(), lvl) for i in range(3)] + [(keyw[i:i + 1].detach(), lvl) for i in range(3)]
        self.last_pushes = pushes                    # with enqueue=False (CUDA-graph capture: the queue pointers live on the
        if enqueue:                                  # host) the caller runs criterionNCE.dequeue_and_enqueue_many(last_pushes)
            # the reference pushes level l's keys right after level l's losses (ppst_model.py:211-218); the levels' queues are
            # independent, so pushing all 24 keys afterwards in one collective leaves the same queues
            self.criterionNCE.dequeue_and_enqueue_many(pushes)
        losses["G_styleContmix"] = style1 * lam["StyleCon"]
        losses["G_styleContrec"] = style2 * lam["StyleCon"]
        if lam["GAN"] > 0:
            losses["G_GAN_rec"] = train.gan_loss(self.D(rec), True) * (lam["GAN"] * 0.5)
            losses["G_GAN_mix"] = train.gan_loss(self.D(mix), True) * (lam["GAN"] * 1.0)
        return losses

    def discriminator_losses(self, real, mask=None):
        """compute_discriminator_losses, training_stage == 2 (ppst_model.py:105-137,68-92): D_real + D_rec (half of the
        batch, reconstructed with the SELF-correlation codes) + D_mix; E1 / E2 / G run without a tape."""
        lam, ng = self.lam, self._noise(real.device)
        B = real.shape[0]
        with torch.no_grad():
            sp = TN.encoder_con(self.E1, real)
            gl, _ = encoder_col(self.E2, real)
            sps = self._features(sp, gl, ng)
            corr = corrm(sps, swap_pairs(sps))
            corr_self = corrm(sps, sps)
            _, gl_w = encoder_col(self.E2, real, corr=corr)
            mix = generator(self.G, swap_pairs(sp), gl_w, ng)
            _, gl = encoder_col(self.E2, real, corr=corr_self)
            rec = generator(self.G, sp[:B // 2], [g[:B // 2] for g in gl], ng)
        return {"D_real": train.gan_loss(self.D(real), True) * lam["GAN"],
                "D_rec": train.gan_loss(self.D(rec), False) * (0.5 * lam["GAN"]),
                "D_mix": train.gan_loss(self.D(mix), False) * (0.5 * lam["GAN"])}

    def g_params(self):
        return list(self.E1.parameters()) + list(self.E2.parameters()) + list(self.G.parameters())
