"""CPU oracle for the PPST swap / inference hot path.

TEST INFRASTRUCTURE ONLY -- never imported by the product package `ppst_b200`.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import it, and there only as the checker / the reported CPU baseline.

This is a *functional restatement* (plain fp32 PyTorch-on-CPU tensor arithmetic, no
nn.Module; autograd only as the differentiator of the restated discriminator, section D) of what the reference computes on the path that
BASELINE.json:north_star names.  Every function cites the reference lines it follows
(paths relative to /root/reference).  The dense arithmetic the reference delegates to
third-party code (ATen conv2d / linear / interpolate, torch 2.x) is delegated to the same
library here; everything the reference itself spells out (FIR resampling, bias-act, the
norm/modulation epilogues, pooling heads, the correlation swap) is written out explicitly.

Pinning: tests/golden/*.npz hold outputs of the UNMODIFIED reference modules
(imported from /root/reference by oracle/gen_golden.py, which is committed) on seeded
inputs; tests/test_oracle_golden.py checks this file against every one of them.  The
reference itself ships no tests, golden vectors or KATs, so those
reference-run fixtures are the pin.

Weights are addressed by the reference's own state_dict keys.
"""
import math
import sys

import numpy as np
import torch
import torch.nn.functional as F

SQRT2 = math.sqrt(2.0)


# --------------------------------------------------------------------------------------
# stylegan2_op operators
# --------------------------------------------------------------------------------------
def upfirdn2d(x, kernel, up=1, down=1, pad=(0, 0)):
    """models/networks/stylegan2_op/upfirdn2d.py:150-159 (entry) and :162-222 (native twin);
    CUDA semantics upfirdn2d_kernel.cu:53-137.  Same pad on x and y (the public signature)."""
    return upfirdn2d_full(x, kernel, up, up, down, down, pad[0], pad[1], pad[0], pad[1])


def upfirdn2d_full(x, kernel, up_x, up_y, down_x, down_y, px0, px1, py0, py1):
    """Zero-insert upsample -> pad (negative pad crops) -> FIR with the FLIPPED kernel (true
    convolution) -> decimate.  Written as an explicit tap loop (shift-and-add), not F.conv2d.
    out = (in*up + p0 + p1 - k)//down + 1   (upfirdn2d.py:107-108)."""
    b, c, in_h, in_w = x.shape
    kh, kw = kernel.shape
    x = x.to(torch.float32)
    up_img = x.new_zeros(b, c, in_h * up_y, in_w * up_x)
    up_img[:, :, ::up_y, ::up_x] = x
    ph, pw = in_h * up_y + py0 + py1, in_w * up_x + px0 + px1
    padded = x.new_zeros(b, c, max(ph, 0), max(pw, 0))
    # destination / source windows for (possibly negative) pads
    sy0, sy1 = max(-py0, 0), in_h * up_y - max(-py1, 0)
    sx0, sx1 = max(-px0, 0), in_w * up_x - max(-px1, 0)
    dy0, dx0 = max(py0, 0), max(px0, 0)
    if sy1 > sy0 and sx1 > sx0:
        padded[:, :, dy0:dy0 + (sy1 - sy0), dx0:dx0 + (sx1 - sx0)] = up_img[:, :, sy0:sy1, sx0:sx1]
    oh_full, ow_full = ph - kh + 1, pw - kw + 1
    acc = x.new_zeros(b, c, max(oh_full, 0), max(ow_full, 0))
    for ky in range(kh):
        for kx in range(kw):
            # F.conv2d(out, flip(kernel)) == sum_k padded[y+ky, x+kx] * kernel[kh-1-ky, kw-1-kx]
            acc += padded[:, :, ky:ky + oh_full, kx:kx + ow_full] * float(kernel[kh - 1 - ky, kw - 1 - kx])
    return acc[:, :, ::down_y, ::down_x].contiguous()


def fused_leaky_relu(x, bias, negative_slope=0.2, scale=SQRT2):
    """stylegan2_op/fused_act.py:89-96 (fallback branch) == fused_bias_act_kernel.cu:19-49 act=3 grad=0."""
    if bias is not None and bias.numel() > 0:
        x = x + bias.view(1, -1, *([1] * (x.dim() - 2)))
    return torch.where(x > 0, x, x * negative_slope) * scale


def fused_leaky_relu_backward(grad_out, out, negative_slope=0.2, scale=SQRT2):
    """fused_act.py:23-43 / kernel case 31: gate on the sign of the saved OUTPUT; grad_bias sums n,h,w."""
    gi = torch.where(out > 0, grad_out, grad_out * negative_slope) * scale
    dims = [0] + list(range(2, gi.dim()))
    return gi, gi.sum(dims)


# --------------------------------------------------------------------------------------
# layer helpers (models/networks/stylegan2_layers.py)
# --------------------------------------------------------------------------------------
def make_kernel(k):
    """stylegan2_layers.py:28-36."""
    k = torch.tensor(k, dtype=torch.float32)
    if k.dim() == 1:
        k = k[None, :] * k[:, None]
    return k / k.sum()


def instance_norm(x, eps=1e-5):
    """nn.InstanceNorm2d(affine=False, track_running_stats=False): biased variance over H*W."""
    mu = x.mean(dim=(2, 3), keepdim=True)
    var = ((x - mu) ** 2).mean(dim=(2, 3), keepdim=True)
    return (x - mu) / torch.sqrt(var + eps)


def equal_conv2d(x, w, b, stride=1, padding=0):
    """EqualConv2d, stylegan2_layers.py:167-202: weight * 1/sqrt(Cin*k*k)."""
    scale = 1.0 / math.sqrt(w.shape[1] * w.shape[2] * w.shape[3])
    return F.conv2d(x, w * scale, b, stride=stride, padding=padding)


def conv_layer(sd, p, x, ksize, downsample=False, blur_kernel=(1, 3, 3, 1), activate=True, norm=None,
               reflection_pad=False):
    """ConvLayer, stylegan2_layers.py:497-555: [Blur] -> [RefPad] -> EqualConv2d -> [IN] -> [Act].
    Conv bias exists only when `bias and not activate`; FusedLeakyReLU bias when `bias and activate`;
    both are discovered from the state_dict keys."""
    w = sd[p + "Conv.weight"]
    cb = sd.get(p + "Conv.bias")
    if downsample:
        pad = (len(blur_kernel) - 2) + (ksize - 1)
        pad0, pad1 = (pad + 1) // 2, pad // 2
        k = sd[p + "Blur.kernel"] if (p + "Blur.kernel") in sd else make_kernel(list(blur_kernel))
        if reflection_pad:                       # Blur.forward, :159-164
            x = F.pad(x, (pad0, pad1, pad0, pad1), mode="reflect")
            x = upfirdn2d(x, k, pad=(0, 0))
        else:
            x = upfirdn2d(x, k, pad=(pad0, pad1))
        x = equal_conv2d(x, w, cb, stride=2, padding=0)
    else:
        padding = ksize // 2
        if reflection_pad:
            if padding > 0:
                x = F.pad(x, (padding,) * 4, mode="reflect")
            padding = 0
        x = equal_conv2d(x, w, cb, stride=1, padding=padding)
    if norm == "in":
        x = instance_norm(x)
    if activate:
        ab = sd.get(p + "Act.bias")
        if ab is not None:
            x = fused_leaky_relu(x, ab)
        else:                                    # ScaledLeakyReLU, :350-359
            x = F.leaky_relu(x, 0.2) * SQRT2
    return x


def res_block(sd, p, x, norm=None, blur_kernel=(1, 2, 1)):
    """ResBlock, stylegan2_layers.py:559-579 as the encoders build it (reflection_pad=True, downsample)."""
    out = conv_layer(sd, p + "conv1.", x, 3, reflection_pad=True, norm=norm)
    out = conv_layer(sd, p + "conv2.", out, 3, downsample=True, blur_kernel=blur_kernel,
                     reflection_pad=True, norm=norm)
    skip = conv_layer(sd, p + "skip.", x, 1, downsample=True, blur_kernel=blur_kernel, activate=False,
                      norm=norm)
    return (out + skip) / SQRT2


def equalized_conv2d(x, w, b, upscale=False):
    """EqualizedConv2d, stylegan2_layers.py:275-348 (use_wscale=False => w_mul = b_mul = 1)."""
    k = w.shape[2]
    if upscale and min(x.shape[2:]) * 2 >= 128:
        wt = w.permute(1, 0, 2, 3)
        wt = F.pad(wt, [1, 1, 1, 1])
        wt = wt[:, :, 1:, 1:] + wt[:, :, :-1, 1:] + wt[:, :, 1:, :-1] + wt[:, :, :-1, :-1]
        x = F.conv_transpose2d(x, wt, stride=2, padding=(wt.size(-1) - 1) // 2)
        return x + b.view(1, -1, 1, 1)
    if upscale:                                  # Upscale2d nearest x2, :88-105
        x = x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)
    return F.conv2d(x, w, b, padding=k // 2)


def style_mod(x, style, lw, lb):
    """StyleMod, :361-374.  LayerEpilogue builds it with use_wscale=True (:417), so EqualizedLinear
    (:249-273, gain=1) multiplies the N(0,1) weight by w_mul = in_dim**-0.5 at run time; b_mul = 1."""
    s = F.linear(style, lw * (lw.shape[1] ** -0.5), lb).view(x.shape[0], 2, x.shape[1], 1, 1)
    return x * (s[:, 0] + 1.0) + s[:, 1]


def styled_conv(sd, p, x, style, upsample=False, noise=None):
    """StyledConv.forward, stylegan2_layers.py:467-475:
    conv -> + noise.weight*noise -> + bias -> fused_leaky_relu(activate.bias) -> IN -> StyleMod."""
    out = equalized_conv2d(x, sd[p + "conv.weight"], sd[p + "conv.bias"], upscale=upsample)
    nw = sd[p + "noise.weight"]
    if noise is not None:
        out = out + nw * noise
    else:
        assert float(nw.abs().max()) == 0.0, "non-zero noise weight needs an explicit noise tensor"
    out = out + sd[p + "bias"]
    out = fused_leaky_relu(out, sd[p + "activate.bias"])
    out = instance_norm(out)
    return style_mod(out, style, sd[p + "epi1.style_mod.lin.weight"], sd[p + "epi1.style_mod.lin.bias"])


def to_rgb(sd, p, x, style):
    """ToRGB.forward (skip is always None on this path), stylegan2_layers.py:477-495."""
    out = equal_conv2d(x, sd[p + "conv.weight"], sd[p + "conv.bias"])
    out = out + sd[p + "bias"]
    out = instance_norm(out)
    return style_mod(out, style, sd[p + "epi1.style_mod.lin.weight"], sd[p + "epi1.style_mod.lin.bias"])


def equal_linear(x, w, b, lr_mul=1.0):
    """EqualLinear (no activation), stylegan2_layers.py:205-247."""
    return F.linear(x, w * (1.0 / math.sqrt(w.shape[1]) * lr_mul), b * lr_mul)


def l2_normalize_rows(v):
    """util.normalize, util/util.py:18-22."""
    if isinstance(v, (list, tuple)):
        return [l2_normalize_rows(t) for t in v]
    return v * torch.rsqrt(torch.sum(v ** 2, dim=1, keepdim=True) + 1e-8)


# --------------------------------------------------------------------------------------
# E1 -- content encoder (models/networks/encoder_con.py:82-92)
# --------------------------------------------------------------------------------------
def encoder_con(sd, x, taps=None):
    x = conv_layer(sd, "FromRGB.", x, 1)
    if taps is not None:
        taps["FromRGB"] = x
    for i in (1, 2, 4):
        x = res_block(sd, "DownToSpatialCode.ResBlockDownBy%d." % i, x, norm="in")
        if taps is not None:
            taps["ResBlockDownBy%d" % i] = x
    x = conv_layer(sd, "ToSpatialCode.0.", x, 1, activate=True, norm="in")
    x = conv_layer(sd, "ToSpatialCode.1.", x, 1, activate=False, norm="in")
    return x


# --------------------------------------------------------------------------------------
# E2 -- colour encoder (models/networks/encoder_col.py:150-251, warp :100-138)
# --------------------------------------------------------------------------------------
def _gap_gmp_head(sd, level, x):
    """GAP || GMP -> conv1x1_<level> -> projector<level> (ReLU-Linear x3) -> F.normalize."""
    gap = x.mean(dim=(2, 3))
    gmp = x.amax(dim=(2, 3))
    v = torch.cat([gap, gmp], 1)
    w = sd["conv1x1_%s.weight" % level]
    v = F.linear(v, w.view(w.shape[0], -1), sd["conv1x1_%s.bias" % level])
    for j in (1, 3, 5):
        v = F.linear(F.relu(v), sd["projector%s.%d.weight" % (level, j)], sd["projector%s.%d.bias" % (level, j)])
    return v / v.norm(dim=1, keepdim=True).clamp_min(1e-12)


def e2_warp(fea, corr, resize=True, scale_factor=1):
    """StyleGAN2ResnetEncodercol.warp, encoder_col.py:100-138 (unfold=False branch)."""
    b, c, h, w = fea.shape
    if resize:
        size = (int(64 * h / w), 64) if h > w else (64, int(64 * w / h))
        feas = F.adaptive_avg_pool2d(fea, size).reshape(b, c, -1).permute(0, 2, 1)
        wf = torch.matmul(corr, feas)
        wf = wf.permute(0, 2, 1).reshape(b, c, -1, 64) if h > w else wf.permute(0, 2, 1).reshape(b, c, 64, -1)
        return F.interpolate(wf, scale_factor=scale_factor, mode="bilinear")
    f = fea.reshape(b, c, -1).permute(0, 2, 1)
    wf = torch.matmul(corr, f)
    return (wf.permute(0, 2, 1).reshape(b, c, -1, 64) if h > w
            else wf.permute(0, 2, 1).reshape(b, c, 64, -1)).contiguous()


def _swap_pairs(x):
    """encoder_col.py:140-148: flip adjacent batch pairs."""
    s = x.shape
    return torch.flip(x.view(s[0] // 2, 2, *s[1:]), [1]).view(*s)


def encoder_col(sd, x, corrmatrix=None, mask=None, taps=None):
    """Returns (vectors, vectors_w) or, with mask, (vectors, projections_m, vectors_w, projections_mw)."""
    if isinstance(corrmatrix, (list, tuple)):      # simple_swapping_evaluator.py:53 passes [corr]
        corrmatrix = corrmatrix[0]
    vectors, vectors_w, proj_m, proj_mw = [], [], [], []
    levels = ["9", "0", "1", "2"]
    x = conv_layer(sd, "FromRGB.", x, 1)
    xx = None
    for li, level in enumerate(levels):
        if li > 0:
            x = res_block(sd, "DownToGlobalCode1.ResBlockDownBy%d." % (2 ** (li - 1)), x, norm=None)
        if taps is not None:
            taps["level" + level] = x
        vectors.append(_gap_gmp_head(sd, level, x))
        if corrmatrix is not None:
            sf = 2 ** (3 - li)                      # 8, 4, 2, (1: resize=False)
            # encoder_col.py:168 warps the first level with the live matrix, :196 the others with corrmatrix.detach()
            xx = e2_warp(x, corrmatrix if li == 0 else corrmatrix.detach(), resize=(li < 3), scale_factor=sf)
            vectors_w.append(_gap_gmp_head(sd, level, xx))
        if mask is not None:
            if li > 0:
                mask = F.max_pool2d(mask, 2, stride=2)
            for i in range(3):
                proj_m.append(_gap_gmp_head(sd, level, x * mask[:, i:i + 1]))
                if corrmatrix is not None:
                    proj_mw.append(_gap_gmp_head(sd, level, xx * _swap_pairs(mask)[:, i:i + 1]))
    if mask is not None:
        return vectors, proj_m, vectors_w, proj_mw
    return vectors, vectors_w


# --------------------------------------------------------------------------------------
# G -- generator (models/networks/generator.py:244-281)
# --------------------------------------------------------------------------------------
HEAD_CH = [(256, 256), (256, 256), (256, 384), (384, 512)]     # generator.py:156-165, nf(0)=512
UP_CH = [(16, 512, 512), (32, 512, 256), (64, 256, 128)]       # :167-175


def _prelu(x, a):
    return torch.where(x >= 0, x, x * a)


def residual_block(sd, p, x):
    """generator.py:10-32 (one PReLU shared by both activations)."""
    a = sd[p + "prelu.weight"]
    out = F.pad(x, (1, 1, 1, 1), mode="replicate")
    out = F.conv2d(out, sd[p + "conv1.weight"], sd[p + "conv1.bias"])
    out = _prelu(instance_norm(out), a)
    out = F.pad(out, (1, 1, 1, 1), mode="replicate")
    out = F.conv2d(out, sd[p + "conv2.weight"], sd[p + "conv2.bias"])
    out = instance_norm(out) + x
    return _prelu(out, a)


def feature_head(sd, p, x, ksize):
    """layer32/64/128 (3x3, ReplicationPad(1)) and layer256 (1x1, ReflectionPad(0)), generator.py:180-231.
    NB the first InstanceNorm runs on the *padded* tensor."""
    pad = ksize // 2
    if pad:
        x = F.pad(x, (pad,) * 4, mode="replicate")
    x = instance_norm(x)
    x = F.conv2d(x, sd[p + "2.weight"], sd[p + "2.bias"])
    x = _prelu(instance_norm(x), sd[p + "4.weight"])
    if pad:
        x = F.pad(x, (pad,) * 4, mode="replicate")
    x = F.conv2d(x, sd[p + "6.weight"], sd[p + "6.bias"])
    return _prelu(instance_norm(x), sd[p + "8.weight"])


def generator(sd, sp, codes, extract_features=False, noises=None, taps=None):
    """StyleGAN2ResnetGenerator.forward.  `noises`: optional dict  '<block>.conv{1,2}' -> [B,1,H,W]."""
    noises = noises or {}
    codes = l2_normalize_rows(list(codes))
    g = codes[-1]
    x = sp * equal_linear(g, sd["SpatialCodeModulation.scale.weight"], sd["SpatialCodeModulation.scale.bias"])[:, :, None, None] \
        + equal_linear(g, sd["SpatialCodeModulation.bias.weight"], sd["SpatialCodeModulation.bias.bias"])[:, :, None, None]
    if taps is not None:
        taps["SpatialCodeModulation"] = x
    for i, (cin, cout) in enumerate(HEAD_CH):
        p = "HeadResnetBlock%d." % i
        skip = conv_layer(sd, p + "skip.", x, 1, activate=False) if cin != cout else x
        r = styled_conv(sd, p + "conv1.", x, g, noise=noises.get(p + "conv1"))
        r = styled_conv(sd, p + "conv2.", r, g, noise=noises.get(p + "conv2"))
        x = (skip + r) / SQRT2
        if taps is not None:
            taps[p[:-1]] = x
    feas = []
    if extract_features:
        pool_hw = x.shape[2:]
        feas.append(feature_head(sd, "layer32.", x.detach(), 3))          # generator.py:257: the heads see x.detach()
    for j, (res, cin, cout) in enumerate(UP_CH):
        p = "UpsamplingResBlock%d." % res
        g = codes[-2 - j]
        skip = conv_layer(sd, p + "skip.", x, 1, activate=True) if cin != cout else x
        skip = F.interpolate(skip, scale_factor=2, mode="bilinear", align_corners=False)
        r = styled_conv(sd, p + "conv1.", x, g, upsample=True, noise=noises.get(p + "conv1"))
        r = styled_conv(sd, p + "conv2.", r, g, noise=noises.get(p + "conv2"))
        x = (skip + r) / SQRT2
        if taps is not None:
            taps[p[:-1]] = x
        if extract_features:
            feas.append(feature_head(sd, "layer%d." % (2 ** (j + 6)), x.detach(), 3 if j < 2 else 1))
    rgb = to_rgb(sd, "ToRGB.", x, codes[0])
    if not extract_features:
        return rgb
    feat = feas[0]
    feat1 = F.interpolate(feas[0], (256, 256), mode="bilinear")
    for f in feas[1:]:
        feat = torch.cat((feat, F.adaptive_avg_pool2d(f, pool_hw)), dim=1)
        feat1 = torch.cat((feat1, F.interpolate(f, (256, 256), mode="bilinear")), dim=1)
    if taps is not None:
        taps["feat_cat"] = feat
        taps["feat1_cat"] = feat1
    for i in range(3):
        feat = residual_block(sd, "layert.%d." % i, feat)
    feat1 = residual_block(sd, "layert1.0.", feat1)
    feat1 = F.conv2d(feat1, sd["layert1.1.weight"], sd["layert1.1.bias"])
    return rgb, feat, feat1


# --------------------------------------------------------------------------------------
# correlation swap (models/ppst_model.py:330-387)
# --------------------------------------------------------------------------------------
_EPS = sys.float_info.epsilon


def rselfcorr(fea):
    """PPSTModel.Rselfcorr, ppst_model.py:330-339.  fea [B,64,256,256] -> [B,256,64,64]:
    per non-overlapping 4x4 patch, centre over channels, L2-normalise over channels, 16x16 Gram
    summed over the channels."""
    b, c, h, w = fea.shape
    ph, pw = h // 4, w // 4
    p = fea.reshape(b, c, ph, 4, pw, 4).permute(0, 1, 2, 4, 3, 5).reshape(b, c, ph * pw, 16)
    p = p - p.mean(dim=1, keepdim=True)
    p = p / (p.pow(2).sum(dim=1, keepdim=True).sqrt() + _EPS)
    gram = torch.einsum("bcli,bclj->blij", p, p).reshape(b, ph * pw, 256)
    return gram.permute(0, 2, 1).reshape(b, 256, 64, 64)


def corrm(fea, fea0):
    """PPSTModel.corrm (match_kernel == 1), ppst_model.py:341-364: softmax((fea0^T fea)/0.01) over
    the last dim; the first 256 channels are centred over the channel dim, all 512 L2-normalised."""
    def prep(f):
        f = f.reshape(f.shape[0], f.shape[1], -1)
        h1 = f[:, :256] - f[:, :256].mean(dim=1, keepdim=True)
        f = torch.cat((h1, f[:, 256:]), dim=1)
        return f / (f.pow(2).sum(dim=1, keepdim=True).sqrt() + _EPS)
    a = prep(fea)                                   # [B,C,L]  (columns)
    q = prep(fea0).permute(0, 2, 1)                 # [B,L0,C] (rows)
    return F.softmax(torch.matmul(q, a) / 0.01, dim=-1)


def model_warp(fea, corr):
    """PPSTModel.warp, ppst_model.py:366-387."""
    b, c, h, w = fea.shape
    l = h * w
    H = corr.shape[1]
    if H != l:
        s = int((l / H) ** 0.5)
        feas = F.unfold(fea, s, stride=s).permute(0, 2, 1)
        wf = torch.matmul(corr, feas).permute(0, 2, 1)
        return F.fold(wf, (h, w), s, stride=s)
    f = fea.reshape(b, c, -1).permute(0, 2, 1)
    return torch.matmul(corr, f).permute(0, 2, 1).reshape(b, c, h, w)


def lerp(a, b, r):
    """util.lerp, util/util.py:32-35."""
    if isinstance(a, (list, tuple)):
        return [lerp(x, y, r) for x, y in zip(a, b)]
    return a * (1 - r) + b * r


def swap_pair(sdE1, sdE2, sdG, content, style, alpha=1.0, noises=None):
    """The evaluator's sequence (evaluation/simple_swapping_evaluator.py:44-59) as one function;
    the `minimal` unit of SURVEY.md section 8(d) C2 (content E1/E2 run once)."""
    sp_c = encoder_con(sdE1, content)
    gl_c = encoder_col(sdE2, content)[0]
    _, fea_c, fea_c1 = generator(sdG, sp_c, gl_c, extract_features=True, noises=noises)
    sp_s = encoder_con(sdE1, style)
    gl_s = encoder_col(sdE2, style)[0]
    _, fea_s, fea_s1 = generator(sdG, sp_s, gl_s, extract_features=True, noises=noises)
    fea_c = torch.cat((fea_c, rselfcorr(fea_c1)), dim=1)
    fea_s = torch.cat((fea_s, rselfcorr(fea_s1)), dim=1)
    corr = corrm(fea_s, fea_c)
    _, gl_w = encoder_col(sdE2, style, corrmatrix=corr)
    code = lerp(gl_c, gl_w, alpha)
    return generator(sdG, sp_c, code, noises=noises)


# --------------------------------------------------------------------------------------
# models/networks/function.py (dead code in the reference, named by north_star)
# --------------------------------------------------------------------------------------
def calc_mean_std(feat, eps=1e-5):
    """function.py:3-11 -- UNBIASED variance + eps."""
    n, c = feat.shape[:2]
    f = feat.reshape(n, c, -1)
    mean = f.mean(dim=2)
    var = ((f - mean[:, :, None]) ** 2).sum(dim=2) / (f.shape[2] - 1) + eps
    return mean.view(n, c, 1, 1), var.sqrt().view(n, c, 1, 1)


def calc_mean(feat):
    """function.py:13-18."""
    n, c = feat.shape[:2]
    return feat.reshape(n, c, -1).mean(dim=2).view(n, c, 1, 1)


def nor_mean_std(feat):
    """function.py:20-24."""
    mean, std = calc_mean_std(feat)
    return (feat - mean) / std


def nor_mean(feat):
    """function.py:26-30."""
    mean = calc_mean(feat)
    return feat - mean, mean


def calc_cov(feat):
    """function.py:32-35."""
    f = feat.flatten(2, 3)
    return torch.bmm(f, f.permute(0, 2, 1)) / f.shape[2]


# --------------------------------------------------------------------------------------
# integer label maps (data/CelebAMask_dataset.py:54-60; g_mask.py:9-28,46)
# --------------------------------------------------------------------------------------
CELEBA_PARTS = ['skin', 'l_brow', 'r_brow', 'l_eye', 'r_eye', 'eye_g', 'l_ear', 'r_ear', 'ear_r',
                'nose', 'mouth', 'u_lip', 'l_lip', 'neck', 'neck_l', 'cloth', 'hair', 'hat']


def aggregate_part_label(part_index):
    """g_mask.py:9-28: the 18 CelebAMask-HQ parts collapse to 1 (face/body) or 2 (hair, hat); bg = 0."""
    return 2 if CELEBA_PARTS[part_index] in ("hair", "hat") else 1


def mask_labels(mask_np, n_labels=3):
    """CelebAMaskDataset._mask_labels: uint8 label map [H,W] -> one-hot float [n_labels,H,W] (bit-exact
    membership test mask_np == i)."""
    labels = np.zeros((n_labels,) + mask_np.shape, dtype=np.float64)
    for i in range(n_labels):
        labels[i][mask_np == i] = 1.0
    return labels


def masked_gap_gmp(x, label, n_labels=3):
    """encoder_col.py:173-190 restated on the integer label map: for region i,
    gap_i = mean_{hw}(x * [label==i]),  gmp_i = max_{hw}(x * [label==i])  (the zeros outside the region
    take part in both the mean's denominator and the max, exactly as `x * mask[:, i]` does)."""
    outs = []
    for i in range(n_labels):
        m = (label == i).to(x.dtype)[:, None]
        xm = x * m
        outs.append((xm.mean(dim=(2, 3)), xm.amax(dim=(2, 3))))
    return outs


# --------------------------------------------------------------------------------------
# D -- discriminator and its training losses (models/networks/stylegan2_layers.py:582-649,
# models/networks/discriminator.py:5-32, models/ppst_model.py:68-92,139-156, models/networks/loss.py:11-18).
# These functions are differentiable torch expressions: gradients for the parity tests come from
# torch.autograd through them, exactly as the reference obtains its own (fallback-op) gradients.
# --------------------------------------------------------------------------------------
D_CHANNELS = {4: 512, 8: 512, 16: 512, 32: 512, 64: 512, 128: 256, 256: 128, 512: 64, 1024: 32}   # channel_multiplier 2


def d_block_names(size):
    """stylegan2_layers.py:603-611: one ResBlock per octave from `size` down to 4x4."""
    log_size = int(math.log(2 ** int(round(math.log(size, 2))), 2))
    out = []
    for i in range(log_size, 2, -1):
        out.append(("convs." + (str(9 - i) if i <= 8 else "%dx%d" % (2 ** i, 2 ** i)), D_CHANNELS[2 ** i], D_CHANNELS[2 ** (i - 1)]))
    return out


def res_block_d(sd, p, x):
    """ResBlock as the discriminator builds it (stylegan2_layers.py:559-579, :609): zero padding, blur [1,3,3,1]."""
    out = conv_layer(sd, p + "conv1.", x, 3)
    out = conv_layer(sd, p + "conv2.", out, 3, downsample=True)
    skip = conv_layer(sd, p + "skip.", x, 1, downsample=True, activate=False)
    return (out + skip) / SQRT2


def discriminator(sd, x, prefix="stylegan2_D."):
    """Discriminator.forward, stylegan2_layers.py:627-645 (minibatch-stddev is commented out upstream)."""
    size = x.shape[-1]
    out = conv_layer(sd, prefix + "convs.0.", x, 1)
    for name, _, _ in d_block_names(size):
        out = res_block_d(sd, prefix + name + ".", out)
    out = conv_layer(sd, prefix + "final_conv.", out, 3)
    out = out.reshape(out.shape[0], -1)
    w0, b0 = sd[prefix + "final_linear.0.weight"], sd[prefix + "final_linear.0.bias"]
    out = fused_leaky_relu(F.linear(out, w0 * (1.0 / math.sqrt(w0.shape[1]))), b0)       # EqualLinear(activation), :231-238
    w1, b1 = sd[prefix + "final_linear.1.weight"], sd[prefix + "final_linear.1.bias"]
    return F.linear(out, w1 * (1.0 / math.sqrt(w1.shape[1])), b1)


def gan_loss(pred, real):
    """loss.py:11-18 (the least-squares form that is live upstream)."""
    return torch.mean((pred - 1) ** 2) if real else torch.mean(pred ** 2)


def d_losses(sd, real, fake, lambda_gan=1.0):
    """compute_image_discriminator_losses, training stage 1 (ppst_model.py:68-92)."""
    return {"D_real": gan_loss(discriminator(sd, real), True) * lambda_gan,
            "D_rec": gan_loss(discriminator(sd, fake), False) * (0.5 * lambda_gan)}


def r1_loss(sd, real, lambda_r1=10.0):
    """compute_R1_loss, ppst_model.py:139-156."""
    real = real.detach().requires_grad_(True)
    pred = discriminator(sd, real).sum()
    grad_real, = torch.autograd.grad(outputs=pred, inputs=[real], create_graph=True, retain_graph=True)
    return grad_real.pow(2).sum(dim=(1, 2, 3)) * (lambda_r1 * 0.5)


# training stage 2 (the reference's default, models/ppst_model.py:24): correlation-warped mix, cycle warp, mask warp, region
# style contrast.  LPIPS (lpips.LPIPS(net='alex'), ppst_model.py:48,178) needs downloaded weights: like the reference-run
# fixture (oracle/ref_harness.py's stub) it is replaced by mean |a - b| on BOTH sides.
def _lpips_stub(a, b):
    return (a - b).abs().mean()


def stage2_inputs():
    """Seeded inputs of the stage-2 fixture (tests/golden/stage2.npz): networks with the NoiseInjection weights zeroed (the
    reference draws fresh noise per call; zero weights make the step deterministic), D at 512 (un-prefixed keys), two
    smooth portraits, blocky label maps -> one-hot masks, deterministic rscl queues."""
    from oracle import rscl_oracle as R
    sdE1, sdE2, sdG = make_networks(0)
    for k in sdG:
        if k.endswith("noise.weight"):
            sdG[k] = torch.zeros_like(sdG[k])
    sdD = make_discriminator(512, seed=0, prefix="")
    real = make_portraits(2, 512, 512, seed=77, smooth=True)
    lab = make_label_map(2, 512, 512, seed=78)
    mask = torch.stack([torch.tensor(mask_labels(lab[b].numpy()), dtype=torch.float) for b in range(2)])
    queues = [R.make_queue(i) for i in range(4)]
    return sdE1, sdE2, sdG, sdD, real, mask, queues


def stage2_features(sdG, sp, gl, noises=None):
    """ppst_model.py:164-166 / 118-120: sps = cat(feas, Rselfcorr(feas1)) of G(sp, gl, extract_features=True)."""
    _, feas, feas1 = generator(sdG, sp, gl, extract_features=True, noises=noises)
    return torch.cat((feas, rselfcorr(feas1)), dim=1)


def stage2_generator_losses(sdE1, sdE2, sdG, sdD, real, mask, queues, lam=None, nce_T=0.07, noises=None):
    """PPSTModel.compute_generator_losses with training_stage == 2 (ppst_model.py:158-236).  mask: one-hot f32 [B,3,H,W];
    queues: the four rscl memories [2048, 128] (networks/rscl.py:26-40).  Returns (losses dict, enqueue list): the keys the
    reference then pushes into the queues, in its order (ppst_model.py:213-218)."""
    from oracle import rscl_oracle as R
    lam = dict({"L1": 3.0, "GAN": 1.0, "StyleCon": 1.0, "Maskwarp": 10.0, "Cycwarp": 5.0}, **(lam or {}))
    losses = {}
    B = real.shape[0]
    sp = encoder_con(sdE1, real)
    gl, _ = encoder_col(sdE2, real)
    sps = stage2_features(sdG, sp, gl, noises)
    corr = corrm(sps, _swap_pairs(sps))
    corr_self = corrm(sps, sps)
    _, gl = encoder_col(sdE2, real, corrmatrix=corr_self)
    _, pro_ms, gl_w, pro_mw = encoder_col(sdE2, real, corrmatrix=corr, mask=mask)
    image_warp = model_warp(real, corr)
    image_rec = model_warp(image_warp, _swap_pairs(corr))
    losses["image_warp_reg"] = _lpips_stub(image_rec, real) * lam["Cycwarp"]
    losses["Mask_warp"] = (model_warp(mask, corr) - _swap_pairs(mask)).abs().mean() * lam["Maskwarp"]
    rec = generator(sdG, sp, gl, noises=noises)
    losses["G_L1"] = (rec - real).abs().mean() * lam["L1"]
    mix = generator(sdG, _swap_pairs(sp), gl_w, noises=noises)
    _, pro_3m, _, _ = encoder_col(sdE2, mix, mask=_swap_pairs(mask))
    _, pro_2m, _, _ = encoder_col(sdE2, rec, mask=mask)
    sp_3 = encoder_con(sdE1, mix)
    half_noises = None if noises is None else {k: v[:B // 2] for k, v in noises.items()}
    cyc = generator(sdG, _swap_pairs(sp_3)[:B // 2], [g[:B // 2] for g in gl], noises=half_noises)
    losses["G_L1_cyc"] = (cyc - real[:B // 2]).abs().mean() * 3
    style1 = style2 = 0.0
    pushes = []
    for lvl in range(4):
        key0 = torch.cat(pro_ms[3 * lvl:3 * lvl + 3], dim=0)
        keyw = torch.cat(pro_mw[3 * lvl:3 * lvl + 3], dim=0)
        query = torch.cat(pro_3m[3 * lvl:3 * lvl + 3], dim=0)
        query_r = torch.cat(pro_2m[3 * lvl:3 * lvl + 3], dim=0)
        # the queue of level `lvl` is read BEFORE this level's keys are pushed (ppst_model.py:211-218)
        style1 = style1 + R.rscl_loss(query, keyw.detach(), queues[lvl], key0.detach(), nce_T)
        style2 = style2 + R.rscl_loss(query_r, key0.detach(), queues[lvl], keyw.detach(), nce_T)
        pushes.append((lvl, [key0[i:i + 1].detach() for i in range(3)] + [keyw[i:i + 1].detach() for i in range(3)]))
    losses["G_styleContmix"] = style1 * lam["StyleCon"]
    losses["G_styleContrec"] = style2 * lam["StyleCon"]
    losses["G_GAN_rec"] = gan_loss(discriminator(sdD, rec), True) * (lam["GAN"] * 0.5)
    losses["G_GAN_mix"] = gan_loss(discriminator(sdD, mix), True) * (lam["GAN"] * 1.0)
    return losses, pushes


def stage2_discriminator_losses(sdE1, sdE2, sdG, sdD, real, lam_gan=1.0, noises=None):
    """PPSTModel.compute_discriminator_losses with training_stage == 2 (ppst_model.py:105-137, 68-92); the caller runs it
    under torch.no_grad() for E1 / E2 / G exactly like the reference's optimizer does (the fakes are detached here)."""
    B = real.shape[0]
    with torch.no_grad():
        sp = encoder_con(sdE1, real)
        gl, _ = encoder_col(sdE2, real)
        sps = stage2_features(sdG, sp, gl, noises)
        corr = corrm(sps, _swap_pairs(sps))
        corr_self = corrm(sps, sps)
        _, gl_w = encoder_col(sdE2, real, corrmatrix=corr)
        mix = generator(sdG, _swap_pairs(sp), gl_w, noises=noises)
        _, gl = encoder_col(sdE2, real, corrmatrix=corr_self)
        half_noises = None if noises is None else {k: v[:B // 2] for k, v in noises.items()}
        rec = generator(sdG, sp[:B // 2], [g[:B // 2] for g in gl], noises=half_noises)
    return {"D_real": gan_loss(discriminator(sdD, real), True) * lam_gan,
            "D_rec": gan_loss(discriminator(sdD, rec), False) * (0.5 * lam_gan),
            "D_mix": gan_loss(discriminator(sdD, mix), False) * (0.5 * lam_gan)}


def _q(t):
    """Round to bfloat16 and back.  Under autograd the gradient is rounded at the same point on its way back."""
    return t.to(torch.bfloat16).to(torch.float32)


def fold_fir(w, k):
    """Stencil product of a conv weight [co,ci,kh,kw] and a FIR kernel: conv(w) o blur(k) == conv(fold_fir(w, k))."""
    co, ci, kh, kw = w.shape
    kf = torch.flip(k.to(w.dtype), [0, 1]).view(1, 1, *k.shape)
    out = F.conv_transpose2d(w.reshape(co * ci, 1, kh, kw), kf)
    return out.view(co, ci, kh + k.shape[0] - 1, kw + k.shape[1] - 1)


def discriminator_bf16_emulation(sd, x, prefix="stylegan2_D.", fold_blur=False):
    """NOT a restatement of the reference: the discriminator with the PRODUCT's bf16 rounding points (weights of the
    tensor-core convolutions, every stored activation incl. the blurred maps, every stored activation gradient; fold_blur:
    the product's alternative with the blur folded into 6x6 / 4x4 stride-2 convolutions), in torch fp32 arithmetic.  Second checker for the training tests: against the fp32 reference a bf16
    network differs mostly by leaky-ReLU branch flips of near-zero activations (several % in rel-L2 after 16 layers);
    against this emulation the flips coincide and what is left is accumulation order."""
    size = x.shape[-1]
    blur = make_kernel([1, 3, 3, 1])
    w = sd[prefix + "convs.0.Conv.weight"]
    out = _q(fused_leaky_relu(F.conv2d(x, w * (1.0 / math.sqrt(3.0))), sd[prefix + "convs.0.Act.bias"]))
    for name, c, co in d_block_names(size):
        p = prefix + name + "."
        s3 = 1.0 / math.sqrt(c * 9)
        y = _q(fused_leaky_relu(F.conv2d(out, _q(sd[p + "conv1.Conv.weight"] * s3), padding=1), sd[p + "conv1.Act.bias"]))
        if fold_blur:
            y = _q(fused_leaky_relu(F.conv2d(y, _q(fold_fir(sd[p + "conv2.Conv.weight"] * s3, blur)), stride=2, padding=2),
                                    sd[p + "conv2.Act.bias"]))
            sk = _q(F.conv2d(out, _q(fold_fir(sd[p + "skip.Conv.weight"] * (1.0 / math.sqrt(c)), blur)), stride=2, padding=1))
        else:
            y = _q(fused_leaky_relu(F.conv2d(_q(upfirdn2d(y, blur, pad=(2, 2))), _q(sd[p + "conv2.Conv.weight"] * s3), stride=2),
                                    sd[p + "conv2.Act.bias"]))
            sk = _q(F.conv2d(_q(upfirdn2d(out, blur, pad=(1, 1))), _q(sd[p + "skip.Conv.weight"] * (1.0 / math.sqrt(c))), stride=2))
        out = _q((y + sk) / SQRT2)
    c = out.shape[1]
    out = _q(fused_leaky_relu(F.conv2d(out, _q(sd[prefix + "final_conv.Conv.weight"] * (1.0 / math.sqrt(c * 9))), padding=1),
                              sd[prefix + "final_conv.Act.bias"]))
    out = out.reshape(out.shape[0], -1)
    w0, b0 = sd[prefix + "final_linear.0.weight"], sd[prefix + "final_linear.0.bias"]
    out = fused_leaky_relu(F.linear(out, w0 * (1.0 / math.sqrt(w0.shape[1]))), b0)
    w1, b1 = sd[prefix + "final_linear.1.weight"], sd[prefix + "final_linear.1.bias"]
    return F.linear(out, w1 * (1.0 / math.sqrt(w1.shape[1])), b1)


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


# --------------------------------------------------------------------------------------
# synthetic weights / inputs; deterministic across machines
# --------------------------------------------------------------------------------------
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
