"""Thin tensor-level wrappers over the C ABI (include/ppst_b200.h).  Device memory, streams and dtypes come
from PyTorch; every computation happens in libppst_b200.so.  No fallbacks: CPU tensors raise."""
import ctypes
import math

import torch

from . import _lib
from ._lib import ApplyDesc, ConvDesc, F32, BF16, check, dt, ptr, require_cuda, stream

SQRT2 = math.sqrt(2.0)
_TORCH_DT = {F32: torch.float32, BF16: torch.bfloat16, _lib.F16: torch.float16}


def _L():
    return _lib.load()


# Optional per-launch timing (bench.py's roofline leg): PROFILE = {} enables it; every wrapped launch is bracketed
# by CUDA events on the launching stream and filed under its kernel name with its algorithmic work
# (FLOPs for GEMM-shaped kernels, bytes for HBM-bound ones).
PROFILE = None
_prof_count = 0


PROFILE_DETAIL = False        # file launches under "name|tag" (tag = shape string) instead of "name"


class _Prof:
    def __init__(self, name, work, tag=None, aux=0.0, alg=None):
        self.name = name if not (PROFILE_DETAIL and tag) else "%s|%s" % (name, tag)
        self.work = work            # EXECUTED work (FLOPs for GEMM-shaped kernels, bytes for HBM-bound ones)
        self.aux = aux              # secondary work figure (conv_tc: algorithmic bytes next to the FLOPs)
        self.alg = work if alg is None else alg     # ALGORITHMIC work of the reference op this launch stands for

    def __enter__(self):
        global _prof_count
        if PROFILE is not None:
            # keep the GPU busy every so often so that the host runs AHEAD of the device: otherwise the interval
            # between the two events also contains the host time spent issuing the launch (Python, ~50-100 us)
            if _prof_count % 64 == 0:
                torch.cuda._sleep(30_000_000)
            _prof_count += 1
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if PROFILE is not None:
            self.e1.record()
            PROFILE.setdefault(self.name, []).append((self.work, self.e0, self.e1, self.aux, self.alg))
        return False


def profile_summary():
    """name -> dict(launches, work, ms) after a torch.cuda.synchronize()."""
    out = {}
    for name, recs in (PROFILE or {}).items():
        ms = sum(r[1].elapsed_time(r[2]) for r in recs)
        out[name] = {"launches": len(recs), "work": float(sum(r[0] for r in recs)), "ms": ms,
                     "aux": float(sum(r[3] for r in recs)), "alg": float(sum(r[4] for r in recs))}
    return out


def _conv_flops(d):
    k = d.kh * d.kw * d.cin * (1.5 if d.split_terms == 3 else 1.0)       # split: 3 segments of cin/2
    return 2.0 * d.n * d.out_h * d.out_w * d.phases_y * d.phases_x * d.cout * k


# ----------------------------------------------------------------------------------------------- layout
def nchw_to_nhwc(x, dtype=None, out=None, c_off=0):
    """[n,c,h,w] -> NHWC [n,h,w,C]; writes channel slice [c_off, c_off+c) of `out` if given."""
    require_cuda(x)
    x = x.contiguous()
    n, c, h, w = x.shape
    if out is None:
        out = torch.empty((n, h, w, c), device=x.device, dtype=dtype or x.dtype)
    check(_L().ppst_nchw_to_nhwc(ptr(out), ptr(x), n, c, h * w, out.shape[3], c_off, dt(x), dt(out), stream()))
    return out


def nhwc_to_nchw(x, c=None, c_off=0, dtype=None, a=None, b=None):
    """NHWC [n,h,w,cs] channel slice -> NCHW [n,c,h,w]; optional per-(n, c) affine a, b (f32 [n, c]) on the way."""
    require_cuda(x, a, b)
    n, h, w, cs = x.shape
    c = cs if c is None else c
    out = torch.empty((n, c, h, w), device=x.device, dtype=dtype or x.dtype)
    if a is not None:
        assert a.dtype == torch.float32 and a.is_contiguous() and a.shape == (n, c)
        assert b is None or (b.dtype == torch.float32 and b.is_contiguous() and b.shape == (n, c))
    check(_L().ppst_nhwc_to_nchw(ptr(out), ptr(x), n, c, h * w, cs, c_off, dt(x), dt(out), ptr(a), ptr(b), stream()))
    return out


# ----------------------------------------------------------------------------------------------- reductions
_scratch = {}


def _get_scratch(device, elems):
    key = (device.index, torch.cuda.current_stream().cuda_stream)
    buf = _scratch.get(key)
    if buf is None or buf.numel() < elems:
        buf = torch.empty(max(int(elems), 1 << 20), device=device, dtype=torch.float32)
        _scratch[key] = buf
    return buf


def plane_reduce(x, channels_last, labels=None, n_regions=1, label_mode=0, shift=None, want=("sum", "sumsq", "max"),
                 border=0):
    """Per-(n, region, c) sum / sumsq / max over H*W.  x: [n,c,h,w] (channels_last False) or [n,h,w,c] (True).
    border > 0 (channels_last): x is [n, h + 2*border, w + 2*border, c] and only its interior is reduced.
    Returns dict of f32 [n, n_regions, c] tensors."""
    require_cuda(x, labels, shift)
    assert x.is_contiguous()
    if channels_last:
        n, h, w, c = x.shape
        h, w = h - 2 * border, w - 2 * border
    else:
        assert border == 0
        n, c, h, w = x.shape
    hw = h * w
    L = _L()
    scratch = _get_scratch(x.device, L.ppst_plane_reduce_scratch_elems(n, c, hw, n_regions))
    outs = {k: torch.empty((n, n_regions, c), device=x.device, dtype=torch.float32) for k in want}
    if labels is not None:
        assert labels.dtype == torch.uint8 and labels.is_contiguous() and labels.numel() == n * hw
    with _Prof("plane_reduce", x.numel() * x.element_size(), "n%d hw%d c%d cl%d %s" % (n, hw, c, int(channels_last), str(x.dtype)[6:])):
        check(L.ppst_plane_reduce(ptr(outs.get("sum")), ptr(outs.get("sumsq")), ptr(outs.get("max")), ptr(x),
                                  ptr(labels), ptr(shift), ptr(scratch), n, c, hw, n_regions, label_mode,
                                  1 if channels_last else 0, dt(x), w, border, stream()))
    return outs


def upsample_reduce(x, c_off, c, scale, want=("sum", "max")):
    """GAP / GMP operands of the bilinear x`scale` upsampling of x[..., c_off:c_off+c] (f32 NHWC [n,lh,lw,cs]) without
    materialising it.  Returns dict of f32 [n, c] ("sum" over the (lh*scale)x(lw*scale) virtual map, "max")."""
    require_cuda(x)
    assert x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 4
    n, lh, lw, cs = x.shape
    L = _L()
    scratch = _get_scratch(x.device, L.ppst_upsample_reduce_scratch_elems(n, lh, c))
    outs = {k: torch.empty((n, c), device=x.device, dtype=torch.float32) for k in want}
    with _Prof("upsample_reduce", n * lh * lw * c * 4, "n%d %dx%d c%d x%d" % (n, lh, lw, c, scale)):
        check(L.ppst_upsample_reduce(ptr(outs.get("sum")), ptr(outs.get("max")), ptr(x), n, lh, lw, cs, c_off, c, scale,
                                     ptr(scratch), stream()))
    return outs


def labels_to_onehot(labels, n_labels=3):
    require_cuda(labels)
    n, h, w = labels.shape
    out = torch.empty((n, n_labels, h, w), device=labels.device, dtype=torch.float32)
    check(_L().ppst_labels_to_onehot(ptr(out), ptr(labels.contiguous()), n, h * w, n_labels, stream()))
    return out


def labels_pool_or(labels, factor, in_mode):
    require_cuda(labels)
    n, h, w = labels.shape
    out = torch.empty((n, h // factor, w // factor), device=labels.device, dtype=torch.uint8)
    check(_L().ppst_labels_pool_or(ptr(out), ptr(labels.contiguous()), n, h, w, factor, in_mode, stream()))
    return out


# ----------------------------------------------------------------------------------------------- convolution
def conv_desc(x_shape, cout, kh, kw, stride=1, off=(0, 0), out_hw=None, phases=(1, 1), phase_off=(0, 0),
              full_hw=None, cout_stride=None, c_off=0, act=False, slope=0.2, gain=1.0, noise_w=0.0,
              in_dtype=F32, out_dtype=F32, in_row_pitch=0, in_img_pitch=0, w_img_rows=0, split_terms=0, out_pad=0,
              bias_img_stride=0, out_pad_trim=0):
    n, in_h, in_w, cin = x_shape
    d = ConvDesc()
    d.n, d.in_h, d.in_w, d.cin = n, in_h, in_w, cin
    d.out_h, d.out_w = out_hw
    d.cout, d.kh, d.kw, d.stride = cout, kh, kw, stride
    d.off_y, d.off_x = off
    d.phases_y, d.phases_x = phases
    d.phase_off_dy, d.phase_off_dx = phase_off
    fh, fw = full_hw if full_hw else (out_hw[0] * phases[0], out_hw[1] * phases[1])
    d.full_h, d.full_w = fh, fw
    d.cout_stride = cout_stride if cout_stride else cout
    d.c_off = c_off
    d.has_act = 1 if act else 0
    d.slope, d.gain, d.noise_w = slope, gain, noise_w
    d.in_dtype, d.out_dtype = in_dtype, out_dtype
    d.in_row_pitch, d.in_img_pitch = in_row_pitch, in_img_pitch
    d.w_img_rows, d.split_terms = w_img_rows, split_terms
    d.out_pad, d.bias_img_stride, d.out_pad_trim = out_pad, bias_img_stride, out_pad_trim
    return d


def conv_simt(out, x, weight_f32, bias, noise, d):
    require_cuda(out, x, weight_f32, bias, noise)
    with _Prof("conv_simt", _conv_flops(d), "n%d %dx%d ci%d co%d k%dx%d s%d" % (d.n, d.out_h, d.out_w, d.cin, d.cout, d.kh, d.kw, d.stride)):
        check(_L().ppst_conv_simt(ptr(out), ptr(x), ptr(weight_f32), ptr(bias), ptr(noise), ctypes.byref(d), stream()))
    return out


def conv_tc_stats_elems(d):
    return int(_L().ppst_conv_tc_stats_elems(ctypes.byref(d)))


def conv_tc(out, x, weight_bf16, bias, noise, stats, d, alg_taps=None):
    """alg_taps: taps of the reference convolution this launch stands for when they differ from kh*kw (a Blur folded
    into a stride-2 convolution executes 5x5 / 3x3 taps for an algorithmic 3x3 / 1x1, SURVEY.md 8(d) counts the latter)."""
    require_cuda(out, x, weight_bf16, bias, noise, stats)
    nbytes = (d.n * d.in_h * d.in_w * d.cin * 2 + d.n * d.full_h * d.full_w * d.cout * (2 if d.out_dtype == BF16 else 4)
              + weight_bf16.numel() * 2)
    flops = _conv_flops(d)
    alg = flops if alg_taps is None else flops * alg_taps / float(d.kh * d.kw)
    with _Prof("conv_tc", flops, "n%d %dx%d ci%d co%d k%dx%d s%d ph%d" % (d.n, d.out_h, d.out_w, d.cin, d.cout, d.kh, d.kw, d.stride, d.phases_y * d.phases_x), aux=nbytes, alg=alg):
        check(_L().ppst_conv_tc(ptr(out), ptr(x), ptr(weight_bf16), ptr(bias), ptr(noise), ptr(stats), ctypes.byref(d),
                                stream()))
    return out


def conv1x1_in_nchw(x_nchw, w, bias, act, out_dtype, slope=0.2, gain=SQRT2, out_border=0):
    """FromRGB: NCHW f32 image -> NHWC [n,h,w,cout] (out_border: the interior of [n, h+2b, w+2b, cout], border unwritten)."""
    require_cuda(x_nchw, w, bias)
    x_nchw = x_nchw.contiguous()
    n, cin, h, wd = x_nchw.shape
    cout = w.shape[0]
    out = torch.empty((n, h + 2 * out_border, wd + 2 * out_border, cout), device=x_nchw.device, dtype=out_dtype)
    with _Prof("conv1x1_in_nchw", x_nchw.numel() * 4 + out.numel() * out.element_size()):
        check(_L().ppst_conv1x1_in_nchw(ptr(out), ptr(x_nchw), ptr(w), ptr(bias), n, cin, h * wd, cout, cout, 0,
                                        1 if act else 0, float(slope), float(gain if act else 1.0), dt(out), wd, out_border,
                                        stream()))
    return out


def conv1x1_thin_out(x_nhwc, w, bias):
    """ToRGB conv: NHWC [n,h,w,cin] -> f32 NHWC [n,h,w,cout]."""
    require_cuda(x_nhwc, w, bias)
    n, h, wd, cin = x_nhwc.shape
    cout = w.shape[0]
    out = torch.empty((n, h, wd, cout), device=x_nhwc.device, dtype=torch.float32)
    with _Prof("conv1x1_thin_out", x_nhwc.numel() * x_nhwc.element_size() + out.numel() * 4):
        check(_L().ppst_conv1x1_thin_out(ptr(out), ptr(x_nhwc), ptr(w), ptr(bias), n * h * wd, cin, cout, dt(x_nhwc),
                                         stream()))
    return out


def conv_tc_available():
    return bool(_L().ppst_conv_tc_available())


# ----------------------------------------------------------------------------------------------- epilogues
def in_coeffs(n, c, sum_=None, sumsq=None, count=1.0, shift=None, style=None, bias=None, n_partials=0, eps=1e-5,
              use_norm=True, device=None):
    device = device or (sum_ if sum_ is not None else style if style is not None else bias).device
    a = torch.empty((n, c), device=device, dtype=torch.float32)
    b = torch.empty((n, c), device=device, dtype=torch.float32)
    style_ld = 0
    if style is not None:
        assert style.stride(1) == 1 and style.shape[1] == 2 * c
        style_ld = style.stride(0)
    check(_L().ppst_in_coeffs(ptr(a), ptr(b), ptr(sum_), ptr(sumsq), ptr(shift), ptr(style), style_ld, ptr(bias), n, c,
                              n_partials, float(count), float(eps), 1 if use_norm else 0, stream()))
    return a, b


def affine_apply(x, c, a=None, b=None, res=None, ra=None, rb=None, pre_slope=1.0, pre_gain=1.0, post_slope=1.0,
                 post_gain=1.0, res_mode=0, up=1, pad=(0, 0, 0, 0), pad_mode=0, out=None, out_dtype=None, out_c_off=0,
                 out_border=0, res_border=0):
    """x NHWC [n,h,w,c] (dense); see ppst_affine_apply.  pad = (top, bottom, left, right).
    out_border: the result goes to the interior of [n, oh + 2b, ow + 2b, c] (border: pad_border); res_border: `res` has
    such a border and is read at its interior."""
    require_cuda(x, a, b, res, ra, rb)
    n, h, w, cs = x.shape
    assert cs == c and x.is_contiguous()
    oh, ow = h * up + pad[0] + pad[1], w * up + pad[2] + pad[3]
    if out is None:
        out = torch.empty((n, oh + 2 * out_border, ow + 2 * out_border, c), device=x.device, dtype=out_dtype or x.dtype)
    assert out.shape[:3] == (n, oh + 2 * out_border, ow + 2 * out_border)
    if res is not None:
        rh, rw = (oh - pad[0] - pad[1], ow - pad[2] - pad[3]) if res_mode == 1 else ((oh - pad[0] - pad[1]) // 2, (ow - pad[2] - pad[3]) // 2)
        assert res.is_contiguous() and tuple(res.shape[1:3]) == (rh + 2 * res_border, rw + 2 * res_border), (res.shape, rh, rw)
    d = ApplyDesc()
    d.n, d.h, d.w, d.c, d.up = n, h, w, c, up
    d.pad_t, d.pad_b, d.pad_l, d.pad_r = pad
    d.pad_mode = pad_mode
    d.pre_slope, d.pre_gain, d.post_slope, d.post_gain = pre_slope, pre_gain, post_slope, post_gain
    d.res_mode = res_mode
    d.x_dtype, d.out_dtype = dt(x), dt(out)
    d.res_dtype = dt(res) if res is not None else dt(x)
    d.out_c_stride, d.out_c_off = out.shape[3], out_c_off
    d.out_border, d.res_border = out_border, res_border
    nbytes = x.numel() * x.element_size() + n * oh * ow * c * out.element_size()
    if res is not None:
        nbytes += (res.numel() if res_mode == 1 else n * oh * ow * c) * res.element_size()
    with _Prof("affine_apply", nbytes, "n%d %dx%d c%d up%d pad%s res%d %s->%s" % (n, h, w, c, up, "".join(map(str, pad)), res_mode, str(x.dtype)[6:], str(out.dtype)[6:])):
        check(_L().ppst_affine_apply(ptr(out), ptr(x), ptr(a), ptr(b), ptr(res), ptr(ra), ptr(rb), ctypes.byref(d),
                                     stream()))
    return out


def affine_apply_stats(x, c, a=None, b=None, res=None, ra=None, rb=None, pre_slope=1.0, pre_gain=1.0, post_slope=1.0,
                       post_gain=1.0, res_mode=0, pad=(0, 0, 0, 0), pad_mode=0, res_border=0):
    """affine_apply that also returns the instance-norm statistics of the (padded) tensor it writes
    (ppst_affine_apply_stats): -> (out bf16 [n, oh, ow, c], stats dict for in_coeffs)."""
    require_cuda(x, a, b, res, ra, rb)
    n, h, w, cs = x.shape
    assert cs == c and x.is_contiguous() and x.dtype == torch.bfloat16
    oh, ow = h + pad[0] + pad[1], w + pad[2] + pad[3]
    out = torch.empty((n, oh, ow, c), device=x.device, dtype=torch.bfloat16)
    sbuf = torch.empty((2, n, oh, c), device=x.device, dtype=torch.float32)
    d = ApplyDesc()
    d.n, d.h, d.w, d.c, d.up = n, h, w, c, 1
    d.pad_t, d.pad_b, d.pad_l, d.pad_r = pad
    d.pad_mode = pad_mode
    d.pre_slope, d.pre_gain, d.post_slope, d.post_gain = pre_slope, pre_gain, post_slope, post_gain
    d.res_mode = res_mode
    d.x_dtype = d.out_dtype = d.res_dtype = dt(x)
    d.out_c_stride, d.out_c_off = c, 0
    d.out_border, d.res_border = 0, res_border
    nbytes = x.numel() * 2 + out.numel() * 2
    if res is not None:
        assert res.is_contiguous() and res.dtype == torch.bfloat16
        nbytes += res.numel() * 2
    with _Prof("affine_apply", nbytes, "n%d %dx%d c%d pad%s res%d +stats" % (n, h, w, c, "".join(map(str, pad)), res_mode)):
        check(_L().ppst_affine_apply_stats(ptr(out), ptr(x), ptr(a), ptr(b), ptr(res), ptr(ra), ptr(rb), ctypes.byref(d),
                                           ptr(sbuf), stream()))
    return out, {"sum": sbuf[0], "sumsq": sbuf[1], "n_partials": oh, "count": oh * ow}


def affine_apply_torgb(x, a, b, res, rgb_w, rgb_bias, post_gain=1.0, write_out=True):
    """The last UpsamplingResBlock tail + ToRGB's 1x1 conv in one pass (ppst_affine_apply_torgb).
    x bf16 NHWC [n,h,w,128], res bf16 [n,h/2,w/2,128] -> (out bf16 [n,h,w,128] or None, rgb f32 [n,h,w,3])."""
    require_cuda(x, a, b, res, rgb_w, rgb_bias)
    n, h, w, c = x.shape
    assert x.is_contiguous() and res.is_contiguous() and rgb_w.is_contiguous() and rgb_w.shape == (3, c)
    out = torch.empty_like(x) if write_out else None
    rgb = torch.empty((n, h, w, 3), device=x.device, dtype=torch.float32)
    d = ApplyDesc()
    d.n, d.h, d.w, d.c, d.up = n, h, w, c, 1
    d.pad_t = d.pad_b = d.pad_l = d.pad_r = 0
    d.pad_mode = 0
    d.pre_slope, d.pre_gain, d.post_slope, d.post_gain = 1.0, 1.0, 1.0, post_gain
    d.res_mode = 2
    d.x_dtype = d.out_dtype = d.res_dtype = dt(x)
    d.out_c_stride, d.out_c_off = c, 0
    nbytes = x.numel() * 2 + res.numel() * 2 + (x.numel() * 2 if write_out else 0) + rgb.numel() * 4
    with _Prof("affine_apply", nbytes, "n%d %dx%d c%d res2+torgb out%d" % (n, h, w, c, int(write_out))):
        check(_L().ppst_affine_apply_torgb(ptr(out), ptr(rgb), ptr(x), ptr(a), ptr(b), ptr(res), ptr(rgb_w),
                                           ptr(rgb_bias), ctypes.byref(d), stream()))
    return out, rgb


def resample(x, out_hw, mode, out=None, out_dtype=None, out_c_off=0, a=None, b=None, slope=1.0, in_border=0,
             out_border=0):
    """mode 0 bilinear (align_corners False), 1 adaptive average.  x NHWC dense.
    a / b (f32 [n, c]) + slope: every source sample is first mapped to lrelu(x*a + b, slope).
    in_border / out_border: x / out carry a border of that many pixels and are addressed at their interiors."""
    require_cuda(x, a, b)
    n, h, w, c = x.shape
    h, w = h - 2 * in_border, w - 2 * in_border
    if out is None:
        out = torch.empty((n, out_hw[0] + 2 * out_border, out_hw[1] + 2 * out_border, c), device=x.device,
                          dtype=out_dtype or x.dtype)
    assert tuple(out.shape[1:3]) == (out_hw[0] + 2 * out_border, out_hw[1] + 2 * out_border)
    if a is not None:
        assert a.dtype == torch.float32 and a.is_contiguous() and a.shape == (n, c)
        assert b is None or (b.dtype == torch.float32 and b.is_contiguous() and b.shape == (n, c))
    with _Prof("resample", x.numel() * x.element_size() + n * out_hw[0] * out_hw[1] * c * out.element_size(),
               "n%d %dx%d->%dx%d c%d m%d" % (n, h, w, out_hw[0], out_hw[1], c, mode)):
        check(_L().ppst_resample(ptr(out), ptr(x), n, h, w, out_hw[0], out_hw[1], c, out.shape[3], out_c_off, mode,
                                 dt(x), dt(out), ptr(a), ptr(b), float(slope), in_border, out_border, stream()))
    return out


# ----------------------------------------------------------------------------------------------- dense / vectors
def linear(x, w, b=None, w_scale=1.0, b_scale=1.0, relu_in=False):
    require_cuda(x, w, b)
    assert x.dtype == torch.float32 and w.dtype == torch.float32 and x.stride(-1) == 1
    m, k = x.shape
    n = w.shape[0]
    y = torch.empty((m, n), device=x.device, dtype=torch.float32)
    ldx = x.stride(0) if m > 1 else k               # the stride of a size-1 dimension is arbitrary
    check(_L().ppst_linear(ptr(y), ptr(x), ptr(w), ptr(b), m, n, k, ldx, float(w_scale), float(b_scale),
                           1 if relu_in else 0, stream()))
    return y


def l2_normalize(x, eps, mode):
    require_cuda(x)
    x = x.contiguous()
    y = torch.empty_like(x)
    check(_L().ppst_l2_normalize(ptr(y), ptr(x), x.shape[0], x.shape[1], float(eps), mode, stream()))
    return y


def lerp(a, b, r):
    require_cuda(a, b)
    a, b = a.contiguous(), b.contiguous()
    y = torch.empty_like(a)
    check(_L().ppst_lerp(ptr(y), ptr(a), ptr(b), float(r), a.numel(), stream()))
    return y


def gemm_f32(a, b, trans_b, alpha=1.0, out=None):
    """a [B,m,k], b [B,k,n] (trans_b False) or [B,n,k] (True), f32 contiguous -> [B,m,n]."""
    require_cuda(a, b)
    assert a.dtype == torch.float32 and b.dtype == torch.float32 and a.is_contiguous() and b.is_contiguous()
    bt, m, k = a.shape
    n = b.shape[1] if trans_b else b.shape[2]
    if out is None:
        out = torch.empty((bt, m, n), device=a.device, dtype=torch.float32)
    ldb = b.shape[2]
    with _Prof("gemm_f32", 2.0 * bt * m * n * k, "b%d m%d n%d k%d t%d" % (bt, m, n, k, int(trans_b))):
        check(_L().ppst_gemm_f32(ptr(out), ptr(a), ptr(b), bt, m, n, k, k, ldb, n, m * k, b.shape[1] * b.shape[2],
                                 m * n, 1 if trans_b else 0, float(alpha), stream()))
    return out


def softmax_rows_(x, divisor=1.0):
    require_cuda(x)
    assert x.dtype == torch.float32 and x.is_contiguous()
    cols = x.shape[-1]
    check(_L().ppst_softmax_rows(ptr(x), x.numel() // cols, cols, float(divisor), stream()))
    return x


# ----------------------------------------------------------------------------------------------- correlation
def rselfcorr(fea_nhwc, out=None, out_c_off=0, out_dtype=None):
    require_cuda(fea_nhwc)
    n, h, w, c = fea_nhwc.shape
    if out is None:
        out = torch.empty((n, h // 4, w // 4, 256), device=fea_nhwc.device, dtype=out_dtype or fea_nhwc.dtype)
    check(_L().ppst_rselfcorr(ptr(out), ptr(fea_nhwc), n, h, w, c, out.shape[3], out_c_off, dt(fea_nhwc), dt(out),
                              stream()))
    return out


def corr_prepare(x_rows, c_center):
    """x_rows: [..., c] channels-last; returns f32 of the same shape, centred (first c_center) + L2-normalised."""
    require_cuda(x_rows)
    x_rows = x_rows.contiguous()
    c = x_rows.shape[-1]
    out = torch.empty(x_rows.shape, device=x_rows.device, dtype=torch.float32)
    check(_L().ppst_corr_prepare(ptr(out), ptr(x_rows), x_rows.numel() // c, c, c_center, dt(x_rows), stream()))
    return out


# ----------------------------------------------------------------------------------------------- split tensor-core GEMM
def corr_prepare_split(x_rows, c_center):
    """corr_prepare with [hi | lo] bf16 output: [..., c] -> bf16 [..., 2c]."""
    require_cuda(x_rows)
    x_rows = x_rows.contiguous()
    c = x_rows.shape[-1]
    out = torch.empty(x_rows.shape[:-1] + (2 * c,), device=x_rows.device, dtype=torch.bfloat16)
    check(_L().ppst_corr_prepare_split(ptr(out), ptr(x_rows), x_rows.numel() // c, c, c_center, dt(x_rows), stream()))
    return out


def split_bf16(x, transpose=False):
    """f32 [b, rows, cols] -> bf16 [b, rows, 2*cols] ([hi | lo]);  transpose: -> bf16 [b, cols, 2*rows]."""
    require_cuda(x)
    assert x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 3
    b, r, c = x.shape
    out = torch.empty((b, c, 2 * r) if transpose else (b, r, 2 * c), device=x.device, dtype=torch.bfloat16)
    check(_L().ppst_split_bf16(ptr(out), ptr(x), b, r, c, 1 if transpose else 0, stream()))
    return out


def softmax_rows_split(logits, divisor):
    """softmax(logits / divisor, -1) as [hi | lo] bf16: f32 [..., cols] -> bf16 [..., 2*cols]."""
    require_cuda(logits)
    assert logits.dtype == torch.float32 and logits.is_contiguous()
    cols = logits.shape[-1]
    out = torch.empty(logits.shape[:-1] + (2 * cols,), device=logits.device, dtype=torch.bfloat16)
    check(_L().ppst_softmax_rows_split(ptr(out), ptr(logits), logits.numel() // cols, cols, float(divisor), stream()))
    return out


def scale_weights(w32, a, taps, cin):
    """Per-sample modulated weights: w32 f32 [rows, taps*cin] (packed), a f32 [n, cin] -> bf16 [n, rows, taps*cin]."""
    require_cuda(w32, a)
    assert w32.dtype == torch.float32 and a.dtype == torch.float32 and w32.is_contiguous()
    a = a.contiguous()
    rows = w32.numel() // (taps * cin)
    out = torch.empty((a.shape[0], rows, taps * cin), device=w32.device, dtype=torch.bfloat16)
    check(_L().ppst_scale_weights(ptr(out), ptr(w32), ptr(a), a.shape[0], rows, int(taps), int(cin), stream()))
    return out


def modulate_weights(w32, a, b, taps, cin, base_bias=None):
    """Lazy affine in front of a zero-padded conv: (bf16 [n, rows, taps*cin] = W*a, f32 [n, rows] = base_bias + sum W*b)."""
    require_cuda(w32, a, b, base_bias)
    assert w32.dtype == torch.float32 and w32.is_contiguous() and a.dtype == torch.float32 and b.dtype == torch.float32
    a, b = a.contiguous(), b.contiguous()
    rows = w32.numel() // (taps * cin)
    n = a.shape[0]
    out_w = torch.empty((n, rows, taps * cin), device=w32.device, dtype=torch.bfloat16)
    out_b = torch.empty((n, rows), device=w32.device, dtype=torch.float32)
    check(_L().ppst_modulate_weights(ptr(out_w), ptr(out_b), ptr(w32), ptr(a), ptr(b), ptr(base_bias), n, rows, int(taps),
                                     int(cin), stream()))
    return out_w, out_b


def fill_border(buf, a, b, pad=1):
    """Border of the padded NHWC bf16 buffer [n, h+2p, w+2p, c] <- -b/a per (sample, channel)."""
    require_cuda(buf, a, b)
    n, hp, wp, c = buf.shape
    assert buf.is_contiguous() and a.is_contiguous() and b.is_contiguous()
    check(_L().ppst_fill_border(ptr(buf), ptr(a), ptr(b), n, hp - 2 * pad, wp - 2 * pad, c, pad, dt(buf), stream()))
    return buf


def pad_border(buf, pad, mode):
    """Border of the padded NHWC buffer [n, pt + h + pb, pl + w + pr, c] from its interior: mode 1 reflect, 2 replicate,
    0 zero.  pad = (top, bottom, left, right)."""
    require_cuda(buf)
    n, hp, wp, c = buf.shape
    assert buf.is_contiguous()
    pt, pb, pl, pr = pad
    check(_L().ppst_pad_border(ptr(buf), n, hp - pt - pb, wp - pl - pr, c, pt, pb, pl, pr, mode, dt(buf), stream()))
    return buf


def guided_smooth(out_img, target, radius=30, eps=(0.02 * 255) ** 2):
    """PPSTModel.decode's target branch (ppst_model.py:288-306; photo_gif.py:13-46) on the device: both images NCHW f32
    [n,3,h,w] in [-1,1]; returns the guided-filtered image in the same convention."""
    require_cuda(out_img, target)
    assert out_img.shape == target.shape and out_img.shape[1] == 3, (out_img.shape, target.shape)
    src = out_img.to(torch.float32).contiguous()
    guide = target.to(torch.float32).contiguous()
    n, _, h, w = src.shape
    ws_bytes = int(_L().ppst_guided_smooth_workspace_bytes(n, h, w))
    ws = torch.empty(ws_bytes, device=src.device, dtype=torch.uint8)
    y = torch.empty_like(src)
    with _Prof("guided_smooth", 2 * src.numel() * 4):
        check(_L().ppst_guided_smooth(ptr(y), ptr(src), ptr(guide), n, h, w, int(radius), float(eps), ptr(ws), ws_bytes,
                                      stream()))
    return y


def gemm_split_tc(a_split, b_split, m_hw, out=None):
    """C[b] = A[b] . B[b]^T with fp32-level accuracy on the tensor cores (3-term bf16 split).
    a_split: bf16 [b, M, 2K] ([hi | lo]), M = m_hw[0] * m_hw[1];  b_split: bf16 [b, N, 2K];  -> f32 [b, M, N].
    Runs ppst_conv_tc as a 1x1 'convolution' over an [b, m_h, m_w, 2K] image with per-sample weights."""
    require_cuda(a_split, b_split)
    bt, M, k2 = a_split.shape
    N = b_split.shape[1]
    assert b_split.shape == (bt, N, k2) and M == m_hw[0] * m_hw[1]
    if out is None:
        out = torch.empty((bt, M, N), device=a_split.device, dtype=torch.float32)
    d = conv_desc((bt, m_hw[0], m_hw[1], k2), N, 1, 1, out_hw=m_hw, in_dtype=BF16, out_dtype=F32, w_img_rows=N,
                  split_terms=3)
    with _Prof("gemm_split_tc", 2.0 * bt * M * N * (k2 // 2) * 3, "b%d m%d n%d k%d" % (bt, M, N, k2 // 2),
               alg=2.0 * bt * M * N * (k2 // 2)):      # 3 bf16 products executed per algorithmic fp32 product
        check(_L().ppst_conv_tc(ptr(out), ptr(a_split), ptr(b_split), None, None, None, ctypes.byref(d), stream()))
    return out
