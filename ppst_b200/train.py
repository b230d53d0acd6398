"""Training-side operators and the discriminator on the B200 kernels (BASELINE config 5).

What the reference does with torch.autograd + cuDNN (F.conv2d / F.conv_transpose2d forward, dgrad, wgrad in fp32) runs
here on the tcgen05 convolution kernel of the inference path:

    forward   y  = conv(x, w)                     ppst_conv_tc                      (bf16 operands, fp32 accumulate)
    dgrad     dx = conv(dy, flip(w)^T)            ppst_conv_tc again: stride 1 = one plain convolution with the rotated,
                                                  transposed weights; stride 2 = four sub-pixel phases (no zero insertion)
    wgrad     dw = sum_pixels dy (x) shift(x)     ppst_conv_wgrad (MN-major tcgen05 operands read straight from NHWC by TMA,
                                                  pixel chunks -> fp32 partials -> fixed-order reduce)

The three maps are bilinear and closed under differentiation, so each is a torch.autograd.Function whose backward calls
the other two: first AND second derivatives (the R1 penalty, models/ppst_model.py:139-156, differentiates the input
gradient of D once more) run on the same kernels.  Activations are NHWC bf16 tensors [n, h, w, c]; parameters stay fp32
in the reference's layout ([co, ci, kh, kw], state_dict keys unchanged) and are packed per call.

torch supplies the tape (autograd), device memory and the optimiser; all activation-sized work is in libppst_b200.so.
No CPU path: CPU tensors raise.
"""
import ctypes
import math

import torch
from torch.autograd import Function

from . import _lib, ops
from ._lib import BF16, F32, WgradDesc, check, dt, ptr, require_cuda, stream

SQRT2 = math.sqrt(2.0)


# ------------------------------------------------------------------------------------------------ raw launches
class Geom:
    """A square-kernel convolution: kernel size k, stride (1 or 2), symmetric zero padding."""

    __slots__ = ("k", "stride", "pad")

    def __init__(self, k, stride=1, pad=0):
        self.k, self.stride, self.pad = int(k), int(stride), int(pad)

    def out_hw(self, h, w):
        return (h + 2 * self.pad - self.k) // self.stride + 1, (w + 2 * self.pad - self.k) // self.stride + 1


def _nhwc(t):
    require_cuda(t)
    if t.dtype != torch.bfloat16:
        t = t.to(torch.bfloat16)
    return t.contiguous()


_slot_cache = {}


def _pack(w, slots_key, groups, slots, src, transpose):
    """f32 [co,ci,k,k] -> bf16 [groups][rows][slots*cols] in one launch (ppst_pack_conv_weights)."""
    co, ci, kh, kw = w.shape
    tab = _slot_cache.get(slots_key)
    if tab is None:
        tab = (ctypes.c_int * len(src))(*src)
        _slot_cache[slots_key] = tab
    rows, cols = (ci, co) if transpose else (co, ci)
    w32 = w.detach()
    if w32.dtype != torch.float32 or not w32.is_contiguous():
        w32 = w32.to(torch.float32).contiguous()
    out = torch.empty((groups, rows, slots * cols), device=w.device, dtype=torch.bfloat16)
    check(_lib.load().ppst_pack_conv_weights(ptr(out), ptr(w32), co, ci, kh, kw, groups, slots, tab, 1 if transpose else 0,
                                             1.0, stream()))
    return out


def _stats_buf(d, device, want):
    return torch.empty(ops.conv_tc_stats_elems(d), device=device, dtype=torch.float32) if want else None


def _conv_raw(x, w, g, bias=None, act=False, noise=None, want_stats=False):
    """x NHWC bf16 [n,h,w,ci], w f32 [co,ci,k,k] -> NHWC bf16 [n,oh,ow,co]; optional fused bias + per-pixel noise
    (f32 [n,oh,ow], already scaled) + leaky-ReLU(0.2)*sqrt2."""
    n, h, wd, ci = x.shape
    co = w.shape[0]
    assert w.shape[1] == ci and w.shape[2] == g.k and w.shape[3] == g.k, (tuple(w.shape), tuple(x.shape))
    oh, ow = g.out_hw(h, wd)
    wp = _pack(w, ("fwd", g.k), 1, g.k * g.k, list(range(g.k * g.k)), False)
    out = torch.empty((n, oh, ow, co), device=x.device, dtype=torch.bfloat16)
    d = ops.conv_desc(x.shape, co, g.k, g.k, stride=g.stride, off=(-g.pad, -g.pad), out_hw=(oh, ow), act=act,
                      slope=0.2, gain=SQRT2 if act else 1.0, noise_w=1.0 if noise is not None else 0.0, in_dtype=BF16,
                      out_dtype=BF16)
    b = None if bias is None else bias.detach().to(torch.float32).contiguous()
    nz = None if noise is None else noise.detach().to(torch.float32).contiguous()
    sbuf = _stats_buf(d, x.device, want_stats)
    ops.conv_tc(out, x, wp, b, nz, sbuf, d)
    return (out, sbuf) if want_stats else out


def _phase_taps(k, pad, r):
    """Stride-2 input gradient, one axis: input index 2a + r receives dy[a - u] * w[2u + r + pad] for the u with a valid
    tap.  Returns (u_hi, u_lo)."""
    us = [u for u in range(-k, k + 1) if 0 <= 2 * u + r + pad < k]
    return (max(us), min(us)) if us else (0, 1)


def _dgrad_raw(gy, w, g, in_hw, bias=None, act=False, noise=None, want_stats=False):
    """dx NHWC bf16 [n, in_h, in_w, ci] of y = conv(x, w) given dy (NHWC bf16 [n,oh,ow,co]) -- equally the FORWARD of
    F.conv_transpose2d(dy, w, stride, padding) (EqualizedConv2d's fused upscale, stylegan2_layers.py:312-321), for which
    the optional bias / noise / leaky-ReLU epilogue exists."""
    n, oh, ow, co = gy.shape
    ci, k = w.shape[1], g.k
    assert w.shape[0] == co
    out = torch.empty((n, in_hw[0], in_hw[1], ci), device=gy.device, dtype=torch.bfloat16)
    epi = dict(act=act, slope=0.2, gain=SQRT2 if act else 1.0, noise_w=1.0 if noise is not None else 0.0)
    b = None if bias is None else bias.detach().to(torch.float32).contiguous()
    nz = None if noise is None else noise.detach().to(torch.float32).contiguous()
    if g.stride == 1:
        # rotated by 180 degrees and transposed: slot (ty', tx') reads tap (k-1-ty', k-1-tx'), rows = ci
        src = [(k - 1 - ty) * k + (k - 1 - tx) for ty in range(k) for tx in range(k)]
        wp = _pack(w, ("dgrad1", k), 1, k * k, src, True)
        off = -(k - 1) + g.pad
        d = ops.conv_desc(gy.shape, ci, k, k, stride=1, off=(off, off), out_hw=tuple(in_hw), in_dtype=BF16, out_dtype=BF16,
                          **epi)
        sbuf = _stats_buf(d, gy.device, want_stats)
        ops.conv_tc(out, gy, wp, b, nz, sbuf, d)
        return (out, sbuf) if want_stats else out
    assert g.stride == 2 and in_hw[0] % 2 == 0 and in_hw[1] % 2 == 0, "stride-2 input gradient needs even input sizes"
    (hi0, lo0), (hi1, lo1) = _phase_taps(k, g.pad, 0), _phase_taps(k, g.pad, 1)
    J = max(hi0 - lo0, hi1 - lo1) + 1
    his = (hi0, hi1)
    src = []
    for ry in range(2):
        for rx in range(2):
            for jy in range(J):
                ty = 2 * (his[ry] - jy) + ry + g.pad
                for jx in range(J):
                    tx = 2 * (his[rx] - jx) + rx + g.pad
                    src.append(ty * k + tx if (0 <= ty < k and 0 <= tx < k) else -1)
    wp = _pack(w, ("dgrad2", k, g.pad), 4, J * J, src, True)
    d = ops.conv_desc(gy.shape, ci, J, J, stride=1, off=(-hi0, -hi0), out_hw=(in_hw[0] // 2, in_hw[1] // 2), phases=(2, 2),
                      phase_off=(hi0 - hi1, hi0 - hi1), full_hw=tuple(in_hw), in_dtype=BF16, out_dtype=BF16, **epi)
    sbuf = _stats_buf(d, gy.device, want_stats)
    ops.conv_tc(out, gy, wp, b, nz, sbuf, d)
    return (out, sbuf) if want_stats else out


_ws = {}


def _workspace(device, nbytes):
    key = (device.index, torch.cuda.current_stream().cuda_stream)
    buf = _ws.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(int(nbytes), device=device, dtype=torch.uint8)
        _ws[key] = buf
    return buf


def _wgrad_raw(x, gy, g, wshape):
    """dw f32 [co,ci,k,k] of y = conv(x, w): x NHWC bf16 [n,h,w,ci], dy NHWC bf16 [n,oh,ow,co]."""
    n, h, wd, ci = x.shape
    _, oh, ow, co = gy.shape
    assert tuple(wshape) == (co, ci, g.k, g.k), (tuple(wshape), co, ci, g.k)
    d = WgradDesc(n, h, wd, ci, oh, ow, co, g.k, g.k, g.stride, -g.pad, -g.pad)
    L = _lib.load()
    nbytes = L.ppst_conv_wgrad_workspace_bytes(ctypes.byref(d))
    if nbytes < 0:
        check(-1)
    ws = _workspace(x.device, nbytes)
    dw = torch.empty((co, g.k * g.k * ci), device=x.device, dtype=torch.float32)
    with ops._Prof("conv_wgrad", 2.0 * n * oh * ow * co * ci * g.k * g.k, "n%d %dx%d ci%d co%d k%d s%d" % (n, oh, ow, ci, co, g.k, g.stride)):
        check(L.ppst_conv_wgrad(ptr(dw), ptr(x), ptr(gy), ctypes.byref(d), ptr(ws), ws.numel(), stream()))
    return dw.view(co, g.k, g.k, ci).permute(0, 3, 1, 2).contiguous()


def _act_grad_raw(gy, y):
    """dy -> d(pre-activation) of y = lrelu(v, 0.2) * sqrt2, from the OUTPUT's sign (fused_act.py:27-35)."""
    out = torch.empty_like(gy)
    check(_lib.load().ppst_fused_bias_act(ptr(out), ptr(gy), None, ptr(y), gy.numel(), 1, 0, 3, 1, 0.2, SQRT2, dt(gy), stream()))
    return out


# ------------------------------------------------------------------------------------------------ autograd closure
class _Conv(Function):
    @staticmethod
    def forward(ctx, x, w, g):
        x = _nhwc(x)
        ctx.save_for_backward(x, w)
        ctx.g = g
        return _conv_raw(x, w, g)

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        gy = _nhwc(gy)
        dx = _ConvT.apply(gy, w, ctx.g, (x.shape[1], x.shape[2])) if ctx.needs_input_grad[0] else None
        dw = _Wgrad.apply(x, gy, ctx.g, tuple(w.shape)) if ctx.needs_input_grad[1] else None
        return dx, dw, None


class _ConvT(Function):
    """The input gradient of _Conv as a function of (dy, w)."""

    @staticmethod
    def forward(ctx, gy, w, g, in_hw):
        gy = _nhwc(gy)
        ctx.save_for_backward(gy, w)
        ctx.g = g
        return _dgrad_raw(gy, w, g, in_hw)

    @staticmethod
    def backward(ctx, gz):
        gy, w = ctx.saved_tensors
        gz = _nhwc(gz)
        dgy = _Conv.apply(gz, w, ctx.g) if ctx.needs_input_grad[0] else None
        dw = _Wgrad.apply(gz, gy, ctx.g, tuple(w.shape)) if ctx.needs_input_grad[1] else None
        return dgy, dw, None, None


class _Wgrad(Function):
    """The weight gradient of _Conv as a function of (x, dy)."""

    @staticmethod
    def forward(ctx, x, gy, g, wshape):
        x, gy = _nhwc(x), _nhwc(gy)
        ctx.save_for_backward(x, gy)
        ctx.g = g
        return _wgrad_raw(x, gy, g, wshape)

    @staticmethod
    def backward(ctx, gdw):
        x, gy = ctx.saved_tensors
        dx = _ConvT.apply(gy, gdw, ctx.g, (x.shape[1], x.shape[2])) if ctx.needs_input_grad[0] else None
        dgy = _Conv.apply(x, gdw, ctx.g) if ctx.needs_input_grad[1] else None
        return dx, dgy, None, None


class _ActGrad(Function):
    """g -> g * (y > 0 ? 1 : 0.2) * sqrt2; linear in g, y only selects the branch (FusedLeakyReLUFunctionBackward,
    fused_act.py:23-52)."""

    @staticmethod
    def forward(ctx, gy, y):
        ctx.save_for_backward(y)
        return _act_grad_raw(gy.contiguous(), y)

    @staticmethod
    def backward(ctx, gg):
        y, = ctx.saved_tensors
        return _ActGrad.apply(gg.to(y.dtype), y), None


class _ActGradSum(Function):
    """(gp, bias_grad) = (leaky-ReLU gradient of g selected by the saved output y, its per-channel sum over all pixels) in
    ONE pass over the tensor (ppst_act_grad_sum); act = False: gp is g itself and only the sum is computed."""

    @staticmethod
    def forward(ctx, g, y, act):
        g = g.contiguous()
        n, h, w, c = g.shape
        L = _lib.load()
        scratch = ops._get_scratch(g.device, L.ppst_act_grad_sum_scratch_elems(n * h * w, c))
        db = torch.empty(c, device=g.device, dtype=torch.float32)
        gp = torch.empty_like(g) if act else g
        with ops._Prof("act_grad_sum", g.numel() * g.element_size() * (3 if act else 1)):
            check(L.ppst_act_grad_sum(ptr(gp) if act else None, ptr(db), ptr(g), ptr(y) if act else None, ptr(scratch),
                                      n * h * w, c, 0.2, SQRT2, 1 if act else 0, dt(g), stream()))
        ctx.save_for_backward(y)
        ctx.act = act
        return gp, db                   # act = False: gp IS g (autograd treats a returned input as an identity view)

    @staticmethod
    def backward(ctx, ggp, gdb):
        y, = ctx.saved_tensors
        gg = ggp
        if gdb is not None:
            b = gdb.to(y.dtype).view(1, 1, 1, -1)
            gg = b.expand(y.shape) if gg is None else gg + b
        if gg is None:
            return None, None, None
        return (_ActGrad.apply(gg.contiguous(), y) if ctx.act else gg), None, None


class _ConvBiasAct(Function):
    """conv (or transposed conv) + bias + optional per-pixel noise + optional leaky-ReLU(0.2) * sqrt2 in ONE launch
    (EqualConv2d + FusedLeakyReLU, stylegan2_layers.py:497-555; StyledConv's conv + NoiseInjection + bias + activate,
    :439-475).  transposed: x is what a strided convolution with weight w would PRODUCE ([n,oh,ow,co]) and the result is
    at its input resolution `out_hw` -- F.conv_transpose2d(x, w, stride, padding)."""

    @staticmethod
    def forward(ctx, x, w, b, noise, g, act, transposed, out_hw, want_stats):
        x = _nhwc(x)
        if transposed:
            r = _dgrad_raw(x, w, g, out_hw, bias=b, act=act, noise=noise, want_stats=want_stats)
        else:
            r = _conv_raw(x, w, g, bias=b, act=act, noise=noise, want_stats=want_stats)
        y, stats = r if want_stats else (r, torch.empty(0, device=x.device))
        ctx.save_for_backward(x, w, y)
        ctx.g, ctx.act, ctx.transposed = g, act, transposed
        ctx.mark_non_differentiable(stats)
        return y, stats                  # stats: the epilogue's per-tile sum / sum of squares of the stored outputs

    @staticmethod
    def backward(ctx, gy, _gstats):
        x, w, y = ctx.saved_tensors
        gy = _nhwc(gy)
        need = ctx.needs_input_grad
        dx = dw = db = dn = None
        if need[2]:
            gp, db = _ActGradSum.apply(gy, y, ctx.act)           # activation gradient + bias gradient in one pass
        else:
            gp = _ActGrad.apply(gy, y) if ctx.act else gy
        if ctx.transposed:
            if need[0]:
                dx = _Conv.apply(gp, w, ctx.g)
            if need[1]:
                dw = _Wgrad.apply(gp, x, ctx.g, tuple(w.shape))
        else:
            if need[0]:
                dx = _ConvT.apply(gp, w, ctx.g, (x.shape[1], x.shape[2]))
            if need[1]:
                dw = _Wgrad.apply(x, gp, ctx.g, tuple(w.shape))
        if need[3]:
            dn = gp.sum(dim=3, dtype=torch.float32)              # d(noise)[n,y,x] = sum over channels (one read of gp)
        return dx, dw, db, dn, None, None, None, None, None


def conv2d(x, w, stride=1, padding=0, bias=None, act=False, noise=None, want_stats=False):
    """F.conv2d on NHWC bf16 activations (x [n,h,w,ci], w f32 [co,ci,k,k]) with the optional fused epilogue
    (+ bias[co], + noise[n,oh,ow], leaky-ReLU(0.2) * sqrt2); differentiable to second order.  want_stats: also return the
    epilogue's partial InstanceNorm statistics of the output (feed them to instance_norm_mod: saves its reduction pass)."""
    g = Geom(w.shape[2], stride, padding)
    if bias is None and not act and noise is None and not want_stats:
        return _Conv.apply(x, w, g)
    if bias is None:
        bias = torch.zeros(w.shape[0], device=w.device, dtype=torch.float32)
    y, stats = _ConvBiasAct.apply(x, w, bias, noise, g, act, False, None, want_stats)
    return (y, stats) if want_stats else y


def conv_transpose2d(x, w, stride, padding, out_hw, bias=None, act=False, noise=None, want_stats=False):
    """F.conv_transpose2d(x, w, stride=stride, padding=padding) on NHWC bf16: x [n,h,w,c_in], w f32 [c_in, c_out, k, k] (torch's
    transposed-conv weight layout), result [n, out_hw, c_out]; same optional epilogue as conv2d."""
    g = Geom(w.shape[2], stride, padding)
    if bias is None:
        bias = torch.zeros(w.shape[1], device=w.device, dtype=torch.float32)
    y, stats = _ConvBiasAct.apply(x, w, bias, noise, g, act, True, tuple(out_hw), want_stats)
    return (y, stats) if want_stats else y


# ------------------------------------------------------------------------------------------------ thin input layer
def _thin_in_raw(img, w2, bias, act):
    return ops.conv1x1_in_nchw(img.to(torch.float32), w2.detach().to(torch.float32).contiguous(),
                               None if bias is None else bias.detach().to(torch.float32).contiguous(), act,
                               torch.bfloat16)


def _thin_out_raw(gp, w2):
    """NHWC [n,h,w,co] x w2 [co, 3] -> d(image) NCHW f32 [n,3,h,w]."""
    r = ops.conv1x1_thin_out(gp, w2.detach().to(torch.float32).t().contiguous(), None)     # [n,h,w,3] f32
    return ops.nhwc_to_nchw(r)


def _thin_wgrad_raw(gp, img):
    n, h, wd, co = gp.shape
    cin = img.shape[1]
    L = _lib.load()
    scratch = ops._get_scratch(gp.device, L.ppst_conv1x1_thin_wgrad_scratch_elems(n, h * wd, co))
    dw = torch.empty((co, cin), device=gp.device, dtype=torch.float32)
    check(L.ppst_conv1x1_thin_wgrad(ptr(dw), ptr(gp), ptr(img), n, cin, h * wd, co, dt(gp), ptr(scratch), stream()))
    return dw


class _ThinIn(Function):
    """1x1 convolution of the NCHW f32 image (3 channels) into NHWC bf16 features, optional bias + FusedLeakyReLU
    (ConvLayer(3, C, 1), stylegan2_layers.py:497-555)."""

    @staticmethod
    def forward(ctx, img, w2, b, act):
        img = img.contiguous()
        y = _thin_in_raw(img, w2, b, act)
        ctx.save_for_backward(img, w2, y)
        ctx.act = act
        return y

    @staticmethod
    def backward(ctx, gy):
        img, w2, y = ctx.saved_tensors
        gy = _nhwc(gy)
        db = None
        if ctx.needs_input_grad[2]:
            gp, db = _ActGradSum.apply(gy, y, ctx.act)
        else:
            gp = _ActGrad.apply(gy, y) if ctx.act else gy
        dimg = _ThinOut.apply(gp, w2) if ctx.needs_input_grad[0] else None
        dw = _ThinWgrad.apply(gp, img) if ctx.needs_input_grad[1] else None
        return dimg, dw, db, None


class _ThinOut(Function):
    @staticmethod
    def forward(ctx, gp, w2):
        gp = _nhwc(gp)
        ctx.save_for_backward(gp, w2)
        return _thin_out_raw(gp, w2)

    @staticmethod
    def backward(ctx, gimg):
        gp, w2 = ctx.saved_tensors
        gimg = gimg.to(torch.float32).contiguous()
        dgp = _ThinIn.apply(gimg, w2, None, False) if ctx.needs_input_grad[0] else None
        dw = _ThinWgrad.apply(gp, gimg) if ctx.needs_input_grad[1] else None
        return dgp, dw


class _ThinWgrad(Function):
    @staticmethod
    def forward(ctx, gp, img):
        gp, img = _nhwc(gp), img.to(torch.float32).contiguous()
        ctx.save_for_backward(gp, img)
        return _thin_wgrad_raw(gp, img)

    @staticmethod
    def backward(ctx, gdw):
        gp, img = ctx.saved_tensors
        dgp = _ThinIn.apply(img, gdw, None, False) if ctx.needs_input_grad[0] else None
        dimg = _ThinOut.apply(gp, gdw) if ctx.needs_input_grad[1] else None
        return dgp, dimg


def conv1x1_from_image(img, w, bias=None, act=False):
    """ConvLayer(3, C, 1): NCHW f32 image -> NHWC bf16 features.  w f32 [C, 3, 1, 1]."""
    return _ThinIn.apply(img, w.reshape(w.shape[0], w.shape[1]), bias, act)


# ------------------------------------------------------------------------------------------------ residual join, linear
class _AddScale(Function):
    """(a + b) * s on NHWC bf16 (ResBlock's `(out + skip) / sqrt(2)`, stylegan2_layers.py:571-579): one pass."""

    @staticmethod
    def forward(ctx, a, b, s):
        ctx.s = s
        a, b = _nhwc(a), _nhwc(b)
        return ops.affine_apply(a, a.shape[3], res=b, res_mode=1, post_gain=s, out_dtype=torch.bfloat16)

    @staticmethod
    def backward(ctx, g):
        gs = _Scale.apply(g, ctx.s)
        return gs, gs, None


class _Scale(Function):
    @staticmethod
    def forward(ctx, g, s):
        ctx.s = s
        g = _nhwc(g)
        return ops.affine_apply(g, g.shape[3], post_gain=s, out_dtype=torch.bfloat16)

    @staticmethod
    def backward(ctx, gg):
        return _Scale.apply(gg, ctx.s), None


def add_scale(a, b, s):
    return _AddScale.apply(a, b, s)


def _lin_dgrad(gy, w):
    m, n = gy.shape
    k = w.shape[1]
    L = _lib.load()
    scratch = ops._get_scratch(gy.device, L.ppst_linear_dgrad_scratch_elems(m, n, k))
    dx = torch.empty((m, k), device=gy.device, dtype=torch.float32)
    check(L.ppst_linear_dgrad(ptr(dx), ptr(gy), ptr(w), ptr(scratch), m, n, k, stream()))
    return dx


def _lin_wgrad(gy, x):
    m, n = gy.shape
    k = x.shape[1]
    dw = torch.empty((n, k), device=gy.device, dtype=torch.float32)
    check(_lib.load().ppst_linear_wgrad(ptr(dw), ptr(gy), ptr(x), m, n, k, stream()))
    return dw


class _Linear(Function):
    """y = x @ w^T (f32; EqualLinear, stylegan2_layers.py:205-247).  x [m,k], w [n,k]; the batch m is tiny, so forward,
    input gradient and weight gradient are three weight-streaming kernels (ppst_linear / _dgrad / _wgrad).  The three maps
    are bilinear and closed, so the backward is itself differentiable (R1 passes through D's head)."""

    @staticmethod
    def forward(ctx, x, w):
        x, w = x.to(torch.float32).contiguous(), w.to(torch.float32).contiguous()
        ctx.save_for_backward(x, w)
        return ops.linear(x, w)

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        gy = gy.to(torch.float32).contiguous()
        dx = _LinearT.apply(gy, w) if ctx.needs_input_grad[0] else None
        dw = _LinearW.apply(gy, x) if ctx.needs_input_grad[1] else None
        return dx, dw


class _LinearT(Function):
    """dx = gy @ w   (gy [m,n], w [n,k])."""

    @staticmethod
    def forward(ctx, gy, w):
        gy, w = gy.to(torch.float32).contiguous(), w.to(torch.float32).contiguous()
        ctx.save_for_backward(gy, w)
        if w.shape[1] % 4:
            return ops.linear(gy, w.t().contiguous())
        return _lin_dgrad(gy, w)

    @staticmethod
    def backward(ctx, gz):
        gy, w = ctx.saved_tensors
        dgy = _Linear.apply(gz, w) if ctx.needs_input_grad[0] else None
        dw = _LinearW.apply(gy, gz) if ctx.needs_input_grad[1] else None
        return dgy, dw


class _LinearW(Function):
    """dw = gy^T @ x   (gy [m,n], x [m,k])."""

    @staticmethod
    def forward(ctx, gy, x):
        gy, x = gy.to(torch.float32).contiguous(), x.to(torch.float32).contiguous()
        ctx.save_for_backward(gy, x)
        if x.shape[1] % 4 or gy.shape[1] > 65535:
            return ops.linear(gy.t().contiguous(), x.t().contiguous())
        return _lin_wgrad(gy, x)

    @staticmethod
    def backward(ctx, gdw):
        gy, x = ctx.saved_tensors
        dgy = _Linear.apply(x, gdw) if ctx.needs_input_grad[0] else None        # [m,k] @ [n,k]^T
        dx = _LinearT.apply(gy, gdw) if ctx.needs_input_grad[1] else None       # [m,n] @ [n,k]
        return dgy, dx


def linear(x, w):
    return _Linear.apply(x, w)


# ------------------------------------------------------------------------------------------------ norm / modulation
def _plane_dot(g, x):
    """(sum_p g, sum_p g*x) per (sample, channel): f32 [n, c] each (ppst_plane_dot)."""
    n, h, w, c = x.shape
    L = _lib.load()
    scratch = ops._get_scratch(x.device, L.ppst_plane_dot_scratch_elems(n, c, h * w))
    sg = torch.empty((n, c), device=x.device, dtype=torch.float32)
    sgx = torch.empty((n, c), device=x.device, dtype=torch.float32)
    with ops._Prof("plane_dot", 2 * x.numel() * x.element_size()):
        check(L.ppst_plane_dot(ptr(sg), ptr(sgx), ptr(g), ptr(x), ptr(scratch), n, c, h * w, dt(x), stream()))
    return sg, sgx


class _NormMod(Function):
    """y = act( InstanceNorm(x) * (1 + scale) + shift ): nn.InstanceNorm2d (biased variance, eps 1e-5) followed by
    StyleMod (stylegan2_layers.py:361-374,414-437: scale / shift f32 [n, c]) or by a per-channel bias and
    FusedLeakyReLU (ConvLayer with norm='in', :527-531: shift = bias, act = True).  Forward: statistics, coefficients,
    ONE apply pass.  Backward: two fused reductions (ppst_plane_dot) and ONE apply pass dx = alpha*g + beta*x + gamma."""

    @staticmethod
    def forward(ctx, x, scale, shift, act, stats):
        x = _nhwc(x)
        n, h, w, c = x.shape
        N = float(h * w)
        if stats is not None and stats.numel():
            # the producing convolution's epilogue already summed its stored outputs per tile: [2][n][tiles][c]
            part = stats.view(2, n, -1, c).sum(dim=2)
            ssum, ssq = part[0], part[1]
        else:
            st = ops.plane_reduce(x, channels_last=True, want=("sum", "sumsq"))
            ssum, ssq = st["sum"].view(n, c), st["sumsq"].view(n, c)
        mean = ssum / N
        var = (ssq / N - mean * mean).clamp_min(0.0)
        rstd = torch.rsqrt(var + 1e-5)
        a = rstd if scale is None else rstd * (1.0 + scale.to(torch.float32))
        b = -mean * a
        if shift is not None:
            b = b + shift.to(torch.float32)
        y = ops.affine_apply(x, c, a=a.contiguous(), b=b.contiguous(), post_slope=0.2 if act else 1.0,
                             post_gain=SQRT2 if act else 1.0, out_dtype=torch.bfloat16)
        ctx.save_for_backward(x, y if act else None, mean, rstd, scale)
        ctx.act, ctx.has_shift = act, shift is not None
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        x, y, mean, rstd, scale = ctx.saved_tensors
        g = _nhwc(g)
        if ctx.act:
            g = _act_grad_raw(g, y)
        n, h, w, c = x.shape
        N = float(h * w)
        sg, sgx = _plane_dot(g, x)
        m1 = sg / N
        gxh = rstd * (sgx - mean * sg)                     # sum_p g * xhat
        m2 = gxh / N
        alpha = rstd if scale is None else rstd * (1.0 + scale.to(torch.float32))
        beta = -alpha * m2 * rstd
        gamma = -alpha * m1 + alpha * mean * rstd * m2
        dx = None
        if ctx.needs_input_grad[0]:
            dx = ops.affine_apply(g, c, a=alpha.contiguous(), b=gamma.contiguous(), res=x, ra=beta.contiguous(), res_mode=1,
                                  out_dtype=torch.bfloat16)
        dscale = gxh if (scale is not None and ctx.needs_input_grad[1]) else None
        dshift = sg if (ctx.has_shift and ctx.needs_input_grad[2]) else None
        return dx, dscale, dshift, None, None


def instance_norm_mod(x, scale=None, shift=None, act=False, stats=None):
    """stats: the second result of conv2d / conv_transpose2d(..., want_stats=True) that produced x (optional)."""
    return _NormMod.apply(x, scale, shift, act, stats)


class _GapGmp(Function):
    """(global average, global max) over H*W of an NHWC tensor -> two f32 [n, c] (the colour encoder's heads,
    encoder_col.py:39-40,160-161): forward ppst_plane_reduce, backward ONE pass (ppst_pool_backward) after counting the
    pixels that attain the maximum (ties share its gradient, torch.amax's rule)."""

    @staticmethod
    def forward(ctx, x):
        x = _nhwc(x)
        n, h, w, c = x.shape
        st = ops.plane_reduce(x, channels_last=True, want=("sum", "max"))
        gap = st["sum"].view(n, c) / float(h * w)
        gmp = st["max"].view(n, c).contiguous()
        ctx.save_for_backward(x, gmp)
        return gap, gmp

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_gap, g_gmp):
        x, gmp = ctx.saved_tensors
        n, h, w, c = x.shape
        L = _lib.load()
        scratch = ops._get_scratch(x.device, L.ppst_plane_dot_scratch_elems(n, c, h * w))
        cnt = torch.empty((n, c), device=x.device, dtype=torch.float32)
        check(L.ppst_plane_count_eq(ptr(cnt), ptr(x), ptr(gmp), ptr(scratch), n, c, h * w, dt(x), stream()))
        ga = (g_gap.to(torch.float32) / float(h * w)).contiguous()
        gm = (g_gmp.to(torch.float32) / cnt).contiguous()
        dx = torch.empty_like(x)
        check(L.ppst_pool_backward(ptr(dx), ptr(x), ptr(ga), ptr(gm), ptr(gmp), n, c, h * w, dt(x), stream()))
        return dx


def gap_gmp(x):
    return _GapGmp.apply(x)


class _Affine(Function):
    """y = x * a + b per (sample, channel) (GeneratorModulation, generator.py:80-91): a, b f32 [n, c]."""

    @staticmethod
    def forward(ctx, x, a, b):
        x = _nhwc(x)
        ctx.save_for_backward(x, a)
        return ops.affine_apply(x, x.shape[3], a=a.to(torch.float32).contiguous(), b=b.to(torch.float32).contiguous(),
                                out_dtype=torch.bfloat16)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        x, a = ctx.saved_tensors
        g = _nhwc(g)
        sg, sgx = _plane_dot(g, x)
        dx = ops.affine_apply(g, g.shape[3], a=a.to(torch.float32).contiguous(), out_dtype=torch.bfloat16) \
            if ctx.needs_input_grad[0] else None
        return dx, sgx, sg


def channel_affine(x, a, b):
    return _Affine.apply(x, a, b)


class _ReflectPad(Function):
    """nn.ReflectionPad2d on NHWC (the encoders' ConvLayer(reflection_pad=True), stylegan2_layers.py:142-164,525):
    forward = one ppst_affine_apply pass with reflect addressing; backward = its adjoint as a gather
    (ppst_reflect_pad_backward)."""

    @staticmethod
    def forward(ctx, x, pad):
        x = _nhwc(x)
        ctx.pad = pad
        return ops.affine_apply(x, x.shape[3], pad=tuple(pad), pad_mode=1, out_dtype=torch.bfloat16)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        t, b, l, r = ctx.pad
        g = _nhwc(g)
        n, c = g.shape[0], g.shape[3]
        h, w = g.shape[1] - t - b, g.shape[2] - l - r
        gin = torch.empty((n, h, w, c), device=g.device, dtype=torch.bfloat16)
        check(_lib.load().ppst_reflect_pad_backward(ptr(gin), ptr(g), n, h, w, c, t, b, l, r, dt(g), stream()))
        return gin, None


def reflect_pad(x, pad):
    """pad = (top, bottom, left, right)."""
    return _ReflectPad.apply(x, tuple(int(p) for p in pad))


class _Blur(Function):
    """Blur (upfirdn2d, up = down = 1) on NHWC bf16: pads (p0 before, p1 after) on both axes.  Linear, so the backward is the
    adjoint blur (unflipped kernel, pads k-1-p) and the backward of that is the blur again (second order for R1)."""

    @staticmethod
    def forward(ctx, x, k, p0, p1, flip):
        x = _nhwc(x)
        n, h, w, c = x.shape
        kh, kw = k.shape
        ctx.save_for_backward(k)
        ctx.cfg = (p0, p1, flip)
        out = torch.empty((n, h + p0 + p1 - kh + 1, w + p0 + p1 - kw + 1, c), device=x.device, dtype=torch.bfloat16)
        with ops._Prof("blur_nhwc", (x.numel() + out.numel()) * 2):
            check(_lib.load().ppst_blur_nhwc(ptr(out), ptr(x), ptr(k), n, h, w, c, kh, kw, p0, p1, p0, p1, 1 if flip else 0,
                                             dt(x), stream()))
        return out

    @staticmethod
    def backward(ctx, g):
        k, = ctx.saved_tensors
        p0, p1, flip = ctx.cfg
        kh = k.shape[0]
        return _Blur.apply(g, k, kh - 1 - p0, kh - 1 - p1, not flip), None, None, None, None


def blur(x, k, pad):
    """upfirdn2d(x, k, pad=(p0, p1)) for an NHWC bf16 tensor and a square FIR kernel k (f32 [kh, kh], device)."""
    return _Blur.apply(x, k, int(pad[0]), int(pad[1]), True)


class _Upsample2x(Function):
    """x2 upsampling on NHWC bf16: bilinear, align_corners=False (the skip of UpsamplingResnetBlock, generator.py:63-77:
    forward ppst_resample) or nearest (Upscale2d, stylegan2_layers.py:88-105: forward ppst_affine_apply with up = 2); the
    backward is the adjoint as a deterministic gather (ppst_upsample2x_backward)."""

    @staticmethod
    def forward(ctx, x, nearest):
        x = _nhwc(x)
        ctx.shape, ctx.nearest = x.shape, nearest
        if nearest:
            return ops.affine_apply(x, x.shape[3], up=2, out_dtype=torch.bfloat16)
        return ops.resample(x, (2 * x.shape[1], 2 * x.shape[2]), 0, out_dtype=torch.bfloat16)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        g = _nhwc(g)
        n, h, w, c = ctx.shape
        gin = torch.empty((n, h, w, c), device=g.device, dtype=torch.bfloat16)
        check(_lib.load().ppst_upsample2x_backward(ptr(gin), ptr(g), n, h, w, c, 1 if ctx.nearest else 0, dt(g), stream()))
        return gin, None


def nearest2x(x):
    return _Upsample2x.apply(x, True)


def bilinear2x(x):
    return _Upsample2x.apply(x, False)


# ------------------------------------------------------------------------------------------------ discriminator
class _Fold(Function):
    """Stencil product of a conv weight with the blur kernel (ppst_fold_fir) -- linear in w, so the backward is the
    adjoint map and the backward of the adjoint is the map itself."""

    @staticmethod
    def forward(ctx, w, k, adjoint, kh, kw):
        w = w.to(torch.float32).contiguous()
        co, ci = w.shape[:2]
        fh, fw = k.shape
        ctx.save_for_backward(k)
        ctx.adjoint, ctx.khw = adjoint, (kh, kw)
        out = torch.empty((co, ci, kh, kw) if adjoint else (co, ci, kh + fh - 1, kw + fw - 1), device=w.device,
                          dtype=torch.float32)
        check(_lib.load().ppst_fold_fir(ptr(out), ptr(w), ptr(k), co * ci, kh, kw, fh, fw, 1 if adjoint else 0, stream()))
        return out

    @staticmethod
    def backward(ctx, g):
        k, = ctx.saved_tensors
        return _Fold.apply(g, k, not ctx.adjoint, ctx.khw[0], ctx.khw[1]), None, None, None, None


def _fold(w, blur_k):
    """Blur (upfirdn2d, zero pad) followed by a stride-2 convolution == ONE stride-2 convolution with the stencil
    product of the two kernels; [co,ci,kh,kw] -> [co,ci,kh+fh-1,kw+fw-1]."""
    return _Fold.apply(w, blur_k, False, w.shape[2], w.shape[3])


class Discriminator(torch.nn.Module):
    """StyleGAN2 discriminator of the reference (stylegan2_layers.py:582-649, wrapped by models/networks/discriminator.py:
    5-32) with the reference's parameter names and shapes: `load_state_dict` takes the reference's `stylegan2_D.*` keys
    unchanged.  forward(x): NCHW f32 image [n,3,S,S] -> [n,1] f32.

    Layer by layer: ConvLayer(3,C,1)+FusedLeakyReLU = one thin kernel; every ConvLayer = one tcgen05 convolution with the
    bias + leaky-ReLU epilogue; Blur([1,3,3,1]) = ppst_blur_nhwc in front of the 3x3 / 1x1 stride-2 convolution
    (fold_blur=True: folded into ONE 6x6 / 4x4 stride-2 convolution instead -- saves the blur's HBM passes but quadruples
    those layers' tensor FLOPs, slower at these channel counts); the residual join is one pass; the two EqualLinear layers
    run on ppst_linear."""

    def __init__(self, size, channel_multiplier=2, blur_kernel=(1, 3, 3, 1), fold_blur=False):
        super().__init__()
        nn = torch.nn
        self.fold_blur = fold_blur
        channels = {4: 512, 8: 512, 16: min(512, int(512 * channel_multiplier)), 32: min(512, int(512 * channel_multiplier)),
                    64: int(256 * channel_multiplier), 128: int(128 * channel_multiplier), 256: int(64 * channel_multiplier),
                    512: int(32 * channel_multiplier), 1024: int(16 * channel_multiplier)}
        original_size = size
        size = 2 ** int(round(math.log(size, 2)))
        log_size = int(math.log(size, 2))
        self.names = []
        P = nn.ParameterDict()

        def param(name, t):
            P[name.replace(".", "/")] = nn.Parameter(t)

        c = channels[size]
        param("convs.0.Conv.weight", torch.randn(c, 3, 1, 1))
        param("convs.0.Act.bias", torch.zeros(c))
        self.blocks = []
        for i in range(log_size, 2, -1):
            co = channels[2 ** (i - 1)]
            name = "convs." + (str(9 - i) if i <= 8 else "%dx%d" % (2 ** i, 2 ** i))
            param(name + ".conv1.Conv.weight", torch.randn(c, c, 3, 3))
            param(name + ".conv1.Act.bias", torch.zeros(c))
            param(name + ".conv2.Conv.weight", torch.randn(co, c, 3, 3))
            param(name + ".conv2.Act.bias", torch.zeros(co))
            param(name + ".skip.Conv.weight", torch.randn(co, c, 1, 1))
            self.blocks.append((name, c, co))
            c = co
        param("final_conv.Conv.weight", torch.randn(channels[4], c, 3, 3))
        param("final_conv.Act.bias", torch.zeros(channels[4]))
        self.side = int(4 * original_size / size)
        param("final_linear.0.weight", torch.randn(channels[4], channels[4] * self.side ** 2))
        param("final_linear.0.bias", torch.zeros(channels[4]))
        param("final_linear.1.weight", torch.randn(1, channels[4]))
        param("final_linear.1.bias", torch.zeros(1))
        self.P = P
        k = torch.tensor(list(blur_kernel), dtype=torch.float32)
        k2 = k[None, :] * k[:, None]
        self.register_buffer("blur", k2 / k2.sum(), persistent=False)
        self.c_last = channels[4]

    # reference state_dict layout <-> ParameterDict keys ('.' is not allowed in ParameterDict names)
    def load_reference_state_dict(self, sd, prefix="stylegan2_D."):
        own = {k.replace("/", "."): v for k, v in self.P.items()}
        with torch.no_grad():
            for k, v in own.items():
                src = sd[prefix + k]
                assert tuple(src.shape) == tuple(v.shape), (k, tuple(src.shape), tuple(v.shape))
                v.copy_(src)
        return self

    def reference_state_dict(self, prefix="stylegan2_D."):
        """Key-complete for the reference's loader: the parameters plus the Blur.kernel buffers of every ResBlock
        (convs.N.conv2.Blur.kernel / convs.N.skip.Blur.kernel, stylegan2_layers.py:142-152)."""
        out = {prefix + k.replace("/", "."): v.detach().clone() for k, v in self.P.items()}
        for name, _, _ in self.blocks:
            for branch in ("conv2", "skip"):
                out[prefix + name + "." + branch + ".Blur.kernel"] = self.blur.detach().clone()
        return out

    def p(self, name):
        return self.P[name.replace(".", "/")]

    def forward(self, img):
        p = self.p
        w = p("convs.0.Conv.weight")
        x = conv1x1_from_image(img, w * (1.0 / math.sqrt(3.0)), p("convs.0.Act.bias"), act=True)
        for name, c, co in self.blocks:
            s3 = 1.0 / math.sqrt(c * 9)
            y = conv2d(x, p(name + ".conv1.Conv.weight") * s3, 1, 1, p(name + ".conv1.Act.bias"), act=True)
            if self.fold_blur:
                w6 = _fold(p(name + ".conv2.Conv.weight") * s3, self.blur)                       # [co, c, 6, 6]
                y = conv2d(y, w6, 2, 2, p(name + ".conv2.Act.bias"), act=True)
                w4 = _fold(p(name + ".skip.Conv.weight") * (1.0 / math.sqrt(c)), self.blur)     # [co, c, 4, 4]
                sk = conv2d(x, w4, 2, 1)
            else:
                # Blur(pad (2,2)) -> 3x3 stride 2 and Blur(pad (1,1)) -> 1x1 stride 2, with ONE more zero row / column of
                # padding after the image: no tap of the strided convolution reads the extra blurred row, and the even size
                # lets the input gradient run as four sub-pixel phases
                y = conv2d(blur(y, self.blur, (2, 3)), p(name + ".conv2.Conv.weight") * s3, 2, 0,
                           p(name + ".conv2.Act.bias"), act=True)
                sk = conv2d(blur(x, self.blur, (1, 2)), p(name + ".skip.Conv.weight") * (1.0 / math.sqrt(c)), 2, 0)
            x = add_scale(y, sk, 1.0 / SQRT2)
        c = x.shape[3]
        x = conv2d(x, p("final_conv.Conv.weight") * (1.0 / math.sqrt(c * 9)), 1, 1, p("final_conv.Act.bias"), act=True)
        n, h, wd, c = x.shape
        # the reference flattens NCHW (`out.view(batch, -1)`): permute the first linear's columns to NHWC order instead
        w0 = p("final_linear.0.weight").view(-1, c, h, wd).permute(0, 2, 3, 1).reshape(-1, h * wd * c)
        from .stylegan2_op import fused_leaky_relu
        hdn = linear(x.reshape(n, -1), w0 * (1.0 / math.sqrt(w0.shape[1])))
        hdn = fused_leaky_relu(hdn, p("final_linear.0.bias"))
        w1 = p("final_linear.1.weight")
        return linear(hdn, w1 * (1.0 / math.sqrt(w1.shape[1]))) + p("final_linear.1.bias")


# ------------------------------------------------------------------------------------------------ losses / step
def gan_loss(pred, real):
    """models/networks/loss.py:11-18 (least squares)."""
    return torch.mean((pred - 1) ** 2) if real else torch.mean(pred ** 2)


def d_losses(D, real, fake, lambda_gan=1.0):
    """compute_image_discriminator_losses, stage 1 (models/ppst_model.py:68-92): D_real + 0.5 * D_rec."""
    # D has no cross-sample statistics (the minibatch-stddev layer is commented out upstream, stylegan2_layers.py:630-637):
    # one pass over [real; fake] gives the same per-sample predictions as the reference's two passes and prepares the
    # weights once
    pred = D(torch.cat([real, fake], dim=0))
    nr = real.shape[0]
    return {"D_real": gan_loss(pred[:nr], True) * lambda_gan, "D_rec": gan_loss(pred[nr:], False) * (0.5 * lambda_gan)}


def r1_loss(D, real, lambda_r1=10.0):
    """compute_R1_loss (models/ppst_model.py:139-156): 0.5 * lambda * |d D(real) / d real|^2 per sample."""
    real = real.detach().requires_grad_(True)
    pred = D(real).sum()
    grad_real, = torch.autograd.grad(outputs=pred, inputs=[real], create_graph=True, retain_graph=True)
    return grad_real.pow(2).sum(dim=(1, 2, 3)) * (lambda_r1 * 0.5)


def allreduce_grads_(params, world=None):
    """See sharding.allreduce_grads_ (one flat bucket, one all-reduce, averaged)."""
    from .sharding import allreduce_grads_ as f
    return f(params, world)


def capture_step(body, zero_grad, generators=(), warmup=2):
    """Capture one training step -- forward, autograd backward (its tape is replayed as recorded), weight preparation,
    noise draws, gradient all-reduce, optimiser (must be `capturable=True`) -- into a CUDA graph and return it; replaying
    costs one host call instead of the ~2500 launches of a generator step.  `body()` must read its batch from tensors that
    stay alive (copy_ the next batch into them before graph.replay()); `zero_grad()` is called before every warm-up
    iteration and before the capture (use set_to_none=True so the gradients are allocated inside the graph's pool);
    `generators`: torch.Generator objects drawn from inside `body` (registered so every replay advances them)."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(warmup):
            zero_grad()
            body()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    for gen in generators:
        if gen is not None:
            graph.register_generator_state(gen)
    zero_grad()
    with torch.cuda.graph(graph):
        body()
    return graph
