"""Network-level executors for the PPST swap / inference path on the ppst_b200 kernels.

The three networks consume the reference's state_dicts unchanged and are re-expressed
as kernel sequences over NHWC activations:

  conv (tcgen05 implicit GEMM, or the CUDA-core one)  ->  [per-(n,c) statistics, free in the conv epilogue]
       ->  ppst_in_coeffs (InstanceNorm + StyleMod / bias collapse to ONE affine per (n,c))
       ->  ppst_affine_apply (affine + activation + residual/skip + padding/upsampling of the NEXT conv's input)

Two numeric modes:
  "fp32": fp32 activations, exact two-pass statistics, CUDA-core convolutions  (tolerance: max-abs 1e-4/layer)
  "bf16": bf16 activations/weights, fp32 accumulation on the tensor cores      (rel-L2 <= 1e-2/layer, PSNR >= 45 dB)

Reference call sites are cited per method (paths relative to the reference repository root).
"""
import math

import torch

from . import ops, packing
from ._lib import BF16, F32
from .packing import SQRT2, make_spec, make_upscale_spec

INV_SQRT2 = 1.0 / SQRT2
HEAD_CH = [(256, 256), (256, 256), (256, 384), (384, 512)]      # generator.py:156-165
UP_CH = [(16, 512, 512), (32, 512, 256), (64, 256, 128)]        # generator.py:167-175
E_CH = [(32, 64), (64, 128), (128, 256)]                        # encoder_con.py:36-41 / encoder_col.py:31-36
LEVELS = ["9", "0", "1", "2"]


class Engine:
    def __init__(self, sdE1, sdE2, sdG, device="cuda", mode="bf16"):
        assert mode in ("bf16", "fp32")
        self.device = torch.device(device)
        self.mode = mode
        self.adt = torch.bfloat16 if mode == "bf16" else torch.float32
        self.acode = BF16 if mode == "bf16" else F32
        self.use_tc = mode == "bf16" and ops.conv_tc_available()
        dev = self.device
        f32 = lambda t: t.detach().to(dev, torch.float32).contiguous()   # noqa: E731
        self.E1 = self._pack_encoder(sdE1, "DownToSpatialCode", norm=True)
        self.E1["tsc0"] = make_spec(sdE1["ToSpatialCode.0.Conv.weight"], dev, scale=1 / math.sqrt(256))
        self.E1["tsc0_actb"] = f32(sdE1["ToSpatialCode.0.Act.bias"])
        self.E1["tsc1"] = make_spec(sdE1["ToSpatialCode.1.Conv.weight"], dev, scale=1 / math.sqrt(256),
                                    bias=sdE1["ToSpatialCode.1.Conv.bias"])
        self.E2 = self._pack_encoder(sdE2, "DownToGlobalCode1", norm=False)
        for lvl in LEVELS:
            w = sdE2["conv1x1_%s.weight" % lvl]
            self.E2["c1x1_%s" % lvl] = (f32(w.reshape(w.shape[0], -1)), f32(sdE2["conv1x1_%s.bias" % lvl]))
            self.E2["proj_%s" % lvl] = [(f32(sdE2["projector%s.%d.weight" % (lvl, j)]),
                                         f32(sdE2["projector%s.%d.bias" % (lvl, j)])) for j in (1, 3, 5)]
        self.G = self._pack_generator(sdG)
        self.padded = True           # produce tensors inside their next consumer's padding border (no padding passes);
                                     # False = the plain sequence with explicit padding passes (bit-identical results)
        self._keep = []              # inputs of forked sections, referenced until the join
        self._side = {}              # side streams for the latency-bound vector chains (code heads, style linears)
        self._side_live = set()

    # ------------------------------------------------------------------------------------------ side streams
    class _Fork:
        """`with engine.fork(i):` -- the enclosed launches go to side stream i, ordered after everything already on the
        current stream; engine.join() makes the current stream wait for every forked section.  Used for the chains of
        tiny dependent kernels (GAP||GMP -> 1x1 -> projector MLP -> normalise; the style linears) whose cost is latency,
        not work: they overlap the next convolutions instead of sitting between them.  Fork/join discipline keeps the
        caching allocator safe in eager mode and under graph capture: a side section never starts before the main
        stream's earlier work, and its inputs stay referenced until the join."""

        def __init__(self, eng, idx):
            self.eng, self.idx = eng, idx

        def __enter__(self):
            e = self.eng
            st = e._side.get(self.idx)
            if st is None:
                st = e._side[self.idx] = torch.cuda.Stream(device=e.device)
            st.wait_stream(torch.cuda.current_stream(e.device))
            e._side_live.add(self.idx)
            self.ctx = torch.cuda.stream(st)
            self.ctx.__enter__()
            return self

        def __exit__(self, *exc):
            return self.ctx.__exit__(*exc)

    def fork(self, idx=0):
        return Engine._Fork(self, idx)

    def join(self):
        cur = torch.cuda.current_stream(self.device)
        for idx in sorted(self._side_live):
            cur.wait_stream(self._side[idx])
        self._side_live.clear()
        self._keep = []

    # ------------------------------------------------------------------------------------------ packing
    def _pack_encoder(self, sd, down, norm):
        """FromRGB + 3 ResBlocks (stylegan2_layers.py:497-579).  The 3x3 [1,2,1] blur of the two downsampling
        branches is folded into their convolutions: conv2 becomes a 5x5 stride-2 conv over the reflect-padded
        input, skip a 3x3 stride-2 conv with one row/col of zero padding on the top/left."""
        dev = self.device
        P = {"norm": norm}
        P["from_rgb"] = make_spec(sd["FromRGB.Conv.weight"], dev, scale=1 / math.sqrt(3), bias=sd["FromRGB.Act.bias"],
                                  act=True)
        for i, (ci, co) in zip((1, 2, 4), E_CH):
            P["res%d" % i] = self._pack_res_block(sd, "%s.ResBlockDownBy%d." % (down, i), norm)
        return P

    def _pack_res_block(self, sd, p, norm):
        """One ResBlock (stylegan2_layers.py:559-579) with reflection padding and a [1,2,1] blur, keys under prefix p."""
        dev = self.device
        ci = sd[p + "conv1.Conv.weight"].shape[1]
        k = sd[p + "conv2.Blur.kernel"].to(torch.float32)
        w1 = sd[p + "conv1.Conv.weight"]
        w2 = packing.fold_fir(sd[p + "conv2.Conv.weight"].to(torch.float32) / math.sqrt(ci * 9), k)
        ws = packing.fold_fir(sd[p + "skip.Conv.weight"].to(torch.float32) / math.sqrt(ci),
                              sd[p + "skip.Blur.kernel"].to(torch.float32))
        b1, b2 = sd[p + "conv1.Act.bias"], sd[p + "conv2.Act.bias"]
        blk = {}
        # norm=True: IN sits between conv and activation, so bias+act move to the apply pass
        blk["conv1"] = make_spec(w1, dev, scale=1 / math.sqrt(ci * 9), bias=None if norm else b1, act=not norm)
        blk["conv2"] = make_spec(w2, dev, stride=2, off=(0, 0), bias=None if norm else b2, act=not norm,
                                 out_hw=lambda h, w: ((h - 5) // 2 + 1, (w - 5) // 2 + 1))
        blk["skip"] = make_spec(ws, dev, stride=2, off=(-1, -1),
                                out_hw=lambda h, w: ((h - 2) // 2 + 1, (w - 2) // 2 + 1))
        blk["conv2"].extra["alg_taps"] = 9          # the reference op is a 3x3 / 1x1 convolution after a Blur
        blk["skip"].extra["alg_taps"] = 1
        blk["b1"] = b1.detach().to(dev, torch.float32).contiguous()
        blk["b2"] = b2.detach().to(dev, torch.float32).contiguous()
        return blk

    def _pack_styled(self, sd, p, upsample_fused_ok=True):
        """StyledConv (stylegan2_layers.py:439-475): all three additive per-channel terms (conv.bias, .bias,
        activate.bias) are pre-summed; noise and leaky-ReLU*sqrt2 run in the conv epilogue."""
        dev = self.device
        bias = sd[p + "conv.bias"].float() + sd[p + "bias"].float().reshape(-1) + sd[p + "activate.bias"].float()
        nw = float(sd[p + "noise.weight"].reshape(-1)[0])
        d = {"plain": make_spec(sd[p + "conv.weight"], dev, bias=bias, act=True, pad=1, noise_w=nw),
             "lin_w": sd[p + "epi1.style_mod.lin.weight"].detach().to(dev, torch.float32).contiguous(),
             "lin_b": sd[p + "epi1.style_mod.lin.bias"].detach().to(dev, torch.float32).contiguous(),
             "noise_w": nw}
        return d

    def _pack_generator(self, sd):
        dev = self.device
        f32 = lambda t: t.detach().to(dev, torch.float32).contiguous()   # noqa: E731
        G = {}
        for n in ("scale", "bias"):
            G["scm_" + n] = (f32(sd["SpatialCodeModulation.%s.weight" % n]), f32(sd["SpatialCodeModulation.%s.bias" % n]))
        for i, (ci, co) in enumerate(HEAD_CH):
            p = "HeadResnetBlock%d." % i
            blk = {"conv1": self._pack_styled(sd, p + "conv1."), "conv2": self._pack_styled(sd, p + "conv2.")}
            if ci != co:
                blk["skip"] = make_spec(sd[p + "skip.Conv.weight"], dev, scale=1 / math.sqrt(ci))
            G["head%d" % i] = blk
        for res, ci, co in UP_CH:
            p = "UpsamplingResBlock%d." % res
            c1 = self._pack_styled(sd, p + "conv1.")
            c1["fused_up"] = make_upscale_spec(sd[p + "conv1.conv.weight"], dev, bias=c1["plain"].bias, act=True,
                                               noise_w=c1["noise_w"])
            blk = {"conv1": c1, "conv2": self._pack_styled(sd, p + "conv2.")}
            if ci != co:
                blk["skip"] = make_spec(sd[p + "skip.Conv.weight"], dev, scale=1 / math.sqrt(ci),
                                        bias=sd[p + "skip.Act.bias"], act=True)
            G["up%d" % res] = blk
        G["rgb"] = make_spec(sd["ToRGB.conv.weight"], dev, scale=1 / math.sqrt(128),
                             bias=sd["ToRGB.conv.bias"].float() + sd["ToRGB.bias"].float().reshape(-1), tc=False)
        G["rgb_lin"] = (f32(sd["ToRGB.epi1.style_mod.lin.weight"]), f32(sd["ToRGB.epi1.style_mod.lin.bias"]))
        for name, k in (("layer32", 3), ("layer64", 3), ("layer128", 3), ("layer256", 1)):
            G[name] = {"k": k,
                       "c1": make_spec(sd[name + ".2.weight"], dev, bias=sd[name + ".2.bias"]),
                       "a1": float(sd[name + ".4.weight"].reshape(-1)[0]),
                       "c2": make_spec(sd[name + ".6.weight"], dev, bias=sd[name + ".6.bias"]),
                       "a2": float(sd[name + ".8.weight"].reshape(-1)[0])}
        for name in ("layert.0", "layert.1", "layert.2", "layert1.0"):
            G[name] = {"c1": make_spec(sd[name + ".conv1.weight"], dev, bias=sd[name + ".conv1.bias"]),
                       "c2": make_spec(sd[name + ".conv2.weight"], dev, bias=sd[name + ".conv2.bias"]),
                       "a": float(sd[name + ".prelu.weight"].reshape(-1)[0])}
        G["layert1.1"] = make_spec(sd["layert1.1.weight"], dev, bias=sd["layert1.1.bias"])
        # every style / modulation vector of a pass comes out of ONE batched linear per texture code
        # (EqualLinear and EqualizedLinear(use_wscale) share the run-time scale in_dim**-0.5)
        groups = {-1: [("scm_scale",) + G["scm_scale"], ("scm_bias",) + G["scm_bias"]], -2: [], -3: [], -4: [], 0: []}
        for i in range(len(HEAD_CH)):
            for c_ in ("conv1", "conv2"):
                P = G["head%d" % i][c_]
                groups[-1].append(("head%d.%s" % (i, c_), P["lin_w"], P["lin_b"]))
        for j, (res, _, _) in enumerate(UP_CH):
            for c_ in ("conv1", "conv2"):
                P = G["up%d" % res][c_]
                groups[-2 - j].append(("up%d.%s" % (res, c_), P["lin_w"], P["lin_b"]))
        groups[0].append(("rgb",) + G["rgb_lin"])
        G["style_groups"] = []
        for idx, items in groups.items():
            w = torch.cat([it[1] for it in items], 0).contiguous()
            b = torch.cat([it[2] for it in items], 0).contiguous()
            slices, o = [], 0
            for it in items:
                slices.append((it[0], o, it[1].shape[0]))
                o += it[1].shape[0]
            G["style_groups"].append((idx, w, b, slices))
        return G

    # ------------------------------------------------------------------------------------------ primitives
    def conv(self, x, spec, noise=None, out=None, c_off=0, out_dtype=None, want_stats=False, w_per_image=None,
             bias_per_image=None, out_pad=0, valid=False, in_border=0, out_pad_trim=0):
        """x: dense NHWC.  Returns (y, stats) with stats = None | dict(partials...) for in_coeffs.
        w_per_image: bf16 [n, cout, K] per-sample weights (ops.scale_weights / modulate_weights) instead of spec.w16,
        bias_per_image: f32 [n, cout] instead of spec.bias (tensor-core path only).
        out_pad (/ out_pad_trim): the result goes into the interior of a new [n, fh+2p-t, fw+2p-t, cout] buffer whose
        top / left border is p and bottom / right border p - t pixels (border left to the caller).
        valid: x is already padded -- tap (0,0) of output (0,0) reads x[0,0] and the output shrinks by k-1.
        in_border: x carries a border of that many pixels which this convolution must NOT see: it reads the interior
        through row / image pitches and pads with zeros as usual (tensor-core path only)."""
        n, h, w, cin = x.shape
        assert cin == spec.cin, (cin, spec.cin)
        pitches = {}
        if in_border:
            h, w = h - 2 * in_border, w - 2 * in_border
            pitches = {"in_row_pitch": x.shape[2] * cin, "in_img_pitch": x.shape[1] * x.shape[2] * cin}
            x = x[:, in_border:in_border + h, in_border:in_border + w, :]      # data_ptr() = first interior pixel
        off = spec.off
        if valid:
            assert spec.stride == 1 and spec.phases == (1, 1)
            off = (0, 0)
            oh, ow = h - spec.kh + 1, w - spec.kw + 1
        else:
            oh, ow = spec.out_hw(h, w)
        fh, fw = oh * spec.phases[0], ow * spec.phases[1]
        odt = out_dtype or self.adt
        if out is None:
            out = torch.empty((n, fh + 2 * out_pad - out_pad_trim, fw + 2 * out_pad - out_pad_trim, spec.cout),
                              device=x.device, dtype=odt)
        tc = self.use_tc and spec.w16 is not None and x.dtype == torch.bfloat16 and out.dtype == torch.bfloat16
        code_in = BF16 if x.dtype == torch.bfloat16 else F32
        code_out = BF16 if out.dtype == torch.bfloat16 else F32
        d = ops.conv_desc((n, h, w, cin), spec.cout, spec.kh, spec.kw, stride=spec.stride, off=off,
                          out_hw=(oh, ow), phases=spec.phases, phase_off=spec.phase_off, full_hw=(fh, fw),
                          cout_stride=out.shape[3], c_off=c_off, act=spec.act, slope=0.2,
                          gain=SQRT2 if spec.act else 1.0, noise_w=spec.noise_w if noise is not None else 0.0,
                          in_dtype=code_in, out_dtype=code_out, w_img_rows=spec.cout if w_per_image is not None else 0,
                          out_pad=out_pad, bias_img_stride=spec.cout if bias_per_image is not None else 0,
                          out_pad_trim=out_pad_trim, **pitches)
        stats = None
        assert (w_per_image is None and bias_per_image is None and out_pad == 0 and in_border == 0) or tc, \
            "per-sample weights / bias, padded outputs and interior views need the tensor-core path"
        if tc:
            sbuf = None
            if want_stats:
                sbuf = torch.empty(ops.conv_tc_stats_elems(d), device=x.device, dtype=torch.float32)
            ops.conv_tc(out, x, spec.w16 if w_per_image is None else w_per_image,
                        spec.bias if bias_per_image is None else bias_per_image, noise, sbuf, d,
                        alg_taps=spec.extra.get("alg_taps"))
            if want_stats:
                half = sbuf.numel() // 2
                stats = {"sum": sbuf[:half], "sumsq": sbuf[half:], "n_partials": half // (n * spec.cout),
                         "count": fh * fw}
        else:
            ops.conv_simt(out, x, spec.w32, spec.bias, noise, d)
        return out, stats

    def coeffs(self, y, stats=None, style=None, bias=None, c=None):
        """InstanceNorm (biased var, eps 1e-5) of NHWC `y` (+ StyleMod style [n,2c], + per-channel bias)
        -> (a, b) f32 [n, c]."""
        n, h, w, cs = y.shape
        c = c or cs
        if stats is not None:
            return ops.in_coeffs(n, c, stats["sum"], stats["sumsq"], count=stats["count"], style=style, bias=bias,
                                 n_partials=stats["n_partials"])
        if self.mode == "fp32" or y.dtype == torch.float32:
            r1 = ops.plane_reduce(y, True, want=("sum",))
            mean = (r1["sum"].view(n, c) / float(h * w)).contiguous()
            r2 = ops.plane_reduce(y, True, shift=mean, want=("sum", "sumsq"))
            return ops.in_coeffs(n, c, r2["sum"].view(n, c), r2["sumsq"].view(n, c), count=h * w, shift=mean,
                                 style=style, bias=bias)
        r = ops.plane_reduce(y, True, want=("sum", "sumsq"))
        return ops.in_coeffs(n, c, r["sum"].view(n, c), r["sumsq"].view(n, c), count=h * w, style=style, bias=bias)

    def apply(self, x, a=None, b=None, **kw):
        return ops.affine_apply(x, x.shape[3], a, b, **kw)

    def to_nhwc(self, x_nchw, dtype=None):
        return ops.nchw_to_nhwc(x_nchw.to(self.device, torch.float32), dtype=dtype or self.adt)

    def from_rgb(self, P, image_nchw, out_border=0):
        """FromRGB = ConvLayer(3, 32, 1) + FusedLeakyReLU straight from the NCHW fp32 image (layout change fused).
        out_border = 1: into the interior of a buffer whose 1-pixel border is then filled by reflection (the padding of
        the first ResBlock's conv1, stylegan2_layers.py:517-523)."""
        spec = P["from_rgb"]
        x = image_nchw.to(self.device, torch.float32)
        y = ops.conv1x1_in_nchw(x, spec.w32[0], spec.bias, True, self.adt, out_border=out_border)
        if out_border:
            ops.pad_border(y, (out_border,) * 4, 1)
        return y

    # ------------------------------------------------------------------------------------------ encoders
    def _res_block(self, blk, x, norm):
        """ResBlock (stylegan2_layers.py:559-579) with reflection padding, on NHWC x."""
        xp = self.apply(x, pad=(1, 1, 1, 1), pad_mode=1)
        c1, st1 = self.conv(xp, blk["conv1"], want_stats=norm)
        if norm:
            a, b = self.coeffs(c1, st1, bias=blk["b1"])
            h1p = self.apply(c1, a, b, pre_slope=0.2, pre_gain=SQRT2, pad=(2, 1, 2, 1), pad_mode=1)
        else:
            h1p = self.apply(c1, pad=(2, 1, 2, 1), pad_mode=1)
        c2, st2 = self.conv(h1p, blk["conv2"], want_stats=norm)
        cs, sts = self.conv(x, blk["skip"], want_stats=norm)
        if norm:
            a2, b2 = self.coeffs(c2, st2, bias=blk["b2"])
            as_, bs = self.coeffs(cs, sts)
            return self.apply(c2, a2, b2, pre_slope=0.2, pre_gain=SQRT2, res=cs, ra=as_, rb=bs, res_mode=1,
                              post_gain=INV_SQRT2)
        return self.apply(c2, res=cs, res_mode=1, post_gain=INV_SQRT2)

    def _res_block_padded(self, blk, xp, norm, last):
        """_res_block without the padding passes (tensor-core path).  xp: the block input inside a buffer with a 1-pixel
        reflected border: conv1 reads it as a valid convolution, the skip branch reads its interior (zero padding by TMA
        out-of-bounds fill).  Without norm, conv1 stores straight into the (2,1,2,1)-padded buffer conv2 reads, and only
        its border is reflected afterwards.  Returns the block output, inside a 1-pixel reflected border unless last."""
        if norm:
            c1, st1 = self.conv(xp, blk["conv1"], want_stats=True)
            a, b = self.coeffs(c1, st1, bias=blk["b1"])
            h1p = self.apply(c1, a, b, pre_slope=0.2, pre_gain=SQRT2, pad=(2, 1, 2, 1), pad_mode=1)
        else:
            h1p, _ = self.conv(xp, blk["conv1"], out_pad=2, out_pad_trim=1)
            ops.pad_border(h1p, (2, 1, 2, 1), 1)
        c2, st2 = self.conv(h1p, blk["conv2"], want_stats=norm)
        cs, sts = self.conv(xp, blk["skip"], want_stats=norm, in_border=1)
        ob = 0 if last else 1
        if norm:
            a2, b2 = self.coeffs(c2, st2, bias=blk["b2"])
            as_, bs = self.coeffs(cs, sts)
            y = self.apply(c2, a2, b2, pre_slope=0.2, pre_gain=SQRT2, res=cs, ra=as_, rb=bs, res_mode=1,
                           post_gain=INV_SQRT2, out_border=ob)
        else:
            y = self.apply(c2, res=cs, res_mode=1, post_gain=INV_SQRT2, out_border=ob)
        if ob:
            ops.pad_border(y, (1, 1, 1, 1), 1)
        return y

    def _padded_ok(self, image_nchw):
        return self.padded and self.use_tc and self.adt == torch.bfloat16 and min(image_nchw.shape[2:]) >= 32

    def encoder_con(self, image_nchw, taps=None):
        """E1: StyleGAN2ResnetEncodercon.forward (encoder_con.py:82-92).  NCHW f32 image -> sp NHWC [n,h/8,w/8,256]."""
        P = self.E1
        if taps is None and self._padded_ok(image_nchw):
            x = self.from_rgb(P, image_nchw, out_border=1)
            for i in (1, 2, 4):
                x = self._res_block_padded(P["res%d" % i], x, True, last=(i == 4))
        else:
            x = self.from_rgb(P, image_nchw)
            if taps is not None:
                taps["FromRGB"] = x
            for i in (1, 2, 4):
                x = self._res_block(P["res%d" % i], x, True)
                if taps is not None:
                    taps["ResBlockDownBy%d" % i] = x
        c, st = self.conv(x, P["tsc0"], want_stats=True)
        a, b = self.coeffs(c, st, bias=P["tsc0_actb"])
        h = self.apply(c, a, b, pre_slope=0.2, pre_gain=SQRT2)
        c, st = self.conv(h, P["tsc1"], want_stats=True)
        a, b = self.coeffs(c, st)
        return self.apply(c, a, b)

    def _code_head(self, lvl, gap_sum, gmax, hw):
        """GAP || GMP -> conv1x1 -> projector MLP -> F.normalize (encoder_col.py:160-166)."""
        v = torch.cat([gap_sum * (1.0 / hw), gmax], dim=1).contiguous()
        w, b = self.E2["c1x1_%s" % lvl]
        v = ops.linear(v, w, b)
        for (pw, pb) in self.E2["proj_%s" % lvl]:
            v = ops.linear(v, pw, pb, relu_in=True)
        return ops.l2_normalize(v, 1e-12, 1)

    def _pooled_head(self, lvl, x, labels=None, label_mode=0, side=None, border=0):
        n, h, w, c = x.shape
        h, w = h - 2 * border, w - 2 * border
        if labels is None:
            r = ops.plane_reduce(x, True, want=("sum", "max"), border=border)
            if side is not None:
                self._keep.append(r)
                with self.fork(side):
                    return self._code_head(lvl, r["sum"].view(n, c), r["max"].view(n, c), h * w)
            return self._code_head(lvl, r["sum"].view(n, c), r["max"].view(n, c), h * w)
        r = ops.plane_reduce(x, True, labels=labels, n_regions=3, label_mode=label_mode, want=("sum", "max"),
                             border=border)
        return [self._code_head(lvl, r["sum"][:, i].contiguous(), r["max"][:, i].contiguous(), h * w) for i in range(3)]

    def _e2_warp(self, x, corr, resize, scale, border=0):
        """StyleGAN2ResnetEncodercol.warp (encoder_col.py:100-138): pool to 64x64, corr @ feat, bilinear up."""
        n, h, w, c = x.shape
        h, w = h - 2 * border, w - 2 * border
        if resize:
            size = (int(64 * h / w), 64) if h > w else (64, int(64 * w / h))
            pooled = ops.resample(x, size, 1, out_dtype=torch.float32, in_border=border)
        else:
            assert border == 0
            size = (h, w)
            pooled = x if x.dtype == torch.float32 else self.apply(x, out_dtype=torch.float32)
        if corr.dtype == torch.bfloat16:
            # corr as [hi | lo] bf16 halves: fp32-accurate product on the tensor cores (3-term split GEMM)
            bsp = ops.split_bf16(pooled.view(n, size[0] * size[1], c), transpose=True)       # [n, c, 2L]
            wf = ops.gemm_split_tc(corr, bsp, (corr.shape[1] // 64, 64))
        else:
            wf = ops.gemm_f32(corr, pooled.view(n, size[0] * size[1], c), False)
        wf = wf.view(n, -1, 64, c) if h > w else wf.view(n, 64, -1, c)
        if resize and scale != 1:
            wf = ops.resample(wf, (wf.shape[1] * scale, wf.shape[2] * scale), 0)
        return wf

    def encoder_col_warp(self, levels, corr):
        """The correlation branch of E2 alone (encoder_col.py:168-178, 100-138) on trunk features kept from an earlier
        encoder_col(..., levels_out=[]) pass over the same images: the warped codes without re-running the trunk.
        Tensor-core mode: the four levels' pooled features sit side by side in ONE B operand, so `corr` (the 537 MB
        operand at batch 8) is read by one split GEMM instead of four, and the bilinear x8/x4/x2 maps that the
        reference interpolates and then only pools are reduced on the fly (ops.upsample_reduce)."""
        borders = [lv[1] if isinstance(lv, tuple) else 0 for lv in levels]     # levels: tensors or (padded tensor, border)
        levels = [lv[0] if isinstance(lv, tuple) else lv for lv in levels]
        n = levels[0].shape[0]
        sizes = []
        for li, x in enumerate(levels):
            h, w = x.shape[1] - 2 * borders[li], x.shape[2] - 2 * borders[li]
            sizes.append(((int(64 * h / w), 64) if h > w else (64, int(64 * w / h))) if li < 3 else (h, w))
        chans = [x.shape[3] for x in levels]
        fused = corr.dtype == torch.bfloat16 and len(set(sizes)) == 1 and sum(chans) <= 512 and \
            all(c % 8 == 0 for c in chans) and \
            all(x.shape[1] - 2 * borders[li] == sizes[0][0] * 2 ** (3 - li) for li, x in enumerate(levels))
        if not fused:
            return [self._pooled_head(lvl, self._e2_warp(levels[li], corr, resize=(li < 3), scale=2 ** (3 - li),
                                                         border=borders[li]))
                    for li, lvl in enumerate(LEVELS)]
        ph, pw = sizes[0]
        pooled = torch.empty((n, ph, pw, 512), device=self.device, dtype=torch.float32)
        off, offs = 0, []
        for li, x in enumerate(levels):
            if li < 3:
                ops.resample(x, (ph, pw), 1, out=pooled, out_c_off=off, in_border=borders[li])
            else:
                ops.affine_apply(x, x.shape[3], out=pooled, out_c_off=off)
            offs.append(off)
            off += chans[li]
        if off < 512:
            pooled[..., off:].zero_()                      # unused B rows (their output columns are never read)
        bsp = ops.split_bf16(pooled.view(n, ph * pw, 512), transpose=True)                 # [n, 512, 2L]
        wf = ops.gemm_split_tc(corr, bsp, (corr.shape[1] // 64, 64)).view(n, ph, pw, 512)
        out = []
        for li, lvl in enumerate(LEVELS):
            scale = 2 ** (3 - li)
            r = ops.upsample_reduce(wf, offs[li], chans[li], scale)
            self._keep.append(r)
            with self.fork(li):                              # the four levels' vector chains run side by side
                out.append(self._code_head(lvl, r["sum"], r["max"], ph * pw * scale * scale))
        self.join()
        return out

    def encoder_col(self, image_nchw, corr=None, labels=None, taps=None, levels_out=None, defer_join=False):
        """E2: StyleGAN2ResnetEncodercol.forward (encoder_col.py:150-251).
        defer_join: the code vectors are still being computed on side stream 0 when this returns; the caller must call
        engine.join() (engine.generator does) before anything on the current stream reads them.
        corr: f32 [n,4096,4096], or its [hi | lo] bf16 split [n,4096,8192] (ops.split_bf16 / corrm(split_out=True)),
        or None.  labels: uint8 [n,h,w] region map or None (training-path mask branch;
        the reference takes the one-hot float mask -- see SwapModel.E2_forward for that signature).
        Returns (vectors, vectors_w) or (vectors, projections_m, vectors_w, projections_mw)."""
        P = self.E2
        vectors, vectors_w, pm, pmw = [], [], [], []
        padded = taps is None and self._padded_ok(image_nchw)
        x = self.from_rgb(P, image_nchw, out_border=1 if padded else 0)
        bits = bits_sw = None
        for li, lvl in enumerate(LEVELS):
            xb = 1 if (padded and li < 3) else 0          # levels that feed another ResBlock sit in a reflected border
            if li > 0:
                blk = P["res%d" % (2 ** (li - 1))]
                x = self._res_block_padded(blk, x, False, last=(li == 3)) if padded else self._res_block(blk, x, False)
            if taps is not None:
                taps["level" + lvl] = x
            if levels_out is not None:
                levels_out.append((x, xb) if padded else x)
            vectors.append(self._pooled_head(lvl, x, side=0, border=xb))
            xx = None
            if corr is not None:
                xx = self._e2_warp(x, corr, resize=(li < 3), scale=2 ** (3 - li), border=xb)
                vectors_w.append(self._pooled_head(lvl, xx))
            if labels is not None:
                bits = ops.labels_pool_or(labels, 1, 0) if li == 0 else ops.labels_pool_or(bits, 2, 1)
                pm.extend(self._pooled_head(lvl, x, labels=bits, label_mode=1, border=xb))
                if corr is not None:
                    n = bits.shape[0]
                    perm = torch.arange(n, device=bits.device).view(n // 2, 2).flip(1).reshape(-1)
                    bits_sw = bits[perm].contiguous()          # encoder_col.py:140-148 (swap adjacent pairs)
                    pmw.extend(self._pooled_head(lvl, xx, labels=bits_sw, label_mode=1))
        if not defer_join:
            self.join()
        if labels is not None:
            return vectors, pm, vectors_w, pmw
        return vectors, vectors_w

    # ------------------------------------------------------------------------------------------ generator
    def styles(self, codes):
        """All StyleMod / SpatialCodeModulation vectors of one generator pass: util.normalize of the four codes
        (generator.py:247) and one batched linear per code."""
        codes = [ops.l2_normalize(c.to(self.device, torch.float32), 1e-8, 0) for c in codes]   # util.normalize
        out = {}
        for idx, w, b, slices in self.G["style_groups"]:
            y = ops.linear(codes[idx], w, b, w_scale=w.shape[1] ** -0.5)
            for key, o, size in slices:
                out[key] = y[:, o:o + size]
        return out

    def styles_async(self, codes, side=0):
        """styles() on a side stream (after whatever already runs there, e.g. the code heads that produce `codes`):
        pass the result as generator(..., styles=...); generator joins before using it."""
        self._keep.append(codes)
        with self.fork(side):
            return self.styles(codes)

    def _styled(self, P, x, s, noise, upsample=False, out_pad=0, lazy=None, in_border=0):
        """StyledConv.forward up to (and excluding) the application of IN+StyleMod: returns (y, a, b).
        s: the layer's style vector [n, 2c] (a row-strided view into the batched linear's output).
        out_pad: y is the interior of a buffer with a 1-pixel border, for a following _styled(..., lazy=(a, b)).
        lazy = (a_in, b_in): x is such a padded, NOT yet normalised buffer: its IN + StyleMod is folded into per-sample
        weights and a per-sample bias, its border set to -b/a (= zero padding of the normalised tensor)."""
        if lazy is not None:
            a_in, b_in = lazy
            spec = P["plain"]
            ops.fill_border(x, a_in, b_in, 1)
            wn, bn = ops.modulate_weights(spec.w32, a_in, b_in, spec.kh * spec.kw, spec.cin, base_bias=spec.bias)
            y, st = self.conv(x, spec, noise=noise, want_stats=True, w_per_image=wn, bias_per_image=bn, valid=True,
                              out_pad=out_pad)
        elif upsample and (min(x.shape[1:3]) - 2 * in_border) * 2 >= 128:
            y, st = self.conv(x, P["fused_up"], noise=noise, want_stats=True, out_pad=out_pad, in_border=in_border)
        else:
            assert in_border == 0
            if upsample:
                x = self.apply(x, up=2)                              # Upscale2d nearest (stylegan2_layers.py:88-105)
            y, st = self.conv(x, P["plain"], noise=noise, want_stats=True, out_pad=out_pad)
        a, b = self.coeffs(y, st, style=s, c=y.shape[3])
        return y, a, b

    def _feature_head(self, P, x, border=0, stats=None):
        """layer32/64/128/256 (generator.py:180-231): [ReplPad] -> IN (on the padded tensor) -> conv -> IN -> PReLU
        -> [ReplPad] -> conv -> IN -> PReLU.  border = 1: x already sits inside its replicated 1-pixel border."""
        pad = (1, 1, 1, 1) if P["k"] == 3 else (0, 0, 0, 0)
        assert border == (1 if P["k"] == 3 else 0) or border == 0
        xp = x if border else (self.apply(x, pad=pad, pad_mode=2) if P["k"] == 3 else x)
        a, b = self.coeffs(xp, stats)
        if self.use_tc and xp.dtype == torch.bfloat16:
            # IN -> conv -> IN: the first normalisation's scale goes into per-sample weights, its shift is a constant
            # per (sample, output channel) that the second normalisation removes -- no pass over the activation
            spec = P["c1"]
            wn = ops.scale_weights(spec.w32, a, spec.kh * spec.kw, spec.cin)
            c1, st = self.conv(xp, spec, want_stats=True, w_per_image=wn)
        else:
            xn = self.apply(xp, a, b)
            c1, st = self.conv(xn, P["c1"], want_stats=True)
        a, b = self.coeffs(c1, st)
        h = self.apply(c1, a, b, pre_slope=P["a1"], pad=pad, pad_mode=2)
        c2, st = self.conv(h, P["c2"], want_stats=True)
        a, b = self.coeffs(c2, st)
        return c2, a, b, P["a2"]

    def _residual_block(self, P, x, border=0, out_border=0):
        """ResidualBlock (generator.py:10-32).  border / out_border = 1: the input / output sits inside a replicated
        1-pixel border (the ReplicationPad2d of this / the next block's conv1), so no padding pass runs for it."""
        xp = x if border else self.apply(x, pad=(1, 1, 1, 1), pad_mode=2)
        c1, st = self.conv(xp, P["c1"], want_stats=True)
        a, b = self.coeffs(c1, st)
        h = self.apply(c1, a, b, pre_slope=P["a"], pad=(1, 1, 1, 1), pad_mode=2)
        c2, st = self.conv(h, P["c2"], want_stats=True)
        a, b = self.coeffs(c2, st)
        y = self.apply(c2, a, b, res=x, res_mode=1, post_slope=P["a"], res_border=border, out_border=out_border)
        if out_border:
            ops.pad_border(y, (out_border,) * 4, 2)
        return y

    def generator(self, sp, codes, extract_features=False, noises=None, taps=None, nchw_out=False, styles=None,
                  need_rgb=True):
        """StyleGAN2ResnetGenerator.forward (generator.py:244-281).  sp NHWC [n,h,w,256] (engine dtype),
        codes: 4 x f32 [n,2048].  Returns rgb NHWC f32 [n,8h,8w,3] (nchw_out: NCHW f32 [n,3,8h,8w], ToRGB's affine
        folded into the layout conversion) (and fea NHWC [n,64,64,256], fea1 NHWC [n,256,256,64] when
        extract_features)."""
        G = self.G
        noises = noises or {}          # device f32 [n,1,H,W] tensors (SwapModel.set_noises moves them once)
        self.join()                    # code vectors / styles that are still in flight on a side stream
        st = styles if styles is not None else self.styles(codes)
        need_rgb = need_rgb or not extract_features
        x = self.apply(sp, st["scm_scale"].contiguous(), st["scm_bias"].contiguous())   # GeneratorModulation, generator.py:80-91
        if taps is not None:
            taps["SpatialCodeModulation"] = x
        tc_fast = self.use_tc and taps is None and x.dtype == torch.bfloat16 and min(x.shape[1:3]) * 2 >= 128
        # extract_features on the tensor-core path: the inputs of the 3x3 feature heads (generator.py:180-231 start with
        # ReplicationPad2d(1)) are produced INSIDE their replicated border, so no padding pass ever copies them
        pad_heads = extract_features and tc_fast and self.padded
        xb = 0                                              # border of x
        for i, (ci, co) in enumerate(HEAD_CH):
            B = G["head%d" % i]
            name = "HeadResnetBlock%d" % i
            skip = self.conv(x, B["skip"])[0] if "skip" in B else x
            y, a, b = self._styled(B["conv1"], x, st["head%d.conv1" % i], noises.get(name + ".conv1"))
            h = self.apply(y, a, b)
            y, a, b = self._styled(B["conv2"], h, st["head%d.conv2" % i], noises.get(name + ".conv2"))
            xb = 1 if (pad_heads and i == len(HEAD_CH) - 1) else 0
            xst = None                                      # statistics of (the padded) x, when its producer emitted them
            if xb:
                x, xst = ops.affine_apply_stats(y, y.shape[3], a, b, res=skip, res_mode=1, post_gain=INV_SQRT2,
                                                pad=(1, 1, 1, 1), pad_mode=2)
            else:
                x = self.apply(y, a, b, res=skip, res_mode=1, post_gain=INV_SQRT2)
            if taps is not None:
                taps[name] = x
        feas = []
        if extract_features:
            pool_hw = (x.shape[1] - 2 * xb, x.shape[2] - 2 * xb)
            feas.append(self._feature_head(G["layer32"], x, border=xb, stats=xst))
        for j, (res, ci, co) in enumerate(UP_CH):
            B = G["up%d" % res]
            name = "UpsamplingResBlock%d" % res
            if "skip" in B:
                skip, sb = self.conv(x, B["skip"], in_border=xb)[0], 0
            else:
                skip, sb = x, xb
            if self.use_tc and taps is None and x.dtype == torch.bfloat16 and (min(x.shape[1:3]) - 2 * xb) * 2 >= 128:
                # conv1's normalisation + style is never applied to the activation: it rides in conv2's weights
                y, a, b = self._styled(B["conv1"], x, st["up%d.conv1" % res], noises.get(name + ".conv1"), upsample=True,
                                       out_pad=1, in_border=xb)
                y, a, b = self._styled(B["conv2"], y, st["up%d.conv2" % res], noises.get(name + ".conv2"), lazy=(a, b))
            else:
                y, a, b = self._styled(B["conv1"], x, st["up%d.conv1" % res], noises.get(name + ".conv1"), upsample=True)
                h = self.apply(y, a, b)
                y, a, b = self._styled(B["conv2"], h, st["up%d.conv2" % res], noises.get(name + ".conv2"))
            yrgb = None
            last = j == len(UP_CH) - 1
            xb = 1 if (pad_heads and not last) else 0       # layer64 / layer128 are 3x3 heads, layer256 is 1x1
            if (last and need_rgb and self.use_tc and taps is None and y.dtype == torch.bfloat16 and y.shape[3] == 128
                    and y.shape[2] % 64 == 0 and y.shape[1] % 2 == 0):
                # last block: ToRGB's 1x1 conv rides on the pass that writes x; plain decode does not write x at all
                x, yrgb = ops.affine_apply_torgb(y, a, b, skip, G["rgb"].w32[0], G["rgb"].bias, post_gain=INV_SQRT2,
                                                 write_out=extract_features)
                xst = None
            elif pad_heads:
                # the pass that forms the block output also replicates its border (layer64 / layer128 start with
                # ReplicationPad2d) and emits the statistics of the padded tensor for the head's first InstanceNorm
                x, xst = ops.affine_apply_stats(y, y.shape[3], a, b, res=skip, res_mode=2, post_gain=INV_SQRT2,
                                                pad=(xb,) * 4, pad_mode=2, res_border=sb)
            else:
                x = self.apply(y, a, b, res=skip, res_mode=2, post_gain=INV_SQRT2, res_border=sb)   # bilinear x2 skip, generator.py:75
                xst = None
            if taps is not None:
                taps[name] = x
            if extract_features:
                feas.append(self._feature_head(G["layer%d" % (2 ** (j + 6))], x, border=xb, stats=xst))
        # ToRGB (stylegan2_layers.py:477-495); fp32 from here on (3-channel IN is sensitive).  need_rgb=False: the caller
        # only wants the feature maps (PPSTModel.extract_feat_from_image discards the image, ppst_model.py:311-315)
        rgb = None
        n = sp.shape[0]
        if need_rgb:
            if yrgb is None:
                yrgb = ops.conv1x1_thin_out(x, G["rgb"].w32[0], G["rgb"].bias)
            a, b = self.coeffs(yrgb, style=st["rgb"])
            rgb = ops.nhwc_to_nchw(yrgb, a=a, b=b) if nchw_out else self.apply(yrgb, a, b)
        if not extract_features:
            return rgb
        # generator.py:272-280: concat pooled / bilinear-resampled heads, then layert / layert1.  Both concatenations
        # are built inside a replicated 1-pixel border (ResidualBlock starts with ReplicationPad2d(1), generator.py:12)
        fb = 1 if (taps is None and self.padded) else 0
        feat = torch.empty((n, pool_hw[0] + 2 * fb, pool_hw[1] + 2 * fb, 256), device=x.device, dtype=self.adt)
        feat1 = torch.empty((n, 256 + 2 * fb, 256 + 2 * fb, 256), device=x.device, dtype=self.adt)
        for i, (c2, a, b, slope) in enumerate(feas):
            # IN + PReLU of the head output applied on load: the normalised map itself is never materialised
            ops.resample(c2, pool_hw, 1, out=feat, out_c_off=64 * i, a=a, b=b, slope=slope, out_border=fb)   # nn.AdaptiveAvgPool2d((h, w))
            ops.resample(c2, (256, 256), 0, out=feat1, out_c_off=64 * i, a=a, b=b, slope=slope, out_border=fb)
        if fb:
            ops.pad_border(feat, (1, 1, 1, 1), 2)
            ops.pad_border(feat1, (1, 1, 1, 1), 2)
        if taps is not None:
            taps["feat_cat"], taps["feat1_cat"] = feat, feat1
        for i in range(3):
            feat = self._residual_block(G["layert.%d" % i], feat, border=fb, out_border=fb if i < 2 else 0)
        feat1 = self._residual_block(G["layert1.0"], feat1, border=fb)
        feat1, _ = self.conv(feat1, G["layert1.1"])
        return rgb, feat, feat1

    # ------------------------------------------------------------------------------------------ correlation
    def rselfcorr_into(self, fea1, out, c_off):
        """PPSTModel.Rselfcorr (ppst_model.py:330-339) written into a channel slice of `out`."""
        return ops.rselfcorr(fea1, out=out, out_c_off=c_off)

    def corrm(self, fea_rows, fea0_rows, split_out=False):
        """PPSTModel.corrm (ppst_model.py:341-364).  fea_rows / fea0_rows: NHWC [n,64,64,512] (any float dtype).
        corr[b, i, j] = softmax_j( <fea0[b,:,i], fea[b,:,j]> / 0.01 )  -> f32 [n, L0, L]
        (split_out: the same numbers as [hi | lo] bf16 halves [n, L0, 2L], the A operand of the warp GEMMs).
        bf16 mode: 3-term split GEMM on tcgen05 (fp32-accurate logits: T = 0.01 turns a 1e-3 cosine error into
        0.1 in the exponent); fp32 mode: exact CUDA-core GEMM."""
        n, c = fea_rows.shape[0], fea_rows.shape[3]
        if self.use_tc:
            a = ops.corr_prepare_split(fea_rows, 256).view(n, -1, 2 * c)
            q = ops.corr_prepare_split(fea0_rows, 256).view(n, -1, 2 * c)
            logits = ops.gemm_split_tc(q, a, (q.shape[1] // 64, 64))
        else:
            a = ops.corr_prepare(fea_rows, 256).view(n, -1, c)
            q = ops.corr_prepare(fea0_rows, 256).view(n, -1, c)
            logits = ops.gemm_f32(q, a, True)
        if split_out:
            return ops.softmax_rows_split(logits, 0.01)
        return ops.softmax_rows_(logits, 0.01)
