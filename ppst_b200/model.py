"""PPSTModel-compatible front end (models/ppst_model.py:264-268,288-315,330-387; models/base_model.py:114-123).

`SwapModel` keeps the reference's inference entry points -- encode / encode2 / decode / extract_feat /
extract_feat_from_image / Rselfcorr / corrm / warp -- with the reference's tensor conventions (NCHW float32 in
and out) and its `model(*args, command="...")` dispatch, so the evaluators' call sequences run unchanged on
top of it.  `swap()` is the same sequence kept on the device in NHWC end to end (the benchmarked path).
"""
import functools

import torch

from . import ops
from .engine import Engine


def _on_own_device(cls):
    """Every public method runs with the model's device current: launches go to the current device's stream with raw
    pointers, so SwapModel(device="cuda:1") must not depend on the caller's torch.cuda.set_device()."""
    def wrap(fn):
        @functools.wraps(fn)
        def inner(self, *a, **kw):
            if torch.cuda.current_device() == self.device.index:
                return fn(self, *a, **kw)
            with torch.cuda.device(self.device):
                return fn(self, *a, **kw)
        return inner
    for name, fn in list(vars(cls).items()):
        if callable(fn) and not name.startswith("_") and not isinstance(fn, (staticmethod, classmethod, property)):
            setattr(cls, name, wrap(fn))
    return cls


def split_state_dict(sd):
    """PPSTModel.state_dict() keys are prefixed E1. / E2. / G.."""
    out = {"E1": {}, "E2": {}, "G": {}}
    for k, v in sd.items():
        head, _, rest = k.partition(".")
        if head in out:
            out[head][rest] = v
    return out["E1"], out["E2"], out["G"]


@_on_own_device
class SwapModel:
    def __init__(self, sdE1, sdE2, sdG, device="cuda", mode="bf16", noise_seed=None):
        if not torch.cuda.is_available():
            raise RuntimeError("ppst_b200.SwapModel needs a CUDA device (there is no CPU fallback)")
        dev = torch.device(device)
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        with torch.cuda.device(dev):
            self.engine = Engine(sdE1, sdE2, sdG, device=dev, mode=mode)
        self.device = self.engine.device
        self._noises = None
        self._noise_gen = None
        self._noise_rep = {}
        self._graphs = {}
        self._pipes = {}
        if noise_seed is not None:
            self.noises = "random"
            self._noise_gen.manual_seed(int(noise_seed))

    @property
    def noises(self):
        """NoiseInjection inputs (stylegan2_layers.py:376-399).
          None      -- the noise term is OFF (exact for random-init / noise.weight == 0 networks; NOT a sample of the
                       reference's output distribution for a trained checkpoint);
          dict      -- '<block>.conv{1,2}' -> [B,1,H,W] explicit maps (parity tests, reproducible outputs);
          "random"  -- what the reference does (its evaluators leave fix_noise commented out): fresh N(0,1) maps on every
                       generator pass, drawn on the device from a generator seeded by `noise_seed` (SwapModel(...,
                       noise_seed=s)).  Eager API only: swap_graph / swap_pipelined replay a captured step and refuse it."""
        return self._noises

    @noises.setter
    def noises(self, value):
        if isinstance(value, str):
            if value != "random":
                raise ValueError("noises: None, a dict of maps, or 'random'")
            self._noises = "random"
            if self._noise_gen is None:
                self._noise_gen = torch.Generator(device=self.device)
        else:
            self._noises = None if value is None else {
                k: v.to(self.device, torch.float32).contiguous() for k, v in value.items()}
        self._graphs.clear()
        self._pipes.clear()
        self._noise_rep = {}

    def _noises_for(self, rows, sp_hw=None):
        """The noise dict for a pass over `rows` samples: the user's [B,...] maps repeated when the content and style
        rows are batched together (rows = 2B); fresh draws in "random" mode (sp_hw = spatial-code size, default 64x64)."""
        if self._noises is None:
            return None
        if isinstance(self._noises, str):
            sh, sw = sp_hw or (64, 64)
            out = {}
            for i in range(4):
                for c in ("conv1", "conv2"):
                    out["HeadResnetBlock%d.%s" % (i, c)] = torch.randn(rows, 1, sh, sw, device=self.device,
                                                                        generator=self._noise_gen)
            for j, res in enumerate((16, 32, 64)):
                f = 2 ** (j + 1)
                for c in ("conv1", "conv2"):
                    out["UpsamplingResBlock%d.%s" % (res, c)] = torch.randn(rows, 1, sh * f, sw * f, device=self.device,
                                                                             generator=self._noise_gen)
            return out
        hit = self._noise_rep.get(rows)
        if hit is None:                                  # repeated once per batch size, not once per pass
            hit = self._noise_rep[rows] = {k: (v if v.shape[0] == rows else v.repeat(rows // v.shape[0], 1, 1, 1))
                                           for k, v in self._noises.items()}
        return hit

    @classmethod
    def from_model_state_dict(cls, sd, **kw):
        return cls(*split_state_dict(sd), **kw)

    # -- BaseModel.forward(command=...) ---------------------------------------------------------------
    def __call__(self, *args, command=None, **kwargs):
        if command is None:
            raise ValueError(command)
        method = getattr(self, command)
        assert callable(method), "[%s] is not a method of %s" % (command, type(self).__name__)
        return method(*args, **kwargs)

    # -- helpers ------------------------------------------------------------------------------------------
    def _in(self, image):
        return image.to(self.device, torch.float32).contiguous()

    @staticmethod
    def _out(x_nhwc):
        return ops.nhwc_to_nchw(x_nhwc, dtype=torch.float32)

    # -- PPSTModel API ------------------------------------------------------------------------------------
    def encode(self, image, extract_features=False, testtime=False):
        """ppst_model.py:264-265 -> (sp [B,256,H/8,W/8], [4 x [B,2048]])."""
        x = self._in(image)
        sp = self.engine.encoder_con(x)
        gl, _ = self.engine.encoder_col(x)
        return self._out(sp), gl

    def encode2(self, image, corrmatrix):
        """ppst_model.py:267-268.  Accepts the tensor or the 1-element list that
        evaluation/simple_swapping_evaluator.py:53 passes."""
        if isinstance(corrmatrix, (list, tuple)):
            corrmatrix = corrmatrix[0]
        corr = self._corr_operand(corrmatrix)
        return self.engine.encoder_col(self._in(image), corr=corr)

    def _corr_operand(self, corrmatrix):
        corr = corrmatrix.to(self.device, torch.float32).contiguous()
        if self.engine.use_tc:
            corr = ops.split_bf16(corr)          # [hi | lo] bf16 halves for the split tensor-core GEMM
        return corr

    def E2_forward(self, image, mask=None, corrmatrix=None):
        """StyleGAN2ResnetEncodercol.forward signature (encoder_col.py:150): `mask` is the one-hot float mask
        [B,3,H,W] of the training path; it is converted to the integer label map the kernels key on."""
        labels = None
        if mask is not None:
            labels = mask.to(self.device).argmax(dim=1).to(torch.uint8).contiguous()
        corr = None
        if corrmatrix is not None:
            if isinstance(corrmatrix, (list, tuple)):
                corrmatrix = corrmatrix[0]
            corr = self._corr_operand(corrmatrix)
        return self.engine.encoder_col(self._in(image), corr=corr, labels=labels)

    def decode(self, spatial_code, global_code, target=None):
        """ppst_model.py:288-306.  With `target` (the content image, as the grid evaluator passes it) the decoded image
        goes through the guided-filter smoothing of photo_gif.GIFSmoothing(r=30, eps=(0.02*255)**2) -- on the device,
        without the reference's host round trip."""
        sp = self.engine.to_nhwc(spatial_code)
        rgb = self.engine.generator(sp, global_code, noises=self._noises_for(sp.shape[0], sp.shape[1:3]), nchw_out=True)
        if target is not None:
            rgb = ops.guided_smooth(rgb, self._in(target))
        return rgb

    def extract_feat(self, spatial_code, global_code):
        """ppst_model.py:308-309 -> (rgb, fea, fea1)."""
        sp = self.engine.to_nhwc(spatial_code)
        rgb, fea, fea1 = self.engine.generator(sp, global_code, extract_features=True,
                                               noises=self._noises_for(sp.shape[0], sp.shape[1:3]), nchw_out=True)
        return rgb, self._out(fea), self._out(fea1)

    def extract_feat_from_image(self, img):
        """ppst_model.py:311-315 -> (fea [B,256,64,64], fea1 [B,64,256,256])."""
        x = self._in(img)
        sp = self.engine.encoder_con(x)
        gl, _ = self.engine.encoder_col(x)
        _, fea, fea1 = self.engine.generator(sp, gl, extract_features=True,
                                             noises=self._noises_for(sp.shape[0], sp.shape[1:3]), need_rgb=False)
        return self._out(fea), self._out(fea1)

    def Rselfcorr(self, fea):
        """ppst_model.py:330-339: [B,64,256,256] -> [B,256,64,64]."""
        x = ops.nchw_to_nhwc(fea.to(self.device, torch.float32))
        return self._out(ops.rselfcorr(x))

    def corrm(self, fea, fea0):
        """ppst_model.py:341-364: ([B,512,64,64], [B,512,64,64]) -> [B,4096,4096]."""
        a = ops.nchw_to_nhwc(fea.to(self.device, torch.float32))
        q = ops.nchw_to_nhwc(fea0.to(self.device, torch.float32))
        return self.engine.corrm(a, q)

    def warp(self, fea, corr):
        """ppst_model.py:366-387 (both the full-resolution and the s x s unfold branch)."""
        fea = fea.to(self.device, torch.float32)
        corr = corr.to(self.device, torch.float32).contiguous()
        b, c, h, w = fea.shape
        H = corr.shape[1]
        x = ops.nchw_to_nhwc(fea)                                        # [b,h,w,c]
        if H != h * w:
            s = int(((h * w) / H) ** 0.5)
            # unfold(s, stride=s): rows = (h/s * w/s) patches, features ordered (c, dy, dx)
            pt = x.view(b, h // s, s, w // s, s, c).permute(0, 1, 3, 5, 2, 4).reshape(b, H, c * s * s).contiguous()
            wf = ops.gemm_f32(corr, pt, False)
            wf = wf.view(b, h // s, w // s, c, s, s).permute(0, 3, 1, 4, 2, 5).reshape(b, c, h, w)
            return wf.contiguous()
        wf = ops.gemm_f32(corr, x.view(b, h * w, c), False).view(b, h, w, c)
        return self._out(wf)

    # -- the fused device-resident swap -------------------------------------------------------------------
    def swap(self, content, style, alpha=1.0, return_nhwc=False):
        """One (content, style) -> swapped image per batch row: the minimal unit of SURVEY.md 8(d) C2
        {E1(c), E2(c), G.extract(c), E1(s), E2(s), G.extract(s), Rselfcorr x2, corrm, E2(s,corr), G.decode},
        i.e. evaluation/simple_swapping_evaluator.py:44-59 without the duplicated content encode.  Eager launches."""
        e = self.engine
        xc, xs = self._in(content), self._in(style)
        n = xc.shape[0]
        if xc.shape != xs.shape:
            raise ValueError("swap: content %s and style %s must have the same shape" % (tuple(xc.shape), tuple(xs.shape)))
        # content and style rows go through E1 / E2 / G.extract as ONE batch of 2n (every normalisation on the path is
        # per sample, so each row's result is bit-identical to a separate pass); E2's trunk features of the style rows
        # are kept for the correlation branch, which the reference recomputes from the image (encode2).
        x2 = torch.cat([xc, xs], 0)
        noises2 = self._noises_for(2 * n)
        # E2 first: its code heads and the style linears (chains of tiny dependent kernels) run on a side stream under E1
        levels = []
        gl2, _ = e.encoder_col(x2, levels_out=levels, defer_join=True)
        st2 = e.styles_async(gl2)
        sp2 = e.encoder_con(x2)
        _, fea2, fea21 = e.generator(sp2, gl2, extract_features=True, noises=noises2, styles=st2, need_rgb=False)
        sp_c, gl_c = sp2[:n], [v[:n] for v in gl2]
        fea_c, fea_s, fea_c1, fea_s1 = fea2[:n], fea2[n:], fea21[:n], fea21[n:]
        fc = torch.empty((n, 64, 64, 512), device=self.device, dtype=torch.float32)
        fs = torch.empty_like(fc)
        ops.affine_apply(fea_c, 256, out=fc, out_c_off=0)
        ops.affine_apply(fea_s, 256, out=fs, out_c_off=0)
        ops.rselfcorr(fea_c1, out=fc, out_c_off=256)
        ops.rselfcorr(fea_s1, out=fs, out_c_off=256)
        corr = e.corrm(fs, fc, split_out=e.use_tc)
        gl_w = e.encoder_col_warp([(x[0][n:], x[1]) if isinstance(x, tuple) else x[n:] for x in levels], corr)
        codes = gl_w if alpha == 1.0 else [ops.lerp(a, b, alpha) for a, b in zip(gl_c, gl_w)]
        return e.generator(sp_c, codes, noises=self._noises_for(n), nchw_out=not return_nhwc)

    def plain_swap(self, content, style, return_nhwc=False):
        """G(E1(content), E2(style)) without the correlation branch: the only swap the reference can run away from
        512^2 (its correlation path is hard-wired to 64x64 / 256x256 feature maps) -- the
        1024^2 configuration of BASELINE.json."""
        e = self.engine
        sp = e.encoder_con(self._in(content))
        gl, _ = e.encoder_col(self._in(style))
        return e.generator(sp, gl, noises=self._noises_for(sp.shape[0], sp.shape[1:3]), nchw_out=not return_nhwc)

    def swap_grid(self, contents, styles, cell_batch=8, smooth=False):
        """Every (content i, style j) cell of evaluation/content_style_grid_generation_evaluator.py:57-91 in one call:
        contents [Nc,3,512,512], styles [Ns,3,512,512] -> [Nc, Ns, 3, 512, 512].
        Per-image work (E1, E2 trunk, G.extract, Rselfcorr) runs ONCE per image instead of once per cell (the
        reference re-encodes the style image for every content); per-cell work {corrm, E2 correlation branch,
        G.decode} is batched `cell_batch` cells at a time.  Each cell is bit-identical to swap(content_i, style_j).
        smooth=True applies decode(target=content_i)'s guided-filter post-process (photo_gif.py) on the device."""
        e = self.engine
        xc, xs = self._in(contents), self._in(styles)
        nc, ns = xc.shape[0], xs.shape[0]
        x2 = torch.cat([xc, xs], 0)
        levels = []
        gl2, _ = e.encoder_col(x2, levels_out=levels, defer_join=True)
        st2 = e.styles_async(gl2)
        sp2 = e.encoder_con(x2)
        _, fea2, fea21 = e.generator(sp2, gl2, extract_features=True, noises=self._noises_for(nc + ns), styles=st2,
                                     need_rgb=False)
        f2 = torch.empty((nc + ns, 64, 64, 512), device=self.device, dtype=torch.float32)
        ops.affine_apply(fea2, 256, out=f2, out_c_off=0)
        ops.rselfcorr(fea21, out=f2, out_c_off=256)
        sp_c = sp2[:nc]
        lv_s = [(x[0][nc:], x[1]) if isinstance(x, tuple) else (x[nc:], 0) for x in levels]
        out = torch.empty((nc, ns, 3) + tuple(xc.shape[2:]), device=self.device, dtype=torch.float32)
        cells = [(i, j) for i in range(nc) for j in range(ns)]
        for k0 in range(0, len(cells), cell_batch):
            chunk = cells[k0:k0 + cell_batch]
            ci = torch.tensor([c[0] for c in chunk], device=self.device)
            sj = torch.tensor([c[1] for c in chunk], device=self.device)
            corr = e.corrm(f2.index_select(0, sj + nc), f2.index_select(0, ci), split_out=e.use_tc)
            gl_w = e.encoder_col_warp([(x.index_select(0, sj), b) for x, b in lv_s], corr)
            rgb = e.generator(sp_c.index_select(0, ci), gl_w, noises=self._noises_for(len(chunk)), nchw_out=True)
            if smooth:
                rgb = ops.guided_smooth(rgb, xc.index_select(0, ci))
            out.view(nc * ns, *out.shape[2:])[k0:k0 + len(chunk)] = rgb
        return out

    def swap_graph(self, content, style, alpha=1.0):
        """swap() replayed from a CUDA graph: the ~600 kernel launches of one step are captured once per
        (batch shape, alpha) and replayed with a single host call.  `content` / `style` may be pinned host tensors
        (the copies into the graph's static inputs are part of the call).  The returned tensor is the graph's static
        output buffer: consume or copy it before the next call with the same shapes."""
        key = (tuple(content.shape), tuple(style.shape), float(alpha))
        entry = self._graphs.get(key)
        if entry is None:
            if isinstance(self._noises, str):
                raise RuntimeError("swap_graph / swap_pipelined replay a captured step: use explicit noise maps or noises=None "
                                   "(noises='random' draws fresh maps per call and belongs to the eager API)")
            c_static = torch.empty(tuple(content.shape), device=self.device, dtype=torch.float32)
            s_static = torch.empty(tuple(style.shape), device=self.device, dtype=torch.float32)
            c_static.copy_(content)
            s_static.copy_(style)
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):                       # warm-up off the capture: attribute setup, tensor maps
                for _ in range(2):
                    self.swap(c_static, s_static, alpha)
            torch.cuda.current_stream(self.device).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out_static = self.swap(c_static, s_static, alpha)
            entry = (graph, c_static, s_static, out_static)
            self._graphs[key] = entry
        graph, c_static, s_static, out_static = entry
        c_static.copy_(content, non_blocking=True)
        s_static.copy_(style, non_blocking=True)
        graph.replay()
        return out_static

    def swap_pipelined(self, content, style, out, alpha=1.0):
        """Streaming form of swap_graph for HOST-resident batches (pinned `content`, `style`, `out`): the host-to-device
        copy of this call's inputs runs on a copy stream into one of two staging buffers while the previous call's graph is
        still executing, and the result leaves on a third stream, so PCIe time hides behind the step instead of adding to
        it.  Every call still moves its own inputs and its own result.  Returns the CUDA event that marks `out` complete
        (also make the launching stream wait for it before timing or reusing `out`)."""
        dev = self.device
        key = (tuple(content.shape), tuple(style.shape), float(alpha))
        if key not in self._graphs:
            self.swap_graph(content, style, alpha)                     # capture (and one un-pipelined step)
        graph, c_static, s_static, out_static = self._graphs[key]
        st = self._pipes
        pp = st.get(key)
        if pp is None:
            pp = {"i": 0, "in": torch.cuda.Stream(device=dev), "out": torch.cuda.Stream(device=dev),
                  "c": [torch.empty_like(c_static) for _ in range(2)], "s": [torch.empty_like(s_static) for _ in range(2)],
                  "o": [torch.empty_like(out_static) for _ in range(2)],
                  "h2d": [torch.cuda.Event() for _ in range(2)], "used": [torch.cuda.Event() for _ in range(2)],
                  "ready": [torch.cuda.Event() for _ in range(2)], "d2h": [torch.cuda.Event() for _ in range(2)]}
            st[key] = pp
        i = pp["i"]
        b = i & 1
        pp["i"] = i + 1
        main = torch.cuda.current_stream(dev)
        with torch.cuda.stream(pp["in"]):
            if i >= 2:
                pp["in"].wait_event(pp["used"][b])                     # staging buffer b was consumed by call i-2
            else:
                pp["in"].wait_stream(main)
            pp["c"][b].copy_(content, non_blocking=True)
            pp["s"][b].copy_(style, non_blocking=True)
            pp["h2d"][b].record(pp["in"])
        main.wait_event(pp["h2d"][b])
        c_static.copy_(pp["c"][b], non_blocking=True)
        s_static.copy_(pp["s"][b], non_blocking=True)
        pp["used"][b].record(main)
        graph.replay()
        if i >= 2:
            main.wait_event(pp["d2h"][b])                              # result buffer b has left for the host
        pp["o"][b].copy_(out_static, non_blocking=True)
        pp["ready"][b].record(main)
        with torch.cuda.stream(pp["out"]):
            pp["out"].wait_event(pp["ready"][b])
            out.copy_(pp["o"][b], non_blocking=True)
            pp["d2h"][b].record(pp["out"])
        return pp["d2h"][b]
