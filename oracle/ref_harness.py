"""Import shims that let the UNMODIFIED reference run outside its own launcher (CPU-only host or GPU box).

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  Used by oracle/gen_golden.py (in the build container, where
/root/reference is mounted) to produce the fixtures under tests/golden/, and -- through the verbatim staged copy
oracle/_ref/reference that oracle/stage_ref.py makes -- by bench.py's reference arm, tools/ref_gpu_ops.py and
tests/test_gpu_reference_modules.py on the GPU box.  Nothing in the product package imports this file.

The shims follow SURVEY.md Appendix B:
  * stub modules for dominate / func_timeout / lpips / GPUtil (off the hot path),
  * util.is_custom_kernel_supported -> False so the reference takes its own pure-PyTorch
    twins (upfirdn2d_native, stylegan2_op/upfirdn2d.py:162-222; the F.leaky_relu branch of
    fused_leaky_relu, stylegan2_op/fused_act.py:93-96),
  * .cuda() neutralised (encoder_con.py:30, encoder_col.py:174,227),
  * networks/rscl.py aliased as models.networks.rscl (models/ppst_model.py:11).
"""
import importlib.util
import sys
import types
from argparse import Namespace

import torch

import os  # noqa: E402

REF_ROOT = "/root/reference"
if not os.path.isdir(os.path.join(REF_ROOT, "models")):
    _staged = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "reference")
    if os.path.isdir(os.path.join(_staged, "models")):
        REF_ROOT = _staged


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "models"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def install(custom_kernels=False):
    """Install the shims and return the imported reference packages.
    custom_kernels=False: the reference's own pure-PyTorch twins (the CPU oracle);
    custom_kernels=True (GPU box): util.is_custom_kernel_supported is left alone, so the reference JIT-builds and uses
    its own CUDA ops (stylegan2_op/*.cu) -- the kernels-to-beat of tools/ref_gpu_ops.py."""
    if not available():
        raise ImportError("the reference tree is neither mounted at /root/reference nor staged under oracle/_ref "
                          "(python oracle/stage_ref.py in the build container)")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    if "dominate" not in sys.modules:
        tags = _stub("dominate.tags", **{t: (lambda *a, **k: None) for t in
                                         ("meta", "h3", "table", "tr", "td", "p", "a", "img", "br")})
        _stub("dominate", tags=tags, document=lambda *a, **k: None)
    if "func_timeout" not in sys.modules:
        _stub("func_timeout", func_timeout=lambda t, f, *a, **k: f(*a, **k), FunctionTimedOut=Exception)
    if "lpips" not in sys.modules:
        class _LPIPS(torch.nn.Module):
            def __init__(self, net=None):
                super().__init__()

            def forward(self, a, b):
                return (a - b).abs().mean()
        _stub("lpips", LPIPS=_LPIPS)
    if "GPUtil" not in sys.modules:
        _stub("GPUtil")
    if not torch.cuda.is_available():
        torch.nn.Module.cuda = lambda self, *a, **k: self
        torch.Tensor.cuda = lambda self, *a, **k: self
    import util  # noqa: E402  (the reference's util package)
    if not custom_kernels:
        util.is_custom_kernel_supported = lambda: False
        util.util.is_custom_kernel_supported = lambda: False
    import models.networks as networks  # noqa: E402
    if "models.networks.rscl" not in sys.modules:
        spec = importlib.util.spec_from_file_location("models.networks.rscl", REF_ROOT + "/networks/rscl.py")
        mod = importlib.util.module_from_spec(spec)
        sys.modules["models.networks.rscl"] = mod
        spec.loader.exec_module(mod)
    return networks


def default_opt(**over):
    """CelebAMaskHQ_default options that the three hot-path networks read."""
    opt = Namespace(
        spatial_code_ch=256, global_code_ch=2048, num_classes=0, use_antialias=True,
        netE_scale_capacity=1.0, netE_num_downsampling_sp=3, netE_num_downsampling_gl=2, netE_nc_steepness=2.0,
        netE2_scale_capacity=1.0, netE2_num_downsampling_gl1=3, netE2_num_downsampling_gl2=0,
        netE2_nc_steepness=2.0,
        netG_scale_capacity=1.0, netG_num_base_resnet_layers=4, netG_use_noise=True, netG_resnet_ch=256,
        match_kernel=1, local_rank=1, num_gpus=0, isTrain=False,
    )
    for k, v in over.items():
        setattr(opt, k, v)
    return opt


def build_networks(opt=None):
    networks = install()
    opt = opt or default_opt()
    E1 = networks.create_network(opt, "StyleGAN2Resnet", "encoder_con", verbose=False)
    E2 = networks.create_network(opt, "StyleGAN2Resnet", "encoder_col", verbose=False)
    G = networks.create_network(opt, "StyleGAN2Resnet", "generator", verbose=False)
    return E1.eval(), E2.eval(), G.eval()


def ppst_model_methods():
    """Unbound PPSTModel.Rselfcorr / corrm / warp bound to a bare object with .opt (Appendix B.6)."""
    install()
    from models.ppst_model import PPSTModel

    class _Bare:
        opt = default_opt()
    b = _Bare()
    return (lambda f: PPSTModel.Rselfcorr(b, f),
            lambda f, f0: PPSTModel.corrm(b, f, f0),
            lambda f, c: PPSTModel.warp(b, f, c))


class RefModel:
    """The reference's own inference model without its launcher: the UNMODIFIED E1 / E2 / G modules (create_network) and
    PPSTModel's own methods (models/ppst_model.py:264-387) borrowed unbound onto this object, dispatched like
    BaseModel.forward(command=...) (models/base_model.py:114-123).  Two host-side accommodations, both listed in SURVEY.md
    Appendix C: the 1-element list that simple_swapping_evaluator.py:53 passes to encode2 is unwrapped (the reference
    raises AttributeError there), and tensors are moved to the device the modules live on."""

    def __init__(self, state_dicts=None, device="cpu", custom_kernels=False, noises=None):
        install(custom_kernels=custom_kernels)
        from models.ppst_model import PPSTModel
        self._cls = PPSTModel
        self.opt = default_opt()
        self.E1, self.E2, self.G = build_networks(self.opt)
        if state_dicts is not None:
            for net, sd in zip((self.E1, self.E2, self.G), state_dicts):
                net.load_state_dict(sd)
        self.device = torch.device(device)
        for net in (self.E1, self.E2, self.G):
            net.to(self.device)
        if noises is not None:
            self.set_noises(noises)

    def set_noises(self, noises):
        for name, m in self.G.named_modules():
            if type(m).__name__ == "NoiseInjection":
                v = noises.get(name.replace(".noise", "")) if noises else None
                m.fixed_noise = None if v is None else v.to(self.device)

    def _to(self, a):
        if isinstance(a, torch.Tensor):
            return a.to(self.device)
        if isinstance(a, (list, tuple)):
            return type(a)(self._to(v) for v in a)
        return a

    def __call__(self, *args, command=None, **kwargs):
        if command is None:
            raise ValueError(command)
        args = [self._to(a) for a in args]
        kwargs = {k: self._to(v) for k, v in kwargs.items()}
        if command == "encode2" and isinstance(args[1], (list, tuple)):
            args[1] = args[1][0]
        method = getattr(self._cls, command)
        with torch.no_grad():
            return method(self, *args, **kwargs)

    def swap_unit(self, content, style):
        """The C2 unit of SURVEY.md 8(d): {E1(c), E2(c), G.extract(c), E1(s), E2(s), G.extract(s), Rselfcorr x2, corrm,
        E2(s, corr), G.decode} through the reference's own methods."""
        sp_c, _ = self(content, command="encode")
        fea_c, fea_c1 = self(content, command="extract_feat_from_image")
        fea_s, fea_s1 = self(style, command="extract_feat_from_image")
        fea_c = torch.cat((fea_c, self(fea_c1, command="Rselfcorr")), dim=1)
        fea_s = torch.cat((fea_s, self(fea_s1, command="Rselfcorr")), dim=1)
        corr = self(fea_s, fea_c, command="corrm")
        _, gl_w = self(style, corr, command="encode2")
        return self(sp_c, gl_w, command="decode")
