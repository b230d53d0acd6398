"""Import shims that let the UNMODIFIED reference (/root/reference) run on a CPU-only host.

TEST INFRASTRUCTURE ONLY.  Used by oracle/gen_golden.py (in the build container, where
/root/reference is mounted) to produce the fixtures under tests/golden/.  Nothing in the
product package imports this file, and it is never used on the GPU box (the reference tree
does not travel there).

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

REF_ROOT = "/root/reference"


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def install():
    """Install the shims and return the imported reference packages."""
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
