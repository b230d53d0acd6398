"""ppst_b200: a B200-native (sm_100a) implementation of the swap / inference hot path of wangxb29/PPST.

    import ppst_b200
    ppst_b200.install()                      # alias models.networks.stylegan2_op -> ppst_b200.stylegan2_op
    m = ppst_b200.SwapModel(sdE1, sdE2, sdG) # reference state_dicts, unchanged
    out = m.swap(content, style)

Everything computes in libppst_b200.so (hand-written CUDA behind the C ABI of include/ppst_b200.h).
There is no CPU or PyTorch fallback: a missing library or a CPU tensor raises.
"""
import sys

__version__ = "0.1.0"


def install():
    """Make the reference's modules pick up the B200 operators: register this package's `stylegan2_op` under the
    reference's module path (models/networks/stylegan2_op/__init__.py:1-2 exports the same three names).  Call
    BEFORE `models.networks.stylegan2_layers` is imported; already-imported layer modules are patched too."""
    from . import stylegan2_op
    sys.modules["models.networks.stylegan2_op"] = stylegan2_op
    sys.modules["models.networks.stylegan2_op.fused_act"] = sys.modules[stylegan2_op.__name__ + ".fused_act"]
    sys.modules["models.networks.stylegan2_op.upfirdn2d"] = sys.modules[stylegan2_op.__name__ + ".upfirdn2d"]
    for name in ("models.networks.stylegan2_layers", "models.networks.encoder_con", "models.networks.encoder_col"):
        mod = sys.modules.get(name)
        if mod is not None:
            for attr in ("upfirdn2d", "fused_leaky_relu", "FusedLeakyReLU"):
                if hasattr(mod, attr):
                    setattr(mod, attr, getattr(stylegan2_op, attr))
    return stylegan2_op


def __getattr__(name):
    if name == "SwapModel":
        from .model import SwapModel
        return SwapModel
    if name == "Engine":
        from .engine import Engine
        return Engine
    raise AttributeError(name)
