"""Drop-in for the reference plugin `models.networks.stylegan2_op` (stylegan2_op/__init__.py:1-2):
same three exports, same signatures, B200-native kernels underneath."""
from .fused_act import FusedLeakyReLU, fused_leaky_relu
from .upfirdn2d import upfirdn2d

__all__ = ["FusedLeakyReLU", "fused_leaky_relu", "upfirdn2d"]
