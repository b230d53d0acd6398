"""The C-ABI library loads on a CPU-only host and exports every symbol include/ppst_b200.h declares
(no compute calls here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "ppst_b200.h")).read()
    return sorted(set(re.findall(r"PPST_API[^;(]*?\b(ppst_\w+)\s*\(", text)))


def test_header_declares_the_plugin_entry_points():
    syms = declared_symbols()
    for must in ("ppst_fused_bias_act", "ppst_upfirdn2d", "ppst_plane_reduce", "ppst_conv_tc", "ppst_affine_apply"):
        assert must in syms
    assert len(syms) >= 25


def test_library_exports_every_declared_symbol():
    from ppst_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(lib, s), "libppst_b200.so does not export %s" % s
    # the ctypes table binds exactly the declared set
    assert sorted(_lib.EXPORTED) == declared_symbols()
    assert _lib.load().ppst_version() >= 100


def test_no_cpu_fallback():
    import torch
    from ppst_b200.stylegan2_op import fused_leaky_relu, upfirdn2d
    with pytest.raises(RuntimeError):
        fused_leaky_relu(torch.zeros(1, 2, 4, 4), torch.zeros(2))
    with pytest.raises(RuntimeError):
        upfirdn2d(torch.zeros(1, 2, 8, 8), torch.ones(3, 3) / 9)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "ppst_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("# oracle", ""), "%s references the oracle" % f
