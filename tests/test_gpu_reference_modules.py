"""The north_star's drop-in claim, exercised with the reference's OWN code: after ppst_b200.install() the unmodified
reference modules (ConvLayer / ResBlock / Discriminator, models/networks/stylegan2_layers.py:497-649) and its
SimpleSwappingEvaluator (evaluation/simple_swapping_evaluator.py:38-76) run on top of ppst_b200 -- compared with the same
reference code on its pure-PyTorch CPU fallbacks (use_custom_kernel = False).

Needs the verbatim staged copy of the reference (oracle/stage_ref.py -> oracle/_ref/reference, git-ignored, ships with
the snapshot); skipped with a message when it is absent."""
import math
import os
import sys
from argparse import Namespace

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness as RH  # noqa: E402

needs_ref = pytest.mark.skipif(not RH.available(), reason="reference tree not staged (python oracle/stage_ref.py)")


@pytest.fixture(scope="module")
def ref_layers():
    """(layers module, restore): the reference's stylegan2_layers imported on its CPU fallbacks; `plugin(True)` switches the
    three operator names it imported to ppst_b200's (what ppst_b200.install() does), `plugin(False)` switches back."""
    RH.install(custom_kernels=False)
    import models.networks.stylegan2_layers as L
    saved = {k: getattr(L, k) for k in ("upfirdn2d", "fused_leaky_relu", "FusedLeakyReLU")}
    saved_mods = {k: sys.modules.get(k) for k in ("models.networks.stylegan2_op", "models.networks.stylegan2_op.fused_act",
                                                  "models.networks.stylegan2_op.upfirdn2d")}
    import ppst_b200

    def plugin(on):
        if on:
            op = ppst_b200.install()
            assert L.upfirdn2d is op.upfirdn2d and L.FusedLeakyReLU is op.FusedLeakyReLU
        else:
            for k, v in saved.items():
                setattr(L, k, v)
            for k, v in saved_mods.items():
                if v is not None:
                    sys.modules[k] = v
    yield L, plugin
    plugin(False)


def _both(L, plugin, make, run):
    """run(module, device) on the same reference module twice with identical parameters: on the CPU with the reference's
    own PyTorch fallbacks, then on cuda with ppst_b200 installed (the layer classes look the three operator names up at
    call time, so the switch must stay in place while each module runs)."""
    plugin(False)
    torch.manual_seed(0)
    m_cpu = make(L)
    for p in m_cpu.parameters():                       # non-trivial biases (they are zero-initialised)
        if p.dim() == 1:
            p.data.normal_(0, 0.1)
    out_cpu = run(m_cpu, "cpu")
    plugin(True)
    m_gpu = make(L)
    m_gpu.load_state_dict(m_cpu.state_dict())
    m_gpu.cuda()
    out_gpu = run(m_gpu, "cuda")
    plugin(False)
    return out_cpu, out_gpu


def _rel(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


@needs_ref
@pytest.mark.parametrize("case", ["conv_down_reflect", "conv_down_zero", "conv_act", "resblock_reflect", "resblock_d"])
def test_reference_layers_run_on_the_plugin(ref_layers, case):
    L, plugin = ref_layers
    make = {
        "conv_down_reflect": lambda L: L.ConvLayer(16, 32, 3, downsample=True, blur_kernel=[1, 2, 1], reflection_pad=True),
        "conv_down_zero": lambda L: L.ConvLayer(16, 32, 3, downsample=True),
        "conv_act": lambda L: L.ConvLayer(3, 32, 1),
        "resblock_reflect": lambda L: L.ResBlock(32, 64, [1, 2, 1], reflection_pad=True, norm="in"),      # encoder_con.py:38
        "resblock_d": lambda L: L.ResBlock(32, 64, [1, 3, 3, 1]),
    }[case]
    cin = 3 if case == "conv_act" else (16 if case.startswith("conv") else 32)
    x = torch.randn(2, cin, 40, 48, generator=torch.Generator().manual_seed(1))
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False

    def run(m, dev):
        xin = x.detach().clone().to(dev).requires_grad_(True)
        y = m(xin)
        go = torch.randn(y.shape, generator=torch.Generator().manual_seed(2)).to(dev)
        y.backward(go)
        return y.detach().cpu(), xin.grad.cpu(), {k: p.grad.detach().cpu() for k, p in m.named_parameters()}
    (y0, gx0, gp0), (y1, gx1, gp1) = _both(L, plugin, make, run)
    assert y1.shape == y0.shape
    assert _rel(y1, y0) <= 1e-5
    assert _rel(gx1, gx0) <= 1e-5
    for k in gp0:
        assert _rel(gp1[k], gp0[k]) <= 1e-4, k


@needs_ref
def test_reference_discriminator_with_r1_double_backward_on_the_plugin(ref_layers):
    """Discriminator (stylegan2_layers.py:582-649) + the R1 penalty of models/ppst_model.py:139-156: needs the plugin's
    double backward (upfirdn2d and fused_leaky_relu both appear under autograd.grad(create_graph=True))."""
    L, plugin = ref_layers
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    x = torch.randn(2, 3, 64, 64, generator=torch.Generator().manual_seed(3))

    def run(m, dev):
        xin = x.detach().clone().to(dev).requires_grad_(True)
        pred = m(xin)
        (g,) = torch.autograd.grad(pred.sum(), [xin], create_graph=True, retain_graph=True)
        r1 = g.pow(2).reshape(2, -1).sum(1).mean()
        r1.backward()
        return pred.detach().cpu(), float(r1.detach()), {k: p.grad.detach().cpu() for k, p in m.named_parameters() if p.grad is not None}
    (p0, r0, g0), (p1, r1v, g1) = _both(L, plugin, lambda L: L.Discriminator(64, 2.0), run)
    assert _rel(p1, p0) <= 1e-4
    assert abs(r1v - r0) <= 1e-3 * abs(r0)
    assert set(g0) == set(g1)
    worst = max(_rel(g1[k], g0[k]) for k in g0 if float(g0[k].norm()) > 0)
    assert worst <= 5e-3, worst


@needs_ref
def test_simple_swapping_evaluator_runs_unchanged_on_swapmodel(tmp_path, ref_layers):
    """evaluation/simple_swapping_evaluator.py:38-76 (file load through data.base_dataset.get_transform, the command
    sequence, ToPILImage + PNG save) driven once by ppst_b200.SwapModel (fp32 mode) and once by the reference's own E1 /
    E2 / G + PPSTModel methods on the CPU fallbacks; the two PNGs must agree."""
    from PIL import Image
    import ppst_b200
    from ppst_b200 import synthetic as S
    L, plugin = ref_layers
    plugin(False)
    from evaluation.simple_swapping_evaluator import SimpleSwappingEvaluator
    sds = S.make_networks(0)
    for name, seed, smooth in (("content", 11, True), ("style", 12, True)):
        img = S.make_portraits(1, 512, 512, seed=seed, smooth=smooth)[0]
        arr = ((img.clamp(-1, 1) + 1) * 127.5).round().byte().permute(1, 2, 0).numpy()
        Image.fromarray(arr).save(str(tmp_path / (name + ".png")))

    def run(model, tag):
        opt = Namespace(input_structure_image=str(tmp_path / "content.png"), input_texture_image=str(tmp_path / "style.png"),
                        texture_mix_alphas=[1.0], preprocess="scale_shortside", load_size=512, crop_size=512, dataroot=str(tmp_path),
                        isTrain=False, no_flip=True, name=tag, result_dir=str(tmp_path / "results"), dataname="synthetic",
                        checkpoints_dir=str(tmp_path), preprocess_crop_padding=None)
        ev = SimpleSwappingEvaluator(opt, "test")
        ev.evaluate(model, None)
        out = os.path.join(ev.output_dir(), "content_style_1.00.png")
        return np.asarray(Image.open(out).convert("RGB")).astype(np.float64)

    ours = ppst_b200.SwapModel(*sds, mode="fp32")          # noise weights are non-zero in the synthetic networks:
    noises = S.make_noises(1, 512, 512, seed=9)              # both sides get the same explicit maps
    ours.noises = noises
    a = run(ours, "ours")
    ref = RH.RefModel(sds, device="cpu", noises=noises)
    b = run(ref, "reference")
    mse = float(((a - b) ** 2).mean())
    psnr = 10 * math.log10(255.0 ** 2 / max(mse, 1e-12))
    print("evaluator PNG: ours (fp32 mode) vs reference CPU: %.1f dB, max |diff| %d" % (psnr, int(np.abs(a - b).max())))
    assert psnr >= 45.0, psnr
