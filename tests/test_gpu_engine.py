"""Network-level parity on the GPU: E1 / E2 / G / correlation swap through ppst_b200 against the CPU oracle
on identical synthetic weights and inputs.

Tolerances are the ones north_star states:
  fp32 mode: max-abs 1e-4 per layer
  bf16 mode: per-layer relative L2 <= 1e-2 and output PSNR >= 45 dB
"""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import ppst_oracle as O  # noqa: E402
from golden_util import check_summary  # noqa: E402

FP32_TOL = 1e-4


def nchw(x_nhwc):
    return x_nhwc.float().permute(0, 3, 1, 2).cpu()


def rel_l2(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


def psnr(a, ref):
    peak = float(ref.max() - ref.min())
    mse = float(((a - ref) ** 2).mean())
    return 10 * math.log10(peak * peak / max(mse, 1e-30))


@pytest.fixture(scope="module")
def world():
    """Oracle outputs for one content/style pair at 512x512 (computed once; ~1 min of CPU)."""
    torch.set_num_threads(max(torch.get_num_threads(), 8))
    sdE1, sdE2, sdG = O.make_networks(0)
    c = O.make_portraits(1, 512, 512, seed=1234)
    s = O.make_portraits(1, 512, 512, seed=4321, smooth=True)
    noises = O.make_noises(1, 512, 512)
    w = dict(sd=(sdE1, sdE2, sdG), c=c, s=s, noises=noises)
    with torch.no_grad():
        w["tE1"], w["tG"], w["tE2"] = {}, {}, {}
        w["sp_c"] = O.encoder_con(sdE1, c, taps=w["tE1"])
        w["gl_c"], _ = O.encoder_col(sdE2, c, taps=w["tE2"])
        w["rgb_c"], w["fea_c"], w["fea_c1"] = O.generator(sdG, w["sp_c"], w["gl_c"], extract_features=True,
                                                          noises=noises, taps=w["tG"])
        w["sp_s"] = O.encoder_con(sdE1, s)
        w["gl_s"], _ = O.encoder_col(sdE2, s)
        _, w["fea_s"], w["fea_s1"] = O.generator(sdG, w["sp_s"], w["gl_s"], extract_features=True, noises=noises)
        w["rs_c"], w["rs_s"] = O.rselfcorr(w["fea_c1"]), O.rselfcorr(w["fea_s1"])
        w["fc"] = torch.cat((w["fea_c"], w["rs_c"]), 1)
        w["fs"] = torch.cat((w["fea_s"], w["rs_s"]), 1)
        w["corr"] = O.corrm(w["fs"], w["fc"])
        _, w["gl_w"] = O.encoder_col(sdE2, s, corrmatrix=w["corr"])
        w["swap"] = O.generator(sdG, w["sp_c"], w["gl_w"], noises=noises)
    return w


@pytest.fixture(scope="module")
def model_fp32(world):
    from ppst_b200 import SwapModel
    m = SwapModel(*world["sd"], mode="fp32")
    m.noises = world["noises"]
    return m


@pytest.fixture(scope="module")
def model_bf16(world):
    from ppst_b200 import SwapModel
    m = SwapModel(*world["sd"], mode="bf16")
    m.noises = world["noises"]
    return m


# ------------------------------------------------------------------------------------------------ fp32 mode
def test_fp32_encoders_per_layer(world, model_fp32, golden):
    e = model_fp32.engine
    x = world["c"].cuda()
    taps = {}
    sp = e.encoder_con(x, taps=taps)
    for k, ref in world["tE1"].items():
        err = (nchw(taps[k]) - ref).abs().max()
        assert err <= FP32_TOL, ("E1." + k, float(err))
    assert (nchw(sp) - world["sp_c"]).abs().max() <= FP32_TOL
    check_summary(golden("nets"), "E1/sp_c", nchw(sp), atol=2e-4)           # the reference's own numbers
    taps = {}
    gl, gw = e.encoder_col(x, taps=taps)
    assert gw == []
    for k, ref in world["tE2"].items():
        err = (nchw(taps[k]) - ref).abs().max()
        assert err <= FP32_TOL, ("E2." + k, float(err))
    g = golden("nets")
    for i in range(4):
        assert (gl[i].cpu() - world["gl_c"][i]).abs().max() <= 1e-5
        assert np.abs(gl[i].cpu().numpy() - g["E2/gl_c/%d" % i]).max() <= 1e-5


def test_fp32_generator_per_layer(world, model_fp32, golden):
    e = model_fp32.engine
    sp = e.to_nhwc(world["sp_c"])
    taps = {}
    rgb, fea, fea1 = e.generator(sp, [t.cuda() for t in world["gl_c"]], extract_features=True,
                                 noises=model_fp32.noises, taps=taps)
    for k in ["SpatialCodeModulation"] + ["HeadResnetBlock%d" % i for i in range(4)] + \
            ["UpsamplingResBlock%d" % r for r in (16, 32, 64)] + ["feat_cat", "feat1_cat"]:
        err = (nchw(taps[k]) - world["tG"][k]).abs().max()
        assert err <= FP32_TOL, ("G." + k, float(err))
    assert (nchw(rgb) - world["rgb_c"]).abs().max() <= 2e-4
    assert (nchw(fea) - world["fea_c"]).abs().max() <= 2e-4
    assert (nchw(fea1) - world["fea_c1"]).abs().max() <= 2e-4
    g = golden("nets")
    check_summary(g, "G/rec_c", nchw(rgb), atol=3e-4)
    check_summary(g, "G/fea_c", nchw(fea), atol=3e-4)
    rgb2 = e.generator(sp, [t.cuda() for t in world["gl_c"]], noises=model_fp32.noises)
    assert torch.equal(rgb2, rgb)                                             # deterministic, extract does not perturb rgb


def test_fp32_correlation_swap(world, model_fp32, golden):
    m = model_fp32
    rs = m.Rselfcorr(world["fea_c1"])
    assert (rs.cpu() - world["rs_c"]).abs().max() <= 1e-5
    corr = m.corrm(world["fs"], world["fc"])
    assert (corr.cpu() - world["corr"]).abs().max() <= 2e-3            # softmax at T=0.01 amplifies 1e-6 cosine noise
    assert (corr.sum(-1) - 1).abs().max() <= 1e-4
    gl, gl_w = m.encode2(world["s"], [world["corr"]])                  # list form (simple_swapping_evaluator.py:53)
    g = golden("nets")
    for i in range(4):
        assert (gl_w[i].cpu() - world["gl_w"][i]).abs().max() <= 2e-5
        assert np.abs(gl_w[i].cpu().numpy() - g["E2/gl_w/%d" % i]).max() <= 5e-5
    out = m.decode(world["sp_c"], [t.cuda() for t in world["gl_w"]])
    assert (out.cpu() - world["swap"]).abs().max() <= 3e-4
    check_summary(g, "G/swap", out.cpu(), atol=8e-4)
    # PPSTModel.warp, both branches
    wf = m.warp(world["sp_s"], world["corr"])
    assert (wf.cpu() - O.model_warp(world["sp_s"], world["corr"])).abs().max() <= 1e-4
    lo = torch.nn.functional.avg_pool2d(world["s"], 2)
    assert (m.warp(lo, world["corr"]).cpu() - O.model_warp(lo, world["corr"])).abs().max() <= 1e-4


def test_fp32_region_branch_integer_labels(world, model_fp32, golden):
    """E2.forward(mask=...) (training path): masked GAP/GMP keyed by the integer label map."""
    g = golden("nets")
    x2 = torch.cat([world["c"], world["s"]], 0)
    lab = O.make_label_map(2, 512, 512)
    onehot = torch.stack([torch.tensor(O.mask_labels(lab[b].numpy()), dtype=torch.float) for b in range(2)])
    corr2 = torch.cat([world["corr"], O.corrm(world["fc"], world["fs"])], 0)
    v, pm, vw, pmw = model_fp32.E2_forward(x2, mask=onehot, corrmatrix=corr2)
    assert len(pm) == 12 and len(pmw) == 12
    for i in range(4):
        assert np.abs(v[i].cpu().numpy() - g["E2mask/v/%d" % i]).max() <= 1e-5
        assert np.abs(vw[i].cpu().numpy() - g["E2mask/vw/%d" % i]).max() <= 5e-5
    for i in range(12):
        assert np.abs(pm[i].cpu().numpy() - g["E2mask/pm/%d" % i]).max() <= 1e-5, i
        assert np.abs(pmw[i].cpu().numpy() - g["E2mask/pmw/%d" % i]).max() <= 5e-5, i


def test_fp32_plain_path_other_sizes(model_fp32, golden):
    """Plain E1 -> E2 -> G works at sizes other than 512x512 (only the correlation path is 512-only)."""
    g = golden("nets")
    c2 = O.make_portraits(1, 256, 384, seed=5)
    model_fp32.noises = O.make_noises(1, 256, 384, seed=6)
    try:
        sp, gl = model_fp32.encode(c2)
        check_summary(g, "plain256x384/sp", sp.cpu(), atol=2e-4)
        rec = model_fp32.decode(sp, gl)
        check_summary(g, "plain256x384/rec", rec.cpu(), atol=3e-4)     # exercises the nearest-upsample StyledConv path
    finally:
        model_fp32.noises = O.make_noises(1, 512, 512)


def test_function_py_helpers(golden):
    from ppst_b200 import function as Fn
    g = golden("function_py")
    x = torch.from_numpy(g["x"]).cuda()
    m, s = Fn.calc_mean_std(x)
    assert np.abs(m.cpu().numpy() - g["mean"]).max() <= 1e-6 and np.abs(s.cpu().numpy() - g["std"]).max() <= 1e-6
    assert np.abs(Fn.nor_mean_std(x).cpu().numpy() - g["nor_mean_std"]).max() <= 1e-5
    assert np.abs(Fn.nor_mean(x)[0].cpu().numpy() - g["nor_mean"]).max() <= 1e-6
    assert np.abs(Fn.calc_cov(x).cpu().numpy() - g["cov"]).max() <= 1e-5


# ------------------------------------------------------------------------------------------------ bf16 mode
def test_bf16_per_layer_and_output(world, model_bf16):
    """Per-layer relative L2 <= 1e-2: every block is fed the ORACLE's input for that block, so the number is
    the error of that layer alone; the drift accumulated through a whole network is bounded separately."""
    e = model_bf16.engine
    assert e.use_tc, "the tcgen05 path must be the one that runs in bf16 mode"
    x = world["c"].cuda()
    # --- E1, block by block from the oracle's activations
    y = e.from_rgb(e.E1, x)
    assert rel_l2(nchw(y), world["tE1"]["FromRGB"]) <= 1e-2
    prev = "FromRGB"
    for i in (1, 2, 4):
        out = e._res_block(e.E1["res%d" % i], e.to_nhwc(world["tE1"][prev]), True)
        r = rel_l2(nchw(out), world["tE1"]["ResBlockDownBy%d" % i])
        assert r <= 1e-2, ("E1.ResBlockDownBy%d" % i, r)
        prev = "ResBlockDownBy%d" % i
    for i in (1, 2, 4):      # E2 blocks (no InstanceNorm)
        src = "level9" if i == 1 else "level%d" % {2: 0, 4: 1}[i]
        dst = "level%d" % {1: 0, 2: 1, 4: 2}[i]
        out = e._res_block(e.E2["res%d" % i], e.to_nhwc(world["tE2"][src]), False)
        r = rel_l2(nchw(out), world["tE2"][dst])
        assert r <= 1e-2, ("E2.ResBlockDownBy%d" % i, r)
    # --- accumulated drift through the whole encoders
    taps = {}
    sp = e.encoder_con(x, taps=taps)
    drift = {k: rel_l2(nchw(taps[k]), ref) for k, ref in world["tE1"].items()}
    drift["sp"] = rel_l2(nchw(sp), world["sp_c"])
    assert max(drift.values()) <= 2e-2, drift
    gl, _ = e.encoder_col(x)
    for i in range(4):
        assert rel_l2(gl[i].cpu(), world["gl_c"][i]) <= 1e-2
    # --- generator: each block from the oracle's input of that block (per-layer error) ...
    G = e.G
    st = e.styles([ops_l2(t) for t in world["gl_c"]])
    names = ["SpatialCodeModulation"] + ["HeadResnetBlock%d" % i for i in range(4)] + \
            ["UpsamplingResBlock%d" % r for r in (16, 32, 64)]
    from ppst_b200.engine import HEAD_CH, UP_CH, INV_SQRT2
    for i in range(4):
        B = G["head%d" % i]
        xin = e.to_nhwc(world["tG"][names[i]])
        skip = e.conv(xin, B["skip"])[0] if "skip" in B else xin
        y, a, b = e._styled(B["conv1"], xin, st["head%d.conv1" % i], world["noises"]["HeadResnetBlock%d.conv1" % i].cuda())
        h = e.apply(y, a, b)
        y, a, b = e._styled(B["conv2"], h, st["head%d.conv2" % i], world["noises"]["HeadResnetBlock%d.conv2" % i].cuda())
        out = e.apply(y, a, b, res=skip, res_mode=1, post_gain=INV_SQRT2)
        r = rel_l2(nchw(out), world["tG"]["HeadResnetBlock%d" % i])
        assert r <= 1e-2, ("G.HeadResnetBlock%d" % i, r)
    for j, (res, ci, co) in enumerate(UP_CH):
        B = G["up%d" % res]
        name = "UpsamplingResBlock%d" % res
        xin = e.to_nhwc(world["tG"][names[4 + j]])
        skip = e.conv(xin, B["skip"])[0] if "skip" in B else xin
        y, a, b = e._styled(B["conv1"], xin, st["up%d.conv1" % res], world["noises"][name + ".conv1"].cuda(), upsample=True)
        h = e.apply(y, a, b)
        y, a, b = e._styled(B["conv2"], h, st["up%d.conv2" % res], world["noises"][name + ".conv2"].cuda())
        out = e.apply(y, a, b, res=skip, res_mode=2, post_gain=INV_SQRT2)
        r = rel_l2(nchw(out), world["tG"][name])
        assert r <= 1e-2, ("G." + name, r)
    # ... and the drift accumulated through the whole generator (up to 14 StyledConvs) from the oracle's codes
    taps = {}
    sp_ref = e.to_nhwc(world["sp_c"])
    rgb, fea, fea1 = e.generator(sp_ref, [t.cuda() for t in world["gl_c"]], extract_features=True,
                                 noises=model_bf16.noises, taps=taps)
    gdrift = {k: rel_l2(nchw(taps[k]), world["tG"][k]) for k in names}
    assert max(gdrift.values()) <= 2e-2, gdrift
    assert psnr(nchw(rgb), world["rgb_c"]) >= 45.0, psnr(nchw(rgb), world["rgb_c"])
    assert rel_l2(nchw(fea), world["fea_c"]) <= 2e-2
    assert rel_l2(nchw(fea1), world["fea_c1"]) <= 2e-2


def ops_l2(t):
    from ppst_b200 import ops
    return ops.l2_normalize(t.cuda().float(), 1e-8, 0)


def test_bf16_end_to_end_swap(world, model_bf16):
    out = model_bf16.swap(world["c"], world["s"])
    p = psnr(out.cpu(), world["swap"])
    assert p >= 45.0, p
    out2 = model_bf16.swap(world["c"], world["s"])
    assert torch.equal(out, out2)                                      # run-to-run deterministic (no fp atomics)
    # the CUDA-graph replay of the same step is bit-identical to the eager launches, also from pinned host inputs
    g1 = model_bf16.swap_graph(world["c"].pin_memory(), world["s"].pin_memory()).clone()
    g2 = model_bf16.swap_graph(world["c"].cuda(), world["s"].cuda()).clone()
    assert torch.equal(g1, out) and torch.equal(g2, out)
    # the streaming form (host buffers in and out, copies on their own streams): alternate two different batches through
    # the double-buffered pipeline; every result must be the bits of the plain swap of ITS inputs
    ca, sa = world["c"].pin_memory(), world["s"].pin_memory()
    cb, sb = world["s"].pin_memory(), world["c"].pin_memory()
    ref_b = model_bf16.swap(world["s"], world["c"]).cpu()
    outs = [torch.empty(out.shape, dtype=torch.float32).pin_memory() for _ in range(5)]
    evs = [model_bf16.swap_pipelined(*((ca, sa) if i % 2 == 0 else (cb, sb)), outs[i]) for i in range(5)]
    for i, ev in enumerate(evs):
        ev.synchronize()
        assert torch.equal(outs[i], out.cpu() if i % 2 == 0 else ref_b), i


def test_batch_rows_are_independent(world, model_bf16):
    """Partitioning pairs over GPUs must not change a sample's result: a row computed inside a batch of 2 equals
    the same row computed alone (InstanceNorm is per sample; no cross-sample arithmetic anywhere)."""
    m = model_bf16
    saved = m.noises
    m.noises = None
    try:
        c2 = torch.cat([world["c"], world["s"]], 0)
        s2 = torch.cat([world["s"], world["c"]], 0)
        both = m.swap(c2, s2)
        one = m.swap(world["c"], world["s"])
        assert torch.equal(both[0:1], one)
    finally:
        m.noises = saved


def test_swap_grid_cells_equal_pairwise_swaps(world, model_bf16):
    """swap_grid (grid evaluator, content_style_grid_generation_evaluator.py:57-91, with per-image work done once):
    every cell is bit-identical to swap(content_i, style_j); a ragged last chunk of cells is covered."""
    m = model_bf16
    contents = torch.cat([world["c"], world["s"].flip(3)], 0)
    styles = torch.cat([world["s"], world["c"].flip(2), world["c"]], 0)
    grid = m.swap_grid(contents, styles, cell_batch=4)
    assert grid.shape == (2, 3, 3, 512, 512)
    for i in range(2):
        for j in range(3):
            one = m.swap(contents[i:i + 1], styles[j:j + 1])
            assert torch.equal(grid[i, j], one[0]), (i, j)
    assert psnr(grid[0, 0].cpu()[None], world["swap"]) >= 45.0


def test_plain_swap_1024_matches_oracle(world):
    """BASELINE config 4: the plain swap G(E1(c), E2(s)) is the only one defined away from 512^2.  Checked at
    1024 x 512 (the oracle needs ~20 s per full 1024^2 generator pass; the kernels are size-agnostic)."""
    from ppst_b200 import SwapModel
    sdE1, sdE2, sdG = world["sd"]
    c = O.make_portraits(1, 1024, 512, seed=77)
    s = O.make_portraits(1, 1024, 512, seed=78, smooth=True)
    noises = O.make_noises(1, 1024, 512)
    with torch.no_grad():
        want = O.generator(sdG, O.encoder_con(sdE1, c), O.encoder_col(sdE2, s)[0], noises=noises)
    m = SwapModel(sdE1, sdE2, sdG, mode="bf16")
    m.noises = noises
    got = m.plain_swap(c, s)
    assert got.shape == want.shape == (1, 3, 1024, 512)
    assert psnr(got.cpu(), want) >= 45.0
    m32 = SwapModel(sdE1, sdE2, sdG, mode="fp32")
    m32.noises = noises
    assert (m32.plain_swap(c, s).cpu() - want).abs().max() <= 2e-3


def test_evaluator_sequence_through_command_dispatch(world, model_fp32):
    """evaluation/simple_swapping_evaluator.py:44-59, verbatim call sequence, through model(..., command=)."""
    model = model_fp32
    structure_image, texture_image = world["c"], world["s"]
    structure_code, source_texture_code = model(structure_image, command="encode")
    fea_c, fea_c1 = model(structure_image, command="extract_feat_from_image")
    fea_s, fea_s1 = model(texture_image, command="extract_feat_from_image")
    fea_c = torch.cat((fea_c, model(fea_c1, command="Rselfcorr")), dim=1)
    fea_s = torch.cat((fea_s, model(fea_s1, command="Rselfcorr")), dim=1)
    corrmatrix = model(fea_s, fea_c, command="corrm")
    _, gl_w = model(texture_image.cuda(), [corrmatrix], command="encode2")
    texture_code = O.lerp(source_texture_code, gl_w, 1.0)
    output_image = model(structure_code, texture_code, target=None, command="decode")
    assert output_image.shape == (1, 3, 512, 512)
    assert (output_image.cpu() - world["swap"]).abs().max() <= 2e-3
    with pytest.raises(ValueError):
        model(structure_image)


def test_grid_evaluator_sequence_with_guided_filter_target(world, model_fp32):
    """evaluation/content_style_grid_generation_evaluator.py:57-88, verbatim call sequence for one (content, style)
    cell, including decode(..., target=structure_image): the guided-filter post-process of ppst_model.py:288-306."""
    from oracle import guided_oracle as G
    model = model_fp32
    style_image, structure_image = world["s"], world["c"]
    f0, f1 = model(style_image, command="extract_feat_from_image")
    fea_ss, fea_ss1 = [f0], [f1]
    sps = [model(structure_image, command="encode")[0]]
    f0, f1 = model(structure_image, command="extract_feat_from_image")
    fea_s, fea_s1 = [f0], [f1]
    fea_s[0] = torch.cat((fea_s[0], model(fea_s1[0], command="Rselfcorr")), dim=1)
    if fea_ss[0].size(1) == 256:
        fea_ss[0] = torch.cat((fea_ss[0], model(fea_ss1[0], command="Rselfcorr")), dim=1)
    corrmatrix = model(fea_ss[0], fea_s[0], command="corrm")
    gl, gl_w = model(style_image.cuda(), corrmatrix, command="encode2")
    out = model(sps[0], gl_w, command="decode", target=structure_image)[0]
    assert out.shape == (3, 512, 512)
    want = G.guided_smooth(world["swap"].numpy(), structure_image.numpy())[0]
    lv = np.abs(out.cpu().numpy() - want) * 127.5                       # in uint8 levels
    assert lv.max() <= 3.0 and lv.mean() <= 0.05, (float(lv.max()), float(lv.mean()))
    # every value is an exact uint8 level mapped back to [-1, 1]
    assert np.abs(np.rint((out.cpu().numpy() + 1) * 127.5) - (out.cpu().numpy() + 1) * 127.5).max() <= 1e-4


@pytest.mark.gpu
def test_random_noise_mode_and_noise_routing_for_other_batch_sizes():
    """(a) noises="random" draws fresh N(0,1) NoiseInjection maps per generator pass like the reference
    (stylegan2_layers.py:376-399), reproducibly from noise_seed; (b) a noise dict set for batch B is repeated for passes over
    another batch size in extract_feat / extract_feat_from_image as well (they used to pass the dict through unchanged)."""
    import ppst_b200
    from ppst_b200 import synthetic as S
    sds = S.make_networks(0)                     # synthetic networks have non-zero noise weights
    c = S.make_portraits(2, 256, 256, seed=3).cuda()
    m1 = ppst_b200.SwapModel(*sds, mode="bf16", noise_seed=7)
    m2 = ppst_b200.SwapModel(*sds, mode="bf16", noise_seed=7)
    sp, gl = m1.encode(c)
    a1, a2 = m1.decode(sp, gl), m1.decode(sp, gl)
    b1 = m2.decode(*m2.encode(c))
    assert not torch.equal(a1, a2)               # fresh noise on every pass
    assert torch.equal(a1, b1)                   # same seed, same stream of draws
    m1.noises = None
    off = m1.decode(sp, gl)
    assert not torch.equal(off, a1)
    m3 = ppst_b200.SwapModel(*sds, mode="bf16")
    m3.noises = S.make_noises(1, 256, 256, seed=5)          # batch-1 maps, passes over batch 2
    f_a = m3.extract_feat_from_image(c)
    f_b = [torch.cat([m3.extract_feat_from_image(c[i:i + 1])[k] for i in range(2)]) for k in range(2)]
    assert all(torch.equal(x, y) for x, y in zip(f_a, f_b))


@pytest.mark.gpu
def test_model_on_a_non_current_device():
    """SwapModel(device="cuda:1") while cuda:0 is current must run on cuda:1 (launches used to go to the current
    device's stream with the other device's pointers); the raw operators refuse tensors of a non-current device."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import ppst_b200
    from ppst_b200 import ops, synthetic as S
    torch.cuda.set_device(0)
    sds = S.make_networks(0)
    c = S.make_portraits(1, 256, 256, seed=3)
    s = S.make_portraits(1, 256, 256, seed=4, smooth=True)
    m0 = ppst_b200.SwapModel(*sds, device="cuda:0", mode="bf16")
    m1 = ppst_b200.SwapModel(*sds, device="cuda:1", mode="bf16")
    assert torch.cuda.current_device() == 0
    out0 = m0.plain_swap(c, s)
    out1 = m1.plain_swap(c, s)
    assert out1.device.index == 1 and torch.cuda.current_device() == 0
    assert torch.equal(out0.cpu(), out1.cpu())
    with pytest.raises(RuntimeError):
        ops.l2_normalize(torch.randn(2, 8, device="cuda:1"), 1e-8, 0)
