"""Pin oracle/ppst_oracle.py against outputs of the unmodified reference (tests/golden/*.npz,
made by oracle/gen_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import ppst_oracle as O
from golden_util import check_summary, sub_state_dict

T = torch.from_numpy


def test_upfirdn2d_all_cases(golden):
    g = golden("ops")
    for name in g["upfirdn2d/names"]:
        up, down, p0, p1 = (int(v) for v in g["upfirdn2d/%s/cfg" % name])
        y = O.upfirdn2d(T(g["upfirdn2d/%s/x" % name]), T(g["upfirdn2d/%s/k" % name]), up, down, (p0, p1))
        want = g["upfirdn2d/%s/y" % name]
        assert tuple(y.shape) == want.shape, name
        assert np.abs(y.numpy() - want).max() <= 2e-6, name


def test_upfirdn2d_hot_size(golden):
    g = golden("ops")
    y = O.upfirdn2d(T(g["upfirdn2d_hot/x"]), O.make_kernel([1, 2, 1]))
    check_summary(g, "upfirdn2d_hot/y", y, atol=2e-6)


def test_fused_leaky_relu_fwd_bwd(golden):
    g = golden("ops")
    for name in g["flrelu/names"]:
        slope, scale = (float(v) for v in g["flrelu/%s/cfg" % name])
        x, b = T(g["flrelu/%s/x" % name]), T(g["flrelu/%s/b" % name])
        y = O.fused_leaky_relu(x, b, slope, scale)
        assert np.array_equal(y.numpy(), g["flrelu/%s/y" % name]), name
        gx, gb = O.fused_leaky_relu_backward(T(g["flrelu/%s/go" % name]), y, slope, scale)
        assert np.abs(gx.numpy() - g["flrelu/%s/gx" % name]).max() <= 1e-6
        assert np.abs(gb.numpy() - g["flrelu/%s/gb" % name]).max() <= 1e-4


LAYER_TOL = 2e-5


def _close(a, want, tol=LAYER_TOL):
    assert tuple(a.shape) == want.shape
    err = np.abs(a.numpy() - want).max()
    assert err <= tol, err


def test_conv_layers(golden):
    g = golden("layers")
    cases = {
        "convlayer_1x1_act": dict(ksize=1),
        "convlayer_3x3_reflect_in": dict(ksize=3, reflection_pad=True, norm="in"),
        "convlayer_down_reflect": dict(ksize=3, downsample=True, blur_kernel=(1, 2, 1), reflection_pad=True),
        "convlayer_skip_down": dict(ksize=1, downsample=True, blur_kernel=(1, 2, 1), activate=False, norm="in"),
        "convlayer_1x1_bias_in_noact": dict(ksize=1, activate=False, norm="in"),
        "convlayer_1x1_skip_plain": dict(ksize=1, activate=False),
    }
    for name, kw in cases.items():
        sd = sub_state_dict(g, name)
        y = O.conv_layer(sd, "", T(g[name + "/in0"]), **kw)
        _close(y, g[name + "/out"])


def test_res_blocks(golden):
    g = golden("layers")
    for name, norm in (("resblock_in", "in"), ("resblock_none", None)):
        y = O.res_block(sub_state_dict(g, name), "", T(g[name + "/in0"]), norm=norm)
        _close(y, g[name + "/out"])


def test_styled_conv_and_torgb(golden):
    g = golden("layers")
    for name, ups in (("styled_plain", False), ("styled_up_small", True), ("styled_up_fused", True)):
        sd = sub_state_dict(g, name)
        y = O.styled_conv(sd, "", T(g[name + "/in0"]), T(g[name + "/in1"]), upsample=ups, noise=T(g[name + "/in2"]))
        _close(y, g[name + "/out"])
    sd = sub_state_dict(g, "torgb")
    _close(O.to_rgb(sd, "", T(g["torgb/in0"]), T(g["torgb/in1"])), g["torgb/out"])


def test_generator_blocks(golden):
    g = golden("layers")
    sd = sub_state_dict(g, "residual_block")
    _close(O.residual_block(sd, "", T(g["residual_block/in0"])), g["residual_block/out"])
    sd = sub_state_dict(g, "generator_modulation")
    x, s = T(g["generator_modulation/in0"]), T(g["generator_modulation/in1"])
    y = x * O.equal_linear(s, sd["scale.weight"], sd["scale.bias"])[:, :, None, None] \
        + O.equal_linear(s, sd["bias.weight"], sd["bias.bias"])[:, :, None, None]
    _close(y, g["generator_modulation/out"])
    name = "respres_block"
    sd = sub_state_dict(g, name)
    x, s, n1, n2 = (T(g["%s/in%d" % (name, i)]) for i in range(4))
    skip = O.conv_layer(sd, "skip.", x, 1, activate=False)
    r = O.styled_conv(sd, "conv1.", x, s, noise=n1)
    r = O.styled_conv(sd, "conv2.", r, s, noise=n2)
    _close((skip + r) / O.SQRT2, g[name + "/out"])
    for name in ("upres_block_small", "upres_block_fused"):
        sd = sub_state_dict(g, name)
        x, s, n1, n2 = (T(g["%s/in%d" % (name, i)]) for i in range(4))
        skip = O.conv_layer(sd, "skip.", x, 1, activate=True)
        skip = torch.nn.functional.interpolate(skip, scale_factor=2, mode="bilinear", align_corners=False)
        r = O.styled_conv(sd, "conv1.", x, s, upsample=True, noise=n1)
        r = O.styled_conv(sd, "conv2.", r, s, noise=n2)
        _close((skip + r) / O.SQRT2, g[name + "/out"])


def test_function_py(golden):
    g = golden("function_py")
    x = T(g["x"])
    m, s = O.calc_mean_std(x)
    _close(m, g["mean"], 1e-6)
    _close(s, g["std"], 1e-6)
    _close(O.nor_mean_std(x), g["nor_mean_std"], 1e-5)
    _close(O.nor_mean(x)[0], g["nor_mean"], 1e-6)
    _close(O.calc_cov(x), g["cov"], 1e-5)


def test_mask_labels_bit_exact():
    lab = O.make_label_map(2, 64, 48)
    assert set(np.unique(lab.numpy())) <= {0, 1, 2}
    oh = O.mask_labels(lab[0].numpy())
    assert oh.shape == (3, 64, 48) and np.array_equal(oh.sum(0), np.ones((64, 48)))
    for i in range(3):
        assert np.array_equal(oh[i] == 1.0, lab[0].numpy() == i)
    assert [O.aggregate_part_label(i) for i in range(18)].count(2) == 2


@pytest.fixture(scope="module")
def full_nets():
    torch.set_num_threads(max(torch.get_num_threads(), 4))
    return O.make_networks(0)


def test_full_networks_512(golden, full_nets):
    """E1 / E2 (plain + corr-warp) / G (plain + extract) / Rselfcorr / corrm / warp at 512x512, B=1,
    against the reference run with the same synthetic state_dict."""
    g = golden("nets")
    sdE1, sdE2, sdG = full_nets
    c = O.make_portraits(1, 512, 512, seed=1234)
    s = O.make_portraits(1, 512, 512, seed=4321, smooth=True)
    noises = O.make_noises(1, 512, 512)
    tol = 1e-4      # north_star fp32 tolerance: max-abs 1e-4 per layer
    with torch.no_grad():
        tE1, tG = {}, {}
        sp_c = O.encoder_con(sdE1, c, taps=tE1)
        check_summary(g, "E1/sp_c", sp_c, atol=tol, stats_tol=1e-4)
        for k in ("FromRGB", "ResBlockDownBy1", "ResBlockDownBy2", "ResBlockDownBy4"):
            key = "tap/E1/" + (k if k == "FromRGB" else "DownToSpatialCode." + k)
            check_summary(g, key, tE1[k], atol=tol)
        gl_c, _ = O.encoder_col(sdE2, c)
        for i in range(4):
            assert np.abs(gl_c[i].numpy() - g["E2/gl_c/%d" % i]).max() <= 1e-6
        rgb, fea_c, fea_c1 = O.generator(sdG, sp_c, gl_c, extract_features=True, noises=noises, taps=tG)
        for k in ["SpatialCodeModulation"] + ["HeadResnetBlock%d" % i for i in range(4)] + \
                ["UpsamplingResBlock%d" % r for r in (16, 32, 64)]:
            check_summary(g, "tap/G/" + k, tG[k], atol=tol)
        check_summary(g, "G/rec_c", rgb, atol=tol, stats_tol=1e-4)
        check_summary(g, "G/fea_c", fea_c, atol=tol)
        check_summary(g, "G/fea_c1", fea_c1, atol=tol)
        sp_s = O.encoder_con(sdE1, s)
        gl_s, _ = O.encoder_col(sdE2, s)
        _, fea_s, fea_s1 = O.generator(sdG, sp_s, gl_s, extract_features=True, noises=noises)
        rs_c, rs_s = O.rselfcorr(fea_c1), O.rselfcorr(fea_s1)
        check_summary(g, "Rselfcorr/c", rs_c, atol=tol)
        fc, fs = torch.cat((fea_c, rs_c), 1), torch.cat((fea_s, rs_s), 1)
        corr = O.corrm(fs, fc)
        check_summary(g, "corrm/corr", corr, atol=2e-3)   # softmax at T=0.01 amplifies 1e-5 cosine noise
        assert (corr.sum(-1) - 1).abs().max() < 1e-4
        gl_s2, gl_w = O.encoder_col(sdE2, s, corrmatrix=[corr])      # list form, as the evaluator passes it
        for i in range(4):
            assert np.abs(gl_w[i].numpy() - g["E2/gl_w/%d" % i]).max() <= 2e-5
            assert np.abs(gl_s2[i].numpy() - g["E2/gl_s/%d" % i]).max() <= 1e-6
        check_summary(g, "G/swap", O.generator(sdG, sp_c, gl_w, noises=noises), atol=5e-4)
        check_summary(g, "G/swap_alpha0.5", O.generator(sdG, sp_c, O.lerp(gl_c, gl_w, 0.5), noises=noises), atol=5e-4)
        check_summary(g, "warp/full", O.model_warp(sp_s, corr), atol=1e-3)
        check_summary(g, "warp/unfold", O.model_warp(torch.nn.functional.avg_pool2d(s, 2), corr), atol=1e-3)
        # mask branch
        x2 = torch.cat([c, s], 0)
        lab = O.make_label_map(2, 512, 512)
        assert int(lab.to(torch.int64).sum()) == int(g["labels/sum"][0])
        onehot = torch.stack([torch.tensor(O.mask_labels(lab[b].numpy()), dtype=torch.float) for b in range(2)])
        corr2 = torch.cat([corr, O.corrm(fc, fs)], 0)
        v, pm, vw, pmw = O.encoder_col(sdE2, x2, mask=onehot, corrmatrix=corr2)
        for i in range(4):
            assert np.abs(v[i].numpy() - g["E2mask/v/%d" % i]).max() <= 1e-6
            assert np.abs(vw[i].numpy() - g["E2mask/vw/%d" % i]).max() <= 2e-5
        for i in range(12):
            assert np.abs(pm[i].numpy() - g["E2mask/pm/%d" % i]).max() <= 1e-6
            assert np.abs(pmw[i].numpy() - g["E2mask/pmw/%d" % i]).max() <= 2e-5


def test_plain_path_non_square(golden, full_nets):
    g = golden("nets")
    sdE1, sdE2, sdG = full_nets
    c2 = O.make_portraits(1, 256, 384, seed=5)
    n2 = O.make_noises(1, 256, 384, seed=6)
    with torch.no_grad():
        sp2 = O.encoder_con(sdE1, c2)
        gl2, _ = O.encoder_col(sdE2, c2)
        check_summary(g, "plain256x384/sp", sp2, atol=1e-4)
        check_summary(g, "plain256x384/rec", O.generator(sdG, sp2, gl2, noises=n2), atol=1e-4)


def test_guided_oracle_box_means_match_opencv_boxfilter():
    """oracle/guided_oracle.py is parity-unpinned (opencv-contrib absent); what CAN be pinned is its box mean, against
    plain OpenCV's cv2.boxFilter with BORDER_REFLECT (the primitive ximgproc's guided filter is built on)."""
    cv2 = pytest.importorskip("cv2")
    from oracle import guided_oracle as G
    rng = np.random.default_rng(0)
    a = rng.integers(0, 65026, (97, 120)).astype(np.int64)
    f = (rng.standard_normal((97, 120)) * 100).astype(np.float32)
    for r in (1, 7, 30):
        k = (2 * r + 1, 2 * r + 1)
        want = cv2.boxFilter(a.astype(np.float32), cv2.CV_32F, k, borderType=cv2.BORDER_REFLECT)
        assert np.array_equal(G.box_mean_int(a, r), want)
        want = cv2.boxFilter(f, cv2.CV_32F, k, borderType=cv2.BORDER_REFLECT)
        assert np.abs(G.box_mean_f32(f, r) - want).max() <= 2e-5


def test_guided_oracle_properties():
    """Known properties of the guided filter: a constant source stays constant; filtering an image with itself as the
    guide at eps -> 0 reproduces it; quantisation follows tensor2im (truncation)."""
    from oracle import guided_oracle as G
    rng = np.random.default_rng(1)
    guide = rng.integers(0, 256, (80, 96, 3)).astype(np.uint8)
    const = np.full((80, 96, 3), 77, np.uint8)
    assert np.array_equal(G.guided_filter_u8(guide, const, 9, np.float32(26.01)), const)
    out = G.guided_filter_u8(guide, guide, 9, np.float32(1e-3))
    assert np.abs(out.astype(int) - guide.astype(int)).max() <= 1
    assert G.quantize_u8(np.array([-1.0, 0.0, 0.999, 1.0, 1.5])).tolist() == [0, 127, 254, 255, 255]


# ---------------------------------------------------------------------------------------------- discriminator (8(f)4)
D_KEYS = ["convs.0.Conv.weight", "convs.0.Act.bias", "convs.1.conv1.Conv.weight", "convs.1.conv2.Conv.weight",
          "convs.1.conv2.Act.bias", "convs.1.skip.Conv.weight", "convs.3.conv1.Conv.weight", "convs.6.conv2.Conv.weight",
          "final_conv.Conv.weight",
          "final_linear.0.weight", "final_linear.0.bias", "final_linear.1.weight"]


def _d_state(requires_grad=True):
    sd = O.make_discriminator(256, seed=0)
    return {k: (v.clone().requires_grad_(True) if (requires_grad and "Blur" not in k) else v) for k, v in sd.items()}


def test_discriminator_forward_losses_and_gradients(golden):
    """oracle.discriminator / d_losses / r1_loss against the reference Discriminator run by oracle/gen_golden.py:
    output, both loss values, parameter gradients of the D step, and the double-backward gradients of the R1 penalty."""
    g = golden("discriminator")
    real, fake = T(g["real"]), T(g["fake"])
    sd = _d_state()
    pred = O.discriminator(sd, real)
    assert np.abs(pred.detach().numpy() - g["pred_real"]).max() <= 1e-5
    losses = O.d_losses(sd, real, fake)
    assert abs(losses["D_real"].item() - float(g["D_real"])) <= 1e-5
    assert abs(losses["D_rec"].item() - float(g["D_rec"])) <= 1e-6
    (losses["D_real"] + losses["D_rec"]).backward()
    for k in D_KEYS:
        grad = sd["stylegan2_D." + k].grad
        scale = float(g["dstep_grad/%s/stats" % k][2])
        check_summary(g, "dstep_grad/" + k, grad, atol=2e-4 * scale + 1e-9)
    sd = _d_state()
    pen = O.r1_loss(sd, real)
    assert np.abs(pen.detach().numpy() - g["r1/penalty"]).max() <= 1e-4 * np.abs(g["r1/penalty"]).max()
    pen.mean().backward()
    for k in D_KEYS:
        grad = sd["stylegan2_D." + k].grad
        if grad is None:                          # biases: the penalty does not depend on them (a.e.)
            assert float(g["r1_grad/%s/stats" % k][2]) == 0.0, k
            continue
        scale = float(g["r1_grad/%s/stats" % k][2])
        check_summary(g, "r1_grad/" + k, grad, atol=1e-3 * scale + 1e-12)


def test_stage2_losses_match_the_reference_run(golden):
    """Training stage 2 (ppst_model.py:105-137,158-236): every term of the generator and discriminator losses of the
    oracle's restatement against tests/golden/stage2.npz (PPSTModel.compute_generator_losses / compute_discriminator_losses
    of the unmodified reference at 512^2, batch 2, lpips stubbed on both sides).  Forward only (about 80 s of CPU)."""
    z = golden("stage2")
    sdE1, sdE2, sdG, sdD, real, mask, queues = O.stage2_inputs()
    sdDp = {"stylegan2_D." + k: v for k, v in sdD.items()}
    with torch.no_grad():
        losses, pushes = O.stage2_generator_losses(sdE1, sdE2, sdG, sdDp, real, mask, queues)
        d = O.stage2_discriminator_losses(sdE1, sdE2, sdG, sdDp, real)
    for k, v in losses.items():
        ref = float(z["G/" + k])
        assert abs(float(v) - ref) <= 1e-4 * max(abs(ref), 1e-2), (k, float(v), ref)
    for k, v in d.items():
        ref = float(z["D/" + k])
        assert abs(float(v) - ref) <= 1e-4 * max(abs(ref), 1e-2), (k, float(v), ref)
    # the keys pushed into the region-contrast queues: level l's first key lands in column 0 of queue l
    from oracle import rscl_oracle as R
    for lvl, keys in pushes:
        q, ptr = queues[lvl], 0
        for kk in keys:
            q, ptr = R.enqueue(q, ptr, kk)
        assert ptr == int(z["queue_ptr/%d" % lvl])
        assert float((q[:, :8] - torch.from_numpy(z["queue_head/%d" % lvl])).abs().max()) <= 1e-5
