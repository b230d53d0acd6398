"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on CPU.

TEST INFRASTRUCTURE ONLY.  Run in the build container (the reference tree does not exist on the
GPU box):      python oracle/gen_golden.py
The fixtures pin oracle/ppst_oracle.py (tests/test_oracle_golden.py) and, on the GPU box, the CUDA
path (tests/test_gpu_*.py).  Everything is seeded; weights come from oracle.ppst_oracle.make_state_dict
so that the same tensors can be rebuilt anywhere without the reference.
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ppst_oracle as O  # noqa: E402
from oracle import ref_harness as RH  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
N_SAMPLE = 4096


def sample_idx(numel, key):
    g = torch.Generator().manual_seed(abs(hash_str(key)) % (2 ** 31))
    return torch.randint(0, numel, (min(N_SAMPLE, numel),), generator=g)


def hash_str(s):
    h = 2166136261
    for ch in s.encode():
        h = ((h ^ ch) * 16777619) & 0xFFFFFFFF
    return h


def summarize(store, key, t):
    """Store a fixed pseudo-random subsample + moments of a big tensor."""
    t = t.detach().to(torch.float32).contiguous()
    flat = t.reshape(-1)
    idx = sample_idx(flat.numel(), key)
    store[key + "/shape"] = np.array(t.shape, dtype=np.int64)
    store[key + "/sample"] = flat[idx].numpy()
    store[key + "/stats"] = np.array([flat.mean().item(), flat.std().item(), flat.abs().max().item()],
                                     dtype=np.float64)


# ----------------------------------------------------------------------------------------------
def gen_ops():
    RH.install()
    import models.networks.stylegan2_layers  # noqa: F401  (pulls in stylegan2_op)
    up_mod = sys.modules["models.networks.stylegan2_op.upfirdn2d"]
    fa_mod = sys.modules["models.networks.stylegan2_op.fused_act"]
    assert not up_mod.use_custom_kernel and not fa_mod.use_custom_kernel
    store = {}
    g = torch.Generator().manual_seed(2024)
    k3 = O.make_kernel([1, 2, 1])
    k4 = O.make_kernel([1, 3, 3, 1])
    k2 = O.make_kernel([1, 1])
    k1 = O.make_kernel([1])
    k43 = torch.rand(4, 3, generator=g)
    k5 = torch.rand(5, 5, generator=g)
    cases = [  # (name, shape, kernel, up, down, pad)
        ("m2_blur_reflect", (2, 3, 19, 19), k3, 1, 1, (0, 0)),
        ("m2_blur_pad10", (1, 4, 16, 16), k3, 1, 1, (1, 0)),
        ("m2_blur_odd", (1, 2, 13, 29), k3, 1, 1, (1, 1)),
        ("m1_blur4", (2, 2, 18, 18), k4, 1, 1, (2, 1)),
        ("m1_asym", (1, 3, 11, 17), k43, 1, 1, (1, 2)),
        ("m3_up2_k4", (1, 3, 9, 9), k4 * 4, 2, 1, (2, 1)),
        ("m4_up2_k2", (2, 2, 8, 12), k2 * 4, 2, 1, (0, 0)),
        ("m5_down2_k4", (1, 3, 16, 16), k4, 1, 2, (1, 1)),
        ("m6_down2_k2", (1, 2, 14, 10), k2, 1, 2, (0, 0)),
        ("k1_identity", (1, 2, 7, 7), k1, 1, 1, (0, 0)),
        ("neg_pad_crop", (1, 2, 12, 12), k3, 1, 1, (-1, -2)),
        ("neg_pos_pad", (1, 2, 12, 12), k3, 1, 1, (2, -1)),
        ("generic_k5", (1, 2, 15, 15), k5, 1, 1, (2, 2)),
        ("generic_up3", (1, 2, 6, 6), k5, 3, 2, (3, 2)),
        ("wide_tile", (1, 1, 40, 150), k3, 1, 1, (1, 1)),
    ]
    names = []
    for name, shape, k, up, down, pad in cases:
        x = torch.randn(*shape, generator=g)
        y = up_mod.upfirdn2d(x, k, up=up, down=down, pad=pad)
        store["upfirdn2d/%s/x" % name] = x.numpy()
        store["upfirdn2d/%s/k" % name] = k.numpy()
        store["upfirdn2d/%s/cfg" % name] = np.array([up, down, pad[0], pad[1]], dtype=np.int64)
        store["upfirdn2d/%s/y" % name] = y.numpy()
        names.append(name)
    store["upfirdn2d/names"] = np.array(names)
    # hot-path sized call (encoder blur): sampled
    x = torch.randn(1, 32, 131, 131, generator=g)
    y = up_mod.upfirdn2d(x, k3, pad=(0, 0))
    store["upfirdn2d_hot/x_seed"] = np.array([0])
    store["upfirdn2d_hot/x"] = x.numpy().astype(np.float32)
    summarize(store, "upfirdn2d_hot/y", y)

    names = []
    for name, shape, slope, scale in [("nchw", (2, 5, 7, 9), 0.2, 2 ** 0.5), ("nc", (4, 16), 0.2, 2 ** 0.5),
                                       ("odd", (1, 3, 5, 5), 0.1, 1.0), ("big", (2, 33, 17, 31), 0.2, 2 ** 0.5)]:
        x = torch.randn(*shape, generator=g)
        b = torch.randn(shape[1], generator=g)
        y = fa_mod.fused_leaky_relu(x, b, slope, scale)
        store["flrelu/%s/x" % name] = x.numpy()
        store["flrelu/%s/b" % name] = b.numpy()
        store["flrelu/%s/cfg" % name] = np.array([slope, scale], dtype=np.float64)
        store["flrelu/%s/y" % name] = y.numpy()
        # autograd of the fallback == the semantics the custom Functions implement
        xg = x.clone().requires_grad_(True)
        bg = b.clone().requires_grad_(True)
        go = torch.randn(*shape, generator=g)
        yy = fa_mod.fused_leaky_relu(xg, bg, slope, scale)
        yy.backward(go)
        store["flrelu/%s/go" % name] = go.numpy()
        store["flrelu/%s/gx" % name] = xg.grad.numpy()
        store["flrelu/%s/gb" % name] = bg.grad.numpy()
        names.append(name)
    store["flrelu/names"] = np.array(names)
    np.savez_compressed(os.path.join(OUT, "ops.npz"), **store)
    print("ops.npz", len(store))


# ----------------------------------------------------------------------------------------------
def randomize_(module, g):
    """Fill every parameter of a reference module with seeded non-degenerate values (buffers kept)."""
    with torch.no_grad():
        for name, p in module.named_parameters():
            if name.endswith("noise.weight"):
                p.copy_(torch.full(p.shape, 0.3))
            elif name.endswith("prelu.weight"):
                p.copy_(torch.full(p.shape, 0.3))
            elif "bias" in name.split(".")[-1]:
                p.copy_(0.2 * torch.randn(p.shape, generator=g))
            else:
                p.copy_(p.detach().std().clamp_min(1e-3) * torch.randn(p.shape, generator=g)
                        if p.numel() > 1 else torch.randn(p.shape, generator=g))


def gen_layers():
    RH.install()
    from models.networks import stylegan2_layers as L
    from models.networks import generator as GM
    store = {}
    g = torch.Generator().manual_seed(777)
    torch.manual_seed(777)          # module constructors draw their initial weights from the global RNG (randomize_ scales
                                    # by the initial std): seeded, so that regenerating the fixture is bit-reproducible

    def put(name, module, inputs, out):
        for k, v in module.state_dict().items():
            store["%s/sd/%s" % (name, k)] = v.numpy()
        for i, t in enumerate(inputs):
            store["%s/in%d" % (name, i)] = t.numpy()
        store["%s/out" % name] = out.detach().numpy()

    # ConvLayer variants (stylegan2_layers.py:497-555)
    for name, kw, shape in [
        ("convlayer_1x1_act", dict(in_channel=3, out_channel=8, kernel_size=1), (2, 3, 12, 12)),
        ("convlayer_3x3_reflect_in", dict(in_channel=6, out_channel=6, kernel_size=3, reflection_pad=True, norm="in"),
         (2, 6, 12, 10)),
        ("convlayer_down_reflect", dict(in_channel=6, out_channel=8, kernel_size=3, downsample=True,
                                        blur_kernel=[1, 2, 1], reflection_pad=True), (1, 6, 16, 16)),
        ("convlayer_skip_down", dict(in_channel=6, out_channel=8, kernel_size=1, downsample=True,
                                     blur_kernel=[1, 2, 1], activate=False, bias=False, norm="in"), (1, 6, 16, 16)),
        ("convlayer_1x1_bias_in_noact", dict(in_channel=8, out_channel=8, kernel_size=1, activate=False, bias=True,
                                             norm="in"), (2, 8, 9, 9)),
        ("convlayer_1x1_skip_plain", dict(in_channel=8, out_channel=12, kernel_size=1, activate=False, bias=False),
         (1, 8, 8, 8)),
    ]:
        m = L.ConvLayer(**kw).eval()
        randomize_(m, g)
        x = torch.randn(*shape, generator=g)
        put(name, m, [x], m(x))
    # ResBlock (encoders)
    for name, norm in (("resblock_in", "in"), ("resblock_none", None)):
        m = L.ResBlock(6, 10, [1, 2, 1], reflection_pad=True, norm=norm).eval()
        randomize_(m, g)
        x = torch.randn(2, 6, 20, 16, generator=g)
        put(name, m, [x], m(x))
    # StyledConv: plain / nearest-upsample path / fused conv_transpose path (min(H,W)*2 >= 128)
    for name, ups, shape in (("styled_plain", False, (2, 6, 10, 12)), ("styled_up_small", True, (2, 6, 10, 12)),
                             ("styled_up_fused", True, (1, 4, 64, 66))):
        m = L.StyledConv(shape[1], 8, 3, 32, upsample=ups).eval()
        randomize_(m, g)
        x = torch.randn(*shape, generator=g)
        s = torch.randn(shape[0], 32, generator=g)
        f = 2 if ups else 1
        noise = torch.randn(shape[0], 1, shape[2] * f, shape[3] * f, generator=g)
        put(name, m, [x, s, noise], m(x, s, noise=noise))
    m = L.ToRGB(8, 32).eval()
    randomize_(m, g)
    x = torch.randn(2, 8, 9, 11, generator=g)
    s = torch.randn(2, 32, generator=g)
    put("torgb", m, [x, s], m(x, s, None))
    m = GM.ResidualBlock(8, 8).eval()
    randomize_(m, g)
    x = torch.randn(2, 8, 10, 10, generator=g)
    put("residual_block", m, [x], m(x.clone()))
    m = GM.GeneratorModulation(32, 8).eval()
    randomize_(m, g)
    x = torch.randn(2, 8, 6, 6, generator=g)
    s = torch.randn(2, 32, generator=g)
    put("generator_modulation", m, [x, s], m(x, s))
    opt = RH.default_opt()
    m = GM.ResolutionPreservingResnetBlock(opt, 6, 8, 32).eval()
    randomize_(m, g)
    x = torch.randn(2, 6, 8, 8, generator=g)
    s = torch.randn(2, 32, generator=g)
    n1, n2 = torch.randn(2, 1, 8, 8, generator=g), torch.randn(2, 1, 8, 8, generator=g)
    m.conv1.noise.fixed_noise, m.conv2.noise.fixed_noise = n1, n2
    put("respres_block", m, [x, s, n1, n2], m(x, s))
    for name, shape in (("upres_block_small", (1, 6, 8, 8)), ("upres_block_fused", (1, 6, 64, 64))):
        m = GM.UpsamplingResnetBlock(6, 4, 32, [1, 3, 3, 1], True).eval()
        randomize_(m, g)
        x = torch.randn(*shape, generator=g)
        s = torch.randn(1, 32, generator=g)
        n1 = torch.randn(1, 1, shape[2] * 2, shape[3] * 2, generator=g)
        n2 = torch.randn(1, 1, shape[2] * 2, shape[3] * 2, generator=g)
        m.conv1.noise.fixed_noise, m.conv2.noise.fixed_noise = n1, n2
        put(name, m, [x, s, n1, n2], m(x, s))
    np.savez_compressed(os.path.join(OUT, "layers.npz"), **store)
    print("layers.npz", len(store))


# ----------------------------------------------------------------------------------------------
def gen_nets():
    E1, E2, G = RH.build_networks()
    sdE1, sdE2, sdG = O.make_networks(0)
    for net, sd, shapes in ((E1, sdE1, O.shapes_E1()), (E2, sdE2, O.shapes_E2()), (G, sdG, O.shapes_G())):
        ref_sd = net.state_dict()
        assert set(ref_sd) == set(sd), (set(ref_sd) ^ set(sd))
        for k in ref_sd:
            assert tuple(ref_sd[k].shape) == tuple(sd[k].shape), k
        net.load_state_dict(sd)
    rself, corrm, mwarp = RH.ppst_model_methods()
    store = {}
    c = O.make_portraits(1, 512, 512, seed=1234)
    s = O.make_portraits(1, 512, 512, seed=4321, smooth=True)
    noises = O.make_noises(1, 512, 512)
    for name, m in G.named_modules():
        if type(m).__name__ == "NoiseInjection":
            m.fixed_noise = noises[name.replace(".noise", "")]
    with torch.no_grad():
        # hooks: per-layer activations of the trunks (layer-level pins at full size)
        taps = {}
        hooks = []
        for net, tag, names in ((E1, "E1", ["FromRGB", "DownToSpatialCode.ResBlockDownBy1", "DownToSpatialCode.ResBlockDownBy2",
                                             "DownToSpatialCode.ResBlockDownBy4"]),
                                (G, "G", ["SpatialCodeModulation"] + ["HeadResnetBlock%d" % i for i in range(4)]
                                 + ["UpsamplingResBlock%d" % r for r in (16, 32, 64)])):
            mods = dict(net.named_modules())
            for n in names:
                hooks.append(mods[n].register_forward_hook(
                    lambda mod, i, o, key="%s/%s" % (tag, n): taps.__setitem__(key, o)))
        sp_c = E1(c)
        summarize(store, "E1/sp_c", sp_c)
        gl_c, _ = E2(c)
        for i, v in enumerate(gl_c):
            store["E2/gl_c/%d" % i] = v.numpy()
        rgb, fea_c, fea_c1 = G(sp_c, gl_c, extract_features=True)
        summarize(store, "G/rec_c", rgb)
        summarize(store, "G/fea_c", fea_c)
        summarize(store, "G/fea_c1", fea_c1)
        for k, v in taps.items():
            summarize(store, "tap/" + k, v)
        for h in hooks:
            h.remove()
        rgb_plain = G(sp_c, gl_c)
        assert torch.equal(rgb_plain, rgb)
        sp_s = E1(s)
        gl_s, _ = E2(s)
        _, fea_s, fea_s1 = G(sp_s, gl_s, extract_features=True)
        rs_c, rs_s = rself(fea_c1), rself(fea_s1)
        summarize(store, "Rselfcorr/c", rs_c)
        fc = torch.cat((fea_c, rs_c), dim=1)
        fs = torch.cat((fea_s, rs_s), dim=1)
        corr = corrm(fs, fc)
        summarize(store, "corrm/corr", corr)
        store["corrm/rowsum_err"] = np.array([(corr.sum(-1) - 1).abs().max().item()])
        store["corrm/mean_rowmax"] = np.array([corr.amax(-1).mean().item()])
        gl_s2, gl_w = E2(s, corrmatrix=corr)
        for i, v in enumerate(gl_w):
            store["E2/gl_w/%d" % i] = v.numpy()
            store["E2/gl_s/%d" % i] = gl_s2[i].numpy()
        swap = G(sp_c, gl_w)
        summarize(store, "G/swap", swap)
        half = G(sp_c, O.lerp(gl_c, gl_w, 0.5))
        summarize(store, "G/swap_alpha0.5", half)
        # PPSTModel.warp on a feature map (both branches)
        summarize(store, "warp/full", mwarp(sp_s, corr))
        summarize(store, "warp/unfold", mwarp(torch.nn.functional.avg_pool2d(s, 2), corr))
        # E2 mask branch (training path) on batch 2 with integer label maps -> one-hot
        x2 = torch.cat([c, s], 0)
        lab = O.make_label_map(2, 512, 512)
        onehot = torch.stack([torch.tensor(O.mask_labels(lab[b].numpy()), dtype=torch.float) for b in range(2)])
        corr2 = torch.cat([corr, corrm(fc, fs)], 0)
        v, pm, vw, pmw = E2(x2, mask=onehot, corrmatrix=corr2)
        for i in range(4):
            store["E2mask/v/%d" % i] = v[i].numpy()
            store["E2mask/vw/%d" % i] = vw[i].numpy()
        for i in range(12):
            store["E2mask/pm/%d" % i] = pm[i].numpy()
            store["E2mask/pmw/%d" % i] = pmw[i].numpy()
        store["labels/sum"] = np.array([int(lab.to(torch.int64).sum())])
    # plain (non-correlation) path at a non-512 size: 256x384
    with torch.no_grad():
        for m in G.modules():
            if type(m).__name__ == "NoiseInjection":
                m.fixed_noise = None
        c2 = O.make_portraits(1, 256, 384, seed=5)
        n2 = O.make_noises(1, 256, 384, seed=6)
        for name, m in G.named_modules():
            if type(m).__name__ == "NoiseInjection":
                m.fixed_noise = n2[name.replace(".noise", "")]
        sp2 = E1(c2)
        gl2, _ = E2(c2)
        summarize(store, "plain256x384/sp", sp2)
        summarize(store, "plain256x384/rec", G(sp2, gl2))
    np.savez_compressed(os.path.join(OUT, "nets.npz"), **store)
    print("nets.npz", len(store))


def gen_function_py():
    RH.install()
    import importlib
    fn = importlib.import_module("models.networks.function")
    g = torch.Generator().manual_seed(31)
    x = torch.randn(2, 5, 9, 7, generator=g) * 2 + 0.5
    store = {"x": x.numpy()}
    m, s = fn.calc_mean_std(x)
    store["mean"], store["std"] = m.numpy(), s.numpy()
    store["nor_mean_std"] = fn.nor_mean_std(x).numpy()
    store["nor_mean"] = fn.nor_mean(x)[0].numpy()
    store["cov"] = fn.calc_cov(x).numpy()
    np.savez_compressed(os.path.join(OUT, "function_py.npz"), **store)
    print("function_py.npz")


def gen_discriminator():
    """The reference Discriminator (stylegan2_layers.py:582-649) at 256x256 (the stage-1 training size, first layer 3 -> 128)
    with oracle.make_discriminator weights: output,
    the stage-1 discriminator losses (ppst_model.py:68-92) and the R1 penalty (:139-156) with their parameter gradients
    -- first and second order through the reference's own (fallback) ops and torch.autograd."""
    RH.install()
    from models.networks import stylegan2_layers as L
    from models.networks import loss as RL
    size = 256
    sd = O.make_discriminator(size, seed=0, prefix="")
    D = L.Discriminator(size, 2.0, blur_kernel=[1, 3, 3, 1])
    missing = D.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(99)
    real = torch.rand(2, 3, size, size, generator=g) * 2 - 1
    fake = torch.rand(2, 3, size, size, generator=g) * 2 - 1
    store = {"real": real.numpy(), "fake": fake.numpy()}
    store["pred_real"] = D(real).detach().numpy()
    keys = ["convs.0.Conv.weight", "convs.0.Act.bias", "convs.1.conv1.Conv.weight", "convs.1.conv2.Conv.weight",
            "convs.1.conv2.Act.bias", "convs.1.skip.Conv.weight", "convs.3.conv1.Conv.weight", "convs.6.conv2.Conv.weight",
            "final_conv.Conv.weight",
            "final_linear.0.weight", "final_linear.0.bias", "final_linear.1.weight"]
    params = dict(D.named_parameters())
    # D step
    D.zero_grad()
    l_real = RL.gan_loss(D(real), True)
    l_rec = RL.gan_loss(D(fake), False) * 0.5
    (l_real + l_rec).backward()
    store["D_real"], store["D_rec"] = np.float64(l_real.item()), np.float64(l_rec.item())
    for k in keys:
        summarize(store, "dstep_grad/" + k, params[k].grad)
    # R1 step (ppst_model.py:139-156, lambda_R1 = 10)
    D.zero_grad()
    r = real.clone().requires_grad_(True)
    pred = D(r).sum()
    grad_real, = torch.autograd.grad(outputs=pred, inputs=[r], create_graph=True, retain_graph=True)
    pen = grad_real.pow(2).sum(dim=(1, 2, 3)) * 5.0
    summarize(store, "r1/grad_real", grad_real)
    store["r1/penalty"] = pen.detach().numpy().astype(np.float64)
    pen.mean().backward()
    for k in keys:
        summarize(store, "r1_grad/" + k, params[k].grad)
    np.savez_compressed(os.path.join(OUT, "discriminator.npz"), **store)
    print("discriminator.npz", store["D_real"], store["D_rec"], store["r1/penalty"])


def gen_rscl():
    """networks/rscl.py (aliased as models.networks.rscl) on deterministic queues / features (oracle.rscl_oracle.make_*):
    loss values with and without feat_k0 on every level, before and after the reference's enqueue sequence
    (ppst_model.py:214-219 pattern: six single-row calls per level), pointers, and a checksum of each queue."""
    RH.install()
    import types
    from oracle import rscl_oracle as RO
    rscl = sys.modules["models.networks.rscl"]
    crit = rscl.rsclLoss(types.SimpleNamespace(nce_T=0.07, num_gpus=1))
    store = {}
    for layer in range(4):
        getattr(crit, "queue_data_A%d" % layer).copy_(RO.make_queue(layer))
    for layer in range(4):
        q, k, k0 = RO.make_feats(3 * layer), RO.make_feats(3 * layer + 1), RO.make_feats(3 * layer + 2)
        store["loss/%d" % layer] = np.float64(crit(q, k, None, layer).item())
        store["loss_k0/%d" % layer] = np.float64(crit(q, k, k0, layer).item())
        for j in range(3):
            crit.dequeue_and_enqueue(k[j:j + 1], layer)
        for j in range(3):
            crit.dequeue_and_enqueue(k0[j:j + 1], layer)
        store["loss_after/%d" % layer] = np.float64(crit(q, k, k0, layer).item())
    for i in range(4):
        qd = getattr(crit, "queue_data_A%d" % i)
        store["queue_head/%d" % i] = qd[:, :8].numpy().copy()                 # the written columns (+2 untouched)
        store["queue_sum/%d" % i] = np.float64(qd.double().sum().item())
        store["ptr/%d" % i] = np.int64(int(getattr(crit, "queue_ptr_A%d" % i)))
    np.savez_compressed(os.path.join(OUT, "rscl.npz"), **store)
    print("rscl.npz", [float(store["loss/%d" % i]) for i in range(4)])


def gen_stage2():
    """Training stage 2 (the reference's default, ppst_model.py:24): PPSTModel.compute_generator_losses /
    compute_discriminator_losses (ppst_model.py:105-137,158-236) of the UNMODIFIED reference on CPU at 512^2, batch 2 (the
    only size its correlation path accepts), with lpips stubbed by mean |a - b| (oracle/ref_harness.py): every loss term, the
    L2 norm of every parameter gradient of the G step (E1 / E2 / G) and of the D step, subsamples of a few of them, and the
    region-contrast queues after the step."""
    RH.install()
    from models.ppst_model import PPSTModel
    opt = RH.default_opt(isTrain=True, continue_train=False, num_gpus=0, lambda_R1=10.0, lambda_L1=3.0, lambda_GAN=1.0,
                         training_stage=2, lambda_StyleCon=1.0, lambda_Maskwarp=10.0, lambda_Cycwarp=5.0, match_kernel=1,
                         nce_T=0.07, netE1="StyleGAN2Resnet", netE2="StyleGAN2Resnet", netG="StyleGAN2Resnet",
                         netD="StyleGAN2", crop_size=512, netD_scale_capacity=1.0, checkpoints_dir="/tmp", name="x")
    m = PPSTModel(opt)
    m.initialize()
    sdE1, sdE2, sdG, sdD, real, mask, queues = O.stage2_inputs()
    m.E1.load_state_dict(sdE1)
    m.E2.load_state_dict(sdE2)
    m.G.load_state_dict(sdG)
    m.D.stylegan2_D.load_state_dict(sdD)
    for i, q in enumerate(queues):
        getattr(m.criterionNCE, "queue_data_A%d" % i).copy_(q)
    store = {}
    # ---- generator step
    for p_ in m.D.parameters():
        p_.requires_grad_(False)
    losses, metrics = m.compute_generator_losses(real, None, None, mask)
    total = sum(v.mean() for v in losses.values())
    total.backward()
    for k, v in losses.items():
        store["G/" + k] = np.float64(v.mean().item())
    store["G/total"] = np.float64(total.item())
    for tag, net in (("E1", m.E1), ("E2", m.E2), ("G", m.G)):
        names, norms = [], []
        for k, p_ in net.named_parameters():
            names.append(k)
            norms.append(0.0 if p_.grad is None else float(p_.grad.double().norm()))
        store["Ggrad/%s/names" % tag] = np.array(names)
        store["Ggrad/%s/norms" % tag] = np.array(norms, dtype=np.float64)
    pick = {"E1": ["FromRGB.Conv.weight", "DownToSpatialCode.ResBlockDownBy4.conv2.Conv.weight", "ToSpatialCode.1.Conv.weight"],
            "E2": ["FromRGB.Conv.weight", "DownToGlobalCode1.ResBlockDownBy1.conv1.Conv.weight", "conv1x1_9.weight",
                   "projector2.5.weight"],
            "G": ["HeadResnetBlock0.conv1.conv.weight", "UpsamplingResBlock64.conv2.conv.weight", "layer32.2.weight",
                  "layer256.6.weight", "layert.2.conv2.weight", "layert1.1.weight", "ToRGB.conv.weight",
                  "UpsamplingResBlock16.conv1.epi1.style_mod.lin.weight"]}
    for tag, net in (("E1", m.E1), ("E2", m.E2), ("G", m.G)):
        params = dict(net.named_parameters())
        for k in pick[tag]:
            summarize(store, "Ggrad/%s/sample/%s" % (tag, k), params[k].grad)
    for i in range(4):
        qd = getattr(m.criterionNCE, "queue_data_A%d" % i)
        store["queue_head/%d" % i] = qd[:, :8].numpy().copy()
        store["queue_ptr/%d" % i] = np.int64(int(getattr(m.criterionNCE, "queue_ptr_A%d" % i)))
    # ---- discriminator step (fresh queues do not matter here)
    for net in (m.E1, m.E2, m.G):
        net.zero_grad()
        for p_ in net.parameters():
            p_.requires_grad_(False)
    for p_ in m.D.parameters():
        p_.requires_grad_(True)
    d_losses, _, _, _ = m.compute_discriminator_losses(real, mask)
    d_total = sum(v.mean() for v in d_losses.values())
    d_total.backward()
    for k, v in d_losses.items():
        store["D/" + k] = np.float64(v.mean().item())
    names, norms = [], []
    for k, p_ in m.D.stylegan2_D.named_parameters():
        names.append(k)
        norms.append(0.0 if p_.grad is None else float(p_.grad.double().norm()))
    store["Dgrad/names"] = np.array(names)
    store["Dgrad/norms"] = np.array(norms, dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "stage2.npz"), **store)
    print("stage2.npz", {k: float(v) for k, v in store.items() if k.startswith(("G/", "D/"))})


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    which = sys.argv[1:] or ["ops", "layers", "function", "nets", "discriminator", "rscl"]
    if "rscl" in which:
        gen_rscl()
    if "stage2" in which:
        gen_stage2()
    if "discriminator" in which:
        gen_discriminator()
    if "ops" in which:
        gen_ops()
    if "layers" in which:
        gen_layers()
    if "function" in which:
        gen_function_py()
    if "nets" in which:
        gen_nets()
