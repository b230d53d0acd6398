"""Discriminator training step (SURVEY.md 8(f)4 / BASELINE config 5, the part that is built): D losses forward + backward,
R1 penalty (double backward), Adam; ours (tcgen05 fwd / dgrad / wgrad, bf16) next to the oracle's torch expression of the
same network run on the same GPU (cuDNN, fp32 / TF32) as context.  Prints one JSON line per size.

    python tools/train_bench.py [--size 256] [--batch 8] [--steps 5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/train_bench.py ...
Under torchrun every rank trains on its own batch (weak scaling) and the gradients go through ONE flat NCCL all-reduce per
step (sharding.allreduce_grads_); times are device times, max over ranks.
"""
import argparse, json, math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":     # NCCL prints its version banner to stdout at VERSION / WARN
    del os.environ["NCCL_DEBUG"]


def timed(fn, steps, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def stage1(args):
    """Training stage 1 (ppst_model.py:113-137,139-156,158-236 with lambda_StyleCon = 0): generator step = E1, E2, G forward
    + backward with the L1 and GAN losses (D forward + input gradient), discriminator step (reconstructions without
    grad, D forward + backward), R1 step; Adam on both parameter sets; one flat gradient all-reduce per step."""
    from oracle import ppst_oracle as O
    from ppst_b200 import train, train_nets as TN, ops, _lib, sharding
    world, rank, local_rank = sharding.init()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    sdE1, sdE2, sdG = O.make_networks(0)
    for k in sdG:
        if k.endswith("noise.weight"):
            sdG[k] = torch.zeros_like(sdG[k])                     # the reference's initial value
    D = train.Discriminator(args.size, 2.0)
    D.load_reference_state_dict(O.make_discriminator(args.size, seed=0))
    m = TN.Stage1(sdE1, sdE2, sdG, D, noise_seed=1234 + rank).to(dev)
    g_params, d_params = m.g_params(), list(D.parameters())
    opt_g = torch.optim.Adam(g_params, lr=2e-3, betas=(0.0, 0.99), fused=True)
    opt_d = torch.optim.Adam(d_params, lr=2e-3, betas=(0.0, 0.99), fused=True)
    real = O.make_portraits(args.batch, args.size, args.size, seed=77 + rank, smooth=True).to(dev)
    bucket = {}
    # gradient exchange: bucketed and overlapped with the rest of the backward pass (sharding.OverlappedGradReducer);
    # PPST_TRAIN_OVERLAP=0: one flat all-reduce after the backward pass (round 1)
    overlap = os.environ.get("PPST_TRAIN_OVERLAP", "1") != "0"
    red_g = sharding.OverlappedGradReducer(g_params, world) if overlap else None
    red_d = sharding.OverlappedGradReducer(d_params, world) if overlap else None

    def backward_and_exchange(loss, params, red, key=None):
        if red is not None and red.active:
            with red:
                loss.backward()
            nbytes = red.bytes_reduced
        else:
            loss.backward()
            nbytes = sharding.allreduce_grads_(params, world)
        if key:
            bucket[key] = nbytes

    def g_step():
        opt_g.zero_grad(set_to_none=True)
        for p in d_params:
            p.requires_grad_(False)
        losses, _ = m.generator_losses(real)
        backward_and_exchange(sum(losses.values()), g_params, red_g, "g")
        for p in d_params:
            p.requires_grad_(True)
        opt_g.step()

    def d_step():
        opt_d.zero_grad(set_to_none=True)
        backward_and_exchange(sum(m.discriminator_losses(real).values()), d_params, red_d, "d")
        opt_d.step()

    def r1_step():
        opt_d.zero_grad(set_to_none=True)
        train.r1_loss(D, real).mean().backward()
        sharding.allreduce_grads_(d_params, world)
        opt_d.step()

    l0 = _lib.launch_count()
    g_step(); d_step(); r1_step()
    launches = _lib.launch_count() - l0
    sharding.barrier()
    ms = {k: sharding.max_over_ranks(timed(f, args.steps), device=dev) for k, f in (("g_step", g_step), ("d_step", d_step), ("r1_step", r1_step))}
    ops.PROFILE = {}
    g_step(); d_step()
    torch.cuda.synchronize()
    prof = ops.profile_summary()
    ops.PROFILE = None
    kern = {k: {"launches": v["launches"], "ms": round(v["ms"], 3),
                "tflops": round(v["work"] / (v["ms"] * 1e-3) / 1e12, 1) if k in ("conv_tc", "conv_wgrad") and v["ms"] > 0 else None,
                "gbs": round(v["work"] / (v["ms"] * 1e-3) / 1e9, 1) if k not in ("conv_tc", "conv_wgrad") and v["ms"] > 0 else None}
            for k, v in prof.items()}
    # the same three steps captured once into CUDA graphs (autograd tape, weight preparation, noise draws, all-reduce and
    # Adam included) and replayed: removes the host gaps between the ~2500 launches of a G step
    ms_graph = None
    try:
        cg = torch.optim.Adam(g_params, lr=2e-3, betas=(0.0, 0.99), fused=True, capturable=True)
        cd = torch.optim.Adam(d_params, lr=2e-3, betas=(0.0, 0.99), fused=True, capturable=True)

        loss_buf = torch.zeros(1, device=dev)

        def body_g():
            losses, _ = m.generator_losses(real)
            total = sum(losses.values())
            backward_and_exchange(total, g_params, red_g)
            loss_buf.copy_(total.detach().reshape(1))
            cg.step()

        def body_d():
            backward_and_exchange(sum(m.discriminator_losses(real).values()), d_params, red_d)
            cd.step()

        def body_r1():
            train.r1_loss(D, real).mean().backward()
            sharding.allreduce_grads_(d_params, world)
            cd.step()

        graphs = {}

        def zero():
            cg.zero_grad(set_to_none=True)
            cd.zero_grad(set_to_none=True)
        for name, body, d_grad in (("g_step", body_g, False), ("d_step", body_d, True), ("r1_step", body_r1, True)):
            for p in d_params:
                p.requires_grad_(d_grad)
            graphs[name] = train.capture_step(body, zero, generators=(m._noise(dev),))
        for p in d_params:
            p.requires_grad_(True)
        sharding.barrier()
        ms_graph = {k: sharding.max_over_ranks(timed(g.replay, args.steps), device=dev) for k, g in graphs.items()}
        # end to end: every iteration copies its batch from pinned host memory into the graphs' static input and reads the
        # generator loss back (R1 on every 16th iteration, as the reference's optimiser schedules it)
        real_pin = real.cpu().pin_memory()
        loss_pin = torch.zeros(1).pin_memory()
        it = [0]

        def iteration():
            real.copy_(real_pin, non_blocking=True)
            graphs["g_step"].replay()
            graphs["d_step"].replay()
            if it[0] % 16 == 0:
                graphs["r1_step"].replay()
            it[0] += 1
            loss_pin.copy_(loss_buf, non_blocking=True)
        ms_graph["iteration_e2e"] = sharding.max_over_ranks(timed(iteration, 16, warm=1), device=dev)
    except Exception as e:                                       # report the eager numbers only
        ms_graph = {"error": str(e).split("\n")[0][:200]}
        for p in d_params:
            p.requires_grad_(True)
    # one iteration of the reference's loop = G step + D step, R1 every 16th iteration (R1_once_every, ppst_optimizer.py)
    best = ms_graph if (ms_graph and "error" not in ms_graph) else ms
    it_ms = best["g_step"] + best["d_step"] + best["r1_step"] / 16.0
    e2e_ms = ms_graph.get("iteration_e2e") if isinstance(ms_graph.get("iteration_e2e"), float) else None
    out = {"workload": "training stage 1 @%d^2, batch %d per GPU: G step (E1+E2+G+D), D step, R1 every 16th" % (args.size, args.batch),
           "n_gpus": world, "scaling": "weak", "images_per_s": round(world * args.batch / (it_ms * 1e-3), 1),
           "ms": {k: round(v, 2) for k, v in ms.items()},
           "graph_ms": {k: (round(v, 2) if isinstance(v, float) else v) for k, v in ms_graph.items()},
           "iteration_ms": round(it_ms, 2), "iteration_e2e_ms": round(e2e_ms, 2) if e2e_ms else None,
           "e2e_images_per_s": round(world * args.batch / (e2e_ms * 1e-3), 1) if e2e_ms else None,
           "h2d_bytes_per_iteration": int(real.numel() * 4), "d2h_bytes_per_iteration": 4,
           "grad_allreduce_bytes": bucket, "launches_g_d_r1": launches, "kernels_g_plus_d": kern}
    if rank != 0:
        return None
    if not args.no_torch and world == 1:
        def leaf(sd):
            return {k: (v.to(dev).requires_grad_(True) if (v.is_floating_point() and not k.endswith("kernel")) else v.to(dev)) for k, v in sd.items()}
        rE1, rE2, rG, rD = leaf(sdE1), leaf(sdE2), leaf(sdG), leaf(O.make_discriminator(args.size, seed=0))
        rp = [v for d in (rE1, rE2, rG) for v in d.values() if v.requires_grad]
        topt = torch.optim.Adam(rp, lr=2e-3, betas=(0.0, 0.99), fused=True)
        tdopt = torch.optim.Adam([v for v in rD.values() if v.requires_grad], lr=2e-3, betas=(0.0, 0.99), fused=True)

        def t_rec(x):
            return O.generator(rG, O.encoder_con(rE1, x), O.encoder_col(rE2, x)[0])

        def t_g():
            topt.zero_grad(set_to_none=True)
            rec = t_rec(real)
            ((rec - real).abs().mean() + 0.5 * O.gan_loss(O.discriminator(rD, rec), True)).backward()
            topt.step()

        def t_d():
            tdopt.zero_grad(set_to_none=True)
            with torch.no_grad():
                rec = t_rec(real[: args.batch // 2])
            sum(O.d_losses(rD, real, rec).values()).backward()
            tdopt.step()
        out["torch_cudnn_tf32_ms"] = {"g_step": round(timed(t_g, max(2, args.steps // 2), warm=1), 2),
                                      "d_step": round(timed(t_d, max(2, args.steps // 2), warm=1), 2)}
    if getattr(args, "return_dict", False):
        return out
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--no-torch", action="store_true")
    ap.add_argument("--stage1", action="store_true", help="whole stage-1 iteration: G step (E1, E2, G, D), D step, R1 step")
    args = ap.parse_args()
    if args.stage1:
        return stage1(args)
    from oracle import ppst_oracle as O
    from ppst_b200 import train, ops, _lib, sharding
    world, rank, local_rank = sharding.init()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    sd = O.make_discriminator(args.size, seed=0)
    D = train.Discriminator(args.size, 2.0).to(dev)
    D.load_reference_state_dict(sd)
    opt = torch.optim.Adam(D.parameters(), lr=2e-3, betas=(0.0, 0.99), fused=True)
    g = torch.Generator().manual_seed(5 + rank)
    real = (torch.rand(args.batch, 3, args.size, args.size, generator=g) * 2 - 1).to(dev)
    fake = (torch.rand(args.batch // 2, 3, args.size, args.size, generator=g) * 2 - 1).to(dev)   # ppst_model.py:128-133

    params = list(D.parameters())
    bucket = [0]

    def d_step():
        opt.zero_grad(set_to_none=True)
        l = train.d_losses(D, real, fake)
        (l["D_real"] + l["D_rec"]).backward()
        bucket[0] = sharding.allreduce_grads_(params, world)
        opt.step()

    def r1_step():
        opt.zero_grad(set_to_none=True)
        train.r1_loss(D, real).mean().backward()
        sharding.allreduce_grads_(params, world)
        opt.step()

    l0 = _lib.launch_count()
    d_step(); r1_step()
    launches = _lib.launch_count() - l0
    sharding.barrier()
    ms_d, ms_r1 = timed(d_step, args.steps), timed(r1_step, args.steps)
    ms_d, ms_r1 = sharding.max_over_ranks(ms_d, device=dev), sharding.max_over_ranks(ms_r1, device=dev)
    # the same two steps captured once into CUDA graphs (autograd tape, weight packing and Adam included) and replayed
    gopt = torch.optim.Adam(D.parameters(), lr=2e-3, betas=(0.0, 0.99), fused=True, capturable=True)
    graphs = []
    for loss_fn in (lambda: sum(train.d_losses(D, real, fake).values()), lambda: train.r1_loss(D, real).mean()):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                gopt.zero_grad(set_to_none=True)
                loss_fn().backward()
                sharding.allreduce_grads_(params, world)
                gopt.step()
        torch.cuda.current_stream().wait_stream(side)
        gr = torch.cuda.CUDAGraph()
        gopt.zero_grad(set_to_none=True)
        with torch.cuda.graph(gr):
            loss = loss_fn()
            loss.backward()
            sharding.allreduce_grads_(params, world)
            gopt.step()
        graphs.append(gr)
    sharding.barrier()
    ms_dg, ms_r1g = timed(graphs[0].replay, args.steps), timed(graphs[1].replay, args.steps)
    ms_dg, ms_r1g = sharding.max_over_ranks(ms_dg, device=dev), sharding.max_over_ranks(ms_r1g, device=dev)
    ops.PROFILE = {}
    d_step(); r1_step()
    torch.cuda.synchronize()
    prof = ops.profile_summary()
    ops.PROFILE = None
    kern = {k: {"launches": v["launches"], "ms": round(v["ms"], 3),
                "tflops": round(v["work"] / (v["ms"] * 1e-3) / 1e12, 1) if k in ("conv_tc", "conv_wgrad") and v["ms"] > 0 else None}
            for k, v in prof.items()}
    imgs = world * (args.batch + args.batch // 2)
    out = {"workload": "discriminator training step @%d^2, batch %d real + %d fake per GPU" % (args.size, args.batch, args.batch // 2),
           "n_gpus": world, "scaling": "weak", "grad_allreduce_bytes": bucket[0],
           "d_step_images_per_s": round(imgs / (min(ms_d, ms_dg) * 1e-3), 1),
           "ours_ms": {"d_step": round(ms_d, 2), "r1_step": round(ms_r1, 2)},
           "ours_graph_ms": {"d_step": round(ms_dg, 2), "r1_step": round(ms_r1g, 2)}, "launches_d_plus_r1": launches, "kernels": kern}
    if rank != 0:
        return
    if not args.no_torch and world == 1:
        sdc = {k: (v.to(dev).requires_grad_(True) if "Blur" not in k else v.to(dev)) for k, v in sd.items()}
        params = [v for k, v in sdc.items() if v.requires_grad]
        topt = torch.optim.Adam(params, lr=2e-3, betas=(0.0, 0.99), fused=True)

        def t_d():
            topt.zero_grad(set_to_none=True)
            l = O.d_losses(sdc, real, fake)
            (l["D_real"] + l["D_rec"]).backward()
            topt.step()

        def t_r1():
            topt.zero_grad(set_to_none=True)
            O.r1_loss(sdc, real).mean().backward()
            topt.step()

        out["torch_cudnn_tf32_ms"] = {"d_step": round(timed(t_d, args.steps), 2), "r1_step": round(timed(t_r1, args.steps), 2)}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()
