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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--no-torch", action="store_true")
    args = ap.parse_args()
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
