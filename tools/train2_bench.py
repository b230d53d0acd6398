"""Training stage 2 (the reference's default: correlation-warped mix, cycle / mask warp, region style contrast;
models/ppst_model.py:105-236) at its only runnable size, 512^2 with 2 images per GPU: device time of the generator step
(losses + backward + gradient all-reduce + Adam), the discriminator step and the R1 step -- eager launches and replayed from
CUDA graphs (the 24 region-contrast keys of a step leave the graph as static tensors and are pushed into the queues by one
eager all_gather after the replay: the queue pointers live on the host) -- with the SM clocks of the run.  Weak scaling:
    python tools/train2_bench.py [iters]
    python -m torch.distributed.run --nproc-per-node N tools/train2_bench.py [iters]   > profiles/rNN_train_stage2_nN.json"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from bench import ClockSampler  # noqa: E402
from ppst_b200 import _lib, ops, sharding, synthetic as S, train  # noqa: E402
from ppst_b200.train_stage2 import Stage2  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 5
world, rank, local_rank = sharding.init()
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
sdE1, sdE2, sdG = S.make_networks(0)
for k in sdG:
    if k.endswith("noise.weight"):
        sdG[k] = torch.zeros_like(sdG[k])
D = train.Discriminator(512, 2.0).cuda()
D.load_reference_state_dict(S.make_discriminator(512, seed=0))
m = Stage2(sdE1, sdE2, sdG, D, noise_seed=1 + rank).cuda()
real = S.make_portraits(2, 512, 512, seed=77 + rank, smooth=True).cuda()
mask = ops.labels_to_onehot(S.make_label_map(2, 512, 512, seed=78 + rank).cuda())
g_params, d_params = m.g_params(), list(D.parameters())
og = torch.optim.Adam(g_params, lr=2e-3, betas=(0.0, 0.99), fused=True, capturable=True)
od = torch.optim.Adam(d_params, lr=2e-3, betas=(0.0, 0.99), fused=True, capturable=True)
loss_buf = torch.zeros(1, device=dev)
bucket = {}
# gradient exchange bucketed and overlapped with the backward pass (PPST_TRAIN_OVERLAP=0: one flat all-reduce after it)
overlap = os.environ.get("PPST_TRAIN_OVERLAP", "1") != "0"
red_g = sharding.OverlappedGradReducer(g_params, world) if overlap else None
red_d = sharding.OverlappedGradReducer(d_params, world) if overlap else None


def backward_and_exchange(loss, params, red, key):
    if red is not None and red.active:
        with red:
            loss.backward()
        bucket[key] = red.bytes_reduced
    else:
        loss.backward()
        bucket[key] = sharding.allreduce_grads_(params, world)


def body_g():
    losses = m.generator_losses(real, mask, enqueue=False)
    total = sum(v.mean() for v in losses.values())
    backward_and_exchange(total, g_params, red_g, "g")
    loss_buf.copy_(total.detach().reshape(1))
    og.step()


def body_d():
    backward_and_exchange(sum(v.mean() for v in m.discriminator_losses(real, mask).values()), d_params, red_d, "d")
    od.step()


def body_r1():
    (train.r1_loss(D, real).mean() * 16).backward()
    sharding.allreduce_grads_(d_params, world)
    od.step()


def zero():
    og.zero_grad(set_to_none=True)
    od.zero_grad(set_to_none=True)


def set_d_grad(flag):
    for p in d_params:
        p.requires_grad_(flag)


def timed(fn, n):
    fn()
    torch.cuda.synchronize()
    sharding.barrier()
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return sharding.max_over_ranks(e0.elapsed_time(e1) / n, device=dev), (_lib.launch_count() - l0) // n


def eager(body, d_grad, push=False):
    def f():
        set_d_grad(d_grad)
        zero()
        body()
        if push:
            m.criterionNCE.dequeue_and_enqueue_many(m.last_pushes)
    return f


sampler = ClockSampler(local_rank)
sampler.start()
res = {}
for name, body, d_grad in (("generator_step", body_g, False), ("discriminator_step", body_d, True), ("r1_step", body_r1, True)):
    res[name + "_eager_ms"], res[name + "_launches"] = timed(eager(body, d_grad, push=(name == "generator_step")), iters)
graph_err = None
try:
    graphs = {}
    for name, body, d_grad in (("generator_step", body_g, False), ("discriminator_step", body_d, True), ("r1_step", body_r1, True)):
        set_d_grad(d_grad)
        graphs[name] = train.capture_step(body, zero, generators=(m._noise(dev),))
    set_d_grad(True)

    def replay_g():
        graphs["generator_step"].replay()
        m.criterionNCE.dequeue_and_enqueue_many(m.last_pushes)       # eager: 1 all_gather + 24 column writes
    res["generator_step_ms"], _ = timed(replay_g, iters)
    res["discriminator_step_ms"], _ = timed(graphs["discriminator_step"].replay, iters)
    res["r1_step_ms"], _ = timed(graphs["r1_step"].replay, iters)
except Exception as e:                                               # keep the eager numbers
    graph_err = str(e).split("\n")[0][:300]
    for k in ("generator_step", "discriminator_step", "r1_step"):
        res[k + "_ms"] = res[k + "_eager_ms"]
if "--trace" in sys.argv and graph_err is None and rank == 0:
    # kernel activity of ONE replayed generator step, by kernel name (stderr)
    from collections import defaultdict
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        graphs["generator_step"].replay()
        torch.cuda.synchronize()
    busy = defaultdict(lambda: [0.0, 0])
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            name = e.name.split("(")[0].replace("ppst::", "").replace("void ", "")[:90]
            busy[name][0] += e.time_range.end - e.time_range.start
            busy[name][1] += 1
    tot = sum(v[0] for v in busy.values())
    sys.stderr.write("generator step: %d kernels, %.2f ms of kernel time\n" % (sum(v[1] for v in busy.values()), tot / 1e3))
    for name, (d, k) in sorted(busy.items(), key=lambda kv: -kv[1][0])[:45]:
        sys.stderr.write("%9.1f us %5.1f%%  x%-4d %s\n" % (d, 100 * d / tot, k, name))
clocks = sampler.stop()
it_ms = res["generator_step_ms"] + res["discriminator_step_ms"] + res["r1_step_ms"] / 16
if rank == 0:
    print(json.dumps(dict(res, **{
        "what": "training stage 2 @512^2, 2 images per GPU, %d GPU(s), weak scaling; *_ms = CUDA-graph replay (max over ranks), "
                "*_eager_ms = eager launches" % world,
        "n_gpus": world, "iteration_ms": it_ms, "images_per_s": 2.0 * world / (it_ms * 1e-3), "graph_capture_error": graph_err,
        "grad_allreduce_bytes": bucket, "loss_last_g_step": float(loss_buf),
        "reference_cpu_seconds_per_step": {"source": "BASELINE.md section 2 (8-core Xeon)", "D": 82, "R1": 21, "G": 82},
        "clocks": clocks, "peak_memory_gb": torch.cuda.max_memory_allocated() / 2 ** 30}), indent=1))
sys.stdout.flush()
if torch.distributed.is_initialized():
    # the N = 8 run printed its line and then sat in process-group teardown until the caller's timeout (graphs holding NCCL
    # work were still alive): synchronise, then leave without tearing the communicator down
    torch.cuda.synchronize()
    sharding.barrier()
    os._exit(0)
