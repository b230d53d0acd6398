"""Training stage 2 (the reference's default: correlation-warped mix, cycle / mask warp, region style contrast;
models/ppst_model.py:105-236) at its only runnable size, 512^2 with 2 images per GPU: device time of the generator step
(losses + backward + Adam), the discriminator step and the R1 step, eager launches, CUDA events, with the SM clocks of the run.
    python tools/train2_bench.py [iters] > profiles/rNN_train_stage2.json"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from bench import ClockSampler  # noqa: E402
from ppst_b200 import _lib, ops, synthetic as S, train  # noqa: E402
from ppst_b200.train_stage2 import Stage2  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 5
torch.cuda.set_device(0)
sdE1, sdE2, sdG = S.make_networks(0)
for k in sdG:
    if k.endswith("noise.weight"):
        sdG[k] = torch.zeros_like(sdG[k])
D = train.Discriminator(512, 2.0).cuda()
D.load_reference_state_dict(S.make_discriminator(512, seed=0))
m = Stage2(sdE1, sdE2, sdG, D, noise_seed=1).cuda()
real = S.make_portraits(2, 512, 512, seed=77, smooth=True).cuda()
mask = ops.labels_to_onehot(S.make_label_map(2, 512, 512, seed=78).cuda())
og = torch.optim.Adam(m.g_params(), lr=2e-3, betas=(0.0, 0.99), fused=True)
od = torch.optim.Adam(list(D.parameters()), lr=2e-3, betas=(0.0, 0.99), fused=True)


def g_step():
    for p in D.parameters():
        p.requires_grad_(False)
    og.zero_grad(set_to_none=True)
    losses = m.generator_losses(real, mask)
    sum(v.mean() for v in losses.values()).backward()
    og.step()
    return losses


def d_step():
    for p in D.parameters():
        p.requires_grad_(True)
    od.zero_grad(set_to_none=True)
    losses = m.discriminator_losses(real, mask)
    sum(v.mean() for v in losses.values()).backward()
    od.step()
    return losses


def r1_step():
    od.zero_grad(set_to_none=True)
    (train.r1_loss(D, real).mean() * 16).backward()
    od.step()


def timed(fn, n):
    fn()
    torch.cuda.synchronize()
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (_lib.launch_count() - l0) // n, out


sampler = ClockSampler(0)
sampler.start()
g_ms, g_l, g_losses = timed(g_step, iters)
d_ms, d_l, d_losses = timed(d_step, iters)
r_ms, r_l, _ = timed(r1_step, iters)
clocks = sampler.stop()
it_ms = g_ms + d_ms + r_ms / 16
print(json.dumps({
    "what": "training stage 2 @512^2, 2 images per GPU, N=1, eager launches (the queue pointers of the region-contrast loss "
            "are read on the host every step, so the step is not graph-captured)",
    "generator_step_ms": g_ms, "discriminator_step_ms": d_ms, "r1_step_ms": r_ms, "iteration_ms": it_ms,
    "images_per_s": 2.0 / (it_ms * 1e-3), "our_kernel_launches": {"g": g_l, "d": d_l, "r1": r_l},
    "reference_cpu_seconds_per_step": {"source": "BASELINE.md section 2 (8-core Xeon)", "D": 82, "R1": 21, "G": 82},
    "losses_last_g_step": {k: float(v) for k, v in g_losses.items()},
    "losses_last_d_step": {k: float(v) for k, v in d_losses.items()}, "clocks": clocks,
    "peak_memory_gb": torch.cuda.max_memory_allocated() / 2 ** 30}, indent=1))
