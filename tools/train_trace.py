"""CUPTI kernel timeline of one discriminator D step + R1 step, or (third argument `stage1`) of the stage-1 generator
step (which kernels the step's GPU time goes to)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from oracle import ppst_oracle as O
from ppst_b200 import train

size, batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256, int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = torch.device("cuda", 0)
D = train.Discriminator(size, 2.0).to(dev)
D.load_reference_state_dict(O.make_discriminator(size, seed=0))
opt = torch.optim.Adam(D.parameters(), lr=2e-3, betas=(0.0, 0.99), fused=True)
g = torch.Generator().manual_seed(5)
real = (torch.rand(batch, 3, size, size, generator=g) * 2 - 1).to(dev)
fake = (torch.rand(batch // 2, 3, size, size, generator=g) * 2 - 1).to(dev)


def step(r1):
    opt.zero_grad(set_to_none=True)
    (train.r1_loss(D, real).mean() if r1 else sum(train.d_losses(D, real, fake).values())).backward()
    opt.step()


if len(sys.argv) > 3 and sys.argv[3] == "stage1":
    from ppst_b200 import train_nets as TN
    sdE1, sdE2, sdG = O.make_networks(0)
    for k in sdG:
        if k.endswith("noise.weight"):
            sdG[k] = torch.zeros_like(sdG[k])
    m = TN.Stage1(sdE1, sdE2, sdG, D, noise_seed=1).to(dev)
    gp = m.g_params()
    gopt = torch.optim.Adam(gp, lr=2e-3, betas=(0.0, 0.99), fused=True)
    for p in D.parameters():
        p.requires_grad_(False)

    def step(_):
        gopt.zero_grad(set_to_none=True)
        losses, _r = m.generator_losses(real)
        sum(losses.values()).backward()
        gopt.step()

for r1 in ((False,) if len(sys.argv) > 3 else (False, True)):
    step(r1); step(r1)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step(r1)
        torch.cuda.synchronize()
    rows = {}
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            r = rows.setdefault(e.name[:90], [0.0, 0])
            r[0] += e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total
            r[1] += 1
    tot = sum(v[0] for v in rows.values())
    print("== %s step: %d kernels, %.2f ms of kernel time" % ("R1" if r1 else "D", sum(v[1] for v in rows.values()), tot / 1e3))
    for name, (us, cnt) in sorted(rows.items(), key=lambda kv: -kv[1][0])[:22]:
        print("  %9.1f us %5.1f%%  x%-4d %s" % (us, 100 * us / tot, cnt, name))
