"""Kernel timeline of ONE CUDA-graph replay of the swap step (torch.profiler / CUPTI activity records): per-kernel
busy time by name, and the idle gaps between consecutive kernels -- what the eager per-launch event profile
(tools/profile_step.py) cannot see.   python tools/trace_step.py [pairs] > profiles/rNN_trace_step.txt"""
import os
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402
from ppst_b200 import synthetic as O  # noqa: E402  (seeded weights / inputs)
import ppst_b200  # noqa: E402

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 8
m = ppst_b200.SwapModel(*O.make_networks(0), mode="bf16")
m.noises = {k: v.expand(pairs, -1, -1, -1).contiguous() for k, v in O.make_noises(1, 512, 512).items()}
c = O.make_portraits(pairs, 512, 512, seed=1).cuda()
s = O.make_portraits(pairs, 512, 512, seed=2, smooth=True).cuda()
for _ in range(4):
    m.swap_graph(c, s)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        m.swap_graph(c, s)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
# keep the last replay: split at the largest gaps is fragile; instead take the last third of the kernel list
n = len(evs) // 3
evs = evs[2 * n:]
t0, t1 = evs[0].time_range.start, evs[-1].time_range.end
busy = defaultdict(lambda: [0.0, 0])
gap_total, gaps = 0.0, []
prev_end = None
for e in evs:
    d = e.time_range.end - e.time_range.start
    name = e.name.split("(")[0].replace("ppst::", "").replace("void ", "")[:70]
    busy[name][0] += d
    busy[name][1] += 1
    if prev_end is not None:
        g = e.time_range.start - prev_end
        if g > 0:
            gap_total += g
            gaps.append((g, name))
    prev_end = max(prev_end or 0, e.time_range.end)
span = t1 - t0
tot = sum(v[0] for v in busy.values())
print("one graph replay: %d kernels, span %.3f ms, kernel busy %.3f ms, idle gaps %.3f ms (%.1f us / kernel)" %
      (len(evs), span / 1e3, tot / 1e3, gap_total / 1e3, gap_total / max(len(evs), 1)))
for name, (d, k) in sorted(busy.items(), key=lambda kv: -kv[1][0]):
    print("%9.1f us %5.1f%%  x%-3d %s" % (d, 100 * d / span, k, name))
print("largest gaps (us, before kernel):")
for g, name in sorted(gaps, reverse=True)[:12]:
    print("  %7.1f  %s" % (g, name))
