"""ONE eager swap step (8 pairs @512^2) between cudaProfilerStart/Stop, for the ncu launch list:
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python tools/one_step.py
and, with -k regex:<kernel> --set full, for single-kernel captures."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from ppst_b200 import synthetic as O  # noqa: E402  (seeded weights / inputs)
import ppst_b200  # noqa: E402

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 8
m = ppst_b200.SwapModel(*O.make_networks(0), mode="bf16")
m.noises = {k: v.expand(pairs, -1, -1, -1).contiguous() for k, v in O.make_noises(1, 512, 512).items()}
c = O.make_portraits(pairs, 512, 512, seed=1).cuda()
s = O.make_portraits(pairs, 512, 512, seed=2, smooth=True).cuda()
for _ in range(2):
    m.swap(c, s)
torch.cuda.synchronize()
torch.cuda.profiler.start()
out = m.swap(c, s)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(tuple(out.shape), float(out.float().abs().mean()))
