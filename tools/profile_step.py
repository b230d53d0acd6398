"""Per-(kernel, shape) timing of one eager swap step (CUDA events around every launch, host kept ahead of the
device).  Development aid; prints a table sorted by time.   python tools/profile_step.py [pairs]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from ppst_b200 import synthetic as O  # noqa: E402  (seeded weights / inputs)
import ppst_b200  # noqa: E402
from ppst_b200 import ops  # noqa: E402

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 8
m = ppst_b200.SwapModel(*O.make_networks(0), mode="bf16")
m.noises = {k: v.expand(pairs, -1, -1, -1).contiguous() for k, v in O.make_noises(1, 512, 512).items()}
c = O.make_portraits(pairs, 512, 512, seed=1).cuda()
s = O.make_portraits(pairs, 512, 512, seed=2, smooth=True).cuda()
for _ in range(2):
    m.swap(c, s)
torch.cuda.synchronize()
ops.PROFILE, ops.PROFILE_DETAIL = {}, True
m.swap(c, s)
torch.cuda.synchronize()
summ = ops.profile_summary()
ops.PROFILE = None
tot = sum(v["ms"] for v in summ.values())
print("total profiled ms %.3f" % tot)
for k, v in sorted(summ.items(), key=lambda kv: -kv[1]["ms"]):
    name = k.split("|")[0]
    rate = v["work"] / (v["ms"] * 1e-3) / (1e12 if name in ("conv_tc", "conv_simt", "gemm_f32") else 1e9)
    print("%7.3f ms %5.1f%% x%-3d %8.1f %s  %s" % (v["ms"], 100 * v["ms"] / tot, v["launches"], rate,
                                                 "TF/s" if name in ("conv_tc", "conv_simt", "gemm_f32") else "GB/s", k))
