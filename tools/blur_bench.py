"""blur_nhwc (training path) alone at the largest stage-1 shape, L2-cold (inputs rotate), CUDA events.
    python tools/blur_bench.py"""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from ppst_b200 import train  # noqa: E402

PEAK = 6549.8
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
k1 = torch.tensor([1., 3., 3., 1.], device="cuda")
k = (k1[None] * k1[:, None] / 64).contiguous()
for (n, h, c) in ((8, 256, 128), (8, 512, 32), (8, 128, 256)):
    xs = [torch.randn(n, h, h, c, device="cuda").to(torch.bfloat16) for _ in range(4)]
    for x in xs:
        y = train.blur(x, k, (2, 2))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for i in range(reps):
        train.blur(xs[i % 4], k, (2, 2))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    byt = (xs[0].numel() + y.numel()) * 2
    print(json.dumps({"op": "blur_nhwc", "shape": [n, h, h, c], "ms": ms, "gbs": byt / ms / 1e6,
                      "frac_of_measured_hbm": byt / ms / 1e6 / PEAK}))
