"""Per-shape timing of ppst_conv_wgrad (graph-timed, like tools/conv_shapes.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ppst_b200 import train

SHAPES = [  # n, h, w, ci, co, k, s, p
    (8, 256, 256, 128, 128, 3, 1, 1), (8, 256, 256, 128, 256, 6, 2, 2), (8, 256, 256, 128, 256, 4, 2, 1),
    (8, 128, 128, 256, 256, 3, 1, 1), (8, 128, 128, 256, 512, 6, 2, 2), (8, 64, 64, 512, 512, 3, 1, 1),
    (8, 64, 64, 512, 512, 6, 2, 2), (8, 16, 16, 512, 512, 3, 1, 1), (8, 512, 512, 128, 128, 3, 1, 1),
    (8, 256, 256, 256, 256, 3, 1, 1), (8, 512, 512, 64, 64, 3, 1, 1),
]
for (n, h, w, ci, co, k, s, p) in SHAPES:
    g = train.Geom(k, s, p)
    oh, ow = g.out_hw(h, w)
    x = torch.randn(n, h, w, ci, device="cuda").to(torch.bfloat16)
    gy = torch.randn(n, oh, ow, co, device="cuda").to(torch.bfloat16)
    for _ in range(2):
        train._wgrad_raw(x, gy, g, (co, ci, k, k))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        train._wgrad_raw(x, gy, g, (co, ci, k, k))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    fl = 2.0 * n * oh * ow * co * ci * k * k
    print("%-44s %8.3f ms  %7.1f TF/s" % (str((n, h, w, ci, co, k, s, p)), ms, fl / ms / 1e9), flush=True)
