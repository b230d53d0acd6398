"""One ppst_conv_wgrad call per listed shape (the launches ncu captures for profiles/)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ppst_b200 import train
for (n, h, w, ci, co, k, s, p) in [(8, 256, 256, 256, 256, 3, 1, 1), (8, 512, 512, 128, 128, 3, 1, 1), (8, 256, 256, 128, 256, 6, 2, 2)]:
    g = train.Geom(k, s, p)
    oh, ow = g.out_hw(h, w)
    x = torch.randn(n, h, w, ci, device="cuda").to(torch.bfloat16)
    gy = torch.randn(n, oh, ow, co, device="cuda").to(torch.bfloat16)
    for _ in range(2):
        train._wgrad_raw(x, gy, g, (co, ci, k, k))
    torch.cuda.synchronize()
