"""One launch of each HBM-bound training kernel at the largest stage-1 shape (8 x 256^2 x 128 channels, bf16) -- what
ncu captures for profiles/ (plane_dot, act_grad_sum, blur_nhwc, the x2 / pad adjoints)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ppst_b200 import train
n, h, w, c = 8, 256, 256, 128
x = torch.randn(n, h, w, c, device="cuda").to(torch.bfloat16)
g = torch.randn(n, h, w, c, device="cuda").to(torch.bfloat16)
k = torch.tensor([1., 3., 3., 1.], device="cuda")
k = (k[None] * k[:, None] / 64).contiguous()
for _ in range(2):
    train._plane_dot(g, x)
    train._ActGradSum.apply(g, x, True)
    train._Blur.apply(x, k, 2, 3, True)
    xs = x[:, :128, :128].contiguous().requires_grad_(True)
    train.bilinear2x(xs).backward(g)
    xp = x.clone().requires_grad_(True)
    train.reflect_pad(xp, (2, 2, 2, 2)).backward(torch.randn(n, h + 4, w + 4, c, device="cuda").to(torch.bfloat16))
torch.cuda.synchronize()
