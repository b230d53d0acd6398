import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ppst_b200.stylegan2_op import upfirdn2d
k3 = torch.tensor([[1., 2, 1], [2, 4, 2], [1, 2, 1]], device="cuda") / 16
x = torch.randn(8, 32, 515, 515, device="cuda")
for _ in range(3):
    y = upfirdn2d(x, k3)
torch.cuda.synchronize()
print(y.shape)
