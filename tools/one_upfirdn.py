"""One upfirdn2d launch per hot-path shape, for `ncu -k regex:upfirdn2d` captures."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from ppst_b200.stylegan2_op import upfirdn2d  # noqa: E402

k3 = torch.tensor([[1., 2, 1], [2, 4, 2], [1, 2, 1]], device="cuda") / 16
dt = torch.bfloat16 if "bf16" in sys.argv else torch.float32
for c, h in ((32, 515), (64, 259), (128, 131)):
    x = torch.randn(8, c, h, h, device="cuda").to(dt)
    for _ in range(3):
        y = upfirdn2d(x, k3)
    torch.cuda.synchronize()
    print(tuple(y.shape))
