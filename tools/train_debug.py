"""Step through the training-side launches one at a time (debug aid)."""
import math, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from ppst_b200 import train

def rel(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm())

def nchw(x):
    return x.float().permute(0, 3, 1, 2).contiguous()

cases = [(3,1,1,32,64,2,24,40),(3,1,1,64,128,3,16,16),(1,1,0,128,32,2,20,12),(6,2,2,64,128,2,32,48),(4,2,1,64,256,2,16,32),(3,1,1,512,512,2,4,4)]
if len(sys.argv) > 1:
    cases = [tuple(int(v) for v in sys.argv[1].split(","))]
for (k,s,p,ci,co,n,h,w) in cases:
    g = torch.Generator().manual_seed(1)
    x = torch.randn(n,h,w,ci,generator=g).to(torch.bfloat16).cuda()
    wt = (torch.randn(co,ci,k,k,generator=g)/math.sqrt(ci*k*k)).to(torch.bfloat16).float().cuda()
    geom = train.Geom(k,s,p)
    oh, ow = geom.out_hw(h,w)
    gy = torch.randn(n,oh,ow,co,generator=g).to(torch.bfloat16).cuda()
    xr = nchw(x).requires_grad_(True); wr = wt.clone().requires_grad_(True)
    with torch.backends.cudnn.flags(allow_tf32=False):
        yr = F.conv2d(xr, wr, stride=s, padding=p); yr.backward(nchw(gy))
    torch.cuda.synchronize()
    print("case", (k,s,p,ci,co,n,h,w), flush=True)
    for name, fn, ref in (("fwd", lambda: nchw(train._conv_raw(x, wt, geom)), yr),
                          ("dgrad", lambda: nchw(train._dgrad_raw(gy, wt, geom, (h,w))), xr.grad),
                          ("wgrad", lambda: train._wgrad_raw(x, gy, geom, wt.shape), wr.grad)):
        try:
            out = fn(); torch.cuda.synchronize()
            print("  %-6s rel %.3e" % (name, rel(out, ref)), flush=True)
        except Exception as e:
            print("  %-6s FAILED: %s" % (name, str(e).split("\n")[0]), flush=True)
            sys.exit(1)
