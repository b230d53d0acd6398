import os, sys, math, torch
sys.path.insert(0, '/root/repo')
from ppst_b200 import ops, packing
from ppst_b200._lib import BF16
from ppst_b200.engine import Engine
eng = Engine.__new__(Engine)
eng.use_tc, eng.adt, eng.device, eng.mode, eng.acode = True, torch.bfloat16, torch.device("cuda"), "bf16", BF16
g = torch.Generator().manual_seed(5)
n,h,w,ci,co = 4,128,128,128,128
x = torch.randn(n,h,w,ci,generator=g).to(torch.bfloat16).cuda()
wt = torch.randn(co,ci,3,3,generator=g)/math.sqrt(ci*9)
spec = packing.make_spec(wt,"cuda",bias=torch.randn(co,generator=g),act=True,pad=1)
os.environ["PPST_CONV_TC_2CTA"]="0"
y0,_ = eng.conv(x,spec); torch.cuda.synchronize(); print("1cta ok")
os.environ["PPST_CONV_TC_2CTA"]="2"
y1,_ = eng.conv(x,spec); torch.cuda.synchronize(); print("2cta ok", torch.equal(y0,y1), float((y0.float()-y1.float()).abs().max()))
