"""Decompose the stage-1 gradient error: each network alone with a fixed random upstream gradient (debug aid)."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import ppst_oracle as O
from ppst_b200 import train_nets as TN

def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))

size, n = int(sys.argv[1]) if len(sys.argv) > 1 else 128, 2
sdE1, sdE2, sdG = O.make_networks(0)
real = O.make_portraits(n, size, size, seed=21, smooth=True)
noises = O.make_noises(n, size, size, seed=5)
g = torch.Generator().manual_seed(3)
def leaf(sd):
    return {k: (v.clone().requires_grad_(True) if (v.is_floating_point() and not k.endswith("kernel")) else v) for k, v in sd.items()}
# --- E1 alone
r = leaf(sdE1); sp_r = O.encoder_con(r, real)
R = torch.randn(sp_r.shape, generator=g)
(sp_r * R).sum().backward()
P = TN.ParamSet(sdE1).cuda(); sp = TN.encoder_con(P, real.cuda())
(sp.float() * R.permute(0, 2, 3, 1).cuda()).sum().backward()
print("E1 fwd", rel(sp.float().permute(0, 3, 1, 2), sp_r))
for k in ["FromRGB.Conv.weight", "FromRGB.Act.bias", "DownToSpatialCode.ResBlockDownBy1.conv1.Conv.weight", "DownToSpatialCode.ResBlockDownBy1.conv1.Act.bias",
          "DownToSpatialCode.ResBlockDownBy1.conv2.Conv.weight", "DownToSpatialCode.ResBlockDownBy1.skip.Conv.weight",
          "DownToSpatialCode.ResBlockDownBy4.conv2.Conv.weight", "ToSpatialCode.0.Conv.weight", "ToSpatialCode.0.Act.bias", "ToSpatialCode.1.Conv.weight", "ToSpatialCode.1.Conv.bias"]:
    print("  E1 %-60s %.4f" % (k, rel(P[k].grad, r[k].grad)))
# --- E2 alone
r = leaf(sdE2); gl_r = O.encoder_col(r, real)[0]
Rs = [torch.randn(v.shape, generator=g) for v in gl_r]
sum((v * q).sum() for v, q in zip(gl_r, Rs)).backward()
P = TN.ParamSet(sdE2).cuda(); gl = TN.encoder_col(P, real.cuda())
sum((v * q.cuda()).sum() for v, q in zip(gl, Rs)).backward()
print("E2 fwd", [round(rel(a, b), 4) for a, b in zip(gl, gl_r)])
for k in ["FromRGB.Conv.weight", "DownToGlobalCode1.ResBlockDownBy1.conv1.Conv.weight", "DownToGlobalCode1.ResBlockDownBy1.conv2.Conv.weight",
          "DownToGlobalCode1.ResBlockDownBy4.skip.Conv.weight", "conv1x1_9.weight", "conv1x1_2.weight", "projector1.3.weight", "projector2.5.bias"]:
    print("  E2 %-60s %.4f" % (k, rel(P[k].grad, r[k].grad)))
# --- G alone (inputs = the oracle's sp / codes)
r = leaf(sdG)
sp_in = O.encoder_con(sdE1, real).detach().requires_grad_(True)
gl_in = [v.detach().requires_grad_(True) for v in O.encoder_col(sdE2, real)[0]]
rec_r = O.generator(r, sp_in, gl_in, noises=noises)
Rr = torch.randn(rec_r.shape, generator=g)
(rec_r * Rr).sum().backward()
P = TN.ParamSet(sdG).cuda()
sp_c = sp_in.detach().permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda().requires_grad_(True)
gl_c = [v.detach().cuda().requires_grad_(True) for v in gl_in]
rec = TN.generator(P, sp_c, gl_c, noises=noises)
(rec * Rr.cuda()).sum().backward()
print("G fwd", rel(rec, rec_r), " d/dsp", rel(sp_c.grad.float().permute(0, 3, 1, 2), sp_in.grad), " d/dcodes", [round(rel(a.grad, b.grad), 4) for a, b in zip(gl_c, gl_in)])
for k in ["SpatialCodeModulation.scale.weight", "HeadResnetBlock0.conv1.conv.weight", "HeadResnetBlock0.conv1.epi1.style_mod.lin.weight", "HeadResnetBlock0.conv1.noise.weight",
          "HeadResnetBlock2.skip.Conv.weight", "HeadResnetBlock3.conv2.conv.weight", "UpsamplingResBlock16.conv1.conv.weight", "UpsamplingResBlock16.conv2.conv.bias",
          "UpsamplingResBlock32.conv1.conv.weight", "UpsamplingResBlock32.skip.Conv.weight", "UpsamplingResBlock64.conv1.conv.weight", "UpsamplingResBlock64.conv2.conv.weight",
          "UpsamplingResBlock64.conv2.bias", "UpsamplingResBlock64.conv2.activate.bias", "ToRGB.conv.weight", "ToRGB.bias", "ToRGB.epi1.style_mod.lin.weight"]:
    print("  G  %-60s %.4f" % (k, rel(P[k].grad, r[k].grad)))
