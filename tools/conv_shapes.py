"""Time selected ppst_conv_tc shapes in isolation (CUDA events, L2-cold inputs rotated) -- development aid and the
command profiled under ncu for profiles/*conv_tc*.   python tools/conv_shapes.py [reps]"""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from ppst_b200 import ops  # noqa: E402
from ppst_b200._lib import BF16  # noqa: E402

SHAPES = [  # n, h, w, cin, cout, k, stride, phases
    (8, 512, 512, 128, 128, 3, 1, 1),
    (8, 256, 256, 256, 256, 3, 1, 1),
    (8, 256, 256, 256, 128, 2, 1, 4),
    (8, 128, 128, 512, 512, 3, 1, 1),
    (8, 64, 64, 256, 256, 3, 1, 1),
    (8, 512, 512, 32, 32, 3, 1, 1),
    (8, 512, 512, 128, 64, 1, 1, 1),
    (8, 256, 256, 32, 64, 5, 2, 1),
    (8, 512, 512, 256, 128, 2, 1, 4),
    (8, 256, 256, 512, 256, 2, 1, 4),
    (8, 256, 256, 64, 64, 3, 1, 1),
    (8, 256, 256, 256, 128, 3, 1, 1),
    (8, 128, 128, 128, 128, 3, 1, 1),
    (8, 64, 64, 512, 512, 3, 1, 1),
]
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
only = int(sys.argv[2]) if len(sys.argv) > 2 else -1
dev = "cuda"
for idx, (n, h, w, ci, co, k, s, ph) in enumerate(SHAPES):
    if only >= 0 and idx != only:
        continue
    if ph == 4:
        ih, iw, oh, ow = h // 2, w // 2, h // 2, w // 2
        phases, off, poff = (2, 2), (-1, -1), (1, 1)
    else:
        oh, ow = h, w
        ih, iw = (h * s + (k - s), w * s + (k - s)) if s == 2 else (h, w)
        phases, off, poff = (1, 1), ((0, 0) if s == 2 else (-(k // 2), -(k // 2))), (0, 0)
    xs = [torch.randn(n, ih, iw, ci, device=dev).bfloat16() for _ in range(3)]
    wt = (torch.randn(ph, co, k * k * ci, device=dev) / math.sqrt(k * k * ci)).bfloat16()
    bias = torch.randn(co, device=dev)
    d = ops.conv_desc((n, ih, iw, ci), co, k, k, stride=s, off=off, out_hw=(oh, ow), phases=phases, phase_off=poff,
                      act=True, gain=math.sqrt(2), in_dtype=BF16, out_dtype=BF16)
    out = torch.empty(n, oh * phases[0], ow * phases[1], co, device=dev, dtype=torch.bfloat16)
    st = torch.empty(ops.conv_tc_stats_elems(d), device=dev) if os.environ.get("NOSTATS") is None else None
    if os.environ.get("NOBIAS") is not None:
        bias = None
    for i in range(2):
        ops.conv_tc(out, xs[i % 3], wt, bias, None, st, d)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        ops.conv_tc(out, xs[i % 3], wt, bias, None, st, d)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    fl = 2.0 * n * oh * ow * ph * co * k * k * ci
    byt = (xs[0].numel() + out.numel()) * 2
    print("%-44s %8.3f ms  %7.1f TF/s  %7.1f GB/s(min traffic)" % (str((n, h, w, ci, co, k, s, ph)), ms, fl / ms / 1e9,
                                                                 byt / ms / 1e6))
