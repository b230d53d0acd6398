"""HBM roofline of the two plugin operators at the reference's hot-path call shapes.
Inputs are rotated over several buffers so that every timed launch reads L2-cold data.
    python tools/op_bench.py > profiles/rNN_op_bench.json"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from ppst_b200.stylegan2_op import fused_leaky_relu, upfirdn2d  # noqa: E402

PEAK = 6549.8
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timeit(fn, bufs, reps=12):
    """Device time per call: the calls are captured into one CUDA graph (the Python wrapper costs ~50-80 us per call,
    more than the small shapes' kernels) and the graph is replayed; inputs rotate so each launch reads L2-cold data."""
    for b in bufs:
        fn(b)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for i in range(reps):
                fn(bufs[i % len(bufs)])
    torch.cuda.synchronize()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (2 * reps)


out = []
k3 = torch.tensor([[1., 2, 1], [2, 4, 2], [1, 2, 1]], device="cuda") / 16
B = 8
for dtype in (torch.float32, torch.bfloat16):
    es = 4 if dtype == torch.float32 else 2
    # the six encoder blur calls (encoder pass at 512^2), batch 8
    for (c, hin, pad) in ((32, 515, (0, 0)), (32, 512, (1, 0)), (64, 259, (0, 0)), (64, 256, (1, 0)), (128, 131, (0, 0)),
                          (128, 128, (1, 0))):
        bufs = [torch.randn(B, c, hin, hin, device="cuda").to(dtype) for _ in range(4)]
        ms = timeit(lambda x: upfirdn2d(x, k3, pad=pad), bufs)
        ho = hin + pad[0] + pad[1] - 2
        byt = B * c * (hin * hin + ho * ho) * es
        out.append({"op": "upfirdn2d", "dtype": str(dtype)[6:], "shape": [B, c, hin, hin], "pad": list(pad), "ms": ms,
                    "gbs": byt / ms / 1e6, "frac_of_measured_hbm": byt / ms / 1e6 / PEAK})
    # fused_leaky_relu at generator sizes
    for (c, h) in ((128, 512), (256, 256), (512, 128), (512, 64), (32, 512)):
        bufs = [torch.randn(B, c, h, h, device="cuda").to(dtype) for _ in range(3)]
        bias = torch.randn(c, device="cuda").to(dtype)
        ms = timeit(lambda x: fused_leaky_relu(x, bias), bufs)
        byt = 2 * B * c * h * h * es
        out.append({"op": "fused_leaky_relu", "dtype": str(dtype)[6:], "shape": [B, c, h, h], "ms": ms,
                    "gbs": byt / ms / 1e6, "frac_of_measured_hbm": byt / ms / 1e6 / PEAK})
# torch copy of the same size as a sanity reference for the achievable read+write rate
x = [torch.randn(B, 128, 512, 512, device="cuda") for _ in range(3)]
y = torch.empty_like(x[0])
ms = timeit(lambda t: y.copy_(t), x)
out.append({"op": "torch.copy_ (reference rate)", "dtype": "float32", "shape": [B, 128, 512, 512], "ms": ms,
            "gbs": 2 * x[0].numel() * 4 / ms / 1e6, "frac_of_measured_hbm": 2 * x[0].numel() * 4 / ms / 1e6 / PEAK})
for r in out:
    print(json.dumps(r))
