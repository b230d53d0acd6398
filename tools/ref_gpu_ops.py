"""The kernels to beat: the reference's OWN CUDA operators -- stylegan2_op/upfirdn2d_kernel.cu:53-137 and
fused_bias_act_kernel.cu:19-49, JIT-built for this GPU by the reference's own torch.utils.cpp_extension.load calls
(stylegan2_op/upfirdn2d.py:9-19, fused_act.py:10-19) from the verbatim staged copy oracle/_ref/reference -- timed next to
ppst_b200's at the six encoder blur shapes and the generator's bias-act shapes (fp32, batch 8, L2-cold inputs, CUDA-graph
replay so that Python call overhead is excluded on both sides), with the SM clocks / throttle reasons of the run.
MEASUREMENT INFRASTRUCTURE: the product never imports the reference.

    python tools/ref_gpu_ops.py > profiles/rNN_ref_gpu_ops.json      (GPU box; first run JIT-builds for ~2 min)"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("TORCH_EXTENSIONS_DIR", os.path.join(ROOT, "oracle", "_ref", "torch_extensions"))
import torch  # noqa: E402
from bench import ClockSampler, load_peaks  # noqa: E402
from oracle import ref_harness as RH  # noqa: E402
from ppst_b200.stylegan2_op import fused_leaky_relu as our_flrelu, upfirdn2d as our_upfirdn2d  # noqa: E402

assert torch.cuda.is_available() and RH.available(), "needs a GPU and the staged reference tree"
t0 = time.time()
sys.stdout.flush()
_saved_fd = os.dup(1)                      # ninja's build log goes to fd 1: keep stdout for the JSON only
os.dup2(2, 1)
try:
    RH.install(custom_kernels=True)
    import models.networks.stylegan2_layers  # noqa: E402,F401  (imports the reference's stylegan2_op -> JIT build)
finally:
    sys.stdout.flush()
    os.dup2(_saved_fd, 1)
    os.close(_saved_fd)
up_mod = sys.modules["models.networks.stylegan2_op.upfirdn2d"]
fa_mod = sys.modules["models.networks.stylegan2_op.fused_act"]
assert up_mod.use_custom_kernel and fa_mod.use_custom_kernel, "the reference fell back to its PyTorch twins"
build_s = time.time() - t0
PEAK = load_peaks()["hbm_gbs"]


def timeit(fn, bufs, reps=12):
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
    for _ in range(4):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (4 * reps)


B = 8
k3 = torch.tensor([[1., 2, 1], [2, 4, 2], [1, 2, 1]], device="cuda") / 16
k4 = torch.tensor([1., 3, 3, 1], device="cuda")
k4 = (k4[:, None] * k4[None, :]) / 64
rows = []
sampler = ClockSampler(0)
sampler.start()
with torch.no_grad():
    cases = [(32, 515, k3, (0, 0)), (32, 512, k3, (1, 0)), (64, 259, k3, (0, 0)), (64, 256, k3, (1, 0)),
             (128, 131, k3, (0, 0)), (128, 128, k3, (1, 0)),
             (64, 512, k4, (2, 1)), (128, 256, k4, (2, 1))]          # + the discriminator's 4x4 blur (stylegan2_layers.py:517)
    for c, hin, k, pad in cases:
        bufs = [torch.randn(B, c, hin, hin, device="cuda") for _ in range(4)]
        y_ref = up_mod.upfirdn2d(bufs[0], k, pad=pad)
        y_our = our_upfirdn2d(bufs[0], k, pad=pad)
        err = float((y_ref - y_our).abs().max())
        ms_ref = timeit(lambda x: up_mod.upfirdn2d(x, k, pad=pad), bufs)
        ms_our = timeit(lambda x: our_upfirdn2d(x, k, pad=pad), bufs)
        ho = hin + pad[0] + pad[1] - (k.shape[0] - 1)
        byt = B * c * (hin * hin + ho * ho) * 4
        rows.append({"op": "upfirdn2d", "shape": [B, c, hin, hin], "kernel": list(k.shape), "pad": list(pad),
                     "ms_reference_cuda": ms_ref, "ms_ppst_b200": ms_our, "speedup": ms_ref / ms_our,
                     "gbs_reference_cuda": byt / ms_ref / 1e6, "gbs_ppst_b200": byt / ms_our / 1e6,
                     "frac_of_measured_hbm": byt / ms_our / 1e6 / PEAK, "max_abs_diff": err})
    for c, h in ((256, 64), (384, 64), (512, 64), (512, 128), (256, 256), (128, 512), (32, 512), (64, 256), (128, 128)):
        bufs = [torch.randn(B, c, h, h, device="cuda") for _ in range(3)]
        bias = torch.randn(c, device="cuda")
        err = float((fa_mod.fused_leaky_relu(bufs[0], bias) - our_flrelu(bufs[0], bias)).abs().max())
        ms_ref = timeit(lambda x: fa_mod.fused_leaky_relu(x, bias), bufs)
        ms_our = timeit(lambda x: our_flrelu(x, bias), bufs)
        byt = 2 * B * c * h * h * 4
        rows.append({"op": "fused_leaky_relu", "shape": [B, c, h, h], "ms_reference_cuda": ms_ref, "ms_ppst_b200": ms_our,
                     "speedup": ms_ref / ms_our, "gbs_reference_cuda": byt / ms_ref / 1e6, "gbs_ppst_b200": byt / ms_our / 1e6,
                     "frac_of_measured_hbm": byt / ms_our / 1e6 / PEAK, "max_abs_diff": err})
clocks = sampler.stop()
print(json.dumps({"what": "reference CUDA ops (JIT, sm_%d%d) vs ppst_b200, fp32, batch 8" % torch.cuda.get_device_capability(),
                  "jit_build_s": build_s, "hbm_peak_gbs": PEAK, "clocks": clocks,
                  "all_faster_or_equal": all(r["speedup"] >= 0.98 for r in rows), "rows": rows}, indent=1))
