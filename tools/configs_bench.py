"""BASELINE.json configs 3 and 4 on one GPU (CUDA events, 3 warm-ups):
  C3  swapping_grid: 16 contents x 16 styles @512^2 through SwapModel.swap_grid (per-image work once, cells batched),
      with and without the guided-filter post-process, against the same 256 cells done as independent swap() pairs;
  C4  plain swap @1024^2, batch 16 (the only swap the reference defines away from 512^2).
Prints one JSON line per measurement:   python tools/configs_bench.py > profiles/rNN_configs.json"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from bench import ClockSampler  # noqa: E402  (nvidia-smi clocks / throttle reasons during the timed regions)
from ppst_b200 import synthetic as O  # noqa: E402  (seeded weights / inputs)
import ppst_b200  # noqa: E402


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


m = ppst_b200.SwapModel(*O.make_networks(0), mode="bf16")
which = sys.argv[1:] or ["c3", "c4"]
sampler = ClockSampler(0)
sampler.start()
if "c3" in which:
    m.noises = O.make_noises(1, 512, 512)
    contents = O.make_portraits(16, 512, 512, seed=11).cuda()
    styles = O.make_portraits(16, 512, 512, seed=12, smooth=True).cuda()
    ms_grid = timed(lambda: m.swap_grid(contents, styles, cell_batch=16), 2)
    ms_smooth = timed(lambda: m.swap_grid(contents, styles, cell_batch=16, smooth=True), 2)
    # the same cells as independent pairs (16 cells per call: one content row against all styles)
    def pairwise():
        for i in range(16):
            m.swap(contents[i:i + 1].expand(16, -1, -1, -1), styles)
    ms_pairs = timed(pairwise, 1)
    gf = timed(lambda: ppst_b200.ops.guided_smooth(contents, styles), 5)
    print(json.dumps({"config": "C3 swapping_grid 16x16 @512^2 (BASELINE configs[2])", "cells": 256,
                      "ms_grid": ms_grid, "cells_per_s": 256 / ms_grid * 1e3,
                      "ms_grid_with_guided_filter": ms_smooth, "cells_per_s_with_guided_filter": 256 / ms_smooth * 1e3,
                      "ms_same_cells_as_independent_swaps": ms_pairs, "encoder_reuse_speedup": ms_pairs / ms_grid,
                      "guided_filter_ms_per_16_images": gf, "dtype": "bf16", "n_gpus": 1, "clocks": sampler.stop()}),
          flush=True)
    sampler = ClockSampler(0)
    sampler.start()
if "c4" in which:
    n = 16
    m.noises = O.make_noises(1, 1024, 1024)
    c = O.make_portraits(n, 1024, 1024, seed=21).cuda()
    s = O.make_portraits(n, 1024, 1024, seed=22, smooth=True).cuda()
    ms = timed(lambda: m.plain_swap(c, s), 3)
    flop = 2.130e12                                   # plain swap @1024^2, conv + GEMM 2*MAC per image
    print(json.dumps({"config": "C4 plain swap @1024^2, batch 16 (BASELINE configs[3]); eager launches", "batch": n,
                      "ms_per_batch": ms, "images_per_s": n / ms * 1e3, "tflops": n * flop / ms / 1e9,
                      "dtype": "bf16", "n_gpus": 1, "clocks": sampler.stop()}), flush=True)
    from ppst_b200 import ops
    ops.PROFILE = {}
    m.plain_swap(c, s)
    torch.cuda.synchronize()
    summ = ops.profile_summary()
    ops.PROFILE = None
    tot = sum(v["ms"] for v in summ.values())
    kern = {}
    for k, v in sorted(summ.items(), key=lambda kv: -kv[1]["ms"]):
        tf = k in ("conv_tc", "conv_simt", "gemm_f32")
        kern[k] = {"ms": v["ms"], "share": v["ms"] / tot, "launches": v["launches"],
                   ("tflops" if tf else "gbs"): v["work"] / (v["ms"] * 1e-3) / (1e12 if tf else 1e9)}
    print(json.dumps({"config": "C4 per-kernel (CUDA events around every launch)", "kernels": kern}), flush=True)
