#!/usr/bin/env python
"""Benchmark of the PPST swap hot path (BASELINE.json metric: swapped images/s @512^2 at 1/2/4/8 B200).

    python bench.py --gpus N --steps K --warmup W            # ours (ppst_b200, bf16 tensor-core mode)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own modules on the host cores (staged copy
                                                             # oracle/_ref/reference; the oracle port if it is absent)

Workload (config.workload): BASELINE configs[1] "batched swap inference at 512^2, batch 64 ... sharded over
8xB200" = 8 (content, style) pairs per GPU per step, weak scaling (per-GPU work fixed, 64 pairs at N=8).
One step = one pass of the whole swap path over one batch of synthetic pairs:
  E1(c) E2(c) G.extract(c)  E1(s) E2(s) G.extract(s)  Rselfcorr x2  corrm  E2(s, corr)  G.decode   (2.143 TFLOP / image,
  SURVEY.md 8(d)).  No collective on the data path; torch.distributed is used for barriers and the max-over-ranks time.

One JSON line on stdout (rank 0).  `value` = device-resident inputs; `e2e` = through SwapModel.swap_pipelined() with
pinned HOST buffers, H2D of both image batches and D2H of the result of every step inside the timed region (the copies run
on their own streams and overlap the neighbouring steps' kernels).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":     # NCCL prints its version banner to STDOUT at the VERSION and
    del os.environ["NCCL_DEBUG"]                               # WARN levels; stdout must carry exactly one JSON line

import torch  # noqa: E402

PAIRS_PER_GPU = 8
SIZE = 512
FLOP_PER_IMAGE = 2.143e12            # conv + GEMM 2*MAC of the minimal swap unit at 512^2
FLOP_PER_IMAGE_RUN = 2.143e12 - 22.65e9   # what the engine runs: E2(style)'s trunk is computed once, not again inside
                                          # E2(style, corr) as encode2 does (E2 plain = 22.65 GF)
METRIC = "swapped images/s @512^2"


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "tf_burst": p["bf16_tflops"], "tf_sustained": p["bf16_tflops_sustained"],
                "source": "measured"}
    except Exception:
        return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def load_traffic():
    """DRAM bytes per conv_tc launch from the committed `ncu --set full` capture of one step (tools/ncu_traffic.py)."""
    import glob
    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_conv_tc_traffic.json")), reverse=True):
        try:
            with open(f) as fh:
                j = json.load(fh)
            return j["traffic_bytes_per_launch"], os.path.basename(f)
        except Exception:
            continue
    return None, None


def make_inputs(pairs, rank):
    from ppst_b200 import synthetic as O    # seeded synthetic portraits
    c = O.make_portraits(pairs, SIZE, SIZE, seed=1234 + rank)
    s = O.make_portraits(pairs, SIZE, SIZE, seed=4321 + rank, smooth=True)
    return c, s


def _cpu_reference(threads):
    """The reference's CPU implementation of one swap unit: a callable() -> image, plus its `kind`.
    "reference": the UNMODIFIED reference modules and PPSTModel methods (oracle/ref_harness.RefModel over the staged copy
    oracle/_ref/reference, use_custom_kernel = False -> upfirdn2d_native / F.leaky_relu, fp32, all host threads) on the
    same seeded weights / inputs as the GPU arm; "port": the oracle's restatement, if the staged tree is absent."""
    import contextlib
    from oracle import ppst_oracle as O
    from oracle import ref_harness as RH
    torch.set_num_threads(threads)
    sds = O.make_networks(0)
    c = O.make_portraits(1, SIZE, SIZE, seed=1234)
    s = O.make_portraits(1, SIZE, SIZE, seed=4321, smooth=True)
    noises = O.make_noises(1, SIZE, SIZE)
    if RH.available():
        with contextlib.redirect_stdout(sys.stderr):           # the reference prints while it initialises its networks
            m = RH.RefModel(sds, device="cpu", noises=noises)
        return (lambda: m.swap_unit(c, s)), "reference"

    def port():
        with torch.no_grad():
            return O.swap_pair(*sds, c, s, noises=noises)
    return port, "port"


def cpu_baseline_sample(threads):
    """The reference's CPU path on a bounded sample: ONE pair at 512^2 (2.143 TFLOP), all host threads.
    Returns (images/s, seconds of CPU work, kind)."""
    fn, kind = _cpu_reference(threads)
    reps = 2
    fn()                                                       # warm-up (allocator, oneDNN primitive caches)
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    dt = (time.perf_counter() - t0) / reps
    return 1.0 / dt, dt * (reps + 1), kind


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (its unmodified modules from the staged copy
    oracle/_ref/reference; the oracle port only if that copy is absent), all host threads, each step a bounded sample of
    the workload (one pair)."""
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    fn, kind = _cpu_reference(threads)
    steps, warm = max(1, min(args.steps, 4)), max(1, min(args.warmup, 1))
    for _ in range(warm):
        fn()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    dt = (time.perf_counter() - t0) / steps
    v = 1.0 / dt
    what = ("the reference's own modules and PPSTModel methods, use_custom_kernel=False (upfirdn2d_native / F.leaky_relu)"
            if kind == "reference" else "oracle port of the reference's pure-PyTorch CPU path (staged reference tree absent)")
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "images/s", "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C2 swap @512^2; reference arm = bounded sample of 1 (content, style) pair per step "
                                   "on host cores: " + what,
                       "pairs_per_step": 1, "size": SIZE},
            "cpu_baseline": {"value": v, "unit": "images/s", "cores": threads, "kind": kind,
                             "sample": "1 pair @512^2 per step, %d step(s)" % steps},
            "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_reference_train(args):
    """--impl reference --workload train: the reference's CPU path for the stage-1 iteration (oracle port: its fp32 torch
    expression of E1 / E2 / G / D under torch.autograd, all host threads), each step a bounded sample: ONE iteration
    (generator step with the L1 + GAN losses, discriminator step) on 2 images at 256^2."""
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ppst_oracle as O
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    size, batch = 256, 2

    def leaf(sd):
        return {k: (v.clone().requires_grad_(True) if (v.is_floating_point() and not k.endswith("kernel")) else v) for k, v in sd.items()}
    sdE1, sdE2, sdG = O.make_networks(0)
    for k in sdG:
        if k.endswith("noise.weight"):
            sdG[k] = torch.zeros_like(sdG[k])
    rE1, rE2, rG, rD = leaf(sdE1), leaf(sdE2), leaf(sdG), leaf(O.make_discriminator(size, seed=0))
    gp = [v for d in (rE1, rE2, rG) for v in d.values() if v.requires_grad]
    dp = [v for v in rD.values() if v.requires_grad]
    og = torch.optim.Adam(gp, lr=2e-3, betas=(0.0, 0.99))
    od = torch.optim.Adam(dp, lr=2e-3, betas=(0.0, 0.99))
    real = O.make_portraits(batch, size, size, seed=77, smooth=True)

    def rec(x):
        return O.generator(rG, O.encoder_con(rE1, x), O.encoder_col(rE2, x)[0])

    def iteration():
        og.zero_grad(set_to_none=True)
        r = rec(real)
        ((r - real).abs().mean() + 0.5 * O.gan_loss(O.discriminator(rD, r), True)).backward()
        og.step()
        od.zero_grad(set_to_none=True)
        with torch.no_grad():
            r = rec(real[: batch // 2])
        sum(O.d_losses(rD, real, r).values()).backward()
        od.step()
    steps, warm = max(1, min(args.steps, 3)), 1
    for _ in range(warm):
        iteration()
    t0 = time.perf_counter()
    for _ in range(steps):
        iteration()
    dt = (time.perf_counter() - t0) / steps
    v = batch / dt
    line = {"impl": "reference", "metric": "training images/s @256^2 (stage 1: G step + D step + R1/16)", "value": v,
            "unit": "images/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C5 stage-1 training @256^2; reference arm = bounded sample of 2 images per iteration on host "
                                   "cores (oracle port of the reference's pure-PyTorch CPU path; R1 step not included)",
                       "batch": batch, "size": size},
            "cpu_baseline": {"value": v, "unit": "images/s", "cores": threads, "kind": "port",
                             "sample": "1 iteration (G step + D step) on 2 images @256^2 per step, %d step(s)" % steps},
            "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_train(args):
    """BASELINE configs[4] (training step, DDP with NCCL gradient all-reduce at 1/2/4/8 GPUs): training stage 1 at 256^2,
    8 images per GPU (weak scaling) -- generator step (E1, E2, G forward + backward, D forward + input gradient), discriminator
    step, R1 step on every 16th iteration, Adam, one flat NCCL all-reduce per step; CUDA-graph replay.  Measured by
    tools/train_bench.py (same code, more detail); this wrapper prints the bench.py-shaped line."""
    import importlib.util
    import types
    spec = importlib.util.spec_from_file_location("train_bench", os.path.join(ROOT, "tools", "train_bench.py"))
    tb = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tb)
    a = types.SimpleNamespace(size=256, batch=8, steps=max(4, min(args.steps, 10)), no_torch=True, stage1=True, return_dict=True)
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    sampler.start()
    out = tb.stage1(a)
    clocks = sampler.stop()
    if out is None:
        return
    line = {"metric": "training images/s @256^2 (stage 1: G step + D step + R1/16)", "value": out["images_per_s"],
            "unit": "images/s", "n_gpus": out["n_gpus"], "steps": a.steps, "warmup": 2, "ms_per_step": out["iteration_ms"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "C5 (BASELINE configs[4]): " + out["workload"], "batch_per_gpu": a.batch, "size": a.size,
                       "parallelism": "data parallel, one flat NCCL gradient all-reduce per step",
                       "grad_allreduce_bytes": out["grad_allreduce_bytes"]},
            "e2e": {"value": out["e2e_images_per_s"], "unit": "images/s", "h2d_bytes_per_step": out["h2d_bytes_per_iteration"],
                    "d2h_bytes_per_step": out["d2h_bytes_per_iteration"], "ms_per_step": out["iteration_e2e_ms"]},
            "gpu_launches": out["launches_g_d_r1"] * a.steps, "step_ms": out["graph_ms"], "eager_step_ms": out["ms"],
            "clocks": clocks, "kernels": out["kernels_g_plus_d"]}
    # roofline of the dominant kernels: every tcgen05 launch of a G step + D step (forward / input-gradient convolutions =
    # ppst_conv_tc, weight gradients = ppst_conv_wgrad), CUDA events per launch in the eager profiling pass
    peaks = load_peaks()
    k = out["kernels_g_plus_d"]
    tc_ms = sum(k[n]["ms"] for n in ("conv_tc", "conv_wgrad") if n in k)
    tc_flop = sum(k[n]["ms"] * 1e-3 * (k[n]["tflops"] or 0.0) * 1e12 for n in ("conv_tc", "conv_wgrad") if n in k)
    tc_launches = sum(k[n]["launches"] for n in ("conv_tc", "conv_wgrad") if n in k)
    if tc_ms > 0:
        ach = tc_flop / (tc_ms * 1e-3) / 1e12
        line["roofline"] = {"bound": "tensor", "kernel": "ppst_conv_tc + ppst_conv_wgrad (every launch of one G step + one D step)",
                            "achieved": ach, "peak": peaks["tf_sustained"], "unit": "TFLOP/s", "frac": ach / peaks["tf_sustained"],
                            "traffic": None, "algorithmic_flop_per_launch": tc_flop / max(tc_launches, 1),
                            "launches_per_iteration": tc_launches,
                            "peak_source": peaks["source"] + " bf16_tflops_sustained (kernels timed inside a long step)",
                            "note": "256^2 training shapes: 32x32 .. 256x256 maps, 32 .. 512 channels, batch 8"}
    line["cpu_baseline"] = None
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=PAIRS_PER_GPU, help="(content, style) pairs per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="swap", choices=["swap", "train"],
                    help="swap: the headline metric (BASELINE configs[1]); train: configs[4], training stage 1 at 256^2")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_train(args) if args.workload == "train" else run_reference(args)
    if args.workload == "train":
        return run_train(args)

    from ppst_b200 import sharding
    world, rank, local_rank = sharding.init()
    assert torch.cuda.is_available(), "bench.py (ours) needs a CUDA device: there is no CPU fallback"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    import ppst_b200
    from ppst_b200 import synthetic as O         # seeded weights / inputs (the oracle is only touched by the CPU legs)
    from ppst_b200 import _lib, ops

    warm = max(3, args.warmup)
    pairs = args.pairs
    sdE1, sdE2, sdG = O.make_networks(0)
    model = ppst_b200.SwapModel(sdE1, sdE2, sdG, device=dev, mode="bf16")
    assert model.engine.use_tc, "tcgen05 path inactive"
    model.noises = {k: v.expand(pairs, -1, -1, -1).contiguous() for k, v in O.make_noises(1, SIZE, SIZE).items()}
    c_host, s_host = make_inputs(pairs, rank)
    c_pin, s_pin = c_host.pin_memory(), s_host.pin_memory()
    out_pin = torch.empty((pairs, 3, SIZE, SIZE), dtype=torch.float32).pin_memory()
    c_dev, s_dev = c_pin.to(dev), s_pin.to(dev)

    def step_device():
        return model.swap_graph(c_dev, s_dev)

    def step_eager():
        return model.swap(c_dev, s_dev)

    last = [None]

    def step_e2e():
        # H2D of both batches (copy stream, overlapping the previous step's kernels), graph replay, D2H of the result
        # (third stream); every step moves its own 50 MB in and 25 MB out
        last[0] = model.swap_pipelined(c_pin, s_pin, out_pin)

    def timed(fn, steps):
        sharding.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _lib.launch_count()
        last[0] = None
        e0.record()
        for _ in range(steps):
            fn()
        if last[0] is not None:
            torch.cuda.current_stream().wait_event(last[0])      # the last result has reached the host buffer
        e1.record()
        torch.cuda.synchronize()
        sharding.barrier()
        ms = sharding.max_over_ranks(e0.elapsed_time(e1), device=dev)
        return ms, _lib.launch_count() - l0

    l0 = _lib.launch_count()
    step_eager()                                   # one eager step: counts the kernels a step launches
    launches_per_step = _lib.launch_count() - l0
    for _ in range(warm):
        step_device()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms, _ = timed(step_device, args.steps)
    launches = launches_per_step * args.steps       # replayed from the captured CUDA graph (same kernels, same order)
    clocks = sampler.stop()
    for _ in range(2):
        step_e2e()
    ms_e2e, _ = timed(step_e2e, args.steps)

    # one replay at a time with the GPU idle in between: what a step costs before the 1 kW power cap pulls the SM clock
    # down (the gap between the sum of kernel times of a single replay and the back-to-back step time is clocks, not idle gaps)
    cool = []
    for _ in range(3):
        torch.cuda.synchronize()
        time.sleep(0.3)
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        step_device()
        c1.record()
        torch.cuda.synchronize()
        cool.append(c0.elapsed_time(c1))
    ms_single_replay = min(cool)

    # per-kernel roofline leg: the same steps with every launch bracketed by CUDA events
    ops.PROFILE = {}
    torch.cuda.synchronize()
    prof_steps = min(args.steps, 2)
    for _ in range(prof_steps):
        step_eager()
    torch.cuda.synchronize()
    prof = ops.profile_summary()
    ops.PROFILE = None

    peaks = load_peaks()
    traffic, traffic_src = load_traffic()
    total_pairs = pairs * world
    value = total_pairs * args.steps / (ms * 1e-3)
    e2e = total_pairs * args.steps / (ms_e2e * 1e-3)
    # every ppst_conv_tc launch of a step: the convolutions and the two split GEMMs (corrm logits, corr @ features).
    # "alg" = ALGORITHMIC FLOPs of the reference op each launch stands for (3x3 / 1x1 taps for the layers
    # whose Blur is folded into 5x5 / 3x3 taps, 1 x K for the 3-term split GEMMs); "work" = executed FLOPs.
    tc = {"work": 0.0, "alg": 0.0, "ms": 0.0, "launches": 0, "aux": 0.0}
    for key in ("conv_tc", "gemm_split_tc"):
        for f in tc:
            tc[f] += prof.get(key, {}).get(f, 0)
    if tc["ms"] <= 0:
        tc["ms"] = 1.0
    step_ms_prof = sum(v["ms"] for v in prof.values()) / max(prof_steps, 1)
    # share of the step spent in the tensor-core kernel UNDER GRAPH REPLAY: kernel activity records (CUPTI through
    # torch.profiler) of two replays; the share, not the absolute (activity tracing perturbs absolute times slightly)
    tc_share = None
    try:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as kprof:
            for _ in range(2):
                step_device()
            torch.cuda.synchronize()
        tc_us = all_us = 0.0
        for e in kprof.events():
            if e.device_type != torch.autograd.DeviceType.CUDA:
                continue
            d_us = e.time_range.end - e.time_range.start
            all_us += d_us
            if "conv_tc" in e.name:
                tc_us += d_us
        if all_us > 0:
            tc_share = tc_us / all_us
    except Exception:
        tc_share = None
    share_source = "kernel activity records of 2 graph replays (torch.profiler / CUPTI)"
    if tc_share is None:                                   # fall back to the eager event profile (wrapped launches only)
        tc_share = (tc["ms"] / max(prof_steps, 1)) / step_ms_prof if step_ms_prof else 0.0
        share_source = "eager per-launch CUDA events (wrapped launches only)"
    achieved_eager = tc["alg"] / (tc["ms"] * 1e-3) / 1e12
    ms_step = ms / args.steps
    tc_ms_in_step = tc_share * ms_step
    achieved_tf = (tc["alg"] / max(prof_steps, 1)) / (tc_ms_in_step * 1e-3) / 1e12 if tc_ms_in_step > 0 else 0.0
    executed_tf = (tc["work"] / max(prof_steps, 1)) / (tc_ms_in_step * 1e-3) / 1e12 if tc_ms_in_step > 0 else 0.0
    kernels = {k: {"launches_per_step": v["launches"] // max(prof_steps, 1), "ms_per_step": v["ms"] / max(prof_steps, 1),
                   "share": (v["ms"] / max(prof_steps, 1)) / step_ms_prof if step_ms_prof else None,
                   ("tflops" if k in ("conv_tc", "conv_simt", "gemm_f32", "gemm_split_tc") else "gbs"):
                       v["alg"] / (v["ms"] * 1e-3) / (1e12 if k in ("conv_tc", "conv_simt", "gemm_f32", "gemm_split_tc") else 1e9)
                       if v["ms"] > 0 else None}
               for k, v in prof.items()}

    if rank != 0:
        return
    line = {
        "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "C2 (BASELINE configs[1]): batched swap inference @512^2, %d (content, style) pairs per "
                               "GPU per step (64 over 8 GPUs), random-init CelebAMaskHQ_default E1/E2/G, full "
                               "correlation swap (2.143 TFLOP/image)" % pairs,
                   "pairs_per_gpu": pairs, "size": SIZE, "parallelism": "pairs partitioned, no collective",
                   "l2": "working set per step >> 126 MB L2 (activations ~ GBs); no explicit flush needed",
                   "flop_per_image": FLOP_PER_IMAGE_RUN,
                   "flop_per_image_note": "SURVEY 8(d) unit (2.143 TF) minus the E2(style) trunk the engine does not recompute",
                   "whole_swap_tflops": value * FLOP_PER_IMAGE_RUN / 1e12 / world,
                   "whole_swap_frac_of_bf16_sustained": value * FLOP_PER_IMAGE_RUN / 1e12 / world / peaks["tf_sustained"]},
        "e2e": {"value": e2e, "unit": "images/s", "h2d_bytes_per_step": int(2 * c_pin.numel() * 4),
                "d2h_bytes_per_step": int(out_pin.numel() * 4), "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "launch_mode": "CUDA graph replay of %d kernels per step (captured from the eager path)" % launches_per_step,
        "ms_single_replay_after_idle": ms_single_replay,
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "ppst_conv_tc (tcgen05 implicit-GEMM convolutions + the two split GEMMs: every "
                                                 "launch of one step)",
                     "achieved": achieved_tf, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                     "frac": achieved_tf / peaks["tf_sustained"], "traffic": traffic, "traffic_unit": "DRAM bytes per launch",
                     "traffic_source": traffic_src,
                     "how": "ALGORITHMIC FLOPs of all launches of one step / (their share of the CUDA-event-timed kernel "
                            "time x the graph-replayed step time of `value`)",
                     "algorithmic_flop_per_launch": tc["alg"] / max(tc["launches"], 1),
                     "executed_flop_per_launch": tc["work"] / max(tc["launches"], 1),
                     "executed_tflops_in_step": executed_tf,
                     "achieved_eager_launches": achieved_eager,
                     "frac_eager_launches": achieved_eager / peaks["tf_sustained"],
                     "algorithmic_bytes_per_launch": tc.get("aux", 0.0) / max(tc["launches"], 1),
                     "peak_source": peaks["source"] + " bf16_tflops_sustained (kernel timed inside a long step)",
                     "launches_per_step": tc["launches"] // max(prof_steps, 1),
                     "share_of_step": tc_share, "share_source": share_source, "ms_in_step": tc_ms_in_step},
        "kernels": kernels,
    }
    if not args.no_cpu_baseline and world == 1:
        threads = os.cpu_count() or 1
        v, secs, kind = cpu_baseline_sample(threads)
        line["cpu_baseline"] = {"value": v, "unit": "images/s", "cores": threads, "kind": kind,
                                "sample": "1 pair @512^2 (2.143 TFLOP) per run; 1 warm-up + 2 timed runs, %.1f s of CPU "
                                          "work in total" % secs}
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()
