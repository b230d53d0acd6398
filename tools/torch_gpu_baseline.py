"""Context number, NOT a bench arm: the oracle's op sequence (the reference's own PyTorch graph, restated) executed by
stock PyTorch / cuDNN / cuBLAS on the same B200 -- what recompiling the reference for sm_100 would give -- next to
ppst_b200 on identical inputs.  8 (content, style) pairs @512^2 per step.
    python tools/torch_gpu_baseline.py > profiles/rNN_torch_gpu_baseline.json"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from oracle import ppst_oracle as O  # noqa: E402

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = "cuda"
sds = [{k: v.to(dev) for k, v in sd.items()} for sd in O.make_networks(0)]
c = O.make_portraits(pairs, 512, 512, seed=1234).to(dev)
s = O.make_portraits(pairs, 512, 512, seed=4321, smooth=True).to(dev)
noises = {k: v.expand(pairs, -1, -1, -1).contiguous().to(dev) for k, v in O.make_noises(1, 512, 512).items()}


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


out = {}
with torch.no_grad():
    for name, tf32 in (("fp32 (TF32 off)", False), ("fp32 with TF32 convolutions / matmuls", True)):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.benchmark = True
        try:
            ms = timed(lambda: O.swap_pair(sds[0], sds[1], sds[2], c, s, noises=noises))
            out[name] = {"ms_per_step": ms, "images_per_s": pairs / ms * 1e3}
        except Exception as e:  # noqa: BLE001
            out[name] = {"error": repr(e)[:300]}
    try:
        with torch.autocast("cuda", dtype=torch.bfloat16):
            ms = timed(lambda: O.swap_pair(sds[0], sds[1], sds[2], c, s, noises=noises))
        out["bf16 autocast (cuDNN / cuBLAS tensor cores)"] = {"ms_per_step": ms, "images_per_s": pairs / ms * 1e3}
    except Exception as e:  # noqa: BLE001
        out["bf16 autocast (cuDNN / cuBLAS tensor cores)"] = {"error": repr(e)[:300]}

import ppst_b200  # noqa: E402
m = ppst_b200.SwapModel(*O.make_networks(0), mode="bf16")
m.noises = noises
ms = timed(lambda: m.swap_graph(c, s), reps=10)
out["ppst_b200 bf16 (this repo, CUDA-graph replay)"] = {"ms_per_step": ms, "images_per_s": pairs / ms * 1e3}
print(json.dumps({"what": "stock PyTorch on the B200 running the oracle's op sequence vs ppst_b200, %d pairs @512^2 per "
                          "step, inputs resident on the device" % pairs, "torch": torch.__version__, "results": out}, indent=1))
