"""Ordered kernel list of ONE CUDA-graph replay of the swap step (CUPTI activity records through torch.profiler's
chrome trace): start offset, duration, grid, block, dynamic smem, kernel name -- in launch order, so every pass of the
engine can be attributed.   python tools/seq_step.py [pairs] > profiles/rNN_seq_step.txt"""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402
import ppst_b200  # noqa: E402
from ppst_b200 import synthetic as S  # noqa: E402

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 8
m = ppst_b200.SwapModel(*S.make_networks(0), mode="bf16")
m.noises = {k: v.expand(pairs, -1, -1, -1).contiguous() for k, v in S.make_noises(1, 512, 512).items()}
c = S.make_portraits(pairs, 512, 512, seed=1).cuda()
s = S.make_portraits(pairs, 512, 512, seed=2, smooth=True).cuda()
for _ in range(4):
    m.swap_graph(c, s)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    m.swap_graph(c, s)
    torch.cuda.synchronize()
path = os.path.join(tempfile.mkdtemp(), "trace.json")
prof.export_chrome_trace(path)
evs = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") == "kernel"]
evs.sort(key=lambda e: e["ts"])
t0 = evs[0]["ts"]
span = evs[-1]["ts"] + evs[-1]["dur"] - t0
print("one graph replay: %d kernels, span %.3f ms, busy %.3f ms" % (len(evs), span / 1e3, sum(e["dur"] for e in evs) / 1e3))
for i, e in enumerate(evs):
    a = e.get("args", {})
    name = e["name"].replace("ppst::", "").replace("void ", "")
    name = name.split("(")[0][:64]
    print("%4d %9.1f %8.1f  g%-14s b%-5s sm%-6s %s" % (i, e["ts"] - t0, e["dur"], "x".join(map(str, a.get("grid", []))),
                                                      a.get("block", ["?"])[0], a.get("shared memory", "?"), name))
