"""Where the gradient exchange's time goes: NCCL all-reduce (AVG) of the stage-1 bucket sizes vs the flatten / scatter copies.
    python -m torch.distributed.run --nproc-per-node N tools/allreduce_probe.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(rank)
dist.init_process_group("nccl")


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for mb in (115, 253):
    flat = torch.randn(mb * 1024 * 1024 // 4, device="cuda")
    parts = [torch.randn(n, device="cuda") for n in [flat.numel() // 250] * 250]
    t_ar = timed(lambda: dist.all_reduce(flat, op=dist.ReduceOp.AVG))
    t_cat = timed(lambda: torch.cat(parts))
    pieces = list(flat[:sum(p.numel() for p in parts)].split([p.numel() for p in parts]))
    t_copy = timed(lambda: torch._foreach_copy_(parts, pieces))
    half = flat.numel() // 2
    t_bf16 = timed(lambda: dist.all_reduce(flat[:half].view(torch.bfloat16), op=dist.ReduceOp.AVG))
    if rank == 0:
        print("%d MB f32, world %d: all_reduce %.3f ms (%.0f GB/s algbw), cat %.3f ms, foreach_copy %.3f ms; same count in bf16: %.3f ms"
              % (mb, dist.get_world_size(), t_ar, mb / 1024 / (t_ar * 1e-3), t_cat, t_copy, t_bf16))
dist.destroy_process_group()
