"""Batch partitioning of independent (content, style) pairs across the GPUs of one box.

The inference path has NO data-path collective: every pair is independent, InstanceNorm is per
sample, weights are replicated.  One process per GPU; torch.distributed is used only for the rendezvous, the
barriers around the timed region and the max-over-ranks reduction of the elapsed time.
"""
import os

import torch
import torch.distributed as dist


def env_world():
    return int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))


def partition(n_items, world, rank):
    """Contiguous, balanced split of `n_items` work items: rank r gets items [lo, hi).  The first (n % world) ranks get
    one extra item; every item is owned by exactly one rank."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank %d/%d" % (rank, world))
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


def init(backend=None):
    """Join the process group described by the torchrun environment (no-op for a single process)."""
    world, rank, local_rank = env_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return world, rank, local_rank


def barrier():
    if dist.is_initialized():
        dist.barrier()


def max_over_ranks(value, device="cpu"):
    """The step time of a multi-GPU job is the slowest rank's."""
    if not dist.is_initialized():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device="cpu"):
    if not dist.is_initialized():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def swap_sharded(model, contents, styles, alpha=1.0):
    """Run this rank's share of the pair list.  contents/styles: [P,3,H,W] host or device tensors holding ALL pairs;
    returns (lo, hi, images[lo:hi])."""
    world, rank, _ = env_world()
    lo, hi = partition(contents.shape[0], world, rank)
    if hi == lo:
        return lo, hi, None
    out = model.swap(contents[lo:hi], styles[lo:hi], alpha=alpha)
    return lo, hi, out


def allreduce_grads_(params, world=None):
    """The gradient exchange of the training-step config (the reference wraps the model in
    torch.nn.parallel.DistributedDataParallel, models/__init__.py:88): every rank holds the gradients of its own batch
    shard; ONE flat f32 bucket, ONE all-reduce (NCCL over NVLink on the GPU box, gloo in the CPU tests), divided by the
    world size, scattered back into the .grad tensors in place.  Parameters without a gradient are skipped (the same
    set on every rank, or the buckets would not line up).  Returns the bucket size in bytes (0 for a single process)."""
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    if world <= 1 or not dist.is_initialized():
        return 0
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return 0
    flat = torch.cat([g.reshape(-1).to(torch.float32) for g in grads])
    if dist.get_backend() == "nccl":
        dist.all_reduce(flat, op=dist.ReduceOp.AVG)          # the division rides in the collective
    else:
        dist.all_reduce(flat)
        flat.div_(world)
    # scatter back with ONE multi-tensor copy instead of a launch per parameter (~250 per step)
    pieces = [v.view_as(g) for v, g in zip(flat.split([g.numel() for g in grads]), grads)]
    torch._foreach_copy_(grads, pieces)
    return flat.numel() * 4
