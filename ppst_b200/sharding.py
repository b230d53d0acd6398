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


def _reduce_bucket_(grads, world):
    """Average one bucket of gradient tensors over the ranks, in place (one flat f32 buffer, one all-reduce)."""
    flat = torch.cat([g.reshape(-1).to(torch.float32) for g in grads])
    if dist.get_backend() == "nccl":
        dist.all_reduce(flat, op=dist.ReduceOp.AVG)
    else:
        dist.all_reduce(flat)
        flat.div_(world)
    torch._foreach_copy_(grads, [v.view_as(g) for v, g in zip(flat.split([g.numel() for g in grads]), grads)])
    return flat.numel() * 4


class OverlappedGradReducer:
    """allreduce_grads_ with the exchange hidden behind the rest of the backward pass (what DistributedDataParallel's
    bucketing does for the reference, models/__init__.py:88): every parameter reports the moment autograd has finished
    its gradient (post-accumulate hook: once per backward() however often the parameter is used); parameters are
    collected in that order and every `bucket_bytes` of them are averaged by one all-reduce on a SIDE stream that waits
    for the main stream at that point only.  The networks here keep most of their parameters in the low-resolution,
    512-channel layers, whose gradients come early in the backward pass of D, E1 and E2 (and, for G, before the whole
    encoder backward), while the time goes into the high-resolution layers -- so all but the last bucket travel while
    the tensor cores are busy.  Deterministic: the bucket sequence is the autograd order, identical on every rank.
    Works inside CUDA-graph capture (the fork / join are event dependencies of the captured graph).

        red = OverlappedGradReducer(params)       # once
        with red:                                  # per step: arm, backward, flush the tail, join
            loss.backward()
        optimiser.step()
    """

    def __init__(self, params, world=None, bucket_bytes=32 << 20):
        if world is None:
            world = dist.get_world_size() if dist.is_initialized() else 1
        self.world = world
        self.active = world > 1 and dist.is_initialized()
        self.bucket_bytes = int(bucket_bytes)
        self.params = [p for p in params]
        self._armed = False
        self._pending, self._pending_bytes = [], 0
        self._side = None
        self.bytes_reduced = 0
        self.buckets = 0
        self._hooks = [p.register_post_accumulate_grad_hook(self._ready) for p in self.params] if self.active else []

    def _ready(self, p):
        if not self._armed or p.grad is None:
            return
        self._pending.append(p.grad)
        self._pending_bytes += p.grad.numel() * 4
        if self._pending_bytes >= self.bucket_bytes:
            self._flush()

    def _flush(self):
        grads, self._pending, self._pending_bytes = self._pending, [], 0
        if not grads:
            return
        self.buckets += 1
        if grads[0].is_cuda:
            main = torch.cuda.current_stream(grads[0].device)
            if self._side is None:
                self._side = torch.cuda.Stream(grads[0].device)
            ready = torch.cuda.Event()
            ready.record(main)
            self._side.wait_event(ready)                      # the side stream sees these gradients, nothing later
            with torch.cuda.stream(self._side):
                self.bytes_reduced += _reduce_bucket_(grads, self.world)
        else:
            self.bytes_reduced += _reduce_bucket_(grads, self.world)

    def __enter__(self):
        self._armed = self.active
        self._pending, self._pending_bytes = [], 0
        self.bytes_reduced, self.buckets = 0, 0
        return self

    def __exit__(self, exc_type, exc, tb):
        if self._armed and exc_type is None:
            self._flush()
            if self._side is not None:
                torch.cuda.current_stream(self._side.device).wait_stream(self._side)
        self._armed = False
        return False

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []
