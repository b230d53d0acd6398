"""Host-side logic of the multi-GPU path on CPU: pair partitioning and the gloo (world_size 2) barrier /
max-over-ranks plumbing that bench.py uses.  No kernels run here."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ppst_b200 import sharding


def test_partition_covers_every_pair_once():
    for n in (0, 1, 7, 8, 64, 65, 256):
        for world in (1, 2, 3, 4, 8):
            seen = []
            sizes = []
            for r in range(world):
                lo, hi = sharding.partition(n, world, r)
                assert 0 <= lo <= hi <= n
                seen.extend(range(lo, hi))
                sizes.append(hi - lo)
            assert seen == list(range(n))
            assert max(sizes) - min(sizes) <= 1
    assert sharding.partition(64, 8, 3) == (24, 32)           # BASELINE configs[1]: 64 pairs, 8 per GPU
    assert sharding.partition(256, 8, 7) == (224, 256)        # configs[2]: 16 x 16 grid cells, 32 per GPU
    with pytest.raises(ValueError):
        sharding.partition(8, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE=str(world), RANK=str(rank),
                      LOCAL_RANK=str(rank))
    w, r, lr = sharding.init(backend="gloo")
    assert (w, r) == (world, rank) and dist.get_backend() == "gloo"
    sharding.barrier()
    slow = sharding.max_over_ranks(10.0 + 5.0 * rank)          # step time = slowest rank
    total = sharding.sum_over_ranks(8)                          # pairs processed by the whole job

    class FakeModel:                                            # stands in for SwapModel on a CPU-only host
        def swap(self, c, s, alpha=1.0):
            return c * 2 + s

    contents = torch.arange(6, dtype=torch.float32).view(6, 1, 1, 1)
    styles = torch.ones(6, 1, 1, 1)
    lo, hi, out = sharding.swap_sharded(FakeModel(), contents, styles)
    gathered = [None] * world
    dist.all_gather_object(gathered, (lo, hi, out.flatten().tolist()))
    sharding.barrier()
    dist.destroy_process_group()
    q.put((rank, slow, total, gathered))


def test_two_rank_gloo_plumbing():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, slow, total, gathered in results:
        assert slow == 15.0 and total == 16.0
        rows = sorted(gathered)
        assert [(lo, hi) for lo, hi, _ in rows] == [(0, 3), (3, 6)]
        merged = [v for _, _, vals in rows for v in vals]
        assert merged == [2 * i + 1 for i in range(6)]          # identical to the single-process result


def _grad_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE=str(world), RANK=str(rank),
                      LOCAL_RANK=str(rank))
    sharding.init(backend="gloo")
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(5, 3)), torch.nn.Parameter(torch.zeros(7)), torch.nn.Parameter(torch.zeros(2))]
    params[0].grad = torch.full((5, 3), float(rank + 1))
    params[1].grad = torch.arange(7, dtype=torch.float32) * (rank + 1)
    # params[2] has no gradient on any rank: skipped
    nbytes = sharding.allreduce_grads_(params)
    dist.destroy_process_group()
    q.put((rank, nbytes, params[0].grad.flatten().tolist(), params[1].grad.tolist(), params[2].grad))


def test_two_rank_gradient_allreduce_is_one_averaged_flat_bucket():
    """Training-step config: gradients averaged over ranks through one flat bucket."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, nbytes, g0, g1, g2 in results:
        assert nbytes == (15 + 7) * 4
        assert g0 == [1.5] * 15                                               # mean of 1 and 2
        assert g1 == [1.5 * i for i in range(7)]
        assert g2 is None
    assert sharding.allreduce_grads_([torch.nn.Parameter(torch.zeros(1))]) == 0   # single process: no-op


def _overlap_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE=str(world), RANK=str(rank),
                      LOCAL_RANK=str(rank))
    sharding.init(backend="gloo")

    def make():
        torch.manual_seed(3)
        return torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.Tanh(), torch.nn.Linear(16, 16), torch.nn.Tanh(),
                                   torch.nn.Linear(16, 4)).double()
    torch.manual_seed(100 + rank)
    x = torch.randn(5, 6, dtype=torch.float64)

    def loss_of(net):
        # the first layer is used TWICE: its gradient must be reported once, after both uses are accumulated
        return (net(x) ** 2).mean() + net[0](x).sum() * 0.1
    ref = make()
    loss_of(ref).backward()
    sharding.allreduce_grads_(list(ref.parameters()))
    net = make()
    red = sharding.OverlappedGradReducer(list(net.parameters()), bucket_bytes=16 * 16 * 4)   # several small buckets
    order = []
    for i, p in enumerate(net.parameters()):
        p.register_post_accumulate_grad_hook(lambda p_, i=i: order.append(i))
    with red:
        loss_of(net).backward()
    same = all(torch.equal(a.grad, b.grad) for a, b in zip(net.parameters(), ref.parameters()))
    # a second step reuses the reducer; a backward outside `with` is left alone
    net.zero_grad(set_to_none=True)
    loss_of(net).backward()
    local = make()
    loss_of(local).backward()                                   # this rank's own gradients, never exchanged
    untouched = all(torch.equal(a.grad, b.grad) for a, b in zip(net.parameters(), local.parameters()))
    with red:
        net.zero_grad(set_to_none=True)
        loss_of(net).backward()
    same2 = all(torch.equal(a.grad, b.grad) for a, b in zip(net.parameters(), ref.parameters()))
    dist.destroy_process_group()
    q.put((rank, same, same2, untouched, red.buckets, red.bytes_reduced, order))


def test_two_rank_overlapped_gradient_reducer_matches_the_flat_allreduce():
    """OverlappedGradReducer: buckets in autograd order, each averaged as soon as it is full; the result equals the
    one-bucket all-reduce bit for bit, parameters used twice are reported once, the bucket sequence is the same on both
    ranks (or the collectives would not line up and this test would hang / mismatch)."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_overlap_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n_param_bytes = (6 * 16 + 16 + 16 * 16 + 16 + 16 * 4 + 4) * 4
    for rank, same, same2, untouched, buckets, nbytes, order in results:
        assert same and same2 and untouched
        assert buckets >= 2 and nbytes == n_param_bytes
        assert sorted(order[:6]) == list(range(6)) and order[0] in (4, 5)    # first backward; last layer first: autograd order
    assert results[0][6] == results[1][6]
    red = sharding.OverlappedGradReducer([torch.nn.Parameter(torch.zeros(1))])     # single process: inert
    assert not red.active
    with red:
        pass
