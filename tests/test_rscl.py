"""Region style contrastive loss (networks/rscl.py) and its queues: the oracle against the reference-run fixture, the product
module against the oracle, and the fused queue update (ONE all_gather for the 24 keys of an iteration) against the reference's
24 sequential all_gathers on two gloo ranks.  CPU only."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import rscl_oracle as RO
from ppst_b200.rscl import RsclLoss


def _load_queues(crit):
    for layer in range(4):
        getattr(crit, "queue_data_A%d" % layer).copy_(RO.make_queue(layer))
        getattr(crit, "queue_ptr_A%d" % layer).zero_()
    return crit


def test_oracle_and_module_match_the_reference_fixture(golden):
    g = golden("rscl")
    crit = _load_queues(RsclLoss(nce_T=0.07))
    assert sorted(crit.state_dict()) == sorted(["queue_data_A%d" % i for i in range(4)] + ["queue_ptr_A%d" % i for i in range(4)])
    for layer in range(4):
        q, k, k0 = RO.make_feats(3 * layer), RO.make_feats(3 * layer + 1), RO.make_feats(3 * layer + 2)
        queue, ptr = RO.make_queue(layer), 0
        for want, kk0 in ((g["loss/%d" % layer], None), (g["loss_k0/%d" % layer], k0)):
            assert abs(RO.rscl_loss(q, k, queue, kk0).item() - float(want)) <= 1e-5 * max(1.0, abs(float(want)))
            assert abs(crit(q, k, kk0, layer).item() - float(want)) <= 1e-6 * max(1.0, abs(float(want)))
        items = [(k[j:j + 1], layer) for j in range(3)] + [(k0[j:j + 1], layer) for j in range(3)]
        for rows, _ in items:
            queue, ptr = RO.enqueue(queue, ptr, rows)
        crit.dequeue_and_enqueue_many(items)                                   # one call instead of six
        want = float(g["loss_after/%d" % layer])
        assert abs(RO.rscl_loss(q, k, queue, k0).item() - want) <= 1e-5 * max(1.0, abs(want))
        assert abs(crit(q, k, k0, layer).item() - want) <= 1e-6 * max(1.0, abs(want))
        qd = getattr(crit, "queue_data_A%d" % layer)
        assert np.array_equal(qd[:, :8].numpy(), g["queue_head/%d" % layer]) and torch.equal(qd, queue)
        assert abs(qd.double().sum().item() - float(g["queue_sum/%d" % layer])) <= 1e-9
        assert int(getattr(crit, "queue_ptr_A%d" % layer)) == int(g["ptr/%d" % layer]) == ptr == 6


def test_loss_is_differentiable_in_the_query_only_through_the_logits():
    crit = _load_queues(RsclLoss())
    q = RO.make_feats(1).requires_grad_(True)
    k = RO.make_feats(2)
    crit(q, k, RO.make_feats(3), 2).backward()
    assert q.grad is not None and torch.isfinite(q.grad).all() and float(q.grad.abs().max()) > 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE=str(world), RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    seq, fused = _load_queues(RsclLoss(world_size=world)), _load_queues(RsclLoss(world_size=world))
    items = []
    for layer in range(4):
        k, k0 = RO.make_feats(10 * rank + 2 * layer), RO.make_feats(10 * rank + 2 * layer + 1)       # rank-dependent keys
        items += [(k[j:j + 1], layer) for j in range(3)] + [(k0[j:j + 1], layer) for j in range(3)]
    for rows, layer in items:                             # the reference: 24 all_gathers of one [1, 2048] row
        seq.dequeue_and_enqueue(rows, layer)
    nbytes = fused.dequeue_and_enqueue_many(items)        # here: one
    same = all(torch.equal(getattr(seq, "queue_data_A%d" % i), getattr(fused, "queue_data_A%d" % i)) and
               int(getattr(seq, "queue_ptr_A%d" % i)) == int(getattr(fused, "queue_ptr_A%d" % i)) == 6 * world for i in range(4))
    # every rank must hold the same queues, rank 0's key before rank 1's
    head = getattr(fused, "queue_data_A0")[:, :2].T.contiguous()
    want = torch.cat([RO.make_feats(0)[0:1], RO.make_feats(10)[0:1]], 0)
    dist.destroy_process_group()
    out.put((rank, bool(same), int(nbytes), bool(torch.equal(head, want))))


def test_fused_queue_update_equals_sequential_all_gathers_on_two_ranks():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, same, nbytes, order_ok in res:
        assert same and order_ok, rank
        assert nbytes == world * 24 * 2048 * 4
