"""Region style contrastive loss of the training step with its memory queues (networks/rscl.py:17-90, used by
models/ppst_model.py:203-224), and the queue update as ONE collective.

The reference enqueues 6 keys per level x 4 levels per iteration, each through its own `torch.distributed.all_gather` of one
[1, 2048] row (24 collectives of 8 KB, rscl.py:9-15,69-72).  `dequeue_and_enqueue_many` gathers all keys of the iteration in
one all_gather ([K, 2048] per rank) and writes the queues in the reference's order, so the buffers end up bit-identical to
the sequential calls.  The loss itself is [n, 2048]-sized arithmetic (n <= 3 x batch rows): it stays a torch expression, kept
operation for operation -- including the upstream quirk that the "current batch" negatives are ALL masked to -10
(`torch.eye(feat_q.size(0))` is evaluated after the view to [1, n, 2048], i.e. eye(1), broadcast over the n x n block).

Buffers keep the reference's names (queue_data_A0..3 [2048, 128], queue_ptr_A0..3), so a checkpointed criterionNCE
state_dict loads unchanged.
"""
import torch
import torch.distributed as dist
import torch.nn.functional as F
from torch import nn


class RsclLoss(nn.Module):
    QUEUE_SIZE = 128
    DIM = 2048
    LEVELS = 4

    def __init__(self, nce_T=0.07, world_size=None, generator=None):
        super().__init__()
        self.nce_T = float(nce_T)
        self.world_size = world_size
        for i in range(self.LEVELS):                       # rscl.py:24-40: unit-norm random columns, pointer 0
            data = F.normalize(torch.randn(self.DIM, self.QUEUE_SIZE, generator=generator), dim=0)
            self.register_buffer("queue_data_A%d" % i, data)
            self.register_buffer("queue_ptr_A%d" % i, torch.zeros(1, dtype=torch.long))

    # -- loss (rscl.py:42-66) -------------------------------------------------------------------------------------------
    def forward(self, feat_q, feat_k, feat_k0=None, layer=-1):
        assert 0 <= layer < self.LEVELS, layer
        l_pos = torch.einsum("nc,nc->n", feat_q, feat_k).unsqueeze(-1)
        queue = getattr(self, "queue_data_A%d" % layer).clone().detach()
        if feat_k0 is not None:
            queue = torch.cat((queue, feat_k0.T), dim=1)
        l_neg2 = torch.einsum("nc,ck->nk", feat_q, queue)
        q3 = feat_q.view(1, -1, self.DIM).contiguous()
        k3 = feat_k.view(1, -1, self.DIM).contiguous()
        l_neg_cur = torch.bmm(q3, k3.transpose(2, 1))
        diagonal = torch.eye(q3.size(0), device=q3.device, dtype=torch.bool)[None, :, :]      # eye(1): masks the whole block
        l_neg_cur = l_neg_cur.masked_fill(diagonal, -10.0)
        l_neg1 = l_neg_cur.view(-1, q3.size(1))
        logits = torch.cat((l_pos, l_neg1, l_neg2), dim=1)
        labels = torch.zeros(logits.size(0), dtype=torch.long, device=logits.device)
        return F.cross_entropy(logits / self.nce_T, labels)

    # -- queues (rscl.py:68-90) ----------------------------------------------------------------------------------------
    def _world(self):
        if not dist.is_initialized():
            return 1
        return self.world_size or dist.get_world_size()

    def _write(self, rows, layer):
        """rows [b, 2048] -> columns [ptr, ptr + b) of level `layer`'s queue; pointer advances modulo the queue size."""
        b = rows.size(0)
        assert self.QUEUE_SIZE % b == 0, b
        ptr_t = getattr(self, "queue_ptr_A%d" % layer)
        ptr = int(ptr_t)
        getattr(self, "queue_data_A%d" % layer)[:, ptr:ptr + b] = rows.T
        ptr_t[0] = (ptr + b) % self.QUEUE_SIZE

    @torch.no_grad()
    def dequeue_and_enqueue(self, keys, layer=-1):
        """The reference's call: one all_gather per call."""
        if dist.is_initialized():
            world = self._world()
            parts = [torch.empty_like(keys) for _ in range(world)]
            dist.all_gather(parts, keys.contiguous())
            keys = torch.cat(parts, dim=0)
        self._write(keys, layer)

    @torch.no_grad()
    def dequeue_and_enqueue_many(self, items):
        """items: [(keys [b_i, 2048], layer_i), ...] in the order the reference would call dequeue_and_enqueue.  ONE
        all_gather for all of them; the queues end up exactly as after the sequential calls."""
        if not items:
            return 0
        sizes = [k.size(0) for k, _ in items]
        flat = torch.cat([k.reshape(-1, self.DIM) for k, _ in items], dim=0).contiguous()
        nbytes = 0
        if dist.is_initialized():
            world = self._world()
            gathered = torch.empty((world,) + tuple(flat.shape), device=flat.device, dtype=flat.dtype)
            dist.all_gather_into_tensor(gathered.view(-1, self.DIM), flat) if dist.get_backend() == "nccl" else \
                self._gather_list(gathered, flat, world)
            nbytes = gathered.numel() * gathered.element_size()
        else:
            gathered = flat.unsqueeze(0)
        off = 0
        for (_, layer), b in zip(items, sizes):
            rows = gathered[:, off:off + b, :].reshape(-1, self.DIM)       # rank-major, like torch.cat of the gathered list
            self._write(rows, layer)
            off += b
        return nbytes

    @staticmethod
    def _gather_list(gathered, flat, world):
        parts = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(parts, flat)
        for r in range(world):
            gathered[r].copy_(parts[r])
