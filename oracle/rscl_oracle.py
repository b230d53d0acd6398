"""CPU restatement of the region style contrastive loss and its queues (networks/rscl.py:17-90).

TEST INFRASTRUCTURE ONLY -- never imported by the product package.  Plain functions over explicit queue tensors; pinned by
tests/golden/rscl.npz, made by oracle/gen_golden.py (`rscl`) from the unmodified reference module.
"""
import torch
import torch.nn.functional as F


def rscl_loss(feat_q, feat_k, queue, feat_k0=None, nce_T=0.07):
    """rsclLoss.forward, rscl.py:42-66.  queue: [2048, 128] (the level's memory).  The positive logit is q.k per row; the
    negatives are (a) the n x n current-batch block, which upstream masks ENTIRELY to -10 (`torch.eye(feat_q.size(0))`
    after `feat_q.view(1, -1, 2048)` is eye(1), broadcast), and (b) the queue columns [+ feat_k0 rows]."""
    n = feat_q.shape[0]
    l_pos = (feat_q * feat_k).sum(dim=1, keepdim=True)
    if feat_k0 is not None:
        queue = torch.cat((queue, feat_k0.T), dim=1)
    l_neg2 = feat_q @ queue
    l_neg1 = torch.full((n, n), -10.0, dtype=feat_q.dtype)
    logits = torch.cat((l_pos, l_neg1, l_neg2), dim=1) / nce_T
    return (torch.logsumexp(logits, dim=1) - logits[:, 0]).mean()          # cross entropy with target 0


def enqueue(queue, ptr, keys):
    """rsclLoss.dequeue_and_enqueue, rscl.py:68-90 (single process): returns (queue, ptr) after writing keys [b, 2048]."""
    b = keys.shape[0]
    assert queue.shape[1] % b == 0
    queue = queue.clone()
    queue[:, ptr:ptr + b] = keys.T
    return queue, (ptr + b) % queue.shape[1]


def make_queue(layer, dim=2048, size=128):
    """Deterministic unit-norm queue columns (no RNG stream involved, so fixtures stay a few numbers)."""
    i = torch.arange(dim * size, dtype=torch.float64).view(dim, size)
    return F.normalize(torch.sin(i * 0.37 + 1.3 * layer) + 0.1 * torch.cos(i * 0.011), dim=0).to(torch.float32)


def make_feats(tag, rows=6, dim=2048):
    """Deterministic unit-norm feature rows [rows, 2048]."""
    i = torch.arange(rows * dim, dtype=torch.float64).view(rows, dim)
    return F.normalize(torch.sin(i * 0.113 + 0.7 * tag) + 0.3 * torch.cos(i * 0.0071 * (tag + 1)), dim=1).to(torch.float32)
