"""Shared helpers for reading the reference-generated fixtures (see oracle/gen_golden.py)."""
import numpy as np
import torch

N_SAMPLE = 4096


def hash_str(s):
    h = 2166136261
    for ch in s.encode():
        h = ((h ^ ch) * 16777619) & 0xFFFFFFFF
    return h


def sample_idx(numel, key):
    g = torch.Generator().manual_seed(abs(hash_str(key)) % (2 ** 31))
    return torch.randint(0, numel, (min(N_SAMPLE, numel),), generator=g)


def check_summary(store, key, t, atol, rtol=0.0, stats_tol=None):
    """Compare tensor `t` with the stored subsample (+moments) of the reference's tensor `key`."""
    t = t.detach().to(torch.float32).cpu().contiguous()
    assert tuple(store[key + "/shape"]) == tuple(t.shape), (key, tuple(store[key + "/shape"]), tuple(t.shape))
    flat = t.reshape(-1)
    got = flat[sample_idx(flat.numel(), key)].numpy()
    want = store[key + "/sample"]
    err = np.abs(got - want)
    bound = atol + rtol * np.abs(want)
    assert (err <= bound).all(), "%s: max abs err %.3e (atol %.1e rtol %.1e)" % (key, err.max(), atol, rtol)
    if stats_tol is not None:
        st = np.array([flat.mean().item(), flat.std().item(), flat.abs().max().item()])
        assert np.allclose(st, store[key + "/stats"], atol=stats_tol, rtol=stats_tol), (key, st, store[key + "/stats"])
    return float(err.max())


def sub_state_dict(store, name):
    pre = name + "/sd/"
    return {k[len(pre):]: torch.from_numpy(store[k]) for k in store.files if k.startswith(pre)}
