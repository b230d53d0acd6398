"""ppst_b200.synthetic (product side: bench.py / tools) and the oracle's generators produce the same tensors."""
import torch

from oracle import ppst_oracle as O
from ppst_b200 import synthetic as S


def test_networks_inputs_and_noises_are_bit_identical():
    for a, b in zip(S.make_networks(0), O.make_networks(0)):
        assert a.keys() == b.keys()
        assert all(torch.equal(a[k], b[k]) for k in a)
    d1, d2 = S.make_discriminator(256, seed=1), O.make_discriminator(256, seed=1)
    assert d1.keys() == d2.keys() and all(torch.equal(d1[k], d2[k]) for k in d1)
    assert torch.equal(S.make_portraits(2, 64, 48, seed=3), O.make_portraits(2, 64, 48, seed=3))
    assert torch.equal(S.make_portraits(2, 64, 48, seed=3, smooth=True), O.make_portraits(2, 64, 48, seed=3, smooth=True))
    assert torch.equal(S.make_label_map(2, 32, 32), O.make_label_map(2, 32, 32))
    n1, n2 = S.make_noises(1, 64, 64), O.make_noises(1, 64, 64)
    assert n1.keys() == n2.keys() and all(torch.equal(n1[k], n2[k]) for k in n1)
