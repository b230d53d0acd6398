"""CPU: the image pre-/post-processing oracle (oracle/imageio_oracle.py) against the real libraries the reference calls
(PIL, run here) and the committed golden vectors; the product's host-side coefficient tables against the oracle's."""
import numpy as np
import pytest

from oracle import imageio_oracle as IO


def test_resize_oracle_matches_golden_and_live_pil(golden):
    g = golden("imageio")
    for name in g["resize/names"]:
        img, (ow, oh) = g["resize/%s/in" % name], g["resize/%s/size" % name]
        got = IO.resize_bicubic_u8(img, (int(ow), int(oh)))
        assert np.array_equal(got, g["resize/%s/out" % name]), name          # bit-exact (integer arithmetic)
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.RandomState(7)
    for (h, w), size in (((23, 41), (64, 48)), ((80, 60), (17, 33)), ((33, 33), (33, 16)), ((5, 4), (11, 9))):
        img = rng.randint(0, 256, size=(h, w, 3)).astype(np.uint8)
        want = np.asarray(Image.fromarray(img).resize(size, Image.BICUBIC))
        assert np.array_equal(IO.resize_bicubic_u8(img, size), want), (h, w, size)


def test_host_coefficient_tables_match_the_oracle():
    from ppst_b200 import imageio as P                                      # host-side logic only: no GPU call here
    for in_size, out_size in ((53, 20), (48, 128), (1024, 768), (683, 512), (7, 2), (512, 512), (300, 301), (2, 9)):
        ks, bounds, kk = P._bicubic_tables(in_size, out_size)
        ks2, bounds2, kk2 = IO.precompute_coeffs(in_size, out_size)
        assert ks == ks2 and np.array_equal(bounds, bounds2) and np.array_equal(kk, kk2), (in_size, out_size)
    assert P.scale_shortside_size(683, 1024, 512) == IO.scale_shortside_size(683, 1024, 512) == (512, 768)
    assert P.make_power_2_size(500, 333) == IO.make_power_2_size(500, 333) == (496, 336)


def test_totensor_and_tensor2im_oracle_match_golden(golden):
    g = golden("imageio")
    assert np.array_equal(IO.to_tensor_normalized(g["totensor/in"]), g["totensor/out"])
    assert np.array_equal(IO.to_tensor_normalized(g["totensor/all_in"]), g["totensor/all_out"])
    assert np.array_equal(IO.tensor2im(g["tensor2im/in"]), g["tensor2im/out"])
