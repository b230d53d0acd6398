"""GPU parity of the image pre-/post-processing row (csrc/imageio.cu through the C ABI): bit-exact against the golden
vectors made by PIL / torchvision / the reference's tensor2im, against the numpy oracle, and against PIL run live."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import imageio_oracle as IO  # noqa: E402

DEV = "cuda"


def test_resize_bit_exact_golden_oracle_and_live_pil(golden):
    from ppst_b200 import imageio as P
    g = golden("imageio")
    for name in g["resize/names"]:
        img, (ow, oh) = g["resize/%s/in" % name], g["resize/%s/size" % name]
        got = P.resize_bicubic_u8(torch.from_numpy(img).to(DEV), (int(ow), int(oh))).cpu().numpy()
        assert np.array_equal(got, g["resize/%s/out" % name]), name
    rng = np.random.RandomState(3)
    # full dataset-size images (CelebAMask-HQ 1024^2 -> 512^2, a 2:3 portrait, an upscale), batched call included
    for (h, w), size in (((1024, 1024), (512, 512)), ((1024, 683), (512, 768)), ((200, 300), (768, 512)), ((64, 64), (64, 64))):
        img = rng.randint(0, 256, size=(h, w, 3)).astype(np.uint8)
        want = IO.resize_bicubic_u8(img, size)
        got = P.resize_bicubic_u8(torch.from_numpy(img).to(DEV), size).cpu().numpy()
        assert np.array_equal(got, want), (h, w, size)
        try:
            from PIL import Image
            assert np.array_equal(got, np.asarray(Image.fromarray(img).resize(size, Image.BICUBIC))), (h, w, size)
        except ImportError:
            pass
    batch = rng.randint(0, 256, size=(3, 90, 70, 3)).astype(np.uint8)
    got = P.resize_bicubic_u8(torch.from_numpy(batch).to(DEV), (35, 45)).cpu().numpy()
    for i in range(3):
        assert np.array_equal(got[i], IO.resize_bicubic_u8(batch[i], (35, 45)))
    with pytest.raises(RuntimeError):
        P.resize_bicubic_u8(torch.from_numpy(batch), (35, 45))               # CPU tensor: no fallback
    with pytest.raises(TypeError):
        P.resize_bicubic_u8(torch.from_numpy(batch).to(DEV).float(), (35, 45))


def test_totensor_normalize_and_tensor2im_bit_exact(golden):
    from ppst_b200 import imageio as P
    g = golden("imageio")
    for key in ("totensor/in", "totensor/all_in"):
        got = P.to_tensor_normalized(torch.from_numpy(g[key]).to(DEV))[0].cpu().numpy()
        assert np.array_equal(got, g[key.replace("in", "out")]), key
    x = torch.from_numpy(g["tensor2im/in"]).to(DEV)
    assert np.array_equal(P.tensor2im(x).cpu().numpy(), g["tensor2im/out"])
    big = (torch.rand(2, 3, 512, 512, generator=torch.Generator().manual_seed(5)) * 2.4 - 1.2)
    got = P.tensor2im(big.to(DEV)).cpu().numpy()
    for i in range(2):
        assert np.array_equal(got[i], IO.tensor2im(big[i].numpy()))
    assert np.array_equal(P.tensor2im(big.to(DEV), normalize=False).cpu().numpy()[0], IO.tensor2im(big[0].numpy(), normalize=False))
    # round trip on exactly representable pixel values: tensor2im(to_tensor(u8)) == u8 except where (v/255*2-1+1)/2*255
    # lands just below the integer (the reference's truncation) -- same as the oracle, value for value
    u8 = torch.arange(256, dtype=torch.uint8).view(1, 16, 16, 1).repeat(1, 1, 1, 3)
    back = P.tensor2im(P.to_tensor_normalized(u8.to(DEV))).cpu().numpy()[0]
    assert np.array_equal(back, IO.tensor2im(IO.to_tensor_normalized(u8[0].numpy())))
    assert np.abs(back.astype(int) - u8[0].numpy().astype(int)).max() <= 1


def test_load_images_matches_the_reference_transform():
    from ppst_b200 import imageio as P
    rng = np.random.RandomState(11)
    imgs = [rng.randint(0, 256, size=(300, 200, 3)).astype(np.uint8), rng.randint(0, 256, size=(1024, 683, 3)).astype(np.uint8)]
    got = P.load_images(imgs, load_size=512).cpu().numpy()
    assert got.shape == (2, 3, 768, 512) and got.dtype == np.float32
    for i, im in enumerate(imgs):
        assert np.array_equal(got[i], IO.load_transform(im, 512)), i
    with pytest.raises(ValueError):
        P.load_images([imgs[0], rng.randint(0, 256, size=(100, 100, 3)).astype(np.uint8)])
