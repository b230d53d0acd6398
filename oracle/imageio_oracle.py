"""CPU restatement (numpy, integer arithmetic) of the image pre-/post-processing around the swap path.

TEST INFRASTRUCTURE ONLY -- imported by tests/ (never by ppst_b200/).

What it follows:
  * data/base_dataset.py:85-172 (get_transform: __scale_shortside / __make_power_2 -> PIL Image.resize(BICUBIC),
    ToTensor, Normalize(0.5, 0.5)) and util/util.py:98-131 (tensor2im).
  * The resize itself is a THIRD-PARTY algorithm: Pillow (requirements.txt pins Pillow==10.4.0), src/libImaging/Resample.c
    -- bicubic_filter (a = -0.5, support 2), precompute_coeffs, normalize_coeffs_8bpc (PRECISION_BITS = 22),
    ImagingResampleHorizontal_8bpc / ImagingResampleVertical_8bpc (horizontal pass first, uint8 intermediate).
    Pinned by running it: tests/test_imageio_oracle.py compares this restatement with the installed PIL (12.2 here; the
    8-bit resampler is unchanged since 10.x) on seeded images and against tests/golden/imageio.npz.
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def bicubic_filter(x):
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def precompute_coeffs(in_size, out_size, support=2.0, filt=bicubic_filter):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc with box = (0, in_size) -> (ksize, bounds [out,2], kk int32)."""
    scale = float(np.float32(in_size) - np.float32(0.0)) / out_size
    filterscale = max(scale, 1.0)
    sup = support * filterscale
    ksize = int(math.ceil(sup)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - sup + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + sup + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [filt((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            k = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + k * (1 << PRECISION_BITS)) if k < 0 else int(0.5 + k * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return ksize, bounds, kk


def _pass(img, out_len, axis):
    """One resampling pass along axis 1 (x) or 0 (y) of a uint8 [h, w, c] image."""
    in_len = img.shape[axis]
    _, bounds, kk = precompute_coeffs(in_len, out_len)
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((out_len,) + src.shape[1:], np.uint8)
    for o in range(out_len):
        lo, cnt = bounds[o]
        ss = (1 << (PRECISION_BITS - 1)) + np.tensordot(kk[o, :cnt].astype(np.int64), src[lo:lo + cnt], axes=(0, 0))
        out[o] = np.clip(ss >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resize_bicubic_u8(img, size):
    """PIL Image.resize(size=(w, h), Image.BICUBIC) of a uint8 [h, w, c] array (horizontal pass first; a pass whose
    length does not change is skipped, as in ImagingResample)."""
    ow, oh = size
    out = img
    if ow != img.shape[1]:
        out = _pass(out, ow, 1)
    if oh != img.shape[0]:
        out = _pass(out, oh, 0)
    return out.copy() if out is img else out


def scale_shortside_size(w, h, load_size):
    """base_dataset.py:168-172."""
    scale = load_size / min(w, h)
    return round(w * scale), round(h * scale)


def make_power_2_size(w, h, base=16):
    """base_dataset.py:150-158."""
    return int(round(w / base) * base), int(round(h / base) * base)


def load_transform(img, load_size=512, base=16):
    """get_transform(preprocess='scale_shortside') without the random parts: resize, make_power_2, ToTensor, Normalize."""
    h, w = img.shape[:2]
    img = resize_bicubic_u8(img, scale_shortside_size(w, h, load_size))
    h, w = img.shape[:2]
    w2, h2 = make_power_2_size(w, h, base)
    if (w2, h2) != (w, h):
        img = resize_bicubic_u8(img, (w2, h2))
    return to_tensor_normalized(img)


def to_tensor_normalized(img, mean=0.5, std=0.5):
    """transforms.ToTensor() + Normalize(mean, std): uint8 [h, w, c] -> float32 [c, h, w], every step in float32."""
    x = img.astype(np.float32) / np.float32(255.0)
    x = (x - np.float32(mean)) / np.float32(std)
    return np.ascontiguousarray(x.transpose(2, 0, 1))


def tensor2im(x, normalize=True):
    """util/util.py:120-128 for one image: float32 [c, h, w] -> uint8 [h, w, c]."""
    x = np.asarray(x, np.float32)
    if normalize:
        y = (np.transpose(x, (1, 2, 0)) + 1) / 2.0 * 255.0
    else:
        y = np.transpose(x, (1, 2, 0)) * 255.0
    return np.clip(y, 0, 255).astype(np.uint8)
