"""Image pre- / post-processing around the swap path on the device.

The reference feeds the networks through PIL + torchvision on the host, one image at a time
(data/base_dataset.py:85-172: __scale_shortside -> Image.resize(BICUBIC), __make_power_2, ToTensor, Normalize(0.5, 0.5))
and turns results back into uint8 with util.tensor2im (util/util.py:98-131).  Here decoded uint8 images go to the GPU
once and stay there: Pillow's fixed-point 8-bit bicubic resampler, the normalisation and tensor2im are C-ABI kernels
(csrc/imageio.cu), bit-exact against PIL / torchvision / numpy.  The per-axis coefficient tables are built on the host
exactly as Pillow's precompute_coeffs / normalize_coeffs_8bpc build them (a few thousand doubles) and cached.
"""
import functools
import math

import numpy as np
import torch

from ._lib import check, load as _L, ptr, require_cuda, stream

PRECISION_BITS = 32 - 8 - 2


@functools.lru_cache(maxsize=256)
def _bicubic_tables(in_size, out_size):
    """Pillow Resample.c precompute_coeffs (bicubic, a = -0.5, support 2, box = whole axis) + normalize_coeffs_8bpc."""
    scale = float(np.float32(in_size)) / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    xx = np.arange(out_size, dtype=np.float64)
    center = (xx + 0.5) * scale
    xmin = np.maximum((center - support + 0.5).astype(np.int64), 0)            # C (int) cast: values are >= -support
    xmin = np.where(center - support + 0.5 < 0, 0, xmin)
    xmax = np.minimum((center + support + 0.5).astype(np.int64), in_size) - xmin
    x = np.arange(ksize, dtype=np.float64)[None, :]
    arg = np.abs((x + xmin[:, None] - center[:, None] + 0.5) * (1.0 / filterscale))
    a = -0.5
    w = np.where(arg < 1.0, ((a + 2.0) * arg - (a + 3.0)) * arg * arg + 1,
                 np.where(arg < 2.0, (((arg - 5) * arg + 8) * arg - 4) * a, 0.0))
    valid = x < xmax[:, None]
    w = np.where(valid, w, 0.0)
    ww = np.zeros(out_size, np.float64)
    for j in range(ksize):                                                    # left-to-right sum, like the C loop
        ww = ww + w[:, j]
    k = np.where(ww[:, None] != 0.0, w / np.where(ww == 0.0, 1.0, ww)[:, None], w)
    scaled = k * float(1 << PRECISION_BITS)
    kk = np.where(k < 0, np.trunc(-0.5 + scaled), np.trunc(0.5 + scaled)).astype(np.int32)
    kk = np.where(valid, kk, 0).astype(np.int32)
    bounds = np.stack([xmin, xmax], 1).astype(np.int32)
    return ksize, bounds, kk


@functools.lru_cache(maxsize=256)
def _device_tables(in_size, out_size, device_index):
    ksize, bounds, kk = _bicubic_tables(in_size, out_size)
    dev = torch.device("cuda", device_index)
    return ksize, torch.from_numpy(bounds).to(dev).contiguous(), torch.from_numpy(kk).to(dev).contiguous()


def resize_bicubic_u8(img, size):
    """PIL `Image.resize(size, Image.BICUBIC)` of uint8 images on the device.
    img: CUDA uint8 [h, w, c] or [n, h, w, c]; size = (width, height) as in PIL.  Returns the same rank."""
    require_cuda(img)
    if img.dtype != torch.uint8:
        raise TypeError("resize_bicubic_u8: uint8 images expected, got %s" % img.dtype)
    squeeze = img.dim() == 3
    x = (img[None] if squeeze else img).contiguous()
    n, h, w, c = x.shape
    ow, oh = int(size[0]), int(size[1])
    dev = x.device.index if x.device.index is not None else torch.cuda.current_device()
    if ow != w:
        ksize, bounds, kk = _device_tables(w, ow, dev)
        y = torch.empty((n, h, ow, c), device=x.device, dtype=torch.uint8)
        check(_L().ppst_resample_u8(ptr(y), ptr(x), ptr(bounds), ptr(kk), n, h, w, ow, c, ksize, 0, stream()))
        x, w = y, ow
    if oh != h:
        ksize, bounds, kk = _device_tables(h, oh, dev)
        y = torch.empty((n, oh, w, c), device=x.device, dtype=torch.uint8)
        check(_L().ppst_resample_u8(ptr(y), ptr(x), ptr(bounds), ptr(kk), n, h, w, oh, c, ksize, 1, stream()))
        x = y
    if x.data_ptr() == img.data_ptr():
        x = x.clone()
    return x[0] if squeeze else x


def scale_shortside_size(w, h, load_size):
    """data/base_dataset.py:168-172 (__scale_shortside): (w, h) after scaling the short side to load_size."""
    scale = load_size / min(w, h)
    return round(w * scale), round(h * scale)


def make_power_2_size(w, h, base=16):
    """data/base_dataset.py:150-158 (__make_power_2)."""
    return int(round(w / base) * base), int(round(h / base) * base)


def to_tensor_normalized(images_u8, mean=0.5, std=0.5):
    """transforms.ToTensor() + Normalize(mean, std) (base_dataset.py:143-148): CUDA uint8 [n, h, w, c] (or [h, w, c])
    -> float32 [n, c, h, w]."""
    require_cuda(images_u8)
    x = (images_u8[None] if images_u8.dim() == 3 else images_u8).contiguous()
    if x.dtype != torch.uint8:
        raise TypeError("to_tensor_normalized: uint8 images expected, got %s" % x.dtype)
    n, h, w, c = x.shape
    out = torch.empty((n, c, h, w), device=x.device, dtype=torch.float32)
    check(_L().ppst_u8_to_nchw_norm(ptr(out), ptr(x), n, h * w, c, float(mean), float(std), stream()))
    return out


def load_images(images, load_size=512, base=16, device="cuda"):
    """The evaluation-time transform of the reference's datasets (preprocess = 'scale_shortside', no crop / flip):
    a list of decoded uint8 [h, w, 3] images (numpy or torch, any sizes) -> float32 [n, 3, H, W] on the device.
    Every image must reach the same (H, W) -- as the reference's batching requires."""
    outs = []
    for im in images:
        t = torch.as_tensor(np.ascontiguousarray(im) if isinstance(im, np.ndarray) else im)
        t = t.to(device, non_blocking=True)
        h, w = t.shape[:2]
        t = resize_bicubic_u8(t, scale_shortside_size(w, h, load_size))
        h, w = t.shape[:2]
        w2, h2 = make_power_2_size(w, h, base)
        if (w2, h2) != (w, h):
            t = resize_bicubic_u8(t, (w2, h2))
        outs.append(t)
    shapes = {tuple(t.shape) for t in outs}
    if len(shapes) != 1:
        raise ValueError("load_images: images end up with different sizes %s" % sorted(shapes))
    return to_tensor_normalized(torch.stack(outs, 0))


def tensor2im(x, normalize=True):
    """util.tensor2im (util/util.py:98-131, tile=False): float32 [n, c, h, w] (or [c, h, w]) on the device
    -> uint8 [n, h, w, c] on the device (copy to the host / encode from there)."""
    require_cuda(x)
    squeeze = x.dim() == 3
    x = (x[None] if squeeze else x).to(torch.float32).contiguous()
    n, c, h, w = x.shape
    out = torch.empty((n, h, w, c), device=x.device, dtype=torch.uint8)
    check(_L().ppst_tensor2im(ptr(out), ptr(x), n, h * w, c, 1 if normalize else 0, stream()))
    return out[0] if squeeze else out
