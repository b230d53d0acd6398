"""CPU restatement of PPSTModel.decode(target=...)'s post-process (models/ppst_model.py:288-306):
tensor2im (util/util.py:98-131) -> photo_gif.GIFSmoothing(r=30, eps=(0.02*255)**2).process (photo_gif.py:13-46) ->
ToTensor -> (x - 0.5) * 2.

TEST INFRASTRUCTURE ONLY (imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs).

**Parity unpinned.**  The arithmetic lives in a third-party dependency that is absent here:
`cv2.ximgproc.guidedFilter`, opencv-contrib-python 4.8.1.78 (requirements.txt:9; call site photo_gif.py:43).  This file
restates its published algorithm -- the colour-guide guided filter of He, Sun & Tang (ECCV 2010 / TPAMI 2013) as
implemented in opencv_contrib/modules/ximgproc/src/guided_filter.cpp: fp32 planes, normalised (2r+1)^2 box means with
BORDER_REFLECT, per-pixel 3x3 covariance inverse by cofactors, output saturate-cast back to uint8 -- and is anchored
only on (i) the reference's call site and argument values and (ii) plain OpenCV's own `cv2.boxFilter`, against which
`box_mean_*` is checked in tests/test_oracle_golden.py.  No golden vector of the contrib function exists in the
reference's tests and the function cannot be executed in this image.
"""
import numpy as np


def quantize_u8(x):
    """util.tensor2im: (x + 1) / 2 * 255 in float32, clip, astype(uint8) (truncation)."""
    x = np.asarray(x, dtype=np.float32)
    v = (x + np.float32(1.0)) / np.float32(2.0) * np.float32(255.0)
    return np.clip(v, 0, 255).astype(np.uint8)


def window_sum(a, r, axis):
    """Sum over [i-r, i+r] along `axis` with cv2.BORDER_REFLECT (fedcba|abcdefgh|hgfedcb = numpy 'symmetric')."""
    a = np.moveaxis(a, axis, -1)
    pad = [(0, 0)] * (a.ndim - 1) + [(r, r)]
    p = np.pad(a, pad, mode="symmetric")
    c = np.cumsum(p, axis=-1, dtype=p.dtype)
    c = np.concatenate([np.zeros(c.shape[:-1] + (1,), dtype=c.dtype), c], axis=-1)
    out = c[..., 2 * r + 1:] - c[..., :-(2 * r + 1)]
    return np.moveaxis(out, -1, axis)


def box_mean_int(a, r):
    """Normalised box mean of an integer plane: exact integer window sums, one rounding to float32."""
    s = window_sum(window_sum(a.astype(np.int64), r, -1), r, -2)
    return (s.astype(np.float64) * (1.0 / float((2 * r + 1) ** 2))).astype(np.float32)


def box_mean_f32(a, r):
    """Normalised box mean of a float32 plane: row sums accumulated in double and stored as float32, column sums in
    double, scaled in double, one final rounding to float32."""
    h = window_sum(a.astype(np.float64), r, -1).astype(np.float32)
    v = window_sum(h.astype(np.float64), r, -2)
    return (v * (1.0 / float((2 * r + 1) ** 2))).astype(np.float32)


def guided_filter_u8(guide, src, r, eps):
    """guide, src: uint8 [H, W, 3] -> uint8 [H, W, 3].  All per-pixel algebra in float32, in a fixed operation order
    (the CUDA kernel uses the same order with contraction disabled)."""
    f32 = np.float32
    I = [guide[..., i].astype(np.int64) for i in range(3)]
    P = [src[..., c].astype(np.int64) for c in range(3)]
    mI = [box_mean_int(I[i], r) for i in range(3)]
    mP = [box_mean_int(P[c], r) for c in range(3)]
    var = {}
    for i in range(3):
        for j in range(i, 3):
            v = box_mean_int(I[i] * I[j], r) - mI[i] * mI[j]
            if i == j:
                v = v + f32(eps)
            var[(i, j)] = v.astype(f32)
    v00, v01, v02, v11, v12, v22 = var[(0, 0)], var[(0, 1)], var[(0, 2)], var[(1, 1)], var[(1, 2)], var[(2, 2)]
    # cofactors of the symmetric 3x3 matrix
    c00 = v11 * v22 - v12 * v12
    c01 = v02 * v12 - v01 * v22
    c02 = v01 * v12 - v02 * v11
    c11 = v00 * v22 - v02 * v02
    c12 = v01 * v02 - v00 * v12
    c22 = v00 * v11 - v01 * v01
    det = (v00 * c00 + v01 * c01) + v02 * c02
    inv = [[c00 / det, c01 / det, c02 / det], [c01 / det, c11 / det, c12 / det], [c02 / det, c12 / det, c22 / det]]
    If = [guide[..., i].astype(f32) for i in range(3)]
    out = np.empty(src.shape, dtype=np.uint8)
    for c in range(3):
        cov = [box_mean_int(I[i] * P[c], r) - mI[i] * mP[c] for i in range(3)]
        a = [(inv[i][0] * cov[0] + inv[i][1] * cov[1]) + inv[i][2] * cov[2] for i in range(3)]
        b = mP[c] - ((a[0] * mI[0] + a[1] * mI[1]) + a[2] * mI[2])
        ma = [box_mean_f32(a[i].astype(f32), r) for i in range(3)]
        mb = box_mean_f32(b.astype(f32), r)
        q = ((ma[0] * If[0] + ma[1] * If[1]) + ma[2] * If[2]) + mb
        out[..., c] = np.clip(np.rint(q), 0, 255).astype(np.uint8)       # saturate_cast<uchar>(cvRound(q))
    return out


def guided_smooth(out_nchw, target_nchw, r=30, eps=(0.02 * 255) ** 2):
    """decode(sp, gl, target): out, target float [-1, 1] NCHW -> float32 NCHW of the smoothed image."""
    out_nchw = np.asarray(out_nchw, dtype=np.float32)
    target_nchw = np.asarray(target_nchw, dtype=np.float32)
    res = np.empty_like(out_nchw)
    for n in range(out_nchw.shape[0]):
        wai = quantize_u8(np.transpose(out_nchw[n], (1, 2, 0)))
        ori = quantize_u8(np.transpose(target_nchw[n], (1, 2, 0)))
        q = guided_filter_u8(ori, wai, r, np.float32(eps))
        t = q.astype(np.float32) / np.float32(255.0)                     # ToTensor
        res[n] = np.transpose((t - np.float32(0.5)) * np.float32(2.0), (2, 0, 1))
    return res
