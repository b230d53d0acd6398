// Normalisation / modulation epilogues on NHWC activations.
//
//  ppst_in_coeffs    : reduced statistics (+ style vector, + bias) -> per-(n,c) scale/shift, so that
//                      InstanceNorm -> StyleMod (stylegan2_layers.py:414-437,361-374) or InstanceNorm ->
//                      FusedLeakyReLU bias (:527-531) collapse into ONE affine per element.
//  ppst_affine_apply : y = post(pre(x*a+b)*g1 + R)*g2, optionally nearest-x2 upsampled, written into a padded
//                      (zero / reflect / replicate) destination, R optionally a bilinear-x2 upsampled skip.
//                      One read of x (+ one of R) and one write per element: the 6-7 full-tensor passes the
//                      reference spends per StyledConv become one.
//  ppst_resample     : bilinear (align_corners=False) / adaptive-average resampling into a channel slice.
#include "common.cuh"

namespace ppst {

__device__ __forceinline__ void coeffs_finish(float* a, float* b, double s, double q, int64_t t, int64_t ni, int64_t ch,
                                              int64_t c, const float* shift, const float* style, int64_t style_ld,
                                              const float* bias, double count, float eps, int use_norm) {
  double scale = 1.0, off = 0.0;
  if (use_norm) {
    const double m = s / count;
    double var = q / count - m * m;
    if (var < 0.0) var = 0.0;
    const double mean = m + (shift ? (double)shift[t] : 0.0);
    const double rstd = 1.0 / sqrt(var + (double)eps);
    scale = rstd;
    off = -mean * rstd;
  }
  if (style) {
    const double s1 = 1.0 + (double)style[ni * style_ld + ch];
    scale *= s1;
    off = off * s1 + (double)style[ni * style_ld + c + ch];
  }
  if (bias) off += (double)bias[ch];
  a[t] = (float)scale;
  b[t] = (float)off;
}

__global__ void in_coeffs_kernel(float* __restrict__ a, float* __restrict__ b, const float* __restrict__ sum,
                                 const float* __restrict__ sumsq, const float* __restrict__ shift,
                                 const float* __restrict__ style, int64_t style_ld, const float* __restrict__ bias,
                                 int64_t n, int64_t c, double count, float eps, int use_norm) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * c) return;
  coeffs_finish(a, b, use_norm ? (double)sum[t] : 0.0, use_norm ? (double)sumsq[t] : 0.0, t, t / c, t % c, c, shift,
                style, style_ld, bias, count, eps, use_norm);
}

// Tile partials [n][P][c] (from the conv epilogue) -> coefficients.  Block = 32 channels x 32 partial lanes; every
// lane accumulates a fixed arithmetic progression of partials in double and the 32 lane sums are combined in a
// fixed order: deterministic, and P = 2048 partials cost 64 dependent steps instead of 2048.
__global__ void __launch_bounds__(1024) in_coeffs_partials_kernel(
    float* __restrict__ a, float* __restrict__ b, const float* __restrict__ sum, const float* __restrict__ sumsq,
    const float* __restrict__ shift, const float* __restrict__ style, int64_t style_ld,
    const float* __restrict__ bias, int64_t n, int64_t c, int64_t n_partials, double count, float eps) {
  __shared__ double ss[32][33], sq[32][33];
  const int64_t ni = blockIdx.y;
  const int64_t ch = (int64_t)blockIdx.x * 32 + threadIdx.x;
  double s = 0.0, q = 0.0;
  if (ch < c) {
    const float* ps = sum + ni * n_partials * c + ch;
    const float* pq = sumsq + ni * n_partials * c + ch;
    // 4 independent accumulator pairs (loads of 4 partials in flight instead of a dependent chain), combined in a
    // fixed order: still deterministic, and independent of the batch
    double s1 = 0.0, s2 = 0.0, s3 = 0.0, q1 = 0.0, q2 = 0.0, q3 = 0.0;
    int64_t p = threadIdx.y;
    for (; p + 96 < n_partials; p += 128) {
      const float a0 = ps[p * c], a1 = ps[(p + 32) * c], a2 = ps[(p + 64) * c], a3 = ps[(p + 96) * c];
      const float b0 = pq[p * c], b1 = pq[(p + 32) * c], b2 = pq[(p + 64) * c], b3 = pq[(p + 96) * c];
      s += (double)a0; s1 += (double)a1; s2 += (double)a2; s3 += (double)a3;
      q += (double)b0; q1 += (double)b1; q2 += (double)b2; q3 += (double)b3;
    }
    for (; p < n_partials; p += 32) {
      s += (double)ps[p * c];
      q += (double)pq[p * c];
    }
    s = (s + s1) + (s2 + s3);
    q = (q + q1) + (q2 + q3);
  }
  ss[threadIdx.y][threadIdx.x] = s;
  sq[threadIdx.y][threadIdx.x] = q;
  __syncthreads();
  if (threadIdx.y == 0 && ch < c) {
    s = 0.0; q = 0.0;
    for (int l = 0; l < 32; ++l) { s += ss[l][threadIdx.x]; q += sq[l][threadIdx.x]; }
    coeffs_finish(a, b, s, q, ni * c + ch, ni, ch, c, shift, style, style_ld, bias, count, eps, 1);
  }
}

__device__ __forceinline__ float lrelu_slope(float v, float slope) { return v >= 0.0f ? v : v * slope; }

__device__ __forceinline__ int map_pad(int u, int size, int mode, bool& zero) {
  if (u >= 0 && u < size) return u;
  if (mode == 0) { zero = true; return 0; }
  if (mode == 1) {                       // reflect (no edge repeat): -1 -> 1, size -> size-2
    if (u < 0) u = -u;
    if (u >= size) u = 2 * (size - 1) - u;
    return min(max(u, 0), size - 1);
  }
  return min(max(u, 0), size - 1);       // replicate
}

struct Bilin2 {
  int i0, i1;
  float l0, l1;
};
// F.interpolate(scale_factor=2, mode='bilinear', align_corners=False): src = max(0, 0.5*(dst+0.5)-0.5)
__device__ __forceinline__ Bilin2 bilin_x2(int dst, int in_size) {
  float src = 0.5f * ((float)dst + 0.5f) - 0.5f;
  if (src < 0.0f) src = 0.0f;
  Bilin2 r;
  r.i0 = (int)src;
  r.i1 = r.i0 + ((r.i0 < in_size - 1) ? 1 : 0);
  r.l1 = src - (float)r.i0;
  r.l0 = 1.0f - r.l1;
  return r;
}

template <typename T, int V>
__device__ __forceinline__ void load_ch(float (&v)[V], const T* p) {
  if (V == 1) {
    v[0] = Cvt<T>::to_f(p[0]);
  } else if (sizeof(T) == 4) {
#pragma unroll
    for (int h = 0; h < V / 4; ++h) {
      const float4 f = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + 4 * h);
      v[4 * h + 0] = f.x; v[4 * h + 1] = f.y; v[4 * h + 2] = f.z; v[4 * h + 3] = f.w;
    }
  } else {
    Vec16<T> r = ld_vec16(p);
#pragma unroll
    for (int e = 0; e < V; ++e) v[e] = Cvt<T>::to_f(r.v[e]);
  }
}

template <typename T, int V>
__device__ __forceinline__ void store_ch(T* p, const float (&v)[V]) {
  if (V == 1) {
    p[0] = Cvt<T>::from_f(v[0]);
  } else if (sizeof(T) == 4) {
#pragma unroll
    for (int h = 0; h < V / 4; ++h)
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(p) + 4 * h) =
          make_float4(v[4 * h + 0], v[4 * h + 1], v[4 * h + 2], v[4 * h + 3]);
  } else {
    Vec16<T> r;
#pragma unroll
    for (int e = 0; e < V; ++e) r.v[e] = Cvt<T>::from_f(v[e]);
    st_vec16(p, r);
  }
}

// V = channels per thread (8 when c % 8 == 0, else 1)
template <typename TX, typename TO, int V>
__global__ void __launch_bounds__(256) affine_apply_kernel(TO* __restrict__ out, const TX* __restrict__ x,
                                                           const float* __restrict__ a, const float* __restrict__ b,
                                                           const TX* __restrict__ res, const float* __restrict__ ra,
                                                           const float* __restrict__ rb, ppst_apply_desc d) {
  const int cv = d.c / V;
  const int hu = d.h * d.up, wu = d.w * d.up;
  const int oh = hu + d.pad_t + d.pad_b, ow = wu + d.pad_l + d.pad_r;
  const int64_t total = (int64_t)d.n * oh * ow * cv;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int cg = (int)(t % cv);
    int64_t r = t / cv;
    const int ox = (int)(r % ow); r /= ow;
    const int oy = (int)(r % oh);
    const int ni = (int)(r / oh);
    const int ch = cg * V;
    bool zero = false;
    const int uy = map_pad(oy - d.pad_t, hu, d.pad_mode, zero);
    const int ux = map_pad(ox - d.pad_l, wu, d.pad_mode, zero);
    float y[V];
    if (zero) {
#pragma unroll
      for (int e = 0; e < V; ++e) y[e] = 0.0f;
    } else {
      const int iy = uy / d.up, ix = ux / d.up;
      float xv[V];
      load_ch<TX, V>(xv, x + (((int64_t)ni * d.h + iy) * d.w + ix) * d.c + ch);
      float av[V], bv[V];
#pragma unroll
      for (int e = 0; e < V; ++e) {
        av[e] = a ? a[(int64_t)ni * d.c + ch + e] : 1.0f;
        bv[e] = b ? b[(int64_t)ni * d.c + ch + e] : 0.0f;
      }
#pragma unroll
      for (int e = 0; e < V; ++e) y[e] = lrelu_slope(fmaf(xv[e], av[e], bv[e]), d.pre_slope) * d.pre_gain;
      if (d.res_mode != 0) {
        float rv[V];
        const int rbd = d.res_border;
        if (d.res_mode == 1) {
          load_ch<TX, V>(rv, res + (((int64_t)ni * (hu + 2 * rbd) + uy + rbd) * (wu + 2 * rbd) + ux + rbd) * d.c + ch);
        } else {
          const int lh = hu / 2, lw = wu / 2, lwp = lw + 2 * rbd;
          const Bilin2 by = bilin_x2(uy, lh), bx = bilin_x2(ux, lw);
          float v00[V], v01[V], v10[V], v11[V];
          const TX* rp = res + ((int64_t)ni * (lh + 2 * rbd) + rbd) * lwp * d.c + (int64_t)rbd * d.c + ch;
          load_ch<TX, V>(v00, rp + ((int64_t)by.i0 * lwp + bx.i0) * d.c);
          load_ch<TX, V>(v01, rp + ((int64_t)by.i0 * lwp + bx.i1) * d.c);
          load_ch<TX, V>(v10, rp + ((int64_t)by.i1 * lwp + bx.i0) * d.c);
          load_ch<TX, V>(v11, rp + ((int64_t)by.i1 * lwp + bx.i1) * d.c);
#pragma unroll
          for (int e = 0; e < V; ++e)
            rv[e] = by.l0 * (bx.l0 * v00[e] + bx.l1 * v01[e]) + by.l1 * (bx.l0 * v10[e] + bx.l1 * v11[e]);
        }
#pragma unroll
        for (int e = 0; e < V; ++e) {
          const float rav = ra ? ra[(int64_t)ni * d.c + ch + e] : 1.0f;
          const float rbv = rb ? rb[(int64_t)ni * d.c + ch + e] : 0.0f;
          y[e] += fmaf(rv[e], rav, rbv);
        }
      }
#pragma unroll
      for (int e = 0; e < V; ++e) y[e] = lrelu_slope(y[e], d.post_slope) * d.post_gain;
    }
    const int ob = d.out_border;
    store_ch<TO, V>(out + (((int64_t)ni * (oh + 2 * ob) + oy + ob) * (ow + 2 * ob) + ox + ob) * d.out_c_stride + d.out_c_off + ch, y);
  }
}

// 8 channels of T kept PACKED in registers (4 regs for bf16, 8 for f32) until they are consumed
template <typename T> struct Raw8 { uint4 q[sizeof(T) / 2]; };
template <typename T>
__device__ __forceinline__ Raw8<T> ld_raw8(const T* p) {
  Raw8<T> r;
#pragma unroll
  for (int i = 0; i < (int)(sizeof(T) / 2); ++i) {
    const uint4* src = reinterpret_cast<const uint4*>(p) + i;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.q[i].x), "=r"(r.q[i].y), "=r"(r.q[i].z), "=r"(r.q[i].w)
                 : "l"(src));
  }
  return r;
}
// same, but allocating in L1: the bilinear-skip neighbours are re-read by the next pixel and the next output row
template <typename T>
__device__ __forceinline__ Raw8<T> ld_raw8_cached(const T* p) {
  Raw8<T> r;
#pragma unroll
  for (int i = 0; i < (int)(sizeof(T) / 2); ++i) {
    const uint4* src = reinterpret_cast<const uint4*>(p) + i;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.q[i].x), "=r"(r.q[i].y), "=r"(r.q[i].z), "=r"(r.q[i].w)
                 : "l"(src));
  }
  return r;
}
__device__ __forceinline__ void cvt8(const Raw8<__nv_bfloat16>& r, float (&f)[8]) {
  const uint32_t w[4] = {r.q[0].x, r.q[0].y, r.q[0].z, r.q[0].w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ void cvt8(const Raw8<float>& r, float (&f)[8]) {
  f[0] = __uint_as_float(r.q[0].x); f[1] = __uint_as_float(r.q[0].y); f[2] = __uint_as_float(r.q[0].z);
  f[3] = __uint_as_float(r.q[0].w); f[4] = __uint_as_float(r.q[1].x); f[5] = __uint_as_float(r.q[1].y);
  f[6] = __uint_as_float(r.q[1].z); f[7] = __uint_as_float(r.q[1].w);
}

// Row-oriented variant for c % 8 == 0 (the hot path): one block per output row of one image; a thread owns ONE group of
// 8 channels (its a/b coefficients stay in registers) and walks along the row.  Per element: one 16-byte load, 8 FMAs,
// one 16-byte store; no integer division, no coefficient reloads.  U pixels are in flight per thread (loads issued
// before any use, data kept packed) and the residual mode is a template parameter so that the common no-residual
// instance stays small enough for 3+ resident blocks per SM.
// Block-level tail of the STATS variants: every thread holds sum / sum of squares of the (rounded) values it stored for its
// 8 channels; the `lanes` threads that share a channel group are combined in a fixed order through shared memory and the
// row's partial is written to stats[(0|1)][n][rows][c] (reduced later by ppst_in_coeffs, n_partials = rows).
__device__ __forceinline__ void row_stats_store(float* __restrict__ stats, const float (&s1)[8], const float (&s2)[8], int c,
                                                int cg, int lane, int lanes, int64_t row_index, int64_t half_elems) {
  __shared__ float red[2][2048];                       // lanes * c == 256 * 8
  if (lane < lanes) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      red[0][lane * c + cg * 8 + e] = s1[e];
      red[1][lane * c + cg * 8 + e] = s2[e];
    }
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < c; ch += 256) {
    float t1 = 0.f, t2 = 0.f;
    for (int l = 0; l < lanes; ++l) { t1 += red[0][l * c + ch]; t2 += red[1][l * c + ch]; }
    stats[row_index * c + ch] = t1;
    stats[half_elems + row_index * c + ch] = t2;
  }
}

template <typename TX, typename TO, int RES, bool STATS = false>
__global__ void __launch_bounds__(256, RES == 0 ? 3 : 2) affine_apply_rows_kernel(TO* __restrict__ out, const TX* __restrict__ x,
                                                                   const float* __restrict__ a, const float* __restrict__ b,
                                                                   const TX* __restrict__ res, const float* __restrict__ ra,
                                                                   const float* __restrict__ rb, ppst_apply_desc d,
                                                                   float* __restrict__ stats = nullptr) {
  constexpr int V = 8;
  constexpr int U = RES == 2 ? 2 : 4;        // RES == 2 also keeps 4 bilinear neighbours per pixel in flight
  const int cv = d.c / V;
  const int lanes = 256 / cv;
  const int cg = threadIdx.x % cv, lane = threadIdx.x / cv;
  if (!STATS && lane >= lanes) return;
  const int oy = blockIdx.x, ni = blockIdx.y;
  const int ch = cg * V;
  const int hu = d.h * d.up, wu = d.w * d.up;
  const int oh = hu + d.pad_t + d.pad_b, ow = wu + d.pad_l + d.pad_r;
  float s1[STATS ? V : 1], s2[STATS ? V : 1];
  if (STATS) {
#pragma unroll
    for (int e = 0; e < V; ++e) { s1[STATS ? e : 0] = 0.f; s2[STATS ? e : 0] = 0.f; }
  }
  float av[V], bv[V];
#pragma unroll
  for (int e = 0; e < V; ++e) { av[e] = 1.0f; bv[e] = 0.0f; }
  if (a) load_ch<float, V>(av, a + (int64_t)ni * d.c + ch);
  if (b) load_ch<float, V>(bv, b + (int64_t)ni * d.c + ch);
  float rav[RES ? V : 1], rbv[RES ? V : 1];
  if (RES) {
#pragma unroll
    for (int e = 0; e < V; ++e) { rav[e] = 1.0f; rbv[e] = 0.0f; }
    if (ra) load_ch<float, V>(*reinterpret_cast<float(*)[V]>(rav), ra + (int64_t)ni * d.c + ch);
    if (rb) load_ch<float, V>(*reinterpret_cast<float(*)[V]>(rbv), rb + (int64_t)ni * d.c + ch);
  }
  bool zero_row = false;
  const int uy = map_pad(oy - d.pad_t, hu, d.pad_mode, zero_row);
  const TX* xrow = x + (((int64_t)ni * d.h + uy / d.up) * d.w) * d.c + ch;
  const int ob = d.out_border, rbd = d.res_border;
  TO* orow = out + (((int64_t)ni * (oh + 2 * ob) + oy + ob) * (ow + 2 * ob) + ob) * d.out_c_stride + d.out_c_off + ch;
  const TX* r0 = nullptr;
  const TX* r1 = nullptr;
  Bilin2 by{0, 0, 1.f, 0.f};
  const int lw = wu / 2;
  if (RES == 1) {
    r0 = res + (((int64_t)ni * (hu + 2 * rbd) + uy + rbd) * (wu + 2 * rbd) + rbd) * d.c + ch;
  } else if (RES == 2) {
    const int lh = hu / 2;
    by = bilin_x2(uy, lh);
    r0 = res + (((int64_t)ni * (lh + 2 * rbd) + by.i0 + rbd) * (lw + 2 * rbd) + rbd) * d.c + ch;
    r1 = res + (((int64_t)ni * (lh + 2 * rbd) + by.i1 + rbd) * (lw + 2 * rbd) + rbd) * d.c + ch;
  }
  const float pre_slope = d.pre_slope, pre_gain = d.pre_gain, post_slope = d.post_slope, post_gain = d.post_gain;
  for (int ox0 = lane; ox0 < ow && lane < lanes; ox0 += U * lanes) {
    Raw8<TX> xr[U], rr[RES == 1 ? U : 1], qq[RES == 2 ? U : 1][4];
    Bilin2 bxs[RES == 2 ? U : 1];
    int uxs[U];
    unsigned live = 0, zero = 0;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int ox = ox0 + u * lanes;
      bool z = zero_row;
      uxs[u] = map_pad(ox - d.pad_l, wu, d.pad_mode, z);
      if (ox < ow) live |= 1u << u;
      if (z) zero |= 1u << u;
      if (ox < ow && !z) {
        xr[u] = ld_raw8<TX>(xrow + (int64_t)(uxs[u] / d.up) * d.c);
        if (RES == 1) rr[u] = ld_raw8<TX>(r0 + (int64_t)uxs[u] * d.c);
        if (RES == 2) {
          const Bilin2 bx = bilin_x2(uxs[u], lw);
          bxs[u] = bx;
          qq[u][0] = ld_raw8_cached<TX>(r0 + (int64_t)bx.i0 * d.c);
          qq[u][1] = ld_raw8_cached<TX>(r0 + (int64_t)bx.i1 * d.c);
          qq[u][2] = ld_raw8_cached<TX>(r1 + (int64_t)bx.i0 * d.c);
          qq[u][3] = ld_raw8_cached<TX>(r1 + (int64_t)bx.i1 * d.c);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!((live >> u) & 1)) continue;
      float y[V];
      if ((zero >> u) & 1) {
#pragma unroll
        for (int e = 0; e < V; ++e) y[e] = 0.0f;
      } else {
        cvt8(xr[u], y);
#pragma unroll
        for (int e = 0; e < V; ++e) y[e] = lrelu_slope(fmaf(y[e], av[e], bv[e]), pre_slope) * pre_gain;
        if (RES == 1) {
          float rv[V];
          cvt8(rr[u], rv);
#pragma unroll
          for (int e = 0; e < V; ++e) y[e] += fmaf(rv[e], rav[e], rbv[e]);
        } else if (RES == 2) {
          const Bilin2 bx = bxs[RES == 2 ? u : 0];
          float t0[V], t1[V];
          cvt8(qq[RES == 2 ? u : 0][0], t0);
          cvt8(qq[RES == 2 ? u : 0][1], t1);
          float rv[V];
#pragma unroll
          for (int e = 0; e < V; ++e) rv[e] = by.l0 * (bx.l0 * t0[e] + bx.l1 * t1[e]);
          cvt8(qq[RES == 2 ? u : 0][2], t0);
          cvt8(qq[RES == 2 ? u : 0][3], t1);
#pragma unroll
          for (int e = 0; e < V; ++e) {
            const float r = rv[e] + by.l1 * (bx.l0 * t0[e] + bx.l1 * t1[e]);
            y[e] += fmaf(r, rav[e], rbv[e]);
          }
        }
#pragma unroll
        for (int e = 0; e < V; ++e) y[e] = lrelu_slope(y[e], post_slope) * post_gain;
      }
      store_ch<TO, V>(orow + (int64_t)(ox0 + u * lanes) * d.out_c_stride, y);
      if (STATS) {
#pragma unroll
        for (int e = 0; e < V; ++e) {
          const float yq = Cvt<TO>::to_f(Cvt<TO>::from_f(y[e]));            // the stored (rounded) value
          s1[STATS ? e : 0] += yq;
          s2[STATS ? e : 0] = fmaf(yq, yq, s2[STATS ? e : 0]);
        }
      }
    }
  }
  if (STATS)
    row_stats_store(stats, *reinterpret_cast<const float(*)[8]>(s1), *reinterpret_cast<const float(*)[8]>(s2), d.c, cg, lane,
                    lanes, (int64_t)ni * oh + oy, (int64_t)d.n * oh * d.c);
}

// Bilinear-x2 skip variant of the row kernel (generator.py:75, the end of every UpsamplingResnetBlock), no padding / no
// nearest upsampling.  A thread owns 8 channels and a run of FOUR consecutive output pixels = two low-res columns: the
// 2 x 4 low-res samples around them are loaded once (8 loads instead of 16), blended vertically in fp32, and every
// output takes two of the four blended columns.  4 independent 16-byte loads of x per thread x 512 threads per SM keep
// ~32 KB of the HBM stream in flight (the generic RES == 2 instance above holds 16 KB and reaches 0.45 of HBM).
//
// RGB != 0 (c == 128 only): the ToRGB 1x1 convolution (stylegan2_layers.py:477-495, 128 -> 3) of the tensor being written
// is computed on the fly from the ROUNDED values -- 24 FMAs per thread, then a 15-shuffle reduction of the 4 pixels x 3
// outputs over the 16 lanes that share a pixel -- so ToRGB never re-reads the 512^2 x 128 activation; RGB == 2 does not
// even write it (plain decode: nothing else consumes the last block's output).
// STATS (RGB == 0 only): d.pad_* = P in {0, 1} is a REPLICATED border written by this kernel (a block per row of the padded
// output; the first / last pixel of a row is stored twice), and every row's per-channel sum / sum of squares of the stored
// values goes to stats[2][n][h + 2P][c] -- the instance-norm statistics of the PADDED tensor that generator.py:180-231
// normalises, without the padding pass and without the reduction pass over it.
template <typename TX, typename TO, bool RAB, int RGB, bool STATS = false>
__global__ void __launch_bounds__(256, 2) affine_apply_bilin_rows_kernel(TO* __restrict__ out, const TX* __restrict__ x,
                                                                         const float* __restrict__ a, const float* __restrict__ b,
                                                                         const TX* __restrict__ res, const float* __restrict__ ra,
                                                                         const float* __restrict__ rb, ppst_apply_desc d,
                                                                         float* __restrict__ rgb_out,
                                                                         const float* __restrict__ rgb_w,
                                                                         const float* __restrict__ rgb_bias,
                                                                         float* __restrict__ stats = nullptr) {
  constexpr int V = 8;
  const int cv = d.c / V;
  const int lanes = 256 / cv;
  const int cg = threadIdx.x % cv, lane = threadIdx.x / cv;
  if (RGB == 0 && !STATS && lane >= lanes) return;    // RGB: cv == 16, every thread is live (warp shuffles below)
  const int P = STATS ? d.pad_t : 0;
  const int h = d.h, w = d.w, lh = h / 2, lw = w / 2;
  const int oyp = blockIdx.x, ni = blockIdx.y;
  const int oy = STATS ? min(max(oyp - P, 0), h - 1) : oyp;
  const int ch = cg * V;
  float s1[STATS ? V : 1], s2[STATS ? V : 1];
  if (STATS) {
#pragma unroll
    for (int e = 0; e < V; ++e) { s1[STATS ? e : 0] = 0.f; s2[STATS ? e : 0] = 0.f; }
  }
  float av[V], bv[V];
#pragma unroll
  for (int e = 0; e < V; ++e) { av[e] = 1.0f; bv[e] = 0.0f; }
  if (a) load_ch<float, V>(av, a + (int64_t)ni * d.c + ch);
  if (b) load_ch<float, V>(bv, b + (int64_t)ni * d.c + ch);
  float rav[RAB ? V : 1], rbv[RAB ? V : 1];
  if (RAB) {
#pragma unroll
    for (int e = 0; e < V; ++e) { rav[e] = 1.0f; rbv[e] = 0.0f; }
    if (ra) load_ch<float, V>(*reinterpret_cast<float(*)[V]>(rav), ra + (int64_t)ni * d.c + ch);
    if (rb) load_ch<float, V>(*reinterpret_cast<float(*)[V]>(rbv), rb + (int64_t)ni * d.c + ch);
  }
  const Bilin2 by = bilin_x2(oy, lh);
  const TX* xrow = x + (((int64_t)ni * h + oy) * w) * d.c + ch;
  const int ob = d.out_border + P, rbd = d.res_border;
  TO* orow = out + (((int64_t)ni * (h + 2 * ob) + oyp + d.out_border) * (w + 2 * ob) + ob) * d.out_c_stride + d.out_c_off + ch;
  const TX* r0 = res + (((int64_t)ni * (lh + 2 * rbd) + by.i0 + rbd) * (lw + 2 * rbd) + rbd) * d.c + ch;
  const TX* r1 = res + (((int64_t)ni * (lh + 2 * rbd) + by.i1 + rbd) * (lw + 2 * rbd) + rbd) * d.c + ch;
  const float pre_slope = d.pre_slope, pre_gain = d.pre_gain, post_slope = d.post_slope, post_gain = d.post_gain;
  float wr[RGB ? 3 : 1][RGB ? V : 1];
  float* rgb_row = nullptr;
  if (RGB) {
#pragma unroll
    for (int co = 0; co < 3; ++co) load_ch<float, V>(*reinterpret_cast<float(*)[V]>(wr[RGB ? co : 0]), rgb_w + co * d.c + ch);
    rgb_row = rgb_out + (((int64_t)ni * h + oy) * w) * 3;
  }
  const int sweeps = (lw + 2 * lanes - 1) / (2 * lanes);       // uniform trip count (the RGB reduction shuffles)
  for (int sw = 0; sw < sweeps; ++sw) {
    const int j0 = 2 * lane + sw * 2 * lanes;
    if (j0 >= lw || (STATS && lane >= lanes)) continue;        // never taken with RGB (w % 64 == 0, host-checked)
    // low-res columns j0-1 .. j0+2 (clamped) serve output pixels 2*j0 .. 2*j0+3
    int col[4];
    col[0] = max(j0 - 1, 0); col[1] = j0; col[2] = min(j0 + 1, lw - 1); col[3] = min(j0 + 2, lw - 1);
    const int ox0 = 2 * j0;
    Raw8<TX> xr[4], q0[4], q1[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (ox0 + u < w) xr[u] = ld_raw8<TX>(xrow + (int64_t)(ox0 + u) * d.c);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      q0[k] = ld_raw8_cached<TX>(r0 + (int64_t)col[k] * d.c);
      q1[k] = ld_raw8_cached<TX>(r1 + (int64_t)col[k] * d.c);
    }
    float t[4][V];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float f0[V], f1[V];
      cvt8(q0[k], f0);
      cvt8(q1[k], f1);
#pragma unroll
      for (int e = 0; e < V; ++e) t[k][e] = by.l0 * f0[e] + by.l1 * f1[e];
    }
    float part[RGB ? 12 : 1];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (RGB) { part[RGB ? 3 * u : 0] = 0.f; part[RGB ? 3 * u + 1 : 0] = 0.f; part[RGB ? 3 * u + 2 : 0] = 0.f; }
      if (ox0 + u >= w) continue;
      const Bilin2 bx = bilin_x2(ox0 + u, lw);
      const int kl = (u + 1) >> 1;                 // u = 0: columns (0,1); 1, 2: (1,2); 3: (2,3)
      float y[V];
      cvt8(xr[u], y);
#pragma unroll
      for (int e = 0; e < V; ++e) {
        y[e] = lrelu_slope(fmaf(y[e], av[e], bv[e]), pre_slope) * pre_gain;
        const float r = bx.l0 * t[kl][e] + bx.l1 * t[kl + 1][e];
        y[e] += RAB ? fmaf(r, rav[RAB ? e : 0], rbv[RAB ? e : 0]) : r;
        y[e] = lrelu_slope(y[e], post_slope) * post_gain;
      }
      if (RGB != 2) store_ch<TO, V>(orow + (int64_t)(ox0 + u) * d.out_c_stride, y);
      if (STATS) {
        float wgt = 1.0f;
        if (P && ox0 + u == 0) { store_ch<TO, V>(orow - (int64_t)d.out_c_stride, y); wgt += 1.0f; }
        if (P && ox0 + u == w - 1) { store_ch<TO, V>(orow + (int64_t)w * d.out_c_stride, y); wgt += 1.0f; }
#pragma unroll
        for (int e = 0; e < V; ++e) {
          const float yq = Cvt<TO>::to_f(Cvt<TO>::from_f(y[e]));            // the stored (rounded) value
          s1[STATS ? e : 0] = fmaf(wgt, yq, s1[STATS ? e : 0]);
          s2[STATS ? e : 0] = fmaf(wgt * yq, yq, s2[STATS ? e : 0]);
        }
      }
      if (RGB) {
#pragma unroll
        for (int e = 0; e < V; ++e) {
          const float yq = Cvt<TO>::to_f(Cvt<TO>::from_f(y[e]));            // what a reader of `out` would see
#pragma unroll
          for (int co = 0; co < 3; ++co)
            part[RGB ? 3 * u + co : 0] = fmaf(yq, wr[RGB ? co : 0][RGB ? e : 0], part[RGB ? 3 * u + co : 0]);
        }
      }
    }
    if (RGB) {
      // 12 partial sums x 16 lanes -> lane quad q = 2*b3 + b2 ends up with the 3 channel totals of pixel q
      const bool b3 = (threadIdx.x & 8) != 0, b2 = (threadIdx.x & 4) != 0;
      float q6[6], r3[3];
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const float mine = b3 ? part[RGB ? 6 + i : 0] : part[RGB ? i : 0];
        const float other = b3 ? part[RGB ? i : 0] : part[RGB ? 6 + i : 0];
        q6[i] = mine + __shfl_xor_sync(0xffffffffu, other, 8);
      }
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const float mine = b2 ? q6[3 + i] : q6[i];
        const float other = b2 ? q6[i] : q6[3 + i];
        r3[i] = mine + __shfl_xor_sync(0xffffffffu, other, 4);
      }
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        r3[i] += __shfl_xor_sync(0xffffffffu, r3[i], 2);
        r3[i] += __shfl_xor_sync(0xffffffffu, r3[i], 1);
      }
      const int u = (b3 ? 2 : 0) + (b2 ? 1 : 0);
      if ((threadIdx.x & 3) == 0 && ox0 + u < w) {
        float* o = rgb_row + (int64_t)(ox0 + u) * 3;
        o[0] = r3[0] + (rgb_bias ? rgb_bias[0] : 0.f);
        o[1] = r3[1] + (rgb_bias ? rgb_bias[1] : 0.f);
        o[2] = r3[2] + (rgb_bias ? rgb_bias[2] : 0.f);
      }
    }
  }
  if (STATS)
    row_stats_store(stats, *reinterpret_cast<const float(*)[8]>(s1), *reinterpret_cast<const float(*)[8]>(s2), d.c, cg, lane,
                    lanes, (int64_t)ni * (h + 2 * P) + oyp, (int64_t)d.n * (h + 2 * P) * d.c);
}

// mode 0 bilinear (align_corners False), mode 1 adaptive average (integer windows).
// One block per output row of one image (grid = (out_h, n)); a thread owns one group of V channels -- its affine
// coefficients stay in registers -- and walks along the row: no integer division per element, 32-bit index math.
template <typename TX, typename TO, int V>
__global__ void __launch_bounds__(256) resample_kernel(TO* __restrict__ out, const TX* __restrict__ x, int64_t n,
                                                       int in_h, int in_w, int out_h, int out_w, int c,
                                                       int out_c_stride, int out_c_off, int mode,
                                                       const float* __restrict__ a, const float* __restrict__ b,
                                                       float slope, int ib, int ob) {
  // ib / ob: border (pixels) of the input / output buffers; both are addressed at their interiors
  const int cv = c / V;
  const int oy = blockIdx.x;
  const int64_t ni = blockIdx.y;
  const int in_wp = in_w + 2 * ib;
  const float sh = (float)in_h / (float)out_h, sw = (float)in_w / (float)out_w;
  TO* orow = out + ((ni * (out_h + 2 * ob) + oy + ob) * (out_w + 2 * ob) + ob) * (int64_t)out_c_stride + out_c_off;
  const TX* img = x + ((ni * (in_h + 2 * ib) + ib) * in_wp + ib) * (int64_t)c;
  const bool tf = (a != nullptr);
  // vertical part of the bilinear stencil is the same for the whole row
  float sy = sh * ((float)oy + 0.5f) - 0.5f;
  if (sy < 0.f) sy = 0.f;
  const int y0 = min((int)sy, in_h - 1);
  const int y1 = y0 + ((y0 < in_h - 1) ? 1 : 0);
  const float ly1 = sy - (float)y0, ly0 = 1.f - ly1;
  const int fh = in_h / out_h, fw = in_w / out_w;          // adaptive-average window (mode 1)
  const int cvb = cv < 256 ? cv : 256;                      // channel groups served concurrently
  const int lanes = 256 / cvb, lane = threadIdx.x / cvb;
  if (lane >= lanes) return;
  for (int cg = threadIdx.x % cvb; cg < cv; cg += cvb) {
    const int ch = cg * V;
    float av[V], bv[V];
    if (tf) {
#pragma unroll
      for (int e = 0; e < V; ++e) { av[e] = a[ni * c + ch + e]; bv[e] = b ? b[ni * c + ch + e] : 0.0f; }
    }
    auto xf = [&](float (&v)[V]) {
      if (tf) {
#pragma unroll
        for (int e = 0; e < V; ++e) v[e] = lrelu_slope(fmaf(v[e], av[e], bv[e]), slope);
      }
    };
    const TX* base = img + ch;
    for (int ox = lane; ox < out_w; ox += lanes) {
      float y[V];
      if (mode == 0) {
        float sx = sw * ((float)ox + 0.5f) - 0.5f;
        if (sx < 0.f) sx = 0.f;
        const int x0 = min((int)sx, in_w - 1);
        const int x1 = x0 + ((x0 < in_w - 1) ? 1 : 0);
        const float lx1 = sx - (float)x0, lx0 = 1.f - lx1;
        float v00[V], v01[V], v10[V], v11[V];
        load_ch<TX, V>(v00, base + ((int64_t)y0 * in_wp + x0) * c);
        load_ch<TX, V>(v01, base + ((int64_t)y0 * in_wp + x1) * c);
        load_ch<TX, V>(v10, base + ((int64_t)y1 * in_wp + x0) * c);
        load_ch<TX, V>(v11, base + ((int64_t)y1 * in_wp + x1) * c);
        xf(v00); xf(v01); xf(v10); xf(v11);
#pragma unroll
        for (int e = 0; e < V; ++e) y[e] = ly0 * (lx0 * v00[e] + lx1 * v01[e]) + ly1 * (lx0 * v10[e] + lx1 * v11[e]);
      } else {
#pragma unroll
        for (int e = 0; e < V; ++e) y[e] = 0.f;
        for (int dy = 0; dy < fh; ++dy)
          for (int dx = 0; dx < fw; ++dx) {
            float v[V];
            load_ch<TX, V>(v, base + ((int64_t)(oy * fh + dy) * in_wp + (ox * fw + dx)) * c);
            xf(v);
#pragma unroll
            for (int e = 0; e < V; ++e) y[e] += v[e];
          }
        const float inv = 1.0f / (float)(fh * fw);
#pragma unroll
        for (int e = 0; e < V; ++e) y[e] *= inv;
      }
      store_ch<TO, V>(orow + (int64_t)ox * out_c_stride + ch, y);
    }
  }
}

static int grid_for(int64_t total) {
  const int64_t b = (total + 255) / 256;
  const int64_t cap = 148LL * 16;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace ppst

extern "C" int ppst_in_coeffs(float* a, float* b, const float* sum, const float* sumsq, const float* shift,
                              const float* style, int64_t style_ld, const float* bias, int64_t n, int64_t c,
                              int64_t n_partials, double count, float eps, int use_norm, void* stream) {
  using namespace ppst;
  PPST_REQUIRE(a && b && n > 0 && c > 0, PPST_ERR_ARG, "in_coeffs: bad args");
  PPST_REQUIRE(!use_norm || (sum && sumsq && count > 0), PPST_ERR_ARG, "in_coeffs: statistics missing");
  if (style_ld <= 0) style_ld = 2 * c;
  if (use_norm && n_partials > 0) {
    PPST_REQUIRE(n <= 65535, PPST_ERR_ARG, "in_coeffs: batch too large");
    dim3 grid(ceil_div(c, 32), (unsigned)n), block(32, 32);
    in_coeffs_partials_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(a, b, sum, sumsq, shift, style, style_ld, bias, n,
                                                                        c, n_partials, count, eps);
  } else {
    in_coeffs_kernel<<<ceil_div(n * c, 256), 256, 0, (cudaStream_t)stream>>>(a, b, sum, sumsq, shift, style, style_ld, bias,
                                                                            n, c, count, eps, use_norm);
  }
  PPST_LAUNCH_CHECK("in_coeffs_kernel");
  return 0;
}

extern "C" int ppst_affine_apply(void* out, const void* x, const float* a, const float* b, const void* res,
                                 const float* ra, const float* rb, const ppst_apply_desc* dp, void* stream) {
  using namespace ppst;
  PPST_REQUIRE(out && x && dp, PPST_ERR_ARG, "affine_apply: null argument");
  const ppst_apply_desc d = *dp;
  PPST_REQUIRE(d.n > 0 && d.h > 0 && d.w > 0 && d.c > 0 && (d.up == 1 || d.up == 2), PPST_ERR_ARG,
               "affine_apply: bad geometry");
  PPST_REQUIRE(d.res_mode == 0 || res != nullptr, PPST_ERR_ARG, "affine_apply: residual missing");
  PPST_REQUIRE(d.res_mode != 2 || ((d.h * d.up) % 2 == 0 && (d.w * d.up) % 2 == 0), PPST_ERR_ARG,
               "affine_apply: bilinear-x2 residual needs even output size");
  PPST_REQUIRE(d.pad_mode >= 0 && d.pad_mode <= 2, PPST_ERR_ARG, "affine_apply: pad_mode");
  PPST_REQUIRE(d.pad_mode != 1 || (d.pad_t < d.h * d.up && d.pad_b < d.h * d.up && d.pad_l < d.w * d.up &&
                                   d.pad_r < d.w * d.up), PPST_ERR_ARG, "affine_apply: reflect pad >= size");
  PPST_REQUIRE(d.res_mode == 0 || d.res_dtype == d.x_dtype, PPST_ERR_UNSUPPORTED,
               "affine_apply: residual dtype must equal x dtype");
  PPST_REQUIRE(d.out_c_off >= 0 && d.out_c_off + d.c <= d.out_c_stride, PPST_ERR_ARG, "affine_apply: channel slice");
  PPST_REQUIRE(d.out_border >= 0 && d.out_border <= 8 && d.res_border >= 0 && d.res_border <= 8, PPST_ERR_ARG,
               "affine_apply: out_border / res_border");
  const bool vec = (d.c % 8 == 0) && (d.out_c_stride % 8 == 0) && (d.out_c_off % 8 == 0) &&
                   ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) |
                     reinterpret_cast<uintptr_t>(res)) & 15) == 0;
  const int64_t total = (int64_t)d.n * (d.h * d.up + d.pad_t + d.pad_b) * (d.w * d.up + d.pad_l + d.pad_r) *
                        (vec ? d.c / 8 : d.c);
  const int grid = grid_for(total);
  cudaStream_t st = (cudaStream_t)stream;
  const int oh_ = d.h * d.up + d.pad_t + d.pad_b;
  const bool rows = vec && (d.c / 8) <= 256 && d.n <= 65535 &&
                    ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(ra) |
                      reinterpret_cast<uintptr_t>(rb)) & 15) == 0;
  const bool bilin_fast = d.res_mode == 2 && d.up == 1 && d.pad_t == 0 && d.pad_b == 0 && d.pad_l == 0 && d.pad_r == 0;
#define PPST_APPLY(TX, TO)                                                                                           \
  do {                                                                                                               \
    if (rows && d.res_mode == 0)                                                                                     \
      affine_apply_rows_kernel<TX, TO, 0><<<dim3(oh_, d.n), 256, 0, st>>>((TO*)out, (const TX*)x, a, b,              \
                                                                          (const TX*)res, ra, rb, d);                \
    else if (rows && d.res_mode == 1)                                                                                \
      affine_apply_rows_kernel<TX, TO, 1><<<dim3(oh_, d.n), 256, 0, st>>>((TO*)out, (const TX*)x, a, b,              \
                                                                          (const TX*)res, ra, rb, d);                \
    else if (rows && bilin_fast && !(ra || rb))                                                                      \
      affine_apply_bilin_rows_kernel<TX, TO, false, 0><<<dim3(oh_, d.n), 256, 0, st>>>(                              \
          (TO*)out, (const TX*)x, a, b, (const TX*)res, ra, rb, d, nullptr, nullptr, nullptr);                       \
    else if (rows && bilin_fast)                                                                                     \
      affine_apply_bilin_rows_kernel<TX, TO, true, 0><<<dim3(oh_, d.n), 256, 0, st>>>(                               \
          (TO*)out, (const TX*)x, a, b, (const TX*)res, ra, rb, d, nullptr, nullptr, nullptr);                       \
    else if (rows)                                                                                                   \
      affine_apply_rows_kernel<TX, TO, 2><<<dim3(oh_, d.n), 256, 0, st>>>((TO*)out, (const TX*)x, a, b,              \
                                                                          (const TX*)res, ra, rb, d);                \
    else if (vec)                                                                                                    \
      affine_apply_kernel<TX, TO, 8><<<grid, 256, 0, st>>>((TO*)out, (const TX*)x, a, b, (const TX*)res, ra, rb, d); \
    else                                                                                                             \
      affine_apply_kernel<TX, TO, 1><<<grid, 256, 0, st>>>((TO*)out, (const TX*)x, a, b, (const TX*)res, ra, rb, d); \
  } while (0)
  if (d.x_dtype == PPST_F32 && d.out_dtype == PPST_F32) PPST_APPLY(float, float);
  else if (d.x_dtype == PPST_BF16 && d.out_dtype == PPST_BF16) PPST_APPLY(__nv_bfloat16, __nv_bfloat16);
  else if (d.x_dtype == PPST_BF16 && d.out_dtype == PPST_F32) PPST_APPLY(__nv_bfloat16, float);
  else if (d.x_dtype == PPST_F32 && d.out_dtype == PPST_BF16) PPST_APPLY(float, __nv_bfloat16);
  else PPST_REQUIRE(false, PPST_ERR_DTYPE, "affine_apply: dtype combination %d -> %d", d.x_dtype, d.out_dtype);
#undef PPST_APPLY
  PPST_LAUNCH_CHECK("affine_apply_kernel");
  return 0;
}

extern "C" int ppst_affine_apply_stats(void* out, const void* x, const float* a, const float* b, const void* res,
                                       const float* ra, const float* rb, const ppst_apply_desc* dp, float* stats,
                                       void* stream) {
  using namespace ppst;
  PPST_REQUIRE(out && x && dp && stats, PPST_ERR_ARG, "affine_apply_stats: null argument");
  const ppst_apply_desc d = *dp;
  PPST_REQUIRE(d.n > 0 && d.n <= 65535 && d.h > 0 && d.w > 0 && d.c > 0 && d.c % 8 == 0 && d.c <= 2048 && d.up == 1,
               PPST_ERR_ARG, "affine_apply_stats: bad geometry (c a multiple of 8, at most 2048; no upsampling)");
  PPST_REQUIRE(d.x_dtype == PPST_BF16 && d.out_dtype == PPST_BF16 && (d.res_mode == 0 || d.res_dtype == PPST_BF16),
               PPST_ERR_DTYPE, "affine_apply_stats: bf16 activations");
  PPST_REQUIRE(d.res_mode == 0 || res != nullptr, PPST_ERR_ARG, "affine_apply_stats: residual missing");
  PPST_REQUIRE(d.out_c_stride % 8 == 0 && d.out_c_off % 8 == 0 && d.out_c_off + d.c <= d.out_c_stride, PPST_ERR_ARG,
               "affine_apply_stats: channel slice");
  PPST_REQUIRE(d.out_border >= 0 && d.out_border <= 8 && d.res_border >= 0 && d.res_border <= 8, PPST_ERR_ARG,
               "affine_apply_stats: out_border / res_border");
  PPST_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(res) |
                 reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(ra) |
                 reinterpret_cast<uintptr_t>(rb)) & 15) == 0, PPST_ERR_ARG, "affine_apply_stats: buffers must be 16-byte aligned");
  PPST_REQUIRE(d.pad_mode >= 0 && d.pad_mode <= 2, PPST_ERR_ARG, "affine_apply_stats: pad_mode");
  cudaStream_t st = (cudaStream_t)stream;
  typedef __nv_bfloat16 T;
  const int oh = d.h + d.pad_t + d.pad_b;
  if (d.res_mode == 2) {
    const int P = d.pad_t;
    PPST_REQUIRE(d.h % 2 == 0 && d.w % 2 == 0 && (P == 0 || P == 1) && d.pad_b == P && d.pad_l == P && d.pad_r == P &&
                 (P == 0 || d.pad_mode == 2), PPST_ERR_UNSUPPORTED,
                 "affine_apply_stats: the bilinear-x2 skip takes no padding or a replicated 1-pixel border");
    if (ra || rb)
      affine_apply_bilin_rows_kernel<T, T, true, 0, true><<<dim3(oh, d.n), 256, 0, st>>>(
          (T*)out, (const T*)x, a, b, (const T*)res, ra, rb, d, nullptr, nullptr, nullptr, stats);
    else
      affine_apply_bilin_rows_kernel<T, T, false, 0, true><<<dim3(oh, d.n), 256, 0, st>>>(
          (T*)out, (const T*)x, a, b, (const T*)res, ra, rb, d, nullptr, nullptr, nullptr, stats);
  } else {
    PPST_REQUIRE(d.pad_mode != 1 || (d.pad_t < d.h && d.pad_b < d.h && d.pad_l < d.w && d.pad_r < d.w), PPST_ERR_ARG,
                 "affine_apply_stats: reflect pad >= size");
    if (d.res_mode == 0)
      affine_apply_rows_kernel<T, T, 0, true><<<dim3(oh, d.n), 256, 0, st>>>((T*)out, (const T*)x, a, b, (const T*)res, ra,
                                                                              rb, d, stats);
    else
      affine_apply_rows_kernel<T, T, 1, true><<<dim3(oh, d.n), 256, 0, st>>>((T*)out, (const T*)x, a, b, (const T*)res, ra,
                                                                              rb, d, stats);
  }
  PPST_LAUNCH_CHECK("affine_apply_stats");
  return 0;
}

extern "C" int ppst_affine_apply_torgb(void* out, float* rgb_out, const void* x, const float* a, const float* b,
                                       const void* res, const float* rgb_w, const float* rgb_bias,
                                       const ppst_apply_desc* dp, void* stream) {
  using namespace ppst;
  PPST_REQUIRE(rgb_out && x && res && rgb_w && dp, PPST_ERR_ARG, "affine_apply_torgb: null argument");
  const ppst_apply_desc d = *dp;
  PPST_REQUIRE(d.n > 0 && d.n <= 65535 && d.h > 0 && d.w > 0 && d.h % 2 == 0 && d.w % 64 == 0, PPST_ERR_ARG,
               "affine_apply_torgb: bad geometry (h even, w a multiple of 64: whole warps stay converged)");
  PPST_REQUIRE(d.c == 128 && d.res_mode == 2 && d.up == 1 && d.pad_t == 0 && d.pad_b == 0 && d.pad_l == 0 && d.pad_r == 0,
               PPST_ERR_UNSUPPORTED, "affine_apply_torgb: c must be 128 with a bilinear-x2 skip and no padding / upsampling");
  PPST_REQUIRE(d.x_dtype == PPST_BF16 && d.res_dtype == PPST_BF16 && d.out_dtype == PPST_BF16, PPST_ERR_DTYPE,
               "affine_apply_torgb: bf16 activations");
  PPST_REQUIRE(out == nullptr || (d.out_c_stride % 8 == 0 && d.out_c_off % 8 == 0 && d.out_c_off + d.c <= d.out_c_stride),
               PPST_ERR_ARG, "affine_apply_torgb: channel slice");
  PPST_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(res) |
                 reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(rgb_w)) & 15) == 0,
               PPST_ERR_ARG, "affine_apply_torgb: buffers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  typedef __nv_bfloat16 T;
  if (out)
    affine_apply_bilin_rows_kernel<T, T, false, 1><<<dim3(d.h, d.n), 256, 0, st>>>(
        (T*)out, (const T*)x, a, b, (const T*)res, nullptr, nullptr, d, rgb_out, rgb_w, rgb_bias);
  else
    affine_apply_bilin_rows_kernel<T, T, false, 2><<<dim3(d.h, d.n), 256, 0, st>>>(
        nullptr, (const T*)x, a, b, (const T*)res, nullptr, nullptr, d, rgb_out, rgb_w, rgb_bias);
  PPST_LAUNCH_CHECK("affine_apply_bilin_rows_kernel(torgb)");
  return 0;
}

// Bilinear upsampling by an integer factor F in {1, 2, 4, 8} (mode 0, out = F x in: the feature heads' x4 / x2 / x1 maps,
// generator.py:252-280, and the x8 / x4 / x2 maps of E2.warp).  The generic kernel above loads and transforms (affine +
// PReLU) four neighbours for EVERY output element -- 31 instructions per element, issue-bound at 1.5-2.8 TB/s -- although
// the F outputs between two input columns share them.  Here a work item is one input column PAIR (j-1, j) x one 16-byte
// channel vector: the four neighbours are loaded and transformed once and the F outputs that lie between the two columns
// (ox = F(j-1) + F/2 .. F(j-1) + F/2 + F - 1) are blended from registers; item j = 0 does the two clamped half segments at
// the row ends.  Same expressions per element as the generic kernel (horizontal blend, then vertical).
namespace ppst {
template <typename TX, typename TO, int F>
__global__ void __launch_bounds__(256) resample_up_kernel(TO* __restrict__ out, const TX* __restrict__ x, int in_h, int in_w,
                                                          int c, int out_c_stride, int out_c_off,
                                                          const float* __restrict__ a, const float* __restrict__ b,
                                                          float slope, int ib, int ob) {
  constexpr int V = 8;
  const int cv = c / V;
  const int out_h = in_h * F, out_w = in_w * F;
  const int oy = blockIdx.x;
  const int64_t ni = blockIdx.y;
  const int in_wp = in_w + 2 * ib;
  TO* orow = out + ((ni * (out_h + 2 * ob) + oy + ob) * (out_w + 2 * ob) + ob) * (int64_t)out_c_stride + out_c_off;
  const TX* img = x + ((ni * (in_h + 2 * ib) + ib) * in_wp + ib) * (int64_t)c;
  const bool tf = (a != nullptr);
  const float sh = 1.0f / (float)F;                          // exact: F is a power of two
  float sy = sh * ((float)oy + 0.5f) - 0.5f;
  if (sy < 0.f) sy = 0.f;
  const int y0 = min((int)sy, in_h - 1);
  const int y1 = y0 + ((y0 < in_h - 1) ? 1 : 0);
  const float ly1 = sy - (float)y0, ly0 = 1.f - ly1;
  const TX* row0 = img + (int64_t)y0 * in_wp * c;
  const TX* row1 = img + (int64_t)y1 * in_wp * c;
  const int cvb = cv < 256 ? cv : 256;                      // channel groups served concurrently
  const int lanes = 256 / cvb, lane = threadIdx.x / cvb;
  if (lane >= lanes) return;
  for (int cg = threadIdx.x % cvb; cg < cv; cg += cvb) {
    const int ch = cg * V;
    float av[V], bv[V];
    if (tf) {
#pragma unroll
      for (int e = 0; e < V; ++e) { av[e] = a[ni * c + ch + e]; bv[e] = b ? b[ni * c + ch + e] : 0.0f; }
    }
    auto ld = [&](const TX* rowp, int xx, float (&v)[V]) {
      load_ch<TX, V>(v, rowp + (int64_t)xx * c + ch);
      if (tf) {
#pragma unroll
        for (int e = 0; e < V; ++e) v[e] = lrelu_slope(fmaf(v[e], av[e], bv[e]), slope);
      }
    };
    auto emit = [&](int ox, const float (&v00)[V], const float (&v01)[V], const float (&v10)[V], const float (&v11)[V], int x0) {
      float sx = sh * ((float)ox + 0.5f) - 0.5f;
      if (sx < 0.f) sx = 0.f;
      const float lx1 = sx - (float)x0, lx0 = 1.f - lx1;
      float y[V];
#pragma unroll
      for (int e = 0; e < V; ++e) y[e] = ly0 * (lx0 * v00[e] + lx1 * v01[e]) + ly1 * (lx0 * v10[e] + lx1 * v11[e]);
      store_ch<TO, V>(orow + (int64_t)ox * out_c_stride + ch, y);
    };
    if (F == 1) {
      // the source sample itself (both blend weights of the other neighbours are exactly 0)
      for (int j = lane; j < in_w; j += lanes) {
        float v[V];
        ld(row0, j, v);
        store_ch<TO, V>(orow + (int64_t)j * out_c_stride + ch, v);
      }
      continue;
    }
    for (int j = lane; j < in_w; j += lanes) {
      if (j > 0) {
        float v00[V], v01[V], v10[V], v11[V];
        ld(row0, j - 1, v00); ld(row0, j, v01); ld(row1, j - 1, v10); ld(row1, j, v11);
        const int ox0 = F * (j - 1) + F / 2;
#pragma unroll
        for (int t = 0; t < F; ++t) emit(ox0 + t, v00, v01, v10, v11, j - 1);
      } else {
        // left end: source coordinate clamped to 0 (lx1 = 0); right end: x0 = x1 = in_w - 1
        float l0[V], l1[V], r0[V], r1[V];
        ld(row0, 0, l0); ld(row1, 0, l1); ld(row0, in_w - 1, r0); ld(row1, in_w - 1, r1);
#pragma unroll
        for (int t = 0; t < F / 2; ++t) {
          emit(t, l0, l0, l1, l1, 0);
          emit(F * (in_w - 1) + F / 2 + t, r0, r0, r1, r1, in_w - 1);
        }
      }
    }
  }
}
}  // namespace ppst

extern "C" int ppst_resample(void* out, const void* x, int64_t n, int in_h, int in_w, int out_h, int out_w, int c,
                             int out_c_stride, int out_c_off, int mode, int in_dtype, int out_dtype, const float* a,
                             const float* b, float slope, int in_border, int out_border, void* stream) {
  using namespace ppst;
  PPST_REQUIRE(out && x && n > 0 && in_h > 0 && in_w > 0 && out_h > 0 && out_w > 0 && c > 0 && in_border >= 0 &&
               out_border >= 0, PPST_ERR_ARG, "resample: bad args");
  PPST_REQUIRE(mode == 0 || (mode == 1 && in_h % out_h == 0 && in_w % out_w == 0), PPST_ERR_UNSUPPORTED,
               "resample: adaptive average needs integer windows (%dx%d -> %dx%d)", in_h, in_w, out_h, out_w);
  PPST_REQUIRE(out_c_off >= 0 && out_c_off + c <= out_c_stride, PPST_ERR_ARG, "resample: channel slice");
  const bool vec = (c % 8 == 0) && (out_c_stride % 8 == 0) && (out_c_off % 8 == 0) &&
                   ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  PPST_REQUIRE(n <= 65535, PPST_ERR_ARG, "resample: batch too large");
  const dim3 grid((unsigned)out_h, (unsigned)n);
  cudaStream_t st = (cudaStream_t)stream;
  {
    // integer-factor bilinear upsampling (x1 = the transform alone): shared neighbours, see resample_up_kernel
    const int f = out_h / in_h;
    const char* e_up = getenv("PPST_RESAMPLE_UP");           // read per call: tests toggle it (0 = generic kernel)
    const int env_up = e_up ? atoi(e_up) : 1;
    if (env_up && mode == 0 && vec && out_h == f * in_h && out_w == f * in_w && (f == 1 || f == 2 || f == 4 || f == 8) &&
        (int64_t)in_w * (c / 8) < (1 << 30)) {
#define PPST_RU(TX, TO, F_)                                                                                          \
  resample_up_kernel<TX, TO, F_><<<grid, 256, 0, st>>>((TO*)out, (const TX*)x, in_h, in_w, c, out_c_stride, out_c_off, \
                                                       a, b, slope, in_border, out_border)
#define PPST_RUF(TX, TO)                                                                       \
  do {                                                                                         \
    if (f == 1) PPST_RU(TX, TO, 1); else if (f == 2) PPST_RU(TX, TO, 2);                       \
    else if (f == 4) PPST_RU(TX, TO, 4); else PPST_RU(TX, TO, 8);                              \
  } while (0)
      if (in_dtype == PPST_F32 && out_dtype == PPST_F32) PPST_RUF(float, float);
      else if (in_dtype == PPST_BF16 && out_dtype == PPST_BF16) PPST_RUF(__nv_bfloat16, __nv_bfloat16);
      else if (in_dtype == PPST_BF16 && out_dtype == PPST_F32) PPST_RUF(__nv_bfloat16, float);
      else if (in_dtype == PPST_F32 && out_dtype == PPST_BF16) PPST_RUF(float, __nv_bfloat16);
      else PPST_REQUIRE(false, PPST_ERR_DTYPE, "resample: dtype combination %d -> %d", in_dtype, out_dtype);
#undef PPST_RUF
#undef PPST_RU
      PPST_LAUNCH_CHECK("resample_up_kernel");
      return 0;
    }
  }
#define PPST_RS(TX, TO)                                                                                        \
  do {                                                                                                         \
    if (vec)                                                                                                   \
      resample_kernel<TX, TO, 8><<<grid, 256, 0, st>>>((TO*)out, (const TX*)x, n, in_h, in_w, out_h, out_w, c, \
                                                       out_c_stride, out_c_off, mode, a, b, slope, in_border,  \
                                                       out_border);                                            \
    else                                                                                                       \
      resample_kernel<TX, TO, 1><<<grid, 256, 0, st>>>((TO*)out, (const TX*)x, n, in_h, in_w, out_h, out_w, c, \
                                                       out_c_stride, out_c_off, mode, a, b, slope, in_border,  \
                                                       out_border);                                            \
  } while (0)
  if (in_dtype == PPST_F32 && out_dtype == PPST_F32) PPST_RS(float, float);
  else if (in_dtype == PPST_BF16 && out_dtype == PPST_BF16) PPST_RS(__nv_bfloat16, __nv_bfloat16);
  else if (in_dtype == PPST_BF16 && out_dtype == PPST_F32) PPST_RS(__nv_bfloat16, float);
  else if (in_dtype == PPST_F32 && out_dtype == PPST_BF16) PPST_RS(float, __nv_bfloat16);
  else PPST_REQUIRE(false, PPST_ERR_DTYPE, "resample: dtype combination %d -> %d", in_dtype, out_dtype);
#undef PPST_RS
  PPST_LAUNCH_CHECK("resample_kernel");
  return 0;
}


// ---------------------------------------------------------------------------------------------------------------------
// Adjoint of the bilinear x2 upsampling (F.interpolate(scale_factor=2, mode="bilinear", align_corners=False),
// generator.py:75) for the training path: gin[n, i, j, :] = sum over the <= 4 x 4 output pixels that read input (i, j),
// as a GATHER (one thread per input pixel and 16-byte channel vector, fixed summation order) -- deterministic, unlike the
// scatter-with-atomics form.  Per axis, output 2i reads inputs (i-1: 0.25, i: 0.75) and output 2i+1 reads (i: 0.75,
// i+1: 0.25), with the source coordinate clamped at both ends, so input i collects
//   0.25 * g[2i-1] + 0.75 * g[2i] + 0.75 * g[2i+1] + 0.25 * g[2i+2]      (+0.25 * g[0] at i = 0, +0.25 * g[2L-1] at i = L-1).
namespace ppst {
template <typename T>
__global__ void __launch_bounds__(256) bilinear2x_backward_kernel(T* __restrict__ gin, const T* __restrict__ g, int64_t n,
                                                                  int h, int w, int c, int nearest) {
  constexpr int VEC = Vec16<T>::N;
  const int cv = c / VEC;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = n * h * w * cv;
  if (idx >= total) return;
  const int v = (int)(idx % cv);
  int64_t r = idx / cv;
  const int j = (int)(r % w); r /= w;
  const int i = (int)(r % h);
  const int64_t ni = r / h;
  const int oh = 2 * h, ow = 2 * w;
  // the (up to) four output rows / columns touching input i / j and their weights
  int ys[4], xs[4];
  float wy[4], wx[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int oy = 2 * i - 1 + t, ox = 2 * j - 1 + t;
    const float base = nearest ? ((t == 1 || t == 2) ? 1.0f : 0.0f) : ((t == 0 || t == 3) ? 0.25f : 0.75f);
    ys[t] = oy; xs[t] = ox;
    wy[t] = (oy >= 0 && oy < oh) ? base : 0.0f;
    wx[t] = (ox >= 0 && ox < ow) ? base : 0.0f;
  }
  // clamping: output 0 puts its whole weight on input 0, output 2L-1 on input L-1
  if (!nearest) {
    if (i == 0) wy[1] = 1.0f;
    if (i == h - 1) wy[2] = 1.0f;
    if (j == 0) wx[1] = 1.0f;
    if (j == w - 1) wx[2] = 1.0f;
  }
  float acc[VEC];
#pragma unroll
  for (int e = 0; e < VEC; ++e) acc[e] = 0.0f;
  const T* gb = g + (ni * oh * ow) * (int64_t)c + (int64_t)v * VEC;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    if (wy[a] == 0.0f) continue;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      if (wx[b] == 0.0f) continue;
      const Vec16<T> gv = ld_vec16(gb + ((int64_t)ys[a] * ow + xs[b]) * c);
      const float wgt = wy[a] * wx[b];
#pragma unroll
      for (int e = 0; e < VEC; ++e) acc[e] = fmaf(wgt, Cvt<T>::to_f(gv.v[e]), acc[e]);
    }
  }
  Vec16<T> o;
#pragma unroll
  for (int e = 0; e < VEC; ++e) o.v[e] = Cvt<T>::from_f(acc[e]);
  st_vec16(gin + ((ni * h + i) * w + j) * (int64_t)c + (int64_t)v * VEC, o);
}
}  // namespace ppst

extern "C" int ppst_upsample2x_backward(void* gin, const void* g, int64_t n, int h, int w, int c, int nearest, int dtype,
                                        void* stream) {
  using namespace ppst;
  PPST_REQUIRE(gin && g && n > 0 && h > 0 && w > 0 && c > 0, PPST_ERR_ARG, "bilinear2x_backward: bad args");
  PPST_REQUIRE(dtype == PPST_BF16 || dtype == PPST_F32, PPST_ERR_DTYPE, "bilinear2x_backward: dtype %d", dtype);
  const int vec = dtype == PPST_BF16 ? 8 : 4;
  PPST_REQUIRE(c % vec == 0, PPST_ERR_UNSUPPORTED, "bilinear2x_backward: c=%d must be a multiple of %d", c, vec);
  const int64_t total = n * h * w * (c / vec);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == PPST_BF16)
    bilinear2x_backward_kernel<__nv_bfloat16><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
        (__nv_bfloat16*)gin, (const __nv_bfloat16*)g, n, h, w, c, nearest);
  else
    bilinear2x_backward_kernel<float><<<(unsigned)((total + 255) / 256), 256, 0, st>>>((float*)gin, (const float*)g, n, h, w, c,
                                                                                     nearest);
  PPST_LAUNCH_CHECK("bilinear2x_backward_kernel");
  return 0;
}

// Adjoint of nn.ReflectionPad2d on NHWC (the encoders' ConvLayer(reflection_pad=True)) as a gather: input pixel (i, j)
// collects its own padded position and the mirrored ones (up to 3 per axis when the image is smaller than twice the pad).
namespace ppst {
__device__ __forceinline__ int reflect_sources(int i, int len, int p0, int p1, int (&q)[3]) {
  int m = 0;
  q[m++] = i + p0;
  if (i >= 1 && i <= p0) q[m++] = p0 - i;
  if (i <= len - 2 && i >= len - 1 - p1) q[m++] = p0 + 2 * (len - 1) - i;
  return m;
}
template <typename T>
__global__ void __launch_bounds__(256) reflect_pad_backward_kernel(T* __restrict__ gin, const T* __restrict__ g, int64_t n,
                                                                   int h, int w, int c, int pt, int pb, int pl, int pr) {
  constexpr int VEC = Vec16<T>::N;
  const int cv = c / VEC;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * h * w * cv) return;
  const int v = (int)(idx % cv);
  int64_t r = idx / cv;
  const int j = (int)(r % w); r /= w;
  const int i = (int)(r % h);
  const int64_t ni = r / h;
  const int ph = h + pt + pb, pw = w + pl + pr;
  int qy[3], qx[3];
  const int my = reflect_sources(i, h, pt, pb, qy), mx = reflect_sources(j, w, pl, pr, qx);
  float acc[VEC];
#pragma unroll
  for (int e = 0; e < VEC; ++e) acc[e] = 0.0f;
  const T* gb = g + ni * ph * pw * (int64_t)c + (int64_t)v * VEC;
  for (int a = 0; a < my; ++a)
    for (int b = 0; b < mx; ++b) {
      const Vec16<T> gv = ld_vec16(gb + ((int64_t)qy[a] * pw + qx[b]) * c);
#pragma unroll
      for (int e = 0; e < VEC; ++e) acc[e] += Cvt<T>::to_f(gv.v[e]);
    }
  Vec16<T> o;
#pragma unroll
  for (int e = 0; e < VEC; ++e) o.v[e] = Cvt<T>::from_f(acc[e]);
  st_vec16(gin + ((ni * h + i) * w + j) * (int64_t)c + (int64_t)v * VEC, o);
}
}  // namespace ppst

extern "C" int ppst_reflect_pad_backward(void* gin, const void* g, int64_t n, int h, int w, int c, int pt, int pb, int pl,
                                         int pr, int dtype, void* stream) {
  using namespace ppst;
  PPST_REQUIRE(gin && g && n > 0 && h > 0 && w > 0 && c > 0, PPST_ERR_ARG, "reflect_pad_backward: bad args");
  PPST_REQUIRE(pt >= 0 && pb >= 0 && pl >= 0 && pr >= 0 && pt < h && pb < h && pl < w && pr < w, PPST_ERR_ARG,
               "reflect_pad_backward: pads must be smaller than the image");
  PPST_REQUIRE(dtype == PPST_BF16 || dtype == PPST_F32, PPST_ERR_DTYPE, "reflect_pad_backward: dtype %d", dtype);
  const int vec = dtype == PPST_BF16 ? 8 : 4;
  PPST_REQUIRE(c % vec == 0, PPST_ERR_UNSUPPORTED, "reflect_pad_backward: c=%d must be a multiple of %d", c, vec);
  const int64_t total = n * h * w * (c / vec);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == PPST_BF16)
    reflect_pad_backward_kernel<__nv_bfloat16><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
        (__nv_bfloat16*)gin, (const __nv_bfloat16*)g, n, h, w, c, pt, pb, pl, pr);
  else
    reflect_pad_backward_kernel<float><<<(unsigned)((total + 255) / 256), 256, 0, st>>>((float*)gin, (const float*)g, n, h, w,
                                                                                      c, pt, pb, pl, pr);
  PPST_LAUNCH_CHECK("reflect_pad_backward_kernel");
  return 0;
}
