// Segmented reductions: per-(sample, region, channel) sum / sum of squares / max over H*W.
// One kernel family serves InstanceNorm statistics, GAP / GMP, the region-masked GAP / GMP of the colour
// encoder keyed by the integer label map, and function.py's mean/std helpers.
//
// HBM-bound: every activation element is read exactly once with 128-bit loads; partial results go to a small
// scratch buffer and are combined by a second tiny kernel in a fixed order (deterministic; no fp atomics).
#include <float.h>

#include "common.cuh"

namespace ppst {

constexpr int kMaxRegions = 4;
constexpr int kReduceThreads = 256;

struct ReducePlan {
  int chunks;        // partial blocks per sample (NHWC) or per plane (NCHW)
  int64_t chunk_len; // pixels per chunk
};

static ReducePlan plan_reduce(int64_t n, int64_t c, int64_t hw, int channels_last) {
  // The split depends ONLY on (hw, layout), never on the batch size: a sample's statistics are then bit-identical
  // whether it is reduced alone or inside a larger batch (the pair-partitioning over GPUs relies on this).
  (void)n; (void)c;
  int64_t chunks = channels_last ? hw / 1024 : hw / 16384;
  const int64_t cap = channels_last ? 128 : 32;
  if (chunks > cap) chunks = cap;
  if (chunks < 1) chunks = 1;
  ReducePlan p;
  p.chunk_len = (hw + chunks - 1) / chunks;
  p.chunk_len = (p.chunk_len + 63) / 64 * 64;          // keep chunk starts 16-byte aligned for all dtypes
  p.chunks = (int)((hw + p.chunk_len - 1) / p.chunk_len);
  return p;
}

// scratch layout: [unit][chunk][region][c_or_1][3]   (3 = sum, sumsq, max)

// ------------------------------------------------------------------------------------------------
// NCHW: grid = (planes, chunks).  Threads stride over the pixels of one plane chunk.
template <typename T, int R>
__global__ void __launch_bounds__(kReduceThreads) plane_reduce_nchw_kernel(
    float* __restrict__ scratch, const T* __restrict__ x, const uint8_t* __restrict__ labels,
    const float* __restrict__ shift, int64_t c, int64_t hw, int64_t chunk_len, int chunks, int label_mode) {
  constexpr int VEC = Vec16<T>::N;
  const int64_t plane = blockIdx.x;
  const int chunk = blockIdx.y;
  const int64_t nidx = plane / c;
  const float sh = shift ? shift[plane] : 0.0f;
  const T* px = x + plane * hw;
  const uint8_t* pl = labels ? labels + nidx * hw : nullptr;
  const int64_t beg = (int64_t)chunk * chunk_len;
  const int64_t end = min(beg + chunk_len, hw);
  float s[R], q[R], m[R];
#pragma unroll
  for (int r = 0; r < R; ++r) { s[r] = 0.f; q[r] = 0.f; m[r] = -FLT_MAX; }
  auto accum = [&](float xv, int L) {
    if (R == 1) {
      const float d = xv - sh;
      s[0] += d; q[0] = fmaf(d, d, q[0]); m[0] = fmaxf(m[0], xv);
    } else {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const bool in = label_mode ? (((L >> r) & 1) != 0) : (L == r);
        const float d = in ? xv - sh : 0.0f;
        s[r] += d; q[r] = fmaf(d, d, q[r]); m[r] = fmaxf(m[r], in ? xv : 0.0f);
      }
    }
  };
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(px + beg) & 15) == 0);
  int64_t scalar_from = beg;                           // first element left to the scalar loop
  if (vec_ok) {
    const int64_t end_vec = beg + (end - beg) / VEC * VEC;
    for (int64_t i = beg + (int64_t)threadIdx.x * VEC; i < end_vec; i += (int64_t)kReduceThreads * VEC) {
      Vec16<T> v = ld_vec16_stream(px + i);
#pragma unroll
      for (int e = 0; e < VEC; ++e) accum(Cvt<T>::to_f(v.v[e]), (R > 1) ? pl[i + e] : 0);
    }
    scalar_from = end_vec;
  }
  for (int64_t j = scalar_from + threadIdx.x; j < end; j += kReduceThreads)
    accum(Cvt<T>::to_f(px[j]), (R > 1) ? pl[j] : 0);
  __shared__ float red[kReduceThreads / 32][R][3];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const float a = warp_sum(s[r]), b = warp_sum(q[r]), mm = warp_max(m[r]);
    if (lane == 0) { red[wid][r][0] = a; red[wid][r][1] = b; red[wid][r][2] = mm; }
  }
  __syncthreads();
  if (threadIdx.x < R) {
    const int r = threadIdx.x;
    float a = 0.f, b = 0.f, mm = -FLT_MAX;
    for (int w = 0; w < kReduceThreads / 32; ++w) { a += red[w][r][0]; b += red[w][r][1]; mm = fmaxf(mm, red[w][r][2]); }
    float* o = scratch + (((int64_t)plane * chunks + chunk) * R + r) * 3;
    o[0] = a; o[1] = b; o[2] = mm;
  }
}

template <int R>
__global__ void plane_reduce_nchw_final(float* __restrict__ sum, float* __restrict__ sumsq, float* __restrict__ vmax,
                                        const float* __restrict__ scratch, int64_t n, int64_t c, int chunks) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // over n*c*R
  if (t >= n * c * R) return;
  const int r = (int)(t % R);
  const int64_t plane = t / R;
  float a = 0.f, b = 0.f, mm = -FLT_MAX;
  for (int k = 0; k < chunks; ++k) {
    const float* o = scratch + ((plane * chunks + k) * R + r) * 3;
    a += o[0]; b += o[1]; mm = fmaxf(mm, o[2]);
  }
  const int64_t nidx = plane / c, ch = plane % c;
  const int64_t oi = (nidx * R + r) * c + ch;
  if (sum) sum[oi] = a;
  if (sumsq) sumsq[oi] = b;
  if (vmax) vmax[oi] = mm;
}

// ------------------------------------------------------------------------------------------------
// NHWC: grid = (n, chunks, channel groups).  A block is TCOLS channel-vector columns x (256/TCOLS) pixel lanes.
template <typename T, int VEC, int R, bool BORDER = false>
__global__ void __launch_bounds__(kReduceThreads) plane_reduce_nhwc_kernel(
    float* __restrict__ scratch, const T* __restrict__ x, const uint8_t* __restrict__ labels,
    const float* __restrict__ shift, int64_t c, int64_t hw, int64_t chunk_len, int chunks, int tcols,
    int label_mode, int w_row, int border) {
  extern __shared__ float red[];                       // [plane_lanes][tcols*VEC][R][3] -> reduced in place
  const int64_t nidx = blockIdx.x;
  const int chunk = blockIdx.y;
  const int col = threadIdx.x % tcols;                 // channel-vector column within the group
  const int lane_p = threadIdx.x / tcols;              // pixel lane
  const int planes = kReduceThreads / tcols;
  const int64_t ch0 = ((int64_t)blockIdx.z * tcols + col) * VEC;
  const bool active = ch0 < c;
  const int64_t beg = (int64_t)chunk * chunk_len;
  const int64_t end = min(beg + chunk_len, hw);
  float s[R][VEC], q[R][VEC], m[R][VEC], sh[VEC];
#pragma unroll
  for (int e = 0; e < VEC; ++e) {
    sh[e] = (shift && active) ? shift[nidx * c + ch0 + e] : 0.0f;
#pragma unroll
    for (int r = 0; r < R; ++r) { s[r][e] = 0.f; q[r][e] = 0.f; m[r][e] = -FLT_MAX; }
  }
  if (active) {
    // border > 0: x is [n, hw/w_row + 2*border, w_row + 2*border, c]; pixel p = (y, x) of the interior
    const int wp = w_row + 2 * border;
    const T* base = BORDER ? x + ((nidx * (hw / w_row + 2 * border) + border) * wp + border) * c + ch0
                           : x + (nidx * hw) * c + ch0;
    const uint8_t* pl = labels ? labels + nidx * hw : nullptr;
    // BORDER: (yy, xx) of pixel p tracked incrementally -- one division per thread, none per pixel (hw < 2^31, host-checked)
    int yy = BORDER ? (int)(beg + lane_p) / w_row : 0;
    int xx = BORDER ? (int)(beg + lane_p) - yy * w_row : 0;
    // UNR pixels per trip: their loads are issued together (one 16-byte load in flight per thread left the stream at
    // 2.5 TB/s), then accumulated in pixel order -- the same order, so the same sums, as one pixel per trip
    constexpr int UNR = (VEC > 1 && R == 1) ? 4 : 1;
    for (int64_t p = beg + lane_p; p < end; p += (int64_t)planes * UNR) {
      Vec16<T> v[UNR];
      T v1[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int64_t pu = p + (int64_t)u * planes;
        int64_t po = pu;
        if (BORDER) {
          po = (int64_t)yy * wp + xx;
          xx += planes;
          while (xx >= w_row) { xx -= w_row; ++yy; }
        }
        if (pu < end) {
          if (VEC > 1) v[u] = ld_vec16_stream(base + po * c);
          else v1[u] = base[po * c];
        }
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int64_t pu = p + (int64_t)u * planes;
        if (pu >= end) break;
        float xv[VEC];
        if (VEC > 1) {
#pragma unroll
          for (int e = 0; e < VEC; ++e) xv[e] = Cvt<T>::to_f(v[u].v[e]);
        } else {
          xv[0] = Cvt<T>::to_f(v1[u]);
        }
        const int L = (R > 1) ? pl[pu] : 0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const bool in = (R == 1) || (label_mode ? (((L >> r) & 1) != 0) : (L == r));
#pragma unroll
          for (int e = 0; e < VEC; ++e) {
            const float d = in ? xv[e] - sh[e] : 0.0f;
            s[r][e] += d; q[r][e] = fmaf(d, d, q[r][e]); m[r][e] = fmaxf(m[r][e], in ? xv[e] : 0.0f);
          }
        }
      }
    }
  }
  // cross-pixel-lane reduction through shared memory, fixed order
  const int width = tcols * VEC * R * 3;
  float* mine = red + (int64_t)lane_p * width + (int64_t)col * VEC * R * 3;
#pragma unroll
  for (int e = 0; e < VEC; ++e)
#pragma unroll
    for (int r = 0; r < R; ++r) {
      mine[(e * R + r) * 3 + 0] = s[r][e];
      mine[(e * R + r) * 3 + 1] = q[r][e];
      mine[(e * R + r) * 3 + 2] = m[r][e];
    }
  __syncthreads();
  for (int idx = threadIdx.x; idx < width; idx += kReduceThreads) {
    const int kind = idx % 3;
    float acc = (kind == 2) ? -FLT_MAX : 0.0f;
    for (int l = 0; l < planes; ++l) {
      const float v = red[(int64_t)l * width + idx];
      acc = (kind == 2) ? fmaxf(acc, v) : acc + v;
    }
    // idx -> (col, e, r, kind)
    const int r = (idx / 3) % R;
    const int ce = idx / (3 * R);                       // col*VEC + e
    const int64_t ch = (int64_t)blockIdx.z * tcols * VEC + ce;
    if (ch < c) scratch[((((int64_t)nidx * chunks + chunk) * R + r) * c + ch) * 3 + kind] = acc;
  }
}

// One WARP per output (n, region, channel): lane l sums chunks l, l + 32, ... left to right, then a fixed shuffle tree --
// deterministic and independent of the batch; a single thread walking 128 chunks cost 16 us of pure latency.
template <int R>
__global__ void __launch_bounds__(256) plane_reduce_nhwc_final(float* __restrict__ sum, float* __restrict__ sumsq,
                                                               float* __restrict__ vmax, const float* __restrict__ scratch,
                                                               int64_t n, int64_t c, int chunks) {
  const int lane = threadIdx.x & 31;
  const int64_t t = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);   // over n*R*c
  if (t >= n * R * c) return;
  const int64_t ch = t % c;
  const int r = (int)((t / c) % R);
  const int64_t nidx = t / (c * R);
  float a = 0.f, b = 0.f, mm = -FLT_MAX;
  const int64_t kstride = (int64_t)R * c * 3;
  const float* o = scratch + ((((int64_t)nidx * chunks) * R + r) * c + ch) * 3;
  for (int k = lane; k < chunks; k += 32) {
    a += o[k * kstride]; b += o[k * kstride + 1]; mm = fmaxf(mm, o[k * kstride + 2]);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, off);
    b += __shfl_xor_sync(0xffffffffu, b, off);
    mm = fmaxf(mm, __shfl_xor_sync(0xffffffffu, mm, off));
  }
  if (lane == 0) {
    if (sum) sum[t] = a;
    if (sumsq) sumsq[t] = b;
    if (vmax) vmax[t] = mm;
  }
}

template <typename T, int R>
static int launch_reduce(float* sum, float* sumsq, float* vmax, const T* x, const uint8_t* labels, const float* shift,
                         float* scratch, int64_t n, int64_t c, int64_t hw, int label_mode, int channels_last, int w_row,
                         int border, cudaStream_t st) {
  const ReducePlan pl = plan_reduce(n, c, hw, channels_last);
  if (!channels_last) {
    dim3 grid((unsigned)(n * c), pl.chunks);
    plane_reduce_nchw_kernel<T, R><<<grid, kReduceThreads, 0, st>>>(scratch, x, labels, shift, c, hw, pl.chunk_len,
                                                                    pl.chunks, label_mode);
    PPST_LAUNCH_CHECK("plane_reduce_nchw_kernel");
    const int64_t tot = n * c * R;
    plane_reduce_nchw_final<R><<<ceil_div(tot, 256), 256, 0, st>>>(sum, sumsq, vmax, scratch, n, c, pl.chunks);
    PPST_LAUNCH_CHECK("plane_reduce_nchw_final");
    return 0;
  }
  constexpr int VECF = Vec16<T>::N;
  const bool vec = (c % VECF == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  const int64_t cols = vec ? c / VECF : c;
  int tcols = 1;
  while (tcols < cols && tcols < kReduceThreads) tcols <<= 1;
  // shared memory: planes * tcols * VEC * R * 3 floats = 256 * VEC * R * 3 floats (<= 256*8*4*3*4 = 96 KB)
  const int vecn = vec ? VECF : 1;
  const size_t smem = sizeof(float) * (size_t)kReduceThreads * vecn * R * 3;
  dim3 grid((unsigned)n, pl.chunks, ceil_div(cols, tcols));
  if (vec && border) {
    auto kern = plane_reduce_nhwc_kernel<T, VECF, R, true>;
    if (smem > 48 * 1024) {
      int rc = cuda_check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                          "cudaFuncSetAttribute(plane_reduce)");
      if (rc) return rc;
    }
    kern<<<grid, kReduceThreads, smem, st>>>(scratch, x, labels, shift, c, hw, pl.chunk_len, pl.chunks, tcols,
                                             label_mode, w_row, border);
  } else if (vec) {
    auto kern = plane_reduce_nhwc_kernel<T, VECF, R>;
    if (smem > 48 * 1024) {
      int rc = cuda_check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                          "cudaFuncSetAttribute(plane_reduce)");
      if (rc) return rc;
    }
    kern<<<grid, kReduceThreads, smem, st>>>(scratch, x, labels, shift, c, hw, pl.chunk_len, pl.chunks, tcols,
                                             label_mode, w_row, border);
  } else if (border) {
    plane_reduce_nhwc_kernel<T, 1, R, true><<<grid, kReduceThreads, smem, st>>>(scratch, x, labels, shift, c, hw,
                                                                                pl.chunk_len, pl.chunks, tcols, label_mode,
                                                                                w_row, border);
  } else {
    plane_reduce_nhwc_kernel<T, 1, R><<<grid, kReduceThreads, smem, st>>>(scratch, x, labels, shift, c, hw,
                                                                          pl.chunk_len, pl.chunks, tcols, label_mode,
                                                                          w_row, border);
  }
  PPST_LAUNCH_CHECK("plane_reduce_nhwc_kernel");
  const int64_t tot = n * c * R;
  plane_reduce_nhwc_final<R><<<ceil_div(tot, 8), 256, 0, st>>>(sum, sumsq, vmax, scratch, n, c, pl.chunks);
  PPST_LAUNCH_CHECK("plane_reduce_nhwc_final");
  return 0;
}

// ------------------------------------------------------------------------------------------------
// GAP / GMP of a bilinear x-scale upsampled map WITHOUT materialising it (encoder_col.py:119-130: the warped 64x64
// features are F.interpolate'd by 8 / 4 / 2 and then only ever pooled).  f32 NHWC input slice [n, lh, lw, c_off:c_off+c]
// of a tensor with channel stride c_stride.  A "cell" is the square between four neighbouring source samples: the
// scale^2 outputs that fall into it share the four corner loads; the half cells along the border (clamped source
// coordinate) hold scale/2 outputs per axis.  Every output value is evaluated with the resample kernel's formula
// (src = (o + 0.5) / scale - 0.5, clamped at 0), so it equals the materialised path value for value.
// grid = (cell rows, n); block = (c/4 channel vectors) x (256 / (c/4) cell lanes); partials per cell row, then
// plane_reduce_nhwc_final<1> combines them in a fixed order.
__global__ void __launch_bounds__(256) upsample_reduce_kernel(float* __restrict__ scratch, const float* __restrict__ x,
                                                              int lh, int lw, int c_stride, int c_off, int c, int scale,
                                                              int rows) {
  extern __shared__ float red[];                       // [lanes][c][2]
  const int cv = c >> 2;
  const int lanes = 256 / cv;
  const int col = threadIdx.x % cv, lane = threadIdx.x / cv;
  const int ni = blockIdx.y;
  const int half = scale >> 1;
  const float inv = 1.0f / (float)scale;
  // cell row -> its output rows [oy_beg, oy_beg + ny)
  const int cy = (int)blockIdx.x - (scale > 1 ? 1 : 0);
  int oy_beg, ny;
  if (scale == 1) { oy_beg = cy; ny = 1; }
  else if (cy < 0) { oy_beg = 0; ny = half; }
  else if (cy == lh - 1) { oy_beg = scale * lh - half; ny = half; }
  else { oy_beg = half + cy * scale; ny = scale; }
  const int y0 = max(cy, 0), y1 = min(y0 + 1, lh - 1);
  float s[4] = {0.f, 0.f, 0.f, 0.f}, m[4] = {-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX};
  if (lane < lanes) {
    const float* base = x + ((int64_t)ni * lh * lw) * c_stride + c_off + col * 4;
    const int ncx = (scale > 1) ? lw + 1 : lw;
    for (int k = lane; k < ncx; k += lanes) {
      const int cx = k - (scale > 1 ? 1 : 0);
      int ox_beg, nx;
      if (scale == 1) { ox_beg = cx; nx = 1; }
      else if (cx < 0) { ox_beg = 0; nx = half; }
      else if (cx == lw - 1) { ox_beg = scale * lw - half; nx = half; }
      else { ox_beg = half + cx * scale; nx = scale; }
      const int x0 = max(cx, 0), x1 = min(x0 + 1, lw - 1);
      const float4 v00 = *reinterpret_cast<const float4*>(base + ((int64_t)y0 * lw + x0) * c_stride);
      const float4 v01 = *reinterpret_cast<const float4*>(base + ((int64_t)y0 * lw + x1) * c_stride);
      const float4 v10 = *reinterpret_cast<const float4*>(base + ((int64_t)y1 * lw + x0) * c_stride);
      const float4 v11 = *reinterpret_cast<const float4*>(base + ((int64_t)y1 * lw + x1) * c_stride);
      const float a00[4] = {v00.x, v00.y, v00.z, v00.w}, a01[4] = {v01.x, v01.y, v01.z, v01.w};
      const float a10[4] = {v10.x, v10.y, v10.z, v10.w}, a11[4] = {v11.x, v11.y, v11.z, v11.w};
      for (int jy = 0; jy < ny; ++jy) {
        float sy = inv * ((float)(oy_beg + jy) + 0.5f) - 0.5f;
        if (sy < 0.f) sy = 0.f;
        const float ly1 = sy - (float)y0, ly0 = 1.f - ly1;
        for (int jx = 0; jx < nx; ++jx) {
          float sx = inv * ((float)(ox_beg + jx) + 0.5f) - 0.5f;
          if (sx < 0.f) sx = 0.f;
          const float lx1 = sx - (float)x0, lx0 = 1.f - lx1;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float v = ly0 * (lx0 * a00[e] + lx1 * a01[e]) + ly1 * (lx0 * a10[e] + lx1 * a11[e]);
            s[e] += v;
            m[e] = fmaxf(m[e], v);
          }
        }
      }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      red[((int64_t)lane * c + col * 4 + e) * 2 + 0] = s[e];
      red[((int64_t)lane * c + col * 4 + e) * 2 + 1] = m[e];
    }
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < c; ch += 256) {
    float a = 0.f, mm = -FLT_MAX;
    for (int l = 0; l < lanes; ++l) {
      a += red[((int64_t)l * c + ch) * 2 + 0];
      mm = fmaxf(mm, red[((int64_t)l * c + ch) * 2 + 1]);
    }
    float* o = scratch + (((int64_t)ni * rows + blockIdx.x) * c + ch) * 3;     // plane_reduce_nhwc_final<1> layout
    o[0] = a; o[1] = 0.f; o[2] = mm;
  }
}

__global__ void labels_to_onehot_kernel(float* __restrict__ onehot, const uint8_t* __restrict__ labels, int64_t n,
                                        int64_t hw, int n_labels) {
  const int64_t total = n * n_labels * hw;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i % hw;
    const int l = (int)((i / hw) % n_labels);
    const int64_t b = i / (hw * n_labels);
    onehot[i] = (labels[b * hw + p] == l) ? 1.0f : 0.0f;
  }
}

__global__ void labels_pool_or_kernel(uint8_t* __restrict__ out, const uint8_t* __restrict__ in, int64_t n, int h, int w,
                                      int f, int in_mode) {
  const int oh = h / f, ow = w / f;
  const int64_t total = n * oh * ow;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ox = (int)(i % ow), oy = (int)((i / ow) % oh);
    const int64_t b = i / ((int64_t)ow * oh);
    unsigned m = 0;
    for (int dy = 0; dy < f; ++dy)
      for (int dx = 0; dx < f; ++dx) {
        const unsigned L = in[(b * h + oy * f + dy) * (int64_t)w + ox * f + dx];
        m |= in_mode ? L : (1u << (L & 7));
      }
    out[i] = (uint8_t)m;
  }
}

}  // namespace ppst

extern "C" int ppst_labels_pool_or(uint8_t* out, const uint8_t* in, int64_t n, int h, int w, int factor, int in_mode,
                                   void* stream) {
  using namespace ppst;
  PPST_REQUIRE(out && in && n > 0 && h > 0 && w > 0 && factor >= 1 && h % factor == 0 && w % factor == 0, PPST_ERR_ARG,
               "labels_pool_or: bad args");
  const int64_t total = n * (h / factor) * (w / factor);
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  labels_pool_or_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(out, in, n, h, w, factor, in_mode);
  PPST_LAUNCH_CHECK("labels_pool_or_kernel");
  return 0;
}

extern "C" int64_t ppst_plane_reduce_scratch_elems(int64_t n, int64_t c, int64_t hw, int n_regions) {
  using namespace ppst;
  if (n <= 0 || c <= 0 || hw <= 0 || n_regions <= 0) return 0;
  // upper bound over both layouts
  const ReducePlan a = plan_reduce(n, c, hw, 0), b = plan_reduce(n, c, hw, 1);
  const int chunks = a.chunks > b.chunks ? a.chunks : b.chunks;
  return n * c * (int64_t)chunks * n_regions * 3;
}

extern "C" int ppst_plane_reduce(float* sum, float* sumsq, float* vmax, const void* x, const uint8_t* labels,
                                 const float* shift, float* scratch, int64_t n, int64_t c, int64_t hw, int n_regions,
                                 int label_mode, int channels_last, int dtype, int w_row, int border, void* stream) {
  using namespace ppst;
  PPST_REQUIRE(x && scratch && n > 0 && c > 0 && hw > 0, PPST_ERR_ARG, "plane_reduce: bad args");
  PPST_REQUIRE(border == 0 || (border > 0 && border <= 8 && channels_last && w_row > 0 && hw % w_row == 0 && hw < (1LL << 31)), PPST_ERR_ARG,
               "plane_reduce: a border needs channels_last and the row width");
  PPST_REQUIRE(n_regions >= 1 && n_regions <= kMaxRegions, PPST_ERR_UNSUPPORTED, "plane_reduce: n_regions=%d (1..%d)",
               n_regions, kMaxRegions);
  PPST_REQUIRE((n_regions == 1) == (labels == nullptr), PPST_ERR_ARG,
               "plane_reduce: a label map needs n_regions >= 2, no label map needs n_regions == 1");
  cudaStream_t st = (cudaStream_t)stream;
#define PPST_RED(R_)                                                                                      \
  PPST_DISPATCH_DTYPE(dtype, T, return launch_reduce<T, R_>(sum, sumsq, vmax, (const T*)x, labels, shift, \
                                                            scratch, n, c, hw, label_mode, channels_last, w_row,   \
                                                            border, st))
  switch (n_regions) {
    case 1: PPST_RED(1);
    case 2: PPST_RED(2);
    case 3: PPST_RED(3);
    default: PPST_RED(4);
  }
#undef PPST_RED
  return 0;
}

extern "C" int64_t ppst_upsample_reduce_scratch_elems(int64_t n, int lh, int c) {
  if (n <= 0 || lh <= 0 || c <= 0) return 0;
  return n * (int64_t)(lh + 1) * c * 3;
}

extern "C" int ppst_upsample_reduce(float* sum, float* vmax, const float* x, int64_t n, int lh, int lw, int c_stride,
                                    int c_off, int c, int scale, float* scratch, void* stream) {
  using namespace ppst;
  PPST_REQUIRE(x && scratch && (sum || vmax) && n > 0 && lh > 0 && lw > 0 && c > 0, PPST_ERR_ARG, "upsample_reduce: bad args");
  PPST_REQUIRE(scale == 1 || scale == 2 || scale == 4 || scale == 8, PPST_ERR_UNSUPPORTED,
               "upsample_reduce: scale %d (1, 2, 4 or 8)", scale);
  PPST_REQUIRE(c % 4 == 0 && c <= 1024 && c_stride % 4 == 0 && c_off % 4 == 0 && c_off >= 0 && c_off + c <= c_stride &&
               (reinterpret_cast<uintptr_t>(x) & 15) == 0, PPST_ERR_UNSUPPORTED,
               "upsample_reduce: channel slice must be 16-byte aligned, c a multiple of 4 and <= 1024");
  PPST_REQUIRE(n <= 65535, PPST_ERR_ARG, "upsample_reduce: batch too large");
  cudaStream_t st = (cudaStream_t)stream;
  const int rows = scale > 1 ? lh + 1 : lh;
  const int cv = c / 4;
  const int lanes = 256 / cv > 0 ? 256 / cv : 1;
  PPST_REQUIRE(cv <= 256, PPST_ERR_UNSUPPORTED, "upsample_reduce: c too large");
  const size_t smem = sizeof(float) * 2 * (size_t)lanes * c;
  upsample_reduce_kernel<<<dim3(rows, (unsigned)n), 256, smem, st>>>(scratch, x, lh, lw, c_stride, c_off, c, scale, rows);
  PPST_LAUNCH_CHECK("upsample_reduce_kernel");
  plane_reduce_nhwc_final<1><<<ceil_div(n * c, 8), 256, 0, st>>>(sum, nullptr, vmax, scratch, n, c, rows);
  PPST_LAUNCH_CHECK("plane_reduce_nhwc_final");
  return 0;
}

extern "C" int ppst_labels_to_onehot(float* onehot, const uint8_t* labels, int64_t n, int64_t hw, int n_labels,
                                     void* stream) {
  using namespace ppst;
  PPST_REQUIRE(onehot && labels && n > 0 && hw > 0 && n_labels > 0, PPST_ERR_ARG, "labels_to_onehot: bad args");
  const int64_t total = n * n_labels * hw;
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  labels_to_onehot_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(onehot, labels, n, hw, n_labels);
  PPST_LAUNCH_CHECK("labels_to_onehot_kernel");
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Backward reductions of InstanceNorm (+ StyleMod) in the training path: per (sample, channel)
//   sum_g = sum_p g[p],   sum_gx = sum_p g[p] * x[p]       over the h*w pixels of NHWC bf16 / f32 tensors g, x.
// Everything the gradient of y = (x - mean) * rstd * (1 + s) + b needs (nn.InstanceNorm2d + StyleMod,
// stylegan2_layers.py:361-374,414-437) follows from these two and the forward statistics.  Two stages, fixed order.
namespace ppst {
constexpr int kDotChunk = 512;           // pixels per block: (hw / 512) x n blocks keep every SM busy at 64x64 and up
template <typename T>
__global__ void __launch_bounds__(256) plane_dot_kernel(float* __restrict__ partials, const T* __restrict__ g,
                                                        const T* __restrict__ x, int64_t c, int64_t hw, int chunks) {
  extern __shared__ float red[];                         // [pixel lanes][c][2]
  constexpr int VEC = Vec16<T>::N;
  const int cols = (int)(c / VEC);                       // vector columns, <= 256
  const int lanes = 256 / cols;
  const int col = threadIdx.x % cols, lane = threadIdx.x / cols;
  const int64_t ni = blockIdx.y;
  const int64_t beg = (int64_t)blockIdx.x * kDotChunk, end = min(beg + (int64_t)kDotChunk, hw);
  float sg[VEC], sx[VEC];
#pragma unroll
  for (int e = 0; e < VEC; ++e) { sg[e] = 0.f; sx[e] = 0.f; }
  if (lane < lanes) {
    const T* gb = g + ni * hw * c + (int64_t)col * VEC;
    const T* xb = x + ni * hw * c + (int64_t)col * VEC;
    // four pixels per iteration: all eight 16-byte loads are issued before the first use
    int64_t p = beg + lane;
    for (; p + 3 * (int64_t)lanes < end; p += 4 * (int64_t)lanes) {
      Vec16<T> gv[4], xv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        gv[u] = ld_vec16_stream(gb + (p + u * (int64_t)lanes) * c);
        xv[u] = ld_vec16_stream(xb + (p + u * (int64_t)lanes) * c);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          const float gf = Cvt<T>::to_f(gv[u].v[e]);
          sg[e] += gf;
          sx[e] = fmaf(gf, Cvt<T>::to_f(xv[u].v[e]), sx[e]);
        }
    }
    for (; p < end; p += lanes) {
      const Vec16<T> gv = ld_vec16_stream(gb + p * c), xv = ld_vec16_stream(xb + p * c);
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        const float gf = Cvt<T>::to_f(gv.v[e]);
        sg[e] += gf;
        sx[e] = fmaf(gf, Cvt<T>::to_f(xv.v[e]), sx[e]);
      }
    }
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      red[((int64_t)lane * c + col * VEC + e) * 2 + 0] = sg[e];
      red[((int64_t)lane * c + col * VEC + e) * 2 + 1] = sx[e];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * c; i += 256) {
    float t = 0.f;
    for (int l = 0; l < lanes; ++l) t += red[(int64_t)l * c * 2 + i];
    partials[((ni * chunks + blockIdx.x) * c) * 2 + i] = t;
  }
}

__global__ void plane_dot_final(float* __restrict__ sum_g, float* __restrict__ sum_gx, const float* __restrict__ partials,
                                int64_t n, int64_t c, int chunks) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * c) return;
  const int64_t ni = t / c, ch = t - ni * c;
  float a = 0.f, b = 0.f;
  for (int k = 0; k < chunks; ++k) {
    a += partials[((ni * chunks + k) * c + ch) * 2 + 0];
    b += partials[((ni * chunks + k) * c + ch) * 2 + 1];
  }
  sum_g[t] = a;
  sum_gx[t] = b;
}
}  // namespace ppst

extern "C" int64_t ppst_plane_dot_scratch_elems(int64_t n, int64_t c, int64_t hw) {
  return n * ((hw + ppst::kDotChunk - 1) / ppst::kDotChunk) * c * 2;
}

extern "C" int ppst_plane_dot(float* sum_g, float* sum_gx, const void* g, const void* x, float* scratch, int64_t n, int64_t c,
                              int64_t hw, int dtype, void* stream) {
  using namespace ppst;
  PPST_REQUIRE(sum_g && sum_gx && g && x && scratch && n > 0 && c > 0 && hw > 0, PPST_ERR_ARG, "plane_dot: bad args");
  PPST_REQUIRE(dtype == PPST_BF16 || dtype == PPST_F32, PPST_ERR_DTYPE, "plane_dot: dtype %d", dtype);
  const int vec = dtype == PPST_BF16 ? 8 : 4;
  PPST_REQUIRE(c % vec == 0 && c / vec <= 256, PPST_ERR_UNSUPPORTED,
               "plane_dot: c=%lld (a multiple of %d, at most %d)", (long long)c, vec, 256 * vec);
  PPST_REQUIRE(((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(x)) & 15) == 0, PPST_ERR_ARG,
               "plane_dot: buffers must be 16-byte aligned");
  const int chunks = (int)((hw + kDotChunk - 1) / kDotChunk);
  const size_t smem = sizeof(float) * (size_t)(256 / (c / vec)) * c * 2;
  dim3 grid((unsigned)chunks, (unsigned)n);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == PPST_BF16) {
    auto kern = plane_dot_kernel<__nv_bfloat16>;
    if (smem > 48 * 1024) {
      int rc = cuda_check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "plane_dot attr");
      if (rc) return rc;
    }
    kern<<<grid, 256, smem, st>>>(scratch, (const __nv_bfloat16*)g, (const __nv_bfloat16*)x, c, hw, chunks);
  } else {
    auto kern = plane_dot_kernel<float>;
    if (smem > 48 * 1024) {
      int rc = cuda_check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "plane_dot attr");
      if (rc) return rc;
    }
    kern<<<grid, 256, smem, st>>>(scratch, (const float*)g, (const float*)x, c, hw, chunks);
  }
  PPST_LAUNCH_CHECK("plane_dot_kernel");
  plane_dot_final<<<ceil_div(n * c, 256), 256, 0, st>>>(sum_g, sum_gx, scratch, n, c, chunks);
  PPST_LAUNCH_CHECK("plane_dot_final");
  return 0;
}


// ------------------------------------------------------------------------------------------------
// Backward of the global average / max pooling pair of the colour encoder's heads (GAP || GMP, encoder_col.py:39-40,
// 160-161) for the training path: dx[n,p,c] = ga[n,c] + (x[n,p,c] == mx[n,c] ? gm[n,c] : 0), where the caller passes
// ga = grad_gap / (h*w) and gm = grad_gmp / count and count = the number of pixels that attain the maximum (torch.amax
// spreads the gradient evenly over ties).  ppst_plane_count_eq produces that count (two stages, fixed order).
namespace ppst {
template <typename T>
__global__ void __launch_bounds__(256) plane_count_eq_kernel(float* __restrict__ partials, const T* __restrict__ x,
                                                             const float* __restrict__ mx, int64_t c, int64_t hw, int chunks) {
  extern __shared__ float red[];                         // [pixel lanes][c]
  constexpr int VEC = Vec16<T>::N;
  const int cols = (int)(c / VEC), lanes = 256 / cols;
  const int col = threadIdx.x % cols, lane = threadIdx.x / cols;
  const int64_t ni = blockIdx.y;
  const int64_t beg = (int64_t)blockIdx.x * kDotChunk, end = min(beg + (int64_t)kDotChunk, hw);
  float cnt[VEC], m[VEC];
#pragma unroll
  for (int e = 0; e < VEC; ++e) { cnt[e] = 0.f; m[e] = (lane < lanes) ? mx[ni * c + (int64_t)col * VEC + e] : 0.f; }
  if (lane < lanes) {
    const T* xb = x + ni * hw * c + (int64_t)col * VEC;
    for (int64_t p = beg + lane; p < end; p += lanes) {
      const Vec16<T> xv = ld_vec16_stream(xb + p * c);
#pragma unroll
      for (int e = 0; e < VEC; ++e) cnt[e] += (Cvt<T>::to_f(xv.v[e]) == m[e]) ? 1.0f : 0.0f;
    }
#pragma unroll
    for (int e = 0; e < VEC; ++e) red[(int64_t)lane * c + col * VEC + e] = cnt[e];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < c; i += 256) {
    float t = 0.f;
    for (int l = 0; l < lanes; ++l) t += red[(int64_t)l * c + i];
    partials[(ni * chunks + blockIdx.x) * c + i] = t;
  }
}

__global__ void plane_count_final(float* __restrict__ cnt, const float* __restrict__ partials, int64_t n, int64_t c, int chunks) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * c) return;
  const int64_t ni = t / c, ch = t - ni * c;
  float a = 0.f;
  for (int k = 0; k < chunks; ++k) a += partials[(ni * chunks + k) * c + ch];
  cnt[t] = a;
}

template <typename T>
__global__ void __launch_bounds__(256) pool_backward_kernel(T* __restrict__ dx, const T* __restrict__ x,
                                                            const float* __restrict__ ga, const float* __restrict__ gm,
                                                            const float* __restrict__ mx, int64_t c, int64_t hw, int64_t total) {
  constexpr int VEC = Vec16<T>::N;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // over n*hw*(c/VEC)
  if (idx >= total) return;
  const int64_t cv = c / VEC;
  const int64_t v = idx % cv, pix = idx / cv, ni = pix / hw;
  const int64_t off = pix * c + v * VEC, co = ni * c + v * VEC;
  const Vec16<T> xv = ld_vec16_stream(x + off);
  Vec16<T> o;
#pragma unroll
  for (int e = 0; e < VEC; ++e)
    o.v[e] = Cvt<T>::from_f(ga[co + e] + ((Cvt<T>::to_f(xv.v[e]) == mx[co + e]) ? gm[co + e] : 0.0f));
  st_vec16(dx + off, o);
}
}  // namespace ppst

extern "C" int ppst_plane_count_eq(float* cnt, const void* x, const float* mx, float* scratch, int64_t n, int64_t c, int64_t hw,
                                   int dtype, void* stream) {
  using namespace ppst;
  PPST_REQUIRE(cnt && x && mx && scratch && n > 0 && c > 0 && hw > 0, PPST_ERR_ARG, "plane_count_eq: bad args");
  PPST_REQUIRE(dtype == PPST_BF16 || dtype == PPST_F32, PPST_ERR_DTYPE, "plane_count_eq: dtype %d", dtype);
  const int vec = dtype == PPST_BF16 ? 8 : 4;
  PPST_REQUIRE(c % vec == 0 && c / vec <= 256, PPST_ERR_UNSUPPORTED, "plane_count_eq: c=%lld", (long long)c);
  const int chunks = (int)((hw + kDotChunk - 1) / kDotChunk);
  const size_t smem = sizeof(float) * (size_t)(256 / (c / vec)) * c;
  dim3 grid((unsigned)chunks, (unsigned)n);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == PPST_BF16)
    plane_count_eq_kernel<__nv_bfloat16><<<grid, 256, smem, st>>>(scratch, (const __nv_bfloat16*)x, mx, c, hw, chunks);
  else
    plane_count_eq_kernel<float><<<grid, 256, smem, st>>>(scratch, (const float*)x, mx, c, hw, chunks);
  PPST_LAUNCH_CHECK("plane_count_eq_kernel");
  plane_count_final<<<ceil_div(n * c, 256), 256, 0, st>>>(cnt, scratch, n, c, chunks);
  PPST_LAUNCH_CHECK("plane_count_final");
  return 0;
}

extern "C" int ppst_pool_backward(void* dx, const void* x, const float* ga, const float* gm, const float* mx, int64_t n, int64_t c,
                                  int64_t hw, int dtype, void* stream) {
  using namespace ppst;
  PPST_REQUIRE(dx && x && ga && gm && mx && n > 0 && c > 0 && hw > 0, PPST_ERR_ARG, "pool_backward: bad args");
  PPST_REQUIRE(dtype == PPST_BF16 || dtype == PPST_F32, PPST_ERR_DTYPE, "pool_backward: dtype %d", dtype);
  const int vec = dtype == PPST_BF16 ? 8 : 4;
  PPST_REQUIRE(c % vec == 0, PPST_ERR_UNSUPPORTED, "pool_backward: c=%lld", (long long)c);
  const int64_t total = n * hw * (c / vec);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == PPST_BF16)
    pool_backward_kernel<__nv_bfloat16><<<(unsigned)((total + 255) / 256), 256, 0, st>>>((__nv_bfloat16*)dx, (const __nv_bfloat16*)x,
                                                                                        ga, gm, mx, c, hw, total);
  else
    pool_backward_kernel<float><<<(unsigned)((total + 255) / 256), 256, 0, st>>>((float*)dx, (const float*)x, ga, gm, mx, c, hw, total);
  PPST_LAUNCH_CHECK("pool_backward_kernel");
  return 0;
}
