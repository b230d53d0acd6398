// Small dense layers, vector normalisation, exact-f32 batched GEMM and row softmax.
#include <float.h>

#include "common.cuh"

namespace ppst {

// ---- skinny linear: y[m, n] = x[m, k] . w[n, k]^T (+ bias), m = batch is tiny ------------------------------------
// Weight-bandwidth bound.  A block owns 4 output columns x 16 batch rows; its 8 warps split K (split-K inside the block
// keeps >= n/4 blocks in flight even for the 512-column style layers), each lane streaming 128-bit weight loads with all
// of its loads issued before the first use.  Partials are combined in a fixed order (lane chain, transposing warp
// butterfly, warps 0..7), so a row's result does not depend on the batch it is in.
constexpr int kLinRows = 16, kLinCols = 4;
__device__ __forceinline__ float dot4(const float4& x, const float4& w, float acc) {
  acc = fmaf(x.x, w.x, acc); acc = fmaf(x.y, w.y, acc);
  acc = fmaf(x.z, w.z, acc); acc = fmaf(x.w, w.w, acc);
  return acc;
}
__global__ void __launch_bounds__(256) linear_kernel(float* __restrict__ y, const float* __restrict__ x,
                                                     const float* __restrict__ w, const float* __restrict__ b,
                                                     int64_t m, int64_t n, int64_t k, int64_t ldx, float w_scale,
                                                     float b_scale, int relu_in) {
  __shared__ float part[8][kLinCols * kLinRows];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t c0 = (int64_t)blockIdx.x * kLinCols;
  const int64_t m0 = (int64_t)blockIdx.y * kLinRows;
  const int rows = (int)((m - m0) < kLinRows ? (m - m0) : kLinRows);
  const float* wp[kLinCols];
#pragma unroll
  for (int c = 0; c < kLinCols; ++c) wp[c] = w + (c0 + c < n ? c0 + c : c0) * k;
  float acc[kLinCols * kLinRows];
#pragma unroll
  for (int i = 0; i < kLinCols * kLinRows; ++i) acc[i] = 0.0f;
  const bool vec = (k % 4 == 0) && (ldx % 4 == 0) &&
                   ((reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(x)) & 15) == 0;
  if (vec) {
    // warp `warp` owns k in [warp*128 + 1024*j, +128): consecutive warps read consecutive 512-byte weight segments
    for (int64_t kk = warp * 128 + lane * 4; kk < k; kk += 1024) {
      float4 wv[kLinCols];
#pragma unroll
      for (int c = 0; c < kLinCols; ++c) wv[c] = *reinterpret_cast<const float4*>(wp[c] + kk);
#pragma unroll
      for (int r = 0; r < kLinRows; ++r) {
        if (r < rows) {
          float4 xv = *reinterpret_cast<const float4*>(x + (m0 + r) * ldx + kk);
          if (relu_in) { xv.x = fmaxf(xv.x, 0.f); xv.y = fmaxf(xv.y, 0.f); xv.z = fmaxf(xv.z, 0.f); xv.w = fmaxf(xv.w, 0.f); }
#pragma unroll
          for (int c = 0; c < kLinCols; ++c) acc[c * kLinRows + r] = dot4(xv, wv[c], acc[c * kLinRows + r]);
        }
      }
    }
  } else {
    for (int64_t kk = warp * 32 + lane; kk < k; kk += 256) {
      float wv[kLinCols];
#pragma unroll
      for (int c = 0; c < kLinCols; ++c) wv[c] = wp[c][kk];
#pragma unroll
      for (int r = 0; r < kLinRows; ++r)
        if (r < rows) {
          float xv = x[(m0 + r) * ldx + kk];
          if (relu_in) xv = fmaxf(xv, 0.f);
#pragma unroll
          for (int c = 0; c < kLinCols; ++c) acc[c * kLinRows + r] = fmaf(xv, wv[c], acc[c * kLinRows + r]);
        }
    }
  }
  // lane L ends up with the warp totals of acc[L] and acc[32 + L]
  const float t0 = warp_col_reduce(*reinterpret_cast<float(*)[32]>(&acc[0]), lane);
  const float t1 = warp_col_reduce(*reinterpret_cast<float(*)[32]>(&acc[32]), lane);
  part[warp][lane] = t0;
  part[warp][32 + lane] = t1;
  __syncthreads();
  if (threadIdx.x < kLinCols * kLinRows) {
    const int c = threadIdx.x / kLinRows, r = threadIdx.x % kLinRows;
    float s = part[0][threadIdx.x];
#pragma unroll
    for (int i = 1; i < 8; ++i) s += part[i][threadIdx.x];
    if (r < rows && c0 + c < n) y[(m0 + r) * n + c0 + c] = s * w_scale + (b ? b[c0 + c] * b_scale : 0.0f);
  }
}

__global__ void __launch_bounds__(256) l2_normalize_kernel(float* __restrict__ y, const float* __restrict__ x,
                                                           int64_t cols, float eps, int mode) {
  const int64_t row = blockIdx.x;
  const float* xr = x + row * cols;
  float acc = 0.f;
  for (int64_t i = threadIdx.x; i < cols; i += blockDim.x) acc = fmaf(xr[i], xr[i], acc);
  __shared__ float sm[8];
  __shared__ float inv;
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += sm[i];
    inv = (mode == 0) ? rsqrtf(t + eps) : 1.0f / fmaxf(sqrtf(t), eps);
  }
  __syncthreads();
  for (int64_t i = threadIdx.x; i < cols; i += blockDim.x) y[row * cols + i] = xr[i] * inv;
}

__global__ void lerp_kernel(float* __restrict__ y, const float* __restrict__ a, const float* __restrict__ b, float r,
                            int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = a[i] * (1.0f - r) + b[i] * r;
}

// ---- batched f32 GEMM, 128x128x16 block tile, 8x8 register tile, 256 threads -------------------------------
constexpr int GBM = 128, GBN = 128, GBK = 16;
template <bool TRANS_B>
__global__ void __launch_bounds__(256) gemm_f32_kernel(float* __restrict__ c, const float* __restrict__ a,
                                                       const float* __restrict__ b, int64_t m, int64_t n, int64_t k,
                                                       int64_t lda, int64_t ldb, int64_t ldc, int64_t stride_a,
                                                       int64_t stride_b, int64_t stride_c, float alpha) {
  __shared__ __align__(16) float As[2][GBK][GBM + 4];
  __shared__ __align__(16) float Bs[2][GBK][GBN + 4];
  const int tid = threadIdx.x;
  a += (int64_t)blockIdx.z * stride_a;
  b += (int64_t)blockIdx.z * stride_b;
  c += (int64_t)blockIdx.z * stride_c;
  const int64_t m0 = (int64_t)blockIdx.y * GBM, n0 = (int64_t)blockIdx.x * GBN;
  const int tx = tid & 15, ty = tid >> 4;                 // 16 x 16 threads, each 8x8 outputs (strided by 16)
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  // loaders: A tile 128 x 16 (row-major, k contiguous): 2048 elems / 256 thr = 8 each: row = tid>>1, k-off = (tid&1)*8
  const int a_row = tid >> 1, a_k = (tid & 1) * 8;
  // B tile: TRANS_B: [n][k] like A;  else [k][n]: 16 x 128: row(k) = tid>>4, n-off = (tid&15)*8
  const int b_row = TRANS_B ? (tid >> 1) : (tid >> 4), b_off = TRANS_B ? (tid & 1) * 8 : (tid & 15) * 8;
  float ra[8], rb[8];
  auto gload = [&](int64_t k0) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int64_t gr = m0 + a_row, gk = k0 + a_k + e;
      ra[e] = (gr < m && gk < k) ? a[gr * lda + gk] : 0.f;
    }
    if (TRANS_B) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int64_t gr = n0 + b_row, gk = k0 + b_off + e;
        rb[e] = (gr < n && gk < k) ? b[gr * ldb + gk] : 0.f;
      }
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int64_t gk = k0 + b_row, gc = n0 + b_off + e;
        rb[e] = (gk < k && gc < n) ? b[gk * ldb + gc] : 0.f;
      }
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int e = 0; e < 8; ++e) As[buf][a_k + e][a_row] = ra[e];
    if (TRANS_B) {
#pragma unroll
      for (int e = 0; e < 8; ++e) Bs[buf][b_off + e][b_row] = rb[e];
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) Bs[buf][b_row][b_off + e] = rb[e];
    }
  };
  gload(0);
  sstore(0);
  __syncthreads();
  int buf = 0;
  for (int64_t k0 = 0; k0 < k; k0 += GBK) {
    const bool more = (k0 + GBK) < k;
    if (more) gload(k0 + GBK);
#pragma unroll
    for (int kk = 0; kk < GBK; ++kk) {
      float av[8], bv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) av[i] = As[buf][kk][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < 8; ++j) bv[j] = Bs[buf][kk][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (more) {
      sstore(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t gr = m0 + ty + 16 * i;
    if (gr >= m) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int64_t gc = n0 + tx + 16 * j;
      if (gc < n) c[gr * ldc + gc] = acc[i][j] * alpha;
    }
  }
}

// ---- in-place row softmax of (x / divisor); one block per row ---------------------------------------------
__global__ void __launch_bounds__(256) softmax_rows_kernel(float* __restrict__ x, int64_t cols, float divisor) {
  float* xr = x + (int64_t)blockIdx.x * cols;
  __shared__ float sm[8];
  __shared__ float bc;
  float mx = -FLT_MAX;
  for (int64_t i = threadIdx.x; i < cols; i += blockDim.x) mx = fmaxf(mx, xr[i] / divisor);
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = sm[0];
    for (int i = 1; i < 8; ++i) t = fmaxf(t, sm[i]);
    bc = t;
  }
  __syncthreads();
  mx = bc;
  float s = 0.f;
  for (int64_t i = threadIdx.x; i < cols; i += blockDim.x) {
    const float e = expf(xr[i] / divisor - mx);
    xr[i] = e;
    s += e;
  }
  s = warp_sum(s);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += sm[i];
    bc = t;
  }
  __syncthreads();
  const float tot = bc;
  for (int64_t i = threadIdx.x; i < cols; i += blockDim.x) xr[i] = xr[i] / tot;
}

// ---- backward of the skinny linear (training): the batch m is tiny, the weights are the traffic ------
// dw[n][k] = sum_m gy[m][n] * x[m][k]: one thread per (n, 4 consecutive k); x (m x k) stays in L1 / L2.
__global__ void __launch_bounds__(256) linear_wgrad_kernel(float* __restrict__ dw, const float* __restrict__ gy,
                                                           const float* __restrict__ x, int m, int64_t n, int64_t k) {
  const int64_t k4 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int64_t row = blockIdx.y;
  if (k4 >= k) return;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = 0; i < m; ++i) {
    const float g = gy[(int64_t)i * n + row];
    const float4 xv = *reinterpret_cast<const float4*>(x + (int64_t)i * k + k4);
    acc.x = fmaf(g, xv.x, acc.x); acc.y = fmaf(g, xv.y, acc.y); acc.z = fmaf(g, xv.z, acc.z); acc.w = fmaf(g, xv.w, acc.w);
  }
  *reinterpret_cast<float4*>(dw + row * k + k4) = acc;
}

// dx[m][k] = sum_n gy[m][n] * w[n][k]: a block owns 128 columns of k (32 lanes x float4), 8 batch rows and one slice of n;
// its 8 warps interleave over the slice, every weight element is read once per 8 rows.  Partials [slice][m][k] are
// summed in slice order by linear_dgrad_final.
constexpr int kDgRows = 8;
__global__ void __launch_bounds__(256) linear_dgrad_kernel(float* __restrict__ partial, const float* __restrict__ gy,
                                                           const float* __restrict__ w, int m, int64_t n, int64_t k,
                                                           int64_t n_per_slice) {
  __shared__ float red[8][kDgRows][128];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t k4 = (int64_t)blockIdx.x * 128 + lane * 4;
  const int64_t n0 = (int64_t)blockIdx.y * n_per_slice, n1 = min(n0 + n_per_slice, n);
  const int m0 = blockIdx.z * kDgRows;
  float acc[kDgRows][4];
#pragma unroll
  for (int r = 0; r < kDgRows; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f;
  if (k4 < k) {
    for (int64_t j = n0 + warp; j < n1; j += 8) {
      const float4 wv = *reinterpret_cast<const float4*>(w + j * k + k4);
#pragma unroll
      for (int r = 0; r < kDgRows; ++r) {
        const float g = (m0 + r < m) ? gy[(int64_t)(m0 + r) * n + j] : 0.f;
        acc[r][0] = fmaf(g, wv.x, acc[r][0]); acc[r][1] = fmaf(g, wv.y, acc[r][1]);
        acc[r][2] = fmaf(g, wv.z, acc[r][2]); acc[r][3] = fmaf(g, wv.w, acc[r][3]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < kDgRows; ++r)
    *reinterpret_cast<float4*>(&red[warp][r][lane * 4]) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
  __syncthreads();
  for (int i = threadIdx.x; i < kDgRows * 128; i += 256) {
    const int r = i / 128, c = i - r * 128;
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += red[q][r][c];
    const int64_t kk = (int64_t)blockIdx.x * 128 + c;
    if (m0 + r < m && kk < k) partial[((int64_t)blockIdx.y * m + m0 + r) * k + kk] = t;
  }
}

__global__ void linear_dgrad_final(float* __restrict__ dx, const float* __restrict__ partial, int64_t mk, int slices) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= mk) return;
  float t = 0.f;
  for (int s = 0; s < slices; ++s) t += partial[(int64_t)s * mk + i];
  dx[i] = t;
}

}  // namespace ppst

static int linear_dgrad_slices(int64_t n) { return n >= 512 ? 8 : (n >= 64 ? 2 : 1); }

extern "C" int64_t ppst_linear_dgrad_scratch_elems(int64_t m, int64_t n, int64_t k) { return linear_dgrad_slices(n) * m * k; }

extern "C" int ppst_linear_dgrad(float* dx, const float* gy, const float* w, float* scratch, int64_t m, int64_t n, int64_t k,
                                 void* stream) {
  using namespace ppst;
  PPST_REQUIRE(dx && gy && w && scratch && m > 0 && n > 0 && k > 0 && k % 4 == 0, PPST_ERR_ARG, "linear_dgrad: bad args (k %% 4)");
  const int slices = linear_dgrad_slices(n);
  const int64_t per = (n + slices - 1) / slices;
  dim3 grid((unsigned)ceil_div(k, 128), (unsigned)slices, (unsigned)ceil_div(m, kDgRows));
  cudaStream_t st = (cudaStream_t)stream;
  linear_dgrad_kernel<<<grid, 256, 0, st>>>(scratch, gy, w, (int)m, n, k, per);
  PPST_LAUNCH_CHECK("linear_dgrad_kernel");
  linear_dgrad_final<<<ceil_div(m * k, 256), 256, 0, st>>>(dx, scratch, m * k, slices);
  PPST_LAUNCH_CHECK("linear_dgrad_final");
  return 0;
}

extern "C" int ppst_linear_wgrad(float* dw, const float* gy, const float* x, int64_t m, int64_t n, int64_t k, void* stream) {
  using namespace ppst;
  PPST_REQUIRE(dw && gy && x && m > 0 && n > 0 && k > 0 && k % 4 == 0 && n <= 65535, PPST_ERR_ARG,
               "linear_wgrad: bad args (k %% 4, n <= 65535)");
  dim3 grid((unsigned)ceil_div(k / 4, 256), (unsigned)n);
  linear_wgrad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dw, gy, x, (int)m, n, k);
  PPST_LAUNCH_CHECK("linear_wgrad_kernel");
  return 0;
}

extern "C" int ppst_linear(float* y, const float* x, const float* w, const float* b, int64_t m, int64_t n, int64_t k,
                           int64_t ldx, float w_scale, float b_scale, int relu_in, void* stream) {
  using namespace ppst;
  PPST_REQUIRE(y && x && w && m > 0 && n > 0 && k > 0 && ldx >= k, PPST_ERR_ARG, "linear: bad args");
  dim3 grid(ceil_div(n, kLinCols), ceil_div(m, kLinRows));
  PPST_REQUIRE(grid.y <= 65535, PPST_ERR_ARG, "linear: too many rows");
  linear_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(y, x, w, b, m, n, k, ldx, w_scale, b_scale, relu_in);
  PPST_LAUNCH_CHECK("linear_kernel");
  return 0;
}

extern "C" int ppst_l2_normalize(float* y, const float* x, int64_t rows, int64_t cols, float eps, int mode,
                                 void* stream) {
  using namespace ppst;
  PPST_REQUIRE(y && x && rows > 0 && cols > 0 && (mode == 0 || mode == 1), PPST_ERR_ARG, "l2_normalize: bad args");
  l2_normalize_kernel<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(y, x, cols, eps, mode);
  PPST_LAUNCH_CHECK("l2_normalize_kernel");
  return 0;
}

extern "C" int ppst_lerp(float* y, const float* a, const float* b, float r, int64_t n, void* stream) {
  using namespace ppst;
  PPST_REQUIRE(y && a && b && n > 0, PPST_ERR_ARG, "lerp: bad args");
  lerp_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(y, a, b, r, n);
  PPST_LAUNCH_CHECK("lerp_kernel");
  return 0;
}

extern "C" int ppst_gemm_f32(float* c, const float* a, const float* b, int64_t batch, int64_t m, int64_t n, int64_t k,
                             int64_t lda, int64_t ldb, int64_t ldc, int64_t stride_a, int64_t stride_b,
                             int64_t stride_c, int trans_b, float alpha, void* stream) {
  using namespace ppst;
  PPST_REQUIRE(c && a && b && batch > 0 && m > 0 && n > 0 && k > 0, PPST_ERR_ARG, "gemm_f32: bad args");
  dim3 grid(ceil_div(n, GBN), ceil_div(m, GBM), (unsigned)batch);
  PPST_REQUIRE(grid.y <= 65535 && batch <= 65535, PPST_ERR_ARG, "gemm_f32: grid too large");
  cudaStream_t st = (cudaStream_t)stream;
  if (trans_b)
    gemm_f32_kernel<true><<<grid, 256, 0, st>>>(c, a, b, m, n, k, lda, ldb, ldc, stride_a, stride_b, stride_c, alpha);
  else
    gemm_f32_kernel<false><<<grid, 256, 0, st>>>(c, a, b, m, n, k, lda, ldb, ldc, stride_a, stride_b, stride_c, alpha);
  PPST_LAUNCH_CHECK("gemm_f32_kernel");
  return 0;
}

extern "C" int ppst_softmax_rows(float* x, int64_t rows, int64_t cols, float divisor, void* stream) {
  using namespace ppst;
  PPST_REQUIRE(x && rows > 0 && cols > 0 && divisor != 0.0f, PPST_ERR_ARG, "softmax_rows: bad args");
  softmax_rows_kernel<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(x, cols, divisor);
  PPST_LAUNCH_CHECK("softmax_rows_kernel");
  return 0;
}
