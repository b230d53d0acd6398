// Correlation-swap helpers (models/ppst_model.py:330-364): Rselfcorr and the corrm pre-normalisation.
#include <float.h>

#include "common.cuh"

namespace ppst {

// One 256-thread block per 4x4 patch.  fea NHWC [n, h, w, 64]; out NHWC slice [n, h/4, w/4, 256]:
//   p[c][i]  (c: channel, i: position in the patch, row-major 4x4)
//   p -= mean_c p ;  p /= (||p||_c + eps) ;  out[i*16 + j] = sum_c p[c][i] * p[c][j]
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) rselfcorr_kernel(TO* __restrict__ out, const TI* __restrict__ fea, int h, int w,
                                                        int c, int out_c_stride, int c_off) {
  extern __shared__ float sp[];                 // [16][c + 1]
  __shared__ float mean[16], inv[16];
  const int pw = w / 4, ph = h / 4;
  const int64_t patch = blockIdx.x;
  const int px = (int)(patch % pw), py = (int)((patch / pw) % ph);
  const int64_t ni = patch / ((int64_t)pw * ph);
  const int pitch = c + 1;
  for (int idx = threadIdx.x; idx < 16 * c; idx += blockDim.x) {
    const int pos = idx / c, ch = idx - pos * c;
    const int yy = py * 4 + pos / 4, xx = px * 4 + pos % 4;
    sp[pos * pitch + ch] = Cvt<TI>::to_f(fea[((ni * h + yy) * w + xx) * (int64_t)c + ch]);
  }
  __syncthreads();
  // 16 positions, 16 threads each: mean over channels then norm of the centred vector
  {
    const int pos = threadIdx.x >> 4, sub = threadIdx.x & 15;
    float s = 0.f;
    for (int ch = sub; ch < c; ch += 16) s += sp[pos * pitch + ch];
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mu = s / (float)c;
    float q = 0.f;
    for (int ch = sub; ch < c; ch += 16) {
      const float d = sp[pos * pitch + ch] - mu;
      q = fmaf(d, d, q);
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    if (sub == 0) {
      mean[pos] = mu;
      inv[pos] = 1.0f / (sqrtf(q) + (float)DBL_EPSILON);
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < 16 * c; idx += blockDim.x) {
    const int pos = idx / c, ch = idx - pos * c;
    sp[pos * pitch + ch] = (sp[pos * pitch + ch] - mean[pos]) * inv[pos];
  }
  __syncthreads();
  const int i = threadIdx.x >> 4, j = threadIdx.x & 15;
  float acc = 0.f;
  for (int ch = 0; ch < c; ++ch) acc = fmaf(sp[i * pitch + ch], sp[j * pitch + ch], acc);
  out[((ni * ph + py) * pw + px) * (int64_t)out_c_stride + c_off + threadIdx.x] = Cvt<TO>::from_f(acc);
}

// c == 64 (the only width PPSTModel.Rselfcorr is called with): HALF A WARP per patch, 8 patches per 128-thread block.
// Thread `pos` of a patch owns pixel `pos` of the 4x4 patch: it loads the pixel's 64 channels with 16-byte loads, centres
// and normalises them in registers (no shuffles), parks them in shared memory as [channel][position], and then computes a
// 4x4 register tile of the 16x16 Gram matrix with two LDS.128 per 16 FMAs (the one-block-per-patch kernel below spends
// two LDS.32 per FMA and 32768 blocks per batch of 8).
template <typename TO> __device__ __forceinline__ void store4(TO* p, const float (&v)[4]);
template <> __device__ __forceinline__ void store4<float>(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <> __device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, const float (&v)[4]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}
template <> __device__ __forceinline__ void store4<__half>(__half* p, const float (&v)[4]) {
  __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(128) rselfcorr64_kernel(TO* __restrict__ out, const TI* __restrict__ fea,
                                                          int64_t patches, int h, int w, int out_c_stride, int c_off) {
  constexpr int C = 64;
  __shared__ __align__(16) float sp[8][C][16];
  const int slot = threadIdx.x >> 4, pos = threadIdx.x & 15;
  const int64_t patch = (int64_t)blockIdx.x * 8 + slot;
  const bool live = patch < patches;
  const int pw = w / 4, ph = h / 4;
  if (live) {
    const int px = (int)(patch % pw), py = (int)((patch / pw) % ph);
    const int64_t ni = patch / ((int64_t)pw * ph);
    const int yy = py * 4 + (pos >> 2), xx = px * 4 + (pos & 3);
    const TI* src = fea + ((ni * h + yy) * w + xx) * (int64_t)C;
    float v[C];
    constexpr int VN = Vec16<TI>::N;
#pragma unroll
    for (int k = 0; k < C / VN; ++k) {
      const Vec16<TI> r = ld_vec16(src + k * VN);
#pragma unroll
      for (int e = 0; e < VN; ++e) v[k * VN + e] = Cvt<TI>::to_f(r.v[e]);
    }
    float s = 0.f;
#pragma unroll
    for (int ch = 0; ch < C; ++ch) s += v[ch];
    const float mu = s / (float)C;
    float q = 0.f;
#pragma unroll
    for (int ch = 0; ch < C; ++ch) { v[ch] -= mu; q = fmaf(v[ch], v[ch], q); }
    const float inv = 1.0f / (sqrtf(q) + (float)DBL_EPSILON);
#pragma unroll
    for (int ch = 0; ch < C; ++ch) sp[slot][ch][pos] = v[ch] * inv;
  }
  __syncwarp();
  if (!live) return;
  const int ti = pos >> 2, tj = pos & 3;
  float acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[r][e] = 0.f;
#pragma unroll 8
  for (int ch = 0; ch < C; ++ch) {
    const float4 a = *reinterpret_cast<const float4*>(&sp[slot][ch][4 * ti]);
    const float4 b = *reinterpret_cast<const float4*>(&sp[slot][ch][4 * tj]);
    const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[r][e] = fmaf(av[r], bv[e], acc[r][e]);
  }
  TO* o = out + patch * (int64_t)out_c_stride + c_off;
#pragma unroll
  for (int r = 0; r < 4; ++r) store4<TO>(o + (4 * ti + r) * 16 + 4 * tj, acc[r]);
}

// One warp per pixel row of c channels: centre the first c_center channels, then L2-normalise all of them.
template <typename TI>
__global__ void __launch_bounds__(256) corr_prepare_kernel(float* __restrict__ out, const TI* __restrict__ in,
                                                           int64_t rows, int c, int c_center) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const TI* p = in + row * c;
  float s = 0.f;
  for (int ch = lane; ch < c_center; ch += 32) s += Cvt<TI>::to_f(p[ch]);
  s = warp_sum(s);
  const float mu = c_center > 0 ? s / (float)c_center : 0.f;
  float q = 0.f;
  for (int ch = lane; ch < c; ch += 32) {
    const float v = Cvt<TI>::to_f(p[ch]) - (ch < c_center ? mu : 0.f);
    q = fmaf(v, v, q);
  }
  q = warp_sum(q);
  const float inv = 1.0f / (sqrtf(q) + (float)DBL_EPSILON);
  for (int ch = lane; ch < c; ch += 32)
    out[row * c + ch] = (Cvt<TI>::to_f(p[ch]) - (ch < c_center ? mu : 0.f)) * inv;
}

__device__ __forceinline__ void split2(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// corr_prepare with the result emitted as [hi | lo] bf16 halves (operand of the 3-term split GEMM).
template <typename TI>
__global__ void __launch_bounds__(256) corr_prepare_split_kernel(__nv_bfloat16* __restrict__ out, const TI* __restrict__ in,
                                                                 int64_t rows, int c, int c_center) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const TI* p = in + row * c;
  float s = 0.f;
  for (int ch = lane; ch < c_center; ch += 32) s += Cvt<TI>::to_f(p[ch]);
  s = warp_sum(s);
  const float mu = c_center > 0 ? s / (float)c_center : 0.f;
  float q = 0.f;
  for (int ch = lane; ch < c; ch += 32) {
    const float v = Cvt<TI>::to_f(p[ch]) - (ch < c_center ? mu : 0.f);
    q = fmaf(v, v, q);
  }
  q = warp_sum(q);
  const float inv = 1.0f / (sqrtf(q) + (float)DBL_EPSILON);
  __nv_bfloat16* o = out + row * 2 * c;
  for (int ch = lane; ch < c; ch += 32) {
    __nv_bfloat16 hi, lo;
    split2((Cvt<TI>::to_f(p[ch]) - (ch < c_center ? mu : 0.f)) * inv, hi, lo);
    o[ch] = hi;
    o[c + ch] = lo;
  }
}

// f32 [rows, cols] -> bf16 [rows, 2*cols] = [hi | lo]
__global__ void split_bf16_kernel(__nv_bfloat16* __restrict__ out, const float* __restrict__ in, int64_t rows, int64_t cols) {
  const int64_t total = rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols, cidx = i - r * cols;
    __nv_bfloat16 hi, lo;
    split2(in[i], hi, lo);
    out[r * 2 * cols + cidx] = hi;
    out[r * 2 * cols + cols + cidx] = lo;
  }
}

// f32 [n, L, C] -> bf16 [n, C, 2L] = per channel row [hi(L) | lo(L)]  (the K-major B operand of corr @ feat)
__global__ void __launch_bounds__(256) split_bf16_t_kernel(__nv_bfloat16* __restrict__ out, const float* __restrict__ in,
                                                           int64_t L, int64_t C) {
  __shared__ float tile[32][33];
  const int64_t n = blockIdx.z;
  const int64_t l0 = (int64_t)blockIdx.x * 32, c0 = (int64_t)blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int j = ty; j < 32; j += 8) {
    const int64_t l = l0 + j, c = c0 + tx;
    tile[j][tx] = (l < L && c < C) ? in[(n * L + l) * C + c] : 0.f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int64_t c = c0 + j, l = l0 + tx;
    if (c < C && l < L) {
      __nv_bfloat16 hi, lo;
      split2(tile[tx][j], hi, lo);
      out[(n * C + c) * 2 * L + l] = hi;
      out[(n * C + c) * 2 * L + L + l] = lo;
    }
  }
}

// block-wide max / sum over 256 threads (result broadcast); `sm` holds 8 partials + the result
__device__ __forceinline__ float block256_reduce(float v, bool is_max, float* sm) {
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = sm[0];
    for (int i = 1; i < 8; ++i) t = is_max ? fmaxf(t, sm[i]) : t + sm[i];
    sm[8] = t;
  }
  __syncthreads();
  return sm[8];
}

// softmax(x / divisor) per row as [hi | lo] bf16 halves, rows of <= 4096 values (the 64x64 correlation rows): the row
// lives in registers (4 float4 per thread), so the logits are read from HBM exactly once and the two halves leave as
// 8-byte stores.  One block per row.
__global__ void __launch_bounds__(256) softmax_rows_split_reg_kernel(__nv_bfloat16* __restrict__ out,
                                                                     const float* __restrict__ x, int64_t cols,
                                                                     float divisor) {
  const float* xr = x + (int64_t)blockIdx.x * cols;
  __nv_bfloat16* o = out + (int64_t)blockIdx.x * 2 * cols;
  __shared__ float sm[9];
  const int nvec = (int)(cols >> 2);
  float4 v[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int i = threadIdx.x + 256 * j;
    if (i < nvec) {
      const Vec16<float> r = ld_vec16_stream(xr + 4 * i);
      v[j] = make_float4(r.v[0] / divisor, r.v[1] / divisor, r.v[2] / divisor, r.v[3] / divisor);
    } else {
      v[j] = make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX);
    }
  }
  float mx = -FLT_MAX;
#pragma unroll
  for (int j = 0; j < 4; ++j) mx = fmaxf(fmaxf(fmaxf(mx, v[j].x), fmaxf(v[j].y, v[j].z)), v[j].w);
  mx = block256_reduce(mx, true, sm);
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const bool on = threadIdx.x + 256 * j < nvec;
    v[j].x = on ? expf(v[j].x - mx) : 0.f; v[j].y = on ? expf(v[j].y - mx) : 0.f;
    v[j].z = on ? expf(v[j].z - mx) : 0.f; v[j].w = on ? expf(v[j].w - mx) : 0.f;
    s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
  }
  const float tot = block256_reduce(s, false, sm);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int i = threadIdx.x + 256 * j;
    if (i < nvec) {
      union { uint2 u; __nv_bfloat16 h[4]; } hi, lo;
      split2(v[j].x / tot, hi.h[0], lo.h[0]);
      split2(v[j].y / tot, hi.h[1], lo.h[1]);
      split2(v[j].z / tot, hi.h[2], lo.h[2]);
      split2(v[j].w / tot, hi.h[3], lo.h[3]);
      *reinterpret_cast<uint2*>(o + 4 * i) = hi.u;
      *reinterpret_cast<uint2*>(o + cols + 4 * i) = lo.u;
    }
  }
}

// general row length: three passes over the row (the re-reads hit L1 / L2); one block per row
__global__ void __launch_bounds__(256) softmax_rows_split_kernel(__nv_bfloat16* __restrict__ out, const float* __restrict__ x,
                                                                 int64_t cols, float divisor) {
  const float* xr = x + (int64_t)blockIdx.x * cols;
  __nv_bfloat16* o = out + (int64_t)blockIdx.x * 2 * cols;
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
  for (int64_t i = threadIdx.x; i < cols; i += blockDim.x) s += expf(xr[i] / divisor - mx);
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
  for (int64_t i = threadIdx.x; i < cols; i += blockDim.x) {
    __nv_bfloat16 hi, lo;
    split2(expf(xr[i] / divisor - mx) / tot, hi, lo);
    o[i] = hi;
    o[cols + i] = lo;
  }
}

}  // namespace ppst

extern "C" int ppst_corr_prepare_split(void* out, const void* in, int64_t rows, int c, int c_center, int in_dtype,
                                       void* stream) {
  using namespace ppst;
  PPST_REQUIRE(out && in && rows > 0 && c > 0 && c_center >= 0 && c_center <= c, PPST_ERR_ARG,
               "corr_prepare_split: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  PPST_DISPATCH_DTYPE(in_dtype, T, corr_prepare_split_kernel<T><<<ceil_div(rows, 8), 256, 0, st>>>(
                                       (__nv_bfloat16*)out, (const T*)in, rows, c, c_center));
  PPST_LAUNCH_CHECK("corr_prepare_split_kernel");
  return 0;
}

extern "C" int ppst_split_bf16(void* out, const float* in, int64_t batch, int64_t rows, int64_t cols, int transpose,
                               void* stream) {
  using namespace ppst;
  PPST_REQUIRE(out && in && batch > 0 && rows > 0 && cols > 0, PPST_ERR_ARG, "split_bf16: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  if (transpose) {
    PPST_REQUIRE(batch <= 65535 && (cols + 31) / 32 <= 65535, PPST_ERR_ARG, "split_bf16: grid too large");
    dim3 grid(ceil_div(rows, 32), ceil_div(cols, 32), (unsigned)batch);
    split_bf16_t_kernel<<<grid, 256, 0, st>>>((__nv_bfloat16*)out, in, rows, cols);
  } else {
    const int64_t total = batch * rows * cols;
    const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    split_bf16_kernel<<<blocks, 256, 0, st>>>((__nv_bfloat16*)out, in, batch * rows, cols);
  }
  PPST_LAUNCH_CHECK("split_bf16_kernel");
  return 0;
}

extern "C" int ppst_softmax_rows_split(void* out, const float* x, int64_t rows, int64_t cols, float divisor,
                                       void* stream) {
  using namespace ppst;
  PPST_REQUIRE(out && x && rows > 0 && cols > 0 && divisor != 0.0f, PPST_ERR_ARG, "softmax_rows_split: bad args");
  if (cols % 4 == 0 && cols <= 4096 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0)
    softmax_rows_split_reg_kernel<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)out, x, cols, divisor);
  else
    softmax_rows_split_kernel<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)out, x, cols, divisor);
  PPST_LAUNCH_CHECK("softmax_rows_split_kernel");
  return 0;
}

extern "C" int ppst_rselfcorr(void* out, const void* fea, int64_t n, int h, int w, int c, int out_c_stride, int c_off,
                              int in_dtype, int out_dtype, void* stream) {
  using namespace ppst;
  PPST_REQUIRE(out && fea && n > 0 && h > 0 && w > 0 && h % 4 == 0 && w % 4 == 0 && c > 0, PPST_ERR_ARG,
               "rselfcorr: bad args");
  PPST_REQUIRE(c_off >= 0 && c_off + 256 <= out_c_stride, PPST_ERR_ARG, "rselfcorr: channel slice");
  const int64_t patches = n * (h / 4) * (w / 4);
  PPST_REQUIRE(patches < (1LL << 31), PPST_ERR_ARG, "rselfcorr: too many patches");
  const size_t smem = sizeof(float) * 16 * (c + 1);
  PPST_REQUIRE(smem <= 48 * 1024, PPST_ERR_UNSUPPORTED, "rselfcorr: too many channels (%d)", c);
  cudaStream_t st = (cudaStream_t)stream;
  const bool fast = (c == 64) && out_c_stride % 4 == 0 && c_off % 4 == 0 &&
                    ((reinterpret_cast<uintptr_t>(fea) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
#define PPST_RSC(TI, TO)                                                                                              \
  do {                                                                                                                \
    if (fast)                                                                                                         \
      rselfcorr64_kernel<TI, TO><<<(unsigned)((patches + 7) / 8), 128, 0, st>>>((TO*)out, (const TI*)fea, patches, h, \
                                                                                 w, out_c_stride, c_off);             \
    else                                                                                                              \
      rselfcorr_kernel<TI, TO><<<(unsigned)patches, 256, smem, st>>>((TO*)out, (const TI*)fea, h, w, c, out_c_stride, \
                                                                     c_off);                                          \
  } while (0)
  if (in_dtype == PPST_F32 && out_dtype == PPST_F32) PPST_RSC(float, float);
  else if (in_dtype == PPST_BF16 && out_dtype == PPST_BF16) PPST_RSC(__nv_bfloat16, __nv_bfloat16);
  else if (in_dtype == PPST_BF16 && out_dtype == PPST_F32) PPST_RSC(__nv_bfloat16, float);
  else if (in_dtype == PPST_F32 && out_dtype == PPST_BF16) PPST_RSC(float, __nv_bfloat16);
  else PPST_REQUIRE(false, PPST_ERR_DTYPE, "rselfcorr: dtype combination");
#undef PPST_RSC
  PPST_LAUNCH_CHECK("rselfcorr_kernel");
  return 0;
}

extern "C" int ppst_corr_prepare(float* out, const void* in, int64_t rows, int c, int c_center, int in_dtype,
                                 void* stream) {
  using namespace ppst;
  PPST_REQUIRE(out && in && rows > 0 && c > 0 && c_center >= 0 && c_center <= c, PPST_ERR_ARG, "corr_prepare: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  PPST_DISPATCH_DTYPE(in_dtype, T,
                      corr_prepare_kernel<T><<<ceil_div(rows, 8), 256, 0, st>>>(out, (const T*)in, rows, c, c_center));
  PPST_LAUNCH_CHECK("corr_prepare_kernel");
  return 0;
}
