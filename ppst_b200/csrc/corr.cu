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

}  // namespace ppst

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
#define PPST_RSC(TI, TO) \
  rselfcorr_kernel<TI, TO><<<(unsigned)patches, 256, smem, st>>>((TO*)out, (const TI*)fea, h, w, c, out_c_stride, c_off)
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
