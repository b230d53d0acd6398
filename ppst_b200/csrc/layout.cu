// NCHW <-> NHWC transposes (with dtype conversion and channel-slice placement), tiled through shared memory so
// that both the read and the write side are coalesced.
#include "common.cuh"

namespace ppst {

// src [n, c, hw] -> dst [n, hw, dst_c] at channel offset c_off
template <typename TS, typename TD>
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(TD* __restrict__ dst, const TS* __restrict__ src, int64_t c,
                                                           int64_t hw, int64_t dst_c, int64_t c_off) {
  __shared__ float tile[32][33];
  const int64_t n = blockIdx.z;
  const int64_t p0 = (int64_t)blockIdx.x * 32, c0 = (int64_t)blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;           // 32 x 8
  for (int j = ty; j < 32; j += 8) {
    const int64_t ch = c0 + j, p = p0 + tx;
    tile[j][tx] = (ch < c && p < hw) ? Cvt<TS>::to_f(src[(n * c + ch) * hw + p]) : 0.0f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int64_t p = p0 + j, ch = c0 + tx;
    if (p < hw && ch < c) dst[(n * hw + p) * dst_c + c_off + ch] = Cvt<TD>::from_f(tile[tx][j]);
  }
}

// src [n, hw, src_c] (channel offset c_off) -> dst [n, c, hw]
template <typename TS, typename TD>
__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(TD* __restrict__ dst, const TS* __restrict__ src, int64_t c,
                                                           int64_t hw, int64_t src_c, int64_t c_off,
                                                           const float* __restrict__ a, const float* __restrict__ b) {
  __shared__ float tile[32][33];
  const int64_t n = blockIdx.z;
  const int64_t p0 = (int64_t)blockIdx.x * 32, c0 = (int64_t)blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int j = ty; j < 32; j += 8) {
    const int64_t p = p0 + j, ch = c0 + tx;
    tile[j][tx] = (p < hw && ch < c) ? Cvt<TS>::to_f(src[(n * hw + p) * src_c + c_off + ch]) : 0.0f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int64_t ch = c0 + j, p = p0 + tx;
    if (ch < c && p < hw) {
      float v = tile[tx][j];
      if (a) v = fmaf(v, a[n * c + ch], b ? b[n * c + ch] : 0.0f);       // per-(n, c) affine (ToRGB's IN + StyleMod)
      dst[(n * c + ch) * hw + p] = Cvt<TD>::from_f(v);
    }
  }
}

// c <= 4 dense f32 (the RGB image that leaves the generator): the tiled transpose above keeps 3 of 32 lanes busy on the read
// side.  Here a thread owns 4 consecutive pixels: C 16-byte loads, the per-(n, c) affine, C 16-byte plane stores.
template <int C>
__global__ void __launch_bounds__(256) nhwc_small_to_nchw_kernel(float* __restrict__ dst, const float* __restrict__ src,
                                                                 int64_t hw, const float* __restrict__ a,
                                                                 const float* __restrict__ b) {
  const int64_t n = blockIdx.y;
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;          // pixel quad
  if (q * 4 >= hw) return;
  float v[4 * C];
  const float4* s4 = reinterpret_cast<const float4*>(src + (n * hw + q * 4) * C);
#pragma unroll
  for (int i = 0; i < C; ++i) {
    const float4 t = s4[i];
    v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
  }
#pragma unroll
  for (int ch = 0; ch < C; ++ch) {
    const float av = a ? a[n * C + ch] : 1.0f, bv = (a && b) ? b[n * C + ch] : 0.0f;
    float4 o;
    o.x = a ? fmaf(v[ch], av, bv) : v[ch];
    o.y = a ? fmaf(v[C + ch], av, bv) : v[C + ch];
    o.z = a ? fmaf(v[2 * C + ch], av, bv) : v[2 * C + ch];
    o.w = a ? fmaf(v[3 * C + ch], av, bv) : v[3 * C + ch];
    *reinterpret_cast<float4*>(dst + (n * C + ch) * hw + q * 4) = o;
  }
}

#define PPST_DISPATCH2(sd, dd, TS, TD, ...)                                        \
  PPST_DISPATCH_DTYPE(sd, TS, PPST_DISPATCH_DTYPE(dd, TD, __VA_ARGS__))

}  // namespace ppst

extern "C" int ppst_nchw_to_nhwc(void* dst, const void* src, int64_t n, int64_t c, int64_t hw, int64_t dst_c,
                                 int64_t c_off, int src_dtype, int dst_dtype, void* stream) {
  using namespace ppst;
  PPST_REQUIRE(dst && src && n > 0 && c > 0 && hw > 0 && c_off >= 0 && c_off + c <= dst_c, PPST_ERR_ARG,
               "nchw_to_nhwc: bad args");
  PPST_REQUIRE(n <= 65535 && (c + 31) / 32 <= 65535, PPST_ERR_ARG, "nchw_to_nhwc: batch/channel count too large");
  dim3 grid(ceil_div(hw, 32), ceil_div(c, 32), (unsigned)n);
  cudaStream_t st = (cudaStream_t)stream;
  PPST_DISPATCH2(src_dtype, dst_dtype, TS, TD,
                 nchw_to_nhwc_kernel<TS, TD><<<grid, 256, 0, st>>>((TD*)dst, (const TS*)src, c, hw, dst_c, c_off));
  PPST_LAUNCH_CHECK("nchw_to_nhwc_kernel");
  return 0;
}

extern "C" int ppst_nhwc_to_nchw(void* dst, const void* src, int64_t n, int64_t c, int64_t hw, int64_t src_c,
                                 int64_t c_off, int src_dtype, int dst_dtype, const float* a, const float* b,
                                 void* stream) {
  using namespace ppst;
  PPST_REQUIRE(dst && src && n > 0 && c > 0 && hw > 0 && c_off >= 0 && c_off + c <= src_c, PPST_ERR_ARG,
               "nhwc_to_nchw: bad args");
  PPST_REQUIRE(n <= 65535 && (c + 31) / 32 <= 65535, PPST_ERR_ARG, "nhwc_to_nchw: batch/channel count too large");
  cudaStream_t st = (cudaStream_t)stream;
  if (c == 3 && src_c == 3 && c_off == 0 && hw % 4 == 0 && src_dtype == PPST_F32 && dst_dtype == PPST_F32 &&
      ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15) == 0) {
    const dim3 g3((unsigned)ceil_div(hw / 4, 256), (unsigned)n);
    nhwc_small_to_nchw_kernel<3><<<g3, 256, 0, st>>>((float*)dst, (const float*)src, hw, a, b);
    PPST_LAUNCH_CHECK("nhwc_small_to_nchw_kernel");
    return 0;
  }
  dim3 grid(ceil_div(hw, 32), ceil_div(c, 32), (unsigned)n);
  PPST_DISPATCH2(src_dtype, dst_dtype, TS, TD,
                 nhwc_to_nchw_kernel<TS, TD><<<grid, 256, 0, st>>>((TD*)dst, (const TS*)src, c, hw, src_c, c_off, a, b));
  PPST_LAUNCH_CHECK("nhwc_to_nchw_kernel");
  return 0;
}

// Per-sample input-channel scaling of packed conv weights: out[n][r][t*cin + c] = bf16(w[r][t*cin + c] * a[n][c]).
// InstanceNorm in FRONT of a convolution whose output is instance-normalised again (the generator's feature heads,
// generator.py:180-231: [ReplPad] -> IN -> conv -> IN) needs no pass over the activation: the scale folds into the
// weights per sample -- StyleGAN2's weight modulation -- and the shift, a per-(sample, output channel) constant because
// nothing is zero-padded after it, is removed by the second normalisation.
namespace ppst {
__global__ void __launch_bounds__(256) scale_weights_kernel(__nv_bfloat16* __restrict__ out, const float* __restrict__ w,
                                                            const float* __restrict__ a, int64_t n, int64_t rows,
                                                            int taps, int cin) {
  const int64_t K = (int64_t)taps * cin, per = rows * K, total = n * per;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < total; i += (int64_t)gridDim.x * blockDim.x * 4) {
    const int64_t ni = i / per, r = i - ni * per;
    const int c = (int)(r % cin);                       // cin % 4 == 0: the 4 elements share a tap
    const float4 wv = *reinterpret_cast<const float4*>(w + r);
    const float4 av = *reinterpret_cast<const float4*>(a + ni * cin + c);
    union { uint2 u; __nv_bfloat16 h[4]; } o;
    o.h[0] = __float2bfloat16_rn(wv.x * av.x); o.h[1] = __float2bfloat16_rn(wv.y * av.y);
    o.h[2] = __float2bfloat16_rn(wv.z * av.z); o.h[3] = __float2bfloat16_rn(wv.w * av.w);
    *reinterpret_cast<uint2*>(out + i) = o.u;
  }
}
}  // namespace ppst

extern "C" int ppst_scale_weights(void* out, const float* w, const float* a, int64_t n, int64_t rows, int taps, int cin,
                                  void* stream) {
  using namespace ppst;
  PPST_REQUIRE(out && w && a && n > 0 && rows > 0 && taps > 0 && cin > 0 && cin % 4 == 0, PPST_ERR_ARG,
               "scale_weights: bad args (cin must be a multiple of 4)");
  PPST_REQUIRE(((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(a)) & 15) == 0,
               PPST_ERR_ARG, "scale_weights: buffers must be 16-byte aligned");
  const int64_t total = n * rows * taps * cin;
  const int64_t blocks = (total / 4 + 255) / 256;
  scale_weights_kernel<<<(unsigned)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, (cudaStream_t)stream>>>(
      (__nv_bfloat16*)out, w, a, n, rows, taps, cin);
  PPST_LAUNCH_CHECK("scale_weights_kernel");
  return 0;
}

// Lazy InstanceNorm + StyleMod in FRONT of a zero-padded convolution (StyledConv conv1 -> conv2 of the generator's
// residual blocks, stylegan2_layers.py:439-475).  With x = a*y + b per (sample, channel) and zero padding of x,
//     conv(x, W)[co] = conv(y', W * a)[co] + sum_{t,c} W[co][t][c] * b[c]
// where y' is y with its padding border set to -b/a (so that a*y' + b = 0 there: ppst_fill_border).  This kernel
// writes the per-sample bf16 weights W*a and the per-sample bias  base_bias[co] + sum W*b  -- the normalised
// activation is never materialised.  One block per (sample, output row of W).
namespace ppst {
__global__ void __launch_bounds__(256) modulate_weights_kernel(__nv_bfloat16* __restrict__ out, float* __restrict__ out_bias,
                                                               const float* __restrict__ w, const float* __restrict__ a,
                                                               const float* __restrict__ b,
                                                               const float* __restrict__ base_bias, int64_t rows, int taps,
                                                               int cin) {
  __shared__ float red[8];
  const int64_t r = blockIdx.x, ni = blockIdx.y;
  const int K = taps * cin;
  const float* wr = w + r * K;
  __nv_bfloat16* o = out + (ni * rows + r) * K;
  const float* av = a + ni * cin;
  const float* bv = b + ni * cin;
  float acc = 0.f;
  for (int i = threadIdx.x * 4; i < K; i += 256 * 4) {
    const int c = i % cin;                               // cin % 4 == 0: the 4 elements share a tap
    const float4 wv = *reinterpret_cast<const float4*>(wr + i);
    const float4 a4 = *reinterpret_cast<const float4*>(av + c);
    const float4 b4 = *reinterpret_cast<const float4*>(bv + c);
    union { uint2 u; __nv_bfloat16 h[4]; } pk;
    pk.h[0] = __float2bfloat16_rn(wv.x * a4.x); pk.h[1] = __float2bfloat16_rn(wv.y * a4.y);
    pk.h[2] = __float2bfloat16_rn(wv.z * a4.z); pk.h[3] = __float2bfloat16_rn(wv.w * a4.w);
    *reinterpret_cast<uint2*>(o + i) = pk.u;
    acc = fmaf(wv.x, b4.x, acc); acc = fmaf(wv.y, b4.y, acc);
    acc = fmaf(wv.z, b4.z, acc); acc = fmaf(wv.w, b4.w, acc);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i];
    out_bias[ni * rows + r] = t + (base_bias ? base_bias[r] : 0.f);
  }
}

// Border of a padded NHWC bf16 buffer [n, h + 2p, w + 2p, c] <- -b/a per (sample, channel): the value the lazy affine
// a*y + b maps to zero, i.e. the zero padding of the normalised tensor.  |a| is clamped away from 0.
__global__ void __launch_bounds__(256) fill_border_kernel(__nv_bfloat16* __restrict__ buf, const float* __restrict__ a,
                                                          const float* __restrict__ b, int h, int w, int c, int pad) {
  const int hp = h + 2 * pad, wp = w + 2 * pad;
  const int border = hp * wp - h * w;                    // border pixels per image
  const int cv = c / 8;
  const int64_t ni = blockIdx.y;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < (int64_t)border * cv;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int cg = (int)(t % cv);
    int k = (int)(t / cv);                               // border pixel index: top rows, bottom rows, then side columns
    int y, x;
    if (k < pad * wp) { y = k / wp; x = k - y * wp; }
    else if (k < 2 * pad * wp) { k -= pad * wp; y = h + pad + k / wp; x = k % wp; }
    else {
      k -= 2 * pad * wp;
      y = pad + k / (2 * pad);
      const int s = k % (2 * pad);
      x = s < pad ? s : w + s;
    }
    union { uint4 u; __nv_bfloat16 v[8]; } pk;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float av = a[ni * c + cg * 8 + e];
      const float bv = b[ni * c + cg * 8 + e];
      if (fabsf(av) < 1e-20f) av = av < 0.f ? -1e-20f : 1e-20f;
      pk.v[e] = __float2bfloat16_rn(-bv / av);
    }
    *reinterpret_cast<uint4*>(buf + ((ni * hp + y) * wp + x) * (int64_t)c + cg * 8) = pk.u;
  }
}

// Border of a padded NHWC buffer from its own interior (reflect / replicate / zero): one 16-byte vector per thread.
// Border pixel index k: the pt top rows, the pb bottom rows (full width), then the pl + pr side pixels of the h middle rows.
template <typename T>
__global__ void __launch_bounds__(256) pad_border_kernel(T* __restrict__ buf, int h, int w, int c, int pt, int pb, int pl,
                                                         int pr, int mode) {
  constexpr int V = 16 / sizeof(T);
  const int hp = h + pt + pb, wp = w + pl + pr;
  const int64_t border = (int64_t)(pt + pb) * wp + (int64_t)h * (pl + pr);
  const int cv = c / V;
  const int64_t ni = blockIdx.y;
  T* img = buf + ni * hp * wp * (int64_t)c;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < border * cv; t += (int64_t)gridDim.x * blockDim.x) {
    const int cg = (int)(t % cv);
    int64_t k = t / cv;
    int y, x;
    if (k < (int64_t)pt * wp) { y = (int)(k / wp); x = (int)(k - (int64_t)y * wp); }
    else if (k < (int64_t)(pt + pb) * wp) { k -= (int64_t)pt * wp; y = pt + h + (int)(k / wp); x = (int)(k % wp); }
    else {
      k -= (int64_t)(pt + pb) * wp;
      y = pt + (int)(k / (pl + pr));
      const int sx = (int)(k % (pl + pr));
      x = sx < pl ? sx : w + sx;
    }
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (mode != 0) {
      int sy = y - pt, sx = x - pl;
      if (mode == 1) {                                   // reflect (no edge repeat)
        sy = sy < 0 ? -sy : (sy >= h ? 2 * (h - 1) - sy : sy);
        sx = sx < 0 ? -sx : (sx >= w ? 2 * (w - 1) - sx : sx);
      } else {                                           // replicate
        sy = sy < 0 ? 0 : (sy >= h ? h - 1 : sy);
        sx = sx < 0 ? 0 : (sx >= w ? w - 1 : sx);
      }
      v = *reinterpret_cast<const uint4*>(img + ((int64_t)(sy + pt) * wp + sx + pl) * c + cg * V);
    }
    *reinterpret_cast<uint4*>(img + ((int64_t)y * wp + x) * c + cg * V) = v;
  }
}
}  // namespace ppst

extern "C" int ppst_pad_border(void* buf, int64_t n, int h, int w, int c, int pt, int pb, int pl, int pr, int mode,
                               int dtype, void* stream) {
  using namespace ppst;
  PPST_REQUIRE(buf && n > 0 && n <= 65535 && h > 0 && w > 0 && c > 0 && pt >= 0 && pb >= 0 && pl >= 0 && pr >= 0 &&
               pt + pb + pl + pr > 0, PPST_ERR_ARG, "pad_border: bad args");
  PPST_REQUIRE(mode >= 0 && mode <= 2, PPST_ERR_ARG, "pad_border: mode %d (0 zero, 1 reflect, 2 replicate)", mode);
  PPST_REQUIRE(mode != 1 || (pt < h && pb < h && pl < w && pr < w), PPST_ERR_ARG, "pad_border: reflect pad >= size");
  PPST_REQUIRE((dtype == PPST_BF16 && c % 8 == 0) || (dtype == PPST_F32 && c % 4 == 0), PPST_ERR_UNSUPPORTED,
               "pad_border: bf16 with c %% 8 == 0 or f32 with c %% 4 == 0");
  PPST_REQUIRE((reinterpret_cast<uintptr_t>(buf) & 15) == 0, PPST_ERR_ARG, "pad_border: buffer must be 16-byte aligned");
  const int64_t border = (int64_t)(pt + pb) * (w + pl + pr) + (int64_t)h * (pl + pr);
  const int64_t vecs = border * (c / (dtype == PPST_BF16 ? 8 : 4));
  const int64_t blocks = (vecs + 255) / 256;
  const dim3 grid((unsigned)(blocks < 2048 ? blocks : 2048), (unsigned)n);
  if (dtype == PPST_BF16)
    pad_border_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)buf, h, w, c, pt, pb, pl, pr, mode);
  else
    pad_border_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((float*)buf, h, w, c, pt, pb, pl, pr, mode);
  PPST_LAUNCH_CHECK("pad_border_kernel");
  return 0;
}

extern "C" int ppst_modulate_weights(void* out_w, float* out_bias, const float* w, const float* a, const float* b,
                                     const float* base_bias, int64_t n, int64_t rows, int taps, int cin, void* stream) {
  using namespace ppst;
  PPST_REQUIRE(out_w && out_bias && w && a && b && n > 0 && n <= 65535 && rows > 0 && taps > 0 && cin > 0 && cin % 4 == 0,
               PPST_ERR_ARG, "modulate_weights: bad args (cin must be a multiple of 4)");
  PPST_REQUIRE(((reinterpret_cast<uintptr_t>(out_w) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(a) |
                 reinterpret_cast<uintptr_t>(b)) & 15) == 0, PPST_ERR_ARG, "modulate_weights: buffers must be 16-byte aligned");
  modulate_weights_kernel<<<dim3((unsigned)rows, (unsigned)n), 256, 0, (cudaStream_t)stream>>>(
      (__nv_bfloat16*)out_w, out_bias, w, a, b, base_bias, rows, taps, cin);
  PPST_LAUNCH_CHECK("modulate_weights_kernel");
  return 0;
}

extern "C" int ppst_fill_border(void* buf, const float* a, const float* b, int64_t n, int h, int w, int c, int pad,
                                int dtype, void* stream) {
  using namespace ppst;
  PPST_REQUIRE(buf && a && b && n > 0 && n <= 65535 && h > 0 && w > 0 && c > 0 && pad >= 1 && pad <= 8, PPST_ERR_ARG,
               "fill_border: bad args");
  PPST_REQUIRE(dtype == PPST_BF16 && c % 8 == 0 && (reinterpret_cast<uintptr_t>(buf) & 15) == 0, PPST_ERR_UNSUPPORTED,
               "fill_border: bf16 buffers with c a multiple of 8");
  const int64_t border = (int64_t)(h + 2 * pad) * (w + 2 * pad) - (int64_t)h * w;
  const int64_t blocks = (border * (c / 8) + 255) / 256;
  fill_border_kernel<<<dim3((unsigned)(blocks < 1024 ? blocks : 1024), (unsigned)n), 256, 0, (cudaStream_t)stream>>>(
      (__nv_bfloat16*)buf, a, b, h, w, c, pad);
  PPST_LAUNCH_CHECK("fill_border_kernel");
  return 0;
}
