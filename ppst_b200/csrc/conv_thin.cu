// Thin 1x1 convolutions that are pure HBM streams, not GEMMs:
//   FromRGB  (ConvLayer 3 -> 32, 1x1 + FusedLeakyReLU, stylegan2_layers.py:497-555; encoder_con.py:26, encoder_col.py:27):
//            reads the NCHW fp32 image planes directly (fusing the NCHW->NHWC conversion) and writes NHWC.
//   ToRGB    (EqualConv2d 128 -> 3, 1x1, stylegan2_layers.py:484): a per-pixel dot product over the channel vector.
#include "common.cuh"

namespace ppst {

// out NHWC [n, hw, cout_stride] (slice c_off..c_off+COUT), x NCHW f32 [n, cin, hw].
// A block stages kPixBlock consecutive pixels of each input plane in shared memory (coalesced 16-byte loads when the
// plane length allows), then thread -> (pixel, group of 8 output channels): COUT/8 consecutive lanes write one pixel's
// channel vector, so a warp store covers 8 pixels x 64 contiguous bytes (bf16) instead of 32 scattered pieces.
constexpr int kPixBlock = 1024;
template <typename TO, int COUT, bool BORDER = false>
__global__ void __launch_bounds__(256) conv1x1_in_nchw_kernel(TO* __restrict__ out, const float* __restrict__ x,
                                                              const float* __restrict__ w, const float* __restrict__ bias,
                                                              int64_t n, int cin, int64_t hw, int cout_stride, int c_off,
                                                              int has_act, float slope, float gain, int w_row, int ob) {
  constexpr int G = COUT / 8;                     // channel groups (= lanes) per pixel
  __shared__ __align__(16) float sx[4][kPixBlock];
  static_assert(256 % G == 0, "a thread must keep its channel group");
  const int g = threadIdx.x % G;
  float rw[8][4], rb[8];                          // this thread's 8 output channels: weights (zero beyond cin), bias
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    rb[e] = bias ? bias[g * 8 + e] : 0.0f;
#pragma unroll
    for (int ci = 0; ci < 4; ++ci) rw[e][ci] = ci < cin ? w[(g * 8 + e) * cin + ci] : 0.0f;
  }
  const int64_t blocks_per_img = (hw + kPixBlock - 1) / kPixBlock;
  const bool vec = (hw % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  for (int64_t blk = blockIdx.x; blk < n * blocks_per_img; blk += gridDim.x) {
    const int64_t ni = blk / blocks_per_img, p0 = (blk - ni * blocks_per_img) * kPixBlock;
    const int np = (int)((hw - p0) < kPixBlock ? (hw - p0) : kPixBlock);
    __syncthreads();                              // previous tile consumed
    for (int ci = cin; ci < 4; ++ci)
      for (int i = threadIdx.x; i < kPixBlock; i += 256) sx[ci][i] = 0.0f;
    for (int ci = 0; ci < cin; ++ci) {
      const float* src = x + (ni * cin + ci) * hw + p0;
      if (vec) {
        const int i = threadIdx.x * 4;
        if (i < np) *reinterpret_cast<float4*>(&sx[ci][i]) = *reinterpret_cast<const float4*>(src + i);
      } else {
        for (int i = threadIdx.x; i < np; i += 256) sx[ci][i] = src[i];
      }
    }
    __syncthreads();
    // ob > 0: the output carries a border of ob pixels; pixel p = (y, x) lands at (y + ob, x + ob) of the padded grid
    const int wp = w_row + 2 * ob;
    TO* obase = BORDER ? out + (ni * (hw / w_row + 2 * ob) * wp) * cout_stride + c_off
                       : out + (ni * hw + p0) * cout_stride + c_off;
    // 256 % G == 0: a thread always serves the same channel group, whose weights / bias live in registers
    for (int item = threadIdx.x; item < np * G; item += 256) {
      const int pix = item / G;
      float xin[4];
#pragma unroll
      for (int ci = 0; ci < 4; ++ci) xin[ci] = sx[ci][pix];
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float acc = rb[e];
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) acc = fmaf(xin[ci], rw[e][ci], acc);
        if (has_act) acc = (acc > 0.f) ? acc : acc * slope;
        v[e] = acc * gain;
      }
      int64_t opix = pix;
      if (BORDER) {                                    // hw < 2^31 (host-checked): 32-bit division
        const int pp = (int)p0 + pix;
        const int yy = pp / w_row;
        opix = (int64_t)(yy + ob) * wp + (pp - yy * w_row) + ob;
      }
      TO* o = obase + opix * cout_stride + g * 8;
      if (sizeof(TO) == 2) {
        Vec16<TO> r;
#pragma unroll
        for (int e = 0; e < 8; ++e) r.v[e] = Cvt<TO>::from_f(v[e]);
        st_vec16(o, r);
      } else {
        float* of = reinterpret_cast<float*>(o);
        *reinterpret_cast<float4*>(of) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(of + 4) = make_float4(v[4], v[5], v[6], v[7]);
      }
    }
  }
}

// out f32 [pixels, COUT], x NHWC [pixels, cin] with cin = 8 * L, L lanes per pixel (L divides 32).
template <typename TI, int COUT>
__global__ void __launch_bounds__(256) conv1x1_thin_out_kernel(float* __restrict__ out, const TI* __restrict__ x,
                                                               const float* __restrict__ w, const float* __restrict__ bias,
                                                               int64_t pixels, int cin) {
  const int L = cin / 8;                         // lanes cooperating on one pixel
  const int lane = threadIdx.x & 31;
  const int cg = lane % L, sub = lane / L;
  const int per_warp = 32 / L;
  float wr[COUT][8];
#pragma unroll
  for (int co = 0; co < COUT; ++co)
#pragma unroll
    for (int e = 0; e < 8; ++e) wr[co][e] = w[co * cin + cg * 8 + e];
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  constexpr int U = 4;                           // pixel groups per iteration: U 16-byte loads in flight per lane
  for (int64_t p0 = warp * per_warp * U; p0 < pixels; p0 += nwarps * per_warp * U) {
    Vec16<TI> rv[U][sizeof(TI) == 2 ? 1 : 2];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t p = p0 + u * per_warp + sub;
      if (p < pixels) {
        if (sizeof(TI) == 2) {
          rv[u][0] = ld_vec16_stream(x + p * cin + cg * 8);
        } else {
          rv[u][0] = ld_vec16_stream(x + p * cin + cg * 8);
          rv[u][sizeof(TI) == 2 ? 0 : 1] = ld_vec16_stream(x + p * cin + cg * 8 + 4);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t p = p0 + u * per_warp + sub;
      float acc[COUT];
#pragma unroll
      for (int co = 0; co < COUT; ++co) acc[co] = 0.f;
      if (p < pixels) {
        float xv[8];
        if (sizeof(TI) == 2) {
#pragma unroll
          for (int e = 0; e < 8; ++e) xv[e] = Cvt<TI>::to_f(rv[u][0].v[e]);
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            xv[e] = Cvt<TI>::to_f(rv[u][0].v[e]);
            xv[4 + e] = Cvt<TI>::to_f(rv[u][sizeof(TI) == 2 ? 0 : 1].v[e]);
          }
        }
#pragma unroll
        for (int co = 0; co < COUT; ++co)
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[co] = fmaf(xv[e], wr[co][e], acc[co]);
      }
      for (int o = L >> 1; o > 0; o >>= 1)
#pragma unroll
        for (int co = 0; co < COUT; ++co) acc[co] += __shfl_xor_sync(0xffffffffu, acc[co], o);
      if (cg == 0 && p < pixels) {
#pragma unroll
        for (int co = 0; co < COUT; ++co) out[p * COUT + co] = acc[co] + (bias ? bias[co] : 0.f);
      }
    }
  }
}

// Weight gradient of the thin input convolution (training): dw[co][ci] = sum_{n,pix} g[n,pix,co] * x[n,ci,pix]
// with g NHWC (bf16 / f32) and x the NCHW f32 image.  Stage 1: one block per (sample, chunk of kWgChunk pixels); thread =
// (group of 8 output channels, pixel lane); block partials are combined through shared memory in a fixed order.
constexpr int kWgChunk = 2048;
template <typename TG>
__global__ void __launch_bounds__(256) thin_wgrad_kernel(float* __restrict__ partials, const TG* __restrict__ g,
                                                         const float* __restrict__ x, int cin, int64_t hw, int cout) {
  __shared__ float s[8192];                          // [pixel lane][cout][4]
  const int V = cout / 8, PL = 256 / V;
  const int v = threadIdx.x % V, pl = threadIdx.x / V;
  const int64_t chunks = (hw + kWgChunk - 1) / kWgChunk;
  const int64_t ni = blockIdx.x / chunks, p0 = (blockIdx.x - ni * chunks) * kWgChunk;
  const int np = (int)((hw - p0) < kWgChunk ? (hw - p0) : kWgChunk);
  float acc[8][4];
#pragma unroll
  for (int e = 0; e < 8; ++e)
#pragma unroll
    for (int ci = 0; ci < 4; ++ci) acc[e][ci] = 0.0f;
  const TG* gp = g + (ni * hw + p0) * cout + v * 8;
  const float* xp = x + ni * cin * hw + p0;
  for (int pix = pl; pix < np; pix += PL) {
    float gv[8];
    if (sizeof(TG) == 2) {
      const Vec16<TG> r = ld_vec16_stream(gp + (int64_t)pix * cout);
#pragma unroll
      for (int e = 0; e < 8; ++e) gv[e] = Cvt<TG>::to_f(r.v[e]);
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) gv[e] = Cvt<TG>::to_f(gp[(int64_t)pix * cout + e]);
    }
    float xv[4];
#pragma unroll
    for (int ci = 0; ci < 4; ++ci) xv[ci] = ci < cin ? xp[ci * hw + pix] : 0.0f;
#pragma unroll
    for (int e = 0; e < 8; ++e)
#pragma unroll
      for (int ci = 0; ci < 4; ++ci) acc[e][ci] = fmaf(gv[e], xv[ci], acc[e][ci]);
  }
#pragma unroll
  for (int e = 0; e < 8; ++e)
    *reinterpret_cast<float4*>(&s[((pl * cout) + v * 8 + e) * 4]) = make_float4(acc[e][0], acc[e][1], acc[e][2], acc[e][3]);
  __syncthreads();
  for (int i = threadIdx.x; i < cout * 4; i += 256) {
    float t = 0.0f;
    for (int l = 0; l < PL; ++l) t += s[l * cout * 4 + i];
    partials[(int64_t)blockIdx.x * cout * 4 + i] = t;
  }
}

__global__ void thin_wgrad_final(float* __restrict__ dw, const float* __restrict__ partials, int64_t J, int cout, int cin) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cout * cin) return;
  const int co = i / cin, ci = i - co * cin;
  float t = 0.0f;
  for (int64_t j = 0; j < J; ++j) t += partials[(j * cout + co) * 4 + ci];
  dw[i] = t;
}

}  // namespace ppst

extern "C" int64_t ppst_conv1x1_thin_wgrad_scratch_elems(int64_t n, int64_t hw, int cout) {
  return n * ((hw + ppst::kWgChunk - 1) / ppst::kWgChunk) * cout * 4;
}

extern "C" int ppst_conv1x1_thin_wgrad(float* dw, const void* g, const float* x, int64_t n, int cin, int64_t hw, int cout,
                                       int g_dtype, float* scratch, void* stream) {
  using namespace ppst;
  PPST_REQUIRE(dw && g && x && scratch && n > 0 && hw > 0 && cin >= 1 && cin <= 4, PPST_ERR_ARG, "conv1x1_thin_wgrad: bad args");
  PPST_REQUIRE(cout % 8 == 0 && cout >= 8 && cout <= 256 && 256 % (cout / 8) == 0, PPST_ERR_UNSUPPORTED,
               "conv1x1_thin_wgrad: cout=%d (8, 16, 32, 64, 128 or 256)", cout);
  const int64_t J = n * ((hw + kWgChunk - 1) / kWgChunk);
  PPST_REQUIRE(J < (1LL << 31), PPST_ERR_ARG, "conv1x1_thin_wgrad: too many blocks");
  cudaStream_t st = (cudaStream_t)stream;
  if (g_dtype == PPST_BF16)
    thin_wgrad_kernel<__nv_bfloat16><<<(unsigned)J, 256, 0, st>>>(scratch, (const __nv_bfloat16*)g, x, cin, hw, cout);
  else if (g_dtype == PPST_F32)
    thin_wgrad_kernel<float><<<(unsigned)J, 256, 0, st>>>(scratch, (const float*)g, x, cin, hw, cout);
  else
    PPST_REQUIRE(false, PPST_ERR_DTYPE, "conv1x1_thin_wgrad: g dtype %d", g_dtype);
  PPST_LAUNCH_CHECK("thin_wgrad_kernel");
  thin_wgrad_final<<<(cout * cin + 127) / 128, 128, 0, st>>>(dw, scratch, J, cout, cin);
  PPST_LAUNCH_CHECK("thin_wgrad_final");
  return 0;
}

extern "C" int ppst_conv1x1_in_nchw(void* out, const float* x, const float* w, const float* bias, int64_t n, int cin,
                                    int64_t hw, int cout, int cout_stride, int c_off, int has_act, float slope,
                                    float gain, int out_dtype, int w_row, int out_border, void* stream) {
  using namespace ppst;
  PPST_REQUIRE(out && x && w && n > 0 && hw > 0 && cin >= 1 && cin <= 4, PPST_ERR_ARG, "conv1x1_in_nchw: bad args");
  PPST_REQUIRE(out_border == 0 || (out_border > 0 && out_border <= 8 && w_row > 0 && hw % w_row == 0), PPST_ERR_ARG,
               "conv1x1_in_nchw: out_border needs the row width");
  PPST_REQUIRE(out_border == 0 || hw < (1LL << 31), PPST_ERR_ARG, "conv1x1_in_nchw: image too large for a bordered output");
  PPST_REQUIRE(cout == 32 || cout == 64 || cout == 128, PPST_ERR_UNSUPPORTED,
               "conv1x1_in_nchw: cout=%d (32: FromRGB of the encoders; 64 / 128: the discriminator's first layer)", cout);
  PPST_REQUIRE(c_off % 8 == 0 && cout_stride % 8 == 0 && c_off + cout <= cout_stride, PPST_ERR_ARG,
               "conv1x1_in_nchw: channel slice");
  const int64_t tiles = n * ((hw + kPixBlock - 1) / kPixBlock);
  const int blocks = (int)(tiles < 148 * 8 ? tiles : 148 * 8);
  cudaStream_t st = (cudaStream_t)stream;
#define PPST_IN_NCHW(TO, CO)                                                                                             \
  do {                                                                                                                   \
    if (out_border)                                                                                                      \
      conv1x1_in_nchw_kernel<TO, CO, true><<<blocks, 256, 0, st>>>((TO*)out, x, w, bias, n, cin, hw, cout_stride, c_off, \
                                                                   has_act, slope, gain, w_row, out_border);             \
    else                                                                                                                 \
      conv1x1_in_nchw_kernel<TO, CO, false><<<blocks, 256, 0, st>>>((TO*)out, x, w, bias, n, cin, hw, cout_stride,       \
                                                                    c_off, has_act, slope, gain, w_row, out_border);     \
  } while (0)
  PPST_REQUIRE(out_dtype == PPST_BF16 || out_dtype == PPST_F32, PPST_ERR_DTYPE, "conv1x1_in_nchw: out dtype %d", out_dtype);
  if (out_dtype == PPST_BF16) {
    if (cout == 32) PPST_IN_NCHW(__nv_bfloat16, 32);
    else if (cout == 64) PPST_IN_NCHW(__nv_bfloat16, 64);
    else PPST_IN_NCHW(__nv_bfloat16, 128);
  } else {
    if (cout == 32) PPST_IN_NCHW(float, 32);
    else if (cout == 64) PPST_IN_NCHW(float, 64);
    else PPST_IN_NCHW(float, 128);
  }
#undef PPST_IN_NCHW
  PPST_LAUNCH_CHECK("conv1x1_in_nchw_kernel");
  return 0;
}

extern "C" int ppst_conv1x1_thin_out(float* out, const void* x, const float* w, const float* bias, int64_t pixels,
                                     int cin, int cout, int in_dtype, void* stream) {
  using namespace ppst;
  PPST_REQUIRE(out && x && w && pixels > 0, PPST_ERR_ARG, "conv1x1_thin_out: bad args");
  PPST_REQUIRE(cout == 3, PPST_ERR_UNSUPPORTED, "conv1x1_thin_out: cout=%d (3 supported: ToRGB)", cout);
  const int L = cin / 8;
  PPST_REQUIRE(cin % 8 == 0 && L >= 1 && L <= 32 && (32 % L) == 0, PPST_ERR_UNSUPPORTED,
               "conv1x1_thin_out: cin=%d must be 8*L with L dividing 32", cin);
  const int per_warp = 4 * (32 / L);             // pixels per warp iteration (4 groups in flight)
  const int64_t warps = (pixels + per_warp - 1) / per_warp;
  const int blocks = (int)((warps + 7) / 8 < 148 * 8 ? (warps + 7) / 8 : 148 * 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (in_dtype == PPST_BF16)
    conv1x1_thin_out_kernel<__nv_bfloat16, 3><<<blocks, 256, 0, st>>>(out, (const __nv_bfloat16*)x, w, bias, pixels, cin);
  else if (in_dtype == PPST_F32)
    conv1x1_thin_out_kernel<float, 3><<<blocks, 256, 0, st>>>(out, (const float*)x, w, bias, pixels, cin);
  else
    PPST_REQUIRE(false, PPST_ERR_DTYPE, "conv1x1_thin_out: in dtype %d", in_dtype);
  PPST_LAUNCH_CHECK("conv1x1_thin_out_kernel");
  return 0;
}
