// CUDA-core implicit-GEMM convolution on NHWC activations (fp32 accumulate).
// The exact-fp32 mode of the engine and the thin layers (FromRGB 3->32, ToRGB 128->3) run here; the bf16 hot
// layers run on the tcgen05 kernel in conv_tc.cu with the same descriptor and weight packing.
//
// GEMM view: M = n*out_h*out_w output pixels of one phase, N = cout, K = kh*kw*cin (tap-major, channel-minor).
// Block tile 64 x 64 x 16, 256 threads, 4 x 4 register tile.
#include "common.cuh"

namespace ppst {

constexpr int CBM = 64, CBN = 64, CBK = 16;

template <typename TI, typename TO, bool VEC>
__global__ void __launch_bounds__(256) conv_simt_kernel(TO* __restrict__ out, const TI* __restrict__ in,
                                                        const float* __restrict__ weight,
                                                        const float* __restrict__ bias,
                                                        const float* __restrict__ noise, ppst_conv_desc d) {
  __shared__ __align__(16) float As[CBK][CBM + 4];
  __shared__ __align__(16) float Bs[CBK][CBN + 4];
  const int tid = threadIdx.x;
  const int phase = blockIdx.z;
  const int py = phase / d.phases_x, px = phase % d.phases_x;
  const int off_y = d.off_y + py * d.phase_off_dy, off_x = d.off_x + px * d.phase_off_dx;
  const int64_t K = (int64_t)d.kh * d.kw * d.cin;
  const int64_t M = (int64_t)d.n * d.out_h * d.out_w;
  const int64_t row_pitch = d.in_row_pitch ? d.in_row_pitch : (int64_t)d.in_w * d.cin;
  const int64_t img_pitch = d.in_img_pitch ? d.in_img_pitch : (int64_t)d.in_h * d.in_w * d.cin;
  const float* wp = weight + (int64_t)phase * d.cout * K;
  const int64_t m0 = (int64_t)blockIdx.x * CBM;
  const int n0 = blockIdx.y * CBN;

  // loader mapping: A: pixel row tid/4, k quad (tid%4)*4 ; B: cout row tid/4, k quad (tid%4)*4
  const int l_row = tid >> 2, l_k = (tid & 3) * 4;
  const int64_t am = m0 + l_row;
  int a_n = 0, a_oy = 0, a_ox = 0;
  const bool a_valid = am < M;
  if (a_valid) {
    a_ox = (int)(am % d.out_w);
    a_oy = (int)((am / d.out_w) % d.out_h);
    a_n = (int)(am / ((int64_t)d.out_w * d.out_h));
  }
  const int b_co = n0 + l_row;

  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int64_t k0 = 0; k0 < K; k0 += CBK) {
    float ra[4], rb[4];
    if (VEC) {
      // cin % 16 == 0: the whole chunk lies in one tap
      const int tap = (int)(k0 / d.cin);
      const int ci = (int)(k0 % d.cin) + l_k;
      const int iy = a_oy * d.stride + off_y + tap / d.kw, ix = a_ox * d.stride + off_x + tap % d.kw;
      if (a_valid && iy >= 0 && iy < d.in_h && ix >= 0 && ix < d.in_w) {
        const TI* p = in + (int64_t)a_n * img_pitch + (int64_t)iy * row_pitch + (int64_t)ix * d.cin + ci;
        if (sizeof(TI) == 4) {
          const float4 v = *reinterpret_cast<const float4*>(p);
          ra[0] = v.x; ra[1] = v.y; ra[2] = v.z; ra[3] = v.w;
        } else {
          union { uint2 u; TI t[4]; } pk;
          pk.u = *reinterpret_cast<const uint2*>(p);
#pragma unroll
          for (int e = 0; e < 4; ++e) ra[e] = Cvt<TI>::to_f(pk.t[e]);
        }
      } else {
        ra[0] = ra[1] = ra[2] = ra[3] = 0.f;
      }
      if (b_co < d.cout) {
        const float4 v = *reinterpret_cast<const float4*>(wp + (int64_t)b_co * K + k0 + l_k);
        rb[0] = v.x; rb[1] = v.y; rb[2] = v.z; rb[3] = v.w;
      } else {
        rb[0] = rb[1] = rb[2] = rb[3] = 0.f;
      }
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int64_t k = k0 + l_k + e;
        float av = 0.f, bv = 0.f;
        if (k < K) {
          const int tap = (int)(k / d.cin), ci = (int)(k % d.cin);
          const int iy = a_oy * d.stride + off_y + tap / d.kw, ix = a_ox * d.stride + off_x + tap % d.kw;
          if (a_valid && iy >= 0 && iy < d.in_h && ix >= 0 && ix < d.in_w)
            av = Cvt<TI>::to_f(in[(int64_t)a_n * img_pitch + (int64_t)iy * row_pitch + (int64_t)ix * d.cin + ci]);
          if (b_co < d.cout) bv = wp[(int64_t)b_co * K + k];
        }
        ra[e] = av;
        rb[e] = bv;
      }
    }
    __syncthreads();                      // previous chunk fully consumed
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      As[l_k + e][l_row] = ra[e];
      Bs[l_k + e][l_row] = rb[e];
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < CBK; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
  }

  // epilogue: + bias + noise_w * noise -> (lrelu) -> * gain -> store at the phase's position
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
    const int ox = (int)(m % d.out_w), oy = (int)((m / d.out_w) % d.out_h);
    const int ni = (int)(m / ((int64_t)d.out_w * d.out_h));
    const int fy = oy * d.phases_y + py, fx = ox * d.phases_x + px;
    const int64_t pix = ((int64_t)ni * d.full_h + fy) * d.full_w + fx;
    const float nz = noise ? d.noise_w * noise[pix] : 0.f;
    TO* o = out + pix * d.cout_stride + d.c_off;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = n0 + tx * 4 + j;
      if (co >= d.cout) continue;
      float v = acc[i][j] + (bias ? bias[co] : 0.f) + nz;
      if (d.has_act) v = (v > 0.f) ? v : v * d.slope;
      o[co] = Cvt<TO>::from_f(v * d.gain);
    }
  }
}

int validate_conv_desc(const ppst_conv_desc& d, const char* who) {
  PPST_REQUIRE(d.n > 0 && d.in_h > 0 && d.in_w > 0 && d.cin > 0 && d.out_h > 0 && d.out_w > 0 && d.cout > 0,
               PPST_ERR_ARG, "%s: bad geometry", who);
  PPST_REQUIRE(d.kh > 0 && d.kw > 0 && d.stride >= 1 && d.phases_y >= 1 && d.phases_x >= 1, PPST_ERR_ARG,
               "%s: bad taps/stride/phases", who);
  PPST_REQUIRE(d.full_h >= (d.out_h - 1) * d.phases_y + d.phases_y && d.full_w >= (d.out_w - 1) * d.phases_x + d.phases_x,
               PPST_ERR_ARG, "%s: output tensor %dx%d too small for %dx%d x phases", who, d.full_h, d.full_w, d.out_h,
               d.out_w);
  PPST_REQUIRE(d.c_off >= 0 && d.c_off + d.cout <= d.cout_stride, PPST_ERR_ARG, "%s: channel slice", who);
  return 0;
}

}  // namespace ppst

extern "C" int ppst_conv_simt(void* out, const void* in, const float* weight, const float* bias, const float* noise,
                              const ppst_conv_desc* dp, void* stream) {
  using namespace ppst;
  PPST_REQUIRE(out && in && weight && dp, PPST_ERR_ARG, "conv_simt: null argument");
  const ppst_conv_desc d = *dp;
  int rc = validate_conv_desc(d, "conv_simt");
  if (rc) return rc;
  PPST_REQUIRE(d.out_pad == 0 && d.bias_img_stride == 0, PPST_ERR_UNSUPPORTED,
               "conv_simt: out_pad / per-sample bias are ppst_conv_tc features");
  const int64_t M = (int64_t)d.n * d.out_h * d.out_w;
  dim3 grid(ceil_div(M, CBM), ceil_div(d.cout, CBN), d.phases_y * d.phases_x);
  PPST_REQUIRE(grid.y <= 65535 && grid.z <= 65535, PPST_ERR_ARG, "conv_simt: grid too large");
  const bool vec = (d.cin % 16 == 0) && (d.in_row_pitch % 4 == 0) && (d.in_img_pitch % 4 == 0) && ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(weight)) & 15) == 0;
  cudaStream_t st = (cudaStream_t)stream;
#define PPST_CONV(TI, TO)                                                                                         \
  do {                                                                                                            \
    if (vec)                                                                                                      \
      conv_simt_kernel<TI, TO, true><<<grid, 256, 0, st>>>((TO*)out, (const TI*)in, weight, bias, noise, d);      \
    else                                                                                                          \
      conv_simt_kernel<TI, TO, false><<<grid, 256, 0, st>>>((TO*)out, (const TI*)in, weight, bias, noise, d);     \
  } while (0)
  if (d.in_dtype == PPST_F32 && d.out_dtype == PPST_F32) PPST_CONV(float, float);
  else if (d.in_dtype == PPST_BF16 && d.out_dtype == PPST_BF16) PPST_CONV(__nv_bfloat16, __nv_bfloat16);
  else if (d.in_dtype == PPST_BF16 && d.out_dtype == PPST_F32) PPST_CONV(__nv_bfloat16, float);
  else if (d.in_dtype == PPST_F32 && d.out_dtype == PPST_BF16) PPST_CONV(float, __nv_bfloat16);
  else PPST_REQUIRE(false, PPST_ERR_DTYPE, "conv_simt: dtype combination %d -> %d", d.in_dtype, d.out_dtype);
#undef PPST_CONV
  PPST_LAUNCH_CHECK("conv_simt_kernel");
  return 0;
}
