// fused bias + activation (+ gain), forward / first-order / second-order variants.
// Replaces fused_bias_act_kernel.cu:19-49 of the reference.  HBM-bound: one read (+ optional ref read)
// and one write per element; 128-bit accesses; the channel index is resolved once per block row instead
// of an integer div+mod per element.
#include "common.cuh"

namespace ppst {

template <int ACT, int GRAD>
__device__ __forceinline__ float bias_act_fn(float x, float ref, float alpha, float scale) {
  float y;
  if (GRAD == 2) {
    y = 0.0f;
  } else if (ACT == PPST_ACT_LRELU) {
    const float gate = (GRAD == 0) ? x : ref;
    y = (gate > 0.0f) ? x : x * alpha;
  } else {
    y = x;
  }
  return y * scale;
}

// Plane kernel: the tensor is viewed as [planes, step_b]; blockIdx.y = plane (-> one bias value per block),
// blockIdx.x tiles the plane.  VEC = elements per 16-byte access when the plane length / base pointers allow
// it, else 1.
template <typename T, int ACT, int GRAD, bool VECTOR>
__global__ void __launch_bounds__(256) bias_act_plane_kernel(T* __restrict__ y, const T* __restrict__ x,
                                                             const T* __restrict__ b, const T* __restrict__ ref,
                                                             int64_t step_b, int64_t size_b, float alpha,
                                                             float scale) {
  constexpr int VEC = Vec16<T>::N;
  constexpr int UNROLL = 4;
  const int64_t plane = blockIdx.x;   // grid.x = planes (up to 2^31-1), grid.y = tiles inside the plane
  const float bias = (b != nullptr) ? Cvt<T>::to_f(b[plane % size_b]) : 0.0f;
  const int64_t base = plane * step_b;
  if (VECTOR) {
    const int64_t nvec = step_b / VEC;
    int64_t i0 = ((int64_t)blockIdx.y * blockDim.x * UNROLL + threadIdx.x);
    Vec16<T> vx[UNROLL], vr[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int64_t i = i0 + (int64_t)u * blockDim.x;
      if (i < nvec) {
        vx[u] = ld_vec16_stream(x + base + i * VEC);
        if (GRAD == 1 && ACT == PPST_ACT_LRELU) vr[u] = ld_vec16_stream(ref + base + i * VEC);
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int64_t i = i0 + (int64_t)u * blockDim.x;
      if (i < nvec) {
        Vec16<T> o;
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          const float r = (GRAD == 1 && ACT == PPST_ACT_LRELU) ? Cvt<T>::to_f(vr[u].v[e]) : 0.0f;
          o.v[e] = Cvt<T>::from_f(bias_act_fn<ACT, GRAD>(Cvt<T>::to_f(vx[u].v[e]) + bias, r, alpha, scale));
        }
        st_vec16(y + base + i * VEC, o);
      }
    }
  } else {
    for (int u = 0; u < UNROLL * VEC; ++u) {
      const int64_t i = ((int64_t)blockIdx.y * UNROLL * VEC + u) * blockDim.x + threadIdx.x;
      if (i < step_b) {
        const float r = (GRAD == 1 && ACT == PPST_ACT_LRELU) ? Cvt<T>::to_f(ref[base + i]) : 0.0f;
        y[base + i] = Cvt<T>::from_f(bias_act_fn<ACT, GRAD>(Cvt<T>::to_f(x[base + i]) + bias, r, alpha, scale));
      }
    }
  }
}

// Flat kernel for short planes ([B, C] activations of EqualLinear, step_b small): div/mod per element.
template <typename T, int ACT, int GRAD>
__global__ void __launch_bounds__(256) bias_act_flat_kernel(T* __restrict__ y, const T* __restrict__ x,
                                                            const T* __restrict__ b, const T* __restrict__ ref,
                                                            int64_t n, int64_t step_b, int64_t size_b, float alpha,
                                                            float scale) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float bias = (b != nullptr) ? Cvt<T>::to_f(b[(i / step_b) % size_b]) : 0.0f;
    const float r = (GRAD == 1 && ACT == PPST_ACT_LRELU) ? Cvt<T>::to_f(ref[i]) : 0.0f;
    y[i] = Cvt<T>::from_f(bias_act_fn<ACT, GRAD>(Cvt<T>::to_f(x[i]) + bias, r, alpha, scale));
  }
}

template <typename T, int ACT, int GRAD>
static int launch_bias_act(T* y, const T* x, const T* b, const T* ref, int64_t n, int64_t step_b, int64_t size_b,
                           float alpha, float scale, cudaStream_t st) {
  constexpr int VEC = Vec16<T>::N;
  constexpr int UNROLL = 4;
  if (b == nullptr) {  // no bias: the whole tensor is one plane
    step_b = n;
    size_b = 1;
  }
  const int64_t planes = n / step_b;
  const int64_t per_block = 256LL * UNROLL * VEC;
  const int64_t tiles = (step_b + per_block - 1) / per_block;
  if (step_b >= 1024 && planes < (1LL << 31) && tiles <= 65535) {
    const bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) |
                           (ref ? reinterpret_cast<uintptr_t>(ref) : 0)) & 15) == 0;
    const bool vec = aligned && (step_b % VEC == 0);
    dim3 grid((unsigned)planes, (unsigned)tiles, 1);
    if (vec)
      bias_act_plane_kernel<T, ACT, GRAD, true><<<grid, 256, 0, st>>>(y, x, b, ref, step_b, size_b, alpha, scale);
    else
      bias_act_plane_kernel<T, ACT, GRAD, false><<<grid, 256, 0, st>>>(y, x, b, ref, step_b, size_b, alpha, scale);
    PPST_LAUNCH_CHECK("bias_act_plane_kernel");
    return 0;
  }
  const int blocks = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
  bias_act_flat_kernel<T, ACT, GRAD><<<blocks > 0 ? blocks : 1, 256, 0, st>>>(y, x, b, ref, n, step_b,
                                                                            size_b > 0 ? size_b : 1, alpha, scale);
  PPST_LAUNCH_CHECK("bias_act_flat_kernel");
  return 0;
}

// grad_bias[c] = sum_{o, s} g[o, c, s]; one block per channel, fp32 accumulation, fixed order.
template <typename T>
__global__ void __launch_bounds__(256) bias_grad_kernel(float* __restrict__ out, const T* __restrict__ g,
                                                        int64_t outer, int64_t size_b, int64_t step_b) {
  const int64_t c = blockIdx.x;
  float acc = 0.0f;
  for (int64_t o = 0; o < outer; ++o) {
    const T* p = g + (o * size_b + c) * step_b;
    for (int64_t i = threadIdx.x; i < step_b; i += blockDim.x) acc += Cvt<T>::to_f(p[i]);
  }
  __shared__ float sm[8];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.0f;
    for (int i = 0; i < 8; ++i) t += sm[i];
    out[c] = t;
  }
}

}  // namespace ppst

extern "C" int ppst_fused_bias_act(void* y, const void* x, const void* b, const void* ref, int64_t n,
                                   int64_t step_b, int64_t size_b, int act, int grad, float alpha, float scale,
                                   int dtype, void* stream) {
  using namespace ppst;
  if (n == 0) return 0;
  PPST_REQUIRE(y && x && n > 0, PPST_ERR_ARG, "fused_bias_act: null buffer or negative size");
  PPST_REQUIRE(act == PPST_ACT_LINEAR || act == PPST_ACT_LRELU, PPST_ERR_UNSUPPORTED,
               "fused_bias_act: act=%d not supported (1 linear, 3 lrelu)", act);
  PPST_REQUIRE(grad >= 0 && grad <= 2, PPST_ERR_ARG, "fused_bias_act: grad=%d", grad);
  if (size_b <= 0) b = nullptr;
  if (b) PPST_REQUIRE(step_b > 0 && n % step_b == 0, PPST_ERR_ARG, "fused_bias_act: step_b=%lld", (long long)step_b);
  PPST_REQUIRE(!(act == PPST_ACT_LRELU && grad == 1 && ref == nullptr), PPST_ERR_ARG,
               "fused_bias_act: grad=1 needs the saved output as `ref`");
  cudaStream_t st = (cudaStream_t)stream;
#define PPST_BA(ACT_, GRAD_)                                                                              \
  PPST_DISPATCH_DTYPE(dtype, T, return launch_bias_act<T, ACT_, GRAD_>((T*)y, (const T*)x, (const T*)b,   \
                                                                        (const T*)ref, n, step_b, size_b, \
                                                                        alpha, scale, st))
  if (act == PPST_ACT_LRELU) {
    if (grad == 0) { PPST_BA(PPST_ACT_LRELU, 0); }
    if (grad == 1) { PPST_BA(PPST_ACT_LRELU, 1); }
    PPST_BA(PPST_ACT_LRELU, 2);
  } else {
    if (grad == 2) { PPST_BA(PPST_ACT_LINEAR, 2); }
    PPST_BA(PPST_ACT_LINEAR, 0);
  }
#undef PPST_BA
  return 0;
}

extern "C" int ppst_bias_grad_reduce(float* out, const void* g, int64_t outer, int64_t size_b, int64_t step_b,
                                     int dtype, void* stream) {
  using namespace ppst;
  PPST_REQUIRE(out && g && outer > 0 && size_b > 0 && step_b > 0, PPST_ERR_ARG, "bias_grad_reduce: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  PPST_DISPATCH_DTYPE(dtype, T, bias_grad_kernel<T><<<(unsigned)size_b, 256, 0, st>>>(out, (const T*)g, outer,
                                                                                     size_b, step_b));
  PPST_LAUNCH_CHECK("bias_grad_kernel");
  return 0;
}


// ---------------------------------------------------------------------------------------------------------------------
// Backward of bias + leaky-ReLU on NHWC activations in ONE pass (training path): gp = g * (y > 0 ? 1 : alpha) * scale
// (FusedLeakyReLUFunctionBackward, fused_act.py:23-43; act = 0: gp = g, nothing is written) together with the bias gradient
// sum_[pixels] gp per channel (`grad_input.sum(dim)`, :37-42) -- the reference reads the tensor twice for these.
// Stage 1: a block owns kAgChunk pixels; thread = (16-byte channel vector, pixel lane); block partials in a fixed order.
namespace ppst {
constexpr int kAgChunk = 512;
template <typename T>
__global__ void __launch_bounds__(256) act_grad_sum_kernel(T* __restrict__ gp, float* __restrict__ partials,
                                                           const T* __restrict__ g, const T* __restrict__ y, int64_t pixels,
                                                           int c, float alpha, float scale, int act) {
  extern __shared__ float red[];                         // [lanes][c]
  constexpr int VEC = Vec16<T>::N;
  const int cols = c / VEC, lanes = 256 / cols;
  const int col = threadIdx.x % cols, lane = threadIdx.x / cols;
  const int64_t beg = (int64_t)blockIdx.x * kAgChunk, end = min(beg + (int64_t)kAgChunk, pixels);
  float acc[VEC];
#pragma unroll
  for (int e = 0; e < VEC; ++e) acc[e] = 0.f;
  if (lane < lanes) {
    // four pixels per iteration: every load is issued before the first use
    for (int64_t p0 = beg + lane; p0 < end; p0 += 4 * (int64_t)lanes) {
      Vec16<T> gv[4], yv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t p = p0 + u * (int64_t)lanes;
        if (p < end) {
          const int64_t off = p * c + (int64_t)col * VEC;
          gv[u] = ld_vec16_stream(g + off);
          if (act) yv[u] = ld_vec16_stream(y + off);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t p = p0 + u * (int64_t)lanes;
        if (p >= end) break;
        const int64_t off = p * c + (int64_t)col * VEC;
        if (act) {
          Vec16<T> o;
#pragma unroll
          for (int e = 0; e < VEC; ++e) {
            const float v = Cvt<T>::to_f(gv[u].v[e]);
            o.v[e] = Cvt<T>::from_f(((Cvt<T>::to_f(yv[u].v[e]) > 0.0f) ? v : v * alpha) * scale);
            acc[e] += Cvt<T>::to_f(o.v[e]);               // the sum of the STORED (rounded) values
          }
          st_vec16(gp + off, o);
        } else {
#pragma unroll
          for (int e = 0; e < VEC; ++e) acc[e] += Cvt<T>::to_f(gv[u].v[e]);
        }
      }
    }
#pragma unroll
    for (int e = 0; e < VEC; ++e) red[(int64_t)lane * c + col * VEC + e] = acc[e];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < c; i += 256) {
    float t = 0.f;
    for (int l = 0; l < lanes; ++l) t += red[(int64_t)l * c + i];
    partials[(int64_t)blockIdx.x * c + i] = t;
  }
}

__global__ void act_grad_sum_final(float* __restrict__ out, const float* __restrict__ partials, int64_t blocks, int c) {
  // one warp-sized group of threads per channel would be faster; the partial count is a few thousand at most
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  float t = 0.f;
  for (int64_t b = 0; b < blocks; ++b) t += partials[b * c + ch];
  out[ch] = t;
}
}  // namespace ppst

extern "C" int64_t ppst_act_grad_sum_scratch_elems(int64_t pixels, int c) {
  return ((pixels + ppst::kAgChunk - 1) / ppst::kAgChunk) * c;
}

extern "C" int ppst_act_grad_sum(void* gp, float* bias_grad, const void* g, const void* y, float* scratch, int64_t pixels, int c,
                                 float alpha, float scale, int act, int dtype, void* stream) {
  using namespace ppst;
  PPST_REQUIRE(bias_grad && g && scratch && pixels > 0 && c > 0 && (!act || (gp && y)), PPST_ERR_ARG, "act_grad_sum: bad args");
  PPST_REQUIRE(dtype == PPST_BF16 || dtype == PPST_F32, PPST_ERR_DTYPE, "act_grad_sum: dtype %d", dtype);
  const int vec = dtype == PPST_BF16 ? 8 : 4;
  PPST_REQUIRE(c % vec == 0 && c / vec <= 256, PPST_ERR_UNSUPPORTED, "act_grad_sum: c=%d (a multiple of %d, at most %d)", c, vec,
               256 * vec);
  const int64_t blocks = (pixels + kAgChunk - 1) / kAgChunk;
  PPST_REQUIRE(blocks < (1LL << 31), PPST_ERR_ARG, "act_grad_sum: too many pixels");
  const size_t smem = sizeof(float) * (size_t)(256 / (c / vec)) * c;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == PPST_BF16)
    act_grad_sum_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, smem, st>>>((__nv_bfloat16*)gp, scratch, (const __nv_bfloat16*)g,
                                                                          (const __nv_bfloat16*)y, pixels, c, alpha, scale, act);
  else
    act_grad_sum_kernel<float><<<(unsigned)blocks, 256, smem, st>>>((float*)gp, scratch, (const float*)g, (const float*)y, pixels,
                                                                  c, alpha, scale, act);
  PPST_LAUNCH_CHECK("act_grad_sum_kernel");
  act_grad_sum_final<<<(c + 127) / 128, 128, 0, st>>>(bias_grad, scratch, blocks, c);
  PPST_LAUNCH_CHECK("act_grad_sum_final");
  return 0;
}
