// upfirdn2d: zero-insert upsample -> pad/crop -> FIR (flipped kernel) -> decimate.
// Replaces upfirdn2d_kernel.cu:53-137 (device) and :140-272 (host dispatch) of the reference.
//
// Data layout [major, H, W, minor]; the Python wrapper always passes minor = 1 (NCHW planes).
// HBM-bound stencil: each block stages an input tile (with halo) in shared memory once, every thread
// produces 4 consecutive outputs along W from registers (sliding window), so each input element is read
// from HBM once and each output written once.  fp32 accumulation for all element types.
#include "common.cuh"

namespace ppst {

__host__ __device__ __forceinline__ int floor_div(int a, int b) {
  int c = a / b;
  if (c * b > a) c--;
  return c;
}

struct UpfirdnParams {
  int64_t major;
  int in_h, in_w, minor, kh, kw;
  int up_x, up_y, down_x, down_y;
  int px0, py0;
  int out_h, out_w;
};

constexpr int kMaxTaps = 32 * 32;

// ---- fast path: up = 1, down = 1, minor = 1, kernel <= KMAX x KMAX (modes 1 and 2 of the reference; the
//      encoder / discriminator blur).  Tile = 32 rows x 128 cols of outputs per 256-thread block.
template <typename T, int KMAX>
__global__ void __launch_bounds__(256) upfirdn2d_blur_kernel(T* __restrict__ out, const T* __restrict__ in,
                                                             const float* __restrict__ kernel, UpfirdnParams p) {
  constexpr int TH = 32, TW = 128, PER = 4;          // outputs per thread along x
  constexpr int IH = TH + KMAX - 1, IW = TW + KMAX - 1;
  constexpr int IWP = 132;                           // row pitch: multiple of 4 floats, >= 124 + 8
  static_assert(IW <= IWP && PER + KMAX - 1 <= 8, "window must fit two float4 reads");
  __shared__ float sk[KMAX][KMAX];
  __shared__ __align__(16) float sx[IH][IWP];
  const int tid = threadIdx.x;
  const int64_t plane = blockIdx.x;
  const int oy0 = blockIdx.y * TH, ox0 = blockIdx.z * TW;
  if (tid < KMAX * KMAX) {
    const int ky = tid / KMAX, kx = tid % KMAX;
    // flipped taps, zero-extended to KMAX x KMAX
    sk[ky][kx] = (ky < p.kh && kx < p.kw) ? kernel[(p.kh - 1 - ky) * p.kw + (p.kw - 1 - kx)] : 0.0f;
  }
  const int iy0 = oy0 - p.py0, ix0 = ox0 - p.px0;
  const T* src = in + plane * (int64_t)p.in_h * p.in_w;
  const int need_h = min(TH, p.out_h - oy0) + p.kh - 1;
  const int need_w = min(TW, p.out_w - ox0) + p.kw - 1;
  for (int idx = tid; idx < IH * IWP; idx += 256) {
    const int ry = idx / IWP, rx = idx - ry * IWP;
    const int y = iy0 + ry, x = ix0 + rx;
    float v = 0.0f;
    if (ry < need_h && rx < need_w && y >= 0 && y < p.in_h && x >= 0 && x < p.in_w)
      v = Cvt<T>::to_f(src[(int64_t)y * p.in_w + x]);
    sx[ry][rx] = v;
  }
  __syncthreads();
  // thread -> (row group, column quad): 32 quads per row, 8 rows per pass
  const int qx = (tid & 31) * PER;
  T* dst = out + plane * (int64_t)p.out_h * p.out_w;
#pragma unroll
  for (int pass = 0; pass < TH / 8; ++pass) {
    const int ry = pass * 8 + (tid >> 5);
    const int oy = oy0 + ry;
    float acc[PER] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int ky = 0; ky < KMAX; ++ky) {
      float win[8];
      const float4 w0 = *reinterpret_cast<const float4*>(&sx[ry + ky][qx]);
      const float4 w1 = *reinterpret_cast<const float4*>(&sx[ry + ky][qx + 4]);
      win[0] = w0.x; win[1] = w0.y; win[2] = w0.z; win[3] = w0.w;
      win[4] = w1.x; win[5] = w1.y; win[6] = w1.z; win[7] = w1.w;
#pragma unroll
      for (int kx = 0; kx < KMAX; ++kx) {
        const float w = sk[ky][kx];
#pragma unroll
        for (int j = 0; j < PER; ++j) acc[j] = fmaf(win[j + kx], w, acc[j]);
      }
    }
    if (oy < p.out_h) {
      const int ox = ox0 + qx;
      T* q = dst + (int64_t)oy * p.out_w + ox;
      if (ox + PER <= p.out_w && (reinterpret_cast<uintptr_t>(q) & (sizeof(T) * PER - 1)) == 0) {
        if (sizeof(T) == 4) {
          *reinterpret_cast<float4*>(q) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        } else {
          union { uint2 u; T t[4]; } pk;
#pragma unroll
          for (int j = 0; j < PER; ++j) pk.t[j] = Cvt<T>::from_f(acc[j]);
          *reinterpret_cast<uint2*>(q) = pk.u;
        }
      } else {
#pragma unroll
        for (int j = 0; j < PER; ++j)
          if (ox + j < p.out_w) q[j] = Cvt<T>::from_f(acc[j]);
      }
    }
  }
}

// ---- tiled path for integer up / down factors (modes 3-6 of the reference and any other small config),
//      minor = 1.  Tile = 16 x 64 outputs; the input tile is staged in dynamic shared memory.
template <typename T>
__global__ void __launch_bounds__(256) upfirdn2d_tiled_kernel(T* __restrict__ out, const T* __restrict__ in,
                                                              const float* __restrict__ kernel, UpfirdnParams p,
                                                              int tile_in_h, int tile_in_w) {
  constexpr int TH = 16, TW = 64;
  extern __shared__ float smem[];
  float* sk = smem;                         // kh*kw flipped taps
  float* sx = smem + p.kh * p.kw;           // tile_in_h x tile_in_w
  const int tid = threadIdx.x;
  const int64_t plane = blockIdx.x;
  const int oy0 = blockIdx.y * TH, ox0 = blockIdx.z * TW;
  for (int t = tid; t < p.kh * p.kw; t += 256) {
    const int ky = t / p.kw, kx = t - ky * p.kw;
    sk[t] = kernel[(p.kh - 1 - ky) * p.kw + (p.kw - 1 - kx)];
  }
  // coordinate of tap 0 of output (oy0, ox0) in the zero-inserted (upsampled) image
  const int mid_y0 = oy0 * p.down_y - p.py0, mid_x0 = ox0 * p.down_x - p.px0;
  // first input row/col that can contribute: ceil(mid/up)
  const int in_y0 = -floor_div(-mid_y0, p.up_y), in_x0 = -floor_div(-mid_x0, p.up_x);
  const T* src = in + plane * (int64_t)p.in_h * p.in_w;
  for (int idx = tid; idx < tile_in_h * tile_in_w; idx += 256) {
    const int ry = idx / tile_in_w, rx = idx - ry * tile_in_w;
    const int y = in_y0 + ry, x = in_x0 + rx;
    float v = 0.0f;
    if (y >= 0 && y < p.in_h && x >= 0 && x < p.in_w) v = Cvt<T>::to_f(src[(int64_t)y * p.in_w + x]);
    sx[idx] = v;
  }
  __syncthreads();
  T* dst = out + plane * (int64_t)p.out_h * p.out_w;
  for (int o = tid; o < TH * TW; o += 256) {
    const int ry = o / TW, rx = o - ry * TW;
    const int oy = oy0 + ry, ox = ox0 + rx;
    if (oy >= p.out_h || ox >= p.out_w) continue;
    const int my = oy * p.down_y - p.py0, mx = ox * p.down_x - p.px0;   // upsampled coordinate of tap 0
    // first tap whose upsampled coordinate is a multiple of `up`
    int ky0 = ((-my) % p.up_y + p.up_y) % p.up_y;
    int kx0 = ((-mx) % p.up_x + p.up_x) % p.up_x;
    float acc = 0.0f;
    for (int ky = ky0; ky < p.kh; ky += p.up_y) {
      const int iy = (my + ky) / p.up_y - in_y0;      // exact division (my+ky multiple of up_y); may be <0 -> zero row
      const int gy = (my + ky) / p.up_y;
      if (gy < 0 || gy >= p.in_h) continue;
      for (int kx = kx0; kx < p.kw; kx += p.up_x) {
        const int gx = (mx + kx) / p.up_x;
        if (gx < 0 || gx >= p.in_w) continue;
        acc = fmaf(sx[iy * tile_in_w + (gx - in_x0)], sk[ky * p.kw + kx], acc);
      }
    }
    dst[(int64_t)oy * p.out_w + ox] = Cvt<T>::from_f(acc);
  }
}

// ---- generic path: any minor, any factor, kernel <= 32x32.  One thread per output element, taps in shared
//      memory, inputs straight from global memory (L1/L2 provide the reuse).
template <typename T>
__global__ void __launch_bounds__(256) upfirdn2d_generic_kernel(T* __restrict__ out, const T* __restrict__ in,
                                                                const float* __restrict__ kernel, UpfirdnParams p) {
  extern __shared__ float sk[];
  for (int t = threadIdx.x; t < p.kh * p.kw; t += blockDim.x) {
    const int ky = t / p.kw, kx = t - ky * p.kw;
    sk[t] = kernel[(p.kh - 1 - ky) * p.kw + (p.kw - 1 - kx)];
  }
  __syncthreads();
  const int64_t total = p.major * (int64_t)p.out_h * p.out_w * p.minor;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i;
    const int mi = (int)(r % p.minor); r /= p.minor;
    const int ox = (int)(r % p.out_w); r /= p.out_w;
    const int oy = (int)(r % p.out_h); r /= p.out_h;
    const int64_t mj = r;
    const int my = oy * p.down_y - p.py0, mx = ox * p.down_x - p.px0;
    const int ky0 = ((-my) % p.up_y + p.up_y) % p.up_y;
    const int kx0 = ((-mx) % p.up_x + p.up_x) % p.up_x;
    float acc = 0.0f;
    for (int ky = ky0; ky < p.kh; ky += p.up_y) {
      const int gy = (my + ky) / p.up_y;
      if ((my + ky) < 0 || gy >= p.in_h) continue;
      for (int kx = kx0; kx < p.kw; kx += p.up_x) {
        const int gx = (mx + kx) / p.up_x;
        if ((mx + kx) < 0 || gx >= p.in_w) continue;
        acc = fmaf(Cvt<T>::to_f(in[((mj * p.in_h + gy) * p.in_w + gx) * p.minor + mi]), sk[ky * p.kw + kx], acc);
      }
    }
    out[i] = Cvt<T>::from_f(acc);
  }
}

template <typename T>
static int launch_upfirdn2d(T* y, const T* x, const float* k, const UpfirdnParams& p, cudaStream_t st) {
  const bool planes_ok = p.minor == 1 && p.major < (1LL << 31);
  if (planes_ok && p.up_x == 1 && p.up_y == 1 && p.down_x == 1 && p.down_y == 1 && p.kh <= 4 && p.kw <= 4) {
    dim3 grid((unsigned)p.major, ceil_div(p.out_h, 32), ceil_div(p.out_w, 128));
    if (grid.y <= 65535 && grid.z <= 65535) {
      if (p.kh <= 3 && p.kw <= 3)
        upfirdn2d_blur_kernel<T, 3><<<grid, 256, 0, st>>>(y, x, k, p);
      else
        upfirdn2d_blur_kernel<T, 4><<<grid, 256, 0, st>>>(y, x, k, p);
      PPST_LAUNCH_CHECK("upfirdn2d_blur_kernel");
      return 0;
    }
  }
  if (planes_ok && p.kh <= 8 && p.kw <= 8 && p.up_x <= 4 && p.up_y <= 4 && p.down_x <= 4 && p.down_y <= 4) {
    const int tile_in_h = ((16 - 1) * p.down_y + p.kh - 1) / p.up_y + 2;
    const int tile_in_w = ((64 - 1) * p.down_x + p.kw - 1) / p.up_x + 2;
    const size_t smem = sizeof(float) * (p.kh * p.kw + (size_t)tile_in_h * tile_in_w);
    dim3 grid((unsigned)p.major, ceil_div(p.out_h, 16), ceil_div(p.out_w, 64));
    if (smem <= 48 * 1024 && grid.y <= 65535 && grid.z <= 65535) {
      upfirdn2d_tiled_kernel<T><<<grid, 256, smem, st>>>(y, x, k, p, tile_in_h, tile_in_w);
      PPST_LAUNCH_CHECK("upfirdn2d_tiled_kernel");
      return 0;
    }
  }
  const int64_t total = p.major * (int64_t)p.out_h * p.out_w * p.minor;
  const int blocks = (int)((total + 255) / 256 < 148 * 32 ? (total + 255) / 256 : 148 * 32);
  upfirdn2d_generic_kernel<T><<<blocks, 256, sizeof(float) * p.kh * p.kw, st>>>(y, x, k, p);
  PPST_LAUNCH_CHECK("upfirdn2d_generic_kernel");
  return 0;
}

}  // namespace ppst

extern "C" int ppst_upfirdn2d(void* y, const void* x, const float* kernel, int64_t major, int in_h, int in_w,
                              int minor, int kh, int kw, int up_x, int up_y, int down_x, int down_y, int px0,
                              int px1, int py0, int py1, int dtype, void* stream) {
  using namespace ppst;
  PPST_REQUIRE(kernel && major >= 0 && in_h > 0 && in_w > 0 && minor > 0, PPST_ERR_ARG, "upfirdn2d: bad shape");
  PPST_REQUIRE(up_x >= 1 && up_y >= 1 && down_x >= 1 && down_y >= 1, PPST_ERR_ARG, "upfirdn2d: up/down must be >= 1");
  PPST_REQUIRE(kh >= 1 && kw >= 1 && kh * kw <= kMaxTaps, PPST_ERR_UNSUPPORTED,
               "upfirdn2d: kernel %dx%d exceeds the supported 32x32 taps", kh, kw);
  UpfirdnParams p;
  p.major = major; p.in_h = in_h; p.in_w = in_w; p.minor = minor; p.kh = kh; p.kw = kw;
  p.up_x = up_x; p.up_y = up_y; p.down_x = down_x; p.down_y = down_y; p.px0 = px0; p.py0 = py0;
  p.out_h = (in_h * up_y + py0 + py1 - kh + down_y) / down_y;
  p.out_w = (in_w * up_x + px0 + px1 - kw + down_x) / down_x;
  PPST_REQUIRE(p.out_h > 0 && p.out_w > 0, PPST_ERR_ARG, "upfirdn2d: empty output %dx%d", p.out_h, p.out_w);
  if (major == 0) return 0;
  PPST_REQUIRE(y && x, PPST_ERR_ARG, "upfirdn2d: null buffer");
  cudaStream_t st = (cudaStream_t)stream;
  PPST_DISPATCH_DTYPE(dtype, T, return launch_upfirdn2d<T>((T*)y, (const T*)x, kernel, p, st));
  return 0;
}
