// Image pre- / post-processing around the swap path, on the device:
//   * the data loader's PIL resize (data/base_dataset.py:85-172: Image.BICUBIC in __scale_shortside / __make_power_2 ...)
//     as Pillow's two-pass fixed-point resampler for 8-bit images (third-party: Pillow, requirements.txt pins 10.4.0,
//     src/libImaging/Resample.c: precompute_coeffs / normalize_coeffs_8bpc / ImagingResampleHorizontal_8bpc /
//     ImagingResampleVertical_8bpc).  The coefficient tables are built on the host (ppst_b200/imageio.py) exactly as
//     Pillow builds them; the kernels do the integer accumulation  ss = 2^21 + sum pixel * k;  out = clip8(ss >> 22).
//     Bit-exact against PIL (integer work).
//   * ToTensor + Normalize(0.5, 0.5) (base_dataset.py:143-148): uint8 HWC -> f32 NCHW, (x / 255 - 0.5) / 0.5 in fp32 with
//     the same operation order as torchvision (bit-exact).
//   * util.tensor2im (util/util.py:98-131): f32 NCHW -> uint8 HWC, ((x + 1) / 2 * 255) clipped to [0, 255] and truncated,
//     in fp32 like the numpy expression (bit-exact).
#include "common.cuh"

namespace ppst {

constexpr int kPrecisionBits = 32 - 8 - 2;

__device__ __forceinline__ uint8_t clip8(int v) {
  v >>= kPrecisionBits;                                   // arithmetic shift, like Pillow's clip8 lookup
  return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// One pass along x: in [n, h, in_w, c] -> out [n, h, out_w, c].  bounds [out_w][2] = (xmin, count), kk [out_w][ksize].
__global__ void __launch_bounds__(256) resample_u8_h_kernel(uint8_t* __restrict__ out, const uint8_t* __restrict__ in,
                                                            const int* __restrict__ bounds, const int* __restrict__ kk,
                                                            int64_t rows, int in_w, int out_w, int c, int ksize) {
  const int64_t total = rows * out_w * c;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int ch = (int)(t % c);
    const int xx = (int)((t / c) % out_w);
    const int64_t row = t / ((int64_t)c * out_w);
    const int xmin = bounds[2 * xx], cnt = bounds[2 * xx + 1];
    const int* k = kk + (int64_t)xx * ksize;
    const uint8_t* src = in + (row * in_w + xmin) * c + ch;
    int ss = 1 << (kPrecisionBits - 1);
    for (int x = 0; x < cnt; ++x) ss += (int)src[(int64_t)x * c] * k[x];
    out[t] = clip8(ss);
  }
}

// One pass along y: in [n, in_h, w, c] -> out [n, out_h, w, c].
__global__ void __launch_bounds__(256) resample_u8_v_kernel(uint8_t* __restrict__ out, const uint8_t* __restrict__ in,
                                                            const int* __restrict__ bounds, const int* __restrict__ kk,
                                                            int64_t n, int in_h, int out_h, int64_t wc, int ksize) {
  const int64_t total = n * out_h * wc;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t col = t % wc;
    const int yy = (int)((t / wc) % out_h);
    const int64_t ni = t / (wc * out_h);
    const int ymin = bounds[2 * yy], cnt = bounds[2 * yy + 1];
    const int* k = kk + (int64_t)yy * ksize;
    const uint8_t* src = in + (ni * in_h + ymin) * wc + col;
    int ss = 1 << (kPrecisionBits - 1);
    for (int y = 0; y < cnt; ++y) ss += (int)src[(int64_t)y * wc] * k[y];
    out[t] = clip8(ss);
  }
}

// uint8 [n, h, w, c] -> f32 [n, c, h, w]:  (x / 255 - mean) / std, each step rounded to fp32 (torchvision ToTensor + Normalize)
__global__ void __launch_bounds__(256) u8_to_nchw_norm_kernel(float* __restrict__ out, const uint8_t* __restrict__ in,
                                                              int64_t n, int64_t hw, int c, float mean, float stdv) {
  const int64_t total = n * c * hw;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = t % hw;
    const int ch = (int)((t / hw) % c);
    const int64_t ni = t / (hw * c);
    const float v = __fdiv_rn((float)in[(ni * hw + p) * c + ch], 255.0f);
    out[t] = __fdiv_rn(__fsub_rn(v, mean), stdv);
  }
}

// f32 [n, c, h, w] -> uint8 [n, h, w, c]:  trunc(clip((x + 1) / 2 * 255, 0, 255))   (normalize = True)
//                                          trunc(clip(x * 255, 0, 255))             (normalize = False)
__global__ void __launch_bounds__(256) tensor2im_kernel(uint8_t* __restrict__ out, const float* __restrict__ in, int64_t n,
                                                        int64_t hw, int c, int normalize) {
  const int64_t total = n * hw * c;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int ch = (int)(t % c);
    const int64_t p = (t / c) % hw;
    const int64_t ni = t / (hw * c);
    float v = in[(ni * c + ch) * hw + p];
    v = normalize ? __fmul_rn(__fdiv_rn(__fadd_rn(v, 1.0f), 2.0f), 255.0f) : __fmul_rn(v, 255.0f);
    v = fminf(fmaxf(v, 0.0f), 255.0f);                    // np.clip; NaN -> 0 here (numpy: undefined cast)
    out[t] = (uint8_t)(int)v;
  }
}

static int blocks_for(int64_t total) {
  const int64_t b = (total + 255) / 256;
  return (int)(b < 148 * 16 ? (b > 0 ? b : 1) : 148 * 16);
}

}  // namespace ppst

extern "C" int ppst_resample_u8(uint8_t* out, const uint8_t* in, const int* bounds, const int* kk, int64_t n, int in_h,
                                int in_w, int out_len, int c, int ksize, int axis, void* stream) {
  using namespace ppst;
  PPST_REQUIRE(out && in && bounds && kk && n > 0 && in_h > 0 && in_w > 0 && out_len > 0 && c > 0 && ksize > 0,
               PPST_ERR_ARG, "resample_u8: bad args");
  PPST_REQUIRE(axis == 0 || axis == 1, PPST_ERR_ARG, "resample_u8: axis must be 0 (x) or 1 (y)");
  cudaStream_t st = (cudaStream_t)stream;
  if (axis == 0) {
    const int64_t rows = n * in_h;
    resample_u8_h_kernel<<<blocks_for(rows * out_len * c), 256, 0, st>>>(out, in, bounds, kk, rows, in_w, out_len, c, ksize);
  } else {
    const int64_t wc = (int64_t)in_w * c;
    resample_u8_v_kernel<<<blocks_for(n * out_len * wc), 256, 0, st>>>(out, in, bounds, kk, n, in_h, out_len, wc, ksize);
  }
  PPST_LAUNCH_CHECK("resample_u8_kernel");
  return 0;
}

extern "C" int ppst_u8_to_nchw_norm(float* out, const uint8_t* in, int64_t n, int64_t hw, int c, float mean, float stdv,
                                    void* stream) {
  using namespace ppst;
  PPST_REQUIRE(out && in && n > 0 && hw > 0 && c > 0 && stdv != 0.0f, PPST_ERR_ARG, "u8_to_nchw_norm: bad args");
  u8_to_nchw_norm_kernel<<<blocks_for(n * hw * c), 256, 0, (cudaStream_t)stream>>>(out, in, n, hw, c, mean, stdv);
  PPST_LAUNCH_CHECK("u8_to_nchw_norm_kernel");
  return 0;
}

extern "C" int ppst_tensor2im(uint8_t* out, const float* in, int64_t n, int64_t hw, int c, int normalize, void* stream) {
  using namespace ppst;
  PPST_REQUIRE(out && in && n > 0 && hw > 0 && c > 0, PPST_ERR_ARG, "tensor2im: bad args");
  tensor2im_kernel<<<blocks_for(n * hw * c), 256, 0, (cudaStream_t)stream>>>(out, in, n, hw, c, normalize);
  PPST_LAUNCH_CHECK("tensor2im_kernel");
  return 0;
}
