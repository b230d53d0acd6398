// Per-step weight preparation for the training path: the parameters stay fp32 in the reference's
// [co, ci, kh, kw] layout (state_dict contract), so every step re-derives what the tensor-core kernels consume.
//   ppst_fold_fir          Blur (upfirdn2d, up = down = 1) followed by a convolution == one convolution with the stencil
//                          product of the two kernels (Blur + EqualConv2d of a downsampling ConvLayer,
//                          stylegan2_layers.py:142-164,497-555); adjoint = 1 maps the gradient of the folded weight back.
//   ppst_pack_conv_weights f32 [co, ci, kh, kw] -> bf16 [groups][rows][slots * cols]: the forward packing (tap-major,
//                          channel-minor), the rotated + transposed packing of the input gradient, and the four sub-pixel
//                          phase kernels of a stride-2 input gradient are all "slot s of group g reads tap src[g][s]".
// Work is proportional to the parameter count (tens of MB), not to the activations.
#include "common.cuh"

namespace ppst {
namespace {

// fold:    out[m][a][b] = sum_{i,j} w[m][i][j] * kf[a - i][b - j]            (a < kh + fh - 1, b < kw + fw - 1)
// adjoint: out[m][i][j] = sum_{a,b} w[m][a][b] * kf[a - i][b - j]            (w is the folded-size tensor)
// kf = flip(k).  m runs over co * ci.
__global__ void __launch_bounds__(256) fold_fir_kernel(float* __restrict__ out, const float* __restrict__ w,
                                                       const float* __restrict__ k, int64_t m, int kh, int kw, int fh,
                                                       int fw, int adjoint) {
  const int oh = kh + fh - 1, ow = kw + fw - 1;
  const int per = adjoint ? kh * kw : oh * ow;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m * per) return;
  const int64_t mi = i / per;
  const int r = (int)(i - mi * per);
  float acc = 0.0f;
  if (!adjoint) {
    const int a = r / ow, b = r - a * ow;
    const float* wp = w + mi * kh * kw;
    for (int y = 0; y < kh; ++y) {
      const int fy = a - y;
      if (fy < 0 || fy >= fh) continue;
      for (int x = 0; x < kw; ++x) {
        const int fx = b - x;
        if (fx < 0 || fx >= fw) continue;
        acc = fmaf(wp[y * kw + x], k[(fh - 1 - fy) * fw + (fw - 1 - fx)], acc);
      }
    }
  } else {
    const int y = r / kw, x = r - y * kw;
    const float* wp = w + mi * oh * ow;
    for (int fy = 0; fy < fh; ++fy)
      for (int fx = 0; fx < fw; ++fx)
        acc = fmaf(wp[(y + fy) * ow + (x + fx)], k[(fh - 1 - fy) * fw + (fw - 1 - fx)], acc);
  }
  out[i] = acc;
}

struct SlotTable {
  int n;            // slots per group
  short src[256];   // [group][slot] -> ty * kw + tx, or -1 (zero)
};

// out[g][row][s * cols + col] = scale * w[transpose ? (col, row) : (row, col)][src tap].  One thread per (row, col): it walks
// the taps of its filter (one contiguous run of `taps` floats), so every sector it touches is used completely, and the
// writes of a warp are consecutive `col`s.
__global__ void __launch_bounds__(256) pack_weights_kernel(__nv_bfloat16* __restrict__ out, const float* __restrict__ w,
                                                           int co, int ci, int taps, int groups, int transpose,
                                                           float scale, const SlotTable tab) {
  const int rows = transpose ? ci : co, cols = transpose ? co : ci;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)rows * cols) return;
  const int col = (int)(i % cols), row = (int)(i / cols);
  const int o = transpose ? col : row, c = transpose ? row : col;
  const float* wp = w + ((int64_t)o * ci + c) * taps;
  for (int g = 0; g < groups; ++g) {
    __nv_bfloat16* op = out + ((int64_t)g * rows + row) * tab.n * cols + col;
    for (int s = 0; s < tab.n; ++s) {
      const int src = tab.src[g * tab.n + s];
      op[(int64_t)s * cols] = __float2bfloat16_rn(src >= 0 ? wp[src] * scale : 0.0f);
    }
  }
}

}  // namespace
}  // namespace ppst

extern "C" int ppst_fold_fir(float* out, const float* w, const float* k, int64_t m, int kh, int kw, int fh, int fw,
                             int adjoint, void* stream) {
  using namespace ppst;
  PPST_REQUIRE(out && w && k && m > 0 && kh > 0 && kw > 0 && fh > 0 && fw > 0, PPST_ERR_ARG, "fold_fir: bad args");
  const int64_t total = m * (adjoint ? kh * kw : (kh + fh - 1) * (kw + fw - 1));
  fold_fir_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(out, w, k, m, kh, kw, fh, fw, adjoint);
  PPST_LAUNCH_CHECK("fold_fir_kernel");
  return 0;
}

extern "C" int ppst_pack_conv_weights(void* out, const float* w, int co, int ci, int kh, int kw, int groups, int slots,
                                      const int* slot_src, int transpose, float scale, void* stream) {
  using namespace ppst;
  PPST_REQUIRE(out && w && slot_src && co > 0 && ci > 0 && kh > 0 && kw > 0 && groups > 0 && slots > 0, PPST_ERR_ARG,
               "pack_conv_weights: bad args");
  PPST_REQUIRE(groups * slots <= 256, PPST_ERR_UNSUPPORTED, "pack_conv_weights: %d x %d slots (256 supported)", groups, slots);
  SlotTable tab;
  tab.n = slots;
  for (int i = 0; i < groups * slots; ++i) {
    PPST_REQUIRE(slot_src[i] >= -1 && slot_src[i] < kh * kw, PPST_ERR_ARG, "pack_conv_weights: slot %d reads tap %d", i, slot_src[i]);
    tab.src[i] = (short)slot_src[i];
  }
  const int64_t total = (int64_t)co * ci;
  pack_weights_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<__nv_bfloat16*>(out), w, co, ci, kh * kw, groups, transpose, scale, tab);
  PPST_LAUNCH_CHECK("pack_weights_kernel");
  return 0;
}
