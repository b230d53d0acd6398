// Weight gradient of a convolution on the tcgen05 path (training step, SURVEY.md 8(f)4: "conv dgrad/wgrad on tcgen05").
// Replaces what torch.autograd computes for F.conv2d's weight in EqualConv2d / EqualizedConv2d / nn.Conv2d
// (stylegan2_layers.py:188,320,331,339-341; generator.py:14-19) -- cuDNN's wgrad in the reference.
//
//   dw[co][ty][tx][ci] = sum_{n,oy,ox} dy[n,oy,ox,co] * x[n, oy*s + ty + off_y, ox*s + tx + off_x, ci]        (x = 0 outside)
//
// is a GEMM whose contraction runs over PIXELS.  NHWC tensors are pixel-major with contiguous channels, i.e. both
// operands are already in the "MN-major" form tcgen05.mma accepts, so the kernel (conv_wgrad_mn_kernel, conv_tc.cu) reads
// dy and x straight through TMA: no transposition pass, no padded copies; a tap is a coordinate offset of the x box, a
// stride is a TMA element stride, zero padding is TMA out-of-bounds fill.  Two launches, both deterministic:
//   1. conv_wgrad_mn_kernel: every (sample, pixel chunk) "j", tap, 128-row co tile and ci tile -> one fp32 tile of
//      partials [j][tap][co][ci]; taps are the fastest work-item index, so the CTAs that stream the same dy chunk (and
//      neighbouring x boxes) run side by side and share it through L2.
//   2. reduce_kernel: fixed-order sum over j into the packed weight layout [co][tap][ci] (no atomics).
// (A first version transposed both operands into pixel-contiguous planes and ran the K-major GEMM of the forward kernel:
//  390-620 TF/s, half of it the transposition.  This one: 1030-1420 TF/s on the layers that matter, profiles/.)
#include <cstring>

#include "common.cuh"

namespace ppst {

int tc_wgrad_mn(float* partials, const void* dy, const void* x, const ppst_wgrad_desc& d, int k_split, cudaStream_t st);

namespace {

// dw[co][tap][ci] = sum_j partials[j][tap][co][ci]   (j ascending: fixed order)
__global__ void __launch_bounds__(256) reduce_kernel(float* __restrict__ dw, const float* __restrict__ partials, int J,
                                                     int co, int taps, int ci) {
  const int64_t per_j = (int64_t)taps * co * ci;
  const int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (e >= per_j) return;
  const int c = (int)(e % ci);
  int64_t r = e / ci;
  const int o = (int)(r % co);
  const int t = (int)(r / co);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int j = 0; j < J; ++j) {
    const float4 v = *reinterpret_cast<const float4*>(partials + (int64_t)j * per_j + e);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  *reinterpret_cast<float4*>(dw + ((int64_t)o * taps + t) * ci + c) = acc;
}

int check_desc(const ppst_wgrad_desc& d) {
  PPST_REQUIRE(d.n > 0 && d.in_h > 0 && d.in_w > 0 && d.out_h > 0 && d.out_w > 0 && d.kh > 0 && d.kw > 0, PPST_ERR_ARG,
               "conv_wgrad: bad geometry");
  PPST_REQUIRE(d.stride == 1 || d.stride == 2, PPST_ERR_UNSUPPORTED, "conv_wgrad: stride %d", d.stride);
  PPST_REQUIRE(d.cin % 8 == 0 && d.cout % 8 == 0 && d.cin > 0 && d.cout > 0, PPST_ERR_UNSUPPORTED,
               "conv_wgrad: cin=%d and cout=%d must be multiples of 8 (16-byte pixel pitch)", d.cin, d.cout);
  return 0;
}

// pixel chunks per sample: about two waves of work items, at least 16 patches (1024 pixels) per chunk, no empty chunk
int k_split(const ppst_wgrad_desc& d) {
  const int patches = ceil_div(d.out_w, 8) * ceil_div(d.out_h, 8);
  const int bn = d.cin > 128 ? 256 : d.cin > 64 ? 128 : 64;
  const int64_t tiles = (int64_t)ceil_div(d.cout, 128) * ceil_div(d.cin, bn) * d.kh * d.kw;
  int64_t ks = (2 * 148 + d.n * tiles - 1) / (d.n * tiles);
  const int64_t ks_max = patches / 16 > 0 ? patches / 16 : 1;
  if (ks > ks_max) ks = ks_max;
  if (ks < 1) ks = 1;
  const int chunk = ceil_div(patches, ks);
  return ceil_div(patches, chunk);
}

}  // namespace
}  // namespace ppst

extern "C" int64_t ppst_conv_wgrad_workspace_bytes(const ppst_wgrad_desc* d) {
  if (!d || ppst::check_desc(*d)) return -1;
  return (int64_t)d->n * ppst::k_split(*d) * d->kh * d->kw * d->cout * d->cin * 4 + 512;
}

extern "C" int ppst_conv_wgrad(float* dw, const void* x, const void* dy, const ppst_wgrad_desc* dp, void* workspace,
                               int64_t workspace_bytes, void* stream) {
  using namespace ppst;
  PPST_REQUIRE(dw && x && dy && dp && workspace, PPST_ERR_ARG, "conv_wgrad: null argument");
  const ppst_wgrad_desc d = *dp;
  int rc = check_desc(d);
  if (rc) return rc;
  const int ks = k_split(d);
  const int taps = d.kh * d.kw;
  const int64_t need = (int64_t)d.n * ks * taps * d.cout * d.cin * 4 + 512;
  PPST_REQUIRE(workspace_bytes >= need, PPST_ERR_ARG, "conv_wgrad: workspace of %lld bytes, %lld needed",
               (long long)workspace_bytes, (long long)need);
  PPST_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dw)) & 15) == 0,
               PPST_ERR_ARG, "conv_wgrad: buffers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  float* partials = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
  rc = tc_wgrad_mn(partials, dy, x, d, ks, st);
  if (rc) return rc;
  const int64_t threads = (int64_t)taps * d.cout * d.cin / 4;
  reduce_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(dw, partials, d.n * ks, d.cout, taps, d.cin);
  PPST_LAUNCH_CHECK("conv_wgrad reduce_kernel");
  return 0;
}
