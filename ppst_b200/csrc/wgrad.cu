// Weight gradient of a convolution on the tcgen05 path (training step, SURVEY.md 8(f)4: "conv dgrad/wgrad on tcgen05").
// Replaces what torch.autograd computes for F.conv2d's weight in EqualConv2d / EqualizedConv2d / nn.Conv2d
// (stylegan2_layers.py:188,320,331,339-341; generator.py:14-19) -- cuDNN's wgrad in the reference.
//
//   dw[co][ty][tx][ci] = sum_{n,oy,ox} dy[n,oy,ox,co] * x[n, oy*s + ty + off_y, ox*s + tx + off_x, ci]        (x = 0 outside)
//
// is a GEMM whose contraction runs over PIXELS, so both operands must be pixel-major ("K-major") for the tensor cores.
// Three steps, all on the caller's stream, all deterministic (no atomics):
//   1. planes_kernel: x (NHWC) -> xt [n][s*s parity planes][column shifts][ci][K], dy -> dyt [n][co][K], K = flattened
//      (row, col) with a common zero-padded row pitch Wp (a multiple of 8).  With that pitch a filter tap is a FLAT
//      shift of xt along K: the zero columns absorb the horizontal wrap-around, TMA's out-of-bounds zero fill the vertical
//      one.  TMA needs a 16-byte aligned start in the innermost dimension, so only the vertical part of the shift
//      (a multiple of Wp) is a coordinate offset; the horizontal part (-1, 0, +1 columns for a 3x3 filter) selects one
//      of the column-shifted copies of the plane that planes_kernel writes.
//   2. tc_wgrad_gemm (conv_tc.cu): one launch of the tcgen05 kernel computes, for every (sample, K chunk) "image" and
//      every tap, the 128 x BN tile  dyt . shift_tap(xt)^T  into fp32 partials; taps are the fastest tile index, so the
//      CTAs that stream the same K range run side by side and share it through L2 (operands leave HBM once).
//   3. reduce_kernel: fixed-order sum of the partials over (sample, chunk) into the packed weight layout [co][tap][ci].
#include <cstring>
#include <cstdlib>

#include "common.cuh"

namespace ppst {

int tc_wgrad_gemm(float* partials, const void* dyt, const void* xt, int samples, int k_split, int kc, int m, int ci,
                  int xt_rows, int n_taps, const int* k_off, const int* row_off, cudaStream_t st);
int tc_wgrad_mn(float* partials, const void* dy, const void* x, const ppst_wgrad_desc& d, int k_split, cudaStream_t st);

namespace {

// in: bf16 [n][H][W][C]  ->  out: bf16 [n][s*s][ncopies][C][kpad];  plane (py, px), copy e, element f holds the plane's
// flat element f + dx0 + e, where flat element y'*wp + x' = in[s*y' + py][s*x' + px] (0 where that pixel does not exist,
// in the pad columns and in the tail up to kpad).
// One block = 64 consecutive f x 64 channels, transposed through shared memory (32-bit words = channel pairs).
__global__ void __launch_bounds__(256) planes_kernel(__nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ in,
                                                     int H, int W, int C, int s, int wp, int64_t kpad, int ncopies, int dx0) {
  __shared__ uint32_t tile[64][33];
  const int c0 = blockIdx.y * 64;
  const int planes = s * s;
  const int zc = blockIdx.z % ncopies, zp = blockIdx.z / ncopies;
  const int img = zp / planes, plane = zp - img * planes;
  const int py = plane / s, px = plane - py * s;
  const int f0 = blockIdx.x * 64;
  const int shift = dx0 + zc;
  const int t = threadIdx.x;
  {
    const int q = t & 7;                                   // 16-byte chunk = channels c0 + 8q .. + 7
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int pl = (t >> 3) + 32 * h;                    // pixel of the tile
      const int f = f0 + pl + shift;
      const int yy = f / wp, xx = f - yy * wp;
      const int sy = s * yy + py, sx = s * xx + px;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (f >= 0 && sy < H && sx < W && c0 + 8 * q < C)
        v = *reinterpret_cast<const uint4*>(in + (((int64_t)img * H + sy) * W + sx) * C + c0 + 8 * q);
      tile[pl][4 * q + 0] = v.x;
      tile[pl][4 * q + 1] = v.y;
      tile[pl][4 * q + 2] = v.z;
      tile[pl][4 * q + 3] = v.w;
    }
  }
  __syncthreads();
  {
    const int w = t >> 5, l = t & 31;
    const int fq = (l & 3) + 4 * (w & 1);                  // 8 consecutive f
    const int cp = (l >> 2) + 8 * (w >> 1);                // channel pair
    uint32_t r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = tile[8 * fq + i][cp];
    uint4 lo, hi;
    lo.x = __byte_perm(r[0], r[1], 0x5410); hi.x = __byte_perm(r[0], r[1], 0x7632);
    lo.y = __byte_perm(r[2], r[3], 0x5410); hi.y = __byte_perm(r[2], r[3], 0x7632);
    lo.z = __byte_perm(r[4], r[5], 0x5410); hi.z = __byte_perm(r[4], r[5], 0x7632);
    lo.w = __byte_perm(r[6], r[7], 0x5410); hi.w = __byte_perm(r[6], r[7], 0x7632);
    const int c = c0 + 2 * cp;
    if (c < C) {
      __nv_bfloat16* o = out + ((((int64_t)img * planes + plane) * ncopies + zc) * C + c) * kpad + f0 + 8 * fq;
      *reinterpret_cast<uint4*>(o) = lo;
      *reinterpret_cast<uint4*>(o + kpad) = hi;            // C is even
    }
  }
}

// dw[co][tap0 + t][ci] = sum_j partials[j][co / 16][t][co % 16][ci]   (j = (sample, chunk), ascending: fixed order)
__global__ void __launch_bounds__(256) reduce_kernel(float* __restrict__ dw, const float* __restrict__ partials, int J,
                                                     int co, int taps, int tap0, int taps_total, int ci) {
  const int64_t per_j = (int64_t)co * taps * ci;
  const int64_t i4 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i4 * 4 >= per_j) return;
  const int64_t e = i4 * 4;
  const int c = (int)(e % ci);
  int64_t r = e / ci;                                       // ((co/16) * taps + t) * 16 + co % 16
  const int lo = (int)(r % 16); r /= 16;
  const int t = (int)(r % taps);
  const int hi = (int)(r / taps);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const float* p = partials + e;
  for (int j = 0; j < J; ++j) {
    const float4 v = *reinterpret_cast<const float4*>(p + (int64_t)j * per_j);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  *reinterpret_cast<float4*>(dw + ((int64_t)(hi * 16 + lo) * taps_total + tap0 + t) * ci + c) = acc;
}

// path 2 (MN-major operands straight from NHWC): dw[co][tap][ci] = sum_j partials[j][tap][co][ci]
__global__ void __launch_bounds__(256) reduce_mn_kernel(float* __restrict__ dw, const float* __restrict__ partials, int J,
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

// pixel chunks per sample for path 2: about two waves of work items, at least 16 patches (1024 pixels) per chunk, no
// empty chunk
int mn_k_split(const ppst_wgrad_desc& d) {
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

bool use_mn_path() {
  static const int v = [] { const char* e = getenv("PPST_WGRAD_PATH"); return (e && e[0] == 'p') ? 0 : 1; }();
  return v != 0;
}

struct Plan {
  int s, planes, wp, k_split, kc, taps, groups, taps_per_group, ncopies, dx0;
  int64_t kpad, xt_elems, dyt_elems, partial_elems, bytes;
};

int floor_div(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

int make_plan(const ppst_wgrad_desc& d, Plan& pl) {
  PPST_REQUIRE(d.n > 0 && d.in_h > 0 && d.in_w > 0 && d.out_h > 0 && d.out_w > 0 && d.kh > 0 && d.kw > 0, PPST_ERR_ARG,
               "conv_wgrad: bad geometry");
  PPST_REQUIRE(d.stride == 1 || d.stride == 2, PPST_ERR_UNSUPPORTED, "conv_wgrad: stride %d", d.stride);
  PPST_REQUIRE(d.cin % 32 == 0 && d.cout % 16 == 0, PPST_ERR_UNSUPPORTED,
               "conv_wgrad: cin=%d must be a multiple of 32 and cout=%d of 16", d.cin, d.cout);
  const int s = d.stride;
  pl.s = s; pl.planes = s * s;
  // horizontal shifts of the taps, in plane columns
  const int dx_min = floor_div(d.off_x, s), dx_max = floor_div(d.kw - 1 + d.off_x, s);
  const int pw = (d.in_w + s - 1) / s, ph = (d.in_h + s - 1) / s;         // plane size
  int wp = pw + (dx_min < 0 ? -dx_min : 0);
  if (d.out_w + (dx_max > 0 ? dx_max : 0) > wp) wp = d.out_w + (dx_max > 0 ? dx_max : 0);
  if (d.out_w - dx_min > wp) wp = d.out_w - dx_min;                        // dy's own zero columns cover a left shift
  pl.wp = (wp + 7) / 8 * 8;                 // vertical shifts stay 16-byte aligned
  wp = pl.wp;
  pl.dx0 = dx_min; pl.ncopies = dx_max - dx_min + 1;
  const int rows = ph > d.out_h ? ph : d.out_h;
  const int64_t kraw = (int64_t)rows * wp;
  pl.taps = d.kh * d.kw;
  pl.taps_per_group = pl.taps <= 16 ? pl.taps : (pl.taps % 12 == 0 ? 12 : (pl.taps % 9 == 0 ? 9 : (pl.taps % 5 == 0 ? 5 : 1)));
  while (pl.taps % pl.taps_per_group) --pl.taps_per_group;
  pl.groups = pl.taps / pl.taps_per_group;
  const int bn = d.cin % 256 == 0 ? 256 : d.cin % 128 == 0 ? 128 : d.cin % 64 == 0 ? 64 : 32;
  const int64_t tiles = (int64_t)((d.cout + 127) / 128) * (d.cin / bn) * pl.taps_per_group;
  int64_t ks = (2 * 148 + d.n * tiles - 1) / (d.n * tiles);
  const int64_t ks_max = kraw / 2048 > 0 ? kraw / 2048 : 1;
  if (ks > ks_max) ks = ks_max;
  if (ks < 1) ks = 1;
  pl.k_split = (int)ks;
  pl.kc = (int)(((kraw + ks - 1) / ks + 63) / 64 * 64);
  pl.kpad = (int64_t)pl.kc * ks;
  pl.xt_elems = (int64_t)d.n * pl.planes * pl.ncopies * d.cin * pl.kpad;
  pl.dyt_elems = (int64_t)d.n * d.cout * pl.kpad;
  pl.partial_elems = (int64_t)d.n * ks * d.cout * pl.taps_per_group * d.cin;
  pl.bytes = (pl.xt_elems + pl.dyt_elems) * 2 + pl.partial_elems * 4 + 512;
  if (use_mn_path()) pl.bytes = (int64_t)d.n * mn_k_split(d) * pl.taps * d.cout * d.cin * 4 + 512;
  return 0;
}

}  // namespace
}  // namespace ppst

extern "C" int64_t ppst_conv_wgrad_workspace_bytes(const ppst_wgrad_desc* d) {
  if (!d) return -1;
  ppst::Plan pl;
  if (ppst::make_plan(*d, pl)) return -1;
  return pl.bytes;
}

extern "C" int ppst_conv_wgrad(float* dw, const void* x, const void* dy, const ppst_wgrad_desc* dp, void* workspace,
                               int64_t workspace_bytes, void* stream) {
  using namespace ppst;
  PPST_REQUIRE(dw && x && dy && dp && workspace, PPST_ERR_ARG, "conv_wgrad: null argument");
  const ppst_wgrad_desc d = *dp;
  Plan pl;
  int rc = make_plan(d, pl);
  if (rc) return rc;
  PPST_REQUIRE(workspace_bytes >= pl.bytes, PPST_ERR_ARG, "conv_wgrad: workspace of %lld bytes, %lld needed",
               (long long)workspace_bytes, (long long)pl.bytes);
  PPST_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dw)) & 15) == 0,
               PPST_ERR_ARG, "conv_wgrad: buffers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  uintptr_t w0 = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255);
  if (use_mn_path()) {
    float* part = reinterpret_cast<float*>(w0);
    const int ks = mn_k_split(d);
    rc = tc_wgrad_mn(part, dy, x, d, ks, st);
    if (rc) return rc;
    const int64_t threads = (int64_t)pl.taps * d.cout * d.cin / 4;
    reduce_mn_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(dw, part, d.n * ks, d.cout, pl.taps, d.cin);
    PPST_LAUNCH_CHECK("conv_wgrad reduce_mn_kernel");
    return 0;
  }
  __nv_bfloat16* xt = reinterpret_cast<__nv_bfloat16*>(w0);
  __nv_bfloat16* dyt = xt + pl.xt_elems;
  float* partials = reinterpret_cast<float*>(dyt + pl.dyt_elems);
  {
    dim3 grid((unsigned)(pl.kpad / 64), (unsigned)((d.cin + 63) / 64), (unsigned)(d.n * pl.planes * pl.ncopies));
    PPST_REQUIRE(grid.z <= 65535, PPST_ERR_ARG, "conv_wgrad: batch too large");
    planes_kernel<<<grid, 256, 0, st>>>(xt, reinterpret_cast<const __nv_bfloat16*>(x), d.in_h, d.in_w, d.cin, pl.s, pl.wp,
                                       pl.kpad, pl.ncopies, pl.dx0);
    PPST_LAUNCH_CHECK("conv_wgrad planes_kernel(x)");
    dim3 grid2((unsigned)(pl.kpad / 64), (unsigned)((d.cout + 63) / 64), (unsigned)d.n);
    planes_kernel<<<grid2, 256, 0, st>>>(dyt, reinterpret_cast<const __nv_bfloat16*>(dy), d.out_h, d.out_w, d.cout, 1,
                                        pl.wp, pl.kpad, 1, 0);
    PPST_LAUNCH_CHECK("conv_wgrad planes_kernel(dy)");
  }
  for (int g = 0; g < pl.groups; ++g) {
    int k_off[16], row_off[16];
    for (int t = 0; t < pl.taps_per_group; ++t) {
      const int tap = g * pl.taps_per_group + t;
      const int ty = tap / d.kw + d.off_y, tx = tap % d.kw + d.off_x;
      const int py = ((ty % pl.s) + pl.s) % pl.s, px = ((tx % pl.s) + pl.s) % pl.s;
      const int dyy = (ty - py) / pl.s, dxx = (tx - px) / pl.s;
      k_off[t] = dyy * pl.wp;                                          // multiple of 8 elements
      row_off[t] = ((py * pl.s + px) * pl.ncopies + (dxx - pl.dx0)) * d.cin;
    }
    rc = tc_wgrad_gemm(partials, dyt, xt, d.n, pl.k_split, pl.kc, d.cout, d.cin, pl.planes * pl.ncopies * d.cin, pl.taps_per_group,
                       k_off, row_off, st);
    if (rc) return rc;
    const int64_t per_j = (int64_t)d.cout * pl.taps_per_group * d.cin;
    const int64_t threads = per_j / 4;
    reduce_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(dw, partials, d.n * pl.k_split, d.cout,
                                                                      pl.taps_per_group, g * pl.taps_per_group, pl.taps,
                                                                      d.cin);
    PPST_LAUNCH_CHECK("conv_wgrad reduce_kernel");
  }
  return 0;
}
