// Guided-filter post-process of PPSTModel.decode(target=...) (models/ppst_model.py:288-306): replaces the reference's
// device -> host copy, uint8 quantisation (util/util.py:98-131), cv2.ximgproc.guidedFilter(guide, src, r=30,
// eps=(0.02*255)^2) on the CPU (photo_gif.py:13-46) and host -> device copy with a device-resident pipeline.
//
// Colour-guide guided filter (He, Sun, Tang) on the uint8-quantised images:
//   1. 21 integer planes per image: I (3), p (3), I_i*I_j (6), I_i*p_c (9)                         gf_products_kernel
//   2. (2r+1)^2 window sums with BORDER_REFLECT, EXACT in int32 (values <= 255^2 * 61^2 < 2^31):
//      row pass = block-wide prefix scan of a row staged in shared memory, column pass = sliding sums  gf_box_h / gf_box_v
//   3. per pixel: fp32 means, 3x3 covariance (+eps) inverse by cofactors, a = inv * cov(I, p_c), b  gf_solve_kernel
//   4. box means of the 12 float planes a (9), b (3): double accumulation, one rounding to fp32     gf_box_h / gf_box_v
//   5. q_c = mean_a . I + mean_b -> round-half-even, saturate to uint8 -> /255 -> (x - 0.5) * 2    gf_final_kernel
// Every fp32 operation is written with explicit round-to-nearest intrinsics (no FMA contraction), in the order the
// oracle uses, so the two agree except where a double-precision window sum lands within an ulp of a rounding boundary.
// All kernels are HBM streams over planar [plane][H][W] buffers in a caller-provided workspace.
#include "common.cuh"

namespace ppst {

constexpr int kGfP1 = 21, kGfP2 = 12;

__device__ __forceinline__ int gf_reflect(int i, int n) { return i < 0 ? -i - 1 : (i >= n ? 2 * n - 1 - i : i); }

__device__ __forceinline__ int gf_quant(float x) {
  float v = __fmul_rn(__fdiv_rn(__fadd_rn(x, 1.0f), 2.0f), 255.0f);
  v = fminf(fmaxf(v, 0.0f), 255.0f);
  return (int)v;                                   // astype(uint8): truncation
}

// src / guide: NCHW f32 in [-1, 1]; out: int32 [n][21][hw]
__global__ void __launch_bounds__(256) gf_products_kernel(int32_t* __restrict__ out, const float* __restrict__ src,
                                                          const float* __restrict__ guide, int64_t n, int64_t hw) {
  const int64_t total = n * hw;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t img = t / hw, pix = t - img * hw;
    int I[3], P[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      I[i] = gf_quant(guide[(img * 3 + i) * hw + pix]);
      P[i] = gf_quant(src[(img * 3 + i) * hw + pix]);
    }
    int32_t* o = out + img * kGfP1 * hw + pix;
    o[0 * hw] = I[0]; o[1 * hw] = I[1]; o[2 * hw] = I[2];
    o[3 * hw] = P[0]; o[4 * hw] = P[1]; o[5 * hw] = P[2];
    o[6 * hw] = I[0] * I[0]; o[7 * hw] = I[0] * I[1]; o[8 * hw] = I[0] * I[2];
    o[9 * hw] = I[1] * I[1]; o[10 * hw] = I[1] * I[2]; o[11 * hw] = I[2] * I[2];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int i = 0; i < 3; ++i) o[(12 + c * 3 + i) * hw] = I[i] * P[c];
  }
}

// Row pass: out[row][x] = sum_{d=-r..r} in[row][reflect(x+d)], r < w.  One block per row at a time: the row is staged in
// shared memory as an inclusive prefix sum (chunked serial scan + warp-shuffle scan of the 128 chunk totals).
template <typename TIN, typename TACC, typename TOUT>
__global__ void __launch_bounds__(128) gf_box_h_kernel(TOUT* __restrict__ out, const TIN* __restrict__ in, int64_t rows,
                                                       int w, int r) {
  extern __shared__ __align__(16) unsigned char gf_smem[];
  TACC* P = reinterpret_cast<TACC*>(gf_smem);
  __shared__ TACC warp_tot[4];
  const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
  const int ch = (w + 127) / 128;
  for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
    const TIN* src = in + row * w;
    for (int x = tid; x < w; x += 128) P[x] = (TACC)src[x];
    __syncthreads();
    const int x0 = min(tid * ch, w), x1 = min(x0 + ch, w);
    TACC run = 0;
    for (int x = x0; x < x1; ++x) { run += P[x]; P[x] = run; }
    // exclusive scan of the chunk totals
    TACC inc = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const TACC up = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += up;
    }
    if (lane == 31) warp_tot[wrp] = inc;
    __syncthreads();
    TACC base = inc - run;
    for (int i = 0; i < wrp; ++i) base += warp_tot[i];
    for (int x = x0; x < x1; ++x) P[x] += base;
    __syncthreads();
    TOUT* dst = out + row * w;
    for (int x = tid; x < w; x += 128) {
      const int lo = x - r - 1, hi = x + r;
      TACC s = P[min(hi, w - 1)] - (lo >= 0 ? P[lo] : (TACC)0);
      if (x - r < 0) s += P[r - x - 1];                           // reflected [x-r, -1] -> [0, r-x-1]
      if (hi >= w) {                                              // reflected [w, hi] -> [2w-1-hi, w-1]
        const int lo2 = 2 * w - 2 - hi;
        s += P[w - 1] - (lo2 >= 0 ? P[lo2] : (TACC)0);
      }
      dst[x] = (TOUT)s;
    }
    __syncthreads();
  }
}

// Column pass: thread = (plane, segment of `seg` rows, column); 2r+1 loads to start, then a sliding sum.
// scale == 0: raw sums; else out = (TOUT)(sum * scale) with the product taken in TACC.
template <typename TIN, typename TACC, typename TOUT>
__global__ void __launch_bounds__(256) gf_box_v_kernel(TOUT* __restrict__ out, const TIN* __restrict__ in, int64_t planes,
                                                       int h, int w, int r, int seg, double scale) {
  const int nseg = (h + seg - 1) / seg;
  const int64_t total = planes * nseg * w;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(t % w);
    const int64_t q = t / w;
    const int sg = (int)(q % nseg);
    const int64_t pl = q / nseg;
    const TIN* src = in + pl * (int64_t)h * w + x;
    TOUT* dst = out + pl * (int64_t)h * w + x;
    const int y0 = sg * seg, y1 = min(y0 + seg, h);
    TACC s = 0;
    for (int d = -r; d <= r; ++d) s += (TACC)src[(int64_t)gf_reflect(y0 + d, h) * w];
    dst[(int64_t)y0 * w] = scale != 0.0 ? (TOUT)((double)s * scale) : (TOUT)s;
    for (int y = y0 + 1; y < y1; ++y) {
      s += (TACC)src[(int64_t)gf_reflect(y + r, h) * w] - (TACC)src[(int64_t)gf_reflect(y - r - 1, h) * w];
      dst[(int64_t)y * w] = scale != 0.0 ? (TOUT)((double)s * scale) : (TOUT)s;
    }
  }
}

// sums: int32 [n][21][hw] window sums -> ab: f32 [n][12][hw]  (a_c[i] at c*3+i, b_c at 9+c)
__global__ void __launch_bounds__(256) gf_solve_kernel(float* __restrict__ ab, const int32_t* __restrict__ sums, int64_t n,
                                                       int64_t hw, double inv_area, float eps) {
  const int64_t total = n * hw;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t img = t / hw, pix = t - img * hw;
    const int32_t* s = sums + img * kGfP1 * hw + pix;
    float m[kGfP1];
#pragma unroll
    for (int i = 0; i < kGfP1; ++i) m[i] = (float)((double)s[i * hw] * inv_area);
    const float* mI = m;
    const float* mP = m + 3;
    float v[6];
    {
      int k = 0;
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = i; j < 3; ++j, ++k) {
          v[k] = __fsub_rn(m[6 + k], __fmul_rn(mI[i], mI[j]));
          if (i == j) v[k] = __fadd_rn(v[k], eps);
        }
    }
    const float v00 = v[0], v01 = v[1], v02 = v[2], v11 = v[3], v12 = v[4], v22 = v[5];
    const float c00 = __fsub_rn(__fmul_rn(v11, v22), __fmul_rn(v12, v12));
    const float c01 = __fsub_rn(__fmul_rn(v02, v12), __fmul_rn(v01, v22));
    const float c02 = __fsub_rn(__fmul_rn(v01, v12), __fmul_rn(v02, v11));
    const float c11 = __fsub_rn(__fmul_rn(v00, v22), __fmul_rn(v02, v02));
    const float c12 = __fsub_rn(__fmul_rn(v01, v02), __fmul_rn(v00, v12));
    const float c22 = __fsub_rn(__fmul_rn(v00, v11), __fmul_rn(v01, v01));
    const float det = __fadd_rn(__fadd_rn(__fmul_rn(v00, c00), __fmul_rn(v01, c01)), __fmul_rn(v02, c02));
    const float inv[3][3] = {{__fdiv_rn(c00, det), __fdiv_rn(c01, det), __fdiv_rn(c02, det)},
                             {__fdiv_rn(c01, det), __fdiv_rn(c11, det), __fdiv_rn(c12, det)},
                             {__fdiv_rn(c02, det), __fdiv_rn(c12, det), __fdiv_rn(c22, det)}};
    float* o = ab + img * kGfP2 * hw + pix;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float cov[3], a[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) cov[i] = __fsub_rn(m[12 + c * 3 + i], __fmul_rn(mI[i], mP[c]));
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        a[i] = __fadd_rn(__fadd_rn(__fmul_rn(inv[i][0], cov[0]), __fmul_rn(inv[i][1], cov[1])), __fmul_rn(inv[i][2], cov[2]));
        o[(c * 3 + i) * hw] = a[i];
      }
      const float dotv = __fadd_rn(__fadd_rn(__fmul_rn(a[0], mI[0]), __fmul_rn(a[1], mI[1])), __fmul_rn(a[2], mI[2]));
      o[(9 + c) * hw] = __fsub_rn(mP[c], dotv);
    }
  }
}

// mab: f32 [n][12][hw] box means of a, b; guide: NCHW f32; out: NCHW f32
__global__ void __launch_bounds__(256) gf_final_kernel(float* __restrict__ out, const float* __restrict__ mab,
                                                       const float* __restrict__ guide, int64_t n, int64_t hw) {
  const int64_t total = n * hw;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t img = t / hw, pix = t - img * hw;
    float I[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) I[i] = (float)gf_quant(guide[(img * 3 + i) * hw + pix]);
    const float* m = mab + img * kGfP2 * hw + pix;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float a0 = m[(c * 3 + 0) * hw], a1 = m[(c * 3 + 1) * hw], a2 = m[(c * 3 + 2) * hw], b = m[(9 + c) * hw];
      const float q = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(a0, I[0]), __fmul_rn(a1, I[1])), __fmul_rn(a2, I[2])), b);
      const float u = fminf(fmaxf(rintf(q), 0.0f), 255.0f);        // saturate_cast<uchar>(cvRound(q)), half to even
      out[(img * 3 + c) * hw + pix] = __fmul_rn(__fsub_rn(__fdiv_rn(u, 255.0f), 0.5f), 2.0f);
    }
  }
}

static int gf_grid(int64_t total) {
  const int64_t b = (total + 255) / 256;
  return (int)(b < 148 * 16 ? (b < 1 ? 1 : b) : 148 * 16);
}

}  // namespace ppst

extern "C" int64_t ppst_guided_smooth_workspace_bytes(int64_t n, int h, int w) {
  if (n <= 0 || h <= 0 || w <= 0) return 0;
  return 2 * n * (int64_t)ppst::kGfP1 * h * w * 4;
}

extern "C" int ppst_guided_smooth(float* out, const float* src, const float* guide, int64_t n, int h, int w, int radius,
                                  float eps, void* workspace, int64_t workspace_bytes, void* stream) {
  using namespace ppst;
  PPST_REQUIRE(out && src && guide && workspace && n > 0 && h > 0 && w > 0, PPST_ERR_ARG, "guided_smooth: bad args");
  PPST_REQUIRE(radius >= 1 && radius < h && radius < w, PPST_ERR_UNSUPPORTED,
               "guided_smooth: radius %d needs an image larger than the radius (%dx%d)", radius, h, w);
  PPST_REQUIRE((2 * radius + 1) <= 181, PPST_ERR_UNSUPPORTED, "guided_smooth: radius %d: int32 window sums overflow", radius);
  PPST_REQUIRE(workspace_bytes >= ppst_guided_smooth_workspace_bytes(n, h, w), PPST_ERR_ARG,
               "guided_smooth: workspace too small");
  PPST_REQUIRE((size_t)w * sizeof(double) <= 96 * 1024, PPST_ERR_UNSUPPORTED, "guided_smooth: rows of %d pixels", w);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t hw = (int64_t)h * w;
  int32_t* bufA = reinterpret_cast<int32_t*>(workspace);
  int32_t* bufB = bufA + n * kGfP1 * hw;
  const double inv_area = 1.0 / ((double)(2 * radius + 1) * (2 * radius + 1));
  const int seg = h >= 512 ? 64 : 32;
  {                                  // per device, so set on every call (a cheap host-side attribute write)
    int rc = cuda_check(cudaFuncSetAttribute(gf_box_h_kernel<float, double, float>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024), "guided attr");
    if (rc) return rc;
    rc = cuda_check(cudaFuncSetAttribute(gf_box_h_kernel<int32_t, int32_t, int32_t>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024), "guided attr");
    if (rc) return rc;
  }
  gf_products_kernel<<<gf_grid(n * hw), 256, 0, st>>>(bufA, src, guide, n, hw);
  PPST_LAUNCH_CHECK("gf_products_kernel");
  {
    const int64_t rows = n * kGfP1 * h;
    const int blocks = (int)(rows < 148 * 16 ? rows : 148 * 16);
    gf_box_h_kernel<int32_t, int32_t, int32_t><<<blocks, 128, (size_t)w * sizeof(int32_t), st>>>(bufB, bufA, rows, w, radius);
    PPST_LAUNCH_CHECK("gf_box_h_kernel<int>");
    const int64_t thr = n * kGfP1 * ((h + seg - 1) / seg) * w;
    gf_box_v_kernel<int32_t, int32_t, int32_t><<<gf_grid(thr), 256, 0, st>>>(bufA, bufB, n * kGfP1, h, w, radius, seg, 0.0);
    PPST_LAUNCH_CHECK("gf_box_v_kernel<int>");
  }
  float* ab = reinterpret_cast<float*>(bufB);                     // [n][12][hw] (fits in the 21-plane buffer)
  gf_solve_kernel<<<gf_grid(n * hw), 256, 0, st>>>(ab, bufA, n, hw, inv_area, eps);
  PPST_LAUNCH_CHECK("gf_solve_kernel");
  float* tmp = reinterpret_cast<float*>(bufA);
  {
    const int64_t rows = n * kGfP2 * h;
    const int blocks = (int)(rows < 148 * 16 ? rows : 148 * 16);
    gf_box_h_kernel<float, double, float><<<blocks, 128, (size_t)w * sizeof(double), st>>>(tmp, ab, rows, w, radius);
    PPST_LAUNCH_CHECK("gf_box_h_kernel<float>");
    const int64_t thr = n * kGfP2 * ((h + seg - 1) / seg) * w;
    gf_box_v_kernel<float, double, float><<<gf_grid(thr), 256, 0, st>>>(ab, tmp, n * kGfP2, h, w, radius, seg, inv_area);
    PPST_LAUNCH_CHECK("gf_box_v_kernel<float>");
  }
  gf_final_kernel<<<gf_grid(n * hw), 256, 0, st>>>(out, ab, guide, n, hw);
  PPST_LAUNCH_CHECK("gf_final_kernel");
  return 0;
}
