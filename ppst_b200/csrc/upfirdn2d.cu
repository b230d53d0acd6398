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

// ---- mbarrier / bulk-copy primitives (1-D TMA: cp.async.bulk, SASS UBLKCP)
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}

// ---- fast path: up = 1, down = 1, minor = 1, kernel <= 4 x 4 (modes 1 and 2 of the reference: the encoder /
//      discriminator blur).  One block = one plane x one strip of R output rows x the full width, so the input rows
//      the strip needs are ONE contiguous span of the plane.  One thread moves the 16-byte-aligned interior of that
//      span into shared memory with a single bulk copy (no LSU wavefronts, no registers, all bytes in flight at once,
//      DRAM sees a ~40 KB burst); the <= 2 ragged ends are copied by a few scalar loads.  Shared memory holds the raw
//      span (element type T, row pitch in_w); the zero padding is applied by predicates, only in threads whose window
//      touches the image border.  Each thread slides its window down RR rows of one column, lanes on consecutive
//      columns: conflict-free LDS, and 128-byte coalesced stores whatever the row alignment (2^n+1-wide rows are the
//      hot case).  KS = 3 / 4: exact square taps in registers; KS = 0: run-time kh, kw <= 4.
template <typename T, int KS>
__global__ void __launch_bounds__(256, 5) upfirdn2d_bulk_kernel(T* __restrict__ out, const T* __restrict__ in,
                                                                const float* __restrict__ kernel, UpfirdnParams p,
                                                                int R, int nstrips) {
  constexpr int RR = 8;                              // output rows per thread (vertical sliding window)
  constexpr int KM = KS ? KS : 4;
  constexpr int VEC = 16 / sizeof(T);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  T* sm = reinterpret_cast<T*>(smem_raw + 16);
  const int tid = threadIdx.x;
  const int strip = blockIdx.x % nstrips;
  const int64_t plane = blockIdx.x / nstrips;
  const int oy0 = strip * R;
  const int rows = min(R, p.out_h - oy0);
  const int kh = KS ? KS : p.kh, kw = KS ? KS : p.kw;
  // staged input rows [iyA, iyB)
  const int iyA = max(0, oy0 - p.py0), iyB = min(p.in_h, oy0 - p.py0 + rows + kh - 1);
  const int64_t n = iyB > iyA ? (int64_t)(iyB - iyA) * p.in_w : 0;
  const T* g = in + (plane * p.in_h + iyA) * (int64_t)p.in_w;      // first staged element
  const int mis = (int)((reinterpret_cast<uintptr_t>(g) / sizeof(T)) & (VEC - 1));
  // element g[i] lives at sm[mis + i]; the bulk copy covers i in [head, head + body)
  const int64_t head = n < ((VEC - mis) & (VEC - 1)) ? n : ((VEC - mis) & (VEC - 1));
  const int64_t body = (n - head) / VEC * VEC;
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_expect_tx(bar, (uint32_t)(body * sizeof(T)));
    if (body) bulk_g2s(sm + mis + head, g + head, (uint32_t)(body * sizeof(T)), bar);
  } else if (tid >= 32 && tid < 32 + 2 * VEC) {
    const int j = tid - 32;
    const int64_t i = j < VEC ? j : head + body + (j - VEC);
    if (j < VEC ? i < head : i < n) sm[mis + i] = g[i];
  }
  // flipped taps in registers
  float k[KM][KM];
#pragma unroll
  for (int ky = 0; ky < KM; ++ky)
#pragma unroll
    for (int kx = 0; kx < KM; ++kx)
      k[ky][kx] = (ky < kh && kx < kw) ? __ldg(kernel + (kh - 1 - ky) * kw + (kw - 1 - kx)) : 0.0f;
  __syncthreads();          // barrier initialised, ragged ends visible
  mbar_wait(bar, 0);        // bulk bytes landed

  const int cols = p.out_w;
  const int nitems = ((rows + RR - 1) / RR) * cols;
  T* dst = out + plane * (int64_t)p.out_h * p.out_w;
  // item = (row group rg, column c), walked 256 items at a time without a division per item
  int rg = tid / cols, c = tid - rg * cols;
  for (int idx = tid; idx < nitems; idx += 256) {
    const int oyb = oy0 + rg * RR;
    const int xb = c - p.px0, yb = oyb - p.py0;           // input coordinate of tap (0, 0) of the item's first output
    const T* sp = sm + mis + (yb - iyA) * p.in_w + xb;   // only dereferenced where the predicates hold
    float acc[RR];
#pragma unroll
    for (int o = 0; o < RR; ++o) acc[o] = 0.0f;
    const bool interior = xb >= 0 && xb + kw <= p.in_w && yb >= iyA && yb + RR + kh - 1 <= iyB;
    if (interior) {
#pragma unroll
      for (int r = 0; r < RR + KM - 1; ++r) {
        if (KS == 0 && r >= RR + kh - 1) break;
        float v[KM];
#pragma unroll
        for (int kx = 0; kx < KM; ++kx)
          if (KS != 0 || kx < kw) v[kx] = Cvt<T>::to_f(sp[r * p.in_w + kx]);
#pragma unroll
        for (int o = 0; o < RR; ++o) {
          const int ky = r - o;
          if (ky >= 0 && ky < KM && (KS != 0 || ky < kh)) {
#pragma unroll
            for (int kx = 0; kx < KM; ++kx)
              if (KS != 0 || kx < kw) acc[o] = fmaf(v[kx], k[ky][kx], acc[o]);
          }
        }
      }
    } else {
      bool cok[KM];
#pragma unroll
      for (int kx = 0; kx < KM; ++kx) cok[kx] = kx < kw && xb + kx >= 0 && xb + kx < p.in_w;
#pragma unroll
      for (int r = 0; r < RR + KM - 1; ++r) {
        const int y = yb + r;
        const bool rok = r < RR + kh - 1 && y >= iyA && y < iyB;
        float v[KM];
#pragma unroll
        for (int kx = 0; kx < KM; ++kx) v[kx] = (rok && cok[kx]) ? Cvt<T>::to_f(sp[r * p.in_w + kx]) : 0.0f;
#pragma unroll
        for (int o = 0; o < RR; ++o) {
          const int ky = r - o;
          if (ky >= 0 && ky < KM && (KS != 0 || ky < kh)) {
#pragma unroll
            for (int kx = 0; kx < KM; ++kx)
              if (KS != 0 || kx < kw) acc[o] = fmaf(v[kx], k[ky][kx], acc[o]);
          }
        }
      }
    }
    T* q = dst + (int64_t)oyb * p.out_w + c;
    const int nvalid = p.out_h - oyb;
#pragma unroll
    for (int o = 0; o < RR; ++o) {
      if (o < nvalid) *q = Cvt<T>::from_f(acc[o]);
      q += p.out_w;
    }
    c += 256;
    while (c >= cols) { c -= cols; ++rg; }
  }
}

// ---- the same strip kernel for 16-bit element types (bf16 / fp16).  The fp32 form above issues one 16-bit LDS, one
//      conversion and KS x KS FFMA per tap row and output: at half the bytes per element it is issue-bound at the same
//      element rate as fp32 (0.34-0.50 of HBM).  Here one thread owns TWO adjacent columns x RR rows: a tap row of the
//      window is KS + 1 elements = three aligned 32-bit shared-memory words (funnel-shifted by 16 bits when the row
//      starts on an odd element: 2^n+3-wide rows alternate), the two columns' accumulators are one float2 and every tap is
//      one packed FFMA2 (sm_100: fma.rn.f32x2), and the two results leave as one 32-bit store.  Per-column arithmetic and
//      its order are those of the scalar kernel, so the results are bit-identical to it.
constexpr int kRows16 = 8;
template <typename T> struct Pack16;
template <> struct Pack16<__nv_bfloat16> {
  __device__ __forceinline__ static float2 unpack(uint32_t w) {
    return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
  }
  __device__ __forceinline__ static uint32_t pack(float2 v) {
    __nv_bfloat162 h = __floats2bfloat162_rn(v.x, v.y);
    return *reinterpret_cast<uint32_t*>(&h);
  }
};
template <> struct Pack16<__half> {
  __device__ __forceinline__ static float2 unpack(uint32_t w) { return __half22float2(*reinterpret_cast<__half2*>(&w)); }
  __device__ __forceinline__ static uint32_t pack(float2 v) {
    __half2 h = __floats2half2_rn(v.x, v.y);
    return *reinterpret_cast<uint32_t*>(&h);
  }
};

// One column x RR rows with every tap predicated (window touches the image border / the strip's staged rows).
template <typename T, int KS, int RR>
__device__ __forceinline__ void upfirdn2d_border_item(T* __restrict__ q, const T* sp, const float (&k)[KS ? KS : 4][KS ? KS : 4],
                                                      int kh, int kw, int xb, int yb, int iyA, int iyB, int in_w, int out_w,
                                                      int nvalid) {
  constexpr int KM = KS ? KS : 4;
  float acc[RR];
#pragma unroll
  for (int o = 0; o < RR; ++o) acc[o] = 0.0f;
  bool cok[KM];
#pragma unroll
  for (int kx = 0; kx < KM; ++kx) cok[kx] = kx < kw && xb + kx >= 0 && xb + kx < in_w;
#pragma unroll
  for (int r = 0; r < RR + KM - 1; ++r) {
    const int y = yb + r;
    const bool rok = r < RR + kh - 1 && y >= iyA && y < iyB;
    float v[KM];
#pragma unroll
    for (int kx = 0; kx < KM; ++kx) v[kx] = (rok && cok[kx]) ? Cvt<T>::to_f(sp[r * in_w + kx]) : 0.0f;
#pragma unroll
    for (int o = 0; o < RR; ++o) {
      const int ky = r - o;
      if (ky >= 0 && ky < KM && (KS != 0 || ky < kh)) {
#pragma unroll
        for (int kx = 0; kx < KM; ++kx)
          if (KS != 0 || kx < kw) acc[o] = fmaf(v[kx], k[ky][kx], acc[o]);
      }
    }
  }
#pragma unroll
  for (int o = 0; o < RR; ++o) {
    if (o < nvalid) *q = Cvt<T>::from_f(acc[o]);
    q += out_w;
  }
}

template <typename T, int KS>
__global__ void __launch_bounds__(256, 4) upfirdn2d_bulk16_kernel(T* __restrict__ out, const T* __restrict__ in,
                                                                  const float* __restrict__ kernel, UpfirdnParams p,
                                                                  int R, int nstrips) {
  static_assert(sizeof(T) == 2, "16-bit element types only");
  constexpr int RR = kRows16;                        // output rows per thread (vertical sliding window)
  constexpr int KM = KS ? KS : 4;
  constexpr int VEC = 8;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  T* sm = reinterpret_cast<T*>(smem_raw + 16);
  const int tid = threadIdx.x;
  const int strip = blockIdx.x % nstrips;
  const int64_t plane = blockIdx.x / nstrips;
  const int oy0 = strip * R;
  const int rows = min(R, p.out_h - oy0);
  const int kh = KS ? KS : p.kh, kw = KS ? KS : p.kw;
  const int iyA = max(0, oy0 - p.py0), iyB = min(p.in_h, oy0 - p.py0 + rows + kh - 1);
  const int64_t n = iyB > iyA ? (int64_t)(iyB - iyA) * p.in_w : 0;
  const T* g = in + (plane * p.in_h + iyA) * (int64_t)p.in_w;
  const int mis = (int)((reinterpret_cast<uintptr_t>(g) / sizeof(T)) & (VEC - 1));
  const int64_t head = n < ((VEC - mis) & (VEC - 1)) ? n : ((VEC - mis) & (VEC - 1));
  const int64_t body = (n - head) / VEC * VEC;
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_expect_tx(bar, (uint32_t)(body * sizeof(T)));
    if (body) bulk_g2s(sm + mis + head, g + head, (uint32_t)(body * sizeof(T)), bar);
  } else if (tid >= 32 && tid < 32 + 2 * VEC) {
    const int j = tid - 32;
    const int64_t i = j < VEC ? j : head + body + (j - VEC);
    if (j < VEC ? i < head : i < n) sm[mis + i] = g[i];
  }
  float k[KM][KM];
#pragma unroll
  for (int ky = 0; ky < KM; ++ky)
#pragma unroll
    for (int kx = 0; kx < KM; ++kx)
      k[ky][kx] = (ky < kh && kx < kw) ? __ldg(kernel + (kh - 1 - ky) * kw + (kw - 1 - kx)) : 0.0f;
  __syncthreads();
  mbar_wait(bar, 0);

  // Columns whose window stays inside the image horizontally form the pair loop: [cl, cl + 2 * cols2).  The few columns
  // left of / right of it (zero padding, and the odd last column) are done one by one afterwards, so that no warp of the
  // pair loop runs both the packed and the predicated code and the items divide evenly over the 256 threads.
  const int cl = min(max(p.px0, 0), p.out_w);
  const int cr = max(cl, min(p.out_w, p.in_w - kw + p.px0 + 1));
  const int cols2 = (cr - cl) >> 1;                  // column pairs
  const int nrg = (rows + RR - 1) / RR;
  const int nitems = nrg * cols2;
  T* dst = out + plane * (int64_t)p.out_h * p.out_w;
  const bool rows_aligned = (reinterpret_cast<uintptr_t>(dst) & 3) == 0 && ((p.out_w | cl) & 1) == 0;
  int rg = cols2 ? tid / cols2 : 0, cp = tid - rg * cols2;
  for (int idx = tid; idx < nitems; idx += 256) {
    const int c = cl + cp * 2;
    const int oyb = oy0 + rg * RR;
    const int xb = c - p.px0, yb = oyb - p.py0;
    const int e = mis + (yb - iyA) * p.in_w + xb;    // element index of tap (0, 0) of the first output in sm
    const int nvalid = p.out_h - oyb;
    T* q = dst + (int64_t)oyb * p.out_w + c;
    const bool interior = yb >= iyA && yb + RR + kh - 1 <= iyB;    // per row group: (almost) warp-uniform
    if (interior) {
      float2 acc[RR];
#pragma unroll
      for (int o = 0; o < RR; ++o) acc[o] = make_float2(0.0f, 0.0f);
      // Row r of the window starts at element e + r * in_w: byte 2 * that, rounded down to a word, plus a 16-bit funnel
      // shift when the element is odd.  With in_w odd the parity alternates, so even and odd rows are two address
      // streams that both advance by 2 rows = 4 * in_w bytes (a whole number of words): one add per row.
      const uint32_t par0 = e & 1, par1 = (e + p.in_w) & 1;
      const uint32_t shv[2] = {par0 << 4, par1 << 4};
      const unsigned char* rowp[2] = {reinterpret_cast<const unsigned char*>(sm) + 2 * (e - (int)par0),
                                      reinterpret_cast<const unsigned char*>(sm) + 2 * (e + p.in_w - (int)par1)};
      const int pitch2 = 4 * p.in_w;
#pragma unroll
      for (int r = 0; r < RR + KM - 1; ++r) {
        if (KS == 0 && r >= RR + kh - 1) break;
        const uint32_t* wp = reinterpret_cast<const uint32_t*>(rowp[r & 1]);
        rowp[r & 1] += pitch2;
        const uint32_t sh = shv[r & 1];
        const uint32_t w0 = wp[0], w1 = wp[1], w2 = wp[2];
        const float2 a = Pack16<T>::unpack(__funnelshift_r(w0, w1, sh));    // elements 0, 1 of the row's window
        const float2 b = Pack16<T>::unpack(__funnelshift_r(w1, w2, sh));    // 2, 3
        float2 v[4];
        v[0] = a;
        v[1] = make_float2(a.y, b.x);
        v[2] = b;
        if (KM == 4) v[3] = make_float2(b.y, Pack16<T>::unpack(w2 >> sh).x);  // 3, 4
#pragma unroll
        for (int o = 0; o < RR; ++o) {
          const int ky = r - o;
          if (ky >= 0 && ky < KM && (KS != 0 || ky < kh)) {
#pragma unroll
            for (int kx = 0; kx < KM; ++kx)
              if (KS != 0 || kx < kw) acc[o] = __ffma2_rn(v[kx], make_float2(k[ky][kx], k[ky][kx]), acc[o]);
          }
        }
      }
      if (nvalid >= RR) {
        if (rows_aligned) {                          // every row starts on a 4-byte boundary: one store per pair
#pragma unroll
          for (int o = 0; o < RR; ++o) {
            *reinterpret_cast<uint32_t*>(q) = Pack16<T>::pack(acc[o]);
            q += p.out_w;
          }
        } else {                                     // odd widths (the 2^n+1 hot case): the alignment alternates
#pragma unroll
          for (int o = 0; o < RR; ++o) {
            const uint32_t w = Pack16<T>::pack(acc[o]);
            reinterpret_cast<unsigned short*>(q)[0] = (unsigned short)(w & 0xffffu);
            reinterpret_cast<unsigned short*>(q)[1] = (unsigned short)(w >> 16);
            q += p.out_w;
          }
        }
      } else {
#pragma unroll
        for (int o = 0; o < RR; ++o) {
          if (o < nvalid) {
            const uint32_t w = Pack16<T>::pack(acc[o]);
            reinterpret_cast<unsigned short*>(q)[0] = (unsigned short)(w & 0xffffu);
            reinterpret_cast<unsigned short*>(q)[1] = (unsigned short)(w >> 16);
          }
          q += p.out_w;
        }
      }
    } else {
      const T* sp = sm + e;                          // only dereferenced where the predicates hold
      upfirdn2d_border_item<T, KS, RR>(q, sp, k, kh, kw, xb, yb, iyA, iyB, p.in_w, p.out_w, nvalid);
      upfirdn2d_border_item<T, KS, RR>(q + 1, sp + 1, k, kh, kw, xb + 1, yb, iyA, iyB, p.in_w, p.out_w, nvalid);
    }
    cp += 256;
    while (cp >= cols2) { cp -= cols2; ++rg; }
  }
  // the columns outside the pair loop
  const int cend = cl + 2 * cols2;
  const int nleft = cl + (p.out_w - cend);
  for (int idx = tid; idx < nleft * nrg; idx += 256) {
    const int g2 = idx / nleft, j = idx - g2 * nleft;
    const int c = j < cl ? j : cend + (j - cl);
    const int oyb = oy0 + g2 * RR;
    const int xb = c - p.px0, yb = oyb - p.py0;
    upfirdn2d_border_item<T, KS, RR>(dst + (int64_t)oyb * p.out_w + c, sm + mis + (yb - iyA) * p.in_w + xb, k, kh, kw, xb, yb,
                                     iyA, iyB, p.in_w, p.out_w, p.out_h - oyb);
  }
}

// ---- register-staged variant of the up = down = 1 path for rows too long for the bulk-copy kernel above
//      (strip x column-chunk tiles, fp32 tile in shared memory with the zero padding materialised).
template <typename T, int KMAX>
__global__ void __launch_bounds__(256, 5) upfirdn2d_strip_kernel(T* __restrict__ out, const T* __restrict__ in,
                                                              const float* __restrict__ kernel, UpfirdnParams p,
                                                              int R, int CW, int nstrips, int nchunks) {
  constexpr int RR = 4;                              // output rows per thread (vertical sliding window)
  constexpr int VEC = 16 / sizeof(T);
  constexpr int UNR = 4;                             // 16-byte loads in flight per lane while staging a row
  extern __shared__ __align__(16) float sx[];        // (R + KMAX - 1) x P
  const int P = CW + KMAX - 1;
  const int tid = threadIdx.x, wrp = tid >> 5, lane = tid & 31;
  unsigned b = blockIdx.x;
  const int strip = b % nstrips; b /= nstrips;
  const int chunk = b % nchunks;
  const int64_t plane = b / nchunks;
  const int oy0 = strip * R, ox0 = chunk * CW;
  const int rows = min(R, p.out_h - oy0), cols = min(CW, p.out_w - ox0);
  // flipped taps, zero-extended to KMAX x KMAX, in registers
  float k[KMAX][KMAX];
#pragma unroll
  for (int ky = 0; ky < KMAX; ++ky)
#pragma unroll
    for (int kx = 0; kx < KMAX; ++kx)
      k[ky][kx] = (ky < p.kh && kx < p.kw) ? __ldg(kernel + (p.kh - 1 - ky) * p.kw + (p.kw - 1 - kx)) : 0.0f;

  // ---- stage: smem (r, s) <- input (oy0 - py0 + r, ox0 - px0 + s), zero outside the image / the needed window
  const T* src = in + plane * (int64_t)p.in_h * p.in_w;
  const int need_h = rows + p.kh - 1, need_w = cols + p.kw - 1;
  const int tile_h = R + KMAX - 1;
  const int s_lo = max(0, p.px0 - ox0);
  const int s_hi = min(need_w, p.in_w + p.px0 - ox0);
  for (int r = wrp; r < tile_h; r += 8) {
    float* row = sx + r * P;
    const int y = oy0 - p.py0 + r;
    if (r >= need_h || y < 0 || y >= p.in_h || s_hi <= s_lo) {
      for (int s = lane; s < P; s += 32) row[s] = 0.0f;
      continue;
    }
    for (int s = lane; s < s_lo; s += 32) row[s] = 0.0f;
    for (int s = s_hi + lane; s < P; s += 32) row[s] = 0.0f;
    const T* rp = src + (int64_t)y * p.in_w + (ox0 - p.px0);        // address of s = 0 (only s in [s_lo, s_hi) is read)
    const int mis = (int)((reinterpret_cast<uintptr_t>(rp + s_lo) & 15) / sizeof(T));
    const int sv0 = s_lo - mis;                                     // s of the first 16-byte aligned vector
    const int nvec = (s_hi - sv0 + VEC - 1) / VEC;
    for (int v0 = 0; v0 < nvec; v0 += 32 * UNR) {
      Vec16<T> q[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int s0 = sv0 + (v0 + u * 32 + lane) * VEC;
        if (s0 >= s_lo && s0 + VEC <= s_hi) {
          q[u] = ld_vec16_stream(rp + s0);
        } else {
#pragma unroll
          for (int i = 0; i < VEC; ++i)
            q[u].v[i] = (s0 + i >= s_lo && s0 + i < s_hi) ? rp[s0 + i] : Cvt<T>::from_f(0.0f);
        }
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int s0 = sv0 + (v0 + u * 32 + lane) * VEC;
#pragma unroll
        for (int i = 0; i < VEC; ++i)
          if (s0 + i >= s_lo && s0 + i < s_hi) row[s0 + i] = Cvt<T>::to_f(q[u].v[i]);
      }
    }
  }
  __syncthreads();

  // ---- compute: item = (row group of RR rows, column); consecutive threads on consecutive columns
  T* dst = out + plane * (int64_t)p.out_h * p.out_w + ox0;
  const int nitems = ((rows + RR - 1) / RR) * cols;
  for (int idx = tid; idx < nitems; idx += 256) {
    const int rg = idx / cols, c = idx - rg * cols;
    const float* sp = sx + (rg * RR) * P + c;
    float acc[RR];
#pragma unroll
    for (int o = 0; o < RR; ++o) acc[o] = 0.0f;
#pragma unroll
    for (int r = 0; r < RR + KMAX - 1; ++r) {
      float v[KMAX];
#pragma unroll
      for (int kx = 0; kx < KMAX; ++kx) v[kx] = sp[r * P + kx];
#pragma unroll
      for (int o = 0; o < RR; ++o) {
        const int ky = r - o;
        if (ky >= 0 && ky < KMAX && ky < p.kh) {
#pragma unroll
          for (int kx = 0; kx < KMAX; ++kx)
            if (kx < p.kw) acc[o] = fmaf(v[kx], k[ky][kx], acc[o]);
        }
      }
    }
    const int oy = oy0 + rg * RR;
#pragma unroll
    for (int o = 0; o < RR; ++o)
      if (oy + o < p.out_h) dst[(int64_t)(oy + o) * p.out_w + c] = Cvt<T>::from_f(acc[o]);
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
    {
      // bulk-copy kernel: R = the largest multiple of 8 rows whose staged span fits ~44 KB (5 blocks per SM)
      const int ks = (p.kh == p.kw && (p.kh == 3 || p.kh == 4)) ? p.kh : 0;
      const size_t row_bytes = sizeof(T) * (size_t)p.in_w;
      int rows = (int)(45000 / row_bytes) - (p.kh - 1);
      rows = rows / 8 * 8;
      rows = rows < 8 ? 8 : (rows > 128 ? 128 : rows);
      rows = (ceil_div(p.out_h, ceil_div(p.out_h, rows)) + 7) / 8 * 8;   // balance the strips of a plane
      const size_t smem = 16 + 32 + row_bytes * (size_t)(rows + p.kh - 1);
      const int nstrips = ceil_div(p.out_h, rows);
      const int64_t blocks = p.major * nstrips;
      if (smem <= 100 * 1024 && blocks < (1LL << 31) && (int64_t)p.in_w * (rows + 8) < (1 << 24)) {
        if constexpr (sizeof(T) == 2) {
          if (p.in_w >= 8) {
            auto kern16 = ks == 3 ? upfirdn2d_bulk16_kernel<T, 3> : ks == 4 ? upfirdn2d_bulk16_kernel<T, 4>
                                                                          : upfirdn2d_bulk16_kernel<T, 0>;
            int rc = cuda_check(cudaFuncSetAttribute(kern16, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024),
                                "upfirdn2d_bulk16_kernel attribute");
            if (rc) return rc;
            kern16<<<(unsigned)blocks, 256, smem, st>>>(y, x, k, p, rows, nstrips);
            PPST_LAUNCH_CHECK("upfirdn2d_bulk16_kernel");
            return 0;
          }
        }
        auto kern = ks == 3 ? upfirdn2d_bulk_kernel<T, 3> : ks == 4 ? upfirdn2d_bulk_kernel<T, 4>
                                                                  : upfirdn2d_bulk_kernel<T, 0>;
        int rc = cuda_check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024),
                            "upfirdn2d_bulk_kernel attribute");
        if (rc) return rc;
        kern<<<(unsigned)blocks, 256, smem, st>>>(y, x, k, p, rows, nstrips);
        PPST_LAUNCH_CHECK("upfirdn2d_bulk_kernel");
        return 0;
      }
    }
    const int kmax = (p.kh <= 3 && p.kw <= 3) ? 3 : 4;
    const int cw = p.out_w <= 1280 ? p.out_w : 1024;               // full-width strips unless the rows are very long
    const int pitch = cw + kmax - 1;
    int rows = (int)(40960 / (sizeof(float) * pitch)) - (kmax - 1);  // ~40 KB of staged input per block
    rows = rows / 4 * 4;
    rows = rows < 8 ? 8 : (rows > 32 ? 32 : rows);
    if (rows > (p.out_h + 3) / 4 * 4) rows = (p.out_h + 3) / 4 * 4;
    const int nstrips = ceil_div(p.out_h, rows), nchunks = ceil_div(p.out_w, cw);
    const int64_t blocks = p.major * nstrips * nchunks;
    const size_t smem = sizeof(float) * (size_t)(rows + kmax - 1) * pitch;
    if (blocks < (1LL << 31)) {
      if (kmax == 3) {
        cudaFuncSetAttribute(upfirdn2d_strip_kernel<T, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        upfirdn2d_strip_kernel<T, 3><<<(unsigned)blocks, 256, smem, st>>>(y, x, k, p, rows, cw, nstrips, nchunks);
      } else {
        cudaFuncSetAttribute(upfirdn2d_strip_kernel<T, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        upfirdn2d_strip_kernel<T, 4><<<(unsigned)blocks, 256, smem, st>>>(y, x, k, p, rows, cw, nstrips, nchunks);
      }
      PPST_LAUNCH_CHECK("upfirdn2d_strip_kernel");
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


// ---------------------------------------------------------------------------------------------------------------------
// Blur on NHWC tensors for the training path (Blur = upfirdn2d with up = down = 1, stylegan2_layers.py:142-164, inside
// ConvLayer(downsample=True)): out[n, oy, ox, :] = sum_{ky,kx} in[n, oy + ky - py0, ox + kx - px0, :] * f[ky][kx], zero
// outside, f = the FLIPPED kernel (flip = 1: true convolution, what upfirdn2d computes) or the kernel as given (flip = 0:
// the adjoint used by the backward pass, UpFirDn2dBackward, upfirdn2d.py:24-90).
namespace ppst {
// One thread = kBlurRows vertically adjacent output pixels x one 16-byte channel vector: the (kBlurRows + kh - 1) x kw input
// vectors it needs are loaded once and shared by the rows (7 loads per output instead of 16 for the 4x4 kernel).
constexpr int kBlurRows = 4;
template <typename T, int KH, int KW>
__global__ void __launch_bounds__(256, 4) blur_nhwc_kernel(T* __restrict__ out, const T* __restrict__ in,
                                                        const float* __restrict__ k, int64_t n, int in_h, int in_w, int c,
                                                        int py0, int px0, int out_h, int out_w, int flip) {
  constexpr int VEC = Vec16<T>::N;
  __shared__ float sk[KH * KW];
  if (threadIdx.x < KH * KW) {
    const int ky = threadIdx.x / KW, kx = threadIdx.x - ky * KW;
    sk[threadIdx.x] = flip ? k[(KH - 1 - ky) * KW + (KW - 1 - kx)] : k[threadIdx.x];
  }
  __syncthreads();
  // grid = (chunks of 256 (column, channel vector) items of a row strip, strips, images): 32-bit index math only
  const int cv = c / VEC;
  const int item = blockIdx.x * 256 + threadIdx.x;
  if (item >= out_w * cv) return;
  const int ox = item / cv, v = item - ox * cv;
  const int oy0 = blockIdx.y * kBlurRows;
  const int64_t ni = blockIdx.z;
  float acc[kBlurRows][VEC];
#pragma unroll
  for (int j = 0; j < kBlurRows; ++j)
#pragma unroll
    for (int e = 0; e < VEC; ++e) acc[j][e] = 0.0f;
  const T* base = in + ni * in_h * in_w * (int64_t)c + (int64_t)v * VEC;
  // branch-free taps: out-of-range rows / columns load a clamped (valid) address and get weight 0, so the KW loads of a
  // row -- and the next row's -- are all in flight together
  int cx[KW];
  float vx[KW];
#pragma unroll
  for (int kx = 0; kx < KW; ++kx) {
    const int ix = ox + kx - px0;
    vx[kx] = (ix >= 0 && ix < in_w) ? 1.0f : 0.0f;
    cx[kx] = min(max(ix, 0), in_w - 1);
  }
#pragma unroll
  for (int rr = 0; rr < kBlurRows + KH - 1; ++rr) {
    const int iy = oy0 + rr - py0;
    const float vy = (iy >= 0 && iy < in_h) ? 1.0f : 0.0f;
    const T* row = base + (int64_t)min(max(iy, 0), in_h - 1) * in_w * c;
    Vec16<T> xv[KW];
#pragma unroll
    for (int kx = 0; kx < KW; ++kx) xv[kx] = ld_vec16(row + (int64_t)cx[kx] * c);
#pragma unroll
    for (int kx = 0; kx < KW; ++kx) {
      const bool valid = (vy * vx[kx]) != 0.0f;
      if constexpr (sizeof(T) == 2) {
        // adjacent channels are the two halves of one 32-bit word: unpack to a float2 and use packed FFMA2 with the tap
        // broadcast -- half the FMA issue slots of the scalar form, same per-channel arithmetic and order
        const uint32_t* wv = reinterpret_cast<const uint32_t*>(&xv[kx].raw);
#pragma unroll
        for (int e2 = 0; e2 < VEC / 2; ++e2) {
          const float2 xf = Pack16<T>::unpack(valid ? wv[e2] : 0u);
#pragma unroll
          for (int j = 0; j < kBlurRows; ++j) {
            const int ky = rr - j;
            if (ky < 0 || ky >= KH) continue;
            const float w = sk[ky * KW + kx];
            const float2 a = __ffma2_rn(make_float2(w, w), xf, make_float2(acc[j][2 * e2], acc[j][2 * e2 + 1]));
            acc[j][2 * e2] = a.x;
            acc[j][2 * e2 + 1] = a.y;
          }
        }
      } else {
        float xf[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) xf[e] = valid ? Cvt<T>::to_f(xv[kx].v[e]) : 0.0f;
#pragma unroll
        for (int j = 0; j < kBlurRows; ++j) {
          const int ky = rr - j;                     // input row rr is tap ky of output row j (compile-time after unrolling)
          if (ky < 0 || ky >= KH) continue;
          const float w = sk[ky * KW + kx];
#pragma unroll
          for (int e = 0; e < VEC; ++e) acc[j][e] = fmaf(w, xf[e], acc[j][e]);
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < kBlurRows; ++j) {
    if (oy0 + j >= out_h) break;
    Vec16<T> o;
#pragma unroll
    for (int e = 0; e < VEC; ++e) o.v[e] = Cvt<T>::from_f(acc[j][e]);
    st_vec16(out + ((ni * out_h + oy0 + j) * out_w + ox) * (int64_t)c + (int64_t)v * VEC, o);
  }
}
}  // namespace ppst

extern "C" int ppst_blur_nhwc(void* out, const void* in, const float* kernel, int64_t n, int in_h, int in_w, int c, int kh,
                              int kw, int py0, int py1, int px0, int px1, int flip, int dtype, void* stream) {
  using namespace ppst;
  PPST_REQUIRE(out && in && kernel && n > 0 && in_h > 0 && in_w > 0 && c > 0, PPST_ERR_ARG, "blur_nhwc: bad args");
  PPST_REQUIRE(kh >= 1 && kw >= 1 && kh * kw <= 16, PPST_ERR_UNSUPPORTED, "blur_nhwc: kernel %dx%d (at most 16 taps)", kh, kw);
  PPST_REQUIRE(dtype == PPST_BF16 || dtype == PPST_F32, PPST_ERR_DTYPE, "blur_nhwc: dtype %d", dtype);
  const int vec = dtype == PPST_BF16 ? 8 : 4;
  PPST_REQUIRE(c % vec == 0, PPST_ERR_UNSUPPORTED, "blur_nhwc: c=%d must be a multiple of %d", c, vec);
  const int out_h = in_h + py0 + py1 - kh + 1, out_w = in_w + px0 + px1 - kw + 1;
  PPST_REQUIRE(out_h > 0 && out_w > 0, PPST_ERR_ARG, "blur_nhwc: empty output");
  PPST_REQUIRE(kh == kw && kh >= 1 && kh <= 4, PPST_ERR_UNSUPPORTED, "blur_nhwc: kernel %dx%d (square, 1..4 taps per axis)", kh, kw);
  const int strips = (out_h + kBlurRows - 1) / kBlurRows;
  PPST_REQUIRE(n <= 65535 && strips <= 65535 && (int64_t)out_w * (c / vec) < (1LL << 31), PPST_ERR_UNSUPPORTED,
               "blur_nhwc: n=%lld / %d row strips exceed the grid limits", (long long)n, strips);
  cudaStream_t st = (cudaStream_t)stream;
  const dim3 blocks((unsigned)(((int64_t)out_w * (c / vec) + 255) / 256), (unsigned)strips, (unsigned)n);
#define PPST_BLUR(T, K) \
  blur_nhwc_kernel<T, K, K><<<blocks, 256, 0, st>>>((T*)out, (const T*)in, kernel, n, in_h, in_w, c, py0, px0, out_h, out_w, flip)
  if (dtype == PPST_BF16) {
    if (kh == 4) PPST_BLUR(__nv_bfloat16, 4); else if (kh == 3) PPST_BLUR(__nv_bfloat16, 3);
    else if (kh == 2) PPST_BLUR(__nv_bfloat16, 2); else PPST_BLUR(__nv_bfloat16, 1);
  } else {
    if (kh == 4) PPST_BLUR(float, 4); else if (kh == 3) PPST_BLUR(float, 3);
    else if (kh == 2) PPST_BLUR(float, 2); else PPST_BLUR(float, 1);
  }
#undef PPST_BLUR
  PPST_LAUNCH_CHECK("blur_nhwc_kernel");
  return 0;
}
