// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a (bf16 in, fp32 accumulate, bf16 out).
//
// GEMM view (per phase): M = output pixels, N = cout, K = taps * cin.
//   * M tile  = 128 output pixels = a TH x TW spatial patch of ONE image.  For tap (ky,kx) the A operand is that
//     patch shifted by the tap offset: ONE 4-D TMA box {BKC channels, TW, TH, 1} over the NHWC tensor.  Out-of-bounds
//     coordinates are zero-filled by the TMA unit, which IS the zero padding of the convolution; stride-2 convs use
//     the tensor map's element strides.  No im2col buffer ever exists.
//   * The box lands in shared memory as 128 rows (pixels) x BKC*2 bytes, 128B/64B-swizzled = the canonical K-major
//     UMMA operand layout, so tcgen05.mma consumes it directly.  B = packed weights [phase*cout][taps*cin], K-major.
//   * Accumulators live in tensor memory (2 stages x BN columns): the epilogue of tile i overlaps the MMAs of
//     tile i+1.  Epilogue: tcgen05.ld -> + bias + noise -> leaky-ReLU * gain -> bf16 -> 16-byte global stores, plus
//     per-channel partial sum / sum-of-squares of the stored values (instance-norm statistics for free).
//   * Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane), warp 2 = TMEM allocator,
//     warps 4-7 and 8-11 = two epilogue groups (TMEM lane quadrant = warp % 4), one per accumulator stage, so a tile's
//     epilogue may take up to two tiles' worth of MMA time before it stalls the tensor pipe.
//     Persistent grid: one CTA per SM, static tile striding.
#include <cuda.h>

#include <mutex>
#include <unordered_map>
#include <string>
#include <cstring>
#include <cstdlib>

#include "common.cuh"

namespace ppst {

int validate_conv_desc(const ppst_conv_desc& d, const char* who);

namespace tc {

constexpr int kThreads = 384;           // 4 control warps + 2 epilogue groups of 4 warps
constexpr int kEpiWarp0 = 4;            // first epilogue warp
constexpr int kBiasSmem = 1024;         // bias values staged in shared memory (larger cout: read from global);
                                        // one copy per epilogue group (a per-sample bias is reloaded per image)
constexpr int BM = 128;

struct Params {
  // geometry
  int n, out_h, out_w, cout, cin;
  int kh, kw, stride;
  int off_y, off_x, phases_y, phases_x, phase_off_dy, phase_off_dx;
  int full_h, full_w, cout_stride, c_off;
  int th, tw, tiles_y, tiles_x;         // spatial tile and tile counts per image
  int n_tiles_n;                        // cout / BN
  int num_tiles;                        // total work items
  int kblocks_per_tap;                  // cin / BKC
  int has_act;
  float slope, gain, noise_w;
  int stats_partials;                   // partial slots per image = tiles_y*tiles_x*phases
  int w_img_rows;                       // weight rows per image (0: one weight matrix for the whole batch)
  int split_seg_kb;                     // 3-term split GEMM: k-blocks per [hi]/[lo] segment (0: plain)
  int out_pad;                          // border (pixels) of the output buffer: stores go to its interior
  int out_pad_trim;                     // the bottom / right border is out_pad - out_pad_trim pixels
  int bias_img_stride;                  // > 0: image i adds bias[i * bias_img_stride + c] (per-sample bias)
  int stages, kps;                      // pipeline depth; k-blocks per pipeline stage (one barrier round trip each)
  // halo kernel only
  int tiles_ys;                         // super-tile rows per image = ceil(tiles_y / MT)
  int a_rows, a_slot_bytes;             // rows of one A box (MT*th + kh - 1) and its 1024-aligned slot size
  int a_box_bytes;                      // bytes of one A box; row stride 2: a slot holds the even-row box, then the odd-row box
  int na, nb, bps;                      // A slots, B slots, taps per B slot (1 or kh)
};

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(addr), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n"
      ".reg .b32 rx;\n"
      ".reg .pred px;\n"
      "elect.sync rx|px, 0xffffffff;\n"
      "selp.b32 %0, 1, 0, px;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// D[tmem] (+)= A[smem] * B[smem], bf16 x bf16 -> f32
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ---- cta_group::2 (CTA pair) variants.  kPeerMask clears the bit that distinguishes the two CTAs of a pair in the
// shared::cluster window, so the address names the EVEN CTA's copy of a barrier from either CTA.
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                                int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & kPeerMask), "r"(c0), "r"(c1), "r"(c2),
        "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & kPeerMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A[smem of each CTA: its 128 rows] * B[smem: each CTA holds half of the N rows]
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this offset in BOTH CTAs of the pair once all previously issued MMAs have completed
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}
// arrive on the EVEN CTA's copy of a barrier (from either CTA of the pair)
__device__ __forceinline__ void mbar_arrive_even_cta(uint64_t* bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerMask) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major shared-memory matrix descriptor (sm_100 "version 1"): rows of BKC*2 bytes, 8-row swizzle atoms.
template <int BKC>
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  constexpr uint64_t layout = (BKC == 64) ? 2ull /*SWIZZLE_128B*/ : 4ull /*SWIZZLE_64B*/;
  constexpr uint64_t sbo = (uint64_t)(8 * BKC * 2) >> 4;      // byte distance between 8-row groups
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);                    // start address, 14 bits
  d |= (uint64_t)1 << 16;                                     // leading byte offset (unused for swizzled K-major)
  d |= sbo << 32;                                             // stride byte offset
  d |= (uint64_t)1 << 46;                                     // descriptor version
  d |= layout << 61;
  return d;
}

template <int BN>
__device__ __forceinline__ uint32_t make_idesc() {
  // kind::f16: D=f32 (1<<4), A=bf16 (1<<7), B=bf16 (1<<10), both K-major, N>>3 at bit 17, M>>4 at bit 24
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// ---------------------------------------------------------------- bf16 epilogue shared by both kernels
constexpr int kEpiStageBytes = 2048;     // per epilogue warp: 32 rows x 64 bytes of packed bf16 (one 32-column chunk)

__device__ __forceinline__ void lds128(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

// One warp's share of the epilogue of a 128 x BN accumulator sub-tile: TMEM lanes [32q, 32q+32) = pixel rows, BN
// columns = output channels.  Per 32-column chunk:
//   tcgen05.ld -> (+ bias + noise, leaky-ReLU, gain) -> bf16 -> the warp's 2 KB staging buffer (16-byte pieces XOR-
//   swizzled by row pair so that both the row-wise writes and the column-wise reads below are bank-conflict free)
//   -> per-channel sum / sum of squares of the STORED values by walking the staged columns (lane = channel pair x row
//      parity: 16 LDS.32 + FMAs per lane instead of a 31-step shuffle transpose per statistic)
//   -> global stores with 4 lanes per pixel: every store instruction writes 8 pixels x 64 contiguous bytes (whole
//      32-byte sectors) instead of 32 scattered 16-byte pieces.
// soff[i] / bit i of svalid describe row 8i + lane/4 of this warp's 32 rows (element offset of the pixel's channel
// slice + 8 * (lane % 4), and whether that pixel exists); nz / valid describe row `lane`.
template <int BN>
__device__ __forceinline__ void epilogue_rows_bf16(uint32_t taddr, uint32_t bias_u32, uint32_t stage_u32, float* st_sum,
                                                   float* st_sq, __nv_bfloat16* __restrict__ out,
                                                   const int64_t (&soff)[4], uint32_t svalid, float nz, bool valid,
                                                   float gain, float slope_eff, int lane) {
  const uint32_t my_row = stage_u32 + (uint32_t)lane * 64u;
  const uint32_t wsw = (uint32_t)(lane >> 1) & 3u;                 // swizzle of the row this lane writes
  const int rpar = lane >> 4, cp = lane & 15;                     // statistics: row parity, channel pair
  // software pipeline over the 32-column chunks: the TMEM load of chunk c0 + 32 is in flight while chunk c0 is converted,
  // staged and stored (tcgen05.wait::ld waits for every outstanding load, so the next one is issued right after the wait)
  uint32_t rn[32];
  tmem_ld32(taddr, rn);
#pragma unroll 1
  for (int c0 = 0; c0 < BN; c0 += 32) {
    uint32_t r[32];
    uint32_t bb[32];
#pragma unroll
    for (int j = 0; j < 8; ++j) lds128(bias_u32 + (uint32_t)(c0 + 4 * j) * 4u, *reinterpret_cast<uint32_t(*)[4]>(&bb[4 * j]));
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) r[j] = rn[j];
    if (c0 + 32 < BN) tmem_ld32(taddr + (uint32_t)(c0 + 32), rn);
    uint32_t pk[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float v[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float a = (__uint_as_float(r[2 * j + h]) + __uint_as_float(bb[2 * j + h])) + nz;
        v[h] = fmaxf(a, a * slope_eff) * gain;                     // leaky-ReLU for slope_eff in (0, 1]; 1 = identity
      }
      pk[j] = valid ? pack_bf16(v[0], v[1]) : 0u;                  // rows outside the image count as 0 in the statistics
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      sts128(my_row + (((uint32_t)j ^ wsw) << 4), pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
    __syncwarp();
    if (st_sum != nullptr) {
      float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const uint32_t word = ((((uint32_t)cp >> 2) ^ (uint32_t)(i & 3)) << 2) | ((uint32_t)cp & 3u);
        const uint32_t w = lds32(stage_u32 + (uint32_t)(2 * i + rpar) * 64u + word * 4u);
        const float fx = __uint_as_float(w << 16), fy = __uint_as_float(w & 0xffff0000u);
        s1a += fx; s1b += fy;
        s2a = fmaf(fx, fx, s2a); s2b = fmaf(fy, fy, s2b);
      }
      s1a += __shfl_xor_sync(0xffffffffu, s1a, 16);
      s1b += __shfl_xor_sync(0xffffffffu, s1b, 16);
      s2a += __shfl_xor_sync(0xffffffffu, s2a, 16);
      s2b += __shfl_xor_sync(0xffffffffu, s2b, 16);
      if (lane < 16) {
        *reinterpret_cast<float2*>(st_sum + c0 + 2 * cp) = make_float2(s1a, s1b);
        *reinterpret_cast<float2*>(st_sq + c0 + 2 * cp) = make_float2(s2a, s2b);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int rr = 8 * i + (lane >> 2);
      uint32_t q4[4];
      lds128(stage_u32 + (uint32_t)rr * 64u + ((((uint32_t)lane & 3u) ^ ((uint32_t)(rr >> 1) & 3u)) << 4), q4);
      if ((svalid >> i) & 1u)
        *reinterpret_cast<uint4*>(out + soff[i] + c0) = make_uint4(q4[0], q4[1], q4[2], q4[3]);
    }
    __syncwarp();
  }
}

// Per-sample bias (bias_img_stride > 0): the 128 threads of an epilogue group swap their bias copy when the image of
// their tile changes (tiles are image-major, so once per image and CTA).  bar_id = 1 + group.
__device__ __forceinline__ void reload_group_bias(float* s_bias_g, const float* __restrict__ bias, int img, int stride,
                                                  int cout, int group, int tid_in_group) {
  if (group == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
  else asm volatile("bar.sync 2, 128;" ::: "memory");
  for (int i = tid_in_group; i < cout; i += 128) s_bias_g[i] = bias[(int64_t)img * stride + i];
  if (group == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
  else asm volatile("bar.sync 2, 128;" ::: "memory");
}

constexpr int kMaxStages = 8;
template <int BN, int BKC>
struct Smem {
  static constexpr int kABytes = BM * BKC * 2;
  static constexpr int kBBytes = BN * BKC * 2;
  static constexpr int kKblockBytes = kABytes + kBBytes;     // one k-block: A tile then B tile (both 1024-aligned)
  static constexpr int kStatsBytes = 2 /*buf*/ * 2 /*sum,sq*/ * 4 /*warps*/ * BN * 4;
  static constexpr int kBiasBytes = 2 * kBiasSmem * 4;
  static constexpr int kEpiBytes = 8 * kEpiStageBytes;        // per-warp staging of the bf16 epilogue
  static constexpr int kFixedBytes = kStatsBytes + kBiasBytes + kEpiBytes + (2 * kMaxStages + 4) * 8 + 16 + 1024 /*align slack*/;
  static constexpr int kMaxDynamic = 232448 - 1024 /*static*/;
};

template <int BN, int BKC, bool OUT_F32>
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               void* __restrict__ out_raw, const float* __restrict__ bias, const float* __restrict__ noise,
               float* __restrict__ stats_sum, float* __restrict__ stats_sq, const Params p) {
  using S = Smem<BN, BKC>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int STAGES = p.stages;
  const int stage_bytes = p.kps * S::kKblockBytes;
  float* s_stats = reinterpret_cast<float*>(smem + STAGES * stage_bytes);
  float* s_bias = reinterpret_cast<float*>(smem + STAGES * stage_bytes + S::kStatsBytes);
  const bool bias_in_smem = (p.cout <= kBiasSmem);        // always true for bf16 output (checked on the host)
  if (bias_in_smem)
    for (int i = threadIdx.x; i < p.cout; i += kThreads) s_bias[i] = s_bias[kBiasSmem + i] = bias ? bias[i] : 0.0f;
  uint8_t* s_epi = smem + STAGES * stage_bytes + S::kStatsBytes + S::kBiasBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_epi + S::kEpiBytes);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tmem_full = empty_bar + kMaxStages;  // [2]
  uint64_t* tmem_empty = tmem_full + 2;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 4);              // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int taps = p.kh * p.kw;
  // split mode (1x1 only): logical K = [hi|hi|lo] x [hi|lo|hi] over physical [hi|lo] operands
  const int kblocks = p.split_seg_kb ? 3 * p.split_seg_kb : taps * p.kblocks_per_tap;
  const int n_phases = p.phases_y * p.phases_x;

  // work item -> (n-tile, phase, spatial tile, image); n-tile fastest so that neighbouring CTAs share A in L2
  auto decode = [&](int tile, int& nt, int& ph, int& tx, int& ty, int& img) {
    nt = tile % p.n_tiles_n; tile /= p.n_tiles_n;
    ph = tile % n_phases; tile /= n_phases;
    tx = tile % p.tiles_x; tile /= p.tiles_x;
    ty = tile % p.tiles_y; img = tile / p.tiles_y;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        int nt, ph, tx, ty, img;
        decode(tile, nt, ph, tx, ty, img);
        const int py = ph / p.phases_x, px = ph % p.phases_x;
        const int iy0 = ty * p.th * p.stride + p.off_y + py * p.phase_off_dy;
        const int ix0 = tx * p.tw * p.stride + p.off_x + px * p.phase_off_dx;
        const int brow = img * p.w_img_rows + ph * p.cout + nt * BN;
        int tap = 0, cb = 0, ky = 0, kx = 0;                 // incremental (tap, channel block) decode: no div/mod
        for (int kb0 = 0; kb0 < kblocks; kb0 += p.kps) {
          const int nk = min(p.kps, kblocks - kb0);
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sbase = smem + stage * stage_bytes;
          mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)(nk * S::kKblockBytes));
          for (int j = 0; j < nk; ++j) {
            int ca, cbk;
            if (p.split_seg_kb) {
              const int kb = kb0 + j;
              const int seg = kb / p.split_seg_kb, r = kb - seg * p.split_seg_kb;
              ca = ((seg == 2 ? p.split_seg_kb : 0) + r) * BKC;        // A: hi, hi, lo
              cbk = ((seg == 1 ? p.split_seg_kb : 0) + r) * BKC;       // B: hi, lo, hi
            } else {
              ca = cb * BKC;
              cbk = tap * p.cin + ca;
            }
            uint8_t* sa = sbase + j * S::kKblockBytes;
            tma_load_4d(sa, &map_a, &full_bar[stage], ca, ix0 + kx, iy0 + ky, img);
            tma_load_2d(sa + S::kABytes, &map_b, &full_bar[stage], cbk, brow);
            if (!p.split_seg_kb && ++cb == p.kblocks_per_tap) {
              cb = 0; ++tap;
              if (++kx == p.kw) { kx = 0; ++ky; }
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      const uint32_t idesc = make_idesc<BN>();
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[as], aphase ^ 1);          // epilogue has drained this accumulator stage
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
        for (int kb0 = 0; kb0 < kblocks; kb0 += p.kps) {
          const int nk = min(p.kps, kblocks - kb0);
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sbase = smem_u32(smem + stage * stage_bytes);
          for (int j = 0; j < nk; ++j) {
            const uint32_t sa = sbase + (uint32_t)(j * S::kKblockBytes);
            const uint64_t da = make_smem_desc<BKC>(sa);
            const uint64_t db = make_smem_desc<BKC>(sa + S::kABytes);
#pragma unroll
            for (int k = 0; k < BKC / 16; ++k) {
              // advance 16 elements (32 bytes) along K inside the swizzle atom: +2 in the 16-byte address field
              umma_bf16(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, ((kb0 + j) | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit(&empty_bar[stage]);                 // frees the smem slot once these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[as]);                      // accumulator ready for the epilogue
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ===================== epilogue =====================
    const int q = warp & 3;                              // TMEM lane quadrant
    const int group = (warp - kEpiWarp0) >> 2;           // epilogue group == accumulator stage it serves
    const int row = q * 32 + lane;                       // row of the 128-pixel tile
    const int ety = row / p.tw, etx = row - ety * p.tw;
    float* s_bias_g = s_bias + group * kBiasSmem;
    const int fh_pad = p.full_h + 2 * p.out_pad - p.out_pad_trim, fw_pad = p.full_w + 2 * p.out_pad - p.out_pad_trim;
    int it = 0, cur_img = -1;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      if ((it & 1) != group) continue;
      int nt, ph, tx, ty, img;
      decode(tile, nt, ph, tx, ty, img);
      if (p.bias_img_stride > 0 && img != cur_img) {
        reload_group_bias(s_bias_g, bias, img, p.bias_img_stride, p.cout, group, (threadIdx.x - kEpiWarp0 * 32) & 127);
        cur_img = img;
      }
      const int py = ph / p.phases_x, px = ph % p.phases_x;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int oy = ty * p.th + ety, ox = tx * p.tw + etx;
      const bool valid = (oy < p.out_h) && (ox < p.out_w);
      const int64_t pix = ((int64_t)img * p.full_h + (oy * p.phases_y + py)) * p.full_w + (ox * p.phases_x + px);
      const float nz = (noise != nullptr && valid) ? p.noise_w * noise[pix] : 0.0f;
      const int64_t opix = ((int64_t)img * fh_pad + (oy * p.phases_y + py) + p.out_pad) * fw_pad + (ox * p.phases_x + px) + p.out_pad;
      const int64_t obase = opix * p.cout_stride + p.c_off + nt * BN;
      float* st_sum = s_stats + (size_t)(as * 2 + 0) * 4 * BN + q * BN;
      float* st_sq = s_stats + (size_t)(as * 2 + 1) * 4 * BN + q * BN;

      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN);
      if constexpr (OUT_F32) {
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(taddr + (uint32_t)c0, r);
          tmem_ld_wait();
          if (valid) {
            float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(out_raw) + obase + c0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float v[4];
#pragma unroll
              for (int h = 0; h < 4; ++h) {
                const int c = nt * BN + c0 + 4 * j + h;
                const float bj = bias ? (bias_in_smem ? s_bias_g[c] : __ldg(bias + c)) : 0.0f;
                float a = __uint_as_float(r[4 * j + h]) + bj + nz;
                if (p.has_act) a = (a > 0.0f) ? a : a * p.slope;
                v[h] = a * p.gain;
              }
              dst[j] = make_float4(v[0], v[1], v[2], v[3]);
            }
          }
        }
      } else {
        // store-side view of this warp's rows: lane -> rows 8i + lane/4, 16-byte piece lane%4
        int64_t soff[4];
        uint32_t svalid = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int srow = q * 32 + 8 * i + (lane >> 2);
          const int sy = srow / p.tw, sx = srow - sy * p.tw;
          const int soy = ty * p.th + sy, sox = tx * p.tw + sx;
          if (soy < p.out_h && sox < p.out_w) svalid |= 1u << i;
          const int64_t spix = ((int64_t)img * fh_pad + (soy * p.phases_y + py) + p.out_pad) * fw_pad + (sox * p.phases_x + px) + p.out_pad;
          soff[i] = spix * p.cout_stride + p.c_off + nt * BN + 8 * (lane & 3);
        }
        epilogue_rows_bf16<BN>(taddr, smem_u32(s_bias_g + nt * BN), smem_u32(s_epi + (warp - kEpiWarp0) * kEpiStageBytes),
                               stats_sum ? st_sum : nullptr, st_sq, reinterpret_cast<__nv_bfloat16*>(out_raw), soff,
                               svalid, nz, valid, p.gain, p.has_act ? p.slope : 1.0f, lane);
      }
      // all TMEM reads of this warp are complete (wait::ld above): hand the accumulator stage back
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[as]);

      if (!OUT_F32 && stats_sum != nullptr) {
        if (group == 0) asm volatile("bar.sync 1, 128;" ::: "memory");     // the 4 warps of this group
        else asm volatile("bar.sync 2, 128;" ::: "memory");
        const int e = (threadIdx.x - kEpiWarp0 * 32) & 127;
        const float* b_sum = s_stats + (size_t)(as * 2 + 0) * 4 * BN;
        const float* b_sq = s_stats + (size_t)(as * 2 + 1) * 4 * BN;
        const int slot = ((ty * p.tiles_x + tx) * n_phases + ph);
        for (int c = e; c < BN; c += 128) {
          const float s = (b_sum[c] + b_sum[BN + c]) + (b_sum[2 * BN + c] + b_sum[3 * BN + c]);
          const float qv = (b_sq[c] + b_sq[BN + c]) + (b_sq[2 * BN + c] + b_sq[3 * BN + c]);
          const int64_t o = ((int64_t)img * p.stats_partials + slot) * p.cout + nt * BN + c;
          stats_sum[o] = s;
          stats_sq[o] = qv;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}


// =====================================================================================================================
// Halo variant: stride-1 convolutions with kh > 1 (the 3x3 layers and the 2x2 phase kernels of the transposed convs).
//
// The plain kernel above fetches one A box per tap, so a 3x3 layer pulls its input tile into the SM nine times and the
// layers with cout <= 128 are bound by the SM's L2 read port (ncu: 13 TB/s of L2->SM traffic at 0.43 of the tensor
// peak).  Here
//   * ONE A box of (MT*th + kh - 1) rows serves the kh taps of a kernel column: tap (ky, kx) is the same box read from
//     row ky on -- a descriptor start address ky * tw * BKC*2 bytes further, a whole number of swizzle atoms because
//     tw >= 8 -- so A is fetched kw * (1 + (kh-1)/(MT*th)) times instead of kh*kw times;
//   * a work item is MT = 2 vertically stacked 128-pixel sub-tiles that share every B tile (two accumulators of BN
//     columns per TMEM stage), halving the weight traffic per FLOP when BN <= 128.
// A and B travel through separate rings (A slots live for kh taps, B slots for one tap or one kernel column).
// K order: channel block, kernel column, kernel row -- fixed, so results do not depend on batch or grid.
// =====================================================================================================================
constexpr int kMaxA = 8, kMaxB = 8;

template <int BN, int MT>
struct HaloSmem {
  static constexpr int kStatsBytes = 2 /*acc stage*/ * MT * 2 /*sum,sq*/ * 4 /*warps*/ * BN * 4;
  static constexpr int kBiasBytes = 2 * kBiasSmem * 4;
  static constexpr int kBarBytes = (2 * kMaxA + 2 * kMaxB + 4) * 8 + 16;
  static constexpr int kEpiBytes = 8 * kEpiStageBytes;
  static constexpr int kFixedBytes = kStatsBytes + kBiasBytes + kEpiBytes + kBarBytes + 1024 /*align slack*/;
  static constexpr int kMaxDynamic = 232448 - 1024;
};

template <int BN, int BKC, int MT>
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_halo_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    __nv_bfloat16* __restrict__ out, const float* __restrict__ bias, const float* __restrict__ noise,
                    float* __restrict__ stats_sum, float* __restrict__ stats_sq, const Params p) {
  using S = HaloSmem<BN, MT>;
  constexpr int kBBytes = BN * BKC * 2;
  constexpr int kSubBytes = BM * BKC * 2;            // one 128-pixel sub-tile of an A box
  constexpr int kAccCols = MT * BN;                  // TMEM columns per accumulator stage
  constexpr uint32_t kTmemCols = (2 * kAccCols <= 32) ? 32 : (2 * kAccCols <= 64) ? 64 : (2 * kAccCols <= 128) ? 128
                                 : (2 * kAccCols <= 256) ? 256 : 512;
  static_assert(2 * kAccCols <= 512, "accumulators exceed tensor memory");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int b_slot_bytes = p.bps * kBBytes;
  uint8_t* a_ring = smem;
  uint8_t* b_ring = smem + p.na * p.a_slot_bytes;
  uint8_t* fixed = b_ring + p.nb * b_slot_bytes;
  float* s_stats = reinterpret_cast<float*>(fixed);
  float* s_bias = reinterpret_cast<float*>(fixed + S::kStatsBytes);
  uint8_t* s_epi = fixed + S::kStatsBytes + S::kBiasBytes;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(s_epi + S::kEpiBytes);
  uint64_t* a_empty = a_full + kMaxA;
  uint64_t* b_full = a_empty + kMaxA;
  uint64_t* b_empty = b_full + kMaxB;
  uint64_t* tmem_full = b_empty + kMaxB;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  for (int i = threadIdx.x; i < p.cout; i += kThreads) s_bias[i] = s_bias[kBiasSmem + i] = bias ? bias[i] : 0.0f;   // cout <= kBiasSmem (host)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.na; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < p.nb; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 4); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_phases = p.phases_y * p.phases_x;

  // work item -> (n-tile, phase, tile column, super-tile row, image)
  auto decode = [&](int tile, int& nt, int& ph, int& tx, int& tys, int& img) {
    nt = tile % p.n_tiles_n; tile /= p.n_tiles_n;
    ph = tile % n_phases; tile /= n_phases;
    tx = tile % p.tiles_x; tile /= p.tiles_x;
    tys = tile % p.tiles_ys; img = tile / p.tiles_ys;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int sa = 0, sb = 0;
      uint32_t pha = 0, phb = 0;
      const uint32_t a_box_bytes = (uint32_t)p.a_box_bytes;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        int nt, ph, tx, tys, img;
        decode(tile, nt, ph, tx, tys, img);
        const int py = ph / p.phases_x, px = ph % p.phases_x;
        const int iy0 = tys * MT * p.th * p.stride + p.off_y + py * p.phase_off_dy;
        const int ix0 = tx * p.tw * p.stride + p.off_x + px * p.phase_off_dx;
        const int brow = img * p.w_img_rows + ph * p.cout + nt * BN;
        for (int cb = 0; cb < p.kblocks_per_tap; ++cb) {
          for (int kx = 0; kx < p.kw; ++kx) {
            mbar_wait(&a_empty[sa], pha ^ 1);
            mbar_arrive_expect_tx(&a_full[sa], a_box_bytes * (uint32_t)p.stride);
            tma_load_4d(a_ring + sa * p.a_slot_bytes, &map_a, &a_full[sa], cb * BKC, ix0 + kx, iy0, img);
            if (p.stride == 2)     // the map skips every other row: second box = the odd rows (taps ky = 1, 3, ...)
              tma_load_4d(a_ring + sa * p.a_slot_bytes + a_box_bytes, &map_a, &a_full[sa], cb * BKC, ix0 + kx, iy0 + 1, img);
            if (++sa == p.na) { sa = 0; pha ^= 1; }
            for (int ky0 = 0; ky0 < p.kh; ky0 += p.bps) {
              mbar_wait(&b_empty[sb], phb ^ 1);
              mbar_arrive_expect_tx(&b_full[sb], (uint32_t)b_slot_bytes);
              for (int j = 0; j < p.bps; ++j)
                tma_load_2d(b_ring + sb * b_slot_bytes + j * kBBytes, &map_b, &b_full[sb],
                            ((ky0 + j) * p.kw + kx) * p.cin + cb * BKC, brow);
              if (++sb == p.nb) { sb = 0; phb ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      const uint32_t idesc = make_idesc<BN>();
      const uint32_t dy_bytes = (uint32_t)(p.tw * BKC * 2);
      int sa = 0, sb = 0, it = 0;
      uint32_t pha = 0, phb = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * kAccCols);
        uint32_t accumulate = 0;
        for (int cb = 0; cb < p.kblocks_per_tap; ++cb) {
          for (int kx = 0; kx < p.kw; ++kx) {
            mbar_wait(&a_full[sa], pha);
            tc_fence_after();
            const uint32_t a_base = smem_u32(a_ring + sa * p.a_slot_bytes);
            for (int ky0 = 0; ky0 < p.kh; ky0 += p.bps) {
              mbar_wait(&b_full[sb], phb);
              tc_fence_after();
              const uint32_t b_base = smem_u32(b_ring + sb * b_slot_bytes);
              for (int j = 0; j < p.bps; ++j) {
                const uint64_t db = make_smem_desc<BKC>(b_base + (uint32_t)(j * kBBytes));
                // tap row ky starts ky rows into the box; row stride 2: ky / 2 rows into the box of its parity
                const int ky = ky0 + j;
                const uint32_t a_tap = p.stride == 2 ? (uint32_t)(ky & 1) * (uint32_t)p.a_box_bytes + (uint32_t)(ky >> 1) * dy_bytes
                                                     : (uint32_t)ky * dy_bytes;
#pragma unroll
                for (int m = 0; m < MT; ++m) {
                  const uint64_t da = make_smem_desc<BKC>(a_base + a_tap + (uint32_t)(m * kSubBytes));
#pragma unroll
                  for (int k = 0; k < BKC / 16; ++k)
                    umma_bf16(tmem_d + (uint32_t)(m * BN), da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc,
                              (accumulate | (uint32_t)k) != 0 ? 1u : 0u);
                }
                accumulate = 1;
              }
              umma_commit(&b_empty[sb]);
              if (++sb == p.nb) { sb = 0; phb ^= 1; }
            }
            umma_commit(&a_empty[sa]);
            if (++sa == p.na) { sa = 0; pha ^= 1; }
          }
        }
        umma_commit(&tmem_full[as]);
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ===================== epilogue =====================
    const int q = warp & 3;
    const int group = (warp - kEpiWarp0) >> 2;
    const int row = q * 32 + lane;
    const int ety = row / p.tw, etx = row - ety * p.tw;
    float* s_bias_g = s_bias + group * kBiasSmem;
    const int fh_pad = p.full_h + 2 * p.out_pad - p.out_pad_trim, fw_pad = p.full_w + 2 * p.out_pad - p.out_pad_trim;
    // store-side rows of this lane (rows 8i + lane/4 of the warp's 32): tile-invariant, so the divisions happen once
    int s_y[4], s_x[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int srow = q * 32 + 8 * i + (lane >> 2);
      s_y[i] = srow / p.tw;
      s_x[i] = srow - s_y[i] * p.tw;
    }
    int it = 0, cur_img = -1;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      if ((it & 1) != group) continue;
      int nt, ph, tx, tys, img;
      decode(tile, nt, ph, tx, tys, img);
      if (p.bias_img_stride > 0 && img != cur_img) {
        reload_group_bias(s_bias_g, bias, img, p.bias_img_stride, p.cout, group, (threadIdx.x - kEpiWarp0 * 32) & 127);
        cur_img = img;
      }
      const int py = ph / p.phases_x, px = ph % p.phases_x;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after();
#pragma unroll 1
      for (int m = 0; m < MT; ++m) {
        const int ty = tys * MT + m;
        const int oy = ty * p.th + ety, ox = tx * p.tw + etx;
        const bool valid = (oy < p.out_h) && (ox < p.out_w);
        const int64_t pix = ((int64_t)img * p.full_h + (oy * p.phases_y + py)) * p.full_w + (ox * p.phases_x + px);
        const float nz = (noise != nullptr && valid) ? p.noise_w * noise[pix] : 0.0f;
        float* b_sum = s_stats + (size_t)((as * MT + m) * 2 + 0) * 4 * BN;
        float* b_sq = s_stats + (size_t)((as * MT + m) * 2 + 1) * 4 * BN;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * kAccCols + m * BN);
        int64_t soff[4];
        uint32_t svalid = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int soy = ty * p.th + s_y[i], sox = tx * p.tw + s_x[i];
          if (soy < p.out_h && sox < p.out_w) svalid |= 1u << i;
          const int64_t spix = ((int64_t)img * fh_pad + (soy * p.phases_y + py) + p.out_pad) * fw_pad + (sox * p.phases_x + px) + p.out_pad;
          soff[i] = spix * p.cout_stride + p.c_off + nt * BN + 8 * (lane & 3);
        }
        epilogue_rows_bf16<BN>(taddr, smem_u32(s_bias_g + nt * BN), smem_u32(s_epi + (warp - kEpiWarp0) * kEpiStageBytes),
                               stats_sum ? b_sum + q * BN : nullptr, b_sq + q * BN, out, soff, svalid, nz, valid, p.gain,
                               p.has_act ? p.slope : 1.0f, lane);
        if (m == MT - 1) {
          // every TMEM read of this warp has completed (wait::ld above): hand the accumulator stage back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[as]);
        }
        if (stats_sum != nullptr) {
          if (group == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
          else asm volatile("bar.sync 2, 128;" ::: "memory");
          if (ty < p.tiles_y) {
            const int e = (threadIdx.x - kEpiWarp0 * 32) & 127;
            const int slot = ((ty * p.tiles_x + tx) * n_phases + ph);
            for (int c = e; c < BN; c += 128) {
              const float s = (b_sum[c] + b_sum[BN + c]) + (b_sum[2 * BN + c] + b_sum[3 * BN + c]);
              const float qv = (b_sq[c] + b_sq[BN + c]) + (b_sq[2 * BN + c] + b_sq[3 * BN + c]);
              const int64_t o = ((int64_t)img * p.stats_partials + slot) * p.cout + nt * BN + c;
              stats_sum[o] = s;
              stats_sq[o] = qv;
            }
          }
          if (group == 0) asm volatile("bar.sync 1, 128;" ::: "memory");   // staging buffer free for its next use
          else asm volatile("bar.sync 2, 128;" ::: "memory");
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}


// =====================================================================================================================
// CTA-pair (cta_group::2) form of the halo kernel for the layers whose 1-CTA form is bound by shared-memory bandwidth
// (cout = 128: every M128 x N128 x K16 MMA reads 8 KB of operands in 64 clocks = the SM's whole 128 B/clk, while TMA writes
// the next stages into the same memory).  Two CTAs on the two SMs of a TPC work on horizontally adjacent pixel tiles
// (tx = 2*txp + rank) of the same image / phase / cout tile:
//   * each CTA loads its own A boxes and only HALF of every B tile (its BN/2 weight rows); tcgen05.mma.cta_group::2,
//     issued by the even CTA alone, multiplies both CTAs' A tiles with the B tile gathered from both halves -- per CTA the
//     operand reads drop to 6 KB per MMA and the B fills halve;
//   * accumulators stay per CTA (its 128 pixel rows x BN columns per sub-tile), so the epilogue is the 1-CTA one and every
//     output element sees the same K order: results are bit-identical to conv_tc_halo_kernel.
// Barriers: "full" barriers live in the even CTA and collect the transaction bytes of BOTH CTAs' TMA loads (the odd CTA's
// loads name them through the peer-bit-masked address); tcgen05.commit multicasts the "empty" / "accumulator ready"
// arrivals to both CTAs; the odd CTA's epilogue warps release an accumulator stage by a remote arrive on the even CTA.
// =====================================================================================================================
template <int BN, int BKC, int MT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
conv_tc_halo2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                     __nv_bfloat16* __restrict__ out, const float* __restrict__ bias, const float* __restrict__ noise,
                     float* __restrict__ stats_sum, float* __restrict__ stats_sq, const Params p) {
  using S = HaloSmem<BN, MT>;
  constexpr int BNH = BN / 2;                        // weight rows this CTA holds of every B tile
  constexpr int kBBytes = BNH * BKC * 2;
  constexpr int kSubBytes = BM * BKC * 2;
  constexpr int kAccCols = MT * BN;
  constexpr uint32_t kTmemCols = (2 * kAccCols <= 32) ? 32 : (2 * kAccCols <= 64) ? 64 : (2 * kAccCols <= 128) ? 128
                                 : (2 * kAccCols <= 256) ? 256 : 512;
  static_assert(2 * kAccCols <= 512, "accumulators exceed tensor memory");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int b_slot_bytes = p.bps * kBBytes;
  uint8_t* a_ring = smem;
  uint8_t* b_ring = smem + p.na * p.a_slot_bytes;
  uint8_t* fixed = b_ring + p.nb * b_slot_bytes;
  float* s_stats = reinterpret_cast<float*>(fixed);
  float* s_bias = reinterpret_cast<float*>(fixed + S::kStatsBytes);
  uint8_t* s_epi = fixed + S::kStatsBytes + S::kBiasBytes;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(s_epi + S::kEpiBytes);
  uint64_t* a_empty = a_full + kMaxA;
  uint64_t* b_full = a_empty + kMaxA;
  uint64_t* b_empty = b_full + kMaxB;
  uint64_t* tmem_full = b_empty + kMaxB;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;    // [2] (the even CTA's copy is the one in use)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  for (int i = threadIdx.x; i < p.cout; i += kThreads) s_bias[i] = s_bias[kBiasSmem + i] = bias ? bias[i] : 0.0f;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.na; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < p.nb; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 8); }   // 4 warps x 2 CTAs
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_2sm(tmem_slot, kTmemCols);
  tc_fence_before();
  cluster_sync_all();                     // barriers of both CTAs are initialised before any remote arrive / TMA signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_phases = p.phases_y * p.phases_x;
  const int tiles_xp = p.tiles_x >> 1;    // tile-column pairs (tiles_x is even, host-checked)

  // cluster work item -> (n-tile, phase, tile-column pair, super-tile row, image)
  auto decode = [&](int tile, int& nt, int& ph, int& tx, int& tys, int& img) {
    nt = tile % p.n_tiles_n; tile /= p.n_tiles_n;
    ph = tile % n_phases; tile /= n_phases;
    tx = 2 * (tile % tiles_xp) + rank; tile /= tiles_xp;
    tys = tile % p.tiles_ys; img = tile / p.tiles_ys;
  };

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (elect_one()) {
      int sa = 0, sb = 0;
      uint32_t pha = 0, phb = 0;
      const uint32_t a_box_bytes = (uint32_t)(p.a_rows * p.tw * BKC * 2);
      for (int tile = cluster_id; tile < p.num_tiles; tile += num_clusters) {
        int nt, ph, tx, tys, img;
        decode(tile, nt, ph, tx, tys, img);
        const int py = ph / p.phases_x, px = ph % p.phases_x;
        const int iy0 = tys * MT * p.th + p.off_y + py * p.phase_off_dy;
        const int ix0 = tx * p.tw + p.off_x + px * p.phase_off_dx;
        const int brow = img * p.w_img_rows + ph * p.cout + nt * BN + rank * BNH;
        for (int cb = 0; cb < p.kblocks_per_tap; ++cb) {
          for (int kx = 0; kx < p.kw; ++kx) {
            mbar_wait(&a_empty[sa], pha ^ 1);
            if (rank == 0) mbar_arrive_expect_tx(&a_full[sa], 2u * a_box_bytes);        // both CTAs' boxes
            tma_load_4d_2sm(a_ring + sa * p.a_slot_bytes, &map_a, &a_full[sa], cb * BKC, ix0 + kx, iy0, img);
            if (++sa == p.na) { sa = 0; pha ^= 1; }
            for (int ky0 = 0; ky0 < p.kh; ky0 += p.bps) {
              mbar_wait(&b_empty[sb], phb ^ 1);
              if (rank == 0) mbar_arrive_expect_tx(&b_full[sb], 2u * (uint32_t)b_slot_bytes);
              for (int j = 0; j < p.bps; ++j)
                tma_load_2d_2sm(b_ring + sb * b_slot_bytes + j * kBBytes, &map_b, &b_full[sb],
                                ((ky0 + j) * p.kw + kx) * p.cin + cb * BKC, brow);
              if (++sb == p.nb) { sb = 0; phb ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (even CTA only) =====================
    if (rank == 0 && elect_one()) {
      // kind::f16: D=f32, A=B=bf16, K-major; N>>3 at bit 17, M>>4 at bit 24 with M = 256 across the pair
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      const uint32_t dy_bytes = (uint32_t)(p.tw * BKC * 2);
      int sa = 0, sb = 0, it = 0;
      uint32_t pha = 0, phb = 0;
      for (int tile = cluster_id; tile < p.num_tiles; tile += num_clusters, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[as], aphase ^ 1);            // both CTAs' epilogues have drained this stage
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * kAccCols);
        uint32_t accumulate = 0;
        for (int cb = 0; cb < p.kblocks_per_tap; ++cb) {
          for (int kx = 0; kx < p.kw; ++kx) {
            mbar_wait(&a_full[sa], pha);
            tc_fence_after();
            const uint32_t a_base = smem_u32(a_ring + sa * p.a_slot_bytes);
            for (int ky0 = 0; ky0 < p.kh; ky0 += p.bps) {
              mbar_wait(&b_full[sb], phb);
              tc_fence_after();
              const uint32_t b_base = smem_u32(b_ring + sb * b_slot_bytes);
              for (int j = 0; j < p.bps; ++j) {
                const uint64_t db = make_smem_desc<BKC>(b_base + (uint32_t)(j * kBBytes));
#pragma unroll
                for (int m = 0; m < MT; ++m) {
                  const uint64_t da = make_smem_desc<BKC>(a_base + (uint32_t)(ky0 + j) * dy_bytes + (uint32_t)(m * kSubBytes));
#pragma unroll
                  for (int k = 0; k < BKC / 16; ++k)
                    umma_bf16_2sm(tmem_d + (uint32_t)(m * BN), da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc,
                                  (accumulate | (uint32_t)k) != 0 ? 1u : 0u);
                }
                accumulate = 1;
              }
              umma_commit_2sm(&b_empty[sb]);
              if (++sb == p.nb) { sb = 0; phb ^= 1; }
            }
            umma_commit_2sm(&a_empty[sa]);
            if (++sa == p.na) { sa = 0; pha ^= 1; }
          }
        }
        umma_commit_2sm(&tmem_full[as]);
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ===================== epilogue (both CTAs, each on its own pixels) =====================
    const int q = warp & 3;
    const int group = (warp - kEpiWarp0) >> 2;
    const int row = q * 32 + lane;
    const int ety = row / p.tw, etx = row - ety * p.tw;
    float* s_bias_g = s_bias + group * kBiasSmem;
    const int fh_pad = p.full_h + 2 * p.out_pad - p.out_pad_trim, fw_pad = p.full_w + 2 * p.out_pad - p.out_pad_trim;
    int it = 0, cur_img = -1;
    for (int tile = cluster_id; tile < p.num_tiles; tile += num_clusters, ++it) {
      if ((it & 1) != group) continue;
      int nt, ph, tx, tys, img;
      decode(tile, nt, ph, tx, tys, img);
      if (p.bias_img_stride > 0 && img != cur_img) {
        reload_group_bias(s_bias_g, bias, img, p.bias_img_stride, p.cout, group, (threadIdx.x - kEpiWarp0 * 32) & 127);
        cur_img = img;
      }
      const int py = ph / p.phases_x, px = ph % p.phases_x;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after();
#pragma unroll 1
      for (int m = 0; m < MT; ++m) {
        const int ty = tys * MT + m;
        const int oy = ty * p.th + ety, ox = tx * p.tw + etx;
        const bool valid = (oy < p.out_h) && (ox < p.out_w);
        const int64_t pix = ((int64_t)img * p.full_h + (oy * p.phases_y + py)) * p.full_w + (ox * p.phases_x + px);
        const float nz = (noise != nullptr && valid) ? p.noise_w * noise[pix] : 0.0f;
        float* b_sum = s_stats + (size_t)((as * MT + m) * 2 + 0) * 4 * BN;
        float* b_sq = s_stats + (size_t)((as * MT + m) * 2 + 1) * 4 * BN;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * kAccCols + m * BN);
        int64_t soff[4];
        uint32_t svalid = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int srow = q * 32 + 8 * i + (lane >> 2);
          const int sy = srow / p.tw, sx = srow - sy * p.tw;
          const int soy = ty * p.th + sy, sox = tx * p.tw + sx;
          if (soy < p.out_h && sox < p.out_w) svalid |= 1u << i;
          const int64_t spix = ((int64_t)img * fh_pad + (soy * p.phases_y + py) + p.out_pad) * fw_pad + (sox * p.phases_x + px) + p.out_pad;
          soff[i] = spix * p.cout_stride + p.c_off + nt * BN + 8 * (lane & 3);
        }
        epilogue_rows_bf16<BN>(taddr, smem_u32(s_bias_g + nt * BN), smem_u32(s_epi + (warp - kEpiWarp0) * kEpiStageBytes),
                               stats_sum ? b_sum + q * BN : nullptr, b_sq + q * BN, out, soff, svalid, nz, valid, p.gain,
                               p.has_act ? p.slope : 1.0f, lane);
        if (m == MT - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_even_cta(&tmem_empty[as]);
        }
        if (stats_sum != nullptr) {
          if (group == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
          else asm volatile("bar.sync 2, 128;" ::: "memory");
          if (ty < p.tiles_y) {
            const int e = (threadIdx.x - kEpiWarp0 * 32) & 127;
            const int slot = ((ty * p.tiles_x + tx) * n_phases + ph);
            for (int c = e; c < BN; c += 128) {
              const float s = (b_sum[c] + b_sum[BN + c]) + (b_sum[2 * BN + c] + b_sum[3 * BN + c]);
              const float qv = (b_sq[c] + b_sq[BN + c]) + (b_sq[2 * BN + c] + b_sq[3 * BN + c]);
              const int64_t o = ((int64_t)img * p.stats_partials + slot) * p.cout + nt * BN + c;
              stats_sum[o] = s;
              stats_sq[o] = qv;
            }
          }
          if (group == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
          else asm volatile("bar.sync 2, 128;" ::: "memory");
        }
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();                     // no CTA leaves while its pair still reads its shared memory / signals its barriers
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, kTmemCols);
  }
}

// =====================================================================================================================
// Weight-gradient kernel with MN-major operands (entry point in wgrad.cu): dw[co][tap][ci] = sum_pixels dy[pix][co] * x[pix + tap][ci].
// The contraction runs over pixels, and NHWC tensors are pixel-major with the channels contiguous -- exactly the
// "MN-major" operand form of tcgen05.mma (instruction-descriptor bits 15 / 16): a TMA box {64 channels, 8 x 8 pixels} lands
// in shared memory as 64 rows (K = pixels) of 128 bytes (64 channels of M or N) in the 128-byte swizzle, which is the
// canonical MN-major SWIZZLE_128B layout ((8,8,m),(8,k)):((1,8,LBO),(64,SBO)) with SBO = 1024 bytes between 8-pixel groups
// and LBO = one box (8192 bytes) between 64-channel blocks.  No transposition pass, no padded copies: a filter tap is a
// coordinate offset of the x box (stride 2 = TMA element strides), zero padding is TMA out-of-bounds fill.
// Work item = (tap, ci tile, co tile, sample, pixel chunk); taps fastest so the CTAs sharing a dy chunk run together.
struct WgParams {
  int n, out_h, out_w, cout, cin, kh, kw, stride, off_y, off_x;
  int tiles_x, patches, chunk, k_split;      // 8x8 patches per row / per sample; patches per chunk; chunks per sample
  int m_tiles, n_tiles, num_tiles, stages;
};

template <int BN>
__global__ void __launch_bounds__(192, 1)
conv_wgrad_mn_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x,
                     float* __restrict__ partials, const WgParams p) {
  constexpr int kBox = 64 * 64 * 2;                 // one {64 ch, 8, 8} box
  constexpr int kABytes = 2 * kBox, kBBytes = (BN / 64) * kBox, kStage = kABytes + kBBytes;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.stages * kStage);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tmem_full = empty_bar + kMaxStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t kTmemCols = (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&map_dy); tma_prefetch_desc(&map_x); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 4); }
    fence_barrier_init();
  }
  __syncwarp();
  if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int taps = p.kh * p.kw;

  auto decode = [&](int tile, int& tap, int& nt, int& mt, int& j) {
    tap = tile % taps; tile /= taps;
    nt = tile % p.n_tiles; tile /= p.n_tiles;
    mt = tile % p.m_tiles; j = tile / p.m_tiles;
  };

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        int tap, nt, mt, j;
        decode(tile, tap, nt, mt, j);
        const int img = j / p.k_split, ch = j - img * p.k_split;
        const int ty_t = tap / p.kw, tx_t = tap - ty_t * p.kw;
        const int p0 = ch * p.chunk, p1 = min(p0 + p.chunk, p.patches);
        int py = p0 / p.tiles_x, px = p0 - py * p.tiles_x;
        for (int pp = p0; pp < p1; ++pp) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * kStage;
          mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)kStage);
#pragma unroll
          for (int b = 0; b < 2; ++b)
            tma_load_4d(sa + b * kBox, &map_dy, &full_bar[stage], mt * 128 + b * 64, px * 8, py * 8, img);
          const int ix = px * 8 * p.stride + tx_t + p.off_x, iy = py * 8 * p.stride + ty_t + p.off_y;
#pragma unroll
          for (int b = 0; b < BN / 64; ++b)
            tma_load_4d(sa + kABytes + b * kBox, &map_x, &full_bar[stage], nt * BN + b * 64, ix, iy, img);
          if (++px == p.tiles_x) { px = 0; ++py; }
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // kind::f16, D = f32, A = B = bf16, BOTH MN-major (bits 15, 16), N >> 3 at bit 17, M >> 4 at bit 24
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(BN >> 3) << 17) |
                             ((uint32_t)(128 >> 4) << 24);
      auto desc = [](uint32_t saddr) {
        uint64_t d = 0;
        d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
        d |= (uint64_t)(kBox >> 4) << 16;              // LBO: next 64-channel block
        d |= (uint64_t)(1024 >> 4) << 32;              // SBO: next group of 8 pixels
        d |= (uint64_t)1 << 46;
        d |= (uint64_t)2 << 61;                        // SWIZZLE_128B
        return d;
      };
      int stage = 0, it = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        int tap, nt, mt, j;
        decode(tile, tap, nt, mt, j);
        const int ch = j % p.k_split;
        const int p0 = ch * p.chunk, p1 = min(p0 + p.chunk, p.patches);
        const int as = it & 1;
        mbar_wait(&tmem_empty[as], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
        for (int pp = p0; pp < p1; ++pp) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * kStage);
          const uint64_t da = desc(sa), db = desc(sa + kABytes);
#pragma unroll
          for (int k = 0; k < 4; ++k)                 // 16 pixels = 2048 bytes per MMA
            umma_bf16(tmem_d, da + (uint64_t)(k * (2048 >> 4)), db + (uint64_t)(k * (2048 >> 4)), idesc,
                      (pp > p0 || k > 0) ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[as]);
      }
    }
  } else {
    // epilogue: warps 2..5, TMEM lane quadrant = warp % 4; lane = output channel row of the 128-row tile
    const int q = warp & 3;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      int tap, nt, mt, j;
      decode(tile, tap, nt, mt, j);
      const int as = it & 1;
      mbar_wait(&tmem_full[as], (it >> 1) & 1);
      tc_fence_after();
      const int co = mt * 128 + q * 32 + lane;
      float* dst = partials + (((int64_t)j * taps + tap) * p.cout + co) * p.cin + nt * BN;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(taddr + (uint32_t)c0, r);
        tmem_ld_wait();
        if (co < p.cout) {
#pragma unroll
          for (int v = 0; v < 8; ++v)
            if (nt * BN + c0 + 4 * v < p.cin)
              *reinterpret_cast<float4*>(dst + c0 + 4 * v) =
                  make_float4(__uint_as_float(r[4 * v]), __uint_as_float(r[4 * v + 1]), __uint_as_float(r[4 * v + 2]),
                              __uint_as_float(r[4 * v + 3]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[as]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}


// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  });
  return fn;
}

struct MapKey {
  uint64_t v[13];
  bool operator==(const MapKey& o) const { return std::memcmp(v, o.v, sizeof(v)) == 0; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    uint64_t h = 1469598103934665603ull;
    for (uint64_t x : k.v) { h ^= x; h *= 1099511628211ull; }
    return (size_t)h;
  }
};
static std::mutex g_map_mu;
static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;

static int encode_map(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_b,
                      const uint32_t* box, const uint32_t* estr, int bkc) {
  MapKey key;
  std::memset(&key, 0, sizeof(key));
  key.v[0] = (uint64_t)base;
  key.v[1] = (uint64_t)rank | ((uint64_t)bkc << 8);
  for (int i = 0; i < rank; ++i) {
    key.v[2 + i] = dims[i];
    key.v[6 + i] = ((uint64_t)box[i] << 32) | estr[i];
  }
  key.v[10] = strides_b[0];
  key.v[11] = rank > 2 ? strides_b[1] : 0;
  key.v[12] = rank > 3 ? strides_b[2] : 0;
  {
    std::lock_guard<std::mutex> lk(g_map_mu);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) { *out = it->second; return 0; }
  }
  EncodeTiledFn enc = get_encode();
  PPST_REQUIRE(enc != nullptr, PPST_ERR_CUDA, "conv_tc: cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t gdim[4], gstr[3];
  cuuint32_t gbox[4], gest[4];
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; gbox[i] = box[i]; gest[i] = estr[i]; }
  for (int i = 0; i < rank - 1; ++i) gstr[i] = strides_b[i];
  const CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr,
                         gbox, gest, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         bkc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PPST_REQUIRE(r == CUDA_SUCCESS, PPST_ERR_CUDA, "conv_tc: cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  std::lock_guard<std::mutex> lk(g_map_mu);
  if (g_maps.size() > 4096) g_maps.clear();
  g_maps[key] = *out;
  return 0;
}

static void pick_tile(int out_h, int out_w, int& th, int& tw) {
  // 128 pixels; prefer a squarish patch (halo reuse in L2), shrink TW for narrow outputs
  tw = 16;
  while (tw > out_w && tw > 1) tw >>= 1;
  if (out_w >= 128 && out_h < 8) tw = 128;
  th = BM / tw;
}

// per-device state: SM count and "dynamic shared memory attribute already set" flags (cudaFuncSetAttribute is per device)
constexpr int kMaxDevices = 64;
static int cur_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < kMaxDevices) ? dev : 0;
}
static int g_num_sms[kMaxDevices] = {0};
static int num_sms() {
  const int dev = cur_device();
  if (g_num_sms[dev] == 0) {
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    g_num_sms[dev] = v > 0 ? v : 148;
  }
  return g_num_sms[dev];
}

template <int BN, int BKC, bool OUT_F32>
static int launch(const ppst_conv_desc& d, void* out, const void* in, const void* weight, const float* bias,
                  const float* noise, float* stats, cudaStream_t st) {
  using S = Smem<BN, BKC>;
  Params p;
  std::memset(&p, 0, sizeof(p));
  p.n = d.n; p.out_h = d.out_h; p.out_w = d.out_w; p.cout = d.cout; p.cin = d.cin;
  p.kh = d.kh; p.kw = d.kw; p.stride = d.stride;
  p.off_y = d.off_y; p.off_x = d.off_x; p.phases_y = d.phases_y; p.phases_x = d.phases_x;
  p.phase_off_dy = d.phase_off_dy; p.phase_off_dx = d.phase_off_dx;
  p.full_h = d.full_h; p.full_w = d.full_w; p.cout_stride = d.cout_stride; p.c_off = d.c_off;
  pick_tile(d.out_h, d.out_w, p.th, p.tw);
  p.tiles_y = ceil_div(d.out_h, p.th); p.tiles_x = ceil_div(d.out_w, p.tw);
  p.n_tiles_n = d.cout / BN;
  const int n_phases = d.phases_y * d.phases_x;
  const int64_t tiles = (int64_t)d.n * p.tiles_y * p.tiles_x * n_phases * p.n_tiles_n;
  PPST_REQUIRE(tiles < (1LL << 31), PPST_ERR_ARG, "conv_tc: too many tiles");
  p.num_tiles = (int)tiles;
  p.kblocks_per_tap = d.cin / BKC;
  {
    // k-blocks per stage: enough bytes per barrier round trip (the single-thread producer / issuer loops cost a few
    // hundred cycles per iteration whatever the tile size), as few iterations as possible, deepest pipeline that fits
    const int kblocks = d.split_terms == 3 ? 3 * ((d.cin / 2) / BKC) : d.kh * d.kw * (d.cin / BKC);
    const int kps_max = 49152 / S::kKblockBytes > 0 ? 49152 / S::kKblockBytes : 1;
    const int iters = (kblocks + kps_max - 1) / kps_max;
    p.kps = (kblocks + iters - 1) / iters;
    const int stage_bytes = p.kps * S::kKblockBytes;
    int stages = (S::kMaxDynamic - S::kFixedBytes) / stage_bytes;
    p.stages = stages > kMaxStages ? kMaxStages : stages;
    PPST_REQUIRE(p.stages >= 2, PPST_ERR_UNSUPPORTED, "conv_tc: pipeline does not fit in shared memory");
  }
  const int smem_total = p.stages * p.kps * S::kKblockBytes + S::kFixedBytes;
  p.w_img_rows = d.w_img_rows;
  p.out_pad = d.out_pad; p.out_pad_trim = d.out_pad_trim; p.bias_img_stride = d.bias_img_stride;
  p.split_seg_kb = d.split_terms == 3 ? (d.cin / 2) / BKC : 0;
  p.has_act = d.has_act; p.slope = d.slope; p.gain = d.gain; p.noise_w = d.noise_w;
  p.stats_partials = p.tiles_y * p.tiles_x * n_phases;

  CUtensorMap map_a, map_b;
  {
    const uint64_t dims[4] = {(uint64_t)d.cin, (uint64_t)d.in_w, (uint64_t)d.in_h, (uint64_t)d.n};
    const uint64_t rowp = d.in_row_pitch ? (uint64_t)d.in_row_pitch : (uint64_t)d.in_w * d.cin;
    const uint64_t imgp = d.in_img_pitch ? (uint64_t)d.in_img_pitch : (uint64_t)d.in_h * d.in_w * d.cin;
    const uint64_t strides[3] = {(uint64_t)d.cin * 2, rowp * 2, imgp * 2};
    const uint32_t box[4] = {(uint32_t)BKC, (uint32_t)(p.tw * p.stride), (uint32_t)(p.th * d.stride), 1};
    const uint32_t est[4] = {1, (uint32_t)p.stride, (uint32_t)d.stride, 1};
    int rc = encode_map(&map_a, in, 4, dims, strides, box, est, BKC);
    if (rc) return rc;
  }
  {
    const uint64_t K = (uint64_t)d.kh * d.kw * d.cin;
    const uint64_t dims[2] = {K, d.w_img_rows ? (uint64_t)d.w_img_rows * d.n : (uint64_t)d.cout * n_phases};
    const uint64_t strides[1] = {K * 2};
    const uint32_t box[2] = {(uint32_t)BKC, (uint32_t)BN};
    const uint32_t est[2] = {1, 1};
    int rc = encode_map(&map_b, weight, 2, dims, strides, box, est, BKC);
    if (rc) return rc;
  }
  auto kern = conv_tc_kernel<BN, BKC, OUT_F32>;
  static bool attr_set[kMaxDevices] = {false};
  if (!attr_set[cur_device()]) {
    int rc = cuda_check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kMaxDynamic),
                        "cudaFuncSetAttribute(conv_tc)");
    if (rc) return rc;
    attr_set[cur_device()] = true;
  }
  const int grid = p.num_tiles < num_sms() ? p.num_tiles : num_sms();
  float* ssum = stats;
  float* ssq = stats ? stats + (int64_t)d.n * p.stats_partials * d.cout : nullptr;
  kern<<<grid, kThreads, smem_total, st>>>(map_a, map_b, out, bias, noise, ssum, ssq, p);
  PPST_LAUNCH_CHECK("conv_tc_kernel");
  return 0;
}


constexpr int kHaloNoFit = -1000;                   // launch_halo only: not an error, never leaves this file
template <int BN, int BKC, int MT>
static int launch_halo(const ppst_conv_desc& d, void* out, const void* in, const void* weight, const float* bias,
                       const float* noise, float* stats, cudaStream_t st) {
  using S = HaloSmem<BN, MT>;
  Params p;
  std::memset(&p, 0, sizeof(p));
  p.n = d.n; p.out_h = d.out_h; p.out_w = d.out_w; p.cout = d.cout; p.cin = d.cin;
  p.kh = d.kh; p.kw = d.kw; p.stride = d.stride;
  p.off_y = d.off_y; p.off_x = d.off_x; p.phases_y = d.phases_y; p.phases_x = d.phases_x;
  p.phase_off_dy = d.phase_off_dy; p.phase_off_dx = d.phase_off_dx;
  p.full_h = d.full_h; p.full_w = d.full_w; p.cout_stride = d.cout_stride; p.c_off = d.c_off;
  pick_tile(d.out_h, d.out_w, p.th, p.tw);
  p.tiles_y = ceil_div(d.out_h, p.th); p.tiles_x = ceil_div(d.out_w, p.tw);
  p.tiles_ys = ceil_div(p.tiles_y, MT);
  p.n_tiles_n = d.cout / BN;
  const int n_phases = d.phases_y * d.phases_x;
  const int64_t tiles = (int64_t)d.n * p.tiles_ys * p.tiles_x * n_phases * p.n_tiles_n;
  PPST_REQUIRE(tiles < (1LL << 31), PPST_ERR_ARG, "conv_tc: too many tiles");
  p.num_tiles = (int)tiles;
  p.kblocks_per_tap = d.cin / BKC;
  p.has_act = d.has_act; p.slope = d.slope; p.gain = d.gain; p.noise_w = d.noise_w;
  p.stats_partials = p.tiles_y * p.tiles_x * n_phases;
  // row stride 2: the tensor map skips every other row, a slot holds the even-row box (taps 0, 2, ..) then the odd-row box
  p.a_rows = MT * p.th + (d.kh - 1) / d.stride;
  p.a_box_bytes = p.a_rows * p.tw * BKC * 2;                         // a whole number of swizzle atoms (tw >= 8)
  p.a_slot_bytes = (d.stride * p.a_box_bytes + 1023) / 1024 * 1024;
  constexpr int kBBytes = BN * BKC * 2;
  p.bps = (d.kh * kBBytes <= 24576) ? d.kh : 1;
  const int b_slot = p.bps * kBBytes;
  const int avail = S::kMaxDynamic - S::kFixedBytes;
  p.na = p.a_slot_bytes <= 24576 ? 4 : 3;
  if (BN <= 64) {
    // thin layers are bound by the latency of the A loads, not by bandwidth (ncu: DRAM 22 %, L2 27 %, tensor 15 % at
    // 8 x 512^2 x 32 -> 32): give the A ring most of the shared memory, several tiles' worth of boxes in flight
    static const int env_na = [] { const char* e = getenv("PPST_CONV_TC_NA"); return e ? atoi(e) : 8; }();
    p.na = env_na < 2 ? 2 : (env_na > kMaxA ? kMaxA : env_na);
    while (p.na > 2 && p.na * p.a_slot_bytes > (avail * 3) / 4) --p.na;
  }
  while (p.na > 2 && avail - p.na * p.a_slot_bytes < 3 * b_slot) --p.na;
  p.nb = (avail - p.na * p.a_slot_bytes) / b_slot;
  if (p.nb > kMaxB) p.nb = kMaxB;
  if (p.nb < 2 || p.na < 2) return kHaloNoFit;      // rings do not fit: the dispatcher tries fewer sub-tiles, then the plain kernel
  const int smem_total = p.na * p.a_slot_bytes + p.nb * b_slot + S::kFixedBytes;

  CUtensorMap map_a, map_b;
  {
    const uint64_t dims[4] = {(uint64_t)d.cin, (uint64_t)d.in_w, (uint64_t)d.in_h, (uint64_t)d.n};
    const uint64_t rowp = d.in_row_pitch ? (uint64_t)d.in_row_pitch : (uint64_t)d.in_w * d.cin;
    const uint64_t imgp = d.in_img_pitch ? (uint64_t)d.in_img_pitch : (uint64_t)d.in_h * d.in_w * d.cin;
    const uint64_t strides[3] = {(uint64_t)d.cin * 2, rowp * 2, imgp * 2};
    const uint32_t box[4] = {(uint32_t)BKC, (uint32_t)(p.tw * p.stride), (uint32_t)(p.a_rows * d.stride), 1};
    const uint32_t est[4] = {1, (uint32_t)p.stride, (uint32_t)d.stride, 1};
    int rc = encode_map(&map_a, in, 4, dims, strides, box, est, BKC);
    if (rc) return rc;
  }
  {
    const uint64_t K = (uint64_t)d.kh * d.kw * d.cin;
    const uint64_t dims[2] = {K, d.w_img_rows ? (uint64_t)d.w_img_rows * d.n : (uint64_t)d.cout * n_phases};
    const uint64_t strides[1] = {K * 2};
    const uint32_t box[2] = {(uint32_t)BKC, (uint32_t)BN};
    const uint32_t est[2] = {1, 1};
    int rc = encode_map(&map_b, weight, 2, dims, strides, box, est, BKC);
    if (rc) return rc;
  }
  p.w_img_rows = d.w_img_rows;
  p.out_pad = d.out_pad; p.out_pad_trim = d.out_pad_trim; p.bias_img_stride = d.bias_img_stride;
  auto kern = conv_tc_halo_kernel<BN, BKC, MT>;
  static bool attr_set[kMaxDevices] = {false};
  if (!attr_set[cur_device()]) {
    int rc = cuda_check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kMaxDynamic),
                        "cudaFuncSetAttribute(conv_tc_halo)");
    if (rc) return rc;
    attr_set[cur_device()] = true;
  }
  const int grid = p.num_tiles < num_sms() ? p.num_tiles : num_sms();
  float* ssum = stats;
  float* ssq = stats ? stats + (int64_t)d.n * p.stats_partials * d.cout : nullptr;
  kern<<<grid, kThreads, smem_total, st>>>(map_a, map_b, reinterpret_cast<__nv_bfloat16*>(out), bias, noise, ssum, ssq, p);
  PPST_LAUNCH_CHECK("conv_tc_halo_kernel");
  return 0;
}


// CTA-pair launch of the halo kernel (see conv_tc_halo2_kernel).  Same geometry / rings as launch_halo, half-size B tiles.
template <int BN, int BKC, int MT>
static int launch_halo2(const ppst_conv_desc& d, void* out, const void* in, const void* weight, const float* bias,
                        const float* noise, float* stats, cudaStream_t st) {
  using S = HaloSmem<BN, MT>;
  Params p;
  std::memset(&p, 0, sizeof(p));
  p.n = d.n; p.out_h = d.out_h; p.out_w = d.out_w; p.cout = d.cout; p.cin = d.cin;
  p.kh = d.kh; p.kw = d.kw; p.stride = 1;
  p.off_y = d.off_y; p.off_x = d.off_x; p.phases_y = d.phases_y; p.phases_x = d.phases_x;
  p.phase_off_dy = d.phase_off_dy; p.phase_off_dx = d.phase_off_dx;
  p.full_h = d.full_h; p.full_w = d.full_w; p.cout_stride = d.cout_stride; p.c_off = d.c_off;
  pick_tile(d.out_h, d.out_w, p.th, p.tw);
  p.tiles_y = ceil_div(d.out_h, p.th); p.tiles_x = ceil_div(d.out_w, p.tw);
  PPST_REQUIRE(p.tiles_x % 2 == 0, PPST_ERR_ARG, "conv_tc(halo2): odd number of tile columns");
  p.tiles_ys = ceil_div(p.tiles_y, MT);
  p.n_tiles_n = d.cout / BN;
  const int n_phases = d.phases_y * d.phases_x;
  const int64_t tiles = (int64_t)d.n * p.tiles_ys * (p.tiles_x / 2) * n_phases * p.n_tiles_n;     // cluster work items
  PPST_REQUIRE(tiles < (1LL << 31), PPST_ERR_ARG, "conv_tc: too many tiles");
  p.num_tiles = (int)tiles;
  p.kblocks_per_tap = d.cin / BKC;
  p.has_act = d.has_act; p.slope = d.slope; p.gain = d.gain; p.noise_w = d.noise_w;
  p.stats_partials = p.tiles_y * p.tiles_x * n_phases;
  p.a_rows = MT * p.th + d.kh - 1;
  p.a_box_bytes = p.a_rows * p.tw * BKC * 2;
  p.a_slot_bytes = (p.a_box_bytes + 1023) / 1024 * 1024;
  constexpr int kBBytes = (BN / 2) * BKC * 2;
  p.bps = (d.kh * kBBytes <= 24576) ? d.kh : 1;
  const int b_slot = p.bps * kBBytes;
  const int avail = S::kMaxDynamic - S::kFixedBytes;
  p.na = p.a_slot_bytes <= 24576 ? 4 : 3;
  while (p.na > 2 && avail - p.na * p.a_slot_bytes < 3 * b_slot) --p.na;
  p.nb = (avail - p.na * p.a_slot_bytes) / b_slot;
  if (p.nb > kMaxB) p.nb = kMaxB;
  PPST_REQUIRE(p.nb >= 2 && p.na >= 2, PPST_ERR_UNSUPPORTED, "conv_tc(halo2): rings do not fit in shared memory");
  const int smem_total = p.na * p.a_slot_bytes + p.nb * b_slot + S::kFixedBytes;

  CUtensorMap map_a, map_b;
  {
    const uint64_t dims[4] = {(uint64_t)d.cin, (uint64_t)d.in_w, (uint64_t)d.in_h, (uint64_t)d.n};
    const uint64_t rowp = d.in_row_pitch ? (uint64_t)d.in_row_pitch : (uint64_t)d.in_w * d.cin;
    const uint64_t imgp = d.in_img_pitch ? (uint64_t)d.in_img_pitch : (uint64_t)d.in_h * d.in_w * d.cin;
    const uint64_t strides[3] = {(uint64_t)d.cin * 2, rowp * 2, imgp * 2};
    const uint32_t box[4] = {(uint32_t)BKC, (uint32_t)p.tw, (uint32_t)p.a_rows, 1};
    const uint32_t est[4] = {1, 1, 1, 1};
    int rc = encode_map(&map_a, in, 4, dims, strides, box, est, BKC);
    if (rc) return rc;
  }
  {
    const uint64_t K = (uint64_t)d.kh * d.kw * d.cin;
    const uint64_t dims[2] = {K, d.w_img_rows ? (uint64_t)d.w_img_rows * d.n : (uint64_t)d.cout * n_phases};
    const uint64_t strides[1] = {K * 2};
    const uint32_t box[2] = {(uint32_t)BKC, (uint32_t)(BN / 2)};
    const uint32_t est[2] = {1, 1};
    int rc = encode_map(&map_b, weight, 2, dims, strides, box, est, BKC);
    if (rc) return rc;
  }
  p.w_img_rows = d.w_img_rows;
  p.out_pad = d.out_pad; p.out_pad_trim = d.out_pad_trim; p.bias_img_stride = d.bias_img_stride;
  auto kern = conv_tc_halo2_kernel<BN, BKC, MT>;
  static bool attr_set[kMaxDevices] = {false};
  if (!attr_set[cur_device()]) {
    int rc = cuda_check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kMaxDynamic),
                        "cudaFuncSetAttribute(conv_tc_halo2)");
    if (rc) return rc;
    attr_set[cur_device()] = true;
  }
  const int pairs = num_sms() / 2;
  const int clusters = p.num_tiles < pairs ? p.num_tiles : pairs;
  float* ssum = stats;
  float* ssq = stats ? stats + (int64_t)d.n * p.stats_partials * d.cout : nullptr;
  kern<<<2 * clusters, kThreads, smem_total, st>>>(map_a, map_b, reinterpret_cast<__nv_bfloat16*>(out), bias, noise, ssum, ssq, p);
  PPST_LAUNCH_CHECK("conv_tc_halo2_kernel");
  return 0;
}

// 0: plain kernel; else MT (1 or 2) of the halo kernel for this layer
static int halo_mode(const ppst_conv_desc& d, int bn) {
  static const int env_halo = [] { const char* e = getenv("PPST_CONV_TC_HALO"); return e ? atoi(e) : 1; }();
  static const int env_mt = [] { const char* e = getenv("PPST_CONV_TC_MT"); return e ? atoi(e) : 0; }();
  const char* e_s2 = getenv("PPST_CONV_TC_HALO_S2");   // read per call: tests toggle it.  0: plain kernel, 1: MT = 1, 2: MT <= 2, 3: MT <= 4
  const int env_s2 = e_s2 ? atoi(e_s2) : 3;
  const bool s2 = d.stride == 2 && env_s2 > 0 && d.phases_y * d.phases_x == 1 && d.kh >= 2 && d.kh <= 6;
  if (!env_halo || !(s2 || (d.stride == 1 && d.kh >= 2 && d.kh <= 4)) || d.split_terms != 0 || d.out_dtype != PPST_BF16)
    return 0;
  int th, tw;
  pick_tile(d.out_h, d.out_w, th, tw);
  if (tw < 8 || tw * th != BM) return 0;       // tap row offsets must be whole swizzle atoms
  if (s2) {
    // a slot holds two boxes of MT*th + (kh-1)/2 rows; MT = 2 when two such slots and a B ring still fit
    const int64_t tiles2 = (int64_t)d.n * ceil_div(d.out_h, th) * ceil_div(d.out_w, tw) * (d.cout / bn);
    if (env_s2 >= 3 && bn <= 64 && ceil_div(d.out_h, th) >= 4 && tiles2 >= 16LL * num_sms()) return 4;
    return (env_s2 >= 2 && bn <= 128 && ceil_div(d.out_h, th) >= 2 && tiles2 >= 4LL * num_sms()) ? 2 : 1;
  }
  if (bn > 128) return 1;
  if (env_mt == 1 || env_mt == 2 || (env_mt == 4 && bn <= 64)) return env_mt;
  const int64_t tiles = (int64_t)d.n * ceil_div(d.out_h, th) * ceil_div(d.out_w, tw) * d.phases_y * d.phases_x * (d.cout / bn);
  static const int env_mt4 = [] { const char* e = getenv("PPST_CONV_TC_MT4"); return e ? atoi(e) : 1; }();
  if (env_mt4 && bn <= 64 && ceil_div(d.out_h, th) >= 4 && tiles >= 16LL * num_sms()) return 4;   // 512-pixel super-tiles
  return (ceil_div(d.out_h, th) >= 2 && tiles >= 4LL * num_sms()) ? 2 : 1;
}

// CTA-pair (cta_group::2) form of the halo kernel?  PPST_CONV_TC_2CTA: 0 never, 1 cout-128 layers only, 2 (default) also the
// cout % 256 == 0 layers (measured on B200, 8 x 256^2 x 256 -> 256: 1348 -> 1551 TF/s; 8 x 128^2 x 512 -> 512: 1436 -> 1644).
static bool pair_mode(const ppst_conv_desc& d, int bn) {
  const char* e = getenv("PPST_CONV_TC_2CTA");        // read per call: tests toggle it
  const int env = e ? atoi(e) : 2;
  if (env == 0 || (bn != 128 && bn != 256) || (bn == 256 && env < 2) || d.stride != 1) return false;
  int th, tw;
  pick_tile(d.out_h, d.out_w, th, tw);
  const int tiles_x = ceil_div(d.out_w, tw);
  const int64_t tiles = (int64_t)d.n * ceil_div(d.out_h, th) * tiles_x * d.phases_y * d.phases_x * (d.cout / bn);
  static const int env_min = [] { const char* e = getenv("PPST_CONV_TC_PAIR_MIN"); return e ? atoi(e) : 2; }();
  return tiles_x % 2 == 0 && tiles >= (int64_t)env_min * num_sms();
}

}  // namespace tc

// MN-major weight-gradient launch (wgrad.cu): partials f32 [n * k_split][taps][cout][cin].
int tc_wgrad_mn(float* partials, const void* dy, const void* x, const ppst_wgrad_desc& d, int k_split, cudaStream_t st) {
  tc::WgParams p;
  std::memset(&p, 0, sizeof(p));
  p.n = d.n; p.out_h = d.out_h; p.out_w = d.out_w; p.cout = d.cout; p.cin = d.cin; p.kh = d.kh; p.kw = d.kw;
  p.stride = d.stride; p.off_y = d.off_y; p.off_x = d.off_x;
  p.tiles_x = ceil_div(d.out_w, 8);
  p.patches = p.tiles_x * ceil_div(d.out_h, 8);
  p.chunk = ceil_div(p.patches, k_split);
  PPST_REQUIRE(ceil_div(p.patches, p.chunk) == k_split, PPST_ERR_ARG, "conv_wgrad: a pixel chunk would be empty");
  p.k_split = k_split;
  p.m_tiles = ceil_div(d.cout, 128);
  const int bn = d.cin > 128 ? 256 : d.cin > 64 ? 128 : 64;
  p.n_tiles = ceil_div(d.cin, bn);
  const int64_t tiles = (int64_t)d.n * k_split * p.m_tiles * p.n_tiles * d.kh * d.kw;
  PPST_REQUIRE(tiles < (1LL << 31), PPST_ERR_ARG, "conv_wgrad: too many tiles");
  p.num_tiles = (int)tiles;
  const int stage_bytes = (2 + bn / 64) * 8192;
  p.stages = (232448 - 1024 - 2048) / stage_bytes;
  if (p.stages > tc::kMaxStages) p.stages = tc::kMaxStages;
  const int smem_total = p.stages * stage_bytes + (2 * tc::kMaxStages + 4) * 8 + 16 + 1024;
  CUtensorMap map_dy, map_x;
  {
    const uint64_t dims[4] = {(uint64_t)d.cout, (uint64_t)d.out_w, (uint64_t)d.out_h, (uint64_t)d.n};
    const uint64_t strides[3] = {(uint64_t)d.cout * 2, (uint64_t)d.out_w * d.cout * 2, (uint64_t)d.out_h * d.out_w * d.cout * 2};
    const uint32_t box[4] = {64, 8, 8, 1};
    const uint32_t est[4] = {1, 1, 1, 1};
    int rc = tc::encode_map(&map_dy, dy, 4, dims, strides, box, est, 64);
    if (rc) return rc;
  }
  {
    const uint64_t dims[4] = {(uint64_t)d.cin, (uint64_t)d.in_w, (uint64_t)d.in_h, (uint64_t)d.n};
    const uint64_t strides[3] = {(uint64_t)d.cin * 2, (uint64_t)d.in_w * d.cin * 2, (uint64_t)d.in_h * d.in_w * d.cin * 2};
    const uint32_t box[4] = {64, (uint32_t)(8 * d.stride), (uint32_t)(8 * d.stride), 1};
    const uint32_t est[4] = {1, (uint32_t)d.stride, (uint32_t)d.stride, 1};
    int rc = tc::encode_map(&map_x, x, 4, dims, strides, box, est, 64);
    if (rc) return rc;
  }
  const int grid = p.num_tiles < tc::num_sms() ? p.num_tiles : tc::num_sms();
#define PPST_WG(BN_)                                                                                              \
  do {                                                                                                            \
    auto kern = tc::conv_wgrad_mn_kernel<BN_>;                                                                    \
    static bool attr_set[tc::kMaxDevices] = {false};                                                              \
    if (!attr_set[tc::cur_device()]) {                                                                            \
      int rc = cuda_check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 1024), \
                          "cudaFuncSetAttribute(conv_wgrad_mn)");                                                 \
      if (rc) return rc;                                                                                          \
      attr_set[tc::cur_device()] = true;                                                                          \
    }                                                                                                             \
    kern<<<grid, 192, smem_total, st>>>(map_dy, map_x, partials, p);                                              \
  } while (0)
  if (bn == 256) PPST_WG(256);
  else if (bn == 128) PPST_WG(128);
  else PPST_WG(64);
#undef PPST_WG
  PPST_LAUNCH_CHECK("conv_wgrad_mn_kernel");
  return 0;
}
}  // namespace ppst

extern "C" int ppst_conv_tc_available(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return (major == 10 && ppst::tc::get_encode() != nullptr) ? 1 : 0;
}

extern "C" int64_t ppst_conv_tc_stats_elems(const ppst_conv_desc* d) {
  if (!d) return 0;
  int th, tw;
  ppst::tc::pick_tile(d->out_h, d->out_w, th, tw);
  const int64_t partials = (int64_t)ppst::ceil_div(d->out_h, th) * ppst::ceil_div(d->out_w, tw) * d->phases_y * d->phases_x;
  return 2 * (int64_t)d->n * partials * d->cout;
}

extern "C" int ppst_conv_tc(void* out, const void* in, const void* weight, const float* bias, const float* noise,
                            float* stats, const ppst_conv_desc* dp, void* stream) {
  using namespace ppst;
  PPST_REQUIRE(out && in && weight && dp, PPST_ERR_ARG, "conv_tc: null argument");
  const ppst_conv_desc d = *dp;
  int rc = validate_conv_desc(d, "conv_tc");
  if (rc) return rc;
  PPST_REQUIRE(d.in_dtype == PPST_BF16 && (d.out_dtype == PPST_BF16 || d.out_dtype == PPST_F32), PPST_ERR_DTYPE,
               "conv_tc: bf16 in, bf16 or f32 out");
  PPST_REQUIRE(d.split_terms == 0 || (d.split_terms == 3 && d.kh == 1 && d.kw == 1 && d.cin % 128 == 0), PPST_ERR_UNSUPPORTED,
               "conv_tc: split_terms=%d needs a 1x1 kernel and cin (= 2 x segment) a multiple of 128", d.split_terms);
  PPST_REQUIRE(d.w_img_rows == 0 || (d.phases_y * d.phases_x == 1 && d.w_img_rows >= d.cout), PPST_ERR_ARG,
               "conv_tc: per-image weights need a single phase and w_img_rows >= cout");
  PPST_REQUIRE(d.out_dtype == PPST_BF16 || stats == nullptr, PPST_ERR_UNSUPPORTED, "conv_tc: statistics need bf16 output");
  PPST_REQUIRE(d.out_pad >= 0 && d.out_pad <= 8 && d.out_pad_trim >= 0 && d.out_pad_trim <= d.out_pad, PPST_ERR_ARG,
               "conv_tc: out_pad %d / out_pad_trim %d", d.out_pad, d.out_pad_trim);
  PPST_REQUIRE(d.bias_img_stride == 0 || (bias != nullptr && d.bias_img_stride >= d.cout && d.cout <= tc::kBiasSmem),
               PPST_ERR_ARG, "conv_tc: per-sample bias needs bias_img_stride >= cout and cout <= %d", tc::kBiasSmem);
  PPST_REQUIRE(d.cin % 32 == 0, PPST_ERR_UNSUPPORTED, "conv_tc: cin=%d must be a multiple of 32", d.cin);
  PPST_REQUIRE(d.cout % 32 == 0, PPST_ERR_UNSUPPORTED, "conv_tc: cout=%d must be a multiple of 32", d.cout);
  PPST_REQUIRE(d.out_dtype == PPST_F32 || d.cout <= tc::kBiasSmem, PPST_ERR_UNSUPPORTED,
               "conv_tc: bf16 output supports cout <= %d (got %d)", tc::kBiasSmem, d.cout);
  PPST_REQUIRE(d.stride == 1 || d.stride == 2, PPST_ERR_UNSUPPORTED, "conv_tc: stride %d", d.stride);
  PPST_REQUIRE(d.in_row_pitch % 8 == 0 && d.in_img_pitch % 8 == 0, PPST_ERR_UNSUPPORTED, "conv_tc: input pitches must be 16-byte multiples");
  PPST_REQUIRE(d.cout_stride % 8 == 0 && d.c_off % 8 == 0, PPST_ERR_UNSUPPORTED, "conv_tc: output channel slice must be 16-byte aligned");
  PPST_REQUIRE(((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(weight) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
               PPST_ERR_ARG, "conv_tc: buffers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const bool k64 = (d.cin % 64 == 0);
#define PPST_TC(BN_, BKC_)                                                                   \
  do {                                                                                       \
    const int hm = tc::halo_mode(d, BN_);                                                    \
    if (hm >= 1 && BKC_ == 64 && tc::pair_mode(d, BN_)) {                                    \
      if (hm == 2 && BN_ == 128)                                                             \
        return tc::launch_halo2<(BN_ == 128 ? 128 : 256), 64, (BN_ == 128 ? 2 : 1)>(d, out, in, weight, bias, noise, stats, st); \
      return tc::launch_halo2<(BN_ == 128 ? 128 : 256), 64, 1>(d, out, in, weight, bias, noise, stats, st); \
    }                                                                                        \
    int hrc = tc::kHaloNoFit;                                                                \
    if (hm == 4 && BN_ <= 64)                                                                \
      hrc = tc::launch_halo<BN_, BKC_, (BN_ <= 64 ? 4 : 1)>(d, out, in, weight, bias, noise, stats, st); \
    if (hrc == tc::kHaloNoFit && hm >= 2 && BN_ <= 128)                                      \
      hrc = tc::launch_halo<BN_, BKC_, (BN_ <= 128 ? 2 : 1)>(d, out, in, weight, bias, noise, stats, st); \
    if (hrc == tc::kHaloNoFit && hm >= 1)                                                    \
      hrc = tc::launch_halo<BN_, BKC_, 1>(d, out, in, weight, bias, noise, stats, st);       \
    if (hrc != tc::kHaloNoFit) return hrc;                                                   \
    if (d.out_dtype == PPST_F32)                                                             \
      return tc::launch<BN_, BKC_, true>(d, out, in, weight, bias, noise, stats, st);        \
    return tc::launch<BN_, BKC_, false>(d, out, in, weight, bias, noise, stats, st);         \
  } while (0)
  if (k64) {
    if (d.cout % 256 == 0) PPST_TC(256, 64);
    if (d.cout % 128 == 0) PPST_TC(128, 64);
    if (d.cout % 64 == 0) PPST_TC(64, 64);
    PPST_TC(32, 64);
  } else {
    if (d.cout % 64 == 0) PPST_TC(64, 32);
    PPST_TC(32, 32);
  }
#undef PPST_TC
}
