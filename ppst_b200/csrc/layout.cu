// NCHW <-> NHWC transposes (with dtype conversion and channel-slice placement), tiled through shared memory so
// that both the read and the write side are coalesced.
#include "common.cuh"

namespace ppst {

// src [n, c, hw] -> dst [n, hw, dst_c] at channel offset c_off
template <typename TS, typename TD>
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(TD* __restrict__ dst, const TS* __restrict__ src, int64_t c,
                                                           int64_t hw, int64_t dst_c, int64_t c_off) {
  __shared__ float tile[32][33];
  const int64_t n = blockIdx.z;
  const int64_t p0 = (int64_t)blockIdx.x * 32, c0 = (int64_t)blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;           // 32 x 8
  for (int j = ty; j < 32; j += 8) {
    const int64_t ch = c0 + j, p = p0 + tx;
    tile[j][tx] = (ch < c && p < hw) ? Cvt<TS>::to_f(src[(n * c + ch) * hw + p]) : 0.0f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int64_t p = p0 + j, ch = c0 + tx;
    if (p < hw && ch < c) dst[(n * hw + p) * dst_c + c_off + ch] = Cvt<TD>::from_f(tile[tx][j]);
  }
}

// src [n, hw, src_c] (channel offset c_off) -> dst [n, c, hw]
template <typename TS, typename TD>
__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(TD* __restrict__ dst, const TS* __restrict__ src, int64_t c,
                                                           int64_t hw, int64_t src_c, int64_t c_off) {
  __shared__ float tile[32][33];
  const int64_t n = blockIdx.z;
  const int64_t p0 = (int64_t)blockIdx.x * 32, c0 = (int64_t)blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int j = ty; j < 32; j += 8) {
    const int64_t p = p0 + j, ch = c0 + tx;
    tile[j][tx] = (p < hw && ch < c) ? Cvt<TS>::to_f(src[(n * hw + p) * src_c + c_off + ch]) : 0.0f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int64_t ch = c0 + j, p = p0 + tx;
    if (ch < c && p < hw) dst[(n * c + ch) * hw + p] = Cvt<TD>::from_f(tile[tx][j]);
  }
}

#define PPST_DISPATCH2(sd, dd, TS, TD, ...)                                        \
  PPST_DISPATCH_DTYPE(sd, TS, PPST_DISPATCH_DTYPE(dd, TD, __VA_ARGS__))

}  // namespace ppst

extern "C" int ppst_nchw_to_nhwc(void* dst, const void* src, int64_t n, int64_t c, int64_t hw, int64_t dst_c,
                                 int64_t c_off, int src_dtype, int dst_dtype, void* stream) {
  using namespace ppst;
  PPST_REQUIRE(dst && src && n > 0 && c > 0 && hw > 0 && c_off >= 0 && c_off + c <= dst_c, PPST_ERR_ARG,
               "nchw_to_nhwc: bad args");
  PPST_REQUIRE(n <= 65535 && (c + 31) / 32 <= 65535, PPST_ERR_ARG, "nchw_to_nhwc: batch/channel count too large");
  dim3 grid(ceil_div(hw, 32), ceil_div(c, 32), (unsigned)n);
  cudaStream_t st = (cudaStream_t)stream;
  PPST_DISPATCH2(src_dtype, dst_dtype, TS, TD,
                 nchw_to_nhwc_kernel<TS, TD><<<grid, 256, 0, st>>>((TD*)dst, (const TS*)src, c, hw, dst_c, c_off));
  PPST_LAUNCH_CHECK("nchw_to_nhwc_kernel");
  return 0;
}

extern "C" int ppst_nhwc_to_nchw(void* dst, const void* src, int64_t n, int64_t c, int64_t hw, int64_t src_c,
                                 int64_t c_off, int src_dtype, int dst_dtype, void* stream) {
  using namespace ppst;
  PPST_REQUIRE(dst && src && n > 0 && c > 0 && hw > 0 && c_off >= 0 && c_off + c <= src_c, PPST_ERR_ARG,
               "nhwc_to_nchw: bad args");
  PPST_REQUIRE(n <= 65535 && (c + 31) / 32 <= 65535, PPST_ERR_ARG, "nhwc_to_nchw: batch/channel count too large");
  dim3 grid(ceil_div(hw, 32), ceil_div(c, 32), (unsigned)n);
  cudaStream_t st = (cudaStream_t)stream;
  PPST_DISPATCH2(src_dtype, dst_dtype, TS, TD,
                 nhwc_to_nchw_kernel<TS, TD><<<grid, 256, 0, st>>>((TD*)dst, (const TS*)src, c, hw, src_c, c_off));
  PPST_LAUNCH_CHECK("nhwc_to_nchw_kernel");
  return 0;
}
