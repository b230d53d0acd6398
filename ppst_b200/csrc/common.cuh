// Shared device/host helpers for the ppst_b200 sm_100a kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/ppst_b200.h"

namespace ppst {

// thread-local last error text (returned by ppst_last_error)
void set_error(const char* fmt, ...);
int cuda_check(cudaError_t e, const char* what);

#define PPST_REQUIRE(cond, code, ...)        \
  do {                                       \
    if (!(cond)) {                           \
      ::ppst::set_error(__VA_ARGS__);        \
      return (code);                         \
    }                                        \
  } while (0)

#define PPST_LAUNCH_CHECK(what)                                  \
  do {                                                           \
    int _rc = ::ppst::cuda_check(cudaGetLastError(), (what));    \
    if (_rc) return _rc;                                         \
  } while (0)

template <typename T> struct Cvt;
template <> struct Cvt<float> {
  __device__ __forceinline__ static float to_f(float v) { return v; }
  __device__ __forceinline__ static float from_f(float v) { return v; }
};
template <> struct Cvt<__half> {
  __device__ __forceinline__ static float to_f(__half v) { return __half2float(v); }
  __device__ __forceinline__ static __half from_f(float v) { return __float2half_rn(v); }
};
template <> struct Cvt<__nv_bfloat16> {
  __device__ __forceinline__ static float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
  __device__ __forceinline__ static __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
};

// 16-byte vector of T (4 x f32 or 8 x 16-bit)
template <typename T> struct Vec16 {
  static constexpr int N = 16 / sizeof(T);
  union {
    uint4 raw;
    T v[N];
  };
};

template <typename T>
__device__ __forceinline__ Vec16<T> ld_vec16(const T* p) {
  Vec16<T> r;
  r.raw = *reinterpret_cast<const uint4*>(p);
  return r;
}
// streaming (read-once) variant: bypass L1 allocation
template <typename T>
__device__ __forceinline__ Vec16<T> ld_vec16_stream(const T* p) {
  Vec16<T> r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.raw.x), "=r"(r.raw.y), "=r"(r.raw.z), "=r"(r.raw.w)
               : "l"(p));
  return r;
}
template <typename T>
__device__ __forceinline__ void st_vec16(T* p, const Vec16<T>& r) {
  *reinterpret_cast<uint4*>(p) = r.raw;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// warp-level transposing reduction: on entry lane L holds v[c] for columns c = 0..31 of ITS row;
// on exit v[0] of lane L = sum over the 32 lanes (rows) of column L.  31 shuffles.
__device__ __forceinline__ float warp_col_reduce(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool hi = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = hi ? v[i] : v[i + off];
      const float keep = hi ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}


inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

inline size_t dtype_size(int dtype) { return dtype == PPST_F32 ? 4 : 2; }

// Dispatch helper on the I/O dtype enum.
#define PPST_DISPATCH_DTYPE(dtype, T, ...)                             \
  switch (dtype) {                                                     \
    case PPST_F32: { using T = float; __VA_ARGS__; break; }            \
    case PPST_F16: { using T = __half; __VA_ARGS__; break; }           \
    case PPST_BF16: { using T = __nv_bfloat16; __VA_ARGS__; break; }   \
    default: ::ppst::set_error("unsupported dtype %d", (int)(dtype));  \
             return PPST_ERR_DTYPE;                                    \
  }

}  // namespace ppst
