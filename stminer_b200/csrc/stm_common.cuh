// Shared helpers for the stminer_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/stminer_b200.h"

namespace stm {

void set_last_error(const char* fmt, ...);

#define STM_CHECK_ARG(cond, msg)                                    \
  do {                                                              \
    if (!(cond)) {                                                  \
      ::stm::set_last_error("%s: invalid argument: %s", __func__, msg); \
      return STM_ERR_INVALID_ARG;                                   \
    }                                                               \
  } while (0)

#define STM_CHECK_CUDA(expr)                                                      \
  do {                                                                            \
    cudaError_t _e = (expr);                                                      \
    if (_e != cudaSuccess) {                                                      \
      ::stm::set_last_error("%s: CUDA error %s at %s:%d", __func__,              \
                            cudaGetErrorString(_e), __FILE__, __LINE__);          \
      return STM_ERR_CUDA;                                                        \
    }                                                                             \
  } while (0)

constexpr unsigned FULL_MASK = 0xffffffffu;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Packed FP32x2 arithmetic (Blackwell FFMA2 / FADD2 / FMUL2): one issue slot and one set of 64-bit
// register-file reads for two FP32 lanes.
typedef float2 f2;
__device__ __forceinline__ f2 f2_make(float lo, float hi) { return make_float2(lo, hi); }
__device__ __forceinline__ f2 f2_fma(f2 a, f2 b, f2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ f2 f2_mul(f2 a, f2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ f2 f2_add(f2 a, f2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ f2 f2_sub(f2 a, f2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
  return v;
}
__device__ __forceinline__ long long warp_sum(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
  return v;
}

}  // namespace stm
