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
