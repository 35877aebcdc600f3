// Library identification and per-thread error string of the stminer_b200 C ABI.
#include <stdarg.h>

#include "stm_common.cuh"

namespace stm {

static thread_local char g_last_error[512] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

}  // namespace stm

extern "C" int stm_version(void) { return 100; /* 0.1.0 */ }

extern "C" const char* stm_last_error_string(void) { return stm::g_last_error; }

// ---------------------------------------------------------------------------------------------
// FP32 FMA-pipe peak micro-benchmark: the measured denominator of the EM kernel's roofline
// (MEASURED_PEAKS.json has no FP32 figure).  Each thread runs `iters` rounds of 16 independent
// FFMA chains; FLOPs = blocks * 256 * iters * 16 * 2.
// ---------------------------------------------------------------------------------------------
namespace stm {
__global__ void __launch_bounds__(256) fma_peak_kernel(int iters, float seed, float* out) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = seed + i + threadIdx.x;
  const float m = 0.999f + seed * 1e-9f, c = 1e-3f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], m, c);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  if (s == 123.456f) out[0] = s;  // keep the chains alive
}
}  // namespace stm

extern "C" int stm_bench_fma(int32_t blocks, int32_t iters, float* out, void* stream) {
  using namespace stm;
  STM_CHECK_ARG(blocks > 0 && iters > 0 && out, "need blocks > 0, iters > 0, out != NULL");
  fma_peak_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(iters, 1.0f, out);
  STM_CHECK_CUDA(cudaGetLastError());
  return STM_OK;
}
