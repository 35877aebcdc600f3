// Micro-benchmark: scalar FFMA vs packed fma.rn.f32x2 throughput under EM-like register pressure.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long pack(float lo, float hi) {
  return ((unsigned long long)__float_as_uint(hi) << 32) | __float_as_uint(lo);
}
// 32 accumulators, 32 distinct multiplicands, 4 x values: every FFMA has 3 distinct registers
__global__ void __launch_bounds__(128) k_scalar(float* out, float s) {
  float acc[32], p[32], x[4];
  for (int i = 0; i < 32; ++i) { acc[i] = threadIdx.x + i; p[i] = s * (1.f + i * 1e-3f); }
  for (int i = 0; i < 4; ++i) x[i] = s * 0.5f + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = fmaf(x[i & 3], p[i], acc[i]);
  }
  float t = 0; for (int i = 0; i < 32; ++i) t += acc[i];
  if (t == 12345.f) out[0] = t;
}
__global__ void __launch_bounds__(128) k_packed(float* out, float s) {
  unsigned long long acc[16], p[16], x[4];
  for (int i = 0; i < 16; ++i) { acc[i] = pack(threadIdx.x + i, i); p[i] = pack(s * (1.f + i * 1e-3f), s * (1.f - i * 1e-3f)); }
  for (int i = 0; i < 4; ++i) x[i] = pack(s * 0.5f + i, s * 0.5f + i);
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = ffma2(x[i & 3], p[i], acc[i]);
  }
  unsigned long long t = 0; for (int i = 0; i < 16; ++i) t ^= acc[i];
  if (t == 12345ull) out[0] = 1.f;
}
// chains with register reuse (the roofline micro-benchmark's shape): acc = acc * m + c
__global__ void __launch_bounds__(128) k_scalar_reuse(float* out, float s) {
  float acc[32]; const float m = 0.999f + s * 1e-9f, c = 1e-3f * s;
  for (int i = 0; i < 32; ++i) acc[i] = threadIdx.x + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = fmaf(acc[i], m, c);
  }
  float t = 0; for (int i = 0; i < 32; ++i) t += acc[i];
  if (t == 12345.f) out[0] = t;
}
template <typename F> float run(F f, int blocks) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(blocks); cudaDeviceSynchronize();
  cudaEventRecord(a); f(blocks); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
int main() {
  float* out; cudaMalloc(&out, 16);
  for (int occ : {4, 8, 16}) {
    const int blocks = 148 * occ;
    const double flops = (double)blocks * 128 * ITERS * 32 * 2;
    float t1 = run([&](int b) { k_scalar<<<b, 128>>>(out, 1.f); }, blocks);
    float t2 = run([&](int b) { k_packed<<<b, 128>>>(out, 1.f); }, blocks);
    float t3 = run([&](int b) { k_scalar_reuse<<<b, 128>>>(out, 1.f); }, blocks);
    printf("CTAs/SM %2d: scalar 3-reg %.1f TFLOP/s | packed f32x2 %.1f TFLOP/s | scalar reuse %.1f TFLOP/s\n", occ,
           flops / t1 / 1e9, flops / t2 / 1e9, flops / t3 / 1e9);
  }
  return 0;
}
