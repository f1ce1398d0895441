// Micro-benchmark: FP32 FMA issue rates on sm_100a — scalar FFMA (3 register operands), packed FFMA2 (fma.rn.f32x2),
// and a 1:1 mix.  Prints FMA/clk/SM.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_rate fma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(float* out, int iters, float a, float b) {
  float x[16];
  unsigned long long y[8];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 0.001f + i;
#pragma unroll
  for (int i = 0; i < 8; ++i) y[i] = (static_cast<unsigned long long>(__float_as_uint(x[2 * i + 1])) << 32) | __float_as_uint(x[2 * i]);
  const unsigned long long ab = (static_cast<unsigned long long>(__float_as_uint(b)) << 32) | __float_as_uint(a);
  const unsigned long long cd = (static_cast<unsigned long long>(__float_as_uint(a)) << 32) | __float_as_uint(b);
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
    } else if (MODE == 1) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(y[i]) : "l"(ab), "l"(cd));
    } else {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(y[i]) : "l"(ab), "l"(cd));
          asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(a), "f"(b));
          if (MODE == 3) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i + 8]) : "f"(a), "f"(b));
        }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) s += __uint_as_float(static_cast<unsigned>(y[i])) + __uint_as_float(static_cast<unsigned>(y[i] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, double fma_per_thread_iter) {
  float* out;
  cudaMalloc(&out, 148 * 512 * 4);
  const int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<MODE><<<148, 512>>>(out, 100, 1.0001f, 0.5f);
  cudaEventRecord(e0);
  k<MODE><<<148, 512>>>(out, iters, 1.0001f, 0.5f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  int clk;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const double fma = fma_per_thread_iter * iters * 512.0;   // per SM
  printf("%-34s %8.3f ms  %7.1f FMA/clk/SM (at %d MHz nominal)\n", name, ms, fma / (ms * 1e-3 * clk * 1e3), clk / 1000);
  cudaFree(out);
}

int main() {
  run<0>("scalar FFMA (3 reg operands)", 64);
  run<1>("packed FFMA2", 128);
  run<2>("FFMA2 + FFMA 1:1", 64 + 32);
  run<3>("FFMA2 + 2 FFMA", 64 + 64);
  return 0;
}
