// Micro-benchmark 2: FFMA2 with the depthwise-stencil operand pattern (12 accumulators, 9 taps, 6 inputs: three distinct
// 64-bit register operands per instruction) and with the bf16x2 -> f32x2 unpack in the loop.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ void ffma2(u64& d, u64 a, u64 b) { asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b)); }
__device__ __forceinline__ u64 unpack(unsigned v) { return (static_cast<u64>(v & 0xFFFF0000u) << 32) | (v << 16); }
__device__ __forceinline__ u64 unpack_prmt(unsigned v) {
  unsigned lo, hi;
  asm volatile("prmt.b32 %0, %1, 0, 0x1044;" : "=r"(lo) : "r"(v));
  asm volatile("prmt.b32 %0, %1, 0, 0x3244;" : "=r"(hi) : "r"(v));
  return (static_cast<u64>(hi) << 32) | lo;
}

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(float* out, const unsigned* in, int iters) {
  u64 acc[12], w[9], x[6];
  unsigned raw[6];
  for (int i = 0; i < 12; ++i) acc[i] = 0ull;
  for (int i = 0; i < 9; ++i) w[i] = (static_cast<u64>(__float_as_uint(0.5f + i)) << 32) | __float_as_uint(0.25f * i);
  for (int i = 0; i < 6; ++i) { raw[i] = in[threadIdx.x + i * 512]; x[i] = unpack(raw[i]); }
  for (int it = 0; it < iters; ++it) {
    if (MODE >= 1) {
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        raw[i] = raw[i] * 3u + 1u;   // something cheap on the ALU/IMAD pipe to keep the unpack live
        x[i] = MODE == 2 ? unpack_prmt(raw[i]) : unpack(raw[i]);
      }
    }
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int oc = 0; oc < 4; ++oc)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) ffma2(acc[dy * 4 + oc], w[dy * 3 + dx], x[oc + dx]);
  }
  float s = 0.f;
  for (int i = 0; i < 12; ++i) s += __uint_as_float(static_cast<unsigned>(acc[i])) + __uint_as_float(static_cast<unsigned>(acc[i] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int threads) {
  float* out; unsigned* in;
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&in, 512 * 6 * 4); cudaMemset(in, 0x3f, 512 * 6 * 4);
  const int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148, threads>>>(out, in, 100);
  cudaEventRecord(e0);
  k<MODE><<<148, threads>>>(out, in, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double cyc = ms * 1e-3 * 1.965e9;
  printf("%-46s thr %3d %8.3f ms  %6.2f clk per 36-FFMA2 row per warp-slot, %6.1f FMA/clk/SM\n", name, threads, ms,
         cyc / iters / (threads / 128.0), 72.0 * iters * threads / cyc);
}
int main() {
  run<0>("FFMA2 stencil pattern (12 acc, 9 w, 6 x)", 512);
  run<0>("FFMA2 stencil pattern (12 acc, 9 w, 6 x)", 256);
  run<1>("  + unpack via LOP3/SHL", 512);
  run<1>("  + unpack via LOP3/SHL", 256);
  run<2>("  + unpack via PRMT", 512);
  run<2>("  + unpack via PRMT", 256);
  return 0;
}
