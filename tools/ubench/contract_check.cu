// Does ptxas contract mul.rn.f32x2 + add.rn.f32x2 into FFMA2 (single rounding)?  a*b+c with a=b=1+2^-23, c=-(1+2^-22):
// two roundings give 0, a fused multiply-add gives 2^-46.
#include <cstdio>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fmul2(u64 a, u64 b) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(0ull)); return d; }
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__global__ void k(const float* in, float* out) {
  const u64 a = (static_cast<u64>(__float_as_uint(in[0])) << 32) | __float_as_uint(in[0]);
  const u64 c = (static_cast<u64>(__float_as_uint(in[1])) << 32) | __float_as_uint(in[1]);
  const u64 r = fadd2(c, fmul2(a, a));
  out[0] = __uint_as_float(static_cast<unsigned>(r));
  out[1] = __fadd_rn(in[1], __fmul_rn(in[0], in[0]));
  out[2] = fmaf(in[0], in[0], in[1]);
}
int main() {
  float h[2] = {1.0f + 1.1920929e-7f, -(1.0f + 2.3841858e-7f)}, *d, *o, r[3];
  cudaMalloc(&d, 8); cudaMalloc(&o, 12);
  cudaMemcpy(d, h, 8, cudaMemcpyHostToDevice);
  k<<<1, 1>>>(d, o);
  cudaMemcpy(r, o, 12, cudaMemcpyDeviceToHost);
  printf("packed mul.rn+add.rn: %g   scalar __fmul_rn/__fadd_rn: %g   fmaf: %g\n", r[0], r[1], r[2]);
  return 0;
}
