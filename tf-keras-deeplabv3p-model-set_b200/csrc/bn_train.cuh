// bn_train.cuh — training-mode batch statistics and normalisation of CustomBatchNormalization /
// SyncBatchNormalization (reference deeplabv3p/models/layers.py:63-70; SURVEY.md §8(c) "SyncBN training"):
//   per replica  sum_x[c], sum_x2[c] over the (N,H,W) rows, row count          (bn_stats_*_kernel)
//   all-reduce(SUM) of [sum_x | sum_x2 | count] over the replicas              (NCCL, by the caller: sharding.py)
//   mean = sum_x / n, var = sum_x2 / n - mean^2 (biased), y = (x - mean) * gamma * rsqrt(var + eps) + beta [ReLU]
// The statistics are the first half of the cfg-5 training step's exchange; the backward kernels are not built yet.
// Memory bound: x is read once per kernel, 4-byte bf16x2 loads (one 128-byte line per warp and row), fixed reduction
// tree (deterministic: no atomics).
#pragma once

#include "sm100_prims.cuh"

namespace dlv3p {

constexpr int kBnBands = 64;   // row bands of the first reduction stage

// grid (ceil(C/64), kBnBands), block 256: warp w of band b strides over the band's rows; lane = channels 2l, 2l+1.
// partial: [kBnBands][2][C] fp32
__global__ void __launch_bounds__(256) bn_stats_partial_kernel(const __nv_bfloat16* __restrict__ x, long long M, int C,
                                                               float* __restrict__ partial) {
  __shared__ float s_red[8][2][64];
  const int chunk = blockIdx.x, band = blockIdx.y;
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const int c0 = chunk * 64 + lane * 2;
  const long long rows_per_band = (M + kBnBands - 1) / kBnBands;
  const long long r0 = band * rows_per_band, r1 = min(M, r0 + rows_per_band);
  float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
  if (c0 < C) {
    const __nv_bfloat16* xb = x + c0;
    for (long long r = r0 + wp; r < r1; r += 8) {
      const uint32_t v = __ldg(reinterpret_cast<const unsigned int*>(xb + r * C));
      const float a = bf16_lo(v), b = bf16_hi(v);
      s0 += a; s1 += b;
      q0 = fmaf(a, a, q0); q1 = fmaf(b, b, q1);
    }
  }
  s_red[wp][0][lane * 2] = s0; s_red[wp][0][lane * 2 + 1] = s1;
  s_red[wp][1][lane * 2] = q0; s_red[wp][1][lane * 2 + 1] = q1;
  __syncthreads();
  if (threadIdx.x < 128) {
    const int which = threadIdx.x >> 6, j = threadIdx.x & 63;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += s_red[w][which][j];
    const int c = chunk * 64 + j;
    if (c < C) partial[(static_cast<size_t>(band) * 2 + which) * C + c] = s;
  }
}
// stats: [2*C + 1] fp32 = sum_x | sum_x2 | row count
__global__ void bn_stats_final_kernel(const float* __restrict__ partial, long long M, int C, float* __restrict__ stats) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 2 * C) {
    const int which = i / C, c = i - which * C;
    float s = 0.f;
    for (int b = 0; b < kBnBands; ++b) s += partial[(static_cast<size_t>(b) * 2 + which) * C + c];
    stats[i] = s;
  }
  if (i == 0) stats[2 * C] = static_cast<float>(M);
}
// y = (x - mean) * gamma * rsqrt(var + eps) + beta, statistics read from DEVICE memory (no host round trip after the
// all-reduce).  One thread = 8 channels of one row (16-byte accesses).
__global__ void __launch_bounds__(256) bn_apply_kernel(const __nv_bfloat16* __restrict__ x, long long M, int C, const float* __restrict__ stats,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int relu,
                                                       __nv_bfloat16* __restrict__ y) {
  const int vecs = C >> 3;
  const long long total = M * vecs;
  const float inv_n = 1.0f / stats[2 * C];
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total; idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int vec = static_cast<int>(idx % vecs);
    const uint4 raw = ldg_nc_v4(x + idx * 8);
    const uint32_t in[4] = {raw.x, raw.y, raw.z, raw.w};
    uint32_t out[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float v[2] = {bf16_lo(in[k]), bf16_hi(in[k])};
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = vec * 8 + k * 2 + h;
        const float mean = stats[c] * inv_n;
        const float var = fmaxf(stats[C + c] * inv_n - mean * mean, 0.0f);
        const float sc = gamma[c] * rsqrtf(var + eps);
        v[h] = (v[h] - mean) * sc + beta[c];
        if (relu) v[h] = fmaxf(v[h], 0.0f);
      }
      out[k] = pack_bf16x2(v[0], v[1]);
    }
    stg_v4(y + idx * 8, make_uint4(out[0], out[1], out[2], out[3]));
  }
}


// ------------------------------------------------------------------------------------------------ vectorised variants
// Column reductions at HBM speed: one thread = 8 channels (16-byte loads), one warp = 256 contiguous channels of a row, the
// 8 warps of a block take 8 rows at a time; grid (ceil(C/256), bands) with bands chosen so that ~4 blocks per SM are in
// flight whatever C is.  Fixed reduction order (warps through smem, bands by the final kernel): deterministic.
__host__ __device__ inline int col_bands(int C) {
  const int chunks = (C + 255) / 256;
  int b = (3 * 148 + chunks - 1) / chunks;
  return b < 8 ? 8 : (b > 1024 ? 1024 : b);
}
// floats of scratch the banded column reductions need for `nq` quantities per channel (monotonic in C)
__host__ __device__ inline size_t col_scratch_floats(int C, int nq) {
  return static_cast<size_t>(nq) * (static_cast<size_t>(4 * 148 + 8) * 256 + static_cast<size_t>(264) * (C > 0 ? C : 0));
}
__device__ __forceinline__ void unpack8_bn(const uint4& r, float (&f)[8]) {
  f[0] = bf16_lo(r.x); f[1] = bf16_hi(r.x); f[2] = bf16_lo(r.y); f[3] = bf16_hi(r.y);
  f[4] = bf16_lo(r.z); f[5] = bf16_hi(r.z); f[6] = bf16_lo(r.w); f[7] = bf16_hi(r.w);
}
// shared tail of the column-reduction kernels: acc[16] per thread (8 channels x 2 quantities) -> partial[band][2][C]
__device__ __forceinline__ void col_reduce_store(float (&acc)[16], float (*s_red)[32][17], int chunk, int band, int C, float* __restrict__ partial) {
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 16; ++k) s_red[wp][lane][k] = acc[k];
  __syncthreads();
  for (int e = threadIdx.x; e < 32 * 16; e += 256) {
    const int l = e >> 4, k = e & 15;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += s_red[w][l][k];
    const int c = chunk * 256 + l * 8 + (k & 7);
    if (c < C) partial[(static_cast<size_t>(band) * 2 + (k >> 3)) * C + c] = s;
  }
}
// x dense [M, C], C % 8 == 0.  partial [bands][2][C] = sum x | sum x^2
__global__ void __launch_bounds__(256) bn_stats_vec_kernel(const __nv_bfloat16* __restrict__ x, long long M, int C, int bands, float* __restrict__ partial) {
  __shared__ float s_red[8][32][17];
  const int chunk = blockIdx.x, band = blockIdx.y;
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const int c0 = chunk * 256 + lane * 8;
  const long long rows_per_band = (M + bands - 1) / bands;
  const long long r0 = band * rows_per_band, r1 = min(M, r0 + rows_per_band);
  float acc[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) acc[k] = 0.f;
  if (c0 < C) {
    const __nv_bfloat16* xb = x + c0;
#pragma unroll 4
    for (long long r = r0 + wp; r < r1; r += 8) {
      float v[8];
      unpack8_bn(ldg_nc_v4(xb + r * C), v);
#pragma unroll
      for (int k = 0; k < 8; ++k) { acc[k] += v[k]; acc[8 + k] = fmaf(v[k], v[k], acc[8 + k]); }
    }
  }
  col_reduce_store(acc, s_red, chunk, band, C, partial);
}
// stats[2C+1] from partial[bands][2][C]
__global__ void bn_stats_vec_final_kernel(const float* __restrict__ partial, int bands, long long M, int C, float* __restrict__ stats) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 2 * C) {
    const int which = i / C, c = i - which * C;
    float s = 0.f;
    for (int b = 0; b < bands; ++b) s += partial[(static_cast<size_t>(b) * 2 + which) * C + c];
    stats[i] = s;
  }
  if (i == 0) stats[2 * C] = static_cast<float>(M);
}
// y[row*ldy + c] = x*scale[c] + shift[c] [ReLU] with scale = gamma * rsqrt(var + eps), shift = beta - mean * scale computed ONCE
// per block into shared memory (dynamic smem: 2*C floats); the element loop is pure streaming.
// Idx = unsigned int whenever M * C/8 < 2^32 (always, at the reference's sizes): 64-bit division per element is what bounded the first version
template <typename Idx>
__global__ void __launch_bounds__(256) bn_apply_vec_kernel(const __nv_bfloat16* __restrict__ x, long long M, int C, const float* __restrict__ stats,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int relu,
                                                           __nv_bfloat16* __restrict__ y, long long ldy) {
  extern __shared__ float s_coef[];   // [C] scale | [C] shift
  {
    const float inv_n = 1.0f / stats[2 * C];
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      const float mean = stats[c] * inv_n;
      const float var = fmaxf(stats[C + c] * inv_n - mean * mean, 0.0f);
      const float sc = gamma[c] * rsqrtf(var + eps);
      s_coef[c] = sc;
      s_coef[C + c] = beta[c] - mean * sc;
    }
  }
  __syncthreads();
  const Idx vecs = static_cast<Idx>(C >> 3);
  const Idx total = static_cast<Idx>(M) * vecs;
  for (Idx idx = blockIdx.x * static_cast<Idx>(blockDim.x) + threadIdx.x; idx < total; idx += static_cast<Idx>(gridDim.x) * blockDim.x) {
    const Idx row = idx / vecs;
    const int vec = static_cast<int>(idx - row * vecs);
    float v[8];
    unpack8_bn(ldg_nc_v4(x + static_cast<size_t>(idx) * 8), v);
    const float4 s0 = *reinterpret_cast<const float4*>(s_coef + vec * 8), s1 = *reinterpret_cast<const float4*>(s_coef + vec * 8 + 4);
    const float4 t0 = *reinterpret_cast<const float4*>(s_coef + C + vec * 8), t1 = *reinterpret_cast<const float4*>(s_coef + C + vec * 8 + 4);
    v[0] = fmaf(v[0], s0.x, t0.x); v[1] = fmaf(v[1], s0.y, t0.y); v[2] = fmaf(v[2], s0.z, t0.z); v[3] = fmaf(v[3], s0.w, t0.w);
    v[4] = fmaf(v[4], s1.x, t1.x); v[5] = fmaf(v[5], s1.y, t1.y); v[6] = fmaf(v[6], s1.z, t1.z); v[7] = fmaf(v[7], s1.w, t1.w);
    if (relu) {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = fmaxf(v[k], 0.0f);
    }
    stg_v4(y + static_cast<long long>(row) * ldy + vec * 8, make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7])));
  }
}

}  // namespace dlv3p
