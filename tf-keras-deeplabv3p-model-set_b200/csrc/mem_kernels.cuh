// mem_kernels.cuh — the memory-bound kernels of the head: 128-bit vectorised NHWC accesses,
// 8 bf16 channels per thread, fp32 arithmetic.
//   aspp_dw_pool_kernel   three dilated depthwise 3x3 (+BN+ReLU) of ASPP in ONE pass over x, plus the
//                         per-channel partial sums of the image-pooling branch   (layers.py:132, :146-153 -> :100-104)
//   pool_proj_kernel      image_pooling 1x1 + BN + ReLU, folded through concat_projection into a per-image
//                         bias vector (a 1x1 -> h x w bilinear resize is a broadcast)  (layers.py:132-138, :155-159)
//   depthwise3x3_kernel   generic dilated depthwise 3x3 + BN + ReLU (standalone operator / unfused path)
//   resize_bilinear_kernel  tf.image.resize bilinear, half-pixel centres (layers.py:48-50, :207)
//   resize_argmax_*       pred_resize fused with the host argmax (model.py:76 + deeplab.py:99)
//   resize_dense_kernel   pred_resize (+ Softmax) materialised, for callers that want the reference output
#pragma once

#include <cuda_fp16.h>

#include "sm100_prims.cuh"

namespace dlv3p {

__device__ __forceinline__ void unpack8(const uint4& r, float (&f)[8]) {
  f[0] = bf16_lo(r.x); f[1] = bf16_hi(r.x);
  f[2] = bf16_lo(r.y); f[3] = bf16_hi(r.y);
  f[4] = bf16_lo(r.z); f[5] = bf16_hi(r.z);
  f[6] = bf16_lo(r.w); f[7] = bf16_hi(r.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}

// ------------------------------------------------------------------------------------------------
__global__ void cast_f32_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, size_t n8) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n8;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(in) + 2 * i);
    const float4 b = __ldg(reinterpret_cast<const float4*>(in) + 2 * i + 1);
    stg_v4(out + 8 * i, make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w)));
  }
}
__global__ void cast_f16_bf16_kernel(const __half* __restrict__ in, __nv_bfloat16* __restrict__ out, size_t n8) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n8;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const uint4 r = ldg_nc_v4(in + 8 * i);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const __half2 h = *reinterpret_cast<const __half2*>(&w[k]);
      const float2 f = __half22float2(h);
      o[k] = pack_bf16x2(f.x, f.y);
    }
    stg_v4(out + 8 * i, make_uint4(o[0], o[1], o[2], o[3]));
  }
}
__global__ void cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, size_t n8) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n8;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float f[8];
    unpack8(ldg_nc_v4(in + 8 * i), f);
    reinterpret_cast<float4*>(out)[2 * i] = make_float4(f[0], f[1], f[2], f[3]);
    reinterpret_cast<float4*>(out)[2 * i + 1] = make_float4(f[4], f[5], f[6], f[7]);
  }
}

// ------------------------------------------------------------------------------------------------
// ASPP: the three dilated depthwise 3x3 convs (+BN+ReLU) and the image-pooling partial sums.
//
// A 3x3 convolution with dilation r only couples pixels of the same phase (i mod r, j mod r): it is r*r independent
// dense 3x3 convolutions on the phase-subsampled images.  One warp owns one (image, 64-channel chunk, rate, phase,
// column segment) and walks down the phase image with a three-row register window, so every input element is
// loaded once per rate (instead of nine times) and all loads/stores are 128-byte coalesced (32 lanes x bf16x2).
// Packed fp32x2 FMAs; taps that fall outside the map are zeros held in registers ('same' padding).
struct AsppDwParams {
  const __nv_bfloat16* x;   // [B,h,w,C]
  const float* w;           // [nrates][9][C]  BN scale folded
  const float* shift;       // [nrates][C]
  __nv_bfloat16* out;       // [nrates][C/64 chunks][B*h*w][64]: K-block-major, every 64-channel slab of a rate is contiguous
                            // (sequential HBM writes here, contiguous 16 KB TMA boxes for the pointwise GEMM's A operand)
  float* pool_partial;      // [B][pool_items][C]  (rate 0 items also reduce their pixels for the pooling branch)
  int B, h, w_, C;
  int nrates;
  int rates[3];
  int nchunks;              // ceil(C / 64)
  int nseg[3];              // column segments per phase image
  int ts_sel[3];            // 0,1,2,3,4 -> 2,3,4,6,8 outputs per segment; 5 / 6 -> whole phase image (<= 2x2 / 3x3) in registers
  int item_off[4];          // prefix sums of items per rate (items = r*r*nseg)
  int pool_items;           // item_off[1]
  long long total_warps;    // B * item_off[nrates] * nchunks
  int debug;                // benchmark aid: bit0 skip output stores, bit1 skip the tap math
  const void* tmap_slab;    // slab kernel: 2D [B*h*w, C] bf16 view of x, box {64, 256}, no swizzle (device memory)
  const uint32_t* item_table;  // [items per image] packed (ri << 28 | seg << 20 | pi << 10 | pj): no div/mod per item (device memory)
};

__device__ __forceinline__ unsigned long long f32x2_from_bf16x2(uint32_t v) {
  return (static_cast<unsigned long long>(v & 0xFFFF0000u) << 32) | (v << 16);
}
__device__ __forceinline__ void ffma2_acc(unsigned long long& d, unsigned long long a, unsigned long long b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}
__device__ __forceinline__ unsigned long long fadd2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ unsigned long long fsub2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// Packed product with its OWN rounding.  ptxas contracts `mul.rn.f32x2` + `add.rn.f32x2` into one FFMA2 (single
// rounding; checked with tools/ubench/contract_check.cu), which would break the oracle's a + (b - a) * t operation
// order.  fma(a, b, +0) rounds exactly like the product and is not contracted with the following add.
__device__ __forceinline__ unsigned long long fmul2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(0ull));
  return d;
}
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  return (static_cast<unsigned long long>(__float_as_uint(hi)) << 32) | __float_as_uint(lo);
}

// kSmem = false: inputs straight from global memory (large maps; every load is a coalesced 128-byte line)
// kSmem = true : inputs from a shared-memory slab [h*w pixels][64 channels] staged once per (image, chunk)
template <int TS, bool kSmem>
__device__ __forceinline__ void aspp_dw_phase_item(const AsppDwParams& P, int b, int chunk, int ri, int pi, int pj, int seg,
                                                   int item_in_rate, int lane, uint32_t slab, const float* s_w, const float* s_shift) {
  const int r = P.rates[ri];
  const int c0 = chunk * 64 + lane * 2;
  const bool ch_ok = c0 < P.C;
  const int cc = ch_ok ? c0 : 0;
  // taps / shift for this lane's two channels
  unsigned long long wt[9];
  unsigned long long sh;
  if (kSmem) {
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const float2 v = *reinterpret_cast<const float2*>(s_w + (ri * 9 + t) * 64 + lane * 2);
      wt[t] = (static_cast<unsigned long long>(__float_as_uint(v.y)) << 32) | __float_as_uint(v.x);
    }
    const float2 shv = *reinterpret_cast<const float2*>(s_shift + ri * 64 + lane * 2);
    sh = (static_cast<unsigned long long>(__float_as_uint(shv.y)) << 32) | __float_as_uint(shv.x);
  } else {
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const float2 v = __ldg(reinterpret_cast<const float2*>(P.w + (static_cast<size_t>(ri) * 9 + t) * P.C + cc));
      wt[t] = (static_cast<unsigned long long>(__float_as_uint(v.y)) << 32) | __float_as_uint(v.x);
    }
    const float2 shv = __ldg(reinterpret_cast<const float2*>(P.shift + static_cast<size_t>(ri) * P.C + cc));
    sh = (static_cast<unsigned long long>(__float_as_uint(shv.y)) << 32) | __float_as_uint(shv.x);
  }

  const size_t img_px = static_cast<size_t>(P.h) * P.w_;
  const __nv_bfloat16* xb = P.x + static_cast<size_t>(b) * img_px * P.C + cc;
  __nv_bfloat16* ob = P.out + ((static_cast<size_t>(ri) * P.nchunks + chunk) * P.B + b) * img_px * 64 + lane * 2;
  const uint32_t slab_lane = slab + lane * 4;
  const int na = pi < P.h ? (P.h - pi + r - 1) / r : 0;   // rows of this phase image
  const int t0 = seg * TS;                                 // first output column (phase coordinates)
  unsigned long long psum = 0ull;                          // packed (0.f, 0.f)

  auto load_row = [&](int q, unsigned long long (&row)[TS + 2]) {
    const int i = pi + r * q;
    const bool row_ok = ch_ok && q >= 0 && i < P.h;
#pragma unroll
    for (int t = 0; t < TS + 2; ++t) {
      const int tt = t0 + t - 1;
      const int j = pj + r * tt;
      uint32_t v = 0u;
      if (row_ok && tt >= 0 && j < P.w_) {
        if (kSmem) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(slab_lane + (i * P.w_ + j) * 128));
        else v = __ldg(reinterpret_cast<const unsigned int*>(xb + (static_cast<size_t>(i) * P.w_ + j) * P.C));
      }
      row[t] = f32x2_from_bf16x2(v);
    }
  };
  auto emit = [&](int a, const unsigned long long (&top)[TS + 2], const unsigned long long (&mid)[TS + 2],
                  const unsigned long long (&bot)[TS + 2]) {
    const int i = pi + r * a;
#pragma unroll
    for (int t = 0; t < TS; ++t) {
      const int j = pj + r * (t0 + t);
      unsigned long long acc = sh;
      ffma2_acc(acc, wt[0], top[t]); ffma2_acc(acc, wt[1], top[t + 1]); ffma2_acc(acc, wt[2], top[t + 2]);
      ffma2_acc(acc, wt[3], mid[t]); ffma2_acc(acc, wt[4], mid[t + 1]); ffma2_acc(acc, wt[5], mid[t + 2]);
      ffma2_acc(acc, wt[6], bot[t]); ffma2_acc(acc, wt[7], bot[t + 1]); ffma2_acc(acc, wt[8], bot[t + 2]);
      if (ch_ok && j < P.w_ && !(P.debug & 1)) {
        const float lo = fmaxf(__uint_as_float(static_cast<uint32_t>(acc)), 0.0f);
        const float hi = fmaxf(__uint_as_float(static_cast<uint32_t>(acc >> 32)), 0.0f);
        *reinterpret_cast<uint32_t*>(ob + (static_cast<size_t>(i) * P.w_ + j) * 64) = pack_bf16x2(lo, hi);
        if (!kSmem) psum = fadd2(psum, mid[t + 1]);
      }
    }
  };

  unsigned long long ra[TS + 2], rb[TS + 2], rc[TS + 2];
  if (na > 0) {
    load_row(-1, ra);
    load_row(0, rb);
    for (int a = 0; a < na; a += 3) {
      load_row(a + 1, rc);
      emit(a, ra, rb, rc);
      if (a + 1 < na) {
        load_row(a + 2, ra);
        emit(a + 1, rb, rc, ra);
      }
      if (a + 2 < na) {
        load_row(a + 3, rb);
        emit(a + 2, rc, ra, rb);
      }
    }
  }
  if (!kSmem && ri == 0 && ch_ok) {
    float2 pv;
    pv.x = __uint_as_float(static_cast<uint32_t>(psum));
    pv.y = __uint_as_float(static_cast<uint32_t>(psum >> 32));
    *reinterpret_cast<float2*>(P.pool_partial + (static_cast<size_t>(b) * P.pool_items + item_in_rate) * P.C + c0) = pv;
  }
}

// Small phase images (at most NA x NT pixels, e.g. rate 18 or 12 on a 32x32 map): the whole phase image lives in
// registers, every output is a fully unrolled sum over the taps that exist at compile time.
template <int NA, int NT, bool kSmem>
__device__ __forceinline__ void aspp_dw_small_item(const AsppDwParams& P, int b, int chunk, int ri, int pi, int pj, int item_in_rate,
                                                   int lane, uint32_t slab, const float* s_w, const float* s_shift) {
  const int r = P.rates[ri];
  const int c0 = chunk * 64 + lane * 2;
  const bool ch_ok = c0 < P.C;
  const int cc = ch_ok ? c0 : 0;
  unsigned long long wt[9], sh;
  if (kSmem) {
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const float2 v = *reinterpret_cast<const float2*>(s_w + (ri * 9 + t) * 64 + lane * 2);
      wt[t] = (static_cast<unsigned long long>(__float_as_uint(v.y)) << 32) | __float_as_uint(v.x);
    }
    const float2 shv = *reinterpret_cast<const float2*>(s_shift + ri * 64 + lane * 2);
    sh = (static_cast<unsigned long long>(__float_as_uint(shv.y)) << 32) | __float_as_uint(shv.x);
  } else {
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const float2 v = __ldg(reinterpret_cast<const float2*>(P.w + (static_cast<size_t>(ri) * 9 + t) * P.C + cc));
      wt[t] = (static_cast<unsigned long long>(__float_as_uint(v.y)) << 32) | __float_as_uint(v.x);
    }
    const float2 shv = __ldg(reinterpret_cast<const float2*>(P.shift + static_cast<size_t>(ri) * P.C + cc));
    sh = (static_cast<unsigned long long>(__float_as_uint(shv.y)) << 32) | __float_as_uint(shv.x);
  }
  const size_t img_px = static_cast<size_t>(P.h) * P.w_;
  const __nv_bfloat16* xb = P.x + static_cast<size_t>(b) * img_px * P.C + cc;
  __nv_bfloat16* ob = P.out + ((static_cast<size_t>(ri) * P.nchunks + chunk) * P.B + b) * img_px * 64 + lane * 2;
  const uint32_t slab_lane = slab + lane * 4;
  unsigned long long x[NA][NT];
  unsigned long long psum = 0ull;
#pragma unroll
  for (int a = 0; a < NA; ++a)
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      const int i = pi + r * a, j = pj + r * t;
      uint32_t v = 0u;
      if (ch_ok && i < P.h && j < P.w_) {
        if (kSmem) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(slab_lane + (i * P.w_ + j) * 128));
        else v = __ldg(reinterpret_cast<const unsigned int*>(xb + (static_cast<size_t>(i) * P.w_ + j) * P.C));
      }
      x[a][t] = f32x2_from_bf16x2(v);
      if (!kSmem) psum = fadd2(psum, x[a][t]);
    }
#pragma unroll
  for (int a = 0; a < NA; ++a)
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      const int i = pi + r * a, j = pj + r * t;
      if (i >= P.h || j >= P.w_) continue;       // warp-uniform
      unsigned long long acc = sh;
#pragma unroll
      for (int u = 0; u < 3; ++u)
#pragma unroll
        for (int v = 0; v < 3; ++v) {
          const int aa = a + u - 1, tt = t + v - 1;
          if (aa >= 0 && aa < NA && tt >= 0 && tt < NT) ffma2_acc(acc, wt[u * 3 + v], x[aa][tt]);   // resolved at compile time
        }
      if (ch_ok && !(P.debug & 1)) {
        const float lo = fmaxf(__uint_as_float(static_cast<uint32_t>(acc)), 0.0f);
        const float hi = fmaxf(__uint_as_float(static_cast<uint32_t>(acc >> 32)), 0.0f);
        *reinterpret_cast<uint32_t*>(ob + (static_cast<size_t>(i) * P.w_ + j) * 64) = pack_bf16x2(lo, hi);
      }
    }
  if (!kSmem && ri == 0 && ch_ok) {
    float2 pv;
    pv.x = __uint_as_float(static_cast<uint32_t>(psum));
    pv.y = __uint_as_float(static_cast<uint32_t>(psum >> 32));
    *reinterpret_cast<float2*>(P.pool_partial + (static_cast<size_t>(b) * P.pool_items + item_in_rate) * P.C + c0) = pv;
  }
}

template <bool kSmem>
__device__ __forceinline__ void aspp_dw_dispatch(const AsppDwParams& P, int b, int chunk, int it, int lane, uint32_t slab,
                                                 const float* s_w, const float* s_shift) {
  const uint32_t e = __ldg(P.item_table + it);
  const int ri = static_cast<int>(e >> 28), seg = static_cast<int>((e >> 20) & 0xFF);
  const int pi = static_cast<int>((e >> 10) & 0x3FF), pj = static_cast<int>(e & 0x3FF);
  const int local = it - P.item_off[ri];
  switch (P.ts_sel[ri]) {
    case 0: aspp_dw_phase_item<2, kSmem>(P, b, chunk, ri, pi, pj, seg, local, lane, slab, s_w, s_shift); break;
    case 1: aspp_dw_phase_item<3, kSmem>(P, b, chunk, ri, pi, pj, seg, local, lane, slab, s_w, s_shift); break;
    case 2: aspp_dw_phase_item<4, kSmem>(P, b, chunk, ri, pi, pj, seg, local, lane, slab, s_w, s_shift); break;
    case 3: aspp_dw_phase_item<6, kSmem>(P, b, chunk, ri, pi, pj, seg, local, lane, slab, s_w, s_shift); break;
    case 5: aspp_dw_small_item<2, 2, kSmem>(P, b, chunk, ri, pi, pj, local, lane, slab, s_w, s_shift); break;
    case 6: aspp_dw_small_item<3, 3, kSmem>(P, b, chunk, ri, pi, pj, local, lane, slab, s_w, s_shift); break;
    default: aspp_dw_phase_item<8, kSmem>(P, b, chunk, ri, pi, pj, seg, local, lane, slab, s_w, s_shift); break;
  }
}

// large maps: one warp per work item, inputs from global memory
__global__ void __launch_bounds__(128, 4) aspp_dw_phase_kernel(const __grid_constant__ AsppDwParams P) {
  const long long gw = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (gw >= P.total_warps) return;
  const int lane = threadIdx.x & 31;
  const int chunk = static_cast<int>(gw % P.nchunks);
  const long long rest = gw / P.nchunks;
  const int items_per_img = P.item_off[P.nrates];
  const int it = static_cast<int>(rest % items_per_img);
  const int b = static_cast<int>(rest / items_per_img);
  aspp_dw_dispatch<false>(P, b, chunk, it, lane, 0u, nullptr, nullptr);
}

// small maps (h*w*128 B fits in shared memory): persistent CTAs, one (image, 64-channel chunk) slab at a time.
// The slab is staged with 16-byte cp.async; all three rates and the pooling partial sums are then computed from
// shared memory (deterministic order), so x is read from HBM exactly once.
constexpr int kSlabThreads = 512;
__global__ void __launch_bounds__(kSlabThreads, 1) aspp_dw_slab_kernel(const __grid_constant__ AsppDwParams P) {
  extern __shared__ __align__(128) uint8_t slab_smem[];
  const int npix = P.h * P.w_;
  const int nbox = (npix + 255) / 256;                 // TMA boxes of 256 pixels x 64 channels (32 KB each)
  uint8_t* s_slab = slab_smem;
  float* s_w = reinterpret_cast<float*>(s_slab + static_cast<size_t>(nbox) * 256 * 128);   // [27][64]
  float* s_shift = s_w + 27 * 64;                                                     // [3][64]
  float* s_red = s_shift + 3 * 64;                                                    // [16][64]
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_red + 16 * 64);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t slab = smem_u32(s_slab);
  const int nitems = P.item_off[P.nrates];
  const int num_slabs = P.B * P.nchunks;
  if (tid == 0) {
    mbar_init(s_bar, 1);
    fence_barrier_init();
  }
  __syncthreads();
  uint32_t slab_phase = 0;
  for (int sid = blockIdx.x; sid < num_slabs; sid += gridDim.x) {
    const int b = sid / P.nchunks, chunk = sid - b * P.nchunks;
    if (tid == 0) {   // the whole slab arrives through TMA (OOB channels / rows are zero filled); x is read from HBM once
      mbar_arrive_expect_tx(s_bar, static_cast<uint32_t>(nbox) * 32768u);
      for (int q = 0; q < nbox; ++q)
        tma_load_2d(s_slab + static_cast<size_t>(q) * 32768, P.tmap_slab, s_bar, chunk * 64, b * npix + q * 256, kEvictFirst);
    }
    for (int i = tid; i < P.nrates * 9 * 64; i += kSlabThreads) {
      const int cc = chunk * 64 + (i & 63);
      s_w[i] = cc < P.C ? __ldg(P.w + static_cast<size_t>(i >> 6) * P.C + cc) : 0.0f;
    }
    for (int i = tid; i < P.nrates * 64; i += kSlabThreads) {
      const int cc = chunk * 64 + (i & 63);
      s_shift[i] = cc < P.C ? __ldg(P.shift + static_cast<size_t>(i >> 6) * P.C + cc) : 0.0f;
    }
    __syncthreads();                  // taps / shifts staged
    mbar_wait(s_bar, slab_phase);     // slab landed
    slab_phase ^= 1;

    // the 16 warps share the (rate, phase, segment) items of the slab, heaviest rate first
    for (int it = warp; it < nitems; it += kSlabThreads / 32) aspp_dw_dispatch<true>(P, b, chunk, it, lane, slab, s_w, s_shift);

    // image-pooling partial sums over the slab: lane = 2 channels, 16 warps stride over the pixels
    {
      float sx = 0.0f, sy = 0.0f;
      for (int p = warp; p < npix; p += kSlabThreads / 32) {
        uint32_t v;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(slab + p * 128 + lane * 4));
        sx += bf16_lo(v);
        sy += bf16_hi(v);
      }
      s_red[warp * 64 + lane * 2] = sx;
      s_red[warp * 64 + lane * 2 + 1] = sy;
    }
    __syncthreads();
    if (tid < 64) {
      float s0 = 0.0f;
#pragma unroll
      for (int q = 0; q < kSlabThreads / 32; ++q) s0 += s_red[q * 64 + tid];
      const int cc = chunk * 64 + tid;
      if (cc < P.C) P.pool_partial[static_cast<size_t>(b) * P.C + cc] = s0;   // pool_items == 1 on this path
    }
    __syncthreads();   // slab / s_red are rewritten by the next iteration
  }
}

// ASPP Lite has no atrous branches: plain per-channel partial sums for the image-pooling branch.
// grid (ceil(C/64), nbands, B), block 256 = 8 warps striding over the band's pixels; lane = 2 channels.
struct PoolParams {
  const __nv_bfloat16* x;   // [B,h*w,C]
  float* pool_partial;      // [B][nbands][C]
  int npix, C, nbands, pix_per_band;
};
__global__ void __launch_bounds__(256) global_pool_kernel(const PoolParams P) {
  __shared__ float s_red[8][64];
  const int chunk = blockIdx.x, band = blockIdx.y, b = blockIdx.z;
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const int c0 = chunk * 64 + lane * 2;
  const int p0 = band * P.pix_per_band, p1 = min(P.npix, p0 + P.pix_per_band);
  float sx = 0.0f, sy = 0.0f;
  if (c0 < P.C) {
    const __nv_bfloat16* xb = P.x + static_cast<size_t>(b) * P.npix * P.C + c0;
    for (int p = p0 + wp; p < p1; p += 8) {
      const uint32_t v = __ldg(reinterpret_cast<const unsigned int*>(xb + static_cast<size_t>(p) * P.C));
      sx += bf16_lo(v);
      sy += bf16_hi(v);
    }
  }
  s_red[wp][lane * 2] = sx;
  s_red[wp][lane * 2 + 1] = sy;
  __syncthreads();
  if (threadIdx.x < 64) {
    float s = 0.0f;
#pragma unroll
    for (int q = 0; q < 8; ++q) s += s_red[q][threadIdx.x];
    const int cc = chunk * 64 + threadIdx.x;
    if (cc < P.C) P.pool_partial[(static_cast<size_t>(b) * P.nbands + band) * P.C + cc] = s;
  }
}

// ------------------------------------------------------------------------------------------------
// image-pooling branch folded into a per-image shift of the concat_projection epilogue.
// grid B, block 1024.
struct PoolProjParams {
  const float* pool_partial;     // [B][nbands][C]
  const __nv_bfloat16* w_ip;     // [C][256]  image_pooling kernel, bf16, k-major rows
  const float* ip_scale;         // [256] image_pooling_BN folded
  const float* ip_shift;
  const __nv_bfloat16* w_proj4;  // [256][256] rows 0..255 of concat_projection (the b4 slice), [k][n]
  const float* proj_scale;       // [256] concat_projection_BN folded
  const float* proj_shift;
  float* img_shift;              // [B][256]  = (b4 . Wproj4) * proj_scale + proj_shift
  float* b4_out;                 // [B][256]  tap: image_pooling output after BN+ReLU (bf16-rounded)
  int C, nbands;
  float inv_count;               // 1 / (h*w)
};

// block 1024 = 128 output-channel pairs x 8 K-slices.  There is no L1 to speak of next to the big GEMM CTAs and the
// weights come from L2, so every thread keeps 16 independent 32-bit loads in flight (two adjacent output channels each).
__global__ void __launch_bounds__(1024) pool_proj_kernel(const PoolProjParams P) {
  extern __shared__ float s_mean[];  // [C] + [256] + [8][256]
  float* s_b4 = s_mean + P.C;
  float* s_part = s_b4 + 256;
  const int b = blockIdx.x, np = threadIdx.x & 127, ks = threadIdx.x >> 7;   // np: channel pair, ks: K slice 0..7
  for (int c = threadIdx.x; c < P.C; c += 1024) {
    float s = 0.0f;
    for (int q = 0; q < P.nbands; ++q) s += P.pool_partial[(static_cast<size_t>(b) * P.nbands + q) * P.C + c];
    s_mean[c] = s * P.inv_count;
  }
  __syncthreads();
  auto gemv = [&](const float* vec, const __nv_bfloat16* w, int K) {   // partial dot products of this K slice -> s_part[ks][2np..]
    const int kq = (K + 7) / 8, k0 = ks * kq, k1 = min(K, k0 + kq);
    const unsigned int* wp = reinterpret_cast<const unsigned int*>(w) + np;   // row k: 128 pairs
    float a0 = 0.0f, a1 = 0.0f, c0 = 0.0f, c1 = 0.0f;
    int k = k0;
    for (; k + 15 < k1; k += 16) {
      uint32_t v[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) v[u] = __ldg(wp + static_cast<size_t>(k + u) * 128);
#pragma unroll
      for (int u = 0; u < 16; u += 2) {
        a0 = fmaf(vec[k + u], bf16_lo(v[u]), a0);
        a1 = fmaf(vec[k + u], bf16_hi(v[u]), a1);
        c0 = fmaf(vec[k + u + 1], bf16_lo(v[u + 1]), c0);
        c1 = fmaf(vec[k + u + 1], bf16_hi(v[u + 1]), c1);
      }
    }
    for (; k < k1; ++k) {
      const uint32_t v = __ldg(wp + static_cast<size_t>(k) * 128);
      a0 = fmaf(vec[k], bf16_lo(v), a0);
      a1 = fmaf(vec[k], bf16_hi(v), a1);
    }
    s_part[ks * 256 + 2 * np] = a0 + c0;
    s_part[ks * 256 + 2 * np + 1] = a1 + c1;
  };
  gemv(s_mean, P.w_ip, P.C);
  __syncthreads();
  if (threadIdx.x < 256) {
    const int n = threadIdx.x;
    float acc = 0.0f;
#pragma unroll
    for (int q = 0; q < 8; ++q) acc += s_part[q * 256 + n];
    float v = fmaxf(fmaf(acc, P.ip_scale[n], P.ip_shift[n]), 0.0f);
    v = __bfloat162float(__float2bfloat16_rn(v));  // activation rounding point, as every other branch
    s_b4[n] = v;
    P.b4_out[static_cast<size_t>(b) * 256 + n] = v;
  }
  __syncthreads();
  gemv(s_b4, P.w_proj4, 256);
  __syncthreads();
  if (threadIdx.x < 256) {
    const int n = threadIdx.x;
    float acc2 = 0.0f;
#pragma unroll
    for (int q = 0; q < 8; ++q) acc2 += s_part[q * 256 + n];
    P.img_shift[static_cast<size_t>(b) * 256 + n] = fmaf(acc2, P.proj_scale[n], P.proj_shift[n]);
  }
}

// ------------------------------------------------------------------------------------------------
// generic depthwise 3x3 (dilation d, 'same') + scale/shift (+ReLU); one thread = one pixel x 8 channels
struct DwParams {
  const __nv_bfloat16* x;  // [B,H,W,C]
  const float* w;          // [9][C] scale folded
  const float* shift;      // [C]
  __nv_bfloat16* out;      // [B,H,W,C]
  int B, H, W, C, rate, relu;
  int wstride;             // row stride of w (>= C; packed taps may be zero padded)
  int flip;                // 1: tap (u,v) reads w[2-u][2-v] — the data gradient of the same convolution (training); shift may be NULL
};
__global__ void __launch_bounds__(256) depthwise3x3_kernel(const DwParams P) {
  const int vecs = P.C >> 3;
  const size_t total = static_cast<size_t>(P.B) * P.H * P.W * vecs;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int vec = static_cast<int>(idx % vecs);
    size_t pix = idx / vecs;
    const int j = static_cast<int>(pix % P.W);
    pix /= P.W;
    const int i = static_cast<int>(pix % P.H);
    const int b = static_cast<int>(pix / P.H);
    const int c0 = vec * 8;
    float acc[8];
    if (P.shift) {
      const float4 s0 = __ldg(reinterpret_cast<const float4*>(P.shift + c0));
      const float4 s1 = __ldg(reinterpret_cast<const float4*>(P.shift + c0 + 4));
      acc[0] = s0.x; acc[1] = s0.y; acc[2] = s0.z; acc[3] = s0.w;
      acc[4] = s1.x; acc[5] = s1.y; acc[6] = s1.z; acc[7] = s1.w;
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = 0.0f;
    }
    const __nv_bfloat16* xb = P.x + static_cast<size_t>(b) * P.H * P.W * P.C + c0;
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      const int ii = i + (u - 1) * P.rate;
      if (ii < 0 || ii >= P.H) continue;
#pragma unroll
      for (int v = 0; v < 3; ++v) {
        const int jj = j + (v - 1) * P.rate;
        if (jj < 0 || jj >= P.W) continue;
        float xv[8];
        unpack8(ldg_nc_v4(xb + (static_cast<size_t>(ii) * P.W + jj) * P.C), xv);
        const int tap = P.flip ? 8 - (u * 3 + v) : u * 3 + v;
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(P.w + tap * P.wstride + c0));
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(P.w + tap * P.wstride + c0 + 4));
        acc[0] = fmaf(xv[0], w0.x, acc[0]); acc[1] = fmaf(xv[1], w0.y, acc[1]);
        acc[2] = fmaf(xv[2], w0.z, acc[2]); acc[3] = fmaf(xv[3], w0.w, acc[3]);
        acc[4] = fmaf(xv[4], w1.x, acc[4]); acc[5] = fmaf(xv[5], w1.y, acc[5]);
        acc[6] = fmaf(xv[6], w1.z, acc[6]); acc[7] = fmaf(xv[7], w1.w, acc[7]);
      }
    }
    if (P.relu) {
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = fmaxf(acc[k], 0.0f);
    }
    stg_v4(P.out + ((static_cast<size_t>(b) * P.H + i) * P.W + j) * P.C + c0, pack8(acc));
  }
}

// ------------------------------------------------------------------------------------------------
// tf.image.resize bilinear (TF2: half-pixel centres, no antialias): same operation order as the oracle
//   src = (dst + 0.5) * (in/out) - 0.5 ; lo = max(floor(src),0) ; hi = min(ceil(src), in-1) ; t = src - floor(src)
//   top = tl + (tr - tl) * tx ; bot = bl + (br - bl) * tx ; out = top + (bot - top) * ty      (no FMA contraction)
__device__ __forceinline__ void resize_coord(int dst, float scale, int n_in, int& lo, int& hi, float& t) {
  const float src = __fsub_rn(__fmul_rn(static_cast<float>(dst) + 0.5f, scale), 0.5f);
  const float fl = floorf(src);
  lo = max(static_cast<int>(fl), 0);
  hi = min(static_cast<int>(ceilf(src)), n_in - 1);
  t = __fsub_rn(src, fl);
}
__device__ __forceinline__ float lerp_nofma(float a, float b, float t) { return __fadd_rn(a, __fmul_rn(__fsub_rn(b, a), t)); }

struct ResizeParams {
  const __nv_bfloat16* x;  // [B,hi,wi,C]
  __nv_bfloat16* out;      // [B,ho,wo,ldo] written at column col_off
  int B, hi, wi, C, ho, wo, ldo, col_off;
  float sy, sx;            // hi/ho, wi/wo computed in fp32 on the host exactly like the oracle
};
// grid (ho, B): one output row per block; threads stride over (X, 8-channel vector) with 32-bit index math
__global__ void __launch_bounds__(256) resize_bilinear_kernel(const ResizeParams P) {
  const int vecs = P.C >> 3;
  const int Y = blockIdx.x, b = blockIdx.y;
  int y0, y1;
  float ty;
  resize_coord(Y, P.sy, P.hi, y0, y1, ty);
  const __nv_bfloat16* row0 = P.x + (static_cast<size_t>(b) * P.hi + y0) * P.wi * P.C;
  const __nv_bfloat16* row1 = P.x + (static_cast<size_t>(b) * P.hi + y1) * P.wi * P.C;
  __nv_bfloat16* orow = P.out + (static_cast<size_t>(b) * P.ho + Y) * P.wo * P.ldo + P.col_off;
  const int total = P.wo * vecs;
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int X = idx / vecs;
    const int vec = idx - X * vecs;
    int x0, x1;
    float tx;
    resize_coord(X, P.sx, P.wi, x0, x1, tx);
    float tl[8], tr[8], bl[8], br[8], o[8];
    unpack8(ldg_nc_v4(row0 + x0 * P.C + vec * 8), tl);
    unpack8(ldg_nc_v4(row0 + x1 * P.C + vec * 8), tr);
    unpack8(ldg_nc_v4(row1 + x0 * P.C + vec * 8), bl);
    unpack8(ldg_nc_v4(row1 + x1 * P.C + vec * 8), br);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float top = lerp_nofma(tl[k], tr[k], tx);
      const float bot = lerp_nofma(bl[k], br[k], tx);
      o[k] = lerp_nofma(top, bot, ty);
    }
    stg_v4(orow + static_cast<size_t>(X) * P.ldo + vec * 8, pack8(o));
  }
}

// x4 upsampling (decoder_resize at OS16: 32x32 -> 128x128): one warp per low-res cell, lane = 8-channel vector.
// The 4x4 outputs Y in [4m+2, 4m+6), X in [4k+2, 4k+6) share the four corner pixels (m, m+1) x (k, k+1): corners are
// loaded and unpacked once, the horizontal lerps once per column.  Same operations in the same order as the generic
// kernel / the oracle (top/bottom lerp, then vertical), so the result is bit identical.
__global__ void __launch_bounds__(256, 2) resize_bilinear_x4_kernel(const ResizeParams P) {
  const int cells_x = P.wi + 1, cells_y = P.hi + 1;
  const int total = P.B * cells_y * cells_x;
  const int lane = threadIdx.x & 31;
  const int vecs = P.C >> 3;
  for (int cell = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); cell < total; cell += gridDim.x * (blockDim.x >> 5)) {
    const int k = cell % cells_x - 1;
    const int r = cell / cells_x;
    const int m = r % cells_y - 1;
    const int b = r / cells_y;
    const int X0 = 4 * k + 2, Y0 = 4 * m + 2;
    float ty[4], tx[4];
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      int lo, hi;
      resize_coord(min(max(Y0 + d, 0), P.ho - 1), P.sy, P.hi, lo, hi, ty[d]);
      resize_coord(min(max(X0 + d, 0), P.wo - 1), P.sx, P.wi, lo, hi, tx[d]);
    }
    const int y0 = max(m, 0), y1 = min(m + 1, P.hi - 1), x0 = max(k, 0), x1 = min(k + 1, P.wi - 1);
    const __nv_bfloat16* xb = P.x + static_cast<size_t>(b) * P.hi * P.wi * P.C;
    for (int vec = lane; vec < vecs; vec += 32) {
      // packed fp32x2 arithmetic (sub / mul / add .rn.f32x2, no FMA): the same rounded operations as lerp_nofma on each
      // element, half the instructions
      unsigned long long tl[4], dt[4], bl[4], db[4];
      {
        const uint4 a = ldg_nc_v4(xb + (static_cast<size_t>(y0) * P.wi + x0) * P.C + vec * 8);
        const uint4 bq = ldg_nc_v4(xb + (static_cast<size_t>(y0) * P.wi + x1) * P.C + vec * 8);
        const uint4 cq = ldg_nc_v4(xb + (static_cast<size_t>(y1) * P.wi + x0) * P.C + vec * 8);
        const uint4 d = ldg_nc_v4(xb + (static_cast<size_t>(y1) * P.wi + x1) * P.C + vec * 8);
        const uint32_t av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {bq.x, bq.y, bq.z, bq.w}, cv[4] = {cq.x, cq.y, cq.z, cq.w}, dv[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          tl[c] = f32x2_from_bf16x2(av[c]);
          dt[c] = fsub2(f32x2_from_bf16x2(bv[c]), tl[c]);      // tr - tl
          bl[c] = f32x2_from_bf16x2(cv[c]);
          db[c] = fsub2(f32x2_from_bf16x2(dv[c]), bl[c]);      // br - bl
        }
      }
#pragma unroll
      for (int dx = 0; dx < 4; ++dx) {
        const int X = X0 + dx;
        if (X < 0 || X >= P.wo) continue;
        const unsigned long long txx = pack2(tx[dx], tx[dx]);
        unsigned long long top[4], dv[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          top[c] = fadd2(tl[c], fmul2(dt[c], txx));
          const unsigned long long bot = fadd2(bl[c], fmul2(db[c], txx));
          dv[c] = fsub2(bot, top[c]);
        }
#pragma unroll
        for (int dy = 0; dy < 4; ++dy) {
          const int Y = Y0 + dy;
          if (Y < 0 || Y >= P.ho) continue;
          const unsigned long long tyy = pack2(ty[dy], ty[dy]);
          uint4 o;
          o.x = f32x2_to_bf16x2(fadd2(top[0], fmul2(dv[0], tyy)));
          o.y = f32x2_to_bf16x2(fadd2(top[1], fmul2(dv[1], tyy)));
          o.z = f32x2_to_bf16x2(fadd2(top[2], fmul2(dv[2], tyy)));
          o.w = f32x2_to_bf16x2(fadd2(top[3], fmul2(dv[3], tyy)));
          stg_v4(P.out + ((static_cast<size_t>(b) * P.ho + Y) * P.wo + X) * P.ldo + P.col_off + vec * 8, o);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// pred_resize + argmax.  logits are PLANAR fp32 [B][NC][hi][wi] so neighbouring threads read neighbouring
// addresses.  First maximum wins (np.argmax; strict '>' scan from class 0, deeplabSegment.cpp:160-167).
struct ArgmaxParams {
  const float* logits;  // [B,NC,hi,wi]
  uint8_t* labels;      // [B,ho,wo]
  int B, NC, hi, wi, ho, wo;
  float sy, sx;
};

// integer scale S (multiple of 4; S = 4 for the decoder models, 8 / 16 / 32 for the *_lite models that resize the
// OS-resolution logits directly): the S x S outputs Y in [S*m + S/2, S*m + 3S/2), X likewise, share the four corner
// pixels (m, m+1) x (k, k+1) (m, k start at -1).  One thread owns a 4 x 4 block of them: corners loaded once per class,
// horizontal lerps once per column.
struct ArgmaxIntParams {
  ArgmaxParams a;
  int S;   // ho / hi == wo / wi
};
__global__ void __launch_bounds__(128) resize_argmax_x4_kernel(const ArgmaxIntParams Q) {
  const ArgmaxParams& P = Q.a;
  const int S = Q.S, sub = S >> 2;
  const int gx_n = (P.wi + 1) * sub, gy_n = (P.hi + 1) * sub;
  const size_t total = static_cast<size_t>(P.B) * gy_n * gx_n;
  const size_t plane = static_cast<size_t>(P.hi) * P.wi;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int gx = static_cast<int>(idx % gx_n);
    size_t r = idx / gx_n;
    const int gy = static_cast<int>(r % gy_n);
    const int b = static_cast<int>(r / gy_n);
    const int k = gx / sub - 1, m = gy / sub - 1;
    const int X0 = S * k + (S >> 1) + 4 * (gx % sub), Y0 = S * m + (S >> 1) + 4 * (gy % sub);
    // per-output coordinates through the generic formula (bit-identical to the generic kernel / oracle)
    int ylo[4], yhi[4], xlo[4], xhi[4];
    float ty[4], tx[4];
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      const int Y = min(max(Y0 + d, 0), P.ho - 1), X = min(max(X0 + d, 0), P.wo - 1);
      resize_coord(Y, P.sy, P.hi, ylo[d], yhi[d], ty[d]);
      resize_coord(X, P.sx, P.wi, xlo[d], xhi[d], tx[d]);
    }
    // all valid outputs of the cell share the same corner pixels
    const int y0 = max(m, 0), y1 = min(m + 1, P.hi - 1), x0 = max(k, 0), x1 = min(k + 1, P.wi - 1);
    float best[16];
    int arg[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { best[i] = -INFINITY; arg[i] = 0; }
    const float* lb = P.logits + static_cast<size_t>(b) * P.NC * plane;
    const int o00 = y0 * P.wi + x0, o01 = y0 * P.wi + x1, o10 = y1 * P.wi + x0, o11 = y1 * P.wi + x1;
    // a + (b - a) * t with the differences hoisted — the very same rounded operations as lerp_nofma — on packed fp32
    // pairs (two columns / two rows per instruction; fmul2 = fma(a, b, +0) so ptxas cannot contract it with the add)
    const unsigned long long tx01 = pack2(tx[0], tx[1]), tx23 = pack2(tx[2], tx[3]);
    const unsigned long long ty01 = pack2(ty[0], ty[1]), ty23 = pack2(ty[2], ty[3]);
    auto scan = [&](int c, float tl, float tr, float bl, float br) {
      const float dtop = __fsub_rn(tr, tl), dbot = __fsub_rn(br, bl);
      const unsigned long long tl2 = pack2(tl, tl), bl2 = pack2(bl, bl), dt2 = pack2(dtop, dtop), db2 = pack2(dbot, dbot);
      unsigned long long top2[2], dv2[2];
      top2[0] = fadd2(tl2, fmul2(dt2, tx01));
      top2[1] = fadd2(tl2, fmul2(dt2, tx23));
      dv2[0] = fsub2(fadd2(bl2, fmul2(db2, tx01)), top2[0]);
      dv2[1] = fsub2(fadd2(bl2, fmul2(db2, tx23)), top2[1]);
#pragma unroll
      for (int dx = 0; dx < 4; ++dx) {
        const float top = __uint_as_float(static_cast<uint32_t>(top2[dx >> 1] >> ((dx & 1) * 32)));
        const float dv = __uint_as_float(static_cast<uint32_t>(dv2[dx >> 1] >> ((dx & 1) * 32)));
        const unsigned long long t2 = pack2(top, top), d2 = pack2(dv, dv);
        const unsigned long long v01 = fadd2(t2, fmul2(d2, ty01)), v23 = fadd2(t2, fmul2(d2, ty23));
        const float v[4] = {__uint_as_float(static_cast<uint32_t>(v01)), __uint_as_float(static_cast<uint32_t>(v01 >> 32)),
                            __uint_as_float(static_cast<uint32_t>(v23)), __uint_as_float(static_cast<uint32_t>(v23 >> 32))};
#pragma unroll
        for (int dy = 0; dy < 4; ++dy)
          if (v[dy] > best[dy * 4 + dx]) { best[dy * 4 + dx] = v[dy]; arg[dy * 4 + dx] = c; }
      }
    };
    int c = 0;
    for (; c + 3 <= P.NC; c += 3) {   // three classes per trip: twelve independent loads in flight
      float q[3][4];
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        const float* lp = lb + static_cast<size_t>(c + u) * plane;
        q[u][0] = __ldg(lp + o00); q[u][1] = __ldg(lp + o01); q[u][2] = __ldg(lp + o10); q[u][3] = __ldg(lp + o11);
      }
#pragma unroll
      for (int u = 0; u < 3; ++u) scan(c + u, q[u][0], q[u][1], q[u][2], q[u][3]);
    }
    for (; c < P.NC; ++c) {
      const float* lp = lb + static_cast<size_t>(c) * plane;
      scan(c, __ldg(lp + o00), __ldg(lp + o01), __ldg(lp + o10), __ldg(lp + o11));
    }
    uint8_t* ob = P.labels + static_cast<size_t>(b) * P.ho * P.wo;
#pragma unroll
    for (int dy = 0; dy < 4; ++dy) {
      const int Y = Y0 + dy;
      if (Y < 0 || Y >= P.ho) continue;
      if (X0 >= 0 && X0 + 3 < P.wo) {
        const uint32_t w = static_cast<uint32_t>(arg[dy * 4]) | (static_cast<uint32_t>(arg[dy * 4 + 1]) << 8) |
                           (static_cast<uint32_t>(arg[dy * 4 + 2]) << 16) | (static_cast<uint32_t>(arg[dy * 4 + 3]) << 24);
        // X0 = 4k+2: 2-byte aligned only -> two 16-bit stores
        *reinterpret_cast<uint16_t*>(ob + static_cast<size_t>(Y) * P.wo + X0) = static_cast<uint16_t>(w & 0xFFFF);
        *reinterpret_cast<uint16_t*>(ob + static_cast<size_t>(Y) * P.wo + X0 + 2) = static_cast<uint16_t>(w >> 16);
      } else {
#pragma unroll
        for (int dx = 0; dx < 4; ++dx) {
          const int X = X0 + dx;
          if (X >= 0 && X < P.wo) ob[static_cast<size_t>(Y) * P.wo + X] = static_cast<uint8_t>(arg[dy * 4 + dx]);
        }
      }
    }
  }
}

// generic scale: one thread per output pixel
__global__ void __launch_bounds__(256) resize_argmax_generic_kernel(const ArgmaxParams P) {
  const size_t total = static_cast<size_t>(P.B) * P.ho * P.wo;
  const size_t plane = static_cast<size_t>(P.hi) * P.wi;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int X = static_cast<int>(idx % P.wo);
    size_t r = idx / P.wo;
    const int Y = static_cast<int>(r % P.ho);
    const int b = static_cast<int>(r / P.ho);
    int y0, y1, x0, x1;
    float ty, tx;
    resize_coord(Y, P.sy, P.hi, y0, y1, ty);
    resize_coord(X, P.sx, P.wi, x0, x1, tx);
    const float* lb = P.logits + static_cast<size_t>(b) * P.NC * plane;
    float best = -INFINITY;
    int arg = 0;
    for (int c = 0; c < P.NC; ++c) {
      const float* lp = lb + static_cast<size_t>(c) * plane;
      const float top = lerp_nofma(__ldg(lp + static_cast<size_t>(y0) * P.wi + x0), __ldg(lp + static_cast<size_t>(y0) * P.wi + x1), tx);
      const float bot = lerp_nofma(__ldg(lp + static_cast<size_t>(y1) * P.wi + x0), __ldg(lp + static_cast<size_t>(y1) * P.wi + x1), tx);
      const float v = lerp_nofma(top, bot, ty);
      if (v > best) { best = v; arg = c; }
    }
    P.labels[idx] = static_cast<uint8_t>(arg);
  }
}

// materialised pred_resize (+ Softmax): fp32 NHWC [B,ho,wo,NC]; one thread per output pixel
struct DenseResizeParams {
  const float* logits;  // planar [B,NC,hi,wi]
  float* out;           // [B,ho,wo,NC]
  int B, NC, hi, wi, ho, wo, softmax;
  float sy, sx;
};
__global__ void __launch_bounds__(256) resize_dense_kernel(const DenseResizeParams P) {
  const size_t total = static_cast<size_t>(P.B) * P.ho * P.wo;
  const size_t plane = static_cast<size_t>(P.hi) * P.wi;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int X = static_cast<int>(idx % P.wo);
    size_t r = idx / P.wo;
    const int Y = static_cast<int>(r % P.ho);
    const int b = static_cast<int>(r / P.ho);
    int y0, y1, x0, x1;
    float ty, tx;
    resize_coord(Y, P.sy, P.hi, y0, y1, ty);
    resize_coord(X, P.sx, P.wi, x0, x1, tx);
    const float* lb = P.logits + static_cast<size_t>(b) * P.NC * plane;
    float* o = P.out + idx * P.NC;
    float mx = -INFINITY;
    for (int c = 0; c < P.NC; ++c) {
      const float* lp = lb + static_cast<size_t>(c) * plane;
      const float top = lerp_nofma(__ldg(lp + static_cast<size_t>(y0) * P.wi + x0), __ldg(lp + static_cast<size_t>(y0) * P.wi + x1), tx);
      const float bot = lerp_nofma(__ldg(lp + static_cast<size_t>(y1) * P.wi + x0), __ldg(lp + static_cast<size_t>(y1) * P.wi + x1), tx);
      const float v = lerp_nofma(top, bot, ty);
      o[c] = v;
      mx = fmaxf(mx, v);
    }
    if (P.softmax) {
      float sum = 0.0f;
      for (int c = 0; c < P.NC; ++c) {
        const float e = expf(o[c] - mx);
        o[c] = e;
        sum += e;
      }
      for (int c = 0; c < P.NC; ++c) o[c] = o[c] / sum;
    }
  }
}

// planar fp32 [B,NC,h,w] -> NHWC fp32 (debug tap)
__global__ void planar_to_nhwc_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int NC, size_t plane) {
  const size_t total = static_cast<size_t>(B) * NC * plane;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % NC);
    size_t r = idx / NC;
    const size_t p = r % plane;
    const size_t b = r / plane;
    out[idx] = in[(b * NC + c) * plane + p];
  }
}

// ------------------------------------------------------------------------------------------------
// Confusion-matrix accumulation of the evaluation loop (reference eval.py:368-373 generate_matrix + the running sum of
// eval.py:403-443): confusion[gt * NC + pred] += 1 for every pixel with gt < NC (255 = ignore label is skipped).
// Integer work: bit exact.  Block-private histogram in shared memory for NC <= 64, global atomics above.
struct ConfusionParams {
  const uint8_t* pred;
  const uint8_t* gt;
  long long n;
  int NC;
  unsigned long long* confusion;   // [NC][NC], accumulated
};
__global__ void __launch_bounds__(256) confusion_matrix_kernel(const ConfusionParams P) {
  extern __shared__ unsigned int s_hist[];   // [NC*NC] when NC <= 64
  const int bins = P.NC * P.NC;
  const bool priv = P.NC <= 64;
  if (priv) {
    for (int i = threadIdx.x; i < bins; i += blockDim.x) s_hist[i] = 0u;
    __syncthreads();
  }
  const long long n16 = P.n >> 4;
  for (long long v = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; v < n16; v += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint4 g = ldg_nc_v4(P.gt + v * 16), q = ldg_nc_v4(P.pred + v * 16);
    const uint32_t gw[4] = {g.x, g.y, g.z, g.w}, pw[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int gt = (gw[k >> 2] >> ((k & 3) * 8)) & 0xFF, pr = (pw[k >> 2] >> ((k & 3) * 8)) & 0xFF;
      if (gt < P.NC && pr < P.NC) {
        if (priv) atomicAdd(&s_hist[gt * P.NC + pr], 1u);
        else atomicAdd(&P.confusion[gt * P.NC + pr], 1ull);
      }
    }
  }
  if (blockIdx.x == 0)   // tail (n % 16 pixels)
    for (long long i = n16 * 16 + threadIdx.x; i < P.n; i += blockDim.x) {
      const int gt = P.gt[i], pr = P.pred[i];
      if (gt < P.NC && pr < P.NC) {
        if (priv) atomicAdd(&s_hist[gt * P.NC + pr], 1u);
        else atomicAdd(&P.confusion[gt * P.NC + pr], 1ull);
      }
    }
  if (priv) {
    __syncthreads();
    for (int i = threadIdx.x; i < bins; i += blockDim.x)
      if (s_hist[i]) atomicAdd(&P.confusion[i], static_cast<unsigned long long>(s_hist[i]));
  }
}


// ------------------------------------------------------------------------------------------------ image pre / post-processing (N4)
// normalize_image (common/data_utils.py:403-416): float32(x) / 127.5 - 1 — an IEEE division and a subtraction, like numpy (no
// reciprocal multiply, no contraction), so the result is bit identical.  out: fp32, or bf16 (the head's input dtype) when asked.
__global__ void __launch_bounds__(256) normalize_image_kernel(const uint8_t* __restrict__ in, size_t n, float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_bf16) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float v = __fsub_rn(__fdiv_rn(static_cast<float>(in[i]), 127.5f), 1.0f);
    if (out_f32) out_f32[i] = v;
    else out_bf16[i] = __float2bfloat16_rn(v);
  }
}
// denormalize_image (:419-433): (x * 127.5 + 127.5).astype(uint8) — separate multiply and add in fp32, truncation toward zero;
// values outside [0, 256) are clamped (numpy's cast is undefined there)
__global__ void __launch_bounds__(256) denormalize_image_kernel(const float* __restrict__ in, size_t n, uint8_t* __restrict__ out) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float v = __fadd_rn(__fmul_rn(in[i], 127.5f), 127.5f);
    out[i] = static_cast<uint8_t>(min(255, max(0, static_cast<int>(v))));
  }
}
// mask_resize (:457-477) = cv2.resize(mask, (wo, ho), interpolation=cv2.INTER_NEAREST): sx = min(floor(x * ifx), wi - 1) with
// ifx = 1 / (wo / wi) in DOUBLE precision, computed on the host exactly like OpenCV's resizeNN; integer work, bit exact.
__global__ void __launch_bounds__(256) mask_resize_nearest_kernel(const uint8_t* __restrict__ in, int B, int hi, int wi, int ho, int wo, double ify, double ifx,
                                                                  uint8_t* __restrict__ out) {
  const size_t total = static_cast<size_t>(B) * ho * wo;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total; idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int X = static_cast<int>(idx % wo);
    size_t r = idx / wo;
    const int Y = static_cast<int>(r % ho);
    const int b = static_cast<int>(r / ho);
    const int sx = min(static_cast<int>(floor(X * ifx)), wi - 1);
    const int sy = min(static_cast<int>(floor(Y * ify)), hi - 1);
    out[idx] = in[(static_cast<size_t>(b) * hi + sy) * wi + sx];
  }
}


// "resize in" of the demo / evaluation loops: preprocess_image (common/data_utils.py:436-454) = PIL Image.resize(size, Image.BICUBIC)
// on the decoded uint8 RGB image.  The arithmetic is Pillow's (third-party dependency of the reference, absent from /root/reference;
// its published algorithm, src/libImaging/Resample.c, restated): a separable convolution with the Keys cubic (a = -0.5) stretched by
// max(scale, 1) (so down-scaling is antialiased), coefficients normalised in DOUBLE on the host and rounded to 22-bit fixed point,
// horizontal pass first into a uint8 intermediate, each pass  clip8((2^21 + sum(pixel * coeff)) >> 22).  Integer work, bit exact.
// One pass over axis `axis_len_in` -> `n_out`: in [outer, n_in, inner] u8, out [outer, n_out, inner]; bounds[i] = {first, count}.
__global__ void __launch_bounds__(256) resample_pass_u8_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, long long outer, int n_in, int n_out,
                                                               int inner, const int2* __restrict__ bounds, const int* __restrict__ kk, int ksize) {
  const long long total = outer * n_out * inner;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total; idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % inner);
    long long r = idx / inner;
    const int i = static_cast<int>(r % n_out);
    const long long o = r / n_out;
    const int2 bd = __ldg(bounds + i);
    const int* k = kk + static_cast<size_t>(i) * ksize;
    const uint8_t* src = in + (o * n_in + bd.x) * inner + c;
    int ss = 1 << 21;
    for (int t = 0; t < bd.y; ++t) ss += static_cast<int>(src[static_cast<size_t>(t) * inner]) * __ldg(k + t);
    ss >>= 22;                                       // arithmetic shift, then Pillow's clip8 table = clamp to 0..255
    out[idx] = static_cast<uint8_t>(ss < 0 ? 0 : (ss > 255 ? 255 : ss));
  }
}

// Present-class set of the native post-process (inference/MNN/deeplabSegment.cpp:171-172): the non-background classes of a label map
// in order of first appearance in raster order.  The device records, per image and class, the smallest pixel index holding the class
// (0xFFFFFFFF = absent) — block-private minima in shared memory, one global atomicMin per class and block; sorting the present classes
// by that index on the host reproduces the reference's emplace_back order exactly.  first: [B][256] uint32, pre-set to 0xFFFFFFFF.
__global__ void __launch_bounds__(256) present_classes_kernel(const uint8_t* __restrict__ labels, long long n_per_image, unsigned int* __restrict__ first) {
  __shared__ unsigned int s_first[256];
  s_first[threadIdx.x] = 0xFFFFFFFFu;
  __syncthreads();
  const uint8_t* img = labels + static_cast<size_t>(blockIdx.y) * n_per_image;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n_per_image; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const unsigned int cls = img[i];
    if (s_first[cls] > static_cast<unsigned int>(i)) atomicMin(&s_first[cls], static_cast<unsigned int>(i));
  }
  __syncthreads();
  const unsigned int v = s_first[threadIdx.x];
  if (v != 0xFFFFFFFFu) atomicMin(first + static_cast<size_t>(blockIdx.y) * 256 + threadIdx.x, v);
}

// ------------------------------------------------------------------------------------------------ Jaccard metric counts (N3)
// The training metric of the reference (deeplabv3p/metrics.py:30-45, compiled in at train.py:141): per image and per class i in
// 0..NC the pixel counts  inter = #(gt == i & pred == i), true = #(gt == i), pred = #(pred == i)  (union = true + pred - inter; a
// pixel with the ignore label 255 belongs to no true class but still counts for the class it was predicted as).  Integer work,
// bit exact; the tiny float part (per-image IoU, means over the images / classes that are present) stays on the host.
// counts: [B][3][NC + 1] uint64, ACCUMULATED.  grid (blocks per image, B); block-private histograms in shared memory.
__global__ void __launch_bounds__(256) jaccard_counts_kernel(const uint8_t* __restrict__ pred, const uint8_t* __restrict__ gt, long long n_per_image, int NC,
                                                             unsigned long long* __restrict__ counts) {
  extern __shared__ unsigned int s_cnt[];   // [3][NC + 1]
  const int bins = NC + 1;
  for (int i = threadIdx.x; i < 3 * bins; i += blockDim.x) s_cnt[i] = 0u;
  __syncthreads();
  const uint8_t* p = pred + static_cast<long long>(blockIdx.y) * n_per_image;
  const uint8_t* g = gt + static_cast<long long>(blockIdx.y) * n_per_image;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n_per_image; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int t = g[i], q = p[i];
    if (t <= NC) atomicAdd(&s_cnt[bins + t], 1u);
    if (q <= NC) atomicAdd(&s_cnt[2 * bins + q], 1u);
    if (t == q && t <= NC) atomicAdd(&s_cnt[t], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * bins; i += blockDim.x)
    if (s_cnt[i]) atomicAdd(&counts[static_cast<long long>(blockIdx.y) * 3 * bins + i], static_cast<unsigned long long>(s_cnt[i]));
}

}  // namespace dlv3p
