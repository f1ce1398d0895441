// aspp_dw_fast.cuh — geometry-specialised ASPP depthwise kernel for small feature maps (the whole h x w map of a
// 32-channel group fits in shared memory twice per SM).
//
// Replaces, in ONE pass over x: the three `aspp{1,2,3}_depthwise` dilated 3x3 convs + `_depthwise_BN` + ReLU
// (reference deeplabv3p/models/layers.py:146-153 -> :100-104) and the AveragePooling2D of the image-pooling branch (:132).
//
// A 3x3 conv with dilation R only couples pixels of equal phase (i mod R, j mod R): it is R*R independent dense 3x3 convs
// on phase images of at most ceil(H/R) x ceil(W/R) pixels.  With H, W and the rates known at compile time a phase image
// lives entirely in registers: every input is read from shared memory ONCE per rate, every address is an immediate
// offset from one base register, border taps are resolved at compile time (zero padding = absent FMAs).
//
//   CTA            one (image, 32-channel group) slab [H*W][32] bf16 = 64 B per pixel, staged by TMA; 2-3 CTAs per SM so
//                  one CTA's HBM load overlaps the others' arithmetic
//   half-warp      one phase image: 16 lanes x bf16x2 = the 32 channels; the two halves of a warp take phases whose
//                  column parity differs, so their 64-byte rows never share a shared-memory bank
//   warp           fetches work units (heaviest rate first) from a shared-memory counter
#pragma once

#include "mem_kernels.cuh"

namespace dlv3p {

constexpr int kAsppFastThreads = 256;


template <int H, int W, int R>
__device__ __forceinline__ void aspp_fast_item(const uint8_t* slab_lane, uint8_t* out_lane, int pi, int pj,
                                               const unsigned long long (&wt)[9], unsigned long long sh, bool store) {
  constexpr int NA = (H + R - 1) / R, NT = (W + R - 1) / R;
  const bool last_row = pi + (NA - 1) * R < H;   // does the last row / column of the phase image exist?
  const bool last_col = pj + (NT - 1) * R < W;
  const uint8_t* src = slab_lane + (pi * W + pj) * 64;
  uint8_t* dst = out_lane + (pi * W + pj) * 128;
  unsigned long long x[NA][NT];
#pragma unroll
  for (int a = 0; a < NA; ++a)
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      const bool ok = (a < NA - 1 || last_row) && (t < NT - 1 || last_col);
      uint32_t v = 0u;
      if (ok) v = *reinterpret_cast<const uint32_t*>(src + (a * R * W + t * R) * 64);
      x[a][t] = f32x2_from_bf16x2(v);
    }
#pragma unroll
  for (int a = 0; a < NA; ++a)
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      unsigned long long acc = sh;
#pragma unroll
      for (int u = 0; u < 3; ++u)
#pragma unroll
        for (int v = 0; v < 3; ++v) {
          const int aa = a + u - 1, tt = t + v - 1;
          if (aa >= 0 && aa < NA && tt >= 0 && tt < NT) ffma2_acc(acc, wt[u * 3 + v], x[aa][tt]);   // compile-time
        }
      const bool ok = (a < NA - 1 || last_row) && (t < NT - 1 || last_col);
      if (ok && store) *reinterpret_cast<uint32_t*>(dst + (a * R * W + t * R) * 128) = f32x2_to_bf16x2_relu(acc);
    }
}

// ---- the reference's rate triples are (r, 2r, 3r) (layers.py:118-125: 6/12/18, 12/24/36, 3/6/9) ------------------------
// The phase image of rate r, x[pi + r*a][pj + r*t], contains the phase images of rates 2r and 3r: on it the three
// atrous convs are 3x3 convs with dilation 1, 2 and 3.  One register-resident phase image therefore yields the outputs
// of ALL THREE rates at its pixels: one shared-memory load and one bf16 unpack per input instead of three, and 3x
// fewer work items.
template <int H, int W, int R, int K>
__device__ __forceinline__ void aspp_fast_emit(const unsigned long long (&x)[(H + R - 1) / R][(W + R - 1) / R], uint8_t* dst,
                                               const float* s_w, const float* s_shift, int l16, bool last_row, bool last_col, bool store) {
  constexpr int NA = (H + R - 1) / R, NT = (W + R - 1) / R;
  unsigned long long wt[9], sh;
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const float2 v = *reinterpret_cast<const float2*>(s_w + ((K - 1) * 9 + t) * 32 + l16 * 2);
    wt[t] = pack2(v.x, v.y);
  }
  {
    const float2 v = *reinterpret_cast<const float2*>(s_shift + (K - 1) * 32 + l16 * 2);
    sh = pack2(v.x, v.y);
  }
#pragma unroll
  for (int a = 0; a < NA; ++a)
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      unsigned long long acc = sh;
#pragma unroll
      for (int u = 0; u < 3; ++u)
#pragma unroll
        for (int v = 0; v < 3; ++v) {
          const int aa = a + (u - 1) * K, tt = t + (v - 1) * K;
          if (aa >= 0 && aa < NA && tt >= 0 && tt < NT) ffma2_acc(acc, wt[u * 3 + v], x[aa][tt]);   // compile-time
        }
      const bool ok = (a < NA - 1 || last_row) && (t < NT - 1 || last_col);
      if (ok && store) *reinterpret_cast<uint32_t*>(dst + (a * R * W + t * R) * 128) = f32x2_to_bf16x2_relu(acc);
    }
}

template <int H, int W, int R>
struct AsppFast3Cfg {
  static_assert(R % 2 == 0 && W % 2 == 0, "phase pairing needs an even rate and width");
  static constexpr int kPix = H * W;
  static constexpr int kBoxes = (kPix + 255) / 256;                 // TMA boxes of 256 pixels x 32 channels (16 KB)
  static constexpr int kItems = R * (R / 2);                        // warp items: pairs of phases of opposite column parity
  static constexpr int kWarps = kItems % 6 == 0 ? 6 : 8;            // 18 items -> 6 warps x 3 (the phase image needs > 100 registers)
  static constexpr int kThreads = kWarps * 32;
  static constexpr int kSmemBytes = kBoxes * 16384 + (27 * 32 + 3 * 32 + kWarps * 32) * 4 + 32;
};

// grid = B * (C / 32) CTAs.  P.tmap_slab: 2D [B*H*W, C] bf16 view of x, box {32, 256}, no swizzle.
template <int H, int W, int R>
__global__ void __maxnreg__(112) aspp_dw_fast3_kernel(const __grid_constant__ AsppDwParams P) {
  using Cfg = AsppFast3Cfg<H, W, R>;
  constexpr int NA = (H + R - 1) / R, NT = (W + R - 1) / R;
  extern __shared__ __align__(128) uint8_t fast_smem[];
  uint8_t* s_slab = fast_smem;
  float* s_w = reinterpret_cast<float*>(s_slab + Cfg::kBoxes * 16384);   // [3][9][32]
  float* s_shift = s_w + 27 * 32;                                         // [3][32]
  float* s_red = s_shift + 3 * 32;                                        // [kWarps][32]
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_red + Cfg::kWarps * 32);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int hw = lane >> 4, l16 = lane & 15;
  const int ngroups = P.C >> 5;
  const int b = blockIdx.x / ngroups, grp = blockIdx.x - b * ngroups;   // 32-channel group
  if (tid == 0) {
    mbar_init(s_bar, 1);
    fence_barrier_init();
    mbar_arrive_expect_tx(s_bar, static_cast<uint32_t>(Cfg::kBoxes) * 16384u);
    for (int q = 0; q < Cfg::kBoxes; ++q)
      tma_load_2d(s_slab + q * 16384, P.tmap_slab, s_bar, grp * 32, b * Cfg::kPix + q * 256, kEvictFirst);
  }
  for (int i = tid; i < 27 * 32; i += Cfg::kThreads) s_w[i] = __ldg(P.w + static_cast<size_t>(i >> 5) * P.C + grp * 32 + (i & 31));
  if (tid < 3 * 32) s_shift[tid] = __ldg(P.shift + static_cast<size_t>(tid >> 5) * P.C + grp * 32 + (tid & 31));
  __syncthreads();
  mbar_wait(s_bar, 0);

  const uint8_t* slab_lane = s_slab + l16 * 4;
  const bool store = !(P.debug & 1);
  const size_t rate_stride = static_cast<size_t>(P.nchunks) * P.B * Cfg::kPix * 128;   // bytes per rate
  uint8_t* out0 = reinterpret_cast<uint8_t*>(P.out) + (static_cast<size_t>(grp >> 1) * P.B + b) * Cfg::kPix * 128 + (grp & 1) * 64 + l16 * 4;
  float sx = 0.0f, sy = 0.0f;   // image-pooling partial sums: every pixel belongs to exactly one phase image
  for (int it = warp; it < Cfg::kItems; it += Cfg::kWarps) {
    const int pi = it / (R / 2), pj = 2 * (it % (R / 2)) + hw;
    const bool last_row = pi + (NA - 1) * R < H;   // does the last row / column of the phase image exist?
    const bool last_col = pj + (NT - 1) * R < W;
    const uint8_t* src = slab_lane + (pi * W + pj) * 64;
    unsigned long long x[NA][NT];
#pragma unroll
    for (int a = 0; a < NA; ++a)
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const bool ok = (a < NA - 1 || last_row) && (t < NT - 1 || last_col);
        uint32_t v = 0u;
        if (ok) v = *reinterpret_cast<const uint32_t*>(src + (a * R * W + t * R) * 64);
        x[a][t] = f32x2_from_bf16x2(v);
        sx += bf16_lo(v);
        sy += bf16_hi(v);
      }
    uint8_t* dst = out0 + (pi * W + pj) * 128;
    aspp_fast_emit<H, W, R, 1>(x, dst, s_w, s_shift, l16, last_row, last_col, store);
    aspp_fast_emit<H, W, R, 2>(x, dst + rate_stride, s_w, s_shift, l16, last_row, last_col, store);
    aspp_fast_emit<H, W, R, 3>(x, dst + 2 * rate_stride, s_w, s_shift, l16, last_row, last_col, store);
  }
  sx += __shfl_xor_sync(0xFFFFFFFFu, sx, 16);
  sy += __shfl_xor_sync(0xFFFFFFFFu, sy, 16);
  if (hw == 0) {
    s_red[warp * 32 + l16 * 2] = sx;
    s_red[warp * 32 + l16 * 2 + 1] = sy;
  }
  __syncthreads();
  if (tid < 32) {
    float s = 0.0f;
#pragma unroll
    for (int q = 0; q < Cfg::kWarps; ++q) s += s_red[q * 32 + tid];
    P.pool_partial[static_cast<size_t>(b) * P.C + grp * 32 + tid] = s;   // pool_items == 1 on this path
  }
}

}  // namespace dlv3p
