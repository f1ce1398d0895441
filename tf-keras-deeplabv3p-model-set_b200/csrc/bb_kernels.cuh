// bb_kernels.cuh — the memory-bound kernels of the Xception backbone:
//   stem_conv_kernel        entry_flow_conv1_1: [normalize_image] -> Conv2D(32, 3x3, strides 2, 'same') -> BN -> ReLU   (deeplabv3p_xception.py:119-123,
//                           common/data_utils.py:403-416 for the uint8 input)
//   bb_depthwise_kernel     the depthwise half of SepConv_BN as the backbone uses it (layers.py:74-111): [ZeroPadding2D] -> [ReLU] ->
//                           DepthwiseConv2D 3x3 (stride 1 'same' | stride 2 'valid' after explicit padding, dilation rate) -> BN -> [ReLU]
//   subsample2_kernel       the stride-2 sampling of a 1x1 shortcut convolution (_conv2d_same with kernel_size 1: no padding, :44-52)
// Activations are NHWC bf16; depthwise outputs are the A operand [pixels, channels] of the pointwise tcgen05 GEMM (bb_gemm.cuh).
#pragma once

#include <cuda.h>

#include "dwpw_gemm.cuh"     // packed fp32x2 helpers
#include "sm100_prims.cuh"

namespace dlv3p {

// ---------------------------------------------------------------------------------------------------------------- stem
struct StemParams {
  const void* img;        // [B, H, W, 3] uint8 (img_f32 == 0: normalised here, x / 127.5 - 1) or fp32 already normalised
  int img_f32;
  const float* w;         // [27][32] fp32, tap-major (ky, kx, cin), Keras HWIO order
  const float* scale;     // [32] folded BN
  const float* shift;     // [32]
  __nv_bfloat16* out;     // [B, Ho, Wo, 32]
  int B, H, W, Ho, Wo, pad_t, pad_l;
};

// thread = one output pixel x 32 output channels; weights in shared memory (warp-uniform broadcast reads).  fp32 arithmetic
// on the exact normalised pixel values: the only rounding is the bf16 store.
__global__ void __launch_bounds__(128) stem_conv_kernel(const StemParams P) {
  __shared__ __align__(16) float s_w[27 * 32];
  __shared__ float s_scale[32], s_shift[32];
  for (int i = threadIdx.x; i < 27 * 32; i += blockDim.x) s_w[i] = P.w[i];
  if (threadIdx.x < 32) {
    s_scale[threadIdx.x] = P.scale[threadIdx.x];
    s_shift[threadIdx.x] = P.shift[threadIdx.x];
  }
  __syncthreads();
  const long long total = static_cast<long long>(P.B) * P.Ho * P.Wo;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  const int ox = static_cast<int>(idx % P.Wo);
  const int oy = static_cast<int>((idx / P.Wo) % P.Ho);
  const int b = static_cast<int>(idx / (static_cast<long long>(P.Wo) * P.Ho));
  float acc[32];
#pragma unroll
  for (int n = 0; n < 32; ++n) acc[n] = 0.0f;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int iy = oy * 2 - P.pad_t + ky;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int ix = ox * 2 - P.pad_l + kx;
      float v[3] = {0.0f, 0.0f, 0.0f};      // zero padding of the NORMALISED image
      if (iy >= 0 && iy < P.H && ix >= 0 && ix < P.W) {
        const size_t off = ((static_cast<size_t>(b) * P.H + iy) * P.W + ix) * 3;
        if (P.img_f32) {
          const float* p = static_cast<const float*>(P.img) + off;
          v[0] = __ldg(p); v[1] = __ldg(p + 1); v[2] = __ldg(p + 2);
        } else {
          const uint8_t* p = static_cast<const uint8_t*>(P.img) + off;
          v[0] = __fsub_rn(__fdiv_rn(static_cast<float>(__ldg(p)), 127.5f), 1.0f);
          v[1] = __fsub_rn(__fdiv_rn(static_cast<float>(__ldg(p + 1)), 127.5f), 1.0f);
          v[2] = __fsub_rn(__fdiv_rn(static_cast<float>(__ldg(p + 2)), 127.5f), 1.0f);
        }
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float4* wr = reinterpret_cast<const float4*>(s_w + ((ky * 3 + kx) * 3 + c) * 32);
#pragma unroll
        for (int n4 = 0; n4 < 8; ++n4) {
          const float4 w4 = wr[n4];
          acc[n4 * 4 + 0] = fmaf(v[c], w4.x, acc[n4 * 4 + 0]);
          acc[n4 * 4 + 1] = fmaf(v[c], w4.y, acc[n4 * 4 + 1]);
          acc[n4 * 4 + 2] = fmaf(v[c], w4.z, acc[n4 * 4 + 2]);
          acc[n4 * 4 + 3] = fmaf(v[c], w4.w, acc[n4 * 4 + 3]);
        }
      }
    }
  }
  __nv_bfloat16* o = P.out + static_cast<size_t>(idx) * 32;
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    uint32_t pk[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float a = fmaxf(fmaf(acc[j + 2 * q], s_scale[j + 2 * q], s_shift[j + 2 * q]), 0.0f);
      const float c = fmaxf(fmaf(acc[j + 2 * q + 1], s_scale[j + 2 * q + 1], s_shift[j + 2 * q + 1]), 0.0f);
      pk[q] = pack_bf16x2(a, c);
    }
    stg_v4(o + j, make_uint4(pk[0], pk[1], pk[2], pk[3]));
  }
}

// ---------------------------------------------------------------------------------------------------------------- depthwise
struct BbDwParams {
  const CUtensorMap* tmap_x;   // 4D {C, W, H, B} bf16, box {64, IW, IH, 1}, no swizzle, OOB -> 0 (= ZeroPadding2D / 'same')
  const float* w;              // [9][Cpad] fp32 taps with the BN scale folded in (Cpad = 64-channel groups, zero padded)
  const float* shift;          // [Cpad]
  __nv_bfloat16* out;          // [B, Ho, Wo, C]
  int B, C, Cpad, Ho, Wo;
  int tiles_x, tiles_y, cgroups;
  int relu_in, relu_out;
};

template <int S, int R, int TH, int TW>
struct BbDwCfg {
  static constexpr int IH = (TH - 1) * S + 2 * R + 1;
  static constexpr int IW = (TW - 1) * S + 2 * R + 1;
  static constexpr int kInBytes = IH * IW * 128;
  static constexpr int kBlocks = (TH / 4) * (TW / 4);          // 4 x 4 output blocks per tile
  static constexpr int kWarps = kBlocks < 8 ? kBlocks : 8;
  static constexpr int kThreads = kWarps * 32;
  static constexpr int kSmemBytes = kInBytes + 128 + 16;       // + alignment slack + the mbarrier
};

// CTA = (image, TH x TW output tile, 64-channel group): ONE TMA box brings the input window (halo included; out-of-bounds
// rows / columns / channels arrive as zeros) into shared memory pixel-major [IH][IW][64 ch].  Warp = a 4 x 4 block of
// output pixels, lane = one channel pair: every shared-memory access of a warp is one conflict-free 128-byte pixel row,
// every global store one full 128-byte line of a pixel.  Rolling window over the input rows of the block: each input value
// is loaded once per block and feeds up to nine packed-fp32 FMAs.
template <int S, int R, int TH, int TW>
__global__ void __launch_bounds__(BbDwCfg<S, R, TH, TW>::kThreads) bb_depthwise_kernel(const __grid_constant__ BbDwParams P) {
  using Cfg = BbDwCfg<S, R, TH, TW>;
  extern __shared__ __align__(128) uint8_t smem_dw[];
  uint8_t* smem_in = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dw) + 127) & ~static_cast<uintptr_t>(127));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_in + Cfg::kInBytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int t = blockIdx.x;
  const int g = t % P.cgroups; t /= P.cgroups;
  const int tx = t % P.tiles_x; t /= P.tiles_x;
  const int ty = t % P.tiles_y;
  const int b = t / P.tiles_y;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, Cfg::kInBytes);
    tma_load_4d(smem_in, P.tmap_x, bar, g * 64, tx * TW * S - R, ty * TH * S - R, b, kEvictNormal);
  }
  // taps + shift of this lane's channel pair (global, L2-resident; overlaps the TMA)
  unsigned long long wt[9], sh;
  {
    const float* wp = P.w + g * 64 + lane * 2;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const float2 w2 = __ldg(reinterpret_cast<const float2*>(wp + static_cast<size_t>(k) * P.Cpad));
      wt[k] = pack_f32x2(w2.x, w2.y);
    }
    const float2 s2 = __ldg(reinterpret_cast<const float2*>(P.shift + g * 64 + lane * 2));
    sh = pack_f32x2(s2.x, s2.y);
  }
  const bool ch_ok = g * 64 + lane * 2 < P.C;
  mbar_wait(bar, 0);
  constexpr int WW = 3 * S + 2 * R + 1;      // input window of a 4 x 4 output block
  for (int bi = warp; bi < Cfg::kBlocks; bi += Cfg::kWarps) {
    const int by = bi / (TW / 4), bx = bi % (TW / 4);
    const uint32_t base = smem_u32(smem_in) + ((by * 4 * S) * Cfg::IW + bx * 4 * S) * 128 + lane * 4;
    unsigned long long acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = sh;
#pragma unroll
    for (int wr = 0; wr < WW; ++wr) {
      unsigned long long x[WW];
#pragma unroll
      for (int cc = 0; cc < WW; ++cc) {
        // only the columns some tap of some output column reads
        bool used = false;
#pragma unroll
        for (int oc = 0; oc < 4; ++oc)
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) used = used || (oc * S + dx * R == cc);
        if (used) {
          uint32_t raw = lds_u32(base + (wr * Cfg::IW + cc) * 128);
          if (P.relu_in) raw = relu_bf16x2(raw);
          x[cc] = bf16x2_to_f32x2(raw);
        } else {
          x[cc] = 0ull;
        }
      }
#pragma unroll
      for (int orow = 0; orow < 4; ++orow) {
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
          if (orow * S + dy * R != wr) continue;
#pragma unroll
          for (int oc = 0; oc < 4; ++oc)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) ffma2(acc[orow][oc], wt[dy * 3 + dx], x[oc * S + dx * R]);
        }
      }
    }
    const int oy0 = ty * TH + by * 4, ox0 = tx * TW + bx * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int oy = oy0 + i;
      if (oy >= P.Ho) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int ox = ox0 + j;
        if (ox >= P.Wo || !ch_ok) continue;
        const uint32_t v = P.relu_out ? f32x2_to_bf16x2_relu(acc[i][j]) : f32x2_to_bf16x2(acc[i][j]);
        uint32_t* o = reinterpret_cast<uint32_t*>(P.out + ((static_cast<size_t>(b) * P.Ho + oy) * P.Wo + ox) * P.C + g * 64) + lane;
        *o = v;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------- subsample
// out[b, i, j, :] = x[b, 2i, 2j, :]  (16-byte vectors; C % 8 == 0)
__global__ void __launch_bounds__(256) subsample2_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, int B, int H, int W, int C,
                                                         int Ho, int Wo) {
  const int cv = C / 8;
  const long long total = static_cast<long long>(B) * Ho * Wo * cv;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total; idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c8 = static_cast<int>(idx % cv);
    long long p = idx / cv;
    const int j = static_cast<int>(p % Wo); p /= Wo;
    const int i = static_cast<int>(p % Ho);
    const int b = static_cast<int>(p / Ho);
    const uint4 v = ldg_nc_v4(x + ((static_cast<size_t>(b) * H + 2 * i) * W + 2 * j) * C + c8 * 8);
    stg_v4(out + static_cast<size_t>(idx) * 8, v);
  }
}

}  // namespace dlv3p
