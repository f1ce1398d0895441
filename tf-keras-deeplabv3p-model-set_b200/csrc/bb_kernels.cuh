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

// CTA = one segment of kStemSeg output pixels of one output row; thread = two of them (t and t + 128) x all 32 output channels.
// The three input rows of the segment are staged in shared memory as fp32 NORMALISED values (uint8 through a 256-entry table of the
// correctly rounded x / 127.5 - 1, so the arithmetic is the reference's), split by column parity and channel so that a warp's reads
// of a stride-2 convolution are consecutive words; zero padding of the normalised image = zeros in the staged rows.  Weights in
// shared memory (warp-uniform 16-byte broadcast reads), every weight read feeds four packed-fp32 FMAs.  fp32 arithmetic in the tap
// order (ky, kx, cin): the only rounding is the bf16 store.
constexpr int kStemSeg = 256;
constexpr int kStemIdx = kStemSeg + 8;      // even-parity entries 0..256 + pad

__global__ void __launch_bounds__(128) stem_conv_kernel(const StemParams P) {
  __shared__ __align__(16) float s_w[27 * 32];
  __shared__ __align__(8) float s_scale[32], s_shift[32];
  __shared__ float s_lut[256];
  __shared__ float s_in[3][3][2][kStemIdx];     // [ky][cin][column parity][column / 2], columns relative to the segment's first tap
  const int t = threadIdx.x;
  pdl_launch_dependents();
  for (int i = t; i < 27 * 32; i += 128) s_w[i] = P.w[i];
  if (t < 32) {
    s_scale[t] = P.scale[t];
    s_shift[t] = P.shift[t];
  }
  s_lut[t] = __fsub_rn(__fdiv_rn(static_cast<float>(t), 127.5f), 1.0f);
  s_lut[t + 128] = __fsub_rn(__fdiv_rn(static_cast<float>(t + 128), 127.5f), 1.0f);
  const int segs = (P.Wo + kStemSeg - 1) / kStemSeg;
  const int seg = static_cast<int>(blockIdx.x) % segs;
  const int row = static_cast<int>(blockIdx.x) / segs;
  const int oy = row % P.Ho, b = row / P.Ho;
  const int x0 = seg * kStemSeg;
  __syncthreads();
  pdl_wait();          // the images may come from the previous kernel of the stream; the output buffer may still be read by it
  // ---- stage: entry e of row ky is input column 2 * x0 + e - pad_l
  const int ncols = 2 * min(kStemSeg, P.Wo - x0) + 1;
  for (int i = t; i < 3 * ncols; i += 128) {
    const int ky = i / ncols, e = i - ky * ncols;
    const int iy = oy * 2 - P.pad_t + ky, ix = 2 * x0 + e - P.pad_l;
    float v0 = 0.0f, v1 = 0.0f, v2 = 0.0f;
    if (iy >= 0 && iy < P.H && ix >= 0 && ix < P.W) {
      const size_t off = ((static_cast<size_t>(b) * P.H + iy) * P.W + ix) * 3;
      if (P.img_f32) {
        const float* p = static_cast<const float*>(P.img) + off;
        v0 = __ldg(p); v1 = __ldg(p + 1); v2 = __ldg(p + 2);
      } else {
        const uint8_t* p = static_cast<const uint8_t*>(P.img) + off;
        v0 = s_lut[__ldg(p)]; v1 = s_lut[__ldg(p + 1)]; v2 = s_lut[__ldg(p + 2)];
      }
    }
    s_in[ky][0][e & 1][e >> 1] = v0;
    s_in[ky][1][e & 1][e >> 1] = v1;
    s_in[ky][2][e & 1][e >> 1] = v2;
  }
  __syncthreads();
  unsigned long long acc[2][16];
#pragma unroll
  for (int n = 0; n < 16; ++n) acc[0][n] = acc[1][n] = 0ull;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float* in = &s_in[ky][c][kx & 1][t + (kx >> 1)];
        const float a0 = in[0], a1 = in[128];      // entries past the segment's columns are never stored (pixels >= Wo)
        const unsigned long long v0 = pack_f32x2(a0, a0), v1 = pack_f32x2(a1, a1);
        const ulonglong2* wr = reinterpret_cast<const ulonglong2*>(s_w + ((ky * 3 + kx) * 3 + c) * 32);
#pragma unroll
        for (int n4 = 0; n4 < 8; ++n4) {
          const ulonglong2 w4 = wr[n4];
          ffma2(acc[0][n4 * 2], v0, w4.x);
          ffma2(acc[0][n4 * 2 + 1], v0, w4.y);
          ffma2(acc[1][n4 * 2], v1, w4.x);
          ffma2(acc[1][n4 * 2 + 1], v1, w4.y);
        }
      }
    }
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int ox = x0 + t + h * 128;
    if (ox >= P.Wo) continue;
    __nv_bfloat16* o = P.out + ((static_cast<size_t>(b) * P.Ho + oy) * P.Wo + ox) * 32;
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      uint32_t pk[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int n = 2 * (j + q);
        const float x = fmaxf(fmaf(f32x2_lo(acc[h][j + q]), s_scale[n], s_shift[n]), 0.0f);
        const float y = fmaxf(fmaf(f32x2_hi(acc[h][j + q]), s_scale[n + 1], s_shift[n + 1]), 0.0f);
        pk[q] = pack_bf16x2(x, y);
      }
      stg_v4(o + j * 2, make_uint4(pk[0], pk[1], pk[2], pk[3]));
    }
  }
}
inline unsigned stem_grid(int B, int Ho, int Wo) { return static_cast<unsigned>(B) * Ho * ((Wo + kStemSeg - 1) / kStemSeg); }

// ---------------------------------------------------------------------------------------------------------------- depthwise
struct BbDwParams {
  const CUtensorMap* tmap_x;   // 4D {C, W, H, B} bf16, box {64, IW, IH, 1}, no swizzle, OOB -> 0 (= ZeroPadding2D / 'same')
  const float* w;              // [9][Cpad] fp32 taps with the BN scale folded in (Cpad = 64-channel groups, zero padded)
  const float* shift;          // [Cpad]
  __nv_bfloat16* out;          // [B, Ho, Wo, C]
  int B, C, Cpad, Ho, Wo;
  int tiles_x, tiles_y, cgroups;
  int relu_in, relu_out;
  int debug;                   // benchmark aid: bit0 skip the global stores, bit1 skip the stencil arithmetic
};

template <int S, int R, int TH, int TW>
struct BbDwCfg {
  static constexpr int IH = (TH - 1) * S + 2 * R + 1;
  static constexpr int IW = (TW - 1) * S + 2 * R + 1;
  static constexpr int kInBytes = IH * IW * 128;
  static constexpr int kStageBytes = kInBytes + 128;           // + the decoded item (tx, ty, b)
  static constexpr int kStages = 2;
  static constexpr int kBlocks = (TH / 4) * (TW / 4);          // 4 x 4 output blocks per tile
  static constexpr int kWarps = kBlocks < 8 ? kBlocks : 8;     // compute warps
  static constexpr int kThreads = (kWarps + 1) * 32;           // + the producer warp
  static constexpr int kSmemBytes = kStages * kStageBytes + 128 + 64;   // + alignment slack + the mbarriers
};

// Persistent CTAs over (image, TH x TW output tile, 64-channel group) work items.  The grid is a multiple of the number of
// channel groups, so a CTA keeps ONE group for its whole life (taps + shift live in registers) and walks the spatial tiles;
// concurrent CTAs read neighbouring 128-byte segments of the same pixels.  A producer warp keeps a 2-stage ring full: per item ONE
// TMA box brings the input window (halo included; out-of-bounds rows / columns / channels arrive as zeros = ZeroPadding2D /
// 'same') into shared memory pixel-major [IH][IW][64 ch].  Compute warp = a 4 x 4 block of output pixels,
// lane = one channel pair: every shared-memory access of a warp is one conflict-free 128-byte pixel row, every global store one
// full 128-byte line of a pixel.  Rolling window over the input rows of the block: each input value is loaded once per block
// and feeds up to nine packed-fp32 FMAs.
template <int S, int R, int TH, int TW>
__global__ void __launch_bounds__(BbDwCfg<S, R, TH, TW>::kThreads) bb_depthwise_kernel(const __grid_constant__ BbDwParams P) {
  using Cfg = BbDwCfg<S, R, TH, TW>;
  extern __shared__ __align__(128) uint8_t smem_dw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dw) + 127) & ~static_cast<uintptr_t>(127));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);   // [kStages]
  uint64_t* empty_bar = full_bar + Cfg::kStages;                                              // [kStages]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = P.B * P.tiles_y * P.tiles_x;
  const int g = static_cast<int>(blockIdx.x) % P.cgroups;                 // gridDim.x % cgroups == 0 (host)
  const int t_first = static_cast<int>(blockIdx.x) / P.cgroups, t_step = static_cast<int>(gridDim.x) / P.cgroups;
  if (threadIdx.x == 0) {
    for (int i = 0; i < Cfg::kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], Cfg::kWarps);
    }
    fence_barrier_init();
  }
  pdl_launch_dependents();
  __syncthreads();
  if (warp == Cfg::kWarps) {
    // ------------------------------------------------------------------ producer warp
    pdl_wait();          // the input is the previous kernel's output; the compute warps are ordered behind this through full_bar
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int t = t_first; t < num_tiles; t += t_step) {
        const int tx = t % P.tiles_x;
        const int r = t / P.tiles_x;
        const int ty = r % P.tiles_y;
        const int b = r / P.tiles_y;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* st = smem + stage * Cfg::kStageBytes;
        *reinterpret_cast<int4*>(st + Cfg::kInBytes) = make_int4(tx, ty, b, 0);   // decoded once: the consumers skip the divisions
        mbar_arrive_expect_tx(&full_bar[stage], Cfg::kInBytes);                  // release: the store above is visible to the waiters
        tma_load_4d(st, P.tmap_x, &full_bar[stage], g * 64, tx * TW * S - R, ty * TH * S - R, b, kEvictNormal);
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
    }
    return;
  }
  // -------------------------------------------------------------------- compute warps
  constexpr int WW = 3 * S + 2 * R + 1;      // input window of a 4 x 4 output block
  // taps + shift of this lane's channel pair: the CTA's group never changes
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
  uint32_t stage = 0, phase = 0;
  const uint32_t pstride = static_cast<uint32_t>(P.C) >> 1;            // one pixel / one output row in 32-bit words
  const uint32_t rstride = static_cast<uint32_t>(P.Wo) * pstride;
  for (int t = t_first; t < num_tiles; t += t_step) {
    const uint8_t* st = smem + stage * Cfg::kStageBytes;
    mbar_wait(&full_bar[stage], phase);
    const int4 it4 = *reinterpret_cast<const int4*>(st + Cfg::kInBytes);
    const int tx = it4.x, ty = it4.y, b = it4.z;
    const bool ch_ok = g * 64 + lane * 2 < P.C;
    for (int bi = warp; bi < Cfg::kBlocks; bi += Cfg::kWarps) {
      const int by = bi / (TW / 4), bx = bi % (TW / 4);
      const int oy0 = ty * TH + by * 4, ox0 = tx * TW + bx * 4;
      if (oy0 >= P.Ho || ox0 >= P.Wo) continue;          // block entirely outside the image (partial tiles)
      const uint32_t base = smem_u32(st) + ((by * 4 * S) * Cfg::IW + bx * 4 * S) * 128 + lane * 4;
      unsigned long long acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = sh;
#pragma unroll
      for (int wr = 0; wr < ((P.debug & 2) ? 0 : WW); ++wr) {
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
      // one 64-bit address per block, 32-bit word offsets from there
      uint32_t* o0 = reinterpret_cast<uint32_t*>(P.out + ((static_cast<size_t>(b) * P.Ho + oy0) * P.Wo + ox0) * P.C + g * 64) + lane;
      if (P.relu_out) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = static_cast<unsigned long long>(f32x2_to_bf16x2_relu(acc[i][j]));
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = static_cast<unsigned long long>(f32x2_to_bf16x2(acc[i][j]));
      }
      if (P.debug & 1) continue;
      if (ch_ok && oy0 + 4 <= P.Ho && ox0 + 4 <= P.Wo) {      // interior block: 16 unpredicated stores
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) o0[i * rstride + j * pstride] = static_cast<uint32_t>(acc[i][j]);
      } else if (ch_ok) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (oy0 + i < P.Ho && ox0 + j < P.Wo) o0[i * rstride + j * pstride] = static_cast<uint32_t>(acc[i][j]);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[stage]);       // this warp is done reading the stage
    if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
  }
}

// ---------------------------------------------------------------------------------------------------------------- subsample
// out[b, i, j, :] = x[b, 2i, 2j, :]  (16-byte vectors; C % 8 == 0)
__global__ void __launch_bounds__(256) subsample2_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, int B, int H, int W, int C,
                                                         int Ho, int Wo) {
  pdl_launch_dependents();
  pdl_wait();
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
