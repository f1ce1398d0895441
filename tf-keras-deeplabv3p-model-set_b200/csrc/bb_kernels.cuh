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

// Persistent CTAs over work items = a segment of kStemSeg output pixels on two consecutive output rows.  Warp = one of the rows x one
// half of the output channels; lane = eight pixels of the segment (lane + 32 j) x 16 channels: every weight word read from shared
// memory (warp-uniform 16-byte broadcast) feeds eight FMAs of the thread and every input word sixteen — broadcast reads deliver one
// word per shared-memory wavefront, so at 2 pixels x 32 channels per thread the kernel was bound by wavefronts, not by the FP32 pipe.
// The five input rows of an item are staged in shared memory as fp32 NORMALISED values (uint8 through a 256-entry table of the
// correctly rounded x / 127.5 - 1, so the arithmetic is the reference's), split by column parity and channel so that a warp's reads
// of the stride-2 convolution are consecutive words; zero padding of the normalised image = zeros in the staged rows.  The raw
// pixels of the NEXT item are on their way into shared memory (cp.async) while the current one is computed (uint8 images), so the
// trip to HBM hides behind the arithmetic.  fp32 FMAs in the tap order (ky, kx, cin): the only rounding is the bf16 store.
constexpr int kStemSeg = 256;
constexpr int kStemIdx = kStemSeg + 8;                     // even-parity entries 0..256 + pad
constexpr int kStemPer = (2 * kStemSeg + 1 + 127) / 128;   // staged entries per thread and input row

struct StemItem { int b, oy0, x0; };
__device__ __forceinline__ StemItem stem_item(const StemParams& P, int item) {
  const int segs = (P.Wo + kStemSeg - 1) / kStemSeg, row_pairs = (P.Ho + 1) / 2;
  const int seg = item % segs, rp = item / segs;
  return StemItem{rp / row_pairs, (rp % row_pairs) * 2, seg * kStemSeg};
}
// The raw bytes of an item's five input rows travel global -> shared memory as asynchronous 4-byte copies (cp.async: no registers,
// nothing waits until the bytes are converted): row r of the item is the byte stream that starts at pixel (2 oy0 - pad_t + r,
// 2 x0 - pad_l), copied from the 4-byte boundary below it (shift s_sh[r]) in whole words.  Bytes outside the image (padding columns,
// rows above / below, the neighbouring row's pixels inside the first / last word) are masked by geometry when the row is converted;
// a word that straddles the ends of the image buffer is copied byte by byte.
constexpr int kStemRawWords = (3 * (2 * kStemSeg + 1) + 3 + 3) / 4 + 3;      // + the words thread 127 reads past the stream
__device__ __forceinline__ void stem_prefetch_u8(const StemParams& P, const StemItem& it, int t, uint32_t (*s_raw)[kStemRawWords], int* s_sh) {
  const int ncols = 2 * min(kStemSeg, P.Wo - it.x0) + 1;
  const long long total = static_cast<long long>(P.B) * P.H * P.W * 3;
  const uint8_t* img = static_cast<const uint8_t*>(P.img);
#pragma unroll
  for (int r = 0; r < 5; ++r) {
    const int iy = it.oy0 * 2 - P.pad_t + r;
    if (iy < 0 || iy >= P.H) continue;
    const long long gs = ((static_cast<long long>(it.b) * P.H + iy) * P.W + (2 * it.x0 - P.pad_l)) * 3;     // stream start (may be -3)
    const int sh = static_cast<int>((reinterpret_cast<uintptr_t>(img) + gs) & 3);
    if (t == 0) s_sh[r] = sh;
    const int words = (sh + 3 * ncols + 3) >> 2;
    const long long first = gs - sh;                                   // byte offset of word 0 in the image buffer
    const uint32_t dst0 = smem_u32(&s_raw[r][0]);
    if (first >= 0 && first + 4ll * words <= total) {                  // all but the first / last row of the buffer
      const uint8_t* src = img + first;
      for (int w = t; w < words; w += 128)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst0 + 4 * w), "l"(src + 4 * w) : "memory");
    } else {
      for (int w = t; w < words; w += 128) {
        const long long off = first + 4ll * w;
        uint32_t v = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (off + k >= 0 && off + k < total) v |= static_cast<uint32_t>(__ldg(img + off + k)) << (8 * k);
        s_raw[r][w] = v;
      }
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

// Warp-specialised: warps 0-3 do nothing but the FMAs and the bf16 stores of an item (one FP32-pipe-bound warp per SM sub-partition),
// warps 4-7 stage the NEXT item into the other half of a double-buffered s_in (copy -> table lookup -> parity-split store), handing
// buffers over through two pairs of mbarriers.  With both jobs in the same warps (any number of CTAs per SM) the FMA pipe idled
// through 70 % of the kernel: the register budget of the 8 x 16 accumulator block leaves two warps per sub-partition, too few to
// cover the staging's dependent shared-memory chains.
template <bool kU8>
__global__ void __launch_bounds__(256, 1) stem_conv_kernel(const StemParams P) {
  __shared__ __align__(16) float s_w[27 * 32];
  __shared__ __align__(8) float s_scale[32], s_shift[32];
  __shared__ float s_lut[256];
  extern __shared__ __align__(16) float s_in_dyn[];
  float (*s_in)[5][3][2][kStemIdx] = reinterpret_cast<float (*)[5][3][2][kStemIdx]>(s_in_dyn);   // [buffer][input row][cin][column parity][column / 2],
                                                                                                 // columns relative to the item's first tap
  __shared__ __align__(16) uint32_t s_raw[kU8 ? 2 : 1][kU8 ? 5 : 1][kStemRawWords];     // copies of item k and k + 1 in flight
  __shared__ int s_sh[2][5];
  __shared__ __align__(8) uint64_t s_full[2], s_empty[2];
  const int t = threadIdx.x & 127, warp = (threadIdx.x >> 5) & 3, lane = threadIdx.x & 31;
  const bool helper = threadIdx.x >= 128;
  pdl_launch_dependents();
  for (int i = threadIdx.x; i < 27 * 32; i += 256) s_w[i] = P.w[i];
  if (threadIdx.x < 32) {
    s_scale[threadIdx.x] = P.scale[threadIdx.x];
    s_shift[threadIdx.x] = P.shift[threadIdx.x];
  }
  s_lut[threadIdx.x] = __fsub_rn(__fdiv_rn(static_cast<float>(threadIdx.x), 127.5f), 1.0f);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 128);
      mbar_init(&s_empty[i], 128);
    }
    fence_barrier_init();
  }
  const int num_items = P.B * ((P.Ho + 1) / 2) * ((P.Wo + kStemSeg - 1) / kStemSeg);
  const int step = static_cast<int>(gridDim.x);
  __syncthreads();
  pdl_wait();          // the images may come from the previous kernel of the stream; the output buffer may still be read by it
  if (helper) {
    // -------------------------------------------------------------------------------------------- staging warps
    int item = static_cast<int>(blockIdx.x);
    if (kU8) {
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        if (item + a * step < num_items) stem_prefetch_u8(P, stem_item(P, item + a * step), t, s_raw[a], s_sh[a]);
        else asm volatile("cp.async.commit_group;" ::: "memory");
      }
    }
    for (uint32_t k = 0; item < num_items; item += step, ++k) {
      const StemItem it = stem_item(P, item);
      const int ncols = 2 * min(kStemSeg, P.Wo - it.x0) + 1;
      float (*in)[3][2][kStemIdx] = s_in[k & 1];
      mbar_wait(&s_empty[k & 1], ((k >> 1) & 1) ^ 1);      // the FMA warps are done with this buffer (first two uses: free)
      if (kU8) {
        asm volatile("cp.async.wait_group 1;" ::: "memory");     // this item's copies (the next item's may still be in flight)
        asm volatile("bar.sync 1, 128;" ::: "memory");      // ... of every staging thread have landed
        uint32_t (*raw)[kStemRawWords] = s_raw[k & 1];
        const int* shp = s_sh[k & 1];
        // thread t converts entries 4t .. 4t+3 of every row (12 bytes = three words of the stream, funnel-shifted from four words of
        // the copy: lanes read words 3 apart -> no bank conflicts), thread 0 also the last entry; twelve independent chains per row
#pragma unroll
        for (int r = 0; r < 5; ++r) {
          const int iy = it.oy0 * 2 - P.pad_t + r;
          const bool row_ok = iy >= 0 && iy < P.H;
          const int sh8 = row_ok ? shp[r] * 8 : 0;
          const uint32_t* rw = &raw[r][3 * t];
          const uint32_t q0 = rw[0], q1 = rw[1], q2 = rw[2], q3 = rw[3];
          const uint32_t wv[3] = {__funnelshift_r(q0, q1, sh8), __funnelshift_r(q1, q2, sh8), __funnelshift_r(q2, q3, sh8)};
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const int e = 4 * t + kk, ix = 2 * it.x0 + e - P.pad_l;
            const bool ok = row_ok && e < ncols && ix >= 0 && ix < P.W;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const int byte = 3 * kk + c;
              const uint32_t u = (wv[byte >> 2] >> (8 * (byte & 3))) & 255u;
              in[r][c][kk & 1][2 * t + (kk >> 1)] = ok ? s_lut[u] : 0.0f;
            }
          }
          if (t == 0) {
            const int e = 4 * 128, ix = 2 * it.x0 + e - P.pad_l;
            const bool ok = row_ok && e < ncols && ix >= 0 && ix < P.W;
            const uint8_t* rb = reinterpret_cast<const uint8_t*>(&raw[r][0]) + (sh8 >> 3) + 3 * e;
#pragma unroll
            for (int c = 0; c < 3; ++c) in[r][c][0][e >> 1] = ok ? s_lut[rb[c]] : 0.0f;
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");      // this copy buffer is free: the bytes of the item after the next may land
        if (item + 2 * step < num_items) stem_prefetch_u8(P, stem_item(P, item + 2 * step), t, raw, s_sh[k & 1]);
        else asm volatile("cp.async.commit_group;" ::: "memory");
      } else {
#pragma unroll 1
        for (int r = 0; r < 5; ++r) {
          const int iy = it.oy0 * 2 - P.pad_t + r;
          const bool row_ok = iy >= 0 && iy < P.H;
          const float* rowp = static_cast<const float*>(P.img) + (static_cast<size_t>(it.b) * P.H + (row_ok ? iy : 0)) * P.W * 3;
          float v[kStemPer][3];
#pragma unroll
          for (int i = 0; i < kStemPer; ++i) {
            const int e = t + i * 128, ix = 2 * it.x0 + e - P.pad_l;
            const bool ok = row_ok && e < ncols && ix >= 0 && ix < P.W;
#pragma unroll
            for (int c = 0; c < 3; ++c) v[i][c] = ok ? __ldg(rowp + ix * 3 + c) : 0.0f;
          }
#pragma unroll
          for (int i = 0; i < kStemPer; ++i) {
            const int e = t + i * 128;
            if (e < ncols) {
#pragma unroll
              for (int c = 0; c < 3; ++c) in[r][c][e & 1][e >> 1] = v[i][c];
            }
          }
        }
      }
      mbar_arrive(&s_full[k & 1]);
    }
    return;
  }
  // ---------------------------------------------------------------------------------------------- FMA warps
  const int half = warp & 1, orow = warp >> 1;      // this warp: channels [16 half, 16 half + 16) of output row oy0 + orow
  float sc[16], sf[16];
#pragma unroll
  for (int n = 0; n < 16; ++n) {
    sc[n] = s_scale[half * 16 + n];
    sf[n] = s_shift[half * 16 + n];
  }
  uint32_t k = 0;
  for (int item = static_cast<int>(blockIdx.x); item < num_items; item += step, ++k) {
    const StemItem it = stem_item(P, item);
    const int oy = it.oy0 + orow;
    mbar_wait(&s_full[k & 1], (k >> 1) & 1);
    float acc[8][16];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int n = 0; n < 16; ++n) acc[j][n] = 0.0f;
    if (oy < P.Ho) {
#pragma unroll 1
      for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float* in = &s_in[k & 1][orow * 2 + ky][c][kx & 1][lane + (kx >> 1)];
            const float4* wr = reinterpret_cast<const float4*>(s_w + ((ky * 3 + kx) * 3 + c) * 32 + half * 16);
            float w[16];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 w4 = wr[q];
              w[4 * q] = w4.x; w[4 * q + 1] = w4.y; w[4 * q + 2] = w4.z; w[4 * q + 3] = w4.w;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float a = in[32 * j];      // entries past the item's columns belong to pixels >= Wo, which are never stored
#pragma unroll
              for (int n = 0; n < 16; ++n) acc[j][n] = fmaf(a, w[n], acc[j][n]);
            }
          }
        }
      }
    }
    mbar_arrive(&s_empty[k & 1]);      // the buffer may be refilled while the results are stored
    if (oy < P.Ho) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int ox = it.x0 + lane + 32 * j;
        if (ox >= P.Wo) continue;
        __nv_bfloat16* o = P.out + ((static_cast<size_t>(it.b) * P.Ho + oy) * P.Wo + ox) * 32 + half * 16;
#pragma unroll
        for (int q4 = 0; q4 < 8; q4 += 4) {
          uint32_t pk[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int n = 2 * (q4 + q);
            const float x = fmaxf(fmaf(acc[j][n], sc[n], sf[n]), 0.0f);
            const float y = fmaxf(fmaf(acc[j][n + 1], sc[n + 1], sf[n + 1]), 0.0f);
            pk[q] = pack_bf16x2(x, y);
          }
          stg_v4(o + q4 * 2, make_uint4(pk[0], pk[1], pk[2], pk[3]));
        }
      }
    }
  }
}
// persistent: one CTA per SM (register bound), never more CTAs than items
constexpr int kStemDynSmem = 2 * 5 * 3 * 2 * kStemIdx * 4;
inline cudaError_t launch_stem(const StemParams& P, int num_sms, cudaStream_t st, bool pdl) {
  static bool attr_done[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_done[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(stem_conv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStemDynSmem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(stem_conv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStemDynSmem);
    if (e != cudaSuccess) return e;
    attr_done[dev & 63] = true;
  }
  const int items = P.B * ((P.Ho + 1) / 2) * ((P.Wo + kStemSeg - 1) / kStemSeg);
  const int grid = items < num_sms ? items : num_sms;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(grid));
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = kStemDynSmem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return P.img_f32 ? cudaLaunchKernelEx(&cfg, stem_conv_kernel<false>, P) : cudaLaunchKernelEx(&cfg, stem_conv_kernel<true>, P);
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
