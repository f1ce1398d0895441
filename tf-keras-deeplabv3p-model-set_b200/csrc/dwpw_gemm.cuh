// dwpw_gemm.cuh — tile geometry, parameters and packed-fp32 helpers of the fused SepConv_BN kernel (dwpw_gemm2.cuh):
//   depthwise 3x3 'same' -> BN -> ReLU -> pointwise 1x1 (C -> 256) -> BN -> ReLU
// (reference deeplabv3p/models/layers.py:74-111, used by Decoder_block :215-218).
//
// The depthwise result never touches HBM: CUDA-core "stencil" warps compute it from a TMA-staged halo
// tile and write it, already BN'd / ReLU'd / rounded to bf16, straight into the 128B-swizzled shared
// memory layout the tcgen05 MMA reads as its A operand.  The pointwise weights (256 x C bf16) stay
// resident in shared memory for the life of the (persistent) CTA.
//
//   tile            8 x 16 output pixels of one image  (= the 128 rows of one UMMA M tile); halo tiles [10][18][64ch] per 64-channel K block
#pragma once

#include <cuda.h>

#include "sm100_prims.cuh"

namespace dlv3p {

constexpr int kDwTH = 8;
constexpr int kDwTW = 16;
constexpr int kDwHaloH = kDwTH + 2;
constexpr int kDwHaloW = kDwTW + 2;
constexpr int kDwInStageBytes = kDwHaloH * kDwHaloW * 128;  // 64 bf16 channels per pixel
constexpr int kDwAStageBytes = 128 * 128;
constexpr int kDwWBlockBytes = 256 * 128;
constexpr int kDwBN = 256;

struct DwPwParams {
  const CUtensorMap* tmap_x;  // 4D {C, W, H, B} bf16, box {64, 18, 10, 1}, no swizzle
  const CUtensorMap* tmap_x2; // optional second input tensor (same geometry): K blocks kb >= kb_split come from it at channel
                              // (kb - kb_split) * 64 — the decoder concat [upsampled ASPP | projected skip] without a concat buffer
  int kb_split;
  const CUtensorMap* tmap_w;  // 2D [256, KB*64] bf16 K-major, box {64, 128}, SWIZZLE_128B
  const float* dw_w;          // [9][KB*64] fp32 depthwise taps with the BN scale folded in, zero padded
  const float* dw_shift;      // [KB*64]    fp32 depthwise BN shift, zero padded
  const float* scale;         // [256] pointwise BN scale
  const float* shift;         // [256] pointwise BN shift
  __nv_bfloat16* out;         // [B, H, W, 256]
  const CUtensorMap* tmap_out;  // 4D {256, W, H, B} bf16, box {64, 16, 2, 1}, SWIZZLE_128B (TMA-store epilogue, KB <= 4)
  int B, H, W;
  int tiles_x, tiles_y, num_tiles;
  int debug;  // benchmark aid: bit0 skip epilogue stores, bit1 skip the stencil math (A tiles left stale), bit2 skip the MMAs
};

__device__ __forceinline__ unsigned long long pack_f32x2(float lo, float hi) {
  return (static_cast<unsigned long long>(__float_as_uint(hi)) << 32) | __float_as_uint(lo);
}
__device__ __forceinline__ float f32x2_lo(unsigned long long v) { return __uint_as_float(static_cast<uint32_t>(v)); }
__device__ __forceinline__ float f32x2_hi(unsigned long long v) { return __uint_as_float(static_cast<uint32_t>(v >> 32)); }
// bf16x2 (two adjacent channels) -> packed fp32x2
__device__ __forceinline__ unsigned long long bf16x2_to_f32x2(uint32_t v) {
  return (static_cast<unsigned long long>(v & 0xFFFF0000u) << 32) | (v << 16);
}
// d = a * b + d on both halves (Blackwell packed fp32 FMA)
__device__ __forceinline__ void ffma2(unsigned long long& d, unsigned long long a, unsigned long long b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}

}  // namespace dlv3p
