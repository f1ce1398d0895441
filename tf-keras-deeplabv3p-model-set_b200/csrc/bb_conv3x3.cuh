// bb_conv3x3.cuh — the one dense 3x3 convolution of the Xception entry flow as an implicit GEMM on tcgen05:
//   entry_flow_conv1_2 = Conv2D(64, 3x3, stride 1, 'same', no bias) on 32 channels -> BN -> ReLU
//   (reference _conv2d_same deeplabv3p/models/deeplabv3p_xception.py:25-41, call site :125-127).
//
// im2col never exists in memory: for an 8 x 16-pixel output tile the A operand of tap (ky, kx) is the 8 x 16 x 32-channel box of
// the NHWC input shifted by (ky-1, kx-1), fetched by ONE TMA load whose out-of-bounds pixels arrive as zeros ('same' padding).
// Each tap is a K = 32 slice (two K = 16 MMAs) of the K = 288 contraction; rows are 64 bytes, so the operand layout is the
// 64-byte swizzle (8-row atoms of 512 bytes) for TMA, for the UMMA descriptors and for the resident weights alike.
//
//   warp 0      TMA producer: weights once (9 taps x [64 out][32 in] = 36 KB), then 9 boxes per tile through a 16-stage ring
//   warp 1      MMA issuer  : tcgen05.mma 128 x 64 x 16, fp32 accumulators in TMEM (2 stages x 64 columns)
//   warps 2..5  epilogue    : tcgen05.ld -> BN scale/shift -> ReLU -> bf16 -> 128B-swizzled smem -> 4D TMA store (2 tile rows per warp)
#pragma once

#include <cuda.h>

#include "sm100_prims.cuh"

namespace dlv3p {

constexpr int kC3Threads = 192;
constexpr int kC3TH = 16, kC3TW = 8;                // output tile: 16 rows of 8 pixels = the 16 eight-row atoms of the M = 128 operand
constexpr int kC3HaloW = kC3TW + 2, kC3HaloH = kC3TH + 2;
constexpr int kC3Stages = 8;
constexpr int kC3HaloBytes = kC3HaloH * kC3HaloW * 64;             // 18 x 10 pixels x 32 channels bf16 = 11520
constexpr int kC3AStageBytes = (kC3HaloBytes + 511) / 512 * 512;   // stages start on the 64-byte swizzle's 512-byte pattern
constexpr int kC3WTapBytes = 64 * 64;               // 64 output channels x 32 input channels bf16
constexpr int kC3WBytes = 9 * kC3WTapBytes;
constexpr int kC3StoreBytes = 4 * 2 * 4096;         // 4 epilogue warps x 2 buffers x [32 pixels x 128 B]
constexpr int kC3SmemBytes = 1024 + kC3WBytes + kC3Stages * kC3AStageBytes + kC3StoreBytes + 256;

struct Conv3x3Params {
  const CUtensorMap* tmap_x;    // 4D {32, W, H, B} bf16, box {32, 10, 18, 1} (tile + halo), SWIZZLE_64B, OOB -> 0
  const CUtensorMap* tmap_w;    // 2D [9 * 64 rows (tap, cout)][32 cin] bf16, box {32, 64}, SWIZZLE_64B
  const CUtensorMap* tmap_out;  // 4D {64, W, H, B} bf16, box {64, 8, 4, 1}, SWIZZLE_128B
  float scale[64];              // folded BN, constant bank
  float shift[64];
  int tiles_x, tiles_y, num_tiles;
};

// K-major operand tile of 64-byte rows under the 64-byte swizzle: 8-row atoms `atom_stride` bytes apart (512: a dense tile)
__device__ __forceinline__ uint64_t make_smem_desc_sw64(uint32_t smem_addr_bytes, uint32_t atom_stride = 512) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr_bytes >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(atom_stride >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(4) << 61;                   // SWIZZLE_64B
  return d;
}

__global__ void __launch_bounds__(kC3Threads, 1) conv3x3_c32_kernel(const __grid_constant__ Conv3x3Params P) {
  extern __shared__ __align__(1024) uint8_t smem_c3[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_c3) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* smem_w = smem;                                   // 9 x [64 rows x 64 B]
  uint8_t* smem_a = smem_w + kC3WBytes;                     // kC3Stages x [128 rows x 64 B]   (36 KB offset: 512-byte aligned)
  uint8_t* smem_o = smem_a + kC3Stages * kC3AStageBytes;    // store staging (1024-byte aligned: 36 KB + 128 KB)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_o + kC3StoreBytes);
  uint64_t* w_full = bars;
  uint64_t* full_bar = bars + 1;
  uint64_t* empty_bar = full_bar + kC3Stages;
  uint64_t* tmem_full = empty_bar + kC3Stages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_per_img = P.tiles_x * P.tiles_y;

  if (warp == 0 && lane == 0) {
    mbar_init(w_full, 1);
    for (int i = 0; i < kC3Stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_base_ptr, 128);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_ptr;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(w_full, kC3WBytes);
      for (int t = 0; t < 9; ++t) tma_load_2d(smem_w + t * kC3WTapBytes, P.tmap_w, w_full, 0, t * 64, kEvictLast);
      uint32_t stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x) {
        const int b = tile / tiles_per_img;
        const int t2 = tile - b * tiles_per_img;
        const int ty = t2 / P.tiles_x, tx = t2 - ty * P.tiles_x;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full_bar[stage], kC3HaloBytes);
        tma_load_4d(smem_a + stage * kC3AStageBytes, P.tmap_x, &full_bar[stage], 0, tx * kC3TW - 1, ty * kC3TH - 1, b, kEvictNormal);
        if (++stage == kC3Stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(128, 64);
    mbar_wait(w_full, 0);
    uint32_t stage = 0, phase = 0, it = 0;
    for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x, ++it) {
      const uint32_t acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tcgen05_fence_after();
      const uint32_t tmem_d = tmem_base + acc * 64;
      mbar_wait(&full_bar[stage], phase);
      tcgen05_fence_after();
      if (elect_one()) {
        // tap (ky, kx) reads the halo tile shifted by ky rows and kx pixels: atom y of the operand = the 8 pixels of halo row y + ky from
        // pixel kx on, 512 contiguous bytes, atoms one halo row (640 bytes) apart.  TMA and UMMA apply the 64-byte swizzle to the same
        // absolute shared-memory address bits, so a start address that is not a multiple of the pattern needs nothing else.
        const uint32_t a0 = smem_u32(smem_a + stage * kC3AStageBytes);
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const uint64_t da = make_smem_desc_sw64(a0 + ((t / 3) * kC3HaloW + (t % 3)) * 64, kC3HaloW * 64);
          const uint64_t db = make_smem_desc_sw64(smem_u32(smem_w + t * kC3WTapBytes));
#pragma unroll
          for (int k = 0; k < 2; ++k)
            umma_bf16_ss(tmem_d, smem_desc_advance(da, k * 32), smem_desc_advance(db, k * 32), idesc, (t > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage]);
        umma_commit(&tmem_full[acc]);
      }
      __syncwarp();
      if (++stage == kC3Stages) { stage = 0; phase ^= 1; }
    }
  } else {
    const int q = warp & 3;
    uint32_t it = 0, store_buf = 0;
    uint8_t* my_o = smem_o + (warp - 2) * 2 * 4096;
    for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x, ++it) {
      const int b = tile / tiles_per_img;
      const int t2 = tile - b * tiles_per_img;
      const int ty = t2 / P.tiles_x, tx = t2 - ty * P.tiles_x;
      const uint32_t acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tmem_full[acc], acc_phase);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * 64;
      if (lane == 0) tma_store_wait_read<1>();
      __syncwarp();
      const uint32_t obuf = smem_u32(my_o + store_buf * 4096) + lane * 128;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int c0 = half * 32;
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr + c0, v);
        tmem_ld_wait();
        if (half == 1) {     // both halves read: the accumulator stage can be refilled
          tcgen05_fence_before();
          mbar_arrive(&tmem_empty[acc]);
        }
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          uint32_t pk[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float a = fmaxf(fmaf(__uint_as_float(v[j + 2 * e]), P.scale[c0 + j + 2 * e], P.shift[c0 + j + 2 * e]), 0.0f);
            const float c = fmaxf(fmaf(__uint_as_float(v[j + 2 * e + 1]), P.scale[c0 + j + 2 * e + 1], P.shift[c0 + j + 2 * e + 1]), 0.0f);
            pk[e] = pack_bf16x2(a, c);
          }
          const uint32_t chunk = static_cast<uint32_t>(half * 4 + (j >> 3)) ^ static_cast<uint32_t>(lane & 7);
          sts_v4(obuf + chunk * 16, make_uint4(pk[0], pk[1], pk[2], pk[3]));
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                         reinterpret_cast<uint64_t>(P.tmap_out)),
                     "r"(smem_u32(my_o + store_buf * 4096)), "r"(0), "r"(tx * kC3TW), "r"(ty * kC3TH + 4 * q), "r"(b)
                     : "memory");
        tma_store_commit();
      }
      store_buf ^= 1;
    }
    if (lane == 0) tma_store_wait_all<0>();
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

}  // namespace dlv3p
