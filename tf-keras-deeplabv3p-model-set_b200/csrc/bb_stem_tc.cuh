// bb_stem_tc.cuh — entry_flow_conv1_1 on the tensor cores for uint8 images:
//   normalize_image (x / 127.5 - 1, common/data_utils.py:403-416) -> Conv2D(32, 3x3, strides 2, 'same', no bias) -> BN -> ReLU
//   (reference deeplabv3p/models/deeplabv3p_xception.py:119-123).
//
// A stride-2 3x3 convolution is a stride-1 2x2 convolution over the 2x2 space-to-depth image: output (oy, ox) reads the four s2d pixels
// (oy + ty, ox + tx), ty, tx in {0, 1}, where s2d pixel (Y, X) holds the image pixels (2Y - pad_t + dy, 2X - pad_l + dx), dy, dx in {0, 1}
// (TensorFlow 'same': pad_t = pad_l = 0 on even sizes, 1 on odd ones) and tap (ky, kx) = (2 ty + dy, 2 tx + dx); taps with ky or kx = 3 get
// zero weights.  Per s2d pixel 12 values (+ 4 zeros) = one K = 16 step.  The arithmetic keeps the reference's accuracy without fp32 FMAs:
//   * the RAW pixel values 0..255 are exact in bf16, so the MMA computes S = sum(u * w) and the epilogue normalises: conv = S / 127.5 - sum(w);
//   * the fp32 weights enter as THREE bf16 pieces (w = w0 + w1 + w2 exactly: 3 x 8 mantissa bits), every product is exact in fp32;
//   * zero padding pads the NORMALISED image, i.e. a padded tap contributes 0, not -w: sum(w) runs over the taps inside the image
//     (a precomputed total for interior pixels, the 9 per-tap sums for the border).
// What differs from the fp32 kernel is the order of the fp32 accumulation (tensor core vs tap order): ~1e-6 relative before the bf16 rounding.
//
//   warp 0        MMA issuer: per tile 3 pieces x 4 taps = 12 tcgen05.mma 128 x 32 x 16; the four taps are descriptors shifted inside ONE
//                 17 x 9 s2d halo tile (32-byte rows under the 32-byte swizzle, atoms one halo row = 288 bytes apart)
//   warps 1..4    epilogue: tcgen05.ld -> normalise -> BN -> ReLU -> bf16 -> 64 bytes per pixel
//   warps 5..12   builders, one pipeline stage each: image bytes -> bf16 s2d halo tile in the swizzled operand layout
#pragma once

#include <cuda.h>

#include "sm100_prims.cuh"

namespace dlv3p {

constexpr int kStcSub = 1;                                   // M = 128 sub-tiles (16 rows x 8 pixels = 16 eight-pixel atoms) side by side in one work item:
                                                             // twice the work per barrier hand-over, one shared halo column
constexpr int kStcTH = 16, kStcTW = 8 * kStcSub;
constexpr int kStcHaloH = kStcTH + 1, kStcHaloW = kStcTW + 1;
constexpr int kStcPix = kStcHaloH * kStcHaloW;               // 153 s2d pixels per tile
constexpr int kStcStageBytes = (kStcPix * 32 + 255) / 256 * 256;      // one 32-byte row (K = 16 bf16) per s2d pixel, 32-byte swizzle (256-byte pattern)
constexpr int kStcStages = 8;
constexpr int kStcBuilders = kStcStages;
constexpr int kStcThreads = (1 + 4 + kStcBuilders) * 32;
constexpr int kStcWBytes = 3 * 4 * 1024;                     // [piece][tap][32 out][16 k] bf16, 32-byte swizzle applied by the host
constexpr int kStcRawRows = 2 * kStcHaloH;                   // image rows under a tile's s2d halo
constexpr int kStcRawChunks = (2 * kStcHaloW * 3 + 15 + 15) / 16;     // 16-byte chunks that cover a row's 2 (TW + 1) pixels at any alignment
constexpr int kStcRawRowBytes = kStcRawChunks * 16;
constexpr int kStcRawBytes = (kStcRawRows * kStcRawRowBytes + 127) / 128 * 128;
constexpr int kStcSmemBytes = 1024 + kStcWBytes + kStcStages * kStcStageBytes + kStcBuilders * kStcRawBytes + 256;

struct StemTcParams {
  const uint8_t* img;          // [B, H, W, 3] uint8
  const uint16_t* w16;         // the 12 B tiles above (host-packed)
  const float* wk;             // [9][32] per-tap channel sums of the fp32 weights, then [32] their total
  const float* scale;          // [32] folded BN
  const float* shift;          // [32]
  __nv_bfloat16* out;          // [B, Ho, Wo, 32]
  int B, H, W, Ho, Wo, pad_t, pad_l;
  int tiles_x, tiles_y, num_tiles;
};

// K-major operand of 32-byte rows (one K = 16 step) under the 32-byte swizzle: the two 16-byte halves of a row swap where address bit 7 is set;
// eight-row atoms `atom_stride` bytes apart.  Like the 64-byte swizzle of bb_conv3x3.cuh the pattern is a function of the absolute shared-memory
// address, so a descriptor may start at any row of a tile that was WRITTEN with the same rule.  (Without a swizzle — core matrices of 8 x 16
// bytes, atoms 144 bytes apart — the kernel was correct but its operand fetches conflicted in shared memory: 58 % of all wavefronts.)
__device__ __forceinline__ uint64_t make_smem_desc_sw32(uint32_t smem_addr_bytes, uint32_t atom_stride) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr_bytes >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>(1) << 16;                             // leading-dimension byte offset: unused (one swizzle row per K step)
  d |= static_cast<uint64_t>((atom_stride >> 4) & 0x3FFF) << 32;   // stride-dimension byte offset
  d |= static_cast<uint64_t>(1) << 46;                             // descriptor version
  d |= static_cast<uint64_t>(6) << 61;                             // SWIZZLE_32B
  return d;
}

__global__ void __launch_bounds__(kStcThreads, 1) stem_tc_kernel(const __grid_constant__ StemTcParams P) {
  extern __shared__ __align__(1024) uint8_t smem_stc[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_stc) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* smem_w = smem;
  uint8_t* smem_a = smem_w + kStcWBytes;
  // small tables in static shared memory (the compiler keeps the shared address space: LDS, not generic loads)
  __shared__ __align__(16) float s_wk[9 * 32];      // per-tap channel sums of the weights
  __shared__ __align__(16) float s_a[32];           // scale / 127.5
  __shared__ __align__(16) float s_b[32];           // shift - scale * (sum of all nine taps): interior pixels
  __shared__ __align__(16) float s_scale[32];
  uint8_t* smem_raw = smem_a + kStcStages * kStcStageBytes;                        // one raw-row buffer per builder warp
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + kStcBuilders * kStcRawBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStcStages;
  uint64_t* tmem_full = empty_bar + kStcStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_per_img = P.tiles_x * P.tiles_y;
  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStcStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 128);      // the four epilogue warps
    }
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < kStcWBytes / 16; i += kStcThreads) sts_v4(smem_u32(smem_w) + i * 16, __ldg(reinterpret_cast<const uint4*>(P.w16) + i));
  for (int i = threadIdx.x; i < 9 * 32; i += kStcThreads) s_wk[i] = __ldg(P.wk + i);
  if (threadIdx.x < 32) {
    const float sc = __ldg(P.scale + threadIdx.x);
    s_scale[threadIdx.x] = sc;
    s_a[threadIdx.x] = sc * (1.0f / 127.5f);
    s_b[threadIdx.x] = fmaf(-sc, __ldg(P.wk + 9 * 32 + threadIdx.x), __ldg(P.shift + threadIdx.x));
  }
  fence_proxy_async_smem();          // the weights were written through the generic proxy and are read by the tensor core
  if (warp == 0) {
    tmem_alloc(tmem_base_ptr, 2 * 32 * kStcSub);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_ptr;
  pdl_wait();          // the images may come from the previous kernel of the stream; the output buffer may still be read by it

  if (warp == 0) {
    // ------------------------------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = make_idesc_bf16(128, 32);
    uint32_t k = 0;
    for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x, ++k) {
      const uint32_t stage = k % kStcStages, phase = (k / kStcStages) & 1;
      const uint32_t acc = k & 1, acc_phase = (k >> 1) & 1;
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      mbar_wait(&full_bar[stage], phase);
      tcgen05_fence_after();
      if (elect_one()) {
        const uint32_t a0 = smem_u32(smem_a + stage * kStcStageBytes);
#pragma unroll
        for (int sub = 0; sub < kStcSub; ++sub) {
          const uint32_t tmem_d = tmem_base + acc * (32 * kStcSub) + sub * 32;
#pragma unroll
          for (int piece = 0; piece < 3; ++piece)
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const uint64_t da = make_smem_desc_sw32(a0 + ((t >> 1) * kStcHaloW + 8 * sub + (t & 1)) * 32, kStcHaloW * 32);
              const uint64_t db = make_smem_desc_sw32(smem_u32(smem_w) + (piece * 4 + t) * 1024, 256);
              umma_bf16_ss(tmem_d, da, db, idesc, (piece > 0 || t > 0) ? 1u : 0u);
            }
        }
        umma_commit(&empty_bar[stage]);
        umma_commit(&tmem_full[acc]);
      }
      __syncwarp();
    }
  } else if (warp < 5) {
    // ------------------------------------------------------------------------------------------ epilogue
    const int q = warp & 3;                              // TMEM lanes 32 q .. 32 q + 31 = tile rows 4 q .. 4 q + 3
    const int r = 4 * q + (lane >> 3), col = lane & 7;
    uint32_t k = 0;
    for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x, ++k) {
      const int b = tile / tiles_per_img, t2 = tile - b * tiles_per_img;
      const int ty = t2 / P.tiles_x, tx = t2 - ty * P.tiles_x;
      const int oy = ty * kStcTH + r;
      const uint32_t acc = k & 1, acc_phase = (k >> 1) & 1;
      mbar_wait(&tmem_full[acc], acc_phase);
      tcgen05_fence_after();
      uint32_t vv[kStcSub][32];
#pragma unroll
      for (int sub = 0; sub < kStcSub; ++sub) tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * (32 * kStcSub) + sub * 32, vv[sub]);
      tmem_ld_wait();
      tcgen05_fence_before();
      mbar_arrive(&tmem_empty[acc]);
#pragma unroll
      for (int sub = 0; sub < kStcSub; ++sub) {
      const uint32_t (&v)[32] = vv[sub];
      const int ox = tx * kStcTW + 8 * sub + col;
      if (oy >= P.Ho || ox >= P.Wo) continue;
      // y = relu(S * (scale / 127.5) + shift - scale * sum of the taps inside the image); a border pixel gets the sums of its taps OUTSIDE the
      // image back (zero padding pads the normalised image: those taps contribute nothing)
      const int iy0 = 2 * oy - P.pad_t, ix0 = 2 * ox - P.pad_l;
      const bool all_in = iy0 >= 0 && iy0 + 2 < P.H && ix0 >= 0 && ix0 + 2 < P.W;
      float corr[32];
#pragma unroll
      for (int n = 0; n < 32; ++n) corr[n] = 0.0f;
      if (!all_in) {
#pragma unroll 1
        for (int t = 0; t < 9; ++t) {
          const int iy = iy0 + t / 3, ix = ix0 + t % 3;
          if (iy >= 0 && iy < P.H && ix >= 0 && ix < P.W) continue;
#pragma unroll
          for (int n4 = 0; n4 < 32; n4 += 4) {
            const float4 w4 = *reinterpret_cast<const float4*>(s_wk + t * 32 + n4);
            corr[n4] += w4.x; corr[n4 + 1] += w4.y; corr[n4 + 2] += w4.z; corr[n4 + 3] += w4.w;
          }
        }
      }
      __nv_bfloat16* o = P.out + ((static_cast<size_t>(b) * P.Ho + oy) * P.Wo + ox) * 32;
#pragma unroll
      for (int n0 = 0; n0 < 32; n0 += 8) {
        float a[8], bb[8], sc[8];
        *reinterpret_cast<float4*>(a) = *reinterpret_cast<const float4*>(s_a + n0);
        *reinterpret_cast<float4*>(a + 4) = *reinterpret_cast<const float4*>(s_a + n0 + 4);
        *reinterpret_cast<float4*>(bb) = *reinterpret_cast<const float4*>(s_b + n0);
        *reinterpret_cast<float4*>(bb + 4) = *reinterpret_cast<const float4*>(s_b + n0 + 4);
        if (!all_in) {
          *reinterpret_cast<float4*>(sc) = *reinterpret_cast<const float4*>(s_scale + n0);
          *reinterpret_cast<float4*>(sc + 4) = *reinterpret_cast<const float4*>(s_scale + n0 + 4);
#pragma unroll
          for (int e = 0; e < 8; ++e) bb[e] = fmaf(sc[e], corr[n0 + e], bb[e]);
        }
        uint32_t pk[4];
#pragma unroll
        for (int e = 0; e < 4; ++e)
          pk[e] = pack_bf16x2(fmaxf(fmaf(__uint_as_float(v[n0 + 2 * e]), a[2 * e], bb[2 * e]), 0.0f),
                              fmaxf(fmaf(__uint_as_float(v[n0 + 2 * e + 1]), a[2 * e + 1], bb[2 * e + 1]), 0.0f));
        stg_v4(o + n0, make_uint4(pk[0], pk[1], pk[2], pk[3]));
      }
      }
    }
  } else {
    // ------------------------------------------------------------------------------------------ builders: one stage each
    // The 34 image rows of a tile (18 pixels = 54 bytes each) come in as 16-byte asynchronous copies from the 16-byte boundary below the row's
    // first byte (five per row cover 54 bytes at any alignment) into the builder's own raw buffer; the lanes then pick their s2d pixels' bytes
    // out of shared memory.  (One byte per load instruction made every instruction touch nine sectors: the builders were the bottleneck.)
    const int bi = warp - 5;
    uint8_t* rawb = smem_raw + bi * kStcRawBytes;
    const uint32_t raw_u32 = smem_u32(rawb);
    const long long total = static_cast<long long>(P.B) * P.H * P.W * 3;
    const long long img_addr = static_cast<long long>(reinterpret_cast<uintptr_t>(P.img));
    uint32_t k = 0;
    for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x, ++k) {
      if (static_cast<int>(k % kStcStages) != bi) continue;
      const uint32_t phase = (k / kStcStages) & 1;
      const int b = tile / tiles_per_img, t2 = tile - b * tiles_per_img;
      const int ty = t2 / P.tiles_x, tx = t2 - ty * P.tiles_x;
      const int iy_first = 2 * ty * kStcTH - P.pad_t, ix_first = 2 * tx * kStcTW - P.pad_l;
      const long long img_off = static_cast<long long>(b) * P.H * P.W * 3;
      __syncwarp();                                      // the previous tile's reads of the raw buffer are done
#pragma unroll 1
      for (int j = lane; j < kStcRawRows * kStcRawChunks; j += 32) {
        const int rr = j / kStcRawChunks, ch = j - rr * kStcRawChunks;
        const int iy = iy_first + rr;
        if (iy < 0 || iy >= P.H) continue;               // rows outside the image are masked when they are read
        const long long s0 = img_off + (static_cast<long long>(iy) * P.W + ix_first) * 3;      // first byte of the row's 18 pixels (may lie 3 bytes before the row)
        const long long a0 = ((img_addr + s0) & ~15ll) - img_addr + 16 * ch;      // 16-byte boundaries of the ADDRESS, whatever the buffer's own alignment
        const uint32_t dst = raw_u32 + rr * kStcRawRowBytes + ch * 16;
        if (a0 >= 0 && a0 + 16 <= total) {
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(P.img + a0) : "memory");
        } else {                                         // a chunk that straddles the ends of the image buffer: byte by byte
          uint32_t wv[4] = {0u, 0u, 0u, 0u};
          for (int e = 0; e < 16; ++e)
            if (a0 + e >= 0 && a0 + e < total) wv[e >> 2] |= static_cast<uint32_t>(__ldg(P.img + a0 + e)) << (8 * (e & 3));
          sts_v4(dst, make_uint4(wv[0], wv[1], wv[2], wv[3]));
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncwarp();
      mbar_wait(&empty_bar[bi], phase ^ 1);              // the MMAs of the tile that used this stage last have retired
      const uint32_t st = smem_u32(smem_a + bi * kStcStageBytes);
#pragma unroll 2
      for (int i = 0; i < (kStcPix + 31) / 32; ++i) {
        const int p = lane + 32 * i;
        if (p < kStcPix) {
          const int yy = p / kStcHaloW, xx = p - yy * kStcHaloW;
          uint32_t val[12];
#pragma unroll
          for (int dy = 0; dy < 2; ++dy) {
            const int rr = 2 * yy + dy, iy = iy_first + rr;
            const bool row_ok = iy >= 0 && iy < P.H;
            const long long s0 = img_off + (static_cast<long long>(iy) * P.W + ix_first) * 3;
            const uint8_t* rowp = rawb + rr * kStcRawRowBytes + static_cast<int>((img_addr + s0) & 15) + 6 * xx;
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
              const int ix = ix_first + 2 * xx + dx;
              const bool ok = row_ok && ix >= 0 && ix < P.W;
#pragma unroll
              for (int c = 0; c < 3; ++c) val[(dy * 2 + dx) * 3 + c] = ok ? static_cast<uint32_t>(rowp[dx * 3 + c]) : 0u;
            }
          }
          uint32_t wds[6];
#pragma unroll
          for (int j = 0; j < 6; ++j)                      // bf16 bits of an integer 0..255: the upper half of its fp32 pattern, exact
            wds[j] = (__float_as_uint(static_cast<float>(val[2 * j])) >> 16) | (__float_as_uint(static_cast<float>(val[2 * j + 1])) & 0xFFFF0000u);
          const uint32_t row = st + p * 32, flip = (row >> 3) & 16u;      // 32-byte swizzle: halves swap where address bit 7 is set
          sts_v4(row + flip, make_uint4(wds[0], wds[1], wds[2], wds[3]));
          sts_v4(row + (flip ^ 16u), make_uint4(wds[4], wds[5], 0u, 0u));
        }
      }
      fence_proxy_async_smem();                          // generic-proxy stores -> visible to the tensor core's reads
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[bi]);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 2 * 32 * kStcSub);
  }
}

}  // namespace dlv3p
