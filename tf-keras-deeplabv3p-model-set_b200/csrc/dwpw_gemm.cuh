// dwpw_gemm.cuh — fused SepConv_BN (depth_activation=True, rate 1):
//   depthwise 3x3 'same' -> BN -> ReLU -> pointwise 1x1 (C -> 256) -> BN -> ReLU
// (reference deeplabv3p/models/layers.py:74-111, used by Decoder_block :215-218).
//
// The depthwise result never touches HBM: CUDA-core "stencil" warps compute it from a TMA-staged halo
// tile and write it, already BN'd / ReLU'd / rounded to bf16, straight into the 128B-swizzled shared
// memory layout the tcgen05 MMA reads as its A operand.  The pointwise weights (256 x C bf16) stay
// resident in shared memory for the life of the (persistent) CTA.
//
//   tile            8 x 16 output pixels of one image  (= the 128 rows of one UMMA M tile)
//   warp 0          TMA producer: halo tiles [10][18][64ch] per 64-channel K block (OOB zero fill = 'same' padding)
//   warp 1          MMA issuer  : tcgen05.mma 128 x 256 x 16, fp32 accumulators in TMEM (2 stages)
//   warps 2..5      epilogue    : tcgen05.ld -> BN scale/shift -> ReLU -> bf16 -> NHWC global
//   warps 6..13     stencil     : 8 warps on the same K block (two row halves); packed fp32x2 FMAs
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
constexpr int kDwThreads = 14 * 32;
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

template <int KB, int AS, int kDwInStages>
struct DwPwCfg {
  static constexpr bool kTmaStore = KB <= 4;                    // room for the epilogue's store staging (4 warps x 4 KB)
  static constexpr int kStoreBytes = kTmaStore ? 4 * 4096 : 0;   // (+1 KB alignment slack, added below)
  static constexpr int kWBytes = KB * kDwWBlockBytes;
  static constexpr int kABytes = AS * kDwAStageBytes;
  static constexpr int kInBytes = kDwInStages * kDwInStageBytes;
  static constexpr int kSmemBytes = kWBytes + kABytes + kInBytes + kStoreBytes + (kTmaStore ? 1024 : 0) + 2048 /*BN scale+shift*/ + 256 /*barriers*/;
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

template <int KB, int AS, int kDwInStages>
__global__ void __launch_bounds__(kDwThreads, 1) dwpw_gemm_kernel(const __grid_constant__ DwPwParams P) {
  using Cfg = DwPwCfg<KB, AS, kDwInStages>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;   // dynamic smem is the only shared memory of this kernel: 1024-byte aligned (checked below)
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  uint8_t* smem_w = smem;                      // KB x [256 rows x 128 B], swizzled
  uint8_t* smem_a = smem_w + Cfg::kWBytes;     // AS x [128 rows x 128 B], swizzled
  uint8_t* smem_in = smem_a + Cfg::kABytes;    // kDwInStages x [10][18][64] bf16
  uint8_t* smem_c = smem_in + Cfg::kInBytes;   // epilogue store staging (1024-byte aligned: all sizes above are multiples of 1024... the halo stages are 22.5 KB, so align up)
  float* s_scale = reinterpret_cast<float*>(smem_c + Cfg::kStoreBytes + (Cfg::kTmaStore ? 1024 : 0));   // [256] pointwise BN scale (L1 is ~0 KB here)
  float* s_shift = s_scale + 256;                                        // [256]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_shift + 256);
  uint64_t* w_full = bars;                     // [1]
  uint64_t* in_full = bars + 1;                // [kDwInStages]
  uint64_t* in_empty = in_full + kDwInStages;  // [kDwInStages]
  uint64_t* a_full = in_empty + kDwInStages;   // [AS]
  uint64_t* a_empty = a_full + AS;             // [AS]
  uint64_t* tmem_full = a_empty + AS;          // [2]
  uint64_t* tmem_empty = tmem_full + 2;        // [2]
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    mbar_init(w_full, 1);
    for (int i = 0; i < kDwInStages; ++i) {
      mbar_init(&in_full[i], 1);
      mbar_init(&in_empty[i], 256);
    }
    for (int i = 0; i < AS; ++i) {
      mbar_init(&a_full[i], 256);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_base_ptr, 512);
    tmem_relinquish();
  }
  if (threadIdx.x >= 64 && threadIdx.x < 64 + 256) {
    s_scale[threadIdx.x - 64] = P.scale[threadIdx.x - 64];
    s_shift[threadIdx.x - 64] = P.shift[threadIdx.x - 64];
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_ptr;
  const int tiles_per_img = P.tiles_x * P.tiles_y;

  const bool free_run = P.debug & 16;   // benchmark aid: stencil warps alone, no synchronisation at all
  if (free_run && warp < 6) {
    // nothing
  } else if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      mbar_arrive_expect_tx(w_full, Cfg::kWBytes);
      for (int kb = 0; kb < KB; ++kb)
{
        tma_load_2d(smem_w + kb * kDwWBlockBytes, P.tmap_w, w_full, kb * 64, 0, kEvictLast);
        tma_load_2d(smem_w + kb * kDwWBlockBytes + 128 * 128, P.tmap_w, w_full, kb * 64, 128, kEvictLast);
      }
      uint32_t c = 0;
      for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x) {
        const int b = tile / tiles_per_img;
        const int t2 = tile - b * tiles_per_img;
        const int ty = t2 / P.tiles_x;
        const int tx = t2 - ty * P.tiles_x;
        for (int kb = 0; kb < KB; ++kb, ++c) {
          const uint32_t si = c % kDwInStages;
          const uint32_t ph = (c / kDwInStages) & 1;
          mbar_wait(&in_empty[si], ph ^ 1);
          mbar_arrive_expect_tx(&in_full[si], kDwInStageBytes);
          const bool second = P.tmap_x2 != nullptr && kb >= P.kb_split;
          tma_load_4d(smem_in + si * kDwInStageBytes, second ? P.tmap_x2 : P.tmap_x, &in_full[si], (second ? kb - P.kb_split : kb) * 64,
                      tx * kDwTW - 1, ty * kDwTH - 1, b, kEvictNormal);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = make_idesc_bf16(128, kDwBN);
    mbar_wait(w_full, 0);
    uint32_t c = 0, it = 0;
    for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x, ++it) {
      const uint32_t acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tcgen05_fence_after();
      const uint32_t tmem_d = tmem_base + acc * kDwBN;
      for (int kb = 0; kb < KB; ++kb, ++c) {
        const uint32_t sa = c % AS;
        const uint32_t ph = (c / AS) & 1;
        mbar_wait(&a_full[sa], ph);
        tcgen05_fence_after();
        if (elect_one()) {
          if (!(P.debug & 4)) {
          const uint64_t da = make_smem_desc_sw128(smem_u32(smem_a + sa * kDwAStageBytes));
          const uint64_t db = make_smem_desc_sw128(smem_u32(smem_w + kb * kDwWBlockBytes));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_ss(tmem_d, smem_desc_advance(da, k * 32), smem_desc_advance(db, k * 32), idesc,
                         (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&a_empty[sa]);
          if (kb == KB - 1) umma_commit(&tmem_full[acc]);
        }
        __syncwarp();
      }
    }
  } else if (warp < 6) {
    // ------------------------------------------------------------------ epilogue
    const int q = warp & 3;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x, ++it) {
      const int b = tile / tiles_per_img;
      const int t2 = tile - b * tiles_per_img;
      const int ty = t2 / P.tiles_x;
      const int tx = t2 - ty * P.tiles_x;
      const uint32_t acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tmem_full[acc], acc_phase);
      tcgen05_fence_after();
      const int m = q * 32 + lane;
      const int gy = ty * kDwTH + (m >> 4);
      const int gx = tx * kDwTW + (m & 15);
      const bool ok = gy < P.H && gx < P.W;
      __nv_bfloat16* o = P.out + ((static_cast<size_t>(b) * P.H + gy) * P.W + gx) * kDwBN;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * kDwBN;
      if constexpr (Cfg::kTmaStore) {
        // bf16 output through 128B-swizzled smem + TMA store: box = this warp's 2 tile rows x 16 pixels x 64 channels;
        // partial tiles are clipped by the TMA unit
        uint8_t* my_c = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_c) + 1023) & ~uintptr_t(1023)) + (warp - 2) * 4096;
        const uint32_t cbuf = smem_u32(my_c) + lane * 128;
#pragma unroll 1
        for (int cb = 0; cb < kDwBN / 64; ++cb) {
          uint32_t pk[32];
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int c0 = cb * 64 + half * 32;
            uint32_t v[32];
            tmem_ld_32x32b_x32(taddr + c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const ulonglong2 s = *reinterpret_cast<const ulonglong2*>(s_scale + c0 + j);   // two packed fp32 pairs
              const ulonglong2 t = *reinterpret_cast<const ulonglong2*>(s_shift + c0 + j);
              pk[half * 16 + j / 2 + 0] = relu_bf16x2(f32x2_to_bf16x2(f32x2_fma(f32x2_make(v[j + 0], v[j + 1]), s.x, t.x)));
              pk[half * 16 + j / 2 + 1] = relu_bf16x2(f32x2_to_bf16x2(f32x2_fma(f32x2_make(v[j + 2], v[j + 3]), s.y, t.y)));
            }
          }
          if (lane == 0) tma_store_wait_read<0>();   // the previous block has been read out of the staging buffer
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t chunk = static_cast<uint32_t>(j) ^ static_cast<uint32_t>(lane & 7);
            sts_v4(cbuf + chunk * 16, make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]));
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0 && !(P.debug & 1)) {
            asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                             reinterpret_cast<uint64_t>(P.tmap_out)),
                         "r"(smem_u32(my_c)), "r"(cb * 64), "r"(tx * kDwTW), "r"(ty * kDwTH + 2 * q), "r"(b)
                         : "memory");
            tma_store_commit();
          }
        }
      } else {
#pragma unroll 1
      for (int c0 = 0; c0 < kDwBN; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr + c0, v);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const ulonglong2 s = *reinterpret_cast<const ulonglong2*>(s_scale + c0 + j);   // two packed fp32 pairs
          const ulonglong2 t = *reinterpret_cast<const ulonglong2*>(s_shift + c0 + j);
          const unsigned long long y01 = f32x2_fma(f32x2_make(v[j + 0], v[j + 1]), s.x, t.x);
          const unsigned long long y23 = f32x2_fma(f32x2_make(v[j + 2], v[j + 3]), s.y, t.y);
          pk[j / 2 + 0] = relu_bf16x2(f32x2_to_bf16x2(y01));
          pk[j / 2 + 1] = relu_bf16x2(f32x2_to_bf16x2(y23));
        }
        if (ok && !(P.debug & 1)) {
#pragma unroll
          for (int j = 0; j < 2; ++j)
            stg_v8(o + c0 + j * 16, make_uint4(pk[8 * j], pk[8 * j + 1], pk[8 * j + 2], pk[8 * j + 3]),
                   make_uint4(pk[8 * j + 4], pk[8 * j + 5], pk[8 * j + 6], pk[8 * j + 7]));
        }
      }
      }
      tcgen05_fence_before();
      mbar_arrive(&tmem_empty[acc]);
    }
    if (Cfg::kTmaStore && lane == 0) tma_store_wait_all<0>();
  } else {
    // ------------------------------------------------------------------ depthwise stencil warps
    // All eight stencil warps work on the SAME K block: warps 6..9 produce output rows 0..3 of the tile, warps 10..13
    // rows 4..7 (each half reads 6 halo rows).  The halo stages are therefore a true ring: the TMA of block c+1 is in
    // flight while block c is being computed.  Every barrier has one producer set and one consumer set that advance
    // together, so each parity wait is at most one phase behind.
    const int g = (warp - 6) >> 2;          // row half 0/1
    const int wg = (warp - 6) & 3;          // warp within the half
    const int v4 = lane & 15;               // which 4-channel slice of the 64-channel K block
    const int cp = wg * 2 + (lane >> 4);    // column pair: output cols 2cp, 2cp+1
    constexpr int kRows = kDwTH / 2;        // output rows per half
    const uint32_t in_base = smem_u32(smem_in);
    const uint32_t a_base = smem_u32(smem_a);
    int my_tiles = 0;
    for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x) ++my_tiles;
    const uint32_t total_c = static_cast<uint32_t>(my_tiles) * KB;   // this CTA's K blocks, numbered over (tile, kb)
    unsigned long long wlo[9], whi[9], sh_lo, sh_hi;
    auto load_taps = [&](uint32_t cn) {   // taps + shift of this thread's 4 channels; L1 is ~0 KB here -> L2 latency, so issue early
      const int ch = static_cast<int>(cn % KB) * 64 + v4 * 4;
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(P.dw_w + t * (KB * 64) + ch));
        wlo[t] = pack_f32x2(w.x, w.y);
        whi[t] = pack_f32x2(w.z, w.w);
      }
      const float4 sh = __ldg(reinterpret_cast<const float4*>(P.dw_shift + ch));
      sh_lo = pack_f32x2(sh.x, sh.y);
      sh_hi = pack_f32x2(sh.z, sh.w);
    };
    if (total_c > 0) load_taps(0);
    for (uint32_t c = 0; c < total_c; ++c) {
      const uint32_t si = c % kDwInStages;
      const uint32_t sa = c % AS;
      if (!free_run) mbar_wait(&in_full[si], (c / kDwInStages) & 1);
      const uint32_t in_addr = in_base + si * kDwInStageBytes + ((g * kRows) * kDwHaloW + 2 * cp) * 128 + v4 * 8;
      const uint32_t a_addr = a_base + sa * kDwAStageBytes;

      // rolling window over this half's 6 halo rows; 3 output rows in flight, 2 output columns, 2 channel pairs.
      // The raw loads of halo row r+1 are issued before the math of row r (software pipelining of the LDS latency).
      unsigned long long acc_lo[3][2], acc_hi[3][2];
      uint2 raw_next[4];
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) raw_next[cc] = lds_v2(in_addr + cc * 128);
#pragma unroll
      for (int r = 0; r < kRows + 2; ++r) {
        unsigned long long x_lo[4], x_hi[4];
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          x_lo[cc] = bf16x2_to_f32x2(raw_next[cc].x);
          x_hi[cc] = bf16x2_to_f32x2(raw_next[cc].y);
        }
        if (r + 1 < kRows + 2) {
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) raw_next[cc] = lds_v2(in_addr + ((r + 1) * kDwHaloW + cc) * 128);
        }
        if (r < kRows) {  // output row r starts with halo row r (dy = 0)
#pragma unroll
          for (int oc = 0; oc < 2; ++oc) {
            acc_lo[r % 3][oc] = sh_lo;
            acc_hi[r % 3][oc] = sh_hi;
          }
        }
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
          const int orow = r - dy;
          if (orow < 0 || orow >= kRows) continue;
#pragma unroll
          for (int oc = 0; oc < 2; ++oc)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
              ffma2(acc_lo[orow % 3][oc], wlo[dy * 3 + dx], x_lo[oc + dx]);
              ffma2(acc_hi[orow % 3][oc], whi[dy * 3 + dx], x_hi[oc + dx]);
            }
        }
        if (r >= 2) {  // output row r-2 of this half is complete
          const int orow = r - 2;
          if (orow == 0 && !free_run) mbar_wait(&a_empty[sa], ((c / AS) & 1) ^ 1);  // the MMA is done with this A stage
#pragma unroll
          for (int oc = 0; oc < 2; ++oc) {
            const unsigned long long lo = acc_lo[orow % 3][oc], hi = acc_hi[orow % 3][oc];
            const uint32_t p0 = relu_bf16x2(f32x2_to_bf16x2(lo));
            const uint32_t p1 = relu_bf16x2(f32x2_to_bf16x2(hi));
            const int m = (g * kRows + orow) * kDwTW + 2 * cp + oc;
            const uint32_t chunk = static_cast<uint32_t>(v4 >> 1) ^ static_cast<uint32_t>(m & 7);
            if (!(P.debug & 2)) sts_v2(a_addr + m * 128 + chunk * 16 + (v4 & 1) * 8, p0, p1);
          }
        }
      }
      if (!free_run) mbar_arrive(&in_empty[si]);   // halo tile fully consumed by this thread
      fence_proxy_async_smem();                    // make the A tile visible to the tensor-core (async) proxy
      if (!free_run) mbar_arrive(&a_full[sa]);
      if (c + 1 < total_c) load_taps(c + 1);       // next block's taps: the L2 latency hides behind the barrier waits
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace dlv3p
