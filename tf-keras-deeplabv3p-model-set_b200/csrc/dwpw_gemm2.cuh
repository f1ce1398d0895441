// dwpw_gemm2.cuh — the fused SepConv_BN of dwpw_gemm.cuh on CTA PAIRS (tcgen05 cta_group::2).
//   depthwise 3x3 'same' -> BN -> ReLU -> pointwise 1x1 (C -> 256) -> BN -> ReLU
// (reference deeplabv3p/models/layers.py:74-111, used by Decoder_block :215-218).
//
// Two CTAs of a cluster (two SMs of a TPC) each own one 8 x 16-pixel tile: each CTA stages its own halo tiles, runs its
// own depthwise stencil into its own A stages, and keeps only HALF of the pointwise weights (128 of the 256 output
// channels) resident.  One thread of the leader CTA issues tcgen05.mma.cta_group::2 (M = 256): the tensor cores of
// both SMs read both weight halves.  Against the 1-CTA kernel this
//   * halves the shared-memory read traffic of the B operand (the MMA of the 1-CTA kernel alone uses 75 % of the
//     128 B/clk shared-memory bandwidth, leaving too little for the stencil's loads and stores),
//   * frees 64-80 KB of shared memory: three A stages and three to four halo stages, so the stencil warps, the TMA and the
//     tensor cores run decoupled instead of in lock step, and every K (<= 320) gets the TMA-store epilogue.
//
//   in_full[s]    (own CTA)   1 arrival + tx bytes : halo tile landed
//   in_empty[s]   (own CTA)   8 arrivals           : one per stencil warp
//   a_full[s]     (leader's)  16 arrivals          : the 8 stencil warps of both CTAs
//   a_empty[s]    (each CTA)  1 arrival            : tcgen05.commit multicast from the leader's MMA thread
//   tmem_full[a]  (each CTA)  1 arrival            : tcgen05.commit multicast
//   tmem_empty[a] (leader's)  256 arrivals         : the 4 epilogue warps of both CTAs
#pragma once

#include <cuda.h>

#include "dwpw_gemm.cuh"
#include "pw_gemm2.cuh"

namespace dlv3p {

template <int KB>
struct DwPw2Cfg {
  static constexpr int kAS = 3;
  static constexpr int kIS = KB <= 4 ? 4 : 3;
  static constexpr int kWHalfBlockBytes = 128 * 128;               // this CTA's 128 output channels x 64 input channels
  static constexpr int kWBytes = KB * kWHalfBlockBytes;
  static constexpr int kABytes = kAS * kDwAStageBytes;
  static constexpr int kInBytes = kIS * kDwInStageBytes;
  static constexpr int kInBytesPad = (kInBytes + 1023) / 1024 * 1024;   // the store staging wants 1024-byte alignment
  static constexpr int kStoreBytes = 4 * 4096;
  static constexpr int kSmemBytes = kWBytes + kABytes + kInBytesPad + kStoreBytes + 2048 /*BN scale+shift*/ + 256 /*barriers*/;
};

template <int KB>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kDwThreads, 1) dwpw_gemm2_kernel(const __grid_constant__ DwPwParams P) {
  using Cfg = DwPw2Cfg<KB>;
  constexpr int AS = Cfg::kAS, IS = Cfg::kIS;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  uint8_t* smem_w = smem;                          // KB x [128 rows x 128 B], swizzled: this CTA's half of the weights
  uint8_t* smem_a = smem_w + Cfg::kWBytes;         // AS x [128 rows x 128 B], swizzled
  uint8_t* smem_in = smem_a + Cfg::kABytes;        // IS x [10][18][64] bf16
  uint8_t* smem_c = smem_in + Cfg::kInBytesPad;    // epilogue store staging, 4 warps x 4 KB
  float* s_scale = reinterpret_cast<float*>(smem_c + Cfg::kStoreBytes);
  float* s_shift = s_scale + 256;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_shift + 256);
  uint64_t* w_full = bars;                   // [1]
  uint64_t* in_full = bars + 1;              // [IS]
  uint64_t* in_empty = in_full + IS;         // [IS]
  uint64_t* a_full = in_empty + IS;          // [AS]
  uint64_t* a_empty = a_full + AS;           // [AS]
  uint64_t* tmem_full = a_empty + AS;        // [2]
  uint64_t* tmem_empty = tmem_full + 2;      // [2]
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int num_items = (P.num_tiles + 1) / 2;               // one item = two tiles, one per CTA of the pair
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int tiles_per_img = P.tiles_x * P.tiles_y;

  if (warp == 0 && lane == 0) {
    mbar_init(w_full, 1);
    for (int i = 0; i < IS; ++i) {
      mbar_init(&in_full[i], 1);
      mbar_init(&in_empty[i], 8);
    }
    for (int i = 0; i < AS; ++i) {
      mbar_init(&a_full[i], 16);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 256);
    }
    fence_barrier_init();
  }
  if (threadIdx.x >= 64 && threadIdx.x < 64 + 256) {
    s_scale[threadIdx.x - 64] = P.scale[threadIdx.x - 64];
    s_shift[threadIdx.x - 64] = P.shift[threadIdx.x - 64];
  }
  cluster_sync_all();                       // barriers of both CTAs exist before anyone signals across the pair
  if (warp == 1) {
    tmem_alloc_2sm(tmem_base_ptr, 512);
    tmem_relinquish_2sm();
  }
  tcgen05_fence_before();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (elect_one()) {
      if (leader) mbar_arrive_expect_tx(w_full, 2 * Cfg::kWBytes);
      for (int kb = 0; kb < KB; ++kb)   // rows [rank*128, rank*128+128) of the [256, K] weight matrix; bytes credited to the leader
        tma_load_2d_2sm(smem_w + kb * Cfg::kWHalfBlockBytes, P.tmap_w, w_full, kb * 64, static_cast<int>(rank) * 128, kEvictLast);
      uint32_t c = 0;
      for (int item = cluster_id; item < num_items; item += num_clusters) {
        const int tile = item * 2 + static_cast<int>(rank);   // tile == num_tiles (odd count): image B is out of bounds -> zero fill
        const int b = tile / tiles_per_img;
        const int t2 = tile - b * tiles_per_img;
        const int ty = t2 / P.tiles_x;
        const int tx = t2 - ty * P.tiles_x;
        for (int kb = 0; kb < KB; ++kb, ++c) {
          const uint32_t si = c % IS;
          const uint32_t ph = (c / IS) & 1;
          mbar_wait(&in_empty[si], ph ^ 1);
          mbar_arrive_expect_tx(&in_full[si], kDwInStageBytes);
          tma_load_4d(smem_in + si * kDwInStageBytes, P.tmap_x, &in_full[si], kb * 64, tx * kDwTW - 1, ty * kDwTH - 1, b, kEvictNormal);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (leader) {
      constexpr uint32_t idesc = make_idesc_bf16(256, kDwBN);
      mbar_wait(w_full, 0);
      uint32_t c = 0, it = 0;
      for (int item = cluster_id; item < num_items; item += num_clusters, ++it) {
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
              const uint64_t db = make_smem_desc_sw128(smem_u32(smem_w + kb * Cfg::kWHalfBlockBytes));
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16_ss_2sm(tmem_d, smem_desc_advance(da, k * 32), smem_desc_advance(db, k * 32), idesc, (kb > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit_2sm(&a_empty[sa]);
            if (kb == KB - 1) umma_commit_2sm(&tmem_full[acc]);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp < 6) {
    // ------------------------------------------------------------------ epilogue (both CTAs, own tile)
    const int q = warp & 3;
    uint32_t it = 0;
    uint8_t* my_c = smem_c + (warp - 2) * 4096;
    const uint32_t cbuf = smem_u32(my_c) + lane * 128;
    for (int item = cluster_id; item < num_items; item += num_clusters, ++it) {
      const int tile = item * 2 + static_cast<int>(rank);
      const int b = tile / tiles_per_img;
      const int t2 = tile - b * tiles_per_img;
      const int ty = t2 / P.tiles_x;
      const int tx = t2 - ty * P.tiles_x;
      const uint32_t acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tmem_full[acc], acc_phase);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * kDwBN;
      // bf16 output through 128B-swizzled smem + TMA store: box = this warp's 2 tile rows x 16 pixels x 64 channels;
      // partial (and phantom) tiles are clipped by the TMA unit
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
            const ulonglong2 s = *reinterpret_cast<const ulonglong2*>(s_scale + c0 + j);
            const ulonglong2 t = *reinterpret_cast<const ulonglong2*>(s_shift + c0 + j);
            pk[half * 16 + j / 2 + 0] = relu_bf16x2(f32x2_to_bf16x2(f32x2_fma(f32x2_make(v[j + 0], v[j + 1]), s.x, t.x)));
            pk[half * 16 + j / 2 + 1] = relu_bf16x2(f32x2_to_bf16x2(f32x2_fma(f32x2_make(v[j + 2], v[j + 3]), s.y, t.y)));
          }
        }
        if (cb == kDwBN / 64 - 1) {   // every column of this accumulator stage is in registers: hand it back early
          tcgen05_fence_before();
          if (leader) mbar_arrive(&tmem_empty[acc]);
          else mbar_arrive_cluster(&tmem_empty[acc], 0);
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
    }
    if (lane == 0) tma_store_wait_all<0>();
  } else {
    // ------------------------------------------------------------------ depthwise stencil warps (both CTAs, own tile)
    // All eight stencil warps work on the SAME K block: warps 6..9 produce output rows 0..3 of the tile, warps 10..13
    // rows 4..7 (each half reads 6 halo rows).
    const int g = (warp - 6) >> 2;          // row half 0/1
    const int wg = (warp - 6) & 3;          // warp within the half
    const int v4 = lane & 15;               // which 4-channel slice of the 64-channel K block
    const int cp = wg * 2 + (lane >> 4);    // column pair: output cols 2cp, 2cp+1
    constexpr int kRows = kDwTH / 2;        // output rows per half
    const uint32_t in_base = smem_u32(smem_in);
    const uint32_t a_base = smem_u32(smem_a);
    int my_items = 0;
    for (int item = cluster_id; item < num_items; item += num_clusters) ++my_items;
    const uint32_t total_c = static_cast<uint32_t>(my_items) * KB;
    unsigned long long wlo[9], whi[9], sh_lo, sh_hi;
    auto load_taps = [&](uint32_t cn) {   // taps + shift of this thread's 4 channels (L2 latency: issued one K block early)
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
      const uint32_t si = c % IS;
      const uint32_t sa = c % AS;
      mbar_wait(&in_full[si], (c / IS) & 1);
      const uint32_t in_addr = in_base + si * kDwInStageBytes + ((g * kRows) * kDwHaloW + 2 * cp) * 128 + v4 * 8;
      const uint32_t a_addr = a_base + sa * kDwAStageBytes;

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
        if (r < kRows) {
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
        if (r >= 2) {
          const int orow = r - 2;
          if (orow == 0) mbar_wait(&a_empty[sa], ((c / AS) & 1) ^ 1);  // the MMA is done with this A stage
#pragma unroll
          for (int oc = 0; oc < 2; ++oc) {
            const uint32_t p0 = relu_bf16x2(f32x2_to_bf16x2(acc_lo[orow % 3][oc]));
            const uint32_t p1 = relu_bf16x2(f32x2_to_bf16x2(acc_hi[orow % 3][oc]));
            const int m = (g * kRows + orow) * kDwTW + 2 * cp + oc;
            const uint32_t chunk = static_cast<uint32_t>(v4 >> 1) ^ static_cast<uint32_t>(m & 7);
            if (!(P.debug & 2)) sts_v2(a_addr + m * 128 + chunk * 16 + (v4 & 1) * 8, p0, p1);
          }
        }
      }
      fence_proxy_async_smem();                    // make this thread's A rows visible to the tensor-core (async) proxy
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&in_empty[si]);                // the warp is done reading the halo tile
        if (leader) mbar_arrive(&a_full[sa]);
        else mbar_arrive_cluster(&a_full[sa], 0);
      }
      if (c + 1 < total_c) load_taps(c + 1);
    }
  }

  tcgen05_fence_before();
  cluster_sync_all();                      // the peer may still be reading this CTA's smem / signalling its barriers
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}

}  // namespace dlv3p
