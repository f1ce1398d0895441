// bb_gemm.cuh — the pointwise (1x1) convolutions of the Xception backbone on tcgen05 CTA pairs:
//   D[M, N] = bf16( acc(A[M, K] * W[K, N]) * scale[n] + shift[n]  [ReLU]  [+ residual[M, N]] )
// = DeeplabConv2D(filters, (1,1)) -> CustomBatchNormalization [-> ReLU] [-> add([residual, shortcut])]
// (reference: SepConv_BN deeplabv3p/models/layers.py:105-109 with depth_activation False / True; _xception_block's shortcut conv
//  + BN and the residual add, deeplabv3p/models/deeplabv3p_xception.py:82-90).
//
// Same machine as pw_gemm2.cuh (two CTAs of a cluster share a 256 x 256 tile through cta_group::2; each stages its own 128 rows
// of A and half of the weight tile; 2 TMEM accumulator stages; TMA-store epilogue) with what the backbone adds:
//   * N up to 2048: work items are (M pair, N tile of 256); consecutive items share the A rows (L2 hits), weights stay L2 resident
//   * K and N that are not multiples of 64 / 256 (728): TMA zero-fills the K tail, the store clips the N tail
//   * the residual / shortcut tensor is added in fp32 in the epilogue, before the single bf16 rounding; it reaches the epilogue
//     warps the way the output leaves them: TMA boxes of 32 rows x 64 columns in the 128-byte swizzle, one block ahead of the
//     arithmetic (per-thread global loads of a row-major tensor cost 21 us per middle-flow GEMM: every load paid L2 latency)
//   * the N tile is a template parameter (256 / 192 / 128): the host picks the one that fills the last round of the 74 CTA pairs
//     best (N = 728 on 128 M pairs: 6 rounds of 256 columns vs 7 rounds of 192)
#pragma once

#include <cuda.h>

#include "dwpw_gemm.cuh"     // packed fp32x2 helpers
#include "sm100_prims.cuh"

namespace dlv3p {

constexpr int kBbBM = 128;
constexpr int kBbBK = 64;
constexpr int kBbMaxBN = 256;
constexpr int kBbThreads = 192;
constexpr int kBbStoreBytes = 4 * 2 * 4096;                          // 4 epilogue warps x 2 buffers x [32 rows x 128 B]

template <int BN, bool kRes>
struct BbCfg {
  static constexpr int kStageBytes = kBbBM * 128 + (BN / 2) * 128;   // A 16 KB + this CTA's half of B
  static constexpr int kResBufs = kRes ? BN / 64 : 0;                 // one residual box per 64-column block of the tile, per epilogue warp
  static constexpr int kResBytes = 4 * kResBufs * 4096;
  static constexpr int kFixed = kBbStoreBytes + kResBytes + 2 * kBbMaxBN * 4 + 512;
  static constexpr int kStages = (232448 - kFixed) / kStageBytes < 8 ? (232448 - kFixed) / kStageBytes : 8;   // with / without residual — 256: 4 / 6, 192: 5 / 6, 128: 6 / 8
  static constexpr int kSmemBytes = kStages * kStageBytes + kFixed;
};

struct BbGemmParams {
  const CUtensorMap* tmap_a;     // [M, K] bf16 row-major, box {64, 128}, SWIZZLE_128B
  const CUtensorMap* tmap_w;     // [Npad, Kpad] bf16 K-major (Npad % BN == 0, zero padded), box {64, BN / 2}, SWIZZLE_128B
  const CUtensorMap* tmap_out;   // [M, N] bf16 (row stride = N), box {64, 32}, SWIZZLE_128B
  const CUtensorMap* tmap_res;   // optional residual [M, N] bf16, same box as tmap_out; nullptr = no residual
  const float* scale;            // [Npad] folded BN
  const float* shift;            // [Npad]
  int M, K, N;
  int relu;
  int m_pairs, n_tiles;
  int debug;                     // benchmark aid: bit0 = skip the stores
  // strided A (the 1x1 stride-2 shortcut convolutions, _conv2d_same with kernel_size 1: deeplabv3p_xception.py:44-52): tmap_a is a 4D map
  // {K, Wo, Ho, B} over every second pixel of every second row of the block input, box {64, min(Wo, 128), 128 / min(Wo, 128), 1}: an M tile
  // of 128 consecutive output pixels is whole rows (or a piece of one) of one image.  a_wo = Wo (0: plain [M, K] operand), a_hw = Ho * Wo.
  int a_wo, a_hw;
};

template <int BN, bool kRes>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kBbThreads, 1) bb_gemm_kernel(const __grid_constant__ BbGemmParams P) {
  using Cfg = BbCfg<BN, kRes>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * (kBbBM * 128);
  uint8_t* smem_c = smem + kStages * Cfg::kStageBytes;
  uint8_t* smem_r = smem_c + kBbStoreBytes;
  float* s_scale = reinterpret_cast<float*>(smem_r + Cfg::kResBytes);
  float* s_shift = s_scale + kBbMaxBN;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_shift + kBbMaxBN);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* tmem_full = bars + 2 * kStages;
  uint64_t* tmem_empty = bars + 2 * kStages + 2;
  uint64_t* res_bar = bars + 2 * kStages + 4;        // [4 warps][kResBufs]
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4 + 4 * Cfg::kResBufs);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int total_items = P.m_pairs * P.n_tiles;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int kblocks = (P.K + kBbBK - 1) / kBbBK;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 2);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 256);
    }
    for (int i = 0; i < 4 * Cfg::kResBufs; ++i) mbar_init(&res_bar[i], 1);
    fence_barrier_init();
  }
  cluster_sync_all();
  if (warp == 1) {
    tmem_alloc_2sm(tmem_base_ptr, 512);
    tmem_relinquish_2sm();
  }
  tcgen05_fence_before();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_ptr;
  pdl_launch_dependents();
  pdl_wait();            // everything above overlapped the previous kernel's tail; A, the residual and the output buffer are its business

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (elect_one()) {
      uint32_t stage = 0, phase = 0;
      for (int item = cluster_id; item < total_items; item += num_clusters) {
        const int nt = item % P.n_tiles;
        const int tile = (item / P.n_tiles) * 2 + static_cast<int>(rank);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
          else mbar_arrive_cluster(&full_bar[stage], 0);
          if (P.a_wo) {
            const int pix = tile * kBbBM, img = pix / P.a_hw, rem = pix - img * P.a_hw, y0 = rem / P.a_wo;
            tma_load_4d_2sm(smem_a + stage * (kBbBM * 128), P.tmap_a, &full_bar[stage], kb * kBbBK, rem - y0 * P.a_wo, y0, img, kEvictNormal);
          } else {
            tma_load_2d_2sm(smem_a + stage * (kBbBM * 128), P.tmap_a, &full_bar[stage], kb * kBbBK, tile * kBbBM, kEvictNormal);
          }
          tma_load_2d_2sm(smem_b + stage * ((BN / 2) * 128), P.tmap_w, &full_bar[stage], kb * kBbBK, nt * BN + static_cast<int>(rank) * (BN / 2), kEvictLast);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (leader) {
      constexpr uint32_t idesc = make_idesc_bf16(256, BN);
      uint32_t stage = 0, phase = 0, it = 0;
      for (int item = cluster_id; item < total_items; item += num_clusters, ++it) {
        const uint32_t acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          if (elect_one()) {
            const uint64_t da = make_smem_desc_sw128(smem_u32(smem_a + stage * (kBbBM * 128)));
            const uint64_t db = make_smem_desc_sw128(smem_u32(smem_b + stage * ((BN / 2) * 128)));
#pragma unroll
            for (int k = 0; k < kBbBK / 16; ++k)
              umma_bf16_ss_2sm(tmem_d, smem_desc_advance(da, k * 32), smem_desc_advance(db, k * 32), idesc, (kb > 0 || k > 0) ? 1u : 0u);
            umma_commit_2sm(&empty_bar[stage]);
            if (kb == kblocks - 1) umma_commit_2sm(&tmem_full[acc]);
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5 of both CTAs)
    const int q = warp & 3;
    uint32_t it = 0;
    uint32_t store_buf = 0, res_phase = 0;
    int ss_nt = -1;
    uint8_t* my_c = smem_c + (warp - 2) * 2 * 4096;
    for (int item = cluster_id; item < total_items; item += num_clusters, ++it) {
      const int nt = item % P.n_tiles;
      const int tile = (item / P.n_tiles) * 2 + static_cast<int>(rank);
      const uint32_t acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int n0 = nt * BN;
      const int ncols = min(BN, P.N - n0);
      const int row = tile * kBbBM + q * 32 + lane;
      const bool row_ok = row < P.M;
      if (nt != ss_nt) {      // this N tile's BN scale / shift into shared memory (no L1 next to ~220 KB of dynamic smem)
        asm volatile("bar.sync 1, 128;" ::: "memory");
        for (int i = (warp - 2) * 32 + lane; i < BN; i += 128) {
          s_scale[i] = __ldg(P.scale + n0 + i);
          s_shift[i] = __ldg(P.shift + n0 + i);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        ss_nt = nt;
      }
      const int nblk = (ncols + 63) / 64;
      constexpr bool has_res = kRes;
      uint8_t* my_r = smem_r + (warp - 2) * Cfg::kResBufs * 4096;
      uint64_t* my_rbar = res_bar + (warp - 2) * Cfg::kResBufs;
      // every residual box of this item is requested before the accumulator is waited for: the loads land while the tensor cores
      // work on the tile (the buffers were last read during the previous item's epilogue, which this warp has left)
      if (has_res && lane == 0) {
        for (int cb = 0; cb < nblk; ++cb) {
          mbar_arrive_expect_tx(&my_rbar[cb], 4096);
          tma_load_2d(my_r + cb * 4096, P.tmap_res, &my_rbar[cb], n0 + cb * 64, tile * kBbBM + q * 32, kEvictNormal);
        }
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
      // the TMEM load of the next 32 columns is in flight while the current 32 are scaled, rounded and staged
      uint32_t vbuf[2][32];
      tmem_ld_32x32b_x32(taddr, vbuf[0]);
      tmem_ld_wait();
#pragma unroll 1
      for (int cb = 0; cb < nblk; ++cb) {
        if (lane == 0) tma_store_wait_read<1>();
        __syncwarp();
        if (has_res) mbar_wait(&my_rbar[cb], (res_phase >> cb) & 1u);
        const uint32_t cbuf = smem_u32(my_c + store_buf * 4096) + lane * 128;
        const uint32_t rbuf = smem_u32(my_r + cb * 4096) + lane * 128;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int c0 = cb * 64 + half * 32;
          uint32_t (&v)[32] = vbuf[half];
          if (half == 0) tmem_ld_32x32b_x32(taddr + c0 + 32, vbuf[1]);
          else if (cb + 1 < nblk) tmem_ld_32x32b_x32(taddr + c0 + 32, vbuf[0]);
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            const ulonglong2 s0 = *reinterpret_cast<const ulonglong2*>(s_scale + c0 + j);
            const ulonglong2 s1 = *reinterpret_cast<const ulonglong2*>(s_scale + c0 + j + 4);
            const ulonglong2 t0 = *reinterpret_cast<const ulonglong2*>(s_shift + c0 + j);
            const ulonglong2 t1 = *reinterpret_cast<const ulonglong2*>(s_shift + c0 + j + 4);
            unsigned long long y0 = f32x2_fma(f32x2_make(v[j + 0], v[j + 1]), s0.x, t0.x);
            unsigned long long y1 = f32x2_fma(f32x2_make(v[j + 2], v[j + 3]), s0.y, t0.y);
            unsigned long long y2 = f32x2_fma(f32x2_make(v[j + 4], v[j + 5]), s1.x, t1.x);
            unsigned long long y3 = f32x2_fma(f32x2_make(v[j + 6], v[j + 7]), s1.y, t1.y);
            const uint32_t chunk = static_cast<uint32_t>(half * 4 + (j >> 3)) ^ static_cast<uint32_t>(lane & 7);
            if (P.relu) {                                    // ReLU belongs to the conv branch, before the add
              y0 = f32x2_make(__float_as_uint(fmaxf(f32x2_lo(y0), 0.f)), __float_as_uint(fmaxf(f32x2_hi(y0), 0.f)));
              y1 = f32x2_make(__float_as_uint(fmaxf(f32x2_lo(y1), 0.f)), __float_as_uint(fmaxf(f32x2_hi(y1), 0.f)));
              y2 = f32x2_make(__float_as_uint(fmaxf(f32x2_lo(y2), 0.f)), __float_as_uint(fmaxf(f32x2_hi(y2), 0.f)));
              y3 = f32x2_make(__float_as_uint(fmaxf(f32x2_lo(y3), 0.f)), __float_as_uint(fmaxf(f32x2_hi(y3), 0.f)));
            }
            if (has_res) {                                   // same swizzled position the output chunk goes to
              const uint4 r = lds_v4(rbuf + chunk * 16);
              const unsigned long long one = f32x2_make(0x3F800000u, 0x3F800000u);
              y0 = f32x2_fma(bf16x2_to_f32x2(r.x), one, y0);
              y1 = f32x2_fma(bf16x2_to_f32x2(r.y), one, y1);
              y2 = f32x2_fma(bf16x2_to_f32x2(r.z), one, y2);
              y3 = f32x2_fma(bf16x2_to_f32x2(r.w), one, y3);
            }
            sts_v4(cbuf + chunk * 16, make_uint4(f32x2_to_bf16x2(y0), f32x2_to_bf16x2(y1), f32x2_to_bf16x2(y2), f32x2_to_bf16x2(y3)));
          }
          tmem_ld_wait();      // the load issued at the top of this half
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0 && !(P.debug & 1)) {
          tma_store_2d(P.tmap_out, my_c + store_buf * 4096, n0 + cb * 64, tile * kBbBM + q * 32);
          tma_store_commit();
        }
        store_buf ^= 1;
      }
      if (has_res) res_phase ^= (1u << nblk) - 1u;      // the barriers armed for this item (a short last N tile arms fewer) completed one phase each
      tcgen05_fence_before();
      if (leader) mbar_arrive(&tmem_empty[acc]);
      else mbar_arrive_cluster(&tmem_empty[acc], 0);
    }
    if (lane == 0) tma_store_wait_all<0>();
  }

  tcgen05_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}

}  // namespace dlv3p
