// pw_gemm2.cuh — the pointwise GEMM of pw_gemm.cuh on CTA PAIRS (tcgen05 cta_group::2), N = 256, bf16 output.
//
// Two CTAs of a cluster (two SMs of one TPC) compute a 256 x 256 output tile together: each CTA stages its own 128 rows
// of A and only HALF of the weight tile (128 of the 256 N rows); one thread of the leader CTA issues
// tcgen05.mma.cta_group::2 (M = 256), the hardware feeds both halves of B to both tensor cores.  Per CTA this halves
// the shared-memory read traffic of B and the L2 -> SM traffic of the weights — the two things that bound the 1-CTA
// kernel — and frees smem for a 6-deep pipeline.
//
//   full_bar[s]   (leader's)  2 arrivals: leader's arrive.expect_tx + peer's remote arrive; TMA bytes of BOTH CTAs
//   empty_bar[s]  (each CTA)  1 arrival : tcgen05.commit multicast from the leader's MMA thread
//   tmem_full[a]  (each CTA)  1 arrival : tcgen05.commit multicast
//   tmem_empty[a] (leader's)  256 arrivals: the 4 epilogue warps of both CTAs
#pragma once

#include <cuda.h>

#include "pw_gemm.cuh"

namespace dlv3p {

constexpr int kPw2BN = 256;
constexpr int kPw2Stages = 6;
constexpr int kPw2StageBytes = kPwBM * 128 + (kPw2BN / 2) * 128;   // A 16 KB + half of B 16 KB per CTA
constexpr int kPw2StoreBytes = 4 * 2 * 4096;
constexpr int kPw2SmemBytes = kPw2Stages * kPw2StageBytes + kPw2StoreBytes + 2 * kPw2BN * 4 + 256;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPwThreads, 1) pw_gemm2_kernel(const __grid_constant__ PwLaunch L) {
  constexpr int BN = kPw2BN;
  constexpr int kStages = kPw2Stages;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  uint8_t* smem_a = smem;                                   // kStages x [128 rows x 128 B]
  uint8_t* smem_b = smem + kStages * (kPwBM * 128);         // kStages x [128 N rows x 128 B]  (this CTA's half of B)
  uint8_t* smem_c = smem + kStages * kPw2StageBytes;        // epilogue staging
  float* s_scale = reinterpret_cast<float*>(smem_c + kPw2StoreBytes);
  float* s_shift = s_scale + BN;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_shift + BN);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* tmem_full = bars + 2 * kStages;
  uint64_t* tmem_empty = bars + 2 * kStages + 2;
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int num_pairs = (L.num_tiles + 1) / 2;
  const int total_items = num_pairs * L.num_problems;        // one item = a pair of M tiles of one problem
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 2);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 256);
    }
    fence_barrier_init();
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
      uint32_t stage = 0, phase = 0;
      for (int item = cluster_id; item < total_items; item += num_clusters) {
        const int p = item % L.num_problems;
        const int tile = (item / L.num_problems) * 2 + static_cast<int>(rank);
        const PwProblem& P = L.prob[p];
        const int kblocks = (P.K + kPwBK - 1) / kPwBK;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * kPw2StageBytes);
          else mbar_arrive_cluster(&full_bar[stage], 0);
          if (P.a_kblock_rows > 0)
            tma_load_2d_2sm(smem_a + stage * (kPwBM * 128), P.tmap_a, &full_bar[stage], 0, kb * P.a_kblock_rows + tile * kPwBM, kEvictFirst);
          else
            tma_load_2d_2sm(smem_a + stage * (kPwBM * 128), P.tmap_a, &full_bar[stage], kb * kPwBK, tile * kPwBM, kEvictFirst);
          tma_load_2d_2sm(smem_b + stage * ((BN / 2) * 128), P.tmap_w, &full_bar[stage], kb * kPwBK, static_cast<int>(rank) * (BN / 2), kEvictLast);
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
        const int p = item % L.num_problems;
        const int kblocks = (L.prob[p].K + kPwBK - 1) / kPwBK;
        const uint32_t acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          if (elect_one()) {
            const uint64_t da = make_smem_desc_sw128(smem_u32(smem_a + stage * (kPwBM * 128)));
            const uint64_t db = make_smem_desc_sw128(smem_u32(smem_b + stage * ((BN / 2) * 128)));
#pragma unroll
            for (int k = 0; k < kPwBK / 16; ++k)
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
    uint32_t store_buf = 0;
    int ss_key = -1;
    for (int item = cluster_id; item < total_items; item += num_clusters, ++it) {
      const int p = item % L.num_problems;
      const int tile = (item / L.num_problems) * 2 + static_cast<int>(rank);
      const PwProblem& P = L.prob[p];
      const uint32_t acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tmem_full[acc], acc_phase);
      tcgen05_fence_after();
      const int row = tile * kPwBM + q * 32 + lane;
      const bool row_ok = row < L.M;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
      const int img = (P.epi == kEpiBf16ImgShift && row_ok) ? row / L.rows_per_img : 0;
      {
        const int row_first = min(tile * kPwBM, L.M - 1), row_last = min(tile * kPwBM + kPwBM, L.M) - 1;
        const int img_first = P.epi == kEpiBf16ImgShift ? row_first / L.rows_per_img : 0;
        const int img_last = P.epi == kEpiBf16ImgShift ? max(row_last, row_first) / L.rows_per_img : 0;
        const int key = (p << 24) | (img_first == img_last ? img_first : 0xFFFFFF);
        if (key != ss_key) {
          asm volatile("bar.sync 1, 128;" ::: "memory");
          const float* gshift = (P.epi == kEpiBf16ImgShift && img_first == img_last) ? P.img_shift + static_cast<size_t>(img_first) * BN : P.shift;
          for (int i = (warp - 2) * 32 + lane; i < BN; i += 128) {
            s_scale[i] = __ldg(P.scale + i);
            s_shift[i] = __ldg(gshift + i);
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
          ss_key = key;
        }
      }
      const bool shift_global = P.epi == kEpiBf16ImgShift && (ss_key & 0xFFFFFF) == 0xFFFFFF;
      const float* shift = shift_global ? P.img_shift + static_cast<size_t>(img) * BN : s_shift;
      uint8_t* my_c = smem_c + (warp - 2) * 2 * 4096;
#pragma unroll 1
      for (int cb = 0; cb < BN / 64; ++cb) {
        if (cb * 64 >= P.N) break;
        if (lane == 0) tma_store_wait_read<1>();
        __syncwarp();
        const uint32_t cbuf = smem_u32(my_c + store_buf * 4096) + lane * 128;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int c0 = cb * 64 + half * 32;
          uint32_t v[32];
          tmem_ld_32x32b_x32(taddr + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            const ulonglong2 s0 = *reinterpret_cast<const ulonglong2*>(s_scale + c0 + j);
            const ulonglong2 s1 = *reinterpret_cast<const ulonglong2*>(s_scale + c0 + j + 4);
            const ulonglong2 t0 = *reinterpret_cast<const ulonglong2*>(shift + c0 + j);
            const ulonglong2 t1 = *reinterpret_cast<const ulonglong2*>(shift + c0 + j + 4);
            uint32_t q0 = f32x2_to_bf16x2(f32x2_fma(f32x2_make(v[j + 0], v[j + 1]), s0.x, t0.x));
            uint32_t q1 = f32x2_to_bf16x2(f32x2_fma(f32x2_make(v[j + 2], v[j + 3]), s0.y, t0.y));
            uint32_t q2 = f32x2_to_bf16x2(f32x2_fma(f32x2_make(v[j + 4], v[j + 5]), s1.x, t1.x));
            uint32_t q3 = f32x2_to_bf16x2(f32x2_fma(f32x2_make(v[j + 6], v[j + 7]), s1.y, t1.y));
            if (P.relu) { q0 = relu_bf16x2(q0); q1 = relu_bf16x2(q1); q2 = relu_bf16x2(q2); q3 = relu_bf16x2(q3); }
            const uint32_t chunk = static_cast<uint32_t>(half * 4 + (j >> 3)) ^ static_cast<uint32_t>(lane & 7);
            sts_v4(cbuf + chunk * 16, make_uint4(q0, q1, q2, q3));
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0 && !(L.debug & 1)) {
          tma_store_2d(P.tmap_out, my_c + store_buf * 4096, cb * 64, tile * kPwBM + q * 32);
          tma_store_commit();
        }
        store_buf ^= 1;
      }
      tcgen05_fence_before();
      if (leader) mbar_arrive(&tmem_empty[acc]);
      else mbar_arrive_cluster(&tmem_empty[acc], 0);
    }
    if (lane == 0) tma_store_wait_all<0>();
  }

  tcgen05_fence_before();
  cluster_sync_all();                      // the peer may still be reading this CTA's smem / signalling its barriers
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}

}  // namespace dlv3p
