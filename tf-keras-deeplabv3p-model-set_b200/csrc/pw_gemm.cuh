// pw_gemm.cuh — 1x1 convolution (pointwise GEMM) + folded BatchNorm + ReLU on tcgen05 tensor cores.
//
// Replaces: DeeplabConv2D(filters,(1,1)) -> CustomBatchNormalization -> ReLU
//           (reference deeplabv3p/models/layers.py:14-21, 63-70; call sites :134-143, :105-109,
//            :157-160, :209-213 and the classifier deeplabv3p/model.py:75).
//
// D[M, N] = epilogue(A[M, K] * W[K, N]);  A = NHWC activations viewed as [pixels, channels] (bf16),
// W packed [Npad, Kpad] K-major bf16.  One CTA owns a 128 x BN output tile:
//   warp 0      TMA producer   (cp.async.bulk.tensor, 128B swizzle, kStages-deep mbarrier ring)
//   warp 1      MMA issuer     (one elected lane issues tcgen05.mma, fp32 accumulators in TMEM,
//                               two accumulator stages so the epilogue overlaps the next tile)
//   warps 2..5  epilogue       (tcgen05.ld -> scale/shift/ReLU -> bf16 -> 128B-swizzled smem -> TMA store, one
//                               32-row x 64-column box per warp, double buffered; fp32 planar logits are stored
//                               directly: consecutive lanes = consecutive pixels of one class plane)
// Persistent over (tile, problem) work items; a launch can carry up to kMaxProblems GEMMs that
// share M (the four ASPP branches: aspp0 + three atrous pointwise convs).
#pragma once

#include <cuda.h>

#include "sm100_prims.cuh"

namespace dlv3p {

constexpr int kPwBM = 128;
constexpr int kPwBK = 64;
constexpr int kPwThreads = 192;
constexpr int kMaxProblems = 4;

enum PwEpilogue : int {
  kEpiBf16 = 0,       // out bf16 [M, ldo] at column col_off
  kEpiBf16ImgShift = 1,  // same, shift taken per image: img_shift[(row / rows_per_img) * BN + n]
  kEpiPlanarF32 = 2   // out fp32 planar [B, N, rows_per_img] (classifier logits), no ReLU unless asked
};

struct PwProblem {
  const CUtensorMap* tmap_a;  // [M, K] bf16, box {64, 128}, SWIZZLE_128B   (device memory)
  const CUtensorMap* tmap_w;  // [Npad, Kpad] bf16, box {64, min(BN, 128)}, SWIZZLE_128B
  const CUtensorMap* tmap_out;  // bf16 modes, BN >= 64: [M, N] view of the output (row stride ldo), box {64, 32}, SWIZZLE_128B
  const float* scale;         // [BN]
  const float* shift;         // [BN]
  const float* img_shift;     // [B, BN] (kEpiBf16ImgShift)
  void* out;
  int K;
  int N;           // valid output columns (<= BN)
  int ldo;         // bf16 modes: leading dimension (elements)
  int col_off;     // bf16 modes: first output column
  int relu;
  int epi;
  int a_kblock_rows;  // 0: A is [M, K] row-major.  > 0: A is K-block-major [K/64][a_kblock_rows][64] (ASPP depthwise outputs)
};

struct PwLaunch {
  PwProblem prob[kMaxProblems];
  int num_problems;
  int M;
  int num_tiles;      // ceil(M / 128)
  int rows_per_img;   // pixels per image (per-image shift, planar output)
  int debug;          // benchmark aid: bit0 = skip the global stores of the epilogue
};

template <int BN>
struct PwCfg {
  static constexpr int kStageBytes = kPwBM * 128 + BN * 128;
  static constexpr int kStages = (BN >= 256) ? 4 : 6;
  static constexpr int kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr int kStoreBytes = (BN >= 64) ? 4 * 2 * 4096 : 0;  // 4 epilogue warps x 2 buffers x [32 rows x 128 B]
  static constexpr int kSmemBytes = kStages * kStageBytes + kStoreBytes + 2 * BN * 4 /*scale+shift*/ + 256 /*barriers*/;
};

template <int BN>
__global__ void __launch_bounds__(kPwThreads, 1) pw_gemm_kernel(const __grid_constant__ PwLaunch L) {
  using Cfg = PwCfg<BN>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment required by the 128B swizzle atoms
  uint8_t* smem = smem_raw;   // dynamic smem is the only shared memory of this kernel: 1024-byte aligned (checked below)
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  uint8_t* smem_a = smem;                                  // kStages x [128 rows x 128 B]
  uint8_t* smem_b = smem + kStages * (kPwBM * 128);        // kStages x [BN rows x 128 B]
  uint8_t* smem_c = smem + kStages * Cfg::kStageBytes;     // epilogue staging (1024-byte aligned)
  float* s_scale = reinterpret_cast<float*>(smem_c + Cfg::kStoreBytes);  // [BN] current problem's BN scale (L1 is ~0 KB here)
  float* s_shift = s_scale + BN;                                          // [BN] shift (per image in kEpiBf16ImgShift)
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_shift + BN);
  uint64_t* full_bar = bars;                    // [kStages]  TMA -> MMA
  uint64_t* empty_bar = bars + kStages;         // [kStages]  MMA -> TMA
  uint64_t* tmem_full = bars + 2 * kStages;     // [2]        MMA -> epilogue
  uint64_t* tmem_empty = bars + 2 * kStages + 2;  // [2]      epilogue -> MMA
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_items = L.num_tiles * L.num_problems;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
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
    tmem_alloc(tmem_base_ptr, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      uint32_t stage = 0, phase = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        const int p = item % L.num_problems;
        const int tile = item / L.num_problems;
        const PwProblem& P = L.prob[p];
        const int kblocks = (P.K + kPwBK - 1) / kPwBK;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          if (P.a_kblock_rows > 0)
            tma_load_2d(smem_a + stage * (kPwBM * 128), P.tmap_a, &full_bar[stage], 0, kb * P.a_kblock_rows + tile * kPwBM, kEvictFirst);
          else
            tma_load_2d(smem_a + stage * (kPwBM * 128), P.tmap_a, &full_bar[stage], kb * kPwBK, tile * kPwBM, kEvictFirst);
          tma_load_2d(smem_b + stage * (BN * 128), P.tmap_w, &full_bar[stage], kb * kPwBK, 0, kEvictLast);
          if constexpr (BN > 128)   // weight boxes hold 128 rows (shared with the CTA-pair kernel)
            tma_load_2d(smem_b + stage * (BN * 128) + 128 * 128, P.tmap_w, &full_bar[stage], kb * kPwBK, 128, kEvictLast);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = make_idesc_bf16(kPwBM, BN);
    uint32_t stage = 0, phase = 0, it = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
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
          const uint64_t db = make_smem_desc_sw128(smem_u32(smem_b + stage * (BN * 128)));
#pragma unroll
          for (int k = 0; k < kPwBK / 16; ++k) {
            umma_bf16_ss(tmem_d, smem_desc_advance(da, k * 32), smem_desc_advance(db, k * 32), idesc,
                         (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);                      // frees the smem slot when the MMAs retire
          if (kb == kblocks - 1) umma_commit(&tmem_full[acc]);  // accumulator complete
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5)
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    uint32_t it = 0;
    uint32_t store_buf = 0;
    int ss_key = -1;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
      const int p = item % L.num_problems;
      const int tile = item / L.num_problems;
      const PwProblem& P = L.prob[p];
      const uint32_t acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tmem_full[acc], acc_phase);
      tcgen05_fence_after();
      const int row = tile * kPwBM + q * 32 + lane;
      const bool row_ok = row < L.M;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
      int img = 0, pix = 0;
      if (P.epi != kEpiBf16) {
        img = row_ok ? row / L.rows_per_img : 0;
        pix = row - img * L.rows_per_img;
      }
      // scale / shift of the current problem (and image, for the per-image shift) live in shared memory: with ~224 KB
      // of dynamic smem there is no L1 left and every __ldg would pay L2 latency inside the epilogue's critical path
      {
        const int row_first = tile * kPwBM, row_last = min(tile * kPwBM + kPwBM, L.M) - 1;
        const int img_first = P.epi == kEpiBf16ImgShift ? row_first / L.rows_per_img : 0;
        const int img_last = P.epi == kEpiBf16ImgShift ? row_last / L.rows_per_img : 0;
        const int key = (p << 24) | (img_first == img_last ? img_first : 0xFFFFFF);
        if (key != ss_key) {
          asm volatile("bar.sync 1, 128;" ::: "memory");   // everyone is done with the previous values
          const float* gshift = (P.epi == kEpiBf16ImgShift && img_first == img_last) ? P.img_shift + static_cast<size_t>(img_first) * BN : P.shift;
          for (int i = (warp - 2) * 32 + lane; i < BN; i += 128) {
            s_scale[i] = __ldg(P.scale + i);
            s_shift[i] = __ldg(gshift + i);
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
          ss_key = key;
        }
      }
      // a tile that straddles two images (per-image shift only) reads its shift from global memory instead
      const bool shift_global = P.epi == kEpiBf16ImgShift && (ss_key & 0xFFFFFF) == 0xFFFFFF;
      const float* shift = shift_global ? P.img_shift + static_cast<size_t>(img) * BN : s_shift;
      bool staged = false;
      if constexpr (BN >= 64) staged = P.epi != kEpiPlanarF32;
      if (staged) {
        if constexpr (BN >= 64) {
          // ---- bf16 output through swizzled smem + TMA store (full 128-byte lines, asynchronous)
          uint8_t* my_c = smem_c + (warp - 2) * 2 * 4096;
#pragma unroll 1
          for (int cb = 0; cb < BN / 64; ++cb) {
            if (cb * 64 >= P.N) break;
            if (lane == 0) tma_store_wait_read<1>();   // the buffer written two blocks ago has been read out
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
                const ulonglong2 s0 = *reinterpret_cast<const ulonglong2*>(s_scale + c0 + j);       // packed fp32 pairs
                const ulonglong2 s1 = *reinterpret_cast<const ulonglong2*>(s_scale + c0 + j + 4);
                const ulonglong2 t0 = *reinterpret_cast<const ulonglong2*>(shift + c0 + j);          // generic: smem, or global when straddling
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
        }
      } else {
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        if (c0 >= P.N) break;  // padded columns: nothing to store (uniform branch)
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr + c0, v);
        tmem_ld_wait();
        float y[32];
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 s = *reinterpret_cast<const float4*>(s_scale + c0 + j);
          const float4 t = *reinterpret_cast<const float4*>(shift + c0 + j);
          y[j + 0] = fmaf(__uint_as_float(v[j + 0]), s.x, t.x);
          y[j + 1] = fmaf(__uint_as_float(v[j + 1]), s.y, t.y);
          y[j + 2] = fmaf(__uint_as_float(v[j + 2]), s.z, t.z);
          y[j + 3] = fmaf(__uint_as_float(v[j + 3]), s.w, t.w);
        }
        if (P.relu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) y[j] = fmaxf(y[j], 0.0f);
        }
        if (P.epi == kEpiPlanarF32) {
          if (row_ok) {
            float* o = reinterpret_cast<float*>(P.out) + (static_cast<size_t>(img) * P.N + c0) * L.rows_per_img + pix;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c0 + j < P.N) o[static_cast<size_t>(j) * L.rows_per_img] = y[j];
          }
        } else if (row_ok && !(L.debug & 1)) {
          __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(P.out) + static_cast<size_t>(row) * P.ldo + P.col_off + c0;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            if (c0 + j < P.N) {
              uint4 w;
              w.x = pack_bf16x2(y[j + 0], y[j + 1]);
              w.y = pack_bf16x2(y[j + 2], y[j + 3]);
              w.z = pack_bf16x2(y[j + 4], y[j + 5]);
              w.w = pack_bf16x2(y[j + 6], y[j + 7]);
              stg_v4(o + j, w);
            }
          }
        }
      }
      }
      tcgen05_fence_before();
      mbar_arrive(&tmem_empty[acc]);
    }
    if (lane == 0) tma_store_wait_all<0>();   // all bulk stores of this warp have completed before smem goes away
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

}  // namespace dlv3p
