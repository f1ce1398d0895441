// tgemm.cuh — the training step's GEMM: D[M, N] = A[M, K] * B[N, K]^T on tcgen05 tensor cores, both operands bf16
// row-major in DEVICE memory with the contraction dimension contiguous ("NT"), fp32 accumulation in TMEM.
//
// One kernel serves the three GEMMs of every 1x1 convolution (DeeplabConv2D, reference deeplabv3p/models/layers.py:14-21)
// in the cfg-5 training step (train.py:143-169: MirroredStrategy fit = forward, backward, update):
//   forward   Y[pixels, N]  = X[pixels, K]   * Wnk[N, K]^T
//   dgrad     dX[pixels, K] = dY[pixels, N]  * Wkn[K, N]^T          (Wkn is the Keras HWIO kernel itself)
//   wgrad     dW[K, N]      = Xt[K, pixels]  * dYt[N, pixels]^T     (operands transposed by transpose_bf16_kernel;
//                                                                    contraction over pixels, split-K partials)
// Structure follows pw_gemm.cuh (warp 0 TMA producer, warp 1 MMA issuer, warps 2..5 epilogue; mbarrier ring, two TMEM
// accumulator stages, persistent over work items) with what inference did not need: N tiling, split-K over the
// contraction (fp32 partials [splits][M][N], reduced in fixed order by tgemm_reduce_kernel: deterministic), tensor maps
// passed as __grid_constant__ kernel parameters (no device allocations per call), edges handled by TMA zero fill.
#pragma once

#include <cuda.h>

#include "sm100_prims.cuh"

namespace dlv3p {

constexpr int kTgBM = 128;
constexpr int kTgBK = 64;
constexpr int kTgThreads = 192;

enum TgOut : int { kTgOutBf16 = 0, kTgOutF32 = 1, kTgOutPartial = 2 };

struct TgLaunch {
  CUtensorMap tmap_a;   // [M, K] bf16, box {64, 128}, SWIZZLE_128B
  CUtensorMap tmap_b;   // [N, K] bf16, box {64, min(BN, 128)}, SWIZZLE_128B
  CUtensorMap tmap_d;   // bf16 output [M, N] (row stride ldd), box {64, 32}, SWIZZLE_128B: the TMA-store epilogue (tma_store != 0)
  void* out;            // bf16 [M, ldd] | fp32 [M, ldd] | fp32 partials [splits][M][N]
  long long ldd;
  int M, N, K;
  int m_tiles, n_tiles, splits, kblocks, kb_per_split;
  int out_mode;
  int tma_store;        // bf16 output through swizzled smem + cp.async.bulk.tensor stores (full 128-byte lines, asynchronous)
};

template <int BN>
struct TgCfg {
  static constexpr int kStageBytes = kTgBM * 128 + BN * 128;
  static constexpr int kStages = (BN >= 256) ? 4 : 6;
  static constexpr int kTmemCols = (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr int kStoreBytes = 4 * 2 * 4096;   // 4 epilogue warps x 2 buffers x [32 rows x 128 B]
  static constexpr int kSmemBytes = kStages * kStageBytes + kStoreBytes + 256;
};

// TN = false: D = A[M,K] * B[N,K]^T (operands K-major).  TN = true: D[M,N] = A[Kc,M]^T * B[Kc,N] — both operands MN-major, read
// straight from the [pixels, channels] activations and activation gradients: the weight gradient without a transpose pass.
// (tensor maps then use boxes of {64 channels, 64 pixels}; a stage holds 2 + BN/64 boxes.)
template <int BN, bool TN = false>
__global__ void __launch_bounds__(kTgThreads, 1) tgemm_kernel(const __grid_constant__ TgLaunch L) {
  using Cfg = TgCfg<BN>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * (kTgBM * 128);
  uint8_t* smem_c = smem + kStages * Cfg::kStageBytes;     // epilogue staging (1024-byte aligned)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_c + Cfg::kStoreBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* tmem_full = bars + 2 * kStages;
  uint64_t* tmem_empty = bars + 2 * kStages + 2;
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles = L.m_tiles * L.n_tiles;
  const int total_items = tiles * L.splits;

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
    tma_prefetch_desc(&L.tmap_a);
    tma_prefetch_desc(&L.tmap_b);
    if (L.tma_store) tma_prefetch_desc(&L.tmap_d);
  }
  if (warp == 1) {
    tmem_alloc(tmem_base_ptr, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_ptr;

  // item -> (split, m tile, n tile): n fastest so that neighbouring CTAs share the A tile through L2
  if (warp == 0) {
    if (elect_one()) {
      uint32_t stage = 0, phase = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        const int split = item / tiles, t = item - split * tiles;
        const int mt = t / L.n_tiles, nt = t - mt * L.n_tiles;
        const int kb0 = split * L.kb_per_split, kb1 = min(L.kblocks, kb0 + L.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          if constexpr (TN) {
#pragma unroll
            for (int g = 0; g < kTgBM / 64; ++g)
              tma_load_2d(smem_a + stage * (kTgBM * 128) + g * 8192, &L.tmap_a, &full_bar[stage], mt * kTgBM + g * 64, kb * kTgBK, kEvictNormal);
#pragma unroll
            for (int g = 0; g < BN / 64; ++g)
              tma_load_2d(smem_b + stage * (BN * 128) + g * 8192, &L.tmap_b, &full_bar[stage], nt * BN + g * 64, kb * kTgBK, kEvictNormal);
          } else {
          tma_load_2d(smem_a + stage * (kTgBM * 128), &L.tmap_a, &full_bar[stage], kb * kTgBK, mt * kTgBM, kEvictNormal);
          tma_load_2d(smem_b + stage * (BN * 128), &L.tmap_b, &full_bar[stage], kb * kTgBK, nt * BN, kEvictNormal);
          if constexpr (BN > 128)
            tma_load_2d(smem_b + stage * (BN * 128) + 128 * 128, &L.tmap_b, &full_bar[stage], kb * kTgBK, nt * BN + 128, kEvictNormal);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = TN ? make_idesc_bf16_mn(kTgBM, BN) : make_idesc_bf16(kTgBM, BN);
    uint32_t stage = 0, phase = 0, it = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
      const int split = item / tiles;
      const int kb0 = split * L.kb_per_split, kb1 = min(L.kblocks, kb0 + L.kb_per_split);
      const uint32_t acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tcgen05_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BN;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tcgen05_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem_a + stage * (kTgBM * 128)), sb = smem_u32(smem_b + stage * (BN * 128));
          const uint64_t da = TN ? make_smem_desc_sw128_mn(sa, 8192) : make_smem_desc_sw128(sa);
          const uint64_t db = TN ? make_smem_desc_sw128_mn(sb, 8192) : make_smem_desc_sw128(sb);
          constexpr uint32_t kstep = TN ? 16 * 128 : 32;   // 16 contraction rows of 128 bytes | 16 elements inside the 128-byte row
#pragma unroll
          for (int k = 0; k < kTgBK / 16; ++k)
            umma_bf16_ss(tmem_d, smem_desc_advance(da, k * kstep), smem_desc_advance(db, k * kstep), idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (kb == kb1 - 1) umma_commit(&tmem_full[acc]);
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    const int q = warp & 3;
    uint32_t it = 0;
    uint32_t store_buf = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
      const int split = item / tiles, t = item - split * tiles;
      const int mt = t / L.n_tiles, nt = t - mt * L.n_tiles;
      const uint32_t acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tmem_full[acc], acc_phase);
      tcgen05_fence_after();
      const int row = mt * kTgBM + q * 32 + lane;
      const bool row_ok = row < L.M;
      const int n0 = nt * BN;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
      if (L.tma_store) {
        // bf16 rows staged in 128B-swizzled smem (16-byte chunk index XOR (row & 7)) and written by TMA: whole 128-byte lines,
        // rows / columns past M / N clipped by the tensor map; double buffered per warp
        uint8_t* my_c = smem_c + (warp - 2) * 2 * 4096;
#pragma unroll 1
        for (int cb = 0; cb < BN / 64; ++cb) {
          if (n0 + cb * 64 >= L.N) break;   // uniform
          if (lane == 0) tma_store_wait_read<1>();
          __syncwarp();
          const uint32_t cbuf = smem_u32(my_c + store_buf * 4096) + lane * 128;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(taddr + cb * 64 + half * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              const uint32_t chunk = static_cast<uint32_t>(half * 4 + (j >> 3)) ^ static_cast<uint32_t>(lane & 7);
              sts_v4(cbuf + chunk * 16, make_uint4(pack_bf16x2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), pack_bf16x2(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])),
                                                  pack_bf16x2(__uint_as_float(v[j + 4]), __uint_as_float(v[j + 5])), pack_bf16x2(__uint_as_float(v[j + 6]), __uint_as_float(v[j + 7]))));
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&L.tmap_d, my_c + store_buf * 4096, n0 + cb * 64, mt * kTgBM + q * 32);
            tma_store_commit();
          }
          store_buf ^= 1;
        }
      } else {
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        if (n0 + c0 >= L.N) break;   // uniform
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr + c0, v);
        tmem_ld_wait();
        const int ncol = n0 + c0;
        if (!row_ok) {
          // nothing to store for rows past M (the tcgen05.ld above stays warp-convergent)
        } else if (L.out_mode == kTgOutBf16) {
          __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(L.out) + static_cast<size_t>(row) * L.ldd + ncol;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            if (ncol + j + 8 <= L.N) {
              stg_v4(o + j, make_uint4(pack_bf16x2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), pack_bf16x2(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])),
                                      pack_bf16x2(__uint_as_float(v[j + 4]), __uint_as_float(v[j + 5])), pack_bf16x2(__uint_as_float(v[j + 6]), __uint_as_float(v[j + 7]))));
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e)
                if (ncol + j + e < L.N) o[j + e] = __float2bfloat16_rn(__uint_as_float(v[j + e]));
            }
          }
        } else {
          float* o = reinterpret_cast<float*>(L.out) + (L.out_mode == kTgOutPartial ? static_cast<size_t>(split) * L.M * L.ldd : 0) +
                     static_cast<size_t>(row) * L.ldd + ncol;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (ncol + j + 4 <= L.N) {
              *reinterpret_cast<float4*>(o + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (ncol + j + e < L.N) o[j + e] = __uint_as_float(v[j + e]);
            }
          }
        }
        __syncwarp();
      }
      }
      tcgen05_fence_before();
      mbar_arrive(&tmem_empty[acc]);
    }
    if (L.tma_store && lane == 0) tma_store_wait_all<0>();   // bulk stores complete before the smem goes away
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// out[m][n] = sum over splits of partial[s][m][n], fixed order (deterministic); fp32 or bf16 result with leading dim ldd
__global__ void __launch_bounds__(256) tgemm_reduce_kernel(const float* __restrict__ partial, int splits, long long M, int N, void* __restrict__ out,
                                                           long long ldd, int out_fp32) {
  const long long total = M * N;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total; idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    float s = 0.0f;
    for (int k = 0; k < splits; ++k) s += partial[static_cast<size_t>(k) * total + idx];
    const long long m = idx / N;
    const int n = static_cast<int>(idx - m * N);
    if (out_fp32) reinterpret_cast<float*>(out)[m * ldd + n] = s;
    else reinterpret_cast<__nv_bfloat16*>(out)[m * ldd + n] = __float2bfloat16_rn(s);
  }
}

// out[c][r] = in[r][c]: bf16 [R, C] (row stride ld_in) -> [C, R] (row stride ld_out); R, C even.  64 x 64 tiles through smem.
__global__ void __launch_bounds__(256) transpose_bf16_kernel(const __nv_bfloat16* __restrict__ in, long long R, int C, long long ld_in,
                                                             __nv_bfloat16* __restrict__ out, long long ld_out) {
  __shared__ uint16_t s[64][66];
  const long long r0 = static_cast<long long>(blockIdx.x) * 64;
  const int c0 = blockIdx.y * 64;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 64; i += 8) {
    const long long r = r0 + i;
    const int c = c0 + 2 * tx;
    uint32_t v = 0u;
    if (r < R && c < C) v = *reinterpret_cast<const uint32_t*>(in + r * ld_in + c);
    s[i][2 * tx] = static_cast<uint16_t>(v & 0xFFFFu);
    s[i][2 * tx + 1] = static_cast<uint16_t>(v >> 16);
  }
  __syncthreads();
  for (int i = ty; i < 64; i += 8) {
    const int c = c0 + i;
    const long long r = r0 + 2 * tx;
    if (c < C && r < R) {
      const uint32_t v = static_cast<uint32_t>(s[2 * tx][i]) | (static_cast<uint32_t>(s[2 * tx + 1][i]) << 16);
      *reinterpret_cast<uint32_t*>(out + static_cast<long long>(c) * ld_out + r) = v;
    }
  }
}

}  // namespace dlv3p
