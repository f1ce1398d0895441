// bb_sepwide.cuh — fused SepConv_BN of the Xception MIDDLE flow (728 -> 728 channels) on a cluster of two SMs:
//   [ReLU] -> depthwise 3x3 stride 1 'same' -> BN -> pointwise 1x1 (C -> N, 656 < N <= 768) -> BN [-> ReLU] [+ residual]
// (reference SepConv_BN deeplabv3p/models/layers.py:74-111 with depth_activation=False; call sites deeplabv3p_xception.py:139-143, the 48
//  separable convolutions of the 16 middle_flow_unit_k blocks, and exit_flow_block1_separable_conv1; the residual is _xception_block's
//  skip_connection_type 'sum', :88-90).
//
// Why a different machine than bb_sepconv.cuh: 728 fp32 accumulator columns do not fit the 512 TMEM columns of one SM, so an M tile
// (8 x 16 pixels = 128 rows) is owned by a CLUSTER OF TWO CTAs that split N: CTA 0 accumulates output channels [0, 384), CTA 1
// [384, N), each with its own tcgen05.mma.cta_group::1 stream over its own half of the pointwise weights (streamed by TMA, 2 stages of
// 384 rows x 64 K).  Both need the SAME A operand (the depthwise + BN result of the tile), and the stencil that makes it is the
// expensive part (CUDA cores, ~2000 clk per 128 px x 64 ch block), so it is computed ONCE per cluster: K block c is produced by CTA
// (c & 1), stencil group ((c >> 1) & 1), and written into A stage (c & 3) of BOTH CTAs — locally with st.shared, into the peer with
// st.shared::cluster (distributed shared memory).  Unfused, the depthwise output (47.7 MB at batch 32) made a round trip through L2 / HBM
// and the depthwise kernel (29 us, issue bound) ran while the tensor cores idled; here the stencil runs under the MMAs.
//
//   w_full[2] / w_empty[2]   (own CTA)    TMA bytes of a weight stage / tcgen05.commit of the MMAs that read it
//   in_full[2] / in_empty[2] (own CTA)    halo tile of stencil group g landed / its 8 warps are done reading it
//   a_full[4]   (BOTH CTAs)  8 arrivals : the 8 warps of the producing group, release.cluster, after their local + remote stores
//   a_empty[4]  (BOTH CTAs)  2 arrivals : tcgen05.commit multicast from the MMA threads of both CTAs (the producer waits on its own copy)
//   tmem_full / tmem_empty   (own CTA)    accumulators complete / 128 epilogue threads have read them (one accumulator stage: 384 of 512 columns)
#pragma once

#include <cuda.h>

#include "dwpw_gemm.cuh"
#include "sm100_prims.cuh"

namespace dlv3p {

constexpr int kWideThreads = 23 * 32;   // weight producer, MMA, 4 epilogue, 16 stencil, halo producer
constexpr int kWideN0 = 384;            // CTA 0: output channels [0, 384); CTA 1: [384, N)
constexpr int kWideWStageBytes = kWideN0 * 128;
constexpr int kWideOffA = 2 * kWideWStageBytes;                  // 4 A stages of [128 rows x 128 B], swizzled
constexpr int kWideOffIn = kWideOffA + 4 * kDwAStageBytes;       // 2 halo stages [10][18][64] bf16 (one per stencil group)
constexpr int kWideOffC = kWideOffIn + 2 * kDwInStageBytes;      // epilogue store staging, 4 warps x 2 KB (512-byte aligned: 64B swizzle)
constexpr int kWideOffBn = kWideOffC + 4 * 2048;                 // pointwise BN scale[384], shift[384] of this CTA's channels
constexpr int kWideOffBar = kWideOffBn + 2 * kWideN0 * 4;
constexpr int kWideSmemBytes = kWideOffBar + 256;
static_assert(kWideOffC % 512 == 0, "store staging must be 512-byte aligned");
static_assert(kWideSmemBytes <= 227 * 1024, "shared memory budget");

struct BbWideParams {
  const CUtensorMap* tmap_x;    // 4D {C, W, H, B} bf16, box {64, 18, 10, 1}, no swizzle (channels / pixels out of bounds read as zero)
  const CUtensorMap* tmap_w;    // 2D [N, K] bf16 K-major, box {64, 128}, SWIZZLE_128B
  const CUtensorMap* tmap_out;  // 4D {N, W, H, B} bf16, box {32, 16, 2, 1}, SWIZZLE_64B
  const float* dw_w;            // [9][KB*64] fp32 depthwise taps with the BN scale folded in, zero padded
  const float* dw_shift;        // [KB*64]    fp32 depthwise BN shift, zero padded
  const float* scale;           // [768] pointwise BN scale, zero padded
  const float* shift;           // [768]
  const __nv_bfloat16* res;     // optional residual [B, H, W, N], added after the BatchNorm (fp32, one rounding)
  int B, H, W, N, KB;
  int tiles_x, tiles_y, num_tiles;
  int relu_out;
  int debug;                    // benchmark aid: bit0 skip the output stores, bit1 skip the stencil math, bit2 skip the MMAs
};

__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void st_cluster_u32(uint32_t cluster_addr, uint32_t v) {
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(cluster_addr), "r"(v) : "memory");
}
// arrive (release at cluster scope) on a barrier given by its shared::cluster address
__device__ __forceinline__ void mbar_arrive_release_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_acq_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_acq_cluster(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait_acq_cluster(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait_acq_cluster(bar, parity)) {
    if (clock64() - t0 > DLV3P_MBAR_TIMEOUT_CYCLES) {
      printf("dlv3p: mbarrier (cluster) timeout block %d thread %d bar %u parity %u\n", blockIdx.x, threadIdx.x, smem_u32(bar), parity);
      __trap();
    }
  }
}
// all state spaces: generic-proxy writes (local and distributed shared memory) -> async proxy (tcgen05 operand reads)
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
// commit of a cta_group::1 MMA stream that arrives on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_both(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(static_cast<uint16_t>(3))
               : "memory");
}

template <bool kReluIn>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kWideThreads, 1) bb_sepwide_kernel(const __grid_constant__ BbWideParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  uint8_t* smem_w = smem;                      // 2 x [384 rows x 128 B], swizzled: this CTA's output channels x 64 input channels
  uint8_t* smem_a = smem + kWideOffA;          // 4 x [128 rows x 128 B], swizzled; stage s is written by CTA (s & 1), group (s >> 1)
  uint8_t* smem_in = smem + kWideOffIn;        // 2 x [10][18][64] bf16 (stage g = stencil group g of this CTA)
  uint8_t* smem_c = smem + kWideOffC;
  float* s_bn = reinterpret_cast<float*>(smem + kWideOffBn);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kWideOffBar);
  uint64_t* w_full = bars;             // [2]
  uint64_t* w_empty = bars + 2;        // [2]
  uint64_t* in_full = bars + 4;        // [2]
  uint64_t* in_empty = bars + 6;       // [2]
  uint64_t* a_full = bars + 8;         // [4]
  uint64_t* a_empty = bars + 12;       // [4]
  uint64_t* tmem_full = bars + 16;     // [1]
  uint64_t* tmem_empty = bars + 17;    // [1]
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(bars + 18);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const uint32_t peer = rank ^ 1u;
  const int KB = P.KB;
  const int n_base = static_cast<int>(rank) * kWideN0;
  const int n_mine = rank == 0 ? kWideN0 : (P.N - kWideN0 + 31) / 32 * 32;    // accumulator columns of this CTA (N = 728: 384 | 352)
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int tiles_per_img = P.tiles_x * P.tiles_y;
  int my_items = 0;
  for (int item = cluster_id; item < P.num_tiles; item += num_clusters) ++my_items;
  const uint32_t total_c = static_cast<uint32_t>(my_items) * KB;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 1);
      mbar_init(&in_full[i], 1);
      mbar_init(&in_empty[i], 8);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&a_full[i], 8);
      mbar_init(&a_empty[i], 2);
    }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, 128);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 2 * kWideN0; i += kWideThreads) {
    const int ch = n_base + (i < kWideN0 ? i : i - kWideN0);
    s_bn[i] = ch < 768 ? __ldg((i < kWideN0 ? P.scale : P.shift) + ch) : 0.0f;
  }
  cluster_sync_all();                       // barriers of both CTAs exist before anyone signals across the pair
  if (warp == 1) {
    tmem_alloc(tmem_base_ptr, 512);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_ptr;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------------------------ weight producer: every K block, this CTA's 384 rows
    if (elect_one()) {
      for (uint32_t c = 0; c < total_c; ++c) {
        const uint32_t wi = c & 1;
        const int kb = static_cast<int>(c % static_cast<uint32_t>(KB));
        mbar_wait(&w_empty[wi], ((c >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&w_full[wi], kWideWStageBytes);
#pragma unroll
        for (int j = 0; j < 3; ++j)   // rows past N (CTA 1: 728 .. 767) are zero filled
          tma_load_2d(smem_w + wi * kWideWStageBytes + j * 16384, P.tmap_w, &w_full[wi], kb * 64, n_base + j * 128, kEvictLast);
      }
    }
  } else if (warp == 22) {
    // ------------------------------------------------------------------ halo producer: the K blocks this CTA's stencil groups compute
    if (elect_one()) {
      for (uint32_t c = rank; c < total_c; c += 2) {
        const uint32_t g = (c >> 1) & 1, j = c >> 2;
        const int it = static_cast<int>(c / static_cast<uint32_t>(KB));
        const int kb = static_cast<int>(c) - it * KB;
        const int tile = cluster_id + it * num_clusters;
        const int b = tile / tiles_per_img;
        const int t2 = tile - b * tiles_per_img;
        const int ty = t2 / P.tiles_x;
        const int tx = t2 - ty * P.tiles_x;
        mbar_wait(&in_empty[g], (j & 1) ^ 1);
        mbar_arrive_expect_tx(&in_full[g], kDwInStageBytes);
        tma_load_4d(smem_in + g * kDwInStageBytes, P.tmap_x, &in_full[g], kb * 64, tx * kDwTW - 1, ty * kDwTH - 1, b, kEvictNormal);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer: M 128 x (256 + n_mine - 256) x 64 per K block
    const uint32_t idesc_lo = make_idesc_bf16(128, 256);
    const uint32_t idesc_hi = make_idesc_bf16(128, n_mine - 256);
    uint32_t c = 0;
    for (int it = 0; it < my_items; ++it) {
      mbar_wait(tmem_empty, (static_cast<uint32_t>(it) & 1) ^ 1);   // the epilogue has drained the previous tile
      tcgen05_fence_after();
      for (int kb = 0; kb < KB; ++kb, ++c) {
        const uint32_t s = c & 3, wi = c & 1;
        mbar_wait_acq_cluster(&a_full[s], (c >> 2) & 1);
        mbar_wait(&w_full[wi], (c >> 1) & 1);
        tcgen05_fence_after();
        if (elect_one()) {
          fence_proxy_async_all();
          if (!(P.debug & 4)) {
            const uint64_t da = make_smem_desc_sw128(smem_u32(smem_a + s * kDwAStageBytes));
            const uint64_t db = make_smem_desc_sw128(smem_u32(smem_w + wi * kWideWStageBytes));
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              umma_bf16_ss(tmem_base, smem_desc_advance(da, k * 32), smem_desc_advance(db, k * 32), idesc_lo, (kb > 0 || k > 0) ? 1u : 0u);
              umma_bf16_ss(tmem_base + 256, smem_desc_advance(da, k * 32), smem_desc_advance(db, 256 * 128 + k * 32), idesc_hi, (kb > 0 || k > 0) ? 1u : 0u);
            }
          }
          umma_commit(&w_empty[wi]);
          umma_commit_both(&a_empty[s]);
          if (kb == KB - 1) umma_commit(tmem_full);
        }
        __syncwarp();
      }
    }
  } else if (warp < 6) {
    // ------------------------------------------------------------------ epilogue: this CTA's channels of the tile
    const int q = warp & 3;
    uint8_t* my_c = smem_c + (warp - 2) * 2048;
    const uint32_t cbuf = smem_u32(my_c) + lane * 64;
    const uint32_t rsw = static_cast<uint32_t>(lane >> 1) & 3u;   // 64B swizzle: 16-byte chunk index ^= (row / 2) % 4
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const int n_blocks = n_mine / 32;
    const int mrow = q * 32 + lane;                               // A row = pixel (mrow / 16, mrow % 16) of the tile
    for (int it = 0; it < my_items; ++it) {
      const int tile = cluster_id + it * num_clusters;
      const int b = tile / tiles_per_img;
      const int t2 = tile - b * tiles_per_img;
      const int ty = t2 / P.tiles_x;
      const int tx = t2 - ty * P.tiles_x;
      const int py = ty * kDwTH + (mrow >> 4), px = tx * kDwTW + (mrow & 15);
      const bool has_res = P.res != nullptr && py < P.H && px < P.W;
      const __nv_bfloat16* rp = P.res + ((static_cast<size_t>(b) * P.H + py) * P.W + px) * P.N + n_base;
      mbar_wait(tmem_full, static_cast<uint32_t>(it) & 1);
      tcgen05_fence_after();
      for (int cb = 0; cb < n_blocks; ++cb) {
        const int c0 = cb * 32;
        uint4 rr[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) rr[j] = (has_res && n_base + c0 + j * 8 < P.N) ? ldg_nc_v4(rp + c0 + j * 8) : make_uint4(0u, 0u, 0u, 0u);
        uint32_t v[32], pk[16];
        tmem_ld_32x32b_x32(taddr + c0, v);
        tmem_ld_wait();
        if (cb == n_blocks - 1) {   // every column of the accumulator has been read: hand it back
          tcgen05_fence_before();
          mbar_arrive(tmem_empty);
        }
        const uint32_t* rw = reinterpret_cast<const uint32_t*>(rr);
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 s4 = *reinterpret_cast<const float4*>(s_bn + c0 + j);
          const float4 t4 = *reinterpret_cast<const float4*>(s_bn + kWideN0 + c0 + j);
          float y0 = fmaf(__uint_as_float(v[j + 0]), s4.x, t4.x), y1 = fmaf(__uint_as_float(v[j + 1]), s4.y, t4.y);
          float y2 = fmaf(__uint_as_float(v[j + 2]), s4.z, t4.z), y3 = fmaf(__uint_as_float(v[j + 3]), s4.w, t4.w);
          if (P.relu_out) { y0 = fmaxf(y0, 0.0f); y1 = fmaxf(y1, 0.0f); y2 = fmaxf(y2, 0.0f); y3 = fmaxf(y3, 0.0f); }
          const uint32_t r01 = rw[j / 2], r23 = rw[j / 2 + 1];
          pk[j / 2 + 0] = pack_bf16x2(y0 + bf16_lo(r01), y1 + bf16_hi(r01));
          pk[j / 2 + 1] = pack_bf16x2(y2 + bf16_lo(r23), y3 + bf16_hi(r23));
        }
        if (lane == 0) tma_store_wait_read<0>();   // the previous block has been read out of the staging buffer
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; ++j)
          sts_v4(cbuf + ((static_cast<uint32_t>(j) ^ rsw) << 4), make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]));
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0 && !(P.debug & 1)) {   // box = this warp's 2 tile rows x 16 pixels x 32 channels; clipped at N and at the image border
          asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                           reinterpret_cast<uint64_t>(P.tmap_out)),
                       "r"(smem_u32(my_c)), "r"(n_base + c0), "r"(tx * kDwTW), "r"(ty * kDwTH + 2 * q), "r"(b)
                       : "memory");
          tma_store_commit();
        }
      }
    }
    if (lane == 0) tma_store_wait_all<0>();
  } else {
    // ------------------------------------------------------------------ depthwise stencil warps
    // warp = a 4 x 4 block of output pixels (row half g, column block cbk), lane = one channel PAIR of the 64-channel K block (the
    // stencil of dwpw_gemm2.cuh).  Group grp of CTA rank owns A stage s = 2 grp + rank in both CTAs and K blocks c = s (mod 4).
    const int grp = (warp - 6) >> 3;
    const int sw = (warp - 6) & 7;
    const int g = sw >> 2;                  // output rows 4g .. 4g+3
    const int cbk = sw & 3;                 // output cols 4cbk .. 4cbk+3
    const uint32_t s = static_cast<uint32_t>(2 * grp) + rank;
    const uint32_t in_addr = smem_u32(smem_in) + grp * kDwInStageBytes + ((4 * g) * kDwHaloW + 4 * cbk) * 128 + lane * 4;
    const uint32_t a_local = smem_u32(smem_a) + s * kDwAStageBytes + (lane & 3) * 4;
    const uint32_t a_remote = mapa_u32(a_local, peer);
    const uint32_t full_local = mapa_u32(smem_u32(&a_full[s]), rank), full_remote = mapa_u32(smem_u32(&a_full[s]), peer);
    const uint32_t jchunk = static_cast<uint32_t>(lane >> 2);
    const int Cpad = KB * 64;
    for (uint32_t c = s; c < total_c; c += 4) {
      const uint32_t j = c >> 2;
      const int kb = static_cast<int>(c % static_cast<uint32_t>(KB));
      // taps + shift of this lane's channel pair (weights: L1 / L2 resident, independent of the previous kernel)
      unsigned long long wt[9], sh;
      {
        const float* tp = P.dw_w + kb * 64 + lane * 2;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const float2 w = __ldg(reinterpret_cast<const float2*>(tp + t * Cpad));
          wt[t] = pack_f32x2(w.x, w.y);
        }
        const float2 w = __ldg(reinterpret_cast<const float2*>(P.dw_shift + kb * 64 + lane * 2));
        sh = pack_f32x2(w.x, w.y);
      }
      mbar_wait(&in_full[grp], j & 1);
      unsigned long long acc[3][4];
      uint32_t raw_next[6];
#pragma unroll
      for (int cc = 0; cc < 6; ++cc) raw_next[cc] = lds_u32(in_addr + cc * 128);
#pragma unroll
      for (int r = 0; r < 6; ++r) {
        unsigned long long x[6];
#pragma unroll
        for (int cc = 0; cc < 6; ++cc) x[cc] = bf16x2_to_f32x2(kReluIn ? relu_bf16x2(raw_next[cc]) : raw_next[cc]);
        if (r + 1 < 6) {
#pragma unroll
          for (int cc = 0; cc < 6; ++cc) raw_next[cc] = lds_u32(in_addr + ((r + 1) * kDwHaloW + cc) * 128);
        }
        if (r < 4) {
#pragma unroll
          for (int oc = 0; oc < 4; ++oc) acc[r % 3][oc] = sh;
        }
        if (!(P.debug & 2)) {
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) {
            const int orow = r - dy;
            if (orow < 0 || orow >= 4) continue;
#pragma unroll
            for (int oc = 0; oc < 4; ++oc)
#pragma unroll
              for (int dx = 0; dx < 3; ++dx) ffma2(acc[orow % 3][oc], wt[dy * 3 + dx], x[oc + dx]);
          }
        }
        if (r >= 2) {  // output row r-2 of this block is complete
          const int orow = r - 2;
          if (orow == 0) mbar_wait(&a_empty[s], (j & 1) ^ 1);  // the MMAs of BOTH CTAs are done with this A stage
#pragma unroll
          for (int oc = 0; oc < 4; ++oc) {
            const uint32_t m = static_cast<uint32_t>((4 * g + orow) * kDwTW + 4 * cbk + oc);
            const uint32_t off = m * 128 + ((jchunk ^ (m & 7u)) << 4);
            const uint32_t val = f32x2_to_bf16x2(acc[orow % 3][oc]);
            sts_u32(a_local + off, val);
            st_cluster_u32(a_remote + off, val);
          }
        }
      }
      fence_proxy_async_all();                     // this thread's A rows (both copies) -> visible to the tensor cores' async proxy
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&in_empty[grp]);               // the warp is done reading the halo tile
        mbar_arrive_release_cluster(full_local);
        mbar_arrive_release_cluster(full_remote);
      }
    }
  }

  tcgen05_fence_before();
  cluster_sync_all();                      // the peer may still be writing this CTA's A stages / signalling its barriers
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace dlv3p
