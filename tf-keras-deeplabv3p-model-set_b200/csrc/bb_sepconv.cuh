// bb_sepconv.cuh — fused SepConv_BN of the Xception ENTRY flow on tcgen05 CTA pairs: the machine of dwpw_gemm2.cuh (TMA halo tiles ->
// 16 stencil warps write the depthwise + BN result as the swizzled A operand -> tcgen05.mma.cta_group::2 against pointwise
// weights resident in shared memory -> constant-bank BN epilogue -> TMA store) generalised to what the backbone needs:
//   [ReLU] -> depthwise 3x3 stride 1 'same' -> BN -> [ReLU] -> pointwise 1x1 (C -> NB, NB = 128 | 256) -> BN -> [ReLU]
// (reference SepConv_BN deeplabv3p/models/layers.py:74-111 with depth_activation=False: ReLU BEFORE the depthwise conv, none after
//  the BatchNorms; call sites deeplabv3p_xception.py:131-138, entry_flow_block1 / block2 separable_conv1 / 2).
// The entry flow's tensors (up to 537 MB at batch 32) do not fit the L2, so unfused every depthwise output makes a round trip
// through HBM; here it never leaves the SM: e.g. entry_flow_block1_separable_conv2 moves 1.07 GB instead of 2.15 GB.
#pragma once

#include <cuda.h>

#include "dwpw_gemm.cuh"
#include "sm100_prims.cuh"

namespace dlv3p {

constexpr int kSepThreads = 22 * 32;   // producer, MMA, 4 epilogue, 16 stencil warps

template <int KB, int NB>
struct BbSepCfg {
  static constexpr int kTapBytes = KB * 10 * 64 * 4;                  // [KB][9 taps + shift][64 channels] fp32
  static constexpr int kWHalfBlockBytes = (NB / 2) * 128;             // this CTA's NB / 2 output channels x 64 input channels
  static constexpr int kWBytes = KB * kWHalfBlockBytes;
  static constexpr int kABytes = 2 * kDwAStageBytes;                  // one A stage per stencil group
  static constexpr int kInBytes = 4 * kDwInStageBytes;                // two halo stages per stencil group
  static constexpr int kInBytesPad = (kInBytes + 1023) / 1024 * 1024;
  static constexpr int kStoreBytes = 4 * 2048;                        // 4 epilogue warps x [32 rows x 64 B], 64B swizzle
  static constexpr int kSmemBytes = kWBytes + kABytes + kInBytesPad + kStoreBytes + kTapBytes + 256 /*barriers*/;
};

// The pointwise BN scale / shift travel in the kernel parameters (constant bank): the epilogue reads them as immediate
// constant operands instead of 128 shared-memory loads per thread and tile.
struct BbSepParams {
  DwPwParams base;                 // tmap_w: 2D [NB, KB*64] bf16 K-major, box {64, NB / 2}
  const CUtensorMap* tmap_out32;   // 4D {NB, W, H, B} bf16, box {32, 16, 2, 1}, SWIZZLE_64B
  float scale_c[256];
  float shift_c[256];
};

// kAct = SepConv_BN's depth_activation (layers.py:98-109): false -> ReLU BEFORE the depthwise conv and none after the BatchNorms (every
// fused layer of the entry flow); true -> ReLU after both BatchNorms.  Compile time: the stencil warps are the bottleneck.
template <int KB, int NB, bool kAct>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kSepThreads, 1) bb_sepconv_kernel(const __grid_constant__ BbSepParams Q) {
  using Cfg = BbSepCfg<KB, NB>;
  const DwPwParams& P = Q.base;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  uint8_t* smem_w = smem;                          // KB x [128 rows x 128 B], swizzled: this CTA's half of the weights
  uint8_t* smem_a = smem_w + Cfg::kWBytes;         // 2 x [128 rows x 128 B], swizzled (stage p = stencil group p)
  uint8_t* smem_in = smem_a + Cfg::kABytes;        // 4 x [10][18][64] bf16 (stages 2p, 2p+1 = stencil group p)
  uint8_t* smem_c = smem_in + Cfg::kInBytesPad;    // epilogue store staging, 4 warps x 2 KB
  float* s_taps = reinterpret_cast<float*>(smem_c + Cfg::kStoreBytes);   // depthwise taps + shift of every K block
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_taps + KB * 10 * 64);
  uint64_t* w_full = bars;                   // [1]
  uint64_t* in_full = bars + 1;              // [4]
  uint64_t* in_empty = in_full + 4;          // [4]
  uint64_t* a_full = in_empty + 4;           // [2]
  uint64_t* a_empty = a_full + 2;            // [2]
  uint64_t* tmem_full = a_empty + 2;         // [2]
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
    for (int i = 0; i < 4; ++i) {
      mbar_init(&in_full[i], 1);
      mbar_init(&in_empty[i], 8);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], 16);
      mbar_init(&a_empty[i], 1);
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 256);
    }
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < KB * 10 * 64; i += kSepThreads) {
    const int kb = i / 640, t = (i - kb * 640) >> 6, ch = kb * 64 + (i & 63);
    s_taps[i] = t < 9 ? __ldg(P.dw_w + t * (KB * 64) + ch) : __ldg(P.dw_shift + ch);
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
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (elect_one()) {
      if (leader) mbar_arrive_expect_tx(w_full, 2 * Cfg::kWBytes);
      for (int kb = 0; kb < KB; ++kb)   // rows [rank*128, rank*128+128) of the [256, K] weight matrix; bytes credited to the leader
        tma_load_2d_2sm(smem_w + kb * Cfg::kWHalfBlockBytes, P.tmap_w, w_full, kb * 64, static_cast<int>(rank) * (NB / 2), kEvictLast);
      uint32_t c = 0;
      for (int item = cluster_id; item < num_items; item += num_clusters) {
        const int tile = item * 2 + static_cast<int>(rank);   // tile == num_tiles (odd count): image B is out of bounds -> zero fill
        const int b = tile / tiles_per_img;
        const int t2 = tile - b * tiles_per_img;
        const int ty = t2 / P.tiles_x;
        const int tx = t2 - ty * P.tiles_x;
        for (int kb = 0; kb < KB; ++kb, ++c) {
          const uint32_t si = (c & 1) * 2 + ((c >> 1) & 1);   // group c&1, its stage (c>>1)&1
          const uint32_t ph = (c >> 2) & 1;
          mbar_wait(&in_empty[si], ph ^ 1);
          mbar_arrive_expect_tx(&in_full[si], kDwInStageBytes);
          const bool second = P.tmap_x2 != nullptr && kb >= P.kb_split;
          tma_load_4d(smem_in + si * kDwInStageBytes, second ? P.tmap_x2 : P.tmap_x, &in_full[si], (second ? kb - P.kb_split : kb) * 64,
                      tx * kDwTW - 1, ty * kDwTH - 1, b, kEvictNormal);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (leader) {
      constexpr uint32_t idesc = make_idesc_bf16(256, NB);
      mbar_wait(w_full, 0);
      uint32_t c = 0, it = 0;
      for (int item = cluster_id; item < num_items; item += num_clusters, ++it) {
        const uint32_t acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + acc * NB;
        for (int kb = 0; kb < KB; ++kb, ++c) {
          const uint32_t sa = c & 1;
          const uint32_t ph = (c >> 1) & 1;
          mbar_wait(&a_full[sa], ph);
          tcgen05_fence_after();
          if (elect_one()) {
            const uint64_t da = make_smem_desc_sw128(smem_u32(smem_a + sa * kDwAStageBytes));
            const uint64_t db = make_smem_desc_sw128(smem_u32(smem_w + kb * Cfg::kWHalfBlockBytes));
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16_ss_2sm(tmem_d, smem_desc_advance(da, k * 32), smem_desc_advance(db, k * 32), idesc, (kb > 0 || k > 0) ? 1u : 0u);
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
    uint8_t* my_c = smem_c + (warp - 2) * 2048;
    const uint32_t cbuf = smem_u32(my_c) + lane * 64;
    const uint32_t rsw = static_cast<uint32_t>(lane >> 1) & 3u;   // 64B swizzle: 16-byte chunk index ^= (row / 2) % 4
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
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * NB;
      // bf16 output through 64B-swizzled smem + TMA store: box = this warp's 2 tile rows x 16 pixels x 32 channels;
      // partial (and phantom) tiles are clipped by the TMA unit
#pragma unroll
      for (int cb = 0; cb < NB / 32; ++cb) {
        const int c0 = cb * 32;
        uint32_t v[32], pk[16];
        tmem_ld_32x32b_x32(taddr + c0, v);
        tmem_ld_wait();
        if (cb == NB / 32 - 1) {   // every column of this accumulator stage has been read: hand it back early
          tcgen05_fence_before();
          if (leader) mbar_arrive(&tmem_empty[acc]);
          else mbar_arrive_cluster(&tmem_empty[acc], 0);
        }
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const unsigned long long s01 = pack_f32x2(Q.scale_c[c0 + j], Q.scale_c[c0 + j + 1]), s23 = pack_f32x2(Q.scale_c[c0 + j + 2], Q.scale_c[c0 + j + 3]);
          const unsigned long long t01 = pack_f32x2(Q.shift_c[c0 + j], Q.shift_c[c0 + j + 1]), t23 = pack_f32x2(Q.shift_c[c0 + j + 2], Q.shift_c[c0 + j + 3]);
          const unsigned long long y01 = f32x2_fma(f32x2_make(v[j + 0], v[j + 1]), s01, t01), y23 = f32x2_fma(f32x2_make(v[j + 2], v[j + 3]), s23, t23);
          pk[j / 2 + 0] = kAct ? f32x2_to_bf16x2_relu(y01) : f32x2_to_bf16x2(y01);
          pk[j / 2 + 1] = kAct ? f32x2_to_bf16x2_relu(y23) : f32x2_to_bf16x2(y23);
        }
        if (lane == 0) tma_store_wait_read<0>();   // the previous block has been read out of the staging buffer
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; ++j)
          sts_v4(cbuf + ((static_cast<uint32_t>(j) ^ rsw) << 4), make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]));
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                           reinterpret_cast<uint64_t>(Q.tmap_out32)),
                       "r"(smem_u32(my_c)), "r"(c0), "r"(tx * kDwTW), "r"(ty * kDwTH + 2 * q), "r"(b)
                       : "memory");
          tma_store_commit();
        }
      }
    }
    if (lane == 0) tma_store_wait_all<0>();
  } else {
    // ------------------------------------------------------------------ depthwise stencil warps (both CTAs, own tile)
    // warp = a 4 x 4 block of output pixels (row half g, column block cbk), lane = one channel PAIR of the 64-channel K
    // block.  Every warp-wide load / store is one 128-byte pixel row: a single conflict-free shared-memory wavefront.
    // Per K block a thread reads its 6 x 6 halo window once (rolling over the rows, next row's loads in flight during
    // this row's FMAs).
    const int grp = (warp - 6) >> 3;        // stencil group: K blocks c with (c & 1) == grp
    const int sw = (warp - 6) & 7;
    const int g = sw >> 2;                  // output rows 4g .. 4g+3
    const int cbk = sw & 3;                 // output cols 4cbk .. 4cbk+3
    const uint32_t in_base = smem_u32(smem_in) + ((4 * g) * kDwHaloW + 4 * cbk) * 128 + lane * 4;
    const uint32_t a_addr = smem_u32(smem_a) + grp * kDwAStageBytes + (lane & 3) * 4;
    const uint32_t jchunk = static_cast<uint32_t>(lane >> 2);
    int my_items = 0;
    for (int item = cluster_id; item < num_items; item += num_clusters) ++my_items;
    const uint32_t total_c = static_cast<uint32_t>(my_items) * KB;
    for (uint32_t c = grp; c < total_c; c += 2) {
      const uint32_t si = grp * 2 + ((c >> 1) & 1);
      // taps + shift of this lane's channel pair: 10 x LDS.64, 256 contiguous bytes per warp
      unsigned long long wt[9], sh;
      {
        const float* tp = s_taps + (c % KB) * 640 + lane * 2;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const float2 w = *reinterpret_cast<const float2*>(tp + t * 64);
          wt[t] = pack_f32x2(w.x, w.y);
        }
        const float2 w = *reinterpret_cast<const float2*>(tp + 9 * 64);
        sh = pack_f32x2(w.x, w.y);
      }
      mbar_wait(&in_full[si], (c >> 2) & 1);
      const uint32_t in_addr = in_base + si * kDwInStageBytes;
      unsigned long long acc[3][4];
      uint32_t raw_next[6];
#pragma unroll
      for (int cc = 0; cc < 6; ++cc) raw_next[cc] = lds_u32(in_addr + cc * 128);
#pragma unroll
      for (int r = 0; r < 6; ++r) {
        unsigned long long x[6];
#pragma unroll
        for (int cc = 0; cc < 6; ++cc) x[cc] = bf16x2_to_f32x2(kAct ? raw_next[cc] : relu_bf16x2(raw_next[cc]));
        if (r + 1 < 6) {
#pragma unroll
          for (int cc = 0; cc < 6; ++cc) raw_next[cc] = lds_u32(in_addr + ((r + 1) * kDwHaloW + cc) * 128);
        }
        if (r < 4) {
#pragma unroll
          for (int oc = 0; oc < 4; ++oc) acc[r % 3][oc] = sh;
        }
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
          const int orow = r - dy;
          if (orow < 0 || orow >= 4) continue;
#pragma unroll
          for (int oc = 0; oc < 4; ++oc)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) ffma2(acc[orow % 3][oc], wt[dy * 3 + dx], x[oc + dx]);
        }
        if (r >= 2) {  // output row r-2 of this block is complete
          const int orow = r - 2;
          if (orow == 0) mbar_wait(&a_empty[grp], ((c >> 1) & 1) ^ 1);  // the MMA is done with this group's A stage
#pragma unroll
          for (int oc = 0; oc < 4; ++oc) {
            const uint32_t m = static_cast<uint32_t>((4 * g + orow) * kDwTW + 4 * cbk + oc);
            sts_u32(a_addr + m * 128 + ((jchunk ^ (m & 7u)) << 4), kAct ? f32x2_to_bf16x2_relu(acc[orow % 3][oc]) : f32x2_to_bf16x2(acc[orow % 3][oc]));
          }
        }
      }
      fence_proxy_async_smem();                    // make this thread's A rows visible to the tensor-core (async) proxy
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&in_empty[si]);                // the warp is done reading the halo tile
        if (leader) mbar_arrive(&a_full[grp]);
        else mbar_arrive_cluster(&a_full[grp], 0);
      }
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
