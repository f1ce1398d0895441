// aspp_dw_gather.cuh — ASPP depthwise kernel for LARGE feature maps (Cityscapes OS8: 128 x 256 map, rates 12/24/36),
// where neither the whole map (aspp_dw_fast.cuh) nor a halo tile (the halo is the dilation) fits in shared memory.
//
// Replaces `aspp{1,2,3}_depthwise` + `_depthwise_BN` + ReLU and the AveragePooling2D partial sums of the image-pooling
// branch (reference deeplabv3p/models/layers.py:146-153 -> :100-104, :132) in one pass over x per rate.
//
// A dilation-r 3x3 conv is r*r independent dense 3x3 convs on the phase images x[pi + r*a, pj + r*t].  A CTA gathers a
// batch of phase images of one (image, 32-channel group, rate) into shared memory — ONE TMA load per phase image
// through a per-phase strided tensor map {C, ceil((w-pj)/r), ceil((h-pi)/r), B} (box origin (-1,-1): the zero border and
// everything past the map edge is TMA zero fill = 'same' padding); a 16-byte cp.async gather of the same data is limited
// by the SM's outstanding-request capacity (1.41 ms at cfg 3 vs the TMA's in-flight depth).  Half-warps (16 lanes x bf16x2 = the 32 channels) then run
// the dense 3x3 conv on one 4-column segment of one phase image each with a three-row register window: every input is
// read from HBM once per rate and from shared memory 1.5 times.  Three CTAs per SM overlap one CTA's gather
// with the others' arithmetic.
#pragma once

#include <cuda.h>

#include "mem_kernels.cuh"
#include "aspp_dw_fast.cuh"

namespace dlv3p {

constexpr int kGatherThreads = 256;
constexpr int kGatherSmemBudget = 70 * 1024; // phase images per CTA batch (3 CTAs per SM)
constexpr int kGatherSmemBytes = kGatherSmemBudget + (9 * 32 + 32 + 16 * 32) * 4 + 16;

struct AsppGatherParams {
  const __nv_bfloat16* x;   // [B,h,w,C]
  const float* w;           // [3][9][C]  BN scale folded
  const float* shift;       // [3][C]
  __nv_bfloat16* out;       // [3][C/64][B*h*w][64]   (K-block-major, same as the other ASPP depthwise kernels)
  float* pool_partial;      // [B][pool_slots][C]: one slot per rate-0 batch
  const uint32_t* batches;  // [num_batches] packed (ri << 28 | nph << 20 | first phase)
  const CUtensorMap* maps;  // [sum r*r] per-phase tensor maps over x, rate-major (device memory)
  int map_off[3];           // first map of each rate
  int B, h, w_, C, nchunks;
  int rates[3];
  int na[3], nt[3];         // phase-image extent ceil(h / r), ceil(w / r)
  int ts[3];                // output columns per segment (4: 36 row registers -> three CTAs per SM; measured better than 6 / 8 with two)
  int nseg[3];              // ceil(nt / ts)
  int num_batches;          // per (image, channel group)
  int pool_slots;           // number of rate-0 batches
  int debug;
};

// one (phase image, TS-column segment) on a half-warp.  img = this lane's view of the zero-bordered phase image
// [na + 2][nseg * TS + 2][32 ch] bf16 in shared memory: no bounds checks, every load is base + immediate offset, and
// everything past the map edge reads as zero (so the pooling sum needs no mask and only the store is predicated).
// The loop is written for the instruction count (the kernel is issue bound: 43 instructions per output before this form, 9 of them
// FFMA2): only the rows that exist are visited, the output pointer advances by one constant per row and every store of a row is
// pointer + immediate; the column predicate is hoisted (kFull: all TS outputs of the segment exist).
template <int RC, int TS, bool kPool, bool kFull>
__device__ __forceinline__ void aspp_gather_rows(const uint8_t* base, int pitch, uint8_t* orow, size_t ostep, int rows, int nvalid, int r,
                                                 const unsigned long long (&wt)[9], unsigned long long sh, unsigned long long& psum) {
  auto load_row = [&](const uint8_t* p, unsigned long long (&row)[TS + 2]) {
#pragma unroll
    for (int t = 0; t < TS + 2; ++t) row[t] = f32x2_from_bf16x2(*reinterpret_cast<const uint32_t*>(p + t * 64));
  };
  auto emit = [&](const unsigned long long (&top)[TS + 2], const unsigned long long (&mid)[TS + 2], const unsigned long long (&bot)[TS + 2]) {
#pragma unroll
    for (int t = 0; t < TS; ++t) {
      unsigned long long acc = sh;
      ffma2_acc(acc, wt[0], top[t]); ffma2_acc(acc, wt[1], top[t + 1]); ffma2_acc(acc, wt[2], top[t + 2]);
      ffma2_acc(acc, wt[3], mid[t]); ffma2_acc(acc, wt[4], mid[t + 1]); ffma2_acc(acc, wt[5], mid[t + 2]);
      ffma2_acc(acc, wt[6], bot[t]); ffma2_acc(acc, wt[7], bot[t + 1]); ffma2_acc(acc, wt[8], bot[t + 2]);
      if (kPool) psum = fadd2(psum, mid[t + 1]);
      const uint32_t v = f32x2_to_bf16x2_relu(acc);
      if (RC) {
        if (kFull || t < nvalid) *reinterpret_cast<uint32_t*>(orow + t * (RC * 128)) = v;
      } else {
        if (kFull || t < nvalid) *reinterpret_cast<uint32_t*>(orow + static_cast<size_t>(t) * r * 128) = v;
      }
    }
    orow += ostep;
  };
  unsigned long long ra[TS + 2], rb[TS + 2], rc[TS + 2];
  load_row(base, ra);                 // padded row 0 = map row -1 of the phase image
  load_row(base + pitch, rb);
  const uint8_t* p = base + 2 * pitch;
  int a = 0;
  for (; a + 3 <= rows; a += 3) {
    load_row(p, rc);
    emit(ra, rb, rc);
    load_row(p + pitch, ra);
    emit(rb, rc, ra);
    load_row(p + 2 * pitch, rb);
    emit(rc, ra, rb);
    p += 3 * pitch;
  }
  if (a < rows) {
    load_row(p, rc);
    emit(ra, rb, rc);
    if (a + 1 < rows) {
      load_row(p + pitch, ra);
      emit(rb, rc, ra);
    }
  }
}

template <int RC, int TS, bool kPool>
__device__ __forceinline__ void aspp_gather_item(const AsppGatherParams& P, const uint8_t* img, int ri, int pi, int pj, int seg,
                                                 uint8_t* out_lane, const unsigned long long (&wt)[9], unsigned long long sh,
                                                 unsigned long long& psum) {
  const int r = RC ? RC : P.rates[ri];
  const int pitch = (P.nseg[ri] * TS + 2) * 64;          // bytes per padded row
  const int t0 = seg * TS;
  const int j0 = pj + r * t0;
  int nvalid = j0 < P.w_ ? (P.w_ - j0 + r - 1) / r : 0;  // outputs of this segment that exist (uniform per half-warp)
  if (nvalid > TS) nvalid = TS;
  if (P.debug & 1) nvalid = 0;
  // rows of this phase image that exist: i = pi + r a < h.  Rows past the map are zero inputs: nothing to store, nothing to pool
  // (a padding segment entirely past the map edge has only zero inputs: skipped)
  const int rows = pi < P.h ? (P.h - pi + r - 1) / r : 0;
  const uint8_t* base = img + t0 * 64;                   // padded row 0 (= row -1), padded column t0 (= column t0 - 1)
  uint8_t* orow = out_lane + (static_cast<size_t>(pi) * P.w_ + j0) * 128;
  const size_t ostep = static_cast<size_t>(r) * P.w_ * 128;
  // the two half-warps of a warp work on different items: pick the code path per WARP (a full and a partial segment side by side
  // would otherwise run both instantiations one after the other)
  const bool full = __all_sync(__activemask(), nvalid == TS);
  if (full) aspp_gather_rows<RC, TS, kPool, true>(base, pitch, orow, ostep, rows, nvalid, r, wt, sh, psum);
  else aspp_gather_rows<RC, TS, kPool, false>(base, pitch, orow, ostep, nvalid > 0 ? rows : 0, nvalid, r, wt, sh, psum);
}

template <int RC>
__device__ __forceinline__ void aspp_gather_body(const AsppGatherParams& P, uint8_t* gather_smem, int batch, int grp, int b, uint32_t e) {
  uint8_t* s_img = gather_smem;
  float* s_w = reinterpret_cast<float*>(s_img + kGatherSmemBudget);   // [9][32] taps of this CTA's rate
  float* s_shift = s_w + 9 * 32;                                       // [32]
  float* s_red = s_shift + 32;                                         // [16][32]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int hw = lane >> 4, l16 = lane & 15;
  const int hwid = warp * 2 + hw;                                      // half-warp 0..15
  const int ri = static_cast<int>(e >> 28), nph = static_cast<int>((e >> 20) & 0xFF), ph0 = static_cast<int>(e & 0xFFFFF);
  const int r = RC ? RC : P.rates[ri];
  const int NA = P.na[ri], NT = P.nt[ri];
  const int ts = P.ts[ri];
  const int rows_per_img = NA + 2, ntp = P.nseg[ri] * ts + 2;   // padded so that every segment's window stays inside its row
  const int img_bytes = rows_per_img * ntp * 64;

  // ---- gather: one TMA load per phase image, all in flight at once
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_red + 16 * 32);
  if (tid == 0) {
    mbar_init(s_bar, 1);
    fence_barrier_init();
    if (!(P.debug & 4)) {   // debug bit 2: no gather (stale shared memory)
      mbar_arrive_expect_tx(s_bar, static_cast<uint32_t>(nph * img_bytes));
      const CUtensorMap* m = P.maps + P.map_off[ri] + ph0;
      for (int ph = 0; ph < nph; ++ph) tma_load_4d(s_img + ph * img_bytes, m + ph, s_bar, grp * 32, -1, -1, b, kEvictNormal);
    } else {
      mbar_arrive(s_bar);
    }
  }
  for (int i = tid; i < 9 * 32; i += kGatherThreads) s_w[i] = __ldg(P.w + (static_cast<size_t>(ri) * 9 + (i >> 5)) * P.C + grp * 32 + (i & 31));
  if (tid < 32) s_shift[tid] = __ldg(P.shift + static_cast<size_t>(ri) * P.C + grp * 32 + tid);
  __syncthreads();
  mbar_wait(s_bar, 0);

  unsigned long long wt[9], sh;
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const float2 v = *reinterpret_cast<const float2*>(s_w + t * 32 + l16 * 2);
    wt[t] = pack2(v.x, v.y);
  }
  {
    const float2 v = *reinterpret_cast<const float2*>(s_shift + l16 * 2);
    sh = pack2(v.x, v.y);
  }
  uint8_t* out_lane = reinterpret_cast<uint8_t*>(P.out) +
                      ((static_cast<size_t>(ri) * P.nchunks + (grp >> 1)) * P.B + b) * P.h * P.w_ * 128 + (grp & 1) * 64 + l16 * 4;
  const int nseg = P.nseg[ri];
  const int items = nph * nseg;
  unsigned long long psum = 0ull;
  for (int it = hwid; it < items && !(P.debug & 2); it += 16) {   // debug bit 1: gather only
    const int ph = it / nseg, seg = it - ph * nseg;
    const int phase = ph0 + ph;
    const int pi = phase / r, pj = phase - pi * r;
    if (ri == 0) aspp_gather_item<RC, 4, true>(P, s_img + ph * img_bytes + l16 * 4, ri, pi, pj, seg, out_lane, wt, sh, psum);
    else aspp_gather_item<RC, 4, false>(P, s_img + ph * img_bytes + l16 * 4, ri, pi, pj, seg, out_lane, wt, sh, psum);
  }
  // ---- image-pooling partial sums: the rate-0 batches cover every pixel exactly once
  if (ri == 0) {
    s_red[hwid * 32 + l16 * 2] = __uint_as_float(static_cast<uint32_t>(psum));
    s_red[hwid * 32 + l16 * 2 + 1] = __uint_as_float(static_cast<uint32_t>(psum >> 32));
    __syncthreads();
    if (tid < 32) {
      float s = 0.0f;
#pragma unroll
      for (int q = 0; q < 16; ++q) s += s_red[q * 32 + tid];
      P.pool_partial[(static_cast<size_t>(b) * P.pool_slots + batch) * P.C + grp * 32 + tid] = s;   // rate-0 batches come first
    }
  }
}

// grid = B * (C / 32) * num_batches, block 256, dynamic smem = kGatherSmemBytes
__global__ void __launch_bounds__(kGatherThreads, 3) aspp_dw_gather_kernel(const __grid_constant__ AsppGatherParams P) {
  extern __shared__ __align__(128) uint8_t gather_smem[];
  const int ngroups = P.C >> 5;
  int bid = blockIdx.x;
  const int batch = bid % P.num_batches; bid /= P.num_batches;
  const int grp = bid % ngroups;
  const int b = bid / ngroups;
  const uint32_t e = __ldg(P.batches + batch);
  switch (P.rates[e >> 28]) {   // the reference's rates (layers.py:118-125) as compile-time constants, anything else at run time
    case 6: aspp_gather_body<6>(P, gather_smem, batch, grp, b, e); break;
    case 12: aspp_gather_body<12>(P, gather_smem, batch, grp, b, e); break;
    case 18: aspp_gather_body<18>(P, gather_smem, batch, grp, b, e); break;
    case 24: aspp_gather_body<24>(P, gather_smem, batch, grp, b, e); break;
    case 36: aspp_gather_body<36>(P, gather_smem, batch, grp, b, e); break;
    default: aspp_gather_body<0>(P, gather_smem, batch, grp, b, e); break;
  }
}

}  // namespace dlv3p
