// p2p_exchange.cuh — the exchanges of the data-parallel training step as ONE-SHOT all-reduces over NVLink peer memory
// (reference: tf.distribute.MirroredStrategy's in-graph NCCL all-reduce, train.py:143-158; SyncBatchNormalization's statistic
// all-reduces, layers.py:63-70).
//
// The SyncBN vectors are <= 54 KB and there are 14 of them per step: an ncclAllReduce costs ~35 us each at 8 GPUs (launch +
// protocol latency), a memory-semantics exchange a few microseconds.  Every replica owns an exchange buffer in device memory that
// its peers map through CUDA IPC (cudaIpcGetMemHandle / cudaIpcOpenMemHandle; NVSwitch gives every pair full bandwidth):
//
//     [ flags: kP2pSlots x kP2pMaxWorld uint32 | payload area ]
//
// Collective number `slot` of step `epoch`, payload = floats [off, off + n) of the payload area:
//   1. the producing kernels of this stream have written the replica's partial sums into ITS OWN payload area;
//   2. one thread per peer publishes  flags[slot][my rank] = epoch  in the PEER's buffer (st.release.sys after a system fence);
//   3. one thread per peer spins on its OWN flags[slot][peer] until it reads >= epoch (ld.acquire.sys; bounded by a clock,
//      then traps instead of hanging the GPU);
//   4. every thread sums the W payloads in RANK ORDER (every replica computes bit-identical results) with 16-byte peer loads and
//      writes the total to a LOCAL output buffer (never in place: a peer may still be reading this replica's payload).
// A payload range is rewritten one whole step later; by then every replica has passed the flags of the collectives in between,
// so no second barrier is needed.  epoch lives in device memory and is advanced by the step's last kernel: a captured CUDA graph
// replays correctly.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace dlv3p {

constexpr int kP2pMaxWorld = 16;
constexpr int kP2pSlots = 64;
constexpr size_t kP2pFlagBytes = static_cast<size_t>(kP2pSlots) * kP2pMaxWorld * sizeof(uint32_t);
#ifndef DLV3P_P2P_TIMEOUT_CYCLES
#define DLV3P_P2P_TIMEOUT_CYCLES (20000000000LL)      // ~10 s: a missing peer traps instead of hanging the GPU
#endif

struct P2pPeers {
  uint8_t* base[kP2pMaxWorld];     // every replica's exchange buffer as mapped in THIS process (base[rank] = the local one)
  int world, rank;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer_v4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

// out[i] = sum over ranks r (in rank order) of payload_r[off + i], i in [0, n); n % 4 == 0, off % 4 == 0.
// Every block waits for the flags itself (block 0 publishes), so any grid size works; the payloads are small: a few blocks.
__global__ void __launch_bounds__(512) p2p_allreduce_kernel(const P2pPeers P, int slot, const uint32_t* __restrict__ epoch_ptr, size_t off, int n,
                                                            float* __restrict__ out) {
  const uint32_t epoch = *epoch_ptr;
  const int W = P.world;
  if (threadIdx.x < W) {
    uint32_t* mine = reinterpret_cast<uint32_t*>(P.base[P.rank]) + slot * kP2pMaxWorld;
    if (blockIdx.x == 0) {
      __threadfence_system();      // this replica's payload (written by earlier kernels of the stream) before the flag
      st_release_sys(reinterpret_cast<uint32_t*>(P.base[threadIdx.x]) + slot * kP2pMaxWorld + P.rank, epoch);
    }
    const long long t0 = clock64();
    while (static_cast<int32_t>(ld_acquire_sys(mine + threadIdx.x) - epoch) < 0) {
      if (clock64() - t0 > DLV3P_P2P_TIMEOUT_CYCLES) {
        printf("dlv3p: p2p all-reduce timeout rank %d waiting for rank %d slot %d epoch %u\n", P.rank, static_cast<int>(threadIdx.x), slot, epoch);
        __trap();
      }
    }
  }
  __syncthreads();
  const int n4 = n >> 2;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < W; ++r) {
      const float4 v = ld_peer_v4(reinterpret_cast<const float*>(P.base[r] + kP2pFlagBytes) + off + 4 * static_cast<size_t>(i));
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    reinterpret_cast<float4*>(out)[i] = acc;
  }
}

__global__ void p2p_advance_epoch_kernel(uint32_t* epoch) { *epoch += 1u; }

// ---- the gradient bucket (12.7 MB at cfg 5): two-shot all-reduce over the same buffers -------------------------------------------------
// publish + wait of collective `slot`; every block waits itself, block 0 publishes: "every earlier kernel of my stream is complete"
__device__ __forceinline__ void p2p_barrier(const P2pPeers& P, int slot, uint32_t epoch) {
  if (threadIdx.x < P.world) {
    uint32_t* mine = reinterpret_cast<uint32_t*>(P.base[P.rank]) + slot * kP2pMaxWorld;
    if (blockIdx.x == 0) {
      __threadfence_system();
      st_release_sys(reinterpret_cast<uint32_t*>(P.base[threadIdx.x]) + slot * kP2pMaxWorld + P.rank, epoch);
    }
    const long long t0 = clock64();
    while (static_cast<int32_t>(ld_acquire_sys(mine + threadIdx.x) - epoch) < 0) {
      if (clock64() - t0 > DLV3P_P2P_TIMEOUT_CYCLES) {
        printf("dlv3p: p2p barrier timeout rank %d waiting for rank %d slot %d epoch %u\n", P.rank, static_cast<int>(threadIdx.x), slot, epoch);
        __trap();
      }
    }
  }
  __syncthreads();
}
// shot 1 (reduce-scatter): replica r sums chunk r of every replica's [src_off, src_off + n) in rank order into its own [red_off + chunk r)
__global__ void __launch_bounds__(512) p2p_reduce_scatter_kernel(const P2pPeers P, int slot, const uint32_t* __restrict__ epoch_ptr, size_t src_off,
                                                                 size_t red_off, size_t n) {
  p2p_barrier(P, slot, *epoch_ptr);
  const size_t n4 = n >> 2, chunk4 = (n4 + P.world - 1) / P.world;
  const size_t lo = chunk4 * P.rank, hi = lo + chunk4 < n4 ? lo + chunk4 : n4;
  float* red = reinterpret_cast<float*>(P.base[P.rank] + kP2pFlagBytes) + red_off;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i0 = lo + blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i0 < hi; i0 += 2 * stride) {   // 2 x W peer loads in flight per thread
    float4 acc[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
    for (int r = 0; r < P.world; ++r) {       // rank order: every replica reduces its chunk, so the order only has to be FIXED
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const size_t i = i0 + u * stride;
        if (i < hi) {
          const float4 v = ld_peer_v4(reinterpret_cast<const float*>(P.base[r] + kP2pFlagBytes) + src_off + 4 * i);
          acc[u].x += v.x; acc[u].y += v.y; acc[u].z += v.z; acc[u].w += v.w;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const size_t i = i0 + u * stride;
      if (i < hi) reinterpret_cast<float4*>(red)[i] = acc[u];
    }
  }
}
// shot 2 (all-gather): every replica copies chunk p of replica p's reduced area into dst (all replicas end with identical bits)
__global__ void __launch_bounds__(512) p2p_all_gather_kernel(const P2pPeers P, int slot, const uint32_t* __restrict__ epoch_ptr, size_t red_off, size_t n,
                                                             float* __restrict__ dst) {
  p2p_barrier(P, slot, *epoch_ptr);
  const size_t n4 = n >> 2, chunk4 = (n4 + P.world - 1) / P.world;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i0 = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i0 < n4; i0 += 4 * stride) {   // four peer loads in flight per thread
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const size_t i = i0 + u * stride;
      if (i < n4) v[u] = ld_peer_v4(reinterpret_cast<const float*>(P.base[static_cast<int>(i / chunk4)] + kP2pFlagBytes) + red_off + 4 * i);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const size_t i = i0 + u * stride;
      if (i < n4) reinterpret_cast<float4*>(dst)[i] = v[u];
    }
  }
}

__global__ void add_u32_kernel(uint32_t* p, uint32_t inc) { *p += inc; }

// Keras moving statistics from the (global) SyncBN sums: moving <- moving * m + batch * (1 - m), biased variance (layers.py:63-70)
__global__ void moving_stats_kernel(const float* __restrict__ stats, const int* __restrict__ ix, int nbn, float momentum, float* __restrict__ mm,
                                    float* __restrict__ mv) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nbn) return;
  const float n = stats[ix[2 * nbn + i]];
  const float mean = stats[ix[i]] / n;
  const float var = fmaxf(stats[ix[nbn + i]] / n - mean * mean, 0.0f);
  mm[i] = mm[i] * momentum + mean * (1.0f - momentum);
  mv[i] = mv[i] * momentum + var * (1.0f - momentum);
}

}  // namespace dlv3p
