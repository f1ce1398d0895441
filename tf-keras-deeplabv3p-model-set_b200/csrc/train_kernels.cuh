// train_kernels.cuh — the memory-bound kernels of the head's TRAINING step (BASELINE cfg 5; reference train.py:143-169
// with MirroredStrategy, SyncBatchNormalization layers.py:63-70, loss deeplabv3p/loss.py:121-156, SGD momentum
// common/model_utils.py:122-123, l2(2e-5) regulariser layers.py:12-21).  The GEMMs are in tgemm.cuh; the forward
// statistics kernels in bn_train.cuh; the depthwise / bilinear forward kernels are shared with inference (mem_kernels.cuh).
//   bn_apply_ld_kernel          y = BN(x) [ReLU], output into a concat slice (row stride ldy)
//   bn_bwd_stats_* / _apply     SyncBN backward: per-replica sum(g), sum(g * xhat) (all-reduced by the caller), then dx
//   dw_wgrad_*                  depthwise 3x3 (dilated) weight gradient: 9 per-channel reductions over all pixels
//   resize_bwd_nhwc_kernel      adjoint of tf.image.resize bilinear (gather form: deterministic, no atomics)
//   softmax_ce_kernel           pred_resize + Softmax + sparse CE with ignore_index, loss and d(logits) at full resolution
//   resize_bwd_planar_kernel    adjoint of pred_resize on the planar fp32 gradient -> bf16 rows for the classifier GEMMs
//   rows_reduce / bcast_rows    image pooling (mean over pixels) and its adjoint, per-image column sums
//   dropout / add / sgd         Dropout(0.5) with a counter-based mask, gradient accumulation, SGD-momentum + L2 update
// Every reduction uses a fixed tree (no atomics): results are bit-reproducible run to run.
#pragma once

#include "bn_train.cuh"
#include "mem_kernels.cuh"

namespace dlv3p {

constexpr int kTrBands = 64;

__device__ __forceinline__ void bn_consts(const float* __restrict__ stats, int C, int c, float eps, float& mean, float& invstd) {
  const float inv_n = 1.0f / stats[2 * C];
  mean = stats[c] * inv_n;
  const float var = fmaxf(stats[C + c] * inv_n - mean * mean, 0.0f);
  invstd = rsqrtf(var + eps);
}

// y[row*ldy + c] = (x - mean) * gamma * invstd + beta [ReLU]; x dense [M, C]; one thread = 8 channels of one row
__global__ void __launch_bounds__(256) bn_apply_ld_kernel(const __nv_bfloat16* __restrict__ x, long long M, int C, const float* __restrict__ stats,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int relu,
                                                          __nv_bfloat16* __restrict__ y, long long ldy) {
  const int vecs = C >> 3;
  const long long total = M * vecs;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total; idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int vec = static_cast<int>(idx % vecs);
    const long long row = idx / vecs;
    float v[8];
    unpack8(ldg_nc_v4(x + idx * 8), v);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = vec * 8 + k;
      float mean, invstd;
      bn_consts(stats, C, c, eps, mean, invstd);
      v[k] = (v[k] - mean) * (gamma[c] * invstd) + beta[c];
      if (relu) v[k] = fmaxf(v[k], 0.0f);
    }
    stg_v4(y + row * ldy + vec * 8, pack8(v));
  }
}

// grid (ceil(C/64), kTrBands), block 256.  g = dy * (y > 0) when relu; partial [kTrBands][2][C] = sum g | sum g*xhat
__global__ void __launch_bounds__(256) bn_bwd_stats_partial_kernel(const __nv_bfloat16* __restrict__ dy, long long ld_dy, const __nv_bfloat16* __restrict__ y,
                                                                   long long ld_y, const __nv_bfloat16* __restrict__ x, long long M, int C,
                                                                   const float* __restrict__ stats, float eps, int relu, float* __restrict__ partial) {
  __shared__ float s_red[8][2][64];
  const int chunk = blockIdx.x, band = blockIdx.y;
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const int c0 = chunk * 64 + lane * 2;
  const long long rows_per_band = (M + kTrBands - 1) / kTrBands;
  const long long r0 = band * rows_per_band, r1 = min(M, r0 + rows_per_band);
  float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
  if (c0 < C) {
    float m0, i0, m1, i1;
    bn_consts(stats, C, c0, eps, m0, i0);
    bn_consts(stats, C, c0 + 1, eps, m1, i1);
    for (long long r = r0 + wp; r < r1; r += 8) {
      const uint32_t dv = __ldg(reinterpret_cast<const unsigned int*>(dy + r * ld_dy + c0));
      float g0 = bf16_lo(dv), g1 = bf16_hi(dv);
      if (relu) {
        const uint32_t yv = __ldg(reinterpret_cast<const unsigned int*>(y + r * ld_y + c0));
        if (!(bf16_lo(yv) > 0.0f)) g0 = 0.0f;
        if (!(bf16_hi(yv) > 0.0f)) g1 = 0.0f;
      }
      const uint32_t xv = __ldg(reinterpret_cast<const unsigned int*>(x + r * C + c0));
      s0 += g0; s1 += g1;
      q0 = fmaf(g0, (bf16_lo(xv) - m0) * i0, q0);
      q1 = fmaf(g1, (bf16_hi(xv) - m1) * i1, q1);
    }
  }
  s_red[wp][0][lane * 2] = s0; s_red[wp][0][lane * 2 + 1] = s1;
  s_red[wp][1][lane * 2] = q0; s_red[wp][1][lane * 2 + 1] = q1;
  __syncthreads();
  if (threadIdx.x < 128) {
    const int which = threadIdx.x >> 6, j = threadIdx.x & 63;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += s_red[w][which][j];
    const int c = chunk * 64 + j;
    if (c < C) partial[(static_cast<size_t>(band) * 2 + which) * C + c] = s;
  }
}
// out[i] = sum over bands of partial[b][i], fixed order: block (32, kFinalRows) — the rows of threads take interleaved bands
// (independent loads in flight), then one pass over shared memory.  extra_index >= 0: out[extra_index] = extra_value (the BN row count).
constexpr int kFinalRows = 32;
__global__ void __launch_bounds__(32 * kFinalRows) bands_final_kernel(const float* __restrict__ partial, int bands, int n, float* __restrict__ out,
                                                                      int extra_index = -1, float extra_value = 0.0f) {
  __shared__ float s_part[kFinalRows][33];
  const int i = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f;
  if (i < n) {
#pragma unroll 4
    for (int b = threadIdx.y; b < bands; b += kFinalRows) s += __ldg(partial + static_cast<size_t>(b) * n + i);
  }
  s_part[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && i < n) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kFinalRows; ++w) t += s_part[w][threadIdx.x];
    out[i] = t;
  }
  if (extra_index >= 0 && blockIdx.x == 0 && threadIdx.x == 0 && threadIdx.y == 0) out[extra_index] = extra_value;
}
// dx = gamma * invstd * (g - S1/n - xhat * S2/n), n = GLOBAL row count (stats[2C], all-reduced), sums all-reduced
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dy, long long ld_dy, const __nv_bfloat16* __restrict__ y,
                                                           long long ld_y, const __nv_bfloat16* __restrict__ x, long long M, int C,
                                                           const float* __restrict__ stats, const float* __restrict__ sums, const float* __restrict__ gamma,
                                                           float eps, int relu, __nv_bfloat16* __restrict__ dx) {
  const int vecs = C >> 3;
  const long long total = M * vecs;
  const float inv_n = 1.0f / stats[2 * C];
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total; idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int vec = static_cast<int>(idx % vecs);
    const long long row = idx / vecs;
    float g[8], xv[8], yv[8];
    unpack8(ldg_nc_v4(dy + row * ld_dy + vec * 8), g);
    unpack8(ldg_nc_v4(x + idx * 8), xv);
    if (relu) unpack8(ldg_nc_v4(y + row * ld_y + vec * 8), yv);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = vec * 8 + k;
      float mean, invstd;
      bn_consts(stats, C, c, eps, mean, invstd);
      float gg = g[k];
      if (relu && !(yv[k] > 0.0f)) gg = 0.0f;
      const float xh = (xv[k] - mean) * invstd;
      g[k] = gamma[c] * invstd * (gg - sums[c] * inv_n - xh * sums[C + c] * inv_n);
    }
    stg_v4(dx + idx * 8, pack8(g));
  }
}

// ------------------------------------------------------------------------------------------------ depthwise wgrad
// dW[u][v][c] = sum_{b,i,j} x[b, i+(u-1)d, j+(v-1)d, c] * dy[b,i,j,c]   ('same' zero padding).
// grid (ceil(C/64), kTrBands), block 256: warp strides over the band's pixels, lane = channel pair.  partial [bands][9][C]
__global__ void __launch_bounds__(256) dw_wgrad_partial_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy, int B, int H, int W,
                                                               int C, int rate, float* __restrict__ partial) {
  __shared__ float s_red[8][9][64];
  const int chunk = blockIdx.x, band = blockIdx.y;
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const int c0 = chunk * 64 + lane * 2;
  const long long npix = static_cast<long long>(B) * H * W;
  const long long per_band = (npix + kTrBands - 1) / kTrBands;
  const long long p0 = band * per_band, p1 = min(npix, p0 + per_band);
  float a0[9], a1[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) { a0[t] = 0.f; a1[t] = 0.f; }
  if (c0 < C) {
    for (long long p = p0 + wp; p < p1; p += 8) {
      const int j = static_cast<int>(p % W);
      const long long r = p / W;
      const int i = static_cast<int>(r % H);
      const long long img = (r / H) * H * W;
      const uint32_t dv = __ldg(reinterpret_cast<const unsigned int*>(dy + p * C + c0));
      const float g0 = bf16_lo(dv), g1 = bf16_hi(dv);
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        const int ii = i + (u - 1) * rate;
        if (ii < 0 || ii >= H) continue;
#pragma unroll
        for (int v = 0; v < 3; ++v) {
          const int jj = j + (v - 1) * rate;
          if (jj < 0 || jj >= W) continue;
          const uint32_t xv = __ldg(reinterpret_cast<const unsigned int*>(x + (img + static_cast<long long>(ii) * W + jj) * C + c0));
          a0[u * 3 + v] = fmaf(bf16_lo(xv), g0, a0[u * 3 + v]);
          a1[u * 3 + v] = fmaf(bf16_hi(xv), g1, a1[u * 3 + v]);
        }
      }
    }
  }
#pragma unroll
  for (int t = 0; t < 9; ++t) { s_red[wp][t][lane * 2] = a0[t]; s_red[wp][t][lane * 2 + 1] = a1[t]; }
  __syncthreads();
  for (int e = threadIdx.x; e < 9 * 64; e += 256) {
    const int t = e >> 6, j = e & 63;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += s_red[w][t][j];
    const int c = chunk * 64 + j;
    if (c < C) partial[(static_cast<size_t>(band) * 9 + t) * C + c] = s;
  }
}

// ------------------------------------------------------------------------------------------------ bilinear adjoint
// outputs Y whose source coordinate falls in (i-1, i+1) touch input i; the range below is conservative and every
// candidate's weight is recomputed with the forward's own resize_coord, so forward and adjoint agree exactly.
__device__ __forceinline__ void adjoint_range(int i, float scale, int n_out, int& a, int& b) {
  a = max(0, static_cast<int>(floorf((static_cast<float>(i) - 0.5f) / scale - 0.5f)) - 1);
  b = min(n_out - 1, static_cast<int>(ceilf((static_cast<float>(i) + 1.5f) / scale - 0.5f)) + 1);
}
__device__ __forceinline__ float adjoint_weight(int Y, float scale, int n_in, int i) {
  int lo, hi;
  float t;
  resize_coord(Y, scale, n_in, lo, hi, t);
  return (lo == i ? 1.0f - t : 0.0f) + (hi == i ? t : 0.0f);
}
// dx[b,i,j,:] = sum_{Y,X} wy(Y,i) wx(X,j) dy[b,Y,X,:]; dy bf16 [B,ho,wo] rows of stride ld_dy, dx dense bf16 [B,hi,wi,C]
struct ResizeBwdParams {
  const __nv_bfloat16* dy;
  __nv_bfloat16* dx;
  long long ld_dy;
  int B, hi, wi, C, ho, wo;
  float sy, sx;
};
__global__ void __launch_bounds__(256) resize_bwd_nhwc_kernel(const ResizeBwdParams P) {
  const int vecs = P.C >> 3;
  const long long total = static_cast<long long>(P.B) * P.hi * P.wi * vecs;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total; idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int vec = static_cast<int>(idx % vecs);
    long long r = idx / vecs;
    const int j = static_cast<int>(r % P.wi);
    r /= P.wi;
    const int i = static_cast<int>(r % P.hi);
    const int b = static_cast<int>(r / P.hi);
    int ya, yb, xa, xb;
    adjoint_range(i, P.sy, P.ho, ya, yb);
    adjoint_range(j, P.sx, P.wo, xa, xb);
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    for (int Y = ya; Y <= yb; ++Y) {
      const float wy = adjoint_weight(Y, P.sy, P.hi, i);
      if (wy == 0.0f) continue;
      const __nv_bfloat16* rowp = P.dy + (static_cast<long long>(b) * P.ho + Y) * P.wo * P.ld_dy + vec * 8;
      for (int X = xa; X <= xb; ++X) {
        const float wgt = wy * adjoint_weight(X, P.sx, P.wi, j);
        if (wgt == 0.0f) continue;
        float g[8];
        unpack8(ldg_nc_v4(rowp + X * P.ld_dy), g);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = fmaf(wgt, g[k], acc[k]);
      }
    }
    stg_v4(P.dx + idx * 8, pack8(acc));
  }
}
// planar fp32 gradient [B, NC, ho, wo] -> bf16 rows [B*hi*wi, ld_dx] (columns >= NC are left untouched: zero padding)
struct ResizeBwdPlanarParams {
  const float* dy;
  __nv_bfloat16* dx;
  long long ld_dx;
  int B, NC, hi, wi, ho, wo;
  float sy, sx;
};
constexpr int kAdjMaxSpan = 24;   // candidate outputs per input and axis held in registers (covers scales up to x8)
__global__ void __launch_bounds__(256) resize_bwd_planar_kernel(const ResizeBwdPlanarParams P) {
  const long long total = static_cast<long long>(P.B) * P.NC * P.hi * P.wi;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total; idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int j = static_cast<int>(idx % P.wi);
    long long r = idx / P.wi;
    const int i = static_cast<int>(r % P.hi);
    r /= P.hi;
    const int c = static_cast<int>(r % P.NC);
    const int b = static_cast<int>(r / P.NC);
    int ya, yb, xa, xb;
    adjoint_range(i, P.sy, P.ho, ya, yb);
    adjoint_range(j, P.sx, P.wo, xa, xb);
    const float* plane = P.dy + (static_cast<long long>(b) * P.NC + c) * P.ho * P.wo;
    float acc = 0.f;
    if (xb - xa < kAdjMaxSpan) {
      float wx[kAdjMaxSpan];   // the horizontal weights once per thread instead of once per (Y, X)
#pragma unroll
      for (int k = 0; k < kAdjMaxSpan; ++k) wx[k] = (xa + k <= xb) ? adjoint_weight(xa + k, P.sx, P.wi, j) : 0.0f;
      for (int Y = ya; Y <= yb; ++Y) {
        const float wy = adjoint_weight(Y, P.sy, P.hi, i);
        if (wy == 0.0f) continue;
        const float* rowp = plane + static_cast<long long>(Y) * P.wo + xa;
        float racc = 0.f;
#pragma unroll
        for (int k = 0; k < kAdjMaxSpan; ++k)
          if (xa + k <= xb && wx[k] != 0.0f) racc = fmaf(wx[k], __ldg(rowp + k), racc);
        acc = fmaf(wy, racc, acc);
      }
    } else {
      for (int Y = ya; Y <= yb; ++Y) {
        const float wy = adjoint_weight(Y, P.sy, P.hi, i);
        if (wy == 0.0f) continue;
        float racc = 0.f;
        for (int X = xa; X <= xb; ++X) {
          const float wgt = adjoint_weight(X, P.sx, P.wi, j);
          if (wgt != 0.0f) racc = fmaf(wgt, __ldg(plane + static_cast<long long>(Y) * P.wo + X), racc);
        }
        acc = fmaf(wy, racc, acc);
      }
    }
    P.dx[((static_cast<long long>(b) * P.hi + i) * P.wi + j) * P.ld_dx + c] = __float2bfloat16_rn(acc);
  }
}

// Separable form of the same adjoint (bilinear weights factor into wy * wx): a vertical pass into an fp32 scratch [B, NC, hi, wo]
// (loads coalesced along X, ~2*scale rows each) and a horizontal pass (~2*scale taps) — 2 x ~10 loads per output instead of ~12 x 12.
// grid (ceil(wo / 256), hi, B * NC): a block works on ONE input row i, so the vertical weights are computed once per block
__global__ void __launch_bounds__(256) resize_bwd_planar_v_kernel(const ResizeBwdPlanarParams P, float* __restrict__ tmp) {
  __shared__ float s_wy[64];
  const int i = blockIdx.y;
  const unsigned bc = blockIdx.z;
  int ya, yb;
  adjoint_range(i, P.sy, P.ho, ya, yb);
  const int span = min(yb - ya + 1, 64);
  if (threadIdx.x < span) s_wy[threadIdx.x] = adjoint_weight(ya + threadIdx.x, P.sy, P.hi, i);
  __syncthreads();
  const int X = blockIdx.x * 256 + threadIdx.x;
  if (X >= P.wo) return;
  const float* col = P.dy + (static_cast<size_t>(bc) * P.ho + ya) * P.wo + X;
  float acc = 0.f;
  for (int k = 0; k < span; ++k) {
    const float wy = s_wy[k];
    if (wy != 0.0f) acc = fmaf(wy, __ldg(col + static_cast<size_t>(k) * P.wo), acc);
  }
  for (int Y = ya + span; Y <= yb; ++Y) {      // scales beyond x32 (never at the reference's sizes)
    const float wy = adjoint_weight(Y, P.sy, P.hi, i);
    if (wy != 0.0f) acc = fmaf(wy, __ldg(P.dy + (static_cast<size_t>(bc) * P.ho + Y) * P.wo + X), acc);
  }
  tmp[(static_cast<size_t>(bc) * P.hi + i) * P.wo + X] = acc;
}
__global__ void __launch_bounds__(256) resize_bwd_planar_h_kernel(const ResizeBwdPlanarParams P, const float* __restrict__ tmp) {
  const unsigned total = static_cast<unsigned>(P.B) * P.NC * P.hi * P.wi;
  for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int j = static_cast<int>(idx % P.wi);
    unsigned r = idx / P.wi;
    const int i = static_cast<int>(r % P.hi);
    r /= P.hi;
    const int c = static_cast<int>(r % P.NC);
    const int b = static_cast<int>(r / P.NC);
    int xa, xb;
    adjoint_range(j, P.sx, P.wo, xa, xb);
    const float* row = tmp + (static_cast<size_t>(b * P.NC + c) * P.hi + i) * P.wo;
    float acc = 0.f;
    for (int X = xa; X <= xb; ++X) {
      const float wx = adjoint_weight(X, P.sx, P.wi, j);
      if (wx != 0.0f) acc = fmaf(wx, __ldg(row + X), acc);
    }
    P.dx[((static_cast<size_t>(b) * P.hi + i) * P.wi + j) * P.ld_dx + c] = __float2bfloat16_rn(acc);
  }
}

// ------------------------------------------------------------------------------------------------ loss
// pred_resize (model.py:76) + Softmax (model.py:86) + SparseCategoricalCrossEntropy with ignore_index (loss.py:121-156):
//   p = softmax(bilinear(logits) + bias); loss_px = -log(clip(p[label], 1e-7, 1 - 1e-7)) * (label != ignore)
//   d(logits_full)[c] = (p[c] - [c == label]) * valid * inv_norm  (zero where the clip is active), inv_norm = 1 / (global B * H * W)
// logits: fp32 rows [B*hi*wi, ldl] (the classifier GEMM output, bias not yet added).  d_full: planar fp32 [B, NC, H, W].
// block_part: [gridDim.x][2] = sum of loss_px | number of valid pixels of the block.
struct LossParams {
  const float* logits;
  const float* bias;
  const uint8_t* labels;
  float* d_full;
  float* block_part;
  long long ldl;
  int B, NC, hi, wi, H, W, ignore;
  float sy, sx, inv_norm;
  int kind;                 // 0 SparseCategoricalCrossEntropy (loss.py:121-156) | 1 Weighted... (:159-192) | 2 SparseSoftmaxFocalLoss (:60-118)
  const float* class_w;     // kind 1: per-class weights [NC]
  float gamma, alpha;       // kind 2
};
// per-pixel loss value and the scalar S with d(loss)/d(logit_k) = S * (p_k - [k == label]) for the three losses of the reference
// (all are functions of the label's probability p only):
//   CE        -log(p)                                          S = 1   (under model.fit the Keras backend sees y_pred come out of a
//             Softmax op and takes its logits path, softmax_cross_entropy_with_logits: no clip_by_value, so confidently wrong pixels
//             keep their gradient; the [1e-7, 1-1e-7] clip only exists on the eager / non-Softmax path)
//   weighted  -w[label] * log(p)                               S = w[label]
//   focal     -alpha * (1-p)^gamma * log(clip(p, 1e-15, .))    S = alpha * ((1-p)^gamma - gamma * (1-p)^(gamma-1) * p * log(p))
__device__ __forceinline__ void loss_terms(const LossParams& P, float pl, int label, float& loss_px, float& S) {
  if (P.kind == 1) {
    const float w = __ldg(P.class_w + label);
    loss_px = -w * __logf(fmaxf(pl, 1.0e-37f));
    S = w;
  } else if (P.kind == 2) {
    const bool clipped = pl < 1e-15f;
    const float p = fmaxf(pl, 1e-15f), q = 1.0f - p, lg = __logf(p);
    const float qg = __powf(q, P.gamma);
    loss_px = -P.alpha * qg * lg;
    const float qg1 = q > 0.0f ? __powf(q, P.gamma - 1.0f) : (P.gamma == 1.0f ? 1.0f : 0.0f);   // pow(0, 0) would be NaN
    S = clipped ? 0.0f : P.alpha * (qg - P.gamma * qg1 * p * lg);
  } else {
    loss_px = -__logf(fmaxf(pl, 1.0e-37f));
    S = 1.0f;
  }
}
__device__ __forceinline__ float interp_logit(const float* p00, const float* p01, const float* p10, const float* p11, int c, float tx, float ty, float bias) {
  const float tl = __ldg(p00 + c), tr = __ldg(p01 + c), bl = __ldg(p10 + c), br = __ldg(p11 + c);
  const float top = lerp_nofma(tl, tr, tx), bot = lerp_nofma(bl, br, tx);
  return lerp_nofma(top, bot, ty) + bias;
}
__global__ void __launch_bounds__(256) softmax_ce_kernel(const LossParams P) {
  __shared__ float s_loss[8], s_cnt[8];
  const long long total = static_cast<long long>(P.B) * P.H * P.W;
  const long long plane = static_cast<long long>(P.H) * P.W;
  float loss_acc = 0.f, cnt_acc = 0.f;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total; idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int X = static_cast<int>(idx % P.W);
    long long r = idx / P.W;
    const int Y = static_cast<int>(r % P.H);
    const int b = static_cast<int>(r / P.H);
    int y0, y1, x0, x1;
    float ty, tx;
    resize_coord(Y, P.sy, P.hi, y0, y1, ty);
    resize_coord(X, P.sx, P.wi, x0, x1, tx);
    const float* base = P.logits + static_cast<long long>(b) * P.hi * P.wi * P.ldl;
    const float* p00 = base + (static_cast<long long>(y0) * P.wi + x0) * P.ldl;
    const float* p01 = base + (static_cast<long long>(y0) * P.wi + x1) * P.ldl;
    const float* p10 = base + (static_cast<long long>(y1) * P.wi + x0) * P.ldl;
    const float* p11 = base + (static_cast<long long>(y1) * P.wi + x1) * P.ldl;
    const int label = P.labels[idx];
    const bool valid = label != P.ignore && label < P.NC;
    float m = -3.0e38f;
    for (int c = 0; c < P.NC; ++c) m = fmaxf(m, interp_logit(p00, p01, p10, p11, c, tx, ty, __ldg(P.bias + c)));
    float s = 0.f;
    for (int c = 0; c < P.NC; ++c) s += __expf(interp_logit(p00, p01, p10, p11, c, tx, ty, __ldg(P.bias + c)) - m);
    const float inv_s = 1.0f / s;
    float pl = 1.0f;
    if (valid) pl = __expf(interp_logit(p00, p01, p10, p11, label, tx, ty, __ldg(P.bias + label)) - m) * inv_s;
    float loss_px = 0.0f, S = 0.0f;
    if (valid) loss_terms(P, pl, label, loss_px, S);
    const float gscale = S * P.inv_norm;
    float* d = P.d_full + static_cast<long long>(b) * P.NC * plane + static_cast<long long>(Y) * P.W + X;
    for (int c = 0; c < P.NC; ++c) {
      const float p = __expf(interp_logit(p00, p01, p10, p11, c, tx, ty, __ldg(P.bias + c)) - m) * inv_s;
      d[c * plane] = (p - (c == label ? 1.0f : 0.0f)) * gscale;
    }
    if (valid) {
      loss_acc += loss_px;
      cnt_acc += 1.0f;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    loss_acc += __shfl_xor_sync(0xFFFFFFFFu, loss_acc, o);
    cnt_acc += __shfl_xor_sync(0xFFFFFFFFu, cnt_acc, o);
  }
  if ((threadIdx.x & 31) == 0) { s_loss[threadIdx.x >> 5] = loss_acc; s_cnt[threadIdx.x >> 5] = cnt_acc; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float l = 0.f, n = 0.f;
    for (int w = 0; w < 8; ++w) { l += s_loss[w]; n += s_cnt[w]; }
    P.block_part[2 * blockIdx.x] = l;
    P.block_part[2 * blockIdx.x + 1] = n;
  }
}
// out[0] = sum of loss_px * inv_norm (this replica's share of the global mean loss), out[1] = valid pixels.  One block of 256:
// thread t adds block partials t, t+256, ... in double, then a fixed tree (deterministic).
__global__ void __launch_bounds__(256) loss_final_kernel(const float* __restrict__ block_part, int nblocks, float inv_norm, float* __restrict__ out) {
  __shared__ double s_l[256], s_n[256];
  double l = 0.0, n = 0.0;
  for (int i = threadIdx.x; i < nblocks; i += 256) { l += block_part[2 * i]; n += block_part[2 * i + 1]; }
  s_l[threadIdx.x] = l; s_n[threadIdx.x] = n;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) { s_l[threadIdx.x] += s_l[threadIdx.x + o]; s_n[threadIdx.x] += s_n[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out[0] = static_cast<float>(s_l[0] * inv_norm);
    out[1] = static_cast<float>(s_n[0]);
  }
}

// NC <= 32: the interpolated logits of a pixel stay in registers — one interpolation and one exp per class instead of three
// interpolations and two exps (the generic kernel recomputes them per pass because it cannot index a register array)
__global__ void __launch_bounds__(256) softmax_ce_small_kernel(const LossParams P) {
  __shared__ float s_loss[8], s_cnt[8];
  const long long total = static_cast<long long>(P.B) * P.H * P.W;
  const long long plane = static_cast<long long>(P.H) * P.W;
  float loss_acc = 0.f, cnt_acc = 0.f;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total; idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int X = static_cast<int>(idx % P.W);
    long long r = idx / P.W;
    const int Y = static_cast<int>(r % P.H);
    const int b = static_cast<int>(r / P.H);
    int y0, y1, x0, x1;
    float ty, tx;
    resize_coord(Y, P.sy, P.hi, y0, y1, ty);
    resize_coord(X, P.sx, P.wi, x0, x1, tx);
    const float* base = P.logits + static_cast<long long>(b) * P.hi * P.wi * P.ldl;
    const float* p00 = base + (static_cast<long long>(y0) * P.wi + x0) * P.ldl;
    const float* p01 = base + (static_cast<long long>(y0) * P.wi + x1) * P.ldl;
    const float* p10 = base + (static_cast<long long>(y1) * P.wi + x0) * P.ldl;
    const float* p11 = base + (static_cast<long long>(y1) * P.wi + x1) * P.ldl;
    const int label = P.labels[idx];
    const bool valid = label != P.ignore && label < P.NC;
    float z[32];
    float m = -3.0e38f;
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      z[c] = c < P.NC ? interp_logit(p00, p01, p10, p11, c, tx, ty, __ldg(P.bias + c)) : -3.0e38f;
      m = fmaxf(m, z[c]);
    }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      z[c] = c < P.NC ? __expf(z[c] - m) : 0.0f;
      s += z[c];
    }
    const float inv_s = 1.0f / s;
    float pl = 1.0f;
#pragma unroll
    for (int c = 0; c < 32; ++c)
      if (valid && c == label) pl = z[c] * inv_s;
    float loss_px = 0.0f, S = 0.0f;
    if (valid) loss_terms(P, pl, label, loss_px, S);
    const float gscale = S * P.inv_norm;
    float* d = P.d_full + static_cast<long long>(b) * P.NC * plane + static_cast<long long>(Y) * P.W + X;
#pragma unroll
    for (int c = 0; c < 32; ++c)
      if (c < P.NC) d[c * plane] = (z[c] * inv_s - (c == label ? 1.0f : 0.0f)) * gscale;
    if (valid) {
      loss_acc += loss_px;
      cnt_acc += 1.0f;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    loss_acc += __shfl_xor_sync(0xFFFFFFFFu, loss_acc, o);
    cnt_acc += __shfl_xor_sync(0xFFFFFFFFu, cnt_acc, o);
  }
  if ((threadIdx.x & 31) == 0) { s_loss[threadIdx.x >> 5] = loss_acc; s_cnt[threadIdx.x >> 5] = cnt_acc; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float l = 0.f, n = 0.f;
    for (int w = 0; w < 8; ++w) { l += s_loss[w]; n += s_cnt[w]; }
    P.block_part[2 * blockIdx.x] = l;
    P.block_part[2 * blockIdx.x + 1] = n;
  }
}

// ------------------------------------------------------------------------------------------------ pooling / broadcast
// out[b][c] = scale * sum_p x[(b*npix + p) * ld + c]   (image pooling: scale = 1/npix, layers.py:132; column sums: scale = 1)
// grid (ceil(C/64), B), block 256; out bf16 [B, C] (or fp32 when out_f32)
__global__ void __launch_bounds__(256) rows_reduce_kernel(const __nv_bfloat16* __restrict__ x, long long ld, int npix, int C, float scale,
                                                          __nv_bfloat16* __restrict__ out_bf16, float* __restrict__ out_f32) {
  __shared__ float s_red[8][64];
  const int chunk = blockIdx.x, b = blockIdx.y;
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const int c0 = chunk * 64 + lane * 2;
  float s0 = 0.f, s1 = 0.f;
  if (c0 < C) {
    const __nv_bfloat16* xb = x + static_cast<long long>(b) * npix * ld + c0;
    for (int p = wp; p < npix; p += 8) {
      const uint32_t v = __ldg(reinterpret_cast<const unsigned int*>(xb + p * ld));
      s0 += bf16_lo(v); s1 += bf16_hi(v);
    }
  }
  s_red[wp][lane * 2] = s0; s_red[wp][lane * 2 + 1] = s1;
  __syncthreads();
  if (threadIdx.x < 64) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += s_red[w][threadIdx.x];
    const int c = chunk * 64 + threadIdx.x;
    if (c < C) {
      if (out_f32) out_f32[static_cast<long long>(b) * C + c] = s * scale;
      else out_bf16[static_cast<long long>(b) * C + c] = __float2bfloat16_rn(s * scale);
    }
  }
}
// dst[(b*npix + p) * ld + c] (+)= scale * src[b*C + c]   (aspp_resize of a 1x1 map = broadcast, layers.py:138; and its use
// as the adjoint of the image pooling mean)
__global__ void __launch_bounds__(256) bcast_rows_kernel(const __nv_bfloat16* __restrict__ src, int B, int npix, int C, float scale,
                                                         __nv_bfloat16* __restrict__ dst, long long ld, int accumulate) {
  const int vecs = C >> 3;
  const long long total = static_cast<long long>(B) * npix * vecs;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total; idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int vec = static_cast<int>(idx % vecs);
    const long long row = idx / vecs;
    const int b = static_cast<int>(row / npix);
    float v[8];
    unpack8(ldg_nc_v4(src + static_cast<long long>(b) * C + vec * 8), v);
    __nv_bfloat16* d = dst + row * ld + vec * 8;
    if (accumulate) {
      float o[8];
      unpack8(*reinterpret_cast<const uint4*>(d), o);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = fmaf(v[k], scale, o[k]);
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] *= scale;
    }
    stg_v4(d, pack8(v));
  }
}

// ------------------------------------------------------------------------------------------------ elementwise
__global__ void __launch_bounds__(256) add_bf16_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b, __nv_bfloat16* __restrict__ out, size_t n8) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n8; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float x[8], y[8];
    unpack8(*reinterpret_cast<const uint4*>(a + 8 * i), x);
    unpack8(*reinterpret_cast<const uint4*>(b + 8 * i), y);
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] += y[k];
    stg_v4(out + 8 * i, pack8(x));
  }
}
// Dropout(rate) (layers.py:161) with a counter-based mask: element e is kept iff fmix32(e * 0x9E3779B1 + seed) >= rate * 2^32;
// kept values are scaled by 1 / (1 - rate).  The same call with the same seed applies the mask to the gradient.
__device__ __forceinline__ uint32_t fmix32(uint32_t h) {
  h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
  return h;
}
__global__ void __launch_bounds__(256) dropout_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, size_t n8, uint32_t seed,
                                                      const uint32_t* __restrict__ d_seed, uint32_t threshold, float keep_scale) {
  if (d_seed) seed = *d_seed;    // seed in device memory: a captured CUDA graph draws a new mask at every replay
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n8; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float v[8];
    unpack8(*reinterpret_cast<const uint4*>(x + 8 * i), v);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const uint32_t e = static_cast<uint32_t>(8 * i + k);
      v[k] = fmix32(e * 0x9E3779B1u + seed) >= threshold ? v[k] * keep_scale : 0.0f;
    }
    stg_v4(out + 8 * i, pack8(v));
  }
}
// SGD with momentum (Keras: v <- m v - lr g; w <- w + v; common/model_utils.py:122-123) with the l2 regulariser's gradient
// (layers.py:12-21: loss += l2 * sum w^2 -> g += 2 l2 w) folded in.  fp32 master weights.
__global__ void __launch_bounds__(256) sgd_kernel(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ v, size_t n, float lr, float momentum,
                                                  float l2, float gscale) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float wi = w[i];
    const float gi = fmaf(2.0f * l2, wi, g[i] * gscale);
    const float vi = momentum * v[i] - lr * gi;
    v[i] = vi;
    w[i] = wi + vi;
  }
}
// fp32 -> bf16 for any n (the weight copies the GEMMs read)
__global__ void __launch_bounds__(256) cast_f32_bf16_any_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, size_t n) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x)
    out[i] = __float2bfloat16_rn(in[i]);
}


// ------------------------------------------------------------------------------------------------ vectorised variants (C % 8 == 0)
// SyncBN backward statistics, same structure as bn_stats_vec_kernel.  partial [bands][2][C] = sum g | sum g * xhat
__global__ void __launch_bounds__(256) bn_bwd_stats_vec_kernel(const __nv_bfloat16* __restrict__ dy, long long ld_dy, const __nv_bfloat16* __restrict__ y,
                                                               long long ld_y, const __nv_bfloat16* __restrict__ x, long long M, int C,
                                                               const float* __restrict__ stats, float eps, int relu, int bands, float* __restrict__ partial) {
  __shared__ float s_red[8][32][17];
  const int chunk = blockIdx.x, band = blockIdx.y;
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const int c0 = chunk * 256 + lane * 8;
  const long long rows_per_band = (M + bands - 1) / bands;
  const long long r0 = band * rows_per_band, r1 = min(M, r0 + rows_per_band);
  float acc[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) acc[k] = 0.f;
  if (c0 < C) {
    float mean[8], invstd[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) bn_consts(stats, C, c0 + k, eps, mean[k], invstd[k]);
#pragma unroll 2
    for (long long r = r0 + wp; r < r1; r += 8) {
      float g[8], xv[8];
      unpack8(ldg_nc_v4(dy + r * ld_dy + c0), g);
      unpack8(ldg_nc_v4(x + r * C + c0), xv);
      if (relu) {
        float yv[8];
        unpack8(ldg_nc_v4(y + r * ld_y + c0), yv);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (!(yv[k] > 0.0f)) g[k] = 0.0f;
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) { acc[k] += g[k]; acc[8 + k] = fmaf(g[k], (xv[k] - mean[k]) * invstd[k], acc[8 + k]); }
    }
  }
  col_reduce_store(acc, s_red, chunk, band, C, partial);
}
// dx = A*g + Bc*x + Cc with per-channel A = gamma*invstd, Bc = -A*invstd*S2/n, Cc = -A*S1/n - Bc*mean (the expansion of
// gamma*invstd*(g - S1/n - xhat*S2/n)), computed once per block into shared memory (dynamic smem: 3*C floats)
template <typename Idx>
__global__ void __launch_bounds__(256) bn_bwd_apply_vec_kernel(const __nv_bfloat16* __restrict__ dy, long long ld_dy, const __nv_bfloat16* __restrict__ y,
                                                               long long ld_y, const __nv_bfloat16* __restrict__ x, long long M, int C,
                                                               const float* __restrict__ stats, const float* __restrict__ sums, const float* __restrict__ gamma,
                                                               float eps, int relu, __nv_bfloat16* __restrict__ dx) {
  extern __shared__ float s_coef[];   // [C] A | [C] Bc | [C] Cc
  {
    const float inv_n = 1.0f / stats[2 * C];
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float mean, invstd;
      bn_consts(stats, C, c, eps, mean, invstd);
      const float A = gamma[c] * invstd;
      const float Bc = -A * invstd * (sums[C + c] * inv_n);
      s_coef[c] = A;
      s_coef[C + c] = Bc;
      s_coef[2 * C + c] = -A * (sums[c] * inv_n) - Bc * mean;
    }
  }
  __syncthreads();
  const Idx vecs = static_cast<Idx>(C >> 3);
  const Idx total = static_cast<Idx>(M) * vecs;
  for (Idx idx = blockIdx.x * static_cast<Idx>(blockDim.x) + threadIdx.x; idx < total; idx += static_cast<Idx>(gridDim.x) * blockDim.x) {
    const Idx row_i = idx / vecs;
    const int vec = static_cast<int>(idx - row_i * vecs);
    const long long row = static_cast<long long>(row_i);
    float g[8], xv[8];
    unpack8(ldg_nc_v4(dy + row * ld_dy + vec * 8), g);
    unpack8(ldg_nc_v4(x + static_cast<size_t>(idx) * 8), xv);
    if (relu) {
      float yv[8];
      unpack8(ldg_nc_v4(y + row * ld_y + vec * 8), yv);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (!(yv[k] > 0.0f)) g[k] = 0.0f;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = vec * 8 + k;
      g[k] = fmaf(s_coef[c], g[k], fmaf(s_coef[C + c], xv[k], s_coef[2 * C + c]));
    }
    stg_v4(dx + static_cast<size_t>(idx) * 8, pack8(g));
  }
}

// depthwise weight gradient, vectorised: one thread = 8 channels of one pixel (16-byte loads), 72 accumulators in registers.
// A block covers min(C/8, 256) channel vectors x ppb pixels per iteration and walks a CONTIGUOUS pixel range, so the x
// neighbours of the same image row come out of L1.  grid (pixel blocks G, ceil(vecs / 256)); partial [G][9][C].
constexpr int kDwWgradBlocks = 296;
__global__ void __launch_bounds__(256) dw_wgrad_vec_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy, int B, int H, int W,
                                                           int C, int rate, float* __restrict__ partial) {
  __shared__ float s_red[256][9];
  const int vecs = C >> 3;
  const int vpb = vecs < 256 ? vecs : 256;          // channel vectors per block
  const int ppb = 256 / vpb;                        // pixels per iteration
  const int tv = threadIdx.x % vpb, tp = threadIdx.x / vpb;
  const int vec = blockIdx.y * 256 + tv;
  const bool active = tp < ppb && vec < vecs;
  const long long npix = static_cast<long long>(B) * H * W;
  const long long per_block = (npix + gridDim.x - 1) / gridDim.x;
  const long long p0 = blockIdx.x * per_block, p1 = min(npix, p0 + per_block);
  float acc[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[t][k] = 0.f;
  if (active) {
    for (long long p = p0 + tp; p < p1; p += ppb) {
      const int j = static_cast<int>(p % W);
      const long long r = p / W;
      const int i = static_cast<int>(r % H);
      const long long img = (r / H) * H * W;
      float g[8];
      unpack8(ldg_nc_v4(dy + p * C + vec * 8), g);
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        const int ii = i + (u - 1) * rate;
        if (ii < 0 || ii >= H) continue;
#pragma unroll
        for (int v = 0; v < 3; ++v) {
          const int jj = j + (v - 1) * rate;
          if (jj < 0 || jj >= W) continue;
          float xv[8];
          unpack8(ldg_nc_v4(x + (img + static_cast<long long>(ii) * W + jj) * C + vec * 8), xv);
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[u * 3 + v][k] = fmaf(xv[k], g[k], acc[u * 3 + v][k]);
        }
      }
    }
  }
  // reduce the ppb pixel lanes (fixed order), one channel of the vector at a time
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    __syncthreads();
#pragma unroll
    for (int t = 0; t < 9; ++t) s_red[threadIdx.x][t] = acc[t][k];
    __syncthreads();
    if (tp == 0 && vec < vecs) {
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        float s = 0.f;
        for (int q = 0; q < ppb; ++q) s += s_red[q * vpb + tv][t];
        partial[(static_cast<size_t>(blockIdx.x) * 9 + t) * C + vec * 8 + k] = s;
      }
    }
  }
}



// ------------------------------------------------------------------------------------------------ band-tiled depthwise (training)
// The generic depthwise kernels read every activation nine times through L2 (dilated taps never share a cache line): they are
// L2-bandwidth bound at ~1 TB/s of useful traffic.  Here a CTA stages ONE band of image rows (+ `rate` halo rows above and
// below, clipped at the image border) of a 32-channel group in shared memory with cp.async and serves all nine taps from
// there: x crosses L2 once (plus the halo).  grid (bands, ceil(C/32), B), block 256 = 64 pixel lanes x 4 channel vectors.
// A whole 32x32 ASPP map is one band (64 KB); the 128-wide decoder maps take 10-row bands.
struct DwBandParams {
  const __nv_bfloat16* x;    // [B,H,W,C] input of the convolution (forward: activations; data gradient: dy)
  const __nv_bfloat16* dy;   // weight gradient only: [B,H,W,C]
  const float* w;            // [9][C] taps (forward / data gradient)
  __nv_bfloat16* out;        // [B,H,W,C]
  float* partial;            // weight gradient: [B * bands][9][C]
  int B, H, W, C, rate, flip, R;   // R: output rows per band
};
__device__ __forceinline__ void cp_async_16(uint32_t smem_addr, const void* gptr) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// stages rows [ylo, yhi) of channel group cg of image b: tile[(row - ylo) * W + px][4 x 16 B]; channels >= C are zero filled
__device__ __forceinline__ void dw_band_load(const DwBandParams& P, uint8_t* tile, int b, int cg, int ylo, int yhi) {
  const int chunks = (yhi - ylo) * P.W * 4;
  const uint32_t tbase = smem_u32(tile);
  for (int e = threadIdx.x; e < chunks; e += blockDim.x) {
    const int v = e & 3, pix = e >> 2;
    const int c = cg * 32 + v * 8;
    if (c < P.C)
      cp_async_16(tbase + e * 16, P.x + (static_cast<long long>(b) * P.H * P.W + static_cast<long long>(ylo) * P.W + pix) * P.C + c);
    else
      *reinterpret_cast<uint4*>(tile + e * 16) = make_uint4(0u, 0u, 0u, 0u);
  }
  cp_async_wait_all();
  __syncthreads();
}
__global__ void __launch_bounds__(256) dw_band_kernel(const DwBandParams P) {
  extern __shared__ __align__(16) uint8_t dw_tile[];
  const int band = blockIdx.x, cg = blockIdx.y, b = blockIdx.z;
  const int y0 = band * P.R, y1 = min(P.H, y0 + P.R);
  const int ylo = max(0, y0 - P.rate), yhi = min(P.H, y1 + P.rate);
  dw_band_load(P, dw_tile, b, cg, ylo, yhi);
  const int tv = threadIdx.x & 3;
  const int c0 = cg * 32 + tv * 8;
  if (c0 >= P.C) return;
  float wt[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const int tap = P.flip ? 8 - t : t;
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(P.w + tap * P.C + c0));
    const float4 w1 = __ldg(reinterpret_cast<const float4*>(P.w + tap * P.C + c0 + 4));
    wt[t][0] = w0.x; wt[t][1] = w0.y; wt[t][2] = w0.z; wt[t][3] = w0.w; wt[t][4] = w1.x; wt[t][5] = w1.y; wt[t][6] = w1.z; wt[t][7] = w1.w;
  }
  const uint32_t tbase = smem_u32(dw_tile);
  const int npx = (y1 - y0) * P.W;
  for (int q = threadIdx.x >> 2; q < npx; q += 64) {
    const int i = y0 + q / P.W, j = q % P.W;
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      const int ii = i + (u - 1) * P.rate;
      if (ii < 0 || ii >= P.H) continue;
#pragma unroll
      for (int v = 0; v < 3; ++v) {
        const int jj = j + (v - 1) * P.rate;
        if (jj < 0 || jj >= P.W) continue;
        float xv[8];
        unpack8(lds_v4(tbase + (((ii - ylo) * P.W + jj) * 4 + tv) * 16), xv);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = fmaf(xv[k], wt[u * 3 + v][k], acc[k]);
      }
    }
    stg_v4(P.out + ((static_cast<long long>(b) * P.H + i) * P.W + j) * P.C + c0, pack8(acc));
  }
}
// weight gradient on the same tiles: x (with halo) from shared memory, dy straight from global (each element is used once).
// 72 accumulators per thread; the 64 pixel lanes are reduced by warp shuffles, then across the 8 warps through shared memory.
__global__ void __launch_bounds__(256) dw_band_wgrad_kernel(const DwBandParams P) {
  extern __shared__ __align__(16) uint8_t dw_tile[];
  const int band = blockIdx.x, cg = blockIdx.y, b = blockIdx.z;
  const int y0 = band * P.R, y1 = min(P.H, y0 + P.R);
  const int ylo = max(0, y0 - P.rate), yhi = min(P.H, y1 + P.rate);
  dw_band_load(P, dw_tile, b, cg, ylo, yhi);
  const int tv = threadIdx.x & 3, lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const int c0 = cg * 32 + tv * 8;
  const bool active = c0 < P.C;
  float acc[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[t][k] = 0.f;
  const uint32_t tbase = smem_u32(dw_tile);
  const int npx = (y1 - y0) * P.W;
  if (active) {
    for (int q = threadIdx.x >> 2; q < npx; q += 64) {
      const int i = y0 + q / P.W, j = q % P.W;
      float g[8];
      unpack8(ldg_nc_v4(P.dy + ((static_cast<long long>(b) * P.H + i) * P.W + j) * P.C + c0), g);
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        const int ii = i + (u - 1) * P.rate;
        if (ii < 0 || ii >= P.H) continue;
#pragma unroll
        for (int v = 0; v < 3; ++v) {
          const int jj = j + (v - 1) * P.rate;
          if (jj < 0 || jj >= P.W) continue;
          float xv[8];
          unpack8(lds_v4(tbase + (((ii - ylo) * P.W + jj) * 4 + tv) * 16), xv);
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[u * 3 + v][k] = fmaf(xv[k], g[k], acc[u * 3 + v][k]);
        }
      }
    }
  }
  // lanes tv, tv+4, ..., tv+28 of a warp hold the same channels: butterfly over lane bits 2..4 (fixed order)
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float v = acc[t][k];
      v += __shfl_xor_sync(0xFFFFFFFFu, v, 4);
      v += __shfl_xor_sync(0xFFFFFFFFu, v, 8);
      v += __shfl_xor_sync(0xFFFFFFFFu, v, 16);
      acc[t][k] = v;
    }
  __syncthreads();                                    // everyone is done with the tile: reuse it as [8 warps][9][32 ch] floats
  float* s_red = reinterpret_cast<float*>(dw_tile);
  if (lane < 4) {
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int k = 0; k < 8; ++k) s_red[(wp * 9 + t) * 32 + lane * 8 + k] = acc[t][k];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 9 * 32; e += 256) {
    const int t = e >> 5, ch = e & 31;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += s_red[(w * 9 + t) * 32 + ch];
    const int c = cg * 32 + ch;
    if (c < P.C) P.partial[(static_cast<size_t>(b * gridDim.x + band) * 9 + t) * P.C + c] = s;
  }
}

}  // namespace dlv3p
