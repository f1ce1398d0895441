// f32_kernels.cuh — the fp32 PRECISION MODE of the whole model (DLV3P_MODEL_FLAG_FP32): plain fp32 arithmetic end to end, the
// reference's default numerics (TensorFlow runs fp32 unless --mixed_precision, train.py:37-46).  BASELINE north_star: "logits must
// agree within ... 1e-4 in fp32".  These are simple CUDA-core kernels (shared-memory tiled SGEMM, direct convolutions): a mode to
// PROVE results, not the performance path — the bf16 tcgen05 kernels are.  Same graph, same layer semantics, same operation
// order inside each operator as the oracle (conv -> BN as scale/shift -> ReLU; TF half-pixel bilinear with top/bottom lerps).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace dlv3p {

// ---- dense KxK convolution, NHWC, zero padding (pad_t, pad_l), stride s: y = relu?(conv(x) * scale + shift).
// img_u8 != nullptr: the input is a uint8 image normalised on the fly (x / 127.5 - 1, common/data_utils.py:403-416).
struct F32ConvParams {
  const float* x; const uint8_t* img_u8; const float* w;   // w: Keras HWIO [k][k][Cin][Cout]
  const float* scale; const float* shift; float* y;
  int B, H, W, Cin, Ho, Wo, Cout, k, stride, pad_t, pad_l, relu;
};
__global__ void __launch_bounds__(256) f32_conv_kernel(const F32ConvParams P) {
  const long long total = static_cast<long long>(P.B) * P.Ho * P.Wo * P.Cout;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total; idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int co = static_cast<int>(idx % P.Cout);
    long long p = idx / P.Cout;
    const int ox = static_cast<int>(p % P.Wo); p /= P.Wo;
    const int oy = static_cast<int>(p % P.Ho);
    const int b = static_cast<int>(p / P.Ho);
    float acc = 0.0f;
    for (int ky = 0; ky < P.k; ++ky) {
      const int iy = oy * P.stride - P.pad_t + ky;
      if (iy < 0 || iy >= P.H) continue;
      for (int kx = 0; kx < P.k; ++kx) {
        const int ix = ox * P.stride - P.pad_l + kx;
        if (ix < 0 || ix >= P.W) continue;
        const size_t xo = ((static_cast<size_t>(b) * P.H + iy) * P.W + ix) * P.Cin;
        const float* wp = P.w + (static_cast<size_t>(ky) * P.k + kx) * P.Cin * P.Cout + co;
        for (int ci = 0; ci < P.Cin; ++ci) {
          const float xv = P.img_u8 ? __fsub_rn(__fdiv_rn(static_cast<float>(P.img_u8[xo + ci]), 127.5f), 1.0f) : P.x[xo + ci];
          acc = fmaf(xv, wp[static_cast<size_t>(ci) * P.Cout], acc);
        }
      }
    }
    float y = fmaf(acc, P.scale[co], P.shift[co]);
    if (P.relu) y = fmaxf(y, 0.0f);
    P.y[idx] = y;
  }
}

// ---- depthwise 3x3: [ReLU] -> conv (stride s, dilation r, zero padding r) -> * scale + shift -> [ReLU]
struct F32DwParams {
  const float* x; const float* w;   // w: Keras [3][3][C]
  const float* scale; const float* shift; float* y;
  int B, H, W, C, Ho, Wo, stride, rate, relu_in, relu_out;
};
__global__ void __launch_bounds__(256) f32_depthwise_kernel(const F32DwParams P) {
  const long long total = static_cast<long long>(P.B) * P.Ho * P.Wo * P.C;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total; idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % P.C);
    long long p = idx / P.C;
    const int ox = static_cast<int>(p % P.Wo); p /= P.Wo;
    const int oy = static_cast<int>(p % P.Ho);
    const int b = static_cast<int>(p / P.Ho);
    float acc = 0.0f;
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      const int iy = oy * P.stride + (u - 1) * P.rate;
      if (iy < 0 || iy >= P.H) continue;
#pragma unroll
      for (int v = 0; v < 3; ++v) {
        const int ix = ox * P.stride + (v - 1) * P.rate;
        if (ix < 0 || ix >= P.W) continue;
        float xv = P.x[((static_cast<size_t>(b) * P.H + iy) * P.W + ix) * P.C + c];
        if (P.relu_in) xv = fmaxf(xv, 0.0f);
        acc = fmaf(xv, P.w[(u * 3 + v) * P.C + c], acc);
      }
    }
    float y = fmaf(acc, P.scale[c], P.shift[c]);
    if (P.relu_out) y = fmaxf(y, 0.0f);
    P.y[idx] = y;
  }
}

// ---- SGEMM: Y[M, ldy @ col_off .. +N) = epi(A[M, lda (row_stride rows apart)] * W[K, N]); epi = * scale + shift [ReLU] [+ residual]
// a_row_step > 1 reads every a_row_step-th pixel of a [B, H, W, K] tensor (the stride-2 1x1 shortcut conv): rows are remapped through
// (Wo, W) so that output pixel (b, i, j) reads input pixel (b, 2i, 2j).
struct F32GemmParams {
  const float* a; const float* w; const float* scale; const float* shift; const float* residual; float* y;
  int M, K, N, lda, ldy, col_off, relu;
  int sub_Ho, sub_Wo, sub_H, sub_W;    // > 0: stride-2 sampling of the A rows
};
__device__ __forceinline__ size_t f32_a_row(const F32GemmParams& P, int m) {
  if (P.sub_Wo <= 0) return static_cast<size_t>(m);
  const int j = m % P.sub_Wo;
  const int t = m / P.sub_Wo;
  const int i = t % P.sub_Ho;
  const int b = t / P.sub_Ho;
  return (static_cast<size_t>(b) * P.sub_H + 2 * i) * P.sub_W + 2 * j;
}
__global__ void __launch_bounds__(256) f32_gemm_kernel(const F32GemmParams P) {
  __shared__ float sa[16][64 + 1];
  __shared__ float sw[16][64 + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < P.K; k0 += 16) {
    for (int i = threadIdx.x; i < 64 * 16; i += 256) {
      const int r = i >> 4, kk = i & 15;
      const int m = m0 + r, k = k0 + kk;
      sa[kk][r] = (m < P.M && k < P.K) ? P.a[f32_a_row(P, m) * P.lda + k] : 0.0f;
    }
    for (int i = threadIdx.x; i < 16 * 64; i += 256) {
      const int kk = i >> 6, c = i & 63;
      const int k = k0 + kk, n = n0 + c;
      sw[kk][c] = (k < P.K && n < P.N) ? P.w[static_cast<size_t>(k) * P.N + n] : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float av[4], wv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { av[i] = sa[kk][ty * 4 + i]; wv[i] = sw[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= P.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= P.N) continue;
      float y = fmaf(acc[i][j], P.scale ? P.scale[n] : 1.0f, P.shift ? P.shift[n] : 0.0f);
      if (P.relu) y = fmaxf(y, 0.0f);
      if (P.residual) y += P.residual[static_cast<size_t>(m) * P.N + n];
      P.y[static_cast<size_t>(m) * P.ldy + P.col_off + n] = y;
    }
  }
}

// ---- tf.image.resize bilinear (half-pixel centres), NHWC fp32 into a slice of rows of stride ldy
__global__ void __launch_bounds__(256) f32_resize_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int hi, int wi, int C, int ho, int wo,
                                                         int ldy, int col_off) {
  const float sy = static_cast<float>(hi) / static_cast<float>(ho), sx = static_cast<float>(wi) / static_cast<float>(wo);
  const long long total = static_cast<long long>(B) * ho * wo * C;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total; idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % C);
    long long p = idx / C;
    const int ox = static_cast<int>(p % wo); p /= wo;
    const int oy = static_cast<int>(p % ho);
    const int b = static_cast<int>(p / ho);
    const float fy = __fsub_rn(__fmul_rn(static_cast<float>(oy) + 0.5f, sy), 0.5f), fx = __fsub_rn(__fmul_rn(static_cast<float>(ox) + 0.5f, sx), 0.5f);
    const float fly = floorf(fy), flx = floorf(fx);
    const int y0 = max(static_cast<int>(fly), 0), y1 = min(static_cast<int>(ceilf(fy)), hi - 1);
    const int x0 = max(static_cast<int>(flx), 0), x1 = min(static_cast<int>(ceilf(fx)), wi - 1);
    const float ty = __fsub_rn(fy, fly), tx = __fsub_rn(fx, flx);
    const float* xb = x + static_cast<size_t>(b) * hi * wi * C + c;
    const float tl = xb[(static_cast<size_t>(y0) * wi + x0) * C], tr = xb[(static_cast<size_t>(y0) * wi + x1) * C];
    const float bl = xb[(static_cast<size_t>(y1) * wi + x0) * C], br = xb[(static_cast<size_t>(y1) * wi + x1) * C];
    const float top = __fadd_rn(tl, __fmul_rn(__fsub_rn(tr, tl), tx)), bot = __fadd_rn(bl, __fmul_rn(__fsub_rn(br, bl), tx));
    y[(static_cast<size_t>(b) * ho * wo + static_cast<size_t>(oy) * wo + ox) * ldy + col_off + c] = __fadd_rn(top, __fmul_rn(__fsub_rn(bot, top), ty));
  }
}

// ---- AveragePooling2D over the whole map: out[b][c] = mean over npix rows
__global__ void __launch_bounds__(256) f32_global_mean_kernel(const float* __restrict__ x, float* __restrict__ out, int B, int npix, int C) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * C) return;
  const int b = idx / C, c = idx % C;
  float s = 0.0f;
  for (int p = 0; p < npix; ++p) s += x[(static_cast<size_t>(b) * npix + p) * C + c];
  out[idx] = s / static_cast<float>(npix);
}
// broadcast of the 1x1 image-pooling map (bilinear resize of a 1x1 map) into columns [col_off, col_off + C) of the concat rows
__global__ void __launch_bounds__(256) f32_bcast_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int npix, int C, int ld, int col_off) {
  const long long total = static_cast<long long>(B) * npix * C;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total; idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % C);
    const long long row = idx / C;
    dst[row * ld + col_off + c] = src[(row / npix) * C + c];
  }
}
// NHWC fp32 [B, npix, NC] -> planar [B, NC, npix] (what the pred_resize + argmax kernel reads)
__global__ void __launch_bounds__(256) f32_to_planar_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int npix, int NC) {
  const long long total = static_cast<long long>(B) * npix * NC;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total; idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % NC);
    const long long row = idx / NC;
    const long long b = row / npix, p = row % npix;
    y[(b * NC + c) * npix + p] = x[idx];
  }
}

}  // namespace dlv3p
