// mem_kernels.cuh — the memory-bound kernels of the head: 128-bit vectorised NHWC accesses,
// 8 bf16 channels per thread, fp32 arithmetic.
//   aspp_dw_pool_kernel   three dilated depthwise 3x3 (+BN+ReLU) of ASPP in ONE pass over x, plus the
//                         per-channel partial sums of the image-pooling branch   (layers.py:132, :146-153 -> :100-104)
//   pool_proj_kernel      image_pooling 1x1 + BN + ReLU, folded through concat_projection into a per-image
//                         bias vector (a 1x1 -> h x w bilinear resize is a broadcast)  (layers.py:132-138, :155-159)
//   depthwise3x3_kernel   generic dilated depthwise 3x3 + BN + ReLU (standalone operator / unfused path)
//   resize_bilinear_kernel  tf.image.resize bilinear, half-pixel centres (layers.py:48-50, :207)
//   resize_argmax_*       pred_resize fused with the host argmax (model.py:76 + deeplab.py:99)
//   resize_dense_kernel   pred_resize (+ Softmax) materialised, for callers that want the reference output
#pragma once

#include "sm100_prims.cuh"

namespace dlv3p {

__device__ __forceinline__ void unpack8(const uint4& r, float (&f)[8]) {
  f[0] = bf16_lo(r.x); f[1] = bf16_hi(r.x);
  f[2] = bf16_lo(r.y); f[3] = bf16_hi(r.y);
  f[4] = bf16_lo(r.z); f[5] = bf16_hi(r.z);
  f[6] = bf16_lo(r.w); f[7] = bf16_hi(r.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}

// ------------------------------------------------------------------------------------------------
__global__ void cast_f32_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, size_t n8) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n8;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(in) + 2 * i);
    const float4 b = __ldg(reinterpret_cast<const float4*>(in) + 2 * i + 1);
    stg_v4(out + 8 * i, make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w)));
  }
}
__global__ void cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, size_t n8) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n8;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float f[8];
    unpack8(ldg_nc_v4(in + 8 * i), f);
    reinterpret_cast<float4*>(out)[2 * i] = make_float4(f[0], f[1], f[2], f[3]);
    reinterpret_cast<float4*>(out)[2 * i + 1] = make_float4(f[4], f[5], f[6], f[7]);
  }
}

// ------------------------------------------------------------------------------------------------
// ASPP: three dilated depthwise convs + pooling partial sums in one pass.
// grid (ceil(C/64), nbands, B), block 256 = 8 channel-vectors x 32 pixel lanes.
struct AsppDwParams {
  const __nv_bfloat16* x;   // [B,h,w,C]
  const float* w;           // [nrates][9][C]  BN scale folded
  const float* shift;       // [nrates][C]
  __nv_bfloat16* out;       // [nrates][B*h*w][C]
  float* pool_partial;      // [B][nbands][C]
  int B, h, w_, C;
  int nrates;               // 3 (ASPP) or 0 (ASPP Lite: pooling only)
  int rates[3];
  int rows_per_band, nbands;
};

__global__ void __launch_bounds__(256) aspp_dw_pool_kernel(const AsppDwParams P) {
  __shared__ float s_w[3 * 9 * 64];
  __shared__ float s_shift[3 * 64];
  __shared__ float s_red[32 * 64];
  const int chunk = blockIdx.x, band = blockIdx.y, b = blockIdx.z;
  const int vec = threadIdx.x & 7, pl = threadIdx.x >> 3;
  const int c0 = chunk * 64 + vec * 8;
  const bool active = c0 < P.C;
  for (int i = threadIdx.x; i < P.nrates * 9 * 64; i += 256) {
    const int rt = i / 64, cc = chunk * 64 + (i & 63);
    s_w[i] = cc < P.C ? P.w[static_cast<size_t>(rt) * P.C + cc] : 0.0f;
  }
  for (int i = threadIdx.x; i < P.nrates * 64; i += 256) {
    const int r = i / 64, cc = chunk * 64 + (i & 63);
    s_shift[i] = cc < P.C ? P.shift[static_cast<size_t>(r) * P.C + cc] : 0.0f;
  }
  __syncthreads();

  const int r0 = band * P.rows_per_band;
  const int r1 = min(P.h, r0 + P.rows_per_band);
  const int npix = (r1 - r0) * P.w_;
  const size_t img_px = static_cast<size_t>(P.h) * P.w_;
  const __nv_bfloat16* xb = P.x + static_cast<size_t>(b) * img_px * P.C + c0;
  float psum[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) psum[k] = 0.0f;

  if (active) {
    for (int p = pl; p < npix; p += 32) {
      const int i = r0 + p / P.w_;
      const int j = p % P.w_;
      float ctr[8];
      unpack8(ldg_nc_v4(xb + (static_cast<size_t>(i) * P.w_ + j) * P.C), ctr);
#pragma unroll
      for (int k = 0; k < 8; ++k) psum[k] += ctr[k];
      for (int r = 0; r < P.nrates; ++r) {
        const int d = P.rates[r];
        float acc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = s_shift[r * 64 + vec * 8 + k];
#pragma unroll
        for (int u = 0; u < 3; ++u) {
          const int ii = i + (u - 1) * d;
          if (ii < 0 || ii >= P.h) continue;
#pragma unroll
          for (int v = 0; v < 3; ++v) {
            const int jj = j + (v - 1) * d;
            if (jj < 0 || jj >= P.w_) continue;
            float xv[8];
            if (u == 1 && v == 1) {
#pragma unroll
              for (int k = 0; k < 8; ++k) xv[k] = ctr[k];
            } else {
              unpack8(ldg_nc_v4(xb + (static_cast<size_t>(ii) * P.w_ + jj) * P.C), xv);
            }
            const float* wt = &s_w[(r * 9 + u * 3 + v) * 64 + vec * 8];
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] = fmaf(xv[k], wt[k], acc[k]);
          }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = fmaxf(acc[k], 0.0f);
        __nv_bfloat16* o = P.out + (static_cast<size_t>(r) * P.B * img_px + static_cast<size_t>(b) * img_px +
                                    static_cast<size_t>(i) * P.w_ + j) * P.C + c0;
        stg_v4(o, pack8(acc));
      }
    }
  }
  // deterministic block reduction of the pooling partial sums over the 32 pixel lanes
#pragma unroll
  for (int k = 0; k < 8; ++k) s_red[pl * 64 + vec * 8 + k] = psum[k];
  __syncthreads();
  if (threadIdx.x < 64) {
    float s = 0.0f;
    for (int q = 0; q < 32; ++q) s += s_red[q * 64 + threadIdx.x];
    const int cc = chunk * 64 + threadIdx.x;
    if (cc < P.C) P.pool_partial[(static_cast<size_t>(b) * P.nbands + band) * P.C + cc] = s;
  }
}

// ------------------------------------------------------------------------------------------------
// image-pooling branch folded into a per-image shift of the concat_projection epilogue.
// grid B, block 256 (one thread per output channel).
struct PoolProjParams {
  const float* pool_partial;     // [B][nbands][C]
  const __nv_bfloat16* w_ip;     // [C][256]  image_pooling kernel, bf16, k-major rows
  const float* ip_scale;         // [256] image_pooling_BN folded
  const float* ip_shift;
  const __nv_bfloat16* w_proj4;  // [256][256] rows 0..255 of concat_projection (the b4 slice), [k][n]
  const float* proj_scale;       // [256] concat_projection_BN folded
  const float* proj_shift;
  float* img_shift;              // [B][256]  = (b4 . Wproj4) * proj_scale + proj_shift
  float* b4_out;                 // [B][256]  tap: image_pooling output after BN+ReLU (bf16-rounded)
  int C, nbands;
  float inv_count;               // 1 / (h*w)
};

__global__ void __launch_bounds__(256) pool_proj_kernel(const PoolProjParams P) {
  extern __shared__ float s_mean[];  // [C] + [256]
  float* s_b4 = s_mean + P.C;
  const int b = blockIdx.x, n = threadIdx.x;
  for (int c = n; c < P.C; c += 256) {
    float s = 0.0f;
    for (int q = 0; q < P.nbands; ++q) s += P.pool_partial[(static_cast<size_t>(b) * P.nbands + q) * P.C + c];
    s_mean[c] = s * P.inv_count;
  }
  __syncthreads();
  float acc = 0.0f;
  for (int k = 0; k < P.C; ++k) acc = fmaf(s_mean[k], __bfloat162float(P.w_ip[static_cast<size_t>(k) * 256 + n]), acc);
  float v = fmaxf(fmaf(acc, P.ip_scale[n], P.ip_shift[n]), 0.0f);
  v = __bfloat162float(__float2bfloat16_rn(v));  // activation rounding point, as every other branch
  s_b4[n] = v;
  P.b4_out[static_cast<size_t>(b) * 256 + n] = v;
  __syncthreads();
  float acc2 = 0.0f;
  for (int k = 0; k < 256; ++k) acc2 = fmaf(s_b4[k], __bfloat162float(P.w_proj4[static_cast<size_t>(k) * 256 + n]), acc2);
  P.img_shift[static_cast<size_t>(b) * 256 + n] = fmaf(acc2, P.proj_scale[n], P.proj_shift[n]);
}

// ------------------------------------------------------------------------------------------------
// generic depthwise 3x3 (dilation d, 'same') + scale/shift (+ReLU); one thread = one pixel x 8 channels
struct DwParams {
  const __nv_bfloat16* x;  // [B,H,W,C]
  const float* w;          // [9][C] scale folded
  const float* shift;      // [C]
  __nv_bfloat16* out;      // [B,H,W,C]
  int B, H, W, C, rate, relu;
  int wstride;             // row stride of w (>= C; packed taps may be zero padded)
};
__global__ void __launch_bounds__(256) depthwise3x3_kernel(const DwParams P) {
  const int vecs = P.C >> 3;
  const size_t total = static_cast<size_t>(P.B) * P.H * P.W * vecs;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int vec = static_cast<int>(idx % vecs);
    size_t pix = idx / vecs;
    const int j = static_cast<int>(pix % P.W);
    pix /= P.W;
    const int i = static_cast<int>(pix % P.H);
    const int b = static_cast<int>(pix / P.H);
    const int c0 = vec * 8;
    float acc[8];
    {
      const float4 s0 = __ldg(reinterpret_cast<const float4*>(P.shift + c0));
      const float4 s1 = __ldg(reinterpret_cast<const float4*>(P.shift + c0 + 4));
      acc[0] = s0.x; acc[1] = s0.y; acc[2] = s0.z; acc[3] = s0.w;
      acc[4] = s1.x; acc[5] = s1.y; acc[6] = s1.z; acc[7] = s1.w;
    }
    const __nv_bfloat16* xb = P.x + static_cast<size_t>(b) * P.H * P.W * P.C + c0;
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      const int ii = i + (u - 1) * P.rate;
      if (ii < 0 || ii >= P.H) continue;
#pragma unroll
      for (int v = 0; v < 3; ++v) {
        const int jj = j + (v - 1) * P.rate;
        if (jj < 0 || jj >= P.W) continue;
        float xv[8];
        unpack8(ldg_nc_v4(xb + (static_cast<size_t>(ii) * P.W + jj) * P.C), xv);
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(P.w + (u * 3 + v) * P.wstride + c0));
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(P.w + (u * 3 + v) * P.wstride + c0 + 4));
        acc[0] = fmaf(xv[0], w0.x, acc[0]); acc[1] = fmaf(xv[1], w0.y, acc[1]);
        acc[2] = fmaf(xv[2], w0.z, acc[2]); acc[3] = fmaf(xv[3], w0.w, acc[3]);
        acc[4] = fmaf(xv[4], w1.x, acc[4]); acc[5] = fmaf(xv[5], w1.y, acc[5]);
        acc[6] = fmaf(xv[6], w1.z, acc[6]); acc[7] = fmaf(xv[7], w1.w, acc[7]);
      }
    }
    if (P.relu) {
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = fmaxf(acc[k], 0.0f);
    }
    stg_v4(P.out + ((static_cast<size_t>(b) * P.H + i) * P.W + j) * P.C + c0, pack8(acc));
  }
}

// ------------------------------------------------------------------------------------------------
// tf.image.resize bilinear (TF2: half-pixel centres, no antialias): same operation order as the oracle
//   src = (dst + 0.5) * (in/out) - 0.5 ; lo = max(floor(src),0) ; hi = min(ceil(src), in-1) ; t = src - floor(src)
//   top = tl + (tr - tl) * tx ; bot = bl + (br - bl) * tx ; out = top + (bot - top) * ty      (no FMA contraction)
__device__ __forceinline__ void resize_coord(int dst, float scale, int n_in, int& lo, int& hi, float& t) {
  const float src = __fsub_rn(__fmul_rn(static_cast<float>(dst) + 0.5f, scale), 0.5f);
  const float fl = floorf(src);
  lo = max(static_cast<int>(fl), 0);
  hi = min(static_cast<int>(ceilf(src)), n_in - 1);
  t = __fsub_rn(src, fl);
}
__device__ __forceinline__ float lerp_nofma(float a, float b, float t) { return __fadd_rn(a, __fmul_rn(__fsub_rn(b, a), t)); }

struct ResizeParams {
  const __nv_bfloat16* x;  // [B,hi,wi,C]
  __nv_bfloat16* out;      // [B,ho,wo,ldo] written at column col_off
  int B, hi, wi, C, ho, wo, ldo, col_off;
  float sy, sx;            // hi/ho, wi/wo computed in fp32 on the host exactly like the oracle
};
__global__ void __launch_bounds__(256) resize_bilinear_kernel(const ResizeParams P) {
  const int vecs = P.C >> 3;
  const size_t total = static_cast<size_t>(P.B) * P.ho * P.wo * vecs;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int vec = static_cast<int>(idx % vecs);
    size_t pix = idx / vecs;
    const int X = static_cast<int>(pix % P.wo);
    pix /= P.wo;
    const int Y = static_cast<int>(pix % P.ho);
    const int b = static_cast<int>(pix / P.ho);
    int y0, y1, x0, x1;
    float ty, tx;
    resize_coord(Y, P.sy, P.hi, y0, y1, ty);
    resize_coord(X, P.sx, P.wi, x0, x1, tx);
    const __nv_bfloat16* xb = P.x + static_cast<size_t>(b) * P.hi * P.wi * P.C + vec * 8;
    float tl[8], tr[8], bl[8], br[8], o[8];
    unpack8(ldg_nc_v4(xb + (static_cast<size_t>(y0) * P.wi + x0) * P.C), tl);
    unpack8(ldg_nc_v4(xb + (static_cast<size_t>(y0) * P.wi + x1) * P.C), tr);
    unpack8(ldg_nc_v4(xb + (static_cast<size_t>(y1) * P.wi + x0) * P.C), bl);
    unpack8(ldg_nc_v4(xb + (static_cast<size_t>(y1) * P.wi + x1) * P.C), br);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float top = lerp_nofma(tl[k], tr[k], tx);
      const float bot = lerp_nofma(bl[k], br[k], tx);
      o[k] = lerp_nofma(top, bot, ty);
    }
    stg_v4(P.out + ((static_cast<size_t>(b) * P.ho + Y) * P.wo + X) * P.ldo + P.col_off + vec * 8, pack8(o));
  }
}

// ------------------------------------------------------------------------------------------------
// pred_resize + argmax.  logits are PLANAR fp32 [B][NC][hi][wi] so neighbouring threads read neighbouring
// addresses.  First maximum wins (np.argmax; strict '>' scan from class 0, deeplabSegment.cpp:160-167).
struct ArgmaxParams {
  const float* logits;  // [B,NC,hi,wi]
  uint8_t* labels;      // [B,ho,wo]
  int B, NC, hi, wi, ho, wo;
  float sy, sx;
};

// integer scale S (even): thread = one low-res cell -> an S x 4 strip handled as (S/4) x ... ; specialised S = 4:
// outputs Y in [4m+2, 4m+6), X in [4k+2, 4k+6) share the corners (m, m+1) x (k, k+1); m, k start at -1.
__global__ void __launch_bounds__(256) resize_argmax_x4_kernel(const ArgmaxParams P) {
  const int cells_x = P.wi + 1, cells_y = P.hi + 1;
  const size_t total = static_cast<size_t>(P.B) * cells_y * cells_x;
  const size_t plane = static_cast<size_t>(P.hi) * P.wi;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(idx % cells_x) - 1;
    size_t r = idx / cells_x;
    const int m = static_cast<int>(r % cells_y) - 1;
    const int b = static_cast<int>(r / cells_y);
    const int X0 = 4 * k + 2, Y0 = 4 * m + 2;
    // per-output coordinates through the generic formula (bit-identical to the generic kernel / oracle)
    int ylo[4], yhi[4], xlo[4], xhi[4];
    float ty[4], tx[4];
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      const int Y = min(max(Y0 + d, 0), P.ho - 1), X = min(max(X0 + d, 0), P.wo - 1);
      resize_coord(Y, P.sy, P.hi, ylo[d], yhi[d], ty[d]);
      resize_coord(X, P.sx, P.wi, xlo[d], xhi[d], tx[d]);
    }
    // all valid outputs of the cell share the same corner pixels
    const int y0 = max(m, 0), y1 = min(m + 1, P.hi - 1), x0 = max(k, 0), x1 = min(k + 1, P.wi - 1);
    float best[16];
    int arg[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { best[i] = -INFINITY; arg[i] = 0; }
    const float* lb = P.logits + static_cast<size_t>(b) * P.NC * plane;
    for (int c = 0; c < P.NC; ++c) {
      const float* lp = lb + static_cast<size_t>(c) * plane;
      const float tl = __ldg(lp + static_cast<size_t>(y0) * P.wi + x0);
      const float tr = __ldg(lp + static_cast<size_t>(y0) * P.wi + x1);
      const float bl = __ldg(lp + static_cast<size_t>(y1) * P.wi + x0);
      const float br = __ldg(lp + static_cast<size_t>(y1) * P.wi + x1);
      float top[4], bot[4];
#pragma unroll
      for (int dx = 0; dx < 4; ++dx) {
        top[dx] = lerp_nofma(tl, tr, tx[dx]);
        bot[dx] = lerp_nofma(bl, br, tx[dx]);
      }
#pragma unroll
      for (int dy = 0; dy < 4; ++dy)
#pragma unroll
        for (int dx = 0; dx < 4; ++dx) {
          const float v = lerp_nofma(top[dx], bot[dx], ty[dy]);
          if (v > best[dy * 4 + dx]) { best[dy * 4 + dx] = v; arg[dy * 4 + dx] = c; }
        }
    }
    uint8_t* ob = P.labels + static_cast<size_t>(b) * P.ho * P.wo;
#pragma unroll
    for (int dy = 0; dy < 4; ++dy) {
      const int Y = Y0 + dy;
      if (Y < 0 || Y >= P.ho) continue;
      if (X0 >= 0 && X0 + 3 < P.wo) {
        const uint32_t w = static_cast<uint32_t>(arg[dy * 4]) | (static_cast<uint32_t>(arg[dy * 4 + 1]) << 8) |
                           (static_cast<uint32_t>(arg[dy * 4 + 2]) << 16) | (static_cast<uint32_t>(arg[dy * 4 + 3]) << 24);
        // X0 = 4k+2: 2-byte aligned only -> two 16-bit stores
        *reinterpret_cast<uint16_t*>(ob + static_cast<size_t>(Y) * P.wo + X0) = static_cast<uint16_t>(w & 0xFFFF);
        *reinterpret_cast<uint16_t*>(ob + static_cast<size_t>(Y) * P.wo + X0 + 2) = static_cast<uint16_t>(w >> 16);
      } else {
#pragma unroll
        for (int dx = 0; dx < 4; ++dx) {
          const int X = X0 + dx;
          if (X >= 0 && X < P.wo) ob[static_cast<size_t>(Y) * P.wo + X] = static_cast<uint8_t>(arg[dy * 4 + dx]);
        }
      }
    }
  }
}

// generic scale: one thread per output pixel
__global__ void __launch_bounds__(256) resize_argmax_generic_kernel(const ArgmaxParams P) {
  const size_t total = static_cast<size_t>(P.B) * P.ho * P.wo;
  const size_t plane = static_cast<size_t>(P.hi) * P.wi;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int X = static_cast<int>(idx % P.wo);
    size_t r = idx / P.wo;
    const int Y = static_cast<int>(r % P.ho);
    const int b = static_cast<int>(r / P.ho);
    int y0, y1, x0, x1;
    float ty, tx;
    resize_coord(Y, P.sy, P.hi, y0, y1, ty);
    resize_coord(X, P.sx, P.wi, x0, x1, tx);
    const float* lb = P.logits + static_cast<size_t>(b) * P.NC * plane;
    float best = -INFINITY;
    int arg = 0;
    for (int c = 0; c < P.NC; ++c) {
      const float* lp = lb + static_cast<size_t>(c) * plane;
      const float top = lerp_nofma(__ldg(lp + static_cast<size_t>(y0) * P.wi + x0), __ldg(lp + static_cast<size_t>(y0) * P.wi + x1), tx);
      const float bot = lerp_nofma(__ldg(lp + static_cast<size_t>(y1) * P.wi + x0), __ldg(lp + static_cast<size_t>(y1) * P.wi + x1), tx);
      const float v = lerp_nofma(top, bot, ty);
      if (v > best) { best = v; arg = c; }
    }
    P.labels[idx] = static_cast<uint8_t>(arg);
  }
}

// materialised pred_resize (+ Softmax): fp32 NHWC [B,ho,wo,NC]; one thread per output pixel
struct DenseResizeParams {
  const float* logits;  // planar [B,NC,hi,wi]
  float* out;           // [B,ho,wo,NC]
  int B, NC, hi, wi, ho, wo, softmax;
  float sy, sx;
};
__global__ void __launch_bounds__(256) resize_dense_kernel(const DenseResizeParams P) {
  const size_t total = static_cast<size_t>(P.B) * P.ho * P.wo;
  const size_t plane = static_cast<size_t>(P.hi) * P.wi;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int X = static_cast<int>(idx % P.wo);
    size_t r = idx / P.wo;
    const int Y = static_cast<int>(r % P.ho);
    const int b = static_cast<int>(r / P.ho);
    int y0, y1, x0, x1;
    float ty, tx;
    resize_coord(Y, P.sy, P.hi, y0, y1, ty);
    resize_coord(X, P.sx, P.wi, x0, x1, tx);
    const float* lb = P.logits + static_cast<size_t>(b) * P.NC * plane;
    float* o = P.out + idx * P.NC;
    float mx = -INFINITY;
    for (int c = 0; c < P.NC; ++c) {
      const float* lp = lb + static_cast<size_t>(c) * plane;
      const float top = lerp_nofma(__ldg(lp + static_cast<size_t>(y0) * P.wi + x0), __ldg(lp + static_cast<size_t>(y0) * P.wi + x1), tx);
      const float bot = lerp_nofma(__ldg(lp + static_cast<size_t>(y1) * P.wi + x0), __ldg(lp + static_cast<size_t>(y1) * P.wi + x1), tx);
      const float v = lerp_nofma(top, bot, ty);
      o[c] = v;
      mx = fmaxf(mx, v);
    }
    if (P.softmax) {
      float sum = 0.0f;
      for (int c = 0; c < P.NC; ++c) {
        const float e = expf(o[c] - mx);
        o[c] = e;
        sum += e;
      }
      for (int c = 0; c < P.NC; ++c) o[c] = o[c] / sum;
    }
  }
}

// planar fp32 [B,NC,h,w] -> NHWC fp32 (debug tap)
__global__ void planar_to_nhwc_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int NC, size_t plane) {
  const size_t total = static_cast<size_t>(B) * NC * plane;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % NC);
    size_t r = idx / NC;
    const size_t p = r % plane;
    const size_t b = r / plane;
    out[idx] = in[(b * NC + c) * plane + p];
  }
}

}  // namespace dlv3p
