// train_api.cuh — C ABI (include/dlv3p_train.h) of the training-step operators; included at the end of dlv3p_api.cu so it
// shares that file's helpers (tensor-map encoding, error strings).  Every entry point is asynchronous on the caller's
// stream and allocates nothing: the SM count is cached per device, tensor maps travel as kernel parameters.
#pragma once

#include "../../include/dlv3p_train.h"
#include "tgemm.cuh"
#include "train_kernels.cuh"
#include "p2p_exchange.cuh"

namespace {

int train_prolog(int device, int* num_sms) {
  static int cached_sms[64] = {};
  if (device < 0 || device >= 64) return fail(nullptr, DLV3P_ERR_INVALID, "bad device index");
  CU_TRY(nullptr, cudaSetDevice(device));
  if (!cached_sms[device]) {
    int major = 0, sms = 0;
    CU_TRY(nullptr, cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    CU_TRY(nullptr, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    if (major != 10) return fail(nullptr, DLV3P_ERR_UNSUPPORTED, fmt("device sm_%d: kernels are sm_100a only", major));
    cached_sms[device] = sms;
  }
  *num_sms = cached_sms[device];
  return 0;
}

template <int BN, bool TN>
cudaError_t launch_tgemm_t(const TgLaunch& L, int num_sms, cudaStream_t st) {
  static bool attr_done[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_done[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(tgemm_kernel<BN, TN>, cudaFuncAttributeMaxDynamicSharedMemorySize, TgCfg<BN>::kSmemBytes);
    if (e != cudaSuccess) return e;
    attr_done[dev & 63] = true;
  }
  const int items = L.m_tiles * L.n_tiles * L.splits;
  const int grid = items < num_sms ? items : num_sms;
  tgemm_kernel<BN, TN><<<grid, kTgThreads, TgCfg<BN>::kSmemBytes, st>>>(L);
  return cudaGetLastError();
}
// 2D bf16 [rows, cols] row-major (ld elements), box {64 cols, 64 rows}, 128B swizzle: MN-major operand boxes
bool encode_2d_sw128_box64(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, std::string* err) {
  return encode_2d_sw128(tm, base, rows, cols, ld, 64, err);
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// band-tiled depthwise kernels: output rows per band R such that (R + 2*rate) rows x W pixels x 64 bytes fit ~96 KB of shared memory
// (two CTAs per SM); 0 = the map is too wide for a useful band (falls back to the generic kernel)
constexpr int kDwBandSmem = 96 * 1024;
int dw_band_rows(int H, int W_, int rate, int* smem_bytes) {
  const int row_bytes = W_ * 64;
  const int max_rows = kDwBandSmem / row_bytes;
  int R;
  if (max_rows >= H) R = H;                       // the whole map in one band (halo clipped by the borders)
  else R = max_rows - 2 * rate;
  if (R < 4 || R * 2 < 2 * rate) return 0;        // halo would dominate
  int rows = R + 2 * rate;
  if (rows > H) rows = H;
  *smem_bytes = rows * row_bytes < 8 * 9 * 32 * 4 ? 8 * 9 * 32 * 4 : rows * row_bytes;
  return R;
}
template <class K>
cudaError_t ensure_smem(K kernel, int bytes) {
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

}  // namespace

extern "C" {

size_t dlv3p_train_gemm_partial_bytes(int64_t M, int N, int splits) {
  return splits > 1 ? static_cast<size_t>(splits) * static_cast<size_t>(M) * static_cast<size_t>(N) * sizeof(float) : 0;
}

static int train_gemm_impl(bool tn, int device, const void* a, int64_t lda, const void* b, int64_t ldb, int64_t M, int N, int64_t K, void* d, int64_t ldd,
                           int out_fp32, int splits, void* d_partial, void* cuda_stream) {
  int sms = 0, r = train_prolog(device, &sms);
  if (r) return r;
  const int64_t amin = tn ? M : K, bmin = tn ? N : K;
  if (!a || !b || !d || M < 1 || N < 1 || K < 8 || K % 8 || lda % 8 || ldb % 8 || lda < amin || ldb < bmin || ldd < N || !aligned16(a) || !aligned16(b) ||
      M > (1ll << 30) || K > (1ll << 30) || (tn && (M % 8 || N % 8)))
    return fail(nullptr, DLV3P_ERR_INVALID, "train_gemm: bad arguments (K, lda, ldb multiples of 8; 16-byte aligned operands)");
  if (!out_fp32 && (ldd % 8 || !aligned16(d))) return fail(nullptr, DLV3P_ERR_INVALID, "train_gemm: bf16 output needs ldd % 8 == 0 and a 16-byte aligned base");
  if (out_fp32 && (ldd % 4 || !aligned16(d))) return fail(nullptr, DLV3P_ERR_INVALID, "train_gemm: fp32 output needs ldd % 4 == 0 and a 16-byte aligned base");
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  const int BN = N > 64 ? 256 : 64;
  TgLaunch L{};
  std::string terr;
  if (tn) {   // operands [K, M] and [K, N]: boxes of 64 channels x 64 contraction rows
    if (!encode_2d_sw128_box64(&L.tmap_a, a, K, M, lda, &terr) || !encode_2d_sw128_box64(&L.tmap_b, b, K, N, ldb, &terr)) return fail(nullptr, DLV3P_ERR_CUDA, terr);
  } else if (!encode_2d_sw128(&L.tmap_a, a, M, K, lda, 128, &terr) || !encode_2d_sw128(&L.tmap_b, b, N, K, ldb, BN >= 128 ? 128 : BN, &terr)) {
    return fail(nullptr, DLV3P_ERR_CUDA, terr);
  }
  L.M = static_cast<int>(M); L.N = N; L.K = static_cast<int>(K);
  L.m_tiles = ceil_div(L.M, kTgBM); L.n_tiles = ceil_div(N, BN);
  L.kblocks = ceil_div(L.K, kTgBK);
  if (splits < 1) splits = 1;
  if (splits > L.kblocks) splits = L.kblocks;
  L.kb_per_split = ceil_div(L.kblocks, splits);
  L.splits = ceil_div(L.kblocks, L.kb_per_split);
  if (L.splits > 1) {
    if (!d_partial || (N % 4)) return fail(nullptr, DLV3P_ERR_INVALID, "train_gemm: split-K needs d_partial and N % 4 == 0");
    L.out = d_partial; L.ldd = N; L.out_mode = kTgOutPartial;
  } else {
    L.out = d; L.ldd = ldd; L.out_mode = out_fp32 ? kTgOutF32 : kTgOutBf16;
    if (!out_fp32 && M >= 32) {   // TMA-store epilogue (tiny M keeps the per-thread stores)
      if (!encode_2d_out(&L.tmap_d, d, M, N, ldd, &terr)) return fail(nullptr, DLV3P_ERR_CUDA, terr);
      L.tma_store = 1;
    }
  }
  cudaError_t e;
  if (tn) e = BN == 256 ? launch_tgemm_t<256, true>(L, sms, st) : launch_tgemm_t<64, true>(L, sms, st);
  else e = BN == 256 ? launch_tgemm_t<256, false>(L, sms, st) : launch_tgemm_t<64, false>(L, sms, st);
  CU_TRY(nullptr, e);
  if (L.splits > 1) {
    tgemm_reduce_kernel<<<grid_for(static_cast<size_t>(M) * N, sms), 256, 0, st>>>(static_cast<const float*>(d_partial), L.splits, M, N, d, ldd, out_fp32);
    CU_TRY(nullptr, cudaGetLastError());
  }
  return DLV3P_OK;
}

int dlv3p_train_gemm_nt(int device, const void* a, int64_t lda, const void* b, int64_t ldb, int64_t M, int N, int64_t K, void* d, int64_t ldd,
                        int out_fp32, int splits, void* d_partial, void* cuda_stream) {
  return train_gemm_impl(false, device, a, lda, b, ldb, M, N, K, d, ldd, out_fp32, splits, d_partial, cuda_stream);
}
int dlv3p_train_gemm_tn(int device, const void* a, int64_t lda, const void* b, int64_t ldb, int64_t M, int N, int64_t K, void* d, int64_t ldd,
                        int out_fp32, int splits, void* d_partial, void* cuda_stream) {
  return train_gemm_impl(true, device, a, lda, b, ldb, M, N, K, d, ldd, out_fp32, splits, d_partial, cuda_stream);
}

int dlv3p_train_transpose(int device, const void* in, int64_t R, int C, int64_t ld_in, void* out, int64_t ld_out, void* cuda_stream) {
  int sms = 0, r = train_prolog(device, &sms);
  if (r) return r;
  if (!in || !out || R < 2 || C < 2 || R % 2 || C % 2 || ld_in % 2 || ld_out % 2 || ld_in < C || ld_out < R)
    return fail(nullptr, DLV3P_ERR_INVALID, "train_transpose: bad arguments (R, C, strides even)");
  transpose_bf16_kernel<<<dim3(static_cast<unsigned>((R + 63) / 64), ceil_div(C, 64)), 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(
      static_cast<const __nv_bfloat16*>(in), R, C, ld_in, static_cast<__nv_bfloat16*>(out), ld_out);
  CU_TRY(nullptr, cudaGetLastError());
  return DLV3P_OK;
}

int dlv3p_train_bn_apply(int device, const void* x, int64_t M, int C, const float* d_stats, const float* d_gamma, const float* d_beta, float eps, int relu,
                         void* y, int64_t ldy, void* cuda_stream) {
  int sms = 0, r = train_prolog(device, &sms);
  if (r) return r;
  if (!x || !d_stats || !d_gamma || !d_beta || !y || M < 1 || C < 8 || C % 8 || ldy % 8 || ldy < C || !aligned16(y))
    return fail(nullptr, DLV3P_ERR_INVALID, "train_bn_apply: bad arguments (C % 8, ldy % 8)");
  if (C > 4096) return fail(nullptr, DLV3P_ERR_UNSUPPORTED, "train_bn_apply: C <= 4096");
  if (static_cast<unsigned long long>(M) * (C / 8) < 0xF0000000ull)
    bn_apply_vec_kernel<unsigned int><<<grid_for(static_cast<size_t>(M) * (C / 8), sms), 256, 2 * C * sizeof(float), static_cast<cudaStream_t>(cuda_stream)>>>(
        static_cast<const __nv_bfloat16*>(x), M, C, d_stats, d_gamma, d_beta, eps, relu, static_cast<__nv_bfloat16*>(y), ldy);
  else
    bn_apply_vec_kernel<unsigned long long><<<grid_for(static_cast<size_t>(M) * (C / 8), sms), 256, 2 * C * sizeof(float), static_cast<cudaStream_t>(cuda_stream)>>>(
        static_cast<const __nv_bfloat16*>(x), M, C, d_stats, d_gamma, d_beta, eps, relu, static_cast<__nv_bfloat16*>(y), ldy);
  CU_TRY(nullptr, cudaGetLastError());
  return DLV3P_OK;
}

size_t dlv3p_train_scratch_bytes(int C) {
  const size_t c = C > 0 ? C : 0;
  size_t a = static_cast<size_t>(kTrBands) * 9 * c, b = col_scratch_floats(C, 2), d = static_cast<size_t>(kDwWgradBlocks) * 9 * c;
  if (b > a) a = b;
  if (d > a) a = d;
  return a * sizeof(float);
}

int dlv3p_train_bn_bwd_stats(int device, const void* dy, int64_t ld_dy, const void* y, int64_t ld_y, const void* x, int64_t M, int C, const float* d_stats,
                             float eps, int relu, float* d_sums, void* d_scratch, void* cuda_stream) {
  int sms = 0, r = train_prolog(device, &sms);
  if (r) return r;
  if (!dy || !x || !d_stats || !d_sums || !d_scratch || (relu && !y) || M < 1 || C < 2 || C % 2 || ld_dy % 2 || ld_y % 2)
    return fail(nullptr, DLV3P_ERR_INVALID, "train_bn_bwd_stats: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  if (C % 8 == 0 && ld_dy % 8 == 0 && ld_y % 8 == 0 && aligned16(dy) && (!relu || aligned16(y))) {
    const int bands = col_bands(C);
    bn_bwd_stats_vec_kernel<<<dim3(ceil_div(C, 256), bands), 256, 0, st>>>(static_cast<const __nv_bfloat16*>(dy), ld_dy, static_cast<const __nv_bfloat16*>(y), ld_y,
                                                                            static_cast<const __nv_bfloat16*>(x), M, C, d_stats, eps, relu, bands,
                                                                            static_cast<float*>(d_scratch));
    bands_final_kernel<<<ceil_div(2 * C, 32), dim3(32, kFinalRows), 0, st>>>(static_cast<const float*>(d_scratch), bands, 2 * C, d_sums);
  } else {
    bn_bwd_stats_partial_kernel<<<dim3(ceil_div(C, 64), kTrBands), 256, 0, st>>>(static_cast<const __nv_bfloat16*>(dy), ld_dy, static_cast<const __nv_bfloat16*>(y), ld_y,
                                                                                  static_cast<const __nv_bfloat16*>(x), M, C, d_stats, eps, relu,
                                                                                  static_cast<float*>(d_scratch));
    bands_final_kernel<<<ceil_div(2 * C, 32), dim3(32, kFinalRows), 0, st>>>(static_cast<const float*>(d_scratch), kTrBands, 2 * C, d_sums);
  }
  CU_TRY(nullptr, cudaGetLastError());
  return DLV3P_OK;
}

int dlv3p_train_bn_bwd_apply(int device, const void* dy, int64_t ld_dy, const void* y, int64_t ld_y, const void* x, int64_t M, int C, const float* d_stats,
                             const float* d_sums, const float* d_gamma, float eps, int relu, void* dx, void* cuda_stream) {
  int sms = 0, r = train_prolog(device, &sms);
  if (r) return r;
  if (!dy || !x || !d_stats || !d_sums || !d_gamma || !dx || (relu && !y) || M < 1 || C < 8 || C % 8 || ld_dy % 8 || ld_y % 8 || !aligned16(dy) ||
      (relu && !aligned16(y)))
    return fail(nullptr, DLV3P_ERR_INVALID, "train_bn_bwd_apply: bad arguments (C % 8, strides % 8, 16-byte aligned slices)");
  if (C > 4096) return fail(nullptr, DLV3P_ERR_UNSUPPORTED, "train_bn_bwd_apply: C <= 4096");
  if (static_cast<unsigned long long>(M) * (C / 8) < 0xF0000000ull)
    bn_bwd_apply_vec_kernel<unsigned int><<<grid_for(static_cast<size_t>(M) * (C / 8), sms), 256, 3 * C * sizeof(float), static_cast<cudaStream_t>(cuda_stream)>>>(
        static_cast<const __nv_bfloat16*>(dy), ld_dy, static_cast<const __nv_bfloat16*>(y), ld_y, static_cast<const __nv_bfloat16*>(x), M, C, d_stats, d_sums,
        d_gamma, eps, relu, static_cast<__nv_bfloat16*>(dx));
  else
    bn_bwd_apply_vec_kernel<unsigned long long><<<grid_for(static_cast<size_t>(M) * (C / 8), sms), 256, 3 * C * sizeof(float), static_cast<cudaStream_t>(cuda_stream)>>>(
        static_cast<const __nv_bfloat16*>(dy), ld_dy, static_cast<const __nv_bfloat16*>(y), ld_y, static_cast<const __nv_bfloat16*>(x), M, C, d_stats, d_sums,
        d_gamma, eps, relu, static_cast<__nv_bfloat16*>(dx));
  CU_TRY(nullptr, cudaGetLastError());
  return DLV3P_OK;
}

int dlv3p_train_depthwise(int device, const void* x, int B, int H, int W_, int C, int rate, const float* d_taps, int flip, void* out, void* cuda_stream) {
  int sms = 0, r = train_prolog(device, &sms);
  if (r) return r;
  if (!x || !d_taps || !out || B < 1 || H < 1 || W_ < 1 || C < 8 || C % 8 || rate < 1 || !aligned16(d_taps))
    return fail(nullptr, DLV3P_ERR_INVALID, "train_depthwise: bad arguments (C % 8)");
  int smem = 0;
  const int R = dw_band_rows(H, W_, rate, &smem);
  if (R > 0) {
    static bool attr_done[64] = {};
    if (!attr_done[device]) { CU_TRY(nullptr, ensure_smem(dw_band_kernel, kDwBandSmem + 1024)); attr_done[device] = true; }
    DwBandParams Q{};
    Q.x = static_cast<const __nv_bfloat16*>(x); Q.w = d_taps; Q.out = static_cast<__nv_bfloat16*>(out);
    Q.B = B; Q.H = H; Q.W = W_; Q.C = C; Q.rate = rate; Q.flip = flip ? 1 : 0; Q.R = R;
    dw_band_kernel<<<dim3(ceil_div(H, R), ceil_div(C, 32), B), 256, smem, static_cast<cudaStream_t>(cuda_stream)>>>(Q);
  } else {
    DwParams P{};
    P.x = static_cast<const __nv_bfloat16*>(x); P.w = d_taps; P.shift = nullptr; P.out = static_cast<__nv_bfloat16*>(out);
    P.B = B; P.H = H; P.W = W_; P.C = C; P.rate = rate; P.relu = 0; P.wstride = C; P.flip = flip ? 1 : 0;
    depthwise3x3_kernel<<<grid_for(static_cast<size_t>(B) * H * W_ * (C / 8), sms), 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(P);
  }
  CU_TRY(nullptr, cudaGetLastError());
  return DLV3P_OK;
}

int dlv3p_train_depthwise_wgrad(int device, const void* x, const void* dy, int B, int H, int W_, int C, int rate, float* d_dw, void* d_scratch,
                                void* cuda_stream) {
  int sms = 0, r = train_prolog(device, &sms);
  if (r) return r;
  if (!x || !dy || !d_dw || !d_scratch || B < 1 || H < 1 || W_ < 1 || C < 2 || C % 2 || rate < 1)
    return fail(nullptr, DLV3P_ERR_INVALID, "train_depthwise_wgrad: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  int smem = 0;
  const int R = C % 8 == 0 ? dw_band_rows(H, W_, rate, &smem) : 0;
  if (R > 0 && static_cast<long long>(B) * ceil_div(H, R) <= kDwWgradBlocks) {
    static bool attr_done[64] = {};
    if (!attr_done[device]) { CU_TRY(nullptr, ensure_smem(dw_band_wgrad_kernel, kDwBandSmem + 1024)); attr_done[device] = true; }
    DwBandParams Q{};
    Q.x = static_cast<const __nv_bfloat16*>(x); Q.dy = static_cast<const __nv_bfloat16*>(dy); Q.partial = static_cast<float*>(d_scratch);
    Q.B = B; Q.H = H; Q.W = W_; Q.C = C; Q.rate = rate; Q.R = R;
    const int bands = ceil_div(H, R);
    dw_band_wgrad_kernel<<<dim3(bands, ceil_div(C, 32), B), 256, smem, st>>>(Q);
    bands_final_kernel<<<ceil_div(9 * C, 32), dim3(32, kFinalRows), 0, st>>>(static_cast<const float*>(d_scratch), B * bands, 9 * C, d_dw);
  } else if (C % 8 == 0) {
    const long long npix = static_cast<long long>(B) * H * W_;
    const int G = npix < kDwWgradBlocks ? static_cast<int>(npix) : kDwWgradBlocks;
    dw_wgrad_vec_kernel<<<dim3(G, ceil_div(C / 8, 256)), 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(dy), B, H, W_, C, rate,
                                                                        static_cast<float*>(d_scratch));
    bands_final_kernel<<<ceil_div(9 * C, 32), dim3(32, kFinalRows), 0, st>>>(static_cast<const float*>(d_scratch), G, 9 * C, d_dw);
  } else {
    dw_wgrad_partial_kernel<<<dim3(ceil_div(C, 64), kTrBands), 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(dy), B, H, W_, C, rate,
                                                                              static_cast<float*>(d_scratch));
    bands_final_kernel<<<ceil_div(9 * C, 32), dim3(32, kFinalRows), 0, st>>>(static_cast<const float*>(d_scratch), kTrBands, 9 * C, d_dw);
  }
  CU_TRY(nullptr, cudaGetLastError());
  return DLV3P_OK;
}

int dlv3p_train_resize(int device, const void* x, int B, int hi, int wi, int C, int ho, int wo, void* out, int64_t ld_out, void* cuda_stream) {
  int sms = 0, r = train_prolog(device, &sms);
  if (r) return r;
  if (!x || !out || C < 8 || C % 8 || hi < 1 || wi < 1 || ho < 1 || wo < 1 || ld_out % 8 || ld_out < C || !aligned16(out))
    return fail(nullptr, DLV3P_ERR_INVALID, "train_resize: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  ResizeParams P{};
  P.x = static_cast<const __nv_bfloat16*>(x); P.out = static_cast<__nv_bfloat16*>(out);
  P.B = B; P.hi = hi; P.wi = wi; P.C = C; P.ho = ho; P.wo = wo; P.ldo = static_cast<int>(ld_out); P.col_off = 0;
  P.sy = static_cast<float>(hi) / static_cast<float>(ho); P.sx = static_cast<float>(wi) / static_cast<float>(wo);
  if (ho == 4 * hi && wo == 4 * wi)
    resize_bilinear_x4_kernel<<<grid_for(static_cast<size_t>(B) * (hi + 1) * (wi + 1) * 32, sms), 256, 0, st>>>(P);
  else
    resize_bilinear_kernel<<<dim3(ho, B), 256, 0, st>>>(P);
  CU_TRY(nullptr, cudaGetLastError());
  return DLV3P_OK;
}

int dlv3p_train_resize_bwd(int device, const void* dy, int64_t ld_dy, int B, int hi, int wi, int C, int ho, int wo, void* dx, void* cuda_stream) {
  int sms = 0, r = train_prolog(device, &sms);
  if (r) return r;
  if (!dy || !dx || C < 8 || C % 8 || hi < 1 || wi < 1 || ho < 1 || wo < 1 || ld_dy % 8 || ld_dy < C || !aligned16(dy))
    return fail(nullptr, DLV3P_ERR_INVALID, "train_resize_bwd: bad arguments");
  ResizeBwdParams P{};
  P.dy = static_cast<const __nv_bfloat16*>(dy); P.dx = static_cast<__nv_bfloat16*>(dx); P.ld_dy = ld_dy;
  P.B = B; P.hi = hi; P.wi = wi; P.C = C; P.ho = ho; P.wo = wo;
  P.sy = static_cast<float>(hi) / static_cast<float>(ho); P.sx = static_cast<float>(wi) / static_cast<float>(wo);
  resize_bwd_nhwc_kernel<<<grid_for(static_cast<size_t>(B) * hi * wi * (C / 8), sms), 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(P);
  CU_TRY(nullptr, cudaGetLastError());
  return DLV3P_OK;
}

namespace { constexpr int kLossBlocks = 148 * 8; }
size_t dlv3p_train_loss_scratch_bytes(void) { return static_cast<size_t>(kLossBlocks) * 2 * sizeof(float); }

int dlv3p_train_softmax_loss(int device, const float* logits, int64_t ldl, const float* bias, const uint8_t* labels, int B, int NC, int hi, int wi, int H, int W_,
                             int ignore_index, float inv_norm, int kind, const float* d_class_weights, float focal_gamma, float focal_alpha, float* d_full,
                             float* d_loss, void* d_scratch, void* cuda_stream) {
  int sms = 0, r = train_prolog(device, &sms);
  if (r) return r;
  if (!logits || !bias || !labels || !d_full || !d_loss || !d_scratch || B < 1 || NC < 1 || NC > 256 || hi < 1 || wi < 1 || H < 1 || W_ < 1 || ldl < NC ||
      kind < 0 || kind > 2 || (kind == 1 && !d_class_weights))
    return fail(nullptr, DLV3P_ERR_INVALID, "train_softmax_loss: bad arguments (kind 0 CE, 1 weighted CE + class weights, 2 focal)");
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  LossParams P{};
  P.logits = logits; P.bias = bias; P.labels = labels; P.d_full = d_full; P.block_part = static_cast<float*>(d_scratch); P.ldl = ldl;
  P.B = B; P.NC = NC; P.hi = hi; P.wi = wi; P.H = H; P.W = W_; P.ignore = ignore_index;
  P.sy = static_cast<float>(hi) / static_cast<float>(H); P.sx = static_cast<float>(wi) / static_cast<float>(W_);
  P.inv_norm = inv_norm;
  P.kind = kind; P.class_w = d_class_weights; P.gamma = focal_gamma; P.alpha = focal_alpha;
  const size_t total = static_cast<size_t>(B) * H * W_;
  int grid = static_cast<int>((total + 255) / 256);
  if (grid > kLossBlocks) grid = kLossBlocks;
  if (NC <= 32) softmax_ce_small_kernel<<<grid, 256, 0, st>>>(P);
  else softmax_ce_kernel<<<grid, 256, 0, st>>>(P);
  loss_final_kernel<<<1, 256, 0, st>>>(static_cast<const float*>(d_scratch), grid, inv_norm, d_loss);
  CU_TRY(nullptr, cudaGetLastError());
  return DLV3P_OK;
}

int dlv3p_train_softmax_ce(int device, const float* logits, int64_t ldl, const float* bias, const uint8_t* labels, int B, int NC, int hi, int wi, int H, int W_,
                           int ignore_index, float inv_norm, float* d_full, float* d_loss, void* d_scratch, void* cuda_stream) {
  return dlv3p_train_softmax_loss(device, logits, ldl, bias, labels, B, NC, hi, wi, H, W_, ignore_index, inv_norm, 0, nullptr, 0.0f, 0.0f, d_full, d_loss, d_scratch,
                                  cuda_stream);
}

size_t dlv3p_train_resize_bwd_planar_scratch_bytes(int B, int NC, int hi, int W_) {
  return static_cast<size_t>(B > 0 ? B : 0) * (NC > 0 ? NC : 0) * (hi > 0 ? hi : 0) * (W_ > 0 ? W_ : 0) * sizeof(float);
}

int dlv3p_train_resize_bwd_planar(int device, const float* d_full, int B, int NC, int hi, int wi, int H, int W_, void* dx, int64_t ld_dx, void* d_scratch,
                                  void* cuda_stream) {
  int sms = 0, r = train_prolog(device, &sms);
  if (r) return r;
  if (!d_full || !dx || B < 1 || NC < 1 || hi < 1 || wi < 1 || H < 1 || W_ < 1 || ld_dx < NC)
    return fail(nullptr, DLV3P_ERR_INVALID, "train_resize_bwd_planar: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  ResizeBwdPlanarParams P{};
  P.dy = d_full; P.dx = static_cast<__nv_bfloat16*>(dx); P.ld_dx = ld_dx; P.B = B; P.NC = NC; P.hi = hi; P.wi = wi; P.ho = H; P.wo = W_;
  P.sy = static_cast<float>(hi) / static_cast<float>(H); P.sx = static_cast<float>(wi) / static_cast<float>(W_);
  const unsigned long long nv = static_cast<unsigned long long>(B) * NC * hi * W_;
  if (d_scratch && nv < 0xF0000000ull && hi <= 65535 && B * NC <= 65535) {   // separable: vertical pass into the scratch, then horizontal
    resize_bwd_planar_v_kernel<<<dim3(ceil_div(W_, 256), hi, B * NC), 256, 0, st>>>(P, static_cast<float*>(d_scratch));
    resize_bwd_planar_h_kernel<<<grid_for(static_cast<size_t>(B) * NC * hi * wi, sms), 256, 0, st>>>(P, static_cast<const float*>(d_scratch));
  } else {
    resize_bwd_planar_kernel<<<grid_for(static_cast<size_t>(B) * NC * hi * wi, sms), 256, 0, st>>>(P);
  }
  CU_TRY(nullptr, cudaGetLastError());
  return DLV3P_OK;
}

int dlv3p_train_rows_reduce(int device, const void* x, int64_t ld, int B, int npix, int C, float scale, void* out, int out_fp32, void* cuda_stream) {
  int sms = 0, r = train_prolog(device, &sms);
  if (r) return r;
  if (!x || !out || B < 1 || npix < 1 || C < 2 || C % 2 || ld % 2 || ld < C) return fail(nullptr, DLV3P_ERR_INVALID, "train_rows_reduce: bad arguments");
  rows_reduce_kernel<<<dim3(ceil_div(C, 64), B), 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(static_cast<const __nv_bfloat16*>(x), ld, npix, C, scale,
                                                                                                 out_fp32 ? nullptr : static_cast<__nv_bfloat16*>(out),
                                                                                                 out_fp32 ? static_cast<float*>(out) : nullptr);
  CU_TRY(nullptr, cudaGetLastError());
  return DLV3P_OK;
}

int dlv3p_train_bcast_rows(int device, const void* src, int B, int npix, int C, float scale, void* dst, int64_t ld, int accumulate, void* cuda_stream) {
  int sms = 0, r = train_prolog(device, &sms);
  if (r) return r;
  if (!src || !dst || B < 1 || npix < 1 || C < 8 || C % 8 || ld % 8 || ld < C || !aligned16(dst) || !aligned16(src))
    return fail(nullptr, DLV3P_ERR_INVALID, "train_bcast_rows: bad arguments");
  bcast_rows_kernel<<<grid_for(static_cast<size_t>(B) * npix * (C / 8), sms), 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(
      static_cast<const __nv_bfloat16*>(src), B, npix, C, scale, static_cast<__nv_bfloat16*>(dst), ld, accumulate);
  CU_TRY(nullptr, cudaGetLastError());
  return DLV3P_OK;
}

int dlv3p_train_add(int device, const void* a, const void* b, void* out, int64_t n, void* cuda_stream) {
  int sms = 0, r = train_prolog(device, &sms);
  if (r) return r;
  if (!a || !b || !out || n < 8 || n % 8) return fail(nullptr, DLV3P_ERR_INVALID, "train_add: bad arguments (n % 8)");
  add_bf16_kernel<<<grid_for(static_cast<size_t>(n / 8), sms), 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(
      static_cast<const __nv_bfloat16*>(a), static_cast<const __nv_bfloat16*>(b), static_cast<__nv_bfloat16*>(out), static_cast<size_t>(n / 8));
  CU_TRY(nullptr, cudaGetLastError());
  return DLV3P_OK;
}

int dlv3p_train_dropout(int device, const void* x, void* out, int64_t n, uint32_t seed, const uint32_t* d_seed, float rate, void* cuda_stream) {
  int sms = 0, r = train_prolog(device, &sms);
  if (r) return r;
  if (!x || !out || n < 8 || n % 8 || n >= (1ll << 32) || !(rate >= 0.0f) || !(rate < 1.0f)) return fail(nullptr, DLV3P_ERR_INVALID, "train_dropout: bad arguments");
  const uint32_t threshold = static_cast<uint32_t>(static_cast<double>(rate) * 4294967296.0);
  dropout_kernel<<<grid_for(static_cast<size_t>(n / 8), sms), 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(out), static_cast<size_t>(n / 8), seed, d_seed, threshold, 1.0f / (1.0f - rate));
  CU_TRY(nullptr, cudaGetLastError());
  return DLV3P_OK;
}

int dlv3p_train_sgd(int device, float* w, const float* g, float* v, int64_t n, float lr, float momentum, float l2, float gscale, void* cuda_stream) {
  int sms = 0, r = train_prolog(device, &sms);
  if (r) return r;
  if (!w || !g || !v || n < 1) return fail(nullptr, DLV3P_ERR_INVALID, "train_sgd: bad arguments");
  sgd_kernel<<<grid_for(static_cast<size_t>(n), sms), 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(w, g, v, static_cast<size_t>(n), lr, momentum, l2, gscale);
  CU_TRY(nullptr, cudaGetLastError());
  return DLV3P_OK;
}

int dlv3p_train_cast_bf16(int device, const float* in, void* out, int64_t n, void* cuda_stream) {
  int sms = 0, r = train_prolog(device, &sms);
  if (r) return r;
  if (!in || !out || n < 1) return fail(nullptr, DLV3P_ERR_INVALID, "train_cast_bf16: bad arguments");
  cast_f32_bf16_any_kernel<<<grid_for(static_cast<size_t>(n), sms), 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(in, static_cast<__nv_bfloat16*>(out),
                                                                                                                      static_cast<size_t>(n));
  CU_TRY(nullptr, cudaGetLastError());
  return DLV3P_OK;
}


// ---------------------------------------------------------------------------------------------- peer-memory exchange
struct dlv3p_p2p {
  int device = 0, world = 1, rank = 0;
  size_t payload_floats = 0;
  uint8_t* local = nullptr;            // [flags | payload]
  uint32_t* epoch = nullptr;           // device counter, starts at 1
  P2pPeers peers{};
  bool opened[kP2pMaxWorld] = {};
};

int dlv3p_p2p_create(int device, int world, int rank, size_t payload_floats, dlv3p_p2p** out, uint8_t handle_out[64]) {
  int sms = 0, r = op_prolog(device, &sms);
  if (r) return r;
  if (!out || !handle_out || world < 1 || world > kP2pMaxWorld || rank < 0 || rank >= world) return fail(nullptr, DLV3P_ERR_INVALID, "p2p_create: bad arguments (world <= 16)");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  dlv3p_p2p* c = new dlv3p_p2p();
  c->device = device; c->world = world; c->rank = rank; c->payload_floats = (payload_floats + 3) / 4 * 4;
  const size_t bytes = kP2pFlagBytes + c->payload_floats * sizeof(float);
  if (cudaMalloc(&c->local, bytes) != cudaSuccess || cudaMalloc(&c->epoch, sizeof(uint32_t)) != cudaSuccess) {
    delete c;
    return fail(nullptr, DLV3P_ERR_NOMEM, "p2p_create: cudaMalloc failed");
  }
  cudaMemset(c->local, 0, bytes);
  const uint32_t one = 1;
  cudaMemcpy(c->epoch, &one, sizeof(one), cudaMemcpyHostToDevice);
  cudaIpcMemHandle_t h;
  if (world > 1) {
    cudaError_t e = cudaIpcGetMemHandle(&h, c->local);
    if (e != cudaSuccess) {
      cudaFree(c->local); cudaFree(c->epoch); delete c;
      return fail(nullptr, DLV3P_ERR_CUDA, fmt("cudaIpcGetMemHandle: %s", cudaGetErrorString(e)));
    }
    std::memcpy(handle_out, &h, 64);
  } else {
    std::memset(handle_out, 0, 64);
  }
  c->peers.world = world; c->peers.rank = rank;
  c->peers.base[rank] = c->local;
  *out = c;
  return DLV3P_OK;
}

// handles: world x 64 bytes in rank order (the entry of this rank is ignored)
int dlv3p_p2p_connect(dlv3p_p2p* c, const uint8_t* handles) {
  if (!c || (!handles && c->world > 1)) return fail(nullptr, DLV3P_ERR_INVALID, "p2p_connect: null argument");
  CU_TRY(nullptr, cudaSetDevice(c->device));
  for (int r = 0; r < c->world; ++r) {
    if (r == c->rank || c->opened[r]) continue;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handles + static_cast<size_t>(r) * 64, 64);
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return fail(nullptr, DLV3P_ERR_CUDA, fmt("cudaIpcOpenMemHandle(rank %d): %s", r, cudaGetErrorString(e)));
    c->peers.base[r] = static_cast<uint8_t*>(p);
    c->opened[r] = true;
  }
  return DLV3P_OK;
}

void dlv3p_p2p_destroy(dlv3p_p2p* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  for (int r = 0; r < c->world; ++r)
    if (c->opened[r]) cudaIpcCloseMemHandle(c->peers.base[r]);
  cudaFree(c->local);
  cudaFree(c->epoch);
  delete c;
}

// device pointer of float `off` of this replica's payload area (the producing kernels write their partial sums there)
void* dlv3p_p2p_payload(dlv3p_p2p* c, size_t off) { return c ? c->local + kP2pFlagBytes + off * sizeof(float) : nullptr; }

int dlv3p_p2p_allreduce(dlv3p_p2p* c, int slot, size_t off, int n, float* d_out, void* cuda_stream) {
  if (!c || !d_out || slot < 0 || slot >= kP2pSlots || n < 1 || (n & 3) || (off & 3) || off + n > c->payload_floats)
    return fail(nullptr, DLV3P_ERR_INVALID, "p2p_allreduce: bad arguments (slot < 64, off and n multiples of 4, inside the payload area)");
  const int blocks = n >= 8192 ? 4 : 1;
  p2p_allreduce_kernel<<<blocks, 512, 0, static_cast<cudaStream_t>(cuda_stream)>>>(c->peers, slot, c->epoch, off, n, d_out);
  CU_TRY(nullptr, cudaGetLastError());
  return DLV3P_OK;
}

// once per step, after the step's last collective: the next step's flags compare against epoch + 1
int dlv3p_p2p_advance(dlv3p_p2p* c, void* cuda_stream) {
  if (!c) return fail(nullptr, DLV3P_ERR_INVALID, "p2p_advance: null argument");
  p2p_advance_epoch_kernel<<<1, 1, 0, static_cast<cudaStream_t>(cuda_stream)>>>(c->epoch);
  CU_TRY(nullptr, cudaGetLastError());
  return DLV3P_OK;
}

}  // extern "C"
