// dlv3p_api.cu — C ABI (include/dlv3p.h) over the sm_100a kernels: context, weight folding/packing,
// TMA descriptors, the forward launch sequence, and the standalone operators used by the parity tests.
//
// Reference sites this file stands in for (paths relative to the reference repo):
//   graph construction  deeplabv3p/models/layers.py:114-219, deeplabv3p/model.py:75-86
//   weight loading      deeplabv3p/model.py:102-103 (Keras layer names / shapes)
//   predict + argmax    deeplab.py:96-99
#include <cuda.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/dlv3p.h"
#include "dwpw_gemm2.cuh"
#include "mem_kernels.cuh"
#include "pw_gemm.cuh"
#include "pw_gemm2.cuh"
#include "aspp_dw_fast.cuh"
#include "aspp_dw_gather.cuh"
#include "bn_train.cuh"
#include "train_kernels.cuh"

using namespace dlv3p;

// =====================================================================================================
// small host utilities
// =====================================================================================================
namespace {

thread_local std::string g_tls_error;

uint16_t f32_to_bf16_rne(float f) {
  uint32_t u;
  std::memcpy(&u, &f, 4);
  if ((u & 0x7F800000u) == 0x7F800000u && (u & 0x007FFFFFu)) return static_cast<uint16_t>((u >> 16) | 0x40);  // NaN
  const uint32_t lsb = (u >> 16) & 1u;
  u += 0x7FFFu + lsb;
  return static_cast<uint16_t>(u >> 16);
}
float bf16_to_f32(uint16_t b) {
  uint32_t u = static_cast<uint32_t>(b) << 16;
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}

std::string fmt(const char* f, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, f);
  vsnprintf(buf, sizeof(buf), f, ap);
  va_end(ap);
  return std::string(buf);
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// cuTensorMapEncodeTiled through the runtime's driver entry point: no link-time dependency on libcuda
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn(std::string* err) {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
    if (err) *err = fmt("cudaGetDriverEntryPoint(cuTensorMapEncodeTiled) failed: %s", cudaGetErrorString(e));
    return nullptr;
  }
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// 2D bf16 [rows, cols] row-major (ld elements), box {64, box_rows}, 128B swizzle
bool encode_2d_sw128(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                     std::string* err) {
  EncodeTiledFn fn = get_encode_fn(err);
  if (!fn) return false;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    if (err) *err = fmt("cuTensorMapEncodeTiled(2d rows=%llu cols=%llu ld=%llu box=%u) -> %d", (unsigned long long)rows,
                        (unsigned long long)cols, (unsigned long long)ld, box_rows, (int)r);
    return false;
  }
  return true;
}
// 2D bf16 [rows, cols] row-major, box {64, 256}, NO swizzle: pixel-major slabs [256 px][64 ch] for the ASPP depthwise kernel
bool encode_2d_slab(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, std::string* err, uint32_t box_cols = 64) {
  EncodeTiledFn fn = get_encode_fn(err);
  if (!fn) return false;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, 256};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    if (err) *err = fmt("cuTensorMapEncodeTiled(slab rows=%llu cols=%llu) -> %d", (unsigned long long)rows, (unsigned long long)cols, (int)r);
    return false;
  }
  return true;
}
// 2D bf16 output view [rows, cols] (row stride ld elements), box {64, 32}, 128B swizzle: the epilogue's TMA store
bool encode_2d_out(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, std::string* err) {
  return encode_2d_sw128(tm, base, rows, cols, ld, 32, err);
}
// 4D bf16 NHWC output {C, W, H, B}, box {64, 16, 2, 1}, 128B swizzle: the fused kernel's TMA-store epilogue
bool encode_4d_out(CUtensorMap* tm, const void* base, uint64_t B, uint64_t H, uint64_t W, uint64_t C, std::string* err, uint32_t box_c = 64) {
  EncodeTiledFn fn = get_encode_fn(err);
  if (!fn) return false;
  cuuint64_t gdim[4] = {C, W, H, B};
  cuuint64_t gstride[3] = {C * 2, W * C * 2, H * W * C * 2};
  cuuint32_t box[4] = {box_c, 16, 2, 1};   // 64 channels: 128B swizzle; 32 channels: 64B swizzle
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, box_c == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    if (err) *err = fmt("cuTensorMapEncodeTiled(4d out) -> %d", (int)r);
    return false;
  }
  return true;
}
// 4D bf16 NHWC {C, W, H, B} (channel stride ldc elements), box {64, bw, bh, 1}, no swizzle, OOB -> 0
bool encode_4d_halo(CUtensorMap* tm, const void* base, uint64_t B, uint64_t H, uint64_t W, uint64_t C, uint64_t ldc,
                    uint32_t bw, uint32_t bh, std::string* err) {
  EncodeTiledFn fn = get_encode_fn(err);
  if (!fn) return false;
  cuuint64_t gdim[4] = {C, W, H, B};
  cuuint64_t gstride[3] = {ldc * 2, W * ldc * 2, H * W * ldc * 2};
  cuuint32_t box[4] = {64, bw, bh, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    if (err) *err = fmt("cuTensorMapEncodeTiled(4d B=%llu H=%llu W=%llu C=%llu) -> %d", (unsigned long long)B,
                        (unsigned long long)H, (unsigned long long)W, (unsigned long long)C, (int)r);
    return false;
  }
  return true;
}

struct BnFold {
  std::vector<float> scale, shift;
};
BnFold fold_bn(const float* gamma, const float* beta, const float* mean, const float* var, int n, float eps) {
  BnFold f;
  f.scale.resize(n);
  f.shift.resize(n);
  for (int i = 0; i < n; ++i) {
    const float inv = gamma[i] / sqrtf(var[i] + eps);
    f.scale[i] = inv;
    f.shift[i] = beta[i] - mean[i] * inv;
  }
  return f;
}

// Keras 1x1 kernel [K][N] fp32 -> bf16 [Npad][Kpad], K-major rows, zero padded
std::vector<uint16_t> pack_pw(const float* w_kn, int K, int N, int k_begin, int k_count, int Npad, int Kpad) {
  std::vector<uint16_t> out(static_cast<size_t>(Npad) * Kpad, 0);
  for (int k = 0; k < k_count; ++k)
    for (int n = 0; n < N; ++n) out[static_cast<size_t>(n) * Kpad + k] = f32_to_bf16_rne(w_kn[static_cast<size_t>(k_begin + k) * N + n]);
  (void)K;
  return out;
}
// Keras depthwise kernel [3][3][C] (x BN scale) -> [9][Cpad]
std::vector<float> pack_dw(const float* w_hwc, const float* scale, int C, int Cpad) {
  std::vector<float> out(static_cast<size_t>(9) * Cpad, 0.0f);
  for (int t = 0; t < 9; ++t)
    for (int c = 0; c < C; ++c) out[static_cast<size_t>(t) * Cpad + c] = w_hwc[static_cast<size_t>(t) * C + c] * (scale ? scale[c] : 1.0f);
  return out;
}

int pick_bn(int N) { return N <= 32 ? 32 : (N <= 64 ? 64 : 256); }
// weight tiles travel in TMA boxes of at most 128 rows (a CTA pair loads one box each, a single CTA two)
uint32_t w_box_rows(int npad) { return static_cast<uint32_t>(npad > 128 ? 128 : npad); }

// (rate, phase, segment) work items of the ASPP depthwise kernels, heaviest rate first, packed for the device
std::vector<uint32_t> build_aspp_items(const AsppDwParams& A) {
  std::vector<uint32_t> t;
  for (int ri = 0; ri < A.nrates; ++ri) {
    const int r = A.rates[ri];
    for (int ph = 0; ph < r * r; ++ph)
      for (int seg = 0; seg < A.nseg[ri]; ++seg)
        t.push_back((static_cast<uint32_t>(ri) << 28) | (static_cast<uint32_t>(seg) << 20) | (static_cast<uint32_t>(ph / r) << 10) | static_cast<uint32_t>(ph % r));
  }
  return t;
}

}  // namespace

// =====================================================================================================
// kernel launch helpers (shared by the context and the standalone operators)
// =====================================================================================================
namespace {

template <int BN>
cudaError_t launch_pw_t(const PwLaunch& L, int num_sms, cudaStream_t st) {
  static bool attr_done[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_done[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(pw_gemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, PwCfg<BN>::kSmemBytes);
    if (e != cudaSuccess) return e;
    attr_done[dev & 63] = true;
  }
  const int items = L.num_tiles * L.num_problems;
  const int grid = items < num_sms ? items : num_sms;
  pw_gemm_kernel<BN><<<grid, kPwThreads, PwCfg<BN>::kSmemBytes, st>>>(L);
  return cudaGetLastError();
}
// CTA-pair (cta_group::2) variant for the 256-wide bf16 GEMMs
cudaError_t launch_pw2(const PwLaunch& L, int num_sms, cudaStream_t st) {
  static bool attr_done[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_done[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(pw_gemm2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPw2SmemBytes);
    if (e != cudaSuccess) return e;
    attr_done[dev & 63] = true;
  }
  const int pairs = (L.num_tiles + 1) / 2 * L.num_problems;
  int grid = 2 * pairs < num_sms ? 2 * pairs : (num_sms & ~1);
  pw_gemm2_kernel<<<grid, kPwThreads, kPw2SmemBytes, st>>>(L);
  return cudaGetLastError();
}

// geometry-specialised ASPP depthwise kernel (aspp_dw_fast.cuh): instantiated for the feature maps the reference's
// 512x512 configurations produce; larger / odd maps run the gather kernel, tiny or unaligned ones the slab / phase kernels
bool aspp_fast_supported(const AsppDwParams& P) {
  if (P.nrates != 3 || P.C % 32 != 0) return false;
  // OS16 maps of 512x512 (the BASELINE configurations), 384x384, 256x256 and 128x128 inputs
  return P.h == P.w_ && (P.h == 32 || P.h == 24 || P.h == 16 || P.h == 8) && P.rates[0] == 6 && P.rates[1] == 12 && P.rates[2] == 18;
}

// large / odd-sized maps: cp.async gather of phase images (aspp_dw_gather.cuh).  Fills the geometry part of the plan and the
// batch table; returns false when a phase image does not fit the shared-memory budget (-> generic kernels).
bool plan_aspp_gather(int h, int w, int C, const int rates[3], AsppGatherParams* G, std::vector<uint32_t>* table) {
  if (C % 32 != 0) return false;
  table->clear();
  G->pool_slots = 0;
  for (int i = 0; i < 3; ++i) {
    const int r = rates[i];
    G->rates[i] = r;
    G->na[i] = ceil_div(h, r);
    G->nt[i] = ceil_div(w, r);
    G->ts[i] = 4;   // four output columns per segment: 36 row registers, three CTAs per SM
    G->nseg[i] = ceil_div(G->nt[i], G->ts[i]);
    const int ntp = G->nseg[i] * G->ts[i] + 2;
    const int img_bytes = (G->na[i] + 2) * ntp * 64;   // zero-bordered phase image, 32 channels
    if (img_bytes > kGatherSmemBudget || r * r > 0xFFFFF || ntp > 256 || G->na[i] + 2 > 256 || r > h || r > w) return false;   // TMA box limits; every phase has pixels
    G->map_off[i] = i == 0 ? 0 : G->map_off[i - 1] + rates[i - 1] * rates[i - 1];
    int cap = kGatherSmemBudget / img_bytes;
    if (cap > 255) cap = 255;
    // phases per CTA: the batch that keeps the 16 half-warps busiest (items = phases x column segments), larger on ties
    int per = 1;
    double best = 0.0;
    for (int n = 1; n <= cap; ++n) {
      const int items = n * G->nseg[i];
      const double util = static_cast<double>(items) / (16.0 * ceil_div(items, 16));
      if (util >= best) { best = util; per = n; }
    }
    for (int ph = 0; ph < r * r; ph += per) {
      const int n = r * r - ph < per ? r * r - ph : per;
      table->push_back((static_cast<uint32_t>(i) << 28) | (static_cast<uint32_t>(n) << 20) | static_cast<uint32_t>(ph));
      if (i == 0) ++G->pool_slots;
    }
  }
  G->h = h; G->w_ = w; G->C = C; G->nchunks = ceil_div(C, 64);
  G->num_batches = static_cast<int>(table->size());
  return true;
}

// per-phase strided tensor maps of the gather kernel over x [B,h,w,C]: {C, ceil((w-pj)/r), ceil((h-pi)/r), B}, element
// (c, t, a, b) = x[b, pi + r*a, pj + r*t, c]; box {32, ntp, na+2, 1} is one zero-bordered phase image of a 32-channel group
bool encode_gather_maps(const AsppGatherParams& G, const void* x, int B, std::vector<CUtensorMap>* maps, std::string* err) {
  EncodeTiledFn fn = get_encode_fn(err);
  if (!fn) return false;
  maps->clear();
  const uint64_t C = static_cast<uint64_t>(G.C), W = static_cast<uint64_t>(G.w_), H = static_cast<uint64_t>(G.h);
  for (int i = 0; i < 3; ++i) {
    const int r = G.rates[i];
    const uint32_t ntp = static_cast<uint32_t>(G.nseg[i] * G.ts[i] + 2);
    for (int ph = 0; ph < r * r; ++ph) {
      const int pi = ph / r, pj = ph % r;
      CUtensorMap tm;
      cuuint64_t gdim[4] = {C, (W - pj + r - 1) / r, (H - pi + r - 1) / r, static_cast<cuuint64_t>(B)};
      cuuint64_t gstride[3] = {static_cast<cuuint64_t>(r) * C * 2, static_cast<cuuint64_t>(r) * W * C * 2, H * W * C * 2};
      cuuint32_t box[4] = {32, ntp, static_cast<cuuint32_t>(G.na[i] + 2), 1};
      cuuint32_t estr[4] = {1, 1, 1, 1};
      const char* base = static_cast<const char*>(x) + (static_cast<uint64_t>(pi) * W + pj) * C * 2;
      CUresult rc = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<char*>(base), gdim, gstride, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (rc != CUDA_SUCCESS) {
        if (err) *err = fmt("cuTensorMapEncodeTiled(gather rate %d phase %d) -> %d", r, ph, (int)rc);
        return false;
      }
      maps->push_back(tm);
    }
  }
  return true;
}
cudaError_t launch_aspp_gather(const AsppGatherParams& G, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(aspp_dw_gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGatherSmemBytes);
  if (e != cudaSuccess) return e;
  const long long blocks = static_cast<long long>(G.B) * (G.C / 32) * G.num_batches;
  aspp_dw_gather_kernel<<<static_cast<unsigned>(blocks), kGatherThreads, kGatherSmemBytes, st>>>(G);
  return cudaGetLastError();
}

template <int H, int W, int R>
cudaError_t launch_aspp_fast3_t(const AsppDwParams& P, cudaStream_t st) {
  using Cfg = AsppFast3Cfg<H, W, R>;
  cudaError_t e = cudaFuncSetAttribute(aspp_dw_fast3_kernel<H, W, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
  if (e != cudaSuccess) return e;
  aspp_dw_fast3_kernel<H, W, R><<<P.B * (P.C / 32), Cfg::kThreads, Cfg::kSmemBytes, st>>>(P);
  return cudaGetLastError();
}
cudaError_t launch_aspp_fast(const AsppDwParams& P, cudaStream_t st) {
  switch (P.h) {   // one phase image for the three rates (r, 2r, 3r)
    case 32: return launch_aspp_fast3_t<32, 32, 6>(P, st);
    case 24: return launch_aspp_fast3_t<24, 24, 6>(P, st);
    case 16: return launch_aspp_fast3_t<16, 16, 6>(P, st);
    case 8: return launch_aspp_fast3_t<8, 8, 6>(P, st);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_pw(int BN, const PwLaunch& L, int num_sms, cudaStream_t st) {
  if (BN == 256) {
    bool all_bf16 = true;
    for (int i = 0; i < L.num_problems; ++i)   // short-K problems are store bound: the pair brings nothing there (measured)
      all_bf16 = all_bf16 && L.prob[i].epi != kEpiPlanarF32 && L.prob[i].N == 256 && L.prob[i].K >= 512;
    if (all_bf16) return launch_pw2(L, num_sms, st);
  }
  switch (BN) {
    case 32: return launch_pw_t<32>(L, num_sms, st);
    case 64: return launch_pw_t<64>(L, num_sms, st);
    case 256: return launch_pw_t<256>(L, num_sms, st);
    default: return cudaErrorInvalidValue;
  }
}

template <int KB>
cudaError_t launch_dwpw2_t(const DwPw2Params& P2, int num_sms, cudaStream_t st) {
  const DwPwParams& P = P2.base;
  static bool attr_done[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_done[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(dwpw_gemm2_kernel<KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, DwPw2Cfg<KB>::kSmemBytes);
    if (e != cudaSuccess) return e;
    attr_done[dev & 63] = true;
  }
  const int items = (P.num_tiles + 1) / 2;
  const int grid = 2 * items < num_sms ? 2 * items : (num_sms & ~1);
  dwpw_gemm2_kernel<KB><<<grid, kDw2Threads, DwPw2Cfg<KB>::kSmemBytes, st>>>(P2);
  return cudaGetLastError();
}
// fused SepConv_BN on CTA pairs (dwpw_gemm2.cuh).  h_scale / h_shift: HOST copies of the pointwise BN scale / shift (256 each)
// for the kernel's constant-bank epilogue
cudaError_t launch_dwpw(int KB, const DwPwParams& P, const CUtensorMap* tmap_out32, const float* h_scale, const float* h_shift, int num_sms, cudaStream_t st) {
  if (!h_scale || !h_shift || !tmap_out32) return cudaErrorInvalidValue;
  DwPw2Params P2;
  P2.base = P;
  P2.tmap_out32 = tmap_out32;
  std::memcpy(P2.scale_c, h_scale, sizeof(P2.scale_c));
  std::memcpy(P2.shift_c, h_shift, sizeof(P2.shift_c));
  switch (KB) {
    case 1: return launch_dwpw2_t<1>(P2, num_sms, st);
    case 2: return launch_dwpw2_t<2>(P2, num_sms, st);
    case 3: return launch_dwpw2_t<3>(P2, num_sms, st);
    case 4: return launch_dwpw2_t<4>(P2, num_sms, st);
    case 5: return launch_dwpw2_t<5>(P2, num_sms, st);
    default: return cudaErrorInvalidValue;
  }
}

int grid_for(size_t items, int num_sms);

// pred_resize + argmax: integer scales that are multiples of 4 share corners per 4x4 output block (mem_kernels.cuh)
cudaError_t launch_resize_argmax(const ArgmaxParams& P, bool force_generic, int num_sms, cudaStream_t st) {
  const int S = P.hi > 0 ? P.ho / P.hi : 0;
  if (!force_generic && S >= 4 && S % 4 == 0 && P.ho == S * P.hi && P.wo == S * P.wi) {
    ArgmaxIntParams Q{P, S};
    const size_t threads = static_cast<size_t>(P.B) * (P.hi + 1) * (S / 4) * (P.wi + 1) * (S / 4);
    resize_argmax_x4_kernel<<<static_cast<unsigned>((threads + 127) / 128), 128, 0, st>>>(Q);
  } else {
    resize_argmax_generic_kernel<<<grid_for(static_cast<size_t>(P.B) * P.ho * P.wo, num_sms), 256, 0, st>>>(P);
  }
  return cudaGetLastError();
}

int grid_for(size_t items, int num_sms) {
  size_t g = (items + 255) / 256;
  const size_t cap = static_cast<size_t>(num_sms) * 8;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace

// =====================================================================================================
// context
// =====================================================================================================
struct WeightSlot {
  std::string layer, var;
  std::vector<int64_t> shape;
  std::vector<float> data;
  bool set = false;
};

enum TmSlot {
  TM_FEAT = 0, TM_DW1, TM_DW2, TM_DW3, TM_CONCAT, TM_SKIP, TM_DECIN, TM_DEC0, TM_CLS_IN,
  TM_W_ASPP0, TM_W_ASPP1, TM_W_ASPP2, TM_W_ASPP3, TM_W_PROJ, TM_W_FP0, TM_W_DEC0, TM_W_DEC1, TM_W_CLS,
  TM_O_ASPP0, TM_O_ASPP1, TM_O_ASPP2, TM_O_ASPP3, TM_O_PROJ, TM_O_FP0, TM_O_DEC0, TM_O_DEC1, TM_FEAT_SLAB, TM_O4_DEC0, TM_O4_DEC1, TM_O4S_DEC0, TM_O4S_DEC1, TM_DECSKIP,
  TM_COUNT
};

struct PwWeights {      // device-side packed 1x1 conv
  uint16_t* w = nullptr;  // bf16 [Npad][Kpad]
  float* scale = nullptr;
  float* shift = nullptr;
  int K = 0, N = 0, Npad = 0, Kpad = 0;
  std::vector<float> h_scale, h_shift;   // host copies, padded to Npad (kernels that take them as parameters)
};
struct DwWeights {      // device-side packed depthwise conv (fused kernel / standalone)
  float* w = nullptr;      // [9][Cpad]
  float* shift = nullptr;  // [Cpad]
  int C = 0, Cpad = 0;
};

struct dlv3p_ctx {
  dlv3p_config cfg{};
  int device = 0;
  bool plan_only = false;  // device == -1: validation, weight inventory and sizes only (host-side tests)
  int num_sms = 148;
  std::string err;
  // derived
  int h = 0, w = 0, hs = 0, ws = 0, ho = 0, wo = 0;  // ho,wo = classifier resolution
  int M1 = 0, M2 = 0, Mc = 0;                        // pixels at ASPP / decoder / classifier resolution
  int Ccat = 0;                                      // channels of the GEMM part of the concat (1024 / 256)
  int rates[3] = {0, 0, 0};
  bool st_aspp = false, st_dec = false, st_tail = false, lite = false;
  float eps = 1e-5f;

  std::vector<WeightSlot> weights;
  std::map<std::string, int> windex;
  bool finalized = false;

  // device memory
  std::vector<void*> allocs;
  std::vector<void*> weight_allocs;   // uploaded at finalize; released and rebuilt when the weights change (no growth per refresh)
  bool in_finalize = false;
  size_t ws_bytes = 0, weight_bytes = 0;
  __nv_bfloat16 *feat_bf16 = nullptr, *skip_bf16 = nullptr;  // cast targets / forward_host staging
  void *in_feat_stage = nullptr, *in_skip_stage = nullptr;   // forward_host raw staging (in_dtype)
  dlv3p_ctx* pipe[2] = {nullptr, nullptr};                   // forward_host pipeline: two quarter-batch child contexts
  bool pipe_failed = false;
  void* out_stage = nullptr;
  __nv_bfloat16* dw_out = nullptr;     // [3][Cin/64][M1][64]  K-block-major
  float* pool_partial = nullptr;       // [B][nbands][Cin]
  float* img_shift = nullptr;          // [B][256]
  float* b4 = nullptr;                 // [B][256]
  __nv_bfloat16* concat = nullptr;     // [M1][Ccat]
  __nv_bfloat16* aspp_out = nullptr;   // [M1][256]
  __nv_bfloat16* dec_in = nullptr;     // [M2][304]  (unfused decoder only)
  __nv_bfloat16* dec_up = nullptr;     // [M2][256]  upsampled ASPP output   } the decoder concat of the fused path: two tensors,
  __nv_bfloat16* dec_skip = nullptr;   // [M2][64]   projected skip (48 valid) } full-line rows, no 608-byte pitch
  __nv_bfloat16* dec_tmp = nullptr;    // [M2][304]  (unfused path: depthwise output)
  __nv_bfloat16* dec0 = nullptr;       // [M2][256]
  __nv_bfloat16* dec1 = nullptr;       // [M2][256]
  float* logits = nullptr;             // [B][NC][ho*wo] planar
  int nbands = 1, pix_per_band = 256;   // pooling partials: ASPP Lite bands, or rate-0 phase items of the ASPP kernel
  AsppDwParams aspp_plan{};
  bool aspp_slab = false;               // small maps: shared-memory slab kernel
  bool aspp_fast = false;               // geometry-specialised kernel (aspp_dw_fast.cuh)
  bool aspp_gather = false;             // cp.async phase-image gather kernel (aspp_dw_gather.cuh)
  AsppGatherParams gather_plan{};
  std::vector<uint32_t> gather_table;
  uint32_t* gather_table_dev = nullptr;
  CUtensorMap* gather_maps_dev = nullptr;   // per-phase tensor maps over the feature buffer (re-encoded when the pointer changes)
  const void* gather_maps_ptr = nullptr;
  size_t aspp_slab_smem = 0;

  // packed weights
  PwWeights pw_aspp[4], pw_proj, pw_fp0, pw_dec0, pw_dec1, pw_cls;
  DwWeights dw_dec0, dw_dec1;
  float *aspp_dw_w = nullptr, *aspp_dw_shift = nullptr;    // [3][9][Cin], [3][Cin]
  uint32_t* aspp_items = nullptr;                          // packed work-item table of the ASPP depthwise kernels
  uint16_t *w_ip = nullptr, *w_proj4 = nullptr;            // [Cin][256], [256][256] bf16
  float *ip_scale = nullptr, *ip_shift = nullptr;

  // tensor maps
  CUtensorMap h_tm[TM_COUNT];
  CUtensorMap* d_tm = nullptr;
  const void* tm_feat_ptr = nullptr;
  const void* tm_skip_ptr = nullptr;

  cudaStream_t own_stream = nullptr;
  cudaStream_t side_stream = nullptr;   // independent branches of the graph (feature_projection0, pool_proj) run beside the main chain
  cudaEvent_t ev_fork = nullptr, ev_dw = nullptr, ev_pool = nullptr, ev_fp0 = nullptr;
  int64_t launches_last = 0, launches_total = 0;

  // profiling
  bool profiling = false;
  std::vector<cudaEvent_t> prof_events;
  std::vector<const char*> prof_names;
};

namespace {

int fail(dlv3p_ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg;
  g_tls_error = msg;
  return code;
}
#define CU_TRY(c, expr)                                                                              \
  do {                                                                                               \
    cudaError_t _e = (expr);                                                                         \
    if (_e != cudaSuccess) return fail((c), DLV3P_ERR_CUDA, fmt("%s: %s", #expr, cudaGetErrorString(_e))); \
  } while (0)

template <class T>
int dev_alloc(dlv3p_ctx* c, T** p, size_t count) {
  void* q = nullptr;
  size_t bytes = align_up(count * sizeof(T) + 256, 256);
  if (c->plan_only) {
    c->ws_bytes += bytes;
    *p = nullptr;
    return 0;
  }
  cudaError_t e = cudaMalloc(&q, bytes);
  if (e != cudaSuccess) return fail(c, DLV3P_ERR_NOMEM, fmt("cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e)));
  if (c->in_finalize) {
    c->weight_allocs.push_back(q);
    c->weight_bytes += bytes;
  } else {
    c->allocs.push_back(q);
  }
  c->ws_bytes += bytes;
  *p = reinterpret_cast<T*>(q);
  return 0;
}
// device buffers of a previous dlv3p_finalize_weights: the stream work that reads them is drained first
void release_weight_allocs(dlv3p_ctx* c) {
  if (c->weight_allocs.empty()) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  for (void* p : c->weight_allocs) cudaFree(p);
  c->weight_allocs.clear();
  c->ws_bytes -= c->weight_bytes;
  c->weight_bytes = 0;
}
void drop_pipe_children(dlv3p_ctx* c) {
  for (dlv3p_ctx*& ch : c->pipe) {
    if (ch) dlv3p_destroy(ch);
    ch = nullptr;
  }
  c->pipe_failed = false;
}
template <class T>
int upload(dlv3p_ctx* c, T** p, const std::vector<T>& h) {
  int r = dev_alloc(c, p, h.size());
  if (r) return r;
  CU_TRY(c, cudaMemcpy(*p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}

void add_w(dlv3p_ctx* c, const std::string& layer, const std::string& var, std::vector<int64_t> shape) {
  WeightSlot s;
  s.layer = layer;
  s.var = var;
  s.shape = std::move(shape);
  c->windex[layer + "/" + var] = static_cast<int>(c->weights.size());
  c->weights.push_back(std::move(s));
}
void add_conv(dlv3p_ctx* c, const std::string& n, int k, int nn, bool bias = false) {
  add_w(c, n, "kernel", {1, 1, k, nn});
  if (bias) add_w(c, n, "bias", {nn});
}
void add_bn(dlv3p_ctx* c, const std::string& n, int ch) {
  for (const char* v : {"gamma", "beta", "moving_mean", "moving_variance"}) add_w(c, n, v, {ch});
}
void add_sep(dlv3p_ctx* c, const std::string& p, int ch, int nn) {
  add_w(c, p + "_depthwise", "depthwise_kernel", {3, 3, ch, 1});
  add_bn(c, p + "_depthwise_BN", ch);
  add_conv(c, p + "_pointwise", ch, nn);
  add_bn(c, p + "_pointwise_BN", nn);
}

const float* W(const dlv3p_ctx* c, const std::string& layer, const std::string& var) {
  auto it = c->windex.find(layer + "/" + var);
  return it == c->windex.end() ? nullptr : c->weights[it->second].data.data();
}
BnFold fold(const dlv3p_ctx* c, const std::string& bn, int n) {
  return fold_bn(W(c, bn, "gamma"), W(c, bn, "beta"), W(c, bn, "moving_mean"), W(c, bn, "moving_variance"), n, c->eps);
}

int make_pw(dlv3p_ctx* c, PwWeights* pw, const float* w_kn, int Ktot, int N, int k_begin, int k_count,
            const std::vector<float>& scale, const std::vector<float>& shift) {
  pw->K = k_count;
  pw->N = N;
  pw->Npad = pick_bn(N);
  pw->Kpad = ceil_div(k_count, 64) * 64;
  std::vector<uint16_t> packed = pack_pw(w_kn, Ktot, N, k_begin, k_count, pw->Npad, pw->Kpad);
  int r = upload(c, &pw->w, packed);
  if (r) return r;
  std::vector<float> s(pw->Npad, 0.0f), t(pw->Npad, 0.0f);
  for (int i = 0; i < N; ++i) {
    s[i] = scale[i];
    t[i] = shift[i];
  }
  pw->h_scale = s;
  pw->h_shift = t;
  if ((r = upload(c, &pw->scale, s))) return r;
  return upload(c, &pw->shift, t);
}

}  // namespace

// =====================================================================================================
// C ABI
// =====================================================================================================
extern "C" {

int dlv3p_abi_version(void) { return DLV3P_ABI_VERSION; }

const char* dlv3p_last_error(const dlv3p_ctx* ctx) { return ctx ? ctx->err.c_str() : g_tls_error.c_str(); }

int dlv3p_device_count(int* n) {
  if (!n) return fail(nullptr, DLV3P_ERR_INVALID, "null argument");
  cudaError_t e = cudaGetDeviceCount(n);
  if (e != cudaSuccess) {
    *n = 0;
    return fail(nullptr, DLV3P_ERR_CUDA, fmt("cudaGetDeviceCount: %s", cudaGetErrorString(e)));
  }
  return DLV3P_OK;
}
int dlv3p_device_info(int device, int* sm_major, int* sm_minor, int* sm_count, size_t* total_mem) {
  cudaDeviceProp p;
  CU_TRY(nullptr, cudaGetDeviceProperties(&p, device));
  if (sm_major) *sm_major = p.major;
  if (sm_minor) *sm_minor = p.minor;
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (total_mem) *total_mem = p.totalGlobalMem;
  return DLV3P_OK;
}
int dlv3p_dev_alloc(int device, size_t bytes, void** d_ptr) {
  if (!d_ptr) return fail(nullptr, DLV3P_ERR_INVALID, "null argument");
  CU_TRY(nullptr, cudaSetDevice(device));
  cudaError_t e = cudaMalloc(d_ptr, bytes ? bytes : 1);
  if (e != cudaSuccess) return fail(nullptr, DLV3P_ERR_NOMEM, fmt("cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e)));
  return DLV3P_OK;
}
int dlv3p_dev_free(int device, void* d_ptr) {
  CU_TRY(nullptr, cudaSetDevice(device));
  CU_TRY(nullptr, cudaFree(d_ptr));
  return DLV3P_OK;
}
int dlv3p_host_alloc_pinned(size_t bytes, void** h_ptr) {
  if (!h_ptr) return fail(nullptr, DLV3P_ERR_INVALID, "null argument");
  cudaError_t e = cudaMallocHost(h_ptr, bytes ? bytes : 1);
  if (e != cudaSuccess) return fail(nullptr, DLV3P_ERR_NOMEM, fmt("cudaMallocHost(%zu): %s", bytes, cudaGetErrorString(e)));
  return DLV3P_OK;
}
int dlv3p_host_free_pinned(void* h_ptr) {
  CU_TRY(nullptr, cudaFreeHost(h_ptr));
  return DLV3P_OK;
}
int dlv3p_memcpy_h2d(int device, void* d_dst, const void* h_src, size_t bytes) {
  CU_TRY(nullptr, cudaSetDevice(device));
  CU_TRY(nullptr, cudaMemcpy(d_dst, h_src, bytes, cudaMemcpyHostToDevice));
  return DLV3P_OK;
}
int dlv3p_memcpy_d2h(int device, void* h_dst, const void* d_src, size_t bytes) {
  CU_TRY(nullptr, cudaSetDevice(device));
  CU_TRY(nullptr, cudaMemcpy(h_dst, d_src, bytes, cudaMemcpyDeviceToHost));
  return DLV3P_OK;
}
int dlv3p_dev_memset(int device, void* d_ptr, int value, size_t bytes) {
  CU_TRY(nullptr, cudaSetDevice(device));
  CU_TRY(nullptr, cudaMemset(d_ptr, value, bytes));
  return DLV3P_OK;
}
int dlv3p_dev_synchronize(int device) {
  CU_TRY(nullptr, cudaSetDevice(device));
  CU_TRY(nullptr, cudaDeviceSynchronize());
  return DLV3P_OK;
}

// ---------------------------------------------------------------------------------------------- create
int dlv3p_create(const dlv3p_config* cfg, int device, dlv3p_ctx** out) {
  if (!cfg || !out) return fail(nullptr, DLV3P_ERR_INVALID, "null argument");
  *out = nullptr;
  dlv3p_ctx* c = new dlv3p_ctx();
  c->cfg = *cfg;
  c->device = device;
  auto bail = [&](int code, const std::string& m) {
    fail(nullptr, code, m);
    dlv3p_destroy(c);
    return code;
  };

  // ---- validate (reference raises ValueError for a bad OS, layers.py:126)
  const dlv3p_config& g = c->cfg;
  c->lite = g.variant == DLV3P_VARIANT_ASPP_LITE;
  if (g.variant != DLV3P_VARIANT_ASPP && g.variant != DLV3P_VARIANT_ASPP_LITE) return bail(DLV3P_ERR_INVALID, "invalid variant");
  int stages = g.stages;
  if (stages == 0) stages = c->lite ? (DLV3P_STAGE_ASPP | DLV3P_STAGE_TAIL) : (DLV3P_STAGE_ASPP | DLV3P_STAGE_DECODER | DLV3P_STAGE_TAIL);
  c->cfg.stages = stages;
  c->st_aspp = stages & DLV3P_STAGE_ASPP;
  c->st_dec = stages & DLV3P_STAGE_DECODER;
  c->st_tail = stages & DLV3P_STAGE_TAIL;
  if (!(c->st_aspp || c->st_dec || c->st_tail)) return bail(DLV3P_ERR_INVALID, "no stage enabled");
  if (g.B < 1 || g.H < 1 || g.W < 1) return bail(DLV3P_ERR_INVALID, "B, H, W must be positive");
  if (c->st_aspp) {
    if (g.OS == 8) { c->rates[0] = 12; c->rates[1] = 24; c->rates[2] = 36; }
    else if (g.OS == 16) { c->rates[0] = 6; c->rates[1] = 12; c->rates[2] = 18; }
    else if (g.OS == 32) { c->rates[0] = 3; c->rates[1] = 6; c->rates[2] = 9; }
    else return bail(DLV3P_ERR_INVALID, fmt("invalid output stride %d", g.OS));
    if (g.Cin < 8 || g.Cin % 8) return bail(DLV3P_ERR_INVALID, "Cin must be a positive multiple of 8");
  } else if (g.OS < 1) {
    c->cfg.OS = 16;
  }
  if (c->st_dec && (g.Cskip < 8 || g.Cskip % 8)) return bail(DLV3P_ERR_INVALID, "Cskip must be a positive multiple of 8");
  if (c->st_tail && (g.NC < 1 || g.NC > 256)) return bail(DLV3P_ERR_INVALID, "NC must be in 1..256");
  if (g.in_dtype != DLV3P_DTYPE_BF16 && g.in_dtype != DLV3P_DTYPE_FP16 && g.in_dtype != DLV3P_DTYPE_FP32)
    return bail(DLV3P_ERR_UNSUPPORTED, "in_dtype: bf16 (0), fp16 (1) and fp32 (2) are implemented");
  if (c->st_tail) {
    if (g.out_mode < DLV3P_OUT_LABELS_U8 || g.out_mode > DLV3P_OUT_LOGITS_FULL) return bail(DLV3P_ERR_INVALID, "out_mode does not match a TAIL stage");
  } else if (g.out_mode != DLV3P_OUT_FEATURES_BF16 && g.out_mode != DLV3P_OUT_FEATURES_FP32) {
    return bail(DLV3P_ERR_INVALID, "without the TAIL stage out_mode must be FEATURES_BF16/FP32");
  }
  c->eps = g.bn_eps > 0 ? g.bn_eps : 1e-5f;
  c->h = g.h > 0 ? g.h : ceil_div(g.H, c->cfg.OS);
  c->w = g.w > 0 ? g.w : ceil_div(g.W, c->cfg.OS);
  c->hs = g.hs > 0 ? g.hs : ceil_div(g.H, 4);
  c->ws = g.ws > 0 ? g.ws : ceil_div(g.W, 4);
  c->ho = c->st_dec ? c->hs : c->h;
  c->wo = c->st_dec ? c->ws : c->w;
  if (!c->st_aspp && !c->st_dec) { c->ho = c->hs; c->wo = c->ws; }  // TAIL only: input is [B,hs,ws,256]
  c->M1 = g.B * c->h * c->w;
  c->M2 = g.B * c->hs * c->ws;
  c->Mc = g.B * c->ho * c->wo;
  c->Ccat = c->lite ? 256 : 1024;

  // ---- device (device == -1: plan-only context, nothing is allocated and nothing can run)
  c->plan_only = device == -1;
  if (!c->plan_only) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return bail(DLV3P_ERR_CUDA, "no CUDA device: libdlv3p has no CPU fallback");
    if (device < 0 || device >= ndev) return bail(DLV3P_ERR_INVALID, fmt("device %d out of range (%d devices)", device, ndev));
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return bail(DLV3P_ERR_CUDA, "cudaGetDeviceProperties failed");
    if (prop.major != 10) return bail(DLV3P_ERR_UNSUPPORTED, fmt("device sm_%d%d: kernels are sm_100a only", prop.major, prop.minor));
    c->num_sms = prop.multiProcessorCount;
    if (cudaSetDevice(device) != cudaSuccess) return bail(DLV3P_ERR_CUDA, "cudaSetDevice failed");
  }

  // ---- weight inventory in Keras creation order
  if (c->st_aspp) {
    add_conv(c, "image_pooling", g.Cin, 256);
    add_bn(c, "image_pooling_BN", 256);
    add_conv(c, "aspp0", g.Cin, 256);
    add_bn(c, "aspp0_BN", 256);
    if (!c->lite)
      for (int i = 1; i <= 3; ++i) add_sep(c, fmt("aspp%d", i), g.Cin, 256);
    add_conv(c, "concat_projection", c->lite ? 512 : 1280, 256);
    add_bn(c, "concat_projection_BN", 256);
  }
  if (c->st_dec) {
    add_conv(c, "feature_projection0", g.Cskip, 48);
    add_bn(c, "feature_projection0_BN", 48);
    add_sep(c, "decoder_conv0", 304, 256);
    add_sep(c, "decoder_conv1", 256, 256);
  }
  if (c->st_tail) add_conv(c, "conv_upsample", 256, g.NC, true);

  // ---- workspace
  int r = 0;
  const size_t B = g.B;
  if (c->st_aspp) {
    const int px = c->h * c->w;
    if (c->lite) {
      c->pix_per_band = 256;
      c->nbands = ceil_div(px, c->pix_per_band);
    } else {
      AsppDwParams& A = c->aspp_plan;
      A.B = g.B; A.h = c->h; A.w_ = c->w; A.C = g.Cin; A.nrates = 3; A.nchunks = ceil_div(g.Cin, 64);
      A.item_off[0] = 0;
      for (int i = 0; i < 3; ++i) {
        const int rr = c->rates[i];
        A.rates[i] = rr;
        const int nt = ceil_div(c->w, rr);
        static const int kTs[5] = {2, 3, 4, 6, 8};
        int sel = 4;
        for (int k = 4; k >= 0; --k)
          if (kTs[k] >= (nt < 8 ? nt : 8)) sel = k;
        const int na_max = ceil_div(c->h, rr);
        if (na_max <= 2 && nt <= 2) sel = 5;
        else if (na_max <= 3 && nt <= 3) sel = 6;
        A.ts_sel[i] = sel;
        A.nseg[i] = sel >= 5 ? 1 : ceil_div(nt, kTs[sel]);
        A.item_off[i + 1] = A.item_off[i] + rr * rr * A.nseg[i];
      }
      c->aspp_slab_smem = static_cast<size_t>(ceil_div(px, 256)) * 32768 + (27 * 64 + 3 * 64 + 16 * 64) * sizeof(float) + 16;
      c->aspp_slab = c->aspp_slab_smem <= 220 * 1024;
      c->aspp_fast = aspp_fast_supported(A);
      c->aspp_gather = !c->aspp_fast && plan_aspp_gather(c->h, c->w, g.Cin, c->rates, &c->gather_plan, &c->gather_table);
      A.pool_items = c->aspp_gather ? c->gather_plan.pool_slots : (c->aspp_slab ? 1 : A.item_off[1]);
      A.total_warps = static_cast<long long>(g.B) * A.item_off[3] * A.nchunks;
      c->nbands = A.pool_items;
    }
    if (!c->lite) {
      const size_t n = 3 * static_cast<size_t>(ceil_div(g.Cin, 64)) * c->M1 * 64;
      if ((r = dev_alloc(c, &c->dw_out, n))) return bail(r, c->err);
      if (!c->plan_only && cudaMemset(c->dw_out, 0, n * 2) != cudaSuccess) return bail(DLV3P_ERR_CUDA, "cudaMemset failed");   // padded channels stay 0
    }
    if ((r = dev_alloc(c, &c->pool_partial, B * c->nbands * g.Cin))) return bail(r, c->err);
    if ((r = dev_alloc(c, &c->img_shift, B * 256))) return bail(r, c->err);
    if ((r = dev_alloc(c, &c->b4, B * 256))) return bail(r, c->err);
    if ((r = dev_alloc(c, &c->concat, static_cast<size_t>(c->M1) * c->Ccat))) return bail(r, c->err);
    if ((r = dev_alloc(c, &c->aspp_out, static_cast<size_t>(c->M1) * 256))) return bail(r, c->err);
    if (g.in_dtype != DLV3P_DTYPE_BF16 && (r = dev_alloc(c, &c->feat_bf16, static_cast<size_t>(c->M1) * g.Cin))) return bail(r, c->err);
  }
  if (c->st_dec) {
    if (g.flags & DLV3P_FLAG_UNFUSED_DECODER) {
      if ((r = dev_alloc(c, &c->dec_in, static_cast<size_t>(c->M2) * 304))) return bail(r, c->err);
    } else {
      if ((r = dev_alloc(c, &c->dec_up, static_cast<size_t>(c->M2) * 256))) return bail(r, c->err);
      if ((r = dev_alloc(c, &c->dec_skip, static_cast<size_t>(c->M2) * 64))) return bail(r, c->err);
      if (!c->plan_only && cudaMemset(c->dec_skip, 0, static_cast<size_t>(c->M2) * 64 * 2) != cudaSuccess) return bail(DLV3P_ERR_CUDA, "cudaMemset failed");
    }
    if ((g.flags & DLV3P_FLAG_UNFUSED_DECODER) && (r = dev_alloc(c, &c->dec_tmp, static_cast<size_t>(c->M2) * 304))) return bail(r, c->err);
    if ((r = dev_alloc(c, &c->dec0, static_cast<size_t>(c->M2) * 256))) return bail(r, c->err);
    if ((r = dev_alloc(c, &c->dec1, static_cast<size_t>(c->M2) * 256))) return bail(r, c->err);
    if (g.in_dtype != DLV3P_DTYPE_BF16 && (r = dev_alloc(c, &c->skip_bf16, static_cast<size_t>(c->M2) * g.Cskip))) return bail(r, c->err);
    if (!c->st_aspp && g.in_dtype != DLV3P_DTYPE_BF16 && (r = dev_alloc(c, &c->feat_bf16, static_cast<size_t>(c->M1) * 256))) return bail(r, c->err);
  }
  if (c->st_tail) {
    if ((r = dev_alloc(c, &c->logits, B * g.NC * c->ho * c->wo))) return bail(r, c->err);
    if (!c->st_aspp && !c->st_dec && g.in_dtype != DLV3P_DTYPE_BF16 && (r = dev_alloc(c, &c->feat_bf16, static_cast<size_t>(c->Mc) * 256))) return bail(r, c->err);
  }
  if ((r = dev_alloc(c, &c->d_tm, TM_COUNT))) return bail(r, c->err);
  std::memset(c->h_tm, 0, sizeof(c->h_tm));
  if (!c->plan_only && cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess) return bail(DLV3P_ERR_CUDA, "cudaStreamCreate failed");
  if (!c->plan_only) {
    if (cudaStreamCreateWithFlags(&c->side_stream, cudaStreamNonBlocking) != cudaSuccess) return bail(DLV3P_ERR_CUDA, "cudaStreamCreate failed");
    for (cudaEvent_t* ev : {&c->ev_fork, &c->ev_dw, &c->ev_pool, &c->ev_fp0})
      if (cudaEventCreateWithFlags(ev, cudaEventDisableTiming) != cudaSuccess) return bail(DLV3P_ERR_CUDA, "cudaEventCreate failed");
  }
  *out = c;
  return DLV3P_OK;
}

void dlv3p_destroy(dlv3p_ctx* c) {
  if (!c) return;
  if (c->plan_only) {
    delete c;
    return;
  }
  cudaSetDevice(c->device);
  for (void* p : c->allocs) cudaFree(p);
  for (void* p : c->weight_allocs) cudaFree(p);
  if (c->in_feat_stage) cudaFree(c->in_feat_stage);
  if (c->in_skip_stage) cudaFree(c->in_skip_stage);
  if (c->out_stage) cudaFree(c->out_stage);
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  if (c->side_stream) cudaStreamDestroy(c->side_stream);
  for (cudaEvent_t ev : {c->ev_fork, c->ev_dw, c->ev_pool, c->ev_fp0})
    if (ev) cudaEventDestroy(ev);
  for (cudaEvent_t e : c->prof_events) cudaEventDestroy(e);
  for (dlv3p_ctx* ch : c->pipe) dlv3p_destroy(ch);
  delete c;
}

// ---------------------------------------------------------------------------------------------- weights
int dlv3p_num_weights(const dlv3p_ctx* c) { return c ? static_cast<int>(c->weights.size()) : DLV3P_ERR_INVALID; }

int dlv3p_weight_info(const dlv3p_ctx* c, int index, const char** layer, const char** var, int64_t shape_out[4], int* rank_out) {
  if (!c || index < 0 || index >= static_cast<int>(c->weights.size())) return fail(nullptr, DLV3P_ERR_INVALID, "bad weight index");
  const WeightSlot& s = c->weights[index];
  if (layer) *layer = s.layer.c_str();
  if (var) *var = s.var.c_str();
  if (rank_out) *rank_out = static_cast<int>(s.shape.size());
  if (shape_out)
    for (size_t i = 0; i < s.shape.size() && i < 4; ++i) shape_out[i] = s.shape[i];
  return DLV3P_OK;
}

int dlv3p_set_weight(dlv3p_ctx* c, const char* layer, const char* var, const float* host, const int64_t* shape, int rank) {
  if (!c || !layer || !var || !host || !shape) return fail(c, DLV3P_ERR_INVALID, "null argument");
  std::string ln(layer);
  if (ln == "logits_semantic") ln = "conv_upsample";  // raw-constructor name of the classifier (deeplabv3p_xception.py:218)
  auto it = c->windex.find(ln + "/" + var);
  if (it == c->windex.end()) return fail(c, DLV3P_ERR_NAME, fmt("unknown weight %s/%s for this configuration", layer, var));
  WeightSlot& s = c->weights[it->second];
  if (rank != static_cast<int>(s.shape.size())) return fail(c, DLV3P_ERR_NAME, fmt("%s/%s: rank %d, expected %zu", layer, var, rank, s.shape.size()));
  size_t n = 1;
  for (int i = 0; i < rank; ++i) {
    if (shape[i] != s.shape[i]) return fail(c, DLV3P_ERR_NAME, fmt("%s/%s: dim %d is %lld, expected %lld", layer, var, i, (long long)shape[i], (long long)s.shape[i]));
    n *= static_cast<size_t>(shape[i]);
  }
  s.data.assign(host, host + n);
  s.set = true;
  c->finalized = false;
  drop_pipe_children(c);   // dlv3p_forward_host's sub-batch contexts hold copies of the old weights: rebuilt lazily
  return DLV3P_OK;
}

int dlv3p_finalize_weights(dlv3p_ctx* c) {
  if (!c) return fail(nullptr, DLV3P_ERR_INVALID, "null context");
  for (const WeightSlot& s : c->weights)
    if (!s.set) return fail(c, DLV3P_ERR_STATE, fmt("weight %s/%s was never set", s.layer.c_str(), s.var.c_str()));
  if (c->plan_only) return fail(c, DLV3P_ERR_STATE, "plan-only context (device -1): nothing can be uploaded or run; there is no CPU path");
  if (c->finalized) return DLV3P_OK;
  CU_TRY(c, cudaSetDevice(c->device));
  release_weight_allocs(c);
  struct Scope { dlv3p_ctx* c; ~Scope() { c->in_finalize = false; } } scope{c};
  c->in_finalize = true;
  const dlv3p_config& g = c->cfg;
  int r = 0;
  std::string terr;
  if (c->st_aspp) {
    const int Cin = g.Cin;
    // image pooling branch: GEMV weights [Cin][256] bf16 + folded BN
    {
      const float* k = W(c, "image_pooling", "kernel");
      std::vector<uint16_t> wip(static_cast<size_t>(Cin) * 256);
      for (size_t i = 0; i < wip.size(); ++i) wip[i] = f32_to_bf16_rne(k[i]);
      if ((r = upload(c, &c->w_ip, wip))) return r;
      BnFold f = fold(c, "image_pooling_BN", 256);
      if ((r = upload(c, &c->ip_scale, f.scale))) return r;
      if ((r = upload(c, &c->ip_shift, f.shift))) return r;
    }
    {
      BnFold f = fold(c, "aspp0_BN", 256);
      if ((r = make_pw(c, &c->pw_aspp[0], W(c, "aspp0", "kernel"), Cin, 256, 0, Cin, f.scale, f.shift))) return r;
    }
    if (!c->lite) {
      std::vector<float> dw(static_cast<size_t>(3) * 9 * Cin), dsh(static_cast<size_t>(3) * Cin);
      for (int i = 1; i <= 3; ++i) {
        const std::string p = fmt("aspp%d", i);
        BnFold fd = fold(c, p + "_depthwise_BN", Cin);
        std::vector<float> one = pack_dw(W(c, p + "_depthwise", "depthwise_kernel"), fd.scale.data(), Cin, Cin);
        std::memcpy(&dw[static_cast<size_t>(i - 1) * 9 * Cin], one.data(), one.size() * sizeof(float));
        std::memcpy(&dsh[static_cast<size_t>(i - 1) * Cin], fd.shift.data(), Cin * sizeof(float));
        BnFold fp = fold(c, p + "_pointwise_BN", 256);
        if ((r = make_pw(c, &c->pw_aspp[i], W(c, p + "_pointwise", "kernel"), Cin, 256, 0, Cin, fp.scale, fp.shift))) return r;
      }
      if ((r = upload(c, &c->aspp_dw_w, dw))) return r;
      if ((r = upload(c, &c->aspp_items, build_aspp_items(c->aspp_plan)))) return r;
      if (c->aspp_gather && (r = upload(c, &c->gather_table_dev, c->gather_table))) return r;
      if ((r = upload(c, &c->aspp_dw_shift, dsh))) return r;
    }
    {
      // concat_projection: rows 0..255 multiply b4 (image pooling) -> per-image shift; the rest is the GEMM (F9)
      const float* k = W(c, "concat_projection", "kernel");
      std::vector<uint16_t> w4(256 * 256);
      for (size_t i = 0; i < w4.size(); ++i) w4[i] = f32_to_bf16_rne(k[i]);
      if ((r = upload(c, &c->w_proj4, w4))) return r;
      BnFold f = fold(c, "concat_projection_BN", 256);
      if ((r = make_pw(c, &c->pw_proj, k, c->Ccat + 256, 256, 256, c->Ccat, f.scale, f.shift))) return r;
    }
  }
  if (c->st_dec) {
    {
      BnFold f = fold(c, "feature_projection0_BN", 48);
      if ((r = make_pw(c, &c->pw_fp0, W(c, "feature_projection0", "kernel"), g.Cskip, 48, 0, g.Cskip, f.scale, f.shift))) return r;
    }
    struct Sep { const char* p; int C; PwWeights* pw; DwWeights* dw; };
    Sep seps[2] = {{"decoder_conv0", 304, &c->pw_dec0, &c->dw_dec0}, {"decoder_conv1", 256, &c->pw_dec1, &c->dw_dec1}};
    for (const Sep& s : seps) {
      const std::string p(s.p);
      BnFold fd = fold(c, p + "_depthwise_BN", s.C);
      s.dw->C = s.C;
      s.dw->Cpad = ceil_div(s.C, 64) * 64;
      std::vector<float> dw = pack_dw(W(c, p + "_depthwise", "depthwise_kernel"), fd.scale.data(), s.C, s.dw->Cpad);
      std::vector<float> sh(s.dw->Cpad, 0.0f);
      std::memcpy(sh.data(), fd.shift.data(), s.C * sizeof(float));
      if ((r = upload(c, &s.dw->w, dw))) return r;
      if ((r = upload(c, &s.dw->shift, sh))) return r;
      BnFold fp = fold(c, p + "_pointwise_BN", 256);
      if ((r = make_pw(c, s.pw, W(c, p + "_pointwise", "kernel"), s.C, 256, 0, s.C, fp.scale, fp.shift))) return r;
    }
  }
  if (c->st_tail) {
    std::vector<float> ones(g.NC, 1.0f), bias(W(c, "conv_upsample", "bias"), W(c, "conv_upsample", "bias") + g.NC);
    if ((r = make_pw(c, &c->pw_cls, W(c, "conv_upsample", "kernel"), 256, g.NC, 0, 256, ones, bias))) return r;
  }

  // ---- tensor maps over weights and workspace buffers
  auto enc2 = [&](int slot, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
    return encode_2d_sw128(&c->h_tm[slot], base, rows, cols, ld, box_rows, &terr);
  };
  bool ok = true;
  if (c->st_aspp) {
    for (int i = 0; i < (c->lite ? 1 : 4); ++i)
      ok = ok && enc2(TM_W_ASPP0 + i, c->pw_aspp[i].w, c->pw_aspp[i].Npad, c->pw_aspp[i].Kpad, c->pw_aspp[i].Kpad, w_box_rows(c->pw_aspp[i].Npad));
    ok = ok && enc2(TM_W_PROJ, c->pw_proj.w, 256, c->pw_proj.Kpad, c->pw_proj.Kpad, w_box_rows(256));
    if (!c->lite)
      for (int i = 0; i < 3; ++i)
        ok = ok && enc2(TM_DW1 + i, c->dw_out + static_cast<size_t>(i) * ceil_div(g.Cin, 64) * c->M1 * 64, static_cast<uint64_t>(ceil_div(g.Cin, 64)) * c->M1, 64, 64, 128);
    ok = ok && enc2(TM_CONCAT, c->concat, c->M1, c->Ccat, c->Ccat, 128);
    for (int i = 0; i < (c->lite ? 1 : 4); ++i)
      ok = ok && encode_2d_out(&c->h_tm[TM_O_ASPP0 + i], c->concat + 256 * i, c->M1, 256, c->Ccat, &terr);
    ok = ok && encode_2d_out(&c->h_tm[TM_O_PROJ], c->aspp_out, c->M1, 256, 256, &terr);
    if (c->feat_bf16) {
      ok = ok && enc2(TM_FEAT, c->feat_bf16, c->M1, g.Cin, g.Cin, 128);
      ok = ok && encode_2d_slab(&c->h_tm[TM_FEAT_SLAB], c->feat_bf16, c->M1, g.Cin, g.Cin, &terr, c->aspp_fast ? 32 : 64);
      c->tm_feat_ptr = c->feat_bf16;
    }
  }
  if (c->st_dec) {
    ok = ok && enc2(TM_W_FP0, c->pw_fp0.w, c->pw_fp0.Npad, c->pw_fp0.Kpad, c->pw_fp0.Kpad, w_box_rows(c->pw_fp0.Npad));
    ok = ok && enc2(TM_W_DEC0, c->pw_dec0.w, 256, c->pw_dec0.Kpad, c->pw_dec0.Kpad, w_box_rows(256));
    ok = ok && enc2(TM_W_DEC1, c->pw_dec1.w, 256, c->pw_dec1.Kpad, c->pw_dec1.Kpad, w_box_rows(256));
    const bool split_concat = !(g.flags & DLV3P_FLAG_UNFUSED_DECODER);
    ok = ok && (split_concat ? encode_2d_out(&c->h_tm[TM_O_FP0], c->dec_skip, c->M2, 48, 64, &terr)
                             : encode_2d_out(&c->h_tm[TM_O_FP0], c->dec_in + 256, c->M2, 48, 304, &terr));
    ok = ok && encode_2d_out(&c->h_tm[TM_O_DEC0], c->dec0, c->M2, 256, 256, &terr);
    ok = ok && encode_2d_out(&c->h_tm[TM_O_DEC1], c->dec1, c->M2, 256, 256, &terr);
    ok = ok && encode_4d_out(&c->h_tm[TM_O4_DEC0], c->dec0, g.B, c->hs, c->ws, 256, &terr);
    ok = ok && encode_4d_out(&c->h_tm[TM_O4_DEC1], c->dec1, g.B, c->hs, c->ws, 256, &terr);
    ok = ok && encode_4d_out(&c->h_tm[TM_O4S_DEC0], c->dec0, g.B, c->hs, c->ws, 256, &terr, 32);
    ok = ok && encode_4d_out(&c->h_tm[TM_O4S_DEC1], c->dec1, g.B, c->hs, c->ws, 256, &terr, 32);
    if (g.flags & DLV3P_FLAG_UNFUSED_DECODER) {
      ok = ok && enc2(TM_DECIN, c->dec_tmp, c->M2, 304, 304, 128);
      ok = ok && enc2(TM_DEC0, c->dec_tmp, c->M2, 256, 256, 128);
    } else {
      ok = ok && encode_4d_halo(&c->h_tm[TM_DECIN], c->dec_up, g.B, c->hs, c->ws, 256, 256, kDwHaloW, kDwHaloH, &terr);
      ok = ok && encode_4d_halo(&c->h_tm[TM_DECSKIP], c->dec_skip, g.B, c->hs, c->ws, 48, 64, kDwHaloW, kDwHaloH, &terr);
      ok = ok && encode_4d_halo(&c->h_tm[TM_DEC0], c->dec0, g.B, c->hs, c->ws, 256, 256, kDwHaloW, kDwHaloH, &terr);
    }
    if (c->skip_bf16) {
      ok = ok && enc2(TM_SKIP, c->skip_bf16, c->M2, g.Cskip, g.Cskip, 128);
      c->tm_skip_ptr = c->skip_bf16;
    }
  }
  if (c->st_tail) {
    ok = ok && enc2(TM_W_CLS, c->pw_cls.w, c->pw_cls.Npad, c->pw_cls.Kpad, c->pw_cls.Kpad, w_box_rows(c->pw_cls.Npad));
    const __nv_bfloat16* cls_in = c->st_dec ? c->dec1 : (c->st_aspp ? c->aspp_out : c->feat_bf16);
    if (cls_in) ok = ok && enc2(TM_CLS_IN, cls_in, c->Mc, 256, 256, 128);
  }
  if (!ok) return fail(c, DLV3P_ERR_CUDA, terr);
  CU_TRY(c, cudaMemcpy(c->d_tm, c->h_tm, sizeof(c->h_tm), cudaMemcpyHostToDevice));
  c->finalized = true;
  return DLV3P_OK;
}

// ---------------------------------------------------------------------------------------------- sizes
int dlv3p_input_bytes(const dlv3p_ctx* c, size_t* feat_bytes, size_t* skip_bytes) {
  if (!c) return fail(nullptr, DLV3P_ERR_INVALID, "null context");
  const size_t es = c->cfg.in_dtype == DLV3P_DTYPE_FP32 ? 4 : 2;
  size_t f = 0;
  if (c->st_aspp) f = static_cast<size_t>(c->M1) * c->cfg.Cin * es;
  else if (c->st_dec) f = static_cast<size_t>(c->M1) * 256 * es;
  else f = static_cast<size_t>(c->Mc) * 256 * es;
  if (feat_bytes) *feat_bytes = f;
  if (skip_bytes) *skip_bytes = c->st_dec ? static_cast<size_t>(c->M2) * c->cfg.Cskip * es : 0;
  return DLV3P_OK;
}
int dlv3p_output_bytes(const dlv3p_ctx* c, size_t* out_bytes) {
  if (!c || !out_bytes) return fail(nullptr, DLV3P_ERR_INVALID, "null argument");
  const dlv3p_config& g = c->cfg;
  const size_t B = g.B;
  switch (g.out_mode) {
    case DLV3P_OUT_LABELS_U8: *out_bytes = B * g.H * g.W; break;
    case DLV3P_OUT_LOGITS_LOWRES: *out_bytes = B * g.NC * c->ho * c->wo * 4; break;
    case DLV3P_OUT_SOFTMAX:
    case DLV3P_OUT_LOGITS_FULL: *out_bytes = B * g.H * g.W * g.NC * 4; break;
    case DLV3P_OUT_FEATURES_BF16: *out_bytes = static_cast<size_t>(c->st_dec ? c->M2 : c->M1) * 256 * 2; break;
    case DLV3P_OUT_FEATURES_FP32: *out_bytes = static_cast<size_t>(c->st_dec ? c->M2 : c->M1) * 256 * 4; break;
    default: return fail(nullptr, DLV3P_ERR_INVALID, "bad out_mode");
  }
  return DLV3P_OK;
}
int dlv3p_workspace_bytes(const dlv3p_ctx* c, size_t* bytes) {
  if (!c || !bytes) return fail(nullptr, DLV3P_ERR_INVALID, "null argument");
  *bytes = c->ws_bytes;
  return DLV3P_OK;
}
int dlv3p_launch_count(const dlv3p_ctx* c, int64_t* last_forward, int64_t* total) {
  if (!c) return fail(nullptr, DLV3P_ERR_INVALID, "null context");
  if (last_forward) *last_forward = c->launches_last;
  if (total) *total = c->launches_total;
  return DLV3P_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------- forward
namespace {
struct Launcher {
  dlv3p_ctx* c;
  cudaStream_t st;
  int rc = 0;
  bool begin(const char* name) {
    if (rc) return false;
    if (c->profiling) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      cudaEventRecord(e, st);
      c->prof_events.push_back(e);
      c->prof_names.push_back(name);
    }
    return true;
  }
  void end(const char* name, cudaError_t e) {
    if (e != cudaSuccess) rc = fail(c, DLV3P_ERR_CUDA, fmt("launch %s: %s", name, cudaGetErrorString(e)));
    else ++c->launches_last;
  }
};
}  // namespace

static int forward_impl(dlv3p_ctx* c, const void* d_feat, const void* d_skip, void* d_out, cudaStream_t st) {
  if (!c) return fail(nullptr, DLV3P_ERR_INVALID, "null context");
  if (c->plan_only) return fail(c, DLV3P_ERR_STATE, "plan-only context (device -1): there is no CPU path");
  if (!c->finalized) return fail(c, DLV3P_ERR_STATE, "dlv3p_finalize_weights has not been called");
  if (!d_feat || !d_out) return fail(c, DLV3P_ERR_INVALID, "null feature / output pointer");
  if (c->st_dec && !d_skip) return fail(c, DLV3P_ERR_INVALID, "decoder stage needs the skip feature");
  CU_TRY(c, cudaSetDevice(c->device));
  const dlv3p_config& g = c->cfg;
  c->launches_last = 0;
  Launcher L{c, st};
  std::string terr;

  // ---- inputs: cast fp32 -> bf16, or (re)encode the TMA descriptors over the caller's bf16 buffers
  const __nv_bfloat16* feat = nullptr;
  const __nv_bfloat16* skip = nullptr;
  const size_t feat_elems = c->st_aspp ? static_cast<size_t>(c->M1) * g.Cin : (c->st_dec ? static_cast<size_t>(c->M1) * 256 : static_cast<size_t>(c->Mc) * 256);
  if (g.in_dtype != DLV3P_DTYPE_BF16) {   // fp32 (a TF backbone's default) or fp16 (the reference's --mixed_precision): cast on the device
    auto cast = [&](const void* src, __nv_bfloat16* dst, size_t n8) {
      if (L.begin("cast_to_bf16")) {
        if (g.in_dtype == DLV3P_DTYPE_FP32) cast_f32_bf16_kernel<<<grid_for(n8, c->num_sms), 256, 0, st>>>(static_cast<const float*>(src), dst, n8);
        else cast_f16_bf16_kernel<<<grid_for(n8, c->num_sms), 256, 0, st>>>(static_cast<const __half*>(src), dst, n8);
        L.end("cast_to_bf16", cudaGetLastError());
      }
    };
    cast(d_feat, c->feat_bf16, feat_elems / 8);
    feat = c->feat_bf16;
    if (c->st_dec) {
      cast(d_skip, c->skip_bf16, static_cast<size_t>(c->M2) * g.Cskip / 8);
      skip = c->skip_bf16;
    }
  } else {
    feat = static_cast<const __nv_bfloat16*>(d_feat);
    skip = static_cast<const __nv_bfloat16*>(d_skip);
    if (c->st_aspp && c->tm_feat_ptr != d_feat) {
      if (!encode_2d_sw128(&c->h_tm[TM_FEAT], feat, c->M1, g.Cin, g.Cin, 128, &terr)) return fail(c, DLV3P_ERR_CUDA, terr);
      if (!encode_2d_slab(&c->h_tm[TM_FEAT_SLAB], feat, c->M1, g.Cin, g.Cin, &terr, c->aspp_fast ? 32 : 64)) return fail(c, DLV3P_ERR_CUDA, terr);
      CU_TRY(c, cudaMemcpyAsync(&c->d_tm[TM_FEAT], &c->h_tm[TM_FEAT], sizeof(CUtensorMap), cudaMemcpyHostToDevice, st));
      CU_TRY(c, cudaMemcpyAsync(&c->d_tm[TM_FEAT_SLAB], &c->h_tm[TM_FEAT_SLAB], sizeof(CUtensorMap), cudaMemcpyHostToDevice, st));
      c->tm_feat_ptr = d_feat;
    }
    if (c->st_dec && c->tm_skip_ptr != d_skip) {
      if (!encode_2d_sw128(&c->h_tm[TM_SKIP], skip, c->M2, g.Cskip, g.Cskip, 128, &terr)) return fail(c, DLV3P_ERR_CUDA, terr);
      CU_TRY(c, cudaMemcpyAsync(&c->d_tm[TM_SKIP], &c->h_tm[TM_SKIP], sizeof(CUtensorMap), cudaMemcpyHostToDevice, st));
      c->tm_skip_ptr = d_skip;
    }
    if (!c->st_aspp && !c->st_dec && c->tm_feat_ptr != d_feat) {  // TAIL only
      if (!encode_2d_sw128(&c->h_tm[TM_CLS_IN], feat, c->Mc, 256, 256, 128, &terr)) return fail(c, DLV3P_ERR_CUDA, terr);
      CU_TRY(c, cudaMemcpyAsync(&c->d_tm[TM_CLS_IN], &c->h_tm[TM_CLS_IN], sizeof(CUtensorMap), cudaMemcpyHostToDevice, st));
      c->tm_feat_ptr = d_feat;
    }
  }

  // independent branches on the side stream (not while profiling kernel by kernel): feature_projection0 only needs the
  // skip feature, pool_proj only the pooling partial sums
  const bool side = c->side_stream && !c->profiling && c->st_aspp && c->st_dec && !c->lite;
  auto launch_fp0 = [&](cudaStream_t s_) {
    if (L.begin("feature_projection0_gemm")) {
      PwLaunch PL{};
      PL.num_problems = 1; PL.M = c->M2; PL.num_tiles = ceil_div(c->M2, kPwBM); PL.rows_per_img = c->hs * c->ws;
      PwProblem& p = PL.prob[0];
      p.tmap_a = &c->d_tm[TM_SKIP]; p.tmap_w = &c->d_tm[TM_W_FP0]; p.tmap_out = &c->d_tm[TM_O_FP0];
      p.scale = c->pw_fp0.scale; p.shift = c->pw_fp0.shift; p.img_shift = nullptr;
      const bool split = !(g.flags & DLV3P_FLAG_UNFUSED_DECODER);
      p.out = split ? c->dec_skip : c->dec_in; p.K = g.Cskip; p.N = 48; p.ldo = split ? 64 : 304; p.col_off = split ? 0 : 256; p.relu = 1; p.epi = kEpiBf16;
      L.end("feature_projection0_gemm", launch_pw(64, PL, c->num_sms, s_));
    }
  };
  if (side) {
    if (cudaEventRecord(c->ev_fork, st) != cudaSuccess || cudaStreamWaitEvent(c->side_stream, c->ev_fork, 0) != cudaSuccess)
      return fail(c, DLV3P_ERR_CUDA, "side stream fork failed");
    launch_fp0(c->side_stream);
    if (cudaEventRecord(c->ev_fp0, c->side_stream) != cudaSuccess) return fail(c, DLV3P_ERR_CUDA, "cudaEventRecord failed");
  }
  const __nv_bfloat16* x256 = feat;  // running 256-channel feature map
  // ------------------------------------------------------------------ ASPP (layers.py:114-196)
  if (c->st_aspp) {
    if (c->lite) {
      if (L.begin("global_pool")) {
        PoolParams P{};
        P.x = feat; P.pool_partial = c->pool_partial; P.npix = c->h * c->w; P.C = g.Cin; P.nbands = c->nbands; P.pix_per_band = c->pix_per_band;
        dim3 grid(ceil_div(g.Cin, 64), c->nbands, g.B);
        global_pool_kernel<<<grid, 256, 0, st>>>(P);
        L.end("global_pool", cudaGetLastError());
      }
    } else if (L.begin("aspp_dw_pool")) {
      AsppDwParams P = c->aspp_plan;
      P.x = feat; P.w = c->aspp_dw_w; P.shift = c->aspp_dw_shift; P.out = c->dw_out; P.pool_partial = c->pool_partial;
      P.tmap_slab = &c->d_tm[TM_FEAT_SLAB]; P.item_table = c->aspp_items;
      if (c->aspp_fast) {
        cudaError_t e = launch_aspp_fast(P, st);
        (void)e;
      } else if (c->aspp_gather) {
        AsppGatherParams G = c->gather_plan;
        G.x = feat; G.w = c->aspp_dw_w; G.shift = c->aspp_dw_shift; G.out = c->dw_out; G.pool_partial = c->pool_partial;
        G.batches = c->gather_table_dev; G.B = g.B; G.debug = (g.flags >> 8) & 7;   // measurement aid (tools/kbench_gather.py)
        if (c->gather_maps_ptr != feat) {   // per-phase tensor maps follow the feature buffer (one-time for a fixed buffer)
          std::vector<CUtensorMap> maps;
          if (!encode_gather_maps(G, feat, g.B, &maps, &terr)) return fail(c, DLV3P_ERR_CUDA, terr);
          if (!c->gather_maps_dev && dev_alloc(c, &c->gather_maps_dev, maps.size())) return DLV3P_ERR_NOMEM;
          CU_TRY(c, cudaMemcpyAsync(c->gather_maps_dev, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice, st));
          CU_TRY(c, cudaStreamSynchronize(st));   // the host vector goes away
          c->gather_maps_ptr = feat;
        }
        G.maps = c->gather_maps_dev;
        cudaError_t e = launch_aspp_gather(G, st);
        (void)e;
      } else if (c->aspp_slab) {
        cudaError_t e = cudaFuncSetAttribute(aspp_dw_slab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(c->aspp_slab_smem));
        const int slabs = g.B * P.nchunks;
        if (e == cudaSuccess) aspp_dw_slab_kernel<<<slabs < c->num_sms ? slabs : c->num_sms, kSlabThreads, c->aspp_slab_smem, st>>>(P);
      } else {
        const long long blocks = (P.total_warps + 3) / 4;
        aspp_dw_phase_kernel<<<static_cast<unsigned>(blocks), 128, 0, st>>>(P);
      }
      L.end("aspp_dw_pool", cudaGetLastError());
    }
    bool pool_on_side = false;
    if (L.begin("pool_proj")) {
      PoolProjParams P{};
      P.pool_partial = c->pool_partial; P.w_ip = reinterpret_cast<const __nv_bfloat16*>(c->w_ip);
      P.ip_scale = c->ip_scale; P.ip_shift = c->ip_shift; P.w_proj4 = reinterpret_cast<const __nv_bfloat16*>(c->w_proj4);
      P.proj_scale = c->pw_proj.scale; P.proj_shift = c->pw_proj.shift; P.img_shift = c->img_shift; P.b4_out = c->b4;
      P.C = g.Cin; P.nbands = c->nbands; P.inv_count = 1.0f / static_cast<float>(c->h * c->w);
      cudaStream_t ps = st;
      if (side && cudaEventRecord(c->ev_dw, st) == cudaSuccess && cudaStreamWaitEvent(c->side_stream, c->ev_dw, 0) == cudaSuccess) ps = c->side_stream;
      pool_proj_kernel<<<g.B, 1024, (g.Cin + 256 + 2048) * sizeof(float), ps>>>(P);
      L.end("pool_proj", cudaGetLastError());
      pool_on_side = ps != st;
      if (pool_on_side && cudaEventRecord(c->ev_pool, ps) != cudaSuccess) L.rc = fail(c, DLV3P_ERR_CUDA, "cudaEventRecord failed");
    }
    if (L.begin("aspp_branches_gemm")) {
      PwLaunch PL{};
      PL.num_problems = c->lite ? 1 : 4;
      PL.M = c->M1; PL.num_tiles = ceil_div(c->M1, kPwBM); PL.rows_per_img = c->h * c->w;
      for (int i = 0; i < PL.num_problems; ++i) {
        PwProblem& p = PL.prob[i];
        p.tmap_a = &c->d_tm[i == 0 ? TM_FEAT : TM_DW1 + (i - 1)];
        p.tmap_w = &c->d_tm[TM_W_ASPP0 + i];
        p.tmap_out = &c->d_tm[TM_O_ASPP0 + i];
        p.scale = c->pw_aspp[i].scale; p.shift = c->pw_aspp[i].shift; p.img_shift = nullptr;
        p.out = c->concat; p.K = g.Cin; p.N = 256; p.ldo = c->Ccat; p.col_off = 256 * i; p.relu = 1; p.epi = kEpiBf16;
        p.a_kblock_rows = i == 0 ? 0 : c->M1;
      }
      L.end("aspp_branches_gemm", launch_pw(256, PL, c->num_sms, st));
    }
    if (pool_on_side && !L.rc && cudaStreamWaitEvent(st, c->ev_pool, 0) != cudaSuccess) L.rc = fail(c, DLV3P_ERR_CUDA, "side stream join failed");
    if (L.begin("concat_projection_gemm")) {
      PwLaunch PL{};
      PL.num_problems = 1; PL.M = c->M1; PL.num_tiles = ceil_div(c->M1, kPwBM); PL.rows_per_img = c->h * c->w;
      PwProblem& p = PL.prob[0];
      p.tmap_a = &c->d_tm[TM_CONCAT]; p.tmap_w = &c->d_tm[TM_W_PROJ]; p.tmap_out = &c->d_tm[TM_O_PROJ];
      p.scale = c->pw_proj.scale; p.shift = c->pw_proj.shift; p.img_shift = c->img_shift;
      p.out = c->aspp_out; p.K = c->Ccat; p.N = 256; p.ldo = 256; p.col_off = 0; p.relu = 1; p.epi = kEpiBf16ImgShift;
      L.end("concat_projection_gemm", launch_pw(256, PL, c->num_sms, st));
    }
    x256 = c->aspp_out;
  }
  // ------------------------------------------------------------------ Decoder (layers.py:199-219)
  if (c->st_dec) {
    if (L.begin("decoder_resize")) {
      ResizeParams P{};
      const bool split = !(g.flags & DLV3P_FLAG_UNFUSED_DECODER);
      P.x = x256; P.out = split ? c->dec_up : c->dec_in; P.B = g.B; P.hi = c->h; P.wi = c->w; P.C = 256; P.ho = c->hs; P.wo = c->ws;
      P.ldo = split ? 256 : 304; P.col_off = 0;
      P.sy = static_cast<float>(c->h) / static_cast<float>(c->hs); P.sx = static_cast<float>(c->w) / static_cast<float>(c->ws);
      if (c->hs == 4 * c->h && c->ws == 4 * c->w)
        resize_bilinear_x4_kernel<<<grid_for(static_cast<size_t>(g.B) * (c->h + 1) * (c->w + 1) * 32, c->num_sms), 256, 0, st>>>(P);
      else
        resize_bilinear_kernel<<<dim3(c->hs, g.B), 256, 0, st>>>(P);
      L.end("decoder_resize", cudaGetLastError());
    }
    if (!side) launch_fp0(st);
    else if (!L.rc && cudaStreamWaitEvent(st, c->ev_fp0, 0) != cudaSuccess) L.rc = fail(c, DLV3P_ERR_CUDA, "cudaStreamWaitEvent failed");
    const int tiles_x = ceil_div(c->ws, kDwTW), tiles_y = ceil_div(c->hs, kDwTH);
    struct SepRun { const char* name; const __nv_bfloat16* in; int C; int tm_x; int tm_w; int tm_o; DwWeights* dw; PwWeights* pw; __nv_bfloat16* out; };
    SepRun runs[2] = {{"decoder_conv0_sepconv", c->dec_in, 304, TM_DECIN, TM_W_DEC0, TM_O_DEC0, &c->dw_dec0, &c->pw_dec0, c->dec0},
                      {"decoder_conv1_sepconv", c->dec0, 256, TM_DEC0, TM_W_DEC1, TM_O_DEC1, &c->dw_dec1, &c->pw_dec1, c->dec1}};
    for (const SepRun& s : runs) {
      if (g.flags & DLV3P_FLAG_UNFUSED_DECODER) {
        if (L.begin("decoder_depthwise")) {
          DwParams P{};
          P.x = s.in; P.w = s.dw->w; P.shift = s.dw->shift; P.out = c->dec_tmp; P.B = g.B; P.H = c->hs; P.W = c->ws; P.C = s.C;
          P.rate = 1; P.relu = 1; P.wstride = s.dw->Cpad;
          depthwise3x3_kernel<<<grid_for(static_cast<size_t>(c->M2) * (s.C / 8), c->num_sms), 256, 0, st>>>(P);
          L.end("decoder_depthwise", cudaGetLastError());
        }
        if (L.begin("decoder_pointwise_gemm")) {
          PwLaunch PL{};
          PL.num_problems = 1; PL.M = c->M2; PL.num_tiles = ceil_div(c->M2, kPwBM); PL.rows_per_img = c->hs * c->ws;
          PwProblem& p = PL.prob[0];
          p.tmap_a = &c->d_tm[s.tm_x]; p.tmap_w = &c->d_tm[s.tm_w]; p.tmap_out = &c->d_tm[s.tm_o];
          p.scale = s.pw->scale; p.shift = s.pw->shift; p.img_shift = nullptr;
          p.out = s.out; p.K = s.C; p.N = 256; p.ldo = 256; p.col_off = 0; p.relu = 1; p.epi = kEpiBf16;
          L.end("decoder_pointwise_gemm", launch_pw(256, PL, c->num_sms, st));
        }
      } else if (L.begin(s.name)) {
        DwPwParams P{};
        P.tmap_x = &c->d_tm[s.tm_x]; P.tmap_w = &c->d_tm[s.tm_w];
        if (s.tm_x == TM_DECIN) { P.tmap_x2 = &c->d_tm[TM_DECSKIP]; P.kb_split = 4; }   // conv0: K blocks 0-3 upsampled ASPP, block 4 projected skip
        P.dw_w = s.dw->w; P.dw_shift = s.dw->shift; P.scale = s.pw->scale; P.shift = s.pw->shift; P.out = s.out;
        P.tmap_out = &c->d_tm[s.tm_o == TM_O_DEC0 ? TM_O4_DEC0 : TM_O4_DEC1];
        P.B = g.B; P.H = c->hs; P.W = c->ws; P.tiles_x = tiles_x; P.tiles_y = tiles_y; P.num_tiles = g.B * tiles_x * tiles_y;
        L.end(s.name, launch_dwpw(s.dw->Cpad / 64, P, &c->d_tm[s.tm_o == TM_O_DEC0 ? TM_O4S_DEC0 : TM_O4S_DEC1], s.pw->h_scale.data(), s.pw->h_shift.data(), c->num_sms, st));
      }
    }
    x256 = c->dec1;
  }
  // ------------------------------------------------------------------ tail (model.py:75-86, deeplab.py:99)
  if (c->st_tail) {
    if (L.begin("classifier_gemm")) {
      PwLaunch PL{};
      PL.num_problems = 1; PL.M = c->Mc; PL.num_tiles = ceil_div(c->Mc, kPwBM); PL.rows_per_img = c->ho * c->wo;
      PwProblem& p = PL.prob[0];
      p.tmap_a = &c->d_tm[TM_CLS_IN]; p.tmap_w = &c->d_tm[TM_W_CLS];
      p.scale = c->pw_cls.scale; p.shift = c->pw_cls.shift; p.img_shift = nullptr;
      p.out = g.out_mode == DLV3P_OUT_LOGITS_LOWRES ? d_out : c->logits;
      p.K = 256; p.N = g.NC; p.ldo = 0; p.col_off = 0; p.relu = 0; p.epi = kEpiPlanarF32;
      L.end("classifier_gemm", launch_pw(c->pw_cls.Npad, PL, c->num_sms, st));
    }
    const float sy = static_cast<float>(c->ho) / static_cast<float>(g.H), sx = static_cast<float>(c->wo) / static_cast<float>(g.W);
    if (g.out_mode == DLV3P_OUT_LABELS_U8) {
      if (L.begin("resize_argmax")) {
        ArgmaxParams P{};
        P.logits = c->logits; P.labels = static_cast<uint8_t*>(d_out); P.B = g.B; P.NC = g.NC; P.hi = c->ho; P.wi = c->wo;
        P.ho = g.H; P.wo = g.W; P.sy = sy; P.sx = sx;
        L.end("resize_argmax", launch_resize_argmax(P, false, c->num_sms, st));
      }
    } else if (g.out_mode == DLV3P_OUT_SOFTMAX || g.out_mode == DLV3P_OUT_LOGITS_FULL) {
      if (L.begin("resize_dense")) {
        DenseResizeParams P{};
        P.logits = c->logits; P.out = static_cast<float*>(d_out); P.B = g.B; P.NC = g.NC; P.hi = c->ho; P.wi = c->wo;
        P.ho = g.H; P.wo = g.W; P.softmax = g.out_mode == DLV3P_OUT_SOFTMAX; P.sy = sy; P.sx = sx;
        resize_dense_kernel<<<grid_for(static_cast<size_t>(g.B) * g.H * g.W, c->num_sms), 256, 0, st>>>(P);
        L.end("resize_dense", cudaGetLastError());
      }
    }
  } else {
    const size_t n = static_cast<size_t>(c->st_dec ? c->M2 : c->M1) * 256;
    if (g.out_mode == DLV3P_OUT_FEATURES_BF16) {
      if (!L.rc) CU_TRY(c, cudaMemcpyAsync(d_out, x256, n * 2, cudaMemcpyDeviceToDevice, st));
    } else if (L.begin("cast_bf16_f32")) {
      cast_bf16_f32_kernel<<<grid_for(n / 8, c->num_sms), 256, 0, st>>>(x256, static_cast<float*>(d_out), n / 8);
      L.end("cast_bf16_f32", cudaGetLastError());
    }
  }
  if (c->profiling) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    c->prof_events.push_back(e);
  }
  c->launches_total += c->launches_last;
  return L.rc;
}

extern "C" {

int dlv3p_forward(dlv3p_ctx* c, const void* d_feat, const void* d_skip, void* d_out, void* cuda_stream) {
  return forward_impl(c, d_feat, d_skip, d_out, static_cast<cudaStream_t>(cuda_stream));
}

// forward_host pipeline: the step is PCIe bound (inputs are ~50x the device time), so the batch goes through in
// kPipeChunks sub-batches on two child contexts (own stream, own staging, own workspace, same weights): the H2D copy of
// chunk i+1 overlaps the kernels of chunk i and the D2H of chunk i-1.  Images are independent, so the result is the
// same as one full-batch forward (tests/test_gpu_2_head.py).
constexpr int kPipeChunks = 4;
static int make_pipe_children(dlv3p_ctx* c) {
  dlv3p_config cfg = c->cfg;
  cfg.B = c->cfg.B / kPipeChunks;
  for (int k = 0; k < 2; ++k) {
    dlv3p_ctx* ch = nullptr;
    int r = dlv3p_create(&cfg, c->device, &ch);
    if (r) return r;
    c->pipe[k] = ch;
    for (const WeightSlot& w : c->weights)
      if ((r = dlv3p_set_weight(ch, w.layer.c_str(), w.var.c_str(), w.data.data(), w.shape.data(), static_cast<int>(w.shape.size())))) return r;
    if ((r = dlv3p_finalize_weights(ch))) return r;
    size_t fb = 0, sb = 0, ob = 0;
    dlv3p_input_bytes(ch, &fb, &sb);
    if ((r = dlv3p_output_bytes(ch, &ob))) return r;
    if (cudaMalloc(&ch->in_feat_stage, fb) != cudaSuccess || (sb && cudaMalloc(&ch->in_skip_stage, sb) != cudaSuccess) ||
        cudaMalloc(&ch->out_stage, ob) != cudaSuccess)
      return DLV3P_ERR_NOMEM;
  }
  return DLV3P_OK;
}

int dlv3p_forward_host(dlv3p_ctx* c, const void* h_feat, const void* h_skip, void* h_out) {
  if (!c || !h_feat || !h_out) return fail(c, DLV3P_ERR_INVALID, "null argument");
  if (c->plan_only) return fail(c, DLV3P_ERR_STATE, "plan-only context (device -1): there is no CPU path");
  if (!c->finalized) return fail(c, DLV3P_ERR_STATE, "dlv3p_finalize_weights has not been called");
  CU_TRY(c, cudaSetDevice(c->device));
  size_t fb = 0, sb = 0, ob = 0;
  dlv3p_input_bytes(c, &fb, &sb);
  int r = dlv3p_output_bytes(c, &ob);
  if (r) return r;
  if (sb && !h_skip) return fail(c, DLV3P_ERR_INVALID, "decoder stage needs the skip feature");

  if (!c->pipe_failed && c->cfg.B >= 2 * kPipeChunks && c->cfg.B % kPipeChunks == 0) {
    if (!c->pipe[0] && make_pipe_children(c) != DLV3P_OK) {   // e.g. out of memory: fall back to the one-shot path below
      for (dlv3p_ctx*& ch : c->pipe) { dlv3p_destroy(ch); ch = nullptr; }
      c->pipe_failed = true;
      CU_TRY(c, cudaSetDevice(c->device));
    }
    if (c->pipe[0]) {
      const size_t cf = fb / kPipeChunks, cs = sb / kPipeChunks, co = ob / kPipeChunks;
      for (int i = 0; i < kPipeChunks; ++i) {
        dlv3p_ctx* ch = c->pipe[i & 1];
        cudaStream_t st = ch->own_stream;   // chunk i-2 on this child is ordered before chunk i by the stream
        CU_TRY(c, cudaMemcpyAsync(ch->in_feat_stage, static_cast<const char*>(h_feat) + i * cf, cf, cudaMemcpyHostToDevice, st));
        if (sb) CU_TRY(c, cudaMemcpyAsync(ch->in_skip_stage, static_cast<const char*>(h_skip) + i * cs, cs, cudaMemcpyHostToDevice, st));
        r = forward_impl(ch, ch->in_feat_stage, sb ? ch->in_skip_stage : nullptr, ch->out_stage, st);
        if (r) return fail(c, r, ch->err);
        CU_TRY(c, cudaMemcpyAsync(static_cast<char*>(h_out) + i * co, ch->out_stage, co, cudaMemcpyDeviceToHost, st));
      }
      CU_TRY(c, cudaStreamSynchronize(c->pipe[0]->own_stream));
      CU_TRY(c, cudaStreamSynchronize(c->pipe[1]->own_stream));
      c->launches_last = kPipeChunks * c->pipe[0]->launches_last;
      c->launches_total += c->launches_last;
      return DLV3P_OK;
    }
  }
  if (!c->in_feat_stage) CU_TRY(c, cudaMalloc(&c->in_feat_stage, fb));
  if (sb && !c->in_skip_stage) CU_TRY(c, cudaMalloc(&c->in_skip_stage, sb));
  if (!c->out_stage) CU_TRY(c, cudaMalloc(&c->out_stage, ob));
  cudaStream_t st = c->own_stream;
  CU_TRY(c, cudaMemcpyAsync(c->in_feat_stage, h_feat, fb, cudaMemcpyHostToDevice, st));
  if (sb) CU_TRY(c, cudaMemcpyAsync(c->in_skip_stage, h_skip, sb, cudaMemcpyHostToDevice, st));
  r = forward_impl(c, c->in_feat_stage, sb ? c->in_skip_stage : nullptr, c->out_stage, st);
  if (r) return r;
  CU_TRY(c, cudaMemcpyAsync(h_out, c->out_stage, ob, cudaMemcpyDeviceToHost, st));
  CU_TRY(c, cudaStreamSynchronize(st));
  return DLV3P_OK;
}

int dlv3p_profile_forward(dlv3p_ctx* c, const void* d_feat, const void* d_skip, void* d_out, void* cuda_stream,
                          const char** names_out, float* ms_out, int max) {
  if (!c) return fail(nullptr, DLV3P_ERR_INVALID, "null context");
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  for (cudaEvent_t e : c->prof_events) cudaEventDestroy(e);
  c->prof_events.clear();
  c->prof_names.clear();
  c->profiling = true;
  int r = forward_impl(c, d_feat, d_skip, d_out, st);
  c->profiling = false;
  if (r) return r;
  CU_TRY(c, cudaStreamSynchronize(st));
  const int n = static_cast<int>(c->prof_names.size());
  for (int i = 0; i < n && i < max; ++i) {
    float ms = 0;
    cudaEventElapsedTime(&ms, c->prof_events[i], c->prof_events[i + 1]);
    if (names_out) names_out[i] = c->prof_names[i];
    if (ms_out) ms_out[i] = ms;
  }
  return n < max ? n : max;
}

int dlv3p_read_tap(dlv3p_ctx* c, const char* name, float* host_out, size_t host_elems) {
  if (!c || !name || !host_out) return fail(c, DLV3P_ERR_INVALID, "null argument");
  if (c->plan_only) return fail(c, DLV3P_ERR_STATE, "plan-only context (device -1): there is no CPU path");
  CU_TRY(c, cudaSetDevice(c->device));
  CU_TRY(c, cudaDeviceSynchronize());
  const std::string n(name);
  const void* src = nullptr;
  size_t elems = 0;
  bool is_f32 = false;
  if (n == "aspp_out" && c->aspp_out) { src = c->aspp_out; elems = static_cast<size_t>(c->M1) * 256; }
  else if (n == "concat" && c->concat) { src = c->concat; elems = static_cast<size_t>(c->M1) * c->Ccat; }
  else if (n == "aspp_depthwise" && c->dw_out) {
    // device layout is K-block-major [3][Cin/64][M1][64]; hand back [3][M1][Cin]
    const int nch = ceil_div(c->cfg.Cin, 64), Cin = c->cfg.Cin;
    const size_t total = static_cast<size_t>(3) * nch * c->M1 * 64;
    if (host_elems < static_cast<size_t>(3) * c->M1 * Cin) return fail(c, DLV3P_ERR_INVALID, "tap buffer too small");
    std::vector<uint16_t> tmp(total);
    CU_TRY(c, cudaMemcpy(tmp.data(), c->dw_out, total * 2, cudaMemcpyDeviceToHost));
    for (int r3 = 0; r3 < 3; ++r3)
      for (int k = 0; k < nch; ++k)
        for (int m = 0; m < c->M1; ++m)
          for (int j = 0; j < 64 && k * 64 + j < Cin; ++j)
            host_out[(static_cast<size_t>(r3) * c->M1 + m) * Cin + k * 64 + j] = bf16_to_f32(tmp[((static_cast<size_t>(r3) * nch + k) * c->M1 + m) * 64 + j]);
    return DLV3P_OK;
  }
  else if (n == "decoder_in" && c->dec_in) { src = c->dec_in; elems = static_cast<size_t>(c->M2) * 304; }
  else if (n == "decoder_in" && c->dec_up) {   // fused path: the concat exists as two tensors
    elems = static_cast<size_t>(c->M2) * 304;
    if (host_elems < elems) return fail(c, DLV3P_ERR_INVALID, "tap buffer too small");
    std::vector<uint16_t> up(static_cast<size_t>(c->M2) * 256), sk(static_cast<size_t>(c->M2) * 64);
    CU_TRY(c, cudaMemcpy(up.data(), c->dec_up, up.size() * 2, cudaMemcpyDeviceToHost));
    CU_TRY(c, cudaMemcpy(sk.data(), c->dec_skip, sk.size() * 2, cudaMemcpyDeviceToHost));
    for (int m = 0; m < c->M2; ++m) {
      for (int j = 0; j < 256; ++j) host_out[static_cast<size_t>(m) * 304 + j] = bf16_to_f32(up[static_cast<size_t>(m) * 256 + j]);
      for (int j = 0; j < 48; ++j) host_out[static_cast<size_t>(m) * 304 + 256 + j] = bf16_to_f32(sk[static_cast<size_t>(m) * 64 + j]);
    }
    return DLV3P_OK;
  }
  else if (n == "decoder_conv0" && c->dec0) { src = c->dec0; elems = static_cast<size_t>(c->M2) * 256; }
  else if (n == "decoder_out" && c->dec1) { src = c->dec1; elems = static_cast<size_t>(c->M2) * 256; }
  else if (n == "logits" && c->logits) { src = c->logits; elems = static_cast<size_t>(c->cfg.B) * c->cfg.NC * c->ho * c->wo; is_f32 = true; }
  else if (n == "image_pooling" && c->b4) { src = c->b4; elems = static_cast<size_t>(c->cfg.B) * 256; is_f32 = true; }
  else return fail(c, DLV3P_ERR_NAME, fmt("no tap named %s in this configuration", name));
  if (host_elems < elems) return fail(c, DLV3P_ERR_INVALID, fmt("tap %s needs %zu elements, buffer has %zu", name, elems, host_elems));
  if (is_f32) {
    CU_TRY(c, cudaMemcpy(host_out, src, elems * 4, cudaMemcpyDeviceToHost));
  } else {
    std::vector<uint16_t> tmp(elems);
    CU_TRY(c, cudaMemcpy(tmp.data(), src, elems * 2, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < elems; ++i) host_out[i] = bf16_to_f32(tmp[i]);
  }
  return DLV3P_OK;
}

}  // extern "C"

// =====================================================================================================
// standalone operators (parity tests drive the same kernels one at a time)
// =====================================================================================================
namespace {
struct TmpDev {
  std::vector<void*> ptrs;
  ~TmpDev() {
    for (void* p : ptrs) cudaFree(p);
  }
  template <class T>
  T* put(const std::vector<T>& h) {
    void* p = nullptr;
    if (cudaMalloc(&p, h.size() * sizeof(T) + 256) != cudaSuccess) return nullptr;
    ptrs.push_back(p);
    cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
    return static_cast<T*>(p);
  }
  template <class T>
  T* alloc(size_t n) {
    void* p = nullptr;
    if (cudaMalloc(&p, n * sizeof(T) + 256) != cudaSuccess) return nullptr;
    ptrs.push_back(p);
    return static_cast<T*>(p);
  }
};
int op_prolog(int device, int* num_sms) {
  static int cached_sms[64] = {};   // per-device SM count (cudaGetDeviceProperties costs ~100 us per call)
  CU_TRY(nullptr, cudaSetDevice(device));
  if (device < 0 || device >= 64 || !cached_sms[device]) {
    cudaDeviceProp p;
    CU_TRY(nullptr, cudaGetDeviceProperties(&p, device));
    if (p.major != 10) return fail(nullptr, DLV3P_ERR_UNSUPPORTED, fmt("device sm_%d%d: kernels are sm_100a only", p.major, p.minor));
    if (device < 0 || device >= 64) { *num_sms = p.multiProcessorCount; return 0; }
    cached_sms[device] = p.multiProcessorCount;
  }
  *num_sms = cached_sms[device];
  return 0;
}
}  // namespace

extern "C" {

int dlv3p_op_pointwise(int device, const void* a_bf16, int64_t M, int K, int N, const float* w_kn, const float* scale,
                       const float* shift, int relu, void* out_bf16, void* cuda_stream) {
  int sms = 0, r = op_prolog(device, &sms);
  if (r) return r;
  if (!a_bf16 || !w_kn || !out_bf16 || M < 1 || K < 8 || K % 8 || N < 8 || N % 8 || N > 256) return fail(nullptr, DLV3P_ERR_INVALID, "op_pointwise: bad arguments (K%8, N%8, N<=256)");
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  const int BN = pick_bn(N), Kpad = ceil_div(K, 64) * 64;
  TmpDev tmp;
  uint16_t* dw = tmp.put(pack_pw(w_kn, K, N, 0, K, BN, Kpad));
  std::vector<float> s(BN, 0.0f), t(BN, 0.0f);
  for (int i = 0; i < N; ++i) { s[i] = scale ? scale[i] : 1.0f; t[i] = shift ? shift[i] : 0.0f; }
  float* ds = tmp.put(s);
  float* dt = tmp.put(t);
  std::string terr;
  std::vector<CUtensorMap> tm(3);
  if (!encode_2d_sw128(&tm[0], a_bf16, M, K, K, 128, &terr) || !encode_2d_sw128(&tm[1], dw, BN, Kpad, Kpad, w_box_rows(BN), &terr) ||
      !encode_2d_out(&tm[2], out_bf16, M, N, N, &terr)) return fail(nullptr, DLV3P_ERR_CUDA, terr);
  CUtensorMap* dtm = tmp.put(tm);
  if (!dw || !ds || !dt || !dtm) return fail(nullptr, DLV3P_ERR_NOMEM, "op_pointwise: cudaMalloc failed");
  PwLaunch PL{};
  PL.num_problems = 1; PL.M = static_cast<int>(M); PL.num_tiles = ceil_div(static_cast<int>(M), kPwBM); PL.rows_per_img = static_cast<int>(M);
  PwProblem& p = PL.prob[0];
  p.tmap_a = &dtm[0]; p.tmap_w = &dtm[1]; p.tmap_out = &dtm[2]; p.scale = ds; p.shift = dt; p.img_shift = nullptr; p.out = out_bf16;
  p.K = K; p.N = N; p.ldo = N; p.col_off = 0; p.relu = relu; p.epi = kEpiBf16;
  CU_TRY(nullptr, launch_pw(BN, PL, sms, st));
  CU_TRY(nullptr, cudaStreamSynchronize(st));
  return DLV3P_OK;
}

int dlv3p_op_depthwise(int device, const void* x_bf16, int B, int H, int W_, int C, int rate, const float* w_hwc,
                       const float* scale, const float* shift, int relu, void* out_bf16, void* cuda_stream) {
  int sms = 0, r = op_prolog(device, &sms);
  if (r) return r;
  if (!x_bf16 || !w_hwc || !out_bf16 || B < 1 || H < 1 || W_ < 1 || C < 8 || C % 8 || rate < 1) return fail(nullptr, DLV3P_ERR_INVALID, "op_depthwise: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  TmpDev tmp;
  float* dw = tmp.put(pack_dw(w_hwc, scale, C, C));
  std::vector<float> sh(C, 0.0f);
  if (shift) sh.assign(shift, shift + C);
  float* dsh = tmp.put(sh);
  if (!dw || !dsh) return fail(nullptr, DLV3P_ERR_NOMEM, "op_depthwise: cudaMalloc failed");
  DwParams P{};
  P.x = static_cast<const __nv_bfloat16*>(x_bf16); P.w = dw; P.shift = dsh; P.out = static_cast<__nv_bfloat16*>(out_bf16);
  P.B = B; P.H = H; P.W = W_; P.C = C; P.rate = rate; P.relu = relu; P.wstride = C;
  depthwise3x3_kernel<<<grid_for(static_cast<size_t>(B) * H * W_ * (C / 8), sms), 256, 0, st>>>(P);
  CU_TRY(nullptr, cudaGetLastError());
  CU_TRY(nullptr, cudaStreamSynchronize(st));
  return DLV3P_OK;
}

int dlv3p_op_sepconv(int device, const void* x_bf16, int B, int H, int W_, int C, int rate, const float* dw_hwc,
                     const float* dw_scale, const float* dw_shift, int N, const float* pw_kn, const float* pw_scale,
                     const float* pw_shift, void* out_bf16, void* cuda_stream) {
  int sms = 0, r = op_prolog(device, &sms);
  if (r) return r;
  if (!x_bf16 || !dw_hwc || !pw_kn || !out_bf16 || C < 8 || C % 8) return fail(nullptr, DLV3P_ERR_INVALID, "op_sepconv: bad arguments");
  if (rate != 1 || N != 256 || C > 320) return fail(nullptr, DLV3P_ERR_UNSUPPORTED, "op_sepconv: fused kernel covers rate 1, N = 256, C <= 320");
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  const int KB = ceil_div(C, 64), Cpad = KB * 64;
  TmpDev tmp;
  float* ddw = tmp.put(pack_dw(dw_hwc, dw_scale, C, Cpad));
  std::vector<float> sh(Cpad, 0.0f);
  if (dw_shift) std::memcpy(sh.data(), dw_shift, C * sizeof(float));
  float* dsh = tmp.put(sh);
  uint16_t* dpw = tmp.put(pack_pw(pw_kn, C, N, 0, C, 256, Cpad));
  std::vector<float> s(256, 1.0f), t(256, 0.0f);
  if (pw_scale) s.assign(pw_scale, pw_scale + 256);
  if (pw_shift) t.assign(pw_shift, pw_shift + 256);
  float* ds = tmp.put(s);
  float* dt = tmp.put(t);
  std::string terr;
  std::vector<CUtensorMap> tm(4);
  if (!encode_4d_halo(&tm[0], x_bf16, B, H, W_, C, C, kDwHaloW, kDwHaloH, &terr) || !encode_2d_sw128(&tm[1], dpw, 256, Cpad, Cpad, w_box_rows(256), &terr) ||
      !encode_4d_out(&tm[2], out_bf16, B, H, W_, 256, &terr) || !encode_4d_out(&tm[3], out_bf16, B, H, W_, 256, &terr, 32))
    return fail(nullptr, DLV3P_ERR_CUDA, terr);
  CUtensorMap* dtm = tmp.put(tm);
  if (!ddw || !dsh || !dpw || !ds || !dt || !dtm) return fail(nullptr, DLV3P_ERR_NOMEM, "op_sepconv: cudaMalloc failed");
  DwPwParams P{};
  P.tmap_x = &dtm[0]; P.tmap_w = &dtm[1]; P.tmap_out = &dtm[2]; P.dw_w = ddw; P.dw_shift = dsh; P.scale = ds; P.shift = dt;
  P.out = static_cast<__nv_bfloat16*>(out_bf16); P.B = B; P.H = H; P.W = W_;
  P.tiles_x = ceil_div(W_, kDwTW); P.tiles_y = ceil_div(H, kDwTH); P.num_tiles = B * P.tiles_x * P.tiles_y;
  CU_TRY(nullptr, launch_dwpw(KB, P, &dtm[3], s.data(), t.data(), sms, st));
  CU_TRY(nullptr, cudaStreamSynchronize(st));
  return DLV3P_OK;
}

int dlv3p_op_resize_bilinear(int device, const void* x_bf16, int B, int hi, int wi, int C, int ho, int wo, void* out_bf16,
                             void* cuda_stream) {
  int sms = 0, r = op_prolog(device, &sms);
  if (r) return r;
  if (!x_bf16 || !out_bf16 || C < 8 || C % 8 || hi < 1 || wi < 1 || ho < 1 || wo < 1) return fail(nullptr, DLV3P_ERR_INVALID, "op_resize_bilinear: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  ResizeParams P{};
  P.x = static_cast<const __nv_bfloat16*>(x_bf16); P.out = static_cast<__nv_bfloat16*>(out_bf16);
  P.B = B; P.hi = hi; P.wi = wi; P.C = C; P.ho = ho; P.wo = wo; P.ldo = C; P.col_off = 0;
  P.sy = static_cast<float>(hi) / static_cast<float>(ho); P.sx = static_cast<float>(wi) / static_cast<float>(wo);
  if (ho == 4 * hi && wo == 4 * wi)
    resize_bilinear_x4_kernel<<<grid_for(static_cast<size_t>(B) * (hi + 1) * (wi + 1) * 32, sms), 256, 0, st>>>(P);
  else
    resize_bilinear_kernel<<<dim3(ho, B), 256, 0, st>>>(P);
  CU_TRY(nullptr, cudaGetLastError());
  CU_TRY(nullptr, cudaStreamSynchronize(st));
  return DLV3P_OK;
}

int dlv3p_op_confusion_matrix(int device, const uint8_t* d_pred, const uint8_t* d_gt, int64_t n, int NC, unsigned long long* d_confusion,
                              void* cuda_stream) {
  int sms = 0, r = op_prolog(device, &sms);
  if (r) return r;
  if (!d_pred || !d_gt || !d_confusion || n < 1 || NC < 1 || NC > 256) return fail(nullptr, DLV3P_ERR_INVALID, "op_confusion_matrix: bad arguments (1 <= NC <= 256)");
  if ((reinterpret_cast<uintptr_t>(d_pred) | reinterpret_cast<uintptr_t>(d_gt)) & 15) return fail(nullptr, DLV3P_ERR_INVALID, "op_confusion_matrix: label buffers must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  ConfusionParams P{d_pred, d_gt, n, NC, d_confusion};
  const size_t smem = NC <= 64 ? static_cast<size_t>(NC) * NC * sizeof(unsigned int) : 0;
  confusion_matrix_kernel<<<grid_for(static_cast<size_t>(n / 16 + 1), sms), 256, smem, st>>>(P);
  CU_TRY(nullptr, cudaGetLastError());
  return DLV3P_OK;   // asynchronous; d_confusion is ACCUMULATED (zero it once per evaluation)
}

int dlv3p_op_jaccard_counts(int device, const uint8_t* d_pred, const uint8_t* d_gt, int B, int64_t n_per_image, int NC, unsigned long long* d_counts,
                            void* cuda_stream) {
  int sms = 0, r = op_prolog(device, &sms);
  if (r) return r;
  if (!d_pred || !d_gt || !d_counts || B < 1 || B > 65535 || n_per_image < 1 || NC < 1 || NC > 254)
    return fail(nullptr, DLV3P_ERR_INVALID, "op_jaccard_counts: bad arguments (1 <= NC <= 254)");
  int bpi = static_cast<int>((n_per_image + 256 * 16 - 1) / (256 * 16));
  const int cap = (8 * sms + B - 1) / B;
  if (bpi > cap) bpi = cap;
  if (bpi < 1) bpi = 1;
  jaccard_counts_kernel<<<dim3(bpi, B), 256, 3 * (NC + 1) * sizeof(unsigned int), static_cast<cudaStream_t>(cuda_stream)>>>(d_pred, d_gt, n_per_image, NC, d_counts);
  CU_TRY(nullptr, cudaGetLastError());
  return DLV3P_OK;   // asynchronous; d_counts is ACCUMULATED
}

// ---- image pre / post-processing of the demo and evaluation loops
int dlv3p_op_normalize_image(int device, const uint8_t* d_img, int64_t n, void* d_out, int out_bf16, void* cuda_stream) {
  int sms = 0, r = op_prolog(device, &sms);
  if (r) return r;
  if (!d_img || !d_out || n < 1) return fail(nullptr, DLV3P_ERR_INVALID, "op_normalize_image: bad arguments");
  normalize_image_kernel<<<grid_for(static_cast<size_t>(n), sms), 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(
      d_img, static_cast<size_t>(n), out_bf16 ? nullptr : static_cast<float*>(d_out), out_bf16 ? static_cast<__nv_bfloat16*>(d_out) : nullptr);
  CU_TRY(nullptr, cudaGetLastError());
  return DLV3P_OK;
}
int dlv3p_op_denormalize_image(int device, const float* d_img, int64_t n, uint8_t* d_out, void* cuda_stream) {
  int sms = 0, r = op_prolog(device, &sms);
  if (r) return r;
  if (!d_img || !d_out || n < 1) return fail(nullptr, DLV3P_ERR_INVALID, "op_denormalize_image: bad arguments");
  denormalize_image_kernel<<<grid_for(static_cast<size_t>(n), sms), 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(d_img, static_cast<size_t>(n), d_out);
  CU_TRY(nullptr, cudaGetLastError());
  return DLV3P_OK;
}
int dlv3p_op_mask_resize_nearest(int device, const uint8_t* d_mask, int B, int hi, int wi, int ho, int wo, uint8_t* d_out, void* cuda_stream) {
  int sms = 0, r = op_prolog(device, &sms);
  if (r) return r;
  if (!d_mask || !d_out || B < 1 || hi < 1 || wi < 1 || ho < 1 || wo < 1) return fail(nullptr, DLV3P_ERR_INVALID, "op_mask_resize_nearest: bad arguments");
  // OpenCV resizeNN: inv_scale = (double)dsize / ssize; ifx = 1. / inv_scale
  const double ifx = 1.0 / (static_cast<double>(wo) / wi), ify = 1.0 / (static_cast<double>(ho) / hi);
  mask_resize_nearest_kernel<<<grid_for(static_cast<size_t>(B) * ho * wo, sms), 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(d_mask, B, hi, wi, ho, wo, ify, ifx, d_out);
  CU_TRY(nullptr, cudaGetLastError());
  return DLV3P_OK;
}

// Pillow's precompute_coeffs + normalize_coeffs_8bpc for the bicubic filter (Resample.c), in double like the original
namespace {
double pil_bicubic(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}
int pil_coeffs(int in_size, int out_size, std::vector<int>* bounds, std::vector<int>* kk) {
  const double scale = static_cast<double>(in_size) / out_size;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = 2.0 * filterscale;
  const int ksize = static_cast<int>(std::ceil(support)) * 2 + 1;
  bounds->assign(static_cast<size_t>(out_size) * 2, 0);
  kk->assign(static_cast<size_t>(out_size) * ksize, 0);
  std::vector<double> k(ksize);
  const double ss = 1.0 / filterscale;
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale;
    double ww = 0.0;
    int xmin = static_cast<int>(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = static_cast<int>(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    for (int x = 0; x < xmax; ++x) {
      const double w = pil_bicubic((x + xmin - center + 0.5) * ss);
      k[x] = w;
      ww += w;
    }
    for (int x = 0; x < xmax; ++x) {
      if (ww != 0.0) k[x] /= ww;
      const double v = k[x] * (1 << 22);
      (*kk)[static_cast<size_t>(xx) * ksize + x] = v < 0 ? static_cast<int>(-0.5 + v) : static_cast<int>(0.5 + v);
    }
    (*bounds)[2 * xx] = xmin;
    (*bounds)[2 * xx + 1] = xmax;
  }
  return ksize;
}
}  // namespace

int dlv3p_pil_bicubic_coeffs(int in_size, int out_size, int* bounds, int* kk, int kk_capacity, int* ksize) {
  if (in_size < 1 || out_size < 1 || !bounds || !kk || !ksize) return fail(nullptr, DLV3P_ERR_INVALID, "pil_bicubic_coeffs: bad arguments");
  std::vector<int> b, k;
  *ksize = pil_coeffs(in_size, out_size, &b, &k);
  if (static_cast<size_t>(kk_capacity) < k.size()) return fail(nullptr, DLV3P_ERR_INVALID, "pil_bicubic_coeffs: kk_capacity < out_size * ksize");
  std::memcpy(bounds, b.data(), b.size() * sizeof(int));
  std::memcpy(kk, k.data(), k.size() * sizeof(int));
  return DLV3P_OK;
}

int dlv3p_op_resize_bicubic_u8(int device, const uint8_t* d_img, int B, int H, int W, int C, int ho, int wo, uint8_t* d_out, void* cuda_stream) {
  int sms = 0, r = op_prolog(device, &sms);
  if (r) return r;
  if (!d_img || !d_out || B < 1 || H < 1 || W < 1 || C < 1 || ho < 1 || wo < 1) return fail(nullptr, DLV3P_ERR_INVALID, "op_resize_bicubic_u8: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  const bool horiz = wo != W, vert = ho != H;      // Pillow skips a pass whose size does not change
  if (!horiz && !vert) {
    CU_TRY(nullptr, cudaMemcpyAsync(d_out, d_img, static_cast<size_t>(B) * H * W * C, cudaMemcpyDeviceToDevice, st));
    return DLV3P_OK;
  }
  std::vector<int> bx, kx, by, ky;
  const int ksx = horiz ? pil_coeffs(W, wo, &bx, &kx) : 0, ksy = vert ? pil_coeffs(H, ho, &by, &ky) : 0;
  const size_t tab_ints = bx.size() + kx.size() + by.size() + ky.size();
  const size_t tmp_bytes = horiz && vert ? static_cast<size_t>(B) * H * wo * C : 0;
  int* d_tab = nullptr;
  uint8_t* d_tmp = nullptr;
  CU_TRY(nullptr, cudaMalloc(&d_tab, tab_ints * sizeof(int)));
  if (tmp_bytes && cudaMalloc(&d_tmp, tmp_bytes) != cudaSuccess) {
    cudaFree(d_tab);
    return fail(nullptr, DLV3P_ERR_NOMEM, "op_resize_bicubic_u8: cudaMalloc of the intermediate image failed");
  }
  std::vector<int> tab;
  tab.reserve(tab_ints);
  tab.insert(tab.end(), bx.begin(), bx.end());      // the {first, count} pairs first: both are read as int2 (even sizes keep them aligned)
  tab.insert(tab.end(), by.begin(), by.end());
  tab.insert(tab.end(), kx.begin(), kx.end());
  tab.insert(tab.end(), ky.begin(), ky.end());
  cudaError_t e = cudaMemcpyAsync(d_tab, tab.data(), tab_ints * sizeof(int), cudaMemcpyHostToDevice, st);
  const int* d_bx = d_tab;
  const int* d_by = d_bx + bx.size();
  const int* d_kx = d_by + by.size();
  const int* d_ky = d_kx + kx.size();
  if (e == cudaSuccess && horiz) {      // rows are the outer axis, pixels the resampled one, channels inner
    uint8_t* dst = vert ? d_tmp : d_out;
    resample_pass_u8_kernel<<<grid_for(static_cast<size_t>(B) * H * wo * C, sms), 256, 0, st>>>(d_img, dst, static_cast<long long>(B) * H, W, wo, C,
                                                                                                 reinterpret_cast<const int2*>(d_bx), d_kx, ksx);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess && vert) {       // images outer, rows resampled, a whole row of pixels x channels inner
    const uint8_t* src = horiz ? d_tmp : d_img;
    resample_pass_u8_kernel<<<grid_for(static_cast<size_t>(B) * ho * wo * C, sms), 256, 0, st>>>(src, d_out, B, H, ho, wo * C, reinterpret_cast<const int2*>(d_by),
                                                                                                  d_ky, ksy);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);      // the tables and the intermediate live only for this call
  cudaFree(d_tab);
  if (d_tmp) cudaFree(d_tmp);
  CU_TRY(nullptr, e);
  return DLV3P_OK;
}

int dlv3p_op_present_classes(int device, const uint8_t* d_labels, int B, int64_t n_per_image, unsigned int* d_first, void* cuda_stream) {
  int sms = 0, r = op_prolog(device, &sms);
  if (r) return r;
  if (!d_labels || !d_first || B < 1 || B > 65535 || n_per_image < 1 || n_per_image > 0xFFFFFFFEll)
    return fail(nullptr, DLV3P_ERR_INVALID, "op_present_classes: bad arguments (n_per_image < 2^32 - 1)");
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  CU_TRY(nullptr, cudaMemsetAsync(d_first, 0xFF, static_cast<size_t>(B) * 256 * sizeof(unsigned int), st));
  int bpi = static_cast<int>((n_per_image + 256 * 16 - 1) / (256 * 16));
  const int cap = (8 * sms + B - 1) / B;
  if (bpi > cap) bpi = cap;
  if (bpi < 1) bpi = 1;
  present_classes_kernel<<<dim3(bpi, B), 256, 0, st>>>(d_labels, n_per_image, d_first);
  CU_TRY(nullptr, cudaGetLastError());
  return DLV3P_OK;   // asynchronous
}

size_t dlv3p_op_bn_scratch_bytes(int C) {
  const size_t a = static_cast<size_t>(kBnBands) * 2 * (C > 0 ? C : 0), b = col_scratch_floats(C, 2);
  return (a > b ? a : b) * sizeof(float);
}

int dlv3p_op_bn_stats(int device, const void* x_bf16, int64_t M, int C, float* d_stats, void* d_scratch, void* cuda_stream) {
  int sms = 0, r = op_prolog(device, &sms);
  if (r) return r;
  if (!x_bf16 || !d_stats || !d_scratch || M < 1 || C < 2 || C % 2) return fail(nullptr, DLV3P_ERR_INVALID, "op_bn_stats: bad arguments (C even)");
  if (M >= (1ll << 24)) return fail(nullptr, DLV3P_ERR_UNSUPPORTED, "op_bn_stats: the row count travels as fp32 (M < 2^24 per replica)");
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  if (C % 8 == 0) {   // 16-byte loads, ~4 blocks per SM in flight
    const int bands = col_bands(C);
    bn_stats_vec_kernel<<<dim3(ceil_div(C, 256), bands), 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x_bf16), M, C, bands, static_cast<float*>(d_scratch));
    bands_final_kernel<<<ceil_div(2 * C, 32), dim3(32, kFinalRows), 0, st>>>(static_cast<const float*>(d_scratch), bands, 2 * C, d_stats, 2 * C, static_cast<float>(M));
  } else {
    bn_stats_partial_kernel<<<dim3(ceil_div(C, 64), kBnBands), 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x_bf16), M, C, static_cast<float*>(d_scratch));
    bn_stats_final_kernel<<<ceil_div(2 * C, 256), 256, 0, st>>>(static_cast<const float*>(d_scratch), M, C, d_stats);
  }
  CU_TRY(nullptr, cudaGetLastError());
  return DLV3P_OK;   // asynchronous: the caller all-reduces d_stats on the same stream / after an event
}

int dlv3p_op_bn_apply(int device, const void* x_bf16, int64_t M, int C, const float* d_stats, const float* d_gamma, const float* d_beta,
                      float eps, int relu, void* y_bf16, void* cuda_stream) {
  int sms = 0, r = op_prolog(device, &sms);
  if (r) return r;
  if (!x_bf16 || !d_stats || !d_gamma || !d_beta || !y_bf16 || M < 1 || C < 8 || C % 8) return fail(nullptr, DLV3P_ERR_INVALID, "op_bn_apply: bad arguments (C % 8)");
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  if (C > 4096) return fail(nullptr, DLV3P_ERR_UNSUPPORTED, "op_bn_apply: C <= 4096");
  bn_apply_vec_kernel<unsigned long long><<<grid_for(static_cast<size_t>(M) * (C / 8), sms), 256, 2 * C * sizeof(float), st>>>(static_cast<const __nv_bfloat16*>(x_bf16), M, C, d_stats, d_gamma,
                                                                                                          d_beta, eps, relu, static_cast<__nv_bfloat16*>(y_bf16), C);
  CU_TRY(nullptr, cudaGetLastError());
  return DLV3P_OK;
}

int dlv3p_op_resize_argmax(int device, const float* logits_planar, int B, int NC, int hi, int wi, int ho, int wo,
                           uint8_t* labels, void* cuda_stream) {
  int sms = 0, r = op_prolog(device, &sms);
  if (r) return r;
  if (!logits_planar || !labels || NC < 1 || NC > 256) return fail(nullptr, DLV3P_ERR_INVALID, "op_resize_argmax: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  ArgmaxParams P{};
  P.logits = logits_planar; P.labels = labels; P.B = B; P.NC = NC; P.hi = hi; P.wi = wi; P.ho = ho; P.wo = wo;
  P.sy = static_cast<float>(hi) / static_cast<float>(ho); P.sx = static_cast<float>(wi) / static_cast<float>(wo);
  CU_TRY(nullptr, launch_resize_argmax(P, false, sms, st));
  CU_TRY(nullptr, cudaStreamSynchronize(st));
  return DLV3P_OK;
}


// Benchmark aid (tools/kbench.py): ms per launch of one operator on synthetic device data, CUDA events, `iters` launches.
int dlv3p_op_time(int device, int op, const int64_t* d, int ndims, int iters, int flags, float* ms_out) {
  int sms = 0, r = op_prolog(device, &sms);
  if (r) return r;
  if (!d || !ms_out || iters < 1) return fail(nullptr, DLV3P_ERR_INVALID, "op_time: bad arguments");
  TmpDev tmp;
  cudaStream_t st = nullptr;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  std::string terr;
  auto fill = [&](size_t n) -> uint16_t* {   // bf16 pattern data (small finite values)
    std::vector<uint16_t> h(n);
    uint32_t sd = 12345u;
    for (size_t i = 0; i < n; ++i) { sd = sd * 1664525u + 1013904223u; h[i] = f32_to_bf16_rne(static_cast<float>((sd >> 16) & 0xFF) / 128.0f - 1.0f); }
    return tmp.put(h);
  };
  auto run = [&](auto&& launch) -> int {
    for (int i = 0; i < 2; ++i) launch();
    if (cudaStreamSynchronize(st) != cudaSuccess) return fail(nullptr, DLV3P_ERR_CUDA, fmt("op_time warmup: %s", cudaGetErrorString(cudaGetLastError())));
    cudaEventRecord(e0, st);
    for (int i = 0; i < iters; ++i) launch();
    cudaEventRecord(e1, st);
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return fail(nullptr, DLV3P_ERR_CUDA, fmt("op_time: %s", cudaGetErrorString(e)));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    *ms_out = ms / iters;
    return DLV3P_OK;
  };
  int rc = DLV3P_ERR_INVALID;
  if (op == 0 && ndims >= 3) {   // pointwise {M,K,N}
    const int M = static_cast<int>(d[0]), K = static_cast<int>(d[1]), N = static_cast<int>(d[2]);
    const int BN = pick_bn(N), Kpad = ceil_div(K, 64) * 64;
    uint16_t* a = fill(static_cast<size_t>(M) * K);
    uint16_t* w = fill(static_cast<size_t>(BN) * Kpad);
    uint16_t* o = tmp.alloc<uint16_t>(static_cast<size_t>(M) * N);
    float* s = tmp.put(std::vector<float>(BN, 1.0f));
    float* t = tmp.put(std::vector<float>(BN, 0.0f));
    if (N % 8) return fail(nullptr, DLV3P_ERR_INVALID, "op_time pointwise: N % 8");
    std::vector<CUtensorMap> tm(3);
    if (!encode_2d_sw128(&tm[0], a, M, K, K, 128, &terr) || !encode_2d_sw128(&tm[1], w, BN, Kpad, Kpad, w_box_rows(BN), &terr) ||
        !encode_2d_out(&tm[2], o, M, N, N, &terr)) return fail(nullptr, DLV3P_ERR_CUDA, terr);
    CUtensorMap* dtm = tmp.put(tm);
    PwLaunch PL{};
    PL.num_problems = 1; PL.M = M; PL.num_tiles = ceil_div(M, kPwBM); PL.rows_per_img = M; PL.debug = flags;
    PwProblem& p = PL.prob[0];
    p.tmap_a = &dtm[0]; p.tmap_w = &dtm[1]; p.tmap_out = &dtm[2]; p.scale = s; p.shift = t; p.out = o; p.K = K; p.N = N; p.ldo = N; p.relu = 1; p.epi = kEpiBf16;
    rc = run([&] { launch_pw(BN, PL, sms, st); });
  } else if (op == 1 && ndims >= 4) {   // fused sepconv {B,H,W,C}
    const int B = static_cast<int>(d[0]), H = static_cast<int>(d[1]), W_ = static_cast<int>(d[2]), C = static_cast<int>(d[3]);
    const int KB = ceil_div(C, 64), Cpad = KB * 64;
    if (KB > 5) return fail(nullptr, DLV3P_ERR_UNSUPPORTED, "op_time sepconv: C <= 320");
    uint16_t* x = fill(static_cast<size_t>(B) * H * W_ * C);
    uint16_t* w = fill(static_cast<size_t>(256) * Cpad);
    uint16_t* o = tmp.alloc<uint16_t>(static_cast<size_t>(B) * H * W_ * 256);
    float* dw = tmp.put(std::vector<float>(static_cast<size_t>(9) * Cpad, 0.1f));
    float* dsh = tmp.put(std::vector<float>(Cpad, 0.0f));
    const std::vector<float> hs(256, 1.0f), ht(256, 0.0f);
    float* s = tmp.put(hs);
    float* t = tmp.put(ht);
    std::vector<CUtensorMap> tm(4);
    if (!encode_4d_halo(&tm[0], x, B, H, W_, C, C, kDwHaloW, kDwHaloH, &terr) || !encode_2d_sw128(&tm[1], w, 256, Cpad, Cpad, w_box_rows(256), &terr) ||
        !encode_4d_out(&tm[2], o, B, H, W_, 256, &terr) || !encode_4d_out(&tm[3], o, B, H, W_, 256, &terr, 32)) return fail(nullptr, DLV3P_ERR_CUDA, terr);
    CUtensorMap* dtm = tmp.put(tm);
    DwPwParams P{};
    P.tmap_x = &dtm[0]; P.tmap_w = &dtm[1]; P.tmap_out = &dtm[2]; P.dw_w = dw; P.dw_shift = dsh; P.scale = s; P.shift = t;
    P.out = reinterpret_cast<__nv_bfloat16*>(o); P.B = B; P.H = H; P.W = W_;
    P.tiles_x = ceil_div(W_, kDwTW); P.tiles_y = ceil_div(H, kDwTH); P.num_tiles = B * P.tiles_x * P.tiles_y; P.debug = flags;
    rc = run([&] { launch_dwpw(KB, P, &dtm[3], hs.data(), ht.data(), sms, st); });
  } else if (op == 2 && ndims >= 6) {   // resize {B,hi,wi,C,ho,wo}
    ResizeParams P{};
    P.B = static_cast<int>(d[0]); P.hi = static_cast<int>(d[1]); P.wi = static_cast<int>(d[2]); P.C = static_cast<int>(d[3]);
    P.ho = static_cast<int>(d[4]); P.wo = static_cast<int>(d[5]); P.ldo = P.C; P.col_off = 0;
    P.sy = static_cast<float>(P.hi) / P.ho; P.sx = static_cast<float>(P.wi) / P.wo;
    P.x = reinterpret_cast<__nv_bfloat16*>(fill(static_cast<size_t>(P.B) * P.hi * P.wi * P.C));
    P.out = reinterpret_cast<__nv_bfloat16*>(tmp.alloc<uint16_t>(static_cast<size_t>(P.B) * P.ho * P.wo * P.C));
    const bool x4 = P.ho == 4 * P.hi && P.wo == 4 * P.wi && !(flags & 1);
    rc = run([&] {
      if (x4) resize_bilinear_x4_kernel<<<grid_for(static_cast<size_t>(P.B) * (P.hi + 1) * (P.wi + 1) * 32, sms), 256, 0, st>>>(P);
      else resize_bilinear_kernel<<<dim3(P.ho, P.B), 256, 0, st>>>(P);
    });
  } else if (op == 3 && ndims >= 6) {   // resize_argmax {B,NC,hi,wi,ho,wo}
    ArgmaxParams P{};
    P.B = static_cast<int>(d[0]); P.NC = static_cast<int>(d[1]); P.hi = static_cast<int>(d[2]); P.wi = static_cast<int>(d[3]);
    P.ho = static_cast<int>(d[4]); P.wo = static_cast<int>(d[5]);
    P.sy = static_cast<float>(P.hi) / P.ho; P.sx = static_cast<float>(P.wi) / P.wo;
    const size_t n = static_cast<size_t>(P.B) * P.NC * P.hi * P.wi;
    std::vector<float> h(n);
    uint32_t sd = 777u;
    for (size_t i = 0; i < n; ++i) { sd = sd * 1664525u + 1013904223u; h[i] = static_cast<float>(sd >> 8) / 8388608.0f - 1.0f; }
    P.logits = tmp.put(h);
    P.labels = tmp.alloc<uint8_t>(static_cast<size_t>(P.B) * P.ho * P.wo);
    rc = run([&] { launch_resize_argmax(P, (flags & 1) != 0, sms, st); });
  } else if (op == 4 && ndims >= 4) {   // ASPP depthwise slab kernel {B,h,w,C} at OS16 rates
    AsppDwParams P{};
    P.B = static_cast<int>(d[0]); P.h = static_cast<int>(d[1]); P.w_ = static_cast<int>(d[2]); P.C = static_cast<int>(d[3]);
    P.nrates = 3; P.nchunks = ceil_div(P.C, 64); P.pool_items = 1; P.debug = flags;
    static const int kTs[5] = {2, 3, 4, 6, 8};
    const int rr[3] = {6, 12, 18};
    for (int i = 0; i < 3; ++i) {
      P.rates[i] = rr[i];
      const int nt = ceil_div(P.w_, rr[i]);
      int sel = 4;
      for (int k = 4; k >= 0; --k)
        if (kTs[k] >= (nt < 8 ? nt : 8)) sel = k;
      const int na_max = ceil_div(P.h, rr[i]);
      if (na_max <= 2 && nt <= 2) sel = 5;
      else if (na_max <= 3 && nt <= 3) sel = 6;
      P.ts_sel[i] = sel; P.nseg[i] = sel >= 5 ? 1 : ceil_div(nt, kTs[sel]); P.item_off[i + 1] = P.item_off[i] + rr[i] * rr[i] * P.nseg[i];
    }
    const size_t n = static_cast<size_t>(P.B) * P.h * P.w_ * P.C;
    P.x = reinterpret_cast<__nv_bfloat16*>(fill(n));
    P.out = reinterpret_cast<__nv_bfloat16*>(tmp.alloc<uint16_t>(3 * static_cast<size_t>(P.nchunks) * P.B * P.h * P.w_ * 64));
    P.w = tmp.put(std::vector<float>(static_cast<size_t>(27) * P.C, 0.1f));
    P.shift = tmp.put(std::vector<float>(static_cast<size_t>(3) * P.C, 0.0f));
    P.pool_partial = tmp.alloc<float>(static_cast<size_t>(P.B) * P.C);
    const bool fast = aspp_fast_supported(P) && !(flags & 8);
    std::vector<CUtensorMap> tm(1);
    if (!encode_2d_slab(&tm[0], P.x, static_cast<uint64_t>(P.B) * P.h * P.w_, P.C, P.C, &terr, fast ? 32 : 64)) return fail(nullptr, DLV3P_ERR_CUDA, terr);
    P.tmap_slab = tmp.put(tm);
    P.item_table = tmp.put(build_aspp_items(P));
    const size_t smem = static_cast<size_t>(ceil_div(P.h * P.w_, 256)) * 32768 + (27 * 64 + 3 * 64 + 16 * 64) * sizeof(float) + 16;
    if (smem > 220 * 1024) return fail(nullptr, DLV3P_ERR_UNSUPPORTED, "op_time aspp_dw: map too large for the slab kernel");
    cudaFuncSetAttribute(aspp_dw_slab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    const int slabs = P.B * P.nchunks;
    if (fast) rc = run([&] { launch_aspp_fast(P, st); });
    else rc = run([&] { aspp_dw_slab_kernel<<<slabs < sms ? slabs : sms, kSlabThreads, smem, st>>>(P); });
  } else {
    rc = fail(nullptr, DLV3P_ERR_INVALID, "op_time: unknown op / too few dims");
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return rc;
}

}  // extern "C"

// training-step operators (include/dlv3p_train.h)
#include "train_api.cuh"
#include "trainer_api.cuh"
