// xception_api.cu — C ABI of the whole DeepLabV3+ Xception model (include/dlv3p_model.h): the modified aligned Xception feature
// extractor as sm_100a kernels (bb_kernels.cuh, bb_conv3x3.cuh, bb_gemm.cuh) in front of the head context of dlv3p_api.cu.
//
// Reference sites (paths relative to the reference repo):
//   Xception_body / _xception_block / _conv2d_same   deeplabv3p/models/deeplabv3p_xception.py:25-163
//   SepConv_BN (backbone use)                        deeplabv3p/models/layers.py:74-111
//   Deeplabv3pXception / get_deeplabv3p_model        deeplabv3p_xception.py:167-239, deeplabv3p/model.py:51-117
//   normalize_image                                  common/data_utils.py:403-416
#include <cuda.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <utility>
#include <vector>

#include "../../include/dlv3p_model.h"
#include "bb_conv3x3.cuh"
#include "bb_gemm.cuh"
#include "bb_kernels.cuh"
#include "bb_sepconv.cuh"
#include "bb_sepwide.cuh"
#include "bb_stem_tc.cuh"
#include "f32_kernels.cuh"

using namespace dlv3p;

namespace {

thread_local std::string g_model_tls_error;

std::string mfmt(const char* f, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, f);
  vsnprintf(buf, sizeof(buf), f, ap);
  va_end(ap);
  return std::string(buf);
}
inline int cdiv(int a, int b) { return (a + b - 1) / b; }

uint16_t bf16_rne(float f) {
  uint32_t u;
  std::memcpy(&u, &f, 4);
  if ((u & 0x7F800000u) == 0x7F800000u && (u & 0x007FFFFFu)) return static_cast<uint16_t>((u >> 16) | 0x40);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return static_cast<uint16_t>(u >> 16);
}
float bf16_f32(uint16_t b) {
  uint32_t u = static_cast<uint32_t>(b) << 16;
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn(std::string* err) {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
    if (err) *err = mfmt("cudaGetDriverEntryPoint(cuTensorMapEncodeTiled) failed: %s", cudaGetErrorString(e));
    return nullptr;
  }
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}
// bf16 tensor map of rank 2..4: dims innermost first, strides in BYTES for dims 1.., zero fill out of bounds
bool tm_encode(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides, const uint32_t* box, CUtensorMapSwizzle sw,
               std::string* err) {
  EncodeTiledFn fn = encode_fn(err);
  if (!fn) return false;
  cuuint64_t gd[4], gs[3];
  cuuint32_t bx[4], es[4] = {1, 1, 1, 1};
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides[i];
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    if (err) *err = mfmt("cuTensorMapEncodeTiled(rank %d, dims %llu %llu, box %u %u) -> %d", rank, (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1], (int)r);
    return false;
  }
  return true;
}
bool tm_2d(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_cols, uint32_t box_rows, CUtensorMapSwizzle sw, std::string* err) {
  const uint64_t d[2] = {cols, rows}, s[1] = {ld * 2};
  const uint32_t b[2] = {box_cols, box_rows};
  return tm_encode(tm, base, 2, d, s, b, sw, err);
}
bool tm_nhwc(CUtensorMap* tm, const void* base, uint64_t B, uint64_t H, uint64_t W, uint64_t C, uint32_t bc, uint32_t bw, uint32_t bh, CUtensorMapSwizzle sw,
             std::string* err) {
  const uint64_t d[4] = {C, W, H, B}, s[3] = {C * 2, W * C * 2, H * W * C * 2};
  const uint32_t b[4] = {bc, bw, bh, 1};
  return tm_encode(tm, base, 4, d, s, b, sw, err);
}

// Launch with programmatic stream serialization: the kernel's prologue (barrier init, TMEM allocation, tap / weight preloads)
// overlaps the tail of the previous kernel of the stream; every kernel of the backbone orders itself behind its producer with
// griddepcontrol.wait before it touches an activation (sm100_prims.cuh).
template <class... KArgs, class... Args>
cudaError_t launch_pdl(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
bool g_pdl = true;      // per launch sequence (dlv3p_model_config.flags bit 1 turns it off for A/B measurements)

// ---- depthwise kernel variants: (stride, rate) -> tile shape
struct DwVariant { int S, R, TH, TW; };
bool dw_variant(int stride, int rate, DwVariant* v) {
  if (stride == 1 && rate == 1) { *v = {1, 1, 8, 32}; return true; }
  if (stride == 2 && rate == 1) { *v = {2, 1, 8, 16}; return true; }
  if (stride == 1 && rate == 2) { *v = {1, 2, 8, 32}; return true; }
  if (stride == 1 && rate == 4) { *v = {1, 4, 4, 32}; return true; }
  return false;
}
template <int S, int R, int TH, int TW>
cudaError_t launch_dw_t(const BbDwParams& P, cudaStream_t st) {
  using Cfg = BbDwCfg<S, R, TH, TW>;
  static bool attr_done[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_done[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(bb_depthwise_kernel<S, R, TH, TW>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) return e;
    attr_done[dev & 63] = true;
  }
  static int sms[64] = {};
  if (!sms[dev & 63]) cudaDeviceGetAttribute(&sms[dev & 63], cudaDevAttrMultiProcessorCount, dev);
  const long long items = static_cast<long long>(P.B) * P.tiles_y * P.tiles_x * P.cgroups;
  const int per_sm = (227 * 1024) / (Cfg::kSmemBytes + 1024) > 0 ? (227 * 1024) / (Cfg::kSmemBytes + 1024) : 1;   // persistent: every CTA resident
  const long long cap = static_cast<long long>(sms[dev & 63]) * (per_sm < 4 ? per_sm : 4);
  // a multiple of the channel groups: every CTA keeps one group (its taps stay in registers) and strides over the spatial tiles
  long long grid = items < cap ? items : cap / P.cgroups * P.cgroups;
  if (grid < P.cgroups) grid = P.cgroups;
  return launch_pdl(g_pdl, bb_depthwise_kernel<S, R, TH, TW>, dim3(static_cast<unsigned>(grid)), dim3(Cfg::kThreads), Cfg::kSmemBytes, st, P);
}
cudaError_t launch_dw(const DwVariant& v, const BbDwParams& P, cudaStream_t st) {
  if (v.S == 1 && v.R == 1 && v.TH == 4) return launch_dw_t<1, 1, 4, 32>(P, st);
  if (v.S == 1 && v.R == 1) return launch_dw_t<1, 1, 8, 32>(P, st);
  if (v.S == 2 && v.R == 1) return launch_dw_t<2, 1, 8, 16>(P, st);
  if (v.S == 1 && v.R == 2) return launch_dw_t<1, 2, 8, 32>(P, st);
  if (v.S == 1 && v.R == 4) return launch_dw_t<1, 4, 4, 32>(P, st);
  return cudaErrorInvalidValue;
}
void dw_box(const DwVariant& v, uint32_t* bw, uint32_t* bh) {
  *bh = static_cast<uint32_t>((v.TH - 1) * v.S + 2 * v.R + 1);
  *bw = static_cast<uint32_t>((v.TW - 1) * v.S + 2 * v.R + 1);
}

template <int BN, bool kRes>
cudaError_t launch_bb_gemm_t(const BbGemmParams& P, int num_sms, cudaStream_t st) {
  static bool attr_done[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_done[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(bb_gemm_kernel<BN, kRes>, cudaFuncAttributeMaxDynamicSharedMemorySize, BbCfg<BN, kRes>::kSmemBytes);
    if (e != cudaSuccess) return e;
    attr_done[dev & 63] = true;
  }
  const int items = P.m_pairs * P.n_tiles;
  const int grid = 2 * items < num_sms ? 2 * items : (num_sms & ~1);
  return launch_pdl(g_pdl, bb_gemm_kernel<BN, kRes>, dim3(grid), dim3(kBbThreads), BbCfg<BN, kRes>::kSmemBytes, st, P);
}
cudaError_t launch_bb_gemm(int BN, const BbGemmParams& P, int num_sms, cudaStream_t st) {
  const bool res = P.tmap_res != nullptr;     // the residual variant trades pipeline stages for the residual staging buffers
  switch (BN) {
    case 256: return res ? launch_bb_gemm_t<256, true>(P, num_sms, st) : launch_bb_gemm_t<256, false>(P, num_sms, st);
    case 192: return res ? launch_bb_gemm_t<192, true>(P, num_sms, st) : launch_bb_gemm_t<192, false>(P, num_sms, st);
    case 128: return res ? launch_bb_gemm_t<128, true>(P, num_sms, st) : launch_bb_gemm_t<128, false>(P, num_sms, st);
    default: return cudaErrorInvalidValue;
  }
}
// N tile of a pointwise GEMM: the candidate that needs the fewest tile-columns over all rounds of the CTA pairs (work items are
// (M pair, N tile); every round costs one tile of BN columns), larger tiles on ties (fewer passes over A)
int pick_bb_bn(int M, int N, int num_sms) {
  const int m_pairs = cdiv(cdiv(M, kBbBM), 2), clusters = num_sms / 2 > 0 ? num_sms / 2 : 1;
  int best = 256;
  long long best_cost = -1;
  for (int bn : {256, 192, 128}) {
    const long long items = static_cast<long long>(m_pairs) * cdiv(N, bn);
    const long long cost = (items + clusters - 1) / clusters * bn;
    if (best_cost < 0 || cost < best_cost) { best = bn; best_cost = cost; }
  }
  return best;
}
template <int KB, int NB>
cudaError_t launch_sepconv_t(const BbSepParams& Q, int num_sms, cudaStream_t st) {
  static bool attr_done[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_done[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(bb_sepconv_kernel<KB, NB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BbSepCfg<KB, NB>::kSmemBytes);
    if (e != cudaSuccess) return e;
    attr_done[dev & 63] = true;
  }
  const int items = (Q.base.num_tiles + 1) / 2;
  const int grid = 2 * items < num_sms ? 2 * items : (num_sms & ~1);
  return launch_pdl(g_pdl, bb_sepconv_kernel<KB, NB, false>, dim3(grid), dim3(kSepThreads), BbSepCfg<KB, NB>::kSmemBytes, st, Q);
}
// fused SepConv_BN of the entry flow: C in {64, 128, 192, 256}, N in {128, 256}
bool sepconv_supported(int C, int N, int stride, int rate, bool act) { return !act && stride == 1 && rate == 1 && C % 64 == 0 && C >= 64 && C <= 256 && (N == 128 || N == 256); }
cudaError_t launch_sepconv(int KB, int NB, const BbSepParams& Q, int num_sms, cudaStream_t st) {
  if (NB == 128) {
    switch (KB) {
      case 1: return launch_sepconv_t<1, 128>(Q, num_sms, st);
      case 2: return launch_sepconv_t<2, 128>(Q, num_sms, st);
      case 3: return launch_sepconv_t<3, 128>(Q, num_sms, st);
      case 4: return launch_sepconv_t<4, 128>(Q, num_sms, st);
    }
  } else if (NB == 256) {
    switch (KB) {
      case 1: return launch_sepconv_t<1, 256>(Q, num_sms, st);
      case 2: return launch_sepconv_t<2, 256>(Q, num_sms, st);
      case 3: return launch_sepconv_t<3, 256>(Q, num_sms, st);
      case 4: return launch_sepconv_t<4, 256>(Q, num_sms, st);
    }
  }
  return cudaErrorInvalidValue;
}
// fused SepConv_BN of the middle flow (bb_sepwide.cuh): one cluster of two CTAs per 8 x 16-pixel tile
bool sepwide_supported(int C, int N, int stride, int rate, bool act) {
  return !act && stride == 1 && rate == 1 && C % 8 == 0 && C >= 64 && C <= 768 && N % 8 == 0 && N > 656 && N <= 768;
}
template <bool kReluIn>
cudaError_t launch_sepwide_t(const BbWideParams& P, int num_sms, cudaStream_t st) {
  static bool attr_done[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_done[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(bb_sepwide_kernel<kReluIn>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWideSmemBytes);
    if (e != cudaSuccess) return e;
    attr_done[dev & 63] = true;
  }
  const int grid = 2 * P.num_tiles < num_sms ? 2 * P.num_tiles : (num_sms & ~1);
  return launch_pdl(g_pdl, bb_sepwide_kernel<kReluIn>, dim3(grid), dim3(kWideThreads), kWideSmemBytes, st, P);
}
cudaError_t launch_sepwide(const BbWideParams& P, bool relu_in, int num_sms, cudaStream_t st) {
  return relu_in ? launch_sepwide_t<true>(P, num_sms, st) : launch_sepwide_t<false>(P, num_sms, st);
}
cudaError_t launch_conv3x3(const Conv3x3Params& P, int num_sms, cudaStream_t st) {
  static bool attr_done[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_done[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(conv3x3_c32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kC3SmemBytes);
    if (e != cudaSuccess) return e;
    attr_done[dev & 63] = true;
  }
  const int grid = P.num_tiles < num_sms ? P.num_tiles : num_sms;
  return launch_pdl(g_pdl, conv3x3_c32_kernel, dim3(grid), dim3(kC3Threads), kC3SmemBytes, st, P);
}

// ---- host-side packing
struct Fold { std::vector<float> scale, shift; };
Fold fold_bn(const float* gamma, const float* beta, const float* mean, const float* var, int n, float eps) {
  Fold f;
  f.scale.resize(n);
  f.shift.resize(n);
  for (int i = 0; i < n; ++i) {
    const float inv = gamma[i] / sqrtf(var[i] + eps);
    f.scale[i] = inv;
    f.shift[i] = beta[i] - mean[i] * inv;
  }
  return f;
}
// Keras 1x1 kernel [K][N] fp32 -> bf16 [Npad][Kpad] K-major, zero padded
std::vector<uint16_t> pack_kn(const float* w_kn, int K, int N, int Npad, int Kpad) {
  std::vector<uint16_t> out(static_cast<size_t>(Npad) * Kpad, 0);
  for (int k = 0; k < K; ++k)
    for (int n = 0; n < N; ++n) out[static_cast<size_t>(n) * Kpad + k] = bf16_rne(w_kn[static_cast<size_t>(k) * N + n]);
  return out;
}
// Keras depthwise kernel [3][3][C] x BN scale -> [9][Cpad]
std::vector<float> pack_taps(const float* w_hwc, const float* scale, int C, int Cpad) {
  std::vector<float> out(static_cast<size_t>(9) * Cpad, 0.0f);
  for (int t = 0; t < 9; ++t)
    for (int c = 0; c < C; ++c) out[static_cast<size_t>(t) * Cpad + c] = w_hwc[static_cast<size_t>(t) * C + c] * (scale ? scale[c] : 1.0f);
  return out;
}
// entry_flow_conv1_1 for the tensor-core stem (bb_stem_tc.cuh): Keras kernel [3][3][3][32] fp32 -> three bf16 pieces of every weight (w0 + w1 + w2 == w
// exactly), laid out [piece][s2d tap (ty, tx)][32 cout][16 k] (32-byte swizzle) with k = (dy, dx, cin), conv tap (2 ty + dy, 2 tx + dx); and the per-tap channel
// sums [9][32] + their total [32] the epilogue subtracts (conv(x / 127.5 - 1) = conv(x) / 127.5 - sum of the weights of the taps inside the image)
float bf16_to_f32_host(uint16_t b) {
  const uint32_t u = static_cast<uint32_t>(b) << 16;
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}
std::vector<uint16_t> pack_stem_tc(const float* w) {
  std::vector<uint16_t> out(static_cast<size_t>(3) * 4 * 2 * 32 * 8, 0);
  for (int t = 0; t < 4; ++t)
    for (int dy = 0; dy < 2; ++dy)
      for (int dx = 0; dx < 2; ++dx) {
        const int ky = 2 * (t >> 1) + dy, kx = 2 * (t & 1) + dx;
        if (ky > 2 || kx > 2) continue;
        for (int c = 0; c < 3; ++c)
          for (int n = 0; n < 32; ++n) {
            const int k = (dy * 2 + dx) * 3 + c;
            float rem = w[((ky * 3 + kx) * 3 + c) * 32 + n];
            for (int piece = 0; piece < 3; ++piece) {
              const uint16_t b = bf16_rne(rem);
              // [piece][tap][cout n][16 k], 32-byte rows with the two 16-byte halves swapped where bit 7 of the row's offset is set (n & 4)
              out[((static_cast<size_t>(piece) * 4 + t) * 32 + n) * 16 + ((((k >> 3) ^ ((n >> 2) & 1)) << 3) | (k & 7))] = b;
              rem -= bf16_to_f32_host(b);      // exact: the difference of a float and its bf16 rounding is a float
            }
          }
      }
  return out;
}
std::vector<float> stem_tap_sums(const float* w) {
  std::vector<float> out(10 * 32, 0.0f);
  for (int n = 0; n < 32; ++n) {
    double tot = 0.0;
    for (int t = 0; t < 9; ++t) {
      double s = 0.0;
      for (int c = 0; c < 3; ++c) s += w[(t * 3 + c) * 32 + n];
      out[t * 32 + n] = static_cast<float>(s);
      tot += s;
    }
    out[9 * 32 + n] = static_cast<float>(tot);
  }
  return out;
}
cudaError_t launch_stem_tc(StemTcParams P, int num_sms, cudaStream_t st, bool pdl) {
  static bool attr_done[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_done[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(stem_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kStcSmemBytes);
    if (e != cudaSuccess) return e;
    attr_done[dev & 63] = true;
  }
  P.tiles_x = cdiv(P.Wo, kStcTW); P.tiles_y = cdiv(P.Ho, kStcTH); P.num_tiles = P.B * P.tiles_x * P.tiles_y;
  const int grid = P.num_tiles < num_sms ? P.num_tiles : num_sms;
  return launch_pdl(pdl, stem_tc_kernel, dim3(grid), dim3(kStcThreads), kStcSmemBytes, st, P);
}
// Keras 3x3 kernel [3][3][32][64] -> bf16 [9 taps][64 cout][32 cin]
std::vector<uint16_t> pack_c3(const float* w) {
  std::vector<uint16_t> out(static_cast<size_t>(9) * 64 * 32);
  for (int t = 0; t < 9; ++t)
    for (int ci = 0; ci < 32; ++ci)
      for (int co = 0; co < 64; ++co) out[(static_cast<size_t>(t) * 64 + co) * 32 + ci] = bf16_rne(w[(static_cast<size_t>(t) * 32 + ci) * 64 + co]);
  return out;
}

}  // namespace

// =====================================================================================================
// model
// =====================================================================================================
struct MWeight {
  std::string layer, var;
  std::vector<int64_t> shape;
  std::vector<float> data;
  bool set = false;
};
struct Tensor {
  __nv_bfloat16* p = nullptr;
  int B = 0, H = 0, W = 0, C = 0;
  size_t elems() const { return static_cast<size_t>(B) * H * W * C; }
  int M() const { return B * H * W; }
};
enum OpKind { OP_STEM, OP_CONV3, OP_DW, OP_PW, OP_SUB, OP_SEP, OP_WIDE };
struct Op {
  OpKind kind;
  std::string name;      // Keras layer name of the convolution the kernel computes
  std::string bn;        // its BatchNormalization layer
  std::string prefix;    // OP_SEP: the SepConv_BN prefix (<prefix>_depthwise / _depthwise_BN / _pointwise / _pointwise_BN)
  int in = -1, out = -1, res = -1;   // tensor indices
  int stride = 1, rate = 1, relu_in = 0, relu_out = 0;
  int a_stride = 1;      // OP_PW: 2 = the A operand is every second pixel of every second row of `in` (strided tensor map, no sampling kernel)
  int K = 0, N = 0, Kpad = 0, Npad = 0, Cpad = 0, BN = 256;
  DwVariant dv{1, 1, 8, 32};
  // device weights
  uint16_t* w16 = nullptr;
  float *wf = nullptr, *scale = nullptr, *shift = nullptr, *dshift = nullptr;
  std::vector<float> h_scale, h_shift;
  int tm0 = -1;          // first tensor-map slot of the op
  double flops = 0, bytes = 0;
};

struct BlockSpec { std::string prefix; int cin; int depth[3]; int shortcut /*0 conv, 1 sum, 2 none*/; int stride, rate; bool act; bool ret_skip; };

struct dlv3p_model {
  dlv3p_model_config cfg{};
  int device = 0;
  bool plan_only = false;
  int num_sms = 148;
  std::string err;
  dlv3p_ctx* head = nullptr;
  std::vector<MWeight> weights;          // backbone only
  std::map<std::string, int> windex;
  std::vector<Tensor> tensors;
  std::map<std::string, int> taps;       // tap name -> tensor index
  std::vector<Op> ops;
  int t_feat = -1, t_skip = -1;
  bool finalized = false;
  std::vector<void*> allocs, weight_allocs;
  size_t ws_bytes = 0;
  std::vector<CUtensorMap> h_tm;
  CUtensorMap* d_tm = nullptr;
  void *in_stage = nullptr, *out_stage = nullptr;
  cudaStream_t own_stream = nullptr;
  int64_t launches_last = 0;
  int pad_t = 0, pad_l = 0;
  // profiling
  bool profiling = false;
  std::vector<cudaEvent_t> prof_events;
  // fp32 precision mode (DLV3P_MODEL_FLAG_FP32): host copies of the head's weights, device fp32 weights, activation buffers, taps
  bool fp32 = false;
  std::map<std::string, std::vector<float>> head_w;
  std::map<std::string, float*> f32_w;
  std::vector<std::pair<float*, size_t>> f32_bufs;
  size_t f32_next = 0;
  struct F32Tap { float* p; int B, H, W, C; };
  std::map<std::string, F32Tap> f32_taps;
  std::vector<BlockSpec> f32_blocks;
};

namespace {

int mfail(dlv3p_model* m, int code, const std::string& msg) {
  if (m) m->err = msg;
  g_model_tls_error = msg;
  return code;
}
#define MCU(m, expr)                                                                                        \
  do {                                                                                                      \
    cudaError_t _e = (expr);                                                                                \
    if (_e != cudaSuccess) return mfail((m), DLV3P_ERR_CUDA, mfmt("%s: %s", #expr, cudaGetErrorString(_e))); \
  } while (0)

int m_alloc(dlv3p_model* m, void** p, size_t bytes, bool weight) {
  bytes = (bytes + 511) / 256 * 256;
  m->ws_bytes += bytes;
  *p = nullptr;
  if (m->plan_only) return 0;
  cudaError_t e = cudaMalloc(p, bytes);
  if (e != cudaSuccess) return mfail(m, DLV3P_ERR_NOMEM, mfmt("cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e)));
  (weight ? m->weight_allocs : m->allocs).push_back(*p);
  return 0;
}
template <class T>
int m_upload(dlv3p_model* m, T** p, const std::vector<T>& h) {
  void* q = nullptr;
  int r = m_alloc(m, &q, h.size() * sizeof(T), true);
  if (r) return r;
  MCU(m, cudaMemcpy(q, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  *p = static_cast<T*>(q);
  return 0;
}

void add_w(dlv3p_model* m, const std::string& layer, const std::string& var, std::vector<int64_t> shape) {
  MWeight s;
  s.layer = layer;
  s.var = var;
  s.shape = std::move(shape);
  m->windex[layer + "/" + var] = static_cast<int>(m->weights.size());
  m->weights.push_back(std::move(s));
}
void add_bn(dlv3p_model* m, const std::string& n, int ch) {
  for (const char* v : {"gamma", "beta", "moving_mean", "moving_variance"}) add_w(m, n, v, {ch});
}
const float* MW(const dlv3p_model* m, const std::string& layer, const std::string& var) {
  auto it = m->windex.find(layer + "/" + var);
  return it == m->windex.end() ? nullptr : m->weights[it->second].data.data();
}
Fold mfold(const dlv3p_model* m, const std::string& bn, int n) {
  return fold_bn(MW(m, bn, "gamma"), MW(m, bn, "beta"), MW(m, bn, "moving_mean"), MW(m, bn, "moving_variance"), n, 1e-3f);   // Keras default epsilon
}

int new_tensor(dlv3p_model* m, int B, int H, int W, int C, const char* tap = nullptr) {
  Tensor t;
  t.B = B; t.H = H; t.W = W; t.C = C;
  m->tensors.push_back(t);
  const int idx = static_cast<int>(m->tensors.size()) - 1;
  if (tap) m->taps[tap] = idx;
  return idx;
}

// the architecture walk of Xception_body (deeplabv3p_xception.py:96-163): registers weights, tensors and ops
int build_plan(dlv3p_model* m) {
  const dlv3p_model_config& g = m->cfg;
  int s16, r16, s32, r32;
  if (g.OS == 8) { s16 = 1; r16 = 2; s32 = 1; r32 = 4; }
  else if (g.OS == 16) { s16 = 2; r16 = 1; s32 = 1; r32 = 2; }
  else if (g.OS == 32) { s16 = 2; r16 = 1; s32 = 2; r32 = 1; }
  else return mfail(m, DLV3P_ERR_INVALID, mfmt("invalid output stride %d", g.OS));
  std::vector<BlockSpec> blocks;
  blocks.push_back({"entry_flow_block1", 64, {128, 128, 128}, 0, 2, 1, false, false});
  blocks.push_back({"entry_flow_block2", 128, {256, 256, 256}, 0, 2, 1, false, true});
  blocks.push_back({"entry_flow_block3", 256, {728, 728, 728}, 0, s16, 1, false, false});
  for (int i = 0; i < 16; ++i) blocks.push_back({mfmt("middle_flow_unit_%d", i + 1), 728, {728, 728, 728}, 1, 1, r16, false, false});
  blocks.push_back({"exit_flow_block1", 728, {728, 1024, 1024}, 0, s32, r16, false, false});
  blocks.push_back({"exit_flow_block2", 1024, {1536, 1536, 2048}, 2, 1, r32, true, false});

  m->f32_blocks = blocks;
  const int B = g.B;
  // entry_flow_conv1_1: Conv2D(32, 3, strides 2, 'same'): TensorFlow pads (total/2, total - total/2) per axis
  int h = cdiv(g.H, 2), w = cdiv(g.W, 2);
  {
    const int ph = std::max((h - 1) * 2 + 3 - g.H, 0), pw = std::max((w - 1) * 2 + 3 - g.W, 0);
    m->pad_t = ph / 2;
    m->pad_l = pw / 2;
  }
  add_w(m, "entry_flow_conv1_1", "kernel", {3, 3, 3, 32});
  add_bn(m, "entry_flow_conv1_1_BN", 32);
  add_w(m, "entry_flow_conv1_2", "kernel", {3, 3, 32, 64});
  add_bn(m, "entry_flow_conv1_2_BN", 64);
  int t1 = new_tensor(m, B, h, w, 32, "entry_flow_conv1_1");
  {
    Op o; o.kind = OP_STEM; o.name = "entry_flow_conv1_1"; o.bn = "entry_flow_conv1_1_BN"; o.out = t1;
    o.flops = 2.0 * B * h * w * 27 * 32; o.bytes = static_cast<double>(B) * g.H * g.W * 3 * (g.img_dtype == DLV3P_IMG_F32 ? 4 : 1) + 2.0 * B * h * w * 32;
    m->ops.push_back(o);
  }
  int x = new_tensor(m, B, h, w, 64, "entry_flow_conv1_2");
  {
    Op o; o.kind = OP_CONV3; o.name = "entry_flow_conv1_2"; o.bn = "entry_flow_conv1_2_BN"; o.in = t1; o.out = x;
    o.flops = 2.0 * B * h * w * 288 * 64; o.bytes = 2.0 * B * h * w * (32 + 64);
    m->ops.push_back(o);
  }
  // 'sum' blocks (the 16 middle-flow units) ping-pong over a small pool of equally sized tensors: the whole flow lives on four
  // buffers that stay L2-resident.  flags bit 0 keeps every intermediate in its own tensor (block-level parity taps).
  int pool_a = -1, pool_b = -1, pool_d = -1, pool_p = -1;
  auto pooled_tensor = [&](int* slot, int hh, int ww, int cc) {
    if (*slot < 0) *slot = new_tensor(m, B, hh, ww, cc);
    return *slot;
  };
  for (const BlockSpec& b : blocks) {
    const int cin = b.cin;
    const Tensor tin = m->tensors[x];
    const int ho = cdiv(tin.H, b.stride), wo = cdiv(tin.W, b.stride);
    const bool pooled = b.shortcut == 1 && !(g.flags & 1);
    std::vector<Op> seq;
    // --- shortcut branch first: [stride-2 sampling ->] 1x1 conv -> BN; the sum happens in the last pointwise GEMM's epilogue
    int res = b.shortcut == 1 ? x : -1;
    if (b.shortcut == 0) {
      int src = x;
      // stride 2: the GEMM reads the sampled pixels itself through a strided tensor map when its 128-pixel M tiles are whole rows (or a
      // piece of one row) of one image; otherwise a sampling kernel makes the dense operand first
      const bool strided_a = b.stride == 2 && (ho * wo) % 128 == 0 && (wo % 128 == 0 || 128 % wo == 0) && !(g.flags & 1);
      if (b.stride == 2 && !strided_a) {
        src = new_tensor(m, B, ho, wo, cin);
        Op s; s.kind = OP_SUB; s.name = b.prefix + "_shortcut_sample"; s.in = x; s.out = src;
        s.bytes = 2.0 * B * ho * wo * cin * 2;
        seq.push_back(s);
      }
      res = new_tensor(m, B, ho, wo, b.depth[2]);
      Op q; q.kind = OP_PW; q.name = b.prefix + "_shortcut"; q.bn = b.prefix + "_shortcut_BN"; q.in = src; q.out = res; q.K = cin; q.N = b.depth[2];
      q.a_stride = strided_a ? 2 : 1;
      q.flops = 2.0 * B * ho * wo * static_cast<double>(cin) * b.depth[2]; q.bytes = 2.0 * B * ho * wo * (cin + b.depth[2]);
      seq.push_back(q);
    }
    // --- three SepConv_BN
    int cur = x, c = cin;
    bool relu_prev = false;      // the producer of `cur` already applied this block's pre-depthwise ReLU
    for (int i = 0; i < 3; ++i) {
      const std::string p = mfmt("%s_separable_conv%d", b.prefix.c_str(), i + 1);
      const int st = i == 2 ? b.stride : 1;
      const Tensor ti = m->tensors[cur];
      const int oh = cdiv(ti.H, st), ow = cdiv(ti.W, st);
      if (!pooled && sepconv_supported(c, b.depth[i], st, b.rate, b.act) && !(i == 2 && res >= 0) && !(g.flags & DLV3P_MODEL_FLAG_UNFUSED_ENTRY)) {
        // entry flow: depthwise -> BN -> pointwise -> BN in ONE kernel, the depthwise output never leaves the SM (bb_sepconv.cuh)
        const int n = b.depth[i];
        const int tp = new_tensor(m, B, oh, ow, n);
        Op f; f.kind = OP_SEP; f.prefix = p; f.name = p + "_sepconv"; f.in = cur; f.out = tp; f.K = c; f.N = n; f.Cpad = c;
        f.relu_in = b.act ? 0 : 1; f.relu_out = b.act ? 1 : 0;
        f.flops = 2.0 * B * oh * ow * (9.0 * c + static_cast<double>(c) * n); f.bytes = 2.0 * B * oh * ow * (c + n);
        seq.push_back(f);
        if (i == 1 && b.ret_skip) { m->t_skip = tp; m->taps["skip"] = tp; }
        cur = tp;
        c = n;
        relu_prev = false;
        continue;
      }
      if (sepwide_supported(c, b.depth[i], st, b.rate, b.act) && (g.flags & DLV3P_MODEL_FLAG_FUSED_MIDDLE)) {
        // middle flow: the same fusion with N split over a cluster of two SMs that share the depthwise result (bb_sepwide.cuh)
        const int n = b.depth[i];
        int tp;
        if (!pooled) tp = new_tensor(m, B, oh, ow, n);
        else if (i == 0) tp = pooled_tensor(&pool_p, oh, ow, n);
        else if (i == 1) tp = pooled_tensor(&pool_d, oh, ow, n);
        else tp = (x == pool_a) ? pooled_tensor(&pool_b, oh, ow, n) : pooled_tensor(&pool_a, oh, ow, n);
        Op f; f.kind = OP_WIDE; f.prefix = p; f.name = p + "_sepconv"; f.in = cur; f.out = tp; f.K = c; f.N = n; f.Cpad = cdiv(c, 64) * 64;
        f.relu_in = relu_prev ? 0 : 1;
        const bool relu_moved = i < 2 && !(i == 1 && b.ret_skip);     // see the pointwise GEMM below
        f.relu_out = relu_moved ? 1 : 0;
        relu_prev = relu_moved;
        f.flops = 2.0 * B * oh * ow * (9.0 * c + static_cast<double>(c) * n); f.bytes = 2.0 * B * oh * ow * (c + n);
        if (i == 2 && res >= 0) { f.res = res; f.bytes += 2.0 * B * oh * ow * n; }
        seq.push_back(f);
        if (i == 1 && b.ret_skip) { m->t_skip = tp; m->taps["skip"] = tp; }
        cur = tp;
        c = n;
        continue;
      }
      const int td = pooled ? pooled_tensor(&pool_d, oh, ow, c) : new_tensor(m, B, oh, ow, c);
      Op d; d.kind = OP_DW; d.name = p + "_depthwise"; d.bn = p + "_depthwise_BN"; d.in = cur; d.out = td; d.stride = st; d.rate = b.rate;
      d.relu_in = (b.act || relu_prev) ? 0 : 1; d.relu_out = b.act ? 1 : 0; d.K = c; d.Cpad = cdiv(c, 64) * 64;
      if (!dw_variant(st, b.rate, &d.dv)) return mfail(m, DLV3P_ERR_UNSUPPORTED, mfmt("depthwise stride %d rate %d is not built", st, b.rate));
      d.flops = 2.0 * B * oh * ow * 9 * c; d.bytes = 2.0 * c * (static_cast<double>(ti.M()) + static_cast<double>(B) * oh * ow);
      seq.push_back(d);
      const int n = b.depth[i];
      int tp;
      if (!pooled) tp = new_tensor(m, B, oh, ow, n);
      else if (i < 2) tp = pooled_tensor(&pool_p, oh, ow, n);
      else tp = (x == pool_a) ? pooled_tensor(&pool_b, oh, ow, n) : pooled_tensor(&pool_a, oh, ow, n);
      Op q; q.kind = OP_PW; q.name = p + "_pointwise"; q.bn = p + "_pointwise_BN"; q.in = td; q.out = tp; q.K = c; q.N = n;
      q.relu_out = b.act ? 1 : 0;
      // depth_activation False: the output of separable_conv1 / 2 is read only by the next depthwise conv, which starts with a ReLU
      // (layers.py:98-99) -> applied here, in the GEMM epilogue (bit identical: rounding is monotonic), unless the tensor is also the
      // decoder's skip feature, which the head takes signed
      const bool relu_moved = !b.act && i < 2 && !(i == 1 && b.ret_skip);
      if (relu_moved) q.relu_out = 1;
      relu_prev = relu_moved;
      q.flops = 2.0 * B * oh * ow * static_cast<double>(c) * n; q.bytes = 2.0 * B * oh * ow * (c + n);
      if (i == 2 && res >= 0) { q.res = res; q.bytes += 2.0 * B * oh * ow * n; }
      seq.push_back(q);
      if (i == 1 && b.ret_skip) { m->t_skip = tp; m->taps["skip"] = tp; }
      cur = tp;
      c = n;
    }
    for (const Op& o : seq) m->ops.push_back(o);
    x = cur;
    m->taps[b.prefix] = x;
  }
  m->t_feat = x;
  m->taps["feature"] = x;
  return DLV3P_OK;
}

// Keras creation order of the backbone's variables (what load_weights(by_name=False) walks): per block the three SepConv_BN,
// then the shortcut conv + BN (deeplabv3p_xception.py:70-85)
void register_block_weights(dlv3p_model* m) {
  // build_plan registered the stem and (out of order) the shortcut kernels; rebuild the list in Keras order
  std::vector<MWeight> stem(m->weights.begin(), m->weights.begin() + 10);
  m->weights = stem;
  m->windex.clear();
  for (size_t i = 0; i < m->weights.size(); ++i) m->windex[m->weights[i].layer + "/" + m->weights[i].var] = static_cast<int>(i);
  std::string cur_block;
  std::vector<const Op*> shortcut;
  auto flush = [&]() {
    for (const Op* o : shortcut) {
      add_w(m, o->name, "kernel", {1, 1, o->K, o->N});
      add_bn(m, o->bn, o->N);
    }
    shortcut.clear();
  };
  for (const Op& o : m->ops) {
    if (o.kind == OP_STEM || o.kind == OP_CONV3 || o.kind == OP_SUB) continue;
    const bool is_shortcut = o.name.size() > 9 && o.name.compare(o.name.size() - 9, 9, "_shortcut") == 0;
    const std::string block = o.name.substr(0, o.name.find(is_shortcut ? "_shortcut" : "_separable_conv"));
    if (block != cur_block) { flush(); cur_block = block; }
    if (is_shortcut) { shortcut.push_back(&o); continue; }
    if (o.kind == OP_SEP || o.kind == OP_WIDE) {
      add_w(m, o.prefix + "_depthwise", "depthwise_kernel", {3, 3, o.K, 1});
      add_bn(m, o.prefix + "_depthwise_BN", o.K);
      add_w(m, o.prefix + "_pointwise", "kernel", {1, 1, o.K, o.N});
      add_bn(m, o.prefix + "_pointwise_BN", o.N);
    } else if (o.kind == OP_DW) {
      add_w(m, o.name, "depthwise_kernel", {3, 3, o.K, 1});
      add_bn(m, o.bn, o.K);
    } else {
      add_w(m, o.name, "kernel", {1, 1, o.K, o.N});
      add_bn(m, o.bn, o.N);
    }
  }
  flush();
}

}  // namespace


// =====================================================================================================
// fp32 precision mode (f32_kernels.cuh): the same graph in plain fp32, weights as given (no bf16 rounding anywhere)
// =====================================================================================================
namespace {

int f32_grid(size_t items) {
  size_t g = (items + 255) / 256;
  if (g > 148 * 32) g = 148 * 32;
  return static_cast<int>(g < 1 ? 1 : g);
}
const std::vector<float>* f32_host(const dlv3p_model* m, const std::string& key) {
  auto it = m->windex.find(key);
  if (it != m->windex.end()) return &m->weights[it->second].data;
  auto ih = m->head_w.find(key);
  return ih == m->head_w.end() ? nullptr : &ih->second;
}
int f32_put(dlv3p_model* m, const std::string& key, const std::vector<float>& h) {
  float* d = nullptr;
  if (cudaMalloc(&d, h.size() * sizeof(float) + 256) != cudaSuccess) return mfail(m, DLV3P_ERR_NOMEM, "fp32 mode: cudaMalloc failed");
  cudaMemcpy(d, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice);
  auto it = m->f32_w.find(key);
  if (it != m->f32_w.end()) cudaFree(it->second);
  m->f32_w[key] = d;
  return 0;
}

int finalize_fp32(dlv3p_model* m) {
  MCU(m, cudaSetDevice(m->device));
  MCU(m, cudaDeviceSynchronize());
  // every variable the whole model expects: backbone inventory + head inventory
  const int n = dlv3p_model_num_weights(m);
  for (int i = 0; i < n; ++i) {
    const char *layer, *var;
    int64_t shape[4];
    int rank;
    dlv3p_model_weight_info(m, i, &layer, &var, shape, &rank);
    const std::string key = std::string(layer) + "/" + var;
    const std::vector<float>* h = f32_host(m, key);
    if (!h || h->empty()) return mfail(m, DLV3P_ERR_STATE, mfmt("weight %s was never set", key.c_str()));
    const std::string v(var);
    if (v == "kernel" || v == "depthwise_kernel" || v == "bias") {
      int r = f32_put(m, key, *h);
      if (r) return r;
    } else if (v == "gamma") {      // fold the BatchNorm once: scale = gamma * rsqrt(var + eps), shift = beta - mean * scale
      const std::string bn(layer);
      const std::vector<float>*be = f32_host(m, bn + "/beta"), *mu = f32_host(m, bn + "/moving_mean"), *va = f32_host(m, bn + "/moving_variance");
      if (!be || !mu || !va || be->empty() || mu->empty() || va->empty()) return mfail(m, DLV3P_ERR_STATE, mfmt("BatchNorm %s is incomplete", bn.c_str()));
      const bool backbone = m->windex.count(key) != 0;
      Fold f = fold_bn(h->data(), be->data(), mu->data(), va->data(), static_cast<int>(h->size()), backbone ? 1e-3f : 1e-5f);
      int r = f32_put(m, bn + "/scale", f.scale);
      if (r || (r = f32_put(m, bn + "/shift", f.shift))) return r;
    }
  }
  m->finalized = true;
  return DLV3P_OK;
}

struct F32Run {
  dlv3p_model* m;
  cudaStream_t st;
  int rc = 0;
  struct T { float* p; int B, H, W, C; int M() const { return B * H * W; } };
  float* buf(size_t n) {
    if (m->f32_next < m->f32_bufs.size() && m->f32_bufs[m->f32_next].second >= n) return m->f32_bufs[m->f32_next++].first;
    float* d = nullptr;
    if (cudaMalloc(&d, n * sizeof(float) + 256) != cudaSuccess) { rc = mfail(m, DLV3P_ERR_NOMEM, "fp32 mode: cudaMalloc failed"); return nullptr; }
    if (m->f32_next < m->f32_bufs.size()) { cudaFree(m->f32_bufs[m->f32_next].first); m->f32_bufs[m->f32_next] = {d, n}; }
    else m->f32_bufs.push_back({d, n});
    ++m->f32_next;
    return d;
  }
  T tensor(int B, int H, int W, int C) { return T{buf(static_cast<size_t>(B) * H * W * C), B, H, W, C}; }
  const float* w(const std::string& key) {
    auto it = m->f32_w.find(key);
    if (it == m->f32_w.end()) { if (!rc) rc = mfail(m, DLV3P_ERR_STATE, mfmt("fp32 mode: missing %s", key.c_str())); return nullptr; }
    return it->second;
  }
  void tap(const std::string& name, const T& t) { m->f32_taps[name] = {t.p, t.B, t.H, t.W, t.C}; }
  void check(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess && !rc) rc = mfail(m, DLV3P_ERR_CUDA, mfmt("fp32 mode launch %s: %s", what, cudaGetErrorString(e)));
    ++m->launches_last;
  }
  T conv(const T& x, const uint8_t* img, const std::string& name, const std::string& bn, int k, int stride, int pad_t, int pad_l, int Cout, int Ho, int Wo) {
    T y = tensor(x.B, Ho, Wo, Cout);
    if (rc) return y;
    F32ConvParams P{x.p, img, w(name + "/kernel"), w(bn + "/scale"), w(bn + "/shift"), y.p, x.B, x.H, x.W, x.C, Ho, Wo, Cout, k, stride, pad_t, pad_l, 1};
    if (!rc) f32_conv_kernel<<<f32_grid(static_cast<size_t>(y.M()) * Cout), 256, 0, st>>>(P);
    check(name.c_str());
    return y;
  }
  T depthwise(const T& x, const std::string& name, const std::string& bn, int stride, int rate, int relu_in, int relu_out) {
    T y = tensor(x.B, cdiv(x.H, stride), cdiv(x.W, stride), x.C);
    if (rc) return y;
    F32DwParams P{x.p, w(name + "/depthwise_kernel"), w(bn + "/scale"), w(bn + "/shift"), y.p, x.B, x.H, x.W, x.C, y.H, y.W, stride, rate, relu_in, relu_out};
    if (!rc) f32_depthwise_kernel<<<f32_grid(static_cast<size_t>(y.M()) * y.C), 256, 0, st>>>(P);
    check(name.c_str());
    return y;
  }
  // y[M, ldy @ col_off] = epi(a[M, K] * W[K, N])
  void gemm(const float* a, int M, int K, int lda, const std::string& name, const float* scale, const float* shift, int relu, const float* residual, float* y,
            int N, int ldy, int col_off, int sub_Ho = 0, int sub_Wo = 0, int sub_H = 0, int sub_W = 0) {
    if (rc) return;
    F32GemmParams P{a, w(name + "/kernel"), scale, shift, residual, y, M, K, N, lda, ldy, col_off, relu, sub_Ho, sub_Wo, sub_H, sub_W};
    if (!rc) f32_gemm_kernel<<<dim3(cdiv(N, 64), cdiv(M, 64)), 256, 0, st>>>(P);
    check(name.c_str());
  }
  T pointwise(const T& x, const std::string& name, const std::string& bn, int N, int relu, const float* residual) {
    T y = tensor(x.B, x.H, x.W, N);
    gemm(x.p, x.M(), x.C, x.C, name, w(bn + "/scale"), w(bn + "/shift"), relu, residual, y.p, N, N, 0);
    return y;
  }
};

int forward_fp32(dlv3p_model* m, const void* d_images, void* d_out, cudaStream_t st) {
  const dlv3p_model_config& g = m->cfg;
  m->launches_last = 0;
  m->f32_next = 0;
  F32Run R{m, st};
  using T = F32Run::T;
  // ---------------- Xception_body (deeplabv3p_xception.py:96-163)
  const int h2 = cdiv(g.H, 2), w2 = cdiv(g.W, 2);
  T img{const_cast<float*>(g.img_dtype == DLV3P_IMG_F32 ? static_cast<const float*>(d_images) : nullptr), g.B, g.H, g.W, 3};
  T x = R.conv(img, g.img_dtype == DLV3P_IMG_U8 ? static_cast<const uint8_t*>(d_images) : nullptr, "entry_flow_conv1_1", "entry_flow_conv1_1_BN", 3, 2,
               m->pad_t, m->pad_l, 32, h2, w2);
  R.tap("entry_flow_conv1_1", x);
  x = R.conv(x, nullptr, "entry_flow_conv1_2", "entry_flow_conv1_2_BN", 3, 1, 1, 1, 64, h2, w2);
  R.tap("entry_flow_conv1_2", x);
  T skip{};
  for (const BlockSpec& b : m->f32_blocks) {
    const T inp = x;
    const float* res = nullptr;
    if (b.shortcut == 0) {      // 1x1 conv (stride-2 sampling inside the GEMM's row addressing) -> BN
      const int ho = cdiv(inp.H, b.stride), wo = cdiv(inp.W, b.stride);
      T sc = R.tensor(inp.B, ho, wo, b.depth[2]);
      R.gemm(inp.p, sc.M(), inp.C, inp.C, b.prefix + "_shortcut", R.w(b.prefix + "_shortcut_BN/scale"), R.w(b.prefix + "_shortcut_BN/shift"), 0, nullptr, sc.p,
             b.depth[2], b.depth[2], 0, b.stride == 2 ? ho : 0, b.stride == 2 ? wo : 0, inp.H, inp.W);
      res = sc.p;
    } else if (b.shortcut == 1) {
      res = inp.p;
    }
    T r = x;
    for (int i = 0; i < 3; ++i) {
      const std::string p = mfmt("%s_separable_conv%d", b.prefix.c_str(), i + 1);
      T d = R.depthwise(r, p + "_depthwise", p + "_depthwise_BN", i == 2 ? b.stride : 1, b.rate, b.act ? 0 : 1, b.act ? 1 : 0);
      r = R.pointwise(d, p + "_pointwise", p + "_pointwise_BN", b.depth[i], b.act ? 1 : 0, i == 2 ? res : nullptr);
      if (i == 1 && b.ret_skip) skip = r;
    }
    x = r;
    R.tap(b.prefix, x);
  }
  R.tap("feature", x);
  R.tap("skip", skip);
  // ---------------- ASPP_block (layers.py:114-163)
  const int M1 = x.M(), npix = x.H * x.W, Cin = x.C;
  int rates[3] = {6, 12, 18};
  if (g.OS == 8) { rates[0] = 12; rates[1] = 24; rates[2] = 36; } else if (g.OS == 32) { rates[0] = 3; rates[1] = 6; rates[2] = 9; }
  T concat = R.tensor(x.B, x.H, x.W, 1280);
  {
    float* pool = R.buf(static_cast<size_t>(x.B) * Cin);
    float* b4 = R.buf(static_cast<size_t>(x.B) * 256);
    if (!R.rc) f32_global_mean_kernel<<<cdiv(x.B * Cin, 256), 256, 0, st>>>(x.p, pool, x.B, npix, Cin);
    R.check("image pooling mean");
    R.gemm(pool, x.B, Cin, Cin, "image_pooling", R.w("image_pooling_BN/scale"), R.w("image_pooling_BN/shift"), 1, nullptr, b4, 256, 256, 0);
    if (!R.rc) f32_bcast_kernel<<<f32_grid(static_cast<size_t>(M1) * 256), 256, 0, st>>>(b4, concat.p, x.B, npix, 256, 1280, 0);   // bilinear resize of a 1x1 map
    R.check("aspp_resize");
  }
  R.gemm(x.p, M1, Cin, Cin, "aspp0", R.w("aspp0_BN/scale"), R.w("aspp0_BN/shift"), 1, nullptr, concat.p, 256, 1280, 256);
  for (int i = 1; i <= 3; ++i) {
    const std::string p = mfmt("aspp%d", i);
    T d = R.depthwise(x, p + "_depthwise", p + "_depthwise_BN", 1, rates[i - 1], 0, 1);
    R.gemm(d.p, M1, Cin, Cin, p + "_pointwise", R.w(p + "_pointwise_BN/scale"), R.w(p + "_pointwise_BN/shift"), 1, nullptr, concat.p, 256, 1280, 256 * (i + 1));
  }
  T aspp = R.tensor(x.B, x.H, x.W, 256);
  R.gemm(concat.p, M1, 1280, 1280, "concat_projection", R.w("concat_projection_BN/scale"), R.w("concat_projection_BN/shift"), 1, nullptr, aspp.p, 256, 256, 0);
  R.tap("aspp_out", aspp);
  // ---------------- Decoder_block (layers.py:199-219)
  T dec_in = R.tensor(skip.B, skip.H, skip.W, 304);
  if (!R.rc) f32_resize_kernel<<<f32_grid(static_cast<size_t>(dec_in.M()) * 256), 256, 0, st>>>(aspp.p, dec_in.p, aspp.B, aspp.H, aspp.W, 256, skip.H, skip.W, 304, 0);
  R.check("decoder_resize");
  R.gemm(skip.p, skip.M(), skip.C, skip.C, "feature_projection0", R.w("feature_projection0_BN/scale"), R.w("feature_projection0_BN/shift"), 1, nullptr, dec_in.p, 48,
         304, 256);
  T d0 = R.depthwise(dec_in, "decoder_conv0_depthwise", "decoder_conv0_depthwise_BN", 1, 1, 0, 1);
  T y0 = R.pointwise(d0, "decoder_conv0_pointwise", "decoder_conv0_pointwise_BN", 256, 1, nullptr);
  T d1 = R.depthwise(y0, "decoder_conv1_depthwise", "decoder_conv1_depthwise_BN", 1, 1, 0, 1);
  T y1 = R.pointwise(d1, "decoder_conv1_pointwise", "decoder_conv1_pointwise_BN", 256, 1, nullptr);
  R.tap("decoder_out", y1);
  // ---------------- tail (model.py:75-86, deeplab.py:99)
  T lg = R.tensor(y1.B, y1.H, y1.W, g.NC);
  R.gemm(y1.p, y1.M(), 256, 256, "conv_upsample", nullptr, R.w("conv_upsample/bias"), 0, nullptr, lg.p, g.NC, g.NC, 0);
  float* planar = g.out_mode == DLV3P_OUT_LOGITS_LOWRES ? static_cast<float*>(d_out) : R.buf(static_cast<size_t>(lg.M()) * g.NC);
  if (!R.rc) f32_to_planar_kernel<<<f32_grid(static_cast<size_t>(lg.M()) * g.NC), 256, 0, st>>>(lg.p, planar, lg.B, lg.H * lg.W, g.NC);
  R.check("logits to planar");
  m->f32_taps["logits"] = {planar, lg.B, g.NC, lg.H, lg.W};
  if (R.rc) return R.rc;
  if (g.out_mode == DLV3P_OUT_LABELS_U8) {
    int r = dlv3p_op_resize_argmax(m->device, planar, lg.B, g.NC, lg.H, lg.W, g.H, g.W, static_cast<uint8_t*>(d_out), st);
    if (r) return mfail(m, r, dlv3p_last_error(nullptr));
    ++m->launches_last;
  }
  return DLV3P_OK;
}

}  // namespace

extern "C" {

const char* dlv3p_model_last_error(const dlv3p_model* m) { return m ? m->err.c_str() : g_model_tls_error.c_str(); }

int dlv3p_model_create(const dlv3p_model_config* cfg, int device, dlv3p_model** out) {
  if (!cfg || !out) return mfail(nullptr, DLV3P_ERR_INVALID, "null argument");
  *out = nullptr;
  dlv3p_model* m = new dlv3p_model();
  m->cfg = *cfg;
  m->device = device;
  m->plan_only = device == -1;
  auto bail = [&](int code, const std::string& msg) {
    mfail(nullptr, code, msg);
    dlv3p_model_destroy(m);
    return code;
  };
  const dlv3p_model_config& g = m->cfg;
  if (g.B < 1 || g.H < 16 || g.W < 16) return bail(DLV3P_ERR_INVALID, "B must be positive, H and W at least 16");
  if (g.NC < 1 || g.NC > 256) return bail(DLV3P_ERR_INVALID, "NC must be in 1..256");
  if (g.img_dtype != DLV3P_IMG_U8 && g.img_dtype != DLV3P_IMG_F32) return bail(DLV3P_ERR_INVALID, "img_dtype: uint8 (0) or fp32 (1)");
  if (g.out_mode < DLV3P_OUT_LABELS_U8 || g.out_mode > DLV3P_OUT_LOGITS_FULL) return bail(DLV3P_ERR_INVALID, "out_mode must be labels / logits / softmax");
  if (!m->plan_only) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return bail(DLV3P_ERR_CUDA, "no CUDA device: libdlv3p has no CPU fallback");
    if (device < 0 || device >= ndev) return bail(DLV3P_ERR_INVALID, mfmt("device %d out of range (%d devices)", device, ndev));
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return bail(DLV3P_ERR_CUDA, "cudaGetDeviceProperties failed");
    if (prop.major != 10) return bail(DLV3P_ERR_UNSUPPORTED, mfmt("device sm_%d%d: kernels are sm_100a only", prop.major, prop.minor));
    m->num_sms = prop.multiProcessorCount;
    if (cudaSetDevice(device) != cudaSuccess) return bail(DLV3P_ERR_CUDA, "cudaSetDevice failed");
  }
  int r = build_plan(m);
  if (r) return bail(r, m->err);
  register_block_weights(m);
  m->fp32 = (g.flags & DLV3P_MODEL_FLAG_FP32) != 0;
  if (m->fp32 && g.out_mode != DLV3P_OUT_LABELS_U8 && g.out_mode != DLV3P_OUT_LOGITS_LOWRES)
    return bail(DLV3P_ERR_UNSUPPORTED, "fp32 precision mode writes labels or low-resolution logits");
  // activations (the fp32 precision mode allocates its own fp32 buffers on its first forward)
  for (Tensor& t : m->tensors) {
    if (m->fp32) break;
    void* p = nullptr;
    if ((r = m_alloc(m, &p, t.elems() * 2, false))) return bail(r, m->err);
    t.p = static_cast<__nv_bfloat16*>(p);
  }
  // head: features [B, h, w, 2048] + skip [B, H/4, W/4, 256] exactly as the backbone leaves them
  dlv3p_config hc{};
  const Tensor& tf = m->tensors[m->t_feat];
  const Tensor& ts = m->tensors[m->t_skip];
  hc.B = g.B; hc.H = g.H; hc.W = g.W; hc.OS = g.OS; hc.h = tf.H; hc.w = tf.W; hc.hs = ts.H; hc.ws = ts.W; hc.Cin = 2048; hc.Cskip = 256; hc.NC = g.NC;
  hc.variant = DLV3P_VARIANT_ASPP; hc.stages = 0; hc.in_dtype = DLV3P_DTYPE_BF16; hc.out_mode = g.out_mode; hc.bn_eps = 0; hc.flags = 0;
  if ((r = dlv3p_create(&hc, device, &m->head))) return bail(r, dlv3p_last_error(nullptr));
  if (!m->plan_only && cudaStreamCreateWithFlags(&m->own_stream, cudaStreamNonBlocking) != cudaSuccess) return bail(DLV3P_ERR_CUDA, "cudaStreamCreate failed");
  *out = m;
  return DLV3P_OK;
}

void dlv3p_model_destroy(dlv3p_model* m) {
  if (!m) return;
  if (m->head) dlv3p_destroy(m->head);
  if (!m->plan_only) {
    cudaSetDevice(m->device);
    for (void* p : m->allocs) cudaFree(p);
    for (void* p : m->weight_allocs) cudaFree(p);
    if (m->d_tm) cudaFree(m->d_tm);
    for (auto& kv : m->f32_w) cudaFree(kv.second);
    for (auto& b : m->f32_bufs) cudaFree(b.first);
    if (m->in_stage) cudaFree(m->in_stage);
    if (m->out_stage) cudaFree(m->out_stage);
    if (m->own_stream) cudaStreamDestroy(m->own_stream);
    for (cudaEvent_t e : m->prof_events) cudaEventDestroy(e);
  }
  delete m;
}

int dlv3p_model_num_weights(const dlv3p_model* m) {
  return m ? static_cast<int>(m->weights.size()) + dlv3p_num_weights(m->head) : DLV3P_ERR_INVALID;
}
int dlv3p_model_weight_info(const dlv3p_model* m, int index, const char** layer, const char** var, int64_t shape_out[4], int* rank_out) {
  if (!m || index < 0) return mfail(nullptr, DLV3P_ERR_INVALID, "bad weight index");
  const int nb = static_cast<int>(m->weights.size());
  if (index >= nb) return dlv3p_weight_info(m->head, index - nb, layer, var, shape_out, rank_out);
  const MWeight& s = m->weights[index];
  if (layer) *layer = s.layer.c_str();
  if (var) *var = s.var.c_str();
  if (rank_out) *rank_out = static_cast<int>(s.shape.size());
  if (shape_out)
    for (size_t i = 0; i < s.shape.size() && i < 4; ++i) shape_out[i] = s.shape[i];
  return DLV3P_OK;
}
int dlv3p_model_set_weight(dlv3p_model* m, const char* layer, const char* var, const float* host, const int64_t* shape, int rank) {
  if (!m || !layer || !var || !host || !shape) return mfail(m, DLV3P_ERR_INVALID, "null argument");
  auto it = m->windex.find(std::string(layer) + "/" + var);
  if (it == m->windex.end()) {
    int r = dlv3p_set_weight(m->head, layer, var, host, shape, rank);
    if (r) return mfail(m, r, dlv3p_last_error(m->head));
    if (m->fp32) {
      size_t n = 1;
      for (int i = 0; i < rank; ++i) n *= static_cast<size_t>(shape[i]);
      std::string ln(layer);
      if (ln == "logits_semantic") ln = "conv_upsample";
      m->head_w[ln + "/" + var].assign(host, host + n);
      m->finalized = false;
    }
    return DLV3P_OK;
  }
  MWeight& s = m->weights[it->second];
  if (rank != static_cast<int>(s.shape.size())) return mfail(m, DLV3P_ERR_NAME, mfmt("%s/%s: rank %d, expected %zu", layer, var, rank, s.shape.size()));
  size_t n = 1;
  for (int i = 0; i < rank; ++i) {
    if (shape[i] != s.shape[i]) return mfail(m, DLV3P_ERR_NAME, mfmt("%s/%s: dim %d is %lld, expected %lld", layer, var, i, (long long)shape[i], (long long)s.shape[i]));
    n *= static_cast<size_t>(shape[i]);
  }
  s.data.assign(host, host + n);
  s.set = true;
  m->finalized = false;
  return DLV3P_OK;
}

int dlv3p_model_finalize_weights(dlv3p_model* m) {
  if (!m) return mfail(nullptr, DLV3P_ERR_INVALID, "null model");
  for (const MWeight& s : m->weights)
    if (!s.set) return mfail(m, DLV3P_ERR_STATE, mfmt("weight %s/%s was never set", s.layer.c_str(), s.var.c_str()));
  if (m->plan_only) return mfail(m, DLV3P_ERR_STATE, "plan-only model (device -1): nothing can be uploaded or run; there is no CPU path");
  if (m->fp32) return finalize_fp32(m);
  int r = dlv3p_finalize_weights(m->head);
  if (r) return mfail(m, r, dlv3p_last_error(m->head));
  if (m->finalized) return DLV3P_OK;
  MCU(m, cudaSetDevice(m->device));
  if (!m->weight_allocs.empty()) {      // a refresh: drain the stream work that reads the old buffers, then release them
    MCU(m, cudaDeviceSynchronize());
    for (void* p : m->weight_allocs) cudaFree(p);
    m->weight_allocs.clear();
  }
  std::string terr;
  m->h_tm.clear();
  auto slot = [&]() { m->h_tm.emplace_back(); return static_cast<int>(m->h_tm.size()) - 1; };
  const dlv3p_model_config& g = m->cfg;
  bool ok = true;
  for (Op& o : m->ops) {
    if (o.kind == OP_STEM) {
      const float* k = MW(m, o.name, "kernel");
      std::vector<float> w(k, k + 27 * 32);
      if ((r = m_upload(m, &o.wf, w))) return r;
      if ((r = m_upload(m, &o.w16, pack_stem_tc(k))) || (r = m_upload(m, &o.dshift, stem_tap_sums(k)))) return r;      // tensor-core stem (uint8 images)
      Fold f = mfold(m, o.bn, 32);
      if ((r = m_upload(m, &o.scale, f.scale)) || (r = m_upload(m, &o.shift, f.shift))) return r;
    } else if (o.kind == OP_CONV3) {
      if ((r = m_upload(m, &o.w16, pack_c3(MW(m, o.name, "kernel"))))) return r;
      Fold f = mfold(m, o.bn, 64);
      o.h_scale = f.scale; o.h_shift = f.shift;
      const Tensor& ti = m->tensors[o.in];
      const Tensor& to = m->tensors[o.out];
      o.tm0 = slot(); slot(); slot();
      ok = ok && tm_nhwc(&m->h_tm[o.tm0], ti.p, ti.B, ti.H, ti.W, 32, 32, kC3HaloW, kC3HaloH, CU_TENSOR_MAP_SWIZZLE_64B, &terr);
      ok = ok && tm_2d(&m->h_tm[o.tm0 + 1], o.w16, 9 * 64, 32, 32, 32, 64, CU_TENSOR_MAP_SWIZZLE_64B, &terr);
      ok = ok && tm_nhwc(&m->h_tm[o.tm0 + 2], to.p, to.B, to.H, to.W, 64, 64, kC3TW, 4, CU_TENSOR_MAP_SWIZZLE_128B, &terr);
    } else if (o.kind == OP_DW) {
      Fold f = mfold(m, o.bn, o.K);
      if ((r = m_upload(m, &o.wf, pack_taps(MW(m, o.name, "depthwise_kernel"), f.scale.data(), o.K, o.Cpad)))) return r;
      std::vector<float> sh(o.Cpad, 0.0f);
      std::memcpy(sh.data(), f.shift.data(), o.K * sizeof(float));
      if ((r = m_upload(m, &o.shift, sh))) return r;
      const Tensor& ti = m->tensors[o.in];
      uint32_t bw, bh;
      dw_box(o.dv, &bw, &bh);
      o.tm0 = slot();
      ok = ok && tm_nhwc(&m->h_tm[o.tm0], ti.p, ti.B, ti.H, ti.W, ti.C, 64, bw, bh, CU_TENSOR_MAP_SWIZZLE_NONE, &terr);
    } else if (o.kind == OP_SEP) {
      Fold fd = mfold(m, o.prefix + "_depthwise_BN", o.K);
      if ((r = m_upload(m, &o.wf, pack_taps(MW(m, o.prefix + "_depthwise", "depthwise_kernel"), fd.scale.data(), o.K, o.K)))) return r;
      if ((r = m_upload(m, &o.shift, fd.shift))) return r;
      if ((r = m_upload(m, &o.w16, pack_kn(MW(m, o.prefix + "_pointwise", "kernel"), o.K, o.N, o.N, o.K)))) return r;
      Fold fp = mfold(m, o.prefix + "_pointwise_BN", o.N);
      o.h_scale.assign(256, 0.0f); o.h_shift.assign(256, 0.0f);
      std::memcpy(o.h_scale.data(), fp.scale.data(), o.N * sizeof(float));
      std::memcpy(o.h_shift.data(), fp.shift.data(), o.N * sizeof(float));
      const Tensor& ti = m->tensors[o.in];
      const Tensor& to = m->tensors[o.out];
      o.tm0 = slot(); slot(); slot();
      ok = ok && tm_nhwc(&m->h_tm[o.tm0], ti.p, ti.B, ti.H, ti.W, ti.C, 64, kDwHaloW, kDwHaloH, CU_TENSOR_MAP_SWIZZLE_NONE, &terr);
      ok = ok && tm_2d(&m->h_tm[o.tm0 + 1], o.w16, o.N, o.K, o.K, 64, o.N / 2, CU_TENSOR_MAP_SWIZZLE_128B, &terr);
      ok = ok && tm_nhwc(&m->h_tm[o.tm0 + 2], to.p, to.B, to.H, to.W, to.C, 32, 16, 2, CU_TENSOR_MAP_SWIZZLE_64B, &terr);
    } else if (o.kind == OP_WIDE) {
      Fold fd = mfold(m, o.prefix + "_depthwise_BN", o.K);
      if ((r = m_upload(m, &o.wf, pack_taps(MW(m, o.prefix + "_depthwise", "depthwise_kernel"), fd.scale.data(), o.K, o.Cpad)))) return r;
      std::vector<float> dsh(o.Cpad, 0.0f);
      std::memcpy(dsh.data(), fd.shift.data(), o.K * sizeof(float));
      float* d_dsh = nullptr;
      if ((r = m_upload(m, &d_dsh, dsh))) return r;
      if ((r = m_upload(m, &o.w16, pack_kn(MW(m, o.prefix + "_pointwise", "kernel"), o.K, o.N, o.N, o.K)))) return r;
      Fold fp = mfold(m, o.prefix + "_pointwise_BN", o.N);
      std::vector<float> s(768, 0.0f), t(768, 0.0f);
      std::memcpy(s.data(), fp.scale.data(), o.N * sizeof(float));
      std::memcpy(t.data(), fp.shift.data(), o.N * sizeof(float));
      if ((r = m_upload(m, &o.scale, s)) || (r = m_upload(m, &o.shift, t))) return r;
      o.dshift = d_dsh;
      const Tensor& ti = m->tensors[o.in];
      const Tensor& to = m->tensors[o.out];
      o.tm0 = slot(); slot(); slot();
      ok = ok && tm_nhwc(&m->h_tm[o.tm0], ti.p, ti.B, ti.H, ti.W, ti.C, 64, kDwHaloW, kDwHaloH, CU_TENSOR_MAP_SWIZZLE_NONE, &terr);
      ok = ok && tm_2d(&m->h_tm[o.tm0 + 1], o.w16, o.N, o.K, o.K, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B, &terr);
      ok = ok && tm_nhwc(&m->h_tm[o.tm0 + 2], to.p, to.B, to.H, to.W, to.C, 32, 16, 2, CU_TENSOR_MAP_SWIZZLE_64B, &terr);
    } else if (o.kind == OP_PW) {
      o.Kpad = cdiv(o.K, 64) * 64;
      o.BN = pick_bb_bn(m->tensors[o.out].M(), o.N, m->num_sms);
      o.Npad = cdiv(o.N, o.BN) * o.BN;
      if ((r = m_upload(m, &o.w16, pack_kn(MW(m, o.name, "kernel"), o.K, o.N, o.Npad, o.Kpad)))) return r;
      Fold f = mfold(m, o.bn, o.N);
      std::vector<float> s(o.Npad, 0.0f), t(o.Npad, 0.0f);
      std::memcpy(s.data(), f.scale.data(), o.N * sizeof(float));
      std::memcpy(t.data(), f.shift.data(), o.N * sizeof(float));
      if ((r = m_upload(m, &o.scale, s)) || (r = m_upload(m, &o.shift, t))) return r;
      const Tensor& ti = m->tensors[o.in];
      const Tensor& to = m->tensors[o.out];
      o.tm0 = slot(); slot(); slot(); slot();
      if (o.a_stride == 2) {
        const uint64_t C = ti.C, bw = to.W < 128 ? to.W : 128;
        const uint64_t d[4] = {C, static_cast<uint64_t>(to.W), static_cast<uint64_t>(to.H), static_cast<uint64_t>(ti.B)};
        const uint64_t sb[3] = {2 * C * 2, 2 * static_cast<uint64_t>(ti.W) * C * 2, static_cast<uint64_t>(ti.H) * ti.W * C * 2};
        const uint32_t bx[4] = {64, static_cast<uint32_t>(bw), static_cast<uint32_t>(128 / bw), 1};
        ok = ok && tm_encode(&m->h_tm[o.tm0], ti.p, 4, d, sb, bx, CU_TENSOR_MAP_SWIZZLE_128B, &terr);
      } else {
        ok = ok && tm_2d(&m->h_tm[o.tm0], ti.p, ti.M(), o.K, o.K, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B, &terr);
      }
      ok = ok && tm_2d(&m->h_tm[o.tm0 + 1], o.w16, o.Npad, o.Kpad, o.Kpad, 64, o.BN / 2, CU_TENSOR_MAP_SWIZZLE_128B, &terr);
      ok = ok && tm_2d(&m->h_tm[o.tm0 + 2], to.p, to.M(), o.N, o.N, 64, 32, CU_TENSOR_MAP_SWIZZLE_128B, &terr);
      if (o.res >= 0) ok = ok && tm_2d(&m->h_tm[o.tm0 + 3], m->tensors[o.res].p, to.M(), o.N, o.N, 64, 32, CU_TENSOR_MAP_SWIZZLE_128B, &terr);
    }
    if (!ok) return mfail(m, DLV3P_ERR_CUDA, terr);
  }
  (void)g;
  if (m->d_tm) { cudaFree(m->d_tm); m->d_tm = nullptr; }
  MCU(m, cudaMalloc(&m->d_tm, m->h_tm.size() * sizeof(CUtensorMap)));
  MCU(m, cudaMemcpy(m->d_tm, m->h_tm.data(), m->h_tm.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice));
  m->finalized = true;
  return DLV3P_OK;
}

int dlv3p_model_input_bytes(const dlv3p_model* m, size_t* bytes) {
  if (!m || !bytes) return mfail(nullptr, DLV3P_ERR_INVALID, "null argument");
  *bytes = static_cast<size_t>(m->cfg.B) * m->cfg.H * m->cfg.W * 3 * (m->cfg.img_dtype == DLV3P_IMG_F32 ? 4 : 1);
  return DLV3P_OK;
}
int dlv3p_model_output_bytes(const dlv3p_model* m, size_t* bytes) {
  if (!m || !bytes) return mfail(nullptr, DLV3P_ERR_INVALID, "null argument");
  return dlv3p_output_bytes(m->head, bytes);
}
int dlv3p_model_workspace_bytes(const dlv3p_model* m, size_t* bytes) {
  if (!m || !bytes) return mfail(nullptr, DLV3P_ERR_INVALID, "null argument");
  size_t hb = 0;
  dlv3p_workspace_bytes(m->head, &hb);
  *bytes = m->ws_bytes + hb;
  return DLV3P_OK;
}
int dlv3p_model_launch_count(const dlv3p_model* m, int64_t* last_forward) {
  if (!m || !last_forward) return mfail(nullptr, DLV3P_ERR_INVALID, "null argument");
  *last_forward = m->launches_last;
  return DLV3P_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------- forward
static int run_backbone(dlv3p_model* m, const void* d_images, cudaStream_t st, int first_op = 0) {
  const dlv3p_model_config& g = m->cfg;
  m->launches_last = 0;
  g_pdl = !(g.flags & DLV3P_MODEL_FLAG_NO_PDL) && !m->profiling;
  for (size_t oi = static_cast<size_t>(first_op); oi < m->ops.size(); ++oi) {
    Op& o = m->ops[oi];
    if (m->profiling) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      cudaEventRecord(e, st);
      m->prof_events.push_back(e);
    }
    cudaError_t e = cudaSuccess;
    if (o.kind == OP_STEM) {
      const Tensor& to = m->tensors[o.out];
      StemParams P{};
      P.img = d_images; P.img_f32 = g.img_dtype == DLV3P_IMG_F32; P.w = o.wf; P.scale = o.scale; P.shift = o.shift; P.out = to.p;
      P.B = g.B; P.H = g.H; P.W = g.W; P.Ho = to.H; P.Wo = to.W; P.pad_t = m->pad_t; P.pad_l = m->pad_l;
      if (g.img_dtype == DLV3P_IMG_F32 || (g.flags & DLV3P_MODEL_FLAG_FP32_STEM)) {
        e = launch_stem(P, m->num_sms, st, g_pdl);      // fp32 FMAs on normalised pixels (float images; A/B flag)
      } else {
        StemTcParams Q{};
        Q.img = static_cast<const uint8_t*>(d_images); Q.w16 = o.w16; Q.wk = o.dshift; Q.scale = o.scale; Q.shift = o.shift; Q.out = to.p;
        Q.B = g.B; Q.H = g.H; Q.W = g.W; Q.Ho = to.H; Q.Wo = to.W; Q.pad_t = m->pad_t; Q.pad_l = m->pad_l;
        e = launch_stem_tc(Q, m->num_sms, st, g_pdl);
      }
    } else if (o.kind == OP_CONV3) {
      const Tensor& to = m->tensors[o.out];
      Conv3x3Params P{};
      P.tmap_x = &m->d_tm[o.tm0]; P.tmap_w = &m->d_tm[o.tm0 + 1]; P.tmap_out = &m->d_tm[o.tm0 + 2];
      std::memcpy(P.scale, o.h_scale.data(), sizeof(P.scale));
      std::memcpy(P.shift, o.h_shift.data(), sizeof(P.shift));
      P.tiles_x = cdiv(to.W, kC3TW); P.tiles_y = cdiv(to.H, kC3TH); P.num_tiles = to.B * P.tiles_x * P.tiles_y;
      e = launch_conv3x3(P, m->num_sms, st);
    } else if (o.kind == OP_DW) {
      const Tensor& to = m->tensors[o.out];
      BbDwParams P{};
      P.tmap_x = &m->d_tm[o.tm0]; P.w = o.wf; P.shift = o.shift; P.out = to.p; P.B = to.B; P.C = to.C; P.Cpad = o.Cpad; P.Ho = to.H; P.Wo = to.W;
      P.tiles_x = cdiv(to.W, o.dv.TW); P.tiles_y = cdiv(to.H, o.dv.TH); P.cgroups = o.Cpad / 64; P.relu_in = o.relu_in; P.relu_out = o.relu_out;
      e = launch_dw(o.dv, P, st);
    } else if (o.kind == OP_SEP) {
      const Tensor& to = m->tensors[o.out];
      BbSepParams Q{};
      DwPwParams& P = Q.base;
      P.tmap_x = &m->d_tm[o.tm0]; P.tmap_x2 = nullptr; P.kb_split = 0; P.tmap_w = &m->d_tm[o.tm0 + 1]; P.dw_w = o.wf; P.dw_shift = o.shift;
      P.scale = nullptr; P.shift = nullptr; P.out = to.p; P.tmap_out = nullptr; P.B = to.B; P.H = to.H; P.W = to.W;
      P.tiles_x = cdiv(to.W, kDwTW); P.tiles_y = cdiv(to.H, kDwTH); P.num_tiles = to.B * P.tiles_x * P.tiles_y; P.debug = 0;
      Q.tmap_out32 = &m->d_tm[o.tm0 + 2];
      std::memcpy(Q.scale_c, o.h_scale.data(), sizeof(Q.scale_c));
      std::memcpy(Q.shift_c, o.h_shift.data(), sizeof(Q.shift_c));
      e = launch_sepconv(o.K / 64, o.N, Q, m->num_sms, st);
    } else if (o.kind == OP_WIDE) {
      const Tensor& to = m->tensors[o.out];
      BbWideParams P{};
      P.tmap_x = &m->d_tm[o.tm0]; P.tmap_w = &m->d_tm[o.tm0 + 1]; P.tmap_out = &m->d_tm[o.tm0 + 2]; P.dw_w = o.wf; P.dw_shift = o.dshift;
      P.scale = o.scale; P.shift = o.shift; P.res = o.res >= 0 ? m->tensors[o.res].p : nullptr; P.B = to.B; P.H = to.H; P.W = to.W; P.N = o.N;
      P.KB = o.Cpad / 64; P.tiles_x = cdiv(to.W, kDwTW); P.tiles_y = cdiv(to.H, kDwTH); P.num_tiles = to.B * P.tiles_x * P.tiles_y;
      P.relu_out = o.relu_out; P.debug = 0;
      e = launch_sepwide(P, o.relu_in != 0, m->num_sms, st);
    } else if (o.kind == OP_PW) {
      const Tensor& to = m->tensors[o.out];
      BbGemmParams P{};
      P.tmap_a = &m->d_tm[o.tm0]; P.tmap_w = &m->d_tm[o.tm0 + 1]; P.tmap_out = &m->d_tm[o.tm0 + 2];
      P.scale = o.scale; P.shift = o.shift; P.tmap_res = o.res >= 0 ? &m->d_tm[o.tm0 + 3] : nullptr;
      P.M = to.M(); P.K = o.K; P.N = o.N; P.relu = o.relu_out; P.m_pairs = cdiv(cdiv(P.M, kBbBM), 2); P.n_tiles = o.Npad / o.BN;
      if (o.a_stride == 2) { P.a_wo = to.W; P.a_hw = to.H * to.W; }
      e = launch_bb_gemm(o.BN, P, m->num_sms, st);
    } else {
      const Tensor& ti = m->tensors[o.in];
      const Tensor& to = m->tensors[o.out];
      const size_t n = to.elems() / 8;
      size_t gsz = (n + 255) / 256;
      if (gsz > static_cast<size_t>(m->num_sms) * 16) gsz = static_cast<size_t>(m->num_sms) * 16;
      e = launch_pdl(g_pdl, subsample2_kernel, dim3(static_cast<unsigned>(gsz)), dim3(256), 0, st, static_cast<const __nv_bfloat16*>(ti.p), to.p, ti.B, ti.H, ti.W,
                     ti.C, to.H, to.W);
    }
    if (e != cudaSuccess) return mfail(m, DLV3P_ERR_CUDA, mfmt("launch %s: %s", o.name.c_str(), cudaGetErrorString(e)));
    ++m->launches_last;
  }
  return DLV3P_OK;
}

static int model_forward_impl(dlv3p_model* m, const void* d_images, void* d_out, cudaStream_t st) {
  if (!m) return mfail(nullptr, DLV3P_ERR_INVALID, "null model");
  if (m->plan_only) return mfail(m, DLV3P_ERR_STATE, "plan-only model (device -1): there is no CPU path");
  if (!m->finalized) return mfail(m, DLV3P_ERR_STATE, "dlv3p_model_finalize_weights has not been called");
  if (!d_images || !d_out) return mfail(m, DLV3P_ERR_INVALID, "null image / output pointer");
  MCU(m, cudaSetDevice(m->device));
  if (m->fp32) return forward_fp32(m, d_images, d_out, st);
  int r = run_backbone(m, d_images, st);
  if (r) return r;
  r = dlv3p_forward(m->head, m->tensors[m->t_feat].p, m->tensors[m->t_skip].p, d_out, st);
  if (r) return mfail(m, r, dlv3p_last_error(m->head));
  int64_t hl = 0;
  dlv3p_launch_count(m->head, &hl, nullptr);
  m->launches_last += hl;
  return DLV3P_OK;
}

extern "C" {

int dlv3p_model_forward(dlv3p_model* m, const void* d_images, void* d_out, void* cuda_stream) {
  return model_forward_impl(m, d_images, d_out, static_cast<cudaStream_t>(cuda_stream));
}

int dlv3p_model_forward_host(dlv3p_model* m, const void* h_images, void* h_out) {
  if (!m || !h_images || !h_out) return mfail(m, DLV3P_ERR_INVALID, "null argument");
  if (m->plan_only) return mfail(m, DLV3P_ERR_STATE, "plan-only model (device -1): there is no CPU path");
  MCU(m, cudaSetDevice(m->device));
  size_t ib = 0, ob = 0;
  dlv3p_model_input_bytes(m, &ib);
  int r = dlv3p_model_output_bytes(m, &ob);
  if (r) return r;
  if (!m->in_stage) MCU(m, cudaMalloc(&m->in_stage, ib));
  if (!m->out_stage) MCU(m, cudaMalloc(&m->out_stage, ob));
  cudaStream_t st = m->own_stream;
  MCU(m, cudaMemcpyAsync(m->in_stage, h_images, ib, cudaMemcpyHostToDevice, st));
  r = model_forward_impl(m, m->in_stage, m->out_stage, st);
  if (r) return r;
  MCU(m, cudaMemcpyAsync(h_out, m->out_stage, ob, cudaMemcpyDeviceToHost, st));
  MCU(m, cudaStreamSynchronize(st));
  return DLV3P_OK;
}

int dlv3p_model_profile_forward(dlv3p_model* m, const void* d_images, void* d_out, void* cuda_stream, const char** names_out, float* ms_out,
                                double* flops_out, double* bytes_out, int max) {
  if (!m) return mfail(nullptr, DLV3P_ERR_INVALID, "null model");
  if (m->plan_only || !m->finalized) return mfail(m, DLV3P_ERR_STATE, "model is not ready");
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  MCU(m, cudaSetDevice(m->device));
  for (cudaEvent_t e : m->prof_events) cudaEventDestroy(e);
  m->prof_events.clear();
  m->profiling = true;
  int r = run_backbone(m, d_images, st);
  m->profiling = false;
  if (r) return r;
  cudaEvent_t e;
  cudaEventCreate(&e);
  cudaEventRecord(e, st);
  m->prof_events.push_back(e);
  const char* hn[64];
  float hms[64];
  const int nh = dlv3p_profile_forward(m->head, m->tensors[m->t_feat].p, m->tensors[m->t_skip].p, d_out, st, hn, hms, 64);
  if (nh < 0) return mfail(m, nh, dlv3p_last_error(m->head));
  MCU(m, cudaStreamSynchronize(st));
  m->launches_last += nh;
  int n = 0;
  for (size_t i = 0; i < m->ops.size() && n < max; ++i, ++n) {
    float ms = 0;
    cudaEventElapsedTime(&ms, m->prof_events[i], m->prof_events[i + 1]);
    if (names_out) names_out[n] = m->ops[i].name.c_str();
    if (ms_out) ms_out[n] = ms;
    if (flops_out) flops_out[n] = m->ops[i].flops;
    if (bytes_out) bytes_out[n] = m->ops[i].bytes;
  }
  for (int i = 0; i < nh && n < max; ++i, ++n) {
    if (names_out) names_out[n] = hn[i];
    if (ms_out) ms_out[n] = hms[i];
    if (flops_out) flops_out[n] = 0;
    if (bytes_out) bytes_out[n] = 0;
  }
  return n;
}

int dlv3p_model_forward_from(dlv3p_model* m, const char* tap, const float* host_fp32, size_t host_elems, void* d_out, void* cuda_stream) {
  if (!m || !tap || !host_fp32 || !d_out) return mfail(m, DLV3P_ERR_INVALID, "null argument");
  if (m->plan_only || !m->finalized) return mfail(m, DLV3P_ERR_STATE, "model is not ready");
  auto it = m->taps.find(tap);
  if (it == m->taps.end()) return mfail(m, DLV3P_ERR_NAME, mfmt("no backbone tap named %s", tap));
  const int ti = it->second;
  const Tensor& t = m->tensors[ti];
  if (host_elems != t.elems()) return mfail(m, DLV3P_ERR_INVALID, mfmt("tap %s has %zu elements, got %zu", tap, t.elems(), host_elems));
  int first = -1;
  for (size_t i = 0; i < m->ops.size(); ++i)
    if (m->ops[i].out == ti) first = static_cast<int>(i) + 1;      // the LAST writer (pooled tensors are written more than once)
  if (first < 0) return mfail(m, DLV3P_ERR_NAME, mfmt("tap %s is not written by a backbone kernel", tap));
  MCU(m, cudaSetDevice(m->device));
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  std::vector<uint16_t> tmp(t.elems());
  for (size_t i = 0; i < tmp.size(); ++i) tmp[i] = bf16_rne(host_fp32[i]);
  MCU(m, cudaMemcpyAsync(t.p, tmp.data(), tmp.size() * 2, cudaMemcpyHostToDevice, st));
  MCU(m, cudaStreamSynchronize(st));
  int r = run_backbone(m, nullptr, st, first);
  if (r) return r;
  r = dlv3p_forward(m->head, m->tensors[m->t_feat].p, m->tensors[m->t_skip].p, d_out, st);
  if (r) return mfail(m, r, dlv3p_last_error(m->head));
  return DLV3P_OK;
}

int dlv3p_model_tap_shape(const dlv3p_model* m, const char* name, int64_t shape_out[4]) {
  if (!m || !name || !shape_out) return mfail(nullptr, DLV3P_ERR_INVALID, "null argument");
  auto it = m->taps.find(name);
  if (it == m->taps.end()) return mfail(nullptr, DLV3P_ERR_NAME, mfmt("no backbone tap named %s", name));
  const Tensor& t = m->tensors[it->second];
  shape_out[0] = t.B; shape_out[1] = t.H; shape_out[2] = t.W; shape_out[3] = t.C;
  return DLV3P_OK;
}

int dlv3p_model_read_tap(dlv3p_model* m, const char* name, float* host_out, size_t host_elems) {
  if (!m || !name || !host_out) return mfail(m, DLV3P_ERR_INVALID, "null argument");
  if (m->plan_only) return mfail(m, DLV3P_ERR_STATE, "plan-only model (device -1): there is no CPU path");
  if (m->fp32) {
    auto ft = m->f32_taps.find(name);
    if (ft == m->f32_taps.end()) return mfail(m, DLV3P_ERR_NAME, mfmt("no fp32 tap named %s (run a forward first)", name));
    const size_t n = static_cast<size_t>(ft->second.B) * ft->second.H * ft->second.W * ft->second.C;
    if (host_elems < n) return mfail(m, DLV3P_ERR_INVALID, mfmt("tap %s needs %zu elements, buffer has %zu", name, n, host_elems));
    MCU(m, cudaSetDevice(m->device));
    MCU(m, cudaDeviceSynchronize());
    MCU(m, cudaMemcpy(host_out, ft->second.p, n * sizeof(float), cudaMemcpyDeviceToHost));
    return DLV3P_OK;
  }
  auto it = m->taps.find(name);
  if (it == m->taps.end()) {
    int r = dlv3p_read_tap(m->head, name, host_out, host_elems);
    if (r) return mfail(m, r, dlv3p_last_error(m->head));
    return DLV3P_OK;
  }
  const Tensor& t = m->tensors[it->second];
  if (host_elems < t.elems()) return mfail(m, DLV3P_ERR_INVALID, mfmt("tap %s needs %zu elements, buffer has %zu", name, t.elems(), host_elems));
  MCU(m, cudaSetDevice(m->device));
  MCU(m, cudaDeviceSynchronize());
  std::vector<uint16_t> tmp(t.elems());
  MCU(m, cudaMemcpy(tmp.data(), t.p, tmp.size() * 2, cudaMemcpyDeviceToHost));
  for (size_t i = 0; i < tmp.size(); ++i) host_out[i] = bf16_f32(tmp[i]);
  return DLV3P_OK;
}

}  // extern "C"

// =====================================================================================================
// standalone operators (unit parity tests)
// =====================================================================================================
namespace {
struct Tmp {
  std::vector<void*> ptrs;
  ~Tmp() {
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
};
int op_begin(int device, int* sms) {
  MCU(nullptr, cudaSetDevice(device));
  cudaDeviceProp p;
  MCU(nullptr, cudaGetDeviceProperties(&p, device));
  if (p.major != 10) return mfail(nullptr, DLV3P_ERR_UNSUPPORTED, mfmt("device sm_%d%d: kernels are sm_100a only", p.major, p.minor));
  *sms = p.multiProcessorCount;
  return 0;
}
}  // namespace

extern "C" {

int dlv3p_op_bb_depthwise(int device, const void* x, int B, int H, int W, int C, int stride, int rate, int relu_in, int relu_out, const float* w_hwc,
                          const float* scale, const float* shift, void* out, void* cuda_stream) {
  int sms = 0, r = op_begin(device, &sms);
  if (r) return r;
  DwVariant v;
  if (!x || !w_hwc || !out || B < 1 || H < 1 || W < 1 || C < 8 || C % 8) return mfail(nullptr, DLV3P_ERR_INVALID, "op_bb_depthwise: bad arguments (C % 8)");
  if (!dw_variant(stride, rate, &v)) return mfail(nullptr, DLV3P_ERR_UNSUPPORTED, "op_bb_depthwise: (stride, rate) in {(1,1), (2,1), (1,2), (1,4)}");
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  const int Cpad = cdiv(C, 64) * 64, Ho = cdiv(H, stride), Wo = cdiv(W, stride);
  Tmp tmp;
  float* dw = tmp.put(pack_taps(w_hwc, scale, C, Cpad));
  std::vector<float> sh(Cpad, 0.0f);
  if (shift) std::memcpy(sh.data(), shift, C * sizeof(float));
  float* dsh = tmp.put(sh);
  std::string terr;
  std::vector<CUtensorMap> tm(1);
  uint32_t bw, bh;
  dw_box(v, &bw, &bh);
  if (!tm_nhwc(&tm[0], x, B, H, W, C, 64, bw, bh, CU_TENSOR_MAP_SWIZZLE_NONE, &terr)) return mfail(nullptr, DLV3P_ERR_CUDA, terr);
  CUtensorMap* dtm = tmp.put(tm);
  if (!dw || !dsh || !dtm) return mfail(nullptr, DLV3P_ERR_NOMEM, "op_bb_depthwise: cudaMalloc failed");
  BbDwParams P{};
  P.tmap_x = dtm; P.w = dw; P.shift = dsh; P.out = static_cast<__nv_bfloat16*>(out); P.B = B; P.C = C; P.Cpad = Cpad; P.Ho = Ho; P.Wo = Wo;
  P.tiles_x = cdiv(Wo, v.TW); P.tiles_y = cdiv(Ho, v.TH); P.cgroups = Cpad / 64; P.relu_in = relu_in; P.relu_out = relu_out;
  MCU(nullptr, launch_dw(v, P, st));
  MCU(nullptr, cudaStreamSynchronize(st));
  return DLV3P_OK;
}

int dlv3p_op_bb_pointwise(int device, const void* a, int64_t M, int K, int N, const float* w_kn, const float* scale, const float* shift, int relu,
                          const void* residual, void* out, void* cuda_stream) {
  int sms = 0, r = op_begin(device, &sms);
  if (r) return r;
  if (!a || !w_kn || !out || M < 1 || K < 8 || K % 8 || N < 8 || N % 8) return mfail(nullptr, DLV3P_ERR_INVALID, "op_bb_pointwise: bad arguments (K % 8, N % 8)");
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  const int BN = pick_bb_bn(static_cast<int>(M), N, sms);
  const int Kpad = cdiv(K, 64) * 64, Npad = cdiv(N, BN) * BN;
  Tmp tmp;
  uint16_t* dw = tmp.put(pack_kn(w_kn, K, N, Npad, Kpad));
  std::vector<float> s(Npad, 0.0f), t(Npad, 0.0f);
  for (int i = 0; i < N; ++i) { s[i] = scale ? scale[i] : 1.0f; t[i] = shift ? shift[i] : 0.0f; }
  float* ds = tmp.put(s);
  float* dt = tmp.put(t);
  std::string terr;
  std::vector<CUtensorMap> tm(4);
  if (!tm_2d(&tm[0], a, M, K, K, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B, &terr) || !tm_2d(&tm[1], dw, Npad, Kpad, Kpad, 64, BN / 2, CU_TENSOR_MAP_SWIZZLE_128B, &terr) ||
      !tm_2d(&tm[2], out, M, N, N, 64, 32, CU_TENSOR_MAP_SWIZZLE_128B, &terr) ||
      (residual && !tm_2d(&tm[3], residual, M, N, N, 64, 32, CU_TENSOR_MAP_SWIZZLE_128B, &terr))) return mfail(nullptr, DLV3P_ERR_CUDA, terr);
  CUtensorMap* dtm = tmp.put(tm);
  if (!dw || !ds || !dt || !dtm) return mfail(nullptr, DLV3P_ERR_NOMEM, "op_bb_pointwise: cudaMalloc failed");
  BbGemmParams P{};
  P.tmap_a = &dtm[0]; P.tmap_w = &dtm[1]; P.tmap_out = &dtm[2]; P.tmap_res = residual ? &dtm[3] : nullptr; P.scale = ds; P.shift = dt;
  P.M = static_cast<int>(M); P.K = K; P.N = N; P.relu = relu; P.m_pairs = cdiv(cdiv(P.M, kBbBM), 2); P.n_tiles = Npad / BN;
  MCU(nullptr, launch_bb_gemm(BN, P, sms, st));
  MCU(nullptr, cudaStreamSynchronize(st));
  return DLV3P_OK;
}

int dlv3p_op_bb_sepwide(int device, const void* x, int B, int H, int W, int C, int N, int relu_in, const float* dw_hwc, const float* dw_scale,
                        const float* dw_shift, const float* w_kn, const float* scale, const float* shift, int relu_out, const void* residual, void* out,
                        void* cuda_stream) {
  int sms = 0, r = op_begin(device, &sms);
  if (r) return r;
  if (!x || !dw_hwc || !w_kn || !out || B < 1 || H < 1 || W < 1) return mfail(nullptr, DLV3P_ERR_INVALID, "op_bb_sepwide: bad arguments");
  if (!sepwide_supported(C, N, 1, 1, false)) return mfail(nullptr, DLV3P_ERR_UNSUPPORTED, "op_bb_sepwide: C % 8 == 0, 64 <= C <= 768, N % 8 == 0, 656 < N <= 768");
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  const int Cpad = cdiv(C, 64) * 64;
  Tmp tmp;
  float* dw = tmp.put(pack_taps(dw_hwc, dw_scale, C, Cpad));
  std::vector<float> sh(Cpad, 0.0f), s(768, 0.0f), t(768, 0.0f);
  if (dw_shift) std::memcpy(sh.data(), dw_shift, C * sizeof(float));
  for (int i = 0; i < N; ++i) { s[i] = scale ? scale[i] : 1.0f; t[i] = shift ? shift[i] : 0.0f; }
  float* dsh = tmp.put(sh);
  float* ds = tmp.put(s);
  float* dt = tmp.put(t);
  uint16_t* w16 = tmp.put(pack_kn(w_kn, C, N, N, C));
  std::string terr;
  std::vector<CUtensorMap> tm(3);
  if (!tm_nhwc(&tm[0], x, B, H, W, C, 64, kDwHaloW, kDwHaloH, CU_TENSOR_MAP_SWIZZLE_NONE, &terr) || !tm_2d(&tm[1], w16, N, C, C, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B, &terr) ||
      !tm_nhwc(&tm[2], out, B, H, W, N, 32, 16, 2, CU_TENSOR_MAP_SWIZZLE_64B, &terr)) return mfail(nullptr, DLV3P_ERR_CUDA, terr);
  CUtensorMap* dtm = tmp.put(tm);
  if (!dw || !dsh || !ds || !dt || !w16 || !dtm) return mfail(nullptr, DLV3P_ERR_NOMEM, "op_bb_sepwide: cudaMalloc failed");
  BbWideParams P{};
  P.tmap_x = &dtm[0]; P.tmap_w = &dtm[1]; P.tmap_out = &dtm[2]; P.dw_w = dw; P.dw_shift = dsh; P.scale = ds; P.shift = dt;
  P.res = static_cast<const __nv_bfloat16*>(residual); P.B = B; P.H = H; P.W = W; P.N = N; P.KB = Cpad / 64;
  P.tiles_x = cdiv(W, kDwTW); P.tiles_y = cdiv(H, kDwTH); P.num_tiles = B * P.tiles_x * P.tiles_y; P.relu_out = relu_out;
  MCU(nullptr, launch_sepwide(P, relu_in != 0, sms, st));
  MCU(nullptr, cudaStreamSynchronize(st));
  return DLV3P_OK;
}

// Benchmark aid (tools/kbench_bb.py): average ms per launch of ONE backbone operator on synthetic device data (CUDA events).
// op 0: pointwise GEMM {M, K, N, residual(0/1), BN (0 = pick)}; op 1: depthwise {B, H, W, C, stride, rate}.
// flags: the kernels' debug bits (GEMM bit0 = skip the stores); results are then meaningless.
int dlv3p_op_bb_time(int device, int op, const int64_t* d, int ndims, int iters, int flags, float* ms_out) {
  int sms = 0, r = op_begin(device, &sms);
  if (r) return r;
  if (!d || !ms_out || iters < 1) return mfail(nullptr, DLV3P_ERR_INVALID, "op_bb_time: bad arguments");
  Tmp tmp;
  cudaStream_t st = nullptr;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  std::string terr;
  auto fill = [&](size_t n) -> uint16_t* {
    std::vector<uint16_t> h(n);
    uint32_t sd = 12345u;
    for (size_t i = 0; i < n; ++i) { sd = sd * 1664525u + 1013904223u; h[i] = bf16_rne(static_cast<float>((sd >> 16) & 0xFF) / 128.0f - 1.0f); }
    return tmp.put(h);
  };
  auto run = [&](auto&& launch) -> int {
    for (int i = 0; i < 3; ++i) launch();
    if (cudaStreamSynchronize(st) != cudaSuccess) return mfail(nullptr, DLV3P_ERR_CUDA, mfmt("op_bb_time warmup: %s", cudaGetErrorString(cudaGetLastError())));
    cudaEventRecord(e0, st);
    for (int i = 0; i < iters; ++i) launch();
    cudaEventRecord(e1, st);
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return mfail(nullptr, DLV3P_ERR_CUDA, mfmt("op_bb_time: %s", cudaGetErrorString(e)));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    *ms_out = ms / iters;
    return DLV3P_OK;
  };
  int rc = DLV3P_ERR_INVALID;
  g_pdl = false;
  if (op == 0 && ndims >= 3) {
    const int M = static_cast<int>(d[0]), K = static_cast<int>(d[1]), N = static_cast<int>(d[2]);
    const bool res = ndims > 3 && d[3];
    const int BN = (ndims > 4 && d[4]) ? static_cast<int>(d[4]) : pick_bb_bn(M, N, sms);
    const int Kpad = cdiv(K, 64) * 64, Npad = cdiv(N, BN) * BN;
    uint16_t* a = fill(static_cast<size_t>(M) * K);
    uint16_t* w = fill(static_cast<size_t>(Npad) * Kpad);
    uint16_t* rr = fill(static_cast<size_t>(M) * N);
    uint16_t* o = fill(static_cast<size_t>(M) * N);
    float* s = tmp.put(std::vector<float>(Npad, 1.0f));
    float* t = tmp.put(std::vector<float>(Npad, 0.0f));
    std::vector<CUtensorMap> tm(4);
    if (!tm_2d(&tm[0], a, M, K, K, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B, &terr) || !tm_2d(&tm[1], w, Npad, Kpad, Kpad, 64, BN / 2, CU_TENSOR_MAP_SWIZZLE_128B, &terr) ||
        !tm_2d(&tm[2], o, M, N, N, 64, 32, CU_TENSOR_MAP_SWIZZLE_128B, &terr) || !tm_2d(&tm[3], rr, M, N, N, 64, 32, CU_TENSOR_MAP_SWIZZLE_128B, &terr))
      return mfail(nullptr, DLV3P_ERR_CUDA, terr);
    CUtensorMap* dtm = tmp.put(tm);
    BbGemmParams P{};
    P.tmap_a = &dtm[0]; P.tmap_w = &dtm[1]; P.tmap_out = &dtm[2]; P.tmap_res = res ? &dtm[3] : nullptr; P.scale = s; P.shift = t;
    P.M = M; P.K = K; P.N = N; P.relu = 0; P.m_pairs = cdiv(cdiv(M, kBbBM), 2); P.n_tiles = Npad / BN; P.debug = flags;
    rc = run([&] { launch_bb_gemm(BN, P, sms, st); });
  } else if (op == 1 && ndims >= 6) {
    const int B = static_cast<int>(d[0]), H = static_cast<int>(d[1]), W = static_cast<int>(d[2]), C = static_cast<int>(d[3]);
    DwVariant v;
    if (!dw_variant(static_cast<int>(d[4]), static_cast<int>(d[5]), &v)) return mfail(nullptr, DLV3P_ERR_UNSUPPORTED, "op_bb_time: depthwise variant");
    if (ndims > 6 && d[6] == 4 && v.S == 1 && v.R == 1) v.TH = 4;      // experiment: half-height tiles, four CTAs per SM
    const int Cpad = cdiv(C, 64) * 64, Ho = cdiv(H, v.S), Wo = cdiv(W, v.S);
    uint16_t* x = fill(static_cast<size_t>(B) * H * W * C);
    uint16_t* o = fill(static_cast<size_t>(B) * Ho * Wo * C);
    float* w = tmp.put(std::vector<float>(static_cast<size_t>(9) * Cpad, 0.1f));
    float* sh = tmp.put(std::vector<float>(Cpad, 0.0f));
    std::vector<CUtensorMap> tm(1);
    uint32_t bw, bh;
    dw_box(v, &bw, &bh);
    if (!tm_nhwc(&tm[0], x, B, H, W, C, 64, bw, bh, CU_TENSOR_MAP_SWIZZLE_NONE, &terr)) return mfail(nullptr, DLV3P_ERR_CUDA, terr);
    CUtensorMap* dtm = tmp.put(tm);
    BbDwParams P{};
    P.tmap_x = dtm; P.w = w; P.shift = sh; P.out = reinterpret_cast<__nv_bfloat16*>(o); P.B = B; P.C = C; P.Cpad = Cpad; P.Ho = Ho; P.Wo = Wo;
    P.tiles_x = cdiv(Wo, v.TW); P.tiles_y = cdiv(Ho, v.TH); P.cgroups = Cpad / 64; P.relu_in = 1; P.relu_out = 0; P.debug = flags;
    rc = run([&] { launch_dw(v, P, st); });
  } else if (op == 2 && ndims >= 5) {      // fused middle-flow SepConv_BN {B, H, W, C, N, residual}
    const int B = static_cast<int>(d[0]), H = static_cast<int>(d[1]), W = static_cast<int>(d[2]), C = static_cast<int>(d[3]), N = static_cast<int>(d[4]);
    if (!sepwide_supported(C, N, 1, 1, false)) return mfail(nullptr, DLV3P_ERR_UNSUPPORTED, "op_bb_time: sepwide shape");
    const bool res = ndims > 5 && d[5];
    const int Cpad = cdiv(C, 64) * 64;
    uint16_t* x = fill(static_cast<size_t>(B) * H * W * C);
    uint16_t* o = fill(static_cast<size_t>(B) * H * W * N);
    uint16_t* rr = fill(static_cast<size_t>(B) * H * W * N);
    uint16_t* w = fill(static_cast<size_t>(N) * C);
    float* dw = tmp.put(std::vector<float>(static_cast<size_t>(9) * Cpad, 0.1f));
    float* sh = tmp.put(std::vector<float>(Cpad, 0.0f));
    float* s = tmp.put(std::vector<float>(768, 1.0f));
    float* t = tmp.put(std::vector<float>(768, 0.0f));
    std::vector<CUtensorMap> tm(3);
    if (!tm_nhwc(&tm[0], x, B, H, W, C, 64, kDwHaloW, kDwHaloH, CU_TENSOR_MAP_SWIZZLE_NONE, &terr) || !tm_2d(&tm[1], w, N, C, C, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B, &terr) ||
        !tm_nhwc(&tm[2], o, B, H, W, N, 32, 16, 2, CU_TENSOR_MAP_SWIZZLE_64B, &terr)) return mfail(nullptr, DLV3P_ERR_CUDA, terr);
    CUtensorMap* dtm = tmp.put(tm);
    BbWideParams P{};
    P.tmap_x = &dtm[0]; P.tmap_w = &dtm[1]; P.tmap_out = &dtm[2]; P.dw_w = dw; P.dw_shift = sh; P.scale = s; P.shift = t;
    P.res = res ? reinterpret_cast<const __nv_bfloat16*>(rr) : nullptr; P.B = B; P.H = H; P.W = W; P.N = N; P.KB = Cpad / 64;
    P.tiles_x = cdiv(W, kDwTW); P.tiles_y = cdiv(H, kDwTH); P.num_tiles = B * P.tiles_x * P.tiles_y; P.relu_out = 0; P.debug = flags;
    rc = run([&] { launch_sepwide(P, true, sms, st); });
  } else {
    rc = mfail(nullptr, DLV3P_ERR_INVALID, "op_bb_time: unknown op / too few dims");
  }
  g_pdl = true;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return rc;
}

int dlv3p_op_conv3x3_c32(int device, const void* x, int B, int H, int W, const float* w_hwio, const float* scale, const float* shift, void* out,
                         void* cuda_stream) {
  int sms = 0, r = op_begin(device, &sms);
  if (r) return r;
  if (!x || !w_hwio || !out || B < 1 || H < 1 || W < 1) return mfail(nullptr, DLV3P_ERR_INVALID, "op_conv3x3_c32: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  Tmp tmp;
  uint16_t* dw = tmp.put(pack_c3(w_hwio));
  std::string terr;
  std::vector<CUtensorMap> tm(3);
  if (!tm_nhwc(&tm[0], x, B, H, W, 32, 32, kC3HaloW, kC3HaloH, CU_TENSOR_MAP_SWIZZLE_64B, &terr) || !tm_2d(&tm[1], dw, 9 * 64, 32, 32, 32, 64, CU_TENSOR_MAP_SWIZZLE_64B, &terr) ||
      !tm_nhwc(&tm[2], out, B, H, W, 64, 64, kC3TW, 4, CU_TENSOR_MAP_SWIZZLE_128B, &terr)) return mfail(nullptr, DLV3P_ERR_CUDA, terr);
  CUtensorMap* dtm = tmp.put(tm);
  if (!dw || !dtm) return mfail(nullptr, DLV3P_ERR_NOMEM, "op_conv3x3_c32: cudaMalloc failed");
  Conv3x3Params P{};
  P.tmap_x = &dtm[0]; P.tmap_w = &dtm[1]; P.tmap_out = &dtm[2];
  for (int i = 0; i < 64; ++i) { P.scale[i] = scale ? scale[i] : 1.0f; P.shift[i] = shift ? shift[i] : 0.0f; }
  P.tiles_x = cdiv(W, kC3TW); P.tiles_y = cdiv(H, kC3TH); P.num_tiles = B * P.tiles_x * P.tiles_y;
  MCU(nullptr, launch_conv3x3(P, sms, st));
  MCU(nullptr, cudaStreamSynchronize(st));
  return DLV3P_OK;
}

int dlv3p_op_stem_conv(int device, const void* img, int img_dtype, int B, int H, int W, const float* w_hwio, const float* scale, const float* shift,
                       void* out, void* cuda_stream) {
  int sms = 0, r = op_begin(device, &sms);
  if (r) return r;
  if (!img || !w_hwio || !out || B < 1 || H < 1 || W < 1) return mfail(nullptr, DLV3P_ERR_INVALID, "op_stem_conv: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  Tmp tmp;
  float* dw = tmp.put(std::vector<float>(w_hwio, w_hwio + 27 * 32));
  std::vector<float> s(32, 1.0f), t(32, 0.0f);
  if (scale) s.assign(scale, scale + 32);
  if (shift) t.assign(shift, shift + 32);
  float* ds = tmp.put(s);
  float* dt = tmp.put(t);
  if (!dw || !ds || !dt) return mfail(nullptr, DLV3P_ERR_NOMEM, "op_stem_conv: cudaMalloc failed");
  StemParams P{};
  P.img = img; P.img_f32 = img_dtype == DLV3P_IMG_F32; P.w = dw; P.scale = ds; P.shift = dt; P.out = static_cast<__nv_bfloat16*>(out);
  P.B = B; P.H = H; P.W = W; P.Ho = cdiv(H, 2); P.Wo = cdiv(W, 2);
  P.pad_t = std::max((P.Ho - 1) * 2 + 3 - H, 0) / 2;
  P.pad_l = std::max((P.Wo - 1) * 2 + 3 - W, 0) / 2;
  if (P.img_f32) {
    MCU(nullptr, launch_stem(P, sms, st, false));
  } else {
    uint16_t* d16 = tmp.put(pack_stem_tc(w_hwio));
    float* dwk = tmp.put(stem_tap_sums(w_hwio));
    if (!d16 || !dwk) return mfail(nullptr, DLV3P_ERR_NOMEM, "op_stem_conv: cudaMalloc failed");
    StemTcParams Q{};
    Q.img = static_cast<const uint8_t*>(img); Q.w16 = d16; Q.wk = dwk; Q.scale = ds; Q.shift = dt; Q.out = static_cast<__nv_bfloat16*>(out);
    Q.B = B; Q.H = H; Q.W = W; Q.Ho = P.Ho; Q.Wo = P.Wo; Q.pad_t = P.pad_t; Q.pad_l = P.pad_l;
    MCU(nullptr, launch_stem_tc(Q, sms, st, false));
  }
  MCU(nullptr, cudaStreamSynchronize(st));
  return DLV3P_OK;
}

}  // extern "C"
