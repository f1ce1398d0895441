// trainer_api.cuh — the WHOLE data-parallel training step of the head behind the C ABI (include/dlv3p_train.h, dlv3p_trainer_*):
// forward in training mode (SyncBatchNormalization on batch statistics, Dropout), loss, backward, the exchanges of the step,
// SGD(momentum) + l2, Keras moving statistics, the Dropout-seed / epoch counters — captured once into ONE CUDA graph and replayed.
// No PyTorch, no NCCL: device memory is cudaMalloc'ed here, the exchanges are the peer-memory collectives of p2p_exchange.cuh.
//
// Reference: train.py:143-169 (model.compile + fit under tf.distribute.MirroredStrategy), graph deeplabv3p/models/layers.py:74-219 +
// model.py:75-86, loss deeplabv3p/loss.py:60-192, optimizer common/model_utils.py:122-123, regulariser layers.py:12-21.
// Included at the end of dlv3p_api.cu after train_api.cuh (uses its operator entry points).
#pragma once

#include <functional>

struct TrOff { size_t off; int d0, d1; };

struct dlv3p_trainer {
  dlv3p_trainer_config cfg{};
  int device = 0, sms = 148;
  std::string err;
  int B = 0, H = 0, W = 0, OS = 16, Cin = 0, Cs = 0, NC = 0, NCp = 0, Bp = 0, h = 0, w = 0, hs = 0, ws = 0, M1 = 0, M2 = 0;
  int rates[3] = {6, 12, 18};
  int world = 1, rank = 0;
  // layout (the flat fp32 buffers: [A: 1x1 kernels [K,N] + classifier bias][B: depthwise taps [9,C]][C: per BN layer beta | gamma])
  std::map<std::string, TrOff> off;
  std::map<std::string, std::pair<size_t, int>> stat_off;
  std::vector<std::tuple<std::string, int, int>> conv_specs;
  std::vector<std::pair<std::string, int>> dw_specs, bn_specs;
  std::vector<std::vector<std::string>> fwd_groups, bwd_groups;
  size_t endA = 0, endB = 0, nparams = 0, nstats = 0, nbn = 0;
  // device state
  std::vector<void*> allocs;
  float *params = nullptr, *grads = nullptr, *velocity = nullptr, *stats = nullptr, *moving_mean = nullptr, *moving_var = nullptr;
  __nv_bfloat16 *w_kn = nullptr, *w_nk = nullptr;
  uint32_t* seed_d = nullptr;
  int* mov_ix = nullptr;
  std::map<std::string, void*> T;
  size_t partial_floats = 0;
  // exchange
  dlv3p_p2p* comm = nullptr;
  size_t pay_stats = 0, pay_bn = 0, pay_grads = 0, pay_red = 0, pay_loss = 0;   // float offsets inside the payload area
  float *stats_send = nullptr, *bn_send = nullptr, *loss_send = nullptr;
  // execution
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_in = nullptr, ev_out = nullptr;
  cudaGraphExec_t gexec = nullptr;
  cudaGraph_t graph = nullptr;
  void *feat_s = nullptr, *skip_s = nullptr, *labels_s = nullptr;
  size_t feat_bytes = 0, skip_bytes = 0, labels_bytes = 0;
  long long step_count = 0, launches = 0, launches_per_step = 0;
  int rc = 0;
  cudaStream_t cur = nullptr;       // stream the step body is being enqueued on
};

namespace {

inline const P2pPeers* dlv3p_p2p_peers(dlv3p_p2p* c) { return &c->peers; }
inline const uint32_t* dlv3p_p2p_epoch(dlv3p_p2p* c) { return c->epoch; }

int tfail(dlv3p_trainer* t, int code, const std::string& msg) {
  if (t) t->err = msg;
  g_tls_error = msg;
  return code;
}
inline size_t rup(size_t v, size_t a) { return (v + a - 1) / a * a; }

void tr_layout(dlv3p_trainer* t) {
  const int Cin = t->Cin, Cs = t->Cs, NCp = t->NCp;
  if (t->cfg.lite) {
    // ASPP_Lite_block (layers.py:166-196) + the tail straight on its output: the *_lite models (deeplabv3p_mobilenetv2.py:326-331)
    t->conv_specs = {{"image_pooling", Cin, 256}, {"aspp0", Cin, 256}, {"concat_projection", 512, 256}, {"conv_upsample", 256, NCp}};
    t->dw_specs = {};
    t->bn_specs = {{"image_pooling_BN", 256}, {"aspp0_BN", 256}, {"concat_projection_BN", 256}};
    t->fwd_groups = {{"image_pooling_BN", "aspp0_BN"}, {"concat_projection_BN"}};
    t->bwd_groups = {{"concat_projection_BN"}, {"aspp0_BN", "image_pooling_BN"}};
  } else {
  t->conv_specs = {{"image_pooling", Cin, 256}, {"aspp0", Cin, 256}, {"aspp1_pointwise", Cin, 256}, {"aspp2_pointwise", Cin, 256},
                   {"aspp3_pointwise", Cin, 256}, {"concat_projection", 1280, 256}, {"feature_projection0", Cs, 48},
                   {"decoder_conv0_pointwise", 304, 256}, {"decoder_conv1_pointwise", 256, 256}, {"conv_upsample", 256, NCp}};
  t->dw_specs = {{"aspp1_depthwise", Cin}, {"aspp2_depthwise", Cin}, {"aspp3_depthwise", Cin}, {"decoder_conv0_depthwise", 304}, {"decoder_conv1_depthwise", 256}};
  t->bn_specs = {{"image_pooling_BN", 256}, {"aspp0_BN", 256}, {"aspp1_depthwise_BN", Cin}, {"aspp1_pointwise_BN", 256}, {"aspp2_depthwise_BN", Cin},
                 {"aspp2_pointwise_BN", 256}, {"aspp3_depthwise_BN", Cin}, {"aspp3_pointwise_BN", 256}, {"concat_projection_BN", 256},
                 {"feature_projection0_BN", 48}, {"decoder_conv0_depthwise_BN", 304}, {"decoder_conv0_pointwise_BN", 256},
                 {"decoder_conv1_depthwise_BN", 256}, {"decoder_conv1_pointwise_BN", 256}};
  // BN layers grouped by data dependence: one exchange per group (7 + 7 collectives per step instead of 28)
  t->fwd_groups = {{"image_pooling_BN", "aspp0_BN", "aspp1_depthwise_BN", "aspp2_depthwise_BN", "aspp3_depthwise_BN", "feature_projection0_BN"},
                   {"aspp1_pointwise_BN", "aspp2_pointwise_BN", "aspp3_pointwise_BN"}, {"concat_projection_BN"}, {"decoder_conv0_depthwise_BN"},
                   {"decoder_conv0_pointwise_BN"}, {"decoder_conv1_depthwise_BN"}, {"decoder_conv1_pointwise_BN"}};
  t->bwd_groups = {{"decoder_conv1_pointwise_BN"}, {"decoder_conv1_depthwise_BN"}, {"decoder_conv0_pointwise_BN"}, {"decoder_conv0_depthwise_BN"},
                   {"feature_projection0_BN", "concat_projection_BN"},
                   {"aspp0_BN", "aspp1_pointwise_BN", "aspp2_pointwise_BN", "aspp3_pointwise_BN", "image_pooling_BN"},
                   {"aspp1_depthwise_BN", "aspp2_depthwise_BN", "aspp3_depthwise_BN"}};
  }
  size_t o = 0;
  for (auto& [name, K, N] : t->conv_specs) {
    t->off[name + "/kernel"] = {o, K, N};
    o = rup(o + static_cast<size_t>(K) * N, 8);
  }
  t->off["conv_upsample/bias"] = {o, NCp, 0};
  o = rup(o + NCp, 8);
  t->endA = o;
  for (auto& [name, C] : t->dw_specs) {
    t->off[name + "/depthwise_kernel"] = {o, 9, C};
    o = rup(o + static_cast<size_t>(9) * C, 8);
  }
  t->endB = o;
  std::map<std::string, int> chan(t->bn_specs.begin(), t->bn_specs.end());
  for (auto& grp : t->bwd_groups)
    for (auto& name : grp) {
      const int C = chan[name];
      t->off[name + "/beta"] = {o, C, 0};
      t->off[name + "/gamma"] = {o + C, C, 0};
      o = rup(o + 2 * static_cast<size_t>(C), 8);
    }
  t->nparams = o;
  size_t so = 0;
  for (auto& grp : t->fwd_groups)
    for (auto& name : grp) {
      t->stat_off[name] = {so, chan[name]};
      so = rup(so + 2 * static_cast<size_t>(chan[name]) + 1, 4);
    }
  t->nstats = so;
  t->nbn = 0;
  for (auto& [n_, C] : t->bn_specs) t->nbn += C;
}

template <class Tp>
int tr_alloc(dlv3p_trainer* t, Tp** p, size_t count, int fill_byte = 0) {
  void* q = nullptr;
  const size_t bytes = rup(count * sizeof(Tp) + 256, 256);
  if (cudaMalloc(&q, bytes) != cudaSuccess) return tfail(t, DLV3P_ERR_NOMEM, fmt("trainer: cudaMalloc(%zu) failed", bytes));
  cudaMemset(q, fill_byte, bytes);
  t->allocs.push_back(q);
  *p = static_cast<Tp*>(q);
  return 0;
}
int tr_buf(dlv3p_trainer* t, const char* name, size_t elems, size_t elem_size) {
  char* p = nullptr;
  int r = tr_alloc(t, &p, elems * elem_size);
  if (r) return r;
  t->T[name] = p;
  return 0;
}

// ---- the step body: every call below is the C-ABI operator the Python trainer used to call one by one
struct Step {
  dlv3p_trainer* t;
  cudaStream_t st;
  int rc = 0;
  void call(int r, int kernels = 1) {
    if (rc) return;
    if (r) { rc = tfail(t, r, g_tls_error); return; }
    t->launches += kernels;
  }
  __nv_bfloat16* bf(const char* n) { return static_cast<__nv_bfloat16*>(t->T.at(n)); }
  float* f32(const char* n) { return static_cast<float*>(t->T.at(n)); }
  float* P(const std::string& key) { return t->params + t->off.at(key).off; }
  float* G(const std::string& key) { return t->grads + t->off.at(key).off; }
  __nv_bfloat16* Wnk(const std::string& name) { return t->w_nk + t->off.at(name + "/kernel").off; }
  __nv_bfloat16* Wkn(const std::string& name) { return t->w_kn + t->off.at(name + "/kernel").off; }

  void gemm(const void* a, int64_t lda, const void* b, int64_t ldb, int64_t M, int N, int64_t K, void* d, int64_t ldd, int out_fp32 = 0, int splits = 1) {
    call(dlv3p_train_gemm_nt(t->device, a, lda, b, ldb, M, N, K, d, ldd, out_fp32, splits, t->T.at("partial"), st), splits > 1 ? 2 : 1);
  }
  void conv_fwd(const std::string& name, const void* x, int64_t ldx, int64_t M, void* out, int64_t ldo, int out_fp32 = 0) {
    const TrOff& o = t->off.at(name + "/kernel");
    gemm(x, ldx, Wnk(name), o.d0, M, o.d1, o.d0, out, ldo, out_fp32);
  }
  void conv_dgrad(const std::string& name, const void* dy, int64_t ld_dy, int64_t M, void* dx, int64_t ldx) {
    const TrOff& o = t->off.at(name + "/kernel");
    gemm(dy, ld_dy, Wkn(name), o.d1, M, o.d0, o.d1, dx, ldx);
  }
  // dW[K,N] = X[M,K]^T dY[M,N]: the MN-major tcgen05 GEMM straight from the [pixels, channels] tensors; the contraction over
  // pixels is split so that ~one wave of CTAs is busy (fp32 partials, fixed-order reduce)
  void conv_wgrad(const std::string& name, const void* x, int64_t ldx, const void* dy, int64_t ld_dy, int64_t M) {
    const TrOff& o = t->off.at(name + "/kernel");
    const int K = o.d0, N = o.d1;
    const int tiles = ceil_div(K, 128) * ceil_div(N, N > 64 ? 256 : 64);
    const long long kblocks = (M + 63) / 64;
    long long splits = std::min<long long>(kblocks, std::min<long long>(t->sms / tiles, static_cast<long long>(t->partial_floats / std::max<size_t>(1, static_cast<size_t>(K) * N))));
    if (splits < 1) splits = 1;
    call(dlv3p_train_gemm_tn(t->device, x, ldx, dy, ld_dy, K, N, M, G(name + "/kernel"), N, 1, static_cast<int>(splits), t->T.at("partial"), st), splits > 1 ? 2 : 1);
  }
  size_t span_end(const std::vector<std::string>& grp) { auto& so = t->stat_off.at(grp.back()); return so.first + 2 * static_cast<size_t>(so.second) + 1; }
  void bn_stats(const std::string& name, const void* x, int64_t M) {
    auto& so = t->stat_off.at(name);
    float* dst = (t->world > 1 ? t->stats_send : t->stats) + so.first;      // several replicas: partial sums go to the payload area
    call(dlv3p_op_bn_stats(t->device, x, M, so.second, dst, t->T.at("bn_scratch"), st), 2);
  }
  void sync_stats(int gi) {      // ONE exchange of the contiguous [sum x | sum x^2 | n] vectors of a group of independent BN layers
    if (t->world == 1) return;
    const auto& grp = t->fwd_groups[gi];
    const size_t b = t->stat_off.at(grp.front()).first, e = span_end(grp);
    call(dlv3p_p2p_allreduce(t->comm, gi, t->pay_stats + b, static_cast<int>(rup(e - b, 4)), t->stats + b, st));
  }
  int fwd_group_of(const std::string& name) {
    for (size_t i = 0; i < t->fwd_groups.size(); ++i)
      for (auto& n : t->fwd_groups[i]) if (n == name) return static_cast<int>(i);
    return -1;
  }
  int bwd_group_of(const std::string& name) {
    for (size_t i = 0; i < t->bwd_groups.size(); ++i)
      for (auto& n : t->bwd_groups[i]) if (n == name) return static_cast<int>(i);
    return -1;
  }
  void bn_apply(const std::string& name, const void* x, int64_t M, void* y, int64_t ldy, int relu = 1) {
    auto& so = t->stat_off.at(name);
    call(dlv3p_train_bn_apply(t->device, x, M, so.second, t->stats + so.first, P(name + "/gamma"), P(name + "/beta"), t->cfg.eps, relu, y, ldy, st));
  }
  void bn_fwd(const std::string& name, const void* x, int64_t M, void* y, int64_t ldy, int relu = 1) {
    bn_stats(name, x, M);
    sync_stats(fwd_group_of(name));
    bn_apply(name, x, M, y, ldy, relu);
  }
  void bn_bwd_stats(const std::string& name, const void* dy, int64_t ld_dy, const void* y, int64_t ld_y, const void* x, int64_t M, int relu = 1) {
    auto& so = t->stat_off.at(name);
    const size_t go = t->off.at(name + "/beta").off;         // grads[go : go+2C] = d(beta) | d(gamma) = sum g | sum g*xhat
    float* dst = t->world > 1 ? t->bn_send + (go - t->endB) : t->grads + go;
    call(dlv3p_train_bn_bwd_stats(t->device, dy, ld_dy, y, ld_y, x, M, so.second, t->stats + so.first, t->cfg.eps, relu, dst, t->T.at("scratch"), st), 2);
  }
  void sync_bn_grads(int gi) {
    if (t->world == 1) return;
    const auto& grp = t->bwd_groups[gi];
    const TrOff& o0 = t->off.at(grp.front() + "/beta");
    const TrOff& o1 = t->off.at(grp.back() + "/beta");
    const size_t b = o0.off, e = o1.off + 2 * static_cast<size_t>(o1.d0);
    call(dlv3p_p2p_allreduce(t->comm, static_cast<int>(t->fwd_groups.size()) + gi, t->pay_bn + (b - t->endB), static_cast<int>(rup(e - b, 4)), t->grads + b, st));
  }
  void bn_bwd_apply(const std::string& name, const void* dy, int64_t ld_dy, const void* y, int64_t ld_y, const void* x, int64_t M, void* dx, int relu = 1) {
    auto& so = t->stat_off.at(name);
    const size_t go = t->off.at(name + "/beta").off;
    call(dlv3p_train_bn_bwd_apply(t->device, dy, ld_dy, y, ld_y, x, M, so.second, t->stats + so.first, t->grads + go, P(name + "/gamma"), t->cfg.eps, relu, dx, st));
  }
  void bn_bwd(const std::string& name, const void* dy, int64_t ld_dy, const void* y, int64_t ld_y, const void* x, int64_t M, void* dx, int relu = 1) {
    bn_bwd_stats(name, dy, ld_dy, y, ld_y, x, M, relu);
    sync_bn_grads(bwd_group_of(name));
    bn_bwd_apply(name, dy, ld_dy, y, ld_y, x, M, dx, relu);
  }
  // SepConv_BN (depth_activation=True, layers.py:98-109) in training mode; keeps d (raw depthwise), a (BN+ReLU), p (raw pointwise)
  void sep_fwd(const std::string& prefix, const void* x, int Bn, int Hh, int Ww, int C, int rate, void* d, void* a, void* p, void* y, int64_t ldy) {
    const int64_t M = static_cast<int64_t>(Bn) * Hh * Ww;
    call(dlv3p_train_depthwise(t->device, x, Bn, Hh, Ww, C, rate, P(prefix + "_depthwise/depthwise_kernel"), 0, d, st));
    bn_fwd(prefix + "_depthwise_BN", d, M, a, C);
    conv_fwd(prefix + "_pointwise", a, C, M, p, 256);
    bn_fwd(prefix + "_pointwise_BN", p, M, y, ldy);
  }
  void sep_bwd(const std::string& prefix, const void* x, int Bn, int Hh, int Ww, int C, int rate, const void* d, const void* a, const void* p, const void* y,
               int64_t ld_y, const void* dy, int64_t ld_dy, void* g_pw, void* g_a, void* g_d, void* dx) {
    const int64_t M = static_cast<int64_t>(Bn) * Hh * Ww;
    bn_bwd(prefix + "_pointwise_BN", dy, ld_dy, y, ld_y, p, M, g_pw);
    conv_wgrad(prefix + "_pointwise", a, C, g_pw, 256, M);
    conv_dgrad(prefix + "_pointwise", g_pw, 256, M, g_a, C);
    bn_bwd(prefix + "_depthwise_BN", g_a, C, a, C, d, M, g_d);
    call(dlv3p_train_depthwise_wgrad(t->device, x, g_d, Bn, Hh, Ww, C, rate, G(prefix + "_depthwise/depthwise_kernel"), t->T.at("scratch"), st), 2);
    call(dlv3p_train_depthwise(t->device, g_d, Bn, Hh, Ww, C, rate, P(prefix + "_depthwise/depthwise_kernel"), 1, dx, st));
  }

  void refresh_bf16() {      // bf16 operand copies of the 1x1 kernels after an update: [K,N] (dgrad) and its transpose [N,K] (forward)
    call(dlv3p_train_cast_bf16(t->device, t->params, t->w_kn, static_cast<int64_t>(t->endA), st));
    for (auto& [name, K, N] : t->conv_specs) {
      const size_t o = t->off.at(name + "/kernel").off;
      call(dlv3p_train_transpose(t->device, t->w_kn + o, K, N, N, t->w_nk + o, K, st));
    }
  }

  void forward_backward(const void* feat, const void* skip, const void* labels) {
    const int B = t->B, Bp = t->Bp, M1 = t->M1, M2 = t->M2, Cin = t->Cin, Cs = t->Cs, NC = t->NC, NCp = t->NCp, h = t->h, w = t->w, hs = t->hs, ws = t->ws;
    const int npix1 = h * w;
    const int dev = t->device;
    auto X = [&](const char* n, size_t elem_off = 0) { return static_cast<void*>(bf(n) + elem_off); };
    char nm[64];
    auto N3 = [&](const char* f, int i) { snprintf(nm, sizeof(nm), f, i); return std::string(nm); };
    if (t->cfg.lite) {
      // ================ ASPP_Lite_block (layers.py:166-196) -> Dropout -> classifier -> pred_resize -> loss; no decoder: hs = h, ws = w
      call(dlv3p_train_rows_reduce(dev, feat, Cin, B, npix1, Cin, 1.0f / npix1, X("pool"), 0, st));
      conv_fwd("image_pooling", X("pool"), Cin, Bp, X("r4"), 256);
      bn_stats("image_pooling_BN", X("r4"), B);
      conv_fwd("aspp0", feat, Cin, M1, X("r0"), 256);
      bn_stats("aspp0_BN", X("r0"), M1);
      sync_stats(0);
      bn_apply("image_pooling_BN", X("r4"), B, X("b4"), 256);
      call(dlv3p_train_bcast_rows(dev, X("b4"), B, npix1, 256, 1.0f, X("concat"), 512, 0, st));      // aspp_resize of a 1 x 1 map
      bn_apply("aspp0_BN", X("r0"), M1, X("concat", 256), 512);
      conv_fwd("concat_projection", X("concat"), 512, M1, X("rp"), 256);
      bn_fwd("concat_projection_BN", X("rp"), M1, X("yproj"), 256);
      const char* aspp_out = "yproj";
      if (t->cfg.dropout > 0) {
        call(dlv3p_train_dropout(dev, X("yproj"), X("aspp_out"), static_cast<int64_t>(M1) * 256, 0, t->seed_d, t->cfg.dropout, st));
        aspp_out = "aspp_out";
      }
      conv_fwd("conv_upsample", X(aspp_out), 256, M1, f32("logits"), NCp, 1);
      const float inv_norm = 1.0f / (static_cast<float>(t->cfg.global_batch) * t->H * t->W);
      call(dlv3p_train_softmax_loss(dev, f32("logits"), NCp, P("conv_upsample/bias"), static_cast<const uint8_t*>(labels), B, NC, h, w, t->H, t->W,
                                    t->cfg.ignore_index, inv_norm, t->cfg.loss_kind, f32("class_w"), t->cfg.focal_gamma, t->cfg.focal_alpha, f32("dfull"),
                                    t->world > 1 ? t->loss_send : f32("loss"), t->T.at("loss_scratch"), st), 2);
      // ================ backward
      call(dlv3p_train_resize_bwd_planar(dev, f32("dfull"), B, NC, h, w, t->H, t->W, X("dlow"), NCp, t->T.at("adj_tmp"), st), 2);
      call(dlv3p_op_bn_stats(dev, X("dlow"), M1, NCp, f32("bias_stats"), t->T.at("bn_scratch"), st), 2);
      if (!rc && cudaMemcpyAsync(G("conv_upsample/bias"), f32("bias_stats"), NCp * sizeof(float), cudaMemcpyDeviceToDevice, st) != cudaSuccess)
        rc = tfail(t, DLV3P_ERR_CUDA, "trainer: bias gradient copy failed");
      conv_wgrad("conv_upsample", X(aspp_out), 256, X("dlow"), NCp, M1);
      conv_dgrad("conv_upsample", X("dlow"), NCp, M1, X("da_out"), 256);
      if (t->cfg.dropout > 0) call(dlv3p_train_dropout(dev, X("da_out"), X("da_out"), static_cast<int64_t>(M1) * 256, 0, t->seed_d, t->cfg.dropout, st));
      bn_bwd("concat_projection_BN", X("da_out"), 256, X("yproj"), 256, X("rp"), M1, X("drp"));
      conv_wgrad("concat_projection", X("concat"), 512, X("drp"), 256, M1);
      conv_dgrad("concat_projection", X("drp"), 256, M1, X("dconcat"), 512);
      call(dlv3p_train_rows_reduce(dev, X("dconcat"), 512, B, npix1, 256, 1.0f, X("db4"), 0, st));
      bn_bwd_stats("aspp0_BN", X("dconcat", 256), 512, X("concat", 256), 512, X("r0"), M1);
      bn_bwd_stats("image_pooling_BN", X("db4"), 256, X("b4"), 256, X("r4"), B);
      sync_bn_grads(1);
      bn_bwd_apply("aspp0_BN", X("dconcat", 256), 512, X("concat", 256), 512, X("r0"), M1, X("g1_256"));
      bn_bwd_apply("image_pooling_BN", X("db4"), 256, X("b4"), 256, X("r4"), B, X("dr4"));
      conv_wgrad("aspp0", feat, Cin, X("g1_256"), 256, M1);
      conv_dgrad("aspp0", X("g1_256"), 256, M1, X("dfeat"), Cin);
      conv_wgrad("image_pooling", X("pool"), Cin, X("dr4"), 256, Bp);
      conv_dgrad("image_pooling", X("dr4"), 256, Bp, X("dpool"), Cin);
      call(dlv3p_train_bcast_rows(dev, X("dpool"), B, npix1, Cin, 1.0f / npix1, X("dfeat"), Cin, 1, st));
      (void)skip; (void)M2; (void)Cs; (void)hs; (void)ws;
      return;
    }
    // ---------------- ASPP_block (layers.py:114-163) + the decoder's skip projection (:209-213), phase by phase
    call(dlv3p_train_rows_reduce(dev, feat, Cin, B, npix1, Cin, 1.0f / npix1, X("pool"), 0, st));
    conv_fwd("image_pooling", X("pool"), Cin, Bp, X("r4"), 256);
    bn_stats("image_pooling_BN", X("r4"), B);
    conv_fwd("aspp0", feat, Cin, M1, X("r0"), 256);
    bn_stats("aspp0_BN", X("r0"), M1);
    for (int i = 1; i <= 3; ++i) {
      call(dlv3p_train_depthwise(dev, feat, B, h, w, Cin, t->rates[i - 1], P(N3("aspp%d_depthwise/depthwise_kernel", i)), 0, X(N3("d%d", i).c_str()), st));
      bn_stats(N3("aspp%d_depthwise_BN", i), X(N3("d%d", i).c_str()), M1);
    }
    conv_fwd("feature_projection0", skip, Cs, M2, X("rs"), 48);
    bn_stats("feature_projection0_BN", X("rs"), M2);
    sync_stats(0);
    bn_apply("image_pooling_BN", X("r4"), B, X("b4"), 256);
    call(dlv3p_train_bcast_rows(dev, X("b4"), B, npix1, 256, 1.0f, X("concat"), 1280, 0, st));
    bn_apply("aspp0_BN", X("r0"), M1, X("concat", 256), 1280);
    for (int i = 1; i <= 3; ++i) bn_apply(N3("aspp%d_depthwise_BN", i), X(N3("d%d", i).c_str()), M1, X(N3("a%d", i).c_str()), Cin);
    bn_apply("feature_projection0_BN", X("rs"), M2, X("dcat", 256), 304);
    for (int i = 1; i <= 3; ++i) {
      conv_fwd(N3("aspp%d_pointwise", i), X(N3("a%d", i).c_str()), Cin, M1, X(N3("p%d", i).c_str()), 256);
      bn_stats(N3("aspp%d_pointwise_BN", i), X(N3("p%d", i).c_str()), M1);
    }
    sync_stats(1);
    for (int i = 1; i <= 3; ++i) bn_apply(N3("aspp%d_pointwise_BN", i), X(N3("p%d", i).c_str()), M1, X("concat", 256 * (i + 1)), 1280);
    conv_fwd("concat_projection", X("concat"), 1280, M1, X("rp"), 256);
    bn_fwd("concat_projection_BN", X("rp"), M1, X("yproj"), 256);
    const char* aspp_out = "yproj";
    if (t->cfg.dropout > 0) {
      call(dlv3p_train_dropout(dev, X("yproj"), X("aspp_out"), static_cast<int64_t>(M1) * 256, 0, t->seed_d, t->cfg.dropout, st));
      aspp_out = "aspp_out";
    }
    // ---------------- Decoder_block (layers.py:199-219)
    call(dlv3p_train_resize(dev, X(aspp_out), B, h, w, 256, hs, ws, X("dcat"), 304, st));
    sep_fwd("decoder_conv0", X("dcat"), B, hs, ws, 304, 1, X("c0d"), X("c0a"), X("c0p"), X("y0"), 256);
    sep_fwd("decoder_conv1", X("y0"), B, hs, ws, 256, 1, X("c1d"), X("c1a"), X("c1p"), X("y1"), 256);
    // ---------------- tail + loss (model.py:75-86, loss.py:121-156)
    conv_fwd("conv_upsample", X("y1"), 256, M2, f32("logits"), NCp, 1);
    const float inv_norm = 1.0f / (static_cast<float>(t->cfg.global_batch) * t->H * t->W);
    call(dlv3p_train_softmax_loss(dev, f32("logits"), NCp, P("conv_upsample/bias"), static_cast<const uint8_t*>(labels), B, NC, hs, ws, t->H, t->W,
                                  t->cfg.ignore_index, inv_norm, t->cfg.loss_kind, f32("class_w"), t->cfg.focal_gamma, t->cfg.focal_alpha, f32("dfull"),
                                  t->world > 1 ? t->loss_send : f32("loss"), t->T.at("loss_scratch"), st), 2);
    // ================ backward
    call(dlv3p_train_resize_bwd_planar(dev, f32("dfull"), B, NC, hs, ws, t->H, t->W, X("dlow"), NCp, t->T.at("adj_tmp"), st), 2);
    // d(bias) = column sums of d(logits): the banded two-stage statistics kernel, first NCp entries
    call(dlv3p_op_bn_stats(dev, X("dlow"), M2, NCp, f32("bias_stats"), t->T.at("bn_scratch"), st), 2);
    if (!rc && cudaMemcpyAsync(G("conv_upsample/bias"), f32("bias_stats"), NCp * sizeof(float), cudaMemcpyDeviceToDevice, st) != cudaSuccess)
      rc = tfail(t, DLV3P_ERR_CUDA, "trainer: bias gradient copy failed");
    conv_wgrad("conv_upsample", X("y1"), 256, X("dlow"), NCp, M2);
    conv_dgrad("conv_upsample", X("dlow"), NCp, M2, X("g256a"), 256);
    sep_bwd("decoder_conv1", X("y0"), B, hs, ws, 256, 1, X("c1d"), X("c1a"), X("c1p"), X("y1"), 256, X("g256a"), 256, X("g256b"), X("g256c"), X("g256b"), X("g256a"));
    sep_bwd("decoder_conv0", X("dcat"), B, hs, ws, 304, 1, X("c0d"), X("c0a"), X("c0p"), X("y0"), 256, X("g256a"), 256, X("g256b"), X("g304a"), X("g304b"), X("g304a"));
    call(dlv3p_train_resize_bwd(dev, X("g304a"), 304, B, h, w, 256, hs, ws, X("da_out"), st));
    if (t->cfg.dropout > 0) call(dlv3p_train_dropout(dev, X("da_out"), X("da_out"), static_cast<int64_t>(M1) * 256, 0, t->seed_d, t->cfg.dropout, st));
    bn_bwd_stats("feature_projection0_BN", X("g304a", 256), 304, X("dcat", 256), 304, X("rs"), M2);
    bn_bwd_stats("concat_projection_BN", X("da_out"), 256, X("yproj"), 256, X("rp"), M1);
    sync_bn_grads(4);
    bn_bwd_apply("feature_projection0_BN", X("g304a", 256), 304, X("dcat", 256), 304, X("rs"), M2, X("drs"));
    bn_bwd_apply("concat_projection_BN", X("da_out"), 256, X("yproj"), 256, X("rp"), M1, X("drp"));
    conv_wgrad("feature_projection0", skip, Cs, X("drs"), 48, M2);
    conv_dgrad("feature_projection0", X("drs"), 48, M2, X("dskip"), Cs);
    conv_wgrad("concat_projection", X("concat"), 1280, X("drp"), 256, M1);
    conv_dgrad("concat_projection", X("drp"), 256, M1, X("dconcat"), 1280);
    call(dlv3p_train_rows_reduce(dev, X("dconcat"), 1280, B, npix1, 256, 1.0f, X("db4"), 0, st));
    const char* names[4] = {"aspp0_BN", "aspp1_pointwise_BN", "aspp2_pointwise_BN", "aspp3_pointwise_BN"};
    const char* raws[4] = {"r0", "p1", "p2", "p3"};
    const char* gout[4] = {"g1_256", "gp1", "gp2", "gp3"};
    for (int k = 0; k < 4; ++k) bn_bwd_stats(names[k], X("dconcat", 256 * (k + 1)), 1280, X("concat", 256 * (k + 1)), 1280, X(raws[k]), M1);
    bn_bwd_stats("image_pooling_BN", X("db4"), 256, X("b4"), 256, X("r4"), B);
    sync_bn_grads(5);
    for (int k = 0; k < 4; ++k) bn_bwd_apply(names[k], X("dconcat", 256 * (k + 1)), 1280, X("concat", 256 * (k + 1)), 1280, X(raws[k]), M1, X(gout[k]));
    bn_bwd_apply("image_pooling_BN", X("db4"), 256, X("b4"), 256, X("r4"), B, X("dr4"));
    conv_wgrad("aspp0", feat, Cin, X("g1_256"), 256, M1);
    conv_dgrad("aspp0", X("g1_256"), 256, M1, X("dfeat"), Cin);
    conv_wgrad("image_pooling", X("pool"), Cin, X("dr4"), 256, Bp);
    conv_dgrad("image_pooling", X("dr4"), 256, Bp, X("dpool"), Cin);
    call(dlv3p_train_bcast_rows(dev, X("dpool"), B, npix1, Cin, 1.0f / npix1, X("dfeat"), Cin, 1, st));
    for (int i = 1; i <= 3; ++i) {
      conv_wgrad(N3("aspp%d_pointwise", i), X(N3("a%d", i).c_str()), Cin, X(N3("gp%d", i).c_str()), 256, M1);
      conv_dgrad(N3("aspp%d_pointwise", i), X(N3("gp%d", i).c_str()), 256, M1, X(N3("ga%d", i).c_str()), Cin);
      bn_bwd_stats(N3("aspp%d_depthwise_BN", i), X(N3("ga%d", i).c_str()), Cin, X(N3("a%d", i).c_str()), Cin, X(N3("d%d", i).c_str()), M1);
    }
    sync_bn_grads(6);
    for (int i = 1; i <= 3; ++i) {
      const std::string name = N3("aspp%d_depthwise", i);
      bn_bwd_apply(name + "_BN", X(N3("ga%d", i).c_str()), Cin, X(N3("a%d", i).c_str()), Cin, X(N3("d%d", i).c_str()), M1, X("gB"));
      call(dlv3p_train_depthwise_wgrad(dev, feat, X("gB"), B, h, w, Cin, t->rates[i - 1], G(name + "/depthwise_kernel"), t->T.at("scratch"), st), 2);
      call(dlv3p_train_depthwise(dev, X("gB"), B, h, w, Cin, t->rates[i - 1], P(name + "/depthwise_kernel"), 1, X("dfeat_tmp"), st));
      call(dlv3p_train_add(dev, X("dfeat"), X("dfeat_tmp"), X("dfeat"), static_cast<int64_t>(M1) * Cin, st));
    }
  }

  // ONE exchange (two-shot all-reduce over peer memory) of the flat fp32 bucket holding every 1x1 kernel, the classifier bias and every
  // depthwise kernel (regions A|B; the loss is normalised by the GLOBAL batch, so the sum is the gradient of the global mean loss —
  // MirroredStrategy semantics, train.py:143-158).  BN gradients are already global.  The loss shares travel with it.
  void all_reduce_gradients() {
    if (t->world == 1 || rc) return;
    const int s0 = static_cast<int>(t->fwd_groups.size() + t->bwd_groups.size());
    const P2pPeers& Pp = *dlv3p_p2p_peers(t->comm);
    const size_t n = rup(t->endB, 4);
    p2p_reduce_scatter_kernel<<<128, 512, 0, st>>>(Pp, s0, dlv3p_p2p_epoch(t->comm), t->pay_grads, t->pay_red, n);
    p2p_all_gather_kernel<<<128, 512, 0, st>>>(Pp, s0 + 1, dlv3p_p2p_epoch(t->comm), t->pay_red, n, t->grads);
    if (cudaGetLastError() != cudaSuccess) rc = tfail(t, DLV3P_ERR_CUDA, "trainer: gradient exchange launch failed");
    t->launches += 2;
    call(dlv3p_p2p_allreduce(t->comm, s0 + 2, t->pay_loss, 4, f32("loss"), st));
  }

  void apply_gradients() {
    const float lr = t->cfg.lr, mom = t->cfg.momentum, l2 = t->cfg.l2;
    call(dlv3p_train_sgd(t->device, t->params, t->grads, t->velocity, static_cast<int64_t>(t->endA), lr, mom, l2, 1.0f, st));
    call(dlv3p_train_sgd(t->device, t->params + t->endA, t->grads + t->endA, t->velocity + t->endA, static_cast<int64_t>(t->nparams - t->endA), lr, mom, 0.0f, 1.0f, st));
    refresh_bf16();
    if (!rc) {
      moving_stats_kernel<<<ceil_div(static_cast<int>(t->nbn), 256), 256, 0, st>>>(t->stats, t->mov_ix, static_cast<int>(t->nbn), t->cfg.bn_momentum, t->moving_mean, t->moving_var);
      add_u32_kernel<<<1, 1, 0, st>>>(t->seed_d, 0x85EBCA6Bu);      // dropout_seed is linear in the step
      t->launches += 2;
      if (t->world > 1) call(dlv3p_p2p_advance(t->comm, st));
      if (cudaGetLastError() != cudaSuccess) rc = tfail(t, DLV3P_ERR_CUDA, "trainer: update launch failed");
    }
  }
};

uint32_t tr_dropout_seed(uint32_t base, uint32_t step, uint32_t rank) { return base * 0x9E3779B1u + step * 0x85EBCA6Bu + rank * 0xC2B2AE35u + 0x27D4EB2Fu; }

}  // namespace

extern "C" {

const char* dlv3p_trainer_last_error(const dlv3p_trainer* t) { return t ? t->err.c_str() : g_tls_error.c_str(); }

void dlv3p_trainer_destroy(dlv3p_trainer* t) {
  if (!t) return;
  cudaSetDevice(t->device);
  cudaDeviceSynchronize();
  if (t->gexec) cudaGraphExecDestroy(t->gexec);
  if (t->graph) cudaGraphDestroy(t->graph);
  if (t->comm) dlv3p_p2p_destroy(t->comm);
  for (void* p : t->allocs) cudaFree(p);
  if (t->ev_in) cudaEventDestroy(t->ev_in);
  if (t->ev_out) cudaEventDestroy(t->ev_out);
  if (t->stream) cudaStreamDestroy(t->stream);
  delete t;
}

int dlv3p_trainer_create(const dlv3p_trainer_config* cfg, int device, dlv3p_trainer** out, uint8_t ipc_handle_out[64]) {
  if (!cfg || !out || !ipc_handle_out) return tfail(nullptr, DLV3P_ERR_INVALID, "null argument");
  *out = nullptr;
  int sms = 0, r = train_prolog(device, &sms);
  if (r) return r;
  const dlv3p_trainer_config& g = *cfg;
  if (g.B < 1 || g.H < 1 || g.W < 1 || g.NC < 1 || g.NC > 256) return tfail(nullptr, DLV3P_ERR_INVALID, "trainer: bad B / H / W / NC");
  if (g.OS != 8 && g.OS != 16 && g.OS != 32) return tfail(nullptr, DLV3P_ERR_INVALID, fmt("invalid output stride %d", g.OS));
  if (g.world < 1 || g.world > kP2pMaxWorld || g.rank < 0 || g.rank >= g.world) return tfail(nullptr, DLV3P_ERR_INVALID, "trainer: bad world / rank (world <= 16)");
  if (g.loss_kind < 0 || g.loss_kind > 2) return tfail(nullptr, DLV3P_ERR_INVALID, "trainer: loss_kind 0 crossentropy, 1 class-weighted, 2 focal");
  dlv3p_trainer* t = new dlv3p_trainer();
  auto bail = [&](int code) { std::string m = t->err; dlv3p_trainer_destroy(t); return tfail(nullptr, code, m); };
  t->cfg = g; t->device = device; t->sms = sms; t->world = g.world; t->rank = g.rank;
  if (t->cfg.global_batch <= 0) t->cfg.global_batch = g.B * g.world;
  if (t->cfg.eps <= 0) t->cfg.eps = 1e-5f;
  t->B = g.B; t->H = g.H; t->W = g.W; t->OS = g.OS; t->Cin = g.Cin; t->Cs = g.Cskip; t->NC = g.NC;
  t->NCp = static_cast<int>(rup(g.NC, 8)); t->Bp = static_cast<int>(rup(g.B, 8));
  t->h = ceil_div(g.H, g.OS); t->w = ceil_div(g.W, g.OS); t->hs = ceil_div(g.H, 4); t->ws = ceil_div(g.W, 4);
  if (g.lite) { t->hs = t->h; t->ws = t->w; t->Cs = 0; }      // no decoder: the classifier works at the feature resolution, there is no skip input
  t->M1 = g.B * t->h * t->w; t->M2 = g.B * t->hs * t->ws;
  if (g.OS == 8) { t->rates[0] = 12; t->rates[1] = 24; t->rates[2] = 36; } else if (g.OS == 32) { t->rates[0] = 3; t->rates[1] = 6; t->rates[2] = 9; }
  if (g.Cin % 8 || g.Cin < 8 || (!g.lite && (g.Cskip % 8 || g.Cskip < 8)) || t->M1 % 8 || t->M2 % 8) {
    t->err = "trainer: Cin, Cskip and the pixel counts per replica must be multiples of 8";
    return bail(DLV3P_ERR_INVALID);
  }
  tr_layout(t);
  const size_t M1 = t->M1, M2 = t->M2, Cin = t->Cin, Cs = t->Cs, NCp = t->NCp, Bp = t->Bp, B = t->B;
  // ---- exchange buffer: [stats partials | BN-gradient partials | the flat gradient buffer | reduced bucket | loss share]
  t->pay_stats = 0; t->pay_bn = rup(t->nstats, 4); t->pay_grads = t->pay_bn + rup(t->nparams - t->endB, 4);
  t->pay_red = t->pay_grads + rup(t->nparams, 4); t->pay_loss = t->pay_red + rup(t->endB, 4);
  if ((r = dlv3p_p2p_create(device, t->world, t->rank, t->pay_loss + 4, &t->comm, ipc_handle_out))) { t->err = g_tls_error; return bail(r); }
  t->stats_send = static_cast<float*>(dlv3p_p2p_payload(t->comm, t->pay_stats));
  t->bn_send = static_cast<float*>(dlv3p_p2p_payload(t->comm, t->pay_bn));
  t->grads = static_cast<float*>(dlv3p_p2p_payload(t->comm, t->pay_grads));       // the backward kernels write straight into peer-visible memory
  t->loss_send = static_cast<float*>(dlv3p_p2p_payload(t->comm, t->pay_loss));
  if ((r = tr_alloc(t, &t->params, t->nparams)) || (r = tr_alloc(t, &t->velocity, t->nparams)) || (r = tr_alloc(t, &t->stats, t->nstats + 4)) ||
      (r = tr_alloc(t, &t->moving_mean, t->nbn)) || (r = tr_alloc(t, &t->moving_var, t->nbn)) || (r = tr_alloc(t, &t->w_kn, t->endA)) ||
      (r = tr_alloc(t, &t->w_nk, t->endA)) || (r = tr_alloc(t, &t->seed_d, 1)) || (r = tr_alloc(t, &t->mov_ix, 3 * t->nbn)))
    return bail(r);
  {
    const uint32_t s0 = tr_dropout_seed(g.seed, 0, static_cast<uint32_t>(g.rank));
    cudaMemcpy(t->seed_d, &s0, 4, cudaMemcpyHostToDevice);
    std::vector<int> ix(3 * t->nbn);
    size_t k = 0;
    for (auto& [name, C] : t->bn_specs) {
      const size_t o = t->stat_off.at(name).first;
      for (int c = 0; c < C; ++c, ++k) { ix[k] = static_cast<int>(o + c); ix[t->nbn + k] = static_cast<int>(o + C + c); ix[2 * t->nbn + k] = static_cast<int>(o + 2 * C); }
    }
    cudaMemcpy(t->mov_ix, ix.data(), ix.size() * sizeof(int), cudaMemcpyHostToDevice);
  }
  struct BS { const char* n; size_t e; size_t sz; };
  const size_t maxC = std::max<size_t>(Cin, 304);
  t->partial_floats = 16 * 1024 * 1024 + static_cast<size_t>(sms) * 2 * 128 * 256;
  const BS bufs[] = {
      {"pool", Bp * Cin, 2}, {"r4", Bp * 256, 2}, {"b4", Bp * 256, 2}, {"r0", M1 * 256, 2}, {"concat", M1 * 1280, 2}, {"rp", M1 * 256, 2}, {"yproj", M1 * 256, 2},
      {"aspp_out", M1 * 256, 2}, {"d1", M1 * Cin, 2}, {"a1", M1 * Cin, 2}, {"p1", M1 * 256, 2}, {"d2", M1 * Cin, 2}, {"a2", M1 * Cin, 2}, {"p2", M1 * 256, 2},
      {"d3", M1 * Cin, 2}, {"a3", M1 * Cin, 2}, {"p3", M1 * 256, 2}, {"dcat", M2 * 304, 2}, {"rs", M2 * 48, 2}, {"c0d", M2 * 304, 2}, {"c0a", M2 * 304, 2},
      {"c0p", M2 * 256, 2}, {"y0", M2 * 256, 2}, {"c1d", M2 * 256, 2}, {"c1a", M2 * 256, 2}, {"c1p", M2 * 256, 2}, {"y1", M2 * 256, 2}, {"logits", M2 * NCp, 4},
      {"dfull", B * t->NC * static_cast<size_t>(t->H) * t->W, 4}, {"loss", 4, 4}, {"class_w", static_cast<size_t>(std::max(t->NC, 1)), 4},
      {"adj_tmp", B * t->NC * static_cast<size_t>(t->hs) * t->W, 4}, {"bias_stats", 2 * NCp + 4, 4}, {"dlow", M2 * NCp, 2}, {"g256a", M2 * 256, 2},
      {"g256b", M2 * 256, 2}, {"g256c", M2 * 256, 2}, {"g304a", M2 * 304, 2}, {"g304b", M2 * 304, 2}, {"drs", M2 * 48, 2}, {"dskip", M2 * Cs, 2},
      {"da_out", M1 * 256, 2}, {"drp", M1 * 256, 2}, {"dconcat", M1 * 1280, 2}, {"gB", M1 * Cin, 2}, {"gp1", M1 * 256, 2}, {"ga1", M1 * Cin, 2},
      {"gp2", M1 * 256, 2}, {"ga2", M1 * Cin, 2}, {"gp3", M1 * 256, 2}, {"ga3", M1 * Cin, 2}, {"dfeat", M1 * Cin, 2}, {"dfeat_tmp", M1 * Cin, 2},
      {"g1_256", M1 * 256, 2}, {"db4", Bp * 256, 2}, {"dr4", Bp * 256, 2}, {"dpool", Bp * Cin, 2}, {"partial", t->partial_floats, 4},
      {"scratch", dlv3p_train_scratch_bytes(static_cast<int>(maxC)) / 4 + 64, 4}, {"bn_scratch", dlv3p_op_bn_scratch_bytes(static_cast<int>(maxC)) / 4 + 64, 4},
      {"loss_scratch", dlv3p_train_loss_scratch_bytes() / 4 + 16, 4}};
  // buffers the lite variant (ASPP_Lite_block, no decoder) never touches stay one element long
  static const char* const kFullOnly[] = {"d1", "a1", "p1", "d2", "a2", "p2", "d3", "a3", "p3", "dcat", "rs", "c0d", "c0a", "c0p", "y0", "c1d", "c1a", "c1p", "y1",
                                          "g256a", "g256b", "g256c", "g304a", "g304b", "drs", "dskip", "gB", "gp1", "ga1", "gp2", "ga2", "gp3", "ga3", "dfeat_tmp"};
  for (const BS& b : bufs) {
    bool unused = false;
    if (g.lite)
      for (const char* n : kFullOnly) unused = unused || !strcmp(n, b.n);
    if ((r = tr_buf(t, b.n, unused ? 1 : b.e, b.sz))) return bail(r);
  }
  {
    std::vector<float> ones(std::max(t->NC, 1), 1.0f), var1(t->nbn, 1.0f);
    cudaMemcpy(t->T.at("class_w"), ones.data(), ones.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(t->moving_var, var1.data(), var1.size() * 4, cudaMemcpyHostToDevice);
  }
  t->feat_bytes = M1 * Cin * 2; t->skip_bytes = M2 * Cs * 2; t->labels_bytes = B * static_cast<size_t>(t->H) * t->W;
  if ((r = tr_alloc(t, reinterpret_cast<char**>(&t->feat_s), t->feat_bytes)) || (r = tr_alloc(t, reinterpret_cast<char**>(&t->skip_s), t->skip_bytes)) ||
      (r = tr_alloc(t, reinterpret_cast<char**>(&t->labels_s), t->labels_bytes)))
    return bail(r);
  if (cudaStreamCreateWithFlags(&t->stream, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreateWithFlags(&t->ev_in, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&t->ev_out, cudaEventDisableTiming) != cudaSuccess) {
    t->err = "trainer: stream / event creation failed";
    return bail(DLV3P_ERR_CUDA);
  }
  *out = t;
  return DLV3P_OK;
}

int dlv3p_trainer_connect(dlv3p_trainer* t, const uint8_t* handles) {
  if (!t) return tfail(nullptr, DLV3P_ERR_INVALID, "null trainer");
  int r = dlv3p_p2p_connect(t->comm, handles);
  return r ? tfail(t, r, g_tls_error) : DLV3P_OK;
}

int dlv3p_trainer_num_params(const dlv3p_trainer* t, int64_t* nparams, int64_t* nbn) {
  if (!t) return tfail(nullptr, DLV3P_ERR_INVALID, "null trainer");
  if (nparams) *nparams = static_cast<int64_t>(t->nparams);
  if (nbn) *nbn = static_cast<int64_t>(t->nbn);
  return DLV3P_OK;
}

// which: 0 master weight, 1 gradient, 2 velocity.  layer / var as in dlv3p_weight_info; the arrays travel in the TRAINER's layouts:
// kernel [K, N] (classifier N = NC), bias [NC], depthwise_kernel [9, C], gamma / beta / moving_mean / moving_variance [C].
static int trainer_locate(dlv3p_trainer* t, const char* layer, const char* var, size_t* off, int* rows, int* cols, int* cols_valid, bool* moving) {
  std::string ln(layer);
  if (ln == "logits_semantic") ln = "conv_upsample";
  const std::string v(var);
  *moving = false;
  if (v == "moving_mean" || v == "moving_variance") {
    size_t o = 0;
    for (auto& [name, C] : t->bn_specs) {
      if (name == ln) { *off = o; *rows = 1; *cols = C; *cols_valid = C; *moving = true; return 0; }
      o += C;
    }
    return tfail(t, DLV3P_ERR_NAME, fmt("trainer: unknown BatchNorm layer %s", layer));
  }
  auto it = t->off.find(ln + "/" + v);
  if (it == t->off.end()) return tfail(t, DLV3P_ERR_NAME, fmt("trainer: unknown variable %s/%s", layer, var));
  *off = it->second.off;
  if (v == "kernel" || v == "depthwise_kernel") { *rows = it->second.d0; *cols = it->second.d1; } else { *rows = 1; *cols = it->second.d0; }
  *cols_valid = (ln == "conv_upsample") ? t->NC : *cols;
  return 0;
}

int dlv3p_trainer_set_weight(dlv3p_trainer* t, const char* layer, const char* var, const float* host, int64_t n) {
  if (!t || !layer || !var || !host) return tfail(t, DLV3P_ERR_INVALID, "null argument");
  size_t off; int rows, cols, cv; bool moving;
  int r = trainer_locate(t, layer, var, &off, &rows, &cols, &cv, &moving);
  if (r) return r;
  if (n != static_cast<int64_t>(rows) * cv) return tfail(t, DLV3P_ERR_NAME, fmt("trainer: %s/%s has %d x %d values, got %lld", layer, var, rows, cv, (long long)n));
  CU_TRY(nullptr, cudaSetDevice(t->device));
  std::vector<float> tmp(static_cast<size_t>(rows) * cols, 0.0f);
  for (int i = 0; i < rows; ++i) std::memcpy(&tmp[static_cast<size_t>(i) * cols], host + static_cast<size_t>(i) * cv, cv * sizeof(float));
  float* dst = moving ? (std::string(var) == "moving_mean" ? t->moving_mean : t->moving_var) + off : t->params + off;
  CU_TRY(nullptr, cudaMemcpy(dst, tmp.data(), tmp.size() * sizeof(float), cudaMemcpyHostToDevice));
  if (t->gexec) { cudaGraphExecDestroy(t->gexec); t->gexec = nullptr; }
  return DLV3P_OK;
}

// after the last dlv3p_trainer_set_weight: velocity <- 0, bf16 operand copies refreshed
int dlv3p_trainer_commit_weights(dlv3p_trainer* t) {
  if (!t) return tfail(nullptr, DLV3P_ERR_INVALID, "null trainer");
  CU_TRY(nullptr, cudaSetDevice(t->device));
  CU_TRY(nullptr, cudaMemset(t->velocity, 0, t->nparams * sizeof(float)));
  Step S{t, t->stream};
  S.refresh_bf16();
  if (S.rc) return S.rc;
  CU_TRY(nullptr, cudaStreamSynchronize(t->stream));
  return DLV3P_OK;
}

int dlv3p_trainer_get(dlv3p_trainer* t, int which, const char* layer, const char* var, float* host, int64_t n) {
  if (!t || !layer || !var || !host) return tfail(t, DLV3P_ERR_INVALID, "null argument");
  size_t off; int rows, cols, cv; bool moving;
  int r = trainer_locate(t, layer, var, &off, &rows, &cols, &cv, &moving);
  if (r) return r;
  if (n != static_cast<int64_t>(rows) * cv) return tfail(t, DLV3P_ERR_NAME, fmt("trainer: %s/%s has %d x %d values, got %lld", layer, var, rows, cv, (long long)n));
  if (moving && which != 0) return tfail(t, DLV3P_ERR_INVALID, "trainer: moving statistics have no gradient / velocity");
  CU_TRY(nullptr, cudaSetDevice(t->device));
  CU_TRY(nullptr, cudaStreamSynchronize(t->stream));
  const float* src = moving ? (std::string(var) == "moving_mean" ? t->moving_mean : t->moving_var) + off
                            : (which == 0 ? t->params : which == 1 ? t->grads : t->velocity) + off;
  std::vector<float> tmp(static_cast<size_t>(rows) * cols);
  CU_TRY(nullptr, cudaMemcpy(tmp.data(), src, tmp.size() * sizeof(float), cudaMemcpyDeviceToHost));
  for (int i = 0; i < rows; ++i) std::memcpy(host + static_cast<size_t>(i) * cv, &tmp[static_cast<size_t>(i) * cols], cv * sizeof(float));
  return DLV3P_OK;
}

int dlv3p_trainer_set_class_weights(dlv3p_trainer* t, const float* host, int n) {
  if (!t || !host || n != t->NC) return tfail(t, DLV3P_ERR_INVALID, "trainer: class_weights must have one entry per class");
  CU_TRY(nullptr, cudaSetDevice(t->device));
  CU_TRY(nullptr, cudaMemcpy(t->T.at("class_w"), host, n * sizeof(float), cudaMemcpyHostToDevice));
  return DLV3P_OK;
}

// lr / momentum / l2 travel by value into the SGD kernels: a change drops the captured graph, the next step re-captures
int dlv3p_trainer_set_hyper(dlv3p_trainer* t, float lr, float momentum, float l2) {
  if (!t) return tfail(nullptr, DLV3P_ERR_INVALID, "null trainer");
  if (lr != t->cfg.lr || momentum != t->cfg.momentum || l2 != t->cfg.l2) {
    t->cfg.lr = lr; t->cfg.momentum = momentum; t->cfg.l2 = l2;
    if (t->gexec) { cudaGraphExecDestroy(t->gexec); t->gexec = nullptr; }
  }
  return DLV3P_OK;
}

static int trainer_inputs(dlv3p_trainer* t, const void* d_feat, const void* d_skip, const void* d_labels, cudaStream_t user) {
  CU_TRY(nullptr, cudaSetDevice(t->device));
  CU_TRY(nullptr, cudaEventRecord(t->ev_in, user));            // the caller's stream produced the inputs
  CU_TRY(nullptr, cudaStreamWaitEvent(t->stream, t->ev_in, 0));
  CU_TRY(nullptr, cudaMemcpyAsync(t->feat_s, d_feat, t->feat_bytes, cudaMemcpyDeviceToDevice, t->stream));
  if (t->skip_bytes) CU_TRY(nullptr, cudaMemcpyAsync(t->skip_s, d_skip, t->skip_bytes, cudaMemcpyDeviceToDevice, t->stream));
  CU_TRY(nullptr, cudaMemcpyAsync(t->labels_s, d_labels, t->labels_bytes, cudaMemcpyDeviceToDevice, t->stream));
  return DLV3P_OK;
}
static int trainer_done(dlv3p_trainer* t, cudaStream_t user) {
  CU_TRY(nullptr, cudaEventRecord(t->ev_out, t->stream));
  CU_TRY(nullptr, cudaStreamWaitEvent(user, t->ev_out, 0));     // later work of the caller's stream sees the step's results
  return DLV3P_OK;
}

// Piecewise entry points (parity tests): forward + loss + backward | gradient exchange | optimizer step.  Eager launches.
int dlv3p_trainer_forward_backward(dlv3p_trainer* t, const void* d_feat, const void* d_skip, const void* d_labels, void* cuda_stream) {
  if (!t || !d_feat || (!d_skip && !t->cfg.lite) || !d_labels) return tfail(t, DLV3P_ERR_INVALID, "null argument");
  cudaStream_t user = static_cast<cudaStream_t>(cuda_stream);
  int r = trainer_inputs(t, d_feat, d_skip, d_labels, user);
  if (r) return tfail(t, r, g_tls_error);
  Step S{t, t->stream};
  S.forward_backward(t->feat_s, t->skip_s, t->labels_s);
  if (S.rc) return S.rc;
  return trainer_done(t, user);
}
int dlv3p_trainer_all_reduce_gradients(dlv3p_trainer* t, void* cuda_stream) {
  if (!t) return tfail(nullptr, DLV3P_ERR_INVALID, "null trainer");
  Step S{t, t->stream};
  S.all_reduce_gradients();
  if (S.rc) return S.rc;
  return trainer_done(t, static_cast<cudaStream_t>(cuda_stream));
}
int dlv3p_trainer_apply_gradients(dlv3p_trainer* t, void* cuda_stream) {
  if (!t) return tfail(nullptr, DLV3P_ERR_INVALID, "null trainer");
  Step S{t, t->stream};
  S.apply_gradients();
  if (S.rc) return S.rc;
  ++t->step_count;
  return trainer_done(t, static_cast<cudaStream_t>(cuda_stream));
}

// One optimizer step (fit's train_step).  The first call launches the ~190 kernels one by one; with use_graph the second call captures
// the whole step — kernels, the peer-memory exchanges, the seed / epoch increments — into ONE CUDA graph that later calls replay
// (the step is launch bound otherwise).  Inputs are copied into static buffers the graph reads.  Asynchronous.
int dlv3p_trainer_step(dlv3p_trainer* t, const void* d_feat, const void* d_skip, const void* d_labels, int use_graph, void* cuda_stream) {
  if (!t || !d_feat || (!d_skip && !t->cfg.lite) || !d_labels) return tfail(t, DLV3P_ERR_INVALID, "null argument");
  cudaStream_t user = static_cast<cudaStream_t>(cuda_stream);
  int r = trainer_inputs(t, d_feat, d_skip, d_labels, user);
  if (r) return tfail(t, r, g_tls_error);
  auto body = [&]() {
    Step S{t, t->stream};
    S.forward_backward(t->feat_s, t->skip_s, t->labels_s);
    S.all_reduce_gradients();
    S.apply_gradients();
    return S.rc;
  };
  if (!use_graph || t->step_count == 0) {
    if ((r = body())) return r;
  } else {
    if (!t->gexec) {
      if (t->graph) { cudaGraphDestroy(t->graph); t->graph = nullptr; }
      const long long n0 = t->launches;
      CU_TRY(nullptr, cudaStreamBeginCapture(t->stream, cudaStreamCaptureModeThreadLocal));
      r = body();
      cudaError_t e = cudaStreamEndCapture(t->stream, &t->graph);
      if (r) return r;
      if (e != cudaSuccess) return tfail(t, DLV3P_ERR_CUDA, fmt("trainer: graph capture failed: %s", cudaGetErrorString(e)));
      CU_TRY(nullptr, cudaGraphInstantiate(&t->gexec, t->graph, 0));
      t->launches_per_step = t->launches - n0;
      t->launches = n0;
    }
    CU_TRY(nullptr, cudaGraphLaunch(t->gexec, t->stream));
    t->launches += t->launches_per_step;
  }
  ++t->step_count;
  return trainer_done(t, user);
}

// global mean loss of the last step (synchronises); [1] = valid pixels over all replicas
int dlv3p_trainer_loss(dlv3p_trainer* t, float* loss_out, float* valid_pixels_out) {
  if (!t || !loss_out) return tfail(t, DLV3P_ERR_INVALID, "null argument");
  CU_TRY(nullptr, cudaSetDevice(t->device));
  CU_TRY(nullptr, cudaStreamSynchronize(t->stream));
  float v[4] = {0, 0, 0, 0};
  CU_TRY(nullptr, cudaMemcpy(v, t->T.at("loss"), sizeof(v), cudaMemcpyDeviceToHost));
  *loss_out = v[0];
  if (valid_pixels_out) *valid_pixels_out = v[1];
  return DLV3P_OK;
}

// named activation / gradient buffer as fp32 on the host ("dfeat" [M1, Cin], "dskip" [M2, Cskip], "logits" fp32 [M2, NCp], ...)
int dlv3p_trainer_read(dlv3p_trainer* t, const char* name, float* host, int64_t n) {
  if (!t || !name || !host) return tfail(t, DLV3P_ERR_INVALID, "null argument");
  auto it = t->T.find(name);
  if (it == t->T.end()) return tfail(t, DLV3P_ERR_NAME, fmt("trainer: no buffer named %s", name));
  CU_TRY(nullptr, cudaSetDevice(t->device));
  CU_TRY(nullptr, cudaStreamSynchronize(t->stream));
  const std::string nm(name);
  const bool f32 = nm == "logits" || nm == "dfull" || nm == "loss" || nm == "class_w" || nm == "adj_tmp" || nm == "bias_stats" || nm == "partial";
  if (f32) {
    CU_TRY(nullptr, cudaMemcpy(host, it->second, static_cast<size_t>(n) * 4, cudaMemcpyDeviceToHost));
  } else {
    std::vector<uint16_t> tmp(static_cast<size_t>(n));
    CU_TRY(nullptr, cudaMemcpy(tmp.data(), it->second, tmp.size() * 2, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < tmp.size(); ++i) host[i] = bf16_to_f32(tmp[i]);
  }
  return DLV3P_OK;
}

int dlv3p_trainer_counters(const dlv3p_trainer* t, int64_t* launches, int64_t* steps, int* graph_captured) {
  if (!t) return tfail(nullptr, DLV3P_ERR_INVALID, "null trainer");
  if (launches) *launches = t->launches;
  if (steps) *steps = t->step_count;
  if (graph_captured) *graph_captured = t->gexec != nullptr;
  return DLV3P_OK;
}

// SHA-free replica check: 64-bit FNV-1a of the fp32 master weights (synchronises)
int dlv3p_trainer_weights_digest(dlv3p_trainer* t, uint64_t* digest) {
  if (!t || !digest) return tfail(t, DLV3P_ERR_INVALID, "null argument");
  CU_TRY(nullptr, cudaSetDevice(t->device));
  CU_TRY(nullptr, cudaStreamSynchronize(t->stream));
  std::vector<uint8_t> h(t->nparams * 4);
  CU_TRY(nullptr, cudaMemcpy(h.data(), t->params, h.size(), cudaMemcpyDeviceToHost));
  uint64_t d = 1469598103934665603ull;
  for (uint8_t b : h) { d ^= b; d *= 1099511628211ull; }
  *digest = d;
  return DLV3P_OK;
}

}  // extern "C"
