/*
 * dlv3p_train.h — C ABI of the TRAINING-step operators of libdlv3p.so (BASELINE cfg 5: DeepLabV3+ head training step
 * with SyncBatchNorm, data parallel over NCCL).  Same conventions as dlv3p.h: extern "C", plain pointers and sizes, 0 or a
 * negative dlv3p_status, nothing throws; every call is ASYNCHRONOUS on `cuda_stream`, every pointer is DEVICE memory owned
 * by the caller, nothing is allocated per call.  There is no CPU path.
 *
 * The reference trains the head through Keras (train.py:143-169 model.compile / fit under tf.distribute.MirroredStrategy):
 * the forward graph of layers.py:74-219 + model.py:75-86 in training mode (batch statistics, Dropout active), the loss
 * deeplabv3p/loss.py:121-156, TensorFlow's autodiff, SGD(momentum 0.9) common/model_utils.py:122-123.  These entry points
 * are the building blocks of that step; dlv3p_b200/train.py (HeadTrainer) strings them together and does the two
 * collectives (SyncBN statistics, gradient all-reduce) with torch.distributed / NCCL in between.
 *
 * Layouts: activations and activation gradients bf16 NHWC viewed as [pixels, channels]; 1x1 kernels in the Keras HWIO
 * layout [K, N]; depthwise kernels [3][3][C]; master weights, weight gradients and optimizer state fp32.
 */
#ifndef DLV3P_TRAIN_H_
#define DLV3P_TRAIN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* D[M, N] = A[M, K] * B[N, K]^T on tcgen05 tensor cores (bf16 operands, fp32 accumulation).  a: [M, K] row stride lda,
 * b: [N, K] row stride ldb (strides in elements, multiples of 8; K a multiple of 8; 16-byte aligned bases).
 * d: bf16 (out_fp32 = 0; ldd % 8 == 0) or fp32 (out_fp32 = 1) [M, N] with row stride ldd.
 * splits > 1 splits the contraction into that many slices (weight gradients: few output tiles, long K); the fp32 partials
 * go to d_partial (dlv3p_train_gemm_partial_bytes) and are reduced in fixed order.  Serves the forward, the data gradient
 * and (on transposed operands) the weight gradient of DeeplabConv2D 1x1 (layers.py:14-21). */
size_t dlv3p_train_gemm_partial_bytes(int64_t M, int N, int splits);
int dlv3p_train_gemm_nt(int device, const void* a_bf16, int64_t lda, const void* b_bf16, int64_t ldb, int64_t M, int N, int64_t K,
                        void* d, int64_t ldd, int out_fp32, int splits, void* d_partial, void* cuda_stream);
/* D[M, N] = A[K, M]^T * B[K, N]: both operands with the CONTRACTION index as the row (bf16 [K, M] row stride lda, [K, N] row stride
 * ldb; M, N, K, strides multiples of 8) — the weight gradient dW[Cin, Cout] = X[pixels, Cin]^T dY[pixels, Cout] read straight from
 * the NHWC activations (MN-major tcgen05 operand descriptors), no transpose pass.  Output / split-K as dlv3p_train_gemm_nt. */
int dlv3p_train_gemm_tn(int device, const void* a_bf16, int64_t lda, const void* b_bf16, int64_t ldb, int64_t M, int N, int64_t K,
                        void* d, int64_t ldd, int out_fp32, int splits, void* d_partial, void* cuda_stream);
/* out[c][r] = in[r][c]; bf16, R and C even. */
int dlv3p_train_transpose(int device, const void* in_bf16, int64_t R, int C, int64_t ld_in, void* out_bf16, int64_t ld_out, void* cuda_stream);

/* Training-mode BatchNorm (+ReLU) output into a buffer with row stride ldy (a concat slice).  x dense [M, C]; d_stats from
 * dlv3p_op_bn_stats after the SyncBN all-reduce (dlv3p.h).  CustomBatchNormalization, layers.py:63-70. */
int dlv3p_train_bn_apply(int device, const void* x_bf16, int64_t M, int C, const float* d_stats, const float* d_gamma, const float* d_beta,
                         float eps, int relu, void* y_bf16, int64_t ldy, void* cuda_stream);
/* SyncBN backward, first half: d_sums[2C] = sum g | sum g * xhat over this replica's rows, g = dy * (y > 0) when relu
 * (y = the BN+ReLU output).  The caller all-reduces d_sums (SUM); they are then d(beta) | d(gamma) of the GLOBAL batch.
 * d_scratch: dlv3p_train_scratch_bytes(C). */
size_t dlv3p_train_scratch_bytes(int C);
int dlv3p_train_bn_bwd_stats(int device, const void* dy_bf16, int64_t ld_dy, const void* y_bf16, int64_t ld_y, const void* x_bf16, int64_t M, int C,
                             const float* d_stats, float eps, int relu, float* d_sums, void* d_scratch, void* cuda_stream);
/* second half: dx = gamma * invstd * (g - S1/n - xhat * S2/n), dense bf16 [M, C]. */
int dlv3p_train_bn_bwd_apply(int device, const void* dy_bf16, int64_t ld_dy, const void* y_bf16, int64_t ld_y, const void* x_bf16, int64_t M, int C,
                             const float* d_stats, const float* d_sums, const float* d_gamma, float eps, int relu, void* dx_bf16, void* cuda_stream);

/* DeeplabDepthwiseConv2D 3x3, dilation `rate`, 'same' (layers.py:24-31) with DEVICE fp32 taps [3][3][C], no BN / ReLU.
 * flip = 0: forward.  flip = 1: the data gradient (same convolution with the taps mirrored). */
int dlv3p_train_depthwise(int device, const void* x_bf16, int B, int H, int W, int C, int rate, const float* d_taps, int flip, void* out_bf16,
                          void* cuda_stream);
/* weight gradient dW[3][3][C] (fp32) = sum over pixels of x(shifted) * dy.  d_scratch: dlv3p_train_scratch_bytes(C). */
int dlv3p_train_depthwise_wgrad(int device, const void* x_bf16, const void* dy_bf16, int B, int H, int W, int C, int rate, float* d_dw,
                                void* d_scratch, void* cuda_stream);

/* tf.image.resize bilinear forward into a concat slice (layers.py:48-60, :207): out rows of stride ld_out. */
int dlv3p_train_resize(int device, const void* x_bf16, int B, int hi, int wi, int C, int ho, int wo, void* out_bf16, int64_t ld_out,
                       void* cuda_stream);
/* its adjoint: dx[B,hi,wi,C] dense from dy[B,ho,wo] rows of stride ld_dy. */
int dlv3p_train_resize_bwd(int device, const void* dy_bf16, int64_t ld_dy, int B, int hi, int wi, int C, int ho, int wo, void* dx_bf16,
                           void* cuda_stream);

/* pred_resize + Softmax + SparseCategoricalCrossEntropy(ignore_index) (model.py:76-86, loss.py:121-156) and its gradient.
 * logits: fp32 rows [B*hi*wi, ldl] (classifier output WITHOUT bias), bias fp32 [NC], labels uint8 [B,H,W].
 * d_full: planar fp32 [B, NC, H, W] = d(loss)/d(full-resolution logits), loss normalised by inv_norm = 1/(global B*H*W).
 * d_loss[0] = this replica's share of the mean loss, d_loss[1] = its valid pixels.  d_scratch: dlv3p_train_loss_scratch_bytes(). */
size_t dlv3p_train_loss_scratch_bytes(void);
int dlv3p_train_softmax_ce(int device, const float* logits, int64_t ldl, const float* bias, const uint8_t* labels, int B, int NC, int hi, int wi,
                           int H, int W, int ignore_index, float inv_norm, float* d_full, float* d_loss, void* d_scratch, void* cuda_stream);
/* The same with the reference's other two losses (train.py:114-138): kind 0 = dlv3p_train_softmax_ce; kind 1 =
 * WeightedSparseCategoricalCrossEntropy (loss.py:159-192: -w[label] * log p, device class weights [NC], no clip); kind 2 =
 * SparseSoftmaxFocalLoss (loss.py:60-118: -alpha * (1-p)^gamma * log(clip(p, 1e-15)); reference defaults gamma 2, alpha 0.25). */
int dlv3p_train_softmax_loss(int device, const float* logits, int64_t ldl, const float* bias, const uint8_t* labels, int B, int NC, int hi, int wi,
                             int H, int W, int ignore_index, float inv_norm, int kind, const float* d_class_weights, float focal_gamma,
                             float focal_alpha, float* d_full, float* d_loss, void* d_scratch, void* cuda_stream);
/* adjoint of pred_resize: planar fp32 [B,NC,H,W] -> bf16 rows [B*hi*wi, ld_dx] (columns >= NC untouched).  With d_scratch
 * (dlv3p_train_resize_bwd_planar_scratch_bytes(B, NC, hi, W) bytes) the adjoint runs as two separable passes; NULL = one pass. */
size_t dlv3p_train_resize_bwd_planar_scratch_bytes(int B, int NC, int hi, int W);
int dlv3p_train_resize_bwd_planar(int device, const float* d_full, int B, int NC, int hi, int wi, int H, int W, void* dx_bf16, int64_t ld_dx,
                                  void* d_scratch, void* cuda_stream);

/* out[b][c] = scale * sum over the image's npix rows of x[., c] (x rows of stride ld): AveragePooling2D over the whole map
 * (layers.py:132) with scale = 1/npix; per-image column sums (adjoint of the aspp_resize broadcast) with scale = 1.
 * out: bf16 [B, C] (out_fp32 = 0) or fp32. */
int dlv3p_train_rows_reduce(int device, const void* x_bf16, int64_t ld, int B, int npix, int C, float scale, void* out, int out_fp32, void* cuda_stream);
/* dst[(b*npix+p)*ld + c] = (accumulate ? dst : 0) + scale * src[b][c]: aspp_resize of the 1x1 pooled map (layers.py:138)
 * and the adjoint of the mean. */
int dlv3p_train_bcast_rows(int device, const void* src_bf16, int B, int npix, int C, float scale, void* dst_bf16, int64_t ld, int accumulate,
                           void* cuda_stream);
/* out = a + b (bf16, n % 8 == 0; in place allowed). */
int dlv3p_train_add(int device, const void* a_bf16, const void* b_bf16, void* out_bf16, int64_t n, void* cuda_stream);
/* Dropout(rate) (layers.py:161, :194) with a counter-based mask (seed, element index); the same call masks the gradient.
 * d_seed != NULL: the seed is read from device memory at execution time (a captured CUDA graph then draws a new mask per replay). */
int dlv3p_train_dropout(int device, const void* x_bf16, void* out_bf16, int64_t n, uint32_t seed, const uint32_t* d_seed, float rate,
                        void* cuda_stream);
/* SGD(momentum) step on fp32 master weights with the l2 regulariser folded in: g' = gscale*g + 2*l2*w; v = m*v - lr*g'; w += v
 * (common/model_utils.py:122-123, layers.py:12-21). */
int dlv3p_train_sgd(int device, float* w, const float* g, float* v, int64_t n, float lr, float momentum, float l2, float gscale, void* cuda_stream);
/* fp32 -> bf16 copy of a weight tensor (any n). */
int dlv3p_train_cast_bf16(int device, const float* in, void* out_bf16, int64_t n, void* cuda_stream);

/* --- the step's exchanges over NVLink peer memory (train.py:143-158: MirroredStrategy's in-graph all-reduces) -------------------
 * One exchange buffer per replica, mapped by its peers through CUDA IPC; a collective is ONE small kernel per replica: publish a
 * flag in every peer's buffer, wait for the peers' flags, sum the W payloads in rank order with peer loads (bit-identical on every
 * replica), write the total to a local buffer.  No NCCL call, no host synchronisation: the whole step stays one CUDA graph.
 *   dlv3p_p2p_create    allocates [flags | payload_floats fp32] on `device`, returns the 64-byte cudaIpcMemHandle_t to hand to the peers
 *                       (any host channel: torch.distributed, MPI, a file)
 *   dlv3p_p2p_connect   handles = world x 64 bytes in rank order; maps the peers' buffers
 *   dlv3p_p2p_payload   device pointer of float `off` of THIS replica's payload area: the producing kernels write their partials there
 *   dlv3p_p2p_allreduce collective number `slot` (< 64, distinct per collective of a step): d_out[0..n) = sum over replicas of
 *                       payload[off .. off+n); off, n multiples of 4; asynchronous on the stream
 *   dlv3p_p2p_advance   once per step after its last collective (advances the device-side epoch the flags carry) */
typedef struct dlv3p_p2p dlv3p_p2p;
int dlv3p_p2p_create(int device, int world, int rank, size_t payload_floats, dlv3p_p2p** out, uint8_t handle_out[64]);
int dlv3p_p2p_connect(dlv3p_p2p* comm, const uint8_t* handles);
void dlv3p_p2p_destroy(dlv3p_p2p* comm);
void* dlv3p_p2p_payload(dlv3p_p2p* comm, size_t off);
int dlv3p_p2p_allreduce(dlv3p_p2p* comm, int slot, size_t off, int n, float* d_out, void* cuda_stream);
int dlv3p_p2p_advance(dlv3p_p2p* comm, void* cuda_stream);

/* --- the whole training step behind one handle (train.py:143-169: model.compile + fit under MirroredStrategy) --------------------
 * dlv3p_trainer owns every device buffer of one replica (fp32 master weights, velocity, activations, gradients, SyncBN statistics,
 * the peer-visible exchange buffer), strings the operators above into forward (training mode) -> loss -> backward -> exchanges ->
 * SGD(momentum) + l2 -> Keras moving statistics, and replays the step as ONE CUDA graph.  No PyTorch, no NCCL: replicas exchange
 * through peer memory (dlv3p_p2p_*); the host only hands the 64-byte IPC handles around once.
 * Scope: the full head (ASPP_block + Decoder_block + tail), backbone frozen / outside (the reference's stage 1, train.py:177-187). */
typedef struct dlv3p_trainer_config {
  int32_t B, H, W, OS;        /* per-replica batch, model input size, output stride */
  int32_t Cin, Cskip, NC;
  int32_t world, rank;        /* data-parallel replicas (one process per GPU) */
  int32_t global_batch;       /* loss normalisation 1 / (global_batch * H * W); 0 -> B * world */
  int32_t ignore_index;       /* 255 in the reference (loss.py:121-156) */
  int32_t loss_kind;          /* 0 sparse CE, 1 class-weighted CE (dlv3p_trainer_set_class_weights), 2 focal (train.py:114-138) */
  uint32_t seed;              /* Dropout mask seed (counter based: seed, step, rank, element) */
  float lr, momentum, l2;     /* SGD(momentum 0.9), lr 1e-2, l2(2e-5) defaults of the reference */
  float bn_momentum, eps;     /* 0.99, 1e-5 */
  float dropout;              /* 0.5 (layers.py:161) */
  float focal_gamma, focal_alpha;
  int32_t lite;               /* 0: ASPP_block + Decoder_block (xception, resnet50, mobilenetv2 / v3 ...);  1: ASPP_Lite_block and no decoder,
                                 the *_lite models (deeplabv3p_mobilenetv2.py:326-331): Cskip is ignored, d_skip may be NULL */
} dlv3p_trainer_config;
typedef struct dlv3p_trainer dlv3p_trainer;

/* ipc_handle_out: this replica's exchange buffer (cudaIpcMemHandle_t); hand it to every peer, then dlv3p_trainer_connect with all
 * handles in rank order (world == 1: nothing to connect). */
int dlv3p_trainer_create(const dlv3p_trainer_config* cfg, int device, dlv3p_trainer** out, uint8_t ipc_handle_out[64]);
int dlv3p_trainer_connect(dlv3p_trainer* t, const uint8_t* handles);
void dlv3p_trainer_destroy(dlv3p_trainer* t);
const char* dlv3p_trainer_last_error(const dlv3p_trainer* t);
/* Weights by Keras layer / variable name (dlv3p_weight_info); arrays in the trainer's layouts: kernel [K, N], bias [NC],
 * depthwise_kernel [9, C], BN vectors [C].  commit after the last set (velocity <- 0, bf16 operand copies). */
int dlv3p_trainer_set_weight(dlv3p_trainer* t, const char* layer, const char* var, const float* host_fp32, int64_t n);
int dlv3p_trainer_commit_weights(dlv3p_trainer* t);
/* which: 0 master weight (or moving statistic), 1 gradient of the last step, 2 velocity */
int dlv3p_trainer_get(dlv3p_trainer* t, int which, const char* layer, const char* var, float* host_fp32, int64_t n);
int dlv3p_trainer_set_class_weights(dlv3p_trainer* t, const float* host_fp32, int n);
int dlv3p_trainer_set_hyper(dlv3p_trainer* t, float lr, float momentum, float l2);   /* LR schedules: the captured graph is rebuilt */
/* d_feat bf16 [B,h,w,Cin], d_skip bf16 [B,H/4,W/4,Cskip], d_labels uint8 [B,H,W]: device pointers, copied into the trainer's static
 * buffers on its own stream (ordered behind `cuda_stream`); asynchronous.  use_graph: replay the captured step from the 2nd call on. */
int dlv3p_trainer_step(dlv3p_trainer* t, const void* d_feat, const void* d_skip, const void* d_labels, int use_graph, void* cuda_stream);
/* the step in three pieces (parity tests): forward + loss + backward | gradient exchange | update */
int dlv3p_trainer_forward_backward(dlv3p_trainer* t, const void* d_feat, const void* d_skip, const void* d_labels, void* cuda_stream);
int dlv3p_trainer_all_reduce_gradients(dlv3p_trainer* t, void* cuda_stream);
int dlv3p_trainer_apply_gradients(dlv3p_trainer* t, void* cuda_stream);
/* global mean loss of the last step and the valid pixels over all replicas (synchronises) */
int dlv3p_trainer_loss(dlv3p_trainer* t, float* loss_out, float* valid_pixels_out);
/* a named activation / gradient buffer as fp32: "dfeat" [B*h*w, Cin], "dskip" [B*hs*ws, Cskip], ... (parity tests, diagnostics) */
int dlv3p_trainer_read(dlv3p_trainer* t, const char* name, float* host_fp32, int64_t n);
int dlv3p_trainer_num_params(const dlv3p_trainer* t, int64_t* nparams, int64_t* nbn);
int dlv3p_trainer_counters(const dlv3p_trainer* t, int64_t* launches, int64_t* steps, int* graph_captured);
int dlv3p_trainer_weights_digest(dlv3p_trainer* t, uint64_t* digest);   /* FNV-1a of the fp32 master weights: replicas-identical check */

#ifdef __cplusplus
}
#endif
#endif /* DLV3P_TRAIN_H_ */
