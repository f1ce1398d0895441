/*
 * dlv3p.h — C ABI of libdlv3p.so: the DeepLabV3+ encoder head (ASPP / ASPP-Lite /
 * Decoder / prediction tail) as hand-written sm_100a CUDA kernels.
 *
 * The reference (david8862/tf-keras-deeplabv3p-model-set) has NO native / FFI
 * interface for this path: it is three Python graph-builder functions plus six
 * lines of tail code executed by TensorFlow.  Each entry point below therefore
 * cites the reference *Python* site it stands in for:
 *
 *   ASPP_block(x, OS)                deeplabv3p/models/layers.py:114-163
 *   ASPP_Lite_block(x)               deeplabv3p/models/layers.py:166-196
 *   Decoder_block(x, skip)           deeplabv3p/models/layers.py:199-219
 *   SepConv_BN(...)                  deeplabv3p/models/layers.py:74-111
 *   tail: conv_upsample/pred_resize/Softmax   deeplabv3p/model.py:75-86
 *   host argmax                      deeplab.py:99, eval.py:35,
 *                                    inference/MNN/deeplabSegment.cpp:160-169
 *   weight loading                   deeplabv3p/model.py:102-103 (Keras layer names,
 *                                    SURVEY.md §8(b) table)
 *
 * Conventions: every function returns 0 (DLV3P_OK) or a negative dlv3p_status;
 * nothing throws, nothing aborts.  A context is bound to one CUDA device and is
 * NOT thread-safe.  All tensors are NHWC.  Device pointers are owned by the
 * caller; kernels are enqueued asynchronously on the stream given.
 * There is no CPU fallback: without a sm_100 device dlv3p_create fails with
 * DLV3P_ERR_CUDA / DLV3P_ERR_UNSUPPORTED.
 */
#ifndef DLV3P_H_
#define DLV3P_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DLV3P_ABI_VERSION 1

typedef enum dlv3p_status {
  DLV3P_OK = 0,
  DLV3P_ERR_INVALID = -1,     /* bad argument / config (reference: ValueError, layers.py:126) */
  DLV3P_ERR_CUDA = -2,        /* CUDA runtime / driver error, see dlv3p_last_error */
  DLV3P_ERR_UNSUPPORTED = -3, /* valid request this build does not implement */
  DLV3P_ERR_STATE = -4,       /* call order (forward before finalize, ...) */
  DLV3P_ERR_NOMEM = -5,
  DLV3P_ERR_NAME = -6         /* unknown layer / variable name, or shape mismatch */
} dlv3p_status;

/* which blocks of the head run (bitmask). Mirrors the reference call sites:
 * constructors call ASPP_block then Decoder_block (deeplabv3p_xception.py:212-215),
 * model.py:75-86 appends the tail. */
enum {
  DLV3P_STAGE_ASPP = 1,    /* ASPP_block or ASPP_Lite_block (variant) */
  DLV3P_STAGE_DECODER = 2, /* Decoder_block */
  DLV3P_STAGE_TAIL = 4     /* conv_upsample + pred_resize (+Softmax / argmax) */
};

enum { /* variant */
  DLV3P_VARIANT_ASPP = 0,     /* full ASPP (layers.py:114) */
  DLV3P_VARIANT_ASPP_LITE = 1 /* ASPP Lite (layers.py:166) */
};

enum { /* in_dtype of the feature / skip tensors handed to forward */
  DLV3P_DTYPE_BF16 = 0,
  DLV3P_DTYPE_FP16 = 1, /* the reference's --mixed_precision activations; converted to bf16 on device */
  DLV3P_DTYPE_FP32 = 2 /* converted to bf16 on device by a cast kernel */
};

enum { /* out_mode: what forward writes to d_out */
  DLV3P_OUT_LABELS_U8 = 0,      /* uint8 [B,H,W]: argmax(resize(logits)), first-max ties  (model.py:76 + deeplab.py:99) */
  DLV3P_OUT_LOGITS_LOWRES = 1,  /* fp32 [B,NC,ho,wo] PLANAR, classifier output before pred_resize (model.py:75) */
  DLV3P_OUT_SOFTMAX = 2,        /* fp32 [B,H,W,NC]: the reference model output 'pred_mask' (model.py:86) */
  DLV3P_OUT_LOGITS_FULL = 3,    /* fp32 [B,H,W,NC]: pred_resize output (model.py:76) */
  DLV3P_OUT_FEATURES_BF16 = 4,  /* bf16 NHWC feature map of the last enabled block (no TAIL stage) */
  DLV3P_OUT_FEATURES_FP32 = 5   /* fp32 NHWC feature map of the last enabled block (no TAIL stage) */
};

enum { /* flags */
  DLV3P_FLAG_UNFUSED_DECODER = 1, /* run the decoder depthwise convs as standalone kernels (A/B + debugging) */
  DLV3P_FLAG_NO_GRAPH = 2         /* reserved */
};

typedef struct dlv3p_config {
  int32_t B;            /* batch */
  int32_t H, W;         /* model input size (pred_resize target; model.py:76) */
  int32_t OS;           /* output stride 8/16/32 -> atrous rates (layers.py:118-126) */
  int32_t h, w;         /* backbone feature size; 0 -> ceil(H/OS), ceil(W/OS) */
  int32_t hs, ws;       /* skip feature size;     0 -> ceil(H/4),  ceil(W/4)  */
  int32_t Cin;          /* backbone feature channels (multiple of 8) */
  int32_t Cskip;        /* skip channels (multiple of 8; ignored without DECODER) */
  int32_t NC;           /* classes (1..256) */
  int32_t variant;      /* DLV3P_VARIANT_* */
  int32_t stages;       /* DLV3P_STAGE_* bitmask; 0 -> ASPP|DECODER|TAIL (ASPP|TAIL for Lite) */
  int32_t in_dtype;     /* DLV3P_DTYPE_* */
  int32_t out_mode;     /* DLV3P_OUT_* */
  float   bn_eps;       /* head BatchNorm epsilon; 0 -> 1e-5 (layers.py:136 etc.) */
  int32_t flags;        /* DLV3P_FLAG_* */
} dlv3p_config;

typedef struct dlv3p_ctx dlv3p_ctx; /* opaque */

/* --- life cycle -------------------------------------------------------------------- */
int dlv3p_abi_version(void);

/* Builds a context on CUDA device `device`; allocates the workspace.  Replaces the graph
 * construction done by ASPP_block/Decoder_block (layers.py:114-219) + model.py:75-86.
 * device == -1 builds a PLAN-ONLY context: config validation, the weight inventory and the size queries
 * work (host-side tests), finalize / forward / read_tap return DLV3P_ERR_STATE — there is no CPU path. */
int dlv3p_create(const dlv3p_config* cfg, int device, dlv3p_ctx** out);
void dlv3p_destroy(dlv3p_ctx* ctx);

/* Thread-local message for ctx==NULL, else the context's last error. Never NULL. */
const char* dlv3p_last_error(const dlv3p_ctx* ctx);

/* --- weights (Keras layout, reference layer names; model.py:102-103) ------------------
 * layer: e.g. "aspp1_depthwise", "aspp1_depthwise_BN", "concat_projection", "conv_upsample"
 *        ("logits_semantic" is accepted as an alias, deeplabv3p_xception.py:218).
 * var:   "kernel" (1,1,K,N) | "depthwise_kernel" (3,3,C,1) | "bias" (N) |
 *        "gamma" | "beta" | "moving_mean" | "moving_variance" (C).
 * host_fp32 is copied; shape is checked against the config. */
int dlv3p_set_weight(dlv3p_ctx* ctx, const char* layer, const char* var,
                     const float* host_fp32, const int64_t* shape, int rank);

/* Number of weight tensors the configured stages expect, and the i-th (layer,var,shape) in
 * Keras creation order — the order load_weights(by_name=False) relies on (model.py:103). */
int dlv3p_num_weights(const dlv3p_ctx* ctx);
int dlv3p_weight_info(const dlv3p_ctx* ctx, int index, const char** layer, const char** var,
                      int64_t shape_out[4], int* rank_out);

/* Folds BatchNorm (inference form), packs weights (bf16, K-major, zero padded), uploads.
 * Fails with DLV3P_ERR_STATE if a weight is missing. */
int dlv3p_finalize_weights(dlv3p_ctx* ctx);

/* --- forward ------------------------------------------------------------------------
 * d_feat: [B,h,w,Cin] (ASPP stage on) or [B,h,w,256] (DECODER first) or [B,hs,ws,256] (TAIL only)
 * d_skip: [B,hs,ws,Cskip] or NULL when the DECODER stage is off
 * d_out : per out_mode; size from dlv3p_output_bytes.
 * Asynchronous on `cuda_stream` (a cudaStream_t; NULL = default stream). */
int dlv3p_forward(dlv3p_ctx* ctx, const void* d_feat, const void* d_skip, void* d_out,
                  void* cuda_stream);

/* Same call with HOST buffers (pageable or pinned): H2D copies, forward, D2H copy, and a
 * stream synchronize — the end-to-end path model.predict()+np.argmax takes (deeplab.py:96-99).
 * Batches of 8 or more (divisible by 4) are pipelined in four sub-batches over two internal contexts so the H2D
 * copy of one sub-batch overlaps the kernels of the previous one (pinned buffers needed for the overlap; images are
 * independent, the result is identical).  dlv3p_read_tap then still refers to the last dlv3p_forward. */
int dlv3p_forward_host(dlv3p_ctx* ctx, const void* h_feat, const void* h_skip, void* h_out);

int dlv3p_input_bytes(const dlv3p_ctx* ctx, size_t* feat_bytes, size_t* skip_bytes);
int dlv3p_output_bytes(const dlv3p_ctx* ctx, size_t* out_bytes);
int dlv3p_workspace_bytes(const dlv3p_ctx* ctx, size_t* bytes);

/* Debug / parity taps: copy an internal activation to the host as fp32 NHWC after a forward.
 * name: "aspp_out" [B,h,w,256], "decoder_in" [B,hs,ws,Cd], "decoder_conv0" / "decoder_out"
 * [B,hs,ws,256], "logits" [B,NC,ho,wo] planar, "image_pooling" [B,256].
 * Layer names follow layers.py; returns DLV3P_ERR_NAME if the tap does not exist. */
int dlv3p_read_tap(dlv3p_ctx* ctx, const char* name, float* host_out, size_t host_elems);

/* Counters for bench.py: kernels launched by the last forward, and cumulative. */
int dlv3p_launch_count(const dlv3p_ctx* ctx, int64_t* last_forward, int64_t* total);

/* Per-kernel device timing of one forward (CUDA events on `cuda_stream`, synchronises).
 * names_out[i] points into static storage; ms_out[i] milliseconds.  Returns the number of
 * kernels (<= max) or a negative status. */
int dlv3p_profile_forward(dlv3p_ctx* ctx, const void* d_feat, const void* d_skip, void* d_out,
                          void* cuda_stream, const char** names_out, float* ms_out, int max);

/* --- device-memory helpers so hosts need neither PyTorch nor cuda-python ------------ */
int dlv3p_device_count(int* n);
int dlv3p_device_info(int device, int* sm_major, int* sm_minor, int* sm_count, size_t* total_mem);
int dlv3p_dev_alloc(int device, size_t bytes, void** d_ptr);
int dlv3p_dev_free(int device, void* d_ptr);
int dlv3p_host_alloc_pinned(size_t bytes, void** h_ptr);
int dlv3p_host_free_pinned(void* h_ptr);
int dlv3p_memcpy_h2d(int device, void* d_dst, const void* h_src, size_t bytes);
int dlv3p_memcpy_d2h(int device, void* h_dst, const void* d_src, size_t bytes);
int dlv3p_dev_memset(int device, void* d_ptr, int value, size_t bytes);
int dlv3p_dev_synchronize(int device);

/* --- standalone operators (unit parity tests; same kernels the forward uses) --------- */
/* D[M,N] = epilogue(A[M,K] * W[K,N]):  y = acc*scale[n] + shift[n], optional ReLU.
 * a_bf16: device bf16 [M,K] row-major (K%8==0). w_kn_fp32: HOST fp32 [K,N] (Keras HWIO 1x1 kernel).
 * out_bf16: device bf16 [M,N].  tcgen05 / TMEM / TMA path (conv 1x1: layers.py:14-21). */
int dlv3p_op_pointwise(int device, const void* a_bf16, int64_t M, int K, int N,
                       const float* w_kn_fp32, const float* scale, const float* shift, int relu,
                       void* out_bf16, void* cuda_stream);
/* Depthwise 3x3, dilation `rate`, 'same' zero padding, + per-channel scale/shift + ReLU
 * (layers.py:100-104).  x,out: device bf16 [B,H,W,C]; w_hwc_fp32: HOST fp32 [3,3,C]. */
int dlv3p_op_depthwise(int device, const void* x_bf16, int B, int H, int W, int C, int rate,
                       const float* w_hwc_fp32, const float* scale, const float* shift, int relu,
                       void* out_bf16, void* cuda_stream);
/* Fused SepConv_BN (depth_activation=True, layers.py:74-111): depthwise 3x3 (rate) + BN + ReLU
 * computed on chip as the A-operand of the tcgen05 pointwise GEMM + BN + ReLU. */
int dlv3p_op_sepconv(int device, const void* x_bf16, int B, int H, int W, int C, int rate,
                     const float* dw_hwc_fp32, const float* dw_scale, const float* dw_shift,
                     int N, const float* pw_kn_fp32, const float* pw_scale, const float* pw_shift,
                     void* out_bf16, void* cuda_stream);
/* tf.image.resize(bilinear, half-pixel centres) (layers.py:48-50). bf16 NHWC in/out. */
int dlv3p_op_resize_bilinear(int device, const void* x_bf16, int B, int hi, int wi, int C,
                             int ho, int wo, void* out_bf16, void* cuda_stream);
/* pred_resize + argmax (model.py:76 + deeplab.py:99): logits fp32 PLANAR [B,NC,hi,wi] ->
 * uint8 labels [B,ho,wo]; first-max tie-break (deeplabSegment.cpp:160-167). */
int dlv3p_op_resize_argmax(int device, const float* logits_planar, int B, int NC, int hi, int wi,
                           int ho, int wo, uint8_t* labels, void* cuda_stream);

/* Confusion-matrix accumulation of the evaluation loop (eval.py:368-373 generate_matrix and its running sum, :403-443):
 * d_confusion[gt * NC + pred] += 1 for every pixel with gt < NC (the ignore label 255 is skipped).  Device uint8 label
 * maps (16-byte aligned), device uint64 [NC*NC] matrix that is ACCUMULATED across calls; asynchronous on the stream.
 * Integer work, bit exact. */
int dlv3p_op_confusion_matrix(int device, const uint8_t* d_pred, const uint8_t* d_gt, int64_t n, int NC,
                              unsigned long long* d_confusion, void* cuda_stream);

/* Pixel counts behind the reference's training metric Jaccard (deeplabv3p/metrics.py:30-45; train.py:141 metrics={'pred_mask': Jaccard}):
 * per image b and class i in 0..NC:  d_counts[b][0][i] = #(gt == i and pred == i), [b][1][i] = #(gt == i), [b][2][i] = #(pred == i)
 * (uint64 [B][3][NC+1], ACCUMULATED; zero it first).  Device uint8 label maps [B, n_per_image]; asynchronous; integer work, bit exact.
 * The float part (IoU per image, mean over the images that contain the class, mean over the classes present) is host arithmetic on
 * 3*(NC+1) numbers per image: dlv3p_b200.metrics.jaccard. */
int dlv3p_op_jaccard_counts(int device, const uint8_t* d_pred, const uint8_t* d_gt, int B, int64_t n_per_image, int NC,
                            unsigned long long* d_counts, void* cuda_stream);

/* Image pre / post-processing of the demo and evaluation loops (deeplab.py:81-109, eval.py:403-443), device
 * pointers, asynchronous on the stream, bit exact against the reference's own numpy / cv2 code (tests/golden/ref_pins.npz):
 *   normalize_image    common/data_utils.py:403-416  uint8 [n] -> float32(x)/127.5 - 1 as fp32, or as bf16 (out_bf16 = 1: the head's input dtype)
 *   denormalize_image  common/data_utils.py:419-433  fp32 [n] -> uint8(x*127.5 + 127.5), truncating
 *   mask_resize        common/data_utils.py:457-477  cv2.resize(mask, (wo, ho), INTER_NEAREST) of B uint8 label maps [hi, wi] -> [ho, wo] */
int dlv3p_op_normalize_image(int device, const uint8_t* d_img_u8, int64_t n, void* d_out, int out_bf16, void* cuda_stream);
int dlv3p_op_denormalize_image(int device, const float* d_img_f32, int64_t n, uint8_t* d_out_u8, void* cuda_stream);
int dlv3p_op_mask_resize_nearest(int device, const uint8_t* d_mask, int B, int hi, int wi, int ho, int wo, uint8_t* d_out, void* cuda_stream);

/* "Resize in" of the demo / evaluation loops: preprocess_image (common/data_utils.py:436-454) = PIL Image.resize((wo, ho), Image.BICUBIC)
 * of B decoded uint8 images [H, W, C] -> [ho, wo, C] (device pointers).  Pillow's arithmetic (dependency of the reference, not vendored;
 * algorithm of src/libImaging/Resample.c restated): Keys cubic a = -0.5 stretched by max(scale, 1) (antialiased down-scaling), double
 * coefficients normalised and rounded to 22-bit fixed point on the host, horizontal pass first into a uint8 intermediate, each pass
 * clip8((2^21 + sum) >> 22).  Integer work, bit exact against PIL (tests).  SYNCHRONOUS: coefficient tables and the intermediate
 * image are allocated, used and freed inside the call (a pre-processing step, not part of the forward). */
/* The coefficient tables dlv3p_op_resize_bicubic_u8 builds for one axis (host arithmetic only, no device needed; parity tests on a CPU box):
 * bounds[2 * out_size] = {first input index, tap count} per output index, kk[out_size * *ksize] = 22-bit fixed-point taps, zero padded. */
int dlv3p_pil_bicubic_coeffs(int in_size, int out_size, int* bounds, int* kk, int kk_capacity, int* ksize);
int dlv3p_op_resize_bicubic_u8(int device, const uint8_t* d_img, int B, int H, int W, int C, int ho, int wo, uint8_t* d_out, void* cuda_stream);

/* Present-class set of the native post-process (inference/MNN/deeplabSegment.cpp:171-172: class_indexes, the non-background classes
 * of the mask in order of first appearance).  d_first: device uint32 [B][256], d_first[b][c] = smallest raster index of a pixel of
 * image b labelled c, 0xFFFFFFFF if absent (the call initialises it); sorting the present classes c != 0 by d_first reproduces the
 * reference's order (dlv3p_b200.ffi.op_present_classes).  Asynchronous; integer work, bit exact. */
int dlv3p_op_present_classes(int device, const uint8_t* d_labels, int B, int64_t n_per_image, unsigned int* d_first, void* cuda_stream);

/* Training-mode batch statistics of CustomBatchNormalization / SyncBatchNormalization (layers.py:63-70) — the SyncBN half of
 * the cfg-5 exchange.  x: device bf16 [M, C] (NHWC as [pixels, channels]).  d_stats: device fp32 [2*C + 1] =
 * sum_x | sum_x2 | row count; with several replicas the caller all-reduces (SUM) d_stats over NCCL (sharding.py), then
 * dlv3p_op_bn_apply normalises with mean = sum/n, var = sum2/n - mean^2 (biased) read from DEVICE memory.
 * Both calls are asynchronous on the stream.  d_scratch: dlv3p_op_bn_scratch_bytes(C) bytes of device memory. */
size_t dlv3p_op_bn_scratch_bytes(int C);
int dlv3p_op_bn_stats(int device, const void* x_bf16, int64_t M, int C, float* d_stats, void* d_scratch, void* cuda_stream);
int dlv3p_op_bn_apply(int device, const void* x_bf16, int64_t M, int C, const float* d_stats, const float* d_gamma,
                      const float* d_beta, float eps, int relu, void* y_bf16, void* cuda_stream);

/* Benchmark aid (tools/kbench.py): average ms per launch of ONE operator on synthetic device data (CUDA events).
 * op 0 pointwise {M,K,N}; 1 fused sepconv {B,H,W,C}; 2 resize {B,hi,wi,C,ho,wo}; 3 resize_argmax {B,NC,hi,wi,ho,wo}.
 * flags: per-kernel debug bits (skip stores / stencil / MMA) to attribute time; results are then meaningless. */
int dlv3p_op_time(int device, int op, const int64_t* dims, int ndims, int iters, int flags, float* ms_out);

#ifdef __cplusplus
}
#endif
#endif /* DLV3P_H_ */
