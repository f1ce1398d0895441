/*
 * dlv3p_model.h — C ABI of the WHOLE DeepLabV3+ Xception model in libdlv3p.so: the modified aligned Xception feature extractor
 *  in front of the encoder head of dlv3p.h, uint8 / fp32 images in, labels / logits / probabilities out.
 *
 * Reference sites each entry point stands in for (paths relative to the reference repo):
 *   Deeplabv3pXception(input_shape, num_classes, OS)    deeplabv3p/models/deeplabv3p_xception.py:167-239
 *     Xception_body                                      :96-163   (entry flow :119-138, 16 middle-flow units :139-143, exit flow :145-152)
 *     _xception_block / _conv2d_same                     :57-93 / :25-52
 *     SepConv_BN (depth_activation False | True, stride 2 = ZeroPadding2D + 'valid', epsilon 1e-3)   deeplabv3p/models/layers.py:74-111
 *   get_deeplabv3p_model('xception', ...)                deeplabv3p/model.py:51-117 (head + tail: dlv3p.h)
 *   preprocess / normalize_image                         common/data_utils.py:403-416 (uint8 -> x/127.5 - 1, fused into the first convolution)
 *   model.predict + np.argmax                            deeplab.py:96-99
 *   model.load_weights                                   deeplabv3p/model.py:102-103 (Keras layer / variable names, creation order)
 *
 * Conventions as in dlv3p.h: 0 or a negative dlv3p_status, nothing throws; a model is bound to one CUDA device and is not
 * thread-safe; tensors are NHWC; device pointers are owned by the caller; kernels are enqueued asynchronously on the stream given.
 * There is no CPU path (device -1 builds a plan-only model: weight inventory and sizes).
 */
#ifndef DLV3P_MODEL_H_
#define DLV3P_MODEL_H_

#include <stddef.h>
#include <stdint.h>

#include "dlv3p.h"

#ifdef __cplusplus
extern "C" {
#endif

enum { /* img_dtype */
  DLV3P_IMG_U8 = 0,  /* uint8 [B,H,W,3] RGB as decoded; normalize_image (x/127.5 - 1) happens inside the first convolution */
  DLV3P_IMG_F32 = 1  /* fp32 [B,H,W,3] already normalised (what preprocess_image returns, data_utils.py:436-454) */
};

enum { /* flags */
  DLV3P_MODEL_FLAG_KEEP_ALL = 1, /* every intermediate keeps its own tensor (block-level parity taps); default: the 16 middle-flow
                                    units ping-pong over four buffers and only the last unit's tap is meaningful */
  DLV3P_MODEL_FLAG_NO_PDL = 2,   /* measurement aid: launch the backbone kernels without programmatic dependent launch */
  DLV3P_MODEL_FLAG_UNFUSED_ENTRY = 4, /* measurement aid: the entry flow's SepConv_BN layers as depthwise kernel + GEMM (like the rest of the backbone) */
  DLV3P_MODEL_FLAG_FUSED_MIDDLE = 16, /* experiment (slower): the middle flow's SepConv_BN layers through the fused two-SM cluster kernel
                                    (bb_sepwide.cuh) instead of depthwise kernel + GEMM */
  DLV3P_MODEL_FLAG_FP32_STEM = 32, /* measurement aid: entry_flow_conv1_1 on uint8 images through the fp32 CUDA-core kernel (the path float images
                                    take) instead of the tensor-core kernel (bb_stem_tc.cuh) */
  DLV3P_MODEL_FLAG_FP32 = 8      /* PRECISION MODE: the whole model in plain fp32 arithmetic (the reference's default numerics, train.py:37-46):
                                    weights as given, fp32 activations, CUDA-core kernels (csrc/f32_kernels.cuh).  Held to 1e-4 relative
                                    against the fp32 oracle; a mode to prove results, not the performance path.  out_mode: labels or
                                    low-resolution logits */
};

typedef struct dlv3p_model_config {
  int32_t B;         /* batch */
  int32_t H, W;      /* model input size */
  int32_t OS;        /* output stride 8 / 16 / 32 (deeplabv3p_xception.py:100-117) */
  int32_t NC;        /* classes */
  int32_t img_dtype; /* DLV3P_IMG_* */
  int32_t out_mode;  /* DLV3P_OUT_LABELS_U8 / LOGITS_LOWRES / SOFTMAX / LOGITS_FULL (dlv3p.h) */
  int32_t flags;     /* DLV3P_MODEL_FLAG_* */
} dlv3p_model_config;

typedef struct dlv3p_model dlv3p_model; /* opaque */

int dlv3p_model_create(const dlv3p_model_config* cfg, int device, dlv3p_model** out);
void dlv3p_model_destroy(dlv3p_model* m);
const char* dlv3p_model_last_error(const dlv3p_model* m);

/* Weights by Keras layer / variable name in Keras creation order: the backbone's (entry_flow_conv1_1 ... exit_flow_block2_*),
 * then the head's (dlv3p_weight_info).  Layouts: Conv2D kernel (kh,kw,Cin,Cout), depthwise_kernel (3,3,C,1), BN vectors (C). */
int dlv3p_model_num_weights(const dlv3p_model* m);
int dlv3p_model_weight_info(const dlv3p_model* m, int index, const char** layer, const char** var, int64_t shape_out[4], int* rank_out);
int dlv3p_model_set_weight(dlv3p_model* m, const char* layer, const char* var, const float* host_fp32, const int64_t* shape, int rank);
/* Folds BatchNorm (backbone epsilon 1e-3, head 1e-5), packs the 1x1 kernels to bf16 K-major, uploads. */
int dlv3p_model_finalize_weights(dlv3p_model* m);

/* d_images: [B,H,W,3] per img_dtype;  d_out: per out_mode (dlv3p_model_output_bytes). */
int dlv3p_model_forward(dlv3p_model* m, const void* d_images, void* d_out, void* cuda_stream);
/* Same with HOST buffers: H2D copy of the images, forward, D2H copy of the result, stream synchronize (model.predict + argmax). */
int dlv3p_model_forward_host(dlv3p_model* m, const void* h_images, void* h_out);

int dlv3p_model_input_bytes(const dlv3p_model* m, size_t* bytes);
int dlv3p_model_output_bytes(const dlv3p_model* m, size_t* bytes);
int dlv3p_model_workspace_bytes(const dlv3p_model* m, size_t* bytes);

/* Parity taps after a forward, fp32 NHWC on the host: "entry_flow_conv1_1", "entry_flow_conv1_2", the output of a block by its
 * prefix ("entry_flow_block1", "middle_flow_unit_7", "exit_flow_block2", ...), "feature" [B,h,w,2048], "skip" [B,H/4,W/4,256];
 * any other name is forwarded to dlv3p_read_tap of the head ("logits", "aspp_out", ...). */
int dlv3p_model_read_tap(dlv3p_model* m, const char* name, float* host_out, size_t host_elems);
int dlv3p_model_tap_shape(const dlv3p_model* m, const char* name, int64_t shape_out[4]);
/* Block-isolated parity: overwrite the backbone tap `tap` with host_fp32 (rounded to bf16) and run ONLY the kernels after the one
 * that produced it (then the head).  With the oracle's own intermediate as input, the next block's tap isolates that block. */
int dlv3p_model_forward_from(dlv3p_model* m, const char* tap, const float* host_fp32, size_t host_elems, void* d_out, void* cuda_stream);

/* Kernels launched by the last forward (backbone + head). */
int dlv3p_model_launch_count(const dlv3p_model* m, int64_t* last_forward);

/* Per-kernel device timing of one forward (CUDA events on the stream, synchronises): name, milliseconds, algorithmic FLOPs
 * (2 per MAC of the convolution the kernel computes; 0 for non-GEMM kernels of the head) and algorithmic HBM bytes (inputs read
 * once + outputs written once) of every launch, backbone first.  Returns the number of kernels (<= max). */
int dlv3p_model_profile_forward(dlv3p_model* m, const void* d_images, void* d_out, void* cuda_stream, const char** names_out, float* ms_out,
                                double* flops_out, double* bytes_out, int max);

/* --- standalone operators of the backbone (unit parity tests; the kernels the forward uses) --------------------------- */
/* [ReLU] -> depthwise 3x3 (stride 1 'same' | stride 2 after ZeroPadding2D(rate), dilation rate) -> scale/shift -> [ReLU]
 * (layers.py:88-104).  x: device bf16 [B,H,W,C] (C % 8 == 0); w_hwc: HOST fp32 [3,3,C]; out: device bf16 [B,Ho,Wo,C],
 * Ho = ceil(H / stride). */
int dlv3p_op_bb_depthwise(int device, const void* x_bf16, int B, int H, int W, int C, int stride, int rate, int relu_in, int relu_out,
                          const float* w_hwc_fp32, const float* scale, const float* shift, void* out_bf16, void* cuda_stream);
/* out[M,N] = bf16(acc(a[M,K] * w[K,N]) * scale + shift [ReLU] [+ residual[M,N]])  (layers.py:105-109, deeplabv3p_xception.py:82-90).
 * a, residual, out: device bf16; w_kn: HOST fp32 (Keras 1x1 kernel); K % 8 == 0, N % 8 == 0. */
int dlv3p_op_bb_pointwise(int device, const void* a_bf16, int64_t M, int K, int N, const float* w_kn_fp32, const float* scale, const float* shift,
                          int relu, const void* residual_bf16, void* out_bf16, void* cuda_stream);
/* Fused SepConv_BN of the middle flow (depth_activation False, layers.py:74-111; residual = _xception_block 'sum', deeplabv3p_xception.py:88-90):
 * out = bf16(BN2(pointwise(bf16(BN1(depthwise3x3(relu_in ? relu(x) : x))))) [ReLU] [+ residual]).  x: device bf16 [B,H,W,C] (C % 8 == 0, C <= 768);
 * dw_hwc: HOST fp32 [3,3,C]; w_kn: HOST fp32 [C,N] (656 < N <= 768, N % 8 == 0); residual / out: device bf16 [B,H,W,N]. */
int dlv3p_op_bb_sepwide(int device, const void* x_bf16, int B, int H, int W, int C, int N, int relu_in, const float* dw_hwc_fp32, const float* dw_scale,
                        const float* dw_shift, const float* w_kn_fp32, const float* scale, const float* shift, int relu_out, const void* residual_bf16,
                        void* out_bf16, void* cuda_stream);
/* Conv2D(64, 3x3, 'same') on 32 channels + scale/shift + ReLU (entry_flow_conv1_2).  x: device bf16 [B,H,W,32]; w_hwio: HOST fp32 [3,3,32,64]. */
int dlv3p_op_conv3x3_c32(int device, const void* x_bf16, int B, int H, int W, const float* w_hwio_fp32, const float* scale, const float* shift,
                         void* out_bf16, void* cuda_stream);
/* [normalize] -> Conv2D(32, 3x3, strides 2, 'same') -> scale/shift -> ReLU (entry_flow_conv1_1).  img: device uint8 or fp32 [B,H,W,3];
 * w_hwio: HOST fp32 [3,3,3,32]; out: device bf16 [B,ceil(H/2),ceil(W/2),32]. */
int dlv3p_op_stem_conv(int device, const void* img, int img_dtype, int B, int H, int W, const float* w_hwio_fp32, const float* scale, const float* shift,
                       void* out_bf16, void* cuda_stream);

/* Benchmark aid (tools/kbench_bb.py): average ms per launch of ONE backbone operator on synthetic device data (CUDA events).
 * op 0 pointwise GEMM {M, K, N, residual, BN (0 = automatic)}; op 1 depthwise {B, H, W, C, stride, rate}; op 2 fused middle-flow
 * SepConv_BN {B, H, W, C, N, residual}.
 * flags: per-kernel debug bits that switch parts of the kernel off to attribute time; results are then meaningless. */
int dlv3p_op_bb_time(int device, int op, const int64_t* dims, int ndims, int iters, int flags, float* ms_out);

#ifdef __cplusplus
}
#endif
#endif /* DLV3P_MODEL_H_ */
