"""
oracle/head_ref.py — CPU restatement of the reference's DeepLabV3+ head.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this file; the product path (libdlv3p.so + the ctypes host layer) never does.

PARITY UNPINNED: the reference ships no tests, no golden vectors and no weights, and its
arithmetic lives in TensorFlow 2.11 (requirements.txt:9), which is neither vendored under
/root/reference nor installable here.  This file restates the published semantics of the TF
ops the reference calls; it is cross-checked op by op against an independent
implementation (torch.nn.functional on CPU) in tests/test_oracle.py, not against TensorFlow.

Reference sites restated (paths relative to /root/reference):
  SepConv_BN          deeplabv3p/models/layers.py:74-111   (head use: stride 1, depth_activation=True)
  ASPP_block          deeplabv3p/models/layers.py:114-163
  ASPP_Lite_block     deeplabv3p/models/layers.py:166-196
  Decoder_block       deeplabv3p/models/layers.py:199-219
  img_resize          deeplabv3p/models/layers.py:48-60    (tf.image.resize bilinear, TF2 half-pixel)
  prediction tail     deeplabv3p/model.py:75-86            (conv_upsample, pred_resize, Softmax)
  host argmax         deeplab.py:99, eval.py:35            (np.argmax, first max wins)
  mIOU                deeplabv3p/metrics.py:10-17
  confusion matrix    eval.py:368-373

Two numeric modes:
  'fp32'  — reference semantics: everything in float32.
  'bf16'  — same graph with the rounding points of the CUDA path (operands of every GEMM and
            every inter-kernel activation rounded to bfloat16, fp32 accumulation, fp32 logits).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

BN_EPS_HEAD = 1e-5  # layers.py:136,142,147,149,152,159,212,216,218

# ------------------------------------------------------------------------------------------
# configuration
# ------------------------------------------------------------------------------------------


@dataclass
class HeadConfig:
    """Static shape description of one head instance (what the Keras graph bakes in)."""
    B: int
    H: int
    W: int
    OS: int
    Cin: int
    Cskip: int
    NC: int
    lite: bool = False          # ASPP_Lite_block instead of ASPP_block
    decoder: bool = True        # Decoder_block present (False for every *_lite model)
    h: int = 0
    w: int = 0
    hs: int = 0
    ws: int = 0
    eps: float = BN_EPS_HEAD

    def __post_init__(self):
        if self.h == 0:
            self.h = -(-self.H // self.OS)
        if self.w == 0:
            self.w = -(-self.W // self.OS)
        if self.hs == 0:
            self.hs = -(-self.H // 4)
        if self.ws == 0:
            self.ws = -(-self.W // 4)

    @property
    def rates(self) -> Tuple[int, int, int]:
        return atrous_rates(self.OS)


def atrous_rates(OS: int) -> Tuple[int, int, int]:
    """layers.py:118-126."""
    if OS == 8:
        return (12, 24, 36)
    if OS == 16:
        return (6, 12, 18)
    if OS == 32:
        return (3, 6, 9)
    raise ValueError('invalid output stride', OS)


# ------------------------------------------------------------------------------------------
# weight inventory: Keras layer names / creation order / shapes
# ------------------------------------------------------------------------------------------

BN_VARS = ('gamma', 'beta', 'moving_mean', 'moving_variance')  # Keras BatchNormalization.weights order


def weight_specs(cfg: HeadConfig) -> List[Tuple[str, str, Tuple[int, ...]]]:
    """(layer, var, shape) in the order the reference creates the layers."""
    specs: List[Tuple[str, str, Tuple[int, ...]]] = []

    def conv(name, k, n, bias=False):
        specs.append((name, 'kernel', (1, 1, k, n)))
        if bias:
            specs.append((name, 'bias', (n,)))

    def bn(name, c):
        for v in BN_VARS:
            specs.append((name, v, (c,)))

    def sep(prefix, c, n):
        specs.append((prefix + '_depthwise', 'depthwise_kernel', (3, 3, c, 1)))
        bn(prefix + '_depthwise_BN', c)
        conv(prefix + '_pointwise', c, n)
        bn(prefix + '_pointwise_BN', n)

    conv('image_pooling', cfg.Cin, 256)          # layers.py:134
    bn('image_pooling_BN', 256)                  # :136
    conv('aspp0', cfg.Cin, 256)                  # :141
    bn('aspp0_BN', 256)                          # :142
    if not cfg.lite:
        for i in (1, 2, 3):                      # :146-153
            sep('aspp%d' % i, cfg.Cin, 256)
    conv('concat_projection', 512 if cfg.lite else 1280, 256)   # :157 / :190
    bn('concat_projection_BN', 256)
    if cfg.decoder:
        conv('feature_projection0', cfg.Cskip, 48)   # :209
        bn('feature_projection0_BN', 48)             # :211
        sep('decoder_conv0', 304, 256)               # :215
        sep('decoder_conv1', 256, 256)               # :217
    conv('conv_upsample', 256, cfg.NC, bias=True)    # model.py:75
    return specs


def make_weights(cfg: HeadConfig, seed: int = 1234) -> Dict[Tuple[str, str], np.ndarray]:
    """Seeded random weights in Keras layout (SURVEY.md §8(c) 'golden vectors' recipe)."""
    rng = np.random.default_rng(seed)
    out: Dict[Tuple[str, str], np.ndarray] = {}
    for layer, var, shape in weight_specs(cfg):
        if var == 'kernel':
            fan_in = shape[2]
            a = rng.normal(0.0, math.sqrt(2.0 / fan_in), size=shape)
        elif var == 'depthwise_kernel':
            a = rng.normal(0.0, 0.3, size=shape)
        elif var == 'bias':
            a = rng.normal(0.0, 0.1, size=shape)
        elif var == 'gamma':
            a = rng.uniform(0.5, 1.5, size=shape)
        elif var in ('beta', 'moving_mean'):
            a = rng.normal(0.0, 0.1, size=shape)
        elif var == 'moving_variance':
            a = rng.uniform(0.5, 1.5, size=shape)
        else:  # pragma: no cover
            raise AssertionError(var)
        out[(layer, var)] = a.astype(np.float32)
    return out


def make_inputs(cfg: HeadConfig, seed: int, relu_feat: bool = True) -> Tuple[np.ndarray, Optional[np.ndarray]]:
    """Synthetic backbone features: feat ~ max(N(0,1),0) for Xception-like
    backbones (ends in ReLU), skip ~ N(0,1)."""
    rng = np.random.default_rng(seed)
    feat = rng.standard_normal((cfg.B, cfg.h, cfg.w, cfg.Cin), dtype=np.float32)
    if relu_feat:
        feat = np.maximum(feat, 0.0)
    skip = None
    if cfg.decoder:
        skip = rng.standard_normal((cfg.B, cfg.hs, cfg.ws, cfg.Cskip), dtype=np.float32)
    return feat, skip


# ------------------------------------------------------------------------------------------
# bfloat16 rounding (round-to-nearest-even), the CUDA path's __float2bfloat16_rn
# ------------------------------------------------------------------------------------------


def bf16_round(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    lsb = (u >> 16) & 1
    r = ((u + 0x7FFF + lsb) >> 16) << 16
    out = (r & 0xFFFFFFFF).astype(np.uint32).view(np.float32)
    nan = np.isnan(x)
    if nan.any():
        out = np.where(nan, x, out)
    return out.reshape(x.shape)


def to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """float32 -> uint16 bf16 bit patterns (RNE)."""
    r = bf16_round(x)
    return (r.view(np.uint32) >> 16).astype(np.uint16)


def from_bf16_bits(b: np.ndarray) -> np.ndarray:
    return (b.astype(np.uint32) << 16).view(np.float32)


# ------------------------------------------------------------------------------------------
# ops — numpy, written from the TF op semantics
# ------------------------------------------------------------------------------------------


def conv1x1(x: np.ndarray, kernel: np.ndarray, bias: Optional[np.ndarray] = None) -> np.ndarray:
    """Conv2D 1x1 'same': y[b,i,j,n] = sum_k x[b,i,j,k] * W[0,0,k,n]  (layers.py:14-21)."""
    k, n = kernel.shape[2], kernel.shape[3]
    y = x.reshape(-1, k).astype(np.float32) @ kernel.reshape(k, n).astype(np.float32)
    if bias is not None:
        y = y + bias.astype(np.float32)
    return y.reshape(x.shape[:-1] + (n,)).astype(np.float32)


def depthwise3x3(x: np.ndarray, dwk: np.ndarray, rate: int) -> np.ndarray:
    """DepthwiseConv2D 3x3, dilation `rate`, 'same', stride 1, no bias (layers.py:100-101):
    zero pad `rate` on all four sides; cross-correlation (no kernel flip)."""
    B, H, W, C = x.shape
    xp = np.zeros((B, H + 2 * rate, W + 2 * rate, C), np.float32)
    xp[:, rate:rate + H, rate:rate + W, :] = x
    y = np.zeros((B, H, W, C), np.float32)
    for u in range(3):
        for v in range(3):
            y += xp[:, u * rate:u * rate + H, v * rate:v * rate + W, :] * dwk[u, v, :, 0].astype(np.float32)
    return y


def bn_inference(x: np.ndarray, gamma, beta, mean, var, eps: float) -> np.ndarray:
    """inv = gamma*rsqrt(var+eps); y = x*inv + (beta - mean*inv)   (layers.py:63-70)."""
    inv = (gamma.astype(np.float32) / np.sqrt(var.astype(np.float32) + np.float32(eps))).astype(np.float32)
    return (x * inv + (beta.astype(np.float32) - mean.astype(np.float32) * inv)).astype(np.float32)


def bn_fold(gamma, beta, mean, var, eps: float) -> Tuple[np.ndarray, np.ndarray]:
    inv = (gamma.astype(np.float32) / np.sqrt(var.astype(np.float32) + np.float32(eps))).astype(np.float32)
    return inv, (beta.astype(np.float32) - mean.astype(np.float32) * inv).astype(np.float32)


def relu(x: np.ndarray) -> np.ndarray:
    return np.maximum(x, np.float32(0))


def global_avg_pool(x: np.ndarray) -> np.ndarray:
    """AveragePooling2D(pool_size=(h,w)) -> (B,1,1,C)  (layers.py:132)."""
    return x.astype(np.float32).mean(axis=(1, 2), keepdims=True, dtype=np.float32)


def _resize_coords(n_in: int, n_out: int):
    """tf.image.resize bilinear, TF2 (half_pixel_centers=True, no antialias)."""
    scale = np.float32(n_in) / np.float32(n_out)
    dst = np.arange(n_out, dtype=np.float32)
    src = (dst + np.float32(0.5)) * scale - np.float32(0.5)
    fl = np.floor(src)
    lo = np.maximum(fl, 0).astype(np.int64)
    hi = np.minimum(np.ceil(src), n_in - 1).astype(np.int64)
    t = (src - fl).astype(np.float32)
    return lo, hi, t


def resize_bilinear(x: np.ndarray, size: Tuple[int, int]) -> np.ndarray:
    """img_resize(mode='bilinear') (layers.py:48-50): lerp order top/bottom then vertical."""
    ho, wo = size
    B, hi, wi, C = x.shape
    x = x.astype(np.float32)
    ylo, yhi, ty = _resize_coords(hi, ho)
    xlo, xhi, tx = _resize_coords(wi, wo)
    tx = tx[None, None, :, None]
    ty = ty[None, :, None, None]
    top_rows = x[:, ylo]
    bot_rows = x[:, yhi]
    tl = top_rows[:, :, xlo]
    tr = top_rows[:, :, xhi]
    bl = bot_rows[:, :, xlo]
    br = bot_rows[:, :, xhi]
    top = tl + (tr - tl) * tx
    bot = bl + (br - bl) * tx
    return (top + (bot - top) * ty).astype(np.float32)


def softmax(x: np.ndarray) -> np.ndarray:
    m = x.max(axis=-1, keepdims=True)
    e = np.exp(x - m, dtype=np.float32)
    return (e / e.sum(axis=-1, keepdims=True, dtype=np.float32)).astype(np.float32)


def argmax_labels(x: np.ndarray) -> np.ndarray:
    """np.argmax(prediction, -1) — first maximum wins (deeplab.py:99; deeplabSegment.cpp:160-167)."""
    return np.argmax(x, axis=-1)


# ------------------------------------------------------------------------------------------
# the head
# ------------------------------------------------------------------------------------------


class _Mode:
    def __init__(self, mode: str):
        if mode not in ('fp32', 'bf16'):
            raise ValueError(mode)
        self.bf16 = mode == 'bf16'

    def act(self, x):      # rounding point for an activation written to HBM by the CUDA path
        return bf16_round(x) if self.bf16 else x

    def wt(self, w):       # rounding point for a GEMM weight
        return bf16_round(w) if self.bf16 else w


def _bn(W, name, x, eps):
    return bn_inference(x, W[(name, 'gamma')], W[(name, 'beta')], W[(name, 'moving_mean')],
                        W[(name, 'moving_variance')], eps)


def sepconv_bn(x, W, prefix, rate, eps, m: _Mode, taps: Optional[dict] = None):
    """SepConv_BN(stride=1, depth_activation=True) (layers.py:74-111)."""
    y = depthwise3x3(x, W[(prefix + '_depthwise', 'depthwise_kernel')], rate)
    y = m.act(relu(_bn(W, prefix + '_depthwise_BN', y, eps)))
    if taps is not None:
        taps[prefix + '_depthwise'] = y
    y = conv1x1(y, m.wt(W[(prefix + '_pointwise', 'kernel')]))
    y = m.act(relu(_bn(W, prefix + '_pointwise_BN', y, eps)))
    if taps is not None:
        taps[prefix + '_pointwise'] = y
    return y


def aspp_block(x, W, cfg: HeadConfig, m: _Mode, taps: dict):
    """ASPP_block / ASPP_Lite_block (layers.py:114-196). concat order [b4,b0,b1,b2,b3] (:155)."""
    eps = cfg.eps
    b4 = global_avg_pool(x)                                                    # :132
    b4 = conv1x1(b4, m.wt(W[('image_pooling', 'kernel')]))                     # :134
    b4 = m.act(relu(_bn(W, 'image_pooling_BN', b4, eps)))                      # :136-137
    taps['image_pooling'] = b4.reshape(cfg.B, 256)
    b4 = resize_bilinear(b4, (cfg.h, cfg.w))                                   # :138 (1x1 -> broadcast)
    b0 = conv1x1(x, m.wt(W[('aspp0', 'kernel')]))                              # :141
    b0 = m.act(relu(_bn(W, 'aspp0_BN', b0, eps)))                              # :142-143
    taps['aspp0'] = b0
    branches = [b4, b0]
    if not cfg.lite:
        for i, r in enumerate(cfg.rates):                                      # :146-153
            branches.append(sepconv_bn(x, W, 'aspp%d' % (i + 1), r, eps, m, taps))
    y = np.concatenate(branches, axis=-1)                                      # :155 / :189
    y = conv1x1(y, m.wt(W[('concat_projection', 'kernel')]))                   # :157
    y = m.act(relu(_bn(W, 'concat_projection_BN', y, eps)))                    # :159-160 ; Dropout = identity
    taps['aspp_out'] = y
    return y


def decoder_block(x, skip, W, cfg: HeadConfig, m: _Mode, taps: dict):
    """Decoder_block (layers.py:199-219). concat order [upsampled x (256), projected skip (48)] (:214)."""
    eps = cfg.eps
    x = m.act(resize_bilinear(x, (cfg.hs, cfg.ws)))                            # :207
    s = conv1x1(skip, m.wt(W[('feature_projection0', 'kernel')]))              # :209
    s = m.act(relu(_bn(W, 'feature_projection0_BN', s, eps)))                  # :211-213
    y = np.concatenate([x, s], axis=-1)                                        # :214
    taps['decoder_in'] = y
    y = sepconv_bn(y, W, 'decoder_conv0', 1, eps, m, taps)                     # :215
    taps['decoder_conv0'] = y
    y = sepconv_bn(y, W, 'decoder_conv1', 1, eps, m, taps)                     # :217
    taps['decoder_out'] = y
    return y


def head_forward(feat: np.ndarray, skip: Optional[np.ndarray], W: Dict[Tuple[str, str], np.ndarray],
                 cfg: HeadConfig, mode: str = 'fp32', want_softmax: bool = False) -> dict:
    """Whole hot path: ASPP(-Lite) -> [Decoder] -> conv_upsample -> pred_resize -> argmax.
    Returns a dict of intermediates keyed by reference layer names plus 'logits' (low-res,
    NHWC), 'logits_full', 'labels' (argmax of the resized LOGITS — see SURVEY §8(c): softmax
    can create fp32 ties the logits do not have) and optionally 'softmax'."""
    m = _Mode(mode)
    taps: dict = {}
    x = m.act(feat.astype(np.float32))
    if skip is not None:
        skip = m.act(skip.astype(np.float32))
    y = aspp_block(x, W, cfg, m, taps)
    if cfg.decoder:
        y = decoder_block(y, skip, W, cfg, m, taps)
    logits = conv1x1(y, m.wt(W[('conv_upsample', 'kernel')]), W[('conv_upsample', 'bias')])   # model.py:75
    taps['logits'] = logits
    full = resize_bilinear(logits, (cfg.H, cfg.W))                                            # model.py:76
    taps['logits_full'] = full
    taps['labels'] = argmax_labels(full).astype(np.uint8 if cfg.NC <= 256 else np.int64)
    if want_softmax:
        taps['softmax'] = softmax(full)                                                       # model.py:86
    return taps


# ------------------------------------------------------------------------------------------
# torch-CPU backend: same graph through torch.nn.functional (an independent implementation of
# the same ops) — used for the large cases and as the timed CPU baseline.
# ------------------------------------------------------------------------------------------


def head_forward_torch(feat, skip, W, cfg: HeadConfig, mode: str = 'fp32', want_labels: bool = True,
                       keep: bool = False):
    """feat/skip: numpy or torch NHWC float32. Returns dict with torch tensors (NHWC)."""
    import torch
    import torch.nn.functional as F

    bf = mode == 'bf16'

    def act(t):
        return t.bfloat16().float() if bf else t

    def wt(a):
        t = torch.from_numpy(np.ascontiguousarray(a))
        return t.bfloat16().float() if bf else t

    def T(a):
        return torch.from_numpy(np.ascontiguousarray(a))

    def to_nchw(a):
        t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))
        return t.float().permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last)

    def conv(x, name, bias=False):
        k = wt(W[(name, 'kernel')]).permute(3, 2, 0, 1).contiguous()
        b = T(W[(name, 'bias')]) if bias else None
        return F.conv2d(x, k, b)

    def bn(x, name):
        s, t = bn_fold(W[(name, 'gamma')], W[(name, 'beta')], W[(name, 'moving_mean')],
                       W[(name, 'moving_variance')], cfg.eps)
        return x * T(s).view(1, -1, 1, 1) + T(t).view(1, -1, 1, 1)

    def sep(x, prefix, rate):
        k = T(W[(prefix + '_depthwise', 'depthwise_kernel')]).permute(2, 3, 0, 1).contiguous()
        y = F.conv2d(x, k, None, stride=1, padding=rate, dilation=rate, groups=x.shape[1])
        y = act(F.relu(bn(y, prefix + '_depthwise_BN')))
        y = conv(y, prefix + '_pointwise')
        return act(F.relu(bn(y, prefix + '_pointwise_BN')))

    out = {}
    with torch.no_grad():
        x = act(to_nchw(feat))
        b4 = x.mean(dim=(2, 3), keepdim=True)
        b4 = act(F.relu(bn(conv(b4, 'image_pooling'), 'image_pooling_BN')))
        b4 = b4.expand(-1, -1, cfg.h, cfg.w)
        b0 = act(F.relu(bn(conv(x, 'aspp0'), 'aspp0_BN')))
        br = [b4, b0]
        if not cfg.lite:
            for i, r in enumerate(cfg.rates):
                br.append(sep(x, 'aspp%d' % (i + 1), r))
        y = torch.cat(br, dim=1)
        y = act(F.relu(bn(conv(y, 'concat_projection'), 'concat_projection_BN')))
        if keep:
            out['aspp_out'] = y.permute(0, 2, 3, 1)
        if cfg.decoder:
            s = act(to_nchw(skip))
            y = act(F.interpolate(y, size=(cfg.hs, cfg.ws), mode='bilinear', align_corners=False))
            s = act(F.relu(bn(conv(s, 'feature_projection0'), 'feature_projection0_BN')))
            y = torch.cat([y, s], dim=1)
            y = sep(y, 'decoder_conv0', 1)
            y = sep(y, 'decoder_conv1', 1)
            if keep:
                out['decoder_out'] = y.permute(0, 2, 3, 1)
        logits = conv(y, 'conv_upsample', bias=True)
        out['logits'] = logits.permute(0, 2, 3, 1)
        full = F.interpolate(logits, size=(cfg.H, cfg.W), mode='bilinear', align_corners=False)
        out['logits_full'] = full.permute(0, 2, 3, 1)
        if want_labels:
            # reference: Softmax (model.py:86) then np.argmax on host (deeplab.py:99)
            prob = torch.softmax(full, dim=1)
            out['labels_softmax'] = prob.argmax(dim=1)
            out['labels'] = full.argmax(dim=1)
    return out


# ------------------------------------------------------------------------------------------
# metrics restated from the reference (used by the example-image parity test)
# ------------------------------------------------------------------------------------------


def mIOU(gt: np.ndarray, preds: np.ndarray) -> float:
    """deeplabv3p/metrics.py:10-17."""
    ulabels = np.unique(gt)
    iou = np.zeros(len(ulabels))
    for k, u in enumerate(ulabels):
        inter = (gt == u) & (preds == u)
        union = (gt == u) | (preds == u)
        iou[k] = inter.sum() / union.sum()
    return float(np.round(iou.mean(), 2))


def generate_matrix(gt_mask: np.ndarray, pre_mask: np.ndarray, num_classes: int) -> np.ndarray:
    """eval.py:368-373."""
    valid = (gt_mask >= 0) & (gt_mask < num_classes)
    label = num_classes * gt_mask[valid].astype('int') + pre_mask[valid]
    count = np.bincount(label, minlength=num_classes ** 2)
    return count.reshape(num_classes, num_classes)


def dataset_mIOU(confusion: np.ndarray) -> float:
    """eval.py:461-470: mean over classes of TP / (TP+FP+FN), nan-safe."""
    with np.errstate(divide='ignore', invalid='ignore'):
        iou = np.diag(confusion) / (confusion.sum(0) + confusion.sum(1) - np.diag(confusion))
    return float(np.nanmean(iou))


def normalize_image(image: np.ndarray) -> np.ndarray:
    """common/data_utils.py:403-417."""
    return image.astype(np.float32) / 127.5 - 1


# ------------------------------------------------------------------------------------------
# stand-in backbone for the example-image fixture: NOT part of the reference.
# strided average pooling + a fixed random 1x1 projection to the channel counts a real backbone
# would produce, so that example JPEGs become (feat, skip) pairs with image structure in them.
# ------------------------------------------------------------------------------------------


def standin_backbone(image_nhwc: np.ndarray, cfg: HeadConfig, seed: int = 99):
    rng = np.random.default_rng(seed)
    B, H, W, _ = image_nhwc.shape

    def pool(x, s, oh, ow):
        ph, pw = oh * s - x.shape[1], ow * s - x.shape[2]
        if ph or pw:
            x = np.pad(x, ((0, 0), (0, ph), (0, pw), (0, 0)), mode='edge')
        return x.reshape(B, oh, s, ow, s, -1).mean(axis=(2, 4))

    def lift(x, c, relu_out):
        # random features of [r,g,b, local contrast terms]
        k = rng.standard_normal((x.shape[-1], c)).astype(np.float32)
        y = x @ k + rng.standard_normal(c).astype(np.float32) * 0.1
        return np.maximum(y, 0) if relu_out else y

    f = pool(image_nhwc, cfg.OS, cfg.h, cfg.w)
    s = pool(image_nhwc, 4, cfg.hs, cfg.ws)
    # add simple gradient features so the maps are not rank-3
    def grads(x):
        gx = np.zeros_like(x); gy = np.zeros_like(x)
        gx[:, :, 1:] = x[:, :, 1:] - x[:, :, :-1]
        gy[:, 1:] = x[:, 1:] - x[:, :-1]
        return np.concatenate([x, gx, gy, x * x], axis=-1)
    feat = lift(grads(f), cfg.Cin, True).astype(np.float32)
    skip = lift(grads(s), cfg.Cskip, False).astype(np.float32) if cfg.decoder else None
    return feat, skip


# ---------------------------------------------------------------------------------------------------------------------
# Training-mode (Sync)BatchNormalization, layers.py:63-70 with train.py:143-158 (SURVEY.md §8(c) "SyncBN training"):
# per replica sum_x, sum_x2 over (N,H,W), count; all-reduce SUM; mean = sum/n, var = sum2/n - mean^2 (biased).
def bn_train_stats(x):
    """x [..., C] -> float64 [2C+1] = sum_x | sum_x2 | rows of ONE replica."""
    C = x.shape[-1]
    x2 = np.asarray(x, np.float64).reshape(-1, C)
    return np.concatenate([x2.sum(0), (x2 * x2).sum(0), [float(x2.shape[0])]])


def sync_bn_train(x_shards, gamma, beta, eps=1e-5, relu=True):
    """List of per-replica tensors [..., C] -> (list of normalised tensors, mean, biased variance) with GLOBAL statistics."""
    C = x_shards[0].shape[-1]
    st = sum(bn_train_stats(s) for s in x_shards)
    mean = st[:C] / st[2 * C]
    var = st[C:2 * C] / st[2 * C] - mean * mean
    inv = np.asarray(gamma, np.float64) / np.sqrt(var + eps)
    outs = []
    for s in x_shards:
        y = (np.asarray(s, np.float64) - mean) * inv + np.asarray(beta, np.float64)
        outs.append(np.maximum(y, 0.0) if relu else y)
    return outs, mean, var
