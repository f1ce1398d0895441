"""
oracle/train_ref.py — CPU restatement of ONE training step of the reference's DeepLabV3+ head.  TEST INFRASTRUCTURE ONLY
(imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs; never by the product path).

PARITY UNPINNED against TensorFlow (not installable here; the reference ships no training fixtures): the step is restated
with torch.nn.functional ops in fp32 and differentiated by torch.autograd — an independent implementation of the same
published semantics — and the restatement itself is checked in tests/test_train_oracle.py against central finite
differences and against a hand-written SyncBN / SGD computation.

Reference sites restated (paths relative to /root/reference):
  graph in training mode     deeplabv3p/models/layers.py:74-219 (SepConv_BN, ASPP_block, Decoder_block), model.py:75-86
  batch normalisation        layers.py:63-70  (batch statistics, biased variance, eps 1e-5; moving <- 0.99 moving + 0.01 batch;
                             SyncBatchNormalization under MirroredStrategy = statistics of the GLOBAL batch)
  Dropout(0.5)               layers.py:161 (after concat_projection; the mask is an input here)
  loss                       deeplabv3p/loss.py:121-156 SparseCategoricalCrossEntropy(ignore_index=255) on the Softmax output:
                             under model.fit (graph mode) K.categorical_crossentropy sees a Softmax op and takes its logits path
                             (softmax_cross_entropy_with_logits, no clip): -log p; pixels with the ignore label contribute 0;
                             Keras averages over ALL pixels of the global batch (train.py:143-158: per-replica sums / global batch)
  regulariser                layers.py:12-21 l2(2e-5) on Conv2D kernels and biases (inert for depthwise kernels, :24-31)
  optimizer                  common/model_utils.py:122-123 SGD(momentum=0.9, nesterov=False): v <- m v - lr g ; w <- w + v
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np

from . import head_ref as R

L2_COEF = 2e-5
BN_MOMENTUM = 0.99

Key = Tuple[str, str]


def _trainable(key: Key) -> bool:
    return key[1] in ('kernel', 'bias', 'depthwise_kernel', 'gamma', 'beta')


def head_train_forward_backward(feat: np.ndarray, skip: np.ndarray, labels: np.ndarray, W: Dict[Key, np.ndarray], cfg: R.HeadConfig,
                                keep_mask: Optional[np.ndarray] = None, drop_rate: float = 0.5, ignore_index: int = 255,
                                dtype=None, mode: str = 'fp32', loss: str = 'crossentropy', class_weights=None, focal_gamma: float = 2.0,
                                focal_alpha: float = 0.25):
    """feat [B,h,w,Cin], skip [B,hs,ws,Cs] fp32 NHWC (the GLOBAL batch), labels uint8 [B,H,W].  keep_mask: bool [B*h*w*256]
    Dropout keep mask in NHWC element order (None = no dropout).  Returns dict with
      loss (float, mean over every pixel of the batch, without the l2 term), grads {(layer,var): array in Keras layout},
      d_feat, d_skip, batch_stats {bn layer: (mean, biased var)}, logits [B,hs,ws,NC].
    mode 'fp32': reference semantics.  mode 'bf16': the same graph with the CUDA path's rounding points — every activation
    written to HBM and every activation GRADIENT written to HBM rounded to bfloat16 (a rounding node whose backward rounds the
    gradient), GEMM weights rounded to bf16 (straight-through: weight gradients stay fp32), fp32 logits / loss / statistics.
    The step is ill-conditioned at random initialisation (rounding the GEMM weights alone moves deep gradients by 5-10 %,
    tests/test_train_oracle.py), so kernel parity is asserted against this mode and the fp32 mode bounds the total drift."""
    import torch
    import torch.nn.functional as F
    dt = dtype or torch.float32
    if cfg.lite == cfg.decoder:
        raise ValueError('the training oracle covers the reference\'s two heads: ASPP_block + Decoder_block, or ASPP_Lite_block without a decoder')

    class _RoundBF16(torch.autograd.Function):
        @staticmethod
        def forward(ctx, t):
            return t.to(torch.bfloat16).to(t.dtype)

        @staticmethod
        def backward(ctx, g):
            return g.to(torch.bfloat16).to(g.dtype)

    if mode == 'bf16':
        rb = _RoundBF16.apply
        rw = lambda k: k + (k.detach().to(torch.bfloat16).to(k.dtype) - k.detach())      # rounded value, unrounded gradient
    else:
        rb = lambda t: t
        rw = lambda k: k
    P = {k: torch.tensor(np.asarray(v), dtype=dt, requires_grad=_trainable(k)) for k, v in W.items()}
    x_in = torch.tensor(feat, dtype=dt, requires_grad=True)
    s_in = torch.tensor(skip if skip is not None else np.zeros((cfg.B, 1, 1, 8), np.float32), dtype=dt, requires_grad=True)
    x = x_in.permute(0, 3, 1, 2)
    s = s_in.permute(0, 3, 1, 2)
    stats = {}
    acts = {}

    def keep(name, t):                                           # intermediate activation whose gradient the tests may inspect
        t.retain_grad()
        acts[name] = t
        return t

    def conv(t, name, bias=False):
        k = rw(P[(name, 'kernel')])                              # (1,1,K,N) HWIO
        y = F.conv2d(t, k.permute(3, 2, 0, 1), P[(name, 'bias')] if bias else None)
        return y if bias else rb(y)                              # the classifier's logits stay fp32

    def bn(t, name, relu=True):
        stats[name] = (t.detach().mean(dim=(0, 2, 3)).numpy().copy(), t.detach().var(dim=(0, 2, 3), unbiased=False).numpy().copy())
        y = F.batch_norm(t, None, None, P[(name, 'gamma')], P[(name, 'beta')], training=True, eps=cfg.eps)
        return rb(F.relu(y) if relu else y)

    def sep(t, prefix, rate):
        k = P[(prefix + '_depthwise', 'depthwise_kernel')]       # (3,3,C,1)
        Cc = k.shape[2]
        t = keep(prefix + '/d', rb(F.conv2d(rb(t), k.permute(2, 3, 0, 1), None, padding=rate, dilation=rate, groups=Cc)))
        t = keep(prefix + '/a', bn(t, prefix + '_depthwise_BN'))
        t = keep(prefix + '/p', conv(t, prefix + '_pointwise'))
        return keep(prefix + '/y', bn(t, prefix + '_pointwise_BN'))

    b4 = bn(conv(rb(rb(x).mean(dim=(2, 3), keepdim=True)), 'image_pooling'), 'image_pooling_BN')
    b4 = b4.expand(-1, -1, cfg.h, cfg.w)                          # bilinear resize of a 1x1 map = broadcast
    b0 = bn(conv(rb(x), 'aspp0'), 'aspp0_BN')
    if cfg.lite:                                                   # ASPP_Lite_block (layers.py:166-196): only the two branches
        bs = []
    else:                                                          # ASPP_block (layers.py:114-163)
        bs = [sep(x, 'aspp%d' % i, cfg.rates[i - 1]) for i in (1, 2, 3)]
    y = bn(conv(rb(torch.cat([b4, b0] + bs, dim=1)), 'concat_projection'), 'concat_projection_BN')
    if keep_mask is not None:
        m = torch.tensor(np.asarray(keep_mask).reshape(cfg.B, cfg.h, cfg.w, 256), dtype=dt).permute(0, 3, 1, 2)
        y = rb(y * m * (1.0 / (1.0 - drop_rate)))
    if cfg.decoder:                                                # Decoder_block (layers.py:199-219)
        up = rb(F.interpolate(y, size=(cfg.hs, cfg.ws), mode='bilinear', align_corners=False))
        sk = bn(conv(rb(s), 'feature_projection0'), 'feature_projection0_BN')
        d = sep(torch.cat([up, sk], dim=1), 'decoder_conv0', 1)
        d = sep(d, 'decoder_conv1', 1)
    else:                                                          # the *_lite models: the classifier reads the ASPP output (deeplabv3p_mobilenetv2.py:326-331)
        d = y
    # tail + loss
    logits = conv(d, 'conv_upsample', bias=True)
    logits = keep('logits', rb(logits) + (logits - rb(logits)).detach() if mode == 'bf16' else logits)   # fp32 value, bf16 gradient
    full = F.interpolate(logits, size=(cfg.H, cfg.W), mode='bilinear', align_corners=False)
    prob = torch.softmax(full, dim=1)
    lab = torch.tensor(np.asarray(labels).astype(np.int64))
    valid = (lab != ignore_index) & (lab < cfg.NC)
    p_lab = prob.gather(1, lab.clamp(max=cfg.NC - 1).unsqueeze(1)).squeeze(1)
    if loss == 'focal':                                            # SparseSoftmaxFocalLoss, loss.py:60-118 (class weights ignored, train.py:131-135)
        pc = p_lab.clamp(1e-15, 1.0 - 1e-15)
        px = focal_alpha * torch.pow(1.0 - pc, focal_gamma) * (-torch.log(pc)) * valid.to(dt)
    elif class_weights is not None:                                # WeightedSparseCategoricalCrossEntropy, loss.py:159-192 (no clip)
        wv = torch.tensor(np.asarray(class_weights), dtype=dt)[lab.clamp(max=cfg.NC - 1)]
        px = -torch.log(p_lab) * wv * valid.to(dt)
    else:                                                          # SparseCategoricalCrossEntropy, loss.py:121-156
        px = -torch.log(p_lab.clamp_min(1e-37)) * valid.to(dt)
    loss = px.sum() / float(cfg.B * cfg.H * cfg.W)
    loss.backward()
    grads = {k: v.grad.numpy().copy() for k, v in P.items() if v.requires_grad and v.grad is not None}
    return {'loss': float(loss.item()), 'grads': grads, 'd_feat': x_in.grad.numpy().copy(), 'd_skip': None if s_in.grad is None else s_in.grad.numpy().copy(),
            'act_grads': {k: v.grad.permute(0, 2, 3, 1).numpy().copy() for k, v in acts.items() if v.grad is not None},
            'batch_stats': stats, 'logits': logits.detach().permute(0, 2, 3, 1).numpy().copy(), 'valid_pixels': int(valid.sum().item())}


def l2_grad(key: Key, w: np.ndarray, l2: float = L2_COEF) -> np.ndarray:
    """Gradient of the regulariser: l2 * sum(w^2) on Conv2D kernels and biases only (layers.py:12-31)."""
    return 2.0 * l2 * w if key[1] in ('kernel', 'bias') else np.zeros_like(w)


def sgd_momentum_update(W: Dict[Key, np.ndarray], grads: Dict[Key, np.ndarray], velocity: Dict[Key, np.ndarray], lr: float = 1e-2,
                        momentum: float = 0.9, l2: float = L2_COEF):
    """Keras SGD(momentum): v <- m v - lr (g + l2 gradient); w <- w + v.  Returns (new weights, new velocity)."""
    Wn, Vn = dict(W), dict(velocity)
    for k, g in grads.items():
        w = np.asarray(W[k], np.float32)
        gt = np.asarray(g, np.float32).reshape(w.shape) + l2_grad(k, w, l2).astype(np.float32)
        v = momentum * np.asarray(velocity.get(k, np.zeros_like(w)), np.float32) - lr * gt
        Vn[k] = v.astype(np.float32)
        Wn[k] = (w + v).astype(np.float32)
    return Wn, Vn


def moving_update(W: Dict[Key, np.ndarray], batch_stats, momentum: float = BN_MOMENTUM):
    Wn = dict(W)
    for name, (mean, var) in batch_stats.items():
        Wn[(name, 'moving_mean')] = (np.asarray(W[(name, 'moving_mean')], np.float32) * momentum + mean * (1 - momentum)).astype(np.float32)
        Wn[(name, 'moving_variance')] = (np.asarray(W[(name, 'moving_variance')], np.float32) * momentum + var * (1 - momentum)).astype(np.float32)
    return Wn


def make_labels(cfg: R.HeadConfig, seed: int, ignore_frac: float = 0.05, ignore_index: int = 255) -> np.ndarray:
    """SURVEY §8(d) cfg 5: labels ~ U{0..NC-1} with 5 % = 255."""
    rng = np.random.default_rng(seed)
    lab = rng.integers(0, cfg.NC, size=(cfg.B, cfg.H, cfg.W)).astype(np.uint8)
    lab[rng.random(lab.shape) < ignore_frac] = ignore_index
    return lab
