"""
oracle/xception_ref.py — CPU restatement of the reference's modified aligned Xception feature extractor (the
next row after the head).  TEST INFRASTRUCTURE ONLY: the checker of the CUDA backbone (csrc/xception_api.cu); nothing in the product
path imports it.

Reference sites restated (paths relative to /root/reference):
  Xception_body       deeplabv3p/models/deeplabv3p_xception.py:95-163  (entry flow, 16 middle-flow units, exit flow; strides / atrous
                      rates per output stride :100-117)
  _xception_block     :55-92   (three SepConv_BN, stride on the third, 'conv' / 'sum' / 'none' shortcut; skip = output of the second)
  _conv2d_same        :25-52   (stride 1: 'same'; stride 2: explicit ZeroPadding2D(pad_beg, pad_end) + 'valid')
  SepConv_BN          deeplabv3p/models/layers.py:74-111 (backbone use: depth_activation False -> ReLU BEFORE the depthwise conv, no
                      ReLU after the BNs; epsilon 1e-3; stride 2 = explicit padding + 'valid')
PINNED by the reference's own published figures (README.md:309): 41.06 M parameters and 102.73 GFLOPs for DeepLabV3+ Xception
512x512 OS16 with 21 classes — tests/test_xception_oracle.py holds the inventory below (plus the head's 3 205 701 parameters) to them.
The arithmetic (TensorFlow ops) is unpinned, as for the head.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np

BN_EPS_BACKBONE = 1e-3        # Keras BatchNormalization default, kept by CustomBatchNormalization (layers.py:63-70) and SepConv_BN's default
BN_VARS = ('gamma', 'beta', 'moving_mean', 'moving_variance')
Key = Tuple[str, str]


def os_plan(OS: int) -> Dict[str, int]:
    """deeplabv3p_xception.py:100-117."""
    if OS == 8:
        return dict(os16_stride=1, os16_rate=2, os32_stride=1, os32_rate=4)
    if OS == 16:
        return dict(os16_stride=2, os16_rate=1, os32_stride=1, os32_rate=2)
    if OS == 32:
        return dict(os16_stride=2, os16_rate=1, os32_stride=2, os32_rate=1)
    raise ValueError('invalid output stride', OS)


def blocks(OS: int) -> List[dict]:
    """The 21 _xception_block calls of Xception_body in order (:131-153): 3 entry-flow, 16 middle-flow, 2 exit-flow."""
    p = os_plan(OS)
    out = [dict(prefix='entry_flow_block1', cin=64, depth=[128, 128, 128], shortcut='conv', stride=2, rate=1, act=False),
           dict(prefix='entry_flow_block2', cin=128, depth=[256, 256, 256], shortcut='conv', stride=2, rate=1, act=False, return_skip=True),
           dict(prefix='entry_flow_block3', cin=256, depth=[728, 728, 728], shortcut='conv', stride=p['os16_stride'], rate=1, act=False)]
    for i in range(16):
        out.append(dict(prefix='middle_flow_unit_%d' % (i + 1), cin=728, depth=[728, 728, 728], shortcut='sum', stride=1, rate=p['os16_rate'], act=False))
    out.append(dict(prefix='exit_flow_block1', cin=728, depth=[728, 1024, 1024], shortcut='conv', stride=p['os32_stride'], rate=p['os16_rate'], act=False))
    out.append(dict(prefix='exit_flow_block2', cin=1024, depth=[1536, 1536, 2048], shortcut='none', stride=1, rate=p['os32_rate'], act=True))
    return out


def weight_specs(OS: int = 16) -> List[Tuple[str, str, Tuple[int, ...]]]:
    """(layer, variable, shape) in Keras creation order (Keras HWIO / (3,3,C,1) layouts)."""
    specs: List[Tuple[str, str, Tuple[int, ...]]] = []

    def bn(name, c):
        for v in BN_VARS:
            specs.append((name, v, (c,)))

    specs.append(('entry_flow_conv1_1', 'kernel', (3, 3, 3, 32)))
    bn('entry_flow_conv1_1_BN', 32)
    specs.append(('entry_flow_conv1_2', 'kernel', (3, 3, 32, 64)))
    bn('entry_flow_conv1_2_BN', 64)
    for b in blocks(OS):
        c = b['cin']
        for i, d in enumerate(b['depth']):
            p = '%s_separable_conv%d' % (b['prefix'], i + 1)
            specs.append((p + '_depthwise', 'depthwise_kernel', (3, 3, c, 1)))
            bn(p + '_depthwise_BN', c)
            specs.append((p + '_pointwise', 'kernel', (1, 1, c, d)))
            bn(p + '_pointwise_BN', d)
            c = d
        if b['shortcut'] == 'conv':
            specs.append((b['prefix'] + '_shortcut', 'kernel', (1, 1, b['cin'], b['depth'][-1])))
            bn(b['prefix'] + '_shortcut_BN', b['depth'][-1])
    return specs


def param_count(OS: int = 16) -> int:
    return int(sum(int(np.prod(s)) for _, _, s in weight_specs(OS)))


def make_weights(OS: int = 16, seed: int = 4321) -> Dict[Key, np.ndarray]:
    rng = np.random.default_rng(seed)
    W: Dict[Key, np.ndarray] = {}
    for layer, var, shape in weight_specs(OS):
        if var == 'kernel':
            fan_in = shape[0] * shape[1] * shape[2]
            a = rng.normal(0.0, np.sqrt(2.0 / fan_in), shape)
        elif var == 'depthwise_kernel':
            a = rng.normal(0.0, 0.3, shape)
        elif var in ('gamma', 'moving_variance'):
            a = rng.uniform(0.5, 1.5, shape)
        else:
            a = rng.normal(0.0, 0.1, shape)
        W[(layer, var)] = a.astype(np.float32)
    return W


def make_calibrated_weights(OS: int = 16, seed: int = 4321, size: int = 96, residual_gamma: float = 0.1) -> Dict[Key, np.ndarray]:
    """Seeded random weights whose BatchNorm moving statistics match the activations they see (one fp32 pass over a seeded
    calibration image): with purely random moving statistics the 66-layer network's activations grow to ~1e4 and every parity
    number is dominated by overflow-scale rounding; a trained network has O(1) activations, which is the regime the tolerances of
    BASELINE.json are written for.  residual_gamma scales the gamma of the LAST BatchNorm of every block that has a shortcut (the
    usual zero-ish initialisation of residual branches): a random residual network with unit-gain branches amplifies any perturbation
    (measured with this oracle, bf16 mode against fp32 mode: 24 % relative L2 at the feature output with gain 1, 2.7 % with 0.1).
    Deterministic: same arguments -> same weights."""
    W = make_weights(OS, seed)
    for b in blocks(OS):
        if b['shortcut'] != 'none':
            k = (b['prefix'] + '_separable_conv3_pointwise_BN', 'gamma')
            W[k] = (W[k] * residual_gamma).astype(np.float32)
    rng = np.random.default_rng(seed + 1)
    img = rng.uniform(-1.0, 1.0, (2, size, size, 3)).astype(np.float32)
    forward_torch(img, W, OS, 'fp32', calibrate=True)
    return W


def _same_pad_stride2(kernel_size: int, rate: int) -> Tuple[int, int]:
    """_conv2d_same / SepConv_BN: kernel_size_effective - 1 split as (beg, end) (:44-48, layers.py:91-95)."""
    ke = kernel_size + (kernel_size - 1) * (rate - 1)
    total = ke - 1
    beg = total // 2
    return beg, total - beg


def depthwise3x3(x, k_hwc1, stride: int, pad: int, rate: int):
    """DepthwiseConv2D 3x3 (cross-correlation, depth multiplier 1) on an NCHW torch tensor as nine shifted multiply-adds
    (tap order (0,0), (0,1), ... like the convolution sum): torch's grouped conv2d on CPU takes ~0.1 s per layer at these shapes
    (8 s per image over the backbone's 63 depthwise layers), this runs at memory speed.  k_hwc1: Keras (3,3,C,1)."""
    import torch.nn.functional as F
    xp = F.pad(x, (pad, pad, pad, pad)) if pad else x
    Ho = (xp.shape[2] - (2 * rate + 1)) // stride + 1
    Wo = (xp.shape[3] - (2 * rate + 1)) // stride + 1
    out = None
    for u in range(3):
        for v in range(3):
            sl = xp[:, :, u * rate: u * rate + (Ho - 1) * stride + 1: stride, v * rate: v * rate + (Wo - 1) * stride + 1: stride]
            term = sl * k_hwc1[u, v, :, 0].view(1, -1, 1, 1)
            out = term if out is None else out + term
    return out


def forward_torch(image_nhwc: np.ndarray, W: Dict[Key, np.ndarray], OS: int = 16, mode: str = 'fp32', taps: Dict[str, np.ndarray] = None,
                  calibrate: bool = False):
    """Inference-mode forward (BN with moving statistics).  image [B,H,W,3] fp32 in [-1, 1] (normalize_image).
    Returns (feature [B,H/OS,W/OS,2048], skip [B,H/4,W/4,256]) as NHWC numpy arrays.
    mode 'fp32' = reference semantics; 'bf16' = the same graph with the CUDA path's rounding points: the first convolution runs in
    fp32 on the fp32 image, every tensor a kernel writes (stem outputs, depthwise+BN outputs, pointwise+BN(+shortcut) outputs) is
    rounded to bf16, pointwise / shortcut / conv1_2 kernels are bf16, accumulation and BN arithmetic fp32.
    taps: optional dict that receives named intermediates (NHWC numpy) for block-level parity tests.
    calibrate: OVERWRITES the moving statistics in W with this batch's statistics layer by layer (see make_calibrated_weights)."""
    import torch
    import torch.nn.functional as F
    bf = mode == 'bf16'
    T = lambda k: torch.from_numpy(np.asarray(W[k], np.float32))
    rb = (lambda t: t.bfloat16().float()) if bf else (lambda t: t)

    def keep(name, t):
        if taps is not None:
            taps[name] = t.permute(0, 2, 3, 1).numpy().copy()
        return t

    def bn(x, name):
        if calibrate:     # fixture generation: moving statistics := the statistics of this very batch (activations stay O(1), as in a trained net)
            W[(name, 'moving_mean')] = x.mean(dim=(0, 2, 3)).numpy().astype(np.float32)
            W[(name, 'moving_variance')] = np.maximum(x.var(dim=(0, 2, 3), unbiased=False).numpy(), 1e-3).astype(np.float32)
        return F.batch_norm(x, T((name, 'moving_mean')), T((name, 'moving_variance')), T((name, 'gamma')), T((name, 'beta')), False, 0.0, BN_EPS_BACKBONE)

    def conv_same(x, name, stride, ksize, rate=1, round_w=True):
        k = T((name, 'kernel')).permute(3, 2, 0, 1)
        if round_w:
            k = rb(k)
        if stride == 1:
            pad = ((ksize - 1) * rate) // 2
            return F.conv2d(x, k, None, 1, pad, rate)
        beg, end = _same_pad_stride2(ksize, rate)
        return F.conv2d(F.pad(x, (beg, end, beg, end)), k, None, stride, 0, rate)

    def sepconv_bn(x, prefix, stride, rate, act, residual=None):
        if stride != 1:
            beg, end = _same_pad_stride2(3, rate)
            x = F.pad(x, (beg, end, beg, end))
            pad = 0
        else:
            pad = rate
        if not act:
            x = F.relu(x)
        x = depthwise3x3(x, T((prefix + '_depthwise', 'depthwise_kernel')), stride, pad, rate)
        x = bn(x, prefix + '_depthwise_BN')
        if act:
            x = F.relu(x)
        x = rb(x)                                                   # the depthwise kernel writes bf16: the GEMM's A operand
        x = F.conv2d(x, rb(T((prefix + '_pointwise', 'kernel')).permute(3, 2, 0, 1)))
        x = bn(x, prefix + '_pointwise_BN')
        if act:
            x = F.relu(x)
        if residual is not None:                                    # add([residual, shortcut]) in the GEMM epilogue, fp32, before rounding
            x = x + residual
        return rb(x)

    x = torch.from_numpy(np.asarray(image_nhwc, np.float32)).permute(0, 3, 1, 2)
    # entry_flow_conv1_1: Conv2D(32, 3, strides 2, padding 'same') — TensorFlow 'same' at stride 2 pads (0, 1) for even sizes (:119-120)
    H, Wd = x.shape[2], x.shape[3]
    ph = max((-(-H // 2) - 1) * 2 + 3 - H, 0)
    pw = max((-(-Wd // 2) - 1) * 2 + 3 - Wd, 0)
    x = F.conv2d(F.pad(x, (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2)), T(('entry_flow_conv1_1', 'kernel')).permute(3, 2, 0, 1), None, 2)
    x = keep('entry_flow_conv1_1', rb(F.relu(bn(x, 'entry_flow_conv1_1_BN'))))
    x = keep('entry_flow_conv1_2', rb(F.relu(bn(conv_same(x, 'entry_flow_conv1_2', 1, 3), 'entry_flow_conv1_2_BN'))))
    skip = None
    for b in blocks(OS):
        inp = x
        if b['shortcut'] == 'conv':
            res = rb(bn(conv_same(inp, b['prefix'] + '_shortcut', b['stride'], 1), b['prefix'] + '_shortcut_BN'))
        elif b['shortcut'] == 'sum':
            res = inp
        else:
            res = None
        r = x
        for i in range(3):
            r = sepconv_bn(r, '%s_separable_conv%d' % (b['prefix'], i + 1), b['stride'] if i == 2 else 1, b['rate'], b['act'], res if i == 2 else None)
            if i == 1 and b.get('return_skip'):
                skip = r
        x = keep(b['prefix'], r)
    return x.permute(0, 2, 3, 1).numpy().copy(), skip.permute(0, 2, 3, 1).numpy().copy()


def conv_flops(H: int, Wd: int, OS: int = 16) -> float:
    """Multiply-accumulate count x 2 of every convolution of the backbone at input H x W (what the reference's FLOPs table counts)."""
    total = 0.0
    h, w = -(-H // 2), -(-Wd // 2)
    total += 2.0 * h * w * 3 * 3 * 3 * 32
    total += 2.0 * h * w * 3 * 3 * 32 * 64
    for b in blocks(OS):
        c = b['cin']
        hin, win = h, w
        for i, d in enumerate(b['depth']):
            s = b['stride'] if i == 2 else 1
            h, w = -(-h // s), -(-w // s)
            total += 2.0 * h * w * 9 * c            # depthwise
            total += 2.0 * h * w * c * d            # pointwise
            c = d
        if b['shortcut'] == 'conv':
            total += 2.0 * h * w * b['cin'] * b['depth'][-1]
        del hin, win
    return total
