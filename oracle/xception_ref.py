"""
oracle/xception_ref.py — CPU restatement of the reference's modified aligned Xception feature extractor (the
next row after the head).  TEST INFRASTRUCTURE ONLY; groundwork for the next round: nothing in the product path uses it and no
CUDA backbone exists yet.

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


def _same_pad_stride2(kernel_size: int, rate: int) -> Tuple[int, int]:
    """_conv2d_same / SepConv_BN: kernel_size_effective - 1 split as (beg, end) (:44-48, layers.py:91-95)."""
    ke = kernel_size + (kernel_size - 1) * (rate - 1)
    total = ke - 1
    beg = total // 2
    return beg, total - beg


def forward_torch(image_nhwc: np.ndarray, W: Dict[Key, np.ndarray], OS: int = 16):
    """Inference-mode forward (BN with moving statistics).  image [B,H,W,3] fp32 in [-1, 1] (normalize_image).
    Returns (feature [B,H/OS,W/OS,2048], skip [B,H/4,W/4,256]) as NHWC numpy arrays."""
    import torch
    import torch.nn.functional as F
    T = lambda k: torch.from_numpy(np.asarray(W[k], np.float32))

    def bn(x, name):
        return F.batch_norm(x, T((name, 'moving_mean')), T((name, 'moving_variance')), T((name, 'gamma')), T((name, 'beta')), False, 0.0, BN_EPS_BACKBONE)

    def conv_same(x, name, stride, ksize, rate=1):
        k = T((name, 'kernel')).permute(3, 2, 0, 1)
        if stride == 1:
            pad = ((ksize - 1) * rate) // 2
            return F.conv2d(x, k, None, 1, pad, rate)
        beg, end = _same_pad_stride2(ksize, rate)
        return F.conv2d(F.pad(x, (beg, end, beg, end)), k, None, stride, 0, rate)

    def sepconv_bn(x, prefix, stride, rate, act):
        if stride != 1:
            beg, end = _same_pad_stride2(3, rate)
            x = F.pad(x, (beg, end, beg, end))
            pad = 0
        else:
            pad = rate
        if not act:
            x = F.relu(x)
        k = T((prefix + '_depthwise', 'depthwise_kernel')).permute(2, 3, 0, 1)
        x = F.conv2d(x, k, None, stride, pad, rate, groups=k.shape[0])
        x = bn(x, prefix + '_depthwise_BN')
        if act:
            x = F.relu(x)
        x = F.conv2d(x, T((prefix + '_pointwise', 'kernel')).permute(3, 2, 0, 1))
        x = bn(x, prefix + '_pointwise_BN')
        return F.relu(x) if act else x

    x = torch.from_numpy(np.asarray(image_nhwc, np.float32)).permute(0, 3, 1, 2)
    # entry_flow_conv1_1: Conv2D(32, 3, strides 2, padding 'same') — TensorFlow 'same' at stride 2 pads (0, 1) for even sizes (:119-120)
    H, Wd = x.shape[2], x.shape[3]
    ph = max((-(-H // 2) - 1) * 2 + 3 - H, 0)
    pw = max((-(-Wd // 2) - 1) * 2 + 3 - Wd, 0)
    x = F.conv2d(F.pad(x, (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2)), T(('entry_flow_conv1_1', 'kernel')).permute(3, 2, 0, 1), None, 2)
    x = F.relu(bn(x, 'entry_flow_conv1_1_BN'))
    x = F.relu(bn(conv_same(x, 'entry_flow_conv1_2', 1, 3), 'entry_flow_conv1_2_BN'))
    skip = None
    for b in blocks(OS):
        inp = x
        r = x
        for i in range(3):
            r = sepconv_bn(r, '%s_separable_conv%d' % (b['prefix'], i + 1), b['stride'] if i == 2 else 1, b['rate'], b['act'])
            if i == 1 and b.get('return_skip'):
                skip = r
        if b['shortcut'] == 'conv':
            sc = bn(conv_same(inp, b['prefix'] + '_shortcut', b['stride'], 1), b['prefix'] + '_shortcut_BN')
            x = r + sc
        elif b['shortcut'] == 'sum':
            x = r + inp
        else:
            x = r
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
