"""Groundwork for SURVEY §8(f) row N1 (the Xception backbone, not built yet): the CPU restatement oracle/xception_ref.py and its
weight inventory, pinned by the reference's own published figures for DeepLabV3+ Xception 512x512 OS16 (README.md:309)."""
import numpy as np
import pytest

from oracle import head_ref as R
from oracle import xception_ref as X


def _whole_model_specs():
    cfg = R.HeadConfig(B=1, H=512, W=512, OS=16, Cin=2048, Cskip=256, NC=21)
    return X.weight_specs(16) + R.weight_specs(cfg)


def test_parameter_count_matches_the_reference_table():
    """README.md:309 'Param 41.06M' = the trainable parameters of backbone + head (Keras model.summary): 41 055 413; with the BN moving
    statistics the model holds 41 258 213 variables."""
    specs = _whole_model_specs()
    total = sum(int(np.prod(s)) for _, _, s in specs)
    non_trainable = sum(int(np.prod(s)) for _, v, s in specs if v in ('moving_mean', 'moving_variance'))
    assert total == 41258213
    assert total - non_trainable == 41055413 and round((total - non_trainable) / 1e6, 2) == 41.06
    assert X.param_count(16) == X.param_count(8) == X.param_count(32) == 38052512     # the output stride changes strides / rates, not shapes


def test_flops_match_the_reference_table():
    """README.md:309 'FLOPS 102.73G': backbone convolutions (this inventory) + head (10.244 G GEMM + 0.386 G other)."""
    total = X.conv_flops(512, 512, 16) / 1e9 + 10.244 + 0.386
    assert abs(total - 102.73) / 102.73 < 2e-3, total


def test_layer_names_and_creation_order():
    specs = X.weight_specs(16)
    names = [l for l, v, _ in specs if v in ('kernel', 'depthwise_kernel')]
    assert names[:4] == ['entry_flow_conv1_1', 'entry_flow_conv1_2', 'entry_flow_block1_separable_conv1_depthwise', 'entry_flow_block1_separable_conv1_pointwise']
    assert 'middle_flow_unit_16_separable_conv3_pointwise' in names and names[-1] == 'exit_flow_block2_separable_conv3_pointwise'
    assert 'exit_flow_block2_shortcut' not in names and 'middle_flow_unit_1_shortcut' not in names       # 'none' / 'sum' shortcuts carry no weights
    assert sum(n.endswith('_shortcut') for n in names) == 4
    assert len(X.blocks(16)) == 21                 # 3 entry-flow + 16 middle-flow + 2 exit-flow blocks


@pytest.mark.parametrize('OS,size', [(16, 64), (8, 64), (32, 64), (16, 96)])
def test_forward_shapes_per_output_stride(OS, size):
    """feature at H/OS with 2048 channels (post-ReLU: exit_flow_block2 has depth_activation=True), skip at H/4 with 256 channels (signed: BN output)."""
    W = X.make_weights(OS, 1)
    img = np.random.default_rng(0).uniform(-1, 1, (1, size, size, 3)).astype(np.float32)
    feat, skip = X.forward_torch(img, W, OS)
    assert feat.shape == (1, size // OS, size // OS, 2048) and skip.shape == (1, size // 4, size // 4, 256)
    assert feat.min() >= 0.0 and skip.min() < 0.0
    assert np.isfinite(feat).all() and np.isfinite(skip).all()
    with pytest.raises(ValueError):
        X.os_plan(4)


def test_backbone_feeds_the_head_oracle():
    """The two tensors the backbone hands over are exactly what the head consumes (SURVEY §8(a) shape table)."""
    cfg = R.HeadConfig(B=1, H=64, W=64, OS=16, Cin=2048, Cskip=256, NC=21)
    feat, skip = X.forward_torch(np.zeros((1, 64, 64, 3), np.float32), X.make_weights(16, 2), 16)
    out = R.head_forward(feat, skip, R.make_weights(cfg, 3), cfg, 'fp32')
    assert out['labels'].shape == (1, 64, 64)


def test_shifted_add_depthwise_equals_grouped_conv():
    """The oracle's fast depthwise formulation against torch's grouped conv2d (stride 1 'same', stride 2 after explicit padding, dilation)."""
    import torch
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 24, 13, 17, generator=g, dtype=torch.float64)
    k = torch.randn(3, 3, 24, 1, generator=g, dtype=torch.float64)
    kt = k.permute(2, 3, 0, 1)
    for stride, rate in ((1, 1), (2, 1), (1, 2), (1, 4)):
        if stride == 1:
            a = X.depthwise3x3(x, k, 1, rate, rate)
            b = F.conv2d(x, kt, None, 1, rate, rate, groups=24)
        else:
            xp = F.pad(x, (rate, rate, rate, rate))
            a = X.depthwise3x3(xp, k, stride, 0, rate)
            b = F.conv2d(xp, kt, None, stride, 0, rate, groups=24)
        assert a.shape == b.shape and float((a - b).abs().max()) < 1e-12
