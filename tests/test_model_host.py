"""CPU tests of the whole-model boundary (include/dlv3p_model.h): every declared symbol is exported and bound, the weight inventory
follows the reference's Keras creation order (the oracle's restatement of deeplabv3p_xception.py), plan-only models cannot compute."""
import ctypes
import os
import re

import numpy as np
import pytest

import dlv3p_b200
from dlv3p_b200 import ffi
from oracle import head_ref as R
from oracle import xception_ref as X

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, 'include', 'dlv3p_model.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(dlv3p_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_model_symbol():
    lib = ffi.load_library()
    declared = _header_symbols()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(lib, name), 'libdlv3p.so does not export %s' % name
    assert {s[0] for s in ffi.MODEL_SYMBOLS} == set(declared)


def test_model_config_struct_layout_matches_header():
    assert ctypes.sizeof(ffi.ModelConfig) == 8 * 4
    text = open(os.path.join(ROOT, 'include', 'dlv3p_model.h')).read()
    body = re.search(r'typedef struct dlv3p_model_config \{(.*?)\} dlv3p_model_config;', text, re.S).group(1)
    body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
    names = []
    for decl in body.split(';'):
        decl = decl.strip()
        if decl:
            names += [n.strip() for n in decl.split(None, 1)[1].split(',')]
    assert names == [f[0] for f in ffi.ModelConfig._fields_]


@pytest.mark.parametrize('OS', [8, 16, 32])
def test_weight_inventory_is_the_keras_creation_order(OS):
    m = dlv3p_b200.DeepLabV3PlusXception((512, 512, 3), 21, OS, batch=1, device=-1)
    specs = m.weight_specs()
    ref = [(a, b, tuple(c)) for a, b, c in X.weight_specs(OS)]
    assert specs[:len(ref)] == ref                                     # backbone: deeplabv3p_xception.py:119-152 in creation order
    head = dlv3p_b200.DeepLabHead(1, 512, 512, OS, 2048, 256, 21, device=-1).weight_specs()
    assert specs[len(ref):] == head                                    # then ASPP, decoder, classifier (layers.py:114-219, model.py:75)
    n = sum(int(np.prod(s)) for _, _, s in specs)
    assert n == X.param_count(OS) + 3205701                           # README.md:312: 41.06 M parameters
    assert abs(n / 1e6 - 41.06) < 0.3


def test_shapes_follow_the_output_stride():
    for OS, hw in ((8, 64), (16, 32), (32, 16)):
        m = dlv3p_b200.DeepLabV3PlusXception((512, 512, 3), 21, OS, batch=2, device=-1)
        assert m.model.tap_shape('feature') == (2, hw, hw, 2048)
        assert m.model.tap_shape('skip') == (2, 128, 128, 256)
        assert m.model.tap_shape('entry_flow_conv1_2') == (2, 256, 256, 64)
    m = dlv3p_b200.DeepLabV3PlusXception((100, 132, 3), 21, 16, batch=1, device=-1)     # odd sizes: ceil at every stride
    assert m.model.tap_shape('feature') == (1, 7, 9, 2048) and m.model.tap_shape('skip') == (1, 25, 33, 256)
    assert m.model.input_bytes() == 100 * 132 * 3 and m.model.output_bytes() == 100 * 132


def test_invalid_arguments_and_plan_only_model():
    with pytest.raises(ValueError):
        dlv3p_b200.DeepLabV3PlusXception((512, 512, 3), 21, 4, device=-1)          # ValueError('invalid output stride', OS)
    with pytest.raises(dlv3p_b200.Dlv3pError):
        ffi.Model(device=-1, B=1, H=512, W=512, OS=16, NC=300)
    with pytest.raises(dlv3p_b200.Dlv3pError):
        ffi.Model(device=-1, B=1, H=8, W=512, OS=16, NC=21)
    m = ffi.Model(device=-1, B=1, H=64, W=64, OS=16, NC=21)
    with pytest.raises(dlv3p_b200.Dlv3pError) as e:
        m.set_weight('entry_flow_conv1_1', 'kernel', np.zeros((3, 3, 3, 16), np.float32))
    assert e.value.status == -6
    with pytest.raises(dlv3p_b200.Dlv3pError) as e:
        m.set_weight('no_such_layer', 'kernel', np.zeros((1, 1, 8, 8), np.float32))
    assert e.value.status == -6
    for layer, var, shape in m.weight_specs():
        m.set_weight(layer, var, np.zeros(shape, np.float32))
    with pytest.raises(dlv3p_b200.Dlv3pError) as e:
        m.finalize()
    assert e.value.status == -4
    with pytest.raises(dlv3p_b200.Dlv3pError):
        m.forward(1, 1)
    if dlv3p_b200.device_count() == 0:
        with pytest.raises(dlv3p_b200.Dlv3pError) as e:
            ffi.Model(device=0, B=1, H=64, W=64, OS=16, NC=21)
        assert 'no CPU fallback' in str(e.value)


def test_backbone_oracle_bf16_mode_and_calibrated_fixture():
    """The checker itself: bf16 mode stays near fp32 mode on the calibrated fixture weights, the fixture is deterministic."""
    W = X.make_calibrated_weights(16, 7, size=64)
    W2 = X.make_calibrated_weights(16, 7, size=64)
    assert all(np.array_equal(W[k], W2[k]) for k in W)
    img = np.random.default_rng(1).uniform(-1, 1, (1, 64, 64, 3)).astype(np.float32)
    f32, s32 = X.forward_torch(img, W, 16, 'fp32')
    f16, s16 = X.forward_torch(img, W, 16, 'bf16')
    assert f32.shape == (1, 4, 4, 2048) and s32.shape == (1, 16, 16, 256)
    assert np.linalg.norm(s16 - s32) / np.linalg.norm(s32) < 2e-2
    assert np.linalg.norm(f16 - f32) / np.linalg.norm(f32) < 8e-2
    assert 0.2 < float(np.sqrt((f32 * f32).mean())) < 5.0             # O(1) activations: the regime of a trained network
