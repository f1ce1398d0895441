"""keras_bridge: reading the head's configuration and weights off a reference tf.keras model BY LAYER NAME.
TensorFlow is not installable here, so the model is a duck-typed stub with the reference's layer names, variable names
('<layer>/<var>:0') and HWIO shapes (SURVEY.md §8(b) weight-layout contract)."""
import numpy as np
import pytest

from oracle import head_ref as R
from tests.common import make_head


class _Var:
    def __init__(self, name, a):
        self.name, self._a, self.shape = name, a, a.shape

    def numpy(self):
        return self._a

    def assign(self, a):
        assert a.shape == self._a.shape
        self._a = np.array(a, np.float32)


class _Tensor:
    def __init__(self, shape):
        self.shape = list(shape)


class _Layer:
    def __init__(self, name, weights, in_shape=None):
        self.name, self.weights = name, weights
        self.input = _Tensor(in_shape) if in_shape else None


class StubKerasModel:
    """get_layer(name) over the head layers of get_deeplabv3p_model (model.py:51-117); raises ValueError like Keras."""

    def __init__(self, cfg, W, classifier_name='conv_upsample', scope=''):
        self.input = _Tensor([None, cfg.H, cfg.W, 3])
        self.layers = {}
        byl = {}
        for (layer, var), a in W.items():
            byl.setdefault(layer, []).append(_Var('%s%s/%s:0' % (scope, layer, var), np.asarray(a, np.float32)))
        for layer, ws in byl.items():
            name = classifier_name if layer == 'conv_upsample' else layer
            in_shape = None
            if layer == 'aspp0':
                in_shape = [None, cfg.h, cfg.w, cfg.Cin]
            elif layer == 'feature_projection0':
                in_shape = [None, cfg.hs, cfg.ws, cfg.Cskip]
            self.layers[name] = _Layer(name, ws, in_shape)

    def get_layer(self, name):
        if name not in self.layers:
            raise ValueError('No such layer: %s' % name)
        return self.layers[name]


CASES = [dict(B=2, H=128, W=128, OS=16, Cin=256, Cskip=64, NC=21),
         dict(B=1, H=128, W=96, OS=8, Cin=64, Cskip=24, NC=19),
         dict(B=1, H=128, W=128, OS=16, Cin=160, Cskip=24, NC=21, lite=True, decoder=False)]


@pytest.mark.parametrize('kw', CASES)
def test_describe_head_reads_the_configuration(kw):
    from dlv3p_b200 import keras_bridge as kb
    cfg = R.HeadConfig(**kw)
    d = kb.describe_head(StubKerasModel(cfg, R.make_weights(cfg, 3)))
    assert (d['H'], d['W'], d['h'], d['w'], d['Cin'], d['NC'], d['OS']) == (cfg.H, cfg.W, cfg.h, cfg.w, cfg.Cin, cfg.NC, cfg.OS)
    assert d['lite'] == cfg.lite and d['decoder'] == cfg.decoder
    if cfg.decoder:
        assert (d['hs'], d['ws'], d['Cskip']) == (cfg.hs, cfg.ws, cfg.Cskip)


@pytest.mark.parametrize('classifier_name,scope', [('conv_upsample', ''), ('logits_semantic', 'model_1/')])
def test_weights_are_read_by_layer_name(classifier_name, scope):
    import dlv3p_b200
    from dlv3p_b200 import keras_bridge as kb
    cfg = R.HeadConfig(**CASES[0])
    W = R.make_weights(cfg, 5)
    model = StubKerasModel(cfg, W, classifier_name, scope)
    plan = dlv3p_b200.DeepLabHead(cfg.B, cfg.H, cfg.W, cfg.OS, cfg.Cin, cfg.Cskip, cfg.NC, device=-1)   # plan-only: no GPU needed
    got = kb.head_weights_from_keras(model, plan)
    assert set(got) == set(W)
    for k in W:
        assert np.array_equal(got[k], np.asarray(W[k], np.float32))
    # error behaviour: a missing layer / variable / wrong shape is reported by name
    del model.layers['aspp2_pointwise_BN']
    with pytest.raises(KeyError, match='aspp2_pointwise_BN'):
        kb.head_weights_from_keras(model, plan)
    plan.close()


def test_not_a_deeplab_model_is_rejected():
    from dlv3p_b200 import keras_bridge as kb

    class Empty:
        input = _Tensor([None, 64, 64, 3])

        def get_layer(self, name):
            raise ValueError(name)
    with pytest.raises(ValueError, match='aspp0'):
        kb.describe_head(Empty())


@pytest.mark.gpu
def test_head_from_keras_equals_direct_head(gpu):
    from dlv3p_b200 import keras_bridge as kb
    cfg = R.HeadConfig(**CASES[0])
    W = R.make_weights(cfg, 9)
    feat, skip = R.make_inputs(cfg, 10)
    a = make_head(cfg, W)(feat, skip)
    hd = kb.head_from_keras(StubKerasModel(cfg, W, 'logits_semantic'), batch=cfg.B)
    b = hd.predict_host(feat, skip)                # fp32 features in, as a TF backbone would hand them over
    assert np.array_equal(a, b)
    hd.close()


def test_head_weights_and_load_into_round_trip():
    """keras_bridge.head_weights(model) -> (training) -> keras_bridge.load_into(model, weights): by layer name, shapes checked."""
    from dlv3p_b200 import keras_bridge as kb
    cfg = R.HeadConfig(B=1, H=128, W=128, OS=16, Cin=64, Cskip=32, NC=21)
    W = R.make_weights(cfg, 5)
    model = StubKerasModel(cfg, W, classifier_name='logits_semantic')
    got = kb.head_weights(model)                       # plan-only context: no GPU needed
    assert set(got) == set(W) and all(np.array_equal(got[k], np.asarray(W[k], np.float32)) for k in W)
    new = {k: v + 1.0 for k, v in got.items()}
    assert kb.load_into(model, new) == len(W)
    again = kb.head_weights(model)
    assert all(np.array_equal(again[k], new[k]) for k in new)
    with pytest.raises(ValueError):
        kb.load_into(model, {('aspp0', 'kernel'): np.zeros((1, 1, 3, 3), np.float32)})
    with pytest.raises(KeyError):
        kb.load_into(model, {('no_such_layer', 'kernel'): np.zeros((1,), np.float32)})


class _H5Dataset:
    """h5py.Dataset stand-in: `ds[()]` returns the array."""

    def __init__(self, a):
        self._a, self.shape = np.asarray(a), np.asarray(a).shape

    def __getitem__(self, idx):
        return self._a


def _h5_tree(cfg, W, scope_twice=True, classifier='conv_upsample'):
    """The group layout Keras writes: /<layer>/<layer>/<var>:0 (plus a backbone layer that must be ignored)."""
    root = {}
    for (layer, var), a in W.items():
        name = classifier if layer == 'conv_upsample' else layer
        g = root.setdefault(name, {})
        if scope_twice:
            g = g.setdefault(name, {})
        g['%s:0' % var] = _H5Dataset(a)
    root['entry_flow_conv1_1'] = {'entry_flow_conv1_1': {'kernel:0': _H5Dataset(np.zeros((3, 3, 3, 32), np.float32))}}
    return root


@pytest.mark.parametrize('scope_twice,classifier', [(True, 'conv_upsample'), (False, 'logits_semantic')])
def test_h5_export_maps_datasets_by_layer_name(tmp_path, scope_twice, classifier):
    """tools/h5_to_npz.py on an h5py-shaped tree (h5py itself is absent here): head datasets by name, backbone ignored, npz round trip."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location('h5_to_npz', os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tools', 'h5_to_npz.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    cfg = R.HeadConfig(B=1, H=128, W=128, OS=16, Cin=64, Cskip=32, NC=21)
    W = R.make_weights(cfg, 3)
    got = mod.head_weights_from_h5(_h5_tree(cfg, W, scope_twice, classifier))
    assert set(got) == {'%s/%s' % k for k in W}
    assert all(np.array_equal(got['%s/%s' % k], np.asarray(v, np.float32)) for k, v in W.items())
    path = str(tmp_path / 'head.npz')
    np.savez(path, **got)
    with np.load(path) as z:
        assert sorted(z.files) == sorted(got)
    with pytest.raises(ValueError):
        mod.head_weights_from_h5({'entry_flow_conv1_1': {'kernel:0': _H5Dataset(np.zeros(3))}})
