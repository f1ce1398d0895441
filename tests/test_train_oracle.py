"""CPU tests of the training-step oracle (oracle/train_ref.py), of the training C ABI's symbol table and of the host-side
bookkeeping of dlv3p_b200.train.  No GPU compute here."""
import os
import re

import numpy as np
import pytest

from oracle import head_ref as R
from oracle import train_ref as TR

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _tiny():
    cfg = R.HeadConfig(B=2, H=32, W=32, OS=16, Cin=8, Cskip=8, NC=5, h=4, w=4, hs=8, ws=8)
    W = R.make_weights(cfg, 7)
    feat, skip = R.make_inputs(cfg, 8)
    labels = TR.make_labels(cfg, 9)
    return cfg, W, feat, skip, labels


def test_train_header_symbols_are_exported_and_bound():
    from dlv3p_b200 import ffi, train_ffi
    text = open(os.path.join(ROOT, 'include', 'dlv3p_train.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    declared = sorted(set(re.findall(r'\b(dlv3p_(?:train|p2p|trainer)_[a-z0-9_]+)\s*\(', text)))
    assert len(declared) >= 20
    lib = ffi.load_library()
    for name in declared:
        assert hasattr(lib, name), 'libdlv3p.so does not export %s' % name
    assert {s[0] for s in train_ffi.SYMBOLS} == set(declared)
    train_ffi.lib()


def test_trainer_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from dlv3p_b200 import train, Dlv3pError
    cfg, W, *_ = _tiny()
    with pytest.raises(Dlv3pError) as e:
        train.HeadTrainer(cfg.B, cfg.H, cfg.W, cfg.OS, cfg.Cin, cfg.Cskip, cfg.NC, W)
    assert 'no CPU path' in str(e.value)


def test_oracle_gradients_match_finite_differences():
    """autograd of the restated graph vs central differences of its own loss in float64 (a few coordinates per tensor)."""
    import torch
    cfg, W, feat, skip, labels = _tiny()
    W64 = {k: np.asarray(v, np.float64) for k, v in W.items()}
    out = TR.head_train_forward_backward(feat.astype(np.float64), skip.astype(np.float64), labels, W64, cfg, keep_mask=None, dtype=torch.float64)
    rng = np.random.default_rng(0)

    def loss_at(Wp):
        return TR.head_train_forward_backward(feat.astype(np.float64), skip.astype(np.float64), labels, Wp, cfg, keep_mask=None, dtype=torch.float64)['loss']

    checked = 0
    for key in [('aspp2_depthwise', 'depthwise_kernel'), ('aspp0', 'kernel'), ('image_pooling', 'kernel'), ('concat_projection_BN', 'gamma'),
                ('decoder_conv0_pointwise', 'kernel'), ('decoder_conv1_depthwise_BN', 'beta'), ('feature_projection0', 'kernel'),
                ('conv_upsample', 'kernel'), ('conv_upsample', 'bias')]:
        g = out['grads'][key]
        for _ in range(2):
            idx = tuple(int(rng.integers(0, s)) for s in g.shape)
            eps = 1e-5
            Wp, Wm = dict(W64), dict(W64)
            Wp[key] = W64[key].copy(); Wp[key][idx] += eps
            Wm[key] = W64[key].copy(); Wm[key][idx] -= eps
            fd = (loss_at(Wp) - loss_at(Wm)) / (2 * eps)
            assert abs(fd - g[idx]) <= 1e-6 + 1e-4 * abs(fd), (key, idx, fd, g[idx])
            checked += 1
    assert checked == 18


def test_oracle_loss_ignores_label_255_but_averages_over_every_pixel():
    cfg, W, feat, skip, labels = _tiny()
    a = TR.head_train_forward_backward(feat, skip, labels, W, cfg)
    lab2 = labels.copy()
    lab2[0, :4] = 255
    b = TR.head_train_forward_backward(feat, skip, lab2, W, cfg)
    assert b['valid_pixels'] < a['valid_pixels'] and b['loss'] < a['loss']      # same denominator B*H*W (loss.py:151-154 + Keras mean)
    allign = np.full_like(labels, 255)
    c = TR.head_train_forward_backward(feat, skip, allign, W, cfg)
    assert c['loss'] == 0.0 and all(np.abs(g).max() == 0.0 for g in c['grads'].values())


def test_oracle_batch_stats_are_global_batch_moments():
    cfg, W, feat, skip, labels = _tiny()
    out = TR.head_train_forward_backward(feat, skip, labels, W, cfg)
    k = np.asarray(W[('aspp0', 'kernel')], np.float32).reshape(cfg.Cin, 256)
    r0 = feat.reshape(-1, cfg.Cin) @ k
    mean, var = out['batch_stats']['aspp0_BN']
    assert np.allclose(mean, r0.mean(0), atol=1e-5) and np.allclose(var, r0.var(0), atol=1e-5)
    Wn = TR.moving_update(W, out['batch_stats'])
    assert np.allclose(Wn[('aspp0_BN', 'moving_mean')], 0.99 * W[('aspp0_BN', 'moving_mean')] + 0.01 * mean, atol=1e-7)


def test_sgd_momentum_and_l2_rule():
    W = {('aspp0', 'kernel'): np.full((1, 1, 2, 2), 0.5, np.float32), ('aspp1_depthwise', 'depthwise_kernel'): np.ones((3, 3, 2, 1), np.float32)}
    g = {k: np.full(v.shape, 0.1, np.float32) for k, v in W.items()}
    W1, V1 = TR.sgd_momentum_update(W, g, {}, lr=0.01, momentum=0.9, l2=2e-5)
    assert np.allclose(W1[('aspp0', 'kernel')], 0.5 - 0.01 * (0.1 + 2 * 2e-5 * 0.5))
    assert np.allclose(W1[('aspp1_depthwise', 'depthwise_kernel')], 1.0 - 0.01 * 0.1)        # regulariser inert for depthwise (layers.py:24-31)
    W2, V2 = TR.sgd_momentum_update(W1, g, V1, lr=0.01, momentum=0.9, l2=0.0)
    assert np.allclose(V2[('aspp1_depthwise', 'depthwise_kernel')], 0.9 * (-0.001) - 0.001)


def test_dropout_mask_restatement():
    from dlv3p_b200 import train_ffi
    from dlv3p_b200.train import dropout_seed
    m = train_ffi.dropout_keep_mask(1 << 16, dropout_seed(3, 0, 0), 0.5)
    assert 0.48 < m.mean() < 0.52
    m2 = train_ffi.dropout_keep_mask(1 << 16, dropout_seed(3, 0, 1), 0.5)
    assert (m != m2).mean() > 0.4                                   # replicas draw different masks
    assert train_ffi.dropout_keep_mask(64, 1, 0.0).all()


@pytest.mark.parametrize('kw', [dict(loss='focal'), dict(class_weights=[0.5, 2.0, 1.0, 0.1, 3.0])])
def test_oracle_loss_variants_match_finite_differences(kw):
    """The focal and the class-weighted losses of the reference (loss.py:60-118, :159-192) through the same oracle."""
    import torch
    cfg, W, feat, skip, labels = _tiny()
    W64 = {k: np.asarray(v, np.float64) for k, v in W.items()}
    run = lambda Wp: TR.head_train_forward_backward(feat.astype(np.float64), skip.astype(np.float64), labels, Wp, cfg, dtype=torch.float64, **kw)
    out = run(W64)
    key, idx, eps = ('conv_upsample', 'kernel'), (0, 0, 3, 2), 1e-5
    Wp, Wm = dict(W64), dict(W64)
    Wp[key] = W64[key].copy(); Wp[key][idx] += eps
    Wm[key] = W64[key].copy(); Wm[key][idx] -= eps
    fd = (run(Wp)['loss'] - run(Wm)['loss']) / (2 * eps)
    assert abs(fd - out['grads'][key][idx]) <= 1e-7 + 1e-4 * abs(fd)
    assert out['loss'] != TR.head_train_forward_backward(feat.astype(np.float64), skip.astype(np.float64), labels, W64, cfg, dtype=torch.float64)['loss']


def test_lite_oracle_gradients_match_finite_differences_and_the_layout_follows():
    """The *_lite head (ASPP_Lite_block, no decoder; layers.py:166-196, deeplabv3p_mobilenetv2.py:326-331): autograd of the restated
    graph against central differences in float64, and the host-side layout of the lite trainer (no depthwise region, two exchange
    groups each way)."""
    import torch
    from dlv3p_b200 import train
    cfg = R.HeadConfig(B=2, H=32, W=32, OS=16, Cin=8, Cskip=0, NC=5, lite=True, decoder=False, h=4, w=4)
    W = R.make_weights(cfg, 17)
    feat, _ = R.make_inputs(cfg, 18)
    labels = TR.make_labels(cfg, 19)
    assert not any(k[0].startswith(('aspp1', 'decoder', 'feature_projection')) for k in W)
    W64 = {k: np.asarray(v, np.float64) for k, v in W.items()}
    f64 = feat.astype(np.float64)
    out = TR.head_train_forward_backward(f64, None, labels, W64, cfg, keep_mask=None, dtype=torch.float64)
    assert out['logits'].shape == (2, 4, 4, 5) and out['d_skip'] is None
    rng = np.random.default_rng(1)
    for key in [('aspp0', 'kernel'), ('image_pooling', 'kernel'), ('concat_projection', 'kernel'), ('concat_projection_BN', 'gamma'), ('aspp0_BN', 'beta'),
                ('conv_upsample', 'kernel'), ('conv_upsample', 'bias')]:
        g = out['grads'][key]
        for _ in range(3):
            idx = tuple(int(rng.integers(0, n)) for n in g.shape)
            Wp, Wm = dict(W64), dict(W64)
            e = 1e-5
            Wp[key] = W64[key].copy(); Wp[key][idx] += e
            Wm[key] = W64[key].copy(); Wm[key][idx] -= e
            fd = (TR.head_train_forward_backward(f64, None, labels, Wp, cfg, keep_mask=None, dtype=torch.float64)['loss'] -
                  TR.head_train_forward_backward(f64, None, labels, Wm, cfg, keep_mask=None, dtype=torch.float64)['loss']) / (2 * e)
            assert abs(fd - g[idx]) <= 1e-6 + 1e-4 * abs(fd), (key, idx, fd, g[idx])
    lay = train.TrainLayout(Cin=320, Cskip=0, NC=21, lite=True)
    assert lay.endA == lay.endB and [n for n, *_ in lay._conv_specs()] == ['image_pooling', 'aspp0', 'concat_projection', 'conv_upsample']
    assert lay.off[('concat_projection', 'kernel')][1] == (512, 256) and lay.nbn == 768
    assert len(lay.FWD_GROUPS) == 2 and len(lay.BWD_GROUPS) == 2 and len(train.TrainLayout.FWD_GROUPS) == 7
