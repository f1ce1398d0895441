"""Bridge from a reference tf.keras DeepLabV3+ model to the B200 head — the reference-side binding of INTEGRATION.md
as a module.  TensorFlow is OPTIONAL and imported lazily: everything except `accelerate()` only needs an object that
quacks like a Keras model (`get_layer(name)`, layer `.weights` with `.name` / `.numpy()`, `.input.shape`), which is what
the unit tests use (TensorFlow is not installable in this image).

The reference model object stays untouched: `model.load_weights(path, by_name=False)` (model.py:102-103),
`load_model(..., custom_objects=get_custom_objects())` (eval.py:566-571) and every tool that addresses layers by name
keep working, because the head's weights are READ from the reference's own layers by name:

    model = get_deeplabv3p_model(...); model.load_weights('xception.h5')         # reference code, unchanged
    fast = dlv3p_b200.keras_bridge.accelerate(model, batch=32)                    # backbone stays in TF
    labels = fast(images)                                                          # uint8 [B,H,W] == np.argmax(model.predict(images), -1)

The head's inputs are found structurally, not per backbone: `aspp0` consumes the ASPP input (layers.py:141) and
`feature_projection0` consumes the skip feature (layers.py:209) in every constructor of deeplabv3p/models/.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np

from . import ffi
from .head import DeepLabHead

_CLASSIFIER_NAMES = ('conv_upsample', 'logits_semantic')   # model.py:75 / deeplabv3p_xception.py:218


def _var_basename(name: str) -> str:
    """'decoder_conv0_depthwise/depthwise_kernel:0' -> 'depthwise_kernel' (TF1/TF2 scopes, optional ':0')."""
    return name.split('/')[-1].split(':')[0]


def _get_layer(model, name: str):
    try:
        return model.get_layer(name)
    except (ValueError, KeyError):
        return None


def _static_shape(t) -> Tuple[int, ...]:
    shp = getattr(t, 'shape', t)
    shp = shp.as_list() if hasattr(shp, 'as_list') else list(shp)
    return tuple(-1 if d is None else int(d) for d in shp)


def describe_head(model) -> Dict[str, object]:
    """Read the head's configuration off a reference model: which variant was built and with which static shapes
    (the reference requires static spatial shapes too: layers.py:129, :205)."""
    aspp0 = _get_layer(model, 'aspp0')
    if aspp0 is None:
        raise ValueError('not a DeepLabV3+ model of the reference: no layer named aspp0')
    cls = next((l for l in (_get_layer(model, n) for n in _CLASSIFIER_NAMES) if l is not None), None)
    if cls is None:
        raise ValueError('no classifier layer (conv_upsample / logits_semantic)')
    fshape = _static_shape(aspp0.input)
    in_shape = _static_shape(model.input)
    fp0 = _get_layer(model, 'feature_projection0')
    cfg = {
        'H': in_shape[1], 'W': in_shape[2], 'h': fshape[1], 'w': fshape[2], 'Cin': fshape[3],
        'lite': _get_layer(model, 'aspp1_depthwise') is None,          # ASPP_Lite_block has no atrous branches (layers.py:166-196)
        'decoder': fp0 is not None,
        'NC': _static_shape(next(w for w in cls.weights if _var_basename(w.name) == 'kernel'))[-1],
        'classifier_layer': cls.name,
    }
    if min(cfg['H'], cfg['W'], cfg['h'], cfg['w']) <= 0:
        raise ValueError('the head needs static spatial shapes (so does the reference: layers.py:129, :205)')
    cfg['OS'] = cfg['H'] // cfg['h']
    if cfg['decoder']:
        sshape = _static_shape(fp0.input)
        cfg.update(hs=sshape[1], ws=sshape[2], Cskip=sshape[3])
    else:
        cfg.update(hs=0, ws=0, Cskip=0)
    return cfg


def head_weights_from_keras(model, head: DeepLabHead) -> Dict[Tuple[str, str], np.ndarray]:
    """{(layer, var): fp32 array} for every weight the head declares, read from the reference model BY LAYER NAME."""
    out = {}
    for layer, var, shape in head.weight_specs():
        names = _CLASSIFIER_NAMES if layer == 'conv_upsample' else (layer,)
        kl = next((l for l in (_get_layer(model, n) for n in names) if l is not None), None)
        if kl is None:
            raise KeyError('the model has no layer named %s' % layer)
        found = next((w for w in kl.weights if _var_basename(w.name) == var), None)
        if found is None:
            raise KeyError('layer %s has no variable %s (has: %s)' % (layer, var, [_var_basename(w.name) for w in kl.weights]))
        a = np.asarray(found.numpy(), np.float32)
        if tuple(a.shape) != tuple(shape):
            raise ValueError('%s/%s has shape %s, the head expects %s' % (layer, var, a.shape, tuple(shape)))
        out[(layer, var)] = a
    return out


def _head_layer(model, layer: str):
    names = _CLASSIFIER_NAMES if layer == 'conv_upsample' else (layer,)
    return next((l for l in (_get_layer(model, n) for n in names) if l is not None), None)


def head_weights(model) -> Dict[Tuple[str, str], np.ndarray]:
    """Every head weight of a reference model, keyed (layer, variable) in the Keras layout — what `DeepLabHead.set_weights` and
    `train.HeadTrainer` take.  Needs no GPU: the inventory comes from a plan-only context (`dlv3p_create(device=-1)`)."""
    d = describe_head(model)
    plan = DeepLabHead(1, d['H'], d['W'], d['OS'], d['Cin'], d['Cskip'], d['NC'], lite=d['lite'], decoder=d['decoder'], device=-1,
                       h=d['h'], w=d['w'], hs=d['hs'], ws=d['ws'])
    try:
        return head_weights_from_keras(model, plan)
    finally:
        plan.close()


def load_into(model, weights: Dict[Tuple[str, str], np.ndarray]) -> int:
    """Write head weights (e.g. `HeadTrainer.get_weights()` after training) back into the reference model's own layers BY NAME, variable
    by variable (`variable.assign`), so `model.save` / `model.save_weights` (train.py:196-203) store them.  Returns the number of
    variables written; raises on a missing layer / variable or a shape mismatch."""
    n = 0
    for (layer, var), a in weights.items():
        kl = _head_layer(model, layer)
        if kl is None:
            raise KeyError('the model has no layer named %s' % layer)
        target = next((w for w in kl.weights if _var_basename(w.name) == var), None)
        if target is None:
            raise KeyError('layer %s has no variable %s' % (layer, var))
        a = np.asarray(a, np.float32)
        if tuple(_static_shape(target)) != tuple(a.shape):
            raise ValueError('%s/%s has shape %s, got %s' % (layer, var, tuple(_static_shape(target)), a.shape))
        target.assign(a)
        n += 1
    return n


def head_from_keras(model, batch: int, device: int = 0, out_mode: int = ffi.OUT_LABELS_U8,
                    in_dtype: int = ffi.DTYPE_FP32) -> DeepLabHead:
    """Build a libdlv3p context for the head of `model` and load the model's own head weights into it.
    in_dtype defaults to fp32 because that is what a TF backbone hands over; the cast to bf16 happens on the device."""
    d = describe_head(model)
    head = DeepLabHead(batch, d['H'], d['W'], d['OS'], d['Cin'], d['Cskip'], d['NC'], lite=d['lite'], decoder=d['decoder'],
                       out_mode=out_mode, in_dtype=in_dtype, device=device, h=d['h'], w=d['w'], hs=d['hs'], ws=d['ws'])
    head.set_weights(head_weights_from_keras(model, head))
    return head


class AcceleratedDeepLab:
    """`model.predict` + `np.argmax` (deeplab.py:96-99) with the head on libdlv3p: the backbone runs in TensorFlow up to
    the two tensors the head consumes, the rest is one dlv3p_forward_host call."""

    def __init__(self, backbone, head: DeepLabHead, decoder: bool):
        self.backbone, self.head, self.decoder = backbone, head, decoder

    def __call__(self, images: np.ndarray) -> np.ndarray:
        feats = self.backbone(images, training=False)
        if not isinstance(feats, (list, tuple)):
            feats = [feats]
        arrs = [np.ascontiguousarray(t.numpy() if hasattr(t, 'numpy') else t, np.float32) for t in feats[:2 if self.decoder else 1]]
        return self._predict_any_batch(arrs)

    def _predict_any_batch(self, arrs) -> np.ndarray:
        """The context has a static batch (head.B): a caller's batch is split into chunks of that size and the last, partial
        chunk (e.g. the tail of an evaluation set) is zero padded — images are independent, the padding rows are dropped."""
        n, B = arrs[0].shape[0], self.head.B
        outs = []
        for i in range(0, n, B):
            chunk = [a[i:i + B] for a in arrs]
            m = chunk[0].shape[0]
            if m < B:
                chunk = [np.concatenate([c, np.zeros((B - m,) + c.shape[1:], c.dtype)], axis=0) for c in chunk]
            y = self.head.predict_host(*[np.ascontiguousarray(c) for c in chunk])
            outs.append(y[:m].copy() if m < B or n > B else y)
        return outs[0] if len(outs) == 1 else np.concatenate(outs, axis=0)

    predict = __call__


def accelerate(model, batch: int, device: int = 0, out_mode: int = ffi.OUT_LABELS_U8) -> AcceleratedDeepLab:
    """Needs TensorFlow (the reference's own dependency): splits `model` at the head's inputs."""
    try:
        import tensorflow as tf   # noqa: WPS433 (optional dependency, imported where it is needed)
    except ImportError as e:       # pragma: no cover - TensorFlow is absent from this image
        raise ImportError('accelerate() runs the backbone in TensorFlow; use head_from_keras() / DeepLabHead without it') from e
    d = describe_head(model)
    outs = [model.get_layer('aspp0').input]
    if d['decoder']:
        outs.append(model.get_layer('feature_projection0').input)
    backbone = tf.keras.Model(model.input, outs, name='dlv3p_backbone')
    return AcceleratedDeepLab(backbone, head_from_keras(model, batch, device=device, out_mode=out_mode), d['decoder'])
