"""Host-side mirror of the reference's head interface, on top of the C ABI.

The reference exposes the hot path as three graph-builder functions and six lines of tail code
(deeplabv3p/models/layers.py:114-219, deeplabv3p/model.py:75-86).  This module mirrors them with
the same names, argument meaning and error behaviour, as objects that hold a libdlv3p context:

    ASPP_block(x_shape, OS)            layers.py:114    -> ASPPBlock
    ASPP_Lite_block(x_shape)           layers.py:166    -> ASPPLiteBlock
    Decoder_block(x_shape, skip_shape) layers.py:199    -> DecoderBlock
    get_deeplabv3p_head(model_type, num_classes, model_input_shape, output_stride, ...)
                                       model.py:51-117  -> DeepLabHead (ASPP[+Decoder]+tail in one context)

Weights use the Keras layer names and (H,W,I,O) layouts of the reference, so a `.h5` exported to
`.npz` (tools in INTEGRATION.md) or the arrays of `model.get_weights()` load unchanged.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import ffi

# backbone -> (feature channels, skip channels) the head receives (e.g.
# deeplabv3p_xception.py:150-152, deeplabv3p_mobilenetv2.py:151-152, deeplabv3p_mobilenetv3.py:590-593)
BACKBONE_CHANNELS = {
    'xception': (2048, 256), 'resnet50': (2048, 256),
    'mobilenetv2': (320, 24), 'mobilenetv3large': (160, 24), 'mobilenetv3small': (96, 16),
    'peleenet': (704, 128), 'ghostnet': (960, 24),
    'mobilevit_s': (640, 64), 'mobilevit_xs': (384, 48), 'mobilevit_xxs': (320, 24),
}
# same keys as deeplab_model_map (model.py:23-48): '<backbone>' = ASPP + Decoder, '<backbone>_lite' = ASPP Lite, no decoder
MODEL_TYPES = sorted(list(BACKBONE_CHANNELS) + [k + '_lite' for k in BACKBONE_CHANNELS if k not in ('xception', 'resnet50')])


def atrous_rates(OS: int) -> Tuple[int, int, int]:
    """layers.py:118-126 (raises ValueError exactly like the reference)."""
    if OS == 8:
        return (12, 24, 36)
    if OS == 16:
        return (6, 12, 18)
    if OS == 32:
        return (3, 6, 9)
    raise ValueError('invalid output stride', OS)


class _Block:
    """Common plumbing: a context, weights by reference layer name, numpy-in / numpy-out helpers."""

    def __init__(self, device: int = 0, **cfg):
        self.ctx = ffi.Context(device=device, **cfg)
        self.device = device
        self._bufs: Dict[str, ffi.DeviceBuffer] = {}

    # -- weights ---------------------------------------------------------------------------
    def weight_specs(self) -> List[Tuple[str, str, Tuple[int, ...]]]:
        """(layer name, variable name, shape) in Keras creation order (what load_weights(by_name=False) walks)."""
        return self.ctx.weight_specs()

    def set_weights(self, weights) -> None:
        """`weights`: dict {(layer, var): array} / {'layer/var': array}, or a flat list in weight_specs() order
        (the layout of keras `model.get_weights()` restricted to the head)."""
        specs = self.weight_specs()
        if isinstance(weights, dict):
            for layer, var, shape in specs:
                key = (layer, var) if (layer, var) in weights else '%s/%s' % (layer, var)
                if key not in weights and layer == 'conv_upsample':
                    key = ('logits_semantic', var) if ('logits_semantic', var) in weights else 'logits_semantic/%s' % var
                if key not in weights:
                    raise KeyError('missing weight %s/%s' % (layer, var))
                self.ctx.set_weight(layer, var, np.asarray(weights[key], np.float32).reshape(shape))
        else:
            weights = list(weights)
            if len(weights) != len(specs):
                raise ValueError('expected %d weight arrays, got %d' % (len(specs), len(weights)))
            for (layer, var, shape), a in zip(specs, weights):
                self.ctx.set_weight(layer, var, np.asarray(a, np.float32).reshape(shape))
        self.ctx.finalize()

    def load_weights_npz(self, path: str) -> None:
        with np.load(path) as z:
            self.set_weights({k: z[k] for k in z.files})

    def load_weights(self, path: str) -> None:
        """model.load_weights(weights_path, by_name=...) (model.py:102-103): a Keras `.h5` / `.hdf5` weight file (model.save() or
        model.save_weights(); read by the dependency-free h5lite, matched by layer / variable name) or an `.npz` with 'layer/variable' keys."""
        if path.endswith('.npz'):
            return self.load_weights_npz(path)
        from . import h5lite
        self.set_weights(h5lite.keras_weights(path))

    # -- execution -------------------------------------------------------------------------
    def _dev(self, name: str, nbytes: int) -> ffi.DeviceBuffer:
        b = self._bufs.get(name)
        if b is None or b.nbytes < nbytes:
            b = ffi.DeviceBuffer(nbytes, self.device)
            self._bufs[name] = b
        return b

    def _prep(self, a: np.ndarray) -> np.ndarray:
        if self.ctx.cfg.in_dtype == ffi.DTYPE_FP32:
            return np.ascontiguousarray(a, np.float32)
        if self.ctx.cfg.in_dtype == ffi.DTYPE_FP16:
            return np.ascontiguousarray(a, np.float16)
        if a.dtype == np.uint16:
            return np.ascontiguousarray(a)
        return ffi.f32_to_bf16_bits(a)

    def _run(self, feat: np.ndarray, skip: Optional[np.ndarray]):
        fb, sb = self.ctx.input_bytes()
        f = self._prep(feat)
        if f.nbytes != fb:
            raise ValueError('feature tensor has %d bytes, context expects %d' % (f.nbytes, fb))
        df = self._dev('feat', fb)
        df.upload(f)
        ds = None
        if sb:
            if skip is None:
                raise ValueError('this block needs a skip feature')
            s = self._prep(skip)
            if s.nbytes != sb:
                raise ValueError('skip tensor has %d bytes, context expects %d' % (s.nbytes, sb))
            ds = self._dev('skip', sb)
            ds.upload(s)
        do = self._dev('out', self.ctx.output_bytes())
        self.ctx.forward(df.ptr, None if ds is None else ds.ptr, do.ptr)
        ffi.synchronize(self.device)
        return do

    def close(self):
        for b in self._bufs.values():
            b.free()
        self._bufs.clear()
        self.ctx.close()


class ASPPBlock(_Block):
    """ASPP_block(x, OS) (layers.py:114-163). __call__(x) -> fp32 [B,h,w,256]."""

    def __init__(self, x_shape: Sequence[int], OS: int, device: int = 0, in_dtype: int = ffi.DTYPE_BF16, lite: bool = False):
        if not lite:
            atrous_rates(OS)  # ValueError('invalid output stride', OS), layers.py:126
        B, h, w, Cin = x_shape
        self.shape_out = (B, h, w, 256)
        super().__init__(device, B=B, H=h * (OS if not lite else 16), W=w * (OS if not lite else 16), OS=OS if not lite else 16,
                         h=h, w=w, Cin=Cin, Cskip=0, NC=1,
                         variant=ffi.VARIANT_ASPP_LITE if lite else ffi.VARIANT_ASPP, stages=ffi.STAGE_ASPP,
                         in_dtype=in_dtype, out_mode=ffi.OUT_FEATURES_FP32)

    def __call__(self, x: np.ndarray) -> np.ndarray:
        return self._run(x, None).download(self.shape_out, np.float32)


class ASPPLiteBlock(ASPPBlock):
    """ASPP_Lite_block(x) (layers.py:166-196)."""

    def __init__(self, x_shape: Sequence[int], device: int = 0, in_dtype: int = ffi.DTYPE_BF16):
        super().__init__(x_shape, 16, device, in_dtype, lite=True)


class DecoderBlock(_Block):
    """Decoder_block(x, skip_feature) (layers.py:199-219). __call__(x, skip) -> fp32 [B,hs,ws,256]."""

    def __init__(self, x_shape: Sequence[int], skip_shape: Sequence[int], device: int = 0, in_dtype: int = ffi.DTYPE_BF16):
        B, h, w, c = x_shape
        if c != 256:
            raise ValueError('Decoder_block input must have 256 channels (ASPP output), got %d' % c)
        B2, hs, ws, Cs = skip_shape
        if B2 != B:
            raise ValueError('batch mismatch between x and skip_feature')
        self.shape_out = (B, hs, ws, 256)
        super().__init__(device, B=B, H=hs * 4, W=ws * 4, OS=16, h=h, w=w, hs=hs, ws=ws, Cin=256, Cskip=Cs, NC=1,
                         variant=ffi.VARIANT_ASPP, stages=ffi.STAGE_DECODER, in_dtype=in_dtype, out_mode=ffi.OUT_FEATURES_FP32)

    def __call__(self, x: np.ndarray, skip: np.ndarray) -> np.ndarray:
        return self._run(x, skip).download(self.shape_out, np.float32)


class DeepLabHead(_Block):
    """ASPP(-Lite) -> [Decoder] -> conv_upsample -> pred_resize -> argmax/softmax in ONE context: everything
    get_deeplabv3p_model (model.py:51-117) places after the backbone, plus the host argmax of deeplab.py:99."""

    def __init__(self, B: int, H: int, W: int, OS: int, Cin: int, Cskip: int, NC: int, lite: bool = False,
                 decoder: Optional[bool] = None, out_mode: int = ffi.OUT_LABELS_U8, in_dtype: int = ffi.DTYPE_BF16,
                 device: int = 0, h: int = 0, w: int = 0, hs: int = 0, ws: int = 0, flags: int = 0):
        if not lite:
            atrous_rates(OS)
        elif OS not in (8, 16, 32):
            raise ValueError('invalid output stride', OS)
        if decoder is None:
            decoder = not lite      # every reference *_lite model has no decoder (deeplabv3p_mobilenetv2.py:326-331)
        stages = ffi.STAGE_ASPP | ffi.STAGE_TAIL | (ffi.STAGE_DECODER if decoder else 0)
        super().__init__(device, B=B, H=H, W=W, OS=OS, h=h, w=w, hs=hs, ws=ws, Cin=Cin, Cskip=Cskip if decoder else 0, NC=NC,
                         variant=ffi.VARIANT_ASPP_LITE if lite else ffi.VARIANT_ASPP, stages=stages,
                         in_dtype=in_dtype, out_mode=out_mode, flags=flags)
        self.B, self.H, self.W, self.NC, self.decoder, self.lite = B, H, W, NC, decoder, lite
        c = self.ctx.cfg
        self.h = h or -(-H // OS)
        self.w = w or -(-W // OS)
        self.hs = hs or -(-H // 4)
        self.ws = ws or -(-W // 4)
        self.ho, self.wo = (self.hs, self.ws) if decoder else (self.h, self.w)
        self.out_mode = out_mode

    def output_shape_dtype(self):
        if self.out_mode == ffi.OUT_LABELS_U8:
            return (self.B, self.H, self.W), np.uint8
        if self.out_mode == ffi.OUT_LOGITS_LOWRES:
            return (self.B, self.NC, self.ho, self.wo), np.float32
        return (self.B, self.H, self.W, self.NC), np.float32

    def __call__(self, feat: np.ndarray, skip: Optional[np.ndarray] = None) -> np.ndarray:
        shape, dt = self.output_shape_dtype()
        return self._run(feat, skip).download(shape, dt)

    def predict_host(self, feat: np.ndarray, skip: Optional[np.ndarray] = None, out: Optional[np.ndarray] = None) -> np.ndarray:
        """model.predict + np.argmax equivalent with host buffers end to end (dlv3p_forward_host)."""
        shape, dt = self.output_shape_dtype()
        fb, sb = self.ctx.input_bytes()
        f = self._prep(feat)
        if f.nbytes != fb:       # dlv3p_forward_host copies exactly fb / sb / output_bytes: a wrong batch or size would read / write past the arrays
            raise ValueError('feature tensor has %d bytes, context expects %d' % (f.nbytes, fb))
        s = None
        if sb:
            if skip is None:
                raise ValueError('this block needs a skip feature')
            s = self._prep(skip)
            if s.nbytes != sb:
                raise ValueError('skip tensor has %d bytes, context expects %d' % (s.nbytes, sb))
        if out is None:
            out = np.empty(shape, dt)
        elif not (isinstance(out, np.ndarray) and out.dtype == np.dtype(dt) and out.flags['C_CONTIGUOUS'] and out.nbytes == self.ctx.output_bytes()):
            raise ValueError('out must be a C-contiguous %s array of shape %s' % (np.dtype(dt).name, tuple(shape)))
        self.ctx.forward_host(f, s, out)
        return out

    def tap(self, name: str) -> np.ndarray:
        shapes = {
            'aspp_out': (self.B, self.h, self.w, 256), 'decoder_in': (self.B, self.hs, self.ws, 304),
            'decoder_conv0': (self.B, self.hs, self.ws, 256), 'decoder_out': (self.B, self.hs, self.ws, 256),
            'logits': (self.B, self.NC, self.ho, self.wo), 'image_pooling': (self.B, 256),
            'concat': (self.B, self.h, self.w, 256 if self.lite else 1024),
            'aspp_depthwise': (3, self.B, self.h, self.w, self.ctx.cfg.Cin),
        }
        return self.ctx.read_tap(name, shapes[name])


def get_deeplabv3p_head(model_type: str, num_classes: int, model_input_shape: Tuple[int, int], output_stride: int,
                        batch: int = 1, out_mode: int = ffi.OUT_LABELS_U8, in_dtype: int = ffi.DTYPE_BF16,
                        weights_path: Optional[str] = None, device: int = 0) -> DeepLabHead:
    """Head of get_deeplabv3p_model(model_type, num_classes, model_input_shape, output_stride, ...) (model.py:51)."""
    lite = model_type.endswith('_lite')
    base = model_type[:-5] if lite else model_type
    if base not in BACKBONE_CHANNELS or (lite and base in ('xception', 'resnet50')):
        raise ValueError('This model type is not supported now')     # model.py:53-54
    cin, cskip = BACKBONE_CHANNELS[base]
    head = DeepLabHead(batch, model_input_shape[0], model_input_shape[1], output_stride, cin, cskip, num_classes,
                       lite=lite, out_mode=out_mode, in_dtype=in_dtype, device=device)
    if weights_path:
        head.load_weights(weights_path)
    return head
