"""Host-side mirror of the reference's whole-model constructor for the Xception backbone, on top of the C ABI
(include/dlv3p_model.h):

    Deeplabv3pXception(input_shape, weights, input_tensor, num_classes, OS)   deeplabv3p/models/deeplabv3p_xception.py:167
    get_deeplabv3p_model('xception', num_classes, model_input_shape, output_stride, ...)   deeplabv3p/model.py:51-117
    DeepLab.predict: model.predict -> np.argmax                               deeplab.py:96-109

Images go in as the reference's loaders produce them — uint8 RGB (normalize_image runs inside the first convolution) or the
float32 [-1, 1] arrays of preprocess_image — and label maps / logits / probabilities come out.  Weights use the Keras layer and
variable names of the reference in Keras creation order (backbone, then head), so `model.get_weights()` of the reference model
or its `.h5` weight file (read by h5lite, no h5py needed) load unchanged.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import ffi


class DeepLabV3PlusXception:
    """Deeplabv3pXception + the prediction tail of model.py:75-86 in ONE native model context."""

    def __init__(self, input_shape: Sequence[int] = (512, 512, 3), num_classes: int = 21, OS: int = 16, batch: int = 1,
                 out_mode: int = ffi.OUT_LABELS_U8, image_dtype=np.uint8, device: int = 0, keep_intermediates: bool = False, flags: int = 0,
                 precision: str = 'bf16'):
        if OS not in (8, 16, 32):
            raise ValueError('invalid output stride', OS)                      # deeplabv3p_xception.py:116-117
        if len(input_shape) == 3 and input_shape[2] != 3:
            raise ValueError('input_shape must be (H, W, 3)')
        self.B, self.H, self.W, self.NC, self.OS = int(batch), int(input_shape[0]), int(input_shape[1]), int(num_classes), int(OS)
        self.image_dtype = np.dtype(image_dtype)
        if self.image_dtype not in (np.dtype(np.uint8), np.dtype(np.float32)):
            raise ValueError('image_dtype must be uint8 (raw RGB) or float32 (normalised to [-1, 1])')
        self.out_mode = out_mode
        self.device = device
        # 'bf16': the tcgen05 performance path (bf16 operands and activations, fp32 accumulation); 'fp32': the precision mode — the
        # whole model in plain fp32 arithmetic like the reference's default TensorFlow numerics (1e-4 against the fp32 oracle)
        if precision not in ('bf16', 'fp32'):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        self.precision = precision
        if precision == 'fp32':
            flags |= ffi.MODEL_FLAG_FP32
        self.model = ffi.Model(device=device, B=self.B, H=self.H, W=self.W, OS=self.OS, NC=self.NC,
                               img_dtype=ffi.IMG_U8 if self.image_dtype == np.uint8 else ffi.IMG_F32, out_mode=out_mode,
                               flags=(ffi.MODEL_FLAG_KEEP_ALL if keep_intermediates else 0) | flags)
        self._bufs: Dict[str, ffi.DeviceBuffer] = {}

    # -- weights ---------------------------------------------------------------------------
    def weight_specs(self) -> List[Tuple[str, str, Tuple[int, ...]]]:
        """(layer, variable, shape) in Keras creation order: what load_weights(by_name=False) walks (model.py:103)."""
        return self.model.weight_specs()

    def set_weights(self, weights) -> None:
        """dict {(layer, var): array} / {'layer/var': array}, or a flat list in weight_specs() order (model.get_weights())."""
        specs = self.weight_specs()
        if isinstance(weights, dict):
            for layer, var, shape in specs:
                key = (layer, var) if (layer, var) in weights else '%s/%s' % (layer, var)
                if key not in weights and layer == 'conv_upsample':
                    key = ('logits_semantic', var) if ('logits_semantic', var) in weights else 'logits_semantic/%s' % var
                if key not in weights:
                    raise KeyError('missing weight %s/%s' % (layer, var))
                self.model.set_weight(layer, var, np.asarray(weights[key], np.float32).reshape(shape))
        else:
            weights = list(weights)
            if len(weights) != len(specs):
                raise ValueError('expected %d weight arrays, got %d' % (len(specs), len(weights)))
            for (layer, var, shape), a in zip(specs, weights):
                self.model.set_weight(layer, var, np.asarray(a, np.float32).reshape(shape))
        self.model.finalize()

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
    def output_shape_dtype(self):
        if self.out_mode == ffi.OUT_LABELS_U8:
            return (self.B, self.H, self.W), np.uint8
        if self.out_mode == ffi.OUT_LOGITS_LOWRES:
            return (self.B, self.NC, -(-self.H // 4), -(-self.W // 4)), np.float32
        return (self.B, self.H, self.W, self.NC), np.float32

    def _images(self, images: np.ndarray) -> np.ndarray:
        a = np.ascontiguousarray(images, self.image_dtype)
        if a.shape != (self.B, self.H, self.W, 3):
            raise ValueError('images must have shape %s, got %s' % ((self.B, self.H, self.W, 3), a.shape))
        return a

    def _dev(self, name: str, nbytes: int) -> ffi.DeviceBuffer:
        b = self._bufs.get(name)
        if b is None or b.nbytes < nbytes:
            b = ffi.DeviceBuffer(nbytes, self.device)
            self._bufs[name] = b
        return b

    def __call__(self, images: np.ndarray) -> np.ndarray:
        """Device-resident forward (upload, forward, download; synchronous)."""
        a = self._images(images)
        shape, dt = self.output_shape_dtype()
        di, do = self._dev('img', a.nbytes), self._dev('out', self.model.output_bytes())
        di.upload(a)
        self.model.forward(di.ptr, do.ptr)
        ffi.synchronize(self.device)
        return do.download(shape, dt)

    def predict(self, images: np.ndarray, out: Optional[np.ndarray] = None) -> np.ndarray:
        """model.predict + np.argmax with host buffers end to end (dlv3p_model_forward_host)."""
        a = self._images(images)
        shape, dt = self.output_shape_dtype()
        if out is None:
            out = np.empty(shape, dt)
        elif not (isinstance(out, np.ndarray) and out.dtype == np.dtype(dt) and out.flags['C_CONTIGUOUS'] and out.nbytes == self.model.output_bytes()):
            raise ValueError('out must be a C-contiguous %s array of shape %s' % (np.dtype(dt).name, tuple(shape)))
        self.model.forward_host(a, out)
        return out

    def segment_mask(self, image) -> np.ndarray:
        """DeepLab.segment_image up to the mask (deeplab.py:81-109) with every array step on the device: preprocess_image's PIL bicubic
        resize (common/data_utils.py:449, dlv3p_op_resize_bicubic_u8 — bit exact against Pillow), normalize_image (inside the first
        convolution), the model, argmax, mask_resize back to the image's own size.  `image`: a PIL RGB image or a uint8 [H, W, 3] array;
        batch-1 models with out_mode labels.  Returns the uint8 mask [H, W]."""
        if self.B != 1 or self.out_mode != ffi.OUT_LABELS_U8 or self.image_dtype != np.uint8:
            raise ValueError('segment_mask needs a batch-1 uint8-image model with label output')
        a = np.ascontiguousarray(np.asarray(image), np.uint8)
        if a.ndim != 3 or a.shape[2] != 3:
            raise ValueError('segment_mask: expected an RGB image, got shape %s' % (a.shape,))
        H0, W0 = a.shape[:2]
        lib = ffi.load_library()
        d_raw, d_img = self._dev('raw', a.nbytes), self._dev('img', self.H * self.W * 3)
        d_lab, d_mask = self._dev('out', self.model.output_bytes()), self._dev('mask', H0 * W0)
        d_raw.upload(a)
        ffi._check(lib.dlv3p_op_resize_bicubic_u8(self.device, d_raw.ptr, 1, H0, W0, 3, self.H, self.W, d_img.ptr, None))
        self.model.forward(d_img.ptr, d_lab.ptr)
        ffi._check(lib.dlv3p_op_mask_resize_nearest(self.device, d_lab.ptr, 1, self.H, self.W, H0, W0, d_mask.ptr, None))
        ffi.synchronize(self.device)
        return d_mask.download((H0, W0), np.uint8)

    def tap(self, name: str) -> np.ndarray:
        """Backbone intermediates by Keras block name ('entry_flow_block1', 'feature', 'skip', ...) or head taps ('logits', ...)."""
        if self.model.tap_shape(name) is not None:
            return self.model.read_tap(name)
        hs, ws = -(-self.H // 4), -(-self.W // 4)
        h, w = self.model.tap_shape('feature')[1:3]
        shapes = {'aspp_out': (self.B, h, w, 256), 'decoder_in': (self.B, hs, ws, 304), 'decoder_conv0': (self.B, hs, ws, 256),
                  'decoder_out': (self.B, hs, ws, 256), 'logits': (self.B, self.NC, hs, ws), 'image_pooling': (self.B, 256)}
        return self.model.read_tap(name, shapes[name])

    def close(self):
        for b in self._bufs.values():
            b.free()
        self._bufs.clear()
        self.model.close()


def get_deeplabv3p_xception(num_classes: int, model_input_shape: Tuple[int, int], output_stride: int, batch: int = 1,
                            weights_path: Optional[str] = None, **kw) -> DeepLabV3PlusXception:
    """get_deeplabv3p_model('xception', num_classes, model_input_shape, output_stride, weights_path) (model.py:51)."""
    m = DeepLabV3PlusXception((model_input_shape[0], model_input_shape[1], 3), num_classes, output_stride, batch, **kw)
    if weights_path:
        m.load_weights(weights_path)
    return m
