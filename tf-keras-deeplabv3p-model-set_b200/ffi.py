"""ctypes binding of libdlv3p.so (include/dlv3p.h) — no PyTorch, no cuda-python.

This is the reference-side binding a maintainer of tf-keras-deeplabv3p-model-set would add (see
INTEGRATION.md): plain pointers and sizes over the C ABI.  There is deliberately NO fallback: if the
shared library is missing, or the machine has no sm_100 GPU, every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libdlv3p.so')

# ---- enums (mirror include/dlv3p.h) ----------------------------------------------------------
STAGE_ASPP, STAGE_DECODER, STAGE_TAIL = 1, 2, 4
VARIANT_ASPP, VARIANT_ASPP_LITE = 0, 1
DTYPE_BF16, DTYPE_FP16, DTYPE_FP32 = 0, 1, 2
OUT_LABELS_U8, OUT_LOGITS_LOWRES, OUT_SOFTMAX, OUT_LOGITS_FULL, OUT_FEATURES_BF16, OUT_FEATURES_FP32 = range(6)
FLAG_UNFUSED_DECODER = 1

STATUS = {0: 'OK', -1: 'ERR_INVALID', -2: 'ERR_CUDA', -3: 'ERR_UNSUPPORTED', -4: 'ERR_STATE', -5: 'ERR_NOMEM', -6: 'ERR_NAME'}


class Dlv3pError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__('dlv3p %s (%d): %s' % (STATUS.get(status, '?'), status, message))
        self.status = status


class Config(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ('B', 'H', 'W', 'OS', 'h', 'w', 'hs', 'ws', 'Cin', 'Cskip', 'NC', 'variant', 'stages', 'in_dtype', 'out_mode')] + \
               [('bn_eps', C.c_float), ('flags', C.c_int32)]


class ModelConfig(C.Structure):
    """dlv3p_model_config (include/dlv3p_model.h)."""
    _fields_ = [(n, C.c_int32) for n in ('B', 'H', 'W', 'OS', 'NC', 'img_dtype', 'out_mode', 'flags')]


IMG_U8, IMG_F32 = 0, 1
MODEL_FLAG_KEEP_ALL = 1
MODEL_FLAG_NO_PDL = 2
MODEL_FLAG_UNFUSED_ENTRY = 4
MODEL_FLAG_FP32 = 8
MODEL_FLAG_FUSED_MIDDLE = 16
MODEL_FLAG_FP32_STEM = 32

# every symbol include/dlv3p.h declares: (name, restype, argtypes)
_vp, _i, _sz = C.c_void_p, C.c_int, C.c_size_t
_fp = C.POINTER(C.c_float)
SYMBOLS = [
    ('dlv3p_abi_version', _i, []),
    ('dlv3p_create', _i, [C.POINTER(Config), _i, C.POINTER(_vp)]),
    ('dlv3p_destroy', None, [_vp]),
    ('dlv3p_last_error', C.c_char_p, [_vp]),
    ('dlv3p_set_weight', _i, [_vp, C.c_char_p, C.c_char_p, _fp, C.POINTER(C.c_int64), _i]),
    ('dlv3p_num_weights', _i, [_vp]),
    ('dlv3p_weight_info', _i, [_vp, _i, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.POINTER(C.c_int64), C.POINTER(_i)]),
    ('dlv3p_finalize_weights', _i, [_vp]),
    ('dlv3p_forward', _i, [_vp, _vp, _vp, _vp, _vp]),
    ('dlv3p_forward_host', _i, [_vp, _vp, _vp, _vp]),
    ('dlv3p_input_bytes', _i, [_vp, C.POINTER(_sz), C.POINTER(_sz)]),
    ('dlv3p_output_bytes', _i, [_vp, C.POINTER(_sz)]),
    ('dlv3p_workspace_bytes', _i, [_vp, C.POINTER(_sz)]),
    ('dlv3p_read_tap', _i, [_vp, C.c_char_p, _fp, _sz]),
    ('dlv3p_launch_count', _i, [_vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    ('dlv3p_profile_forward', _i, [_vp, _vp, _vp, _vp, _vp, C.POINTER(C.c_char_p), _fp, _i]),
    ('dlv3p_device_count', _i, [C.POINTER(_i)]),
    ('dlv3p_device_info', _i, [_i, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), C.POINTER(_sz)]),
    ('dlv3p_dev_alloc', _i, [_i, _sz, C.POINTER(_vp)]),
    ('dlv3p_dev_free', _i, [_i, _vp]),
    ('dlv3p_host_alloc_pinned', _i, [_sz, C.POINTER(_vp)]),
    ('dlv3p_host_free_pinned', _i, [_vp]),
    ('dlv3p_memcpy_h2d', _i, [_i, _vp, _vp, _sz]),
    ('dlv3p_memcpy_d2h', _i, [_i, _vp, _vp, _sz]),
    ('dlv3p_dev_memset', _i, [_i, _vp, _i, _sz]),
    ('dlv3p_dev_synchronize', _i, [_i]),
    ('dlv3p_op_pointwise', _i, [_i, _vp, C.c_int64, _i, _i, _fp, _fp, _fp, _i, _vp, _vp]),
    ('dlv3p_op_depthwise', _i, [_i, _vp, _i, _i, _i, _i, _i, _fp, _fp, _fp, _i, _vp, _vp]),
    ('dlv3p_op_sepconv', _i, [_i, _vp, _i, _i, _i, _i, _i, _fp, _fp, _fp, _i, _fp, _fp, _fp, _vp, _vp]),
    ('dlv3p_op_resize_bilinear', _i, [_i, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    ('dlv3p_op_resize_argmax', _i, [_i, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    ('dlv3p_op_time', _i, [_i, _i, C.POINTER(C.c_int64), _i, _i, _i, _fp]),
    ('dlv3p_op_confusion_matrix', _i, [_i, _vp, _vp, C.c_int64, _i, _vp, _vp]),
    ('dlv3p_op_jaccard_counts', _i, [_i, _vp, _vp, _i, C.c_int64, _i, _vp, _vp]),
    ('dlv3p_op_normalize_image', _i, [_i, _vp, C.c_int64, _vp, _i, _vp]),
    ('dlv3p_op_denormalize_image', _i, [_i, _vp, C.c_int64, _vp, _vp]),
    ('dlv3p_op_mask_resize_nearest', _i, [_i, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    ('dlv3p_op_resize_bicubic_u8', _i, [_i, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    ('dlv3p_pil_bicubic_coeffs', _i, [_i, _i, _vp, _vp, _i, _vp]),
    ('dlv3p_op_present_classes', _i, [_i, _vp, _i, C.c_int64, _vp, _vp]),
    ('dlv3p_op_bn_scratch_bytes', C.c_size_t, [_i]),
    ('dlv3p_op_bn_stats', _i, [_i, _vp, C.c_int64, _i, _vp, _vp, _vp]),
    ('dlv3p_op_bn_apply', _i, [_i, _vp, C.c_int64, _i, _vp, _vp, _vp, C.c_float, _i, _vp, _vp]),
]

# every symbol include/dlv3p_model.h declares
_dp = C.POINTER(C.c_double)
MODEL_SYMBOLS = [
    ('dlv3p_model_create', _i, [C.POINTER(ModelConfig), _i, C.POINTER(_vp)]),
    ('dlv3p_model_destroy', None, [_vp]),
    ('dlv3p_model_last_error', C.c_char_p, [_vp]),
    ('dlv3p_model_num_weights', _i, [_vp]),
    ('dlv3p_model_weight_info', _i, [_vp, _i, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.POINTER(C.c_int64), C.POINTER(_i)]),
    ('dlv3p_model_set_weight', _i, [_vp, C.c_char_p, C.c_char_p, _fp, C.POINTER(C.c_int64), _i]),
    ('dlv3p_model_finalize_weights', _i, [_vp]),
    ('dlv3p_model_forward', _i, [_vp, _vp, _vp, _vp]),
    ('dlv3p_model_forward_host', _i, [_vp, _vp, _vp]),
    ('dlv3p_model_input_bytes', _i, [_vp, C.POINTER(_sz)]),
    ('dlv3p_model_output_bytes', _i, [_vp, C.POINTER(_sz)]),
    ('dlv3p_model_workspace_bytes', _i, [_vp, C.POINTER(_sz)]),
    ('dlv3p_model_read_tap', _i, [_vp, C.c_char_p, _fp, _sz]),
    ('dlv3p_model_tap_shape', _i, [_vp, C.c_char_p, C.POINTER(C.c_int64)]),
    ('dlv3p_model_forward_from', _i, [_vp, C.c_char_p, _fp, _sz, _vp, _vp]),
    ('dlv3p_model_launch_count', _i, [_vp, C.POINTER(C.c_int64)]),
    ('dlv3p_model_profile_forward', _i, [_vp, _vp, _vp, _vp, C.POINTER(C.c_char_p), _fp, _dp, _dp, _i]),
    ('dlv3p_op_bb_depthwise', _i, [_i, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _fp, _fp, _fp, _vp, _vp]),
    ('dlv3p_op_bb_pointwise', _i, [_i, _vp, C.c_int64, _i, _i, _fp, _fp, _fp, _i, _vp, _vp, _vp]),
    ('dlv3p_op_bb_sepwide', _i, [_i, _vp, _i, _i, _i, _i, _i, _i, _fp, _fp, _fp, _fp, _fp, _fp, _i, _vp, _vp, _vp]),
    ('dlv3p_op_conv3x3_c32', _i, [_i, _vp, _i, _i, _i, _fp, _fp, _fp, _vp, _vp]),
    ('dlv3p_op_stem_conv', _i, [_i, _vp, _i, _i, _i, _i, _fp, _fp, _fp, _vp, _vp]),
    ('dlv3p_op_bb_time', _i, [_i, _i, C.POINTER(C.c_int64), _i, _i, _i, _fp]),
]

_lib = None


def load_library(path: Optional[str] = None) -> C.CDLL:
    """dlopen libdlv3p.so and type every entry point. Raises if the extension has not been built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise ImportError('libdlv3p.so not found at %s — build it first: python -c "import __graft_entry__ as g; g.build()" '
                          '(there is no CPU / PyTorch fallback for this path)' % p)
    lib = C.CDLL(p)
    for name, res, args in SYMBOLS + MODEL_SYMBOLS:
        fn = getattr(lib, name)          # AttributeError here == header/library mismatch
        fn.restype = res
        fn.argtypes = args
    if lib.dlv3p_abi_version() != 1:
        raise ImportError('libdlv3p.so ABI version %d, binding expects 1' % lib.dlv3p_abi_version())
    if path is None:
        _lib = lib
    return lib


def _check(status: int, ctx=None):
    if status < 0:
        msg = load_library().dlv3p_last_error(ctx)
        raise Dlv3pError(status, msg.decode() if msg else '')
    return status


def device_count() -> int:
    n = C.c_int(0)
    st = load_library().dlv3p_device_count(C.byref(n))
    return n.value if st == 0 else 0


def device_info(device: int = 0) -> dict:
    ma, mi, sms, mem = C.c_int(), C.c_int(), C.c_int(), C.c_size_t()
    _check(load_library().dlv3p_device_info(device, C.byref(ma), C.byref(mi), C.byref(sms), C.byref(mem)))
    return {'sm': (ma.value, mi.value), 'sm_count': sms.value, 'total_mem': mem.value}


# ---- bf16 <-> numpy (uint16 bit patterns; round to nearest even) ------------------------------
def f32_to_bf16_bits(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    r = (u + 0x7FFF + ((u >> 16) & 1)) >> 16
    return (r & 0xFFFF).astype(np.uint16).reshape(x.shape)


def bf16_bits_to_f32(b: np.ndarray) -> np.ndarray:
    return (np.ascontiguousarray(b, dtype=np.uint16).astype(np.uint32) << 16).view(np.float32).reshape(b.shape)


class DeviceBuffer:
    """A cudaMalloc'ed buffer owned by Python (freed on GC)."""

    def __init__(self, nbytes: int, device: int = 0):
        self.device, self.nbytes = device, int(nbytes)
        p = C.c_void_p()
        _check(load_library().dlv3p_dev_alloc(device, self.nbytes, C.byref(p)))
        self.ptr = p.value

    @classmethod
    def from_numpy(cls, a: np.ndarray, device: int = 0) -> 'DeviceBuffer':
        a = np.ascontiguousarray(a)
        buf = cls(a.nbytes, device)
        buf.upload(a)
        return buf

    def upload(self, a: np.ndarray):
        a = np.ascontiguousarray(a)
        assert a.nbytes <= self.nbytes
        _check(load_library().dlv3p_memcpy_h2d(self.device, self.ptr, a.ctypes.data, a.nbytes))

    def download(self, shape: Sequence[int], dtype) -> np.ndarray:
        out = np.empty(shape, dtype=dtype)
        assert out.nbytes <= self.nbytes
        _check(load_library().dlv3p_memcpy_d2h(self.device, out.ctypes.data, self.ptr, out.nbytes))
        return out

    def memset(self, value: int = 0):
        _check(load_library().dlv3p_dev_memset(self.device, self.ptr, value, self.nbytes))

    def free(self):
        if getattr(self, 'ptr', None):
            load_library().dlv3p_dev_free(self.device, self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class PinnedBuffer:
    """cudaMallocHost'ed host memory exposed as a numpy array."""

    def __init__(self, nbytes: int):
        p = C.c_void_p()
        _check(load_library().dlv3p_host_alloc_pinned(int(nbytes), C.byref(p)))
        self.ptr, self.nbytes = p.value, int(nbytes)
        self.array = np.ctypeslib.as_array((C.c_uint8 * self.nbytes).from_address(self.ptr))

    def view(self, dtype, shape):
        return self.array[:int(np.prod(shape)) * np.dtype(dtype).itemsize].view(dtype).reshape(shape)

    def free(self):
        if getattr(self, 'ptr', None):
            self.array = None
            load_library().dlv3p_host_free_pinned(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def synchronize(device: int = 0):
    _check(load_library().dlv3p_dev_synchronize(device))


class Context:
    """One dlv3p_ctx: a head instance with static shapes on one device."""

    def __init__(self, device: int = 0, **cfg):
        self.lib = load_library()
        self.cfg = Config()
        for k, v in cfg.items():
            if not hasattr(self.cfg, k):
                raise TypeError('unknown config field %r' % k)
            setattr(self.cfg, k, v)
        self.device = device
        h = C.c_void_p()
        _check(self.lib.dlv3p_create(C.byref(self.cfg), device, C.byref(h)))
        self.handle = h

    # -- weights --------------------------------------------------------------------------
    def weight_specs(self) -> List[Tuple[str, str, Tuple[int, ...]]]:
        out = []
        for i in range(self.lib.dlv3p_num_weights(self.handle)):
            layer, var = C.c_char_p(), C.c_char_p()
            shape = (C.c_int64 * 4)()
            rank = C.c_int()
            _check(self.lib.dlv3p_weight_info(self.handle, i, C.byref(layer), C.byref(var), shape, C.byref(rank)), self.handle)
            out.append((layer.value.decode(), var.value.decode(), tuple(int(shape[j]) for j in range(rank.value))))
        return out

    def set_weight(self, layer: str, var: str, value: np.ndarray):
        a = np.ascontiguousarray(value, dtype=np.float32)
        shape = (C.c_int64 * a.ndim)(*a.shape)
        _check(self.lib.dlv3p_set_weight(self.handle, layer.encode(), var.encode(),
                                         a.ctypes.data_as(_fp), shape, a.ndim), self.handle)

    def set_weights(self, weights: Dict[Tuple[str, str], np.ndarray], finalize: bool = True):
        for (layer, var), v in weights.items():
            self.set_weight(layer, var, v)
        if finalize:
            self.finalize()

    def finalize(self):
        _check(self.lib.dlv3p_finalize_weights(self.handle), self.handle)

    # -- sizes ----------------------------------------------------------------------------
    def input_bytes(self) -> Tuple[int, int]:
        f, s = C.c_size_t(), C.c_size_t()
        _check(self.lib.dlv3p_input_bytes(self.handle, C.byref(f), C.byref(s)), self.handle)
        return f.value, s.value

    def output_bytes(self) -> int:
        o = C.c_size_t()
        _check(self.lib.dlv3p_output_bytes(self.handle, C.byref(o)), self.handle)
        return o.value

    def workspace_bytes(self) -> int:
        o = C.c_size_t()
        _check(self.lib.dlv3p_workspace_bytes(self.handle, C.byref(o)), self.handle)
        return o.value

    # -- execution ------------------------------------------------------------------------
    def forward(self, d_feat: int, d_skip: Optional[int], d_out: int, stream: int = 0):
        _check(self.lib.dlv3p_forward(self.handle, d_feat, d_skip, d_out, stream), self.handle)

    def forward_host(self, h_feat: np.ndarray, h_skip: Optional[np.ndarray], h_out: np.ndarray):
        _check(self.lib.dlv3p_forward_host(self.handle, h_feat.ctypes.data, None if h_skip is None else h_skip.ctypes.data,
                                           h_out.ctypes.data), self.handle)

    def profile(self, d_feat: int, d_skip: Optional[int], d_out: int, stream: int = 0) -> List[Tuple[str, float]]:
        names = (C.c_char_p * 64)()
        ms = (C.c_float * 64)()
        n = _check(self.lib.dlv3p_profile_forward(self.handle, d_feat, d_skip, d_out, stream, names, ms, 64), self.handle)
        return [(names[i].decode(), float(ms[i])) for i in range(n)]

    def read_tap(self, name: str, shape: Sequence[int]) -> np.ndarray:
        out = np.empty(shape, dtype=np.float32)
        _check(self.lib.dlv3p_read_tap(self.handle, name.encode(), out.ctypes.data_as(_fp), out.size), self.handle)
        return out

    def launch_count(self) -> Tuple[int, int]:
        a, b = C.c_int64(), C.c_int64()
        _check(self.lib.dlv3p_launch_count(self.handle, C.byref(a), C.byref(b)), self.handle)
        return a.value, b.value

    def close(self):
        if getattr(self, 'handle', None):
            self.lib.dlv3p_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Model:
    """One dlv3p_model: the whole DeepLabV3+ Xception network (backbone + head) with static shapes on one device."""

    def __init__(self, device: int = 0, **cfg):
        self.lib = load_library()
        self.cfg = ModelConfig()
        for k, v in cfg.items():
            if not hasattr(self.cfg, k):
                raise TypeError('unknown model config field %r' % k)
            setattr(self.cfg, k, v)
        self.device = device
        h = C.c_void_p()
        st = self.lib.dlv3p_model_create(C.byref(self.cfg), device, C.byref(h))
        if st < 0:
            raise Dlv3pError(st, (self.lib.dlv3p_model_last_error(None) or b'').decode())
        self.handle = h

    def _check(self, st: int) -> int:
        if st < 0:
            raise Dlv3pError(st, (self.lib.dlv3p_model_last_error(self.handle) or b'').decode())
        return st

    def weight_specs(self) -> List[Tuple[str, str, Tuple[int, ...]]]:
        out = []
        for i in range(self.lib.dlv3p_model_num_weights(self.handle)):
            layer, var = C.c_char_p(), C.c_char_p()
            shape = (C.c_int64 * 4)()
            rank = C.c_int()
            self._check(self.lib.dlv3p_model_weight_info(self.handle, i, C.byref(layer), C.byref(var), shape, C.byref(rank)))
            out.append((layer.value.decode(), var.value.decode(), tuple(int(shape[j]) for j in range(rank.value))))
        return out

    def set_weight(self, layer: str, var: str, value: np.ndarray):
        a = np.ascontiguousarray(value, dtype=np.float32)
        shape = (C.c_int64 * a.ndim)(*a.shape)
        self._check(self.lib.dlv3p_model_set_weight(self.handle, layer.encode(), var.encode(), a.ctypes.data_as(_fp), shape, a.ndim))

    def finalize(self):
        self._check(self.lib.dlv3p_model_finalize_weights(self.handle))

    def input_bytes(self) -> int:
        o = C.c_size_t()
        self._check(self.lib.dlv3p_model_input_bytes(self.handle, C.byref(o)))
        return o.value

    def output_bytes(self) -> int:
        o = C.c_size_t()
        self._check(self.lib.dlv3p_model_output_bytes(self.handle, C.byref(o)))
        return o.value

    def workspace_bytes(self) -> int:
        o = C.c_size_t()
        self._check(self.lib.dlv3p_model_workspace_bytes(self.handle, C.byref(o)))
        return o.value

    def forward(self, d_images: int, d_out: int, stream: int = 0):
        self._check(self.lib.dlv3p_model_forward(self.handle, d_images, d_out, stream))

    def forward_host(self, h_images: np.ndarray, h_out: np.ndarray):
        self._check(self.lib.dlv3p_model_forward_host(self.handle, h_images.ctypes.data, h_out.ctypes.data))

    def profile(self, d_images: int, d_out: int, stream: int = 0, max_kernels: int = 512):
        """[(kernel name, ms, algorithmic flops, algorithmic bytes)] of one forward, backbone first."""
        names = (C.c_char_p * max_kernels)()
        ms = (C.c_float * max_kernels)()
        fl = (C.c_double * max_kernels)()
        by = (C.c_double * max_kernels)()
        n = self._check(self.lib.dlv3p_model_profile_forward(self.handle, d_images, d_out, stream, names, ms, fl, by, max_kernels))
        return [(names[i].decode(), float(ms[i]), float(fl[i]), float(by[i])) for i in range(n)]

    def tap_shape(self, name: str) -> Optional[Tuple[int, ...]]:
        shape = (C.c_int64 * 4)()
        if self.lib.dlv3p_model_tap_shape(self.handle, name.encode(), shape) < 0:
            return None
        return tuple(int(v) for v in shape)

    def read_tap(self, name: str, shape: Optional[Sequence[int]] = None) -> np.ndarray:
        shape = self.tap_shape(name) if shape is None else tuple(shape)
        if shape is None:
            raise KeyError('tap %r is not a backbone tap: pass its shape (head taps, see dlv3p_read_tap)' % name)
        out = np.empty(shape, dtype=np.float32)
        self._check(self.lib.dlv3p_model_read_tap(self.handle, name.encode(), out.ctypes.data_as(_fp), out.size))
        return out

    def forward_from(self, tap: str, value: np.ndarray, d_out: int, stream: int = 0):
        """Overwrite a backbone tap with `value` (fp32, rounded to bf16 on the way) and run only the kernels after it."""
        a = np.ascontiguousarray(value, np.float32)
        self._check(self.lib.dlv3p_model_forward_from(self.handle, tap.encode(), a.ctypes.data_as(_fp), a.size, d_out, stream))

    def launch_count(self) -> int:
        a = C.c_int64()
        self._check(self.lib.dlv3p_model_launch_count(self.handle, C.byref(a)))
        return a.value

    def close(self):
        if getattr(self, 'handle', None):
            self.lib.dlv3p_model_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- standalone operators (unit parity tests) -------------------------------------------------
def _f(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32).ctypes.data_as(_fp)


def op_pointwise(a_bits: np.ndarray, w_kn: np.ndarray, scale=None, shift=None, relu=True, device=0) -> np.ndarray:
    """a_bits: uint16 bf16 [M,K]; w_kn fp32 [K,N] -> uint16 bf16 [M,N]."""
    M, K = a_bits.shape
    N = w_kn.shape[1]
    da = DeviceBuffer.from_numpy(a_bits, device)
    do = DeviceBuffer(M * N * 2, device)
    w = np.ascontiguousarray(w_kn, np.float32)
    s = None if scale is None else np.ascontiguousarray(scale, np.float32)
    t = None if shift is None else np.ascontiguousarray(shift, np.float32)
    _check(load_library().dlv3p_op_pointwise(device, da.ptr, M, K, N, _f(w), _f(s), _f(t), int(relu), do.ptr, None))
    return do.download((M, N), np.uint16)


def op_depthwise(x_bits: np.ndarray, w_hwc: np.ndarray, rate: int, scale=None, shift=None, relu=True, device=0) -> np.ndarray:
    B, H, W, Cc = x_bits.shape
    dx = DeviceBuffer.from_numpy(x_bits, device)
    do = DeviceBuffer(x_bits.nbytes, device)
    w = np.ascontiguousarray(w_hwc, np.float32)
    s = None if scale is None else np.ascontiguousarray(scale, np.float32)
    t = None if shift is None else np.ascontiguousarray(shift, np.float32)
    _check(load_library().dlv3p_op_depthwise(device, dx.ptr, B, H, W, Cc, rate, _f(w), _f(s), _f(t), int(relu), do.ptr, None))
    return do.download(x_bits.shape, np.uint16)


def op_sepconv(x_bits: np.ndarray, dw_hwc, dw_scale, dw_shift, pw_kn, pw_scale, pw_shift, rate: int = 1, device=0) -> np.ndarray:
    B, H, W, Cc = x_bits.shape
    N = pw_kn.shape[1]
    dx = DeviceBuffer.from_numpy(x_bits, device)
    do = DeviceBuffer(B * H * W * N * 2, device)
    arrs = [np.ascontiguousarray(a, np.float32) for a in (dw_hwc, dw_scale, dw_shift, pw_kn, pw_scale, pw_shift)]
    _check(load_library().dlv3p_op_sepconv(device, dx.ptr, B, H, W, Cc, rate, _f(arrs[0]), _f(arrs[1]), _f(arrs[2]), N,
                                           _f(arrs[3]), _f(arrs[4]), _f(arrs[5]), do.ptr, None))
    return do.download((B, H, W, N), np.uint16)


def op_bb_depthwise(x_bits: np.ndarray, w_hwc: np.ndarray, stride: int = 1, rate: int = 1, relu_in: bool = True, relu_out: bool = False,
                    scale=None, shift=None, device=0) -> np.ndarray:
    """Backbone depthwise half of SepConv_BN (layers.py:88-104): x uint16 bf16 [B,H,W,C] -> uint16 bf16 [B,ceil(H/s),ceil(W/s),C]."""
    B, H, W, Cc = x_bits.shape
    Ho, Wo = -(-H // stride), -(-W // stride)
    dx = DeviceBuffer.from_numpy(x_bits, device)
    do = DeviceBuffer(B * Ho * Wo * Cc * 2, device)
    _check(load_library().dlv3p_op_bb_depthwise(device, dx.ptr, B, H, W, Cc, stride, rate, int(relu_in), int(relu_out), _f(w_hwc), _f(scale), _f(shift), do.ptr, None))
    return do.download((B, Ho, Wo, Cc), np.uint16)


def op_bb_pointwise(a_bits: np.ndarray, w_kn: np.ndarray, scale=None, shift=None, relu: bool = False, residual_bits: Optional[np.ndarray] = None,
                    device=0) -> np.ndarray:
    """Backbone 1x1 conv + BN [+ReLU] [+residual]: a uint16 bf16 [M,K], w fp32 [K,N] (any N % 8 == 0) -> uint16 bf16 [M,N]."""
    M, K = a_bits.shape
    N = w_kn.shape[1]
    da = DeviceBuffer.from_numpy(a_bits, device)
    dr = None if residual_bits is None else DeviceBuffer.from_numpy(residual_bits, device)
    do = DeviceBuffer(M * N * 2, device)
    _check(load_library().dlv3p_op_bb_pointwise(device, da.ptr, M, K, N, _f(w_kn), _f(scale), _f(shift), int(relu), None if dr is None else dr.ptr, do.ptr, None))
    return do.download((M, N), np.uint16)


def op_bb_sepwide(x_bits: np.ndarray, dw_hwc: np.ndarray, w_kn: np.ndarray, relu_in: bool = True, dw_scale=None, dw_shift=None, scale=None, shift=None,
                  relu_out: bool = False, residual_bits: Optional[np.ndarray] = None, device=0) -> np.ndarray:
    """Fused middle-flow SepConv_BN (layers.py:74-111, depth_activation False) [+ residual]: x uint16 bf16 [B,H,W,C] -> uint16 bf16 [B,H,W,N]."""
    B, H, W, Cc = x_bits.shape
    N = w_kn.shape[1]
    dx = DeviceBuffer.from_numpy(x_bits, device)
    dr = None if residual_bits is None else DeviceBuffer.from_numpy(residual_bits, device)
    do = DeviceBuffer(B * H * W * N * 2, device)
    _check(load_library().dlv3p_op_bb_sepwide(device, dx.ptr, B, H, W, Cc, N, int(relu_in), _f(dw_hwc), _f(dw_scale), _f(dw_shift), _f(w_kn), _f(scale), _f(shift),
                                              int(relu_out), None if dr is None else dr.ptr, do.ptr, None))
    return do.download((B, H, W, N), np.uint16)


def op_conv3x3_c32(x_bits: np.ndarray, w_hwio: np.ndarray, scale=None, shift=None, device=0) -> np.ndarray:
    """entry_flow_conv1_2: x uint16 bf16 [B,H,W,32], w fp32 [3,3,32,64] -> uint16 bf16 [B,H,W,64] (BN + ReLU)."""
    B, H, W, Cc = x_bits.shape
    assert Cc == 32 and tuple(w_hwio.shape) == (3, 3, 32, 64)
    dx = DeviceBuffer.from_numpy(x_bits, device)
    do = DeviceBuffer(B * H * W * 64 * 2, device)
    _check(load_library().dlv3p_op_conv3x3_c32(device, dx.ptr, B, H, W, _f(w_hwio), _f(scale), _f(shift), do.ptr, None))
    return do.download((B, H, W, 64), np.uint16)


def op_stem_conv(img: np.ndarray, w_hwio: np.ndarray, scale=None, shift=None, device=0) -> np.ndarray:
    """entry_flow_conv1_1: uint8 (normalised on the device) or fp32 [B,H,W,3], w fp32 [3,3,3,32] -> uint16 bf16 [B,ceil(H/2),ceil(W/2),32]."""
    B, H, W, _ = img.shape
    f32 = img.dtype != np.uint8
    a = np.ascontiguousarray(img, np.float32 if f32 else np.uint8)
    dx = DeviceBuffer.from_numpy(a, device)
    Ho, Wo = -(-H // 2), -(-W // 2)
    do = DeviceBuffer(B * Ho * Wo * 32 * 2, device)
    _check(load_library().dlv3p_op_stem_conv(device, dx.ptr, IMG_F32 if f32 else IMG_U8, B, H, W, _f(w_hwio), _f(scale), _f(shift), do.ptr, None))
    return do.download((B, Ho, Wo, 32), np.uint16)


def op_resize_bilinear(x_bits: np.ndarray, ho: int, wo: int, device=0) -> np.ndarray:
    B, hi, wi, Cc = x_bits.shape
    dx = DeviceBuffer.from_numpy(x_bits, device)
    do = DeviceBuffer(B * ho * wo * Cc * 2, device)
    _check(load_library().dlv3p_op_resize_bilinear(device, dx.ptr, B, hi, wi, Cc, ho, wo, do.ptr, None))
    return do.download((B, ho, wo, Cc), np.uint16)


def op_resize_argmax(logits_planar: np.ndarray, ho: int, wo: int, device=0) -> np.ndarray:
    """logits fp32 [B,NC,hi,wi] planar -> uint8 [B,ho,wo]."""
    B, NC, hi, wi = logits_planar.shape
    dl = DeviceBuffer.from_numpy(np.ascontiguousarray(logits_planar, np.float32), device)
    do = DeviceBuffer(B * ho * wo, device)
    _check(load_library().dlv3p_op_resize_argmax(device, dl.ptr, B, NC, hi, wi, ho, wo, do.ptr, None))
    return do.download((B, ho, wo), np.uint8)


def op_confusion_matrix(pred: np.ndarray, gt: np.ndarray, num_classes: int, device=0, repeat: int = 1) -> np.ndarray:
    """generate_matrix (eval.py:368-373) on the device: uint8 label maps -> int64 [NC, NC]; `repeat` accumulates the
    same maps several times (the evaluation loop's running sum)."""
    p = np.ascontiguousarray(pred, np.uint8).reshape(-1)
    g = np.ascontiguousarray(gt, np.uint8).reshape(-1)
    dp, dg = DeviceBuffer.from_numpy(p, device), DeviceBuffer.from_numpy(g, device)
    dc = DeviceBuffer.from_numpy(np.zeros(num_classes * num_classes, np.uint64), device)
    for _ in range(repeat):
        _check(load_library().dlv3p_op_confusion_matrix(device, dp.ptr, dg.ptr, p.size, num_classes, dc.ptr, None))
    synchronize(device)
    return dc.download((num_classes, num_classes), np.uint64).astype(np.int64)


def op_normalize_image(image_u8: np.ndarray, out_bf16: bool = False, device=0) -> np.ndarray:
    """normalize_image (common/data_utils.py:403-416) on the device: uint8 -> float32 x/127.5 - 1 (or bf16 bit patterns as uint16)."""
    a = np.ascontiguousarray(image_u8, np.uint8)
    din = DeviceBuffer.from_numpy(a.reshape(-1), device)
    dout = DeviceBuffer(a.size * (2 if out_bf16 else 4), device)
    _check(load_library().dlv3p_op_normalize_image(device, din.ptr, a.size, dout.ptr, int(out_bf16), None))
    synchronize(device)
    return dout.download(a.shape, np.uint16 if out_bf16 else np.float32)


def op_denormalize_image(image_f32: np.ndarray, device=0) -> np.ndarray:
    """denormalize_image (common/data_utils.py:419-433) on the device: float32 -> uint8."""
    a = np.ascontiguousarray(image_f32, np.float32)
    din = DeviceBuffer.from_numpy(a.reshape(-1), device)
    dout = DeviceBuffer(a.size, device)
    _check(load_library().dlv3p_op_denormalize_image(device, din.ptr, a.size, dout.ptr, None))
    synchronize(device)
    return dout.download(a.shape, np.uint8)


def op_mask_resize(mask: np.ndarray, target_size: Tuple[int, int], device=0) -> np.ndarray:
    """mask_resize (common/data_utils.py:457-477): cv2 INTER_NEAREST resize of uint8 label maps [..., hi, wi] to target_size = (width, height)."""
    a = np.ascontiguousarray(mask, np.uint8)
    hi, wi = a.shape[-2:]
    wo, ho = int(target_size[0]), int(target_size[1])
    B = a.size // (hi * wi)
    din = DeviceBuffer.from_numpy(a.reshape(-1), device)
    dout = DeviceBuffer(B * ho * wo, device)
    _check(load_library().dlv3p_op_mask_resize_nearest(device, din.ptr, B, hi, wi, ho, wo, dout.ptr, None))
    synchronize(device)
    return dout.download(a.shape[:-2] + (ho, wo), np.uint8)


def op_resize_bicubic(image: np.ndarray, size_hw: Tuple[int, int], device=0) -> np.ndarray:
    """The resize of preprocess_image (common/data_utils.py:449: image.resize(model_input_shape[::-1], Image.BICUBIC)) on the device:
    uint8 images [..., H, W, C] -> [..., ho, wo, C], bit exact against Pillow."""
    a = np.ascontiguousarray(image, np.uint8)
    if a.ndim < 3:
        raise ValueError('op_resize_bicubic: expected [..., H, W, C] uint8')
    H, W, Cc = a.shape[-3:]
    ho, wo = int(size_hw[0]), int(size_hw[1])
    B = a.size // (H * W * Cc)
    din = DeviceBuffer.from_numpy(a.reshape(-1), device)
    dout = DeviceBuffer(B * ho * wo * Cc, device)
    _check(load_library().dlv3p_op_resize_bicubic_u8(device, din.ptr, B, H, W, Cc, ho, wo, dout.ptr, None))
    synchronize(device)
    return dout.download(a.shape[:-3] + (ho, wo, Cc), np.uint8)


def pil_bicubic_coeffs(in_size: int, out_size: int):
    """(bounds [out_size, 2], kk [out_size, ksize]) of one resampling axis as the library computes them on the host (no GPU needed)."""
    import math
    ks = int(math.ceil(2.0 * max(in_size / out_size, 1.0))) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ks), np.int32)
    ksize = C.c_int()
    _check(load_library().dlv3p_pil_bicubic_coeffs(in_size, out_size, bounds.ctypes.data_as(C.c_void_p), kk.ctypes.data_as(C.c_void_p), kk.size, C.byref(ksize)))
    assert ksize.value == ks
    return bounds, kk


def op_present_classes(labels: np.ndarray, device=0) -> List[List[int]]:
    """class_indexes of the native post-process (inference/MNN/deeplabSegment.cpp:171-172) for uint8 label maps [B, H, W] (or [H, W]):
    per image the classes != 0 that occur, in order of first appearance in raster order."""
    a = np.ascontiguousarray(labels, np.uint8)
    if a.ndim == 2:
        a = a[None]
    B, n = a.shape[0], int(a.shape[1] * a.shape[2])
    din = DeviceBuffer.from_numpy(a.reshape(-1), device)
    dfirst = DeviceBuffer(B * 256 * 4, device)
    _check(load_library().dlv3p_op_present_classes(device, din.ptr, B, n, dfirst.ptr, None))
    synchronize(device)
    first = dfirst.download((B, 256), np.uint32)
    return [[int(c) for c in np.argsort(first[b], kind='stable') if c != 0 and first[b, c] != 0xFFFFFFFF] for b in range(B)]


def bn_stats(x_ptr: int, M: int, Cc: int, stats_ptr: int, scratch_ptr: int, stream=None, device=0) -> None:
    """Device pointers in and out, asynchronous: stats fp32 [2C+1] = sum_x | sum_x2 | rows (all-reduce it across replicas)."""
    _check(load_library().dlv3p_op_bn_stats(device, x_ptr, M, Cc, stats_ptr, scratch_ptr, stream))


def bn_apply(x_ptr: int, M: int, Cc: int, stats_ptr: int, gamma_ptr: int, beta_ptr: int, eps: float, relu: bool, y_ptr: int,
             stream=None, device=0) -> None:
    _check(load_library().dlv3p_op_bn_apply(device, x_ptr, M, Cc, stats_ptr, gamma_ptr, beta_ptr, eps, int(relu), y_ptr, stream))


def bn_scratch_bytes(Cc: int) -> int:
    return int(load_library().dlv3p_op_bn_scratch_bytes(Cc))


def op_bn_train(x_bits: np.ndarray, gamma: np.ndarray, beta: np.ndarray, eps: float = 1e-5, relu: bool = True, device=0):
    """Single-replica training-mode BN on host arrays: returns (y bf16 bits [M,C], stats fp32 [2C+1])."""
    M, Cc = x_bits.shape
    dx = DeviceBuffer.from_numpy(x_bits, device)
    dst = DeviceBuffer((2 * Cc + 1) * 4, device)
    dsc = DeviceBuffer(bn_scratch_bytes(Cc), device)
    dg = DeviceBuffer.from_numpy(np.ascontiguousarray(gamma, np.float32), device)
    db = DeviceBuffer.from_numpy(np.ascontiguousarray(beta, np.float32), device)
    dy = DeviceBuffer(M * Cc * 2, device)
    bn_stats(dx.ptr, M, Cc, dst.ptr, dsc.ptr, None, device)
    bn_apply(dx.ptr, M, Cc, dst.ptr, dg.ptr, db.ptr, eps, relu, dy.ptr, None, device)
    synchronize(device)
    return dy.download((M, Cc), np.uint16), dst.download((2 * Cc + 1,), np.float32)


def op_bb_time(op: int, dims: Sequence[int], iters: int = 20, flags: int = 0, device: int = 0) -> float:
    """ms per launch of one BACKBONE operator on synthetic data (benchmark aid, see include/dlv3p_model.h)."""
    arr = (C.c_int64 * len(dims))(*dims)
    ms = C.c_float()
    _check(load_library().dlv3p_op_bb_time(device, op, arr, len(dims), iters, flags, C.byref(ms)))
    return ms.value


def op_time(op: int, dims: Sequence[int], iters: int = 20, flags: int = 0, device: int = 0) -> float:
    """ms per launch of one operator on synthetic data (benchmark aid, see include/dlv3p.h)."""
    arr = (C.c_int64 * len(dims))(*dims)
    ms = C.c_float()
    _check(load_library().dlv3p_op_time(device, op, arr, len(dims), iters, flags, C.byref(ms)))
    return ms.value
