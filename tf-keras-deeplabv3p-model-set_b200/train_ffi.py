"""ctypes binding of the training-step operators of libdlv3p.so (include/dlv3p_train.h).

Plain device pointers (ints) and sizes in, nothing out: every call is asynchronous on the CUDA stream given.  There is no
fallback: a missing library or a non-sm_100 device raises Dlv3pError."""
from __future__ import annotations

import ctypes as C

from . import ffi

_vp, _i, _i64, _sz, _f, _u32 = C.c_void_p, C.c_int, C.c_int64, C.c_size_t, C.c_float, C.c_uint32

# every symbol include/dlv3p_train.h declares: (name, restype, argtypes)
SYMBOLS = [
    ('dlv3p_train_gemm_partial_bytes', _sz, [_i64, _i, _i]),
    ('dlv3p_train_gemm_nt', _i, [_i, _vp, _i64, _vp, _i64, _i64, _i, _i64, _vp, _i64, _i, _i, _vp, _vp]),
    ('dlv3p_train_gemm_tn', _i, [_i, _vp, _i64, _vp, _i64, _i64, _i, _i64, _vp, _i64, _i, _i, _vp, _vp]),
    ('dlv3p_train_transpose', _i, [_i, _vp, _i64, _i, _i64, _vp, _i64, _vp]),
    ('dlv3p_train_bn_apply', _i, [_i, _vp, _i64, _i, _vp, _vp, _vp, _f, _i, _vp, _i64, _vp]),
    ('dlv3p_train_scratch_bytes', _sz, [_i]),
    ('dlv3p_train_bn_bwd_stats', _i, [_i, _vp, _i64, _vp, _i64, _vp, _i64, _i, _vp, _f, _i, _vp, _vp, _vp]),
    ('dlv3p_train_bn_bwd_apply', _i, [_i, _vp, _i64, _vp, _i64, _vp, _i64, _i, _vp, _vp, _vp, _f, _i, _vp, _vp]),
    ('dlv3p_train_depthwise', _i, [_i, _vp, _i, _i, _i, _i, _i, _vp, _i, _vp, _vp]),
    ('dlv3p_train_depthwise_wgrad', _i, [_i, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    ('dlv3p_train_resize', _i, [_i, _vp, _i, _i, _i, _i, _i, _i, _vp, _i64, _vp]),
    ('dlv3p_train_resize_bwd', _i, [_i, _vp, _i64, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    ('dlv3p_train_loss_scratch_bytes', _sz, []),
    ('dlv3p_train_softmax_ce', _i, [_i, _vp, _i64, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp, _vp]),
    ('dlv3p_train_softmax_loss', _i, [_i, _vp, _i64, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _f, _i, _vp, _f, _f, _vp, _vp, _vp, _vp]),
    ('dlv3p_train_resize_bwd_planar_scratch_bytes', _sz, [_i, _i, _i, _i]),
    ('dlv3p_train_resize_bwd_planar', _i, [_i, _vp, _i, _i, _i, _i, _i, _i, _vp, _i64, _vp, _vp]),
    ('dlv3p_train_rows_reduce', _i, [_i, _vp, _i64, _i, _i, _i, _f, _vp, _i, _vp]),
    ('dlv3p_train_bcast_rows', _i, [_i, _vp, _i, _i, _i, _f, _vp, _i64, _i, _vp]),
    ('dlv3p_train_add', _i, [_i, _vp, _vp, _vp, _i64, _vp]),
    ('dlv3p_train_dropout', _i, [_i, _vp, _vp, _i64, _u32, _vp, _f, _vp]),
    ('dlv3p_train_sgd', _i, [_i, _vp, _vp, _vp, _i64, _f, _f, _f, _f, _vp]),
    ('dlv3p_train_cast_bf16', _i, [_i, _vp, _vp, _i64, _vp]),
    ('dlv3p_p2p_create', _i, [_i, _i, _i, _sz, C.POINTER(_vp), C.c_char_p]),
    ('dlv3p_p2p_connect', _i, [_vp, C.c_char_p]),
    ('dlv3p_p2p_destroy', None, [_vp]),
    ('dlv3p_p2p_payload', _vp, [_vp, _sz]),
    ('dlv3p_p2p_allreduce', _i, [_vp, _i, _sz, _i, _vp, _vp]),
    ('dlv3p_p2p_advance', _i, [_vp, _vp]),
    ('dlv3p_trainer_create', _i, [_vp, _i, C.POINTER(_vp), C.c_char_p]),
    ('dlv3p_trainer_connect', _i, [_vp, C.c_char_p]),
    ('dlv3p_trainer_destroy', None, [_vp]),
    ('dlv3p_trainer_last_error', C.c_char_p, [_vp]),
    ('dlv3p_trainer_set_weight', _i, [_vp, C.c_char_p, C.c_char_p, C.POINTER(C.c_float), _i64]),
    ('dlv3p_trainer_commit_weights', _i, [_vp]),
    ('dlv3p_trainer_get', _i, [_vp, _i, C.c_char_p, C.c_char_p, C.POINTER(C.c_float), _i64]),
    ('dlv3p_trainer_set_class_weights', _i, [_vp, C.POINTER(C.c_float), _i]),
    ('dlv3p_trainer_set_hyper', _i, [_vp, _f, _f, _f]),
    ('dlv3p_trainer_step', _i, [_vp, _vp, _vp, _vp, _i, _vp]),
    ('dlv3p_trainer_forward_backward', _i, [_vp, _vp, _vp, _vp, _vp]),
    ('dlv3p_trainer_all_reduce_gradients', _i, [_vp, _vp]),
    ('dlv3p_trainer_apply_gradients', _i, [_vp, _vp]),
    ('dlv3p_trainer_loss', _i, [_vp, C.POINTER(_f), C.POINTER(_f)]),
    ('dlv3p_trainer_read', _i, [_vp, C.c_char_p, C.POINTER(C.c_float), _i64]),
    ('dlv3p_trainer_num_params', _i, [_vp, C.POINTER(_i64), C.POINTER(_i64)]),
    ('dlv3p_trainer_counters', _i, [_vp, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i)]),
    ('dlv3p_trainer_weights_digest', _i, [_vp, C.POINTER(C.c_uint64)]),
]


class TrainerConfig(C.Structure):
    """dlv3p_trainer_config (include/dlv3p_train.h)."""
    _fields_ = [(n, C.c_int32) for n in ('B', 'H', 'W', 'OS', 'Cin', 'Cskip', 'NC', 'world', 'rank', 'global_batch', 'ignore_index', 'loss_kind')] + \
               [('seed', C.c_uint32)] + [(n, C.c_float) for n in ('lr', 'momentum', 'l2', 'bn_momentum', 'eps', 'dropout', 'focal_gamma', 'focal_alpha')] + [('lite', C.c_int32)]

_typed = False


def lib() -> C.CDLL:
    global _typed
    L = ffi.load_library()
    if not _typed:
        for name, res, args in SYMBOLS:
            fn = getattr(L, name)          # AttributeError == header / library mismatch
            fn.restype = res
            fn.argtypes = args
        _typed = True
    return L


def call(name: str, *args) -> None:
    """Invoke an int-returning entry point; raises Dlv3pError with the library's message on a negative status."""
    ffi._check(getattr(lib(), name)(*args))


def scratch_bytes(C_: int) -> int:
    return int(lib().dlv3p_train_scratch_bytes(C_))


def loss_scratch_bytes() -> int:
    return int(lib().dlv3p_train_loss_scratch_bytes())


def gemm_partial_bytes(M: int, N: int, splits: int) -> int:
    return int(lib().dlv3p_train_gemm_partial_bytes(M, N, splits))


def dropout_keep_mask(n: int, seed: int, rate: float):
    """The mask dlv3p_train_dropout applies, restated in numpy (host-side bookkeeping and tests): element e is kept iff
    fmix32(e * 0x9E3779B1 + seed) >= rate * 2^32."""
    import numpy as np
    e = np.arange(n, dtype=np.uint64)
    h = (e * 0x9E3779B1 + (seed & 0xFFFFFFFF)) & 0xFFFFFFFF
    h ^= h >> 16
    h = (h * 0x85EBCA6B) & 0xFFFFFFFF
    h ^= h >> 13
    h = (h * 0xC2B2AE35) & 0xFFFFFFFF
    h ^= h >> 16
    return h >= int(float(rate) * 4294967296.0)


class P2pExchange:
    """One replica's end of the peer-memory exchange (dlv3p_p2p_*): create -> hand `handle` (64 bytes) to every peer ->
    connect(all handles in rank order) -> allreduce(slot, off, n, out_ptr, stream) ... advance(stream) once per step."""

    def __init__(self, device: int, world: int, rank: int, payload_floats: int):
        self.world, self.rank = world, rank
        h = C.c_void_p()
        buf = C.create_string_buffer(64)
        ffi._check(lib().dlv3p_p2p_create(device, world, rank, payload_floats, C.byref(h), buf))
        self.handle_ptr, self.handle = h, buf.raw

    def connect(self, handles) -> None:
        blob = b''.join(bytes(x) for x in handles)
        assert len(blob) == 64 * self.world
        ffi._check(lib().dlv3p_p2p_connect(self.handle_ptr, blob))

    def payload(self, off: int = 0) -> int:
        return int(lib().dlv3p_p2p_payload(self.handle_ptr, off))

    def allreduce(self, slot: int, off: int, n: int, out_ptr: int, stream) -> None:
        ffi._check(lib().dlv3p_p2p_allreduce(self.handle_ptr, slot, off, n, out_ptr, stream))

    def advance(self, stream) -> None:
        ffi._check(lib().dlv3p_p2p_advance(self.handle_ptr, stream))

    def close(self) -> None:
        if getattr(self, 'handle_ptr', None):
            lib().dlv3p_p2p_destroy(self.handle_ptr)
            self.handle_ptr = None
