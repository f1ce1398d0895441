"""Training step of the DeepLabV3+ head on B200 (BASELINE cfg 5): forward in training mode (SyncBatchNormalization with
batch statistics, Dropout active), sparse softmax cross-entropy with ignore_index, backward, gradient all-reduce and
SGD-momentum update — the whole step lives behind the C ABI (include/dlv3p_train.h, dlv3p_trainer_*): this module is a thin
ctypes host of it, one process per GPU.  No PyTorch here: device pointers in (anything with .data_ptr(), a
__cuda_array_interface__, an ffi.DeviceBuffer or an int), numpy out.

Reference: train.py:143-169 (model.compile + fit under tf.distribute.MirroredStrategy: per-replica forward / backward,
SyncBN statistics and gradients summed over the replicas), graph deeplabv3p/models/layers.py:74-219 + model.py:75-86,
loss deeplabv3p/loss.py:121-156 (SparseCategoricalCrossEntropy, ignore_index 255, Keras mean over every pixel), optimizer
common/model_utils.py:122-123 (SGD, momentum 0.9, lr 1e-2 default train.py:280-286), regulariser l2(2e-5) on the conv
kernels and biases (layers.py:12-21; inert for depthwise kernels, :24-31), BN momentum 0.99 / eps 1e-5.

The two exchanges of the step — the per-layer SyncBN statistics (forward: sum x | sum x^2 | n, backward: sum g | sum g*xhat) and ONE
all-reduce of the flat fp32 gradient bucket — run inside the library over NVLink peer memory (csrc/p2p_exchange.cuh); the host's
only job is to hand the replicas' 64-byte CUDA-IPC handles around once (`exchange_handles`, any channel: torch.distributed if the
program already runs it, MPI, a file).

Scope: both head variants of the reference — ASPP_block + Decoder_block + tail (the Xception / ResNet50 / MobileNetV2 / V3 models) and,
with lite=True, ASPP_Lite_block + tail without a decoder (the *_lite models, deeplabv3p_mobilenetv2.py:326-331; no skip input).
"""
from __future__ import annotations

import ctypes as C
import sys
from typing import Callable, Dict, Optional, Sequence, Tuple

import numpy as np

from . import ffi
from . import train_ffi as tf_
from .head import atrous_rates

L2_COEF = 2e-5       # layers.py:12
BN_MOMENTUM = 0.99   # Keras BatchNormalization default, kept by CustomBatchNormalization (layers.py:63-70)


def _rup(v: int, a: int) -> int:
    return (v + a - 1) // a * a


def dropout_seed(base_seed: int, step: int, rank: int) -> int:
    """Seed of the Dropout mask replica `rank` draws at step `step` (tests rebuild the mask with train_ffi.dropout_keep_mask)."""
    return (base_seed * 0x9E3779B1 + step * 0x85EBCA6B + rank * 0xC2B2AE35 + 0x27D4EB2F) & 0xFFFFFFFF


class TrainLayout:
    """Host-side layout of the training state (no CUDA needed): where every parameter / gradient lives in the flat fp32 buffers, where
    every BN layer's SyncBN vector lives, and which contiguous spans each collective of the step exchanges.  The world_size-2 gloo
    test on CPU (tests/test_dist_cpu.py) drives exactly these spans."""

    def __init__(self, Cin: int, Cskip: int, NC: int, lite: bool = False):
        self.Cin, self.Cs, self.NC, self.lite = Cin, (0 if lite else Cskip), NC, bool(lite)
        self.NCp = _rup(NC, 8)
        if self.lite:      # ASPP_Lite_block (layers.py:166-196): image pooling + one 1x1 branch, no depthwise convolutions, no decoder
            self.FWD_GROUPS = [['image_pooling_BN', 'aspp0_BN'], ['concat_projection_BN']]
            self.BWD_GROUPS = [['concat_projection_BN'], ['aspp0_BN', 'image_pooling_BN']]
        self._layout_params()

    def stats_span(self, group) -> Tuple[int, int]:
        """[begin, end) in the forward-statistics buffer of one group of independent BN layers: ONE all-reduce."""
        o1, c1 = self.stat_off[group[-1]]
        return self.stat_off[group[0]][0], o1 + 2 * c1 + 1

    def bn_grad_span(self, group) -> Tuple[int, int]:
        """[begin, end) in the gradient buffer of the d(beta) | d(gamma) vectors of one backward group: ONE all-reduce."""
        g1, (c1,) = self.off[(group[-1], 'beta')]
        return self.off[(group[0], 'beta')][0], g1 + 2 * c1

    def bucket_span(self) -> Tuple[int, int]:
        """[begin, end) of the gradient bucket all-reduced once per step (1x1 kernels, classifier bias, depthwise kernels)."""
        return 0, self.endB

    # ------------------------------------------------------------------------------------------------ parameters
    def _conv_specs(self):
        Cin, Cs, NCp = self.Cin, self.Cs, self.NCp
        if self.lite:
            return [('image_pooling', Cin, 256), ('aspp0', Cin, 256), ('concat_projection', 512, 256), ('conv_upsample', 256, NCp)]
        return [('image_pooling', Cin, 256), ('aspp0', Cin, 256), ('aspp1_pointwise', Cin, 256), ('aspp2_pointwise', Cin, 256),
                ('aspp3_pointwise', Cin, 256), ('concat_projection', 1280, 256), ('feature_projection0', Cs, 48),
                ('decoder_conv0_pointwise', 304, 256), ('decoder_conv1_pointwise', 256, 256), ('conv_upsample', 256, NCp)]

    def _dw_specs(self):
        if self.lite:
            return []
        return [('aspp1_depthwise', self.Cin), ('aspp2_depthwise', self.Cin), ('aspp3_depthwise', self.Cin),
                ('decoder_conv0_depthwise', 304), ('decoder_conv1_depthwise', 256)]

    def _bn_specs(self):
        Cin = self.Cin
        if self.lite:
            return [('image_pooling_BN', 256), ('aspp0_BN', 256), ('concat_projection_BN', 256)]
        return [('image_pooling_BN', 256), ('aspp0_BN', 256),
                ('aspp1_depthwise_BN', Cin), ('aspp1_pointwise_BN', 256), ('aspp2_depthwise_BN', Cin), ('aspp2_pointwise_BN', 256),
                ('aspp3_depthwise_BN', Cin), ('aspp3_pointwise_BN', 256), ('concat_projection_BN', 256), ('feature_projection0_BN', 48),
                ('decoder_conv0_depthwise_BN', 304), ('decoder_conv0_pointwise_BN', 256), ('decoder_conv1_depthwise_BN', 256),
                ('decoder_conv1_pointwise_BN', 256)]

    # BN layers grouped by data dependence (one all-reduce per group)
    FWD_GROUPS = [['image_pooling_BN', 'aspp0_BN', 'aspp1_depthwise_BN', 'aspp2_depthwise_BN', 'aspp3_depthwise_BN', 'feature_projection0_BN'],
                  ['aspp1_pointwise_BN', 'aspp2_pointwise_BN', 'aspp3_pointwise_BN'], ['concat_projection_BN'],
                  ['decoder_conv0_depthwise_BN'], ['decoder_conv0_pointwise_BN'], ['decoder_conv1_depthwise_BN'], ['decoder_conv1_pointwise_BN']]
    BWD_GROUPS = [['decoder_conv1_pointwise_BN'], ['decoder_conv1_depthwise_BN'], ['decoder_conv0_pointwise_BN'], ['decoder_conv0_depthwise_BN'],
                  ['feature_projection0_BN', 'concat_projection_BN'],
                  ['aspp0_BN', 'aspp1_pointwise_BN', 'aspp2_pointwise_BN', 'aspp3_pointwise_BN', 'image_pooling_BN'],
                  ['aspp1_depthwise_BN', 'aspp2_depthwise_BN', 'aspp3_depthwise_BN']]

    def _layout_params(self):
        """Flat fp32 layout: [A: 1x1 kernels [K,N] + classifier bias (l2-regularised)] [B: depthwise taps [9,C]]
        [C: per BN layer beta | gamma].  A and B are all-reduced; C's gradients come out of the SyncBN backward exchange
        already summed over the global batch."""
        off = 0
        self.off: Dict[Tuple[str, str], Tuple[int, Tuple[int, ...]]] = {}
        for name, K, N in self._conv_specs():
            self.off[(name, 'kernel')] = (off, (K, N))
            off = _rup(off + K * N, 8)
        self.off[('conv_upsample', 'bias')] = (off, (self.NCp,))
        off = _rup(off + self.NCp, 8)
        self.endA = off
        for name, Cc in self._dw_specs():
            self.off[(name, 'depthwise_kernel')] = (off, (9, Cc))
            off = _rup(off + 9 * Cc, 8)
        self.endB = off
        # BN parameters / gradients in the order the BACKWARD pass meets the layers, and forward statistics in the order the
        # FORWARD pass does: BN layers whose inputs are independent sit next to each other and share ONE all-reduce of their
        # SyncBN vectors (14 collectives per step instead of 28)
        chan = dict(self._bn_specs())
        for name in [n for grp in self.BWD_GROUPS for n in grp]:
            Cc = chan[name]
            self.off[(name, 'beta')] = (off, (Cc,))
            self.off[(name, 'gamma')] = (off + Cc, (Cc,))
            off = _rup(off + 2 * Cc, 8)
        self.nparams = off
        soff = 0
        self.stat_off: Dict[str, Tuple[int, int]] = {}
        for name in [n for grp in self.FWD_GROUPS for n in grp]:
            self.stat_off[name] = (soff, chan[name])
            soff = _rup(soff + 2 * chan[name] + 1, 4)
        self.nstats = soff
        self.nbn = sum(c for _, c in self._bn_specs())


def _ptr(x) -> int:
    """Device pointer of a torch tensor / cupy array / ffi.DeviceBuffer / int."""
    if isinstance(x, int):
        return x
    if hasattr(x, 'data_ptr'):
        return int(x.data_ptr())
    if hasattr(x, '__cuda_array_interface__'):
        return int(x.__cuda_array_interface__['data'][0])
    if hasattr(x, 'ptr'):
        return int(x.ptr)
    raise TypeError('expected a device pointer (int, .data_ptr(), __cuda_array_interface__ or ffi.DeviceBuffer), got %r' % type(x))


def _running_process_group(process_group=None):
    """torch.distributed, if the HOST PROGRAM already imported and initialised it (this package never imports torch): used only
    to pass the 64-byte IPC handles around and to learn world size / rank."""
    dist = sys.modules.get('torch.distributed')
    if dist is None or not (dist.is_available() and dist.is_initialized()):
        return None
    return dist


class _HostArray:
    """numpy snapshot of a device buffer with the few tensor-like accessors callers chain (.float().cpu().numpy())."""

    def __init__(self, a: np.ndarray):
        self._a = a

    def float(self):
        return self

    def cpu(self):
        return self

    def numpy(self) -> np.ndarray:
        return self._a


class _Buffers:
    def __init__(self, tr):
        self._tr = tr

    def __getitem__(self, name: str) -> _HostArray:
        return _HostArray(self._tr.tensor(name))


class HeadTrainer(TrainLayout):
    """One replica of the data-parallel training step: a handle on a dlv3p_trainer.  Feed bf16 NHWC features and uint8 labels
    already resident in HBM (the backbone is outside this path; d(loss)/d(feat), d(loss)/d(skip) are left for it in the buffers
    'dfeat' / 'dskip')."""

    def __init__(self, B: int, H: int, W: int, OS: int, Cin: int, Cskip: int, NC: int, weights: Dict[Tuple[str, str], np.ndarray],
                 device: int = 0, lr: float = 1e-2, momentum: float = 0.9, l2: float = L2_COEF, bn_momentum: float = BN_MOMENTUM,
                 eps: float = 1e-5, dropout: float = 0.5, seed: int = 0, ignore_index: int = 255, global_batch: Optional[int] = None,
                 process_group=None, graph: bool = True, loss: str = 'crossentropy', class_weights=None,
                 focal_gamma: float = 2.0, focal_alpha: float = 0.25, world: Optional[int] = None, rank: Optional[int] = None,
                 exchange_handles: Optional[Callable[[bytes], Sequence[bytes]]] = None, stream: int = 0, lite: bool = False, **_ignored):
        self.lib = tf_.lib()
        if ffi.device_count() == 0:
            raise ffi.Dlv3pError(-2, 'HeadTrainer needs a CUDA device: there is no CPU path')
        self.rates = atrous_rates(OS)
        self.B, self.H, self.W, self.OS, self.Cin, self.Cs, self.NC = B, H, W, OS, Cin, Cskip, NC
        self.h, self.w = -(-H // OS), -(-W // OS)
        self.hs, self.ws = (self.h, self.w) if lite else (-(-H // 4), -(-W // 4))      # lite: no decoder, logits at the feature resolution
        self.M1, self.M2 = B * self.h * self.w, B * self.hs * self.ws
        if Cin % 8 or (not lite and Cskip % 8) or self.M1 % 8 or self.M2 % 8:
            raise ffi.Dlv3pError(-1, 'HeadTrainer: Cin, Cskip and the pixel counts per replica must be multiples of 8')
        TrainLayout.__init__(self, Cin, Cskip, NC, lite)
        # train.py:114-138: --loss crossentropy (optionally class weighted, --weighted_type balanced) | focal (ignores the weights)
        if loss not in ('crossentropy', 'focal'):
            raise ValueError('invalid loss type {}'.format(loss))
        cw = None if class_weights is None else np.ascontiguousarray(class_weights, np.float32).reshape(-1)
        if cw is not None and cw.size != NC:
            raise ValueError('class_weights must have one entry per class')
        dist = None
        if world is None:
            dist = _running_process_group(process_group)
            world = dist.get_world_size(process_group) if dist else 1
            rank = dist.get_rank(process_group) if dist else 0
            if dist and world > 1 and exchange_handles is None:
                def exchange_handles(h, _d=dist, _g=process_group, _w=world):
                    out = [None] * _w
                    _d.all_gather_object(out, h, group=_g)
                    return out
        self.world, self.rank = int(world), int(rank or 0)
        if self.world > 1 and exchange_handles is None:
            raise ValueError('world > 1 needs exchange_handles: a callable that takes this replica\'s IPC handle (bytes) and returns all of them in rank order')
        self.device, self.dev = device, device
        self.stream = stream
        self.use_graph = graph
        self.global_batch = global_batch if global_batch is not None else B * self.world
        self.seed, self.drop_rate, self.ignore, self.eps, self.bn_momentum = seed, dropout, ignore_index, eps, bn_momentum
        cfg = tf_.TrainerConfig(B=B, H=H, W=W, OS=OS, Cin=Cin, Cskip=Cskip, NC=NC, world=self.world, rank=self.rank, global_batch=self.global_batch,
                                ignore_index=ignore_index, loss_kind=2 if loss == 'focal' else (1 if cw is not None else 0), seed=seed & 0xFFFFFFFF,
                                lr=lr, momentum=momentum, l2=l2, bn_momentum=bn_momentum, eps=eps, dropout=dropout, focal_gamma=focal_gamma,
                                focal_alpha=focal_alpha, lite=int(bool(lite)))
        self._hyper = [lr, momentum, l2]
        h = C.c_void_p()
        buf = C.create_string_buffer(64)
        ffi._check(self.lib.dlv3p_trainer_create(C.byref(cfg), device, C.byref(h), buf))
        self.handle = h
        if self.world > 1:
            handles = list(exchange_handles(buf.raw))
            if len(handles) != self.world:
                raise ValueError('exchange_handles returned %d handles for a world of %d' % (len(handles), self.world))
            self._check(self.lib.dlv3p_trainer_connect(self.handle, b''.join(bytes(x) for x in handles)))
            if dist:
                dist.barrier(group=process_group)
        if cw is not None:
            self._check(self.lib.dlv3p_trainer_set_class_weights(self.handle, cw.ctypes.data_as(C.POINTER(C.c_float)), NC))
        self.T = _Buffers(self)
        self.set_weights(weights)

    def _check(self, status: int) -> int:
        if status < 0:
            msg = self.lib.dlv3p_trainer_last_error(self.handle)
            raise ffi.Dlv3pError(status, msg.decode() if msg else '')
        return status

    # lr / momentum / l2 travel by value into the SGD kernels, so a captured CUDA graph has them baked in: changing one
    # (ReduceLROnPlateau, the cosine / poly decay schedules of the reference's train.py:49-66, :192-215) makes the library drop the
    # captured graph; the next train_step re-captures with the new value.
    def _hyper_prop(i):
        def get(self):
            return self._hyper[i]

        def set_(self, v):
            self._hyper[i] = float(v)
            self._check(self.lib.dlv3p_trainer_set_hyper(self.handle, *self._hyper))
        return property(get, set_)

    lr, momentum, l2 = _hyper_prop(0), _hyper_prop(1), _hyper_prop(2)
    del _hyper_prop

    # ------------------------------------------------------------------------------------------------ weights
    def _specs(self):
        """(layer, variable, trainer-layout shape, Keras shape)."""
        out = []
        for name, K, N in self._conv_specs():
            n = self.NC if name == 'conv_upsample' else N
            out.append((name, 'kernel', (K, n), (1, 1, K, n)))
        out.append(('conv_upsample', 'bias', (self.NC,), (self.NC,)))
        for name, Cc in self._dw_specs():
            out.append((name, 'depthwise_kernel', (9, Cc), (3, 3, Cc, 1)))
        for name, Cc in self._bn_specs():
            for v in ('gamma', 'beta', 'moving_mean', 'moving_variance'):
                out.append((name, v, (Cc,), (Cc,)))
        return out

    def set_weights(self, W: Dict[Tuple[str, str], np.ndarray]) -> None:
        """Keras-layout arrays keyed (layer, variable) or "layer/variable" (the npz of tools/h5_to_npz.py) — the inventory of
        dlv3p_weight_info / SURVEY §8(b); `logits_semantic` is accepted for `conv_upsample`."""
        W = {(tuple(k.split('/', 1)) if isinstance(k, str) else tuple(k)): v for k, v in W.items()}
        W = {(('conv_upsample', k[1]) if k[0] == 'logits_semantic' else k): v for k, v in W.items()}
        fp = C.POINTER(C.c_float)
        for layer, var, shape, _ in self._specs():
            a = np.ascontiguousarray(np.asarray(W[(layer, var)], np.float32).reshape(shape))
            self._check(self.lib.dlv3p_trainer_set_weight(self.handle, layer.encode(), var.encode(), a.ctypes.data_as(fp), a.size))
        self._check(self.lib.dlv3p_trainer_commit_weights(self.handle))

    def _get(self, which: int, keras_shapes: bool, with_moving: bool) -> Dict[Tuple[str, str], np.ndarray]:
        fp = C.POINTER(C.c_float)
        out = {}
        for layer, var, shape, kshape in self._specs():
            if var.startswith('moving') and not with_moving:
                continue
            a = np.empty(shape, np.float32)
            self._check(self.lib.dlv3p_trainer_get(self.handle, which, layer.encode(), var.encode(), a.ctypes.data_as(fp), a.size))
            out[(layer, var)] = a.reshape(kshape) if keras_shapes else a
        return out

    def get_weights(self) -> Dict[Tuple[str, str], np.ndarray]:
        """Current fp32 master weights and moving statistics in the Keras layout (what model.save would write)."""
        return self._get(0, True, True)

    def get_grads(self) -> Dict[Tuple[str, str], np.ndarray]:
        return self._get(1, False, False)

    def get_velocity(self) -> Dict[Tuple[str, str], np.ndarray]:
        return self._get(2, False, False)

    # ------------------------------------------------------------------------------------------------ the step
    def forward_backward(self, feat, skip, labels) -> None:
        """feat bf16 [B,h,w,Cin], skip bf16 [B,hs,ws,Cs], labels uint8 [B,H,W] (device memory).  Leaves this replica's share of the
        global mean loss, all weight gradients (BN gradients global, the rest per replica until all_reduce_gradients) and
        d(loss)/d(feat), d(loss)/d(skip) in the buffers 'dfeat' / 'dskip'."""
        self._check(self.lib.dlv3p_trainer_forward_backward(self.handle, _ptr(feat), 0 if self.lite else _ptr(skip), _ptr(labels), self.stream))

    def all_reduce_gradients(self) -> None:
        """ONE exchange of the flat fp32 bucket holding every 1x1 kernel, the classifier bias and every depthwise kernel
        (MirroredStrategy semantics, train.py:143-158: the loss is normalised by the GLOBAL batch, so the sum is the gradient of
        the global mean loss).  BN gradients are already global."""
        self._check(self.lib.dlv3p_trainer_all_reduce_gradients(self.handle, self.stream))

    def apply_gradients(self) -> None:
        self._check(self.lib.dlv3p_trainer_apply_gradients(self.handle, self.stream))

    def train_step(self, feat, skip, labels) -> None:
        """One optimizer step (fit's train_step): forward, loss, backward, gradient exchange, SGD update.  Asynchronous; read the
        loss with .loss() (synchronises).  The first call launches the ~190 kernels one by one; with graph=True the second call
        captures the whole step into ONE CUDA graph inside the library that later calls replay."""
        self._check(self.lib.dlv3p_trainer_step(self.handle, _ptr(feat), 0 if self.lite else _ptr(skip), _ptr(labels), int(self.use_graph), self.stream))

    def loss(self) -> float:
        """Global mean loss of the last step (synchronises)."""
        v, n = C.c_float(), C.c_float()
        self._check(self.lib.dlv3p_trainer_loss(self.handle, C.byref(v), C.byref(n)))
        return float(v.value)

    def tensor(self, name: str) -> np.ndarray:
        """A named activation / gradient buffer as fp32 numpy ('dfeat' [B*h*w, Cin], 'dskip' [B*hs*ws, Cskip], 'logits' [B*hs*ws, NCp])."""
        shapes = {'dfeat': (self.M1, self.Cin), 'dskip': (self.M2, self.Cs), 'logits': (self.M2, self.NCp), 'yproj': (self.M1, 256),
                  'y0': (self.M2, 256), 'y1': (self.M2, 256), 'dlow': (self.M2, self.NCp), 'concat': (self.M1, 512 if self.lite else 1280)}
        a = np.empty(shapes[name], np.float32)
        self._check(self.lib.dlv3p_trainer_read(self.handle, name.encode(), a.ctypes.data_as(C.POINTER(C.c_float)), a.size))
        return a

    def _counters(self):
        a, b, g = C.c_int64(), C.c_int64(), C.c_int()
        self._check(self.lib.dlv3p_trainer_counters(self.handle, C.byref(a), C.byref(b), C.byref(g)))
        return a.value, b.value, bool(g.value)

    @property
    def launches(self) -> int:
        return self._counters()[0]

    @property
    def step_count(self) -> int:
        return self._counters()[1]

    @property
    def graph_captured(self) -> bool:
        return self._counters()[2]

    def weights_digest(self) -> int:
        d = C.c_uint64()
        self._check(self.lib.dlv3p_trainer_weights_digest(self.handle, C.byref(d)))
        return int(d.value)

    def comm_backend(self) -> str:
        return 'none' if self.world == 1 else 'NVLink peer memory (dlv3p_p2p): one-shot all-reduces of the SyncBN vectors, two-shot all-reduce of the gradient bucket; no NCCL'

    def close(self) -> None:
        if getattr(self, 'handle', None):
            self.lib.dlv3p_trainer_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
