"""Training step of the DeepLabV3+ head on B200 (BASELINE cfg 5): forward in training mode (SyncBatchNormalization with
batch statistics, Dropout active), sparse softmax cross-entropy with ignore_index, backward, gradient all-reduce and
SGD-momentum update — every operator a libdlv3p.so kernel (include/dlv3p_train.h), one process per GPU.

Reference: train.py:143-169 (model.compile + fit under tf.distribute.MirroredStrategy: per-replica forward / backward,
SyncBN statistics and gradients summed over the replicas), graph deeplabv3p/models/layers.py:74-219 + model.py:75-86,
loss deeplabv3p/loss.py:121-156 (SparseCategoricalCrossEntropy, ignore_index 255, Keras mean over every pixel), optimizer
common/model_utils.py:122-123 (SGD, momentum 0.9, lr 1e-2 default train.py:280-286), regulariser l2(2e-5) on the conv
kernels and biases (layers.py:12-21; inert for depthwise kernels, :24-31), BN momentum 0.99 / eps 1e-5.

torch is plumbing only: device memory (tensors), the CUDA stream, and torch.distributed (NCCL over NVLink) for the two
exchanges of the step — the per-layer SyncBN statistics (forward: sum x | sum x^2 | n, backward: sum g | sum g*xhat) and ONE
all-reduce of the flat fp32 gradient bucket.  No torch operator touches an activation.

Scope: the full head (ASPP_block + Decoder_block + tail) — the Xception / ResNet50 / MobileNetV2 non-lite models.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

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

    def __init__(self, Cin: int, Cskip: int, NC: int):
        self.Cin, self.Cs, self.NC = Cin, Cskip, NC
        self.NCp = _rup(NC, 8)
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
        return [('image_pooling', Cin, 256), ('aspp0', Cin, 256), ('aspp1_pointwise', Cin, 256), ('aspp2_pointwise', Cin, 256),
                ('aspp3_pointwise', Cin, 256), ('concat_projection', 1280, 256), ('feature_projection0', Cs, 48),
                ('decoder_conv0_pointwise', 304, 256), ('decoder_conv1_pointwise', 256, 256), ('conv_upsample', 256, NCp)]

    def _dw_specs(self):
        return [('aspp1_depthwise', self.Cin), ('aspp2_depthwise', self.Cin), ('aspp3_depthwise', self.Cin),
                ('decoder_conv0_depthwise', 304), ('decoder_conv1_depthwise', 256)]

    def _bn_specs(self):
        Cin = self.Cin
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


class HeadTrainer(TrainLayout):
    """One replica of the data-parallel training step.  All tensors live on `device`; feed bf16 NHWC features and uint8
    labels already resident in HBM (the backbone is outside this path; d_feat / d_skip are returned for it)."""

    def __init__(self, B: int, H: int, W: int, OS: int, Cin: int, Cskip: int, NC: int, weights: Dict[Tuple[str, str], np.ndarray],
                 device: int = 0, lr: float = 1e-2, momentum: float = 0.9, l2: float = L2_COEF, bn_momentum: float = BN_MOMENTUM,
                 eps: float = 1e-5, dropout: float = 0.5, seed: int = 0, ignore_index: int = 255, global_batch: Optional[int] = None,
                 process_group=None, graph: bool = True, wgrad_tn: bool = True, loss: str = 'crossentropy', class_weights=None,
                 focal_gamma: float = 2.0, focal_alpha: float = 0.25, exchange: str = 'p2p'):
        import torch
        import torch.distributed as dist
        self.torch = torch
        if not torch.cuda.is_available():
            raise ffi.Dlv3pError(-2, 'HeadTrainer needs a CUDA device: there is no CPU path')
        tf_.lib()
        self.B, self.H, self.W, self.OS, self.Cin, self.Cs, self.NC = B, H, W, OS, Cin, Cskip, NC
        self.rates = atrous_rates(OS)
        self.h, self.w = -(-H // OS), -(-W // OS)
        self.hs, self.ws = -(-H // 4), -(-W // 4)
        self.M1, self.M2 = B * self.h * self.w, B * self.hs * self.ws
        if Cin % 8 or Cskip % 8 or self.M1 % 8 or self.M2 % 8:
            raise ffi.Dlv3pError(-1, 'HeadTrainer: Cin, Cskip and the pixel counts per replica must be multiples of 8')
        TrainLayout.__init__(self, Cin, Cskip, NC)
        self.Bp = _rup(B, 8)
        self.dev = device
        self.tdev = torch.device('cuda', device)
        self.lr, self.momentum, self.l2, self.bn_momentum, self.eps = lr, momentum, l2, bn_momentum, eps
        self.drop_rate, self.seed, self.ignore = dropout, seed, ignore_index
        self.pg = process_group
        self.dist = dist if (dist.is_available() and dist.is_initialized() and dist.get_world_size(process_group) > 1) else None
        self.world = dist.get_world_size(process_group) if self.dist else 1
        self.rank = dist.get_rank(process_group) if self.dist else 0
        self.global_batch = global_batch if global_batch is not None else B * self.world
        self.step_count = 0
        self.launches = 0
        self.debug_taps = None          # set to {} to snapshot intermediate gradients (diagnostics / tests)
        self.use_graph, self.wgrad_tn = graph, wgrad_tn
        # train.py:114-138: --loss crossentropy (optionally class weighted, --weighted_type balanced) | focal (ignores the weights)
        if loss not in ('crossentropy', 'focal'):
            raise ValueError('invalid loss type {}'.format(loss))
        self.loss_kind = 2 if loss == 'focal' else (1 if class_weights is not None else 0)
        self.focal_gamma, self.focal_alpha = focal_gamma, focal_alpha
        self._class_weights = None if class_weights is None else np.asarray(class_weights, np.float32).reshape(-1)
        if self._class_weights is not None and self._class_weights.size != NC:
            raise ValueError('class_weights must have one entry per class')
        self.graph_error_mode = 'global' if self.world == 1 else 'thread_local'   # NCCL's watchdog thread polls events during capture
        self._graph = None
        self._static_in = None
        # SyncBN exchanges: 'p2p' = one-shot all-reduces over NVLink peer memory (dlv3p_p2p_*: ~5 us per collective instead of ~35 us
        # of ncclAllReduce latency for <= 54 KB vectors, 14 per step); 'nccl' = torch.distributed all_reduce.  The 12.7 MB gradient
        # bucket stays on NCCL either way (bandwidth bound: a ring / tree beats W peer reads of the whole bucket).
        if exchange not in ('p2p', 'nccl'):
            raise ValueError("exchange must be 'p2p' or 'nccl'")
        self.xchg = None
        self._alloc()
        if self.dist and exchange == 'p2p':
            self.xchg = tf_.P2pExchange(device, self.world, self.rank, self.nstats + (self.nparams - self.endB))
            handles = [None] * self.world
            dist.all_gather_object(handles, self.xchg.handle, group=process_group)
            self.xchg.connect(handles)
            dist.barrier(group=process_group)
        self.set_weights(weights)

    # lr / momentum / l2 travel by value into dlv3p_train_sgd, so a captured CUDA graph has them baked in: changing one
    # (ReduceLROnPlateau, the cosine / poly decay schedules of the reference's train.py:49-66, :192-215) drops the captured
    # graph; the next train_step re-captures with the new value.
    def _hyper(name):
        def get(self):
            return self.__dict__['_' + name]

        def set_(self, v):
            if self.__dict__.get('_' + name) != v:
                self.__dict__['_' + name] = v
                if self.__dict__.get('_graph') is not None:
                    self.__dict__['_graph'] = None
        return property(get, set_)

    lr, momentum, l2 = _hyper('lr'), _hyper('momentum'), _hyper('l2')
    del _hyper

    def _alloc(self):
        t = self.torch
        dv = self.tdev
        bf, f32 = t.bfloat16, t.float32
        z = lambda *s, dtype=bf: t.zeros(*s, dtype=dtype, device=dv)
        self.params, self.grads, self.velocity = z(self.nparams, dtype=f32), z(self.nparams, dtype=f32), z(self.nparams, dtype=f32)
        self.w_kn = z(self.endA)                 # bf16 copy of region A (Keras [K,N] layout: the dgrad operand)
        self.w_nk = z(self.endA)                 # transposed copies [N,K]: the forward operand
        self.stats = z(self.nstats, dtype=f32)
        # Dropout seed of the CURRENT step in device memory (int32 bit pattern of dropout_seed(seed, step, rank)); advanced on the device
        # at the end of every step so that a captured CUDA graph draws a fresh mask per replay
        s0 = dropout_seed(self.seed, 0, self.rank)
        self.seed_t = t.tensor([s0 - (1 << 32) if s0 >= (1 << 31) else s0], dtype=t.int32, device=dv)
        self.moving_mean, self.moving_var = z(self.nbn, dtype=f32), t.ones(self.nbn, dtype=f32, device=dv)
        B, Bp, M1, M2, Cin, Cs, NCp = self.B, self.Bp, self.M1, self.M2, self.Cin, self.Cs, self.NCp
        T = {}
        T['pool'], T['r4'], T['b4'] = z(Bp, Cin), z(Bp, 256), z(Bp, 256)
        T['r0'], T['concat'], T['rp'], T['yproj'], T['aspp_out'] = z(M1, 256), z(M1, 1280), z(M1, 256), z(M1, 256), z(M1, 256)
        for i in (1, 2, 3):
            T['d%d' % i], T['a%d' % i], T['p%d' % i] = z(M1, Cin), z(M1, Cin), z(M1, 256)
        T['dcat'], T['rs'] = z(M2, 304), z(M2, 48)
        T['c0d'], T['c0a'], T['c0p'], T['y0'] = z(M2, 304), z(M2, 304), z(M2, 256), z(M2, 256)
        T['c1d'], T['c1a'], T['c1p'], T['y1'] = z(M2, 256), z(M2, 256), z(M2, 256), z(M2, 256)
        T['logits'] = z(M2, NCp, dtype=f32)
        T['dfull'] = z(B, self.NC, self.H, self.W, dtype=f32)
        T['loss'] = z(2, dtype=f32)
        T['class_w'] = t.ones(max(self.NC, 1), dtype=f32, device=dv) if self._class_weights is None else t.from_numpy(self._class_weights).to(dv)
        T['adj_tmp'] = z(B, self.NC, self.hs, self.W, dtype=f32)      # vertical pass of the separable pred_resize adjoint
        T['bias_stats'] = z(2 * NCp + 4, dtype=f32)
        # backward
        T['dlow'] = z(M2, NCp)
        T['g256a'], T['g256b'], T['g256c'] = z(M2, 256), z(M2, 256), z(M2, 256)   # gradient buffers at decoder resolution
        T['g304a'], T['g304b'] = z(M2, 304), z(M2, 304)
        T['drs'], T['dskip'] = z(M2, 48), z(M2, Cs)
        T['da_out'], T['drp'], T['dconcat'] = z(M1, 256), z(M1, 256), z(M1, 1280)
        T['gB'] = z(M1, Cin)
        for i in (1, 2, 3):
            T['gp%d' % i], T['ga%d' % i] = z(M1, 256), z(M1, Cin)
        T['dfeat'], T['dfeat_tmp'] = z(M1, Cin), z(M1, Cin)
        T['g1_256'] = z(M1, 256)
        T['db4'], T['dr4'], T['dpool'] = z(Bp, 256), z(Bp, 256), z(Bp, Cin)
        if not self.wgrad_tn:
            T['xT'] = z(max(Cin * M1, 304 * M2, 1280 * M1, Cs * M2, Cin * Bp))
            T['dyT'] = z(max(256 * M2, 256 * M1, 256 * Bp, NCp * M2))
        self.sms = ffi.device_info(self.dev)['sm_count']
        T['partial'] = z(16 * 1024 * 1024 + self.sms * 2 * 128 * 256, dtype=f32)
        maxC = max(Cin, 304, 256)
        T['scratch'] = z(tf_.scratch_bytes(maxC) // 4 + 64, dtype=f32)
        T['bn_scratch'] = z(ffi.bn_scratch_bytes(maxC) // 4 + 64, dtype=f32)
        T['loss_scratch'] = z(tf_.loss_scratch_bytes() // 4 + 16, dtype=f32)
        self.T = T
        # gather indices for the vectorised moving-statistics update
        sx, sq, nn = [], [], []
        for name, Cc in self._bn_specs():
            o, _ = self.stat_off[name]
            sx += list(range(o, o + Cc)); sq += list(range(o + Cc, o + 2 * Cc)); nn += [o + 2 * Cc] * Cc
        self._ix = tuple(t.tensor(v, dtype=t.long, device=dv) for v in (sx, sq, nn))

    def view(self, flat, key):
        o, shape = self.off[key]
        n = int(np.prod(shape))
        return flat[o:o + n].view(*shape)

    def set_weights(self, W: Dict[Tuple[str, str], np.ndarray]) -> None:
        """Keras-layout arrays keyed (layer, variable) or "layer/variable" (the npz of tools/h5_to_npz.py) — the inventory of
        dlv3p_weight_info / SURVEY §8(b); `logits_semantic` is accepted for `conv_upsample`."""
        t = self.torch
        W = {(tuple(k.split('/', 1)) if isinstance(k, str) else tuple(k)): v for k, v in W.items()}
        W = {(('conv_upsample', k[1]) if k[0] == 'logits_semantic' else k): v for k, v in W.items()}
        host = np.zeros(self.nparams, np.float32)

        def put(key, arr):
            o, shape = self.off[key]
            host[o:o + arr.size] = np.asarray(arr, np.float32).reshape(-1)

        for name, K, N in self._conv_specs():
            k = np.asarray(W[(name, 'kernel')], np.float32).reshape(K, -1)
            if name == 'conv_upsample':
                k = np.concatenate([k, np.zeros((K, self.NCp - self.NC), np.float32)], axis=1)
            put((name, 'kernel'), k)
        put(('conv_upsample', 'bias'), np.concatenate([np.asarray(W[('conv_upsample', 'bias')], np.float32), np.zeros(self.NCp - self.NC, np.float32)]))
        for name, Cc in self._dw_specs():
            put((name, 'depthwise_kernel'), np.asarray(W[(name, 'depthwise_kernel')], np.float32).reshape(9, Cc))
        mm, mv = [], []
        for name, Cc in self._bn_specs():
            put((name, 'beta'), W[(name, 'beta')]); put((name, 'gamma'), W[(name, 'gamma')])
            mm.append(np.asarray(W[(name, 'moving_mean')], np.float32)); mv.append(np.asarray(W[(name, 'moving_variance')], np.float32))
        self.params.copy_(t.from_numpy(host))
        self.moving_mean.copy_(t.from_numpy(np.concatenate(mm))); self.moving_var.copy_(t.from_numpy(np.concatenate(mv)))
        self.velocity.zero_()
        self._refresh_bf16()
        t.cuda.synchronize(self.tdev)

    def get_weights(self) -> Dict[Tuple[str, str], np.ndarray]:
        """Current fp32 master weights and moving statistics in the Keras layout (what model.save would write)."""
        host = self.params.detach().cpu().numpy()
        out = {}
        for key, (o, shape) in self.off.items():
            a = host[o:o + int(np.prod(shape))].reshape(shape).copy()
            name, var = key
            if var == 'kernel':
                a = a[:, :self.NC] if name == 'conv_upsample' else a
                a = a.reshape(1, 1, *a.shape)
            elif var == 'bias':
                a = a[:self.NC]
            elif var == 'depthwise_kernel':
                a = a.reshape(3, 3, shape[1], 1)
            out[key] = a
        mm, mv = self.moving_mean.cpu().numpy(), self.moving_var.cpu().numpy()
        o = 0
        for name, Cc in self._bn_specs():
            out[(name, 'moving_mean')] = mm[o:o + Cc].copy(); out[(name, 'moving_variance')] = mv[o:o + Cc].copy()
            o += Cc
        return out

    def get_grads(self) -> Dict[Tuple[str, str], np.ndarray]:
        host = self.grads.detach().cpu().numpy()
        out = {}
        for key, (o, shape) in self.off.items():
            a = host[o:o + int(np.prod(shape))].reshape(shape).copy()
            if key == ('conv_upsample', 'kernel'):
                a = a[:, :self.NC]
            elif key == ('conv_upsample', 'bias'):
                a = a[:self.NC]
            out[key] = a
        return out

    # ------------------------------------------------------------------------------------------------ helpers
    def _s(self):
        return self.torch.cuda.current_stream(self.tdev).cuda_stream

    # kernels behind one entry point (two-stage reductions launch a partial and a final kernel)
    _KERNELS_PER_CALL = {'dlv3p_op_bn_stats': 2, 'dlv3p_train_bn_bwd_stats': 2, 'dlv3p_train_depthwise_wgrad': 2, 'dlv3p_train_softmax_loss': 2,
                         'dlv3p_train_softmax_ce': 2, 'dlv3p_train_resize_bwd_planar': 2}

    def _call(self, name, *args):
        tf_.call(name, self.dev, *args, self._s())
        self.launches += self._KERNELS_PER_CALL.get(name, 1)

    @staticmethod
    def _p(tensor, off_elems: int = 0) -> int:
        return tensor.data_ptr() + off_elems * tensor.element_size()

    def _wp(self, flat, key) -> int:
        return self._p(flat, self.off[key][0])

    def _allreduce(self, tensor):
        if self.dist:
            self.dist.all_reduce(tensor, op=self.dist.ReduceOp.SUM, group=self.pg)

    def _refresh_bf16(self):
        """bf16 operand copies of the 1x1 kernels after an update: [K,N] (dgrad) and its transpose [N,K] (forward)."""
        self._call('dlv3p_train_cast_bf16', self._p(self.params), self._p(self.w_kn), self.endA)
        for name, K, N in self._conv_specs():
            o = self.off[(name, 'kernel')][0]
            self._call('dlv3p_train_transpose', self._p(self.w_kn, o), K, N, N, self._p(self.w_nk, o), K)

    def _gemm(self, a, lda, b, ldb, M, N, K, d, ldd, out_fp32=0, splits=1):
        self._call('dlv3p_train_gemm_nt', a, lda, b, ldb, M, N, K, d, ldd, out_fp32, splits, self._p(self.T['partial']))
        self.launches += 1 if splits > 1 else 0          # the split-K reduction

    def _conv_fwd(self, name, x_ptr, ldx, M, out_ptr, ldo, out_fp32=0):
        _, (K, N) = self.off[(name, 'kernel')]
        self._gemm(x_ptr, ldx, self._wp(self.w_nk, (name, 'kernel')), K, M, N, K, out_ptr, ldo, out_fp32)

    def _conv_dgrad(self, name, dy_ptr, ld_dy, M, dx_ptr, ldx):
        _, (K, N) = self.off[(name, 'kernel')]
        self._gemm(dy_ptr, ld_dy, self._wp(self.w_kn, (name, 'kernel')), N, M, K, N, dx_ptr, ldx)

    def _conv_wgrad(self, name, x_ptr, ldx, dy_ptr, ld_dy, M):
        """dW[K,N] = X[M,K]^T dY[M,N]: the MN-major tcgen05 GEMM reads both straight from the [pixels, channels] tensors;
        the long contraction over pixels is split so that ~one wave of CTAs is busy (fp32 partials, fixed-order reduce)."""
        _, (K, N) = self.off[(name, 'kernel')]
        T = self.T
        tiles = -(-K // 128) * -(-N // (256 if N > 64 else 64))
        kblocks = -(-M // 64)
        splits = max(1, min(kblocks, self.sms // tiles, (T['partial'].numel()) // max(1, K * N)))
        if self.wgrad_tn:
            self._call('dlv3p_train_gemm_tn', x_ptr, ldx, dy_ptr, ld_dy, K, N, M, self._wp(self.grads, (name, 'kernel')), N, 1, splits, self._p(T['partial']))
            self.launches += 1 if splits > 1 else 0      # the split-K reduction
        else:   # A/B path: explicit transposes + the K-major kernel
            self._call('dlv3p_train_transpose', x_ptr, M, K, ldx, self._p(T['xT']), M)
            self._call('dlv3p_train_transpose', dy_ptr, M, N, ld_dy, self._p(T['dyT']), M)
            self._gemm(self._p(T['xT']), M, self._p(T['dyT']), M, K, N, M, self._wp(self.grads, (name, 'kernel')), N, 1, splits)

    def _bn_stats(self, name, x, M):
        o, Cc = self.stat_off[name]
        dst = self.xchg.payload(o) if self.xchg else self._p(self.stats, o)     # peer exchange: partial sums go to this replica's payload area
        self._call('dlv3p_op_bn_stats', self._p(x), M, Cc, dst, self._p(self.T['bn_scratch']))

    def _sync_stats(self, group):
        """ONE all-reduce (SUM) of the contiguous [sum x | sum x^2 | n] vectors of a group of independent BN layers."""
        b, e = self.stats_span(group)
        if self.xchg:
            self.xchg.allreduce(self.FWD_GROUPS.index(group), b, _rup(e - b, 4), self._p(self.stats, b), self._s())
            self.launches += 1
        else:
            self._allreduce(self.stats[b:e])

    def _bn_apply(self, name, x, M, y_ptr, ldy, relu=1):
        o, Cc = self.stat_off[name]
        self._call('dlv3p_train_bn_apply', self._p(x), M, Cc, self._p(self.stats, o), self._wp(self.params, (name, 'gamma')),
                   self._wp(self.params, (name, 'beta')), self.eps, relu, y_ptr, ldy)

    def _bn_fwd(self, name, x, M, y_ptr, ldy, relu=1):
        self._bn_stats(name, x, M)
        self._sync_stats([name])
        self._bn_apply(name, x, M, y_ptr, ldy, relu)

    def _bn_bwd_stats(self, name, dy_ptr, ld_dy, y_ptr, ld_y, x, M, relu=1):
        o, Cc = self.stat_off[name]
        go = self.off[(name, 'beta')][0]                      # grads[go : go+2C] = d(beta) | d(gamma) = sum g | sum g*xhat
        dst = self.xchg.payload(self.nstats + go - self.endB) if self.xchg else self._p(self.grads, go)
        self._call('dlv3p_train_bn_bwd_stats', dy_ptr, ld_dy, y_ptr, ld_y, self._p(x), M, Cc, self._p(self.stats, o), self.eps, relu,
                   dst, self._p(self.T['scratch']))

    def _sync_bn_grads(self, group):
        b, e = self.bn_grad_span(group)
        if self.xchg:
            self.xchg.allreduce(len(self.FWD_GROUPS) + self.BWD_GROUPS.index(group), self.nstats + b - self.endB, _rup(e - b, 4), self._p(self.grads, b), self._s())
            self.launches += 1
        else:
            self._allreduce(self.grads[b:e])

    def _bn_bwd_apply(self, name, dy_ptr, ld_dy, y_ptr, ld_y, x, M, dx, relu=1):
        o, Cc = self.stat_off[name]
        go = self.off[(name, 'beta')][0]
        self._call('dlv3p_train_bn_bwd_apply', dy_ptr, ld_dy, y_ptr, ld_y, self._p(x), M, Cc, self._p(self.stats, o), self._p(self.grads, go),
                   self._wp(self.params, (name, 'gamma')), self.eps, relu, self._p(dx))

    def _bn_bwd(self, name, dy_ptr, ld_dy, y_ptr, ld_y, x, M, dx, relu=1):
        self._bn_bwd_stats(name, dy_ptr, ld_dy, y_ptr, ld_y, x, M, relu)
        self._sync_bn_grads([name])
        self._bn_bwd_apply(name, dy_ptr, ld_dy, y_ptr, ld_y, x, M, dx, relu)

    def _sep_fwd(self, prefix, x, Bn, Hh, Ww, Cc, rate, d, a, p, y_ptr, ldy):
        """SepConv_BN (depth_activation=True, layers.py:98-109) in training mode; keeps d (raw depthwise), a (BN+ReLU), p (raw pointwise)."""
        M = Bn * Hh * Ww
        self._call('dlv3p_train_depthwise', self._p(x), Bn, Hh, Ww, Cc, rate, self._wp(self.params, (prefix + '_depthwise', 'depthwise_kernel')), 0, self._p(d))
        self._bn_fwd(prefix + '_depthwise_BN', d, M, self._p(a), Cc)
        self._conv_fwd(prefix + '_pointwise', self._p(a), Cc, M, self._p(p), 256)
        self._bn_fwd(prefix + '_pointwise_BN', p, M, y_ptr, ldy)

    def _tap(self, name, buf):
        if self.debug_taps is not None:
            self.debug_taps[name] = buf.detach().float().cpu().numpy().copy()

    def _sep_bwd(self, prefix, x, Bn, Hh, Ww, Cc, rate, d, a, p, y_ptr, ld_y, dy_ptr, ld_dy, g_pw, g_a, g_d, dx):
        """Backward of _sep_fwd.  g_pw [M,256], g_a / g_d [M,Cc] are scratch gradients; dx [M,Cc] receives d(loss)/d(x)."""
        M = Bn * Hh * Ww
        self._bn_bwd(prefix + '_pointwise_BN', dy_ptr, ld_dy, y_ptr, ld_y, p, M, g_pw)
        self._tap(prefix + '/p', g_pw)
        self._conv_wgrad(prefix + '_pointwise', self._p(a), Cc, self._p(g_pw), 256, M)
        self._conv_dgrad(prefix + '_pointwise', self._p(g_pw), 256, M, self._p(g_a), Cc)
        self._tap(prefix + '/a', g_a)
        self._bn_bwd(prefix + '_depthwise_BN', self._p(g_a), Cc, self._p(a), Cc, d, M, g_d)
        self._tap(prefix + '/d', g_d)
        self._call('dlv3p_train_depthwise_wgrad', self._p(x), self._p(g_d), Bn, Hh, Ww, Cc, rate,
                   self._wp(self.grads, (prefix + '_depthwise', 'depthwise_kernel')), self._p(self.T['scratch']))
        self._call('dlv3p_train_depthwise', self._p(g_d), Bn, Hh, Ww, Cc, rate, self._wp(self.params, (prefix + '_depthwise', 'depthwise_kernel')), 1, self._p(dx))

    # ------------------------------------------------------------------------------------------------ the step
    def forward_backward(self, feat, skip, labels):
        """feat bf16 [B,h,w,Cin], skip bf16 [B,hs,ws,Cs], labels uint8 [B,H,W] (CUDA tensors).  Leaves the loss in
        self.T['loss'][0] (this replica's share of the global mean), all weight gradients in self.grads (BN gradients global,
        the rest per replica until all_reduce_gradients) and d(loss)/d(feat), d(loss)/d(skip) in self.T['dfeat'] / ['dskip']."""
        T = self.T
        P = self._p
        B, Bp, M1, M2, Cin, Cs, NC, NCp = self.B, self.Bp, self.M1, self.M2, self.Cin, self.Cs, self.NC, self.NCp
        h, w, hs, ws = self.h, self.w, self.hs, self.ws
        npix1 = h * w
        assert feat.is_contiguous() and skip.is_contiguous() and labels.is_contiguous()
        # ---------------- ASPP_block (layers.py:114-163) + the decoder's skip projection (:209-213), phase by phase:
        # phase 1: everything that depends only on the inputs — raw outputs, their statistics, ONE exchange, then the six BN+ReLU
        self._call('dlv3p_train_rows_reduce', P(feat), Cin, B, npix1, Cin, 1.0 / npix1, P(T['pool']), 0)
        self._conv_fwd('image_pooling', P(T['pool']), Cin, Bp, P(T['r4']), 256)
        self._bn_stats('image_pooling_BN', T['r4'], B)
        self._conv_fwd('aspp0', P(feat), Cin, M1, P(T['r0']), 256)
        self._bn_stats('aspp0_BN', T['r0'], M1)
        for i in (1, 2, 3):
            self._call('dlv3p_train_depthwise', P(feat), B, h, w, Cin, self.rates[i - 1], self._wp(self.params, ('aspp%d_depthwise' % i, 'depthwise_kernel')), 0,
                       P(T['d%d' % i]))
            self._bn_stats('aspp%d_depthwise_BN' % i, T['d%d' % i], M1)
        self._conv_fwd('feature_projection0', P(skip), Cs, M2, P(T['rs']), 48)
        self._bn_stats('feature_projection0_BN', T['rs'], M2)
        self._sync_stats(self.FWD_GROUPS[0])
        self._bn_apply('image_pooling_BN', T['r4'], B, P(T['b4']), 256)
        self._call('dlv3p_train_bcast_rows', P(T['b4']), B, npix1, 256, 1.0, P(T['concat']), 1280, 0)
        self._bn_apply('aspp0_BN', T['r0'], M1, P(T['concat'], 256), 1280)
        for i in (1, 2, 3):
            self._bn_apply('aspp%d_depthwise_BN' % i, T['d%d' % i], M1, P(T['a%d' % i]), Cin)
        self._bn_apply('feature_projection0_BN', T['rs'], M2, P(T['dcat'], 256), 304)
        # phase 2: the three atrous pointwise convs
        for i in (1, 2, 3):
            self._conv_fwd('aspp%d_pointwise' % i, P(T['a%d' % i]), Cin, M1, P(T['p%d' % i]), 256)
            self._bn_stats('aspp%d_pointwise_BN' % i, T['p%d' % i], M1)
        self._sync_stats(self.FWD_GROUPS[1])
        for i in (1, 2, 3):
            self._bn_apply('aspp%d_pointwise_BN' % i, T['p%d' % i], M1, P(T['concat'], 256 * (i + 1)), 1280)
        self._conv_fwd('concat_projection', P(T['concat']), 1280, M1, P(T['rp']), 256)
        self._bn_fwd('concat_projection_BN', T['rp'], M1, P(T['yproj']), 256)
        if self.drop_rate > 0:
            self._call('dlv3p_train_dropout', P(T['yproj']), P(T['aspp_out']), M1 * 256, 0, P(self.seed_t), self.drop_rate)
            aspp_out = T['aspp_out']
        else:
            aspp_out = T['yproj']
        # ---------------- Decoder_block (layers.py:199-219)
        self._call('dlv3p_train_resize', P(aspp_out), B, h, w, 256, hs, ws, P(T['dcat']), 304)
        self._sep_fwd('decoder_conv0', T['dcat'], B, hs, ws, 304, 1, T['c0d'], T['c0a'], T['c0p'], P(T['y0']), 256)
        self._sep_fwd('decoder_conv1', T['y0'], B, hs, ws, 256, 1, T['c1d'], T['c1a'], T['c1p'], P(T['y1']), 256)
        # ---------------- tail + loss (model.py:75-86, loss.py:121-156)
        self._conv_fwd('conv_upsample', P(T['y1']), 256, M2, P(T['logits']), NCp, 1)
        inv_norm = 1.0 / (float(self.global_batch) * self.H * self.W)
        self._call('dlv3p_train_softmax_loss', P(T['logits']), NCp, self._wp(self.params, ('conv_upsample', 'bias')), P(labels), B, NC, hs, ws, self.H, self.W,
                   self.ignore, inv_norm, self.loss_kind, P(T['class_w']), self.focal_gamma, self.focal_alpha, P(T['dfull']), P(T['loss']), P(T['loss_scratch']))
        # ================ backward
        self._call('dlv3p_train_resize_bwd_planar', P(T['dfull']), B, NC, hs, ws, self.H, self.W, P(T['dlow']), NCp, P(T['adj_tmp']))
        # d(bias) = column sums of d(logits): the banded two-stage statistics kernel (sum x | sum x^2 | n), first NCp entries
        self._call('dlv3p_op_bn_stats', P(T['dlow']), M2, NCp, P(T['bias_stats']), P(T['bn_scratch']))
        self.view(self.grads, ('conv_upsample', 'bias')).copy_(T['bias_stats'][:NCp])
        self._conv_wgrad('conv_upsample', P(T['y1']), 256, P(T['dlow']), NCp, M2)
        self._conv_dgrad('conv_upsample', P(T['dlow']), NCp, M2, P(T['g256a']), 256)
        self._tap('logits', T['dlow'])
        self._tap('decoder_conv1/y', T['g256a'])
        # decoder_conv1: dy = g256a, x = y0; scratch g256b / g256c; dx overwrites g256a (dy is dead after the first BN backward)
        self._sep_bwd('decoder_conv1', T['y0'], B, hs, ws, 256, 1, T['c1d'], T['c1a'], T['c1p'], P(T['y1']), 256, P(T['g256a']), 256,
                      T['g256b'], T['g256c'], T['g256b'], T['g256a'])
        # decoder_conv0: dy = g256a, x = dcat (304 channels); dx overwrites g304a
        self._sep_bwd('decoder_conv0', T['dcat'], B, hs, ws, 304, 1, T['c0d'], T['c0a'], T['c0p'], P(T['y0']), 256, P(T['g256a']), 256,
                      T['g256b'], T['g304a'], T['g304b'], T['g304a'])
        # decoder_resize adjoint and the Dropout mask give d(loss)/d(concat_projection output); together with the skip projection
        # (dy = g304a[:, 256:304]) that is one group of two independent BN backwards: statistics, ONE exchange, then both applies
        self._call('dlv3p_train_resize_bwd', P(T['g304a']), 304, B, h, w, 256, hs, ws, P(T['da_out']))
        if self.drop_rate > 0:
            self._call('dlv3p_train_dropout', P(T['da_out']), P(T['da_out']), M1 * 256, 0, P(self.seed_t), self.drop_rate)
        self._bn_bwd_stats('feature_projection0_BN', P(T['g304a'], 256), 304, P(T['dcat'], 256), 304, T['rs'], M2)
        self._bn_bwd_stats('concat_projection_BN', P(T['da_out']), 256, P(T['yproj']), 256, T['rp'], M1)
        self._sync_bn_grads(self.BWD_GROUPS[4])
        self._bn_bwd_apply('feature_projection0_BN', P(T['g304a'], 256), 304, P(T['dcat'], 256), 304, T['rs'], M2, T['drs'])
        self._bn_bwd_apply('concat_projection_BN', P(T['da_out']), 256, P(T['yproj']), 256, T['rp'], M1, T['drp'])
        self._conv_wgrad('feature_projection0', P(skip), Cs, P(T['drs']), 48, M2)
        self._conv_dgrad('feature_projection0', P(T['drs']), 48, M2, P(T['dskip']), Cs)
        self._conv_wgrad('concat_projection', P(T['concat']), 1280, P(T['drp']), 256, M1)
        self._conv_dgrad('concat_projection', P(T['drp']), 256, M1, P(T['dconcat']), 1280)
        # the five BN layers that feed the concat: aspp0, the three atrous pointwise convs, the image pooling branch
        self._call('dlv3p_train_rows_reduce', P(T['dconcat']), 1280, B, npix1, 256, 1.0, P(T['db4']), 0)
        names = ['aspp0_BN', 'aspp1_pointwise_BN', 'aspp2_pointwise_BN', 'aspp3_pointwise_BN']
        raws = [T['r0'], T['p1'], T['p2'], T['p3']]
        gout = [T['g1_256'], T['gp1'], T['gp2'], T['gp3']]
        for k, (name, raw) in enumerate(zip(names, raws)):
            self._bn_bwd_stats(name, P(T['dconcat'], 256 * (k + 1)), 1280, P(T['concat'], 256 * (k + 1)), 1280, raw, M1)
        self._bn_bwd_stats('image_pooling_BN', P(T['db4']), 256, P(T['b4']), 256, T['r4'], B)
        self._sync_bn_grads(self.BWD_GROUPS[5])
        for k, (name, raw, g) in enumerate(zip(names, raws, gout)):
            self._bn_bwd_apply(name, P(T['dconcat'], 256 * (k + 1)), 1280, P(T['concat'], 256 * (k + 1)), 1280, raw, M1, g)
        self._bn_bwd_apply('image_pooling_BN', P(T['db4']), 256, P(T['b4']), 256, T['r4'], B, T['dr4'])
        # aspp0 and the image pooling branch end here
        self._conv_wgrad('aspp0', P(feat), Cin, P(T['g1_256']), 256, M1)
        self._conv_dgrad('aspp0', P(T['g1_256']), 256, M1, P(T['dfeat']), Cin)
        self._conv_wgrad('image_pooling', P(T['pool']), Cin, P(T['dr4']), 256, Bp)
        self._conv_dgrad('image_pooling', P(T['dr4']), 256, Bp, P(T['dpool']), Cin)
        self._call('dlv3p_train_bcast_rows', P(T['dpool']), B, npix1, Cin, 1.0 / npix1, P(T['dfeat']), Cin, 1)
        # atrous branches: pointwise gradients, then the three depthwise BN backwards as one group
        for i in (1, 2, 3):
            self._tap('aspp%d/p' % i, T['gp%d' % i])
            self._conv_wgrad('aspp%d_pointwise' % i, P(T['a%d' % i]), Cin, P(T['gp%d' % i]), 256, M1)
            self._conv_dgrad('aspp%d_pointwise' % i, P(T['gp%d' % i]), 256, M1, P(T['ga%d' % i]), Cin)
            self._tap('aspp%d/a' % i, T['ga%d' % i])
            self._bn_bwd_stats('aspp%d_depthwise_BN' % i, P(T['ga%d' % i]), Cin, P(T['a%d' % i]), Cin, T['d%d' % i], M1)
        self._sync_bn_grads(self.BWD_GROUPS[6])
        for i in (1, 2, 3):
            name = 'aspp%d_depthwise' % i
            self._bn_bwd_apply(name + '_BN', P(T['ga%d' % i]), Cin, P(T['a%d' % i]), Cin, T['d%d' % i], M1, T['gB'])
            self._tap('aspp%d/d' % i, T['gB'])
            self._call('dlv3p_train_depthwise_wgrad', P(feat), P(T['gB']), B, h, w, Cin, self.rates[i - 1], self._wp(self.grads, (name, 'depthwise_kernel')),
                       P(T['scratch']))
            self._call('dlv3p_train_depthwise', P(T['gB']), B, h, w, Cin, self.rates[i - 1], self._wp(self.params, (name, 'depthwise_kernel')), 1, P(T['dfeat_tmp']))
            self._call('dlv3p_train_add', P(T['dfeat']), P(T['dfeat_tmp']), P(T['dfeat']), M1 * Cin)

    def all_reduce_gradients(self):
        """ONE all-reduce (SUM) of the flat fp32 bucket holding every 1x1 kernel, the classifier bias and every depthwise kernel
        (regions A|B; the loss is normalised by the GLOBAL batch, so the sum is the gradient of the global mean loss —
        MirroredStrategy semantics, train.py:143-158).  BN gradients are already global."""
        b, e = self.bucket_span()
        self._allreduce(self.grads[b:e])

    def apply_gradients(self):
        P = self._p
        self._call('dlv3p_train_sgd', P(self.params), P(self.grads), P(self.velocity), self.endA, self.lr, self.momentum, self.l2, 1.0)
        n = self.nparams - self.endA
        self._call('dlv3p_train_sgd', P(self.params, self.endA), P(self.grads, self.endA), P(self.velocity, self.endA), n, self.lr, self.momentum, 0.0, 1.0)
        self._refresh_bf16()
        # moving statistics (Keras: moving <- moving * m + batch * (1 - m), biased variance); a handful of vector ops on <= 10k floats
        sx, sq, nn = self._ix
        n_ = self.stats[nn]
        mean = self.stats[sx] / n_
        var = (self.stats[sq] / n_ - mean * mean).clamp_(min=0)
        m = self.bn_momentum
        self.moving_mean.mul_(m).add_(mean, alpha=1 - m)
        self.moving_var.mul_(m).add_(var, alpha=1 - m)

    def _advance_seed(self):
        self.seed_t.add_(0x85EBCA6B - (1 << 32))        # dropout_seed is linear in the step: + 0x85EBCA6B (mod 2^32)
        if self.xchg:
            self.xchg.advance(self._s())                 # the flags of the next step's collectives carry the next epoch
            self.launches += 1

    def comm_backend(self) -> str:
        if not self.dist:
            return 'none'
        return 'SyncBN vectors: one-shot all-reduce over NVLink peer memory (dlv3p_p2p); gradient bucket: NCCL' if self.xchg else 'NCCL (torch.distributed)'

    def close(self):
        self._graph = None
        if self.xchg:
            self.torch.cuda.synchronize(self.tdev)
            self.xchg.close()
            self.xchg = None

    def _step_body(self, feat, skip, labels):
        self.forward_backward(feat, skip, labels)
        self.all_reduce_gradients()
        self.apply_gradients()
        self._advance_seed()

    def train_step(self, feat, skip, labels) -> None:
        """One optimizer step (fit's train_step): forward, loss, backward, gradient all-reduce, SGD update.  Asynchronous; read
        the loss with .loss() (synchronises).  The first call runs the ~150 kernels one by one; with graph=True the second call
        captures the whole step (kernels, NCCL all-reduces, the seed increment) into ONE CUDA graph that later calls replay: the
        step is launch bound otherwise (average kernel 25 us, ~35 us of host work per launch through ctypes).  Inputs are copied
        into static buffers the graph reads."""
        t = self.torch
        if not self.use_graph:
            self._step_body(feat, skip, labels)
        elif self.step_count == 0:
            self._static_in = (t.empty_like(feat), t.empty_like(skip), t.empty_like(labels))
            for dst, src in zip(self._static_in, (feat, skip, labels)):
                dst.copy_(src)
            self._step_body(*self._static_in)
        else:
            for dst, src in zip(self._static_in, (feat, skip, labels)):
                if dst.data_ptr() != src.data_ptr():
                    dst.copy_(src)
            if self._graph is None:
                t.cuda.synchronize(self.tdev)
                g = t.cuda.CUDAGraph()
                n0 = self.launches
                with t.cuda.graph(g, capture_error_mode=self.graph_error_mode):
                    self._step_body(*self._static_in)
                self._graph, self.launches_per_step = g, self.launches - n0
                self.launches = n0
            self._graph.replay()
            self.launches += self.launches_per_step
        self.step_count += 1

    def loss(self) -> float:
        """Global mean loss of the last step (sums the replicas' shares; synchronises)."""
        v = self.T['loss'][:1].clone()
        self._allreduce(v)
        return float(v.item())
