#!/usr/bin/env python3
"""bench.py — headline benchmark (BASELINE.json: images/sec, DeepLabV3+ Xception OS16 512x512 forward at 1/2/4/8 B200; ASPP %roofline).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (libdlv3p.so via ctypes)
  python bench.py --impl reference --gpus N --steps K ...   # the reference path's CPU restatement (oracle), host cores
  torchrun --nproc-per-node N ... bench.py --gpus N ...     # one rank per GPU; batch-sharded, no data-path collective

One "step" = one forward of the WHOLE model over one batch of synthetic uint8 images (configs[1]: Xception OS16 512x512, VOC 21
classes, batch 32 per GPU, bf16): normalize -> Xception backbone -> ASPP -> Decoder -> classifier -> pred_resize -> argmax, uint8
labels out.  Sub-records on the same line: `aspp` (the metric's ASPP %roofline), `other_configs` (configs 0/2/3), `train` (configs[4],
the head's data-parallel training step with its collectives).

PyTorch is used here only as plumbing: CUDA events on the launch stream, device buffers for the synthetic inputs and
torch.distributed (NCCL) for the barrier / max-over-ranks.  The measured path is the C ABI.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# ---- workload: BASELINE.json configs[1] ---------------------------------------------------------------------------
CFG = dict(B=32, H=512, W=512, OS=16, Cin=2048, Cskip=256, NC=21)
WORKLOAD = 'DeepLabV3+ Xception OS16 512x512 VOC, WHOLE MODEL (Xception backbone + ASPP + decoder + classifier + resize + argmax), batch 32/GPU, bf16 forward'
METRIC = 'images/sec DeepLabV3+ Xception OS16 512^2 fwd'

# algorithmic work per image at this config (SURVEY.md §8(d), 2 FLOP per MAC; DESIGN.md "Measurement")
GEMM_FLOP_PER_IMG = {
    'aspp_branches_gemm': 4 * 2 * 1024 * 2048 * 256,          # aspp0 + three atrous pointwise convs
    'concat_projection_gemm': 2 * 1024 * 1024 * 256,           # K = 1024: the b4 slice is a per-image bias
    'feature_projection0_gemm': 2 * 16384 * 256 * 48,
    'decoder_conv0_sepconv': 2 * 16384 * 304 * 256,
    'decoder_conv1_sepconv': 2 * 16384 * 256 * 256,
    'decoder_pointwise_gemm': 2 * 16384 * 280 * 256,           # unfused path only (mean of K=304 and K=256)
    'classifier_gemm': 2 * 16384 * 256 * 21,
}
# algorithmic HBM bytes per image for the memory-bound kernels (read inputs once + write outputs once)
HBM_BYTES_PER_IMG = {
    'aspp_dw_pool': 1024 * 2048 * 2 * (1 + 3),                  # read x once, write three depthwise maps
    'decoder_resize': 1024 * 256 * 2 + 16384 * 256 * 2,
    'feature_projection0_gemm': 16384 * 256 * 2 + 16384 * 48 * 2,
    'classifier_gemm': 16384 * 256 * 2 + 16384 * 21 * 4,
    'resize_argmax': 16384 * 21 * 4 + 512 * 512,
    'pool_proj': 2048 * 4 + 256 * 4,
}


# DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum, MB) from one `ncu --set full` capture of this
# very command at batch 32 (profiles/r01c_ncu_full_step_summary.csv): the `traffic` field of the roofline line.
NCU_DRAM_MB_PER_LAUNCH_B32 = {
    'decoder_conv0_sepconv': 335.77 + 228.12,
    'decoder_conv1_sepconv': 268.63 + 224.47,
    'aspp_dw_pool': 134.87 + 350.28,
    'aspp_branches_gemm': 541.21 + 14.98,
    'concat_projection_gemm': 67.69 + 2.27,
    'feature_projection0_gemm': 268.51 + 10.40,
    'classifier_gemm': 268.48 + 7.82,
    'decoder_resize': 16.86 + 215.53,
    'resize_argmax': 44.06 + 0.21,
    'pool_proj': 1.46,
}


# DRAM traffic per step (dram__bytes_read.sum + dram__bytes_write.sum summed over a kernel function's launches, bytes) from one
# `ncu --set full` capture of this command at batch 32; filled in from profiles/ once captured
NCU_TRAFFIC = {   # profiles/r03g_ncu_per_kernel_dram_and_time.txt (every template variant of a function summed over one step, MB)
    'bb_gemm_kernel': 5831.7 * 1e6,
    'bb_depthwise_kernel': 4503.6 * 1e6,
    'bb_sepconv_kernel': 2606.3 * 1e6,
    'conv3x3_c32_kernel': 340.2 * 1e6,
    'stem_tc_kernel': 98.7 * 1e6,
}
NCU_TRAFFIC_NOTE = 'bytes per step (dram__bytes_read.sum + dram__bytes_write.sum over the function\'s launches of one forward, ncu, cold caches: profiles/r03g_ncu_per_kernel_dram_and_time.txt)'

# algorithmic HBM bytes per image of the fused SepConv kernels (read the input once, write the 256-channel output once)
ALGO_BYTES_PER_IMG = {
    'decoder_conv0_sepconv': 16384 * 304 * 2 + 16384 * 256 * 2,
    'decoder_conv1_sepconv': 16384 * 256 * 2 + 16384 * 256 * 2,
    'aspp_branches_gemm': 4 * 1024 * 2048 * 2 + 4 * 1024 * 256 * 2,
    'concat_projection_gemm': 1024 * 1024 * 2 + 1024 * 256 * 2,
}


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {'hbm_gbs': d['hbm_gbs'], 'tflops_sustained': d.get('bf16_tflops_sustained', d['bf16_tflops']),
                'tflops_burst': d['bf16_tflops'], 'src': 'measured (MEASURED_PEAKS.json)'}
    return {'hbm_gbs': 6650.0, 'tflops_sustained': 1400.0, 'tflops_burst': 1590.0, 'src': 'fallback (B200_PROFILING.md)'}


def random_weights(specs, seed=1234):
    """Seeded random-init weights in Keras layout for ctx.weight_specs() (same recipe as SURVEY §8(c))."""
    rng = np.random.default_rng(seed)
    out = {}
    for layer, var, shape in specs:
        if var == 'kernel':
            a = rng.normal(0.0, np.sqrt(2.0 / shape[2]), size=shape)
        elif var == 'depthwise_kernel':
            a = rng.normal(0.0, 0.3, size=shape)
        elif var in ('gamma', 'moving_variance'):
            a = rng.uniform(0.5, 1.5, size=shape)
        else:
            a = rng.normal(0.0, 0.1, size=shape)
        out[(layer, var)] = a.astype(np.float32)
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = 'index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, gpu_index):
        self.idx, self.lines, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.idx), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'power_w_max': max(pw) if pw else None, 'samples': len(sm), 'reasons': sorted(reasons)}


def aspp_block(prof, B, peaks):
    """BASELINE.json's metric asks for 'ASPP %roofline': the ASPP block (layers.py:114-163) = depthwise + pooling pass, the
    four-branch GEMM and the projection GEMM (pool_proj runs on a side stream beside them)."""
    names = [k for k in ('aspp_dw_pool', 'aspp_branches_gemm', 'concat_projection_gemm') if k in prof]
    if not names:
        return None
    ms = sum(prof[k] for k in names)
    flop = sum(GEMM_FLOP_PER_IMG.get(k, 0) for k in names) * B
    tf = flop / (ms / 1000.0) / 1e12
    return {'kernels': names, 'ms': ms, 'gemm_tflops': tf, 'frac_of_tensor_peak': tf / peaks['tflops_burst'], 'peak': peaks['tflops_burst'],
            'branches_gemm_frac': (GEMM_FLOP_PER_IMG['aspp_branches_gemm'] * B / (prof['aspp_branches_gemm'] / 1000.0) / 1e12 / peaks['tflops_burst'])
            if 'aspp_branches_gemm' in prof else None}


OTHER_CONFIGS = {   # the BASELINE.json configurations that are parity-test cases, timed here so they reach a driver record
    'cfg1_mobilenetv2_os16_512_b1': dict(B=1, H=512, W=512, OS=16, Cin=320, Cskip=24, NC=21, lite=False, decoder=True),
    'cfg3_xception_os8_1024x2048_b8': dict(B=8, H=1024, W=2048, OS=8, Cin=2048, Cskip=256, NC=19, lite=False, decoder=True),
    'cfg4a_mobilenetv3large_lite_b64': dict(B=64, H=512, W=512, OS=16, Cin=160, Cskip=24, NC=21, lite=True, decoder=False),
    'cfg4b_mobilenetv3large_lite_decoder_b64': dict(B=64, H=512, W=512, OS=16, Cin=160, Cskip=24, NC=21, lite=True, decoder=True),
}
GEMM_FLOP_OTHER = {'cfg3_xception_os8_1024x2048_b8': 200.99e9, 'cfg1_mobilenetv2_os16_512_b1': 6.254e9,
                   'cfg4a_mobilenetv3large_lite_b64': 0.363e9, 'cfg4b_mobilenetv3large_lite_decoder_b64': 5.264e9}   # per image, SURVEY §8(d)


def time_other_configs(local_rank, steps=10):
    """Head-only device time of the other BASELINE configurations (N=1 only): ms/step, images/s, dominant kernel."""
    import torch
    import dlv3p_b200
    out = {}
    sp = torch.cuda.current_stream().cuda_stream
    for name, c in OTHER_CONFIGS.items():
        try:
            head = dlv3p_b200.DeepLabHead(c['B'], c['H'], c['W'], c['OS'], c['Cin'], c['Cskip'], c['NC'], lite=c['lite'], decoder=c['decoder'], device=local_rank)
            head.set_weights(random_weights(head.weight_specs()))
            h, w = c['H'] // c['OS'], c['W'] // c['OS']
            feat = torch.randn((c['B'], h, w, c['Cin']), device='cuda').clamp_(min=0).to(torch.bfloat16)
            skip = torch.randn((c['B'], c['H'] // 4, c['W'] // 4, c['Cskip']), device='cuda').to(torch.bfloat16) if c['decoder'] else None
            o = torch.empty((c['B'], c['H'], c['W']), device='cuda', dtype=torch.uint8)
            sk = skip.data_ptr() if skip is not None else 0
            for _ in range(3):
                head.ctx.forward(feat.data_ptr(), sk, o.data_ptr(), sp)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                head.ctx.forward(feat.data_ptr(), sk, o.data_ptr(), sp)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            prof = {}
            for _ in range(3):
                for k, t in head.ctx.profile(feat.data_ptr(), sk, o.data_ptr(), sp):
                    prof.setdefault(k, []).append(t)
            prof = {k: float(np.mean(v)) for k, v in prof.items()}
            dom = max(prof, key=prof.get)
            out[name] = {'ms_per_step': ms, 'images_per_s': c['B'] / ms * 1000.0, 'batch': c['B'],
                         'gemm_tflops': GEMM_FLOP_OTHER[name] * c['B'] / (ms / 1000.0) / 1e12,
                         'dominant_kernel': dom, 'dominant_kernel_ms': prof[dom], 'scope': 'head only'}
            head.close()
            del feat, skip, o
            torch.cuda.empty_cache()
        except Exception as e:      # a failing side configuration must not take the headline line down
            out[name] = {'error': repr(e)[:200]}
    return out


def time_train_step(rank, world, local_rank, steps=10, warmup=3, batch=8):
    """BASELINE configs[4]: the head's training step (forward in training mode, loss, backward, SyncBN exchanges, gradient
    all-reduce, SGD) at 8 images per GPU, replayed as one CUDA graph; max-over-ranks device time; replicas-identical check."""
    import torch
    import torch.distributed as dist
    from dlv3p_b200 import train
    specs = [(l, v, s) for l, v, s in __import__('dlv3p_b200').DeepLabHead(1, 512, 512, 16, CFG['Cin'], CFG['Cskip'], CFG['NC'], device=-1).weight_specs()]
    W = random_weights(specs)
    tr = train.HeadTrainer(batch, 512, 512, 16, CFG['Cin'], CFG['Cskip'], CFG['NC'], W, device=local_rank, seed=7, graph=True)
    g = torch.Generator(device='cuda').manual_seed(4321 + rank)
    feat = torch.randn(batch, 32, 32, CFG['Cin'], device='cuda', generator=g).clamp_(min=0).to(torch.bfloat16)
    skip = torch.randn(batch, 128, 128, CFG['Cskip'], device='cuda', generator=g).to(torch.bfloat16)
    labels = torch.randint(0, 21, (batch, 512, 512), device='cuda', generator=g, dtype=torch.uint8)
    for _ in range(max(3, warmup)):
        tr.train_step(feat, skip, labels)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    l0 = tr.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        tr.train_step(feat, skip, labels)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device='cuda', dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    loss = tr.loss()
    # replicas hold identical weights after every step: compare a digest of the fp32 master weights over the ranks
    digest = tr.weights_digest() >> 2
    same = True
    if world > 1:
        d = torch.tensor([digest], device='cuda', dtype=torch.int64)
        lo, hi = d.clone(), d.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        same = bool(lo.item() == hi.item())
    spans = [tr.stats_span(g_) for g_ in tr.FWD_GROUPS] + [tr.bn_grad_span(g_) for g_ in tr.BWD_GROUPS]
    rec = {'workload': 'cfg5: Xception OS16 512x512 head training step, SyncBN + Dropout + sparse CE + SGD, %d img/GPU, global batch %d, bf16' % (batch, batch * world),
           'ms_per_step': ms / steps, 'images_per_s': batch * world * steps / (ms / 1000.0), 'steps': steps, 'n_gpus': world,
           'gpu_launches_per_step': (tr.launches - l0) // steps, 'cuda_graph': True, 'loss': loss,
           'collectives_per_step': {'syncbn_allreduce': len(spans), 'grad_allreduce': 1, 'backend': tr.comm_backend()},
           'bytes_per_collective': {'syncbn': [4 * (e - b) for b, e in spans], 'grad_bucket': 4 * (tr.bucket_span()[1] - tr.bucket_span()[0])},
           'replicas_identical': same, 'scope': 'head only (backbone frozen / outside, train.py stage 1)',
           'host': 'dlv3p_trainer_step (C ABI): one CUDA graph per step, no torch / NCCL call inside the step'}
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    tr.close()
    return rec


KERNEL_OF = [('_shortcut_sample', 'subsample2_kernel'), ('_sepconv', 'bb_sepconv_kernel'), ('_depthwise', 'bb_depthwise_kernel'), ('_pointwise', 'bb_gemm_kernel'),
             ('_shortcut', 'bb_gemm_kernel'), ('entry_flow_conv1_1', 'stem_tc_kernel'), ('entry_flow_conv1_2', 'conv3x3_c32_kernel')]
HEAD_GEMMS = ('aspp_branches_gemm', 'concat_projection_gemm', 'feature_projection0_gemm', 'decoder_conv0_sepconv', 'decoder_conv1_sepconv', 'classifier_gemm')


def kernel_of(op_name):
    if not op_name.startswith(('entry_flow', 'middle_flow', 'exit_flow')):
        return op_name      # head launches carry their own names
    for suffix, k in KERNEL_OF:
        if op_name.endswith(suffix):
            return k
    return op_name      # head launches carry their own names


def summarize_profile(prof_runs, B, peaks):
    """prof_runs: list of per-forward [(name, ms, flops, bytes)].  Aggregates by kernel function; FLOPs / bytes are the algorithmic
    figures the library reports per launch (backbone) or the table above (head)."""
    n = len(prof_runs)
    agg = {}
    per_op = {}
    for run in prof_runs:
        for name, ms, fl, by in run:
            if fl == 0 and name in GEMM_FLOP_PER_IMG:
                fl = GEMM_FLOP_PER_IMG[name] * B
            if by == 0:
                by = (HBM_BYTES_PER_IMG.get(name) or ALGO_BYTES_PER_IMG.get(name, 0)) * B
            k = kernel_of(name)
            a = agg.setdefault(k, {'ms': 0.0, 'flops': 0.0, 'bytes': 0.0, 'launches': 0})
            a['ms'] += ms / n; a['flops'] += fl / n; a['bytes'] += by / n; a['launches'] += 1.0 / n
            per_op[name] = per_op.get(name, 0.0) + ms / n
    return agg, per_op


def run_whole_model(args, rank, world, local_rank):
    """The headline: DeepLabV3+ Xception OS16 512x512 forward, WHOLE MODEL (backbone + head), uint8 images in, uint8 labels out."""
    import torch
    import dlv3p_b200
    from dlv3p_b200 import ffi
    from tools import torch_plumbing
    B = args.batch
    if args.strong:
        B = max(1, dlv3p_b200.sharding.shard_batch(args.batch, world, rank)[1])
    net = dlv3p_b200.DeepLabV3PlusXception((CFG['H'], CFG['W'], 3), CFG['NC'], CFG['OS'], batch=B, device=local_rank)
    net.set_weights(random_weights(net.weight_specs()))
    mdl = net.model
    g = torch.Generator(device='cuda').manual_seed(1236 + rank)
    images = torch.randint(0, 256, (B, CFG['H'], CFG['W'], 3), generator=g, device='cuda', dtype=torch.uint8)
    out = torch.empty((B, CFG['H'], CFG['W']), device='cuda', dtype=torch.uint8)
    stream = torch.cuda.current_stream()
    sp = stream.cuda_stream

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def step():
        mdl.forward(images.data_ptr(), out.data_ptr(), sp)

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    launches_per_step = mdl.launch_count()
    prof_runs = [mdl.profile(images.data_ptr(), out.data_ptr(), sp) for _ in range(max(3, min(5, args.steps)))]
    # ---- end to end through the public host API: pinned uint8 images -> H2D -> backbone + head -> D2H labels, every step
    ib, ob = mdl.input_bytes(), mdl.output_bytes()
    pin_i, pin_o = ffi.PinnedBuffer(ib), ffi.PinnedBuffer(ob)
    hi, ho = pin_i.view(np.uint8, (B, CFG['H'], CFG['W'], 3)), pin_o.view(np.uint8, (B, CFG['H'], CFG['W']))
    hi[:] = images.cpu().numpy()
    e2e_steps = max(3, min(args.steps, 10))
    mdl.forward_host(hi, ho)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        mdl.forward_host(hi, ho)
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1000.0
    labels_ok = bool(np.array_equal(ho, out.cpu().numpy()))
    ms, e2e_ms = torch_plumbing.max_over_ranks([ms, e2e_ms], device='cuda')
    ws = mdl.workspace_bytes()
    net.close()
    del images, out
    torch.cuda.empty_cache()
    return dict(B=B, ms=ms, clocks=clocks, launches_per_step=launches_per_step, prof_runs=prof_runs, ib=ib, ob=ob, e2e_steps=e2e_steps, e2e_ms=e2e_ms,
                labels_ok=labels_ok, workspace_bytes=ws)


def time_whole_model_cfg3(local_rank, steps=5):
    """BASELINE configs[2], whole model: Xception OS8 1024x2048, 19 classes, batch 8 (N=1 only)."""
    import torch
    import dlv3p_b200
    try:
        net = dlv3p_b200.DeepLabV3PlusXception((1024, 2048, 3), 19, 8, batch=8, device=local_rank)
        net.set_weights(random_weights(net.weight_specs()))
        images = torch.randint(0, 256, (8, 1024, 2048, 3), device='cuda', dtype=torch.uint8)
        out = torch.empty((8, 1024, 2048), device='cuda', dtype=torch.uint8)
        sp = torch.cuda.current_stream().cuda_stream
        for _ in range(3):
            net.model.forward(images.data_ptr(), out.data_ptr(), sp)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            net.model.forward(images.data_ptr(), out.data_ptr(), sp)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        prof = net.model.profile(images.data_ptr(), out.data_ptr(), sp)
        flops = sum(p[2] for p in prof) + 200.99e9 * 8
        agg = {}
        for name, t, fl, by in prof:
            agg[kernel_of(name)] = agg.get(kernel_of(name), 0.0) + t
        dom = max(agg, key=agg.get)
        rec = {'ms_per_step': ms, 'images_per_s': 8 / ms * 1000.0, 'batch': 8, 'conv_tflops': flops / (ms / 1000.0) / 1e12, 'dominant_kernel': dom,
               'dominant_kernel_ms': agg[dom], 'scope': 'whole model (Xception OS8 backbone + head)', 'workspace_gb': net.model.workspace_bytes() / 1e9}
        net.close()
        del images, out
        torch.cuda.empty_cache()
        return rec
    except Exception as e:
        return {'error': repr(e)[:200]}


def cpu_whole_model_throughput(budget_s=20.0, batch=1, max_iters=8):
    """The reference path restated with torch-CPU ops (oracle/xception_ref.py + oracle/head_ref.py), fp32, all host threads:
    normalize_image -> Xception_body -> ASPP -> Decoder -> tail -> argmax.  'Restatement, not TensorFlow' (not installable here)."""
    import torch
    from oracle import head_ref as R
    from oracle import xception_ref as X
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    Wb = X.make_weights(CFG['OS'], 4321)
    cfg = R.HeadConfig(B=batch, H=CFG['H'], W=CFG['W'], OS=CFG['OS'], Cin=CFG['Cin'], Cskip=CFG['Cskip'], NC=CFG['NC'])
    Wh = R.make_weights(cfg, 1234)
    img = np.random.default_rng(5).integers(0, 256, (batch, CFG['H'], CFG['W'], 3)).astype(np.uint8)

    def once():
        f, s = X.forward_torch(R.normalize_image(img), Wb, CFG['OS'], 'fp32')
        return R.head_forward_torch(f, s, Wh, cfg, 'fp32')

    once()
    t0 = time.perf_counter()
    it = 0
    while it < max_iters and (it < 2 or time.perf_counter() - t0 < budget_s):
        once()
        it += 1
    dt = time.perf_counter() - t0
    return {'value': batch * it / dt, 'unit': 'images/s', 'cores': torch.get_num_threads(), 'kind': 'port',
            'sample': '%d iterations of batch %d of the same workload (whole model), fp32 torch-CPU restatement incl. softmax+argmax (%.1f s)' % (it, batch, dt),
            'host_cpus': os.cpu_count()}


def run_reference(args, rank, world):
    if rank != 0:
        return
    steps = max(1, args.steps)
    res = cpu_whole_model_throughput(budget_s=min(120.0, 6.0 * steps), batch=1, max_iters=max(2, steps))
    line = {'impl': 'reference', 'metric': METRIC, 'value': res['value'], 'unit': 'images/s', 'n_gpus': args.gpus, 'steps': steps,
            'warmup': args.warmup, 'ms_per_step': 1000.0 / res['value'], 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'reference_arm': 'oracle port (whole model) on host cores, bounded sample: batch 1 per step'},
            'cpu_baseline': res,
            'e2e': {'value': res['value'], 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='dlv3p', choices=['dlv3p', 'reference'])
    ap.add_argument('--batch', type=int, default=CFG['B'], help='per-GPU batch (default: the BASELINE config)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-train', action='store_true', help='skip the cfg-5 training-step sub-record')
    ap.add_argument('--no-other-configs', action='store_true', help='skip the timings of BASELINE configs 1/3/4 (N=1 only)')
    ap.add_argument('--strong', action='store_true', help='strong scaling: the global batch stays --batch, each rank takes batch / N images')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))

    if args.impl == 'reference':
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import dlv3p_b200

    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device — the hot path has no CPU fallback (use --impl reference for the CPU arm)')
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')   # keep stdout to the ONE JSON line (NCCL prints its banner there)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))

    R_ = run_whole_model(args, rank, world, local_rank)
    B, ms = R_['B'], R_['ms']
    other = None
    if world == 1 and not args.no_other_configs and not args.strong:
        other = time_other_configs(local_rank)
        other['cfg3_xception_os8_1024x2048_b8_whole_model'] = time_whole_model_cfg3(local_rank)
    train_rec = None
    if not args.no_train and not args.strong:
        try:
            train_rec = time_train_step(rank, world, local_rank, steps=max(3, min(args.steps, 10)), warmup=args.warmup)
        except Exception as e:
            train_rec = {'error': repr(e)[:300]}

    if rank == 0:
        peaks = load_peaks()
        step_s = ms / args.steps / 1000.0
        value = (args.batch * args.steps / (ms / 1000.0)) if args.strong else dlv3p_b200.sharding.aggregate_throughput(B, args.steps, world, ms)
        agg, per_op = summarize_profile(R_['prof_runs'], B, peaks)
        sum_ms = sum(a['ms'] for a in agg.values())
        dom = max(agg, key=lambda k: agg[k]['ms'])
        d = agg[dom]
        tensor_bound = dom in ('bb_gemm_kernel', 'conv3x3_c32_kernel', 'bb_sepconv_kernel') or dom in HEAD_GEMMS
        if tensor_bound:
            ach = d['flops'] / (d['ms'] / 1000.0) / 1e12
            roof = {'kernel': dom, 'bound': 'tensor', 'achieved': ach, 'peak': peaks['tflops_burst'], 'unit': 'TFLOP/s', 'frac': ach / peaks['tflops_burst'],
                    'frac_of_sustained': ach / peaks['tflops_sustained'],
                    'peak_source': peaks['src'] + ', burst bf16 (kernels event-timed inside a sub-second region)'}
        else:
            ach = d['bytes'] / (d['ms'] / 1000.0) / 1e9
            roof = {'kernel': dom, 'bound': 'hbm', 'achieved': ach, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s', 'frac': ach / peaks['hbm_gbs'], 'peak_source': peaks['src']}
        roof.update({'launches_per_step': round(d['launches']), 'kernel_ms_per_step': d['ms'], 'avg_launch_ms': d['ms'] / max(d['launches'], 1),
                     'algorithmic_flops_per_step': d['flops'], 'algorithmic_bytes_per_step': d['bytes'], 'share_of_step': d['ms'] / sum_ms,
                     'traffic': NCU_TRAFFIC.get(dom), 'traffic_note': NCU_TRAFFIC_NOTE if NCU_TRAFFIC.get(dom) else 'no ncu capture for this kernel yet'})
        total_flops = sum(a['flops'] for a in agg.values())
        kernels = {}
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1]['ms']):
            e = {'ms': round(a['ms'], 4), 'launches': round(a['launches'])}
            if a['flops']:
                e['tflops'] = round(a['flops'] / (a['ms'] / 1000.0) / 1e12, 1)
            if a['bytes']:
                e['algorithmic_gbs'] = round(a['bytes'] / (a['ms'] / 1000.0) / 1e9, 1)
            kernels[k] = e
        head_ms = sum(per_op.get(k, 0.0) for k in per_op if kernel_of(k) == k)
        line = {
            'metric': METRIC, 'value': value, 'unit': 'images/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'strong' if args.strong else 'weak', 'vs_baseline': None, 'dtype': 'bf16',
            'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'batch_per_gpu': B, 'global_batch': args.batch if args.strong else B * world,
                       'sharding': 'batch (independent images, no collective)',
                       'l2': 'activations of one step (%.1f GB workspace) far exceed the 126 MB L2; every step streams them again (no flush needed)' % (R_['workspace_bytes'] / 1e9),
                       'input': 'uint8 RGB images [B,512,512,3], normalised inside the first convolution', 'output': 'uint8 label maps [B,512,512]'},
            'e2e': {'value': B * world * R_['e2e_steps'] / (R_['e2e_ms'] / 1000.0), 'unit': 'images/s', 'h2d_bytes_per_step': R_['ib'],
                    'd2h_bytes_per_step': R_['ob'], 'steps': R_['e2e_steps'], 'labels_equal_device_path': R_['labels_ok']},
            'gpu_launches': int(R_['launches_per_step'] * args.steps),
            'clocks': R_['clocks'],
            'roofline': roof,
            'whole_step': {'conv_tflops': total_flops / step_s / 1e12, 'frac_of_tensor_peak': total_flops / step_s / 1e12 / peaks['tflops_burst'],
                           'algorithmic_flops_per_image': total_flops / B, 'sum_kernel_ms': sum_ms, 'launches': R_['launches_per_step'],
                           'backbone_ms': sum_ms - head_ms, 'head_ms': head_ms},
            'aspp': aspp_block(per_op, B, peaks),
            'kernels': kernels,
        }
        if other is not None:
            line['other_configs'] = other
        if train_rec is not None:
            line['train'] = train_rec
        if world == 1 and not args.no_cpu_baseline:
            line['cpu_baseline'] = cpu_whole_model_throughput()
        print(json.dumps(line))
        sys.stdout.flush()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        dist.destroy_process_group()


if __name__ == '__main__':
    if '--train' in sys.argv:      # cfg-5 training step of the head: tools/trainbench.py (same launch conventions; its own JSON line)
        import runpy
        sys.argv = [a for a in sys.argv if a != '--train']
        keep = [sys.argv[0]]
        i = 1
        while i < len(sys.argv):
            if sys.argv[i] in ('--steps', '--warmup', '--batch') and i + 1 < len(sys.argv):
                keep += sys.argv[i:i + 2]
                i += 2
            else:
                i += 2 if sys.argv[i] == '--gpus' else 1
        sys.argv = keep
        runpy.run_path(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'tools', 'trainbench.py'), run_name='__main__')
    else:
        main()
