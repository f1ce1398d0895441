#!/usr/bin/env python3
"""Per-launch device times of the whole model (dlv3p_model_profile_forward): ms, algorithmic TFLOP/s and GB/s per kernel launch.
tools/modelprof.py [--batch 32] [--size 512] [--os 16] [--reps 5]   -> gpurun_out/modelprof.txt"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dlv3p_b200  # noqa: E402
from bench import kernel_of, random_weights  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=32)
    ap.add_argument('--size', type=int, default=512)
    ap.add_argument('--width', type=int, default=0)
    ap.add_argument('--os', type=int, default=16)
    ap.add_argument('--reps', type=int, default=5)
    ap.add_argument('--tag', default='')
    ap.add_argument('--no-pdl', action='store_true', help='A/B: launch the backbone kernels without programmatic dependent launch')
    a = ap.parse_args()
    H, W = a.size, a.width or a.size
    net = dlv3p_b200.DeepLabV3PlusXception((H, W, 3), 21, a.os, batch=a.batch, device=0, flags=2 if a.no_pdl else 0)
    net.set_weights(random_weights(net.weight_specs()))
    img = torch.randint(0, 256, (a.batch, H, W, 3), device='cuda', dtype=torch.uint8)
    out = torch.empty((a.batch, H, W), device='cuda', dtype=torch.uint8)
    sp = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        net.model.forward(img.data_ptr(), out.data_ptr(), sp)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        net.model.forward(img.data_ptr(), out.data_ptr(), sp)
    e1.record()
    torch.cuda.synchronize()
    step = e0.elapsed_time(e1) / 10
    runs = [net.model.profile(img.data_ptr(), out.data_ptr(), sp) for _ in range(a.reps)]
    lines = ['%s B=%d %dx%d OS%d: %.4f ms/step = %.1f img/s; sum of kernels %.4f ms' % (a.tag, a.batch, H, W, a.os, step, a.batch / step * 1e3,
                                                                                         sum(np.mean([r[i][1] for r in runs]) for i in range(len(runs[0]))))]
    for i, (name, _, fl, by) in enumerate(runs[0]):
        ms = float(np.mean([r[i][1] for r in runs]))
        lines.append('%-52s %-22s %8.4f ms %8.1f TFLOP/s %8.1f GB/s' % (name, kernel_of(name), ms, fl / ms / 1e9, by / ms / 1e6))
    txt = '\n'.join(lines)
    print(txt)
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(ROOT, 'gpurun_out', 'modelprof%s.txt' % (('_' + a.tag) if a.tag else '')), 'w') as f:
        f.write(txt + '\n')


if __name__ == '__main__':
    main()
