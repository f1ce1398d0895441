#!/usr/bin/env python3
"""Attribution runs for the backbone kernels (dlv3p_op_bb_time): GEMM tile widths, residual, stores off; depthwise stores / stencil off.
Writes gpurun_out/kbench_bb.txt."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dlv3p_b200 import ffi  # noqa: E402

lines = []


def say(s):
    print(s, flush=True)
    lines.append(s)


M = 32768
for K, N in ((728, 728), (1536, 2048), (256, 728), (128, 128)):
    Mx = M if K > 200 else 32 * 256 * 256
    for bn in (0, 256, 192, 128):
        for res in (0, 1):
            for flags in (0, 1):
                print('-> gemm', Mx, K, N, bn, res, flags, flush=True)
                ms = ffi.op_bb_time(0, [Mx, K, N, res, bn], 20, flags)
                say('gemm M=%d K=%d N=%d BN=%s res=%d flags=%d: %.4f ms  %.1f TFLOP/s' % (Mx, K, N, bn or 'auto', res, flags, ms, 2.0 * Mx * K * N / ms / 1e9))
for (B, H, W, C, s, r) in ((32, 32, 32, 728, 1, 1), (32, 32, 32, 1536, 1, 2), (32, 256, 256, 128, 1, 1), (32, 256, 256, 128, 2, 1)):
    for flags in (0, 1, 2, 3):
        ms = ffi.op_bb_time(1, [B, H, W, C, s, r], 20, flags)
        by = 2.0 * B * C * (H * W + (H // s) * (W // s))
        say('dw B=%d %dx%d C=%d s=%d r=%d flags=%d: %.4f ms  %.1f GB/s' % (B, H, W, C, s, r, flags, ms, by / ms / 1e6))
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
open(os.path.join(ROOT, 'gpurun_out', 'kbench_bb.txt'), 'w').write('\n'.join(lines) + '\n')
