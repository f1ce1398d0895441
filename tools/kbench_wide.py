#!/usr/bin/env python3
"""Attribution runs for the fused middle-flow SepConv_BN (dlv3p_op_bb_time op 2) beside the two kernels it replaces.
debug flags of the fused kernel: 1 = no output stores, 2 = no stencil math, 4 = no MMAs.  Writes gpurun_out/kbench_wide.txt."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dlv3p_b200 import ffi  # noqa: E402

lines = []


def say(s):
    print(s, flush=True)
    lines.append(s)


B, H, W, C, N = 32, 32, 32, 728, 728
M = B * H * W
for res in (0, 1):
    g = ffi.op_bb_time(0, [M, C, N, res, 0], 20, 0)
    d = ffi.op_bb_time(1, [B, H, W, C, 1, 1], 20, 0)
    say('unfused res=%d: depthwise %.4f ms + gemm %.4f ms = %.4f ms' % (res, d, g, d + g))
    for flags in (0, 1, 2, 4, 6, 7):
        ms = ffi.op_bb_time(2, [B, H, W, C, N, res], 20, flags)
        say('fused   res=%d flags=%d: %.4f ms  %.1f TFLOP/s (GEMM flops)' % (res, flags, ms, 2.0 * M * C * N / ms / 1e9))
for Bx in (8, 16, 37, 64):
    ms = ffi.op_bb_time(2, [Bx, H, W, C, N, 1], 20, 0)
    g = ffi.op_bb_time(0, [Bx * H * W, C, N, 1, 0], 20, 0)
    d = ffi.op_bb_time(1, [Bx, H, W, C, 1, 1], 20, 0)
    say('batch %d: fused %.4f ms, unfused %.4f ms' % (Bx, ms, d + g))
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
open(os.path.join(ROOT, 'gpurun_out', 'kbench_wide.txt'), 'w').write('\n'.join(lines) + '\n')
