#!/usr/bin/env python3
"""Per-operator micro-benchmarks through dlv3p_op_time (CUDA events inside the library, synthetic data).
Flags are the kernels' debug bits (skip stores / stencil / MMA) used to attribute time; writes gpurun_out/kbench.txt."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dlv3p_b200 import ffi  # noqa: E402

OUT = os.path.join(ROOT, 'gpurun_out')


def main():
    os.makedirs(OUT, exist_ok=True)
    lines = []

    def rec(name, op, dims, flags=0, flop=0.0, byts=0.0, iters=20):
        try:
            ms = ffi.op_time(op, dims, iters, flags)
        except Exception as e:  # noqa: BLE001
            lines.append('%-44s flags %d  FAILED %s' % (name, flags, e))
            print(lines[-1], flush=True)
            return
        s = '%-44s flags %d  %8.4f ms' % (name, flags, ms)
        if flop:
            s += '  %7.1f TFLOP/s' % (flop / ms / 1e9)
        if byts:
            s += '  %7.1f GB/s' % (byts / ms / 1e6)
        lines.append(s)
        print(s, flush=True)

    sel = sys.argv[1:] or ['pw', 'sep', 'mem', 'aspp']
    if 'pw' in sel:
        for (M, K, N) in [(32768, 2048, 256), (32768, 1024, 256), (524288, 256, 256), (524288, 304, 256), (524288, 256, 48)]:
            for fl in (0, 8):
                rec('pointwise M=%d K=%d N=%d' % (M, K, N), 0, [M, K, N], fl, flop=2.0 * M * K * N, byts=2.0 * M * (K + N))
    if 'sep' in sel:
        for C in (256, 304):
            M = 32 * 128 * 128
            for fl in ((0, 1, 4, 5, 5 + 32, 5 + 64, 8) if C == 256 else (0, 8)):
                rec('sepconv B=32 128x128 C=%d' % C, 1, [32, 128, 128, C], fl, flop=2.0 * M * C * 256, byts=2.0 * M * (C + 256))
    if 'aspp' in sel:
        n = 32 * 32 * 32 * 2048 * 2.0
        for fl in (0, 1, 8, 9):
            rec('aspp_dw slab B=32 32x32 C=2048', 4, [32, 32, 32, 2048], fl, byts=4 * n)
    if 'mem' in sel:
        rec('resize x4 32x32x256 -> 128x128 (B=32)', 2, [32, 32, 32, 256, 128, 128], 0, byts=32 * (1024 * 512 + 16384 * 512.0))
        rec('resize generic (same shape)', 2, [32, 32, 32, 256, 128, 128], 1, byts=32 * (1024 * 512 + 16384 * 512.0))
        rec('resize_argmax x4 21cls 128->512 (B=32)', 3, [32, 21, 128, 128, 512, 512], 0, byts=32 * (16384 * 84 + 262144.0))
        rec('resize_argmax generic (same shape)', 3, [32, 21, 128, 128, 512, 512], 1, byts=32 * (16384 * 84 + 262144.0))
    with open(os.path.join(OUT, 'kbench.txt'), 'a') as f:
        f.write('\n'.join(lines) + '\n')


if __name__ == '__main__':
    main()
