#!/usr/bin/env python3
"""Attribution of aspp_dw_gather_kernel at the Cityscapes geometry (cfg 3: 128 x 256 x 2048 feature map, batch 8, rates 12/24/36):
config flag bits 8-10 are the kernel's debug bits (1 = no stores, 2 = gather only, 4 = no gather).  Writes gpurun_out/kbench_gather.txt."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dlv3p_b200 import ffi  # noqa: E402
from oracle import head_ref as R  # noqa: E402  (weight shapes only)

B, h, w, C, OS = 8, 128, 256, 2048, 8
lines = []
feat = ffi.DeviceBuffer(B * h * w * C * 2, 0)
out = ffi.DeviceBuffer(B * h * w * 256 * 4, 0)
DBGS = [int(a) for a in sys.argv[1:]] or [0, 1, 2, 4, 6]
for dbg in DBGS:
    ctx = ffi.Context(device=0, B=B, H=h * OS, W=w * OS, OS=OS, h=h, w=w, Cin=C, Cskip=0, NC=1, variant=ffi.VARIANT_ASPP, stages=ffi.STAGE_ASPP,
                      in_dtype=ffi.DTYPE_BF16, out_mode=ffi.OUT_FEATURES_FP32, flags=dbg << 8)
    rng = np.random.default_rng(0)
    for layer, var, shape in ctx.weight_specs():
        a = rng.standard_normal(shape).astype(np.float32) * 0.05
        if var == 'moving_variance':
            a = np.abs(a) + 1.0
        ctx.set_weight(layer, var, a)
    ctx.finalize()
    for _ in range(2):
        ctx.forward(feat.ptr, None, out.ptr)
    ffi.synchronize(0)
    runs = [dict(ctx.profile(feat.ptr, None, out.ptr)) for _ in range(5)]
    ms = float(np.mean([r['aspp_dw_pool'] for r in runs]))
    by = 2.0 * B * h * w * C * 4
    s = 'gather debug=%d: aspp_dw_pool %.4f ms  (%.0f GB/s algorithmic)  | branches gemm %.4f ms' % (dbg, ms, by / ms / 1e6, float(np.mean([r['aspp_branches_gemm'] for r in runs])))
    print(s, flush=True)
    lines.append(s)
    ctx.close()
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
open(os.path.join(ROOT, 'gpurun_out', 'kbench_gather.txt'), 'w').write('\n'.join(lines) + '\n')
