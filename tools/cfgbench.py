#!/usr/bin/env python3
"""Per-kernel device times of the head for every BASELINE.json configuration (synthetic inputs resident in HBM).
Writes gpurun_out/cfgbench.txt.  cfgbench.py [cfg ...]   cfg in 1 2 3 4a 4b"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dlv3p_b200  # noqa: E402
from bench import random_weights  # noqa: E402

CFGS = {
    '1': dict(B=1, H=512, W=512, OS=16, Cin=320, Cskip=24, NC=21, lite=False, decoder=True),
    '2': dict(B=32, H=512, W=512, OS=16, Cin=2048, Cskip=256, NC=21, lite=False, decoder=True),
    '3': dict(B=8, H=1024, W=2048, OS=8, Cin=2048, Cskip=256, NC=19, lite=False, decoder=True),
    '4a': dict(B=64, H=512, W=512, OS=16, Cin=160, Cskip=24, NC=21, lite=True, decoder=False),
    '4b': dict(B=64, H=512, W=512, OS=16, Cin=160, Cskip=24, NC=21, lite=True, decoder=True),
}


def main():
    sel = sys.argv[1:] or list(CFGS)
    lines = []
    for name in sel:
        c = CFGS[name]
        head = dlv3p_b200.DeepLabHead(c['B'], c['H'], c['W'], c['OS'], c['Cin'], c['Cskip'], c['NC'], lite=c['lite'], decoder=c['decoder'], device=0)
        head.set_weights(random_weights(head.weight_specs()))
        ctx = head.ctx
        h, w = c['H'] // c['OS'], c['W'] // c['OS']
        feat = torch.randn((c['B'], h, w, c['Cin']), device='cuda').clamp_(min=0).to(torch.bfloat16)
        skip = torch.randn((c['B'], c['H'] // 4, c['W'] // 4, c['Cskip']), device='cuda').to(torch.bfloat16) if c['decoder'] else None
        out = torch.empty((c['B'], c['H'], c['W']), device='cuda', dtype=torch.uint8)
        sp = torch.cuda.current_stream().cuda_stream
        sk = skip.data_ptr() if skip is not None else 0
        for _ in range(3):
            ctx.forward(feat.data_ptr(), sk, out.data_ptr(), sp)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 10
        e0.record()
        for _ in range(n):
            ctx.forward(feat.data_ptr(), sk, out.data_ptr(), sp)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        prof = {}
        for _ in range(5):
            for k, t in ctx.profile(feat.data_ptr(), sk, out.data_ptr(), sp):
                prof.setdefault(k, []).append(t)
        s = 'cfg %-3s B=%d %dx%d OS%d Cin=%d: %.4f ms/step  %.1f img/s | ' % (name, c['B'], c['H'], c['W'], c['OS'], c['Cin'], ms, c['B'] / ms * 1000.0)
        s += '  '.join('%s %.4f' % (k, float(np.mean(v))) for k, v in sorted(prof.items(), key=lambda kv: -np.mean(kv[1])))
        print(s, flush=True)
        lines.append(s)
        head.close()
        del feat, skip, out
        torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(ROOT, 'gpurun_out', 'cfgbench.txt'), 'a') as f:
        f.write('\n'.join(lines) + '\n')


if __name__ == '__main__':
    main()
