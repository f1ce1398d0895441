#!/usr/bin/env python3
"""Determinism stress: N forwards of a configuration on fixed inputs, every result compared bit for bit with the first
(a race in a barrier protocol shows up as a rare mismatch).  stress.py [iters]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dlv3p_b200  # noqa: E402
from bench import random_weights  # noqa: E402
from tools.cfgbench import CFGS  # noqa: E402


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    for name in ('2', '1', '3', '4a', '4b'):
        c = CFGS[name]
        head = dlv3p_b200.DeepLabHead(c['B'], c['H'], c['W'], c['OS'], c['Cin'], c['Cskip'], c['NC'], lite=c['lite'], decoder=c['decoder'], device=0)
        head.set_weights(random_weights(head.weight_specs()))
        h, w = c['H'] // c['OS'], c['W'] // c['OS']
        feat = torch.randn((c['B'], h, w, c['Cin']), device='cuda').clamp_(min=0).to(torch.bfloat16)
        skip = torch.randn((c['B'], c['H'] // 4, c['W'] // 4, c['Cskip']), device='cuda').to(torch.bfloat16) if c['decoder'] else None
        out = torch.empty((c['B'], c['H'], c['W']), device='cuda', dtype=torch.uint8)
        sp = torch.cuda.current_stream().cuda_stream
        sk = skip.data_ptr() if skip is not None else 0
        head.ctx.forward(feat.data_ptr(), sk, out.data_ptr(), sp)
        torch.cuda.synchronize()
        first = out.clone()
        n = iters if name == '2' else max(10, iters // 5)
        bad = 0
        for _ in range(n):
            out.zero_()
            head.ctx.forward(feat.data_ptr(), sk, out.data_ptr(), sp)
            torch.cuda.synchronize()
            bad += int(not torch.equal(out, first))
        print('cfg %-3s %4d forwards, %d differ from the first' % (name, n, bad), flush=True)
        head.close()
        assert bad == 0


if __name__ == '__main__':
    main()
