"""Diagnostics for the training step: per-tensor relative L2 error of every gradient (weights and intermediate activations)
against the fp32 oracle, for random labels (cfg 5's synthetic labels) and for spatially coherent labels."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import head_ref as R, train_ref as TR
from dlv3p_b200 import train, train_ffi


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a.reshape(-1) - b.reshape(-1)) / max(np.linalg.norm(b), 1e-30))


def run(mode, drop):
    cfg = R.HeadConfig(B=2, H=320, W=320, OS=16, Cin=64, Cskip=32, NC=21)
    W = R.make_weights(cfg, 21)
    feat, skip = R.make_inputs(cfg, 22)
    feat, skip = R.bf16_round(feat), R.bf16_round(skip)
    labels = TR.make_labels(cfg, 23)
    if mode == 'coherent':
        yy, xx = np.mgrid[0:cfg.H, 0:cfg.W]
        labels = np.stack([((yy // 32 + xx // 32 + b) % cfg.NC).astype(np.uint8) for b in range(cfg.B)])
    tr = train.HeadTrainer(cfg.B, cfg.H, cfg.W, cfg.OS, cfg.Cin, cfg.Cskip, cfg.NC, W, device=0, seed=5, dropout=drop)
    bf = lambda a: torch.from_numpy(a).cuda().to(torch.bfloat16).contiguous()
    tr.forward_backward(bf(feat), bf(skip), torch.from_numpy(labels).cuda())
    torch.cuda.synchronize()
    keep = train_ffi.dropout_keep_mask(cfg.B * cfg.h * cfg.w * 256, train.dropout_seed(5, 0, 0), drop) if drop > 0 else None
    ref = TR.head_train_forward_backward(feat, skip, labels, W, cfg, keep_mask=keep, drop_rate=drop, mode=os.environ.get('ORACLE_MODE', 'bf16'))
    print('== labels %s dropout %.1f: loss %.6f vs %.6f' % (mode, drop, tr.loss(), ref['loss']))
    g = tr.get_grads()
    for k, v in ref['grads'].items():
        print('  w   %-40s %.4f' % ('/'.join(k), rel(g[k], v)))
    print('  d_feat %.4f  d_skip %.4f' % (rel(tr.T['dfeat'].float().cpu().numpy(), ref['d_feat']), rel(tr.T['dskip'].float().cpu().numpy(), ref['d_skip'])))


if __name__ == '__main__':
    run('random', 0.5)
    run('coherent', 0.0)
