"""Shared helpers for the parity tests (test infrastructure: may import oracle/)."""
import ast
import os

import numpy as np

from oracle import head_ref as R

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
HEAD_CASES = ['head_small_full', 'head_small_os8', 'head_small_os32', 'head_small_lite', 'head_small_lite_dec', 'head_odd_size']


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name + '.npz'))
    kw = ast.literal_eval(str(z['cfg'][0]))
    cfg = R.HeadConfig(**kw)
    W = R.make_weights(cfg, int(z['wseed']))
    feat, skip = R.make_inputs(cfg, int(z['iseed']))
    return cfg, W, feat, skip, z


def rel_err(a, ref):
    """max|a-ref| / max|ref| per tensor (element-wise relative error is meaningless near zero)."""
    ref = np.asarray(ref, np.float64)
    return float(np.abs(np.asarray(a, np.float64) - ref).max() / max(np.abs(ref).max(), 1e-30))


def make_head(cfg, W, out_mode=0, in_dtype=0, flags=0, device=0):
    """CUDA head for an oracle HeadConfig."""
    import dlv3p_b200
    hd = dlv3p_b200.DeepLabHead(cfg.B, cfg.H, cfg.W, cfg.OS, cfg.Cin, cfg.Cskip, cfg.NC, lite=cfg.lite, decoder=cfg.decoder,
                                out_mode=out_mode, in_dtype=in_dtype, device=device, h=cfg.h, w=cfg.w, hs=cfg.hs, ws=cfg.ws,
                                flags=flags)
    hd.set_weights(W)
    return hd


def planar_to_nhwc(a):
    return np.ascontiguousarray(np.transpose(a, (0, 2, 3, 1)))


def label_agreement(labels, ref_labels, ref_logits_full, logit_tol=1e-2):
    """Label parity that is honest about near-ties (with random-init weights ~1 % of pixels have a
    top1-top2 margin below the logits tolerance itself).  Returns
      overall    fraction of identical labels,
      decided    fraction of identical labels among pixels whose oracle margin exceeds the logits tolerance
                 (logit_tol * max|logit|) — a mismatch there would be a real error,
      worst      largest oracle margin (relative to max|logit|) at a mismatching pixel."""
    ref_logits_full = np.asarray(ref_logits_full, np.float32)
    srt = np.sort(ref_logits_full, axis=-1)
    margin = (srt[..., -1] - srt[..., -2]) / max(float(np.abs(ref_logits_full).max()), 1e-30)
    same = np.asarray(labels) == np.asarray(ref_labels)
    decided = margin > logit_tol
    overall = float(same.mean())
    dec = float(same[decided].mean()) if decided.any() else 1.0
    worst = float(margin[~same].max()) if (~same).any() else 0.0
    return overall, dec, worst
