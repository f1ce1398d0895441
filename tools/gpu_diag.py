#!/usr/bin/env python3
"""Staged GPU diagnostics: every stage runs in its own subprocess under a timeout (a trapped / wedged kernel must not
take the rest of the session with it) and appends to gpurun_out/diag.txt.  Usage: python tools/gpu_diag.py [stage ...]"""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'gpurun_out')

PRE = '''
import sys, numpy as np
sys.path.insert(0, %r)
import dlv3p_b200
from dlv3p_b200 import ffi
from oracle import head_ref as R
np.set_printoptions(linewidth=200, precision=4, suppress=True)
def blockmap(err, rb, cb, thr):
    M, N = err.shape
    m = (err > thr)
    mm = m[:M // rb * rb, :N // cb * cb].reshape(M // rb, rb, N // cb, cb).any(axis=(1, 3))
    return mm.astype(int)
''' % ROOT

STAGES = {
    'info': '''
print(dlv3p_b200.device_info(0))
''',
    'gemm_small': '''
for (M, K, N) in [(128, 64, 256), (128, 128, 256), (256, 256, 256), (128, 64, 32), (128, 64, 64), (384, 2048, 256)]:
    rng = np.random.default_rng(0)
    a = R.bf16_round(rng.standard_normal((M, K)).astype(np.float32))
    w = R.bf16_round(rng.standard_normal((K, N)).astype(np.float32) * 0.1)
    got = ffi.bf16_bits_to_f32(ffi.op_pointwise(R.to_bf16_bits(a), w, None, None, relu=False))
    ref = a @ w
    err = np.abs(got - ref)
    print('GEMM', M, K, N, 'max err', err.max(), 'ref max', np.abs(ref).max(), 'bad frac', (err > 0.05).mean())
    if (err > 0.05).any():
        print(' bad block map (8 rows x 32 cols):')
        print(blockmap(err, 8, min(32, N), 0.05)[:16])
        print(' got[0,:8]', got[0, :8], ' ref[0,:8]', ref[0, :8])
        print(' got[1,:8]', got[1, :8], ' ref[1,:8]', ref[1, :8])
''',
    'mem_ops': '''
rng = np.random.default_rng(1)
x = R.bf16_round(rng.standard_normal((2, 32, 32, 64)).astype(np.float32))
k = rng.standard_normal((3, 3, 64, 1)).astype(np.float32) * 0.3
for rate in (1, 6, 18):
    got = ffi.bf16_bits_to_f32(ffi.op_depthwise(R.to_bf16_bits(x), k[..., 0], rate, None, None, relu=False))
    ref = R.depthwise3x3(x, k, rate)
    print('depthwise rate', rate, 'max err', np.abs(got - ref).max())
got = ffi.op_resize_bilinear(R.to_bf16_bits(x), 128, 128)
print('resize exact:', np.array_equal(got, R.to_bf16_bits(R.resize_bilinear(x, (128, 128)))))
lg = rng.standard_normal((2, 32, 32, 21)).astype(np.float32)
ref = R.argmax_labels(R.resize_bilinear(lg, (128, 128))).astype(np.uint8)
got = ffi.op_resize_argmax(np.ascontiguousarray(lg.transpose(0, 3, 1, 2)), 128, 128)
print('argmax x4 exact:', np.array_equal(got, ref), (got != ref).sum())
ref = R.argmax_labels(R.resize_bilinear(lg, (100, 90))).astype(np.uint8)
got = ffi.op_resize_argmax(np.ascontiguousarray(lg.transpose(0, 3, 1, 2)), 100, 90)
print('argmax generic exact:', np.array_equal(got, ref), (got != ref).sum())
''',
    'sepconv': '''
for (B, H, W, C) in [(1, 8, 16, 64), (1, 8, 16, 256), (1, 16, 32, 304), (2, 13, 21, 128)]:
    rng = np.random.default_rng(2)
    x = R.bf16_round(rng.standard_normal((B, H, W, C)).astype(np.float32))
    dk = rng.standard_normal((3, 3, C, 1)).astype(np.float32) * 0.3
    pk = rng.standard_normal((C, 256)).astype(np.float32) * np.float32(np.sqrt(2.0 / C))
    one, zero = np.ones(C, np.float32), np.zeros(C, np.float32)
    got = ffi.bf16_bits_to_f32(ffi.op_sepconv(R.to_bf16_bits(x), dk[..., 0], one, zero, pk, np.ones(256, np.float32), np.zeros(256, np.float32)))
    mid = R.bf16_round(np.maximum(R.depthwise3x3(x, dk, 1), 0))
    ref = np.maximum(mid.reshape(-1, C) @ R.bf16_round(pk), 0).reshape(B, H, W, 256)
    err = np.abs(got - ref)
    print('SEPCONV', B, H, W, C, 'max err', err.max(), 'ref max', np.abs(ref).max(), 'bad frac', (err > 0.1).mean())
    if (err > 0.1).any():
        e2 = err.reshape(-1, 256)
        print(blockmap(e2, 8, 32, 0.1)[:32])
''',
    'head': '''
from tests.common import load_case, make_head, planar_to_nhwc, rel_err
for name in ['head_small_full', 'head_small_lite', 'head_odd_size']:
    cfg, W, feat, skip, z = load_case(name)
    ref = R.head_forward(feat, skip, W, cfg, 'bf16')
    hd = make_head(cfg, W)
    labels = hd(feat, skip)
    for tap in ['image_pooling', 'aspp_out'] + (['decoder_in', 'decoder_conv0', 'decoder_out'] if cfg.decoder else []):
        print(name, tap, 'rel err', rel_err(hd.tap(tap), ref[tap]))
    print(name, 'logits rel err', rel_err(planar_to_nhwc(hd.tap('logits')), ref['logits']), 'labels agree', (labels == ref['labels']).mean())
    hd.close()
''',
}


def main():
    os.makedirs(OUT, exist_ok=True)
    stages = sys.argv[1:] or list(STAGES)
    with open(os.path.join(OUT, 'diag.txt'), 'a') as log:
        for s in stages:
            code = PRE + textwrap.dedent(STAGES[s])
            try:
                r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=240, cwd=ROOT)
                txt = '==== %s (rc %d)\n%s%s\n' % (s, r.returncode, r.stdout, r.stderr[-3000:])
            except subprocess.TimeoutExpired as e:
                txt = '==== %s TIMEOUT\n%s\n' % (s, (e.stdout or b'')[-2000:] if isinstance(e.stdout, bytes) else e.stdout)
            log.write(txt)
            log.flush()
            print(txt)


if __name__ == '__main__':
    main()
