"""Generates the committed golden fixtures under tests/golden/.  Run HERE (the container that has
/root/reference); the GPU box only reads the .npz files.

  python tests/golden/make_golden.py

1. head_*.npz     — seeded inputs/weights are regenerated from the seeds stored in the file; the file
                    pins the ORACLE's outputs (fp32 and bf16 modes): low-res logits, labels, tap checksums.
2. example_*.npz  — the reference's example images (example/2007_000039.jpg, 2007_000346.jpg + their VOC
                    label PNGs) pre-processed like the reference does (PIL BICUBIC resize, /127.5-1:
                    common/data_utils.py:436-454) at 256x256 and stored as uint8, plus NEAREST-resized labels.
                    A stand-in backbone (oracle.head_ref.standin_backbone — NOT the reference's) turns them
                    into (feat, skip); the test asserts the oracle and the CUDA path give identical mIoU.

The reference itself pins nothing for this path (no tests, no weights; TensorFlow is not installed):
these vectors pin the restatement, not TensorFlow.  See oracle/head_ref.py header ("parity unpinned").
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import head_ref as R  # noqa: E402

CASES = {
    # name: (HeadConfig kwargs, weight seed, input seed)
    'head_small_full': (dict(B=2, H=64, W=64, OS=16, Cin=64, Cskip=16, NC=21), 11, 12),
    'head_small_os8': (dict(B=1, H=96, W=128, OS=8, Cin=72, Cskip=24, NC=19), 21, 22),
    'head_small_os32': (dict(B=1, H=128, W=128, OS=32, Cin=96, Cskip=16, NC=5), 31, 32),
    'head_small_lite': (dict(B=2, H=128, W=64, OS=16, Cin=96, Cskip=16, NC=21, lite=True, decoder=False), 41, 42),
    'head_small_lite_dec': (dict(B=1, H=64, W=64, OS=16, Cin=160, Cskip=24, NC=21, lite=True, decoder=True), 51, 52),
    'head_odd_size': (dict(B=1, H=100, W=76, OS=16, Cin=64, Cskip=16, NC=21), 61, 62),
}


def checksum(a: np.ndarray) -> np.ndarray:
    a = a.astype(np.float64)
    return np.array([a.sum(), np.abs(a).sum(), (a * a).sum()], np.float64)


def make_head_case(name, kw, wseed, iseed):
    cfg = R.HeadConfig(**kw)
    W = R.make_weights(cfg, wseed)
    feat, skip = R.make_inputs(cfg, iseed)
    out = {'cfg': np.array([repr(kw)]), 'wseed': np.int64(wseed), 'iseed': np.int64(iseed)}
    for mode in ('fp32', 'bf16'):
        t = R.head_forward(feat, skip, W, cfg, mode)
        out['logits_' + mode] = t['logits']
        out['labels_' + mode] = t['labels']
        out['aspp_out_sum_' + mode] = checksum(t['aspp_out'])
        if cfg.decoder:
            out['decoder_out_sum_' + mode] = checksum(t['decoder_out'])
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print(name, {k: getattr(v, 'shape', None) for k, v in out.items() if k.startswith('l')})


def make_example_images(size=256):
    from PIL import Image
    ref = '/root/reference/example'
    for stem in ('2007_000039', '2007_000346'):
        img = Image.open(os.path.join(ref, stem + '.jpg')).convert('RGB')
        img = img.resize((size, size), Image.BICUBIC)                       # preprocess_image, data_utils.py:450
        lab = Image.open(os.path.join(ref, stem + '.png'))
        lab = lab.resize((size, size), Image.NEAREST)                       # mask resize uses nearest (data_utils.py:457-477)
        np.savez_compressed(os.path.join(HERE, 'example_%s.npz' % stem), image=np.asarray(img, np.uint8),
                            label=np.asarray(lab, np.uint8))
        print(stem, np.unique(np.asarray(lab)))


if __name__ == '__main__':
    for name, (kw, ws, is_) in CASES.items():
        make_head_case(name, kw, ws, is_)
    make_example_images()
