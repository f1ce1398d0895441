"""Export the DeepLabV3+ head weights of a reference Keras `.h5` file to the `.npz` the B200 head loads
(`DeepLabHead.load_weights_npz`, `get_deeplabv3p_head(weights_path=...)`, `HeadTrainer(weights=dict(np.load(...)))`).

Run it where `h5py` exists (it ships with the reference's TensorFlow install; it is not in this image):
    python tools/h5_to_npz.py model.h5 head.npz

Both layouts Keras writes are handled: `model.save()` files keep the weights under `/model_weights`, `model.save_weights()` files at
the root (model.py:102-103 loads either by topology, deeplabv3p_xception.py:237 by name).  Every dataset is visited and matched BY
NAME: a dataset path ending in `<layer>/<variable>:0` (any scope prefix, e.g. `aspp0/aspp0/kernel:0`) whose layer is one of the
head's layers (`logits_semantic` is stored as `conv_upsample`).  Keys of the npz: `"<layer>/<variable>"`.
"""
from __future__ import annotations

import sys
from typing import Dict, Iterable, Tuple

import numpy as np

HEAD_LAYER_PREFIXES = ('image_pooling', 'aspp0', 'aspp1_', 'aspp2_', 'aspp3_', 'concat_projection', 'feature_projection0',
                       'decoder_conv0_', 'decoder_conv1_', 'conv_upsample', 'logits_semantic')
VARIABLES = ('kernel', 'bias', 'depthwise_kernel', 'gamma', 'beta', 'moving_mean', 'moving_variance')


def is_head_layer(name: str) -> bool:
    return any(name == p or name == p + '_BN' or (p.endswith('_') and name.startswith(p)) for p in HEAD_LAYER_PREFIXES)


def walk_datasets(group, prefix: str = '') -> Iterable[Tuple[str, object]]:
    """(path, dataset) for every dataset below an h5py-like group (anything with .keys() and item access; datasets have .shape)."""
    for key in group.keys():
        item = group[key]
        path = prefix + '/' + key if prefix else key
        if hasattr(item, 'keys'):
            yield from walk_datasets(item, path)
        else:
            yield path, item


def head_weights_from_h5(root) -> Dict[str, np.ndarray]:
    out: Dict[str, np.ndarray] = {}
    for path, ds in walk_datasets(root):
        parts = path.split('/')
        if len(parts) < 2:
            continue
        var = parts[-1].split(':')[0]
        layer = parts[-2]
        if var not in VARIABLES or not is_head_layer(layer):
            continue
        if layer == 'logits_semantic':
            layer = 'conv_upsample'
        key = '%s/%s' % (layer, var)
        a = np.asarray(ds[()] if hasattr(ds, '__getitem__') and not isinstance(ds, np.ndarray) else ds, np.float32)
        if key in out and not np.array_equal(out[key], a):
            raise ValueError('two different datasets map to %s' % key)
        out[key] = a
    if 'aspp0/kernel' not in out:
        raise ValueError('no DeepLabV3+ head found (no aspp0/kernel dataset)')
    return out


def main(argv) -> int:
    if len(argv) != 3:
        print(__doc__)
        return 2
    try:
        import h5py
    except ImportError:
        print('h5py is not installed here; run this where the reference (TensorFlow + h5py) is installed', file=sys.stderr)
        return 1
    with h5py.File(argv[1], 'r') as f:
        root = f['model_weights'] if 'model_weights' in f else f
        W = head_weights_from_h5(root)
    np.savez(argv[2], **W)
    print('wrote %d tensors to %s' % (len(W), argv[2]))
    return 0


if __name__ == '__main__':
    sys.exit(main(sys.argv))
