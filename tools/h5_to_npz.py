"""Export the DeepLabV3+ head weights of a reference Keras `.h5` file to the `.npz` the B200 head loads
(`DeepLabHead.load_weights_npz`, `get_deeplabv3p_head(weights_path=...)`, `HeadTrainer(weights=dict(np.load(...)))`).

    python tools/h5_to_npz.py model.h5 head.npz            # the head's layers
    python tools/h5_to_npz.py --all model.h5 model.npz     # every layer (backbone + head: DeepLabV3PlusXception.load_weights)
h5py is used when it is installed; otherwise the file is read by the dependency-free `dlv3p_b200.h5lite` (contiguous float datasets,
the layout Keras writes).  `DeepLabHead.load_weights` / `DeepLabV3PlusXception.load_weights` read `.h5` files directly the same way.

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
        # Keras: /<layer>/<scope>/<variable>:0 — the top-level group is the layer name; the inner scope repeats it, with a '_1' suffix
        # when the model was built twice in one session
        layer = parts[0] if is_head_layer(parts[0]) else parts[-2]
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


def open_h5(path):
    """h5py.File when h5py is importable, else the in-repo reader (same keys() / item interface)."""
    try:
        import h5py
        return h5py.File(path, 'r')
    except ImportError:
        import os
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from dlv3p_b200 import h5lite
        return h5lite.File(path)


def main(argv) -> int:
    every = '--all' in argv
    argv = [a for a in argv if a != '--all']
    if len(argv) != 3:
        print(__doc__)
        return 2
    with open_h5(argv[1]) as f:
        root = f['model_weights'] if 'model_weights' in f else f
        if every:
            from dlv3p_b200 import h5lite
            W = h5lite.keras_weights(f) if isinstance(f, h5lite.Group) else all_weights_from_h5(root)
        else:
            W = head_weights_from_h5(root)
    np.savez(argv[2], **W)
    print('wrote %d tensors to %s' % (len(W), argv[2]))
    return 0


def all_weights_from_h5(root) -> Dict[str, np.ndarray]:
    out: Dict[str, np.ndarray] = {}
    for path, ds in walk_datasets(root):
        parts = path.split('/')
        var = parts[-1].split(':')[0]
        if len(parts) >= 2 and var in VARIABLES:
            out['%s/%s' % (parts[0], var)] = np.asarray(ds[()], np.float32)
    return out


if __name__ == '__main__':
    sys.exit(main(sys.argv))
