"""dlv3p_b200.h5lite (the dependency-free HDF5 reader behind load_weights / tools/h5_to_npz.py, reference model.py:102-103) on files
laid out the way libhdf5 lays out Keras weight files: version-0 superblock, symbol-table groups with multi-level B-trees, version-1
object headers, contiguous float datasets, string-array and variable-length string attributes (tests/h5_writer.py)."""
import importlib.util
import os

import numpy as np
import pytest

from dlv3p_b200 import ffi, h5lite
from oracle import head_ref as R
from tests.h5_writer import write_h5

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _keras_tree(W, scope_suffix=''):
    """{(layer, var): array} -> the tree Keras writes: /<layer>/<layer><suffix>/<var>:0, plus layer_names / weight_names attributes."""
    tree, attrs = {}, {}
    layers = []
    for (layer, var), a in W.items():
        if layer not in tree:
            tree[layer] = {layer + scope_suffix: {}}
            layers.append(layer)
        tree[layer][layer + scope_suffix][var + ':0'] = np.asarray(a, np.float32)
    for layer in layers:
        names = ['%s%s/%s' % (layer, scope_suffix, k) for k in tree[layer][layer + scope_suffix]]
        attrs['/model_weights/' + layer] = {'weight_names': np.array([n.encode() for n in names])}
    attrs['/model_weights'] = {'layer_names': np.array([n.encode() for n in layers]), 'backend': 'tensorflow', 'keras_version': '2.4.0'}
    return {'model_weights': tree}, attrs


def test_roundtrip_types_shapes_attributes(tmp_path):
    rng = np.random.default_rng(0)
    tree = {
        'g': {'f32': rng.standard_normal((3, 3, 7, 5)).astype(np.float32), 'f64': rng.standard_normal((4, 2)), 'scalar': np.float32(2.5),
              'empty_group': {}, 'deep': {'deeper': {'i32': np.arange(-5, 7, dtype=np.int32), 'u8': np.arange(200, dtype=np.uint8)}}},
        'big_endian': rng.standard_normal(9).astype('>f4'),
        'f16': rng.standard_normal(6).astype(np.float16),
    }
    attrs = {'/': {'backend': 'tensorflow', 'names': np.array([b'alpha', b'be', b'gamma_long_name'])}, '/g': {'n': np.int64(7), 'w': np.float32(0.5)}}
    p = str(tmp_path / 't.h5')
    write_h5(p, tree, attrs)
    with h5lite.File(p) as f:
        assert f.keys() == ['big_endian', 'f16', 'g']
        assert f.attrs['backend'] == 'tensorflow'
        assert [x.decode() for x in f.attrs['names']] == ['alpha', 'be', 'gamma_long_name']
        g = f['g']
        assert g.attrs['n'] == 7 and g.attrs['w'] == np.float32(0.5)
        assert g['f32'].shape == (3, 3, 7, 5) and g['f32'].dtype == np.float32
        np.testing.assert_array_equal(g['f32'][()], tree['g']['f32'])
        np.testing.assert_array_equal(f['g/f64'][()], tree['g']['f64'])
        assert f['g/scalar'][()] == np.float32(2.5) and f['g/scalar'].shape == ()
        np.testing.assert_array_equal(f['g/deep/deeper/i32'][()], tree['g']['deep']['deeper']['i32'])
        np.testing.assert_array_equal(f['/g/deep/deeper/u8'][()], tree['g']['deep']['deeper']['u8'])
        np.testing.assert_array_equal(f['big_endian'][()].astype(np.float32), tree['big_endian'].astype(np.float32))
        np.testing.assert_array_equal(f['f16'][()], tree['f16'])
        assert f['g/empty_group'].keys() == [] and 'nope' not in f and 'g/deep' in f
        with pytest.raises(KeyError):
            f['g/missing']
        assert sorted(k for k, _ in h5lite.walk_datasets(f)) == sorted(['big_endian', 'f16', 'g/f32', 'g/f64', 'g/scalar', 'g/deep/deeper/i32', 'g/deep/deeper/u8'])


@pytest.mark.parametrize('n', [1, 8, 9, 257, 700])
def test_groups_larger_than_one_symbol_node_and_one_btree_level(tmp_path, n):
    """8 links fill a symbol node, 256 one level-0 B-tree node: 700 links need a two-level tree (a whole Xception has ~330 layers)."""
    tree = {'layer_%04d' % i: {'kernel:0': np.full((2, 3), i, np.float32)} for i in range(n)}
    data = write_h5(None, tree)
    f = h5lite.File(data)
    assert len(f.keys()) == n and len(f) == n
    for i in (0, n // 2, n - 1):
        np.testing.assert_array_equal(f['layer_%04d/kernel:0' % i][()], np.full((2, 3), i, np.float32))


def test_unsupported_layouts_and_corrupt_files_fail_loudly(tmp_path):
    a = np.ones((4, 4), np.float32)
    f = h5lite.File(write_h5(None, {'d': a, 'c': a}, chunked=('/c',)))
    np.testing.assert_array_equal(f['d'][()], a)
    with pytest.raises(h5lite.H5Error, match='chunked'):
        f['c']
    with pytest.raises(h5lite.H5Error, match='not an HDF5 file'):
        h5lite.File(b'PK\x03\x04' + bytes(600))
    good = write_h5(None, {'d': a})
    with pytest.raises(h5lite.H5Error):
        h5lite.File(good[:len(good) // 2])['d'][()]
    with pytest.raises(h5lite.H5Error, match='read-only'):
        h5lite.File(good, 'w')


@pytest.mark.parametrize('scope_suffix,save_kind', [('', 'model'), ('_1', 'weights')])
def test_keras_weight_file_of_the_head_by_name(tmp_path, scope_suffix, save_kind):
    """A head weight file in both Keras layouts (model.save: /model_weights/...; save_weights: at the root; inner scopes may carry a
    '_1' suffix after a second model build) -> keras_weights -> exactly the head's inventory; tools/h5_to_npz.py on the same file."""
    cfg = R.HeadConfig(B=1, H=64, W=64, OS=16, Cin=320, Cskip=24, NC=21)
    W = R.make_weights(cfg, 11)
    W[('entry_flow_conv1_1', 'kernel')] = np.zeros((3, 3, 3, 32), np.float32)      # a backbone layer: ignored by the head export
    tree, attrs = _keras_tree(W, scope_suffix)
    if save_kind == 'weights':
        tree = tree['model_weights']
        attrs = {k.replace('/model_weights', '') or '/': v for k, v in attrs.items()}
    p = str(tmp_path / 'w.h5')
    write_h5(p, tree, attrs)
    got = h5lite.keras_weights(p)
    assert set(got) == {'%s/%s' % k for k in W}
    for (layer, var), a in W.items():
        np.testing.assert_array_equal(got['%s/%s' % (layer, var)], a)
    f = h5lite.File(p)
    root = f['model_weights'] if save_kind == 'model' else f
    assert [x.decode() for x in root.attrs['layer_names']][:2] == list(dict.fromkeys(k[0] for k in W))[:2]
    assert root.attrs['keras_version'] == '2.4.0'
    spec = importlib.util.spec_from_file_location('h5_to_npz', os.path.join(ROOT, 'tools', 'h5_to_npz.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    out = str(tmp_path / 'head.npz')
    assert mod.main(['h5_to_npz', p, out]) == 0
    with np.load(out) as z:
        assert 'entry_flow_conv1_1/kernel' not in z.files and 'aspp0/kernel' in z.files
        for k in z.files:
            layer, var = k.split('/')
            np.testing.assert_array_equal(z[k], W[(layer, var)])
    out2 = str(tmp_path / 'all.npz')
    assert mod.main(['h5_to_npz', '--all', p, out2]) == 0
    with np.load(out2) as z:
        assert 'entry_flow_conv1_1/kernel' in z.files


def test_whole_model_inventory_from_an_h5_file(tmp_path):
    """Every variable of DeepLabV3+ Xception (backbone + head, ~330 layers: a two-level group B-tree) written as a Keras .h5 and read
    back by name: the file covers the library's weight inventory exactly (plan-only model: no GPU needed)."""
    m = ffi.Model(device=-1, B=1, H=64, W=64, OS=16, NC=21, img_dtype=0, out_mode=ffi.OUT_LABELS_U8, flags=0)
    specs = m.weight_specs()
    rng = np.random.default_rng(5)
    W = {(layer, var): rng.standard_normal(shape).astype(np.float32) for layer, var, shape in specs}
    tree, attrs = _keras_tree(W)
    p = str(tmp_path / 'xception.h5')
    write_h5(p, tree, attrs)
    got = h5lite.keras_weights(p)
    assert len(got) == len(specs) and len({l for l, _, _ in specs}) > 256
    for layer, var, shape in specs:
        a = got['%s/%s' % (layer, var)]
        assert a.shape == tuple(shape)
        np.testing.assert_array_equal(a, W[(layer, var)])
    m.close()
