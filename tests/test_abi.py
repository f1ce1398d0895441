"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/dlv3p.h
declares, and fails loudly (no fallback) without a GPU. No compute calls here."""
import ctypes
import os
import re

import numpy as np
import pytest

import dlv3p_b200
from dlv3p_b200 import ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, 'include', 'dlv3p.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(dlv3p_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    lib = ffi.load_library()
    declared = _header_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), 'libdlv3p.so does not export %s' % name
    bound = {s[0] for s in ffi.SYMBOLS}
    assert bound == set(declared), 'ctypes binding and header disagree: %s' % (bound ^ set(declared))
    assert lib.dlv3p_abi_version() == 1


def test_config_struct_layout_matches_header():
    # 15 int32 + float + int32, no padding
    assert ctypes.sizeof(ffi.Config) == 17 * 4
    text = open(os.path.join(ROOT, 'include', 'dlv3p.h')).read()
    body = re.search(r'typedef struct dlv3p_config \{(.*?)\} dlv3p_config;', text, re.S).group(1)
    body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
    names = []
    for decl in body.split(';'):
        decl = decl.strip()
        if not decl:
            continue
        typ, rest = decl.split(None, 1)
        names += [n.strip() for n in rest.split(',')]
    assert names == [f[0] for f in ffi.Config._fields_]


def test_no_libcuda_link_dependency():
    # the library must load on a machine without the CUDA driver (this container)
    import subprocess
    out = subprocess.run(['ldd', ffi.LIB_PATH], capture_output=True, text=True).stdout
    assert 'libcuda.so' not in out and 'libcudart' not in out and 'torch' not in out


def test_create_fails_loudly_without_gpu_or_bad_config():
    if dlv3p_b200.device_count() == 0:
        with pytest.raises(dlv3p_b200.Dlv3pError) as e:
            ffi.Context(device=0, B=1, H=64, W=64, OS=16, Cin=64, Cskip=16, NC=21)
        assert 'no CPU fallback' in str(e.value)
    # reference: ValueError('invalid output stride', OS) (layers.py:126)
    with pytest.raises(dlv3p_b200.Dlv3pError) as e:
        ffi.Context(device=-1, B=1, H=64, W=64, OS=4, Cin=64, Cskip=16, NC=21)
    assert e.value.status == -1 and 'output stride' in str(e.value)
    with pytest.raises(dlv3p_b200.Dlv3pError):
        ffi.Context(device=-1, B=1, H=64, W=64, OS=16, Cin=60, Cskip=16, NC=21)     # Cin % 8
    with pytest.raises(dlv3p_b200.Dlv3pError):
        ffi.Context(device=-1, B=1, H=64, W=64, OS=16, Cin=64, Cskip=16, NC=300)    # NC > 256
    with pytest.raises(dlv3p_b200.Dlv3pError):
        ffi.Context(device=-1, B=0, H=64, W=64, OS=16, Cin=64, Cskip=16, NC=21)


def test_plan_only_context_cannot_compute():
    c = ffi.Context(device=-1, B=1, H=64, W=64, OS=16, Cin=64, Cskip=16, NC=21)
    for layer, var, shape in c.weight_specs():
        c.set_weight(layer, var, np.zeros(shape, np.float32))
    with pytest.raises(dlv3p_b200.Dlv3pError) as e:
        c.finalize()
    assert e.value.status == -4 and 'no CPU path' in str(e.value)
    with pytest.raises(dlv3p_b200.Dlv3pError):
        c.forward(1, 1, 1)


def test_set_weight_name_and_shape_errors():
    c = ffi.Context(device=-1, B=1, H=64, W=64, OS=16, Cin=64, Cskip=16, NC=21)
    with pytest.raises(dlv3p_b200.Dlv3pError) as e:
        c.set_weight('no_such_layer', 'kernel', np.zeros((1, 1, 64, 256), np.float32))
    assert e.value.status == -6
    with pytest.raises(dlv3p_b200.Dlv3pError) as e:
        c.set_weight('aspp0', 'kernel', np.zeros((1, 1, 32, 256), np.float32))
    assert e.value.status == -6
    c.set_weight('logits_semantic', 'bias', np.zeros((21,), np.float32))     # alias of conv_upsample


def test_sizes_cfg2():
    c = ffi.Context(device=-1, B=32, H=512, W=512, OS=16, Cin=2048, Cskip=256, NC=21)
    assert c.input_bytes() == (32 * 32 * 32 * 2048 * 2, 32 * 128 * 128 * 256 * 2)
    assert c.output_bytes() == 32 * 512 * 512
    assert 1.0e9 < c.workspace_bytes() < 3.0e9
