"""Build libdlv3p.so in-tree with nvcc for sm_100a (cross-compiles without a GPU): one object per translation unit, compiled in
parallel, linked into ONE shared library."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
UNITS = ['dlv3p_api.cu', 'xception_api.cu']          # head + training operators | Xception backbone + whole model
OUT = os.path.join(HERE, 'libdlv3p.so')
OBJ_DIR = os.path.join(HERE, 'build')
INCLUDE = os.path.join(os.path.dirname(HERE), 'include')

ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
CFLAGS = ARCH + ['-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC']
LDFLAGS = ARCH + ['-shared', '-Xcompiler', '-fPIC', '-cudart', 'static']


def find_nvcc() -> str:
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found')


def _deps():
    d = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(('.cu', '.cuh'))]
    return d + [os.path.join(INCLUDE, f) for f in sorted(os.listdir(INCLUDE)) if f.endswith('.h')]


def up_to_date() -> bool:
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(d) <= t for d in _deps())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return OUT
    nvcc = find_nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)

    def compile_one(unit):
        obj = os.path.join(OBJ_DIR, unit.replace('.cu', '.o'))
        cmd = [nvcc] + CFLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', '-o', obj, os.path.join(CSRC, unit)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        return unit, obj, res

    with ThreadPoolExecutor(max_workers=len(UNITS)) as ex:
        results = list(ex.map(compile_one, UNITS))
    objs = []
    for unit, obj, res in results:
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError('nvcc failed compiling %s' % unit)
        if verbose:
            sys.stderr.write(res.stderr)
        objs.append(obj)
    res = subprocess.run([nvcc] + LDFLAGS + ['-o', OUT] + objs, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError('nvcc failed linking libdlv3p.so')
    return OUT


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
