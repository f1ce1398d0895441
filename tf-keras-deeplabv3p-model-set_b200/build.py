"""Build libdlv3p.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'csrc', 'dlv3p_api.cu')
OUT = os.path.join(HERE, 'libdlv3p.so')
DEPS = [os.path.join(HERE, 'csrc', f) for f in ('dlv3p_api.cu', 'pw_gemm.cuh', 'pw_gemm2.cuh', 'dwpw_gemm2.cuh', 'aspp_dw_fast.cuh', 'aspp_dw_gather.cuh', 'bn_train.cuh', 'tgemm.cuh', 'train_kernels.cuh', 'train_api.cuh', 'dwpw_gemm.cuh', 'mem_kernels.cuh', 'sm100_prims.cuh')] + \
       [os.path.join(os.path.dirname(HERE), 'include', 'dlv3p.h'), os.path.join(os.path.dirname(HERE), 'include', 'dlv3p_train.h')]

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-std=c++17', '-lineinfo',
              '-shared', '-Xcompiler', '-fPIC', '-cudart', 'static']


def find_nvcc() -> str:
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found')


def up_to_date() -> bool:
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(d) <= t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return OUT
    cmd = [find_nvcc()] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-o', OUT, SRC]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError('nvcc failed building libdlv3p.so')
    if verbose:
        sys.stderr.write(res.stderr)
    return OUT


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
