"""Build libproteus_b200.so in-tree with nvcc for sm_100a.

    python proteus_b200/build.py [--force] [--verbose]

(run it as a script: `python -m proteus_b200.build` imports the package first, and the package
refuses to import when the library is missing or older than the header)

The shared library is the only compiled artefact of the package.  It is built
next to this file so that it travels with the source tree (it is git-ignored,
not gpurun-ignored).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libproteus_b200.so')
SOURCES = [os.path.join(CSRC, 'pb200_api.cu')]
HEADERS = [os.path.join(CSRC, 'pb200_kernels.cuh'),
           os.path.join(CSRC, 'pb200_fused.cuh'),
           os.path.join(CSRC, 'pb200_fused_row.inc'),
           os.path.join(CSRC, 'pb200_stream.cuh'),
           os.path.join(CSRC, 'pb200_sweep.cuh'),
           os.path.join(CSRC, 'pb200_hillshade.cuh'),
           os.path.join(CSRC, 'pb200_cover.cuh'),
           os.path.join(CSRC, 'pb200_landcover.cuh'),
           os.path.join(CSRC, 'pb200_device.cuh'),
           os.path.join(CSRC, 'pb200_comm.cuh'),
           os.path.join(HERE, '..', 'include', 'proteus_b200.h')]

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-O3', '-lineinfo', '-std=c++17',
    '-fmad=false',            # float32 shadow chain must not contract (D:4255-4265)
    '-prec-div=true', '-prec-sqrt=true', '-ftz=false',
    '-shared', '-Xcompiler', '-fPIC,-O2,-Wall',
    '-cudart', 'static',
]


def find_nvcc():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'),
                 '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError('nvcc not found (set NVCC=/path/to/nvcc)')


def is_stale():
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(f) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False, extra_flags=()):
    """Compile the library if it is missing or older than its sources."""
    if not force and not is_stale():
        return LIB
    cmd = [find_nvcc(), *NVCC_FLAGS, *extra_flags, '-o', LIB, *SOURCES]
    if verbose:
        cmd.insert(1, '-Xptxas')
        cmd.insert(2, '-v')
        print(' '.join(cmd), flush=True)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode:
        raise RuntimeError('nvcc failed building libproteus_b200.so')
    return LIB


if __name__ == '__main__':
    build(force='--force' in sys.argv, verbose='--verbose' in sys.argv)
    print(LIB)
