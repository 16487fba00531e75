"""In-tree build of libd2d_b200.so (nvcc, sm_100a only).  Cross-compiles without a GPU."""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / 'csrc'
LIB = PKG / 'libd2d_b200.so'
SOURCES = ['d2d_abi.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC,-fvisibility=hidden', '-shared']


def _nvcc() -> str:
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError('nvcc not found: libd2d_b200.so can only be built with the CUDA toolkit')


def needs_build() -> bool:
    if not LIB.exists():
        return True
    newest = max(p.stat().st_mtime for p in list(CSRC.glob('*.cu*')) + [PKG.parent / 'include' / 'd2d_b200.h'])
    return LIB.stat().st_mtime < newest


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every CUDA source of the package into gym_d2d_b200/libd2d_b200.so."""
    if not force and not needs_build():
        return LIB
    cmd = [_nvcc(), *NVCC_FLAGS, '-o', str(LIB), *[str(CSRC / s) for s in SOURCES]]
    if verbose:
        cmd.insert(1, '-Xptxas=-v')
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f'nvcc failed ({res.returncode}):\n{res.stdout}\n{res.stderr}')
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == '__main__':
    print(build(force=True, verbose=True))
