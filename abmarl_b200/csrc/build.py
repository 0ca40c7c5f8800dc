"""Build abmarl_b200/csrc/libbgw.so (the C-ABI of include/bgw.h) with nvcc for sm_100a.

    python -m abmarl_b200.csrc.build [--force] [--verbose]

In-tree build: the .so sits next to the sources so it travels with the repository snapshot.  nvcc
cross-compiles without a GPU.  -ffp-contract=off on the host side: the line-of-sight rays of
bgw_los_mask are (a/b)*t in IEEE float64 (utils.py:45-115); the device side uses explicit _rn intrinsics.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SRC = os.path.join(HERE, 'bgw.cu')
OUT = os.path.join(HERE, 'libbgw.so')
DEPS = [SRC, os.path.join(HERE, 'bgw_dev.cuh'), os.path.join(HERE, 'bgw_fast.cuh'), os.path.join(ROOT, 'include', 'bgw.h'),
        os.path.join(ROOT, 'include', 'bgw_philox.h'), os.path.abspath(__file__)]
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')


def build(force=False, verbose=False):
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in DEPS):
        return OUT
    cmd = [NVCC, '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
           '-shared', '-Xcompiler', '-fPIC,-ffp-contract=off,-fno-fast-math', '--fmad=false',
           '-I', os.path.join(ROOT, 'include'), '-o', OUT, SRC, '-lcudart']
    if os.environ.get('BGW_PROFILE'):
        cmd.insert(1, '-DBGW_PROFILE')
    for d in os.environ.get('BGW_DEFINES', '').split():          # kernel variants for A/B measurements
        cmd.insert(1, '-D' + d)
    if os.environ.get('BGW_OUT'):
        cmd[cmd.index('-o') + 1] = os.environ['BGW_OUT']
    if verbose:
        cmd.insert(1, '-Xptxas=-v')
        print(' '.join(cmd))
    subprocess.check_call(cmd)
    return OUT


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
