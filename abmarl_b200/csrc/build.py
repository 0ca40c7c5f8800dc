"""Build abmarl_b200/csrc/libbgw.so (the C-ABI of include/bgw.h) with nvcc for sm_100a.

    python -m abmarl_b200.csrc.build [--force] [--verbose]

In-tree build: the .so sits next to the sources so it travels with the repository snapshot.  nvcc
cross-compiles without a GPU.  Three translation units compiled in parallel (bgw.cu: host side and small kernels,
bgw_general.cu: the general step kernel's instantiations, bgw_fastk.cu: the specialised kernel's); an object is
rebuilt when its source, any header of this directory / include/, this script or the variant flags changed.
-ffp-contract=off on the host side: the line-of-sight rays of bgw_los_mask are (a/b)*t in IEEE float64
(utils.py:45-115); the device side uses explicit _rn intrinsics.
"""
import glob
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, 'libbgw.so')
UNITS = ('bgw.cu', 'bgw_general.cu', 'bgw_fastk.cu')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')


# headers a unit includes (bgw_general.cu does not see the specialised kernel: editing bgw_fast.cuh leaves it alone)
UNIT_HEADERS = {'bgw.cu': ('bgw_dev.cuh', 'bgw_fast.cuh', 'bgw_maze.cuh', 'bgw_jit.h'), 'bgw_general.cu': ('bgw_dev.cuh',),
                'bgw_fastk.cu': ('bgw_dev.cuh', 'bgw_fast.cuh')}


def _headers(unit):
    return [os.path.join(HERE, h) for h in UNIT_HEADERS[unit]] + sorted(glob.glob(os.path.join(ROOT, 'include', '*.h'))) + [os.path.abspath(__file__)]


def build(force=False, verbose=False):
    out = os.environ.get('BGW_OUT') or OUT
    flags = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
             '-Xcompiler', '-fPIC,-ffp-contract=off,-fno-fast-math', '--fmad=false', '-I', os.path.join(ROOT, 'include')]
    if os.environ.get('BGW_PROFILE'):
        flags.append('-DBGW_PROFILE')
    flags += ['-D' + d for d in os.environ.get('BGW_DEFINES', '').split()]          # kernel variants for A/B measurements
    flags += os.environ.get('BGW_NVCC_FLAGS', '').split()
    if verbose:
        flags.append('-Xptxas=-v')
    # BGW_FAST_DEFINES: defines for the specialised kernel's unit only (e.g. BGW_JITTER): the other two units are shared with
    # the plain build instead of being recompiled
    unit_flags = {u: list(flags) for u in UNITS}
    unit_flags['bgw_fastk.cu'] += ['-D' + d for d in os.environ.get('BGW_FAST_DEFINES', '').split()]
    jobs, objs = [], []
    for u in UNITS:
        tag = hashlib.sha1(' '.join(unit_flags[u]).encode()).hexdigest()[:10]       # objects of different variants do not mix
        objdir = os.path.join(HERE, '_obj', tag)
        os.makedirs(objdir, exist_ok=True)
        src, obj = os.path.join(HERE, u), os.path.join(objdir, u[:-3] + '.o')
        objs.append(obj)
        newest = max([os.path.getmtime(src)] + [os.path.getmtime(h) for h in _headers(u)])
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < newest:
            jobs.append([NVCC] + unit_flags[u] + ['-c', src, '-o', obj])
    if not jobs and os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(o) for o in objs):
        return out

    def run(cmd):
        if verbose:
            print(' '.join(cmd), flush=True)
        subprocess.check_call(cmd)

    if jobs:
        with ThreadPoolExecutor(len(jobs)) as pool:
            list(pool.map(run, jobs))
    run([NVCC, '-gencode', 'arch=compute_100a,code=sm_100a', '-shared', '-o', out] + objs + ['-lcudart', '-ldl'])
    return out


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
