"""The C-ABI library builds, loads and exports every symbol include/bgw.h declares (no GPU needed)."""
import ctypes as C
import os
import re

import numpy as np

from abmarl_b200 import _capi as K, philox

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    from abmarl_b200.csrc.build import build
    build()
    return K.load()


def test_every_declared_symbol_is_exported():
    header = open(os.path.join(ROOT, 'include', 'bgw.h')).read()
    declared = set(re.findall(r'\b(bgw_[a-z_]+)\s*\(', header))
    assert declared == set(K.EXPORTS), declared ^ set(K.EXPORTS)
    lib = _lib()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.bgw_abi_version() == K.BGW_ABI_VERSION


def test_struct_layouts_match_the_header():
    """ctypes mirrors of BgwSpec / BgwState / BgwDims have the C sizes (checked against the oracle build,
    which includes the same header)."""
    from oracle.oracle import lib
    assert C.sizeof(K.BgwSpec) == lib().bgwo_sizeof(0) == 26 * 4 + 3 * 8 + 8 * K.BGW_RW_COUNT + 17 * 8
    assert C.sizeof(K.BgwState) == lib().bgwo_sizeof(1) == 13 * 8
    assert C.sizeof(K.BgwDims) == lib().bgwo_sizeof(2) == 14 * 4


def test_host_rng_draw_matches_python_philox():
    lib = _lib()
    out = (C.c_uint32 * 4)()
    rng = np.random.default_rng(1)
    for _ in range(50):
        key = (int(rng.integers(0, 2**63)), int(rng.integers(0, 2**32)), int(rng.integers(0, 2**32)),
               int(rng.integers(0, 2**32)), int(rng.integers(0, 8)), int(rng.integers(0, 4096)), int(rng.integers(0, 65536)))
        assert lib.bgw_rng_draw(*key, C.byref(out)) == 0
        assert tuple(out) == philox.draw4(*key)


def test_create_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    from tests import scenarios
    from abmarl_b200.spec import compile_sim
    lib = _lib()
    spec = compile_sim(scenarios.build_tb_c2(scenarios.mirror_api()))
    h = C.c_void_p()
    rc = lib.bgw_create(C.byref(spec.c_struct()), 0, C.byref(h))
    assert rc != 0 and b'no CUDA device' in lib.bgw_last_error()
    import pytest
    from abmarl_b200.engine import BatchedGridWorld
    with pytest.raises(RuntimeError):
        BatchedGridWorld(spec)
