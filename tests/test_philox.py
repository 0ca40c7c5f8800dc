"""Keyed Philox4x32-10: Python == C oracle == the published Random123 known answers."""
import ctypes as C

import numpy as np

from abmarl_b200 import philox
from oracle.oracle import lib


def _raw(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    out = (C.c_uint32 * 4)()
    lib().bgwo_philox_raw(c, k, out)
    return tuple(out)


def test_random123_known_answers():
    # Random123 kat_vectors: philox4x32 10 rounds
    kats = [
        ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
         (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    for ctr, key, want in kats:
        assert philox.philox4x32_10(ctr, key) == want
        assert _raw(ctr, key) == want


def test_python_and_c_streams_agree():
    rng = np.random.default_rng(0)
    out = (C.c_uint32 * 4)()
    for _ in range(200):
        seed = int(rng.integers(0, 2**63)) * 2 + int(rng.integers(0, 2))
        env, ep, step = (int(rng.integers(0, 2**32)) for _ in range(3))
        site, slot, k = int(rng.integers(0, 8)), int(rng.integers(0, 4096)), int(rng.integers(0, 65536))
        lib().bgwo_rng_draw(C.c_uint64(seed), env, ep, step, site, slot, k, out)
        assert tuple(out) == philox.draw4(seed, env, ep, step, site, slot, k)


def test_mappings():
    assert philox.u01(0) == 0.0 and philox.u01(2**32 - 1) < 1.0
    for n in (1, 2, 3, 7, 121, 4096):
        for x in (0, 1, 2**31, 2**32 - 1, 123456789):
            assert philox.index(x, n) == int(philox.u01(x) * n)     # floor(u*n): exact, u*n < 2^53
