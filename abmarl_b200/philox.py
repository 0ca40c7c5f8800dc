"""Pure-Python Philox4x32-10 with the engine's key layout (include/bgw_philox.h).

Host-side users: layout generators (abmarl_b200.layouts) and the RNG replay shim of the test harness.
Bit-identical to the device stream (tests/test_philox.py checks Python == C oracle == Random123 KAT,
and the gpu tests check == bgw_rng_draw).
"""
M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF


def philox4x32_10(ctr, key):
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & MASK, p1 & MASK, ((p0 >> 32) ^ c3 ^ k1) & MASK, p0 & MASK
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return c0, c1, c2, c3


def draw4(seed, env, episode, step, site, slot, k):
    ctr = (env & MASK, episode & MASK, step & MASK, ((site << 28) | ((slot & 0xFFF) << 16) | (k & 0xFFFF)) & MASK)
    return philox4x32_10(ctr, (seed & MASK, (seed >> 32) & MASK))


def draw(seed, env, episode, step, site, slot, k=0):
    return draw4(seed, env, episode, step, site, slot, k)[0]


def u01(x):
    """uniform in [0,1) as float64 (exact)."""
    return x * (1.0 / 4294967296.0)


def index(x, n):
    """floor(u01(x) * n) in exact integer arithmetic."""
    return (x * n) >> 32
