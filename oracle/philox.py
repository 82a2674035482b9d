"""TEST INFRASTRUCTURE ONLY -- numpy Philox4x32-10 and the draw-stream contract.

This restates, on the CPU, the counter-based stream that the CUDA kernels in
``mjhmc_b200/csrc`` use in PHILOX mode (DESIGN.md "Random streams"), so a GPU
run in PHILOX mode can be compared draw-for-draw with the oracle.

The reference has no counterpart (it uses the global ``np.random`` Mersenne
Twister: hmc_state.py:26,126; markov_jump_hmc.py:125,132,138,326; utils.py:42);
Philox4x32-10 itself is the published Random123 algorithm (Salmon et al., SC'11).

Stream contract (key = 64-bit seed; 128-bit counter):
    c0,c1 = global particle index (lo, hi 32 bits)
    c2    = attempt index (sampling iterations, counting failed back-off attempts)
    c3    = slot:  0 -> words (w0,w1)->u[0], (w2,w3)->u[1]
                   1 -> words (w0,w1)->u[2]
                   2+j -> Box-Muller pair j -> normals 2j, 2j+1
    batch-wide coin of the discrete samplers: particle = 2**64-1, slot 0, u[0]
    uniform = ((wa >> 5) * 2**26 + (wb >> 6)) / 2**53          (same as numpy's legacy double)
    normal pair: r = sqrt(-2 log(1 - u1)); z0 = r cos(2 pi u2); z1 = r sin(2 pi u2)
"""
import numpy as np

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)
_S32 = np.uint64(32)

COIN_PARTICLE = (1 << 64) - 1


def philox4x32_10(c0, c1, c2, c3, seed):
    """Vectorised Philox4x32-10.  All counters broadcastable integer arrays."""
    c0, c1, c2, c3 = np.broadcast_arrays(*(np.asarray(c, dtype=np.uint64) for c in (c0, c1, c2, c3)))
    c0, c1, c2, c3 = (c & _MASK for c in (c0, c1, c2, c3))
    k0 = int(seed) & 0xFFFFFFFF
    k1 = (int(seed) >> 32) & 0xFFFFFFFF
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> _S32, p0 & _MASK
        hi1, lo1 = p1 >> _S32, p1 & _MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)), lo1, (hi0 ^ c3 ^ np.uint64(k1)), lo0
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def _u53(wa, wb):
    return ((wa >> np.uint64(5)).astype(np.float64) * 67108864.0
            + (wb >> np.uint64(6)).astype(np.float64)) / 9007199254740992.0


def _split(particles):
    p = np.asarray(particles, dtype=np.uint64)
    return p & _MASK, p >> _S32


def uniforms(seed, attempt, particles):
    """u[0..2] for every particle -> array (3, n)."""
    lo, hi = _split(particles)
    w = philox4x32_10(lo, hi, attempt, 0, seed)
    w1 = philox4x32_10(lo, hi, attempt, 1, seed)
    return np.stack([_u53(w[0], w[1]), _u53(w[2], w[3]), _u53(w1[0], w1[1])])


def coin(seed, attempt):
    w = philox4x32_10(COIN_PARTICLE & 0xFFFFFFFF, COIN_PARTICLE >> 32, attempt, 0, seed)
    return float(_u53(w[0], w[1]))


def normals(seed, attempt, particles, ndims):
    """Standard normals (ndims, n) for every particle."""
    lo, hi = _split(particles)
    out = np.empty((ndims, lo.shape[0]), dtype=np.float64)
    for j in range((ndims + 1) // 2):
        w = philox4x32_10(lo, hi, attempt, 2 + j, seed)
        u1 = _u53(w[0], w[1])
        u2 = _u53(w[2], w[3])
        r = np.sqrt(-2.0 * np.log(1.0 - u1))
        out[2 * j] = r * np.cos(2.0 * np.pi * u2)
        if 2 * j + 1 < ndims:
            out[2 * j + 1] = r * np.sin(2.0 * np.pi * u2)
    return out
