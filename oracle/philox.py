"""ORACLE (test infrastructure, not product code) — Philox4x32-10 in numpy.

PARITY UNPINNED: the reference ships no test/golden vector for this path and never seeds its RNG
(global legacy np.random, /root/reference/monte_carlo.py:90-92).  This generator is the frozen spec of docs/MODEL.md §5;
it is pinned instead against the published Random123 known-answer vectors (tests/test_oracle_philox.py).

Algorithm: Salmon, Moraes, Dror, Shaw, "Parallel random numbers: as easy as 1, 2, 3", SC'11 — 4 words of 32 bits,
10 rounds, multipliers 0xD2511F53 / 0xCD9E8D57, Weyl key increments 0x9E3779B9 / 0xBB67AE85.  (numpy's own
`np.random.Philox` is the 4x64 variant and does not match.)
"""
from __future__ import annotations

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)
SHIFT32 = np.uint64(32)

TAG_STAR = 0x53544152  # "STAR"  (MODEL.md §5)
TAG_PARA = 0x50415241  # "PARA"  (MODEL.md §6)


def philox4x32_10(c0, c1, c2, c3, k0: int, k1: int):
    """Vectorised Philox4x32-10. c0..c3: broadcastable uint32-valued arrays; k0,k1: python ints. Returns 4 uint32 arrays."""
    c0, c1, c2, c3 = np.broadcast_arrays(*(np.asarray(c, dtype=np.uint64) & MASK32 for c in (c0, c1, c2, c3)))
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0            # < 2^64, exact in uint64
        p1 = M1 * c2
        hi0, lo0 = p0 >> SHIFT32, p0 & MASK32
        hi1, lo1 = p1 >> SHIFT32, p1 & MASK32
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


M2 = np.uint64(0xD256D193)   # Philox2x32 multiplier (Random123)


def philox2x32_10(c0, c1, k0):
    """Vectorised Philox2x32-10: counter (c0, c1), ONE 32-bit key word (array or int). Returns 2 uint32 arrays."""
    c0, c1, k = np.broadcast_arrays(*(np.asarray(c, dtype=np.uint64) & MASK32 for c in (c0, c1, k0)))
    for _ in range(10):
        p = M2 * c0
        c0, c1 = (p >> SHIFT32) ^ k ^ c1, p & MASK32
        k = (k + np.uint64(W0)) & MASK32
    return c0.astype(np.uint32), c1.astype(np.uint32)


def star_words(seed: int, trial_idx, star_id):
    """MODEL.md §5: the two 32-bit words behind one star's centroid noise.
    Per trial a key pair (ka, kb) = first two words of Philox4x32-10(ctr = (trial_lo, trial_hi, 0, "STAR"), key = seed);
    per star (x0, x1) = Philox2x32-10(ctr = (star_id, kb), key = ka)."""
    k0, k1 = split_seed(seed)
    tl, th = split_index(trial_idx)
    ka, kb, _, _ = philox4x32_10(tl, th, 0, TAG_STAR, k0, k1)
    return philox2x32_10(np.asarray(star_id, dtype=np.uint64), kb, ka)


def u01(w):
    """(w + 0.5) * 2^-32 in float64: strictly inside (0, 1), exact."""
    return (np.asarray(w, dtype=np.float64) + 0.5) * (1.0 / 4294967296.0)


def box_muller(wa, wb):
    """MODEL.md §5: R = sqrt(-2 ln u1); z1 = R cos(2 pi u2); z2 = R sin(2 pi u2)."""
    u1, u2 = u01(wa), u01(wb)
    r = np.sqrt(-2.0 * np.log(u1))
    ang = 2.0 * np.pi * u2
    return r * np.cos(ang), r * np.sin(ang)


def split_seed(seed: int):
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return seed & 0xFFFFFFFF, seed >> 32


def split_index(idx):
    idx = np.asarray(idx, dtype=np.uint64)
    return idx & MASK32, idx >> SHIFT32
