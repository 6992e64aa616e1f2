"""Philox4x32-10 oracle against the published Random123 known-answer vectors (kat_vectors of the Random123 distribution;
the same generator as cuRAND's curand_philox4x32_x.h).  This is the one part of the path with external golden vectors."""
import numpy as np

from oracle import philox

KAT = [
    ((0x00000000, 0x00000000, 0x00000000, 0x00000000), (0x00000000, 0x00000000),
     (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff),
     (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_known_answer_vectors():
    for ctr, key, want in KAT:
        got = philox.philox4x32_10(*ctr, *key)
        assert tuple(int(g) for g in got) == want


KAT2 = [((0x00000000, 0x00000000), 0x00000000, (0xff1dae59, 0x6cd10df2)),
        ((0xffffffff, 0xffffffff), 0xffffffff, (0x2c3f628b, 0xab4fd7ad)),
        ((0x243f6a88, 0x85a308d3), 0x13198a2e, (0xdd7ce038, 0xf62a4c12))]


def test_philox2x32_known_answer_vectors():
    for ctr, key, want in KAT2:
        got = philox.philox2x32_10(*ctr, key)
        assert tuple(int(g) for g in got) == want


def test_star_words_composition():
    tr = np.array([0, 5, (7 << 32) | 9], dtype=np.uint64)
    x0, x1 = philox.star_words((3 << 32) | 11, tr, np.array([1, 2, 3]))
    for i in range(3):
        ka, kb, _, _ = philox.philox4x32_10(int(tr[i]) & 0xffffffff, int(tr[i]) >> 32, 0, philox.TAG_STAR, 11, 3)
        a, b = philox.philox2x32_10(i + 1, int(kb), int(ka))
        assert (int(x0[i]), int(x1[i])) == (int(a), int(b))


def test_vectorised_matches_scalar():
    rng = np.random.default_rng(0)
    c = rng.integers(0, 2**32, size=(4, 64), dtype=np.uint64)
    got = philox.philox4x32_10(c[0], c[1], c[2], c[3], 123, 456)
    for i in range(64):
        one = philox.philox4x32_10(int(c[0, i]), int(c[1, i]), int(c[2, i]), int(c[3, i]), 123, 456)
        assert [int(g[i]) for g in got] == [int(o) for o in one]


def test_u01_open_interval_and_box_muller_moments():
    assert 0.0 < philox.u01(0) < philox.u01(2**32 - 1) < 1.0
    idx = np.arange(200_000, dtype=np.uint64)
    x0, x1 = philox.star_words(11 | (22 << 32), idx, 7)
    z1, z2 = philox.box_muller(x0, x1)
    z = np.concatenate([z1, z2])
    assert abs(z.mean()) < 4 / np.sqrt(z.size)
    assert abs(z.std() - 1) < 4 / np.sqrt(2 * z.size)
    assert abs(np.corrcoef(z1, z2)[0, 1]) < 4 / np.sqrt(z1.size)
    assert np.abs(z).max() <= np.sqrt(-2 * np.log(0.5 * 2.0**-32)) + 1e-12   # bound used by the CUDA pre-filter margin


def test_seed_and_index_split():
    assert philox.split_seed((5 << 32) | 9) == (9, 5)
    lo, hi = philox.split_index(np.array([(3 << 32) | 4], dtype=np.uint64))
    assert int(lo[0]) == 4 and int(hi[0]) == 3
