"""Streaming-statistics oracle and the host-side C finalizer (stmpm_stats_finalize needs no GPU)."""
import numpy as np
import pytest

from oracle import stats as ostats


def sample(n=50_000, seed=0):
    rng = np.random.default_rng(seed)
    x = np.abs(rng.normal(0, 120.0, n)) + rng.uniform(0, 1e-3, n)
    x[rng.uniform(size=n) < 0.01] = -1.0
    return x, rng.integers(0, 40, n).astype(np.int32)


def test_bin_index_matches_edges():
    for k in (4, 8):
        edges = ostats.bin_edges(k)
        assert len(edges) == ostats.n_bins(k) + 1 and np.all(np.diff(edges) > 0)
        x = np.exp(np.random.default_rng(1).uniform(np.log(edges[0]), np.log(edges[-1] * 0.999), 20_000))
        assert np.array_equal(ostats.bin_index(x, k), np.searchsorted(edges, x, side="right") - 1)
        assert ostats.bin_index(edges[:-1], k).tolist() == list(range(ostats.n_bins(k)))


def test_packed_stats_and_percentiles():
    x, nd = sample()
    s = ostats.packed_stats(x, nd)
    ok = x >= 0
    assert s[0] == ok.sum() and s[1] == (~ok).sum() and s[4] == nd.sum()
    assert s[ostats.STATS_HEAD:].sum() + s[6] + s[7] == s[0]
    r = ostats.summarize(s)
    assert r["std"] == pytest.approx(x[ok].std(ddof=1), rel=1e-12)
    assert r["sigma_halfnormal"] == pytest.approx(x[ok].std(ddof=1) / np.sqrt(1 - 2 / np.pi), rel=1e-12)  # driver.py:324
    for name, q in (("p50", 50), ("p99", 99), ("p999", 99.9)):
        assert r[name] == pytest.approx(np.percentile(x[ok], q), rel=5e-3)   # bins are <= 0.39 % wide


def test_c_finalizer_matches_oracle(lib_built):
    from startrackermpm_b200 import engine
    for k in (5, 8):
        x, nd = sample(seed=k)
        s = ostats.packed_stats(x, nd, k)
        assert engine.stats_len(k) == s.size
        got, want = engine.summarize(s, k), ostats.summarize(s, k)
        for key, v in want.items():
            assert got[key] == pytest.approx(v, rel=1e-12, abs=1e-300), key


def test_toolplot_reductions_on_half_normal():
    x = np.abs(np.random.default_rng(3).normal(0, 50.0, 200_000))
    r = ostats.toolplot_reductions(x)
    assert r["rms"] == pytest.approx(50.0, rel=0.01)      # toolplot.py:82
    assert r["mode"] < 2.0                                 # half-normal peaks at 0 -> first bin's left edge
    assert r["sigma_about_mode"] == pytest.approx(50.0, rel=0.03)
