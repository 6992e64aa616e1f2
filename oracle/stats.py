"""ORACLE (test infrastructure, not product code) — numpy restatement of the streaming statistics (docs/MODEL.md §9).

PARITY UNPINNED for the trial itself (see oracle/trial.py).  The *reductions* here are pinned by present reference code:
half-normal sigma/mean /root/reference/driver.py:290-291,323-324; failure rate driver.py:321-322; RMS toolplot.py:82;
6-sigma clip + 50-bin histogram + mode + sigma-about-mode toolplot.py:82-99.
"""
from __future__ import annotations

import numpy as np

STATS_HEAD = 8
HIST_OCTAVES = 34
HIST_MIN_EXP = -14


def n_bins(log2_sub: int = 8) -> int:
    return HIST_OCTAVES << log2_sub


def bin_index(x, log2_sub: int = 8):
    """Log-linear bin of x > 0 [arcsec]: exponent and top `log2_sub` mantissa bits of the float64 (exact, integer-only)."""
    bits = np.asarray(x, dtype=np.float64).view(np.int64)
    return (bits >> (52 - log2_sub)) - ((1023 + HIST_MIN_EXP) << log2_sub)


def bin_edges(log2_sub: int = 8):
    sub = 1 << log2_sub
    oct_base = np.ldexp(1.0, HIST_MIN_EXP + np.arange(HIST_OCTAVES))
    edges = (oct_base[:, None] * (1.0 + np.arange(sub)[None, :] / sub)).reshape(-1)
    return np.append(edges, np.ldexp(1.0, HIST_MIN_EXP + HIST_OCTAVES))


def packed_stats(err_arcsec, n_det, log2_sub: int = 8):
    """[n_ok, n_fail, sum x, sum x^2, sum n_det, x_max, under, over, hist...] exactly as the CUDA accumulators define it."""
    err = np.asarray(err_arcsec, dtype=np.float64)
    ok = err >= 0.0
    x = err[ok]
    nb = n_bins(log2_sub)
    out = np.zeros(STATS_HEAD + nb)
    out[0], out[1] = ok.sum(), (~ok).sum()
    out[2], out[3] = x.sum(), (x * x).sum()
    out[4] = np.asarray(n_det, dtype=np.int64).sum()
    out[5] = x.max() if x.size else 0.0
    b = bin_index(x, log2_sub)
    out[6], out[7] = (b < 0).sum(), (b >= nb).sum()
    inside = (b >= 0) & (b < nb)
    out[STATS_HEAD:] = np.bincount(b[inside], minlength=nb)
    return out


def percentile(stats, q, log2_sub: int = 8):
    n_ok, under = stats[0], stats[6]
    hist = stats[STATS_HEAD:]
    target = q * n_ok
    cum = under
    if target <= cum:
        return np.ldexp(1.0, HIST_MIN_EXP)
    edges = bin_edges(log2_sub)
    for b in np.nonzero(hist)[0]:
        c = hist[b]
        before = under + hist[:b].sum()
        if before + c >= target:
            return edges[b] + (edges[b + 1] - edges[b]) * (target - before) / c
    return edges[-1]


def summarize(stats, log2_sub: int = 8):
    n, nf = stats[0], stats[1]
    out = {"n_ok": n, "n_fail": nf, "fail_rate": nf / (n + nf) if n + nf else 0.0,   # driver.py:321-322
           "mean_ndet": stats[4] / (n + nf) if n + nf else 0.0, "x_max": stats[5]}
    if n > 0:
        out["mean"] = stats[2] / n
        out["rms"] = np.sqrt(stats[3] / n)                                           # toolplot.py:82
        var = max(0.0, (stats[3] - stats[2] ** 2 / n) / (n - 1)) if n > 1 else 0.0   # pandas Series.std(): ddof = 1
        out["std"] = np.sqrt(var)
        out["sigma_halfnormal"] = out["std"] / np.sqrt(1 - 2 / np.pi)                # driver.py:290,324
        out["mean_halfnormal"] = out["sigma_halfnormal"] * np.sqrt(2 / np.pi)        # driver.py:291,323
        for name, q in (("p50", .5), ("p90", .9), ("p99", .99), ("p999", .999)):
            out[name] = percentile(stats, q, log2_sub)
    return out


def toolplot_reductions(x, bins: int = 50):
    """toolplot.py:82-99 on an error sample: RMS, 6-sigma clip, density histogram, mode (left edge of the fullest bin),
    sigma about the mode."""
    x = np.asarray(x, dtype=np.float64)
    x = x[x >= 0.0]                                         # failed trials are rows the reference dropped (driver.py:321)
    rms = np.sqrt((x ** 2).sum() / x.size)                  # toolplot.py:82
    kept = x[x <= 6 * rms]                                  # toolplot.py:83
    h = np.histogram(kept, bins=bins, density=True)         # toolplot.py:89
    hmax = int(np.argmax(h[0]))                             # toolplot.py:90
    mode = h[1][hmax]                                       # toolplot.py:96
    newsig = np.sqrt(((kept - mode) ** 2).sum() / kept.size)  # toolplot.py:99
    return {"n": x.size, "rms": rms, "n_kept": kept.size, "kept_min": kept.min(), "kept_max": kept.max(), "mode": mode,
            "mode_bin": hmax, "sigma_about_mode": newsig, "density": h[0]}
