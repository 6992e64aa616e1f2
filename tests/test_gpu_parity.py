"""GPU parity tests proper (-m gpu): the CUDA path, called through the C ABI (ctypes), against the numpy oracle on the
same seeded inputs, against the committed golden fixture, and — at BASELINE sizes — through size-independent properties.

Bars (BASELINE.json north_star): per-trial |d error| <= 1e-9 rad in fp64 mode, <= 1e-6 rad in fp32 mode; star-detection
counts and failure flags bit-exact.  PARITY UNPINNED: the oracle is this repo's restatement (the reference's trial code
is absent), so these tests prove CUDA == oracle == docs/MODEL.md, not CUDA == reference."""
import numpy as np
import pytest
import torch

from conftest import BASIC_MEAN, BASIC_SIGMA, THERMAL_COEF, ZERO_SIGMA
from oracle import stats as ostats
from oracle import trial
from startrackermpm_b200.engine import TrialEngine, make_dist

pytestmark = pytest.mark.gpu
ARCSEC = trial.RAD2ARCSEC
TOL_RAD = {64: 1e-9, 32: 1e-6}
F = 3500.0


def compare(e_gpu, d_gpu, e_ref, d_ref, prec):
    assert np.array_equal(d_gpu, d_ref), f"{(d_gpu != d_ref).sum()} detection-count mismatches"
    assert np.array_equal(e_gpu < 0, e_ref < 0)
    ok = e_ref >= 0
    if ok.any():
        worst = np.abs(e_gpu[ok] - e_ref[ok]).max() / ARCSEC
        assert worst <= TOL_RAD[prec], f"max |d err| = {worst:.3e} rad > {TOL_RAD[prec]:.0e}"
    assert (e_gpu[~ok] == -1.0).all()


@pytest.mark.parametrize("prec", [64, 32])
def test_presampled_matches_oracle_default_config(engine, catalog, prec):
    xyz, mag = catalog
    n = 30_000
    cols = trial.sample_params(BASIC_MEAN, BASIC_SIGMA, THERMAL_COEF, 1234, 0, n)
    e_ref, d_ref = trial.run_trials(cols, xyz, mag, F, 1234, 0)
    e, d = engine.run_presampled(cols, 1234, 0, precision=prec)
    compare(e, d, e_ref, d_ref, prec)


@pytest.mark.parametrize("prec", [64, 32])
def test_golden_fixture(engine, golden, prec):
    e, d = engine.run_presampled(golden["cols"], int(golden["seed"]), int(golden["trial0"]), precision=prec)
    compare(e, d, golden["err_arcsec"], golden["n_det"], prec)
    for name in np.unique(golden["scenario"]):
        m = golden["scenario"] == name
        assert np.array_equal(d[m], golden["n_det"][m]), name


@pytest.mark.parametrize("prec", [64, 32])
def test_independent_golden_v2(engine, golden_v2, prec):
    """CUDA against the fixture of the independent scalar restatement (tests/golden/independent_trial.py: scipy
    align_vectors, libm Box-Muller, own Philox — no code shared with oracle/ or csrc/), 64-bit seed, trial indices
    across 2^32."""
    g = golden_v2
    e, d = engine.run_presampled(g["cols"], int(g["seed"]), int(g["trial0"]), precision=prec)
    compare(e, d, g["err_arcsec"], g["n_det"], prec)
    e2, d2, m2 = engine.run_presampled(g["cols"], int(g["seed"]), int(g["trial0"]), precision=prec, want_maxmag=True)
    assert np.array_equal(d2, g["n_det"]) and np.array_equal(m2, g["faintest_mag"])     # star_mag objective, bit-exact


@pytest.mark.parametrize("prec", [64, 32])
def test_stress_configs(engine, catalog, prec):
    """Sweep-like extremes: 4x sigmas, wide FOV (f = 1000 px), everything visible (MAX_MAGNITUDE = 20, sensitivity.py:112),
    nothing visible, poles / RA wrap, negative BASE_DEV (sensitivity.py:46-47)."""
    xyz, mag = catalog
    n = 1500
    big = {k: 4 * v for k, v in BASIC_SIGMA.items()}
    cases = {
        "big_sigma": trial.sample_params(BASIC_MEAN, big, THERMAL_COEF, 5, 0, n),
        "wide": trial.sample_params(dict(BASIC_MEAN, FOCAL_LENGTH=1000.0), BASIC_SIGMA, 0.0, 5, 0, 300),
        "all_visible": trial.sample_params(dict(BASIC_MEAN, MAX_MAGNITUDE=20.0), BASIC_SIGMA, 0.0, 5, 0, 600),
        "none_visible": trial.sample_params(dict(BASIC_MEAN, MAX_MAGNITUDE=-3.0), ZERO_SIGMA, 0.0, 5, 0, 600),
        "few": trial.sample_params(dict(BASIC_MEAN, MAX_MAGNITUDE=4.2), BASIC_SIGMA, 0.0, 5, 0, n),
        "neg_dev": trial.sample_params(dict(BASIC_MEAN, BASE_DEV_X=-0.3, BASE_DEV_Y=-0.2), ZERO_SIGMA, 0.0, 5, 0, 600),
        # tilts beyond the small-angle series (> 0.1 rad) and beyond the candidate lists' cone -> sky-grid band scan
        "big_tilt": trial.sample_params(dict(BASIC_MEAN, F_ARR_PHI=7.0, F_ARR_THETA=-6.5, F_ARR_PSI=30.0), BASIC_SIGMA, 0.0, 5, 0, 400),
        "inf_mag": trial.sample_params(dict(BASIC_MEAN, MAX_MAGNITUDE=np.inf), ZERO_SIGMA, 0.0, 5, 0, 200),
        "big_shift": trial.sample_params(dict(BASIC_MEAN, F_ARR_EPS_X=60.0, F_ARR_EPS_Y=-45.0), BASIC_SIGMA, 0.0, 5, 0, 400),
    }
    polar = trial.sample_params(BASIC_MEAN, BASIC_SIGMA, 0.0, 5, 0, 720)
    polar[1] = np.linspace(-np.pi / 2, np.pi / 2, 720)
    polar[0] = np.linspace(-1.0, 2 * np.pi + 1.0, 720)
    cases["polar"] = polar
    for name, cols in cases.items():
        e_ref, d_ref = trial.run_trials(cols, xyz, mag, F, 5, 0)
        e, d = engine.run_presampled(cols, 5, 0, precision=prec)
        compare(e, d, e_ref, d_ref, prec)
    assert (engine.run_presampled(cases["none_visible"], 5)[0] == -1).all()


def test_analytic_known_answers_on_gpu(engine):
    cols = trial.sample_params(BASIC_MEAN, ZERO_SIGMA, 0.0, 3, 0, 4096)
    e, d = engine.run_presampled(cols, 3)
    assert (e >= 0).all() and (e / ARCSEC).max() < 1e-12                       # ideal tracker
    cols[9] = 0.2                                                                # pure psi rotation -> exactly psi
    e, _ = engine.run_presampled(cols, 3)
    assert np.allclose(e / ARCSEC, np.deg2rad(0.2), rtol=0, atol=1e-12)
    cols[9] = 0.0; cols[4] = 2.0                                                 # lateral shift -> atan(eps / f)
    e, _ = engine.run_presampled(cols, 3)
    assert np.allclose(e, np.arctan(2.0 / F) * ARCSEC, rtol=0.02)


def test_ragged_and_empty_sizes(engine, catalog):
    xyz, mag = catalog
    cols = trial.sample_params(BASIC_MEAN, BASIC_SIGMA, THERMAL_COEF, 8, 0, 1537)
    e_ref, d_ref = trial.run_trials(cols, xyz, mag, F, 8, 0)
    for n in (0, 1, 31, 32, 33, 511, 512, 513, 1537):
        e, d = engine.run_presampled(np.ascontiguousarray(cols[:, :n]), 8, 0)
        assert e.shape == (n,)
        compare(e, d, e_ref[:n], d_ref[:n], 64)


def test_trial_index_sharding_is_exact(engine):
    """Any shard [a, b) with trial0 = a reproduces the same records as the whole run: the basis of multi-GPU sharding."""
    cols = trial.sample_params(BASIC_MEAN, BASIC_SIGMA, THERMAL_COEF, 21, 1000, 5000)
    e, d = engine.run_presampled(cols, 21, 1000)
    for a, b in ((0, 1), (17, 2300), (2300, 5000)):
        es, ds = engine.run_presampled(np.ascontiguousarray(cols[:, a:b]), 21, 1000 + a)
        assert np.array_equal(es, e[a:b]) and np.array_equal(ds, d[a:b])
    big = (1 << 40) + 12345                                                     # 64-bit trial indices
    c2 = trial.sample_params(BASIC_MEAN, BASIC_SIGMA, THERMAL_COEF, 21, big, 200)
    from startrackermpm_b200.catalog import load_catalog
    xyz, mag = load_catalog()
    e_ref, d_ref = trial.run_trials(c2, xyz, mag, F, 21, big)
    compare(*engine.run_presampled(c2, 21, big), e_ref, d_ref, 64)


@pytest.mark.parametrize("prec", [64, 32])
def test_native_rng_equals_generator_plus_presampled(engine, catalog, prec):
    """stmpm_trials_rng == stmpm_fill_presampled -> stmpm_trials_presampled, and the generator == the oracle sampler."""
    xyz, mag = catalog
    n, seed, t0 = 20_000, 99, 7_000_000
    dist = make_dist(BASIC_MEAN, BASIC_SIGMA, THERMAL_COEF)
    dcols = torch.empty((13, n), dtype=torch.float64, device="cuda")
    engine.fill_presampled_device(dcols, dist, seed=seed, trial0=t0)
    torch.cuda.synchronize()
    cols = dcols.cpu().numpy()
    ref_cols = trial.sample_params(BASIC_MEAN, BASIC_SIGMA, THERMAL_COEF, seed, t0, n)
    assert np.abs(cols - ref_cols).max() < 1e-11
    stats, e_rng, d_rng = engine.run_rng(dist, n, seed, t0, precision=prec, records=True)
    e_pre, d_pre = engine.run_presampled(cols, seed, t0, precision=prec)
    assert np.array_equal(d_rng, d_pre)
    ok = e_pre >= 0
    assert np.abs(e_rng[ok] - e_pre[ok]).max() / ARCSEC < (1e-12 if prec == 64 else 1e-6)
    e_ref, d_ref = trial.run_trials(ref_cols, xyz, mag, F, seed, t0)
    compare(e_rng, d_rng, e_ref, d_ref, prec)
    # streaming statistics == oracle statistics of the same records
    want = ostats.packed_stats(e_rng, d_rng)
    got = stats[0]
    assert np.array_equal(got[[0, 1, 4, 6, 7]], want[[0, 1, 4, 6, 7]])
    assert np.array_equal(got[8:], want[8:])                                    # histogram bit-exact
    assert got[5] == want[5]
    assert np.allclose(got[[2, 3]], want[[2, 3]], rtol=1e-12)
    s = engine.summarize(got)
    assert s["p99"] == pytest.approx(np.percentile(e_rng[ok], 99), rel=5e-3)


def test_segments_are_independent_parameter_sets(engine):
    """Sweep launcher: segment s of a multi-segment launch == a single-segment launch with dists[s] at the same indices."""
    tps, seed = 3000, 31
    dists = [make_dist(dict(BASIC_MEAN, F_ARR_PSI=0.02 * k), dict(BASIC_SIGMA, BASE_DEV_X=0.05 * k, BASE_DEV_Y=0.05 * k),
                       THERMAL_COEF) for k in range(5)]
    stats, e, d = engine.run_rng(dists, tps, seed, 500, records=True)
    only_stats = engine.run_rng(dists, tps, seed, 500)
    for k in range(5):
        sk, ek, dk = engine.run_rng(dists[k], tps, seed, 500 + k * tps, records=True)
        assert np.array_equal(ek, e[k * tps:(k + 1) * tps]) and np.array_equal(dk, d[k * tps:(k + 1) * tps])
        for st in (stats[k], only_stats[k]):
            assert np.array_equal(st[[0, 1, 4, 5]], sk[0][[0, 1, 4, 5]]) and np.array_equal(st[8:], sk[0][8:])
            assert np.allclose(st[[2, 3]], sk[0][[2, 3]], rtol=1e-12)
    means = [engine.summarize(stats[k])["mean"] for k in range(5)]
    assert means[4] > means[0]                                                   # psi bias grows along the sweep


def test_statistical_parity_with_native_rng(engine, catalog):
    """Second correctness leg of the north_star: with GPU-native RNG at a DIFFERENT seed than the oracle sample, mean,
    std and p99 agree within Monte Carlo confidence intervals."""
    xyz, mag = catalog
    n_ref = 40_000
    cols = trial.sample_params(BASIC_MEAN, BASIC_SIGMA, THERMAL_COEF, 777, 0, n_ref)
    e_ref, _ = trial.run_trials(cols, xyz, mag, F, 777, 0)
    x = e_ref[e_ref >= 0]
    n_gpu = 4_000_000
    s = engine.summarize(engine.run_rng(make_dist(BASIC_MEAN, BASIC_SIGMA, THERMAL_COEF), n_gpu, 4242)[0])
    se_mean = x.std(ddof=1) / np.sqrt(x.size)
    assert abs(s["mean"] - x.mean()) < 4.5 * se_mean
    m4 = ((x - x.mean()) ** 4).mean()
    se_std = np.sqrt((m4 - x.var() ** 2) / x.size) / (2 * x.std())
    assert abs(s["std"] - x.std(ddof=1)) < 4.5 * se_std
    lo, hi = np.percentile(x, [99 - 4.5 * 100 * np.sqrt(.99 * .01 / x.size), 99 + 4.5 * 100 * np.sqrt(.99 * .01 / x.size)])
    assert lo * 0.995 <= s["p99"] <= hi * 1.005                                  # binomial order-statistic bound + bin width
    assert abs(s["fail_rate"] - (e_ref < 0).mean()) < 4.5 * np.sqrt(max((e_ref < 0).mean(), 1e-5) / n_ref) + 1e-4


def test_full_size_properties(engine):
    """BASELINE config 2 scale (2^26 trials per launch): size-independent properties instead of an oracle run."""
    n = 1 << 26
    dist0 = make_dist(BASIC_MEAN, ZERO_SIGMA, 0.0)
    dcols = torch.empty((13, n), dtype=torch.float64, device="cuda")
    err = torch.empty(n, dtype=torch.float64, device="cuda")
    ndet = torch.empty(n, dtype=torch.int32, device="cuda")
    stats = torch.zeros(engine.stats_len, dtype=torch.float64, device="cuda")
    engine.fill_presampled_device(dcols, dist0, seed=6)
    engine.run_presampled_device(dcols, err, ndet, stats, seed=6)
    torch.cuda.synchronize()
    ok = err >= 0
    assert float((err[ok] / ARCSEC).max()) < 1e-9                               # ideal tracker at every one of 6.7e7 boresights
    st = stats.cpu().numpy()
    assert st[0] + st[1] == n and st[0] == int(ok.sum()) and st[4] == int(ndet.sum(dtype=torch.int64))
    assert st[8:].sum() + st[6] + st[7] == st[0]
    assert int((ndet[~ok] >= 3).sum()) == 0 and int((ndet[ok] < 3).sum()) == 0
    # default config: idempotent (same seed -> identical records), seed-sensitive, stats consistent with records
    dist = make_dist(BASIC_MEAN, BASIC_SIGMA, THERMAL_COEF)
    engine.fill_presampled_device(dcols, dist, seed=6)
    err2 = torch.empty_like(err); ndet2 = torch.empty_like(ndet)
    engine.run_presampled_device(dcols, err, ndet, None, seed=6)
    engine.run_presampled_device(dcols, err2, ndet2, None, seed=6)
    torch.cuda.synchronize()
    assert torch.equal(err, err2) and torch.equal(ndet, ndet2)
    engine.run_presampled_device(dcols, err2, ndet2, None, seed=7)
    torch.cuda.synchronize()
    assert not torch.equal(err, err2)
    stats.zero_()
    engine.run_rng_device(dist, n, stats, seed=6)
    torch.cuda.synchronize()
    st = stats.cpu().numpy()
    ok = err >= 0
    assert st[0] == int(ok.sum()) and st[4] == int(ndet.sum(dtype=torch.int64))
    assert st[2] == pytest.approx(float(err[ok].sum()), rel=1e-9) and st[5] == float(err.max())


def test_custom_catalog_with_non_identity_star_ids(lib_built, catalog):
    """A catalog file in arbitrary row order: star ids (Philox keys) are file rows, not sky-grid positions."""
    from startrackermpm_b200.engine import TrialEngine
    xyz, mag = catalog
    perm = np.random.default_rng(5).permutation(len(mag))[:6000]             # also a different catalog size
    eng = TrialEngine(0, catalog=(xyz[perm], mag[perm]), f_ideal=F)
    assert not np.array_equal(eng.grid.star_id, np.arange(len(perm)))
    cols = trial.sample_params(dict(BASIC_MEAN, MAX_MAGNITUDE=6.6), BASIC_SIGMA, THERMAL_COEF, 77, 0, 6000)
    e_ref, d_ref = trial.run_trials(cols, xyz[perm], mag[perm], F, 77, 0)
    for prec in (64, 32):
        compare(*eng.run_presampled(cols, 77, 0, precision=prec), e_ref, d_ref, prec)
    eng.close()


def test_star_mag_output_is_bit_exact(engine, catalog):
    """star_mag objective: faintest detected star's catalog magnitude (a max over float32 catalog values: exact)."""
    xyz, mag = catalog
    cols = trial.sample_params(dict(BASIC_MEAN, MAX_MAGNITUDE=4.6), BASIC_SIGMA, THERMAL_COEF, 41, 0, 8000)
    e_ref, d_ref, aux = trial.run_trials(cols, xyz, mag, F, 41, 0, return_aux=True)
    for prec in (64, 32):
        e, d, m = engine.run_presampled(cols, 41, 0, precision=prec, want_maxmag=True)
        compare(e, d, e_ref, d_ref, prec)
        assert np.array_equal(m, aux["maxmag"])
    assert np.isneginf(m[d_ref == 0]).all() and (d_ref == 0).any()


def test_toolplot_reductions_on_device(engine):
    """toolplot.py:82-99 (RMS, 6-sigma clip, 50-bin density histogram, mode, sigma about the mode) from device records."""
    n = 400_000
    dist = make_dist(BASIC_MEAN, BASIC_SIGMA, THERMAL_COEF)
    err = torch.empty(n, dtype=torch.float64, device="cuda")
    ndet = torch.empty(n, dtype=torch.int32, device="cuda")
    stats = torch.zeros(engine.stats_len, dtype=torch.float64, device="cuda")
    engine.run_rng_device(dist, n, stats, err, ndet, seed=8)
    torch.cuda.synchronize()
    x = err.cpu().numpy()
    x[:50] = 1e5                                                       # outliers beyond the 6-sigma clip
    x[50:60] = -1.0
    err.copy_(torch.from_numpy(x))
    for bins in (50, 7):
        got = engine.toolplot_reduce(err, bins=bins)
        want = ostats.toolplot_reductions(x, bins=bins)
        assert got["n"] == want["n"] and got["n_kept"] == want["n_kept"] and got["n_kept"] < got["n"]
        assert got["kept_min"] == want["kept_min"] and got["kept_max"] == want["kept_max"]
        assert got["mode_bin"] == want["mode_bin"] and got["mode"] == pytest.approx(want["mode"], rel=1e-14)
        assert got["rms"] == pytest.approx(want["rms"], rel=1e-12)
        assert got["sigma_about_mode"] == pytest.approx(want["sigma_about_mode"], rel=1e-10)
        assert np.allclose(got["density"], want["density"], rtol=1e-12, atol=0)      # identical bin counts


@pytest.mark.timeout(120)
def test_garbage_inputs_do_not_hang_or_fault(engine):
    """NaN / inf / absurd magnitudes in any column: the kernel must return (no hang, no fault); such trials simply fail
    or produce non-finite errors.  Good rows in the same launch are unaffected."""
    n = 4096
    good = trial.sample_params(BASIC_MEAN, BASIC_SIGMA, THERMAL_COEF, 13, 0, n)
    e_good, d_good = engine.run_presampled(good, 13, 0)
    bad = good.copy()
    vals = [np.nan, np.inf, -np.inf, 1e300, -1e300, 1e-300, 1e9, -1e9]
    rng = np.random.default_rng(0)
    rows = rng.choice(n, 1200, replace=False)
    for k, r in enumerate(rows):
        bad[k % 13, r] = vals[(k // 13) % len(vals)]
    for prec in (64, 32):
        e, d = engine.run_presampled(bad, 13, 0, precision=prec)
        assert e.shape == (n,) and d.shape == (n,)
        keep = np.ones(n, bool); keep[rows] = False
        assert np.array_equal(d[keep], d_good[keep])
        if prec == 64:
            assert np.array_equal(e[keep], e_good[keep])
    st = engine.run_rng(make_dist(dict(BASIC_MEAN, FOCAL_LENGTH=np.nan), BASIC_SIGMA, 0.0), 2048, 1)   # NaN distribution
    assert st[0][0] + st[0][1] == 2048


def test_two_engines_with_different_catalog_sizes_coexist(engine, catalog):
    """Regression: the dynamic-shared-memory attribute is per kernel function, not per handle — a second engine with a
    smaller catalog must not break launches of the first."""
    from startrackermpm_b200.engine import TrialEngine
    xyz, mag = catalog
    cols = trial.sample_params(BASIC_MEAN, BASIC_SIGMA, THERMAL_COEF, 3, 0, 600)
    e0, d0 = engine.run_presampled(cols, 3, 0)
    small = TrialEngine(0, catalog=(xyz[:3000], mag[:3000]), f_ideal=F)
    small.run_presampled(cols, 3, 0)
    e1, d1 = engine.run_presampled(cols, 3, 0)            # the big-catalog engine again
    small.close()
    assert np.array_equal(e0, e1) and np.array_equal(d0, d1)


def test_records_do_not_depend_on_the_sort_window(engine):
    """The work-key sort and the dynamic group scheduling only reorder WHICH lanes run a trial: per-trial records are
    bit-identical for every window size (stmpm_set_tuning's window is the kernel's tuning knob), in both modes, and the streamed
    moments / histogram account for every trial exactly once."""
    n = 300_001                                            # ragged: last window / group partially filled
    cols = trial.sample_params(BASIC_MEAN, BASIC_SIGMA, THERMAL_COEF, 33, 7, n)
    dist = make_dist(BASIC_MEAN, BASIC_SIGMA, THERMAL_COEF)
    ref = ref_rng = None
    for win in (32, 96, 256, 1024, 2048):
        engine.set_tuning(window=win)
        e, d, st = engine.run_presampled(cols, 33, 7, want_stats=True)
        if ref is None:
            ref = (e, d)
        assert np.array_equal(e, ref[0]) and np.array_equal(d, ref[1]), f"window {win} changed per-trial records"
        ok = e >= 0
        assert st[0] == ok.sum() and st[1] == (~ok).sum() and st[4] == d.sum()
        assert st[6] + st[7] + st[8:].sum() == ok.sum()                         # every successful trial in one bin
        assert abs(st[2] - e[ok].sum()) <= 1e-9 * e[ok].sum() and st[5] == e[ok].max()
        # several segments in one launch: per-segment statistics flushed at the segment changes
        s3, e3, d3 = engine.run_rng([dist, dist, dist], 70_001, 5, 11, records=True)
        if ref_rng is None:
            ref_rng = (e3, d3)
        assert np.array_equal(e3, ref_rng[0]) and np.array_equal(d3, ref_rng[1])
        for k in range(3):
            ek, dk = e3[k * 70_001:(k + 1) * 70_001], d3[k * 70_001:(k + 1) * 70_001]
            assert s3[k, 0] == (ek >= 0).sum() and s3[k, 1] == (ek < 0).sum() and s3[k, 4] == dk.sum()
    engine.set_tuning(window=0)


def test_compact_records_and_host_pipeline_knob(engine):
    """stmpm_run_rng_host_compact: n_det as uint8 (9-byte records) — same errors, same counts (saturated at 255); the
    host-pipeline chunk size never changes a record; the copy-only roofline call moves the bytes it says."""
    dist = make_dist(dict(BASIC_MEAN, MAX_MAGNITUDE=20.0), BASIC_SIGMA, THERMAL_COEF)       # ~60 detections per trial
    n = 70_001
    st, e, d = engine.run_rng(dist, n, 3, 11, records=True)
    for chunk in (4096, 0):
        engine.set_host_pipeline(chunk)
        st8, e8, d8 = engine.run_rng(dist, n, 3, 11, records=True, compact=True)
        assert d8.dtype == np.uint8 and np.array_equal(e8, e) and np.array_equal(d8, np.minimum(d, 255).astype(np.uint8))
        assert np.array_equal(st8[0, :2], st[0, :2]) and np.array_equal(st8[0, 4:], st[0, 4:])
    st5, e5, d5 = engine.run_rng(dist, n, 3, 11, records=True, compact=5)                          # 5-byte records: float32 error
    assert e5.dtype == np.float32 and np.array_equal(e5, e.astype(np.float32)) and np.array_equal(d5, d8)
    assert np.array_equal(st5[0, :2], st[0, :2]) and np.array_equal(st5[0, 4:], st[0, 4:])         # statistics from the float64 error
    wide = make_dist(dict(BASIC_MEAN, MAX_MAGNITUDE=20.0, FOCAL_LENGTH=700.0), ZERO_SIGMA, 0.0)   # > 255 stars in the field
    _, _, dw = engine.run_rng(wide, 2000, 3, 0, records=True)
    _, _, dw8 = engine.run_rng(wide, 2000, 3, 0, records=True, compact=True)
    assert dw.max() > 255 and dw8.max() == 255 and np.array_equal(dw8, np.minimum(dw, 255).astype(np.uint8))
    src = torch.empty(104 * n, dtype=torch.uint8, pin_memory=True).numpy()
    dst = torch.empty(12 * n, dtype=torch.uint8, pin_memory=True).numpy()
    assert engine.host_copy_roofline(n, 104, 12, src, dst) > 0 and engine.host_copy_roofline(n, 0, 9, None, dst) > 0


def test_sorted_parameter_staging_changes_no_record(engine):
    """Round 2's sorted parameter staging (pre-sampled windows re-written in key order to an L2 scratch, read coalesced) only
    changes HOW a warp gets its thirteen parameters: records and statistics are bit-identical to the gather path, for
    ragged sizes and every window width."""
    n = 1_400_003                                          # above the staging threshold (4 x 2048 trials per SM), ragged
    d = make_dist(BASIC_MEAN, BASIC_SIGMA, THERMAL_COEF)
    s = torch.cuda.current_stream().cuda_stream
    cols = torch.empty((13, n), dtype=torch.float64, device="cuda")
    engine.fill_presampled_device(cols, d, seed=91, trial0=5, stream=s)
    ref_e = torch.empty(n, dtype=torch.float64, device="cuda")
    ref_d = torch.empty(n, dtype=torch.int32, device="cuda")
    ref_s = torch.zeros(engine.stats_len, dtype=torch.float64, device="cuda")
    engine.set_tuning(stage_columns=False)
    engine.run_presampled_device(cols, ref_e, ref_d, ref_s, seed=91, trial0=5, stream=s)
    torch.cuda.synchronize()
    try:
        for win in (0, 2048, 1024, 480, 64):
            for prec in (64, 32):
                engine.set_tuning(window=win, stage_columns=True, key_table=1 - (win // 64) % 2)
                e = torch.full((n,), -7.0, dtype=torch.float64, device="cuda")
                nd = torch.full((n,), -7, dtype=torch.int32, device="cuda")
                st = torch.zeros(engine.stats_len, dtype=torch.float64, device="cuda")
                engine.run_presampled_device(cols, e, nd, st, seed=91, trial0=5, precision=prec, stream=s)
                torch.cuda.synchronize()
                assert torch.equal(nd, ref_d), (win, prec)
                if prec == 64:
                    assert torch.equal(e, ref_e), win
                    assert torch.equal(st[:2], ref_s[:2]) and torch.equal(st[4:], ref_s[4:])      # counts, n_det, max, histogram
                else:
                    ok = ref_e >= 0
                    assert torch.equal(e < 0, ~ok) and float((e[ok] - ref_e[ok]).abs().max()) / ARCSEC < 1e-6
    finally:
        engine.set_tuning()


@pytest.mark.parametrize("prec", [64, 32])
def test_split_pipeline_equals_single_kernel(engine, catalog, prec):
    """Round 2's two-kernel pipeline (candidate scan -> counting sort by survivor count -> star loop + solve over warps of
    identical star counts) must reproduce the single persistent kernel: detection counts and failure flags identical,
    errors within 1e-13 rad (the same functions inlined into two kernels: the compiler's fma contraction may differ in the
    last bit; trials the scan hands to the sky-grid walk sum B in another order); statistics account for every trial.  Also against the oracle, and on the workloads that trigger the device-side fall-back."""
    xyz, mag = catalog
    s = torch.cuda.current_stream().cuda_stream
    cases = {
        "default": (BASIC_MEAN, BASIC_SIGMA, 300_001),
        "sigma_x5": (BASIC_MEAN, {k: 5 * v for k, v in BASIC_SIGMA.items()}, 200_003),        # ~0.3 % slow trials (wide cones)
        "all_visible": (dict(BASIC_MEAN, MAX_MAGNITUDE=20.0), BASIC_SIGMA, 100_000),          # every trial overflows: fall-back
        "few": (dict(BASIC_MEAN, MAX_MAGNITUDE=4.0), BASIC_SIGMA, 100_000),                   # many failed / zero-star trials
    }
    try:
        for name, (mean, sigma, n) in cases.items():
            d = make_dist(mean, sigma, THERMAL_COEF)
            cols = torch.empty((13, n), dtype=torch.float64, device="cuda")
            engine.fill_presampled_device(cols, d, seed=77, trial0=3, stream=s)
            out = {}
            for pipe in (0, 1):
                engine.set_pipeline(pipe)
                e = torch.full((n,), -7.0, dtype=torch.float64, device="cuda")
                nd = torch.full((n,), -7, dtype=torch.int32, device="cuda")
                st = torch.zeros(engine.stats_len, dtype=torch.float64, device="cuda")
                engine.run_presampled_device(cols, e, nd, st, seed=77, trial0=3, precision=prec, stream=s)
                torch.cuda.synchronize()
                engine.check()
                out[pipe] = (e.cpu().numpy(), nd.cpu().numpy(), st.cpu().numpy())
            (e0, d0, s0), (e1, d1, s1) = out[0], out[1]
            assert np.array_equal(d0, d1), name
            assert np.array_equal(e0 < 0, e1 < 0), name
            ok = e0 >= 0
            diff = np.abs(e0[ok] - e1[ok]) / ARCSEC
            assert diff.max() <= 1e-13, (name, diff.max())
            assert np.array_equal(s0[:2], s1[:2]) and s0[4] == s1[4] and s1[6] + s1[7] + s1[8:].sum() == ok.sum()
            if name != "all_visible":
                assert np.array_equal(s0[5:], s1[5:]) or diff.max() > 0                         # same maximum and histogram
            m = min(n, 20_000)
            e_ref, d_ref = trial.run_trials(cols[:, :m].cpu().numpy(), xyz, mag, F, 77, 3, grid=trial.candidate_grid(xyz))
            compare(e1[:m], d1[:m], e_ref, d_ref, prec)
    finally:
        engine.set_pipeline(-1)
