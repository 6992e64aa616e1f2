"""Study drivers (configs 3-5 callers): distribution builders on CPU, launches on the GPU."""
import numpy as np
import pandas as pd
import pytest

from conftest import BASIC_MEAN, BASIC_SIGMA, THERMAL_COEF


def test_grid_dists_layout(lib_built):
    from startrackermpm_b200 import studies
    ea, pa, da = studies.grid_axes(1.0, 0.05, 0.1, 4, 3, 2)
    assert ea[0] == 0 and ea[-1] == pytest.approx(3.0) and pa[-1] == pytest.approx(0.15) and da[-1] == pytest.approx(0.3)
    dists = studies.sensitivity_grid_dists(BASIC_MEAN, BASIC_SIGMA, THERMAL_COEF, ea, pa, da)
    assert len(dists) == 24
    d = dists[(2 * 3 + 1) * 2 + 1]                                     # C order (eps, phi, dev)
    names = studies.NORMAL_PARAMS
    assert d.mean[names.index("F_ARR_EPS_Y")] == pytest.approx(ea[2]) and d.sigma[names.index("F_ARR_EPS_Y")] == 0
    assert d.mean[names.index("F_ARR_PSI")] == pytest.approx(pa[1]) and d.mean[names.index("BASE_DEV_X")] == pytest.approx(da[1])
    assert d.sigma[names.index("FOCAL_LENGTH")] == 2.0 and d.thermal_coef == THERMAL_COEF


def test_leo_schedule(lib_built):
    from startrackermpm_b200 import studies
    dists, sched = studies.leo_lifetime_dists(BASIC_MEAN, BASIC_SIGMA, THERMAL_COEF, n_steps=60)
    assert len(dists) == 60 and sched[0]["t_years"] == 0 and sched[-1]["t_years"] == pytest.approx(5.0)
    mags = [s["MAX_MAGNITUDE"] for s in sched]
    assert mags[0] == 6.0 and all(a > b for a, b in zip(mags, mags[1:]))
    assert abs(max(s["D_TEMP"] for s in sched)) <= 15.0 and sched[-1]["dev_sigma"] > sched[0]["dev_sigma"]


@pytest.mark.gpu
def test_sensitivity_grid_on_gpu(engine):
    from startrackermpm_b200 import sharding, studies
    ea, pa, da = studies.grid_axes(1.0, 0.05, 0.1, 3, 3, 2)
    tps = 4000
    stats = studies.sensitivity_grid(engine, BASIC_MEAN, BASIC_SIGMA, 0.0, ea, pa, da, tps, seed=3)
    assert stats.shape == (18, engine.stats_len) and (stats[:, 0] + stats[:, 1] == tps).all()
    s = studies.summaries(stats)
    mean = np.array([x["mean"] for x in s]).reshape(3, 3, 2)
    assert mean[0, 0, 0] < 40                                             # only focal-length + magnitude scatter left
    assert (np.diff(mean[:, 0, 0]) > 0).all() and (np.diff(mean[0, :, 0]) > 0).all() and (mean[0, 0, 1] > mean[0, 0, 0])
    # rank-sharded subsets reproduce their grid points exactly (per-point statistics never need a reduction)
    for r in range(2):
        seg = list(sharding.shard_segments(18, r, 2))
        part = studies.sensitivity_grid(engine, BASIC_MEAN, BASIC_SIGMA, 0.0, ea, pa, da, tps, seed=3, segments=seg)
        assert np.array_equal(part[:, [0, 1, 4, 5]], stats[seg][:, [0, 1, 4, 5]]) and np.array_equal(part[:, 8:], stats[seg][:, 8:])


@pytest.mark.gpu
def test_leo_lifetime_on_gpu(engine):
    from startrackermpm_b200 import studies
    dists, sched = studies.leo_lifetime_dists(BASIC_MEAN, BASIC_SIGMA, THERMAL_COEF, n_steps=12, mag_loss_per_year=0.4)
    s = studies.summaries(studies.run_segments(engine, dists, 20_000, seed=4))
    assert s[-1]["mean_ndet"] < 0.5 * s[0]["mean_ndet"]                    # radiation: fewer detectable stars
    assert s[-1]["fail_rate"] > s[0]["fail_rate"]


@pytest.mark.gpu
def test_monte_carlo_stop_rule_matches_reference_loop(engine):
    """monte_carlo.py:48-76 replayed on per-trial records with pandas must stop at the same batch with the same sigma."""
    from startrackermpm_b200 import studies
    from startrackermpm_b200.engine import make_dist
    tol, min_rows, batch, seed = 2e-3, 3000, 100, 12
    got = studies.monte_carlo_converged(engine, BASIC_MEAN, BASIC_SIGMA, THERMAL_COEF, batch=batch, seed=seed,
                                        min_rows=min_rows, tol=tol)
    _, err, _ = engine.run_rng(make_dist(BASIC_MEAN, BASIC_SIGMA, THERMAL_COEF), got["trials"] + 5 * batch, seed, 0,
                               records=True)
    k = np.sqrt(1 - 2 / np.pi)
    fin, prev, ratio, count = pd.Series(dtype=float), 1.0, 1.0, 0
    while ratio > tol or len(fin) < min_rows:                              # the reference's loop, verbatim structure
        chunk = pd.Series(err[count * batch:(count + 1) * batch])
        count += 1
        fin = pd.concat([fin, chunk[chunk >= 0]])
        ratio = abs((prev - (fin / k).std()) / prev)
        prev = fin.std() / k
    assert count == got["count"] and got["n_ok"] == len(fin)
    assert got["sigma_halfnormal"] == pytest.approx(prev, rel=1e-10)


@pytest.mark.gpu
def test_run_sh_study_matrix_on_gpu(engine):
    """run.sh's nine studies (MC + diagonal sweep each) and the analytic shapes of their sweeps."""
    from startrackermpm_b200 import studies
    res = studies.study_matrix(engine, BASIC_MEAN, BASIC_SIGMA, mc_trials=5000, sweep_repeats=100, seed=7)
    assert set(res) == set(studies.STUDY_TAGS)
    psi = res["PHI_LONG"]["sweep"]
    assert np.allclose(psi["mean"], np.abs(psi["values"]["F_ARR_PSI"]) * 3600.0, atol=1e-6)     # pure psi -> |psi| exactly
    lat = res["EPS_LAT"]["sweep"]
    r = np.hypot(lat["values"]["F_ARR_EPS_X"], lat["values"]["F_ARR_EPS_Y"])
    assert np.allclose(lat["mean"], np.arctan(r / 3500.0) * 206264.806, rtol=0.03, atol=0.5)     # README.md:13 "consistent bias"
    cen = res["CENTROID"]["sweep"]
    assert cen["mean"][0] < 1e-6 and np.all(np.diff(cen["mean"][::10]) > 0)                     # grows with centroid sigma
    stars = res["STARS"]["sweep"]
    assert stars["fail_rate"][0] > 0.9 and stars["fail_rate"][-1] == 0.0                        # mag 3 -> 9
    assert res["EPS_LONG"]["mc"]["mean"] < res["EPS_LAT"]["mc"]["mean"]                         # vertical shift hurts less
    assert res["PHI"]["mc"]["sigma_halfnormal"] > res["CENTROID"]["mc"]["sigma_halfnormal"]
