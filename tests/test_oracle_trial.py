"""Known-answer tests of the numpy oracle (the reference ships none: SURVEY.md §4, §8c lists these in lieu)."""
import numpy as np
import pytest
from scipy.spatial.transform import Rotation

from conftest import BASIC_MEAN, BASIC_SIGMA, THERMAL_COEF, ZERO_SIGMA
from oracle import trial

F = 3500.0
ARCSEC = trial.RAD2ARCSEC


def ideal_cols(n, seed=3, **over):
    c = trial.sample_params(BASIC_MEAN, ZERO_SIGMA, 0.0, seed, 0, n)
    for k, v in over.items():
        c[trial.COLUMNS.index(k)] = v
    return c


def test_rotation_primitives_match_reference_constants():
    # /root/reference/constants.py:81-94 — active, right-handed
    a = 0.3
    assert np.allclose(trial.Rx(a) @ [0, 1, 0], [0, np.cos(a), np.sin(a)])
    assert np.allclose(trial.Ry(a) @ [0, 0, 1], [np.sin(a), 0, np.cos(a)])
    assert np.allclose(trial.Rz(a) @ [1, 0, 0], [np.cos(a), np.sin(a), 0])


def test_truth_attitude_boresight_and_orthonormal():
    ra, dec, roll = np.array([0.4, 5.1]), np.array([-0.7, 1.2]), np.array([2.0, -3.0])
    C = trial.truth_attitude(ra, dec, roll)
    assert np.allclose(C[:, :, 2], np.stack([np.cos(dec) * np.cos(ra), np.cos(dec) * np.sin(ra), np.sin(dec)], -1))
    assert np.allclose(C @ np.swapaxes(C, 1, 2), np.eye(3), atol=1e-15)
    assert np.allclose(np.linalg.det(C), 1.0)


def test_ideal_tracker_error_below_1e12_rad(catalog):           # (i) thesis "BASE / Ideal Star Tracker" case
    xyz, mag = catalog
    e, d = trial.run_trials(ideal_cols(300), xyz, mag, F, seed=1)
    assert (e >= 0).all() and (d >= 3).all()
    assert (e / ARCSEC).max() < 1e-12


def test_wahba_recovers_known_rotation_vs_scipy():               # (ii)
    rng = np.random.default_rng(5)
    r = rng.normal(size=(12, 3)); r /= np.linalg.norm(r, axis=1, keepdims=True)
    R = Rotation.from_rotvec([0.3, -1.1, 0.7])
    b = R.apply(r)                                                 # b = A r
    B = (b[:, :, None] * r[:, None, :]).mean(0)
    q, lam = trial.q_method(B[None])
    assert lam[0] == pytest.approx(1.0, abs=1e-14)
    assert np.allclose(trial.quat_to_dcm(q)[0], R.as_matrix(), atol=1e-14)
    est, _ = Rotation.align_vectors(b, r)
    assert np.allclose(est.as_matrix(), R.as_matrix(), atol=1e-9)
    q2, _ = trial.quest(B[None])
    assert np.allclose(np.abs(q2[0] @ q[0]), 1.0, atol=1e-12)


def test_pure_psi_rotation_gives_exactly_psi(catalog):           # (iii)
    xyz, mag = catalog
    for psi_deg in (0.01, 0.2, -1.0):
        e, _ = trial.run_trials(ideal_cols(40, F_ARR_PSI=psi_deg), xyz, mag, F, seed=1)
        assert np.allclose(e / ARCSEC, np.deg2rad(abs(psi_deg)), rtol=0, atol=1e-12)


def test_lateral_translation_bias_is_atan_eps_over_f(catalog):   # (iv) README.md:13 "consistent bias"
    xyz, mag = catalog
    for eps in (0.5, 2.0, 8.0):
        e, _ = trial.run_trials(ideal_cols(200, F_ARR_EPS_X=eps), xyz, mag, F, seed=1)
        want = np.arctan(eps / F) * ARCSEC
        assert np.allclose(e, want, rtol=0.02)
        assert e.std() / e.mean() < 0.01


def test_focal_error_with_symmetric_field_has_no_bias():         # (v)
    t = np.deg2rad(4.0)
    body = np.array([[np.sin(t), 0, np.cos(t)], [-np.sin(t), 0, np.cos(t)],
                     [0, np.sin(t), np.cos(t)], [0, -np.sin(t), np.cos(t)]])
    cols = ideal_cols(1, RIGHT_ASCENSION=1.0, DECLINATION=0.3, ROLL=0.5, FOCAL_LENGTH=F + 30.0)
    C = trial.truth_attitude(cols[0], cols[1], cols[2])[0]
    xyz = body @ C.T                                               # inertial vectors of stars symmetric about boresight
    e, d = trial.run_trials(cols, xyz, np.zeros(4, np.float32), F, seed=1, brute=True)
    assert d[0] == 4 and e[0] / ARCSEC < 1e-12


def test_magnitude_threshold_extremes(catalog):                  # (vi) sensitivity.py:112 uses MAX_MAGNITUDE = 20
    xyz, mag = catalog
    e, d = trial.run_trials(ideal_cols(100, MAX_MAGNITUDE=float(mag.min()) - 0.1), xyz, mag, F, seed=1)
    assert (e == -1.0).all() and (d == 0).all()
    e, d = trial.run_trials(ideal_cols(100, MAX_MAGNITUDE=20.0), xyz, mag, F, seed=1)
    assert (e >= 0).all() and d.min() >= 20


def test_mag_compare_is_float32_promoted(catalog):
    xyz, mag = catalog
    i = int(np.argsort(mag)[len(mag) // 2])
    cols = ideal_cols(1, MAX_MAGNITUDE=float(mag[i]))
    C = trial.truth_attitude(cols[0], cols[1], cols[2])[0]
    one = (C[:, 2])[None, :]                                       # a single star exactly on the boresight
    _, d_eq = trial.run_trials(cols, one, mag[i:i + 1], F, seed=1, brute=True)
    cols[12] = np.nextafter(float(mag[i]), -np.inf)
    _, d_lt = trial.run_trials(cols, one, mag[i:i + 1], F, seed=1, brute=True)
    assert d_eq[0] == 1 and d_lt[0] == 0


def test_cone_prefilter_equals_brute_force(catalog):
    xyz, mag = catalog
    s = {k: 4 * v for k, v in BASIC_SIGMA.items()}
    cols = trial.sample_params(BASIC_MEAN, s, THERMAL_COEF, 9, 0, 96)
    e1, d1 = trial.run_trials(cols, xyz, mag, F, seed=9)
    e2, d2 = trial.run_trials(cols, xyz, mag, F, seed=9, brute=True, chunk=16)
    assert (d1 == d2).all() and np.array_equal(e1, e2)


def test_solvers_agree(catalog):                                 # q-method (eigh) vs QUEST Newton in the body frame
    xyz, mag = catalog
    cols = trial.sample_params(BASIC_MEAN, BASIC_SIGMA, THERMAL_COEF, 4, 0, 1500)
    e1, _ = trial.run_trials(cols, xyz, mag, F, seed=4)
    e2, _ = trial.run_trials(cols, xyz, mag, F, seed=4, solver="quest_body")
    ok = e1 >= 0
    assert np.abs(e1[ok] - e2[ok]).max() / ARCSEC < 1e-11


def test_invariances(catalog):                                   # (vii)
    xyz, mag = catalog
    cols = trial.sample_params(BASIC_MEAN, dict(BASIC_SIGMA, BASE_DEV_X=0.0, BASE_DEV_Y=0.0), 0.0, 6, 0, 64)
    e0, d0 = trial.run_trials(cols, xyz, mag, F, seed=6)
    perm = np.random.default_rng(0).permutation(len(mag))          # star ordering (noise-free: ids only key the noise)
    e1, d1 = trial.run_trials(cols, xyz[perm], mag[perm], F, seed=6)
    assert (d0 == d1).all() and np.abs(e0 - e1).max() / ARCSEC < 1e-12
    c2 = cols.copy(); c2[0] += 2 * np.pi                            # RA re-parameterisation
    e2, d2 = trial.run_trials(c2, xyz, mag, F, seed=6)
    assert (d0 == d2).all() and np.abs(e0 - e2).max() / ARCSEC < 1e-11
    B = np.random.default_rng(1).normal(size=(5, 3, 3))
    qa, _ = trial.q_method(B); qb, _ = trial.q_method(3.7 * B)      # weight rescaling
    assert np.allclose(np.abs(np.einsum("ni,ni->n", qa, qb)), 1.0, atol=1e-12)


def test_trial_index_keys_the_noise(catalog):
    xyz, mag = catalog
    cols = trial.sample_params(BASIC_MEAN, BASIC_SIGMA, 0.0, 2, 100, 50)
    e_all, d_all = trial.run_trials(cols, xyz, mag, F, seed=2, trial0=100)
    e_sub, d_sub = trial.run_trials(cols[:, 20:30], xyz, mag, F, seed=2, trial0=120)   # any shard reproduces its slice
    assert np.array_equal(e_all[20:30], e_sub) and np.array_equal(d_all[20:30], d_sub)
    e_other, _ = trial.run_trials(cols[:, 20:30], xyz, mag, F, seed=2, trial0=0)
    assert not np.array_equal(e_other, e_sub)


def test_sample_params_ranges_and_thermal_update():
    c = trial.sample_params(BASIC_MEAN, BASIC_SIGMA, THERMAL_COEF, 8, 0, 50_000)
    fov = np.deg2rad(10)                                           # monte_carlo.py:87-92
    assert fov <= c[0].min() and c[0].max() <= 2 * np.pi - fov
    assert -np.pi / 2 + fov <= c[1].min() and c[1].max() <= np.pi / 2 - fov
    assert -np.pi <= c[2].min() and c[2].max() <= np.pi
    assert abs(c[4].std() - 1.0) < 0.02 and abs(c[12].mean() - 6.0) < 0.01
    # f = F (1 + D_TEMP coef): sigma_f^2 = 2^2 + (3500 * 5 * 2e-5)^2     (sensitivity.py:115-117)
    assert c[3].std() == pytest.approx(np.hypot(2.0, 3500 * 5 * THERMAL_COEF), rel=0.02)


def test_oracle_reproduces_golden_fixture(catalog, golden):
    xyz, mag = catalog
    e, d = trial.run_trials(golden["cols"], xyz, mag, float(golden["f_ideal"]), int(golden["seed"]), int(golden["trial0"]))
    assert np.array_equal(d, golden["n_det"])
    assert np.array_equal(e < 0, golden["err_arcsec"] < 0)
    ok = e >= 0
    assert np.abs(e[ok] - golden["err_arcsec"][ok]).max() / ARCSEC < 1e-10


def test_oracle_reproduces_the_independent_golden_v2(catalog, golden_v2):
    """trial_golden_v2.npz comes from tests/golden/independent_trial.py — a scalar per-row restatement of MODEL.md that
    shares no function with this oracle (own Philox, libm Box-Muller, scipy's SVD align_vectors).  Counts and failure
    flags must be identical, errors within 1e-12 rad (two different Wahba solvers in float64)."""
    xyz, mag = catalog
    g = golden_v2
    e, d = trial.run_trials(g["cols"], xyz, mag, float(g["f_ideal"]), int(g["seed"]), int(g["trial0"]), brute=True, chunk=16)
    assert np.array_equal(d, g["n_det"])
    assert np.array_equal(e < 0, g["err_arcsec"] < 0)
    ok = e >= 0
    assert np.abs(e[ok] - g["err_arcsec"][ok]).max() / trial.RAD2ARCSEC < 1e-12
    assert set(np.unique(g["scenario"])) >= {"basic", "ideal", "all_visible", "few_stars", "sigma_x5", "poles_wrap"}
    assert (g["err_arcsec"] < 0).sum() > 20 and g["n_det"].max() > 90          # failures and star-rich fields are in it


def test_oracle_sampler_matches_the_independent_sampler():
    """MODEL.md §6 restated twice: oracle.sample_params (vectorised numpy) vs independent_trial.draw_rows (Python ints)."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import independent_trial as it
    from conftest import BASIC_MEAN, BASIC_SIGMA
    seed, t0, n = 0xABCDEF0123456789, (1 << 32) - 7, 40
    a = trial.sample_params(BASIC_MEAN, BASIC_SIGMA, 2e-5, seed, t0, n)
    b = it.draw_rows(BASIC_MEAN, BASIC_SIGMA, 2e-5, seed, t0, n)
    assert np.abs(a - b).max() < 1e-11
    assert np.array_equal(a[:3], b[:3])                                         # the uniform draws are exact


def test_independent_restatement_known_answers(catalog):
    """The second restatement is itself pinned: an ideal tracker has zero error, a pure psi tilt returns psi."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import independent_trial as it
    xyz, mag = catalog
    row = [1.0, 0.3, -2.0, 3500.0, 0, 0, 0, 0, 0, 0, 0, 0, 6.0]
    e, n, _ = it.one_trial(row, xyz, mag, 3500.0, 1, 0)
    assert n >= 3 and e / trial.RAD2ARCSEC < 1e-12
    row[9] = 0.25                                                               # psi = 0.25 deg about the boresight
    e, n, _ = it.one_trial(row, xyz, mag, 3500.0, 1, 0)
    assert abs(e / trial.RAD2ARCSEC - np.deg2rad(0.25)) < 1e-12


def test_sky_grid_candidate_search_equals_brute_force(catalog):
    """The CPU arm's sky-grid candidate lists (oracle.trial.CandidateGrid: the fair baseline bench.py times) and the
    noise-only-for-live-pairs shortcut change no result: identical counts and errors to brute force, including trials
    whose cone exceeds the lists' (wide FOV, 5x sigmas) and boresights on the poles / RA wrap."""
    from conftest import BASIC_MEAN, BASIC_SIGMA
    xyz, mag = catalog
    g = trial.candidate_grid(xyz)
    cases = [trial.sample_params(BASIC_MEAN, BASIC_SIGMA, 2e-5, 3, 0, 600),
             trial.sample_params(BASIC_MEAN, {k: 5 * v for k, v in BASIC_SIGMA.items()}, 2e-5, 3, 0, 300),
             trial.sample_params(dict(BASIC_MEAN, FOCAL_LENGTH=1400.0), BASIC_SIGMA, 0.0, 3, 0, 60),
             trial.sample_params(dict(BASIC_MEAN, MAX_MAGNITUDE=20.0), BASIC_SIGMA, 0.0, 3, 0, 100)]
    polar = trial.sample_params(BASIC_MEAN, BASIC_SIGMA, 0.0, 3, 0, 90)
    polar[1] = np.linspace(-np.pi / 2, np.pi / 2, 90)
    polar[0] = np.linspace(-0.5, 2 * np.pi + 0.5, 90)
    for cols in cases + [polar]:
        e0, d0 = trial.run_trials(cols, xyz, mag, 3500.0, 3, 0, brute=True, chunk=32)
        e1, d1, aux = trial.run_trials(cols, xyz, mag, 3500.0, 3, 0, grid=g, return_aux=True)
        assert np.array_equal(d0, d1) and np.array_equal(e0, e1)
    assert aux["n_cand"] / 90 < 400                                              # ~170 listed stars per trial, not 9 110
