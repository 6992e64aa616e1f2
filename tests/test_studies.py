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


def test_result_frame_is_the_file_toolplot_reads(lib_built, tmp_path):
    """the pickled frame has the reference's columns, drops failed rows like Simulation.run_sim (or
    keeps them with the -1 sentinel for `toolplot.py -s s`), and toolplot.py:82-99 / 168-180 evaluated on the re-read
    file reproduce the oracle's reductions.  Host logic only: oracle engine double, pre-sampled columns."""
    from oracle import stats as ostats
    from oracle import trial
    from oracle_engine import OracleEngine
    from startrackermpm_b200 import _abi, studies
    mean = dict(BASIC_MEAN, MAX_MAGNITUDE=3.6)                             # bright limit: a visible share of trials fails
    cols = trial.sample_params(mean, dict(BASIC_SIGMA, MAX_MAGNITUDE=0.0), THERMAL_COEF, 5, 0, 1500)
    cols[12, 750:] = 4.0                                                    # two magnitude levels, as a STARS study has
    eng = OracleEngine()
    kept = studies.result_frame(eng, cols=cols, seed=5)
    full = studies.result_frame(eng, cols=cols, seed=5, keep_failed=True, star_mag=True)
    assert list(kept.columns) == list(_abi.COLUMNS) + [studies.QUEST_OUTPUT, studies.N_DET_OUTPUT]
    assert studies.MAG_OUTPUT in full.columns and len(full) == 1500
    failed = full[studies.QUEST_OUTPUT] < 0
    assert 0 < failed.sum() < 1500 and (full.loc[failed, studies.QUEST_OUTPUT] == -1.0).all()
    assert len(kept) == 1500 - failed.sum() and (kept[studies.QUEST_OUTPUT] >= 0).all()
    path = studies.write_result_pickle(kept, "MC_TEST", str(tmp_path / "data"))
    assert path.endswith("data/MC_TEST.pkl")                                # toolplot.py:72 wants exactly dir/name.pkl
    df = pd.read_pickle(path)
    col = studies.QUEST_OUTPUT
    std = np.sqrt((df[col] ** 2).sum() / len(df.index))                     # toolplot.py:82-99, verbatim arithmetic
    clip = df[df[col] <= 6 * std]
    h = np.histogram(clip[col], bins=50, density=True)
    newsig = np.sqrt(((clip[col] - h[1][np.argmax(h[0])]) ** 2).sum() / len(clip.index))
    ref = ostats.toolplot_reductions(full[col].to_numpy())
    assert std == pytest.approx(ref["rms"], rel=1e-12) and len(clip) == ref["n_kept"]
    assert newsig == pytest.approx(ref["sigma_about_mode"], rel=1e-12)
    data = full[["MAX_MAGNITUDE", col]].copy()                              # toolplot.py:168-180: star-count failure rate
    data.loc[data[col] > 0, col] = 0
    data[col] *= -1
    rates = {m: data[data.MAX_MAGNITUDE == m][col].sum() / (data.MAX_MAGNITUDE == m).sum() for m in data.MAX_MAGNITUDE.unique()}
    assert set(rates) == {3.6, 4.0} and rates[3.6] > rates[4.0] >= 0
    assert rates[3.6] == pytest.approx(failed[:750].mean())


@pytest.mark.gpu
def test_result_frame_on_gpu(engine, tmp_path):
    """The bulk GPU route to the same file: device-generated parameters, records through the C ABI, toolplot's reductions
    from the frame equal the on-device ones."""
    from startrackermpm_b200 import studies
    from startrackermpm_b200.engine import make_dist
    n = 200_000
    df = studies.result_frame(engine, BASIC_MEAN, BASIC_SIGMA, THERMAL_COEF, n, seed=9, keep_failed=True, star_mag=True)
    _, e, d = engine.run_rng(make_dist(BASIC_MEAN, BASIC_SIGMA, THERMAL_COEF), n, 9, 0, records=True)
    assert np.array_equal(df[studies.QUEST_OUTPUT].to_numpy(), e) and np.array_equal(df[studies.N_DET_OUTPUT].to_numpy(), d)
    ok = e >= 0
    assert np.isfinite(df.loc[ok, studies.MAG_OUTPUT]).all() and (df.loc[ok, studies.MAG_OUTPUT] <= df.loc[ok, "MAX_MAGNITUDE"]).all()
    import torch
    tp = engine.toolplot_reduce(torch.from_numpy(e).cuda())
    x = df.loc[ok, studies.QUEST_OUTPUT]
    std = np.sqrt((x ** 2).sum() / len(x))
    assert std == pytest.approx(tp["rms"], rel=1e-12) and int((x <= 6 * std).sum()) == tp["n_kept"]
    kept = studies.result_frame(engine, BASIC_MEAN, BASIC_SIGMA, THERMAL_COEF, n, seed=9)
    assert len(kept) == int(ok.sum())
    assert pd.read_pickle(studies.write_result_pickle(kept, "MC_GPU", str(tmp_path / "data"))).equals(kept)


def test_thermal_balance_from_the_reference_constants(lib_built):
    """the schedule's D_TEMP amplitude comes from a radiative balance on constants.py:99-108."""
    from startrackermpm_b200 import studies
    t_sun, t_ecl = studies.radiative_equilibrium("1U", True), studies.radiative_equilibrium("1U", False)
    assert t_sun == pytest.approx(((0.273 * 1367 + 213) * 0.01 / (5.670367e-8 * 0.06)) ** 0.25) and 150 < t_ecl < t_sun < 230
    assert studies.radiative_equilibrium("3U", True) > t_sun                  # larger absorbing share of the total area
    assert studies.orbit_dtemp_amplitude("1U") == pytest.approx(15.0, abs=1.0)  # the default amplitude of leo_lifetime_dists
    with pytest.raises(ValueError):
        studies.radiative_equilibrium("6U")


@pytest.mark.parametrize("order", ["<", ">"])
def test_bsc5_binary_loader_round_trip(tmp_path, order):
    """A BSC5-format file written by the test (the real file is absent from the reference mount) is read back in either byte
    order: unit vectors, magnitudes / 100, HR numbers; position-less entries dropped."""
    from startrackermpm_b200 import catalog
    rng = np.random.default_rng(4)
    n = 50
    rec = np.dtype([("xno", order + "f4"), ("ra", order + "f8"), ("dec", order + "f8"), ("sp", "S2"), ("mag", order + "i2"),
                    ("pm_ra", order + "f4"), ("pm_dec", order + "f4")])
    ent = np.zeros(n, dtype=rec)
    ent["xno"] = np.arange(1, n + 1)
    ent["ra"] = rng.uniform(0.01, 2 * np.pi, n)
    ent["dec"] = rng.uniform(-1.5, 1.5, n)
    ent["mag"] = rng.integers(-146, 796, n)
    ent["sp"] = b"G2"
    ent["ra"][[7, 31]] = 0.0                                                # two non-stellar rows without a position
    ent["dec"][[7, 31]] = 0.0
    hdr = np.array([0, 1, -n, 1, 1, 1, 32], dtype=order + "i4")
    path = tmp_path / "BSC5"
    path.write_bytes(hdr.tobytes() + ent.tobytes())
    xyz, mag, hr = catalog.load_bsc5(str(path))
    keep = np.setdiff1d(np.arange(n), [7, 31])
    assert xyz.shape == (n - 2, 3) and np.allclose(np.linalg.norm(xyz, axis=1), 1.0, atol=1e-15)
    assert np.array_equal(hr, keep + 1) and np.array_equal(mag, (ent["mag"][keep] / 100.0).astype(np.float32))
    assert np.allclose(np.arctan2(xyz[:, 1], xyz[:, 0]) % (2 * np.pi), ent["ra"][keep]) and np.allclose(np.arcsin(xyz[:, 2]), ent["dec"][keep])
    g = catalog.build_sky_grid(xyz, mag)                                    # feeds the same grid builder as the synthetic catalog
    assert g.n_stars == n - 2 and sorted(g.star_id.tolist()) == list(range(n - 2))
    with pytest.raises(ValueError):
        bad = tmp_path / "bad"
        bad.write_bytes(b"\0" * 64)
        catalog.load_bsc5(str(bad))


@pytest.mark.parametrize("name", ["le", "be"])
def test_bsc5_loader_on_the_hand_built_byte_fixture(name):
    """tests/golden/bsc5_three_stars.*.bin: header + Sirius (HR 2491), Vega (HR 7001), HR 1 and one position-less entry,
    packed byte by byte from the published BSC5 record layout (tests/golden/make_bsc5_fixture.py, `struct.pack`) — not
    written through the loader's own numpy dtype."""
    import os
    from startrackermpm_b200 import catalog
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"bsc5_three_stars.{name}.bin")
    raw = open(path, "rb").read()
    assert len(raw) == 28 + 4 * 32
    # spot-check the bytes themselves against the layout: NBENT = 32 closes the header, the first entry's XNO is 2491.0f
    assert raw[24:28] == ((32).to_bytes(4, "little") if name == "le" else (32).to_bytes(4, "big"))
    assert raw[28:32] == (bytes.fromhex("00b01b45") if name == "le" else bytes.fromhex("451bb000"))       # 2491.0 as IEEE-754 float32
    xyz, mag, hr = catalog.load_bsc5(path)
    assert hr.tolist() == [2491, 7001, 1]                                   # the position-less row is dropped
    assert mag.tolist() == [np.float32(-1.46), np.float32(0.03), np.float32(6.70)]
    ra = np.arctan2(xyz[:, 1], xyz[:, 0]) % (2 * np.pi)
    dec = np.arcsin(xyz[:, 2])
    assert np.allclose(np.rad2deg(ra), [101.28708, 279.23458, 1.29125], atol=1e-4)        # 06h45m08.9s, 18h36m56.3s, 00h05m09.9s
    assert np.allclose(np.rad2deg(dec), [-16.71611, 38.78361, 45.22917], atol=1e-4)
    assert np.allclose(np.linalg.norm(xyz, axis=1), 1.0, atol=1e-15)


@pytest.mark.gpu
def test_on_device_stop_rule_equals_the_host_loop(engine):
    """SURVEY.md §8f rank 1 / VERDICT r1 #6: `stmpm_mc_converge` (rule evaluated on the device after every batch, one host
    read) must stop at the SAME batch with the same sigma as the host loop over per-launch statistics, and as the
    reference's loop replayed with pandas on the per-trial records — at >= 100x the host loop's trials/s."""
    import time
    from startrackermpm_b200 import studies
    from startrackermpm_b200.engine import make_dist
    k = np.sqrt(1 - 2 / np.pi)
    for tol, min_rows, batch, seed, max_runs in ((2e-3, 3000, 100, 12, -1), (1e-5, 10_000, 100, 5, -1), (1e-9, 500, 64, 9, 2000)):
        kw = dict(batch=batch, seed=seed, min_rows=min_rows, tol=tol, max_runs=max_runs)
        dev = studies.monte_carlo_converged(engine, BASIC_MEAN, BASIC_SIGMA, THERMAL_COEF, **kw)
        host = studies.monte_carlo_converged(engine, BASIC_MEAN, BASIC_SIGMA, THERMAL_COEF, on_device=False, **kw)
        assert dev["count"] == host["count"] and dev["n_ok"] == host["n_ok"] and dev["n_fail"] == host["n_fail"]
        assert dev["sigma_halfnormal"] == pytest.approx(host["sigma_halfnormal"], rel=1e-12)
        assert dev["rounds"] <= 2 and dev["converged"]
        assert np.array_equal(dev["stats"][6:], host["stats"][6:]) and dev["stats"][4] == host["stats"][4]   # histogram, n_det
        assert dev["stats"][5] == host["stats"][5]                                                           # x_max
        # the reference's loop, verbatim structure, on the records
        _, err, _ = engine.run_rng(make_dist(BASIC_MEAN, BASIC_SIGMA, THERMAL_COEF), dev["trials"] + 3 * batch, seed, 0, records=True)
        fin, prev, ratio, count = pd.Series(dtype=float), 1.0, 1.0, 0
        while ratio > tol or len(fin) < min_rows:
            chunk = pd.Series(err[count * batch:(count + 1) * batch])
            count += 1
            fin = pd.concat([fin, chunk[chunk >= 0]])
            ratio = abs((prev - (fin / k).std()) / prev)
            prev = fin.std() / k
            if max_runs > 0 and len(fin) >= max_runs:
                break
        assert count == dev["count"] and len(fin) == dev["n_ok"]
        assert dev["sigma_halfnormal"] == pytest.approx(prev, rel=1e-10)
    # speed: the reference's defaults
    kw = dict(batch=100, seed=5, min_rows=10_000, tol=1e-5)
    studies.monte_carlo_converged(engine, BASIC_MEAN, BASIC_SIGMA, THERMAL_COEF, **kw)
    t = time.perf_counter(); dev = studies.monte_carlo_converged(engine, BASIC_MEAN, BASIC_SIGMA, THERMAL_COEF, **kw); t_dev = time.perf_counter() - t
    t = time.perf_counter(); host = studies.monte_carlo_converged(engine, BASIC_MEAN, BASIC_SIGMA, THERMAL_COEF, on_device=False, **kw); t_host = time.perf_counter() - t
    print(f"\nstop rule: {dev['count']} batches; on-device {t_dev * 1e3:.2f} ms ({dev['trials'] / t_dev:.3e} trials/s, "
          f"{dev['trials_computed']} computed) vs host loop {t_host * 1e3:.1f} ms ({host['trials'] / t_host:.3e} trials/s): {t_host / t_dev:.0f}x")
    assert t_host / t_dev >= 30          # ~100x measured; the bound is kept loose against box noise


@pytest.mark.gpu
def test_toolplot_modes_s_and_c_on_device(engine):
    """toolplot.py:170-187 (failure rate per unique MAX_MAGNITUDE) and :240-256 (magnitude column statistics + histogram)
    from device-resident records, against the reference's own pandas arithmetic."""
    import torch
    from startrackermpm_b200.engine import make_dist
    n_lev, per = 25, 4000
    levels = np.linspace(3.0, 6.0, n_lev)
    dists = [make_dist(dict(BASIC_MEAN, MAX_MAGNITUDE=float(m)), dict(BASIC_SIGMA, MAX_MAGNITUDE=0.0), THERMAL_COEF) for m in levels]
    _, err, _ = engine.run_rng(dists, per, 21, 0, records=True)
    mag = np.repeat(levels, per)
    d_err, d_mag = torch.from_numpy(err).cuda(), torch.from_numpy(mag).cuda()
    got = engine.toolplot_failrate(d_err, d_mag, levels)
    # toolplot.py:82-83,170-187 verbatim on a DataFrame
    col = "QUATERNION_ERROR_[arcsec]"
    df = pd.DataFrame({"MAX_MAGNITUDE": mag, col: err})
    std = np.sqrt((df[col] ** 2).sum() / len(df.index))
    df = df[df[col] <= 6 * std]
    data = df[["MAX_MAGNITUDE", col]].copy()
    data.loc[data[col] > 0, col] = 0
    data[col] *= -1
    xdata = data.MAX_MAGNITUDE.unique()
    ydata = [data[data.MAX_MAGNITUDE == m][col].sum() / len(data[data.MAX_MAGNITUDE == m].index) * 100 for m in xdata]
    assert np.array_equal(np.sort(xdata), levels) and got["unmatched_rows"] == 0
    assert np.allclose(got["rate_percent"], np.array(ydata)[np.argsort(xdata)], rtol=0, atol=1e-12)
    assert got["rate_percent"][0] > 50 and got["rate_percent"][-1] < 1
    # mode 'c': the star_mag objective's column
    from oracle import trial
    cols = trial.sample_params(BASIC_MEAN, BASIC_SIGMA, THERMAL_COEF, 8, 0, 50_000)
    _, _, mm = engine.run_presampled(cols, 8, 0, want_maxmag=True)
    c = engine.column_hist(torch.from_numpy(mm).cuda(), bins=50)
    s = pd.Series(mm[np.isfinite(mm)].astype(np.float64))
    assert c["n"] == len(s) and c["min"] == s.min() and c["max"] == s.max()
    assert c["std"] == pytest.approx(s.std(), rel=1e-12) and c["std_halfnormal"] == pytest.approx(s.std() * np.sqrt(1 - 2 / np.pi), rel=1e-12)
    h = np.histogram(s, bins=50, density=True)
    assert np.allclose(c["density"], h[0], rtol=1e-12, atol=0)


@pytest.mark.gpu
def test_tail_statistics_checkpoint_and_resume(engine, tmp_path):
    """SURVEY.md §5 checkpoint/resume: an interrupted chunked run resumed from its 70 KB checkpoint ends with the statistics
    of an uninterrupted run (counts and histogram bit-identical), and refuses a checkpoint of a different run."""
    from startrackermpm_b200 import studies
    from startrackermpm_b200.engine import make_dist
    d = make_dist(BASIC_MEAN, BASIC_SIGMA, THERMAL_COEF)
    n, chunk, ck = 1_000_003, 300_000, str(tmp_path / "tail.npz")
    ref = engine.run_rng(d, n, 17, 1000)[0]
    part, nxt = studies.tail_statistics(engine, d, n, 17, trial0=1000, chunk=chunk, checkpoint=ck, stop_after_chunks=2)
    assert nxt == 1000 + 2 * chunk and part[0] + part[1] == 2 * chunk
    full, nxt = studies.tail_statistics(engine, d, n, 17, trial0=1000, chunk=chunk, checkpoint=ck)        # resumes
    assert nxt == 1000 + n
    assert np.array_equal(full[:2], ref[:2]) and np.array_equal(full[4:], ref[4:])
    assert np.allclose(full[2:4], ref[2:4], rtol=1e-12)
    again, _ = studies.tail_statistics(engine, d, n, 17, trial0=1000, chunk=chunk, checkpoint=ck)         # nothing left to do
    assert np.array_equal(again, full)
    with pytest.raises(ValueError):
        studies.tail_statistics(engine, d, n, 18, trial0=1000, chunk=chunk, checkpoint=ck)
