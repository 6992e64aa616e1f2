"""Bulk study drivers on the CUDA engine — the callers either side of the trial.

* `monte_carlo_converged`  — the stop rule of MonteCarlo.run_sim (/root/reference/monte_carlo.py:48-76) evaluated on
  streaming moments: no DataFrame, no O(n^2) pd.concat.
* `sensitivity_grid`       — BASELINE config 3: translation x rotation x centroid-sigma Cartesian grid, one segment per
  grid point (a generalisation of Sense.run_sim's 1-D diagonal sweep, sensitivity.py:39-55).
* `leo_lifetime`           — BASELINE config 4: one segment per orbit step with thermal / radiation parameter overrides
  (thermal focal update sensitivity.py:115-118; radiation -> image noise -> detection threshold README.md:12).

* `result_frame` / `write_result_pickle` — the result files of driver.py:337-339 (`data/<name>.pkl`: a DataFrame of the
  sampled parameter columns + the objective column) from a bulk GPU run, so `toolplot.py -s m/s/c` reads GPU output
  unchanged.

Nothing numerical about the trial happens here: these build `stmpm_dist` arrays and call `TrialEngine.run_rng*`.
"""
from __future__ import annotations

from typing import Mapping, Sequence

import numpy as np

from . import _abi
from .engine import NORMAL_PARAMS, TrialEngine, make_dist, summarize

HALF_NORMAL_K = np.sqrt(1.0 - 2.0 / np.pi)     # driver.py:290


def _merge(total: np.ndarray, part: np.ndarray) -> None:
    xmax = max(total[5], part[5])
    total += part
    total[5] = xmax


def monte_carlo_converged(engine: TrialEngine, means: Mapping[str, float], sigmas: Mapping[str, float],
                          thermal_coef: float = 0.0, *, batch: int = 100, seed: int = 0, max_runs: int = -1,
                          min_rows: int = 10_000, tol: float = 1e-5, precision: int = 64,
                          batches_per_launch: int = 1, on_device: bool = True) -> dict:
    """MonteCarlo.run_sim's loop (monte_carlo.py:48-76): batches of `batch` trials until the kept rows number at least
    `min_rows` AND the relative change of the half-normal sigma is <= tol, or kept rows >= max_runs > 0.
    sigma_hat = std(x, ddof=1) / sqrt(1 - 2/pi) over all successful trials so far (monte_carlo.py:60-61).

    on_device=True (default): `stmpm_mc_converge` — the rule is evaluated on the GPU after every batch, the host reads the
    state once.  on_device=False: round 1's host loop, one launch + synchronisation per
    `batches_per_launch` batches on merged packed statistics (kept as the cross-check; the reference's rule is
    batches_per_launch = 1)."""
    dist = make_dist(means, sigmas, thermal_coef)
    if on_device:
        if batches_per_launch != 1:
            raise ValueError("batches_per_launch applies to the host loop only (on_device=False)")
        r = engine.mc_converge(dist, batch=batch, seed=seed, min_rows=min_rows, tol=tol, max_runs=max_runs, precision=precision)
        out = summarize(r["stats"], engine.hist_log2_sub)
        out.update(count=r["count"], trials=r["trials"], std_ratio=r["std_ratio"], stats=r["stats"],
                   sigma_halfnormal=r["sigma_halfnormal"], mean_halfnormal=r["mean_halfnormal"],
                   converged=bool(r["converged"]), rounds=r["rounds"], trials_computed=r["trials_computed"])
        return out
    total = np.zeros(engine.stats_len)
    prev, ratio, count, trial0 = 1.0, 1.0, 0, 0                      # monte_carlo.py:45-46
    per_launch = int(batch) * int(batches_per_launch)
    while ratio > tol or total[0] < min_rows:
        count += batches_per_launch
        part = engine.run_rng(dist, per_launch, seed, trial0, precision, )[0]
        trial0 += per_launch
        _merge(total, part)
        n = total[0]
        std = np.sqrt(max(0.0, (total[3] - total[2] ** 2 / n) / (n - 1))) if n > 1 else 0.0
        sigma_hat = std / HALF_NORMAL_K
        ratio = abs((prev - sigma_hat) / prev)
        prev = sigma_hat
        if max_runs > 0 and n >= max_runs:
            break
    out = summarize(total, engine.hist_log2_sub)
    out.update(count=count, trials=trial0, std_ratio=ratio, stats=total)
    return out


def grid_axes(sig_eps: float, sig_phi: float, sig_dev: float, n_eps: int = 64, n_phi: int = 64, n_dev: int = 16):
    """Config 3 axes: eps scale, phi scale, centroid sigma, each linspace(0, 3 sigma, n)."""
    return (np.linspace(0.0, 3.0 * sig_eps, n_eps), np.linspace(0.0, 3.0 * sig_phi, n_phi),
            np.linspace(0.0, 3.0 * sig_dev, n_dev))


def sensitivity_grid_dists(means: Mapping[str, float], sigmas: Mapping[str, float], thermal_coef: float,
                           eps_axis: Sequence[float], phi_axis: Sequence[float], dev_axis: Sequence[float]):
    """One stmpm_dist per grid point, C order (eps, phi, dev).  At a grid point the swept quantities are FIXED (sigma 0),
    like the reference overwrites swept columns with the linspace values (sensitivity.py:54-55): axis 1 sets
    F_ARR_EPS_X/Y/Z jointly, axis 2 F_ARR_PHI/THETA/PSI jointly, axis 3 BASE_DEV_X = BASE_DEV_Y."""
    dists = []
    for e in eps_axis:
        for p in phi_axis:
            for d in dev_axis:
                m, s = dict(means), dict(sigmas)
                for k in ("F_ARR_EPS_X", "F_ARR_EPS_Y", "F_ARR_EPS_Z"):
                    m[k], s[k] = float(e), 0.0
                for k in ("F_ARR_PHI", "F_ARR_THETA", "F_ARR_PSI"):
                    m[k], s[k] = float(p), 0.0
                for k in ("BASE_DEV_X", "BASE_DEV_Y"):
                    m[k], s[k] = float(d), 0.0
                dists.append(make_dist(m, s, thermal_coef))
    return dists


def sensitivity_grid(engine: TrialEngine, means, sigmas, thermal_coef, eps_axis, phi_axis, dev_axis,
                     trials_per_point: int, seed: int = 3, precision: int = 64, segments=None) -> np.ndarray:
    """Runs the grid (or the subset `segments` of flat point indices, e.g. this rank's share from
    sharding.shard_segments) and returns packed statistics [n_points_run, stats_len].  Global trial index of trial i of
    point p is p * trials_per_point + i regardless of which rank runs it."""
    dists = sensitivity_grid_dists(means, sigmas, thermal_coef, eps_axis, phi_axis, dev_axis)
    return run_segments(engine, dists, trials_per_point, seed, precision, segments)


# /root/reference/constants.py:99-108 — the thermal-balance constants the reference carries (its balance code is absent)
SOLARFLUX, EARTHFLUX, STEFBOLTZ, C1U, C3U, TEMP_GAMMA = 1367.0, 213.0, 5.670367e-8, 0.01, 0.03, 0.273


def radiative_equilibrium(form_factor: str = "1U", sunlit: bool = True) -> float:
    """[FREE, SURVEY.md §8f rank 4] Radiative-equilibrium temperature [K] of a CubeSat from the reference's constants:
    absorbed = TEMP_GAMMA (= alpha / epsilon) x SOLARFLUX on one face (sunlit arcs only) + EARTHFLUX (Earth IR) on one face;
    emitted = STEFBOLTZ T^4 over all faces.  Face areas C1U (10 cm x 10 cm) / C3U (10 cm x 30 cm)."""
    if form_factor == "1U":
        a_sun = a_earth = C1U
        a_tot = 6.0 * C1U
    elif form_factor == "3U":
        a_sun = a_earth = C3U
        a_tot = 2.0 * C1U + 4.0 * C3U
    else:
        raise ValueError("form_factor must be '1U' or '3U'")
    q = (TEMP_GAMMA * SOLARFLUX * a_sun if sunlit else 0.0) + EARTHFLUX * a_earth
    return float((q / (STEFBOLTZ * a_tot)) ** 0.25)


def orbit_dtemp_amplitude(form_factor: str = "1U", damping: float = 0.65) -> float:
    """Half the sunlit / eclipse equilibrium swing, damped by the structure's thermal inertia (fraction of the swing the
    optics actually follow within one ~90 min orbit): the D_TEMP amplitude [K] `leo_lifetime_dists` defaults to (~15 K)."""
    return 0.5 * damping * (radiative_equilibrium(form_factor, True) - radiative_equilibrium(form_factor, False))


def leo_lifetime_dists(means: Mapping[str, float], sigmas: Mapping[str, float], thermal_coef: float, n_steps: int = 60,
                       years: float = 5.0, dtemp_amp: float | None = None, dtemp_sigma: float = 5.0,
                       mag_loss_per_year: float = 0.12, dev_growth_per_year: float = 0.15):
    """Config 4 schedule [FREE, SURVEY.md §8d]: monthly steps over `years`.  Thermal: D_TEMP mean follows a seasonal
    sinusoid of amplitude dtemp_amp K (beta-angle cycle), sigma dtemp_sigma.  Radiation: accumulated dose raises the image
    noise floor, i.e. lowers the detection threshold MAX_MAGNITUDE linearly by mag_loss_per_year and inflates the
    centroiding sigma by dev_growth_per_year (fractional).  dtemp_amp defaults to `orbit_dtemp_amplitude()` — the
    radiative balance on the reference's own constants (constants.py:99-108), not a hard-coded number."""
    if dtemp_amp is None:
        dtemp_amp = orbit_dtemp_amplitude()
    dists, sched = [], []
    for k in range(n_steps):
        t = years * k / max(1, n_steps - 1)
        m, s = dict(means), dict(sigmas)
        m["D_TEMP"] = dtemp_amp * np.sin(2.0 * np.pi * t)
        s["D_TEMP"] = dtemp_sigma
        m["MAX_MAGNITUDE"] = means["MAX_MAGNITUDE"] - mag_loss_per_year * t
        grow = 1.0 + dev_growth_per_year * t
        s["BASE_DEV_X"] = sigmas.get("BASE_DEV_X", 0.0) * grow
        s["BASE_DEV_Y"] = sigmas.get("BASE_DEV_Y", 0.0) * grow
        dists.append(make_dist(m, s, thermal_coef))
        sched.append({"t_years": t, "D_TEMP": m["D_TEMP"], "MAX_MAGNITUDE": m["MAX_MAGNITUDE"], "dev_sigma": s["BASE_DEV_X"]})
    return dists, sched


def run_segments(engine: TrialEngine, dists: Sequence[_abi.Dist], trials_per_segment: int, seed: int,
                 precision: int = 64, segments=None, max_segments_per_launch: int = 4096) -> np.ndarray:
    """Statistics-only multi-segment run; `segments` selects flat indices (default all).  Consecutive selected indices
    are batched into single launches (segment s always uses global trial indices s * trials_per_segment + i)."""
    idx = list(range(len(dists))) if segments is None else [int(i) for i in segments]
    out = np.zeros((len(idx), engine.stats_len))
    i = 0
    while i < len(idx):
        j = i + 1                                                  # grow a run of consecutive segment indices
        while j < len(idx) and idx[j] == idx[j - 1] + 1 and j - i < max_segments_per_launch:
            j += 1
        out[i:j] = engine.run_rng([dists[k] for k in idx[i:j]], trials_per_segment, seed,
                                  idx[i] * int(trials_per_segment), precision)
        i = j
    return out


# ---------------------------------------------------------------- run.sh study matrix
# study name -> swept / perturbed columns, exactly run.sh:11-42's tag table
STUDY_TAGS = {
    "STARS": ("MAX_MAGNITUDE",),                                             # run.sh:16
    "EPS_LAT": ("F_ARR_EPS_X", "F_ARR_EPS_Y"),                               # run.sh:12
    "EPS_LONG": ("F_ARR_EPS_Z",),                                            # run.sh:14
    "EPS": ("F_ARR_EPS_X", "F_ARR_EPS_Y", "F_ARR_EPS_Z"),                    # run.sh:35
    "PHI_LAT": ("F_ARR_PHI", "F_ARR_THETA"),                                 # run.sh:11
    "PHI_LONG": ("F_ARR_PSI",),                                              # run.sh:13
    "PHI": ("F_ARR_PHI", "F_ARR_THETA", "F_ARR_PSI"),                        # run.sh:38
    "FOCLEN": ("FOCAL_LENGTH", "MAX_MAGNITUDE"),                             # run.sh:15
    "CENTROID": ("BASE_DEV_X", "BASE_DEV_Y"),                                # run.sh:17
}
SWEEP_POINTS, SWEEP_REPEATS = 100, 200                                        # sensitivity.py:19-20


def study_matrix(engine: TrialEngine, means: Mapping[str, float], sigmas: Mapping[str, float], *, studies=None,
                 mc_trials: int = 20_000, sweep_repeats: int = SWEEP_REPEATS, seed: int = 0, precision: int = 64) -> dict:
    """The thesis' "univariate effect analysis" (run.sh:50-131) on the GPU engine, per study name:

    * `mc`    — Monte Carlo with ONLY the study's parameters perturbed (their sigma from `sigmas`, everything else ideal):
                what run.sh does with `-cam data/<STUDY>.json -sw i -n 20000` (run.sh:61-63,120).
    * `sweep` — Sense.run_sim's 1-D diagonal sweep (sensitivity.py:39-55): 100 values, every swept column at the same
                index: linspace(ideal - 3 sigma, ideal + 3 sigma); BASE_DEV_* linspace(0, -3 sigma) (sensitivity.py:46-47);
                MAX_MAGNITUDE linspace(ideal - 3, ideal + 3) (sensitivity.py:49-52), or ~ N(ideal, 1) per trial when
                swept together with FOCAL_LENGTH (sensitivity.py:61-62); all other hardware ideal and, unless swept,
                MAX_MAGNITUDE = 20 (sensitivity.py:112).  `sweep_repeats` trials per value (200 in the reference).
    One multi-segment launch per sweep.  Returns {study: {"mc": summary, "sweep": {"values": {...}, "mean": [...], ...}}}."""
    out = {}
    ideal_sig = {k: 0.0 for k in NORMAL_PARAMS}
    for si, name in enumerate(studies or STUDY_TAGS):
        tags = STUDY_TAGS[name]
        s_mc = dict(ideal_sig)
        for t in tags:
            s_mc[t] = float(sigmas.get(t, 0.0))
        mc = summarize(engine.run_rng(make_dist(means, s_mc, 0.0), mc_trials, seed + 2 * si, 0, precision)[0],
                       engine.hist_log2_sub)
        values, dists = {}, []
        for t in tags:
            up = 3.0 * float(sigmas.get(t, 0.0))
            if t in ("BASE_DEV_X", "BASE_DEV_Y"):
                values[t] = np.linspace(0.0, -up, SWEEP_POINTS)
            elif t == "MAX_MAGNITUDE":
                values[t] = np.linspace(means[t] - 3.0, means[t] + 3.0, SWEEP_POINTS)
            else:
                values[t] = np.linspace(means[t] - up, means[t] + up, SWEEP_POINTS)
        both = "FOCAL_LENGTH" in tags and "MAX_MAGNITUDE" in tags
        for i in range(SWEEP_POINTS):
            m, s = dict(means), dict(ideal_sig)
            m["MAX_MAGNITUDE"] = 20.0
            for t in tags:
                m[t] = float(values[t][i])
            if both:
                m["MAX_MAGNITUDE"], s["MAX_MAGNITUDE"] = means["MAX_MAGNITUDE"], 1.0
            dists.append(make_dist(m, s, 0.0))
        st = engine.run_rng(dists, sweep_repeats, seed + 2 * si + 1, 0, precision)
        ss = summaries(st, engine.hist_log2_sub)
        out[name] = {"mc": mc, "sweep": {"values": values, "mean": np.array([x.get("mean", np.nan) for x in ss]),
                                         "sigma_halfnormal": np.array([x.get("sigma_halfnormal", np.nan) for x in ss]),
                                         "fail_rate": np.array([x["fail_rate"] for x in ss]), "stats": st}}
    return out


QUEST_OUTPUT = "QUATERNION_ERROR_[arcsec]"      # toolplot.py:71
N_DET_OUTPUT = "N_STARS_DETECTED"
MAG_OUTPUT = "MAXIMUM MAGNITUDE VISIBLE"            # toolplot.py:241


def result_frame(engine, means: Mapping[str, float] | None = None, sigmas: Mapping[str, float] | None = None,
                 thermal_coef: float = 0.0, n: int = 0, seed: int = 0, *, trial0: int = 0, precision: int = 64,
                 keep_failed: bool = False, star_mag: bool = False, cols: np.ndarray | None = None):
    """The DataFrame driver.py accumulates and pickles (driver.py:286-339), from one bulk run.

    Columns: the 13 parameter columns under the reference's names, `QUATERNION_ERROR_[arcsec]` (toolplot.py:71),
    `N_STARS_DETECTED`, and with star_mag=True `MAXIMUM MAGNITUDE VISIBLE` (toolplot.py:241, objective `-func M`).
    Parameters are drawn on the device by the seeded generator (`fill_presampled_device`: exactly native-RNG mode's
    draws) unless pre-sampled `cols` [13, n] are given.  Failed trials (fewer than n_min detections) are dropped, as
    Simulation.run_sim drops them (the driver's row shortfall is its failure count, driver.py:321); keep_failed=True keeps
    them with the -1 sentinel, which is what `toolplot.py -s s` turns into its star-count failure rate (toolplot.py:168-180)."""
    import pandas as pd
    if cols is None:
        import torch
        dist = make_dist(means, sigmas, thermal_coef)
        dev = torch.empty((len(_abi.COLUMNS), int(n)), dtype=torch.float64, device=f"cuda:{engine.device}")
        engine.fill_presampled_device(dev, dist, seed=seed, trial0=trial0)
        torch.cuda.synchronize()
        cols = dev.cpu().numpy()
    cols = np.ascontiguousarray(cols, dtype=np.float64)
    out = engine.run_presampled(cols, seed, trial0, precision, want_maxmag=star_mag)
    err, ndet = out[0], out[1]
    df = pd.DataFrame({name: cols[i] for i, name in enumerate(_abi.COLUMNS)})
    df[QUEST_OUTPUT] = err
    df[N_DET_OUTPUT] = ndet
    if star_mag:
        df[MAG_OUTPUT] = out[2]
    return df if keep_failed else df[err >= 0].reset_index(drop=True)


def write_result_pickle(df, name: str, data_dir: str = "data") -> str:
    """driver.py:337-339: `pd.to_pickle(fin_df, <repo>/data/<name>.pkl)` — the file `toolplot.py -f data/<name>.pkl` reads."""
    import os

    import pandas as pd
    os.makedirs(data_dir, exist_ok=True)
    path = os.path.join(data_dir, name + ".pkl")
    pd.to_pickle(df, path)
    return path


def tail_statistics(engine: TrialEngine, dist: _abi.Dist, n_trials: int, seed: int, *, trial0: int = 0,
                    chunk: int = 1 << 30, checkpoint: str | None = None, precision: int = 64, stop_after_chunks: int = 0):
    """Config-5-style statistics-only run of `n_trials` global trial indices [trial0, trial0 + n_trials) in chunks, with
    checkpoint / resume (trial i is a pure function of (seed, i), so a checkpoint is just
    (seed, next trial index, packed statistics) — 70 KB).  After every chunk the state is written atomically to
    `checkpoint` (.npz); a later call with the same arguments resumes after the last finished chunk and ends with exactly
    the statistics an uninterrupted run produces (integer entries bit-identical; the two float sums to rounding order).
    `stop_after_chunks` > 0 returns early (tests / time-boxed jobs).  Returns (packed stats, next trial index)."""
    import os
    total = np.zeros(engine.stats_len)
    nxt = int(trial0)
    if checkpoint and os.path.exists(checkpoint):
        with np.load(checkpoint) as ck:
            if int(ck["seed"]) != int(seed) & (2**64 - 1) or int(ck["trial0"]) != int(trial0) or int(ck["hist_log2_sub"]) != engine.hist_log2_sub \
                    or not np.array_equal(ck["dist"], np.frombuffer(bytes(dist), dtype=np.float64)):
                raise ValueError(f"{checkpoint} belongs to a different run (seed / trial0 / distribution / histogram differ)")
            total, nxt = ck["stats"].copy(), int(ck["next_trial"])
    end, done = int(trial0) + int(n_trials), 0
    while nxt < end:
        m = min(int(chunk), end - nxt)
        _merge(total, engine.run_rng(dist, m, seed, nxt, precision)[0])
        nxt += m
        done += 1
        if checkpoint:
            tmp = checkpoint + ".tmp.npz"
            np.savez(tmp, stats=total, next_trial=np.int64(nxt), seed=np.uint64(int(seed) & (2**64 - 1)), trial0=np.int64(trial0),
                     hist_log2_sub=np.int64(engine.hist_log2_sub), dist=np.frombuffer(bytes(dist), dtype=np.float64))
            os.replace(tmp, checkpoint)
        if stop_after_chunks and done >= stop_after_chunks:
            break
    return total, nxt


def summaries(stats: np.ndarray, hist_log2_sub: int = 8) -> list[dict]:
    return [summarize(s, hist_log2_sub) for s in np.atleast_2d(stats)]


__all__ = ["monte_carlo_converged", "grid_axes", "sensitivity_grid_dists", "sensitivity_grid", "leo_lifetime_dists",
           "run_segments", "summaries", "study_matrix", "STUDY_TAGS", "NORMAL_PARAMS", "result_frame",
           "write_result_pickle", "tail_statistics", "QUEST_OUTPUT", "N_DET_OUTPUT", "MAG_OUTPUT", "radiative_equilibrium", "orbit_dtemp_amplitude"]
