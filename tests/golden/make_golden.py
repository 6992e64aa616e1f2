"""Generates tests/golden/trial_golden_v1.npz — run from the repo root: python tests/golden/make_golden.py

The reference cannot produce golden vectors for this path (its trial code is absent from /root/reference and it ships
no fixtures: SURVEY.md §0, §4), so these are outputs of THIS repo's numpy oracle (oracle/trial.py, brute-force star
search) on seeded inputs.  PARITY UNPINNED: they pin the oracle against regressions and give the GPU tests a fixture
that travels to the GPU box; they are not reference outputs.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import trial  # noqa: E402
from startrackermpm_b200.catalog import load_catalog  # noqa: E402

SEED = 20231105
F_IDEAL = 3500.0
BASIC_MEAN = dict(FOCAL_LENGTH=3500.0, F_ARR_EPS_X=0, F_ARR_EPS_Y=0, F_ARR_EPS_Z=0, F_ARR_PHI=0, F_ARR_THETA=0,
                  F_ARR_PSI=0, BASE_DEV_X=0, BASE_DEV_Y=0, MAX_MAGNITUDE=6.0, D_TEMP=0.0)
BASIC_SIGMA = dict(FOCAL_LENGTH=2.0, F_ARR_EPS_X=1, F_ARR_EPS_Y=1, F_ARR_EPS_Z=1, F_ARR_PHI=.05, F_ARR_THETA=.05,
                   F_ARR_PSI=.05, BASE_DEV_X=.1, BASE_DEV_Y=.1, MAX_MAGNITUDE=.25, D_TEMP=5.0)


def scenarios():
    out = []
    t0 = 0
    c = trial.sample_params(BASIC_MEAN, BASIC_SIGMA, 2e-5, SEED, t0, 160); out.append(("basic", t0, c)); t0 += 160
    c = trial.sample_params(BASIC_MEAN, {k: 0.0 for k in BASIC_SIGMA}, 0.0, SEED, t0, 32); out.append(("ideal", t0, c)); t0 += 32
    m = dict(BASIC_MEAN, FOCAL_LENGTH=1200.0)
    c = trial.sample_params(m, BASIC_SIGMA, 0.0, SEED, t0, 24); out.append(("wide_fov", t0, c)); t0 += 24
    m = dict(BASIC_MEAN, MAX_MAGNITUDE=20.0)
    c = trial.sample_params(m, BASIC_SIGMA, 0.0, SEED, t0, 24); out.append(("all_visible", t0, c)); t0 += 24
    m = dict(BASIC_MEAN, MAX_MAGNITUDE=4.2)
    c = trial.sample_params(m, BASIC_SIGMA, 0.0, SEED, t0, 32); out.append(("few_stars", t0, c)); t0 += 32
    c = trial.sample_params(BASIC_MEAN, BASIC_SIGMA, 0.0, SEED, t0, 24)
    c[1] = np.linspace(-np.pi / 2, np.pi / 2, 24)          # declinations up to the poles (beyond the reference's range)
    c[0] = np.linspace(-0.3, 2 * np.pi + 0.3, 24)          # RA outside [0, 2 pi)
    out.append(("poles_wrap", t0, c)); t0 += 24
    s = {k: 5 * v for k, v in BASIC_SIGMA.items()}
    c = trial.sample_params(BASIC_MEAN, s, 2e-5, SEED, t0, 32); out.append(("large_sigma", t0, c)); t0 += 32
    return out


def main():
    xyz, mag = load_catalog()
    cols, err, ndet, names, starts = [], [], [], [], []
    for name, t0, c in scenarios():
        e, d = trial.run_trials(c, xyz, mag, F_IDEAL, SEED, t0, brute=True, chunk=16)
        cols.append(c); err.append(e); ndet.append(d); names += [name] * c.shape[1]; starts.append((name, t0, c.shape[1]))
        print(f"{name:12s} n={c.shape[1]:4d} fail={int((e < 0).sum()):3d} ndet mean={d.mean():6.2f} err mean={e[e >= 0].mean() if (e >= 0).any() else float('nan'):.4f}")
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "trial_golden_v1.npz")
    np.savez_compressed(path, cols=np.concatenate(cols, 1), err_arcsec=np.concatenate(err), n_det=np.concatenate(ndet),
                        scenario=np.array(names), seed=np.int64(SEED), f_ideal=np.float64(F_IDEAL), trial0=np.int64(0))
    print("wrote", path)


if __name__ == "__main__":
    main()
