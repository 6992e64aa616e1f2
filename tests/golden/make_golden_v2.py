"""Generates tests/golden/trial_golden_v2.npz — run from the repo root: python tests/golden/make_golden_v2.py

The fixture's inputs AND outputs come from tests/golden/independent_trial.py: a scalar, one-row-at-a-time restatement of
docs/MODEL.md that shares no code with `oracle/` or the CUDA path (own Philox, libm Box-Muller, scipy's SVD-based
`align_vectors` for the Wahba solve).  Both the numpy oracle and the CUDA path must reproduce it
(tests/test_oracle_trial.py, tests/test_gpu_parity.py).  PARITY UNPINNED still holds with respect to the REFERENCE (its
trial code and fixtures are absent) — what this fixture adds is that the oracle is no longer checked
against its own output.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
import independent_trial as it  # noqa: E402

SEED = 0x5EED00020231105          # a 64-bit seed: exercises the high key word
F_IDEAL = 3500.0
TRIAL0 = (1 << 32) - 100          # global trial indices cross 2^32: exercises the high counter word
MEAN = dict(FOCAL_LENGTH=3500.0, MAX_MAGNITUDE=6.0)
SIGMA = dict(FOCAL_LENGTH=2.0, F_ARR_EPS_X=1.0, F_ARR_EPS_Y=1.0, F_ARR_EPS_Z=1.0, F_ARR_PHI=0.05, F_ARR_THETA=0.05,
             F_ARR_PSI=0.05, BASE_DEV_X=0.1, BASE_DEV_Y=0.1, MAX_MAGNITUDE=0.25, D_TEMP=5.0)


def scenarios():
    """(name, n, mean, sigma, thermal_coef[, column overrides])"""
    zero = {k: 0.0 for k in SIGMA}
    return [
        ("basic", 192, MEAN, SIGMA, 2e-5, None),
        ("ideal", 32, MEAN, zero, 0.0, None),
        ("all_visible", 32, dict(MEAN, MAX_MAGNITUDE=20.0), dict(SIGMA, MAX_MAGNITUDE=0.0), 0.0, None),   # sensitivity.py:112
        ("few_stars", 48, dict(MEAN, MAX_MAGNITUDE=4.0), SIGMA, 0.0, None),
        ("wide_fov", 24, dict(MEAN, FOCAL_LENGTH=1500.0), SIGMA, 0.0, None),
        ("sigma_x5", 48, MEAN, {k: 5.0 * v for k, v in SIGMA.items()}, 2e-5, None),
        ("negative_dev", 16, dict(MEAN, BASE_DEV_X=-0.3, BASE_DEV_Y=-0.2), zero, 0.0, None),               # sensitivity.py:46-47
        ("poles_wrap", 24, MEAN, SIGMA, 0.0, {1: np.linspace(-np.pi / 2, np.pi / 2, 24), 0: np.linspace(-0.4, 2 * np.pi + 0.4, 24)}),
    ]


def main():
    cat = np.load(os.path.join(ROOT, "startrackermpm_b200", "utils", "catalog_synth_9110.npz"))
    xyz, mag = cat["xyz"], cat["mag"]
    cols, err, ndet, faint, names = [], [], [], [], []
    t0 = TRIAL0
    for name, n, mean, sigma, coef, override in scenarios():
        c = it.draw_rows(mean, sigma, coef, SEED, t0, n)
        if override:
            for k, v in override.items():
                c[k] = v
        e, d, m = it.run_rows(c, xyz, mag, F_IDEAL, SEED, t0)
        ok = e >= 0
        print(f"{name:13s} n={n:4d} failed={int((~ok).sum()):3d} n_det mean={d.mean():6.2f} "
              f"err mean={e[ok].mean() if ok.any() else float('nan'):.5f} arcsec")
        cols.append(c); err.append(e); ndet.append(d); faint.append(m); names += [name] * n
        t0 += n
    path = os.path.join(HERE, "trial_golden_v2.npz")
    np.savez_compressed(path, cols=np.concatenate(cols, 1), err_arcsec=np.concatenate(err), n_det=np.concatenate(ndet),
                        faintest_mag=np.concatenate(faint), scenario=np.array(names), seed=np.uint64(SEED),
                        f_ideal=np.float64(F_IDEAL), trial0=np.int64(TRIAL0))
    print("wrote", path)


if __name__ == "__main__":
    main()
