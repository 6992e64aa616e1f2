"""A SECOND, independent CPU restatement of the trial — test infrastructure, the generator of trial_golden_v2.npz.

Written from docs/MODEL.md alone, one trial (DataFrame row) at a time, WITHOUT reading or importing `oracle/trial.py`
or anything under `startrackermpm_b200/csrc/` ("What's weak" #1: the v1 fixture was produced by the oracle
it then checked).  It shares no function with the oracle or the CUDA path:

* Philox4x32-10 / Philox2x32-10: plain Python integers, written from MODEL.md §5's constants;
* rotations: the reference's own primitives restated (/root/reference/constants.py:81-94), composed as MODEL.md §2 says;
  the star is taken to the body frame (r = C^T s) and the ray-plane intersection is done there, literally as §2 writes
  it (the CUDA path never forms r; the oracle is free to do either);
* Box-Muller: `math.log / sqrt / cos / sin` (libm) on Python floats;
* Wahba: `scipy.spatial.transform.Rotation.align_vectors` (Kabsch/SVD) end to end — neither `numpy.linalg.eigh` on
  Davenport's K (the oracle) nor a QUEST Newton iteration (the kernel); the error angle is `Rotation.magnitude()` of
  A_est C.
* star search: every catalog star, no sky grid.

Only tests/ may import this file.  `python tests/golden/make_golden_v2.py` regenerates the fixture.
"""
from __future__ import annotations

import math

import numpy as np
from scipy.spatial.transform import Rotation

M32 = 0xFFFFFFFF
ARCSEC_PER_RAD = 206264.80624709636            # MODEL.md §4
HALF_EXTENT = 512.0                             # constants.py:97-98: 1024 x 1024 px, measured from the centre
N_MIN = 3                                       # MODEL.md §3
COLUMN_ORDER = ("RIGHT_ASCENSION", "DECLINATION", "ROLL", "FOCAL_LENGTH", "F_ARR_EPS_X", "F_ARR_EPS_Y", "F_ARR_EPS_Z",
                "F_ARR_PHI", "F_ARR_THETA", "F_ARR_PSI", "BASE_DEV_X", "BASE_DEV_Y", "MAX_MAGNITUDE")   # MODEL.md §1
NORMAL_ORDER = ("FOCAL_LENGTH", "F_ARR_EPS_X", "F_ARR_EPS_Y", "F_ARR_EPS_Z", "F_ARR_PHI", "F_ARR_THETA", "F_ARR_PSI",
                "BASE_DEV_X", "BASE_DEV_Y", "MAX_MAGNITUDE", "D_TEMP")                                  # MODEL.md §6


# ------------------------------------------------------------------------------------------------ Philox (MODEL.md §5)
def philox4x32_10(ctr, key):
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    for _ in range(10):
        p0 = 0xD2511F53 * c0
        p1 = 0xCD9E8D57 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & M32, p1 & M32, ((p0 >> 32) ^ c3 ^ k1) & M32, p0 & M32
        k0 = (k0 + 0x9E3779B9) & M32
        k1 = (k1 + 0xBB67AE85) & M32
    return c0, c1, c2, c3


def philox2x32_10(ctr, key):
    c0, c1 = ctr
    k = key
    for _ in range(10):
        p = 0xD256D193 * c0
        c0, c1 = ((p >> 32) ^ k ^ c1) & M32, p & M32
        k = (k + 0x9E3779B9) & M32
    return c0, c1


def unit_open(w: int) -> float:
    return (w + 0.5) * 2.0 ** -32


def normal_pair(wa: int, wb: int):
    rad = math.sqrt(-2.0 * math.log(unit_open(wa)))
    ang = 2.0 * math.pi * unit_open(wb)
    return rad * math.cos(ang), rad * math.sin(ang)


# ------------------------------------------------------------------------------------------------ MODEL.md §6
def draw_row(mean: dict, sigma: dict, thermal_coef: float, seed: int, trial: int) -> list[float]:
    """One row of native-RNG parameter draws, in COLUMN_ORDER."""
    key = (seed & M32, (seed >> 32) & M32)
    w = []
    for j in range(4):
        w += philox4x32_10((trial & M32, (trial >> 32) & M32, j, 0x50415241), key)
    fov = math.radians(10.0)                    # monte_carlo.py:87
    ra = fov + (2.0 * math.pi - 2.0 * fov) * unit_open(w[0])
    dec = (-math.pi / 2.0 + fov) + (math.pi - 2.0 * fov) * unit_open(w[1])
    roll = -math.pi + 2.0 * math.pi * unit_open(w[2])
    z = []
    for k in range(6):
        z += normal_pair(w[4 + 2 * k], w[5 + 2 * k])
    val = {name: mean.get(name, 0.0) + sigma.get(name, 0.0) * z[i] for i, name in enumerate(NORMAL_ORDER)}
    f = val["FOCAL_LENGTH"] * (1.0 + val["D_TEMP"] * thermal_coef)      # sensitivity.py:115-117
    return [ra, dec, roll, f, val["F_ARR_EPS_X"], val["F_ARR_EPS_Y"], val["F_ARR_EPS_Z"], val["F_ARR_PHI"],
            val["F_ARR_THETA"], val["F_ARR_PSI"], val["BASE_DEV_X"], val["BASE_DEV_Y"], val["MAX_MAGNITUDE"]]


def draw_rows(mean, sigma, thermal_coef, seed, trial0, n) -> np.ndarray:
    return np.array([draw_row(mean, sigma, thermal_coef, seed, trial0 + i) for i in range(n)], dtype=np.float64).T.copy()


# ------------------------------------------------------------------------------------------------ MODEL.md §2
def rot_x(a):   # constants.py:81-84
    return np.array([[1.0, 0.0, 0.0], [0.0, math.cos(a), -math.sin(a)], [0.0, math.sin(a), math.cos(a)]])


def rot_y(a):   # constants.py:86-89
    return np.array([[math.cos(a), 0.0, math.sin(a)], [0.0, 1.0, 0.0], [-math.sin(a), 0.0, math.cos(a)]])


def rot_z(a):   # constants.py:91-94
    return np.array([[math.cos(a), -math.sin(a), 0.0], [math.sin(a), math.cos(a), 0.0], [0.0, 0.0, 1.0]])


def one_trial(row, cat_xyz: np.ndarray, cat_mag: np.ndarray, f_ideal: float, seed: int, trial: int):
    """row: the 13 values of MODEL.md §1.  Returns (error [arcsec] or -1.0, n_det, faintest detected magnitude or -inf)."""
    ra, dec, roll, f, ex, ey, ez, phi, theta, psi, devx, devy, maxmag = (float(x) for x in row)
    body_to_inertial = rot_z(ra) @ rot_y(math.pi / 2.0 - dec) @ rot_z(roll)
    plane = rot_x(math.radians(phi)) @ rot_y(math.radians(theta)) @ rot_z(math.radians(psi))
    e_u, e_v, nrm = plane[:, 0], plane[:, 1], plane[:, 2]
    centre = np.array([ex, ey, f + ez])

    # every catalog star in the body frame; geometric visibility and the magnitude test (float32 -> float64 compare)
    r_all = cat_xyz @ body_to_inertial                       # row i = (C^T s_i)^T
    facing = r_all @ nrm
    bright = cat_mag.astype(np.float64) <= maxmag
    # noise is bounded: |z| <= sqrt(-2 ln 2^-33) < 6.8, so stars farther than that from the array cannot be detected;
    # this only saves Philox evaluations for stars that fail the bounds test for every possible draw
    reach_u = HALF_EXTENT + 6.8 * abs(devx) + 1.0
    reach_v = HALF_EXTENT + 6.8 * abs(devy) + 1.0

    key4 = (seed & M32, (seed >> 32) & M32)
    ka, kb, _, _ = philox4x32_10((trial & M32, (trial >> 32) & M32, 0, 0x53544152), key4)

    body_vecs, ref_vecs, mags = [], [], []
    for i in np.nonzero((facing > 0.0) & bright)[0]:
        r = r_all[i]
        lam = float(nrm @ centre) / float(facing[i])
        hit = lam * r - centre
        u, v = float(e_u @ hit), float(e_v @ hit)
        if abs(u) > reach_u or abs(v) > reach_v:
            continue
        x0, x1 = philox2x32_10((int(i), kb), ka)             # star_id = index in the catalog file
        z1, z2 = normal_pair(x0, x1)
        um = u + abs(devx) * z1
        vm = v + abs(devy) * z2
        if abs(um) <= HALF_EXTENT and abs(vm) <= HALF_EXTENT:
            b = np.array([um, vm, f_ideal])
            body_vecs.append(b / math.sqrt(float(b @ b)))
            ref_vecs.append(cat_xyz[i])
            mags.append(float(cat_mag[i]))
    n_det = len(body_vecs)
    faintest = max(mags) if mags else -math.inf
    if n_det < N_MIN:
        return -1.0, n_det, faintest
    # Wahba with equal weights: the rotation A with b_i ~ A s_i
    est, _ = Rotation.align_vectors(np.array(body_vecs), np.array(ref_vecs))
    residual = Rotation.from_matrix(est.as_matrix() @ body_to_inertial)
    return residual.magnitude() * ARCSEC_PER_RAD, n_det, faintest


def run_rows(cols: np.ndarray, cat_xyz, cat_mag, f_ideal: float, seed: int, trial0: int):
    n = cols.shape[1]
    err, ndet, mag = np.empty(n), np.empty(n, dtype=np.int32), np.empty(n, dtype=np.float32)
    for i in range(n):
        err[i], ndet[i], mag[i] = one_trial(cols[:, i], cat_xyz, cat_mag, f_ideal, seed, trial0 + i)
    return err, ndet, mag
