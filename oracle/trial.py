"""ORACLE (test infrastructure, not product code) — numpy float64 restatement of the StarTrackerMPM trial.

PARITY UNPINNED.  The reference's own trial code (`simObjects/Simulation.py`, `AttitudeEstimation.py`, ...) is absent
from /root/reference and the reference holds no tests, golden vectors or fixtures, so
this file cannot be checked against reference outputs.  It restates docs/MODEL.md, which is constrained by the facts
the present reference files DO pin; each function cites them:

  * rotation primitives Rx/Ry/Rz ............ /root/reference/constants.py:81-94
  * sensor 1024 x 1024 px ................... /root/reference/constants.py:97-98
  * boresight sampling ranges ............... /root/reference/monte_carlo.py:87-92
  * thermal focal-length update ............. /root/reference/sensitivity.py:115-118
  * output column / arcsec .................. /root/reference/toolplot.py:71
  * failure = row dropped / -1 sentinel ..... /root/reference/driver.py:321-322, toolplot.py:170-172
  * Wahba / Davenport q-method / QUEST ...... Wahba 1965; Davenport 1968; Shuster & Oh 1981 (published algorithms)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package.
"""
from __future__ import annotations

import numpy as np

from . import philox

SENSOR_WIDTH = 1024    # constants.py:97
SENSOR_HEIGHT = 1024   # constants.py:98
HALF_U = SENSOR_WIDTH / 2.0
HALF_V = SENSOR_HEIGHT / 2.0
N_MIN = 3
RAD2ARCSEC = 206264.80624709636
FOV_EXCL = np.deg2rad(10)  # monte_carlo.py:87

COLUMNS = ("RIGHT_ASCENSION", "DECLINATION", "ROLL", "FOCAL_LENGTH",
           "F_ARR_EPS_X", "F_ARR_EPS_Y", "F_ARR_EPS_Z",
           "F_ARR_PHI", "F_ARR_THETA", "F_ARR_PSI",
           "BASE_DEV_X", "BASE_DEV_Y", "MAX_MAGNITUDE")
# order of the normal-distributed parameters in native-RNG mode (MODEL.md §6)
NORMAL_PARAMS = ("FOCAL_LENGTH", "F_ARR_EPS_X", "F_ARR_EPS_Y", "F_ARR_EPS_Z", "F_ARR_PHI", "F_ARR_THETA",
                 "F_ARR_PSI", "BASE_DEV_X", "BASE_DEV_Y", "MAX_MAGNITUDE", "D_TEMP")


# ---------------------------------------------------------------- rotations (constants.py:81-94), batched over trials
def Rx(a):
    a = np.asarray(a, dtype=np.float64)
    c, s, o, z = np.cos(a), np.sin(a), np.ones_like(a), np.zeros_like(a)
    return np.stack([np.stack([o, z, z], -1), np.stack([z, c, -s], -1), np.stack([z, s, c], -1)], -2)


def Ry(a):
    a = np.asarray(a, dtype=np.float64)
    c, s, o, z = np.cos(a), np.sin(a), np.ones_like(a), np.zeros_like(a)
    return np.stack([np.stack([c, z, s], -1), np.stack([z, o, z], -1), np.stack([-s, z, c], -1)], -2)


def Rz(a):
    a = np.asarray(a, dtype=np.float64)
    c, s, o, z = np.cos(a), np.sin(a), np.ones_like(a), np.zeros_like(a)
    return np.stack([np.stack([c, -s, z], -1), np.stack([s, c, z], -1), np.stack([z, z, o], -1)], -2)


def truth_attitude(ra, dec, roll):
    """Body->inertial C = Rz(ra) Ry(pi/2 - dec) Rz(roll)  (MODEL.md §2). [n,3,3]"""
    return Rz(ra) @ Ry(0.5 * np.pi - np.asarray(dec, dtype=np.float64)) @ Rz(roll)


def plane_rotation(phi_deg, theta_deg, psi_deg):
    """R_p = Rx(phi) Ry(theta) Rz(psi), angles in degrees (units: toolplot.py:53-55). [n,3,3]"""
    return Rx(np.deg2rad(phi_deg)) @ Ry(np.deg2rad(theta_deg)) @ Rz(np.deg2rad(psi_deg))


# ---------------------------------------------------------------- Wahba solvers
def davenport_K(B):
    """K(B) with quaternion ordered (vector, scalar); A(q) maps reference -> body."""
    S = B + np.swapaxes(B, -1, -2)
    sigma = np.trace(B, axis1=-2, axis2=-1)
    Z = np.stack([B[..., 1, 2] - B[..., 2, 1], B[..., 2, 0] - B[..., 0, 2], B[..., 0, 1] - B[..., 1, 0]], -1)
    K = np.zeros(B.shape[:-2] + (4, 4))
    K[..., :3, :3] = S - sigma[..., None, None] * np.eye(3)
    K[..., :3, 3] = Z
    K[..., 3, :3] = Z
    K[..., 3, 3] = sigma
    return K


def q_method(B):
    """Davenport q-method: eigenvector of the largest eigenvalue of K(B). Returns q [...,4] (x,y,z,w), lam_max."""
    w, v = np.linalg.eigh(davenport_K(B))
    return v[..., :, -1], w[..., -1]


def quest(B, lam0=1.0, iters=8):
    """Shuster & Oh 1981 QUEST: Newton on the quartic characteristic equation from lam0 = sum of weights."""
    S = B + np.swapaxes(B, -1, -2)
    sigma = np.trace(B, axis1=-2, axis2=-1)
    Z = np.stack([B[..., 1, 2] - B[..., 2, 1], B[..., 2, 0] - B[..., 0, 2], B[..., 0, 1] - B[..., 1, 0]], -1)
    S2 = S @ S
    kappa = 0.5 * (np.trace(S, axis1=-2, axis2=-1) ** 2 - np.trace(S2, axis1=-2, axis2=-1))  # tr adj S
    delta = np.linalg.det(S)
    SZ = np.einsum("...ij,...j->...i", S, Z)
    S2Z = np.einsum("...ij,...j->...i", S2, Z)
    a = sigma ** 2 - kappa
    b = sigma ** 2 + np.einsum("...i,...i", Z, Z)
    c = delta + np.einsum("...i,...i", Z, SZ)
    d = np.einsum("...i,...i", Z, S2Z)
    lam = np.full_like(sigma, lam0)
    for _ in range(iters):
        f = lam ** 4 - (a + b) * lam ** 2 - c * lam + (a * b + c * sigma - d)
        fp = 4 * lam ** 3 - 2 * (a + b) * lam - c
        lam = lam - f / fp
    alpha = lam ** 2 - sigma ** 2 + kappa
    beta = lam - sigma
    gamma = (lam + sigma) * alpha - delta
    x = alpha[..., None] * Z + beta[..., None] * SZ + S2Z
    q = np.concatenate([x, gamma[..., None]], -1)
    return q / np.linalg.norm(q, axis=-1, keepdims=True), lam


def quat_to_dcm(q):
    """A(q) = (w^2 - |v|^2) I + 2 v v^T - 2 w [v x]   (reference -> body)."""
    v, w = q[..., :3], q[..., 3]
    vx = np.zeros(q.shape[:-1] + (3, 3))
    vx[..., 0, 1], vx[..., 0, 2] = -v[..., 2], v[..., 1]
    vx[..., 1, 0], vx[..., 1, 2] = v[..., 2], -v[..., 0]
    vx[..., 2, 0], vx[..., 2, 1] = -v[..., 1], v[..., 0]
    vv = v[..., :, None] * v[..., None, :]
    s = (w * w - np.einsum("...i,...i", v, v))[..., None, None]
    return s * np.eye(3) + 2.0 * vv - 2.0 * w[..., None, None] * vx


def rotation_angle(dA):
    """Angle of a rotation matrix, accurate for tiny angles: atan2(|vee(dA - dA^T)|/2, (tr - 1)/2)."""
    vee = np.stack([dA[..., 2, 1] - dA[..., 1, 2], dA[..., 0, 2] - dA[..., 2, 0], dA[..., 1, 0] - dA[..., 0, 1]], -1)
    return np.arctan2(0.5 * np.linalg.norm(vee, axis=-1), 0.5 * (np.trace(dA, axis1=-2, axis2=-1) - 1.0))


# ---------------------------------------------------------------- the trial
def candidate_cos_limit(cols):
    """Generous per-trial cone (cosine of its half-angle) that contains every star that can land on the sensor.
    Pure speed-up of the brute-force search: detection below never depends on it (tests check against brute=True)."""
    f = cols[3]
    eps_lat = np.hypot(cols[4], cols[5])
    tilt = np.deg2rad(np.abs(cols[7]) + np.abs(cols[8]) + np.abs(cols[9]))
    noise = 40.0 * np.maximum(np.abs(cols[10]), np.abs(cols[11]))
    reach = np.sqrt(2.0) * (max(HALF_U, HALF_V) + noise) + eps_lat
    depth = f + cols[6] - reach * np.sin(np.minimum(tilt, 0.5))
    ang = np.arctan2(reach, np.maximum(depth, 1.0)) + tilt + np.deg2rad(1.0)
    return np.cos(np.minimum(ang, np.pi))


def run_trials(cols, cat_xyz, cat_mag, f_ideal, seed, trial0=0, *, brute=False, chunk=512, n_min=N_MIN,
               solver="q_method", return_aux=False, star_ids=None):
    """The hot path (reference: Simulation.run_sim + quest_objective, called at monte_carlo.py:52 / sensitivity.py:77;
    bodies absent) per docs/MODEL.md §2-§5.

    cols: float64 [13, n] in COLUMNS order (pre-sampled rows).  Returns (err_arcsec f64 [n] with -1 = failed,
    n_det int32 [n]).  Trial i uses global trial index trial0 + i for its Philox stream; star_ids (default: row index)
    are the catalog-file indices that key each star's noise, independent of the order of cat_xyz's rows.
    """
    cols = np.ascontiguousarray(np.asarray(cols, dtype=np.float64))
    assert cols.shape[0] == 13
    n = cols.shape[1]
    cat_xyz = np.asarray(cat_xyz, dtype=np.float64)
    cat_mag64 = np.asarray(cat_mag, dtype=np.float32).astype(np.float64)
    err = np.empty(n, dtype=np.float64)
    ndet = np.empty(n, dtype=np.int32)
    aux = {"n_cand": 0, "lam": np.empty(n), "maxmag": np.empty(n, dtype=np.float32)} if return_aux else None
    star_ids = np.arange(cat_xyz.shape[0], dtype=np.uint64) if star_ids is None else np.asarray(star_ids, np.uint64)
    for s0 in range(0, n, chunk):
        sl = slice(s0, min(n, s0 + chunk))
        e, d, a = _run_chunk(cols[:, sl], cat_xyz, cat_mag64, star_ids, float(f_ideal), seed,
                             np.uint64(trial0) + np.arange(sl.start, sl.stop, dtype=np.uint64), brute, n_min, solver)
        err[sl], ndet[sl] = e, d
        if return_aux:
            aux["n_cand"] += a["n_cand"]
            aux["lam"][sl] = a["lam"]
            aux["maxmag"][sl] = a["maxmag"]
    return (err, ndet, aux) if return_aux else (err, ndet)


def _run_chunk(cols, cat_xyz, cat_mag64, star_ids, f_ideal, seed, trial_idx, brute, n_min, solver):
    m = cols.shape[1]
    ra, dec, roll, f, ex, ey, ez, phi, theta, psi, devx, devy, maxmag = cols
    C = truth_attitude(ra, dec, roll)                      # [m,3,3] body->inertial
    Rp = plane_rotation(phi, theta, psi)                   # [m,3,3]
    e_u, e_v, nrm = Rp[:, :, 0], Rp[:, :, 1], Rp[:, :, 2]
    p0 = np.stack([ex, ey, f + ez], -1)

    bore = C[:, :, 2]
    dots = bore @ cat_xyz.T                                # [m,S]
    if brute:
        cand = np.ones_like(dots, dtype=bool)
    else:
        cand = dots >= candidate_cos_limit(cols)[:, None]
    ti, si = np.nonzero(cand)                              # flat (trial, star) pairs, trial-major
    s = cat_xyz[si]                                        # inertial reference vectors
    r = np.einsum("pji,pj->pi", C[ti], s)                  # r = C^T s
    n_dot_r = np.einsum("pi,pi->p", nrm[ti], r)
    vis = n_dot_r > 0.0
    lam = np.einsum("pi,pi->p", nrm[ti], p0[ti]) / np.where(vis, n_dot_r, 1.0)
    P = lam[:, None] * r - p0[ti]
    u = np.einsum("pi,pi->p", e_u[ti], P)
    v = np.einsum("pi,pi->p", e_v[ti], P)

    x0, x1 = philox.star_words(seed, trial_idx[ti], star_ids[si])
    z1, z2 = philox.box_muller(x0, x1)
    um = u + np.abs(devx[ti]) * z1
    vm = v + np.abs(devy[ti]) * z2

    det = vis & (cat_mag64[si] <= maxmag[ti]) & (np.abs(um) <= HALF_U) & (np.abs(vm) <= HALF_V)
    ndet = np.bincount(ti[det], minlength=m).astype(np.int32)

    td, sd = ti[det], s[det]
    bh = np.stack([um[det], vm[det], np.full(td.shape, f_ideal)], -1)
    bh /= np.linalg.norm(bh, axis=-1, keepdims=True)
    w = 1.0 / np.maximum(ndet, 1)[td]
    outer = (w[:, None, None] * bh[:, :, None]) * sd[:, None, :]   # w b s^T
    B = np.zeros((m, 3, 3))
    np.add.at(B, td, outer)

    ok = ndet >= n_min
    Bok = np.where(ok[:, None, None], B, np.eye(3))
    if solver == "quest_body":
        # the CUDA path's formulation: references r_i = C^T s_i  =>  B' = B C, optimum A' = A_est C (near identity,
        # far from QUEST's 180-degree singularity); error = rotation angle of A' = 2 atan2(|q_v|, |q_w|)
        q, lam_max = quest(Bok @ C)
        theta_err = 2.0 * np.arctan2(np.linalg.norm(q[:, :3], axis=-1), np.abs(q[:, 3]))
    else:
        q, lam_max = q_method(Bok) if solver == "q_method" else quest(Bok)
        dA = quat_to_dcm(q) @ C                            # A_est * A_true^T,  A_true = C^T
        theta_err = rotation_angle(dA)
    err = np.where(ok, theta_err * RAD2ARCSEC, -1.0)
    # star_mag objective (driver.py:246-247; output 'MAXIMUM MAGNITUDE VISIBLE', toolplot.py:241): catalog magnitude of
    # the faintest detected star, -inf when nothing is detected
    maxmag = np.full(m, -np.inf, dtype=np.float32)
    np.maximum.at(maxmag, td, cat_mag64[si[det]].astype(np.float32))
    return err, ndet, {"n_cand": int(cand.sum()), "lam": lam_max, "maxmag": maxmag}


# ---------------------------------------------------------------- native-RNG parameter sampling (MODEL.md §6)
def sample_params(means, sigmas, thermal_coef, seed, trial0, n):
    """means/sigmas: dict or sequence in NORMAL_PARAMS order (11 entries). Returns cols float64 [13, n]."""
    if isinstance(means, dict):
        means = [means[k] for k in NORMAL_PARAMS]
    if isinstance(sigmas, dict):
        sigmas = [sigmas[k] for k in NORMAL_PARAMS]
    k0, k1 = philox.split_seed(seed)
    idx = np.uint64(trial0) + np.arange(n, dtype=np.uint64)
    tl, th = philox.split_index(idx)
    w = []
    for j in range(4):
        w.extend(philox.philox4x32_10(tl, th, j, philox.TAG_PARA, k0, k1))
    fov = FOV_EXCL
    ra = fov + (2.0 * np.pi - 2.0 * fov) * philox.u01(w[0])          # monte_carlo.py:90
    dec = (-0.5 * np.pi + fov) + (np.pi - 2.0 * fov) * philox.u01(w[1])  # monte_carlo.py:91
    roll = -np.pi + 2.0 * np.pi * philox.u01(w[2])                   # monte_carlo.py:92
    z = []
    for k in range(6):
        za, zb = philox.box_muller(w[4 + 2 * k], w[5 + 2 * k])
        z += [za, zb]
    p = [means[i] + sigmas[i] * z[i] for i in range(11)]
    f = p[0] * (1.0 + p[10] * thermal_coef)                          # sensitivity.py:115-117
    return np.stack([ra, dec, roll, f] + p[1:10], 0)
