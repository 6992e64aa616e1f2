"""Property tests (hypothesis) of the oracle's Wahba machinery, and the bench.py reference-arm JSON contract."""
import json
import os
import subprocess
import sys

import numpy as np
from hypothesis import given, settings, strategies as st
from scipy.spatial.transform import Rotation

from conftest import ROOT
from oracle import trial

angles = st.floats(-np.pi, np.pi, allow_nan=False)


def star_field(seed, n, spread):
    rng = np.random.default_rng(seed)
    v = np.stack([rng.normal(0, spread, n), rng.normal(0, spread, n), np.ones(n)], -1)
    return v / np.linalg.norm(v, axis=1, keepdims=True)


@settings(max_examples=60, deadline=None)
@given(angles, st.floats(-1.4, 1.4), angles, st.integers(0, 10_000), st.integers(3, 40), st.floats(1e-7, 1e-3))
def test_wahba_error_is_invariant_to_attitude_ordering_and_weights(ra, dec, roll, seed, n, noise):
    """The error angle depends only on the body-frame measurement errors: not on the truth attitude (incl. any roll
    re-parameterisation), the star order, or a common weight rescaling (SURVEY.md §8c vii)."""
    r = star_field(seed, n, 0.08)                                       # body-frame truth vectors, ~9 deg field
    rng = np.random.default_rng(seed + 1)
    b = r + rng.normal(0, noise, r.shape)
    b /= np.linalg.norm(b, axis=1, keepdims=True)

    def err(C, order, w):
        s = r[order] @ C.T                                              # inertial references s = C r
        B = (w * b[order][:, :, None] * s[:, None, :]).sum(0)
        q, _ = trial.q_method(B[None])
        return trial.rotation_angle(trial.quat_to_dcm(q) @ C[None])[0]

    C0 = np.eye(3)
    C1 = trial.truth_attitude(np.array([ra]), np.array([dec]), np.array([roll]))[0]
    base = err(C0, np.arange(n), 1.0 / n)
    assert abs(err(C1, np.arange(n), 1.0 / n) - base) < 1e-11
    assert abs(err(C1, np.random.default_rng(seed + 2).permutation(n), 1.0 / n) - base) < 1e-11
    assert abs(err(C1, np.arange(n), 7.3) - base) < 1e-11
    assert base < 2000 * noise + 1e-12     # sane magnitude (roll about the boresight is amplified ~1/field-width, more for 3 stars)


@settings(max_examples=40, deadline=None)
@given(st.integers(0, 10_000), st.integers(3, 30))
def test_quest_newton_equals_q_method_near_identity(seed, n):
    """QUEST (Newton on the quartic, lambda0 = 1) == Davenport eigh in the body-frame formulation the CUDA path uses."""
    r = star_field(seed, n, 0.1)
    small = Rotation.from_rotvec(np.random.default_rng(seed).normal(0, 1e-3, 3)).as_matrix()
    b = r @ small.T + np.random.default_rng(seed + 5).normal(0, 1e-5, r.shape)
    b /= np.linalg.norm(b, axis=1, keepdims=True)
    B = (b[:, :, None] * r[:, None, :]).mean(0)
    q1, l1 = trial.q_method(B[None])
    q2, l2 = trial.quest(B[None], iters=4)
    assert abs(abs(q1[0] @ q2[0]) - 1.0) < 1e-13 and abs(l1[0] - l2[0]) < 1e-13


def test_bench_reference_arm_json_contract():
    """`bench.py --impl reference` (CPU arm, oracle port on all cores): one JSON line with the contract's keys."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1", "--cpu-trials-per-core", "512"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "mc_trials_per_sec" and line["unit"] == "trials/s"
    assert line["higher_is_better"] is True and line["dtype"] == "f64" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"]
