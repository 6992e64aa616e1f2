"""Large-sample parity and scheduler stress as `-m gpu` tests (round 1 ran them by hand: tests/tools_*.py).

* parity at scale: 2·10^6 trials (5·10^5 for the star-rich case) per configuration x {fp64, fp32}: CUDA through the C ABI
  against the numpy oracle evaluated on all host cores.  Bars: detection counts and failure flags bit-exact,
  |d error| <= 1e-9 rad (fp64 mode) / 1e-6 rad (fp32 mode) — BASELINE.json north_star.
* scheduler stress: ~10 s of launches of random size / sort window / mode; every record must equal one reference run's
  and the packed statistics must account for every trial exactly once (the lock-free window pipeline of trials_kernel).
"""
import multiprocessing as mp
import os
import time

import numpy as np
import pytest
import torch

from conftest import BASIC_MEAN, BASIC_SIGMA, THERMAL_COEF
from oracle import trial
from startrackermpm_b200.engine import make_dist

pytestmark = pytest.mark.gpu
SEED = 424242
CONFIGS = {"default": (BASIC_MEAN, BASIC_SIGMA, 2_000_000),
           "sigma_x5": (BASIC_MEAN, {k: 5 * v for k, v in BASIC_SIGMA.items()}, 2_000_000),
           "all_visible": (dict(BASIC_MEAN, MAX_MAGNITUDE=20.0), BASIC_SIGMA, 500_000)}      # sensitivity.py:112


def _oracle(args):
    name, t0, n = args
    from oracle import trial as tr
    from startrackermpm_b200.catalog import load_catalog
    xyz, mag = load_catalog()
    m, s, _ = CONFIGS[name]
    cols = tr.sample_params(m, s, THERMAL_COEF, SEED, t0, n)
    e, d = tr.run_trials(cols, xyz, mag, 3500.0, SEED, t0, grid=tr.candidate_grid(xyz))
    return cols, e, d


@pytest.fixture(scope="module")
def pool():
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    p = mp.get_context("spawn").Pool(os.cpu_count() or 1)
    yield p
    p.close()


@pytest.mark.timeout(900)
@pytest.mark.parametrize("name", list(CONFIGS))
def test_parity_at_scale(engine, pool, name):
    n = int(CONFIGS[name][2] * float(os.environ.get("STMPM_SCALE_TEST_FRACTION", "1")))
    per = 4096
    t = time.perf_counter()
    parts = pool.map(_oracle, [(name, t0, min(per, n - t0)) for t0 in range(0, n, per)])
    cols = np.concatenate([p[0] for p in parts], 1)
    e_ref = np.concatenate([p[1] for p in parts])
    d_ref = np.concatenate([p[2] for p in parts])
    t_oracle = time.perf_counter() - t
    ok = e_ref >= 0
    for prec, bar in ((64, 1e-9), (32, 1e-6)):
        e, d = engine.run_presampled(cols, SEED, 0, precision=prec)
        assert int((d != d_ref).sum()) == 0, f"{name}/{prec}: detection-count mismatches"
        assert np.array_equal(e < 0, ~ok), f"{name}/{prec}: failure flags differ"
        worst = float(np.abs(e[ok] - e_ref[ok]).max() / trial.RAD2ARCSEC)
        print(f"\nparity at scale: {name} fp{prec}: {n} trials, {int(d_ref.sum())} detections, {int((~ok).sum())} failed, "
              f"max |d err| {worst:.2e} rad (bar {bar:.0e}); oracle {t_oracle:.0f} s on {os.cpu_count()} cores")
        assert worst <= bar


@pytest.mark.timeout(300)
def test_scheduler_stress(engine):
    budget = float(os.environ.get("STMPM_STRESS_SECONDS", "10"))
    d = make_dist(BASIC_MEAN, BASIC_SIGMA, THERMAL_COEF)
    N = 3_000_000
    s = torch.cuda.current_stream().cuda_stream
    cols = torch.empty((13, N), dtype=torch.float64, device="cuda")
    engine.fill_presampled_device(cols, d, seed=77, stream=s)
    ref_e = torch.empty(N, dtype=torch.float64, device="cuda")
    ref_d = torch.empty(N, dtype=torch.int32, device="cuda")
    engine.set_tuning(window=256)
    try:
        engine.run_presampled_device(cols, ref_e, ref_d, None, seed=77, stream=s)
        torch.cuda.synchronize()
        rng = np.random.default_rng(0)
        e = torch.empty(N, dtype=torch.float64, device="cuda")
        nd = torch.empty(N, dtype=torch.int32, device="cuda")
        stats = torch.zeros(engine.stats_len, dtype=torch.float64, device="cuda")
        t0 = time.perf_counter()
        launches = trials = 0
        while time.perf_counter() - t0 < budget:
            win = int(rng.choice([32, 64, 96, 128, 160, 256, 512, 1024, 2048]))
            n = int(rng.choice([rng.integers(1, 300), rng.integers(300, 40_000), rng.integers(40_000, N)]))
            a = int(rng.integers(0, N - n + 1))
            engine.set_tuning(window=win)
            e.fill_(-7.0); nd.fill_(-7); stats.zero_()
            if rng.random() < 0.5:      # pre-sampled slice [a, a+n) with trial0 = a
                engine.run_presampled_device(cols[:, a:a + n], e[:n], nd[:n], stats, seed=77, trial0=a, stream=s)
            else:                       # native RNG on the same index range: must reproduce the same records
                engine.run_rng_device(d, n, stats, e[:n], nd[:n], seed=77, trial0=a, stream=s)
            torch.cuda.synchronize()
            assert torch.equal(e[:n], ref_e[a:a + n]) and torch.equal(nd[:n], ref_d[a:a + n]), (win, n, a)
            st = stats.cpu().numpy()
            ok = int((ref_e[a:a + n] >= 0).sum().item())
            assert st[0] == ok and st[1] == n - ok and st[4] == int(ref_d[a:a + n].sum().item()), (win, n, a, st[:5])
            assert st[6] + st[7] + st[8:].sum() == ok
            launches += 1
            trials += n
    finally:
        engine.set_tuning(window=0)
    print(f"\nscheduler stress: {launches} launches, {trials:.3e} trials, all records and statistics exact")
    assert launches > 50
