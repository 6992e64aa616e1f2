"""Scheduler stress (test tooling): many launches of random size / sort window / mode; every per-trial record must equal
the records of one reference run, and the packed statistics must account for every trial.  Shakes the lock-free window
pipeline of the trial kernel (claim counter, `built` flag, `reads` counters) over ~10^4 window builds with every
combination of partially filled windows and groups.  Usage: python tests/tools_scheduler_stress.py [seconds]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from startrackermpm_b200.engine import TrialEngine, make_dist

MEAN = dict(FOCAL_LENGTH=3500.0, F_ARR_EPS_X=0.0, F_ARR_EPS_Y=0.0, F_ARR_EPS_Z=0.0, F_ARR_PHI=0.0, F_ARR_THETA=0.0,
            F_ARR_PSI=0.0, BASE_DEV_X=0.0, BASE_DEV_Y=0.0, MAX_MAGNITUDE=6.0, D_TEMP=0.0)
SIGMA = dict(FOCAL_LENGTH=2.0, F_ARR_EPS_X=1.0, F_ARR_EPS_Y=1.0, F_ARR_EPS_Z=1.0, F_ARR_PHI=0.05, F_ARR_THETA=0.05,
             F_ARR_PSI=0.05, BASE_DEV_X=0.1, BASE_DEV_Y=0.1, MAX_MAGNITUDE=0.25, D_TEMP=5.0)


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 45.0
    eng = TrialEngine(0, f_ideal=3500.0)
    d = make_dist(MEAN, SIGMA, 2e-5)
    N = 3_000_000
    s = torch.cuda.current_stream().cuda_stream
    cols = torch.empty((13, N), dtype=torch.float64, device="cuda")
    eng.fill_presampled_device(cols, d, seed=77, stream=s)
    ref_e = torch.empty(N, dtype=torch.float64, device="cuda"); ref_d = torch.empty(N, dtype=torch.int32, device="cuda")
    os.environ["STMPM_WIN"] = "256"
    eng.run_presampled_device(cols, ref_e, ref_d, None, seed=77, stream=s)
    torch.cuda.synchronize()
    rng = np.random.default_rng(0)
    e = torch.empty(N, dtype=torch.float64, device="cuda"); nd = torch.empty(N, dtype=torch.int32, device="cuda")
    stats = torch.zeros(eng.stats_len, dtype=torch.float64, device="cuda")
    t0 = time.perf_counter(); launches = trials = 0
    while time.perf_counter() - t0 < budget:
        win = int(rng.choice([32, 64, 96, 128, 160, 256, 512, 1024, 2048]))
        n = int(rng.choice([rng.integers(1, 300), rng.integers(300, 40_000), rng.integers(40_000, N)]))
        a = int(rng.integers(0, N - n + 1))
        os.environ["STMPM_WIN"] = str(win)
        e.fill_(-7.0); nd.fill_(-7); stats.zero_()
        if rng.random() < 0.5:      # pre-sampled slice [a, a+n) with trial0 = a
            eng.run_presampled_device(cols[:, a:a + n], e[:n], nd[:n], stats, seed=77, trial0=a, stream=s)
        else:                       # native RNG on the same index range: must reproduce the same records
            eng.run_rng_device(d, n, stats, e[:n], nd[:n], seed=77, trial0=a, stream=s)
        torch.cuda.synchronize()
        assert torch.equal(e[:n], ref_e[a:a + n]) and torch.equal(nd[:n], ref_d[a:a + n]), (win, n, a)
        st = stats.cpu().numpy(); ok = int((ref_e[a:a + n] >= 0).sum().item())
        assert st[0] == ok and st[1] == n - ok and st[4] == int(ref_d[a:a + n].sum().item()), (win, n, a, st[:5])
        assert st[6] + st[7] + st[8:].sum() == ok
        launches += 1; trials += n
    print(f"scheduler stress: {launches} launches, {trials:.3e} trials, all records and statistics exact "
          f"({time.perf_counter() - t0:.1f} s)")


if __name__ == "__main__":
    main()
