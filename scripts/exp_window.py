"""Experiment: device-resident throughput against the CTA sort-window size, with and without the per-trial record stores,
and for the native-RNG mode (no column gathers at all) — separates the cost of the permuted gathers / scattered stores from
the benefit of the wider sort.  Usage: python scripts/exp_window.py [windows...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from startrackermpm_b200.engine import TrialEngine, make_dist
from bench import BASIC_MEAN, BASIC_SIGMA, THERMAL_COEF, F_IDEAL, SEED
eng = TrialEngine(0, f_ideal=F_IDEAL)
d = make_dist(BASIC_MEAN, BASIC_SIGMA, THERMAL_COEF)
N = 1 << 26
cols = torch.empty((13, N), dtype=torch.float64, device="cuda")
err = torch.empty(N, dtype=torch.float64, device="cuda"); nd = torch.empty(N, dtype=torch.int32, device="cuda")
stats = torch.zeros(eng.stats_len, dtype=torch.float64, device="cuda")
s = torch.cuda.current_stream().cuda_stream
eng.fill_presampled_device(cols, d, seed=SEED, stream=s)
def timeit(fn):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): fn()
    e1.record(); torch.cuda.synchronize()
    return 3 * N / (e0.elapsed_time(e1) * 1e-3)
for w in [int(x) for x in sys.argv[1:]] or [128, 512, 2048]:
    eng.set_tuning(window=w)
    a = timeit(lambda: eng.run_presampled_device(cols, err, nd, stats, seed=SEED, stream=s))
    b = timeit(lambda: eng.run_presampled_device(cols, None, None, stats, seed=SEED, stream=s))
    c = timeit(lambda: eng.run_rng_device(d, N, stats, err, nd, seed=SEED, stream=s))
    e = timeit(lambda: eng.run_rng_device(d, N, stats, None, None, seed=SEED, stream=s))
    print(f"win={w:5d}  presampled+records {a:.4e}  presampled stats-only {b:.4e}  rng+records {c:.4e}  rng stats-only {e:.4e}", flush=True)
