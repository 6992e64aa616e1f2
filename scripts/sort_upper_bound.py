"""Experiment: how much would perfect work-sorting buy?  Times the presampled kernel on (a) the normal column order and
(b) the same trials globally pre-sorted by MAX_MAGNITUDE (every 128-trial window then holds near-identical work)."""
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
s = torch.cuda.current_stream().cuda_stream
eng.fill_presampled_device(cols, d, seed=SEED, stream=s)
def timeit(c):
    for _ in range(3): eng.run_presampled_device(c, err, nd, None, seed=SEED, stream=s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): eng.run_presampled_device(c, err, nd, None, seed=SEED, stream=s)
    e1.record(); torch.cuda.synchronize()
    return 5 * N / (e0.elapsed_time(e1) * 1e-3)
print("as generated      : %.4e trials/s" % timeit(cols))
order = torch.argsort(cols[12])
sorted_cols = cols[:, order].contiguous()
print("sorted by max_mag : %.4e trials/s" % timeit(sorted_cols))
nd_sorted = nd.clone()
order2 = torch.argsort(nd_sorted.to(torch.int64) * 1000 + (sorted_cols[12] * 10).to(torch.int64))   # by actual detections
sorted2 = sorted_cols[:, order2].contiguous()
print("sorted by n_det   : %.4e trials/s" % timeit(sorted2))
