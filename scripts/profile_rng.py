"""Native-RNG mode workload for profiling (config 5's kernel): a few launches of 2e7 trials, statistics only."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from startrackermpm_b200.engine import TrialEngine, make_dist
from bench import BASIC_MEAN, BASIC_SIGMA, THERMAL_COEF, F_IDEAL, SEED
eng = TrialEngine(0, f_ideal=F_IDEAL)
d = make_dist(BASIC_MEAN, BASIC_SIGMA, THERMAL_COEF)
stats = torch.zeros(eng.stats_len, dtype=torch.float64, device="cuda")
s = torch.cuda.current_stream().cuda_stream
for _ in range(5):
    eng.run_rng_device(d, 20_000_000, stats, None, None, seed=SEED, stream=s)
torch.cuda.synchronize()
print("n_ok", float(stats[0].item()))
