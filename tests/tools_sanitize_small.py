"""Small invocation of every kernel path for compute-sanitizer (memcheck / racecheck): both modes, both precisions,
ragged sizes, the band-scan fallback, multi-segment statistics."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # repo root
import numpy as np
from oracle import trial
from startrackermpm_b200.engine import TrialEngine, make_dist
from bench import BASIC_MEAN, BASIC_SIGMA, THERMAL_COEF
eng = TrialEngine(0, f_ideal=3500.0)
for n in (1, 33, 700, 5000):
    cols = trial.sample_params(BASIC_MEAN, BASIC_SIGMA, THERMAL_COEF, 3, 0, n)
    for prec in (64, 32):
        e, d, st = eng.run_presampled(cols, 3, 0, precision=prec, want_stats=True)
        assert st[0] + st[1] == n
wide = trial.sample_params(dict(BASIC_MEAN, FOCAL_LENGTH=1000.0, MAX_MAGNITUDE=20.0), BASIC_SIGMA, 0.0, 3, 0, 300)
eng.run_presampled(wide, 3)
dists = [make_dist(dict(BASIC_MEAN, F_ARR_PSI=0.1 * k), BASIC_SIGMA, THERMAL_COEF) for k in range(3)]
st, e, d = eng.run_rng(dists, 1500, 9, 100, records=True)
st2 = eng.run_rng(dists, 1500, 9, 100, precision=32)
print("sanitize_small ok", st[:, 0], st2[:, 0], eng.work_counters(dists[0], 4096))
