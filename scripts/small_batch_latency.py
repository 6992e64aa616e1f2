"""Per-call latency of the host-buffer path at the reference's batch sizes (100 rows per run_sim call)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from startrackermpm_b200.engine import TrialEngine
eng = TrialEngine(0, f_ideal=3500.0)
rng = np.random.default_rng(0)
for n in (1, 100, 1000, 10000, 100000):
    cols = np.zeros((13, n)); cols[0] = rng.uniform(0.2, 6, n); cols[1] = rng.uniform(-1.3, 1.3, n); cols[2] = rng.uniform(-3, 3, n)
    cols[3] = 3500; cols[10] = cols[11] = 0.1; cols[12] = 6.0
    for _ in range(20): eng.run_presampled(cols, 1, 0)
    reps = 300 if n <= 10000 else 50
    t = time.perf_counter()
    for _ in range(reps): eng.run_presampled(cols, 1, 0)
    dt = (time.perf_counter() - t) / reps
    print(f"n={n:7d}  {dt*1e6:9.1f} us/call  {n/dt:12.3e} trials/s")
