"""Device-resident throughput of every mode x precision (CUDA events, 3 warm-ups, 5 timed launches of 2^26 trials)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from startrackermpm_b200.engine import TrialEngine, make_dist
from bench import BASIC_MEAN, BASIC_SIGMA, THERMAL_COEF, F_IDEAL, SEED
eng = TrialEngine(0, f_ideal=F_IDEAL)
if os.environ.get("MM_KEY_TABLE"): eng.set_tuning(key_table=int(os.environ["MM_KEY_TABLE"]))   # experiment knob
d = make_dist(BASIC_MEAN, BASIC_SIGMA, THERMAL_COEF)
N = 1 << 26
cols = torch.empty((13, N), dtype=torch.float64, device="cuda")
err = torch.empty(N, dtype=torch.float64, device="cuda"); nd = torch.empty(N, dtype=torch.int32, device="cuda")
stats = torch.zeros(eng.stats_len, dtype=torch.float64, device="cuda")
s = torch.cuda.current_stream().cuda_stream
eng.fill_presampled_device(cols, d, seed=SEED, stream=s)
out = {}
for prec in (64, 32):
    for mode in ("presampled_records", "presampled_records_stats", "rng_stats", "rng_records_stats"):
        def run():
            if mode == "presampled_records": eng.run_presampled_device(cols, err, nd, None, seed=SEED, precision=prec, stream=s)
            elif mode == "presampled_records_stats": eng.run_presampled_device(cols, err, nd, stats, seed=SEED, precision=prec, stream=s)
            elif mode == "rng_stats": eng.run_rng_device(d, N, stats, seed=SEED, precision=prec, stream=s)
            else: eng.run_rng_device(d, N, stats, err, nd, seed=SEED, precision=prec, stream=s)
        for _ in range(3): run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): run()
        e1.record(); torch.cuda.synchronize()
        out[f"{mode}_fp{prec}"] = 5 * N / (e0.elapsed_time(e1) * 1e-3)
print(json.dumps({k: float(f"{v:.4e}") for k, v in out.items()}))
