"""Ad-hoc first GPU check: parity vs oracle + rough throughput. Run under gpurun."""
import sys, time, json, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from startrackermpm_b200.engine import TrialEngine, make_dist
from startrackermpm_b200.catalog import load_catalog
from oracle import trial as otrial

means = dict(FOCAL_LENGTH=3500, MAX_MAGNITUDE=6.0)
sig = dict(FOCAL_LENGTH=2, F_ARR_EPS_X=1, F_ARR_EPS_Y=1, F_ARR_EPS_Z=1, F_ARR_PHI=.05, F_ARR_THETA=.05, F_ARR_PSI=.05,
           BASE_DEV_X=.1, BASE_DEV_Y=.1, MAX_MAGNITUDE=.25, D_TEMP=5.0)
coef = 2e-5
xyz, mag = load_catalog()
eng = TrialEngine(0)
print("device", eng.device_info())
dist = make_dist(means, sig, coef)
seed = 1234
n = 20000
m_full = {k: means.get(k, 0.0) for k in otrial.NORMAL_PARAMS}
s_full = {k: sig.get(k, 0.0) for k in otrial.NORMAL_PARAMS}
cols = otrial.sample_params(m_full, s_full, coef, seed, 0, n)
t = time.perf_counter(); oe, od = otrial.run_trials(cols, xyz, mag, 3500.0, seed, 0); print("oracle s", time.perf_counter() - t)

# 1. generator kernel vs oracle sampler
dcols = torch.empty((13, n), dtype=torch.float64, device="cuda")
eng.fill_presampled_device(dcols, dist, seed=seed)
torch.cuda.synchronize()
print("sampler max abs diff", float((dcols.cpu().numpy() - cols).__abs__().max()))

for prec in (64, 32):
    e, d = eng.run_presampled(cols, seed, 0, precision=prec)
    ok = (oe >= 0) & (e >= 0)
    print(f"prec {prec}: count mismatches {(d != od).sum()} fail mismatch {((oe < 0) != (e < 0)).sum()} "
          f"max |derr| rad {np.abs(e[ok] - oe[ok]).max() / otrial.RAD2ARCSEC:.3e}  mean err {e[ok].mean():.3f} fail {(e<0).mean():.4f}")

# 2. rng mode == presampled(fill)
st, e2, d2 = eng.run_rng(dist, n, seed, 0, 64, records=True)
e1, d1 = eng.run_presampled(cols, seed, 0, 64)
print("rng-vs-presampled max diff", np.abs(e2 - e1).max(), (d2 != d1).sum())
print("summary", eng.summarize(st[0]))
print("oracle mean/std", oe[oe >= 0].mean(), oe[oe >= 0].std(ddof=1), "p99", np.percentile(oe[oe >= 0], 99))

# 3. throughput (device resident)
for prec in (64, 32):
    N = 1 << 24
    dc = torch.empty((13, N), dtype=torch.float64, device="cuda")
    eng.fill_presampled_device(dc, dist, seed=seed)
    err = torch.empty(N, dtype=torch.float64, device="cuda"); nd = torch.empty(N, dtype=torch.int32, device="cuda")
    stats = torch.zeros(eng.stats_len, dtype=torch.float64, device="cuda")
    for mode in ("presampled", "rng_stats", "rng_records"):
        for it in range(3):
            torch.cuda.synchronize(); ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
            ev0.record()
            s = torch.cuda.current_stream().cuda_stream
            if mode == "presampled": eng.run_presampled_device(dc, err, nd, None, seed=seed, precision=prec, stream=s)
            elif mode == "rng_stats": eng.run_rng_device(dist, N, stats, seed=seed, precision=prec, stream=s)
            else: eng.run_rng_device(dist, N, stats, err, nd, seed=seed, precision=prec, stream=s)
            ev1.record(); torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1)
        print(f"prec {prec} {mode}: {ms:.2f} ms for {N} trials -> {N / ms * 1e3:.3e} trials/s")
print("fma peak fp64 TF", eng.fma_peak(64), "fp32 TF", eng.fma_peak(32))
