"""Times BASELINE.json configs 3, 4, 5 (native-RNG, statistics-only) on the local GPU(s).  Under torchrun each rank takes
its share (grid points / orbit steps round-robin; config 5: a contiguous trial-index range + one stats all-reduce).
Usage: python scripts/run_configs.py [--scale 1.0] [--configs 3 4 5]   -> JSON lines on rank 0"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from startrackermpm_b200 import sharding, studies
from startrackermpm_b200.engine import TrialEngine, make_dist, summarize
from bench import BASIC_MEAN, BASIC_SIGMA, THERMAL_COEF, F_IDEAL

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=float, default=1.0, help="fraction of each config's trial count")
ap.add_argument("--configs", type=int, nargs="*", default=[3, 4, 5])
ap.add_argument("--precision", type=int, default=64)
ap.add_argument("--all-visible", action="store_true",
                help="config 3 with the reference sweep's own convention MAX_MAGNITUDE = 20 on every row (sensitivity.py:112)")
args = ap.parse_args()
world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", 1), ("RANK", 0), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

def sync():
    if world > 1: dist.barrier()
    torch.cuda.synchronize()

def emit(**kw):
    if rank == 0: print(json.dumps(kw), flush=True)

if 3 in args.configs:      # sensitivity grid 64 x 64 x 16, 1e6 trials per point, coarse histograms (34 * 32 bins)
    eng = TrialEngine(local, f_ideal=F_IDEAL, hist_log2_sub=5)
    tps = max(1000, int(1e6 * args.scale))
    ea, pa, da = studies.grid_axes(BASIC_SIGMA["F_ARR_EPS_X"], BASIC_SIGMA["F_ARR_PHI"], BASIC_SIGMA["BASE_DEV_X"])
    seg = list(range(*sharding.shard_range(64 * 64 * 16, rank, world)))   # contiguous block: 4096 points per launch
    mean3, sig3 = (dict(BASIC_MEAN, MAX_MAGNITUDE=20.0), dict(BASIC_SIGMA, MAX_MAGNITUDE=0.0)) if args.all_visible else (BASIC_MEAN, BASIC_SIGMA)
    sync(); t = time.perf_counter()
    st = studies.sensitivity_grid(eng, mean3, sig3, THERMAL_COEF, ea, pa, da, tps, seed=3,
                                  precision=args.precision, segments=seg)
    sync(); dt = time.perf_counter() - t
    s0, s1 = summarize(st[0], 5), summarize(st[-1], 5)
    emit(config=3, n_gpus=world, grid=[64, 64, 16], trials_per_point=tps, trials=65536 * tps, seconds=dt,
         trials_per_sec=65536 * tps / dt, first_point_mean=s0["mean"], last_point_mean=s1["mean"], precision=args.precision,
         max_magnitude=20.0 if args.all_visible else BASIC_MEAN["MAX_MAGNITUDE"], mean_ndet=s0["mean_ndet"])
    eng.close()
if 4 in args.configs:      # LEO lifetime: 60 monthly steps, 1e7 trials per step
    eng = TrialEngine(local, f_ideal=F_IDEAL)
    tps = max(1000, int(1e7 * args.scale))
    dists, sched = studies.leo_lifetime_dists(BASIC_MEAN, BASIC_SIGMA, THERMAL_COEF, n_steps=60)
    seg = list(sharding.shard_segments(60, rank, world))
    sync(); t = time.perf_counter()
    st = studies.run_segments(eng, dists, tps, seed=4, precision=args.precision, segments=seg)
    sync(); dt = time.perf_counter() - t
    ss = studies.summaries(st)
    emit(config=4, n_gpus=world, steps=60, trials_per_step=tps, trials=60 * tps, seconds=dt, trials_per_sec=60 * tps / dt,
         sigma_first=ss[0]["sigma_halfnormal"], sigma_last=ss[-1]["sigma_halfnormal"], fail_first=ss[0]["fail_rate"],
         fail_last=ss[-1]["fail_rate"], precision=args.precision)
    eng.close()
if 5 in args.configs:      # tail statistics: 1e10 trials over the ranks, p99.9 sketch, ONE stats all-reduce
    eng = TrialEngine(local, f_ideal=F_IDEAL)
    n = int(1e10 * args.scale)
    b, e = sharding.shard_range(n, rank, world)
    d = make_dist(BASIC_MEAN, BASIC_SIGMA, THERMAL_COEF)
    stats = torch.zeros(eng.stats_len, dtype=torch.float64, device="cuda")
    sync(); t = time.perf_counter()
    eng.run_rng_device(d, e - b, stats, seed=5, trial0=b, precision=args.precision, stream=torch.cuda.current_stream().cuda_stream)
    sharding.allreduce_stats(stats)
    sync(); dt = time.perf_counter() - t
    s = summarize(stats.cpu().numpy())
    n_ok = s["n_ok"]; p = 0.999
    ci = 3.0 * np.sqrt(p * (1 - p) / n_ok)           # +/- 3 sigma binomial band on the quantile level
    emit(config=5, n_gpus=world, trials=n, seconds=dt, trials_per_sec=n / dt, p999=s["p999"], p99=s["p99"], mean=s["mean"],
         sigma_halfnormal=s["sigma_halfnormal"], fail_rate=s["fail_rate"], p999_level_ci=[p - ci, p + ci], precision=args.precision)
if world > 1:
    dist.destroy_process_group()
