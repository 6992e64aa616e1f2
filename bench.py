#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on BASELINE.json's config (contract: task brief §④).

Metric   : MC trials/s (= attitude solves/s), whole job over N GPUs.
Workload : configs[1] — default sensor config, 10^8 trials per GPU per step, pre-sampled fp64 parameter columns
           resident in HBM (written by the seeded generator kernel), per-trial (error f64, n_det i32) records out,
           streaming statistics accumulated; fp64 mode (the reference computes in float64).
A "step" : one pass of the hot path (stmpm_trials_presampled) over that batch.  Inputs are 10.4 GB per step, far
           larger than the 126 MB L2, so no explicit L2 flush is needed between steps.
value    : device-resident throughput (CUDA events on the launching stream, barrier + synchronize both sides, max over ranks).
e2e      : the same batch through the host-buffer C-ABI call (stmpm_run_presampled_host: pinned host columns -> H2D ->
           kernel -> D2H records), wall-clocked around the synchronous call.
--impl reference : the CPU arm — the numpy oracle port (the reference's own trial code is absent from its mount,
           SURVEY.md §0) on all host cores, bounded sample per step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "mc_trials_per_sec"
UNIT = "trials/s"
SEED = 20231105
F_IDEAL = 3500.0
THERMAL_COEF = 2e-5
BASIC_MEAN = dict(FOCAL_LENGTH=3500.0, F_ARR_EPS_X=0.0, F_ARR_EPS_Y=0.0, F_ARR_EPS_Z=0.0, F_ARR_PHI=0.0,
                  F_ARR_THETA=0.0, F_ARR_PSI=0.0, BASE_DEV_X=0.0, BASE_DEV_Y=0.0, MAX_MAGNITUDE=6.0, D_TEMP=0.0)
BASIC_SIGMA = dict(FOCAL_LENGTH=2.0, F_ARR_EPS_X=1.0, F_ARR_EPS_Y=1.0, F_ARR_EPS_Z=1.0, F_ARR_PHI=0.05,
                   F_ARR_THETA=0.05, F_ARR_PSI=0.05, BASE_DEV_X=0.1, BASE_DEV_Y=0.1, MAX_MAGNITUDE=0.25, D_TEMP=5.0)
# --workload: "default" = BASELINE configs[1]; "all_visible" = the reference's sensitivity-sweep regime, where
# Sense.create_data sets MAX_MAGNITUDE = 20 for every row (sensitivity.py:112): every star in the field is detected (~60)
WORKLOADS = {
    "default": (BASIC_MEAN, BASIC_SIGMA,
                "configs[1]: default sensor config (cameraBasic + simpleCentroid), pre-sampled fp64 parameter columns -> "
                "per-trial error records + streaming statistics"),
    "all_visible": (dict(BASIC_MEAN, MAX_MAGNITUDE=20.0), dict(BASIC_SIGMA, MAX_MAGNITUDE=0.0),
                    "star-rich fields: default sensor config with MAX_MAGNITUDE = 20 on every row (sensitivity.py:112, the "
                    "reference's sweep regime, ~60 detections per trial), pre-sampled fp64 columns -> records + statistics"),
}
BYTES_IN_PER_TRIAL = 13 * 8          # pre-sampled columns
BYTES_OUT_PER_TRIAL = 8 + 4          # error f64 + n_det i32


def work_flops_per_trial(n_cand: float, n_star: float) -> float:
    """Algorithmic work per pre-sampled trial: FMA = 2 flop, every transcendental / div /
    rsqrt = 1 flop.  F = 430 + 7 N_cand + 75 N_star + 6 N_star (Box-Muller arithmetic of the 2 N_star centroid normals)."""
    return 430.0 + 7.0 * n_cand + 81.0 * n_star


# ------------------------------------------------------------------------------------------------ CPU arm (oracle port)
# Three baselines, all on the numpy fp64 oracle with the SAME sky-grid candidate search
# idea the GPU uses (oracle.trial.CandidateGrid: ~170 listed stars per trial instead of all 9 110):
#   C1 "faithful": the reference's own driver.py -n 10000, unmodified, one core, its own pandas batch-of-100 loop and clock
#   C2 "fair-1"  : the vectorised oracle on one core
#   C3 "fair-N"  : the vectorised oracle, one process per host core
def _cpu_worker(args):
    t0, n = args
    from oracle import trial
    from startrackermpm_b200.catalog import load_catalog
    global _CPU_CAT
    try:
        xyz, mag, grid = _CPU_CAT
    except NameError:
        xyz, mag = load_catalog()
        grid = trial.candidate_grid(xyz)
        _CPU_CAT = (xyz, mag, grid)
    mean, sigma, _ = WORKLOADS[os.environ.get("STMPM_BENCH_WORKLOAD", "default")]
    cols = trial.sample_params(mean, sigma, THERMAL_COEF, SEED, t0, n)
    e, d = trial.run_trials(cols, xyz, mag, F_IDEAL, SEED, t0, grid=grid)
    return int((e >= 0).sum()), float(e[e >= 0].sum())


def cpu_trials_per_sec(per_core: int, cores: int, pool=None):
    """Times the numpy fp64 oracle (kind 'port') on `cores` processes, per_core trials each. Returns (trials/s, seconds)."""
    import multiprocessing as mp
    own = pool is None
    if own:
        os.environ.setdefault("OMP_NUM_THREADS", "1")    # one process per core; no BLAS oversubscription
        pool = mp.get_context("spawn").Pool(cores)
        pool.map(_cpu_worker, [(i * 64, 64) for i in range(cores)])      # warm: imports, catalog load, grid build
    t = time.perf_counter()
    res = pool.map(_cpu_worker, [(i * per_core, per_core) for i in range(cores)])
    dt = time.perf_counter() - t
    if own:
        pool.close()
    assert sum(r[0] for r in res) > 0
    return cores * per_core / dt, dt


def cpu_c1_faithful(n: int = 10000):
    """C1: unmodified reference driver.py -n 10000 (configs[0]) in a fresh single-threaded process; its own clock."""
    env = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", MKL_NUM_THREADS="1")
    try:
        r = subprocess.run([sys.executable, "-m", "oracle.run_reference_driver", str(n)], cwd=ROOT, env=env,
                           capture_output=True, text=True, timeout=600)
        d = json.loads(r.stdout.strip().splitlines()[-1])
    except Exception as ex:                                               # noqa: BLE001 — a baseline must not sink the bench
        return {"unavailable": f"{type(ex).__name__}: {str(ex)[:120]}"}
    if "unavailable" in d:
        return d
    return {"value": d["trials_per_s"], "unit": UNIT, "cores": 1, "rows": d["rows"], "seconds": d["seconds"],
            "sigma_arcsec": d["sigma_arcsec"],
            "what": "unmodified reference driver.py -n 10000 (-b 100: 100 batches through its own pandas loop) on the drop-in "
                    "simObjects with the numpy oracle as engine; timed by driver.py:280,319"}


def cpu_baselines(per_core_all: int, cores: int, per_core_one: int = 100_000, pool=None):
    """cpu_baseline object: headline value = C3 (all cores); C1 and C2 beside it."""
    c3, dt3 = cpu_trials_per_sec(per_core_all, cores, pool)
    c2, dt2 = cpu_trials_per_sec(per_core_one, 1)
    return {"value": c3, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{cores * per_core_all} trials of the same workload ({per_core_all} per process x {cores} processes, "
                      f"numpy fp64 oracle with sky-grid candidate lists, {dt3:.1f} s)",
            "c1_faithful": cpu_c1_faithful(),
            "c2_vectorised_1core": {"value": c2, "unit": UNIT, "cores": 1, "sample": f"{per_core_one} trials, {dt2:.1f} s"},
            "c3_all_cores": {"value": c3, "unit": UNIT, "cores": cores},
            "candidate_search": "sky-grid lists (oracle.trial.CandidateGrid, ~170 stars per trial); round 1 timed the "
                                "brute-force 9 110-star product"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    per_core = args.cpu_trials_per_core
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    pool = mp.get_context("spawn").Pool(cores)
    pool.map(_cpu_worker, [(i * 64, 64) for i in range(cores)])
    for _ in range(args.warmup):
        cpu_trials_per_sec(max(256, per_core // 8), cores, pool)
    t = time.perf_counter()
    for _ in range(args.steps):
        cpu_trials_per_sec(per_core, cores, pool)
    dt = time.perf_counter() - t
    n = cores * per_core
    v = n * args.steps / dt
    c2, dt2 = cpu_trials_per_sec(max(20_000, per_core), 1)
    pool.close()
    sample = f"{n} trials per step ({per_core} per process x {cores} processes) of the same default-config workload"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(n, args.gpus, args.workload),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "c1_faithful": cpu_c1_faithful(),
                         "c2_vectorised_1core": {"value": c2, "unit": UNIT, "cores": 1, "sample": f"{max(20_000, per_core)} trials, {dt2:.1f} s"},
                         "c3_all_cores": {"value": v, "unit": UNIT, "cores": cores},
                         "candidate_search": "sky-grid lists (oracle.trial.CandidateGrid, ~170 stars per trial)"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "numpy fp64 oracle port (oracle/trial.py): the reference's own trial code is absent from its mount "
                "(SURVEY.md §0), so no unmodified implementation of the trial exists to install or time; its entry scripts "
                "ARE staged (baseline/_ref) and drive C1",
    }))
    return 0


def workload_config(trials_per_gpu_step: int, n_gpus: int, workload: str = "default"):
    return {"workload": WORKLOADS[workload][2],
            "trials_per_gpu_per_step": int(trials_per_gpu_step), "catalog": "synthetic 9110 stars, seed 20231105",
            "f_ideal_px": F_IDEAL, "sensor_px": 1024, "precision": 64, "sharding": f"trial-index ranges x{n_gpus}",
            "l2_policy": "inputs (10.4 GB/step at 1e8 trials) larger than L2; no flush needed"}


# ------------------------------------------------------------------------------------------------ clocks sampler
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [l.split(", ") for t, l in self.lines if t0 <= t <= t1] or [l.split(", ") for _, l in self.lines[-3:]]
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for name, flag in zip(names, r[4:8]):
                    if flag.strip().lower() == "active":
                        reasons.add(name)
            except (ValueError, IndexError):
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def bind_to_gpu_numa_node(local: int):
    """Best effort: pin this rank's host threads — and, through first-touch, the pinned buffers it allocates afterwards — to
    the CPUs of the NUMA node its GPU hangs off (sysfs `local_cpulist` of the GPU's PCI function).  Returns what was found
    and done, for the JSON line.  On this pool the 8-GPU box exposes ONE node (numa_node 0 / -1 and the same CPU list for
    every GPU), so this is a recorded no-op there."""
    info = {"numa_node": None, "cpus": None, "bound": False, "nodes_visible": None}
    try:
        import glob
        info["nodes_visible"] = len(glob.glob("/sys/devices/system/node/node[0-9]*"))
    except Exception:                                                      # noqa: BLE001
        pass
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        dev = f"/sys/bus/pci/devices/{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        info["numa_node"] = int(open(dev + "/numa_node").read().strip())
        cpulist = open(dev + "/local_cpulist").read().strip()
        info["cpus"] = cpulist
        cpus = set()
        for part in cpulist.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        target = cpus & allowed
        if target and target != allowed:
            os.sched_setaffinity(0, target)
            info["bound"] = True
    except Exception as ex:                                                # noqa: BLE001 — sysfs layout / permissions vary
        info["error"] = f"{type(ex).__name__}: {str(ex)[:80]}"
    return info


# ------------------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    import ctypes

    from startrackermpm_b200 import _abi, build, sharding
    from startrackermpm_b200.engine import TrialEngine, make_dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the trial loop has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank == 0:
        build.build()
    if world > 1:
        dist.barrier()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n = int(args.trials)
    eng = TrialEngine(local, f_ideal=F_IDEAL)
    if args.window or args.key_radius_deg >= 0 or args.stage or args.key_table >= 0:
        eng.set_tuning(args.window, args.key_radius_deg, args.stage, args.key_table)
    if args.host_chunk:
        eng.set_host_pipeline(args.host_chunk)
    if args.pipeline >= 0:
        eng.set_pipeline(args.pipeline)
    d = make_dist(WORKLOADS[args.workload][0], WORKLOADS[args.workload][1], THERMAL_COEF)
    trial0 = rank * n                                      # weak scaling: every rank its own global trial-index range
    cols = torch.empty((13, n), dtype=torch.float64, device="cuda")
    err = torch.empty(n, dtype=torch.float64, device="cuda")
    ndet = torch.empty(n, dtype=torch.int32, device="cuda")
    stats = torch.zeros(eng.stats_len, dtype=torch.float64, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    eng.fill_presampled_device(cols, d, seed=SEED, trial0=trial0, stream=stream)

    def step():
        eng.run_presampled_device(cols, err, ndet, stats, seed=SEED, trial0=trial0, precision=64, stream=stream)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    stats.zero_()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    l0 = eng.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    t_wall0 = time.perf_counter()
    ev[0].record()
    for k in range(args.steps):
        step()
        ev[k + 1].record()
    if world > 1:
        sharding.allreduce_stats(stats)                     # the one collective of the path
    barrier()
    t_wall1 = time.perf_counter()
    launches = eng.launch_count() - l0
    clocks = sampler.stop(t_wall0, t_wall1)
    ms_total = ev[0].elapsed_time(ev[-1])
    kernel_ms = float(np.mean([ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]))
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * n * args.steps / (ms_total * 1e-3)

    # ---- the fp32 instantiation on the same device-resident batch (north_star: FP32 and FP64 pipe fractions)
    fp32_ms = None
    if rank == 0 or world > 1:
        for _ in range(2):
            eng.run_presampled_device(cols, err, ndet, None, seed=SEED, trial0=trial0, precision=32, stream=stream)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(2):
            eng.run_presampled_device(cols, err, ndet, None, seed=SEED, trial0=trial0, precision=32, stream=stream)
        f1.record()
        torch.cuda.synchronize()
        fp32_ms = f0.elapsed_time(f1) / 2
        eng.run_presampled_device(cols, err, ndet, None, seed=SEED, trial0=trial0, precision=64, stream=stream)   # records back to fp64

    # ---- e2e: host pinned buffers through the host C-ABI call, copies inside the timed region
    n_e2e = min(n, int(args.e2e_trials))
    e2e = None
    try:
        h_cols = torch.empty((13, n_e2e), dtype=torch.float64, pin_memory=True)
        h_cols.copy_(cols[:, :n_e2e])
        # ONE pinned result buffer of 13 bytes per trial, carved into the views the calls below need (no rank pins more than
        # 117 bytes per trial: eight ranks share one host)
        h_out = torch.empty(13 * n_e2e, dtype=torch.uint8, pin_memory=True).numpy()
        he = h_out[:8 * n_e2e].view(np.float64)
        hn = h_out[8 * n_e2e:12 * n_e2e].view(np.int32)
        h_nd8 = h_out[12 * n_e2e:13 * n_e2e]
        h_dst = h_out[:12 * n_e2e]
        hc = h_cols.numpy()
        eng.run_presampled(hc, SEED, trial0, 64, out_err=he, out_ndet=hn)       # warm: staging buffers allocate here
        barrier()
        e2e_steps = max(1, min(args.steps, 3))
        tw = time.perf_counter()
        for _ in range(e2e_steps):
            eng.run_presampled(hc, SEED, trial0, 64, out_err=he, out_ndet=hn)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - tw], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * n_e2e * e2e_steps / float(dt.item()), "unit": UNIT,
               "h2d_bytes_per_step": BYTES_IN_PER_TRIAL * n_e2e, "d2h_bytes_per_step": BYTES_OUT_PER_TRIAL * n_e2e,
               "trials_per_gpu_per_step": n_e2e, "steps": e2e_steps,
               "api": "stmpm_run_presampled_host (pinned host columns in, host records out, chunked H2D/kernel/D2H pipeline)",
               "records_match_device_run": bool(np.array_equal(he, err[:n_e2e].cpu().numpy())),
               "host_numa": numa}

        def max_over_ranks(seconds):
            t_ = torch.tensor([seconds], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(t_, op=dist.ReduceOp.MAX)
            return float(t_.item())

        def host_roofline(h2d, d2h, src, dst):
            """copies only, same pipeline (stmpm_host_copy_roofline), all ranks at once; best of 2"""
            best = None
            for _ in range(2):
                barrier()
                sec = max_over_ranks(eng.host_copy_roofline(n_e2e, h2d, d2h, src, dst))
                best = sec if best is None else min(best, sec)
            return world * n_e2e / best

        roof = host_roofline(BYTES_IN_PER_TRIAL, BYTES_OUT_PER_TRIAL, hc, h_dst)
        solo = None
        if world > 1:     # one rank alone on the host fabric (the others idle at the barrier): what a single PCIe link gives here
            barrier()
            sec = eng.host_copy_roofline(n_e2e, BYTES_IN_PER_TRIAL, BYTES_OUT_PER_TRIAL, hc, h_dst) if rank == 0 else 0.0
            barrier()
            solo = n_e2e / max_over_ranks(sec)
        e2e["host_roofline"] = {"value": roof, "unit": UNIT, "frac": e2e["value"] / roof,
                                "gb_per_s": roof * (BYTES_IN_PER_TRIAL + BYTES_OUT_PER_TRIAL) / 1e9,
                                "one_rank_alone": None if solo is None else {
                                    "value": solo, "unit": UNIT, "gb_per_s": solo * (BYTES_IN_PER_TRIAL + BYTES_OUT_PER_TRIAL) / 1e9,
                                    "ranks_x_alone": world * solo,
                                    "what": "rank 0's copies with the other ranks idle; ranks x this = the fabric-unlimited aggregate"},
                                "what": "the same host pipeline with the kernel removed: 104 B/trial H2D + 12 B/trial D2H through "
                                        "the same chunks, streams and staging buffers, all ranks concurrently, max over ranks"}
        # the same workload through the native-RNG host entry (MonteCarlo.create_data + run_sim fused on the device):
        # nothing but the distribution goes up, the per-trial records come down (12 B, or 9 B with n_det as uint8)
        native = {}
        n_chk = min(n_e2e, 1 << 20)
        nd_sample = None
        he32 = h_out[:4 * n_e2e].view(np.float32)              # the 5-byte format's float32 errors (aliases he: used last)
        err_sample = None
        for label, compact, eb, nd, nbytes in (("records_12B", False, he, hn, 12), ("records_9B", True, he, h_nd8, 9),
                                               ("records_5B", 5, he32, h_nd8, 5)):
            eng.run_rng(d, n_e2e, SEED, trial0, 64, out_err=eb, out_ndet=nd, compact=compact)
            barrier()
            tw = time.perf_counter()
            for _ in range(e2e_steps):
                eng.run_rng(d, n_e2e, SEED, trial0, 64, out_err=eb, out_ndet=nd, compact=compact)
            torch.cuda.synchronize()
            v = world * n_e2e * e2e_steps / max_over_ranks(time.perf_counter() - tw)
            if nbytes == 12:                                  # (the copy-only pass below reuses these buffers)
                nd_sample = np.minimum(hn[:n_chk], 255).astype(np.uint8)
                err_sample = he[:n_chk].astype(np.float32)
            else:
                assert np.array_equal(nd_sample, h_nd8[:n_chk]), "compact records differ from the int32 records"
                if nbytes == 5:
                    assert np.array_equal(err_sample, he32[:n_chk]), "float32 records differ from the rounded float64 records"
            roof = host_roofline(0, nbytes, None, h_dst)
            native[label] = {"value": v, "unit": UNIT, "d2h_bytes_per_step": nbytes * n_e2e,
                             "host_roofline": {"value": roof, "unit": UNIT, "frac": v / roof, "gb_per_s": roof * nbytes / 1e9}}
        e2e["native_rng"] = {"value": native["records_12B"]["value"], "unit": UNIT,
                             "h2d_bytes_per_step": ctypes.sizeof(_abi.Dist), "d2h_bytes_per_step": BYTES_OUT_PER_TRIAL * n_e2e,
                             "api": "stmpm_run_rng_host (parameters sampled on the device, host records + statistics out)",
                             **native}
    except RuntimeError as ex:                                                   # pinned allocation refused
        e2e = {"value": None, "unit": UNIT, "error": str(ex)[:200]}

    out = None
    if rank == 0:
        st = stats.cpu().numpy()
        summ = eng.summarize(st)
        wc = eng.work_counters(d, 1 << 20, SEED, 0)
        flops_trial = work_flops_per_trial(wc["visits"], wc["detected"])
        peak64 = eng.fma_peak(64)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        k_s = kernel_ms * 1e-3
        achieved_tf = flops_trial * n / k_s / 1e12
        achieved_gbs = (BYTES_IN_PER_TRIAL + BYTES_OUT_PER_TRIAL) * n / k_s / 1e9
        traffic, traffic_src, pipes = None, None, None
        try:   # static: DRAM bytes and pipe utilisation come from the committed ncu capture of this kernel, not from this run
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            traffic = tj.get("dram_bytes_per_trial")
            traffic = traffic * n if traffic is not None else None
            traffic_src = f"profiles/traffic.json (ncu --set full capture {tj.get('capture', '?')}; static, not measured in this run)"
            pipes = tj.get("pipes")
        except (OSError, ValueError):
            pass
        peak32 = eng.fma_peak(32)
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(n, world, args.workload),
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {
                "bound": "fp64", "kernel": "trials_kernel<presampled, fp64>",
                "achieved": achieved_tf, "peak": peak64, "unit": "TFLOP/s", "frac": achieved_tf / peak64 if peak64 else None,
                "peak_source": "stmpm_fma_peak: FP64 FMA-chain microbenchmark measured live on this GPU "
                               "(MEASURED_PEAKS.json holds no FP64 figure)",
                "flops_per_trial": flops_trial, "work_model": "430 + 7*N_visit + 81*N_star (DESIGN.md §4)",
                "n_visit": wc["visits"], "n_mag_pass": wc["mag_pass"], "n_survivors": wc["survivors"],
                "n_star": wc["detected"], "kernel_ms": kernel_ms, "traffic": traffic, "traffic_source": traffic_src,
                "pipes_pct_of_peak": pipes,
                "peak_fp32": {"value": peak32, "unit": "TFLOP/s", "source": "stmpm_fma_peak(32), measured live"},
                "fp32_mode": None if not fp32_ms else {
                    "kernel": "trials_kernel<presampled, fp32>", "kernel_ms": fp32_ms, "trials_per_s": n / (fp32_ms * 1e-3),
                    "vs_fp64_mode": kernel_ms / fp32_ms,
                    "achieved": flops_trial * n / (fp32_ms * 1e-3) / 1e12, "unit": "TFLOP/s",
                    "frac_of_fp64_peak": flops_trial * n / (fp32_ms * 1e-3) / 1e12 / peak64 if peak64 else None,
                    "frac_of_fp32_peak": flops_trial * n / (fp32_ms * 1e-3) / 1e12 / peak32 if peak32 else None,
                    "note": "same algorithmic flops; the dot products, projection scale and B stay fp64, the noise draw is fp32 (DESIGN.md §6)"},
                "hbm": {"achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": achieved_gbs / hbm_peak,
                        "bytes_per_trial": BYTES_IN_PER_TRIAL + BYTES_OUT_PER_TRIAL,
                        "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)"},
            },
            "result": {k: summ[k] for k in ("n_ok", "n_fail", "fail_rate", "mean", "std", "sigma_halfnormal", "p99",
                                            "p999", "mean_ndet")},
        }
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baselines(args.cpu_trials_per_core, os.cpu_count() or 1)
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--trials", type=float, default=1e8, help="trials per GPU per step (configs[1]: 1e8)")
    ap.add_argument("--e2e-trials", type=float, default=1e8, help="trials per GPU per e2e step")
    ap.add_argument("--cpu-trials-per-core", type=int, default=None,
                    help="oracle-port trials per host process: default 262144 for the cpu_baseline leg (~13 s), 65536 per step "
                         "for --impl reference (~3 s per step)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="default", choices=sorted(WORKLOADS),
                    help="default = BASELINE configs[1]; all_visible = MAX_MAGNITUDE 20 (sensitivity.py:112), ~60 stars per trial")
    ap.add_argument("--host-chunk", type=int, default=0, help="tuning sweep: trials per stage of the host pipelines (0 = 2^22)")
    ap.add_argument("--pipeline", type=int, default=-1, help="tuning sweep: -1 auto, 0 single kernel, 1 split pipeline")
    ap.add_argument("--stage", action="store_true", help="tuning sweep: stage the pre-sampled columns in key order (measured slower)")
    ap.add_argument("--key-table", type=int, default=-1, help="tuning sweep: -1 per mode, 0 disc, 1 array footprint")
    ap.add_argument("--window", type=int, default=0, help="tuning sweep: CTA sort window in trials (0 = default)")
    ap.add_argument("--key-radius-deg", type=float, default=-1.0, help="tuning sweep: work-key disc radius (<0 = default)")
    args = ap.parse_args()
    os.environ["STMPM_BENCH_WORKLOAD"] = args.workload          # the spawned CPU-arm workers read it
    if args.cpu_trials_per_core is None:
        args.cpu_trials_per_core = 65536 if args.impl == "reference" else 262144
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
