"""Large-sample parity run (test tooling, may import oracle/): CUDA path vs numpy oracle on N trials per configuration,
oracle evaluated on all host cores.  Prints one JSON line per (config, precision) — committed as profiles/r1_parity_at_scale.jsonl.
Usage: python tests/tools_parity_at_scale.py [N]"""
import json, multiprocessing as mp, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

SEED = 424242
MEAN = dict(FOCAL_LENGTH=3500.0, F_ARR_EPS_X=0.0, F_ARR_EPS_Y=0.0, F_ARR_EPS_Z=0.0, F_ARR_PHI=0.0, F_ARR_THETA=0.0,
            F_ARR_PSI=0.0, BASE_DEV_X=0.0, BASE_DEV_Y=0.0, MAX_MAGNITUDE=6.0, D_TEMP=0.0)
SIGMA = dict(FOCAL_LENGTH=2.0, F_ARR_EPS_X=1.0, F_ARR_EPS_Y=1.0, F_ARR_EPS_Z=1.0, F_ARR_PHI=0.05, F_ARR_THETA=0.05,
             F_ARR_PSI=0.05, BASE_DEV_X=0.1, BASE_DEV_Y=0.1, MAX_MAGNITUDE=0.25, D_TEMP=5.0)
CONFIGS = {"default": (MEAN, SIGMA), "sigma_x5": (MEAN, {k: 5 * v for k, v in SIGMA.items()}),
           "all_visible": (dict(MEAN, MAX_MAGNITUDE=20.0), SIGMA)}


def _oracle(args):
    name, t0, n = args
    from oracle import trial
    from startrackermpm_b200.catalog import load_catalog
    xyz, mag = load_catalog()
    m, s = CONFIGS[name]
    cols = trial.sample_params(m, s, 2e-5, SEED, t0, n)
    e, d = trial.run_trials(cols, xyz, mag, 3500.0, SEED, t0)
    return cols, e, d


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2_000_000
    from oracle import trial
    from startrackermpm_b200.engine import TrialEngine
    eng = TrialEngine(0, f_ideal=3500.0)
    cores = os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    pool = mp.get_context("spawn").Pool(cores)
    for name in CONFIGS:
        nn = n if name != "all_visible" else n // 4
        per = 4096
        t = time.perf_counter()
        parts = pool.map(_oracle, [(name, t0, min(per, nn - t0)) for t0 in range(0, nn, per)])
        cols = np.concatenate([p[0] for p in parts], 1); e_ref = np.concatenate([p[1] for p in parts]); d_ref = np.concatenate([p[2] for p in parts])
        t_oracle = time.perf_counter() - t
        for prec in (64, 32):
            e, d = eng.run_presampled(cols, SEED, 0, precision=prec)
            ok = e_ref >= 0
            print(json.dumps({"config": name, "precision": prec, "trials": nn, "star_detections": int(d_ref.sum()),
                              "detection_count_mismatches": int((d != d_ref).sum()),
                              "failure_flag_mismatches": int(((e < 0) != (e_ref < 0)).sum()),
                              "max_abs_err_diff_rad": float(np.abs(e[ok] - e_ref[ok]).max() / trial.RAD2ARCSEC),
                              "bar_rad": 1e-9 if prec == 64 else 1e-6, "n_failed_trials": int((~ok).sum()),
                              "oracle_seconds": round(t_oracle, 1), "oracle_cores": cores}), flush=True)
    pool.close()


if __name__ == "__main__":
    main()
