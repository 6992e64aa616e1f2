"""CPU baseline C1 ("faithful", BASELINE.md §3 / SURVEY.md §8d): the reference's own `driver.py`, UNMODIFIED, run as
__main__ on this repo's drop-in `simObjects` with the numpy oracle as the trial engine — BASELINE.json configs[0]
("driver.py default sensor config, 10^4 Monte Carlo trials on CPU").  One process, one thread, the reference's own
batch-of-100 pandas loop (monte_carlo.py:48-76) and its own clock (driver.py:280,319).  Test/bench infrastructure only.

usage: python -m oracle.run_reference_driver [N]   -> one JSON line {rows, seconds, trials_per_s, count, sigma_arcsec}
"""
import json
import logging
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_dir():
    for d in ("/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if os.path.exists(os.path.join(d, "driver.py")):
            return d
    return None


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
    ref = reference_dir()
    if ref is None:
        print(json.dumps({"unavailable": "reference scripts not staged (run __graft_entry__.build() where /root/reference exists)"}))
        return 0
    import numpy as np
    sys.path[:0] = [ref, os.path.join(ROOT, "startrackermpm_b200"), os.path.join(ROOT, "tests", "stubs"), ROOT]
    from simObjects.Simulation import Simulation
    from oracle.engine_double import OracleEngine
    Simulation.engine_factory = staticmethod(lambda f_ideal: OracleEngine(f_ideal))
    OracleEngine(3500.0)                                   # catalog + candidate grid built outside the reference's timed region
    np.random.seed(0)
    sys.argv = ["driver.py", "-n", str(n), "-log", "CRITICAL"]
    logging.disable(logging.CRITICAL)
    g = runpy.run_path(os.path.join(ref, "driver.py"), run_name="__main__")
    rows, dt = len(g["fin_df"].index), float(g["delta"])
    print(json.dumps({"rows": rows, "seconds": dt, "trials_per_s": g["count"] * 100 / dt, "count": int(g["count"]),
                      "sigma_arcsec": float(g["true_std"]), "reference_dir": ref}))
    return 0


if __name__ == "__main__":
    sys.exit(main())
