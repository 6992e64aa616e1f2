"""ORACLE-backed engine double (test infrastructure, not product code): the surface of
`startrackermpm_b200.engine.TrialEngine.run_presampled`, computed by the numpy oracle on the CPU.

Two users, both allowed by the contract: tests/ (host logic of `simObjects` and the unmodified reference scripts on a box
without a GPU) and bench.py's CPU arm (`cpu_baseline` C1: the reference's own `driver.py -n 10000` loop on one host core).
The product path never imports this (tests/test_simobjects_host.py::test_product_path_has_no_oracle_or_cpu_fallback)."""
import numpy as np

from . import trial


class OracleEngine:
    def __init__(self, f_ideal=3500.0, catalog=None, use_grid=True, **_):
        self.f_ideal = float(f_ideal)
        if catalog is None:
            from startrackermpm_b200.catalog import load_catalog
            catalog = load_catalog()
        self.xyz, self.mag = catalog
        self.grid = trial.candidate_grid(self.xyz) if use_grid else None
        self.calls = 0

    def run_presampled(self, cols, seed, trial0=0, precision=64, want_stats=False, want_maxmag=False, **_):
        self.calls += 1
        e, d, aux = trial.run_trials(np.asarray(cols, dtype=np.float64), self.xyz, self.mag, self.f_ideal, seed, trial0,
                                     return_aux=True, grid=self.grid)
        return (e, d, aux["maxmag"]) if want_maxmag else (e, d)
