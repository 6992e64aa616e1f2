"""TEST-ONLY engine double: same surface as startrackermpm_b200.engine.TrialEngine.run_presampled, backed by the numpy
oracle.  Lets the host logic (simObjects, the unmodified reference scripts) run on a box without a GPU — BASELINE config 1,
"the reference numpy path".  Never imported by product code."""
import numpy as np

from oracle import trial
from startrackermpm_b200.catalog import load_catalog


class OracleEngine:
    def __init__(self, f_ideal=3500.0, **_):
        self.f_ideal = float(f_ideal)
        self.xyz, self.mag = load_catalog()
        self.calls = 0

    def run_presampled(self, cols, seed, trial0=0, precision=64, want_stats=False, want_maxmag=False, **_):
        self.calls += 1
        e, d, aux = trial.run_trials(np.asarray(cols, dtype=np.float64), self.xyz, self.mag, self.f_ideal, seed, trial0,
                                     return_aux=True)
        return (e, d, aux["maxmag"]) if want_maxmag else (e, d)
