"""Simulation — base class of the reference's MonteCarlo / Sense strategies (/root/reference/monte_carlo.py:19-24,
sensitivity.py:24-30).  `run_sim` is THE hot path (called at monte_carlo.py:52 and sensitivity.py:77): here it marshals
the DataFrame's 13 columns to the CUDA trial engine (libstmpm.so through ctypes) and returns the rows that produced an
attitude solution, with the output column appended.

No trial arithmetic happens in Python and there is no CPU fallback: without libstmpm.so + a GPU, run_sim raises.
"""
from __future__ import annotations

import logging
import os

import numpy as np
import pandas as pd

from .Orbit import Orbit
from .Parameter import Parameter
from .Software import Software
from .StarTracker import StarTracker

logger = logging.getLogger(__name__)

QUEST_OUTPUT = "QUATERNION_ERROR_[arcsec]"      # toolplot.py:71
N_DET_OUTPUT = "N_STARS_DETECTED"
MAG_OUTPUT = "MAXIMUM MAGNITUDE VISIBLE"            # toolplot.py:241

_COLUMN_DEFAULT_FROM = {"RIGHT_ASCENSION": 0.0, "DECLINATION": 0.0, "ROLL": 0.0}


def _default_engine_factory(f_ideal: float):
    from startrackermpm_b200.engine import TrialEngine
    return TrialEngine(device=int(os.environ.get("STMPM_DEVICE", "0")), f_ideal=f_ideal)


class Simulation:
    # Hook for tests only (they inject an oracle-backed engine to exercise the host logic without a GPU).  The default —
    # and the only thing product code ever sets — is the CUDA engine.
    engine_factory = staticmethod(_default_engine_factory)

    def __init__(self, *, camera: StarTracker, software: Software, orbit: Orbit, num_runs: int) -> None:
        self.camera = camera
        self.software = software
        self.orbit = orbit
        self.num_runs = int(num_runs)
        self.params: dict[str, Parameter] = {**camera.params, **software.params, **orbit.params}
        self.sim_data = pd.DataFrame()
        self.output_name: str | None = None
        self.count = 0
        self.type = "SIM"
        self.precision = int(os.environ.get("STMPM_PRECISION", "64"))
        self._engine = None
        self._seed: int | None = None
        self._next_trial = 0

    def __repr__(self) -> str:
        return f"{self.num_runs} runs/batch; {self.camera.cam_name}; {self.software!r}"

    # ---------------------------------------------------------------- engine plumbing
    @property
    def engine(self):
        if self._engine is None:
            self._engine = type(self).engine_factory(self.camera.f_len.ideal)
        return self._engine

    def _philox_seed(self) -> int:
        # the reference draws everything from the global legacy numpy RNG (never seeded: monte_carlo.py:90-92);
        # the per-star Philox streams are keyed from it, so np.random.seed(k) makes a whole run reproducible
        if self._seed is None:
            self._seed = int(np.random.randint(0, 2**31 - 1)) | (int(np.random.randint(0, 2**31 - 1)) << 32)
        return self._seed

    def _columns(self, data: pd.DataFrame) -> np.ndarray:
        from startrackermpm_b200.engine import COLUMNS
        n = len(data.index)
        cols = np.empty((len(COLUMNS), n), dtype=np.float64)
        for i, name in enumerate(COLUMNS):
            if name in data.columns:
                cols[i] = data[name].to_numpy(dtype=np.float64)
            elif name in self.params:
                cols[i] = self.params[name].ideal
            else:
                cols[i] = _COLUMN_DEFAULT_FROM.get(name, 0.0)
        return cols

    def _run_engine(self, data: pd.DataFrame, want_maxmag: bool = False):
        cols = self._columns(data)
        n = cols.shape[1]
        out = self.engine.run_presampled(cols, self._philox_seed(), self._next_trial, self.precision,
                                         want_maxmag=want_maxmag)
        self._next_trial += n
        return out

    # ---------------------------------------------------------------- objectives (selected at driver.py:240-247)
    def quest_objective(self, data):
        """QUEST first-principles objective.  Accepts a DataFrame (vectorised: the whole batch is one kernel launch) or a
        single row (Series); returns the attitude error in arcsec (-1 = fewer than 3 detected stars)."""
        self.output_name = QUEST_OUTPUT
        if isinstance(data, pd.Series):
            err, _ = self._run_engine(data.to_frame().T)
            return float(err[0])
        err, _ = self._run_engine(data)
        return err

    def star_mag(self, data):
        """Magnitude objective (-func M, driver.py:246-247): catalog magnitude of the faintest star the trial detects —
        the column 'MAXIMUM MAGNITUDE VISIBLE' that toolplot.py:241 histograms.  A by-product of the same kernel."""
        self.output_name = MAG_OUTPUT
        if isinstance(data, pd.Series):
            return float(self._run_engine(data.to_frame().T, want_maxmag=True)[2][0])
        return self._run_engine(data, want_maxmag=True)[2]

    def phi_model(self, data):
        raise NotImplementedError(
            "phi_model (-func P, 'Phi Model Analysis', driver.py:243-244): its body is absent from the reference mount "
            "and its semantics cannot be recovered (SURVEY.md §8a R9)")

    # ---------------------------------------------------------------- the hot path
    def run_sim(self, params, obj_func, trueData: pd.DataFrame | None = None, **kwargs) -> pd.DataFrame:
        """Evaluate obj_func over every row of trueData (or self.sim_data).  Returns the input columns plus
        `self.output_name`; rows whose trial failed (fewer than 3 detected stars) are dropped, so the caller's row
        shortfall is the failure count (driver.py:321)."""
        data = self.sim_data if trueData is None else trueData
        if len(data.index) == 0:
            raise ValueError("run_sim: no rows (call create_data first)")
        func = getattr(obj_func, "__func__", obj_func)
        if func is Simulation.quest_objective:
            self.output_name = QUEST_OUTPUT
            err, ndet = self._run_engine(data)
            out = data.copy()
            out[QUEST_OUTPUT] = err
            out[N_DET_OUTPUT] = ndet
            out = out[err >= 0.0]
        elif func is Simulation.star_mag:
            self.output_name = MAG_OUTPUT
            err, ndet, mmag = self._run_engine(data, want_maxmag=True)
            out = data.copy()
            out[MAG_OUTPUT] = mmag.astype(np.float64)
            out[N_DET_OUTPUT] = ndet
            out = out[ndet > 0]                      # no detected star -> no magnitude: row dropped
        elif func is Simulation.phi_model:
            obj_func(data)   # raises NotImplementedError with the explanation
            raise AssertionError("unreachable")
        else:   # a user-supplied per-row callable
            out = data.copy()
            self.output_name = self.output_name or "CALC_ACCURACY"
            out[self.output_name] = [obj_func(row) for _, row in data.iterrows()]
        logger.debug("run_sim: %d rows in, %d out", len(data.index), len(out.index))
        return out

    # ---------------------------------------------------------------- statistics (driver.py:290-291,323-324)
    @staticmethod
    def half_normal_std(x):
        """sigma of the underlying normal from the std of a half-normal sample; scalar or Series (monte_carlo.py:60-61)."""
        return x / np.sqrt(1 - 2 / np.pi)

    @classmethod
    def half_normal_mean(cls, x):
        return cls.half_normal_std(x) * np.sqrt(2 / np.pi)

    def plot_data(self, data: pd.DataFrame, *args) -> None:
        """Histogram of the output column with the half-normal overlay (driver.py:346).  Plotting is outside the hot
        path; it needs matplotlib, which this environment does not ship."""
        try:
            import matplotlib.pyplot as plt
        except ImportError as e:   # pragma: no cover
            raise RuntimeError("plot_data needs matplotlib") from e
        col = self.output_name or QUEST_OUTPUT
        fig, ax = plt.subplots()
        ax.hist(data[col], bins=50, density=True, label="Simulation Data")
        if args:
            s = float(args[0])
            x = np.linspace(0, 1.05 * float(data[col].max()), 1000)
            ax.plot(x, np.sqrt(2) / (s * np.sqrt(np.pi)) * np.exp(-x**2 / (2 * s**2)), "r", label=f"half-normal {s:.3f}\"")
        ax.set_xlabel(col.replace("_", " ").title())
        ax.set_ylabel("Probability Density")
        ax.legend()
