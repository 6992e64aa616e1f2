"""Orbit — LEO environment draws (constructed at /root/reference/driver.py:265; sampled at monte_carlo.py:95 — result
discarded there — and sensitivity.py:109, where the D_TEMP column drives the thermal focal update :115-118).
Only the environment *parameters* are on the hot path; orbital mechanics are out of scope."""
from __future__ import annotations

import pandas as pd

from . import _config
from .Parameter import Parameter


class Orbit:
    def __init__(self, orbit_json: str = "orbitLEO.json") -> None:
        _, params = _config.load_params(orbit_json)
        params.setdefault("D_TEMP", Parameter(0.0, 0.0, "D_TEMP", "K"))
        self.params: dict[str, Parameter] = params

    def __repr__(self) -> str:
        return "Orbit: " + "; ".join(repr(p) for p in self.params.values())

    def randomize(self, num: int = 1) -> pd.DataFrame:
        return _config.frame(self.params, num, True)

    def ideal(self, num: int = 1) -> pd.DataFrame:
        return _config.frame(self.params, num, False)
