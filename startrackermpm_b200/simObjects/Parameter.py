"""Parameter — a scalar random parameter with `.ideal` and `.stddev` (used at /root/reference/sensitivity.py:43-44,50,118).
Sampling is normal(ideal, stddev) from the global legacy numpy RNG, like the rest of the reference (monte_carlo.py:90-92)."""
from __future__ import annotations

import numpy as np


class Parameter:
    def __init__(self, ideal: float, stddev: float, name: str = "", units: str = "") -> None:
        self.ideal = float(ideal)
        self.stddev = float(stddev)
        self.name = name
        self.units = units

    def __repr__(self) -> str:
        return f"{self.name}: {self.ideal} (+/- {self.stddev}) [{self.units}]"

    def randomize(self, num: int = 1) -> np.ndarray:
        if self.stddev == 0.0:
            return np.full(num, self.ideal, dtype=np.float64)
        return np.random.normal(self.ideal, self.stddev, num)

    def modulate(self, num: int = 1) -> np.ndarray:   # alias
        return self.randomize(num)

    def idealize(self, num: int = 1) -> np.ndarray:
        return np.full(num, self.ideal, dtype=np.float64)
