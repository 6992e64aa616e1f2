"""AttitudeEstimation — in the reference this module held `QUEST` and `RandomProjection` (only commented-out imports
survive: /root/reference/driver.py:14, monte_carlo.py:11, sensitivity.py:13).  Here the projection and the QUEST solve
are fused into the CUDA trial kernel (startrackermpm_b200/csrc/stmpm_device.cuh); these thin classes expose them for
single frames through the same engine, so nothing numerical lives in Python."""
from __future__ import annotations

import numpy as np


class QUEST:
    """Attitude error of one or more frames via the engine (pre-sampled mode)."""

    def __init__(self, engine) -> None:
        self.engine = engine

    def error_arcsec(self, cols: np.ndarray, seed: int = 0, trial0: int = 0, precision: int = 64):
        return self.engine.run_presampled(np.ascontiguousarray(cols, dtype=np.float64), seed, trial0, precision)


class RandomProjection(QUEST):
    """Kept for import compatibility; projection happens inside the same kernel as the solve."""
