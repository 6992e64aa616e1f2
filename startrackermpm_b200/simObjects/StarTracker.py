"""StarTracker — camera hardware model (constructed at /root/reference/driver.py:111-127; sampled at
monte_carlo.py:93 and sensitivity.py:107; attributes f_len / mag_sensor read at sensitivity.py:50,62,118)."""
from __future__ import annotations

import os

import pandas as pd

from . import _config
from .Parameter import Parameter

HARDWARE_COLUMNS = ("FOCAL_LENGTH", "F_ARR_EPS_X", "F_ARR_EPS_Y", "F_ARR_EPS_Z", "F_ARR_PHI", "F_ARR_THETA",
                    "F_ARR_PSI", "MAX_MAGNITUDE", "FOCAL_THERMAL_COEFFICIENT")
_DEFAULTS = {"FOCAL_LENGTH": (3500.0, "px"), "MAX_MAGNITUDE": (6.0, "mag"), "FOCAL_THERMAL_COEFFICIENT": (0.0, "1/K")}


class StarTracker:
    def __init__(self, cam_json: str, cam_name: str | None = None) -> None:
        doc, params = _config.load_params(cam_json)
        self.cam_json = cam_json
        self.cam_name = cam_name or doc.get("cam_name") or os.path.splitext(os.path.basename(cam_json))[0]
        for col in HARDWARE_COLUMNS:   # a study JSON may list only the parameters it perturbs (run.sh:63)
            if col not in params:
                ideal, units = _DEFAULTS.get(col, (0.0, "px" if "EPS" in col else "deg"))
                params[col] = Parameter(ideal, 0.0, col, units)
        self.params: dict[str, Parameter] = params
        self.f_len: Parameter = params["FOCAL_LENGTH"]
        self.mag_sensor: Parameter = params["MAX_MAGNITUDE"]

    def __repr__(self) -> str:
        return f"{self.cam_name}: " + "; ".join(repr(p) for p in self.params.values())

    def randomize(self, num: int = 1) -> pd.DataFrame:
        return _config.frame(self.params, num, True)

    def ideal(self, num: int = 1) -> pd.DataFrame:
        return _config.frame(self.params, num, False)
