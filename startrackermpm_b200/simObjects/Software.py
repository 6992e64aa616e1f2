"""Software — centroiding + identification model (constructed at /root/reference/driver.py:147-157; sampled at
monte_carlo.py:94 and sensitivity.py:108).  Columns BASE_DEV_X / BASE_DEV_Y [px] (driver.py:210-211; run.sh:17) are the
per-trial centroiding accuracy; per-star centroid errors are drawn inside the trial (docs/MODEL.md §3)."""
from __future__ import annotations

import pandas as pd

from . import _config
from .Parameter import Parameter


class Software:
    def __init__(self, ctr_json: str, id_json: str | None = None) -> None:
        _, params = _config.load_params(ctr_json)
        if id_json is not None:
            _, idp = _config.load_params(id_json)
            params.update(idp)
        for col in ("BASE_DEV_X", "BASE_DEV_Y"):
            params.setdefault(col, Parameter(0.0, 0.0, col, "px"))
        self.ctr_json, self.id_json = ctr_json, id_json
        self.params: dict[str, Parameter] = params

    def __repr__(self) -> str:
        return "Software: " + "; ".join(repr(p) for p in self.params.values())

    def randomize(self, num: int = 1) -> pd.DataFrame:
        return _config.frame(self.params, num, True)

    def ideal(self, num: int = 1) -> pd.DataFrame:
        return _config.frame(self.params, num, False)
