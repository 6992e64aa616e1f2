"""JSON config loading shared by StarTracker / Software / Orbit.

Schema (docs/MODEL.md §7; the reference's own JSON files are absent):
    {"cam_name": "...", "params": {"<COLUMN_NAME>": {"ideal": x, "stddev": s, "units": "..."}, ...}}
A path that does not exist is looked up by basename among this package's defaults, so the paths
/root/reference/constants.py:10-21 builds (relative to the reference checkout, where utils/ is absent) resolve."""
from __future__ import annotations

import json
import os

import numpy as np
import pandas as pd

from startrackermpm_b200 import UTILS_DIR
from .Parameter import Parameter


def resolve(path: str) -> str:
    if path and os.path.exists(path):
        return path
    cand = os.path.join(UTILS_DIR, os.path.basename(path or ""))
    if path and os.path.exists(cand):
        return cand
    raise FileNotFoundError(f"config {path!r} not found (also looked for {cand!r})")


def load_params(path: str) -> tuple[dict, dict]:
    with open(resolve(path)) as f:
        doc = json.load(f)
    params = {name: Parameter(v["ideal"], v.get("stddev", 0.0), name, v.get("units", ""))
              for name, v in doc.get("params", {}).items()}
    return doc, params


def frame(params: dict, num: int, randomize: bool) -> pd.DataFrame:
    num = int(num)
    return pd.DataFrame({k: (p.randomize(num) if randomize else p.idealize(num)) for k, p in params.items()},
                        index=np.arange(num))
