"""Drop-in `simObjects` package: the first-party package the reference's driver.py / monte_carlo.py / sensitivity.py
import (driver.py:14-19) but which is absent from the reference mount.  Put
`startrackermpm_b200.dropin_path()` on sys.path and the unmodified reference scripts run on the CUDA trial engine."""
import os as _os
import sys as _sys

# When imported as the top-level package `simObjects` (the reference scripts' spelling) only this package's parent
# directory is on sys.path; make the `startrackermpm_b200` package importable too.
_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _root not in _sys.path:
    _sys.path.append(_root)
