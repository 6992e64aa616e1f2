"""startrackermpm_b200 — B200-native Monte Carlo trial loop of the StarTrackerMPM measurement-process model.

Only the hot path: hand-written sm_100a CUDA kernels behind a flat C ABI (`libstmpm.so`, declared in
`include/stmpm.h`), a ctypes binding (`_abi`), the bulk host API (`engine`) and the drop-in `simObjects` package the
reference's unmodified `driver.py` / `monte_carlo.py` / `sensitivity.py` import.  There is no CPU fallback: every
compute entry point raises if the CUDA library or a GPU is missing.
"""
import os

__version__ = "0.1.0"

PACKAGE_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(PACKAGE_DIR)
UTILS_DIR = os.path.join(PACKAGE_DIR, "utils")
LIB_PATH = os.path.join(PACKAGE_DIR, "libstmpm.so")


def dropin_path() -> str:
    """Directory to put on sys.path/PYTHONPATH so that `import simObjects` resolves to this repo's drop-in package."""
    return PACKAGE_DIR
