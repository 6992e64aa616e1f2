import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REFERENCE_DIR = "/root/reference"
STUBS_DIR = os.path.join(ROOT, "tests", "stubs")
GOLDEN = os.path.join(ROOT, "tests", "golden", "trial_golden_v1.npz")      # regression pin (made by the oracle itself)
GOLDEN_V2 = os.path.join(ROOT, "tests", "golden", "trial_golden_v2.npz")   # made by tests/golden/independent_trial.py
# the reference's entry scripts: /root/reference here; on the GPU box the copy __graft_entry__.build() staged
# (baseline/_ref is git-ignored but travels with the snapshot)
STAGED_REFERENCE_DIR = os.path.join(ROOT, "baseline", "_ref")
if not os.path.exists(os.path.join(REFERENCE_DIR, "driver.py")) and os.path.exists(os.path.join(STAGED_REFERENCE_DIR, "driver.py")):
    REFERENCE_DIR = STAGED_REFERENCE_DIR

BASIC_MEAN = dict(FOCAL_LENGTH=3500.0, F_ARR_EPS_X=0.0, F_ARR_EPS_Y=0.0, F_ARR_EPS_Z=0.0, F_ARR_PHI=0.0,
                  F_ARR_THETA=0.0, F_ARR_PSI=0.0, BASE_DEV_X=0.0, BASE_DEV_Y=0.0, MAX_MAGNITUDE=6.0, D_TEMP=0.0)
BASIC_SIGMA = dict(FOCAL_LENGTH=2.0, F_ARR_EPS_X=1.0, F_ARR_EPS_Y=1.0, F_ARR_EPS_Z=1.0, F_ARR_PHI=0.05,
                   F_ARR_THETA=0.05, F_ARR_PSI=0.05, BASE_DEV_X=0.1, BASE_DEV_Y=0.1, MAX_MAGNITUDE=0.25, D_TEMP=5.0)
ZERO_SIGMA = {k: 0.0 for k in BASIC_SIGMA}
THERMAL_COEF = 2e-5


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA GPU (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def catalog():
    from startrackermpm_b200.catalog import load_catalog
    return load_catalog()


@pytest.fixture(scope="session")
def golden():
    with np.load(GOLDEN) as f:
        return {k: f[k] for k in f.files}


@pytest.fixture(scope="session")
def golden_v2():
    """Inputs and outputs of the independent scalar restatement (scipy align_vectors, libm, own Philox)."""
    with np.load(GOLDEN_V2) as f:
        return {k: f[k] for k in f.files}


@pytest.fixture(scope="session")
def lib_built():
    """Build libstmpm.so if it is stale (nvcc cross-compiles without a GPU)."""
    from startrackermpm_b200 import build
    return build.build()


@pytest.fixture(scope="session")
def engine(lib_built):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from startrackermpm_b200.engine import TrialEngine
    return TrialEngine(0, f_ideal=3500.0)
