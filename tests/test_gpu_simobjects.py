"""The drop-in boundary on the real CUDA engine: simObjects.Simulation.run_sim -> ctypes -> libstmpm.so."""
import os
import sys

import numpy as np
import pandas as pd
import pytest

import startrackermpm_b200
from conftest import REFERENCE_DIR, STUBS_DIR
from oracle import trial

pytestmark = pytest.mark.gpu


@pytest.fixture()
def Simulation(monkeypatch, lib_built):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    monkeypatch.syspath_prepend(startrackermpm_b200.dropin_path())
    monkeypatch.syspath_prepend(STUBS_DIR)
    for m in [m for m in sys.modules if m.split(".")[0] in ("simObjects", "monte_carlo", "sensitivity", "driver")]:
        monkeypatch.delitem(sys.modules, m)
    from simObjects.Simulation import Simulation
    return Simulation


def test_run_sim_on_cuda_matches_oracle(Simulation, catalog):
    from simObjects.Orbit import Orbit
    from simObjects.Software import Software
    from simObjects.StarTracker import StarTracker
    np.random.seed(0)
    n = 2000
    cam = StarTracker(cam_json="utils/cameraBasic.json", cam_name="Basic Camera")
    sw = Software(ctr_json="utils/simpleCentroid.json", id_json="utils/typicalIdentification.json")
    sim = Simulation(camera=cam, software=sw, orbit=Orbit(), num_runs=n)
    fov = np.deg2rad(10)
    q = pd.DataFrame({"RIGHT_ASCENSION": np.random.uniform(fov, 2 * np.pi - fov, n),       # monte_carlo.py:90-92
                      "DECLINATION": np.random.uniform(-np.pi / 2 + fov, np.pi / 2 - fov, n),
                      "ROLL": np.random.uniform(-np.pi, np.pi, n)})
    sim.sim_data = pd.concat([q, cam.randomize(n), sw.randomize(n)], axis=1)
    sim.sim_data.loc[:4, "MAX_MAGNITUDE"] = -9.0
    out = sim.run_sim(params=["FOCAL_LENGTH"], obj_func=sim.quest_objective)
    assert type(sim.engine).__module__ == "startrackermpm_b200.engine"          # the CUDA engine, not a double
    xyz, mag = catalog
    cols = np.stack([sim.sim_data[c].to_numpy() for c in trial.COLUMNS])
    e_ref, d_ref = trial.run_trials(cols, xyz, mag, 3500.0, sim._philox_seed(), 0)
    keep = e_ref >= 0
    assert len(out) == keep.sum() and not keep[:5].any()
    assert np.abs(out[sim.output_name].to_numpy() - e_ref[keep]).max() / trial.RAD2ARCSEC < 1e-9
    assert np.array_equal(out["N_STARS_DETECTED"].to_numpy(), d_ref[keep])


@pytest.mark.skipif(not os.path.exists(os.path.join(REFERENCE_DIR, "monte_carlo.py")),
                    reason="reference checkout does not travel to the GPU box")
def test_unmodified_monte_carlo_on_cuda(Simulation, monkeypatch):
    monkeypatch.syspath_prepend(REFERENCE_DIR)
    import constants as c
    import monte_carlo
    from simObjects.Orbit import Orbit
    from simObjects.Software import Software
    from simObjects.StarTracker import StarTracker
    np.random.seed(0)
    sim = monte_carlo.MonteCarlo(StarTracker(cam_json=c.BASIC_CAM), Software(ctr_json=c.SIMPLE_CENTROID), Orbit(), 100)
    fin = sim.run_sim(params=["FOCAL_LENGTH"], obj_func=sim.quest_objective, maxRuns=2000)
    assert sim.count == 20 and 1980 <= len(fin) <= 2000
