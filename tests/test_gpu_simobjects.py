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


def test_reference_scripts_are_present():
    """The reference's entry scripts must be on this box: /root/reference in the build container, the unmodified copy
    `__graft_entry__.build()` stages into baseline/_ref on the GPU box.  A skip here would leave the drop-in claim
    unobserved (VERDICT r1 "What's missing" #1), so this is an assertion, not a skipif."""
    for f in ("driver.py", "monte_carlo.py", "sensitivity.py", "toolplot.py", "constants.py"):
        assert os.path.exists(os.path.join(REFERENCE_DIR, f)), f"{f} not staged: run __graft_entry__.build() before gpurun"


def test_unmodified_monte_carlo_on_cuda(Simulation, monkeypatch):
    """MonteCarlo.run_sim (monte_carlo.py:30-80), unmodified, on simObjects -> libstmpm.so."""
    monkeypatch.syspath_prepend(REFERENCE_DIR)
    import constants as c
    import monte_carlo
    from simObjects.Orbit import Orbit
    from simObjects.Software import Software
    from simObjects.StarTracker import StarTracker
    np.random.seed(0)
    sim = monte_carlo.MonteCarlo(StarTracker(cam_json=c.BASIC_CAM), Software(ctr_json=c.SIMPLE_CENTROID), Orbit(), 100)
    fin = sim.run_sim(params=["FOCAL_LENGTH"], obj_func=sim.quest_objective, maxRuns=2000)
    assert type(sim.engine).__module__ == "startrackermpm_b200.engine"
    assert sim.count == 20 and 1980 <= len(fin) <= 2000


def test_unmodified_sensitivity_sweep_on_cuda(Simulation, monkeypatch, catalog):
    """Sense.run_sim (sensitivity.py:32-97), unmodified and at its full size: 20 000 rows in 200 chunks of 100, every
    star visible (MAX_MAGNITUDE = 20, sensitivity.py:112 — the star-rich regime), checked row by row against the oracle."""
    monkeypatch.syspath_prepend(REFERENCE_DIR)
    import constants as c
    import sensitivity
    from simObjects.Orbit import Orbit
    from simObjects.Software import Software
    from simObjects.StarTracker import StarTracker
    np.random.seed(1)
    s = sensitivity.Sense(StarTracker(cam_json=c.BASIC_CAM, cam_name="Basic Camera"),
                          Software(ctr_json=c.SIMPLE_CENTROID, id_json=c.TYP_IDENT), Orbit(), 100)
    out = s.run_sim(params=["F_ARR_PHI", "F_ARR_THETA"], obj_func=s.quest_objective)      # run.sh:11 PHI_LAT
    assert type(s.engine).__module__ == "startrackermpm_b200.engine"
    assert s.count == 200 and len(out) == 20_000 and (out["MAX_MAGNITUDE"] == 20).all()
    phi = out["F_ARR_PHI"].to_numpy()
    assert phi.min() == pytest.approx(-0.15) and phi.max() == pytest.approx(0.15) and np.array_equal(phi, out["F_ARR_THETA"])
    assert out["N_STARS_DETECTED"].mean() > 50                                   # ~60 detections per trial
    # per-row parity: the frame holds the inputs; no row was dropped, so row k ran as trial index k of the sim's seed
    xyz, mag = catalog
    cols = np.stack([out[k].to_numpy() for k in trial.COLUMNS])
    worst = 0.0
    for k in range(0, 20_000, 4000):                                            # five chunks of the 200 are re-run on the CPU
        sl = slice(k, k + 100)
        e_ref, d_ref = trial.run_trials(cols[:, sl], xyz, mag, 3500.0, s._philox_seed(), k)
        assert np.array_equal(out["N_STARS_DETECTED"].to_numpy()[sl], d_ref)
        worst = max(worst, np.abs(out[s.output_name].to_numpy()[sl] - e_ref).max() / trial.RAD2ARCSEC)
    assert worst < 1e-9


def test_unmodified_driver_config1_on_cuda(Simulation, monkeypatch, caplog):
    """BASELINE config 1 — `python driver.py -n 10000` (defaults: -cam B -sw B -func Q -sim M -b 100), the reference's
    file run as __main__, on the CUDA engine.  Prints the wall time the reference itself measures (driver.py:280,319)."""
    import logging
    import runpy
    import time
    monkeypatch.syspath_prepend(REFERENCE_DIR)
    np.random.seed(0)
    monkeypatch.setattr(sys, "argv", ["driver.py", "-n", "10000", "-log", "WARNING"])
    t0 = time.perf_counter()
    with caplog.at_level(logging.CRITICAL):
        g = runpy.run_path(os.path.join(REFERENCE_DIR, "driver.py"), run_name="__main__")
    wall = time.perf_counter() - t0
    assert type(g["sim"].engine).__module__ == "startrackermpm_b200.engine"
    fin = g["fin_df"]
    assert g["count"] == 100 and 9950 <= len(fin) <= 10000 and g["acc_col"] == "QUATERNION_ERROR_[arcsec]"
    assert 140 < g["true_std"] < 185                                  # default config: sigma ~ 162 arcsec
    assert any("FAILURE RATE" in r.getMessage() for r in caplog.records)
    print(f"\nconfig 1 on CUDA: driver.py -n 10000 -> {len(fin)} rows in {g['delta']:.3f} s "
          f"({len(fin) / g['delta']:.0f} trials/s through the reference's own pandas loop; wall {wall:.3f} s), "
          f"sigma {g['true_std']:.2f} arcsec")
