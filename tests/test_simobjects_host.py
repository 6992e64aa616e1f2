"""Host logic of the drop-in `simObjects` package, and BASELINE config 1: the UNMODIFIED reference driver.py /
monte_carlo.py / sensitivity.py running on it (with the numpy oracle injected as the engine — tests only)."""
import logging
import os
import runpy
import sys

import numpy as np
import pandas as pd
import pytest

from conftest import REFERENCE_DIR, ROOT, STUBS_DIR

import startrackermpm_b200
from oracle_engine import OracleEngine

HAVE_REF = os.path.exists(os.path.join(REFERENCE_DIR, "driver.py"))


@pytest.fixture()
def dropin(monkeypatch):
    """sys.path as INTEGRATION.md prescribes + the oracle engine double."""
    monkeypatch.syspath_prepend(startrackermpm_b200.dropin_path())
    monkeypatch.syspath_prepend(STUBS_DIR)
    if HAVE_REF:
        monkeypatch.syspath_prepend(REFERENCE_DIR)
    for m in [m for m in sys.modules if m.split(".")[0] in ("simObjects", "monte_carlo", "sensitivity", "driver")]:
        monkeypatch.delitem(sys.modules, m)
    from simObjects.Simulation import Simulation
    monkeypatch.setattr(Simulation, "engine_factory", staticmethod(lambda f_ideal: OracleEngine(f_ideal)))
    return Simulation


def make_sim(Simulation, n=50):
    from simObjects.Orbit import Orbit
    from simObjects.Software import Software
    from simObjects.StarTracker import StarTracker
    cam = StarTracker(cam_json="utils/cameraBasic.json", cam_name="Basic Camera")   # resolved among packaged defaults
    sw = Software(ctr_json="utils/simpleCentroid.json", id_json="utils/typicalIdentification.json")
    return Simulation(camera=cam, software=sw, orbit=Orbit(), num_runs=n)


def test_parameter_and_component_surface(dropin):
    sim = make_sim(dropin)
    assert sim.camera.cam_name == "Basic Camera"
    assert sim.camera.f_len.ideal == 3500.0 and sim.camera.mag_sensor.ideal == 6.0
    assert {"FOCAL_LENGTH", "F_ARR_EPS_X", "F_ARR_PSI", "MAX_MAGNITUDE", "FOCAL_THERMAL_COEFFICIENT"} <= set(sim.camera.randomize(4).columns)
    assert list(sim.software.ideal(3).columns) == ["BASE_DEV_X", "BASE_DEV_Y"]
    assert "D_TEMP" in sim.orbit.randomize(2).columns
    for k in ("FOCAL_LENGTH", "BASE_DEV_X", "D_TEMP"):
        assert hasattr(sim.params[k], "ideal") and hasattr(sim.params[k], "stddev")     # sensitivity.py:43-44
    ideal = sim.camera.ideal(5)
    assert (ideal["FOCAL_LENGTH"] == 3500.0).all() and (ideal["F_ARR_PHI"] == 0).all()
    with pytest.raises(FileNotFoundError):
        from simObjects.StarTracker import StarTracker
        StarTracker(cam_json="utils/noSuchCamera.json")


def test_half_normal_helpers_accept_scalar_and_series(dropin):
    sim = make_sim(dropin)
    k = np.sqrt(1 - 2 / np.pi)                                       # driver.py:290
    assert sim.half_normal_std(2.0) == pytest.approx(2.0 / k)
    s = pd.Series([1.0, 2.0, 3.0])
    assert np.allclose(sim.half_normal_std(s), s / k)                # monte_carlo.py:60 passes a Series
    assert sim.half_normal_mean(2.0) == pytest.approx(2.0 / k * np.sqrt(2 / np.pi))   # driver.py:291


def test_run_sim_contract(dropin):
    np.random.seed(0)
    sim = make_sim(dropin, n=60)
    q = pd.DataFrame({"RIGHT_ASCENSION": np.random.uniform(0.2, 6.0, 60), "DECLINATION": np.random.uniform(-1.3, 1.3, 60),
                      "ROLL": np.random.uniform(-np.pi, np.pi, 60)})
    sim.sim_data = pd.concat([q, sim.camera.randomize(60), sim.software.randomize(60)], axis=1)
    sim.sim_data.loc[:9, "MAX_MAGNITUDE"] = -5.0                     # force 10 failures
    out = sim.run_sim(params=["FOCAL_LENGTH"], obj_func=sim.quest_objective)
    assert sim.output_name == "QUATERNION_ERROR_[arcsec]"            # toolplot.py:71
    assert len(out) == 50 and (out[sim.output_name] >= 0).all()      # failed rows dropped (driver.py:321)
    assert set(sim.sim_data.columns) <= set(out.columns)
    out2 = sim.run_sim(params=[], obj_func=sim.quest_objective, trueData=sim.sim_data.iloc[10:20].copy())
    assert len(out2) == 10
    assert not np.array_equal(out2[sim.output_name].to_numpy(), out[sim.output_name].to_numpy()[:10])  # new trial indices
    with pytest.raises(NotImplementedError):
        sim.run_sim(params=[], obj_func=sim.phi_model)
    mag_out = sim.run_sim(params=[], obj_func=sim.star_mag)          # -func M (driver.py:246-247)
    assert sim.output_name == "MAXIMUM MAGNITUDE VISIBLE" and len(mag_out) == 50   # toolplot.py:241
    assert (mag_out[sim.output_name] <= mag_out["MAX_MAGNITUDE"] + 1e-6).all() and (mag_out[sim.output_name] > 3).all()
    assert isinstance(sim.quest_objective(sim.sim_data.iloc[30]), float)


def test_product_path_has_no_oracle_or_cpu_fallback():
    """The product package must never import oracle/ (only tests, smoke() and bench's cpu legs may)."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "startrackermpm_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


@pytest.mark.skipif(not HAVE_REF, reason="reference checkout not present (it does not travel to the GPU box)")
def test_unmodified_reference_driver_monte_carlo(dropin, monkeypatch, caplog):
    """BASELINE config 1: `python driver.py -n 1000` (default -cam B -sw B -func Q -sim M -b 100), reference files untouched."""
    np.random.seed(0)
    monkeypatch.setattr(sys, "argv", ["driver.py", "-n", "1000", "-log", "WARNING"])
    with caplog.at_level(logging.CRITICAL):
        g = runpy.run_path(os.path.join(REFERENCE_DIR, "driver.py"), run_name="__main__")
    fin = g["fin_df"]
    assert g["count"] == 10 and 990 <= len(fin) <= 1000
    assert g["acc_col"] == "QUATERNION_ERROR_[arcsec]"
    assert 100 < g["true_std"] < 250                                 # default config: sigma ~ 160 arcsec
    assert any("FAILURE RATE" in r.getMessage() for r in caplog.records)
    assert {"RIGHT_ASCENSION", "FOCAL_LENGTH", "BASE_DEV_X", "MAX_MAGNITUDE"} <= set(fin.columns)


@pytest.mark.skipif(not HAVE_REF, reason="reference checkout not present")
def test_unmodified_reference_sensitivity_sweep(dropin, monkeypatch):
    """run.sh's PHI_LONG study: `driver.py -sim s -par F_ARR_PSI` — 20 000 rows in 200 chunks of 100 (sensitivity.py:19-21,71)."""
    import sensitivity
    monkeypatch.setattr(sensitivity, "PER_PARAM", 4)                 # 400 rows instead of 20 000: same code path, CPU-sized
    np.random.seed(1)
    from simObjects.Orbit import Orbit
    from simObjects.Software import Software
    from simObjects.StarTracker import StarTracker
    import constants as c
    cam = StarTracker(cam_json=c.BASIC_CAM, cam_name="Basic Camera")
    sw = Software(ctr_json=c.IDEAL_CENTROID, id_json=c.IDEAL_IDENT)
    s = sensitivity.Sense(cam, sw, Orbit(), 100)
    out = s.run_sim(params=["F_ARR_PSI"], obj_func=s.quest_objective)
    assert len(out) == 400 and s.count == 4
    psi = out["F_ARR_PSI"].to_numpy()
    assert psi.min() == pytest.approx(-0.15) and psi.max() == pytest.approx(0.15)      # ideal +/- 3 sigma (sensitivity.py:43-44)
    err = out["QUATERNION_ERROR_[arcsec]"].to_numpy()
    assert np.allclose(err, np.abs(psi) * 3600.0, atol=1e-6)          # pure psi -> error = |psi| exactly (MAX_MAGNITUDE = 20)
    assert (out["MAX_MAGNITUDE"] == 20).all() and "D_FOCAL_LENGTH" in out.columns      # sensitivity.py:112,118


@pytest.mark.skipif(not HAVE_REF, reason="reference checkout not present")
def test_unmodified_toolplot_reads_the_bulk_result_file(monkeypatch, tmp_path, capsys):
    """/root/reference/toolplot.py, unmodified, on the pickle `studies.write_result_pickle` writes.
    It prints its statistics (toolplot.py:80-96) before it starts plotting, where the matplotlib stub stops it; the numbers
    must be the oracle's toolplot reductions."""
    from oracle import stats as ostats
    from oracle import trial
    from startrackermpm_b200 import studies
    from conftest import BASIC_MEAN, BASIC_SIGMA, THERMAL_COEF
    monkeypatch.syspath_prepend(STUBS_DIR)
    monkeypatch.syspath_prepend(REFERENCE_DIR)
    cols = trial.sample_params(BASIC_MEAN, BASIC_SIGMA, THERMAL_COEF, 12, 0, 3000)
    df = studies.result_frame(OracleEngine(), cols=cols, seed=12)
    path = studies.write_result_pickle(df, "FULL_TEST", str(tmp_path / "data"))
    ref = ostats.toolplot_reductions(df[studies.QUEST_OUTPUT].to_numpy())
    for mode in ("m", "s"):
        monkeypatch.setattr(sys, "argv", ["toolplot.py", "-f", path, "-s", mode, "-p"])    # absolute -f wins os.path.join
        with pytest.raises(RuntimeError, match="matplotlib stub"):
            runpy.run_path(os.path.join(REFERENCE_DIR, "toolplot.py"), run_name="__main__")
        out = capsys.readouterr().out.split()
        assert out[:2] == ["bins:", "50"] and int(out[2]) == len(df) and int(out[3]) == ref["n_kept"]
        if mode == "m":
            assert out[4:7] == ["HMAX", "IDX:", str(ref["mode_bin"])]
            assert float(out[7].split(":")[1]) == pytest.approx(ref["mode"], rel=1e-12)
