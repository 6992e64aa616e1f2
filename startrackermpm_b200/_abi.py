"""ctypes binding of libstmpm.so — one-to-one with include/stmpm.h.  No compute happens in Python; there is no fallback:
a missing library or GPU raises."""
from __future__ import annotations

import ctypes as C
import os

from . import LIB_PATH

NCOLS = 13
NNORMAL = 11
STATS_HEAD = 8
HIST_OCTAVES = 34
HIST_MIN_EXP = -14

COLUMNS = ("RIGHT_ASCENSION", "DECLINATION", "ROLL", "FOCAL_LENGTH",
           "F_ARR_EPS_X", "F_ARR_EPS_Y", "F_ARR_EPS_Z",
           "F_ARR_PHI", "F_ARR_THETA", "F_ARR_PSI",
           "BASE_DEV_X", "BASE_DEV_Y", "MAX_MAGNITUDE")
NORMAL_PARAMS = ("FOCAL_LENGTH", "F_ARR_EPS_X", "F_ARR_EPS_Y", "F_ARR_EPS_Z", "F_ARR_PHI", "F_ARR_THETA",
                 "F_ARR_PSI", "BASE_DEV_X", "BASE_DEV_Y", "MAX_MAGNITUDE", "D_TEMP")

# every symbol include/stmpm.h declares (tests check that the library exports each one)
EXPORTS = ("stmpm_last_error", "stmpm_version", "stmpm_create", "stmpm_destroy", "stmpm_check", "stmpm_device_info",
           "stmpm_set_catalog", "stmpm_set_sensor", "stmpm_set_tuning", "stmpm_set_pipeline", "stmpm_stats_len", "stmpm_stats_finalize",
           "stmpm_trials_presampled", "stmpm_trials_presampled_ex", "stmpm_trials_rng", "stmpm_fill_presampled", "stmpm_run_presampled_host", "stmpm_run_presampled_host_ex",
           "stmpm_run_rng_host", "stmpm_run_rng_host_compact", "stmpm_run_rng_host_compact32", "stmpm_set_host_pipeline", "stmpm_host_copy_roofline", "stmpm_mc_converge", "stmpm_toolplot_failrate", "stmpm_column_hist", "stmpm_toolplot_reduce", "stmpm_fma_peak", "stmpm_work_counters", "stmpm_debug_box_muller", "stmpm_debug_box_muller_star", "stmpm_debug_box_muller_f32", "stmpm_launch_count")


class SensorCfg(C.Structure):
    _fields_ = [("f_ideal", C.c_double), ("half_u", C.c_double), ("half_v", C.c_double),
                ("n_min", C.c_int32), ("hist_log2_sub", C.c_int32), ("scan_pad_px", C.c_double)]


class Dist(C.Structure):
    _fields_ = [("mean", C.c_double * NNORMAL), ("sigma", C.c_double * NNORMAL), ("thermal_coef", C.c_double)]


class Summary(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("n_ok", "n_fail", "fail_rate", "mean", "std", "rms", "sigma_halfnormal",
                                          "mean_halfnormal", "p50", "p90", "p99", "p999", "x_max", "mean_ndet")]


class Toolplot(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("n", "rms", "n_kept", "kept_min", "kept_max", "mode", "sigma_about_mode")] + \
               [("mode_bin", C.c_int32), ("reserved", C.c_int32)]


class McResult(C.Structure):
    _fields_ = [(k, C.c_int64) for k in ("count", "trials", "converged", "rounds", "trials_computed")] + \
               [(k, C.c_double) for k in ("n_ok", "n_fail", "fail_rate", "sigma_halfnormal", "mean_halfnormal", "std_ratio")]


class ColumnStats(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("n", "min", "max", "mean", "std", "std_halfnormal")]


class StmpmError(RuntimeError):
    pass


_lib = None


def load():
    """Load libstmpm.so (built in-tree by startrackermpm_b200.build). Raises if it is missing: no fallback exists."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise StmpmError(f"{LIB_PATH} not found: build it with `python -m startrackermpm_b200.build` "
                         "(there is no CPU fallback for the trial loop)")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, u64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_double
    P = C.POINTER
    lib.stmpm_last_error.restype = C.c_char_p
    lib.stmpm_last_error.argtypes = []
    lib.stmpm_version.restype = C.c_int
    lib.stmpm_create.argtypes = [P(vp), C.c_int]
    lib.stmpm_destroy.argtypes = [vp]
    lib.stmpm_check.argtypes = [vp]
    lib.stmpm_device_info.argtypes = [vp, P(i32), P(i64), P(i32), P(i32)]
    lib.stmpm_set_catalog.argtypes = [vp, vp, vp, vp, i32, i32, vp, vp]
    lib.stmpm_set_sensor.argtypes = [vp, P(SensorCfg)]
    lib.stmpm_set_tuning.argtypes = [vp, i32, dbl, i32, i32]
    lib.stmpm_set_pipeline.argtypes = [vp, i32]
    lib.stmpm_stats_len.argtypes = [i32, P(i64)]
    lib.stmpm_stats_finalize.argtypes = [i32, vp, P(Summary)]
    lib.stmpm_trials_presampled.argtypes = [vp, C.c_int, i64, u64, i64, P(vp), vp, vp, vp, vp]
    lib.stmpm_trials_presampled_ex.argtypes = [vp, C.c_int, i64, u64, i64, P(vp), vp, vp, vp, vp, vp]
    lib.stmpm_run_presampled_host_ex.argtypes = [vp, C.c_int, i64, u64, i64, P(vp), vp, vp, vp, vp]
    lib.stmpm_trials_rng.argtypes = [vp, C.c_int, i64, i64, u64, i64, P(Dist), vp, vp, vp, vp]
    lib.stmpm_fill_presampled.argtypes = [vp, i64, u64, i64, P(Dist), P(vp), vp]
    lib.stmpm_run_presampled_host.argtypes = [vp, C.c_int, i64, u64, i64, P(vp), vp, vp, vp]
    lib.stmpm_run_rng_host.argtypes = [vp, C.c_int, i64, i64, u64, i64, P(Dist), vp, vp, vp]
    lib.stmpm_run_rng_host_compact.argtypes = [vp, C.c_int, i64, i64, u64, i64, P(Dist), vp, vp, vp]
    lib.stmpm_run_rng_host_compact32.argtypes = [vp, C.c_int, i64, i64, u64, i64, P(Dist), vp, vp, vp]
    lib.stmpm_set_host_pipeline.argtypes = [vp, i64]
    lib.stmpm_host_copy_roofline.argtypes = [vp, i64, i32, i32, vp, vp, P(dbl)]
    lib.stmpm_mc_converge.argtypes = [vp, C.c_int, P(Dist), i32, u64, i64, i64, dbl, i64, i64, P(McResult), vp]
    lib.stmpm_toolplot_failrate.argtypes = [vp, vp, vp, i64, vp, i32, vp, vp, vp]
    lib.stmpm_column_hist.argtypes = [vp, vp, i64, i32, P(ColumnStats), vp, vp]
    lib.stmpm_toolplot_reduce.argtypes = [vp, vp, i64, i32, P(Toolplot), vp, vp]
    lib.stmpm_fma_peak.argtypes = [vp, C.c_int, P(dbl)]
    lib.stmpm_work_counters.argtypes = [vp, i64, u64, i64, P(Dist), P(dbl)]
    lib.stmpm_debug_box_muller.argtypes = [vp, vp, i64, vp, vp]
    lib.stmpm_debug_box_muller_star.argtypes = [vp, vp, i64, vp, vp]
    lib.stmpm_debug_box_muller_f32.argtypes = [vp, vp, i64, vp, vp]
    lib.stmpm_launch_count.argtypes = [vp, P(i64)]
    for name in EXPORTS:
        if name not in ("stmpm_last_error",):
            getattr(lib, name).restype = C.c_int
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().stmpm_last_error()
        raise StmpmError(f"libstmpm error {rc}: {msg.decode() if msg else '?'}")
