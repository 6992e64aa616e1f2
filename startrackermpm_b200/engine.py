"""Bulk host API over the C ABI (include/stmpm.h): the call sites `simObjects.Simulation.run_sim` and bench.py use.

Everything numerical happens in libstmpm.so on the GPU.  numpy arrays carry host buffers, torch CUDA tensors carry
device buffers (`tensor.data_ptr()`); neither does any arithmetic of the trial.
"""
from __future__ import annotations

import ctypes as C
from typing import Mapping, Sequence

import numpy as np

from . import _abi
from .catalog import SkyGrid, build_sky_grid, load_catalog

COLUMNS = _abi.COLUMNS
NORMAL_PARAMS = _abi.NORMAL_PARAMS
SENSOR_WIDTH = 1024    # /root/reference/constants.py:97
SENSOR_HEIGHT = 1024   # /root/reference/constants.py:98
N_MIN = 3


def make_dist(means: Mapping[str, float] | Sequence[float], sigmas: Mapping[str, float] | Sequence[float],
              thermal_coef: float = 0.0) -> _abi.Dist:
    """Parameter distributions of native-RNG mode, in NORMAL_PARAMS order (docs/MODEL.md §6)."""
    if isinstance(means, Mapping):
        means = [float(means.get(k, 0.0)) for k in NORMAL_PARAMS]
    if isinstance(sigmas, Mapping):
        sigmas = [float(sigmas.get(k, 0.0)) for k in NORMAL_PARAMS]
    if len(means) != _abi.NNORMAL or len(sigmas) != _abi.NNORMAL:
        raise ValueError(f"need {_abi.NNORMAL} means and sigmas in order {NORMAL_PARAMS}")
    d = _abi.Dist()
    for i in range(_abi.NNORMAL):
        d.mean[i] = float(means[i])
        d.sigma[i] = float(sigmas[i])
    d.thermal_coef = float(thermal_coef)
    return d


def _ptr(a) -> int:
    """Address of a numpy array / torch tensor / None."""
    if a is None:
        return 0
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return int(a.data_ptr())


class TrialEngine:
    """One handle = one GPU + one catalog + one sensor description."""

    def __init__(self, device: int = 0, catalog=None, f_ideal: float = 3500.0, half_u: float = SENSOR_WIDTH / 2,
                 half_v: float = SENSOR_HEIGHT / 2, n_min: int = N_MIN, hist_log2_sub: int = 8,
                 cell_deg: float = 3.0, scan_pad_px: float = 0.0):
        self._lib = _abi.load()
        self._h = C.c_void_p()
        _abi.check(self._lib.stmpm_create(C.byref(self._h), int(device)))
        self.device = int(device)
        xyz, mag = catalog if catalog is not None else load_catalog()
        self.set_catalog(build_sky_grid(xyz, mag, cell_deg))
        self.set_sensor(f_ideal, half_u, half_v, n_min, hist_log2_sub, scan_pad_px)

    # ---------------------------------------------------------------- configuration
    def set_catalog(self, grid: SkyGrid) -> None:
        self.grid = grid
        xyz = np.ascontiguousarray(grid.xyz, dtype=np.float64)
        mag = np.ascontiguousarray(grid.mag, dtype=np.float32)
        sid = np.ascontiguousarray(grid.star_id, dtype=np.int32)
        n_ra = np.ascontiguousarray(grid.n_ra, dtype=np.int32)
        cs = np.ascontiguousarray(grid.cell_start, dtype=np.int32)
        _abi.check(self._lib.stmpm_set_catalog(self._h, _ptr(xyz), _ptr(mag), _ptr(sid), xyz.shape[0], grid.n_bands,
                                               _ptr(n_ra), _ptr(cs)))

    def set_sensor(self, f_ideal: float, half_u: float = SENSOR_WIDTH / 2, half_v: float = SENSOR_HEIGHT / 2,
                   n_min: int = N_MIN, hist_log2_sub: int = 8, scan_pad_px: float = 0.0) -> None:
        cfg = _abi.SensorCfg(float(f_ideal), float(half_u), float(half_v), int(n_min), int(hist_log2_sub),
                             float(scan_pad_px))
        _abi.check(self._lib.stmpm_set_sensor(self._h, C.byref(cfg)))
        self.f_ideal = float(f_ideal)
        self.hist_log2_sub = int(hist_log2_sub)
        self.stats_len = stats_len(hist_log2_sub)

    def set_tuning(self, window: int = 0, key_radius_deg: float = -1.0, stage_columns: bool = False, key_table: int = -1) -> None:
        """Performance knobs (results never depend on them): CTA sort window in trials (0 = default), work-key disc
        radius in degrees (< 0 = default), sorted parameter staging of pre-sampled columns (default off: measured slower),
        work-key table (-1 = per mode, 0 = disc, 1 = array footprint).  Replaces round 1's STMPM_WIN /
        STMPM_KEY_RADIUS_DEG environment variables."""
        _abi.check(self._lib.stmpm_set_tuning(self._h, int(window), float(key_radius_deg), int(bool(stage_columns)),
                                              int(key_table)))

    def set_pipeline(self, pipeline: int = -1) -> None:
        """-1 auto, 0 single persistent kernel, 1 two-kernel split pipeline (pre-sampled, one segment); records identical."""
        _abi.check(self._lib.stmpm_set_pipeline(self._h, int(pipeline)))

    def check(self) -> None:
        """After synchronising a stream handed to the *_device calls: raises if a kernel hit its scheduler watchdog."""
        _abi.check(self._lib.stmpm_check(self._h))

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.stmpm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---------------------------------------------------------------- info
    def device_info(self) -> dict:
        sm, smem, skhz, mkhz = C.c_int32(), C.c_int64(), C.c_int32(), C.c_int32()
        _abi.check(self._lib.stmpm_device_info(self._h, C.byref(sm), C.byref(smem), C.byref(skhz), C.byref(mkhz)))
        return {"sm_count": sm.value, "smem_per_block": smem.value, "sm_khz": skhz.value, "mem_khz": mkhz.value}

    def launch_count(self) -> int:
        n = C.c_int64()
        _abi.check(self._lib.stmpm_launch_count(self._h, C.byref(n)))
        return int(n.value)

    def fma_peak(self, precision: int = 64) -> float:
        t = C.c_double()
        _abi.check(self._lib.stmpm_fma_peak(self._h, int(precision), C.byref(t)))
        return float(t.value)

    def work_counters(self, dist: _abi.Dist, n: int = 1 << 20, seed: int = 1, trial0: int = 0) -> dict:
        """Per-trial means of the instrumented kernel: stars visited in the sky grid, magnitude passes, pre-filter
        survivors, detected stars (DESIGN.md §4 work model)."""
        out = (C.c_double * 4)()
        _abi.check(self._lib.stmpm_work_counters(self._h, int(n), int(seed), int(trial0), C.byref(dist), out))
        return {"visits": out[0], "mag_pass": out[1], "survivors": out[2], "detected": out[3]}

    def _check_dev(self, t, dtype, numel: int, name: str, optional: bool = False) -> None:
        """A device buffer handed to the C ABI must be a contiguous CUDA tensor of the right dtype and size on THIS engine's
        device: the library writes numel * itemsize bytes through the raw pointer."""
        if t is None:
            if optional:
                return
            raise ValueError(f"{name} is required")
        if isinstance(t, np.ndarray) or not getattr(t, "is_cuda", False):
            raise ValueError(f"{name} must be a CUDA tensor (host buffers go through run_presampled / run_rng)")
        if t.device.index != self.device:
            raise ValueError(f"{name} lives on cuda:{t.device.index}, the engine on cuda:{self.device}")
        if t.dtype != dtype or t.numel() != int(numel) or not t.is_contiguous():
            raise ValueError(f"{name} must be a contiguous {dtype} tensor of {int(numel)} elements "
                             f"(got {t.dtype}, {t.numel()} elements, contiguous={t.is_contiguous()})")

    # ---------------------------------------------------------------- pre-sampled mode
    @staticmethod
    def _col_ptrs(cols, n_expected=None):
        """cols: [13, n] array/tensor (rows = COLUMNS) or a sequence of 13 1-D float64 arrays/tensors."""
        rows = [cols[i] for i in range(_abi.NCOLS)]
        n = int(rows[0].shape[0])
        for r in rows:
            if int(r.shape[0]) != n or r.dtype not in (np.float64, _torch_f64()):
                raise ValueError("every column must be float64 with the same length")
            if isinstance(r, np.ndarray):
                if not r.flags["C_CONTIGUOUS"]:
                    raise ValueError("columns must be contiguous")
            elif not r.is_contiguous():
                raise ValueError("columns must be contiguous")
        arr = (C.c_void_p * _abi.NCOLS)(*[_ptr(r) for r in rows])
        return arr, n

    def run_presampled(self, cols, seed: int, trial0: int = 0, precision: int = 64, want_stats: bool = False,
                       out_err: np.ndarray | None = None, out_ndet: np.ndarray | None = None, want_maxmag: bool = False):
        """HOST buffers in, HOST buffers out (H2D, kernel, D2H inside the call) — `Simulation.run_sim`'s path.
        cols: float64 [13, n] in COLUMNS order.  Returns (err_arcsec f64 [n] with -1 = failed, n_det i32 [n][, stats])."""
        if not isinstance(cols, np.ndarray) and not isinstance(cols, (list, tuple)):
            cols = cols.numpy()
        if isinstance(cols, np.ndarray):
            cols = np.ascontiguousarray(cols, dtype=np.float64)
        ptrs, n = self._col_ptrs(cols)
        _check_host_out(out_err, np.float64, n, "out_err")
        _check_host_out(out_ndet, np.int32, n, "out_ndet")
        err = out_err if out_err is not None else np.empty(n, dtype=np.float64)
        ndet = out_ndet if out_ndet is not None else np.empty(n, dtype=np.int32)
        stats = np.zeros(self.stats_len, dtype=np.float64) if want_stats else None
        maxmag = np.empty(n, dtype=np.float32) if want_maxmag else None
        _abi.check(self._lib.stmpm_run_presampled_host_ex(self._h, int(precision), n, int(seed) & (2**64 - 1),
                                                          int(trial0), ptrs, _ptr(err), _ptr(ndet), _ptr(maxmag),
                                                          _ptr(stats)))
        out = (err, ndet)
        if want_maxmag:
            out += (maxmag,)       # faintest detected star's catalog magnitude, -inf if none (star_mag objective)
        if want_stats:
            out += (stats,)
        return out

    def run_presampled_device(self, cols, err, ndet, stats=None, *, seed: int, trial0: int = 0, precision: int = 64,
                              stream: int = 0) -> None:
        """DEVICE buffers (torch CUDA tensors), stream-ordered, asynchronous."""
        ptrs, n = self._col_ptrs(cols)
        for r in (cols[i] for i in range(_abi.NCOLS)):
            self._check_dev(r, _torch_f64(), n, "cols")
        self._check_dev(err, _torch_f64(), n, "err", optional=True)
        self._check_dev(ndet, _torch_dtype("int32"), n, "ndet", optional=True)
        self._check_dev(stats, _torch_f64(), self.stats_len, "stats", optional=True)
        _abi.check(self._lib.stmpm_trials_presampled(self._h, int(precision), n, int(seed) & (2**64 - 1), int(trial0),
                                                     ptrs, _ptr(err), _ptr(ndet), _ptr(stats), int(stream)))

    def fill_presampled_device(self, cols, dist: _abi.Dist, *, seed: int, trial0: int = 0, stream: int = 0) -> None:
        """The seeded generator kernel of config 2: fills 13 device columns with native-RNG mode's draws."""
        ptrs, n = self._col_ptrs(cols)
        for r in (cols[i] for i in range(_abi.NCOLS)):
            self._check_dev(r, _torch_f64(), n, "cols")
        _abi.check(self._lib.stmpm_fill_presampled(self._h, n, int(seed) & (2**64 - 1), int(trial0), C.byref(dist), ptrs,
                                                   int(stream)))

    # ---------------------------------------------------------------- native-RNG mode
    @staticmethod
    def _dist_array(dists):
        if isinstance(dists, _abi.Dist):
            dists = [dists]
        arr = (_abi.Dist * len(dists))(*dists)
        return arr, len(dists)

    def run_rng(self, dists, trials_per_segment: int, seed: int, trial0: int = 0, precision: int = 64,
                records: bool = False, out_err=None, out_ndet=None, compact=False):
        """HOST results.  Returns stats [n_segments, stats_len] (and err, n_det [n_segments*trials_per_segment] if records).
        out_err / out_ndet: caller-owned (e.g. pinned) record buffers; giving them implies records.
        compact: False = (float64 err, int32 n_det), 12 bytes per trial over PCIe; True or 9 = n_det as uint8 saturated at
        255 (9 bytes); 5 = the error as float32 too (5 bytes; the statistics still come from the float64 error)."""
        arr, ns = self._dist_array(dists)
        n = ns * int(trials_per_segment)
        if (out_err is None) != (out_ndet is None):
            raise ValueError("out_err and out_ndet go together")
        fmt = {False: 12, True: 9, 12: 12, 9: 9, 5: 5}.get(compact)
        if fmt is None:
            raise ValueError("compact must be False, True / 9, or 5")
        nd_type = np.int32 if fmt == 12 else np.uint8
        err_type = np.float32 if fmt == 5 else np.float64
        if out_err is not None:
            _check_host_out(out_err, err_type, n, "out_err")
            _check_host_out(out_ndet, nd_type, n, "out_ndet")
            records = True
        err = out_err if out_err is not None else (np.empty(n, dtype=err_type) if records else None)
        ndet = out_ndet if out_ndet is not None else (np.empty(n, dtype=nd_type) if records else None)
        stats = np.zeros((ns, self.stats_len), dtype=np.float64)
        fn = {12: self._lib.stmpm_run_rng_host, 9: self._lib.stmpm_run_rng_host_compact, 5: self._lib.stmpm_run_rng_host_compact32}[fmt]
        _abi.check(fn(self._h, int(precision), ns, int(trials_per_segment), int(seed) & (2**64 - 1), int(trial0), arr,
                      _ptr(err), _ptr(ndet), _ptr(stats)))
        return (stats, err, ndet) if records else stats

    def set_host_pipeline(self, chunk_trials: int = 0) -> None:
        """Trials per stage of the host pipelines (H2D / kernel / D2H overlap); 0 = default 2^22."""
        _abi.check(self._lib.stmpm_set_host_pipeline(self._h, int(chunk_trials)))

    def host_copy_roofline(self, n: int, h2d_bytes_per_trial: int, d2h_bytes_per_trial: int, src=None, dst=None) -> float:
        """Seconds the host pipeline needs for the COPIES alone (same chunking, streams, staging buffers; no kernel):
        the host-side roofline bench.py quotes e2e against.  src / dst: pinned host arrays of n * bytes_per_trial bytes."""
        sec = C.c_double()
        _abi.check(self._lib.stmpm_host_copy_roofline(self._h, int(n), int(h2d_bytes_per_trial), int(d2h_bytes_per_trial),
                                                      _ptr(src), _ptr(dst), C.byref(sec)))
        return float(sec.value)

    def run_rng_device(self, dists, trials_per_segment: int, stats, err=None, ndet=None, *, seed: int, trial0: int = 0,
                       precision: int = 64, stream: int = 0) -> None:
        """DEVICE buffers, stream-ordered, asynchronous. stats (+=) must be zeroed by the caller: [n_segments, stats_len]."""
        arr, ns = self._dist_array(dists)
        n = ns * int(trials_per_segment)
        self._check_dev(stats, _torch_f64(), ns * self.stats_len, "stats", optional=True)
        self._check_dev(err, _torch_f64(), n, "err", optional=True)
        self._check_dev(ndet, _torch_dtype("int32"), n, "ndet", optional=True)
        _abi.check(self._lib.stmpm_trials_rng(self._h, int(precision), ns, int(trials_per_segment),
                                              int(seed) & (2**64 - 1), int(trial0), arr, _ptr(err), _ptr(ndet),
                                              _ptr(stats), int(stream)))

    def toolplot_reduce(self, err, bins: int = 50, stream: int = 0) -> dict:
        """toolplot.py:82-99 on the device: `err` is a float64 CUDA tensor of per-trial errors [arcsec] (-1 = failed).
        Returns rms, n_kept (<= 6 rms), mode (left edge of the fullest of `bins` bins), sigma_about_mode, density[bins]."""
        n = int(err.shape[0])
        out = _abi.Toolplot()
        dens = np.empty(int(bins), dtype=np.float64)
        _abi.check(self._lib.stmpm_toolplot_reduce(self._h, _ptr(err), n, int(bins), C.byref(out), _ptr(dens), int(stream)))
        d = {k: getattr(out, k) for k, _ in _abi.Toolplot._fields_ if k != "reserved"}
        d["density"] = dens
        return d

    def mc_converge(self, dist: _abi.Dist, *, batch: int = 100, seed: int = 0, trial0: int = 0, min_rows: int = 10_000,
                    tol: float = 1e-5, max_runs: int = -1, max_batches: int = 0, precision: int = 64) -> dict:
        """MonteCarlo.run_sim's convergence loop (monte_carlo.py:45-76) with the rule evaluated on the device after every
        batch; one host synchronisation per round (a default run is one round).  Returns count, trials, sigma_halfnormal, ... and
        the packed statistics of the accepted batches."""
        res = _abi.McResult()
        stats = np.zeros(self.stats_len, dtype=np.float64)
        _abi.check(self._lib.stmpm_mc_converge(self._h, int(precision), C.byref(dist), int(batch), int(seed) & (2**64 - 1),
                                               int(trial0), int(min_rows), float(tol), int(max_runs), int(max_batches),
                                               C.byref(res), _ptr(stats)))
        out = {k: getattr(res, k) for k, _ in _abi.McResult._fields_}
        out["stats"] = stats
        return out

    def toolplot_failrate(self, err, max_magnitude, levels, stream: int = 0) -> dict:
        """toolplot.py:170-187 (mode 's') on the device: err / max_magnitude are float64 CUDA tensors [n] (failures kept as
        -1), levels the sorted unique magnitudes.  Returns rows, failed, rate_percent per level (+ unmatched rows)."""
        n = int(err.shape[0])
        self._check_dev(err, _torch_f64(), n, "err")
        self._check_dev(max_magnitude, _torch_f64(), n, "max_magnitude")
        lev = np.ascontiguousarray(levels, dtype=np.float64)
        rows, failed = np.empty(lev.size + 1), np.empty(lev.size + 1)
        _abi.check(self._lib.stmpm_toolplot_failrate(self._h, _ptr(err), _ptr(max_magnitude), n, _ptr(lev), lev.size,
                                                     _ptr(rows), _ptr(failed), int(stream)))
        with np.errstate(invalid="ignore", divide="ignore"):
            rate = 100.0 * failed[:-1] / rows[:-1]
        return {"levels": lev, "rows": rows[:-1], "failed": failed[:-1], "rate_percent": rate, "unmatched_rows": rows[-1]}

    def column_hist(self, x, bins: int = 50, stream: int = 0) -> dict:
        """toolplot.py:240-256 (mode 'c') on the device: x = the star_mag objective's float32 CUDA column."""
        n = int(x.shape[0])
        self._check_dev(x, _torch_dtype("float32"), n, "x")
        out = _abi.ColumnStats()
        dens = np.empty(int(bins), dtype=np.float64)
        _abi.check(self._lib.stmpm_column_hist(self._h, _ptr(x), n, int(bins), C.byref(out), _ptr(dens), int(stream)))
        d = {k: getattr(out, k) for k, _ in _abi.ColumnStats._fields_}
        d["density"] = dens
        return d

    # ---------------------------------------------------------------- statistics
    def summarize(self, stats: np.ndarray) -> dict:
        return summarize(stats, self.hist_log2_sub)


def debug_box_muller(wa: np.ndarray, wb: np.ndarray, star: bool = False):
    """Host build of the kernels' table-based Box-Muller (diagnostic; no GPU needed): the parameter sampler's full-accuracy
    draw, or (star=True) the star loop's shorter evaluation of the same draw."""
    wa = np.ascontiguousarray(wa, dtype=np.uint32)
    wb = np.ascontiguousarray(wb, dtype=np.uint32)
    z1, z2 = np.empty(wa.size), np.empty(wa.size)
    fn = _abi.load().stmpm_debug_box_muller_star if star else _abi.load().stmpm_debug_box_muller
    _abi.check(fn(_ptr(wa), _ptr(wb), wa.size, _ptr(z1), _ptr(z2)))
    return z1, z2


def stats_len(hist_log2_sub: int = 8) -> int:
    n = C.c_int64()
    _abi.check(_abi.load().stmpm_stats_len(int(hist_log2_sub), C.byref(n)))
    return int(n.value)


def summarize(stats: np.ndarray, hist_log2_sub: int = 8) -> dict:
    """Final reductions (driver.py:321-324; toolplot.py:82) of one packed stats buffer — host-side C, no GPU needed."""
    stats = np.ascontiguousarray(stats, dtype=np.float64).reshape(-1)
    if stats.shape[0] != stats_len(hist_log2_sub):
        raise ValueError(f"packed stats buffer must have {stats_len(hist_log2_sub)} doubles")
    s = _abi.Summary()
    _abi.check(_abi.load().stmpm_stats_finalize(int(hist_log2_sub), _ptr(stats), C.byref(s)))
    return {k: getattr(s, k) for k, _ in _abi.Summary._fields_}


def _check_host_out(a, dtype, n: int, name: str) -> None:
    if a is None:
        return
    if not isinstance(a, np.ndarray) or a.dtype != dtype or a.size != int(n) or not a.flags.c_contiguous \
            or not a.flags.writeable:
        raise ValueError(f"{name} must be a writeable C-contiguous numpy {np.dtype(dtype).name}[{int(n)}] array")


def _torch_dtype(name: str):
    import torch
    return getattr(torch, name)


def _torch_f64():
    try:
        import torch
        return torch.float64
    except Exception:  # torch absent: only numpy host buffers are usable
        return None
