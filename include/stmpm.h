/* stmpm.h — C ABI of libstmpm.so: the B200-native (sm_100a) Monte Carlo trial loop of StarTrackerMPM.
 *
 * The reference has no FFI: its hot path is the pure-Python `simObjects` package (absent from the mount).
 * Each entry point below cites the reference interface it replaces (file:line in /root/reference) — the call sites that
 * pin the behaviour.  Trial semantics are frozen in docs/MODEL.md.
 *
 * Conventions: every function returns 0 on success, non-zero on failure (a cudaError_t value, or STMPM_E_* below) and
 * stmpm_last_error() then holds a message.  `stream` is a cudaStream_t passed as void* (NULL = default stream); calls
 * are stream-ordered.  Pointers named *_dev are device pointers (e.g. torch.Tensor.data_ptr()), *_host host pointers.
 * No torch / C++ types cross this boundary.  There is no CPU fallback anywhere behind it.
 */
#ifndef STMPM_H
#define STMPM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STMPM_E_ARG      10001   /* bad argument */
#define STMPM_E_STATE    10002   /* catalog / sensor not set */
#define STMPM_E_NOGPU    10003   /* no CUDA device */
#define STMPM_E_WATCHDOG 10004   /* a trial kernel's internal scheduler wait exceeded 20 s of wall time: results incomplete */

#define STMPM_NCOLS      13      /* pre-sampled input columns, order below */
#define STMPM_NNORMAL    11      /* normal-distributed parameters in native-RNG mode */
#define STMPM_STATS_HEAD 8       /* [n_ok, n_fail, sum_x, sum_x2, sum_ndet, x_max, under, over] then the histogram */
#define STMPM_HIST_OCTAVES 34    /* x in [2^-14, 2^20) arcsec */
#define STMPM_HIST_MIN_EXP (-14)

/* Column order of pre-sampled inputs == DataFrame columns the reference builds in MonteCarlo.create_data
 * (monte_carlo.py:85-105) from StarTracker.randomize / Software.randomize (monte_carlo.py:93-94); names pinned by
 * run.sh:11-17:  RIGHT_ASCENSION, DECLINATION, ROLL [rad], FOCAL_LENGTH [px], F_ARR_EPS_X/Y/Z [px],
 * F_ARR_PHI/THETA/PSI [deg], BASE_DEV_X/Y [px], MAX_MAGNITUDE. */
enum { STMPM_COL_RA = 0, STMPM_COL_DEC, STMPM_COL_ROLL, STMPM_COL_F, STMPM_COL_EPS_X, STMPM_COL_EPS_Y, STMPM_COL_EPS_Z,
       STMPM_COL_PHI, STMPM_COL_THETA, STMPM_COL_PSI, STMPM_COL_DEV_X, STMPM_COL_DEV_Y, STMPM_COL_MAX_MAG };

typedef struct stmpm_handle stmpm_t;

/* Software-side camera knowledge + detection rule.  Replaces the attributes quest_objective reads from
 * Simulation.camera (StarTracker.f_len.ideal: sensitivity.py:118; SENSOR_WIDTH/HEIGHT: constants.py:97-98). */
typedef struct {
    double  f_ideal;      /* px; back-projection focal length (MODEL.md §4) */
    double  half_u;       /* SENSOR_WIDTH / 2  = 512 */
    double  half_v;       /* SENSOR_HEIGHT / 2 = 512 */
    int32_t n_min;        /* minimum detected stars for a solve (3) */
    int32_t hist_log2_sub;/* histogram sub-bins per octave = 2^k, 4 <= k <= 8 (8 -> 8704 bins) */
    double  scan_pad_px;  /* candidate lists cover the ideal array grown by this many px (0 = default 32); trials
                             whose perturbations + noise exceed it still run, on the slower sky-grid scan */
} stmpm_sensor_cfg;

/* Parameter distributions for native-RNG mode (MODEL.md §6): normal(mean, sigma) in the order
 * FOCAL_LENGTH, F_ARR_EPS_X, _Y, _Z, F_ARR_PHI, _THETA, _PSI, BASE_DEV_X, _Y, MAX_MAGNITUDE, D_TEMP.
 * Replaces Parameter(.ideal,.stddev) sampling in StarTracker/Software/Orbit.randomize (monte_carlo.py:93-95;
 * sensitivity.py:43-44) and the thermal focal update (sensitivity.py:115-118). */
typedef struct {
    double mean[STMPM_NNORMAL];
    double sigma[STMPM_NNORMAL];
    double thermal_coef;  /* FOCAL_THERMAL_COEFFICIENT */
} stmpm_dist;

/* Final reductions: driver.py:321-324 (failure rate, half-normal sigma/mean), toolplot.py:82 (RMS). */
typedef struct {
    double n_ok, n_fail, fail_rate;
    double mean, std, rms;
    double sigma_halfnormal, mean_halfnormal;
    double p50, p90, p99, p999, x_max;
    double mean_ndet;
} stmpm_summary;

/* toolplot.py:82-99 (mode 'm'): RMS about 0, 6-sigma clip, bins-bin density histogram over [min, max] of the kept values,
 * mode = left edge of the fullest bin, sigma about the mode. */
typedef struct {
    double n, rms, n_kept, kept_min, kept_max, mode, sigma_about_mode;
    int32_t mode_bin, reserved;
} stmpm_toolplot;

/* Result of the on-device MonteCarlo.run_sim loop (monte_carlo.py:30-80). */
typedef struct {
    int64_t count;            /* batches run == MonteCarlo.count (monte_carlo.py:50,79) */
    int64_t trials;           /* count * batch */
    int64_t converged;        /* 1 = the stop rule fired, 0 = max_batches reached first */
    int64_t rounds;           /* host synchronisations it took (1 for a default run) */
    int64_t trials_computed;  /* trials the device evaluated, speculative ones after the stopping batch included */
    double  n_ok, n_fail, fail_rate;
    double  sigma_halfnormal; /* monte_carlo.py:61 `prev` after the last batch == driver.py:324 true_std */
    double  mean_halfnormal;  /* driver.py:323 */
    double  std_ratio;        /* monte_carlo.py:60 after the last batch */
} stmpm_mc_result;

/* toolplot.py:240-256 (mode 'c') statistics of a magnitude column. */
typedef struct {
    double n, min, max, mean, std, std_halfnormal;
} stmpm_column_stats;

const char* stmpm_last_error(void);
int stmpm_version(void);

int stmpm_create(stmpm_t** h, int device);
int stmpm_destroy(stmpm_t* h);
/* After the caller has synchronised a stream it passed to the asynchronous entry points: 0, or STMPM_E_WATCHDOG if a kernel of
 * this handle gave up on its internal scheduler (the host-buffer entry points below check this themselves). */
int stmpm_check(stmpm_t* h);
/* sm_count, max dynamic shared memory per block, SM clock kHz, memory clock kHz */
int stmpm_device_info(stmpm_t* h, int32_t* sm_count, int64_t* smem_per_block, int32_t* sm_khz, int32_t* mem_khz);

/* Catalog sorted by sky-grid cell (host pointers; uploaded once).  n_stars is bounded by the shared-memory staging of the
 * fp32 catalog: (227 KB - fixed regions) / 16 B = about 9 250 stars on B200 (checked here, with the exact limit in the
 * error message); n_bands <= 256.  xyz is S x 3 row-major float64 unit vectors,
 * star_id the Philox key of each row (MODEL.md §5).  Band b (n_bands equal declination bands from -pi/2) has n_ra[b]
 * equal RA cells; cell_start has sum(n_ra)+1 entries.  Replaces the BSC5 catalog the reference loads from
 * utils/BSC5PKL.pkl (constants.py:32-33). */
int stmpm_set_catalog(stmpm_t* h, const double* xyz_host, const float* mag_host, const int32_t* star_id_host,
                      int32_t n_stars, int32_t n_bands, const int32_t* n_ra_host, const int32_t* cell_start_host);
int stmpm_set_sensor(stmpm_t* h, const stmpm_sensor_cfg* cfg);
/* Performance knobs of the trial kernel; results never depend on them (tests/test_gpu_parity.py checks bit-identity).
 * window: trials per CTA sort window (multiple of 32, <= 2048; 0 = the measured default per mode).  key_radius_deg: radius
 * of the work-key disc (< 0 = default, 0 = whole candidate list); changing it synchronises the device and rebuilds the
 * key table.  stage_columns: non-zero = pre-sampled launches re-write each sort window's columns in key order to an
 * L2-resident scratch (allocated per launch in stream order) and read them coalesced; 0 (default: staging measured
 * slower) = gather them.  key_table: -1 (default) = disc table for pre-sampled launches, array-footprint
 * table for native-RNG launches; 0 / 1 force the disc / footprint table.
 * Not to be called while launches on this handle are in flight. */
int stmpm_set_tuning(stmpm_t* h, int32_t window, double key_radius_deg, int32_t stage_columns, int32_t key_table);

/* Pre-sampled launches of one segment run either as ONE persistent kernel or as the two-kernel split pipeline (candidate scan
 * -> counting sort of the trials by survivor count -> exact star loop + solve over warps of identical star counts; DESIGN.md
 * §4).  The split measured SLOWER than the single kernel (the duplicated set-up and the record traffic outweigh the idle
 * lanes it removes), so -1 (default) and 0 = the single kernel; 1 = the split, kept as a tested experiment (chunks it cannot
 * hand over — star-rich / wide-cone workloads — fall back to the single kernel, decided on the device).  Records agree to
 * the last bits (1e-13 rad). */
int stmpm_set_pipeline(stmpm_t* h, int32_t pipeline);

/* Packed statistics buffer (MODEL.md §9): STMPM_STATS_HEAD doubles then 34 * 2^hist_log2_sub histogram bins; every entry
 * but x_max is a sum, so ranks merge with one allreduce(sum).  Host-only helpers (no GPU needed). */
int stmpm_stats_len(int32_t hist_log2_sub, int64_t* n_doubles);
int stmpm_stats_finalize(int32_t hist_log2_sub, const double* packed_host, stmpm_summary* out);

/* (1) Pre-sampled mode == Simulation.run_sim(params, obj_func=quest_objective[, trueData]) over n DataFrame rows
 * (monte_carlo.py:52; sensitivity.py:77).  cols_dev: STMPM_NCOLS device pointers to n float64 each.
 * err_arcsec_dev[n] = QUATERNION_ERROR_[arcsec] (toolplot.py:71), -1 = failed trial; n_det_dev[n] = detected stars.
 * stats_dev (optional) is accumulated into (+=), not cleared.  precision = 64 or 32. */
int stmpm_trials_presampled(stmpm_t* h, int precision, int64_t n, uint64_t seed, int64_t trial0,
                            const double* const* cols_dev, double* err_arcsec_dev, int32_t* n_det_dev,
                            double* stats_dev, void* stream);

/* Same, plus the star_mag objective's output (driver.py:246-247, column 'MAXIMUM MAGNITUDE VISIBLE', toolplot.py:241):
 * maxmag_dev[n] (optional) = catalog magnitude of the faintest detected star of each trial, -inf if none was detected. */
int stmpm_trials_presampled_ex(stmpm_t* h, int precision, int64_t n, uint64_t seed, int64_t trial0,
                               const double* const* cols_dev, double* err_arcsec_dev, int32_t* n_det_dev,
                               float* maxmag_dev, double* stats_dev, void* stream);

/* (2) Native-RNG mode == MonteCarlo.create_data + Simulation.run_sim fused (monte_carlo.py:48-58): parameters are drawn
 * on the device from Philox streams keyed by (seed, global trial index).  n_segments independent parameter sets (a
 * sensitivity grid point / an orbit step: sensitivity.py:39-55) of trials_per_segment trials each; segment s uses
 * dists_host[s], global trial indices trial0 + s*trials_per_segment + i, and accumulates into
 * stats_dev[s*stats_len ...].  Records (optional, may be NULL) are n_segments*trials_per_segment long.
 * dists_host is consumed before the call returns; the device copy is allocated and freed in stream order per call
 * (cudaMallocAsync), so calls on different streams of one handle may overlap. */
int stmpm_trials_rng(stmpm_t* h, int precision, int64_t n_segments, int64_t trials_per_segment, uint64_t seed,
                     int64_t trial0, const stmpm_dist* dists_host, double* err_arcsec_dev, int32_t* n_det_dev,
                     double* stats_dev, void* stream);

/* The seeded generator of config 2: fills STMPM_NCOLS device columns with exactly the parameters native-RNG mode draws
 * for trials trial0 .. trial0+n-1 (so presampled(fill(...)) == rng(...) trial for trial). */
int stmpm_fill_presampled(stmpm_t* h, int64_t n, uint64_t seed, int64_t trial0, const stmpm_dist* dist_host,
                          double* const* cols_dev, void* stream);

/* Host-buffer entry points (the call Simulation.run_sim makes): chunked, double-buffered H2D -> kernel -> D2H on
 * internal streams; returns after the results are in the host buffers.  Host buffers should be pinned for full PCIe
 * rate (pageable works, slower).  stats_host (optional) receives the packed statistics of this call. */
int stmpm_run_presampled_host(stmpm_t* h, int precision, int64_t n, uint64_t seed, int64_t trial0,
                              const double* const* cols_host, double* err_arcsec_host, int32_t* n_det_host,
                              double* stats_host);
int stmpm_run_presampled_host_ex(stmpm_t* h, int precision, int64_t n, uint64_t seed, int64_t trial0,
                                 const double* const* cols_host, double* err_arcsec_host, int32_t* n_det_host,
                                 float* maxmag_host, double* stats_host);
int stmpm_run_rng_host(stmpm_t* h, int precision, int64_t n_segments, int64_t trials_per_segment, uint64_t seed,
                       int64_t trial0, const stmpm_dist* dists_host, double* err_arcsec_host, int32_t* n_det_host,
                       double* stats_host);

/* stmpm_run_rng_host with compact records: n_det as uint8 saturated at 255 (9 bytes per trial instead of 12 come back over
 * PCIe; the default sensor detects ~20 stars per trial, MAX_MAGNITUDE = 20 ~60). */
int stmpm_run_rng_host_compact(stmpm_t* h, int precision, int64_t n_segments, int64_t trials_per_segment, uint64_t seed,
                               int64_t trial0, const stmpm_dist* dists_host, double* err_arcsec_host, uint8_t* n_det_u8_host,
                               double* stats_host);
/* ... and with 5-byte records: the error rounded to float32 as well.  The statistics are still accumulated from the float64
 * error on the device; only the per-trial record loses precision (half an ulp of a float32 number of arcseconds: 3e-5" =
 * 1.5e-10 rad at 1000", inside the 1e-9 rad parity bar up to 0.9 degrees of error).  For hosts whose device-to-host fabric is
 * the limit (94 GB/s for eight B200s on this pool): records per second scale with 1 / bytes. */
int stmpm_run_rng_host_compact32(stmpm_t* h, int precision, int64_t n_segments, int64_t trials_per_segment, uint64_t seed,
                                 int64_t trial0, const stmpm_dist* dists_host, float* err_arcsec_f32_host, uint8_t* n_det_u8_host,
                                 double* stats_host);
/* Trials per stage of the host pipelines (0 = default 2^22; staging buffers are re-allocated on the next host call). */
int stmpm_set_host_pipeline(stmpm_t* h, int64_t chunk_trials);
/* Host-side roofline of the host-buffer entry points: n trials' worth of bytes through the SAME chunking, streams and
 * staging buffers, copies only, no kernel.  src_host: n * h2d_bytes_per_trial bytes (pinned), dst_host:
 * n * d2h_bytes_per_trial bytes.  seconds: wall time of the pipeline.  (bench.py reports e2e against this.) */
int stmpm_host_copy_roofline(stmpm_t* h, int64_t n, int32_t h2d_bytes_per_trial, int32_t d2h_bytes_per_trial,
                             const void* src_host, void* dst_host, double* seconds);

/* MonteCarlo.run_sim's convergence loop on the device (monte_carlo.py:45-76): batches of `batch`
 * native-RNG trials (global trial indices trial0 + k*batch + i) until the successful rows number at least min_rows AND
 * |prev - sigma_hat| / prev <= tol (sigma_hat = std(ddof=1)/sqrt(1-2/pi) over all successful trials so far, prev starts
 * at 1), or successful rows >= max_runs > 0, or max_batches (> 0) batches ran.  The rule is evaluated after EVERY batch,
 * on the device; the host reads the state once per round (round 1 = twice the batches min_rows needs, then x4).  stats_host (optional)
 * receives the packed statistics of exactly the batches the loop accepted.  The reference's defaults: batch 100
 * (driver.py:21,74), min_rows 10000, tol 1e-5 (monte_carlo.py:48), max_runs = driver.py's -n. */
int stmpm_mc_converge(stmpm_t* h, int precision, const stmpm_dist* dist_host, int32_t batch, uint64_t seed, int64_t trial0,
                      int64_t min_rows, double tol, int64_t max_runs, int64_t max_batches, stmpm_mc_result* out,
                      double* stats_host);

/* toolplot.py:170-187 (mode 's'): "Quantity-Based Star Failure Rate" per unique MAX_MAGNITUDE value, from n rows of
 * (error with the -1 failure sentinel kept, MAX_MAGNITUDE), after the 6-sigma clip of toolplot.py:82-83.  levels_host:
 * the n_levels sorted unique magnitudes; rows_host / failed_host [n_levels + 1]: rows and failed rows per level (last
 * slot: rows matching no level).  rate[%] = 100 * failed / rows. */
int stmpm_toolplot_failrate(stmpm_t* h, const double* err_arcsec_dev, const double* max_magnitude_dev, int64_t n,
                            const double* levels_host, int32_t n_levels, double* rows_host, double* failed_host,
                            void* stream);
/* toolplot.py:240-256 (mode 'c'): min / max / std of the 'MAXIMUM MAGNITUDE VISIBLE' column (the star_mag objective's
 * float32 output; -inf = no star detected, skipped) and its `bins`-bin density histogram, np.histogram-identical. */
int stmpm_column_hist(stmpm_t* h, const float* x_dev, int64_t n, int32_t bins, stmpm_column_stats* out,
                      double* density_host, void* stream);

/* toolplot.py's reductions on the device from n per-trial error records (err < 0 = failed trial = a row the reference
 * dropped).  density_host (optional) receives the `bins` values of np.histogram(kept, bins, density=True)[0]. */
int stmpm_toolplot_reduce(stmpm_t* h, const double* err_arcsec_dev, int64_t n, int32_t bins, stmpm_toolplot* out,
                          double* density_host, void* stream);

/* FMA-chain microbenchmark: measured FP64 / FP32 FMA peak of this GPU in TFLOP/s (roofline denominator). */
int stmpm_fma_peak(stmpm_t* h, int precision, double* tflops);
/* Work-model calibration: runs n native-RNG trials through an instrumented build of the trial kernel and
 * returns per-trial means [sky-grid stars visited, magnitude passes, pre-filter survivors, detected stars]. */
int stmpm_work_counters(stmpm_t* h, int64_t n, uint64_t seed, int64_t trial0, const stmpm_dist* dist_host,
                        double* per_trial4_host);
/* Diagnostic, host-only (no GPU): the table-based Box-Muller (MODEL.md §5) of the parameter sampler, compiled for the host
 * from the same source the kernels use, so its accuracy (|dz| < 1e-12) can be checked against libm without a GPU. */
int stmpm_debug_box_muller(const uint32_t* wa, const uint32_t* wb, int64_t n, double* z1, double* z2);
/* Same for the star loop's shorter evaluation of that draw (centroid noise, fp64 mode): |dz| < 1e-10, where the 1e-9 rad
 * bar needs ~3e-7 (exp X). */
int stmpm_debug_box_muller_star(const uint32_t* wa, const uint32_t* wb, int64_t n, double* z1, double* z2);
/* Same for fp32 mode's Box-Muller (host build: log2f / sqrtf stand in for the MUFU approximations). */
int stmpm_debug_box_muller_f32(const uint32_t* wa, const uint32_t* wb, int64_t n, float* z1, float* z2);
/* Number of kernels this library has launched on this handle (bench.py's gpu_launches claim). */
int stmpm_launch_count(const stmpm_t* h, int64_t* n);

#ifdef __cplusplus
}
#endif
#endif /* STMPM_H */
