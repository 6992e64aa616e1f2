// stmpm.cu — kernels + C ABI (include/stmpm.h) of the B200-native StarTrackerMPM Monte Carlo trial loop.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -shared -Xcompiler -fPIC  (see build.py)
//
// Kernel design: persistent CTAs, one per SM, 384 threads, one trial per thread.  The fp32 copy of the
// sky-grid-sorted catalog (16 B/star), the Box-Muller tables, the per-thread survivor lists and the CTA's sort window +
// group scheduler live in shared memory; the exact fp64 star records (one 32 B sector each) are gathered from L2 only for
// pre-filter survivors; the error histogram is accumulated with fire-and-forget REDs in L2.
#include "../../include/stmpm.h"
#include "stmpm_device.cuh"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

using namespace stmpm;


// ------------------------------------------------------------------------------------------------ kernel arguments
struct KArgs {
    const float4* cat32;
    const StarRec* cat64;
    const uint16_t* cell_start;   // n_cells + 1 (padded)
    const int2* band_tab;         // (cell_base, n_ra) per band
    int n_stars, n_cells_pad, n_bands;
    float inv_band_h;
    // per-boresight-cell candidate lists (magnitude-sorted star indices), the fast phase-A path
    const int2* lk_band_tab;      // lookup grid: (cell_base, n_ra) per band
    const uint32_t* lk_off;       // n_lk_cells + 1 offsets into lk_list (multiples of 8 entries)
    const uint16_t* lk_list;      // star indices into the sorted catalog; each list ends with index n_stars (the +inf dummy)
    int lk_bands;
    float lk_inv_band_h, lk_cone; // a trial whose scan cone exceeds lk_cone takes the sky-grid scan instead
    float lk_cone_tan;            // tan(lk_cone - 0.002): what a trial's cone_t is compared with
    const MathTables* tables;
    const uint8_t* lk_key;        // [KEY_LEVELS][n_lk_cells][key_rolls]: stars with mag <= KEY_M0 + (k+1)/16 on the array footprint, saturated at 255
                                  // (level-major: a run's MAX_MAGNITUDE sits in a few levels, whose slices — 41 KB disc, 330 KB footprint — stay cached)
    int lk_cells;                 // n_lk_cells
    int key_rolls;                // orientation bins of the table handed to this launch (1 = the disc table, KEY_ROLLS = the footprint table)
    float key_roll_inv;           // 1 / symmetry period of the footprint's orientation (2/pi for a square array, 1/pi otherwise)
    int win;                      // trials per CTA sort window (multiple of 32, <= WIN_MAX)
    int windows_per_unit;         // ceil(part_len / win)
    double f_ideal, f_ideal2, half_u, half_v;
    int n_min, hist_shift, n_bins;   // hist_shift = 52 - log2(sub-bins per octave)
    long long n_segments, trials_per_segment, trial0, part_len;
    int parts_per_segment;
    uint32_t k0, k1;
    const double* cols[STMPM_NCOLS];
    const DistDev* dists;         // per-segment distributions (n_segments > 1)
    DistDev dist0;                // the single segment's distribution (n_segments == 1): no device buffer to race on
    double* err;
    int* ndet;
    float* maxmag;                // optional: magnitude of the faintest detected star (-inf if none) — star_mag objective
    double* stats;
    long long stats_len;
    unsigned long long* counters;   // INSTR only: [visits, mag_pass, survivors, detected]
    double* stage;                  // MODE 0: per-CTA scratch [gridDim][2][STMPM_NCOLS][WIN_MAX] — the window's columns in key order
    uint8_t* ndet8;                 // optional compact record: detected stars saturated at 255 (9-byte records with err)
    float* err32;                   // optional compact record: the error rounded to float32 (5-byte records with ndet8)
    int* status;                    // host-mapped word: a scheduler wait that exceeds WATCHDOG_NS writes STMPM_E_WATCHDOG here
    const int* cancel;              // optional: a non-zero word makes the launch a no-op (on-device MonteCarlo stop rule)
};

// Shared-memory layout: every region at a COMPILE-TIME offset (fixed-size regions first, the variable-size catalog
// last), so ptxas addresses them as [reg + immediate] instead of re-deriving bases from kernel parameters per access.
// CTA-wide sort window: two permutation buffers (window w uses buffer w & 1: w + 1 is built while w is consumed), the keys
// and the 256 bin counters (16 bit, two per atomic word) of the window being built, and the scheduler words.
constexpr size_t WIN_BYTES = 2 * 2 * (size_t)WIN_MAX /*perm u16 x2*/ + WIN_MAX /*keys u8*/ + 2 * KEY_BINS /*counters u16*/ + 32 /*sched*/ + 64 /*window meta x2*/;
static_assert(WIN_MAX <= 4096 && WIN_MAX % 128 == 0 && KEY_BINS == 256, "16-bit bin counters; the key pass runs 128 trials per round");
#if !defined(STMPM_KEY_SLOTS)
#define STMPM_KEY_SLOTS 4
#endif
// Which instantiations run two stars per star-loop iteration (exp W, same box): pre-sampled fp64 +4.6 % with it; fp32 mode -5 %
// (3.08 -> 2.92e9: its loop spills at two stars) and native RNG -0.7 % — those keep the single-star loop.
#if !defined(STMPM_TWO_STARS_F32)
#define STMPM_TWO_STARS_F32 0
#endif
#if !defined(STMPM_TWO_STARS_RNG)
#define STMPM_TWO_STARS_RNG 0
#endif
constexpr int KS = STMPM_KEY_SLOTS;    // key pass: 32 KS trials per round (the round's loads are issued before their first use)
constexpr int LK_BANDS = 180;          // 1-degree lookup grid
constexpr int MAX_BANDS = 256;         // sky-grid declination bands (fallback scan)
constexpr size_t OFF_LIST = 0;
constexpr size_t OFF_WIN = OFF_LIST + (size_t)LIST_CAP * TPB * 2;
constexpr size_t OFF_TAB = OFF_WIN + (WIN_BYTES + 15) / 16 * 16;
constexpr size_t OFF_MBAR = OFF_TAB + (sizeof(MathTables) + 15) / 16 * 16;
constexpr size_t OFF_LKBAND = OFF_MBAR + 16;
constexpr size_t OFF_BAND = OFF_LKBAND + (size_t)LK_BANDS * 8;
constexpr size_t OFF_CAT = (OFF_BAND + (size_t)MAX_BANDS * 8 + 127) / 128 * 128;
constexpr size_t CAT_ENTRY_BYTES = 16;
static size_t smem_total(int n_stars) { return OFF_CAT + (size_t)(n_stars + 1) * CAT_ENTRY_BYTES; }   // + the +inf terminator star
// the real catalog (BSC5: 9 110 stars, 9 096 after load_bsc5 drops the non-stellar entries) must fit the 227 KB a CTA can opt in to
static_assert(OFF_CAT + (9110 + 1) * CAT_ENTRY_BYTES <= 227 * 1024, "a 9110-star catalog no longer fits in shared memory");

// One star record = one 32-byte sector = ONE load: sm_100's 256-bit LDG.  The lanes of a warp gather 32 different records,
// and the LSU works such an instruction off one tag look-up per distinct line — two 128-bit loads per record would keep
// it busy twice as long, and shared-memory loads queue behind it in the same pipe.
__device__ __forceinline__ double4 ld_rec(const StarRec* p) {
    double4 v;
    asm("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
    return v;
}
// fire-and-forget fp64 add (RED, no response): atomicAdd with an unused result compiles to ATOMG ... RZ here
__device__ __forceinline__ void red_add_f64(double* p, double v) {
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "d"(v) : "memory");
}
// staged-column accesses: cache-global by default; STMPM_STAGE_CS = streaming (evict-first) hints, an experiment against
// L2 pollution by the 63 MB of windows in flight
#if defined(STMPM_STAGE_CS)
#define STAGE_ST __stcs
#define STAGE_LD __ldcs
#else
#define STAGE_ST __stcg
#define STAGE_LD __ldcg
#endif

// Scheduler watchdog: the window / group waits below are bounded by WALL TIME (%globaltimer), not by a poll count — a poll
// budget can run out under compute-sanitizer, cuda-gdb or preemption without any deadlock.  On expiry the warp records
// STMPM_E_WATCHDOG in a host-mapped status word and leaves its trial loop cleanly (no __trap: a trap is a sticky,
// context-fatal error that would also take the caller's torch context down); the host entry points report the code.
constexpr unsigned long long WATCHDOG_NS = 20ull * 1000000000ull;
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// Evaluates the eight candidates of one pack WITHOUT an exit between them: one basic block, so ptxas can issue the eight
// catalog reads (and the caller's next-pack load) ahead of the arithmetic — with an exit after every candidate it sinks
// that load below the exits, to just before its first use, where nothing hides its latency.  The list is magnitude-
// sorted, so once a star fails the magnitude test every later one fails it too: "the last one passed" (the return value)
// is the continue condition, and the stars after the first failure cost a few predicated-off instructions.
template <int INSTR>
__device__ __forceinline__ bool eval_pack(const Ctx32& c, uint32_t sb_cat, const uint4 pk, uint32_t sb_row, int& cnt,
                                          unsigned long long* ins) {
    const uint32_t w4[4] = {pk.x, pk.y, pk.z, pk.w};
    bool m = true;
    uint32_t pass = 0;                       // bit e: candidate e survives (pushes deferred: an intervening st.shared
#pragma unroll                               // would pin every later catalog read behind it — ptxas must assume aliasing)
    for (int e = 0; e < 8; ++e) {
        const uint32_t idx = (w4[e >> 1] >> (16 * (e & 1))) & 0xffffu;
        const float4 sc = lds_f4(sb_cat + idx * 16u);
        if (INSTR == 1) ins[0] += m ? 1 : 0;
        m = sc.w <= c.magf;
        if (INSTR == 1) ins[1] += m ? 1 : 0;
        const bool g = prefilter_geom(c, sc);        // evaluated regardless of m: no branch
        pass |= (m & g) ? (1u << e) : 0u;
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        if (pass & (1u << e)) {
            sts_u16(sb_row + (uint32_t)cnt * (TPB * 2u), (w4[e >> 1] >> (16 * (e & 1))) & 0xffffu);
            ++cnt;
            if (INSTR == 1) ++ins[2];
        }
    }
    return m;
}

struct Pack { uint4 a; };
constexpr int PACK_N = 8;
__device__ __forceinline__ Pack ld_pack(const uint16_t* p) { Pack v; v.a = __ldg(reinterpret_cast<const uint4*>(p)); return v; }

// Walks the magnitude-sorted candidate list of the boresight's lookup cell from lp on, pushing pre-filter survivors to
// row[cnt * TPB]; the first star fainter than MAX_MAGNITUDE ends the scan (returns true).  Returns false while candidates
// remain and fewer than PACK_N row slots are free: the caller drains the row and calls again.  Every list ends with a
// pack of entries that point at a dummy catalog star of magnitude +inf (index n_stars), so the magnitude test alone
// terminates the walk: no per-entry sentinel or bounds check, and the next pack can always be loaded (slack at the end).
// pk_next = the pack at lp, loaded by the caller (the first one of a trial long before the scan starts).
template <int INSTR>
__device__ __forceinline__ bool scan_list(const Ctx32& c, uint32_t sb_cat, const uint16_t* __restrict__& lp, Pack pk_next,
                                          uint32_t sb_row, int& cnt, unsigned long long* ins) {
    while (cnt <= LIST_CAP - PACK_N) {
        const Pack pk = pk_next;
        lp += PACK_N;
        pk_next = ld_pack(lp);
        if (!eval_pack<INSTR>(c, sb_cat, pk.a, sb_row, cnt, ins)) return true;
    }
    return false;
}

// ---- TMA bulk copy (cp.async.bulk, global -> shared, completion on an mbarrier): stages the 146 KB fp32 catalog
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}

__device__ __forceinline__ int lookup_cell(const int2* __restrict__ s_lkband, int lk_bands, float lk_inv_band_h, float ra0,
                                           float dec0) {
    const int lb = min(lk_bands - 1, max(0, (int)floorf((dec0 + 1.57079633f) * lk_inv_band_h)));
    const int2 bt = s_lkband[lb];
    return bt.x + min(bt.y - 1, max(0, (int)floorf(ra0 * 0.159154943f * (float)bt.y)));
}

// work key of a trial: stars brighter than its MAX_MAGNITUDE level on the array footprint centred on its lookup cell, in its
// roll bin (host table, rebuild_lists).  Any finite input gives a valid index; the key only orders the work.
__device__ __forceinline__ int work_key(const KArgs& a, const int2* __restrict__ s_lkband, float raf, float decf, float rollf, float mm) {
    raf -= 6.28318531f * floorf(raf * 0.159154943f);
    decf = fminf(fmaxf(decf, -1.57079633f), 1.57079633f);
    const int cell = lookup_cell(s_lkband, a.lk_bands, a.lk_inv_band_h, raf, decf);
    float t = rollf * a.key_roll_inv;
    t -= floorf(t);
    const int rb = min(a.key_rolls - 1, max(0, (int)(t * (float)a.key_rolls)));
    const int lev = min(KEY_LEVELS - 1, (int)floorf((mm - KEY_M0) * 16.0f));
    return lev < 0 ? 0 : (int)__ldg(a.lk_key + ((size_t)lev * a.lk_cells + cell) * a.key_rolls + rb);
}

// ------------------------------------------------------------------------------------------------ fallback trial
// A trial whose scan cone is wider than the candidate lists were built for (huge translations, tilts or centroid noise, tiny
// focal lengths) is evaluated here in full: conservative fp32 walk over the sky-grid bands that overlap the cone, survivors
// through the same exact per-star evaluation.  Never taken with the reference's configurations; out of line so that its
// state (seven walk registers) stays out of the hot path's register budget.
template <int PREC, int INSTR>
__device__ __noinline__ void band_walk_trial(const Ctx64* cp, const Ctx32* cfp, const StarRec* __restrict__ cat64,
                                             const uint16_t* __restrict__ g_cell, int n_bands, float inv_band_h, double f_ideal,
                                             double half_u, double half_v, uint32_t sb, uint32_t ka, uint32_t kb, double* B_out,
                                             int* n_det_out, float* mag_out, unsigned long long* ins) {
    const Ctx64 c = *cp;
    Ctx32 cf = *cfp;
    cf.cone = cone_angle(cf);
    unsigned char* smem = reinterpret_cast<unsigned char*>(__cvta_shared_to_generic((size_t)sb));
    const float4* s_cat = reinterpret_cast<const float4*>(smem + OFF_CAT);
    const int2* s_band = reinterpret_cast<const int2*>(smem + OFF_BAND);
    uint16_t* s_list = reinterpret_cast<uint16_t*>(smem + OFF_LIST) + threadIdx.x;   // entry j at s_list[j * TPB]
    const TabSmem tabs{sb + (uint32_t)OFF_TAB};
    double B[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    int n_det = 0;
    float mag_faintest = -INFINITY;

    int b = max(0, (int)floorf((cf.dec0 - cf.cone + 1.57079633f) * inv_band_h));
    const int b_hi = min(n_bands - 1, (int)floorf((cf.dec0 + cf.cone + 1.57079633f) * inv_band_h));
    float dalpha = -1.0f;                           // < 0: the cone touches a pole, take whole bands
    if (fabsf(cf.dec0) + cf.cone < 1.5607963f) dalpha = asinf(fminf(sinf(cf.cone) / cosf(cf.dec0), 1.0f)) + 0.002f;
    int i = 0, i_end = 0, pi = 0, p_end = 0;
    bool scanning = true;
    while (scanning) {
        int cnt = 0;
        while (cnt < LIST_CAP) {
            if (i >= i_end) {
                if (pi < p_end) { i = pi; i_end = p_end; pi = p_end = 0; continue; }
                if (b > b_hi) { scanning = false; break; }
                const int2 bt = s_band[b];
                ++b;
                const int base = bt.x, n = bt.y;
                if (dalpha < 0.0f) { i = g_cell[base]; i_end = g_cell[base + n]; continue; }
                const float inv_w = (float)n * 0.159154943f;
                const int c_lo = (int)floorf((cf.ra0 - dalpha) * inv_w);
                const int c_hi = (int)floorf((cf.ra0 + dalpha) * inv_w);
                const int count = c_hi - c_lo + 1;
                if (count >= n) { i = g_cell[base]; i_end = g_cell[base + n]; continue; }
                int lo = c_lo % n;
                if (lo < 0) lo += n;
                if (lo + count <= n) { i = g_cell[base + lo]; i_end = g_cell[base + lo + count]; }
                else {
                    i = g_cell[base + lo]; i_end = g_cell[base + n];
                    pi = g_cell[base]; p_end = g_cell[base + lo + count - n];
                }
                continue;
            }
            const float4 sc = s_cat[i];
            if (INSTR == 1) { ++ins[0]; if (sc.w <= cf.magf) ++ins[1]; }
            if (sc.w <= cf.magf && prefilter_geom(cf, sc)) {
                s_list[cnt * TPB] = (uint16_t)i; ++cnt; if (INSTR == 1) ++ins[2];
            }
            ++i;
        }
        for (int j = 0; j < cnt; ++j) {
            const StarRec s = cat64[s_list[j * TPB]];
            uint32_t x0, x1;
            philox2x32_10(s.star_id, kb, ka, x0, x1);
            bool det;
            if (PREC == 64) det = star_exact(c, tabs, s, x0, x1, f_ideal, f_ideal * f_ideal, half_u, half_v, B);
            else det = star_f32(c, cf, tabs, s, x0, x1, f_ideal, f_ideal * f_ideal, half_u, half_v, B);
            n_det += det ? 1 : 0;
            if (INSTR == 2 && det) mag_faintest = fmaxf(mag_faintest, s.mag);
        }
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) B_out[k] = B[k];
    *n_det_out = n_det;
    *mag_out = mag_faintest;
}

// Warp reduction of the per-thread moments into one segment's packed statistics (six atomics per warp), then reset.
__device__ __forceinline__ void flush_moments(double* st, int lane, unsigned int& acc_ok, unsigned int& acc_fail,
                                              unsigned long long& acc_ndet, double& acc_x, double& acc_xx, double& acc_max) {
    double v[5] = {(double)acc_ok, (double)acc_fail, acc_x, acc_xx, (double)acc_ndet};
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
        for (int k = 0; k < 5; ++k) v[k] += __shfl_xor_sync(0xffffffffu, v[k], off);
        acc_max = fmax(acc_max, __shfl_xor_sync(0xffffffffu, acc_max, off));
    }
    if (lane < 5) {
        double r = v[0];
#pragma unroll
        for (int k = 1; k < 5; ++k) r = (lane == k) ? v[k] : r;
        if (r != 0.0) atomicAdd(st + lane, r);
    } else if (lane == 5) {
        atomicMax(reinterpret_cast<unsigned long long*>(st + 5), (unsigned long long)__double_as_longlong(acc_max));
    }
    acc_ok = acc_fail = 0; acc_ndet = 0; acc_x = acc_xx = acc_max = 0.0;
}

// ------------------------------------------------------------------------------------------------ the trial kernel
// MODE 0: pre-sampled columns.  MODE 1: native RNG.  PREC 64 / 32: per-star arithmetic.
// INSTR 1: also count candidate visits / magnitude passes / pre-filter survivors (work-model calibration, not timed).
// INSTR 2: also record the faintest detected star's magnitude (star_mag objective).
template <int MODE, int PREC, int INSTR>
__global__ void __launch_bounds__(TPB, 1) trials_kernel(const KArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    if (a.cancel != nullptr && *reinterpret_cast<const volatile int*>(a.cancel) != 0) return;   // converged in an earlier chunk
    float4* s_cat = reinterpret_cast<float4*>(smem + OFF_CAT);
    int2* s_band = reinterpret_cast<int2*>(smem + OFF_BAND);
    int2* s_lkband = reinterpret_cast<int2*>(smem + OFF_LKBAND);
    const MathTables* s_tab = reinterpret_cast<const MathTables*>(smem + OFF_TAB);
    // 32-bit shared-memory addresses of the hot regions, kept in registers (see lds_f4 / TabSmem)
    const uint32_t sb = smem_u32(smem);
    const uint32_t sb_cat = sb + (uint32_t)OFF_CAT;
    uint32_t sb_row = sb + (uint32_t)OFF_LIST + threadIdx.x * 2u;
    asm volatile("" : "+r"(sb_row));   // keep it: ptxas otherwise re-reads %tid.x and re-adds at every list push
    const TabSmem tabs{sb + (uint32_t)OFF_TAB};
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // the CTA's sort window: two permutation buffers, keys + bin counters of the window being built, scheduler words
    uint16_t* s_perm = reinterpret_cast<uint16_t*>(smem + OFF_WIN);                       // [2][WIN_MAX]
    uint8_t* w_key = smem + OFF_WIN + 4 * (size_t)WIN_MAX;                                // [WIN_MAX]
    unsigned int* w_hist = reinterpret_cast<unsigned int*>(w_key + WIN_MAX);              // 256 16-bit counters, 2 per word
    int* s_sched = reinterpret_cast<int*>(w_hist + KEY_BINS / 2);                         // claim, built, reads[2]
    volatile int* v_sched = s_sched;
    long long* s_meta = reinterpret_cast<long long*>(s_sched + 8);                        // per window buffer: rec0, seg, nw

    // ---- stage the catalog + tables into shared memory (once per CTA; the CTA is persistent).  The 146 KB fp32 catalog
    // goes through the TMA unit: one elected thread issues cp.async.bulk chunks that complete on an mbarrier while the
    // other threads load the small tables; nobody touches s_cat before the barrier phase flips.
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + OFF_MBAR);
    if (threadIdx.x == 0) mbar_init(s_bar, 1);
    if (threadIdx.x < 8) s_sched[threadIdx.x] = 0;
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t total = (uint32_t)(a.n_stars + 1) * 16u;
        mbar_expect_tx(s_bar, total);
        const unsigned char* src = reinterpret_cast<const unsigned char*>(a.cat32);
        unsigned char* dst = reinterpret_cast<unsigned char*>(s_cat);
        for (uint32_t off = 0; off < total; off += 32768u) bulk_g2s(dst + off, src + off, min(32768u, total - off), s_bar);
    }
    {
        for (int i = threadIdx.x; i < a.n_bands; i += TPB) s_band[i] = __ldg(a.band_tab + i);
        for (int i = threadIdx.x; i < a.lk_bands; i += TPB) s_lkband[i] = __ldg(a.lk_band_tab + i);
        const double* gt = reinterpret_cast<const double*>(a.tables);
        double* stb = reinterpret_cast<double*>(smem + OFF_TAB);
        for (int i = threadIdx.x; i < (int)(sizeof(MathTables) / 8); i += TPB) stb[i] = __ldg(gt + i);
    }
    while (!mbar_try_wait(s_bar, 0)) {}
    __syncthreads();

    const float hu32 = (float)a.half_u, hv32 = (float)a.half_v;
    const bool want_stats = a.stats != nullptr;
    // star loop: two stars per iteration (exp U) or one; per instantiation, by measurement (exp V, W)
#if defined(STMPM_ONE_STAR)
    constexpr bool TWO_STARS = false;
#else
    constexpr bool TWO_STARS = (PREC == 64 || STMPM_TWO_STARS_F32) && (MODE == 0 || STMPM_TWO_STARS_RNG);
#endif

    // ---- work distribution.  The CTA owns units blockIdx.x, blockIdx.x + gridDim.x, ... (unit = a contiguous range of
    // one segment's trials); a unit is cut into windows of `win` trials.  Each window is ordered by work key ONCE for the
    // whole CTA (counting sort by one warp, into the permutation buffer the previous-but-one window has vacated), and the
    // 12 warps then claim its groups of 32 consecutive sorted trials from a shared counter.  The wider the window, the
    // more alike the 32 trials a warp runs in lock-step (2 048 instead of 128 trials: ~8 % of the issue slots, lanes idle in
    // the star loop) — but the wider the 32 lanes' column gathers scatter (launch_trials picks the window per mode).
    // No CTA barrier: window w + 1 is built by the warp that claims the first group of window w, while the others consume w.
    const int win = a.win;                                 // trials per window, multiple of 32, <= WIN_MAX
    const int gpw = win >> 5;                              // groups per window
    const int gpw_shift = (gpw & (gpw - 1)) == 0 ? __ffs(gpw) - 1 : -1;
    const long long n_units = a.n_segments * (long long)a.parts_per_segment;
    const long long my_units = (n_units - (long long)blockIdx.x + gridDim.x - 1) / gridDim.x;
    const int n_windows = (int)(my_units * a.windows_per_unit);
    const int n_groups = n_windows * gpw;                  // launch_trials keeps this below 2^31

    // per-thread streaming accumulators (MODEL.md §9), flushed whenever the warp moves on to another segment
    unsigned int acc_ok = 0, acc_fail = 0;
    unsigned long long acc_ndet = 0;
    double acc_x = 0.0, acc_xx = 0.0, acc_max = 0.0;
    unsigned long long ins[4] = {0, 0, 0, 0};
    long long cur_seg = -1;

    int build_w = (warp == 0 && n_windows > 0) ? 0 : -1;   // pending window to build (warp 0 builds the first one)
    bool have_group = false, aborted = false;
    int grp_w = 0, grp_r = 0, grp_q = 0;
    long long grp_rec0 = 0, grp_seg = 0;
    int grp_nw = 0;
    for (;;) {
        if (build_w >= 0) {
            // ================= key pass + counting sort of window build_w into permutation buffer build_w & 1
            const int bw = build_w;
            build_w = -1;
            uint16_t* w_perm = s_perm + (bw & 1) * WIN_MAX;
            {   // the buffer is free once every group of window bw - 2 has copied its permutation entries out
                const int need = gpw * (bw >> 1);
                if (v_sched[2 + (bw & 1)] < need) {
                    const unsigned long long t0 = global_ns();
                    while (v_sched[2 + (bw & 1)] < need) {
                        __nanosleep(32);
                        if (global_ns() - t0 > WATCHDOG_NS) { aborted = true; break; }
                    }
                }
                if (aborted) break;
                __threadfence_block();
            }
            const long long ku = bw / a.windows_per_unit;
            const long long unit = (long long)blockIdx.x + ku * gridDim.x;
            const long long seg = unit / a.parts_per_segment;
            const long long part = unit - seg * a.parts_per_segment;
            const long long t_begin = part * a.part_len;
            const long long t_end = min(t_begin + a.part_len, a.trials_per_segment);
            const long long wb = t_begin + (long long)(bw - ku * a.windows_per_unit) * win;
            const int nw = (int)max(0LL, min((long long)win, t_end - wb));     // valid trials in this window
            const long long rec0 = seg * a.trials_per_segment + wb;
            const DistDev* dist = (MODE == 1) ? ((a.n_segments == 1) ? &a.dist0 : a.dists + seg) : nullptr;
#if !defined(STMPM_NO_WIN_META)
            // the window's place in the launch, for its groups: they read three words instead of redoing ~100 instructions
            // of 64-bit index arithmetic (two divisions) per group (exp S)
            if (lane == 0) { s_meta[(bw & 1) * 4] = rec0; s_meta[(bw & 1) * 4 + 1] = seg; s_meta[(bw & 1) * 4 + 2] = nw; }
#endif
#pragma unroll
            for (int q = 0; q < KEY_BINS / 2 / 32; ++q) w_hist[lane + 32 * q] = 0u;
            __syncwarp();
            if (MODE == 0) {
                // 128 trials per round: the four slots' column loads are issued before the first use, then the four
                // key-table loads — two exposed memory latencies per round
                for (int r0 = 0; r0 < win; r0 += 32 * KS) {
                    double kra[KS], kdec[KS], kroll[KS], kmm[KS];
#pragma unroll
                    for (int q = 0; q < KS; ++q) {
                        const int r = r0 + lane + 32 * q;
                        kra[q] = kdec[q] = kroll[q] = 0.0; kmm[q] = -100.0;
                        if (r < nw) {
                            kra[q] = __ldg(a.cols[0] + rec0 + r); kdec[q] = __ldg(a.cols[1] + rec0 + r);
                            if (a.key_rolls > 1) kroll[q] = __ldg(a.cols[2] + rec0 + r);
                            kmm[q] = __ldg(a.cols[12] + rec0 + r);
                        }
                    }
                    // pull the round's other ten columns into L2 (one 32-byte sector = 4 trials per lane): the groups gather
                    // them in key order, 32 scattered sectors per load — as L2 hits those gathers hold their L1 miss slots
                    // a quarter as long as DRAM misses do, and the LSU pipe they share with shared memory stays fluid
                    if (lane < 8 * KS && r0 + 4 * lane < nw) {          // (with staging too: the copy below then reads L2 hits)
#pragma unroll
                        for (int cI = 2; cI < 12; ++cI) prefetch_l2(a.cols[cI] + rec0 + r0 + 4 * lane);
                    }
                    int kkey[KS];
#pragma unroll
                    for (int q = 0; q < KS; ++q)
                        kkey[q] = work_key(a, s_lkband, (float)kra[q], (float)kdec[q], (float)kroll[q], (float)kmm[q]);
#pragma unroll
                    for (int q = 0; q < KS; ++q) {
                        const int r = r0 + lane + 32 * q;
                        if (r < win) {
                            const int key = r < nw ? kkey[q] : KEY_BINS - 1;   // padding slots sort last
                            w_key[r] = (uint8_t)key;
                            atomicAdd(w_hist + (key >> 1), 1u << (16 * (key & 1)));
                        }
                    }
                }
            } else {
                for (int r = lane; r < win; r += 32) {
                    int key = KEY_BINS - 1;                               // padding slots sort last
                    if (r < nw) {
                        double ra, dec, roll, mm;
                        const unsigned long long gi = (unsigned long long)(a.trial0 + rec0 + r);
                        sample_key_params(*dist, s_tab, (uint32_t)gi, (uint32_t)(gi >> 32), a.k0, a.k1, ra, dec, roll, mm);
                        key = work_key(a, s_lkband, (float)ra, (float)dec, (float)roll, (float)mm);
                    }
                    w_key[r] = (uint8_t)key;
                    atomicAdd(w_hist + (key >> 1), 1u << (16 * (key & 1)));
                }
            }
            __syncwarp();
            {   // exclusive scan of the 256 16-bit counters: lane owns bins 8 lane .. 8 lane + 7 (four words)
                uint32_t wv[4], hv[8], tot = 0;
#pragma unroll
                for (int q = 0; q < 4; ++q) { wv[q] = w_hist[4 * lane + q]; hv[2 * q] = wv[q] & 0xffffu; hv[2 * q + 1] = wv[q] >> 16; }
#pragma unroll
                for (int q = 0; q < 8; ++q) tot += hv[q];
                uint32_t incl = tot;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const uint32_t o = __shfl_up_sync(0xffffffffu, incl, off);
                    if (lane >= off) incl += o;
                }
                uint32_t run = incl - tot;
                __syncwarp();
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t lo = run; run += hv[2 * q];
                    const uint32_t hi = run; run += hv[2 * q + 1];
                    w_hist[4 * lane + q] = lo | (hi << 16);
                }
            }
            __syncwarp();
            if (MODE == 0 && a.stage != nullptr) {
                // sorted parameter staging: the window's 13 columns are re-written in KEY ORDER to an L2-resident scratch
                // (coalesced column reads, rank-scattered 8-byte writes, by this one warp), so that the 32 lanes of a group
                // later read 13 x 256 contiguous bytes instead of gathering 13 x 32 scattered words — which is what kept
                // pre-sampled windows at 256 trials in round 1 (the gathers saturated the LSU pipe at wider windows)
                double* sg = a.stage + ((size_t)blockIdx.x * 2 + (size_t)(bw & 1)) * (size_t)(STMPM_NCOLS * WIN_MAX);
                for (int r = lane; r < win; r += 32) {
                    const int key = w_key[r];
                    const unsigned int old = atomicAdd(w_hist + (key >> 1), 1u << (16 * (key & 1)));
                    const unsigned int rank = (old >> (16 * (key & 1))) & 0xffffu;
                    w_perm[rank] = (uint16_t)r;
                    if (r < nw) {
                        double v[STMPM_NCOLS];
#pragma unroll
                        for (int cI = 0; cI < STMPM_NCOLS; ++cI) v[cI] = __ldcs(a.cols[cI] + rec0 + r);
#pragma unroll
                        for (int cI = 0; cI < STMPM_NCOLS; ++cI) STAGE_ST(sg + (size_t)cI * WIN_MAX + rank, v[cI]);
                    }
                }
                __threadfence();
            } else {
                for (int r = lane; r < win; r += 32) {
                    const int key = w_key[r];
                    const unsigned int old = atomicAdd(w_hist + (key >> 1), 1u << (16 * (key & 1)));
                    w_perm[(old >> (16 * (key & 1))) & 0xffffu] = (uint16_t)r;
                }
            }
            __threadfence_block();
            __syncwarp();
            if (lane == 0) v_sched[1] = bw + 1;            // publish: windows 0 .. bw are built
        }

        if (!have_group) {
            int g = 0;
            if (lane == 0) g = atomicAdd(s_sched, 1);
            g = __shfl_sync(0xffffffffu, g, 0);
            if (g >= n_groups) break;
#if !defined(STMPM_NO_WIN_META)
            grp_w = gpw_shift >= 0 ? (g >> gpw_shift) : g / gpw;   // windows of 2^k groups (the defaults): no division
#else
            grp_w = g / gpw;
#endif
            const int gi = g - grp_w * gpw;
            if (v_sched[1] <= grp_w) {
                const unsigned long long t0 = global_ns();
                while (v_sched[1] <= grp_w) {
                    __nanosleep(64);
                    if (global_ns() - t0 > WATCHDOG_NS) { aborted = true; break; }
                }
            }
            if (aborted) break;
            __threadfence_block();
            grp_q = gi * 32 + lane;                                   // position in key order
            grp_r = s_perm[(grp_w & 1) * WIN_MAX + grp_q];
#if !defined(STMPM_NO_WIN_META)
            grp_rec0 = s_meta[(grp_w & 1) * 4]; grp_seg = s_meta[(grp_w & 1) * 4 + 1]; grp_nw = (int)s_meta[(grp_w & 1) * 4 + 2];
#endif
            __threadfence_block();                        // the entry is in a register before the slot is reported free
            __syncwarp();
            // (with sorted parameter staging the buffer also holds the group's staged columns: it is reported free only
            // after those are in registers, below)
            if (!(MODE == 0 && a.stage != nullptr) && lane == 0) atomicAdd(s_sched + 2 + (grp_w & 1), 1);
            have_group = true;
            if (gi == 0 && grp_w + 1 < n_windows) { build_w = grp_w + 1; continue; }   // build the next window first
        }
        have_group = false;

        // ================= one group: 32 trials of window grp_w, consecutive in key order
        {
#if !defined(STMPM_NO_WIN_META)
            const long long seg = grp_seg, rec0 = grp_rec0;
            const int nw = grp_nw;
#else
            const long long ku = grp_w / a.windows_per_unit;
            const long long unit = (long long)blockIdx.x + ku * gridDim.x;
            const long long seg = unit / a.parts_per_segment;
            const long long part = unit - seg * a.parts_per_segment;
            const long long t_begin = part * a.part_len;
            const long long t_end = min(t_begin + a.part_len, a.trials_per_segment);
            const long long wb = t_begin + (long long)(grp_w - ku * a.windows_per_unit) * win;
            const int nw = (int)max(0LL, min((long long)win, t_end - wb));
            const long long rec0 = seg * a.trials_per_segment + wb;
#endif
            const DistDev* dist = (MODE == 1) ? ((a.n_segments == 1) ? &a.dist0 : a.dists + seg) : nullptr;
            if (seg != cur_seg) {
                if (want_stats && cur_seg >= 0)
                    flush_moments(a.stats + cur_seg * a.stats_len, lane, acc_ok, acc_fail, acc_ndet, acc_x, acc_xx, acc_max);
                cur_seg = seg;
            }
            {
                const int r = grp_r;
                const bool valid = r < nw;
                if (!__any_sync(0xffffffffu, valid)) {
                    if (MODE == 0 && a.stage != nullptr && lane == 0) atomicAdd(s_sched + 2 + (grp_w & 1), 1);
                    continue;
                }
                const long long rec = rec0 + r;
                const unsigned long long gidx = (unsigned long long)(a.trial0 + rec);
                const uint32_t t_lo = (uint32_t)gidx, t_hi = (uint32_t)(gidx >> 32);

                TrialParams p;
                if (MODE == 0 && a.stage != nullptr) {
                    if (valid) {                                // coalesced: lane q reads word q of each staged column
                        const double* sg = a.stage + ((size_t)blockIdx.x * 2 + (size_t)(grp_w & 1)) * (size_t)(STMPM_NCOLS * WIN_MAX) + grp_q;
                        p.ra = STAGE_LD(sg);                      p.dec = STAGE_LD(sg + WIN_MAX);
                        p.roll = STAGE_LD(sg + 2 * WIN_MAX);      p.f = STAGE_LD(sg + 3 * WIN_MAX);
                        p.ex = STAGE_LD(sg + 4 * WIN_MAX);        p.ey = STAGE_LD(sg + 5 * WIN_MAX);
                        p.ez = STAGE_LD(sg + 6 * WIN_MAX);        p.phi = STAGE_LD(sg + 7 * WIN_MAX);
                        p.theta = STAGE_LD(sg + 8 * WIN_MAX);     p.psi = STAGE_LD(sg + 9 * WIN_MAX);
                        p.devx = STAGE_LD(sg + 10 * WIN_MAX);     p.devy = STAGE_LD(sg + 11 * WIN_MAX);
                        p.maxmag = STAGE_LD(sg + 12 * WIN_MAX);
                    } else {
                        p.ra = p.dec = p.roll = 0.0; p.f = 1000.0; p.ex = p.ey = p.ez = p.phi = p.theta = p.psi = 0.0;
                        p.devx = p.devy = 0.0; p.maxmag = -100.0;
                    }
                    __threadfence_block();                      // the thirteen words are in registers: the buffer may be rebuilt
                    __syncwarp();
                    if (lane == 0) atomicAdd(s_sched + 2 + (grp_w & 1), 1);
                } else if (MODE == 0) {
                    if (valid) {
                        // ld.global.cg: the 32 lanes gather from 32 different lines of the window; cached in L1 each of the 13
                        // loads would claim 32 of its ~220 lines at once and stall the LSU pipe (shared-memory loads queue
                        // behind it) on line allocation.  L2 holds the window (prefetched by the key pass).
                        p.ra = __ldcg(a.cols[0] + rec);     p.dec = __ldcg(a.cols[1] + rec);
                        p.roll = __ldcg(a.cols[2] + rec);   p.f = __ldcg(a.cols[3] + rec);
                        p.ex = __ldcg(a.cols[4] + rec);     p.ey = __ldcg(a.cols[5] + rec);
                        p.ez = __ldcg(a.cols[6] + rec);     p.phi = __ldcg(a.cols[7] + rec);
                        p.theta = __ldcg(a.cols[8] + rec);  p.psi = __ldcg(a.cols[9] + rec);
                        p.devx = __ldcg(a.cols[10] + rec);  p.devy = __ldcg(a.cols[11] + rec);
                        p.maxmag = __ldcg(a.cols[12] + rec);
                    } else {
                        p.ra = p.dec = p.roll = 0.0; p.f = 1000.0; p.ex = p.ey = p.ez = p.phi = p.theta = p.psi = 0.0;
                        p.devx = p.devy = 0.0; p.maxmag = -100.0;
                    }
                } else {
                    sample_params(*dist, s_tab, t_lo, t_hi, a.k0, a.k1, p);
                }
                // ---- phase-A set-up, ordered so that the two dependent memory latencies in front of the candidate scan
                // (list offset of the lookup cell, first pack of the list) are spent under the fp64 set-up instead of at the
                // scan's first use.  For (ra, dec) in their principal ranges the boresight IS (ra, dec), so the cell comes
                // from the raw angles (fp32 rounding moves it by < 4e-5 rad; the lists carry a > 3e-3 rad margin).
                Ctx32 cf;
                window_and_cone(p, hu32, hv32, cf);
                const bool use_list = cf.cone_t <= a.lk_cone_tan;
                const float ra_f = (float)p.ra, dec_f = (float)p.dec;
                const bool early = use_list && fabsf(dec_f) <= 1.57079633f && fabsf(ra_f) <= 256.0f;
                uint32_t l_off = 0;
                if (early) {
                    const float ra_w = ra_f - 6.28318531f * floorf(ra_f * 0.159154943f);
                    l_off = __ldg(a.lk_off + lookup_cell(s_lkband, a.lk_bands, a.lk_inv_band_h, ra_w, dec_f));
                }
                Ctx64 c;
                setup64(p, c);
                const uint16_t* __restrict__ lp = a.lk_list + l_off;
                Pack pk0 = ld_pack(lp);
                geom32(c, cf);
                uint32_t ka, kb;
                star_key(t_lo, t_hi, a.k0, a.k1, ka, kb);
                if (use_list && !early) {                       // folded / far-out angles: cell from the rotated boresight
                    boresight32(c, cf);
                    lp = a.lk_list + __ldg(a.lk_off + lookup_cell(s_lkband, a.lk_bands, a.lk_inv_band_h, cf.ra0, cf.dec0));
                    pk0 = ld_pack(lp);
                }

                // The truth attitude C (18 registers) is needed again only by the solve: park it in local memory (L1) for the
                // duration of the candidate scan and the star loop.  At the register cap (128 with 512 threads, 168 with 384 and
                // two stars per iteration) ptxas otherwise spills values of its own choosing inside the loops and sinks the next
                // star record's load to just before its use (round 2, fp32 instantiation: 19 % of all stall samples on the first
                // instruction that reuses the load's registers; two stars without parking: -6.5 %, exp U).
#if !defined(STMPM_NO_PARK)
                volatile double c_park[9];
#pragma unroll
                for (int k = 0; k < 9; ++k) c_park[k] = c.C[k];
#endif

                bool scanning = valid && use_list;
                int cnt = 0, n_det = 0;
                float mag_faintest = -INFINITY;
                double B[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
#if !defined(STMPM_SCAN_ONCE)
                if (scanning && scan_list<INSTR>(cf, sb_cat, lp, pk0, sb_row, cnt, ins)) scanning = false;
#endif

                for (;;) {
#if defined(STMPM_SCAN_ONCE)
                    // one copy of the eight-way unrolled pack evaluation for the first scan and the re-scans (exp P: the second
                    // copy is 248 instructions of a hot stream that is larger than the 32 KB L1.5 instruction cache)
                    if (scanning && scan_list<INSTR>(cf, sb_cat, lp, pk0, sb_row, cnt, ins)) scanning = false;
#endif
                    __syncwarp();
                    // ---- phase B: exact evaluation of the listed stars (lanes in lock-step).
                    // Software-pipelined: the 32-byte record of star j+1 is in flight while star j is evaluated.
#if defined(STMPM_PINGPONG)
                    // experiment (round 2, exp J): two record buffers used alternately by a two-way unrolled loop, instead of one
                    // landing buffer copied into the working one at the end of every iteration (8 MOVs of ~165 instructions)
                    auto star_step = [&](const double4& cur) {
                        StarRec s;
                        s.x = cur.x; s.y = cur.y; s.z = cur.z;
                        s.star_id = (uint32_t)(__double_as_longlong(cur.w) & 0xffffffffll);
                        if (INSTR == 2) s.mag = __int_as_float((int)(__double_as_longlong(cur.w) >> 32));
                        const bool mag_ok = __int_as_float(__double2hiint(cur.w)) <= cf.magf;
                        uint32_t x0, x1;
                        philox2x32_10(s.star_id, kb, ka, x0, x1);
                        bool det;
                        if (PREC == 64) det = star_exact(c, tabs, s, x0, x1, a.f_ideal, a.f_ideal2, a.half_u, a.half_v, B);
                        else det = star_f32(c, cf, tabs, s, x0, x1, a.f_ideal, a.f_ideal2, a.half_u, a.half_v, B);
                        n_det += (det && mag_ok) ? 1 : 0;
                        if (INSTR == 2 && det) mag_faintest = fmaxf(mag_faintest, s.mag);
                    };
                    double4 ra = make_double4(0.0, 0.0, 0.0, 0.0), rb = ra;
                    if (cnt > 0) ra = ld_rec(a.cat64 + lds_u16(sb_row));
                    for (int j = 0; j < cnt;) {
                        if (j + 1 < cnt) rb = ld_rec(a.cat64 + lds_u16(sb_row + (uint32_t)(j + 1) * (TPB * 2u)));
                        star_step(ra);
                        if (++j >= cnt) break;
                        if (j + 1 < cnt) ra = ld_rec(a.cat64 + lds_u16(sb_row + (uint32_t)(j + 1) * (TPB * 2u)));
                        star_step(rb);
                        ++j;
                    }
#elif defined(STMPM_REBALANCE)
                    // experiment (round 2, exp N): in-warp work donation.  The 32 trials of a group differ in survivor count (key
                    // sorting leaves the star loop at ~0.77 lane occupancy with 256-trial windows); once every list is complete,
                    // the lanes run their own stars up to K = 0.9 x the group's mean count, and at that warp-uniform point each
                    // lane that has finished takes over the second half of the remaining stars of one lane that has not: the
                    // m-th finished lane pairs with the lane of m-th largest remainder (rank by bit-sliced ballots, partner by
                    // match.any).  The helper's plane context, Philox key and thresholds are overwritten by the donor's through
                    // shuffles (its own are dead after its last star), its own B is parked; afterwards the partial B and count
                    // go back by shuffle.  Oracle simulation: occupancy 0.77 -> 0.89.  B's summation order then depends on the
                    // group: records are reproducible for a given launch geometry but no longer bit-identical across geometries.
                    {
                        constexpr unsigned FULL = 0xffffffffu;
                        bool do_rebal = !__any_sync(FULL, scanning);
                        int end = cnt;
                        if (do_rebal) {
                            const int K = ((int)__reduce_add_sync(FULL, (unsigned)cnt) * 29) >> 10;
                            do_rebal = K >= 6;
                            if (do_rebal) end = min(cnt, K);
                        }
                        int j = 0;
                        unsigned role = (unsigned)lane;          // partner lane | helper << 8 | donor << 9
                        volatile double b_park[9];
                        volatile int nd_park = 0;
                        volatile float mf_park = 0.0f;
                        double4 rr = make_double4(0.0, 0.0, 0.0, 0.0);
                        if (end > 0) rr = ld_rec(a.cat64 + lds_u16(sb_row));
#pragma unroll 1
                        for (int ph = 0; ph < 2; ++ph) {
                            if (ph == 1) {
                                const int rem = cnt - end;                                   // own stars not done yet
                                const bool idle = rem == 0 && (use_list || !valid);         // (a fall-back trial keeps its context)
                                const bool busy = rem >= 2;
                                const unsigned idle_m = __ballot_sync(FULL, idle), busy_m = __ballot_sync(FULL, busy);
                                j = end; end = cnt;
                                if (idle_m != 0u && busy_m != 0u) {
                                    unsigned eq = busy_m, gt = 0u;
#pragma unroll
                                    for (int b = 5; b >= 0; --b) {
                                        const bool mine = (rem >> b) & 1;
                                        const unsigned bm = __ballot_sync(FULL, busy && mine);
                                        if (!mine) gt |= eq & bm;
                                        eq &= mine ? bm : ~bm;
                                    }
                                    const unsigned lt_mask = (1u << lane) - 1u;
                                    const int rank = __popc(gt) + __popc(eq & ~lt_mask & ~(1u << lane));
                                    const int v = idle ? __popc(idle_m & lt_mask) : (busy ? rank : 64 + lane);
                                    const unsigned other = __match_any_sync(FULL, v) & ~(1u << lane);
                                    const int src = other ? (__ffs(other) - 1) : lane;
                                    const bool helper = idle && other != 0u, donor = busy && other != 0u;
                                    role = (unsigned)src | (helper ? 0x100u : 0u) | (donor ? 0x200u : 0u);
                                    const int p_cnt = __shfl_sync(FULL, cnt, src), p_rem = __shfl_sync(FULL, rem, src);
                                    const uint32_t p_row = __shfl_sync(FULL, sb_row, src);
#define STMPM_XCHG(x) { const auto t_ = __shfl_sync(FULL, x, src); if (helper) x = t_; }
#pragma unroll
                                    for (int k = 0; k < 3; ++k) { STMPM_XCHG(c.NI[k]) STMPM_XCHG(c.EU[k]) STMPM_XCHG(c.EV[k]) }
                                    STMPM_XCHG(c.np0) STMPM_XCHG(c.cu) STMPM_XCHG(c.cv) STMPM_XCHG(c.sx) STMPM_XCHG(c.sy)
                                    STMPM_XCHG(ka) STMPM_XCHG(kb) STMPM_XCHG(cf.magf)
                                    if (PREC == 32) { STMPM_XCHG(cf.sxf) STMPM_XCHG(cf.syf) STMPM_XCHG(cf.guard_hi) }
#undef STMPM_XCHG
                                    if (helper) {
#pragma unroll
                                        for (int k = 0; k < 9; ++k) { b_park[k] = B[k]; B[k] = 0.0; }
                                        nd_park = n_det; n_det = 0;
                                        if (INSTR == 2) { mf_park = mag_faintest; mag_faintest = -INFINITY; }
                                        sb_row = p_row; end = p_cnt; j = p_cnt - (p_rem >> 1);
                                    } else if (donor) {
                                        end = cnt - (rem >> 1);
                                    }
                                }
                                if (j < end) rr = ld_rec(a.cat64 + lds_u16(sb_row + (uint32_t)j * (TPB * 2u)));
                            }
                            for (; j < end; ++j) {
                                StarRec s;
                                s.x = rr.x; s.y = rr.y; s.z = rr.z;
                                s.star_id = (uint32_t)(__double_as_longlong(rr.w) & 0xffffffffll);
                                if (INSTR == 2) s.mag = __int_as_float((int)(__double_as_longlong(rr.w) >> 32));
                                const bool mag_ok = __int_as_float(__double2hiint(rr.w)) <= cf.magf;   // see the default loop below
                                if (j + 1 < end) rr = ld_rec(a.cat64 + lds_u16(sb_row + (uint32_t)(j + 1) * (TPB * 2u)));
                                uint32_t x0, x1;
                                philox2x32_10(s.star_id, kb, ka, x0, x1);
                                bool det;
                                if (PREC == 64) det = star_exact(c, tabs, s, x0, x1, a.f_ideal, a.f_ideal2, a.half_u, a.half_v, B);
                                else det = star_f32(c, cf, tabs, s, x0, x1, a.f_ideal, a.f_ideal2, a.half_u, a.half_v, B);
                                n_det += (det && mag_ok) ? 1 : 0;
                                if (INSTR == 2 && det) mag_faintest = fmaxf(mag_faintest, s.mag);
                            }
                            __syncwarp();
                            if (!do_rebal) break;
                        }
                        if (__any_sync(FULL, (role & 0x300u) != 0u)) {          // partial sums back to their trials
                            const int src = (int)(role & 31u);
                            const bool helper = (role & 0x100u) != 0u, donor = (role & 0x200u) != 0u;
#pragma unroll
                            for (int k = 0; k < 9; ++k) {
                                const double t = __shfl_sync(FULL, B[k], src);
                                if (donor) B[k] += t;
                            }
                            const int tn = __shfl_sync(FULL, n_det, src);
                            if (donor) n_det += tn;
                            if (INSTR == 2) {
                                const float tm = __shfl_sync(FULL, mag_faintest, src);
                                if (donor) mag_faintest = fmaxf(mag_faintest, tm);
                            }
                            if (helper) {
#pragma unroll
                                for (int k = 0; k < 9; ++k) B[k] = b_park[k];
                                n_det = nd_park;
                                if (INSTR == 2) mag_faintest = mf_park;
                                sb_row = sb + (uint32_t)OFF_LIST + threadIdx.x * 2u;
                            }
                        }
                    }
#else
                    if (TWO_STARS) {
                    // Two stars per iteration (round 2, exp U): two independent Philox / Box-Muller / projection chains side by
                    // side supply the instruction-level parallelism that three resident warps per scheduler do not; B is
                    // accumulated in star order, so every record is what the single-star loop (below, -DSTMPM_ONE_STAR) writes.
                    // The records of stars j+2, j+3 are in flight while j, j+1 are evaluated.
                        double4 r0 = make_double4(0.0, 0.0, 0.0, 0.0), r1 = r0;
                        if (cnt > 0) r0 = ld_rec(a.cat64 + lds_u16(sb_row));
                        if (cnt > 1) r1 = ld_rec(a.cat64 + lds_u16(sb_row + (uint32_t)(TPB * 2u)));
#if defined(STMPM_STAR_UNROLL)
#pragma unroll 2
#endif
                        for (int j = 0; j < cnt; j += 2) {
                            const bool has1 = j + 1 < cnt;     // odd count: the second slot re-evaluates a stale record, result dropped
                            StarRec s0, s1;
                            s0.x = r0.x; s0.y = r0.y; s0.z = r0.z;
                            s0.star_id = (uint32_t)(__double_as_longlong(r0.w) & 0xffffffffll);
                            s1.x = r1.x; s1.y = r1.y; s1.z = r1.z;
                            s1.star_id = (uint32_t)(__double_as_longlong(r1.w) & 0xffffffffll);
                            if (INSTR == 2) {
                                s0.mag = __int_as_float((int)(__double_as_longlong(r0.w) >> 32));
                                s1.mag = __int_as_float((int)(__double_as_longlong(r1.w) >> 32));
                            }
                            // the magnitude word takes part in the decision: see the single-star loop (keeps it alive)
                            const bool ok0 = __int_as_float(__double2hiint(r0.w)) <= cf.magf;
                            const bool ok1 = __int_as_float(__double2hiint(r1.w)) <= cf.magf;
#if defined(STMPM_UNCOND_PREFETCH)
                            // experiment (round 2, exp Y): unconditional prefetch of clamped list entries.  A predicated load has
                            // to preserve its destination when the predicate is false, and slot 1's predicate (j + 3 < cnt) is not
                            // the loop's: ptxas copies the record out and back around the load (16 moves per iteration)
                            r0 = ld_rec(a.cat64 + lds_u16(sb_row + (uint32_t)min(j + 2, cnt - 1) * (TPB * 2u)));
                            r1 = ld_rec(a.cat64 + lds_u16(sb_row + (uint32_t)min(j + 3, cnt - 1) * (TPB * 2u)));
#else
                            if (j + 2 < cnt) r0 = ld_rec(a.cat64 + lds_u16(sb_row + (uint32_t)(j + 2) * (TPB * 2u)));
                            if (j + 3 < cnt) r1 = ld_rec(a.cat64 + lds_u16(sb_row + (uint32_t)(j + 3) * (TPB * 2u)));
#endif
                            uint32_t x00, x01, x10, x11;
                            philox2x32_10(s0.star_id, kb, ka, x00, x01);
                            philox2x32_10(s1.star_id, kb, ka, x10, x11);
                            double b0x, b0y, b0z, b1x, b1y, b1z;
                            bool d0, d1;
                            if (PREC == 64) {
                                d0 = star_eval64(c, tabs, s0, x00, x01, a.f_ideal, a.f_ideal2, a.half_u, a.half_v, b0x, b0y, b0z);
                                d1 = star_eval64(c, tabs, s1, x10, x11, a.f_ideal, a.f_ideal2, a.half_u, a.half_v, b1x, b1y, b1z);
                            } else {
                                bool g0, g1;
                                d0 = star_eval32(c, cf, tabs, s0, x00, x01, a.f_ideal, a.f_ideal2, a.half_u, a.half_v, b0x, b0y, b0z, g0);
                                d1 = star_eval32(c, cf, tabs, s1, x10, x11, a.f_ideal, a.f_ideal2, a.half_u, a.half_v, b1x, b1y, b1z, g1);
                                if (g0) d0 = star_eval64(c, tabs, s0, x00, x01, a.f_ideal, a.f_ideal2, a.half_u, a.half_v, b0x, b0y, b0z);
                                if (g1) d1 = star_eval64(c, tabs, s1, x10, x11, a.f_ideal, a.f_ideal2, a.half_u, a.half_v, b1x, b1y, b1z);
                            }
                            d1 = d1 && has1;
                            if (d0) star_accumulate(B, s0, b0x, b0y, b0z);
                            if (d1) star_accumulate(B, s1, b1x, b1y, b1z);
                            n_det += ((d0 && ok0) ? 1 : 0) + ((d1 && ok1) ? 1 : 0);
                            if (INSTR == 2 && d0) mag_faintest = fmaxf(mag_faintest, s0.mag);
                            if (INSTR == 2 && d1) mag_faintest = fmaxf(mag_faintest, s1.mag);
                        }
                    } else {
                    double4 rr = make_double4(0.0, 0.0, 0.0, 0.0);
                    if (cnt > 0) rr = ld_rec(a.cat64 + lds_u16(sb_row));
#if defined(STMPM_IDX_AHEAD)
                    // experiment (round 2, exp G): the survivor INDEX of star j+2 read a whole iteration before the record load
                    // that needs it, to take the shared-memory latency out of the load's address chain (3 - 4.5 % of the stall
                    // samples).  Measured -6 % (2.70 vs 2.86e9): the extra live register costs more than the wait.  Off.
                    uint32_t i_nx = 0u;
                    if (cnt > 1) i_nx = lds_u16(sb_row + (uint32_t)(TPB * 2u));
#endif
                    for (int j = 0; j < cnt; ++j) {
                        StarRec s;
                        s.x = rr.x; s.y = rr.y; s.z = rr.z;
                        s.star_id = (uint32_t)(__double_as_longlong(rr.w) & 0xffffffffll);
                        if (INSTR == 2) s.mag = __int_as_float((int)(__double_as_longlong(rr.w) >> 32));
                        // The record's last word (the magnitude) takes part in the detection decision below.  Besides being the
                        // exact magnitude test once more, this keeps that word ALIVE: when it was unused, ptxas treated the
                        // corresponding destination register of the NEXT record's 256-bit load as dead and reused it as a Philox
                        // temporary right behind the load — a write-after-write hazard that exposed the full L2 latency in every
                        // iteration (round 2, fp32 instantiation: 19 % of all stall samples on one IADD3).
                        const bool mag_ok = __int_as_float(__double2hiint(rr.w)) <= cf.magf;   // (an immediate compare instead: +-0, exp H)
#if defined(STMPM_IDX_AHEAD)
                        if (j + 1 < cnt) rr = ld_rec(a.cat64 + i_nx);
                        if (j + 2 < cnt) i_nx = lds_u16(sb_row + (uint32_t)(j + 2) * (TPB * 2u));
#else
                        if (j + 1 < cnt) rr = ld_rec(a.cat64 + lds_u16(sb_row + (uint32_t)(j + 1) * (TPB * 2u)));
#endif
                        // (hand-pipelining Philox one star ahead, or two stars per iteration, both measured slower AT 128
                        // REGISTERS — extra live state spills; with 168 registers the two-star loop above wins)
                        uint32_t x0, x1;
                        philox2x32_10(s.star_id, kb, ka, x0, x1);
                        bool det;
                        if (PREC == 64) det = star_exact(c, tabs, s, x0, x1, a.f_ideal, a.f_ideal2, a.half_u, a.half_v, B);
                        else det = star_f32(c, cf, tabs, s, x0, x1, a.f_ideal, a.f_ideal2, a.half_u, a.half_v, B);
                        n_det += (det && mag_ok) ? 1 : 0;
                        if (INSTR == 2 && det) mag_faintest = fmaxf(mag_faintest, s.mag);
                    }
                    }
#endif
                    cnt = 0;
                    if (!__any_sync(0xffffffffu, scanning)) break;
                    if (scanning) {                             // survivor row was full: re-derive the fp32 geometry, go on
                        geom32(c, cf);
#if defined(STMPM_SCAN_ONCE)
                        pk0 = ld_pack(lp);
#else
                        if (scan_list<INSTR>(cf, sb_cat, lp, ld_pack(lp), sb_row, cnt, ins))
                            scanning = false;
#endif
                    }
                }
#if !defined(STMPM_NO_PARK)
#pragma unroll
                for (int k = 0; k < 9; ++k) c.C[k] = c_park[k];
#endif
                // ---- fallback: scan cone wider than the lists were built for -> whole trial on the sky-grid band walk
                if (valid && !use_list) {
                    Ctx64 c2 = c;
                    Ctx32 cf2 = cf;
                    geom32(c, cf2);
                    boresight32(c, cf2);
                    double B2[9];
                    int nd2 = 0;
                    float mf2 = -INFINITY;
                    band_walk_trial<PREC, INSTR>(&c2, &cf2, a.cat64, a.cell_start, a.n_bands, a.inv_band_h, a.f_ideal, a.half_u,
                                                 a.half_v, sb, ka, kb, B2, &nd2, &mf2, INSTR == 1 ? ins : nullptr);
#pragma unroll
                    for (int k = 0; k < 9; ++k) B[k] = B2[k];
                    n_det = nd2;
                    mag_faintest = mf2;
                }

                // ---- solve + record
                double err = -1.0;
                if (n_det >= a.n_min) err = quest_error_arcsec(B, c.C, n_det);
                if (valid) {
                    if (INSTR == 1) ins[3] += (unsigned long long)n_det;
                    if (a.err) a.err[rec] = err;
                    if (a.ndet) a.ndet[rec] = n_det;
                    if (a.ndet8) a.ndet8[rec] = (uint8_t)min(n_det, 255);
                    if (a.err32) a.err32[rec] = (float)err;
                    if (INSTR == 2 && a.maxmag) a.maxmag[rec] = mag_faintest;
                    if (want_stats) {
                        acc_ndet += (unsigned long long)n_det;
                        if (err >= 0.0) {
                            ++acc_ok;
                            acc_x += err;
                            acc_xx = fma(err, err, acc_xx);
                            acc_max = fmax(acc_max, err);
                            const long long e = (__double_as_longlong(err) >> a.hist_shift)
                                                - ((long long)(1023 + STMPM_HIST_MIN_EXP) << (52 - a.hist_shift));
                            // log-linear histogram: one fire-and-forget RED.ADD.F64 per successful trial, straight into
                            // the packed statistics in L2 (bins are spread over ~10^3 addresses: no hot spot)
                            double* st = a.stats + seg * a.stats_len;
                            const int slot = e < 0 ? 6 : (e >= a.n_bins ? 7 : STMPM_STATS_HEAD + (int)e);
                            red_add_f64(st + slot, 1.0);
                        } else {
                            ++acc_fail;
                        }
                    }
                }
            }
        }
    }

    if (aborted && a.status != nullptr && lane == 0) *reinterpret_cast<volatile int*>(a.status) = STMPM_E_WATCHDOG;
    if (INSTR == 1) {
#pragma unroll
        for (int k = 0; k < 4; ++k) atomicAdd(a.counters + k, ins[k]);
    }
    if (want_stats && cur_seg >= 0)
        flush_moments(a.stats + cur_seg * a.stats_len, lane, acc_ok, acc_fail, acc_ndet, acc_x, acc_xx, acc_max);
}

#include "stmpm_split.cuh"

// ------------------------------------------------------------------------------------------------ generator kernel
__global__ void __launch_bounds__(256) fill_presampled_kernel(long long n, long long trial0, uint32_t k0, uint32_t k1,
                                                              DistDev dist, const MathTables* __restrict__ tab, double* c0, double* c1, double* c2,
                                                              double* c3, double* c4, double* c5, double* c6,
                                                              double* c7, double* c8, double* c9, double* c10,
                                                              double* c11, double* c12) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const unsigned long long g = (unsigned long long)(trial0 + t);
        TrialParams p;
        sample_params(dist, tab, (uint32_t)g, (uint32_t)(g >> 32), k0, k1, p);
        c0[t] = p.ra; c1[t] = p.dec; c2[t] = p.roll; c3[t] = p.f; c4[t] = p.ex; c5[t] = p.ey; c6[t] = p.ez;
        c7[t] = p.phi; c8[t] = p.theta; c9[t] = p.psi; c10[t] = p.devx; c11[t] = p.devy; c12[t] = p.maxmag;
    }
}

// ------------------------------------------------------------------------------------------------ toolplot reductions
// toolplot.py:82-99 on the device, from the per-trial error records (failed trials, err < 0, are the rows the reference
// dropped).  acc layout (doubles): [0] n, [1] sum x^2 | [2] n_kept, [3] sum x, [4] sum x^2, [5] min bits, [6] max bits.
__global__ void __launch_bounds__(256) toolplot_pass1(const double* __restrict__ x, long long n, double* acc) {
    double c = 0.0, s2 = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double v = __ldcs(x + i);
        if (v >= 0.0) { c += 1.0; s2 = fma(v, v, s2); }
    }
    for (int off = 16; off > 0; off >>= 1) { c += __shfl_xor_sync(0xffffffffu, c, off); s2 += __shfl_xor_sync(0xffffffffu, s2, off); }
    if ((threadIdx.x & 31) == 0 && c > 0.0) { atomicAdd(acc + 0, c); atomicAdd(acc + 1, s2); }
}
__global__ void __launch_bounds__(256) toolplot_pass2(const double* __restrict__ x, long long n, double* acc) {
    const double clip = 6.0 * sqrt(acc[1] / acc[0]);                       // toolplot.py:82-83
    double c = 0.0, s1 = 0.0, s2 = 0.0, mn = INFINITY, mx = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double v = __ldcs(x + i);
        if (v >= 0.0 && v <= clip) { c += 1.0; s1 += v; s2 = fma(v, v, s2); mn = fmin(mn, v); mx = fmax(mx, v); }
    }
    for (int off = 16; off > 0; off >>= 1) {
        c += __shfl_xor_sync(0xffffffffu, c, off); s1 += __shfl_xor_sync(0xffffffffu, s1, off);
        s2 += __shfl_xor_sync(0xffffffffu, s2, off);
        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, off)); mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    }
    if ((threadIdx.x & 31) == 0 && c > 0.0) {
        atomicAdd(acc + 2, c); atomicAdd(acc + 3, s1); atomicAdd(acc + 4, s2);
        atomicMin(reinterpret_cast<unsigned long long*>(acc + 5), (unsigned long long)__double_as_longlong(mn));   // x >= 0:
        atomicMax(reinterpret_cast<unsigned long long*>(acc + 6), (unsigned long long)__double_as_longlong(mx));   // bit order == value order
    }
}
__global__ void __launch_bounds__(256) toolplot_pass3(const double* __restrict__ x, long long n, const double* acc, int bins,
                                                      unsigned long long* hist) {
    extern __shared__ unsigned int s_h[];
    for (int i = threadIdx.x; i < bins; i += blockDim.x) s_h[i] = 0u;
    __syncthreads();
    const double clip = 6.0 * sqrt(acc[1] / acc[0]);
    const double lo = __longlong_as_double((long long)reinterpret_cast<const unsigned long long*>(acc)[5]);
    const double hi = __longlong_as_double((long long)reinterpret_cast<const unsigned long long*>(acc)[6]);
    const double step = (hi - lo) / (double)bins;                          // numpy.linspace(lo, hi, bins + 1)
    const double norm = hi > lo ? (double)bins / (hi - lo) : 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double v = __ldcs(x + i);
        if (v >= 0.0 && v <= clip) {
            // numpy.histogram: provisional index from the scaled position, then corrected against the actual edges
            int b = (int)((v - lo) * norm);
            if (b >= bins) b = bins - 1;
            const double e0 = lo + (double)b * step, e1 = (b + 1 == bins) ? hi : lo + (double)(b + 1) * step;
            if (v < e0 && b > 0) --b;
            else if (v >= e1 && b != bins - 1) ++b;
            atomicAdd(&s_h[b], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < bins; i += blockDim.x) if (s_h[i]) atomicAdd(hist + i, (unsigned long long)s_h[i]);
}


// ------------------------------------------------------------------------------------------------ MonteCarlo stop rule
// MonteCarlo.run_sim's convergence loop (/root/reference/monte_carlo.py:45-76) evaluated ON THE DEVICE.  The trials of a
// chunk of batches are computed speculatively by one trials_kernel launch (they are independent of the rule); this
// single-CTA kernel then replays the reference's sequential rule over the chunk's batches in parallel: per-batch moments
// (thread = batch, fixed summation order), a block-wide prefix sum with the carry of the earlier chunks, sigma_hat after
// every batch, the first batch k with  not(ratio_k > tol or n_k < min_rows)  or  (max_runs > 0 and n_k >= max_runs).
// Trials after the stopping batch are discarded: the statistics cover exactly the batches the reference's loop ran.
struct McState {
    double n_ok, n_fail, sum, sum2, sum_ndet, x_max;   // totals over the batches accepted so far
    double prev, ratio;                                // monte_carlo.py:46,60-61
    long long count;                                   // batches run (MonteCarlo.count, monte_carlo.py:50,79)
    int done, pad;
};
constexpr int MC_MAX_BATCHES = 1024;                   // batches per chunk = threads of the rule kernel

__global__ void __launch_bounds__(MC_MAX_BATCHES) mc_rule_kernel(const double* __restrict__ err, const int* __restrict__ ndet,
                                                                 int n_batches, int batch, McState* st, double* stats, int hist_shift,
                                                                 int n_bins, long long min_rows, double tol, long long max_runs) {
    __shared__ double s_n[32], s_s1[32], s_s2[32];
    __shared__ double s_sig[MC_MAX_BATCHES];
    __shared__ int s_stop;
    if (st->done) return;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    double n = 0.0, s1 = 0.0, s2 = 0.0;
    if (t < n_batches) {
        const double* e = err + (long long)t * batch;
        for (int i = 0; i < batch; ++i) {
            const double v = e[i];
            if (v >= 0.0) { n += 1.0; s1 += v; s2 = fma(v, v, s2); }
        }
    }
    if (t == 0) s_stop = n_batches;
    // inclusive prefix sums over the batches of the chunk (fixed order: warp scan, then the warp totals)
    double pn = n, p1 = s1, p2 = s2;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const double a = __shfl_up_sync(0xffffffffu, pn, off), b = __shfl_up_sync(0xffffffffu, p1, off),
                     c = __shfl_up_sync(0xffffffffu, p2, off);
        if (lane >= off) { pn += a; p1 += b; p2 += c; }
    }
    if (lane == 31) { s_n[warp] = pn; s_s1[warp] = p1; s_s2[warp] = p2; }
    __syncthreads();
    double cn = st->n_ok, c1 = st->sum, c2 = st->sum2;
    for (int w = 0; w < warp; ++w) { cn += s_n[w]; c1 += s_s1[w]; c2 += s_s2[w]; }
    pn += cn; p1 += c1; p2 += c2;                       // totals after batch t
    // sigma_hat = std(ddof = 1) / sqrt(1 - 2/pi)   (monte_carlo.py:60-61; driver.py:290)
    const double var = pn > 1.0 ? fmax(0.0, (p2 - p1 * p1 / pn) / (pn - 1.0)) : 0.0;
    const double sig = sqrt(var) / 0.6028102749890869;  // sqrt(1 - 2/pi)
    if (t < n_batches) s_sig[t] = sig;
    __syncthreads();
    const double prev = t == 0 ? st->prev : s_sig[t - 1];
    const double ratio = fabs((prev - sig) / prev);
    bool stop = false;
    if (t < n_batches) stop = !(ratio > tol || pn < (double)min_rows) || (max_runs > 0 && pn >= (double)max_runs);
    if (stop) atomicMin(&s_stop, t);
    __syncthreads();
    const int k = s_stop;                               // first stopping batch, or n_batches
    const int last = min(k, n_batches - 1);             // last batch accepted from this chunk
    // statistics of the accepted records: failures, n_det, maximum, histogram (integer counts: order-independent)
    double nf = 0.0, snd = 0.0, xm = 0.0;
    const long long n_rec = (long long)(last + 1) * batch;
    for (long long i = t; i < n_rec; i += blockDim.x) {
        const double v = err[i];
        snd += (double)ndet[i];
        if (v >= 0.0) {
            xm = fmax(xm, v);
            const long long e = (__double_as_longlong(v) >> hist_shift) - ((long long)(1023 + STMPM_HIST_MIN_EXP) << (52 - hist_shift));
            const int slot = e < 0 ? 6 : (e >= n_bins ? 7 : STMPM_STATS_HEAD + (int)e);
            atomicAdd(stats + slot, 1.0);
        } else {
            nf += 1.0;
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        nf += __shfl_xor_sync(0xffffffffu, nf, off); snd += __shfl_xor_sync(0xffffffffu, snd, off);
        xm = fmax(xm, __shfl_xor_sync(0xffffffffu, xm, off));
    }
    __syncthreads();                                    // s_n / s_s1 / s_s2 are re-used below
    if (lane == 0) { s_n[warp] = nf; s_s1[warp] = snd; s_s2[warp] = xm; }
    __syncthreads();
    if (t == last) {                                    // this thread holds the totals after the last accepted batch
        double tf = st->n_fail, td = st->sum_ndet, tm = st->x_max;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { tf += s_n[w]; td += s_s1[w]; tm = fmax(tm, s_s2[w]); }
        st->n_ok = pn; st->sum = p1; st->sum2 = p2; st->n_fail = tf; st->sum_ndet = td; st->x_max = tm;
        st->prev = sig; st->ratio = ratio; st->count += last + 1;
        stats[0] = pn; stats[1] = tf; stats[2] = p1; stats[3] = p2; stats[4] = td; stats[5] = tm;
        __threadfence();
        if (k < n_batches) st->done = 1;
    }
}

// ------------------------------------------------------------------------------------------------ toolplot modes s, c
// toolplot.py:170-187 (mode 's'): per unique MAX_MAGNITUDE value, the share of rows that failed (error stored as -1),
// among the rows the 6-sigma clip of toolplot.py:82-83 keeps (std = RMS over ALL rows, sentinels included, as written
// there).  levels: the sorted unique magnitudes; rows whose magnitude matches no level are counted in slot n_levels.
__global__ void __launch_bounds__(256) toolplot_rms_all(const double* __restrict__ x, long long n, double* acc) {
    double s2 = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double v = __ldcs(x + i);
        s2 = fma(v, v, s2);
    }
    for (int off = 16; off > 0; off >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, off);
    if ((threadIdx.x & 31) == 0) atomicAdd(acc, s2);
}
__global__ void __launch_bounds__(256) toolplot_failrate_kernel(const double* __restrict__ x, const double* __restrict__ mag, long long n,
                                                                const double* __restrict__ levels, int n_levels, const double* acc,
                                                                unsigned long long* rows, unsigned long long* failed) {
    const double clip = 6.0 * sqrt(acc[0] / (double)n);                    // toolplot.py:82-83
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double v = __ldcs(x + i);
        if (!(v <= clip)) continue;
        const double m = __ldcs(mag + i);
        int lo = 0, hi = n_levels;                                         // first level >= m
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (levels[mid] < m) lo = mid + 1; else hi = mid; }
        const int slot = (lo < n_levels && levels[lo] == m) ? lo : n_levels;
        atomicAdd(rows + slot, 1ull);
        if (v < 0.0) atomicAdd(failed + slot, 1ull);                       // toolplot.py:170-172: -1 -> 1, everything else -> 0
    }
}
// toolplot.py:240-256 (mode 'c'): min, max, std and the 50-bin density histogram of a magnitude column (float32 on the
// device: the star_mag objective's output; -inf = no star detected = a row the reference never had).
__global__ void __launch_bounds__(256) column_pass1(const float* __restrict__ x, long long n, double* acc) {
    double c = 0.0, s1 = 0.0, s2 = 0.0, mn = INFINITY, mx = -INFINITY;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double v = (double)__ldcs(x + i);
        if (v > -INFINITY && v < INFINITY) { c += 1.0; s1 += v; s2 = fma(v, v, s2); mn = fmin(mn, v); mx = fmax(mx, v); }
    }
    for (int off = 16; off > 0; off >>= 1) {
        c += __shfl_xor_sync(0xffffffffu, c, off); s1 += __shfl_xor_sync(0xffffffffu, s1, off);
        s2 += __shfl_xor_sync(0xffffffffu, s2, off);
        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, off)); mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    }
    if ((threadIdx.x & 31) == 0 && c > 0.0) {
        atomicAdd(acc + 0, c); atomicAdd(acc + 1, s1); atomicAdd(acc + 2, s2);
        // order-preserving key of a double: flip all bits of negatives, the sign bit of non-negatives
        const unsigned long long bn = (unsigned long long)__double_as_longlong(mn), bx = (unsigned long long)__double_as_longlong(mx);
        atomicMin(reinterpret_cast<unsigned long long*>(acc + 3), (bn >> 63) ? ~bn : (bn | 0x8000000000000000ull));
        atomicMax(reinterpret_cast<unsigned long long*>(acc + 4), (bx >> 63) ? ~bx : (bx | 0x8000000000000000ull));
    }
}
__device__ __forceinline__ double key_to_double(unsigned long long k) {
    return __longlong_as_double((long long)((k >> 63) ? (k & 0x7fffffffffffffffull) : ~k));
}
__global__ void __launch_bounds__(256) column_pass2(const float* __restrict__ x, long long n, const double* acc, int bins,
                                                    unsigned long long* hist) {
    extern __shared__ unsigned int s_h[];
    for (int i = threadIdx.x; i < bins; i += blockDim.x) s_h[i] = 0u;
    __syncthreads();
    const double lo = key_to_double(reinterpret_cast<const unsigned long long*>(acc)[3]);
    const double hi = key_to_double(reinterpret_cast<const unsigned long long*>(acc)[4]);
    const double step = (hi - lo) / (double)bins;                          // numpy.linspace(lo, hi, bins + 1)
    const double norm = hi > lo ? (double)bins / (hi - lo) : 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double v = (double)__ldcs(x + i);
        if (v > -INFINITY && v < INFINITY) {
            int b = (int)((v - lo) * norm);                                // numpy.histogram's index, corrected against the edges
            if (b >= bins) b = bins - 1;
            const double e0 = lo + (double)b * step, e1 = (b + 1 == bins) ? hi : lo + (double)(b + 1) * step;
            if (v < e0 && b > 0) --b;
            else if (v >= e1 && b != bins - 1) ++b;
            atomicAdd(&s_h[b], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < bins; i += blockDim.x) if (s_h[i]) atomicAdd(hist + i, (unsigned long long)s_h[i]);
}

// ------------------------------------------------------------------------------------------------ FMA peak kernels
template <typename T>
__global__ void __launch_bounds__(256) fma_peak_kernel(T* out, int iters, T a, T b) {
    T x[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = (T)(threadIdx.x + k);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = x[k] * a + b;   // contracted to FMA
    }
    T s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += x[k];
    if (s == (T)12345.678) out[0] = s;   // keep the chain alive without a store in practice
}

// ------------------------------------------------------------------------------------------------ host side
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return fail((int)e_, std::string(#call) + ": " + cudaGetErrorString(e_));                  \
    } while (0)

// Every entry point runs on the handle's device and puts the caller's current device back on the way out (the library
// shares the CUDA runtime with torch: silently changing the current device would redirect the caller's next allocation).
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != device) cudaSetDevice(device); else prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};

static constexpr long long HOST_CHUNK = 1LL << 22;   // default trials per stage of the host pipelines (stmpm_set_host_pipeline)

struct stmpm_handle {
    int device = 0;
    cudaDeviceProp prop{};
    // catalog
    float4* cat32 = nullptr;
    StarRec* cat64 = nullptr;
    uint16_t* cell_start = nullptr;
    int2* band_tab = nullptr;
    int n_stars = 0, n_cells_pad = 0, n_bands = 0;
    std::vector<float> h_xyzf;    // host copy of the sorted catalog directions + magnitudes (candidate-list builds)
    std::vector<float> h_mag;
    int2* lk_band_tab = nullptr;
    uint32_t* lk_off = nullptr;
    uint16_t* lk_list = nullptr;
    int lk_bands = 0;
    float lk_cone = -1.0f;
    MathTables* tables = nullptr;
    uint8_t* lk_key = nullptr;    // footprint work-key table [KEY_LEVELS][n_lk_cells][KEY_ROLLS]
    uint8_t* lk_key_disc = nullptr;   // disc work-key table [KEY_LEVELS][n_lk_cells]
    int lk_cells = 0;
    int tune_key_table = -1;      // -1 = per mode (disc for pre-sampled, footprint for native RNG), 0 = disc, 1 = footprint
    float key_roll_inv = 0.63661977f;
    bool have_catalog = false, have_sensor = false;
    stmpm_sensor_cfg cfg{};
    int n_bins = 0;
    // host pipeline
    static constexpr int NBUF = 3;
    cudaStream_t streams[NBUF] = {nullptr, nullptr, nullptr};
    double* stage_cols[NBUF] = {nullptr, nullptr, nullptr};
    double* stage_err[NBUF] = {nullptr, nullptr, nullptr};
    int* stage_ndet[NBUF] = {nullptr, nullptr, nullptr};
    float* stage_mag[NBUF] = {nullptr, nullptr, nullptr};
    long long stage_cap = 0;
    double* stage_stats = nullptr;
    long long stage_stats_cap = 0;
    long long launches = 0;
    unsigned long long* counters = nullptr;
    // tuning (stmpm_set_tuning): 0 / negative = the measured defaults
    int tune_window = 0;
    double tune_key_radius_deg = -1.0;
    int tune_stage = 0;                   // sorted parameter staging for pre-sampled launches: measured slower, off by default
    int* status_host = nullptr;           // pinned + mapped: [0] the kernels' watchdog word (stmpm_check), [1] split-pipeline fallbacks seen
    int* status_dev = nullptr;
    int tune_pipeline = -1;               // -1 auto, 0 single kernel, 1 two-kernel split pipeline (pre-sampled, one segment)
    cudaMemPool_t pool = nullptr;         // stream-ordered per-launch storage (staging scratch, distribution arrays)
    long long host_chunk = HOST_CHUNK;    // trials per stage of the host pipelines
};

// Tables of box_muller_tab (stmpm_device.cuh), computed with the host libm.
static void stmpm_host_tables(MathTables* t) {
    for (int j = 0; j < 256; ++j) {
        const double c = 1.0 + (j + 0.5) / 256.0;
        t->log_tab[j][0] = 1.0 / c;
        t->log_tab[j][1] = -2.0 * std::log(c);
    }
    for (int k = 0; k < 34; ++k) t->k_ln2[k] = 2.0 * k * M_LN2;
    for (int j = 0; j < 1024; ++j) {
        const double a = 2.0 * M_PI * (j + 0.5) / 1024.0;
        t->sc_tab[j][0] = std::sin(a);
        t->sc_tab[j][1] = std::cos(a);
    }
    for (int j = 0; j < 128; ++j) {
        const double a = 2.0 * M_PI * (j + 0.5) / 128.0;
        t->sc32[j][0] = (float)std::sin(a);
        t->sc32[j][1] = (float)std::cos(a);
    }
}

// Candidate lists: for every cell of a ~1 degree boresight lookup grid, the indices (into the sorted catalog) of all stars
// within lk_cone + cell radius of the cell centre, ordered by magnitude (brightest first).  A trial whose boresight falls
// in the cell and whose scan cone is <= lk_cone finds every star that can reach its sensor in this list.
static int rebuild_lists(stmpm_handle* h) {
    if (!h->have_catalog || !h->have_sensor) return 0;
    const stmpm_sensor_cfg& cfg = h->cfg;
    const double pad = cfg.scan_pad_px > 0 ? cfg.scan_pad_px : 32.0;
    const double cone = std::atan(std::sqrt(2.0) * (std::max(cfg.half_u, cfg.half_v) + pad) / cfg.f_ideal) + 0.002;
    const int S = h->n_stars;
    // Two work-key tables (the key = the number of stars a trial is expected to evaluate; it only orders the work):
    //  * DISC table [cell][level] (5 MB, L2-resident): stars inside a disc of 1.07 x the radius of the array's area around the
    //    cell centre (measured optimum, flat within +-10 %: scripts/sweep_keyradius.sh; stmpm_set_tuning overrides the radius,
    //    0 = count the whole candidate list).  Used by pre-sampled launches: their 256-trial windows turn over every eight
    //    groups, so the key pass is latency-critical and must hit L2.
    //  * FOOTPRINT table [cell][KEY_ROLLS][level] (42 MB): the stars ON THE ARRAY — the ideal array centred on the cell, in
    //    KEY_ROLLS orientation bins of the roll angle (a square array repeats every 90 degrees).  The disc / square mismatch is
    //    the larger part of the spread of the actual star count at a given key (star-loop lane occupancy in 2 048-trial
    //    windows, simulated on the oracle: 0.82 -> 0.86).  Used by native-RNG launches (2 048-trial windows: one key pass per
    //    64 groups, its DRAM misses are hidden): +2.4 % measured; on pre-sampled launches it measured -4.7 % (round 2).
    double key_radius = std::atan(1.07 * std::sqrt(4.0 * cfg.half_u * cfg.half_v / M_PI) / cfg.f_ideal);
    if (h->tune_key_radius_deg >= 0.0) key_radius = h->tune_key_radius_deg * M_PI / 180.0;
    const double roll_period = (cfg.half_u == cfg.half_v) ? M_PI / 2 : M_PI;
    h->key_roll_inv = (float)(1.0 / roll_period);
    float roll_c[KEY_ROLLS], roll_s[KEY_ROLLS];
    for (int k = 0; k < KEY_ROLLS; ++k) {
        roll_c[k] = (float)std::cos((k + 0.5) * roll_period / KEY_ROLLS);
        roll_s[k] = (float)std::sin((k + 0.5) * roll_period / KEY_ROLLS);
    }
    const float tan_u = (float)(cfg.half_u / cfg.f_ideal), tan_v = (float)(cfg.half_v / cfg.f_ideal);
    // stars in magnitude order, SoA floats
    std::vector<int> order(S);
    for (int i = 0; i < S; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return h->h_mag[a] < h->h_mag[b]; });
    std::vector<float> sx(S), sy(S), sz(S), dots(S);
    for (int i = 0; i < S; ++i) {
        sx[i] = h->h_xyzf[3 * order[i]]; sy[i] = h->h_xyzf[3 * order[i] + 1]; sz[i] = h->h_xyzf[3 * order[i] + 2];
    }
    const int nb = LK_BANDS;
    const double bh = M_PI / nb;
    std::vector<int2> bt(nb);
    long long n_cells = 0;
    for (int b = 0; b < nb; ++b) {
        const double dm = -M_PI / 2 + (b + 0.5) * bh;
        const int n_ra = std::max(1, (int)std::lround(2.0 * M_PI * std::cos(dm) / bh));
        bt[b] = make_int2((int)n_cells, n_ra);
        n_cells += n_ra;
    }
    std::vector<uint32_t> off(n_cells + 1, 0);
    std::vector<uint16_t> list;
    list.reserve((size_t)n_cells * 160);
    std::vector<uint8_t> key_tab((size_t)n_cells * KEY_ROLLS * KEY_LEVELS, 0);   // footprint stars per cell, roll bin and magnitude level
    std::vector<uint8_t> key_disc((size_t)n_cells * KEY_LEVELS, 0);               // disc stars per cell and magnitude level
    std::vector<float> smag(S);
    for (int i = 0; i < S; ++i) smag[i] = h->h_mag[order[i]];
    for (int b = 0; b < nb; ++b) {
        const double d0 = -M_PI / 2 + b * bh, d1 = d0 + bh, dm = 0.5 * (d0 + d1);
        const int n_ra = bt[b].y;
        const double w = 2.0 * M_PI / n_ra;
        // cell radius: farthest boundary point from the centre (same for every cell of the band), sampled + 10 %
        double rho = 0.0;
        for (int e = 0; e <= 16; ++e) {
            const double ra = -0.5 * w + w * e / 16.0;
            for (double dd : {d0, d1}) {
                const double cd = std::cos(dm) * std::cos(dd) * std::cos(ra) + std::sin(dm) * std::sin(dd);
                rho = std::max(rho, std::acos(std::min(1.0, std::max(-1.0, cd))));
            }
            const double dd = d0 + bh * e / 16.0;
            const double cd = std::cos(dm) * std::cos(dd) * std::cos(0.5 * w) + std::sin(dm) * std::sin(dd);
            rho = std::max(rho, std::acos(std::min(1.0, std::max(-1.0, cd))));
        }
        const double reach = std::min(M_PI, cone + 1.1 * rho + 0.002);
        const float thresh = (float)std::cos(reach) - 4e-6f;
        // (disc variant of the key: stars inside a disc of `key_radius` around the cell centre)
        const float key_thresh = key_radius > 0 ? (float)std::cos(std::min(reach, key_radius)) : thresh;
        for (int rc = 0; rc < n_ra; ++rc) {
            const double ra = (rc + 0.5) * w;
            const float cx = (float)(std::cos(dm) * std::cos(ra)), cy = (float)(std::cos(dm) * std::sin(ra)),
                        cz = (float)std::sin(dm);
            for (int i = 0; i < S; ++i) dots[i] = cx * sx[i] + cy * sy[i] + cz * sz[i];
            const size_t start = list.size();
            {
                // body axes of a boresight at the cell centre with roll 0: C = Rz(ra) Ry(pi/2 - dec) Rz(0) (MODEL.md §2)
                const float sd = (float)std::sin(dm), cdm = (float)std::cos(dm), sra = (float)std::sin(ra), cra = (float)std::cos(ra);
                const float xb[3] = {cra * sd, sra * sd, -cdm}, yb[3] = {-sra, cra, 0.0f};
                uint8_t* kt = key_tab.data() + (size_t)(bt[b].x + rc) * KEY_ROLLS * KEY_LEVELS;
                uint8_t* kd = key_disc.data() + (size_t)(bt[b].x + rc) * KEY_LEVELS;
                int lev = 0, n_in[KEY_ROLLS] = {0}, n_disc = 0;
                auto close_levels = [&](float m) {   // stars come in magnitude order: close every level this star is fainter than
                    while (lev < KEY_LEVELS && m > KEY_M0 + (lev + 1) / 16.0f) {
                        for (int k = 0; k < KEY_ROLLS; ++k) kt[(size_t)k * KEY_LEVELS + lev] = (uint8_t)std::min(n_in[k], 255);
                        kd[lev] = (uint8_t)std::min(n_disc, 255);
                        ++lev;
                    }
                };
                for (int i = 0; i < S; ++i) {
                    if (dots[i] < thresh) continue;
                    close_levels(smag[i]);
                    list.push_back((uint16_t)order[i]);
                    if (dots[i] >= key_thresh) ++n_disc;
                    if (dots[i] > 0.05f) {
                        // gnomonic coordinates (tangent of the field angles) on the roll-0 axes, rotated into each bin
                        const float inv = 1.0f / dots[i];
                        const float x0 = (xb[0] * sx[i] + xb[1] * sy[i] + xb[2] * sz[i]) * inv, y0 = (yb[0] * sx[i] + yb[1] * sy[i]) * inv;
                        for (int k = 0; k < KEY_ROLLS; ++k) {
                            const float xr = x0 * roll_c[k] + y0 * roll_s[k], yr = y0 * roll_c[k] - x0 * roll_s[k];
                            if (std::fabs(xr) <= tan_u && std::fabs(yr) <= tan_v) ++n_in[k];
                        }
                    }
                }
                close_levels(INFINITY);
            }
            list.push_back((uint16_t)S);                                    // terminator: the +inf dummy star
            while ((list.size() - start) % PACK_N) list.push_back((uint16_t)S);
            if (list.size() > 0xfffffff0ull) return fail(STMPM_E_ARG, "candidate lists exceed 2^32 entries: field of view too wide");
            off[bt[b].x + rc + 1] = (uint32_t)list.size();
        }
    }
    for (int k = 0; k < 2 * PACK_N; ++k) list.push_back((uint16_t)S);             // slack for the unconditional load / prefetch ahead
    cudaFree(h->lk_band_tab); cudaFree(h->lk_off); cudaFree(h->lk_list);
    h->lk_band_tab = nullptr; h->lk_off = nullptr; h->lk_list = nullptr;
    CK(cudaMalloc(&h->lk_band_tab, sizeof(int2) * nb));
    CK(cudaMalloc(&h->lk_off, sizeof(uint32_t) * off.size()));
    CK(cudaMalloc(&h->lk_list, sizeof(uint16_t) * list.size()));
    CK(cudaMemcpy(h->lk_band_tab, bt.data(), sizeof(int2) * nb, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->lk_off, off.data(), sizeof(uint32_t) * off.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->lk_list, list.data(), sizeof(uint16_t) * list.size(), cudaMemcpyHostToDevice));
    h->lk_bands = nb;
    h->lk_cone = (float)cone;
    cudaFree(h->lk_key); h->lk_key = nullptr;
    cudaFree(h->lk_key_disc); h->lk_key_disc = nullptr;
    {   // built cell-major (one pass over each cell's stars in magnitude order), stored level-major (see KArgs::lk_key)
        std::vector<uint8_t> t_tab(key_tab.size()), t_disc(key_disc.size());
        for (long long c = 0; c < n_cells; ++c)
            for (int lev = 0; lev < KEY_LEVELS; ++lev) {
                for (int k = 0; k < KEY_ROLLS; ++k)
                    t_tab[((size_t)lev * n_cells + c) * KEY_ROLLS + k] = key_tab[((size_t)c * KEY_ROLLS + k) * KEY_LEVELS + lev];
                t_disc[(size_t)lev * n_cells + c] = key_disc[(size_t)c * KEY_LEVELS + lev];
            }
        CK(cudaMalloc(&h->lk_key, t_tab.size()));
        CK(cudaMemcpy(h->lk_key, t_tab.data(), t_tab.size(), cudaMemcpyHostToDevice));
        CK(cudaMalloc(&h->lk_key_disc, t_disc.size()));
        CK(cudaMemcpy(h->lk_key_disc, t_disc.data(), t_disc.size(), cudaMemcpyHostToDevice));
    }
    h->lk_cells = (int)n_cells;
    return 0;
}

extern "C" const char* stmpm_last_error(void) { return g_err.c_str(); }
extern "C" int stmpm_version(void) { return 100; }

extern "C" int stmpm_create(stmpm_t** out, int device) {
    if (!out) return fail(STMPM_E_ARG, "stmpm_create: null out pointer");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(STMPM_E_NOGPU, std::string("stmpm_create: no CUDA device (") + cudaGetErrorString(e) +
                                       "); this library has no CPU fallback");
    if (device < 0 || device >= n) return fail(STMPM_E_ARG, "stmpm_create: bad device index");
    DeviceGuard guard(device);
    stmpm_handle* h = new stmpm_handle();
    h->device = device;
    const int rc = [&]() -> int {
        CK(cudaGetDeviceProperties(&h->prop, device));
        for (int i = 0; i < stmpm_handle::NBUF; ++i) CK(cudaStreamCreateWithFlags(&h->streams[i], cudaStreamNonBlocking));
        MathTables mt;
        stmpm_host_tables(&mt);
        CK(cudaMalloc(&h->tables, sizeof(MathTables)));
        CK(cudaMemcpy(h->tables, &mt, sizeof(MathTables), cudaMemcpyHostToDevice));
        CK(cudaHostAlloc(&h->status_host, 2 * sizeof(int), cudaHostAllocMapped));
        h->status_host[0] = h->status_host[1] = 0;
        CK(cudaHostGetDevicePointer(&h->status_dev, h->status_host, 0));
        // the handle's own stream-ordered pool: blocks freed in stream order are kept for the next launch, not returned to the OS
        cudaMemPoolProps pp{};
        pp.allocType = cudaMemAllocationTypePinned;
        pp.handleTypes = cudaMemHandleTypeNone;
        pp.location.type = cudaMemLocationTypeDevice;
        pp.location.id = device;
        CK(cudaMemPoolCreate(&h->pool, &pp));
        unsigned long long keep = ~0ull;
        CK(cudaMemPoolSetAttribute(h->pool, cudaMemPoolAttrReleaseThreshold, &keep));
        return 0;
    }();
    if (rc) { stmpm_destroy(h); return rc; }          // nothing leaks on a failed create
    *out = h;
    return 0;
}

static void free_stage(stmpm_handle* h) {
    for (int i = 0; i < stmpm_handle::NBUF; ++i) {
        cudaFree(h->stage_cols[i]); cudaFree(h->stage_err[i]); cudaFree(h->stage_ndet[i]); cudaFree(h->stage_mag[i]);
        h->stage_cols[i] = nullptr; h->stage_err[i] = nullptr; h->stage_ndet[i] = nullptr; h->stage_mag[i] = nullptr;
    }
    h->stage_cap = 0;
}

extern "C" int stmpm_destroy(stmpm_t* h) {
    if (!h) return 0;
    DeviceGuard guard(h->device);
    cudaFree(h->cat32); cudaFree(h->cat64); cudaFree(h->cell_start); cudaFree(h->band_tab);
    cudaFree(h->stage_stats); cudaFree(h->counters);
    cudaFree(h->lk_band_tab); cudaFree(h->lk_off); cudaFree(h->lk_list); cudaFree(h->tables);
    cudaFree(h->lk_key); cudaFree(h->lk_key_disc);
    free_stage(h);
    for (int i = 0; i < stmpm_handle::NBUF; ++i) if (h->streams[i]) cudaStreamDestroy(h->streams[i]);
    if (h->pool) { cudaDeviceSynchronize(); cudaMemPoolDestroy(h->pool); }
    if (h->status_host) cudaFreeHost(h->status_host);
    delete h;
    return 0;
}

extern "C" int stmpm_device_info(stmpm_t* h, int32_t* sm_count, int64_t* smem_per_block, int32_t* sm_khz,
                                 int32_t* mem_khz) {
    if (!h) return fail(STMPM_E_ARG, "null handle");
    if (sm_count) *sm_count = h->prop.multiProcessorCount;
    if (smem_per_block) *smem_per_block = (int64_t)h->prop.sharedMemPerBlockOptin;
    int v = 0;
    if (sm_khz) { cudaDeviceGetAttribute(&v, cudaDevAttrClockRate, h->device); *sm_khz = v; }
    if (mem_khz) { cudaDeviceGetAttribute(&v, cudaDevAttrMemoryClockRate, h->device); *mem_khz = v; }
    return 0;
}

extern "C" int stmpm_set_catalog(stmpm_t* h, const double* xyz, const float* mag, const int32_t* star_id,
                                 int32_t n_stars, int32_t n_bands, const int32_t* n_ra, const int32_t* cell_start) {
    if (!h || !xyz || !mag || !star_id || !n_ra || !cell_start) return fail(STMPM_E_ARG, "stmpm_set_catalog: null");
    if (n_stars <= 0 || n_stars > 65534)
        return fail(STMPM_E_ARG, "stmpm_set_catalog: n_stars must be in [1, 65534] (16-bit candidate indices + one terminator)");
    if (n_bands <= 0 || n_bands > MAX_BANDS) return fail(STMPM_E_ARG, "stmpm_set_catalog: n_bands must be in [1, 256]");
    // the whole fp32 catalog is staged into shared memory by every CTA: refuse here, with the true limit, not at the first launch
    if (smem_total(n_stars) > (size_t)h->prop.sharedMemPerBlockOptin) {
        const long long cap = ((long long)h->prop.sharedMemPerBlockOptin - (long long)OFF_CAT) / (long long)CAT_ENTRY_BYTES - 1;
        return fail(STMPM_E_ARG, "stmpm_set_catalog: " + std::to_string(n_stars) + " stars do not fit the shared-memory catalog (limit " +
                                     std::to_string(cap) + " stars on this device)");
    }
    DeviceGuard guard(h->device);
    long long n_cells = 0;
    std::vector<int2> bt(n_bands);
    for (int b = 0; b < n_bands; ++b) {
        if (n_ra[b] <= 0) return fail(STMPM_E_ARG, "stmpm_set_catalog: n_ra must be positive");
        bt[b] = make_int2((int)n_cells, n_ra[b]);
        n_cells += n_ra[b];
    }
    if (cell_start[0] != 0 || cell_start[n_cells] != n_stars)
        return fail(STMPM_E_ARG, "stmpm_set_catalog: cell_start must run from 0 to n_stars");
    const int n_pad = (int)((n_cells + 1 + 7) / 8 * 8);
    std::vector<uint16_t> cs(n_pad, (uint16_t)n_stars);
    for (long long c = 0; c <= n_cells; ++c) {
        if (c && cell_start[c] < cell_start[c - 1]) return fail(STMPM_E_ARG, "stmpm_set_catalog: cell_start not sorted");
        cs[c] = (uint16_t)cell_start[c];
    }
    std::vector<float4> c32(n_stars + 1);
    c32[n_stars] = make_float4(0.0f, 0.0f, 1.0f, INFINITY);                // dummy list terminator: never passes a magnitude test
    std::vector<StarRec> c64(n_stars);
    for (int i = 0; i < n_stars; ++i) {
        c32[i] = make_float4((float)xyz[3 * i], (float)xyz[3 * i + 1], (float)xyz[3 * i + 2], mag[i]);
        c64[i].x = xyz[3 * i]; c64[i].y = xyz[3 * i + 1]; c64[i].z = xyz[3 * i + 2];
        c64[i].star_id = (uint32_t)star_id[i]; c64[i].mag = mag[i];
    }
    cudaFree(h->cat32); cudaFree(h->cat64); cudaFree(h->cell_start); cudaFree(h->band_tab);
    h->cat32 = nullptr; h->cat64 = nullptr; h->cell_start = nullptr; h->band_tab = nullptr;
    CK(cudaMalloc(&h->cat32, sizeof(float4) * (n_stars + 1)));
    CK(cudaMalloc(&h->cat64, sizeof(StarRec) * n_stars));
    CK(cudaMalloc(&h->cell_start, sizeof(uint16_t) * n_pad));
    CK(cudaMalloc(&h->band_tab, sizeof(int2) * n_bands));
    CK(cudaMemcpy(h->cat32, c32.data(), sizeof(float4) * (n_stars + 1), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->cat64, c64.data(), sizeof(StarRec) * n_stars, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->cell_start, cs.data(), sizeof(uint16_t) * n_pad, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->band_tab, bt.data(), sizeof(int2) * n_bands, cudaMemcpyHostToDevice));
    h->n_stars = n_stars; h->n_cells_pad = n_pad; h->n_bands = n_bands;

    h->h_xyzf.resize((size_t)3 * n_stars);
    h->h_mag.assign(mag, mag + n_stars);
    for (int i = 0; i < 3 * n_stars; ++i) h->h_xyzf[i] = (float)xyz[i];
    h->have_catalog = true;
    return rebuild_lists(h);
}

extern "C" int stmpm_set_sensor(stmpm_t* h, const stmpm_sensor_cfg* cfg) {
    if (!h || !cfg) return fail(STMPM_E_ARG, "stmpm_set_sensor: null");
    if (!(cfg->f_ideal > 0) || !(cfg->half_u > 0) || !(cfg->half_v > 0) || cfg->n_min < 2)
        return fail(STMPM_E_ARG, "stmpm_set_sensor: f_ideal, half_u, half_v must be > 0 and n_min >= 2");
    if (cfg->hist_log2_sub < 4 || cfg->hist_log2_sub > 8)
        return fail(STMPM_E_ARG, "stmpm_set_sensor: hist_log2_sub must be in [4, 8]");
    if (!(cfg->scan_pad_px >= 0)) return fail(STMPM_E_ARG, "stmpm_set_sensor: scan_pad_px must be >= 0 (0 = default 32)");
    h->cfg = *cfg;
    h->n_bins = STMPM_HIST_OCTAVES << cfg->hist_log2_sub;
    h->have_sensor = true;
    if (h->status_host) h->status_host[1] = 0;          // a new sensor: the split pipeline gets another chance
    return rebuild_lists(h);
}

extern "C" int stmpm_set_pipeline(stmpm_t* h, int32_t pipeline) {
    if (!h) return fail(STMPM_E_ARG, "stmpm_set_pipeline: null handle");
    if (pipeline < -1 || pipeline > 1) return fail(STMPM_E_ARG, "stmpm_set_pipeline: -1 (auto), 0 (single kernel) or 1 (split pipeline)");
    h->tune_pipeline = pipeline;
    h->status_host[1] = 0;
    return 0;
}

extern "C" int stmpm_set_tuning(stmpm_t* h, int32_t window, double key_radius_deg, int32_t stage_columns, int32_t key_table) {
    if (!h) return fail(STMPM_E_ARG, "stmpm_set_tuning: null handle");
    if (key_table < -1 || key_table > 1) return fail(STMPM_E_ARG, "stmpm_set_tuning: key_table must be -1 (per mode), 0 (disc) or 1 (footprint)");
    h->tune_stage = stage_columns != 0;
    h->tune_key_table = key_table;
    if (window < 0 || window > WIN_MAX || window % 32) return fail(STMPM_E_ARG, "stmpm_set_tuning: window must be 0 (default) or a multiple of 32 up to 2048");
    const bool rebuild = key_radius_deg != h->tune_key_radius_deg;
    h->tune_window = window;
    h->tune_key_radius_deg = key_radius_deg;
    if (!rebuild) return 0;
    DeviceGuard guard(h->device);
    CK(cudaDeviceSynchronize());                       // the work-key table is about to be replaced
    return rebuild_lists(h);
}

extern "C" int stmpm_set_host_pipeline(stmpm_t* h, int64_t chunk_trials) {
    if (!h) return fail(STMPM_E_ARG, "stmpm_set_host_pipeline: null handle");
    if (chunk_trials != 0 && (chunk_trials < 1024 || chunk_trials > (1LL << 28)))
        return fail(STMPM_E_ARG, "stmpm_set_host_pipeline: chunk_trials must be 0 (default 2^22) or in [1024, 2^28]");
    DeviceGuard guard(h->device);
    for (int i = 0; i < stmpm_handle::NBUF; ++i) CK(cudaStreamSynchronize(h->streams[i]));
    h->host_chunk = chunk_trials ? chunk_trials : HOST_CHUNK;
    free_stage(h);                                       // re-allocated at the new size by the next host call
    return 0;
}

extern "C" int stmpm_stats_len(int32_t hist_log2_sub, int64_t* n) {
    if (!n) return fail(STMPM_E_ARG, "stmpm_stats_len: null");
    if (hist_log2_sub < 4 || hist_log2_sub > 8) return fail(STMPM_E_ARG, "hist_log2_sub must be in [4, 8]");
    *n = STMPM_STATS_HEAD + ((int64_t)STMPM_HIST_OCTAVES << hist_log2_sub);
    return 0;
}

// percentile from the log-linear histogram: bins are linear inside each octave
static double hist_percentile(const double* hist, int n_bins, int log2_sub, double under, double n_ok, double q) {
    const double target = q * n_ok;
    double cum = under;
    if (target <= cum) return std::ldexp(1.0, STMPM_HIST_MIN_EXP);
    for (int b = 0; b < n_bins; ++b) {
        const double c = hist[b];
        if (c > 0 && cum + c >= target) {
            const int oct = b >> log2_sub, sub = b & ((1 << log2_sub) - 1);
            const double base = std::ldexp(1.0, STMPM_HIST_MIN_EXP + oct);
            const double w = base / (double)(1 << log2_sub);
            return base + w * ((double)sub + (target - cum) / c);
        }
        cum += c;
    }
    return std::ldexp(1.0, STMPM_HIST_MIN_EXP + STMPM_HIST_OCTAVES);
}

extern "C" int stmpm_stats_finalize(int32_t hist_log2_sub, const double* s, stmpm_summary* o) {
    if (!s || !o) return fail(STMPM_E_ARG, "stmpm_stats_finalize: null");
    if (hist_log2_sub < 4 || hist_log2_sub > 8) return fail(STMPM_E_ARG, "hist_log2_sub must be in [4, 8]");
    const int n_bins = STMPM_HIST_OCTAVES << hist_log2_sub;
    std::memset(o, 0, sizeof(*o));
    const double n = s[0], nf = s[1];
    o->n_ok = n; o->n_fail = nf;
    o->fail_rate = (n + nf) > 0 ? nf / (n + nf) : 0.0;          // driver.py:321-322
    o->mean_ndet = (n + nf) > 0 ? s[4] / (n + nf) : 0.0;
    o->x_max = s[5];
    if (n > 0) {
        o->mean = s[2] / n;
        o->rms = std::sqrt(s[3] / n);                            // toolplot.py:82
        const double var = n > 1 ? std::max(0.0, (s[3] - s[2] * s[2] / n) / (n - 1.0)) : 0.0;   // pandas .std(): ddof=1
        o->std = std::sqrt(var);
        o->sigma_halfnormal = o->std / std::sqrt(1.0 - 2.0 / M_PI);      // driver.py:290,324
        o->mean_halfnormal = o->sigma_halfnormal * std::sqrt(2.0 / M_PI); // driver.py:291,323
        const double* hist = s + STMPM_STATS_HEAD;
        o->p50 = hist_percentile(hist, n_bins, hist_log2_sub, s[6], n, 0.50);
        o->p90 = hist_percentile(hist, n_bins, hist_log2_sub, s[6], n, 0.90);
        o->p99 = hist_percentile(hist, n_bins, hist_log2_sub, s[6], n, 0.99);
        o->p999 = hist_percentile(hist, n_bins, hist_log2_sub, s[6], n, 0.999);
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------ launch helper
template <int MODE, int PREC, int INSTR = 0>
static int launch_trials(stmpm_handle* h, KArgs& a, cudaStream_t st) {
    const size_t smem_bytes = smem_total(a.n_stars);
    if (smem_bytes > (size_t)h->prop.sharedMemPerBlockOptin)
        return fail(STMPM_E_ARG, "catalog too large for shared-memory staging (" + std::to_string(smem_bytes) + " B)");
    if (a.n_bands > MAX_BANDS || a.lk_bands != LK_BANDS) return fail(STMPM_E_ARG, "sky grid has too many declination bands");
    auto kern = trials_kernel<MODE, PREC, INSTR>;
    // the attribute belongs to the FUNCTION (per device), not to a handle: keep a per-instantiation, per-device maximum so a
    // second engine with a smaller catalog can never lower it under a live one
    static std::atomic<size_t> attr_set[64];            // zero-initialised
    static std::mutex attr_mutex;                       // taken only to RAISE the attribute (first launch per catalog size)
    std::atomic<size_t>& cached = attr_set[h->device & 63];
    if (cached.load(std::memory_order_acquire) < smem_bytes) {
        std::lock_guard<std::mutex> attr_lock(attr_mutex);
        if (cached.load(std::memory_order_relaxed) < smem_bytes) {
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
            // leave the rest of the unified array to L1: register spills and the permuted input gathers live there
            const int pct = (int)std::min<size_t>(100, (smem_bytes + 1024) * 100 / (size_t)h->prop.sharedMemPerMultiprocessor + 1);
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
            cached.store(smem_bytes, std::memory_order_release);
        }
    }
    const int n_sm = h->prop.multiProcessorCount;
    // work units: (segment, part).  A unit is consumed in CTA-wide sort windows of `win` trials; small launches shrink
    // the window so that more CTAs have work.
    const bool footprint_key = h->tune_key_table < 0 ? (MODE == 1) : (h->tune_key_table == 1);
    a.lk_key = footprint_key ? h->lk_key : h->lk_key_disc;
    a.key_rolls = footprint_key ? KEY_ROLLS : 1;
    const long long tps = a.trials_per_segment;
    const long long total = tps * a.n_segments;
    // sorted parameter staging (pre-sampled mode): worth its scratch from a few windows per CTA on
    const bool stage = MODE == 0 && h->tune_stage && total >= 4LL * WIN_STAGED * n_sm;
    long long win = MODE == 0 ? (stage ? WIN_STAGED : WIN_PRESAMPLED) : WIN_RNG;
    if (h->tune_window > 0) win = std::max(32, std::min(WIN_MAX, h->tune_window / 32 * 32));   // stmpm_set_tuning
    const long long per_cta = (total + n_sm - 1) / n_sm;
    if (per_cta < win) win = std::max(32LL, (per_cta + 31) / 32 * 32);
    a.win = (int)win;
    const long long rounds_per_seg = (tps + win - 1) / win;          // windows per segment
    long long parts = 1;
    if (a.n_segments < 8LL * n_sm) parts = std::max(1LL, std::min(rounds_per_seg, (8LL * n_sm + a.n_segments - 1) / a.n_segments));   // 8 units per CTA, strided: inputs with a trend still balance
    const long long rounds_per_part = (rounds_per_seg + parts - 1) / parts;
    parts = (rounds_per_seg + rounds_per_part - 1) / rounds_per_part;
    a.parts_per_segment = (int)parts;
    a.part_len = rounds_per_part * win;
    a.windows_per_unit = (int)rounds_per_part;
    const long long n_units = a.n_segments * parts;
    const int grid = (int)std::min<long long>(n_units, n_sm);
    if (grid <= 0) return 0;
    if (rounds_per_part > 0x7fffffffLL || ((n_units + grid - 1) / grid) * rounds_per_part * (win / 32) > 0x7fffffffLL)
        return fail(STMPM_E_ARG, "launch too large: more than 2^31 trial groups per CTA (split the call)");
    void* scratch = nullptr;
    a.stage = nullptr;
    if (stage) {   // [grid][2 buffers][13 columns][WIN_MAX] doubles, allocated and freed in stream order: concurrent launches never share it
        CK(cudaMallocFromPoolAsync(&scratch, (size_t)grid * 2 * STMPM_NCOLS * WIN_MAX * sizeof(double), h->pool, st));
        a.stage = static_cast<double*>(scratch);
    }
    kern<<<grid, TPB, smem_bytes, st>>>(a);
    const cudaError_t le = cudaGetLastError();
    if (scratch) cudaFreeAsync(scratch, st);
    if (le != cudaSuccess) return fail((int)le, std::string("trials_kernel launch: ") + cudaGetErrorString(le));
    ++h->launches;
    return 0;
}

// ------------------------------------------------------------------------------------------------ split pipeline launcher
// Pre-sampled, one segment, no instrumentation: scan kernel -> counting sort by survivor count -> solve kernel, in chunks of
// 2^24 trials (3.8 GB of records from the handle's pool), all in stream order.  The single kernel is enqueued behind every
// chunk with a device-side cancel word: it runs only when the scan could not hand over more than 1/64 of the chunk (star-
// rich or wide-cone workloads), and the host then stops choosing the split for this catalog / sensor.
template <int PREC>
static int launch_split(stmpm_handle* h, const KArgs& a0, cudaStream_t st) {
    const size_t smem_bytes = smem_total(a0.n_stars);
    if (smem_bytes > (size_t)h->prop.sharedMemPerBlockOptin) return fail(STMPM_E_ARG, "catalog too large for shared-memory staging");
    static std::atomic<size_t> attr_set[64];
    static std::mutex attr_mutex;
    std::atomic<size_t>& cached = attr_set[h->device & 63];
    if (cached.load(std::memory_order_acquire) < smem_bytes) {
        std::lock_guard<std::mutex> lock(attr_mutex);
        if (cached.load(std::memory_order_relaxed) < smem_bytes) {
            CK(cudaFuncSetAttribute(split_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
            cached.store(smem_bytes, std::memory_order_release);
        }
    }
    const int n_sm = h->prop.multiProcessorCount;
    const long long n = a0.trials_per_segment;
    const long long chunk = std::min<long long>(n, 1LL << 24);
    struct PoolBuf {
        void* p = nullptr; cudaStream_t st;
        ~PoolBuf() { if (p) cudaFreeAsync(p, st); }
    } b_rec{nullptr, st}, b_perm{nullptr, st}, b_ctl{nullptr, st};
    const long long n_chunks = (n + chunk - 1) / chunk;
    CK(cudaMallocFromPoolAsync(&b_rec.p, sizeof(TrialRec) * (size_t)chunk, h->pool, st));
    CK(cudaMallocFromPoolAsync(&b_perm.p, sizeof(uint32_t) * (size_t)chunk, h->pool, st));
    CK(cudaMallocFromPoolAsync(&b_ctl.p, sizeof(SplitCtl) * (size_t)n_chunks, h->pool, st));
    CK(cudaMemsetAsync(b_ctl.p, 0, sizeof(SplitCtl) * (size_t)n_chunks, st));
    TrialRec* rec = static_cast<TrialRec*>(b_rec.p);
    uint32_t* perm = static_cast<uint32_t*>(b_perm.p);
    for (long long ci = 0; ci < n_chunks; ++ci) {
        const long long off = ci * chunk, m = std::min<long long>(chunk, n - off);
        SplitCtl* ctl = static_cast<SplitCtl*>(b_ctl.p) + ci;
        KArgs k = a0;
        for (int c = 0; c < STMPM_NCOLS; ++c) k.cols[c] = a0.cols[c] + off;
        if (k.err) k.err += off;
        if (k.ndet) k.ndet += off;
        if (k.ndet8) k.ndet8 += off;
        k.trial0 = a0.trial0 + off;
        k.trials_per_segment = m;
        const int grid_scan = (int)std::min<long long>(n_sm, (m + TPB_SCAN - 1) / TPB_SCAN);
        split_scan_kernel<<<grid_scan, TPB_SCAN, smem_bytes, st>>>(k, rec, perm, m, ctl);
        split_decide_kernel<<<1, 1, 0, st>>>(ctl, m, h->status_dev + 1);
        const int grid_solve = (int)std::min<long long>(n_sm, (m + TPB - 1) / TPB);
        split_solve_kernel<PREC><<<grid_solve, TPB, 0, st>>>(k, rec, perm, m, ctl);
        CK(cudaGetLastError());
        h->launches += 3;
        KArgs kl = k;                                    // the single kernel, behind a cancel word: runs only on a fall-back
        kl.cancel = &ctl->skip_single;
        if (int rc = launch_trials<0, PREC, 0>(h, kl, st)) return rc;
    }
    return 0;
}

static bool want_split(const stmpm_handle* h, const KArgs& a) {
    // measured slower than the single kernel: opt-in only
    return h->tune_pipeline == 1 && a.n_segments == 1 && a.maxmag == nullptr;
}

static int fill_common(stmpm_handle* h, KArgs& a, int precision) {
    if (!h) return fail(STMPM_E_ARG, "null handle");
    if (!h->have_catalog || !h->have_sensor)
        return fail(STMPM_E_STATE, "call stmpm_set_catalog and stmpm_set_sensor before running trials");
    if (precision != 64 && precision != 32) return fail(STMPM_E_ARG, "precision must be 64 or 32");
    std::memset(&a, 0, sizeof(a));
    a.cat32 = h->cat32; a.cat64 = h->cat64; a.cell_start = h->cell_start; a.band_tab = h->band_tab;
    a.n_stars = h->n_stars; a.n_cells_pad = h->n_cells_pad; a.n_bands = h->n_bands;
    a.inv_band_h = (float)((double)h->n_bands / M_PI);
    a.lk_band_tab = h->lk_band_tab; a.lk_off = h->lk_off; a.lk_list = h->lk_list; a.lk_bands = h->lk_bands;
    a.lk_inv_band_h = (float)((double)h->lk_bands / M_PI); a.lk_cone = h->lk_cone; a.tables = h->tables;
    a.lk_cone_tan = (float)(std::tan(std::min((double)h->lk_cone - 0.002, 1.55)) * (1.0 - 1e-6));
    a.lk_key = h->lk_key; a.lk_cells = h->lk_cells; a.key_rolls = KEY_ROLLS; a.key_roll_inv = h->key_roll_inv;   // launch_trials picks the table per mode
    a.f_ideal = h->cfg.f_ideal; a.f_ideal2 = h->cfg.f_ideal * h->cfg.f_ideal; a.half_u = h->cfg.half_u; a.half_v = h->cfg.half_v; a.n_min = h->cfg.n_min;
    a.hist_shift = 52 - h->cfg.hist_log2_sub; a.n_bins = h->n_bins;
    a.stats_len = STMPM_STATS_HEAD + h->n_bins;
    a.status = h->status_dev;
    return 0;
}

// after a synchronisation: did a kernel of this handle hit its scheduler watchdog?
static int check_status(stmpm_handle* h) {
    const int code = *reinterpret_cast<volatile int*>(h->status_host);
    if (code == 0) return 0;
    *h->status_host = 0;
    return fail(code, "trial kernel scheduler watchdog: a window / group wait exceeded 20 s (results of this call are incomplete)");
}

extern "C" int stmpm_check(stmpm_t* h) {
    if (!h) return fail(STMPM_E_ARG, "stmpm_check: null handle");
    return check_status(h);
}

extern "C" int stmpm_trials_presampled(stmpm_t* h, int precision, int64_t n, uint64_t seed, int64_t trial0,
                                       const double* const* cols, double* err, int32_t* ndet, double* stats,
                                       void* stream) {
    return stmpm_trials_presampled_ex(h, precision, n, seed, trial0, cols, err, ndet, nullptr, stats, stream);
}

extern "C" int stmpm_trials_presampled_ex(stmpm_t* h, int precision, int64_t n, uint64_t seed, int64_t trial0,
                                          const double* const* cols, double* err, int32_t* ndet, float* maxmag,
                                          double* stats, void* stream) {
    KArgs a;
    if (int rc = fill_common(h, a, precision)) return rc;
    if (n < 0 || !cols) return fail(STMPM_E_ARG, "stmpm_trials_presampled: bad n / cols");
    if (n == 0) return 0;
    DeviceGuard guard(h->device);
    for (int i = 0; i < STMPM_NCOLS; ++i) {
        if (!cols[i]) return fail(STMPM_E_ARG, "stmpm_trials_presampled: null column pointer");
        a.cols[i] = cols[i];
    }
    a.n_segments = 1; a.trials_per_segment = n; a.trial0 = trial0;
    a.k0 = (uint32_t)seed; a.k1 = (uint32_t)(seed >> 32);
    a.err = err; a.ndet = ndet; a.maxmag = maxmag; a.stats = stats;
    if (maxmag) return precision == 64 ? launch_trials<0, 64, 2>(h, a, (cudaStream_t)stream)
                                       : launch_trials<0, 32, 2>(h, a, (cudaStream_t)stream);
    if (want_split(h, a)) return precision == 64 ? launch_split<64>(h, a, (cudaStream_t)stream)
                                                 : launch_split<32>(h, a, (cudaStream_t)stream);
    return precision == 64 ? launch_trials<0, 64>(h, a, (cudaStream_t)stream)
                           : launch_trials<0, 32>(h, a, (cudaStream_t)stream);
}

static void to_dev_dist(const stmpm_dist& s, DistDev& d) {
    for (int i = 0; i < STMPM_NNORMAL; ++i) { d.mean[i] = s.mean[i]; d.sigma[i] = s.sigma[i]; }
    d.thermal_coef = s.thermal_coef;
}

extern "C" int stmpm_trials_rng(stmpm_t* h, int precision, int64_t n_segments, int64_t tps, uint64_t seed,
                                int64_t trial0, const stmpm_dist* dists, double* err, int32_t* ndet, double* stats,
                                void* stream) {
    KArgs a;
    if (int rc = fill_common(h, a, precision)) return rc;
    if (n_segments < 0 || tps < 0 || !dists) return fail(STMPM_E_ARG, "stmpm_trials_rng: bad arguments");
    if (n_segments == 0 || tps == 0) return 0;
    DeviceGuard guard(h->device);
    static_assert(sizeof(DistDev) == sizeof(stmpm_dist), "dist layout");
    DistDev* d_dists = nullptr;
    if (n_segments == 1) {
        to_dev_dist(dists[0], a.dist0);
    } else {
        // every call owns its distribution array: allocated, filled and freed in stream order on the caller's stream, so
        // concurrent multi-segment calls on different streams never share storage (the host array is pageable memory:
        // cudaMemcpyAsync has staged it before it returns, the caller may reuse it at once)
        CK(cudaMallocFromPoolAsync(&d_dists, sizeof(DistDev) * n_segments, h->pool, (cudaStream_t)stream));
        cudaError_t e = cudaMemcpyAsync(d_dists, dists, sizeof(DistDev) * n_segments, cudaMemcpyHostToDevice, (cudaStream_t)stream);
        if (e != cudaSuccess) { cudaFreeAsync(d_dists, (cudaStream_t)stream); return fail((int)e, std::string("dists upload: ") + cudaGetErrorString(e)); }
        a.dists = d_dists;
    }
    a.n_segments = n_segments; a.trials_per_segment = tps; a.trial0 = trial0;
    a.k0 = (uint32_t)seed; a.k1 = (uint32_t)(seed >> 32);
    a.err = err; a.ndet = ndet; a.stats = stats;
    const int rc = precision == 64 ? launch_trials<1, 64>(h, a, (cudaStream_t)stream)
                                   : launch_trials<1, 32>(h, a, (cudaStream_t)stream);
    if (d_dists) cudaFreeAsync(d_dists, (cudaStream_t)stream);
    return rc;
}

extern "C" int stmpm_fill_presampled(stmpm_t* h, int64_t n, uint64_t seed, int64_t trial0, const stmpm_dist* dist,
                                     double* const* cols, void* stream) {
    if (!h || !dist || !cols || n < 0) return fail(STMPM_E_ARG, "stmpm_fill_presampled: bad arguments");
    if (n == 0) return 0;
    DeviceGuard guard(h->device);
    for (int i = 0; i < STMPM_NCOLS; ++i) if (!cols[i]) return fail(STMPM_E_ARG, "stmpm_fill_presampled: null column");
    DistDev d;
    to_dev_dist(*dist, d);
    const int grid = (int)std::min<long long>((n + 255) / 256, (long long)h->prop.multiProcessorCount * 8);
    fill_presampled_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        n, trial0, (uint32_t)seed, (uint32_t)(seed >> 32), d, h->tables, cols[0], cols[1], cols[2], cols[3], cols[4], cols[5],
        cols[6], cols[7], cols[8], cols[9], cols[10], cols[11], cols[12]);
    CK(cudaGetLastError());
    ++h->launches;
    return 0;
}

// ------------------------------------------------------------------------------------------------ host pipelines
static int ensure_stage(stmpm_handle* h, long long chunk, bool need_cols) {
    if (h->stage_cap >= chunk && (!need_cols || h->stage_cols[0])) return 0;
    free_stage(h);
    for (int i = 0; i < stmpm_handle::NBUF; ++i) {
        if (need_cols) CK(cudaMalloc(&h->stage_cols[i], sizeof(double) * STMPM_NCOLS * chunk));
        CK(cudaMalloc(&h->stage_err[i], sizeof(double) * chunk));
        CK(cudaMalloc(&h->stage_ndet[i], sizeof(int) * chunk));
        CK(cudaMalloc(&h->stage_mag[i], sizeof(float) * chunk));
    }
    h->stage_cap = chunk;
    return 0;
}
static int ensure_stats(stmpm_handle* h, long long n_doubles) {
    if (h->stage_stats && h->stage_stats_cap >= n_doubles) return 0;
    cudaFree(h->stage_stats); h->stage_stats = nullptr; h->stage_stats_cap = 0;
    CK(cudaMalloc(&h->stage_stats, sizeof(double) * n_doubles));
    h->stage_stats_cap = n_doubles;
    return 0;
}


extern "C" int stmpm_run_presampled_host(stmpm_t* h, int precision, int64_t n, uint64_t seed, int64_t trial0,
                                         const double* const* cols, double* err, int32_t* ndet, double* stats) {
    return stmpm_run_presampled_host_ex(h, precision, n, seed, trial0, cols, err, ndet, nullptr, stats);
}

extern "C" int stmpm_run_presampled_host_ex(stmpm_t* h, int precision, int64_t n, uint64_t seed, int64_t trial0,
                                            const double* const* cols, double* err, int32_t* ndet, float* maxmag,
                                            double* stats) {
    KArgs probe;
    if (int rc = fill_common(h, probe, precision)) return rc;
    if (n < 0 || !cols || !err || !ndet) return fail(STMPM_E_ARG, "stmpm_run_presampled_host: bad arguments");
    for (int i = 0; i < STMPM_NCOLS; ++i) if (!cols[i]) return fail(STMPM_E_ARG, "null column pointer");
    if (n == 0) return 0;
    DeviceGuard guard(h->device);
    const long long chunk = std::min<long long>(h->host_chunk, n);
    if (int rc = ensure_stage(h, chunk, true)) return rc;
    const long long slen = probe.stats_len;
    if (stats) {
        if (int rc = ensure_stats(h, slen)) return rc;
        CK(cudaMemsetAsync(h->stage_stats, 0, sizeof(double) * slen, h->streams[0]));
        CK(cudaStreamSynchronize(h->streams[0]));
    }
    int k = 0;
    for (long long off = 0; off < n; off += chunk, ++k) {
        const int b = k % stmpm_handle::NBUF;
        const long long m = std::min<long long>(chunk, n - off);
        cudaStream_t st = h->streams[b];
        const double* dcols[STMPM_NCOLS];
        bool contiguous = (off == 0 && m == n);            // one [13, n] block (what Simulation.run_sim passes): one copy
        for (int c = 1; c < STMPM_NCOLS && contiguous; ++c) contiguous = cols[c] == cols[0] + (long long)c * n;
        if (contiguous) {
            CK(cudaMemcpyAsync(h->stage_cols[b], cols[0], sizeof(double) * STMPM_NCOLS * m, cudaMemcpyHostToDevice, st));
            for (int c = 0; c < STMPM_NCOLS; ++c) dcols[c] = h->stage_cols[b] + (long long)c * m;
        } else {
            for (int c = 0; c < STMPM_NCOLS; ++c) {
                double* dst = h->stage_cols[b] + (long long)c * h->stage_cap;
                CK(cudaMemcpyAsync(dst, cols[c] + off, sizeof(double) * m, cudaMemcpyHostToDevice, st));
                dcols[c] = dst;
            }
        }
        if (int rc = stmpm_trials_presampled_ex(h, precision, m, seed, trial0 + off, dcols, h->stage_err[b],
                                                h->stage_ndet[b], maxmag ? h->stage_mag[b] : nullptr,
                                                stats ? h->stage_stats : nullptr, st)) return rc;
        CK(cudaMemcpyAsync(err + off, h->stage_err[b], sizeof(double) * m, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(ndet + off, h->stage_ndet[b], sizeof(int) * m, cudaMemcpyDeviceToHost, st));
        if (maxmag) CK(cudaMemcpyAsync(maxmag + off, h->stage_mag[b], sizeof(float) * m, cudaMemcpyDeviceToHost, st));
    }
    for (int i = 0; i < stmpm_handle::NBUF; ++i) CK(cudaStreamSynchronize(h->streams[i]));
    if (stats) CK(cudaMemcpy(stats, h->stage_stats, sizeof(double) * slen, cudaMemcpyDeviceToHost));
    return check_status(h);
}

// record format: 12 = (f64 err, i32 n_det), 9 = (f64 err, u8 n_det), 5 = (f32 err, u8 n_det)
static int run_rng_host_impl(stmpm_t* h, int precision, int64_t n_segments, int64_t tps, uint64_t seed, int64_t trial0,
                             const stmpm_dist* dists, void* err, void* ndet, int format, double* stats) {
    KArgs probe;
    if (int rc = fill_common(h, probe, precision)) return rc;
    if (n_segments < 0 || tps < 0 || !dists) return fail(STMPM_E_ARG, "stmpm_run_rng_host: bad arguments");
    if ((err == nullptr) != (ndet == nullptr)) return fail(STMPM_E_ARG, "err and n_det must both be given or both NULL");
    if (n_segments == 0 || tps == 0) return 0;
    DeviceGuard guard(h->device);
    const long long slen = probe.stats_len;
    cudaStream_t s0 = h->streams[0];
    if (stats) {
        if (int rc = ensure_stats(h, slen * n_segments)) return rc;
        CK(cudaMemsetAsync(h->stage_stats, 0, sizeof(double) * slen * n_segments, s0));
    }
    if (!err) {   // stats only: one launch over all segments
        if (int rc = stmpm_trials_rng(h, precision, n_segments, tps, seed, trial0, dists, nullptr, nullptr,
                                      stats ? h->stage_stats : nullptr, s0)) return rc;
    } else {      // records: per segment, chunked + multi-buffered D2H
        const long long chunk = std::min<long long>(h->host_chunk, tps);
        if (int rc = ensure_stage(h, chunk, false)) return rc;
        CK(cudaStreamSynchronize(s0));
        const size_t nd_bytes = format == 12 ? sizeof(int) : 1, err_bytes = format == 5 ? sizeof(float) : sizeof(double);
        int k = 0;
        for (long long s = 0; s < n_segments; ++s) {
            for (long long off = 0; off < tps; off += chunk, ++k) {
                const int b = k % stmpm_handle::NBUF;
                const long long m = std::min<long long>(chunk, tps - off);
                cudaStream_t st = h->streams[b];
                KArgs a;
                if (int rc = fill_common(h, a, precision)) return rc;
                to_dev_dist(dists[s], a.dist0);
                a.n_segments = 1; a.trials_per_segment = m; a.trial0 = trial0 + s * tps + off;
                a.k0 = (uint32_t)seed; a.k1 = (uint32_t)(seed >> 32);
                if (format == 5) a.err32 = reinterpret_cast<float*>(h->stage_err[b]); else a.err = h->stage_err[b];
                if (format != 12) a.ndet8 = reinterpret_cast<uint8_t*>(h->stage_ndet[b]); else a.ndet = h->stage_ndet[b];
                a.stats = stats ? h->stage_stats + s * slen : nullptr;
                const int rc = precision == 64 ? launch_trials<1, 64>(h, a, st) : launch_trials<1, 32>(h, a, st);
                if (rc) return rc;
                CK(cudaMemcpyAsync(static_cast<char*>(err) + (size_t)(s * tps + off) * err_bytes, h->stage_err[b], err_bytes * m,
                                   cudaMemcpyDeviceToHost, st));
                CK(cudaMemcpyAsync(static_cast<char*>(ndet) + (size_t)(s * tps + off) * nd_bytes, h->stage_ndet[b], nd_bytes * m,
                                   cudaMemcpyDeviceToHost, st));
            }
        }
    }
    for (int i = 0; i < stmpm_handle::NBUF; ++i) CK(cudaStreamSynchronize(h->streams[i]));
    if (stats) CK(cudaMemcpy(stats, h->stage_stats, sizeof(double) * slen * n_segments, cudaMemcpyDeviceToHost));
    return check_status(h);
}

extern "C" int stmpm_run_rng_host(stmpm_t* h, int precision, int64_t n_segments, int64_t tps, uint64_t seed,
                                  int64_t trial0, const stmpm_dist* dists, double* err, int32_t* ndet, double* stats) {
    return run_rng_host_impl(h, precision, n_segments, tps, seed, trial0, dists, err, ndet, 12, stats);
}

extern "C" int stmpm_run_rng_host_compact(stmpm_t* h, int precision, int64_t n_segments, int64_t tps, uint64_t seed,
                                          int64_t trial0, const stmpm_dist* dists, double* err, uint8_t* ndet_u8, double* stats) {
    return run_rng_host_impl(h, precision, n_segments, tps, seed, trial0, dists, err, ndet_u8, 9, stats);
}

extern "C" int stmpm_run_rng_host_compact32(stmpm_t* h, int precision, int64_t n_segments, int64_t tps, uint64_t seed,
                                            int64_t trial0, const stmpm_dist* dists, float* err_f32, uint8_t* ndet_u8, double* stats) {
    return run_rng_host_impl(h, precision, n_segments, tps, seed, trial0, dists, err_f32, ndet_u8, 5, stats);
}

// Host-side roofline of the host-buffer entry points: the SAME chunking, streams and staging buffers, copies only — no
// kernel.  h2d_bytes_per_trial / d2h_bytes_per_trial: 104 / 12 (pre-sampled), 0 / 12 or 0 / 9 (native RNG, compact).
extern "C" int stmpm_host_copy_roofline(stmpm_t* h, int64_t n, int32_t h2d_bytes_per_trial, int32_t d2h_bytes_per_trial,
                                        const void* src_host, void* dst_host, double* seconds) {
    if (!h || n <= 0 || !seconds || h2d_bytes_per_trial < 0 || h2d_bytes_per_trial > 8 * STMPM_NCOLS || d2h_bytes_per_trial < 0 ||
        d2h_bytes_per_trial > 12 || (h2d_bytes_per_trial && !src_host) || (d2h_bytes_per_trial && !dst_host))
        return fail(STMPM_E_ARG, "stmpm_host_copy_roofline: bad arguments");
    DeviceGuard guard(h->device);
    const long long chunk = std::min<long long>(h->host_chunk, n);
    if (int rc = ensure_stage(h, chunk, h2d_bytes_per_trial > 0)) return rc;
    for (int i = 0; i < stmpm_handle::NBUF; ++i) CK(cudaStreamSynchronize(h->streams[i]));
    const int rc = [&]() -> int {
        const auto t0 = std::chrono::steady_clock::now();
        int k = 0;
        for (long long off = 0; off < n; off += chunk, ++k) {
            const int b = k % stmpm_handle::NBUF;
            const long long m = std::min<long long>(chunk, n - off);
            cudaStream_t st = h->streams[b];
            if (h2d_bytes_per_trial)
                CK(cudaMemcpyAsync(h->stage_cols[b], static_cast<const char*>(src_host) + (size_t)off * h2d_bytes_per_trial,
                                   (size_t)m * h2d_bytes_per_trial, cudaMemcpyHostToDevice, st));
            if (d2h_bytes_per_trial) {
                const size_t b8 = std::min<size_t>(8, d2h_bytes_per_trial), b4 = d2h_bytes_per_trial - b8;
                CK(cudaMemcpyAsync(static_cast<char*>(dst_host) + (size_t)off * b8, h->stage_err[b], (size_t)m * b8, cudaMemcpyDeviceToHost, st));
                if (b4) CK(cudaMemcpyAsync(static_cast<char*>(dst_host) + (size_t)n * b8 + (size_t)off * b4, h->stage_ndet[b], (size_t)m * b4,
                                           cudaMemcpyDeviceToHost, st));
            }
        }
        for (int i = 0; i < stmpm_handle::NBUF; ++i) CK(cudaStreamSynchronize(h->streams[i]));
        *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        return 0;
    }();
    return rc;
}

// ------------------------------------------------------------------------------------------------ misc
extern "C" int stmpm_fma_peak(stmpm_t* h, int precision, double* tflops) {
    if (!h || !tflops) return fail(STMPM_E_ARG, "stmpm_fma_peak: null");
    if (precision != 64 && precision != 32) return fail(STMPM_E_ARG, "precision must be 64 or 32");
    DeviceGuard guard(h->device);
    void* out = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    double best = 0.0;
    const int rc = [&]() -> int {
        CK(cudaMalloc(&out, 64));
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        const int grid = h->prop.multiProcessorCount * 8, iters = precision == 64 ? 1 << 15 : 1 << 16;
        for (int rep = 0; rep < 4; ++rep) {
            CK(cudaEventRecord(e0, h->streams[0]));
            if (precision == 64) fma_peak_kernel<double><<<grid, 256, 0, h->streams[0]>>>((double*)out, iters, 1.0000001, 1e-9);
            else fma_peak_kernel<float><<<grid, 256, 0, h->streams[0]>>>((float*)out, iters, 1.0001f, 1e-6f);
            CK(cudaEventRecord(e1, h->streams[0]));
            CK(cudaEventSynchronize(e1));
            ++h->launches;
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            const double flops = 2.0 * 8.0 * (double)iters * 256.0 * (double)grid;
            if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
        }
        return 0;
    }();
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    cudaFree(out);
    if (rc) return rc;
    *tflops = best;
    return 0;
}

extern "C" int stmpm_work_counters(stmpm_t* h, int64_t n, uint64_t seed, int64_t trial0, const stmpm_dist* dist,
                                   double* per_trial4) {
    KArgs a;
    if (int rc = fill_common(h, a, 64)) return rc;
    if (n <= 0 || !dist || !per_trial4) return fail(STMPM_E_ARG, "stmpm_work_counters: bad arguments");
    DeviceGuard guard(h->device);
    if (!h->counters) CK(cudaMalloc(&h->counters, 4 * sizeof(unsigned long long)));
    CK(cudaMemsetAsync(h->counters, 0, 4 * sizeof(unsigned long long), h->streams[0]));
    to_dev_dist(*dist, a.dist0);
    a.n_segments = 1; a.trials_per_segment = n; a.trial0 = trial0;
    a.k0 = (uint32_t)seed; a.k1 = (uint32_t)(seed >> 32);
    a.counters = h->counters;
    if (int rc = launch_trials<1, 64, 1>(h, a, h->streams[0])) return rc;
    unsigned long long c[4];
    CK(cudaMemcpyAsync(c, h->counters, sizeof(c), cudaMemcpyDeviceToHost, h->streams[0]));
    CK(cudaStreamSynchronize(h->streams[0]));
    for (int i = 0; i < 4; ++i) per_trial4[i] = (double)c[i] / (double)n;
    return 0;
}

extern "C" int stmpm_debug_box_muller_f32(const uint32_t* wa, const uint32_t* wb, int64_t n, float* z1, float* z2) {
    if (!wa || !wb || !z1 || !z2 || n < 0) return fail(STMPM_E_ARG, "stmpm_debug_box_muller_f32: bad arguments");
    MathTables mt;
    stmpm_host_tables(&mt);
    for (int64_t i = 0; i < n; ++i) box_muller_f32(TabPtr{&mt}, wa[i], wb[i], z1[i], z2[i]);
    return 0;
}

extern "C" int stmpm_debug_box_muller(const uint32_t* wa, const uint32_t* wb, int64_t n, double* z1, double* z2) {
    if (!wa || !wb || !z1 || !z2 || n < 0) return fail(STMPM_E_ARG, "stmpm_debug_box_muller: bad arguments");
    MathTables mt;
    stmpm_host_tables(&mt);
    for (int64_t i = 0; i < n; ++i) box_muller_tab(&mt, wa[i], wb[i], z1[i], z2[i]);
    return 0;
}

extern "C" int stmpm_debug_box_muller_star(const uint32_t* wa, const uint32_t* wb, int64_t n, double* z1, double* z2) {
    if (!wa || !wb || !z1 || !z2 || n < 0) return fail(STMPM_E_ARG, "stmpm_debug_box_muller_star: bad arguments");
    MathTables mt;
    stmpm_host_tables(&mt);
    for (int64_t i = 0; i < n; ++i) box_muller_star_tab(&mt, wa[i], wb[i], z1[i], z2[i]);
    return 0;
}

extern "C" int stmpm_toolplot_reduce(stmpm_t* h, const double* err_dev, int64_t n, int32_t bins, stmpm_toolplot* out,
                                     double* density_host, void* stream) {
    if (!h || !err_dev || !out || n <= 0 || bins < 1 || bins > 4096)
        return fail(STMPM_E_ARG, "stmpm_toolplot_reduce: bad arguments");
    DeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    struct DevBuf { void* p = nullptr; ~DevBuf() { cudaFree(p); } } acc_buf;       // freed on every exit path
    CK(cudaMalloc(&acc_buf.p, sizeof(double) * 8 + sizeof(unsigned long long) * bins));
    double* acc = static_cast<double*>(acc_buf.p);
    unsigned long long* hist = reinterpret_cast<unsigned long long*>(acc + 8);
    CK(cudaMemsetAsync(acc, 0, sizeof(double) * 8 + sizeof(unsigned long long) * bins, st));
    const unsigned long long inf_bits = 0x7FF0000000000000ull;
    CK(cudaMemcpyAsync(acc + 5, &inf_bits, 8, cudaMemcpyHostToDevice, st));
    const int grid = (int)std::min<long long>((n + 255) / 256, (long long)h->prop.multiProcessorCount * 16);
    toolplot_pass1<<<grid, 256, 0, st>>>(err_dev, n, acc);
    toolplot_pass2<<<grid, 256, 0, st>>>(err_dev, n, acc);
    toolplot_pass3<<<grid, 256, sizeof(unsigned int) * bins, st>>>(err_dev, n, acc, bins, hist);
    CK(cudaGetLastError());
    h->launches += 3;
    std::vector<double> a(8);
    std::vector<unsigned long long> hh(bins);
    CK(cudaMemcpyAsync(a.data(), acc, sizeof(double) * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(hh.data(), hist, sizeof(unsigned long long) * bins, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    std::memset(out, 0, sizeof(*out));
    out->n = a[0];
    if (a[0] <= 0) return 0;
    out->rms = std::sqrt(a[1] / a[0]);                                     // toolplot.py:82
    out->n_kept = a[2];
    if (a[2] <= 0) return 0;
    double lo, hi;
    std::memcpy(&lo, &a[5], 8); std::memcpy(&hi, &a[6], 8);
    out->kept_min = lo; out->kept_max = hi;
    int arg = 0;
    for (int i = 1; i < bins; ++i) if (hh[i] > hh[arg]) arg = i;          // numpy.argmax: first maximum
    out->mode_bin = arg;
    const double step = (hi - lo) / (double)bins;
    out->mode = lo + arg * step;                                           // h[1][hmax]: left edge (toolplot.py:90,96)
    const double m = out->mode;
    out->sigma_about_mode = std::sqrt(std::max(0.0, (a[4] - 2.0 * m * a[3] + a[2] * m * m) / a[2]));   // toolplot.py:99
    if (density_host) {
        const double w = hi > lo ? step : 1.0;
        for (int i = 0; i < bins; ++i) density_host[i] = (double)hh[i] / (a[2] * w);
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------ MonteCarlo.run_sim on the device
// monte_carlo.py:30-80 without a DataFrame, without a launch + synchronise per batch: the host enqueues rounds of chunks
// (trials_kernel + mc_rule_kernel per chunk of up to 1024 batches) and reads the state ONCE per round; every launch after
// the rule fired is a no-op (cancel word).  The first round covers twice the batches min_rows needs, later rounds grow
// fourfold, so a default run (min_rows 10 000, stop some hundred batches later) is one round: two launches, one
// synchronisation, one 70 KB read.
extern "C" int stmpm_mc_converge(stmpm_t* h, int precision, const stmpm_dist* dist, int32_t batch, uint64_t seed, int64_t trial0,
                                 int64_t min_rows, double tol, int64_t max_runs, int64_t max_batches, stmpm_mc_result* out,
                                 double* stats_host) {
    KArgs a;
    if (int rc = fill_common(h, a, precision)) return rc;
    if (!dist || !out || batch <= 0 || batch > (1 << 20) || max_batches < 0)
        return fail(STMPM_E_ARG, "stmpm_mc_converge: bad arguments (batch must be in [1, 2^20])");
    DeviceGuard guard(h->device);
    cudaStream_t st = h->streams[0];
    const long long slen = a.stats_len;
    const int chunk_batches = (int)std::max<long long>(1, std::min<long long>(MC_MAX_BATCHES, (1LL << 22) / batch));
    const long long chunk_trials = (long long)chunk_batches * batch;
    if (int rc = ensure_stage(h, chunk_trials, false)) return rc;
    if (int rc = ensure_stats(h, slen + (long long)(sizeof(McState) + 7) / 8)) return rc;
    double* d_stats = h->stage_stats;
    McState* d_state = reinterpret_cast<McState*>(h->stage_stats + slen);
    McState init{};
    init.prev = 1.0; init.ratio = 1.0;                                      // monte_carlo.py:45-46
    CK(cudaMemsetAsync(d_stats, 0, sizeof(double) * slen, st));
    CK(cudaMemcpyAsync(d_state, &init, sizeof(init), cudaMemcpyHostToDevice, st));
    to_dev_dist(*dist, a.dist0);
    a.n_segments = 1; a.k0 = (uint32_t)seed; a.k1 = (uint32_t)(seed >> 32);
    a.err = h->stage_err[0]; a.ndet = h->stage_ndet[0]; a.stats = nullptr;
    a.cancel = &d_state->done;
    McState res{};
    // first round: twice the batches min_rows needs, at least 32 (the rule cannot fire before min_rows, and typically fires
    // within that span); every later round covers four times as many batches
    long long done_batches = 0, rounds = 0;
    long long round_batches = std::max<long long>(32, (2 * std::max<long long>(min_rows, 0) + batch - 1) / batch);
    const long long limit = max_batches > 0 ? max_batches : (1LL << 40);
    for (;;) {
        const long long round_end = std::min<long long>(limit, done_batches + round_batches);
        while (done_batches < round_end) {
            const int nb = (int)std::min<long long>(chunk_batches, round_end - done_batches);
            KArgs k = a;
            k.trials_per_segment = (long long)nb * batch;
            k.trial0 = trial0 + done_batches * batch;
            const int rc = precision == 64 ? launch_trials<1, 64>(h, k, st) : launch_trials<1, 32>(h, k, st);
            if (rc) return rc;
            mc_rule_kernel<<<1, MC_MAX_BATCHES, 0, st>>>(h->stage_err[0], h->stage_ndet[0], nb, batch, d_state, d_stats, a.hist_shift,
                                                         a.n_bins, (long long)min_rows, tol, (long long)max_runs);
            CK(cudaGetLastError());
            ++h->launches;
            done_batches += nb;
        }
        ++rounds;
        CK(cudaMemcpyAsync(&res, d_state, sizeof(res), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (res.done || done_batches >= limit) break;
        round_batches = std::min<long long>(round_batches * 4, 4096LL * chunk_batches);
    }
    if (stats_host) CK(cudaMemcpy(stats_host, d_stats, sizeof(double) * slen, cudaMemcpyDeviceToHost));
    std::memset(out, 0, sizeof(*out));
    out->count = res.count; out->trials = res.count * batch; out->converged = res.done;
    out->n_ok = res.n_ok; out->n_fail = res.n_fail; out->std_ratio = res.ratio;
    out->sigma_halfnormal = res.prev;                                        // monte_carlo.py:61 (prev = sigma_hat of the last batch)
    out->mean_halfnormal = res.prev * std::sqrt(2.0 / M_PI);                 // driver.py:291,323
    out->fail_rate = (res.n_ok + res.n_fail) > 0 ? res.n_fail / (res.n_ok + res.n_fail) : 0.0;
    out->rounds = rounds; out->trials_computed = done_batches * batch;
    return check_status(h);
}

extern "C" int stmpm_toolplot_failrate(stmpm_t* h, const double* err_dev, const double* mag_dev, int64_t n,
                                       const double* levels_host, int32_t n_levels, double* rows_host, double* failed_host,
                                       void* stream) {
    if (!h || !err_dev || !mag_dev || !levels_host || !rows_host || !failed_host || n <= 0 || n_levels < 1 || n_levels > (1 << 20))
        return fail(STMPM_E_ARG, "stmpm_toolplot_failrate: bad arguments");
    for (int i = 1; i < n_levels; ++i)
        if (!(levels_host[i] > levels_host[i - 1])) return fail(STMPM_E_ARG, "stmpm_toolplot_failrate: levels must be strictly increasing");
    DeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    struct DevBuf { void* p = nullptr; ~DevBuf() { cudaFree(p); } } buf;
    const size_t n_slots = (size_t)n_levels + 1;
    const size_t bytes = 8 + sizeof(double) * n_levels + 2 * sizeof(unsigned long long) * n_slots;
    CK(cudaMalloc(&buf.p, bytes));
    double* acc = static_cast<double*>(buf.p);
    double* lev = acc + 1;
    unsigned long long* rows = reinterpret_cast<unsigned long long*>(lev + n_levels);
    unsigned long long* failed = rows + n_slots;
    CK(cudaMemsetAsync(buf.p, 0, bytes, st));
    CK(cudaMemcpyAsync(lev, levels_host, sizeof(double) * n_levels, cudaMemcpyHostToDevice, st));
    const int grid = (int)std::min<long long>((n + 255) / 256, (long long)h->prop.multiProcessorCount * 16);
    toolplot_rms_all<<<grid, 256, 0, st>>>(err_dev, n, acc);
    toolplot_failrate_kernel<<<grid, 256, 0, st>>>(err_dev, mag_dev, n, lev, n_levels, acc, rows, failed);
    CK(cudaGetLastError());
    h->launches += 2;
    std::vector<unsigned long long> hr(2 * n_slots);
    CK(cudaMemcpyAsync(hr.data(), rows, sizeof(unsigned long long) * 2 * n_slots, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    for (size_t i = 0; i < n_slots; ++i) { rows_host[i] = (double)hr[i]; failed_host[i] = (double)hr[n_slots + i]; }
    return 0;
}

extern "C" int stmpm_column_hist(stmpm_t* h, const float* x_dev, int64_t n, int32_t bins, stmpm_column_stats* out,
                                 double* density_host, void* stream) {
    if (!h || !x_dev || !out || n <= 0 || bins < 1 || bins > 4096) return fail(STMPM_E_ARG, "stmpm_column_hist: bad arguments");
    DeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    struct DevBuf { void* p = nullptr; ~DevBuf() { cudaFree(p); } } buf;
    const size_t bytes = sizeof(double) * 8 + sizeof(unsigned long long) * bins;
    CK(cudaMalloc(&buf.p, bytes));
    double* acc = static_cast<double*>(buf.p);
    unsigned long long* hist = reinterpret_cast<unsigned long long*>(acc + 8);
    CK(cudaMemsetAsync(buf.p, 0, bytes, st));
    const unsigned long long all_ones = ~0ull;                              // minimum key starts at the top
    CK(cudaMemcpyAsync(acc + 3, &all_ones, 8, cudaMemcpyHostToDevice, st));
    const int grid = (int)std::min<long long>((n + 255) / 256, (long long)h->prop.multiProcessorCount * 16);
    column_pass1<<<grid, 256, 0, st>>>(x_dev, n, acc);
    column_pass2<<<grid, 256, sizeof(unsigned int) * bins, st>>>(x_dev, n, acc, bins, hist);
    CK(cudaGetLastError());
    h->launches += 2;
    double a[8];
    std::vector<unsigned long long> hh(bins);
    CK(cudaMemcpyAsync(a, acc, sizeof(a), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(hh.data(), hist, sizeof(unsigned long long) * bins, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    std::memset(out, 0, sizeof(*out));
    out->n = a[0];
    if (a[0] <= 0) return 0;
    auto unkey = [](double d) { unsigned long long k; std::memcpy(&k, &d, 8); k = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
                                double r; std::memcpy(&r, &k, 8); return r; };
    out->min = unkey(a[3]); out->max = unkey(a[4]);
    out->mean = a[1] / a[0];
    out->std = a[0] > 1 ? std::sqrt(std::max(0.0, (a[2] - a[1] * a[1] / a[0]) / (a[0] - 1.0))) : 0.0;    // pandas .std(): ddof = 1
    out->std_halfnormal = out->std * std::sqrt(1.0 - 2.0 / M_PI);                                         // toolplot.py:248,254
    if (density_host) {
        const double w = out->max > out->min ? (out->max - out->min) / bins : 1.0;
        for (int i = 0; i < bins; ++i) density_host[i] = (double)hh[i] / (a[0] * w);
    }
    return 0;
}

extern "C" int stmpm_launch_count(const stmpm_t* h, int64_t* n) {
    if (!h || !n) return fail(STMPM_E_ARG, "null");
    *n = h->launches;
    return 0;
}
