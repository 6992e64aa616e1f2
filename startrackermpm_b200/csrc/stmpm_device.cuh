// stmpm_device.cuh — device-side building blocks of the trial (docs/MODEL.md). sm_100a only.
//
// Replaces the per-row body of Simulation.run_sim + quest_objective (called at /root/reference/monte_carlo.py:52 and
// sensitivity.py:77; bodies absent from the mount).  One trial per thread:
//   key pass : per CTA, a window of up to 2048 trials is ordered by its work key (the number of stars the trial is
//              expected to evaluate: a table look-up from boresight cell and MAX_MAGNITUDE) — counting sort by one warp
//              while the others consume the previous window — so the 32 trials a warp runs together do the same
//              amount of work; keys never affect results
//   set-up (fp64) -> phase A: candidate scan with a conservative fp32 pre-filter out of shared memory
//                 -> phase B: exact per-star projection + Philox centroid noise + detection + B accumulation
//                 -> QUEST Newton solve in the body frame + error angle.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>

#if defined(__CUDACC__)
#define STMPM_HD __host__ __device__ __forceinline__
#else
#define STMPM_HD inline
#endif

namespace stmpm {

// 384 threads (12 warps, 168 registers: no spills) running TWO stars per star-loop iteration beat 512 threads (16 warps, 128
// registers) running one: +4.7 % (round 2, exp U) — the loop is bound by dependency latency, and a second independent chain per
// thread hides more of it than a fourth warp per scheduler.  -DSTMPM_TPB=512 -DSTMPM_ONE_STAR -DSTMPM_LIST_CAP=48 rebuilds v16.
#if !defined(STMPM_TPB)
#define STMPM_TPB 384
#endif
constexpr int TPB = STMPM_TPB;          // threads per CTA (1 CTA / SM: the fp32 catalog lives in shared memory)
#if !defined(STMPM_LIST_CAP)
#define STMPM_LIST_CAP 64         // 48 KB at 384 threads; 48: -2.4 %, 32: -4.4 % (more drain rounds); larger does not fit beside the catalog
#endif
constexpr int LIST_CAP = STMPM_LIST_CAP;   // per-thread survivor list in shared memory (phase A -> phase B hand-off), drained in rounds
constexpr int WIN_MAX = 2048;     // largest CTA sort window (trials)
constexpr int WIN_PRESAMPLED = 256;   // default window, pre-sampled columns: wider windows sort better but every column gather then
                                      // touches 32 different lines, and the LSU pipe (shared with shared memory) saturates
constexpr int WIN_STAGED = 2048;      // default window, pre-sampled columns staged in key order (round 2): as wide as native RNG
constexpr int WIN_RNG = 2048;         // default window, native RNG: no gathers, so the widest sort wins
constexpr int KEY_LEVELS = 128;   // work-key table: candidate visits per lookup cell at magnitude levels KEY_M0 + (k+1)/16
constexpr float KEY_M0 = 1.0f;
constexpr int KEY_ROLLS = 8;      // work-key table: orientation bins of the array footprint (roll modulo its symmetry period)
constexpr int KEY_BINS = 256;     // sort bins (= candidate visits, saturated)
constexpr uint32_t PHILOX_M0 = 0xD2511F53u, PHILOX_M1 = 0xCD9E8D57u, PHILOX_W0 = 0x9E3779B9u, PHILOX_W1 = 0xBB67AE85u;
constexpr uint32_t PHILOX2_M = 0xD256D193u;
constexpr uint32_t TAG_STAR = 0x53544152u, TAG_PARA = 0x50415241u;
constexpr double RAD2ARCSEC = 206264.80624709636;
constexpr double DEG2RAD = 0.017453292519943295;
constexpr double TWO_M32 = 2.3283064365386963e-10;   // 2^-32

struct __align__(32) StarRec {    // exactly one 32-byte sector per star
    double x, y, z;
    uint32_t star_id;
    float mag;
};

struct TrialParams {              // one DataFrame row (MODEL.md §1)
    double ra, dec, roll, f, ex, ey, ez, phi, theta, psi, devx, devy, maxmag;
};

// ------------------------------------------------------------------------------------------------ Philox (MODEL.md §5)
STMPM_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

// Philox4x32-10 (Salmon et al., SC'11).  Ten rounds, key bumped between rounds.
STMPM_HD void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                            uint32_t& o0, uint32_t& o1, uint32_t& o2, uint32_t& o3) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = mulhi32(PHILOX_M0, c0), lo0 = PHILOX_M0 * c0;
        const uint32_t hi1 = mulhi32(PHILOX_M1, c2), lo1 = PHILOX_M1 * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += PHILOX_W0;
        k1 += PHILOX_W1;
    }
    o0 = c0; o1 = c1; o2 = c2; o3 = c3;
}

// Philox2x32-10: one multiply per round — the per-star generator of the hot loop.
STMPM_HD void philox2x32_10(uint32_t c0, uint32_t c1, uint32_t k, uint32_t& o0, uint32_t& o1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
#if defined(__CUDA_ARCH__) && defined(STMPM_PHILOX_WIDE)
        // experiment (round 2, exp V): force ONE IMAD.WIDE per round (ptxas splits the second chain of the two-star loop
        // into IMAD + IMAD.HI)
        uint32_t hi, lo;
        asm("{\n\t.reg .u64 t;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {%0, %1}, t;\n\t}" : "=r"(lo), "=r"(hi) : "r"(c0), "r"(PHILOX2_M));
#else
        const uint32_t hi = mulhi32(PHILOX2_M, c0), lo = PHILOX2_M * c0;
#endif
        c0 = hi ^ k ^ c1;
        c1 = lo;
        k += PHILOX_W0;
    }
    o0 = c0; o1 = c1;
}

// per-trial key pair of the star-noise streams
STMPM_HD void star_key(uint32_t t_lo, uint32_t t_hi, uint32_t k0, uint32_t k1, uint32_t& ka, uint32_t& kb) {
    uint32_t x2, x3;
    philox4x32_10(t_lo, t_hi, 0u, TAG_STAR, k0, k1, ka, kb, x2, x3);
}

STMPM_HD double u01(uint32_t w) { return ((double)w + 0.5) * TWO_M32; }

// ------------------------------------------------------------------------------------------------ table-based Box-Muller
// MODEL.md §5: R = sqrt(-2 ln u1), z1 = R cos(2 pi u2), z2 = R sin(2 pi u2), u = (w + 0.5) 2^-32.
// Evaluated to ~1e-15 absolute with integer argument reduction (no int->double conversions: those go to the XU pipe)
// and small tables instead of libm's long polynomials: the star loop is FP64-pipe bound.
//   ln:  2w+1 = m 2^p, m in [1,2); j = top 8 mantissa bits; r = m / c_j - 1, |r| <= 2^-9; ln m = ln c_j + log1p(r)
//   sin/cos: 1024 table intervals over the full circle (top 10 bits of w); |d| <= pi/1024 remainder by Taylor.
// Dense tables keep the polynomials — i.e. the dependent fp64 chain of the star loop — short.
struct __align__(16) MathTables {
    double log_tab[256][2];                 // {1 / c_j, -2 ln c_j},  c_j = 1 + (j + 0.5) / 256   (one 16-byte load)
    double sc_tab[1024][2];                 // {sin, cos} of 2 pi (j + 0.5) / 1024: the full circle, no quadrant logic
    double k_ln2[34];                       // 2 k ln 2
    float sc32[128][2];                     // fp32 mode: {sin, cos} of 2 pi (j + 0.5) / 128   (1 KB: what shared memory has left)
};
STMPM_HD void load2(const double* p, double& a, double& b) {
#if defined(__CUDA_ARCH__)
    const double2 v = *reinterpret_cast<const double2*>(p);
    a = v.x; b = v.y;
#else
    a = p[0]; b = p[1];
#endif
}

STMPM_HD double bits_to_double(unsigned long long b) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)b);
#else
    double d;
    memcpy(&d, &b, 8);
    return d;
#endif
}
STMPM_HD unsigned long long double_to_bits(double d) {
#if defined(__CUDA_ARCH__)
    return (unsigned long long)__double_as_longlong(d);
#else
    unsigned long long b;
    memcpy(&b, &d, 8);
    return b;
#endif
}
// a/b (and, in seed_rsqrt-based code below, 1/sqrt(x)) for normal, finite operands (every use below: no zero / inf / denormal / NaN handling needed, which
// is what libm's expansions spend their branches on): MUFU seed (rel. error ~2^-22) + ONE Newton step -> ~2^-44 = 6e-14.
// That is 5 orders of magnitude inside the parity bar for each use: the noise radius R (|dz| < 4e-13), the projection
// scale n.p0 / n.r (centroid error < 1e-10 px) and the back-projection norm (a common scale of b, i.e. a 6e-14 relative
// change of one star's Wahba weight).  The star loop is a single dependent chain, so every step removed is latency saved.
STMPM_HD double fast_div(double a, double b) {
#if defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    return a * (r * fma(-b, r, 2.0));
#else
    return a / b;
#endif
}

STMPM_HD double seed_rsqrt(double x) {   // ~2^-22 relative; x normal, positive, finite
#if defined(__CUDA_ARCH__)
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
#else
    return (double)(float)(1.0 / sqrt(x));
#endif
}

// Table accessors: a plain pointer (host build, global memory) or a 32-bit shared-memory address (the trial kernel: explicit
// ld.shared on a base kept in a register, so ptxas does not rebuild the shared-window base inside the star loop).
struct TabPtr {
    const MathTables* T;
    STMPM_HD void log(int j, double& a, double& b) const { load2(T->log_tab[j], a, b); }
    STMPM_HD void sc(int j, double& a, double& b) const { load2(T->sc_tab[j], a, b); }
    STMPM_HD double k2(int k) const { return T->k_ln2[k]; }
    STMPM_HD void sc32(int j, float& a, float& b) const { a = T->sc32[j][0]; b = T->sc32[j][1]; }
};
#if defined(__CUDACC__)
__device__ __forceinline__ void lds_d2(uint32_t addr, double& a, double& b) {
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "r"(addr));
}
struct TabSmem {
    uint32_t base;   // shared address of the MathTables copy
    __device__ __forceinline__ void log(int j, double& a, double& b) const { lds_d2(base + (uint32_t)j * 16u, a, b); }
    __device__ __forceinline__ void sc(int j, double& a, double& b) const { lds_d2(base + 4096u + (uint32_t)j * 16u, a, b); }
    __device__ __forceinline__ double k2(int k) const {
        double v;
        asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(base + 4096u + 16384u + (uint32_t)k * 8u));
        return v;
    }
    __device__ __forceinline__ void sc32(int j, float& a, float& b) const {
        asm("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(a), "=f"(b) : "r"(base + 4096u + 16384u + 272u + (uint32_t)j * 8u));
    }
};
static_assert(offsetof(MathTables, sc32) == 4096 + 16384 + 272, "TabSmem::sc32 offset");
#endif
STMPM_HD void box_muller_tab(const MathTables* __restrict__ T, uint32_t wa, uint32_t wb, double& z1, double& z2);

// STAR = false: the full-accuracy draw (|dz| < 2e-13) — the parameter sampler (MODEL.md §6: columns reproduce the oracle's to
//   1e-11) and the debug entry.
// STAR = true: the per-star centroid-noise draw of the star loop, three FP64-pipe instructions shorter (round 2, exp X:
//   +1.2 %): the r^5/5 term of log1p (< 5.6e-15) and the d^4/24 term of cos d (< 3.7e-12) dropped, the angle scale folded
//   into the remainder's coefficients.  |dz| < 2.5e-11, i.e. a centroid moved by < 2.5e-11 sigma_c px — the 1e-9 rad bar
//   needs |dz| <~ 3e-7 (sigma_c * dz over a roll lever arm of ~100 px); parity at 2e6 trials x 3 configs is unchanged by it
//   (max |d error| 6.9e-13 / 1.3e-12 / 9e-14 rad, tests/test_gpu_scale.py: the Wahba solve's rounding dominates).
//   -DSTMPM_NO_TRIM64 gives the star loop the full draw again.
template <bool STAR, class Tab>
STMPM_HD void box_muller_t(const Tab T, uint32_t wa, uint32_t wb, double& z1, double& z2) {
#if defined(STMPM_NO_TRIM64)
    constexpr bool TRIM = false;
#else
    constexpr bool TRIM = STAR;
#endif
    // ---- L = -2 ln u1,  u1 = (2 wa + 1) 2^-33.  Exact int -> double through the 2^52 magic number (one DADD, no
    // conversion instruction, no clz / shift normalisation); exponent and mantissa then come from the bit pattern.
    const unsigned long long v = 2ull * wa + 1ull;                         // 33-bit odd integer
    const double tv = bits_to_double(0x4330000000000000ull | v) - 4503599627370496.0;   // == (double)v
    const unsigned long long tb = double_to_bits(tv);
    const int k = 1056 - (int)(tb >> 52);                                  // 33 - unbiased exponent, in [1, 33]
    const double m = bits_to_double((tb & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull);   // mantissa in [1, 2)
    const int j = (int)(tb >> 44) & 0xFF;                                  // top 8 mantissa bits
    double inv_c, m2ln_c;
    T.log(j, inv_c, m2ln_c);                                               // {1 / c_j, -2 ln c_j}
    const double r = fma(m, inv_c, -1.0);
    // log1p(r) = r - r^2/2 + r^3/3 - r^4/4 + r^5/5, |r| <= 2^-9 (truncation r^6/6 < 1e-17), Estrin form: depth 3
    const double r2 = r * r;
    double p;
#if defined(STMPM_HI_COEF)
    // experiment (round 2, exp J): 1/3 and 1/5 truncated to their high words (relative 2.4e-7 on terms below 2.5e-9), so that
    // they encode as DFMA immediates instead of two register moves each
    p = fma(r2, fma(r2, fma(r, 0.19999992847442627, -0.25), fma(r, 0.33333325386047363, -0.5)), r);
#else
    if (TRIM) p = fma(r2, fma(r2, -0.25, fma(r, 1.0 / 3.0, -0.5)), r);
    else p = fma(r2, fma(r2, fma(r, 0.2, -0.25), fma(r, 1.0 / 3.0, -0.5)), r);
#endif
#if !defined(STMPM_K2_TABLE)
    // -2 ln m + 2 k ln 2, the second term by arithmetic (k exact through the 2^52 magic number): the table read it replaces
    // was the loop's largest single short-scoreboard wait (4 % of the stall samples; round 2 exp H: +0.3 %)
    const double kd = bits_to_double(0x4330000000000000ull | (unsigned long long)(unsigned)k) - 4503599627370496.0;
    const double L = fma(kd, 1.3862943611198906, fma(-2.0, p, m2ln_c));
#else
    const double L = fma(-2.0, p, m2ln_c) + T.k2(k);                   // -2 ln m + 2 k ln 2  (tables hold the factor 2)
#endif
    // R = sqrt(L): seed y ~ 1/sqrt(L), R0 = L y, one Newton step on the square root: R = R0 + (L - R0^2) y / 2
    const double y = seed_rsqrt(L);
    const double R0 = L * y;
    const double R = fma(fma(-R0, R0, L), 0.5 * y, R0);
    // ---- angle 2 pi u2 = 2 pi (j2 + 0.5) / 1024 + d,  j2 = top 10 bits of wb
    const int j2 = (int)(wb >> 22);
    const uint32_t rem = wb & 0x3FFFFFu;
    // d = (pi/512) ((rem + 0.5) / 2^22 - 0.5), |d| <= pi/1024, via the 2^52 magic number (exact, no conversion instruction)
    const double t = bits_to_double(0x4330000000000000ull | (unsigned long long)(2u * rem + 1u));
    constexpr double kc = 7.314590396335798e-10;                         // pi / 512 / 2^23
    double sd, cd;
    if (TRIM) {
        // e = t - (2^52 + 2^22), d = c e:  sin d = e (c - c^3 e^2 / 6),  cos d = 1 - c^2 e^2 / 2 — five instead of seven
        // FP64-pipe instructions, and cos d does not wait for d
        const double e = t - 4503599631564800.0;
        const double q = e * e;
        sd = e * fma(q, -(kc * kc * kc) / 6.0, kc);
        cd = fma(q, -(kc * kc) / 2.0, 1.0);
    } else {
        const double d = (t - 4503599631564800.0) * kc;
        const double d2 = d * d;
#if defined(STMPM_HI_COEF)
        sd = fma(d * d2, -0.16666662693023682, d);                       // high-word-only 1/6: 1.2e-15 on top of the truncation
        cd = fma(d2, fma(d2, 0.041666656732559204, -0.5), 1.0);
#else
        sd = fma(d * d2, -1.0 / 6.0, d);                                 // truncation d^5/120 < 2.3e-15
        cd = fma(d2, fma(d2, 1.0 / 24.0, -0.5), 1.0);                    // truncation d^6/720 < 1.2e-18
#endif
    }
    double S, C;
    T.sc(j2, S, C);
    const double sq = fma(S, cd, C * sd);
    const double cq = fma(C, cd, -S * sd);
    z1 = R * cq;
    z2 = R * sq;
}
STMPM_HD void box_muller_tab(const MathTables* __restrict__ T, uint32_t wa, uint32_t wb, double& z1, double& z2) {
    box_muller_t<false>(TabPtr{T}, wa, wb, z1, z2);
}
STMPM_HD void box_muller_star_tab(const MathTables* __restrict__ T, uint32_t wa, uint32_t wb, double& z1, double& z2) {
    box_muller_t<true>(TabPtr{T}, wa, wb, z1, z2);
}

// ------------------------------------------------------------------------------------------------ fp32-mode Box-Muller
// The same draw (MODEL.md §5) in fp32, |dz| <~ 3e-6: two orders inside what the 1e-6 rad bar needs (a centroid moved by
// sigma_c * 3e-6 px).  No int -> float conversions (XU pipe): integers enter through the 2^23 magic number.  Two MUFU
// operations (lg2, sqrt); sin / cos from a 128-entry fp32 table + a cubic / quartic remainder.
STMPM_HD float uint_as_float(uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(b);
#else
    float f;
    memcpy(&f, &b, 4);
    return f;
#endif
}
STMPM_HD float lg2_fast(float x) {
#if defined(__CUDA_ARCH__)
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#else
    return log2f(x);
#endif
}
STMPM_HD float sqrt_fast(float x) {
#if defined(__CUDA_ARCH__)
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#else
    return sqrtf(x);
#endif
}
template <class Tab>
STMPM_HD void box_muller_f32(const Tab T, uint32_t wa, uint32_t wb, float& z1, float& z2) {
    // u1 = (wa + 0.5) 2^-32 from two exact pieces (wa >> 9, (wa & 511) + 0.5), one rounding
    const float hi = uint_as_float(0x4B000000u | (wa >> 9)) - 8388608.0f;
    const float lo = uint_as_float(0x4B000000u | (wa & 511u)) - 8388607.5f;
    const float u1 = fmaf(hi, 512.0f, lo) * 2.3283064365386963e-10f;
    float L = -1.3862943611198906f * lg2_fast(u1);                       // -2 ln u1
    // within 2^-10 of one the nearest float has lost eps = 1 - u1: take eps = (2^32 - wa - 0.5) 2^-32 exactly, and
    // -2 ln(1 - eps) = eps (2 + eps (1 + 2 eps / 3)) + O(eps^4)
    const uint32_t c = ~wa;
    const float eps = (uint_as_float(0x4B000000u | (c & 0x3FFFFFu)) - 8388607.5f) * 2.3283064365386963e-10f;
    const float Ls = eps * fmaf(eps, fmaf(eps, 0.66666669f, 1.0f), 2.0f);
    L = (c < (1u << 22)) ? Ls : L;
    const float R = sqrt_fast(L);
    // angle 2 pi u2 = 2 pi (j + 0.5) / 128 + d,  j = top 7 bits of wb, d from the next 23 bits (the last 2 bits move the
    // angle by < 6e-9 rad), |d| <= pi / 128
    const int j = (int)(wb >> 25);
    const float t = uint_as_float(0x4B000000u | ((wb >> 2) & 0x7FFFFFu)) - 12582912.0f;     // in [-2^22, 2^22)
    const float d = fmaf(t, 5.8516723170686385e-9f, 2.9258361585343192e-9f);               // (t + 0.5) * 8 pi / 2^32
    const float d2 = d * d;
    const float sd = fmaf(d * d2, -0.16666667f, d);                         // d^5 / 120 < 8e-11
    const float cm = d2 * fmaf(d2, 0.041666668f, -0.5f);                    // cos d - 1; d^6 / 720 < 4e-13
    float S, C;
    T.sc32(j, S, C);
    z1 = R * fmaf(C, cm, fmaf(-S, sd, C));
    z2 = R * fmaf(S, cm, fmaf(C, sd, S));
}

#if defined(__CUDACC__)
// float -> double by integer operations (exact for normal floats; a zero or denormal input becomes ~1e-38, harmless as
// an additive centroid noise): F2F conversions run on the 16-lane XU pipe, six of them per star made fp32 mode slower
// than fp64 mode in round 1.
__device__ __forceinline__ double f2d_bits(float f) {
    const uint32_t b = __float_as_uint(f);
    const uint32_t hi = (((b & 0x7fffffffu) >> 3) + 0x38000000u) | (b & 0x80000000u);
    return __hiloint2double((int)hi, (int)(b << 29));
}

// Everything phase B needs.  Inertial-frame plane vectors: NI = C n, EU = C e_u, EV = C e_v, so that for a catalog vector
// s:  n.r = NI.s,  u = (n.p0 / NI.s) EU.s - e_u.p0,  v likewise  (r = C^T s never materialised).
struct Ctx64 {
    double C[9];                  // body -> inertial, row-major
    double NI[3], EU[3], EV[3];
    double np0, cu, cv;           // n.p0, e_u.p0, e_v.p0
    double sx, sy;                // |BASE_DEV_X|, |BASE_DEV_Y|
};
// Everything phase A (the conservative fp32 candidate scan) needs.
struct Ctx32 {
    float NI[3], EU[3], EV[3], np0, cu, cv, limu, limv, magf;
    float sxf, syf;               // fp32 mode: |BASE_DEV_X|, |BASE_DEV_Y|
    int guard_hi;                 // fp32 mode: high word of the guard distance around the array edges (see star_f32)
    float ra0, dec0, cone;        // boresight (lookup-cell key) and scan-cone half angle (cone: see cone_angle)
    float cone_t;                 // tangent of the scan-cone half angle without its 0.002 rad margin; +inf = no usable cone
};

__device__ __forceinline__ float opaque(float x) {   // stop ptxas re-deriving an fp32 copy from its fp64 source at every
    asm volatile("" : "+f"(x));                       // use (F2F conversions run on the 16-lane XU pipe)
    return x;
}

// sin / cos of a small angle by Taylor series: |x| <= 0.1 rad -> truncation < 3e-19 (sin), 3e-17 (cos).  The plane tilts
// are fractions of a degree; libm's sincos spends ~4x the instructions on range reduction they do not need.
__device__ __forceinline__ void sincos_small(double x, double* s, double* c) {
    const double x2 = x * x;
    *s = x * fma(x2, fma(x2, fma(x2, fma(x2, 1.0 / 362880.0, -1.0 / 5040.0), 1.0 / 120.0), -1.0 / 6.0), 1.0);
    *c = fma(x2, fma(x2, fma(x2, fma(x2, fma(x2, -1.0 / 3628800.0, 1.0 / 40320.0), -1.0 / 720.0), 1.0 / 24.0), -0.5), 1.0);
}

// Experiment (round 2, exp I; STMPM_SMALL_CODE): the trial kernel's 16 warps sit in different phases of a ~100 KB instruction
// stream and ncu shows them waiting for instruction fetch (`no_instruction` 0.23 per issue in fp64 mode, 0.51 in fp32 mode), so
// libm's sincos / atan2 and fp32 mode's exact fall-back were tried as ONE out-of-line copy each.  Measured WORSE: fp64 -2.5 %,
// fp32 -50 % (an out-of-line callee takes the plane context and B by address, which moves them to local memory inside the star
// loop).  Inlined by default.
#if defined(STMPM_SMALL_CODE)
#define STMPM_LIBM_INLINE __noinline__
#else
#define STMPM_LIBM_INLINE __forceinline__
#endif
__device__ STMPM_LIBM_INLINE void sincos_libm(double x, double* s, double* c) { sincos(x, s, c); }
__device__ STMPM_LIBM_INLINE double atan2_libm(double y, double x) { return atan2(y, x); }

// sin / cos of an attitude angle, |x| <= 1000: three-term Cody-Waite reduction by pi/2 (exact with FMA for |k| < 2^10), fdlibm's
// kernel polynomials on |r| <= pi/4, quadrant by the low bits of the rounded quotient.  <= 2.3e-16 absolute error (checked against
// exact-arithmetic emulation over +-1000): the same 1-ulp class as libm, at ~40 instructions instead of libm's ~76 executed (its
// fast path carries the hand-over to a Payne-Hanek slow path that attitude angles never reach).  Round 2, exp R.
__device__ __forceinline__ void sincos_attitude(double x, double* sn, double* cs) {
    const double t = fma(x, 0.6366197723675814, 6755399441055744.0);             // 2/pi; 1.5 * 2^52: t's low word = rint(x * 2/pi)
    const double j = t - 6755399441055744.0;
    const int q = __double2loint(t);
    double r = fma(-j, 1.5707963267948966, x);
    r = fma(-j, 6.123233995736766e-17, r);
    r = fma(-j, -1.4973849048591698e-33, r);
    const double z = r * r;
    double ps = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
    ps = fma(z, ps, 2.75573137070700676789e-06);
    ps = fma(z, ps, -1.98412698298579493134e-04);
    ps = fma(z, ps, 8.33333333332248946124e-03);
    ps = fma(z, ps, -1.66666666666666324348e-01);
    const double s = fma(r * z, ps, r);
    double pc = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
    pc = fma(z, pc, -2.75573143513906633035e-07);
    pc = fma(z, pc, 2.48015872894767294178e-05);
    pc = fma(z, pc, -1.38888888888741095749e-03);
    pc = fma(z, pc, 4.16666666666666019037e-02);
    const double c = fma(z * z, pc, fma(z, -0.5, 1.0));
    const double a = (q & 1) ? c : s, b = (q & 1) ? s : c;                        // q = 0: (s, c)  1: (c, -s)  2: (-s, -c)  3: (-c, s)
    *sn = (q & 2) ? -a : a;
    *cs = ((q + 1) & 2) ? -b : b;
}

__device__ __forceinline__ void setup64(const TrialParams& p, Ctx64& c) {
    double sa, ca, sd, cd, sr, cr;
#if defined(STMPM_SINCOS_LOOP)
    {   // ONE copy of libm's sincos (~160 instructions) run three times instead of three inlined copies: the per-trial code is
        // executed once per group by 16 warps in 16 different places, and its footprint decides whether the hot instruction
        // stream (~35 KB) fits the 32 KB L1.5 instruction cache (round 2, exp O)
        // (through local arrays: no_instruction 0.24 -> 0.12 per issue but long_scoreboard +0.23 and -2.9 %; this form rotates
        // the results through registers instead)
        double x = p.ra;
        sa = ca = sd = cd = sr = cr = 0.0;
#pragma unroll 1
        for (int i = 0; i < 3; ++i) {
            double s_, c_;
            sincos(x, &s_, &c_);
            sa = sd; ca = cd; sd = sr; cd = cr; sr = s_; cr = c_;
            x = i == 0 ? p.dec : p.roll;
        }
    }
#elif !defined(STMPM_LIBM_SINCOS)
    if (fmax(fabs(p.ra), fmax(fabs(p.dec), fabs(p.roll))) <= 1000.0) {     // practically always, and warp-uniform
        sincos_attitude(p.ra, &sa, &ca);
        sincos_attitude(p.dec, &sd, &cd);
        sincos_attitude(p.roll, &sr, &cr);
    } else {                                                                // one cold copy of libm's sincos for far-out angles
        double x = p.ra;
        sa = ca = sd = cd = sr = cr = 0.0;
#pragma unroll 1
        for (int i = 0; i < 3; ++i) {
            double s_, c_;
            sincos(x, &s_, &c_);
            sa = sd; ca = cd; sd = sr; cd = cr; sr = s_; cr = c_;
            x = i == 0 ? p.dec : p.roll;
        }
    }
#else
    sincos_libm(p.ra, &sa, &ca);
    sincos_libm(p.dec, &sd, &cd);
    sincos_libm(p.roll, &sr, &cr);
#endif
    // C = Rz(ra) Ry(pi/2 - dec) Rz(roll)   (MODEL.md §2; primitives: constants.py:81-94)
    c.C[0] = ca * sd * cr - sa * sr;  c.C[1] = -ca * sd * sr - sa * cr;  c.C[2] = ca * cd;
    c.C[3] = sa * sd * cr + ca * sr;  c.C[4] = -sa * sd * sr + ca * cr;  c.C[5] = sa * cd;
    c.C[6] = -cd * cr;                c.C[7] = cd * sr;                  c.C[8] = sd;
    double sp, cp, st, ct, ss, cs;
    const double phi = p.phi * DEG2RAD, theta = p.theta * DEG2RAD, psi = p.psi * DEG2RAD;
    if (fmax(fabs(phi), fmax(fabs(theta), fabs(psi))) <= 0.1) {   // practically always, and warp-uniform
        sincos_small(phi, &sp, &cp);
        sincos_small(theta, &st, &ct);
        sincos_small(psi, &ss, &cs);
    } else {
        sincos_libm(phi, &sp, &cp);
        sincos_libm(theta, &st, &ct);
        sincos_libm(psi, &ss, &cs);
    }
    // R_p = Rx(phi) Ry(theta) Rz(psi): columns e_u, e_v, n
    const double eu[3] = {ct * cs, cp * ss + sp * st * cs, sp * ss - cp * st * cs};
    const double ev[3] = {-ct * ss, cp * cs - sp * st * ss, sp * cs + cp * st * ss};
    const double nn[3] = {st, -sp * ct, cp * ct};
    const double p0[3] = {p.ex, p.ey, p.f + p.ez};
    c.np0 = nn[0] * p0[0] + nn[1] * p0[1] + nn[2] * p0[2];
    c.cu = eu[0] * p0[0] + eu[1] * p0[1] + eu[2] * p0[2];
    c.cv = ev[0] * p0[0] + ev[1] * p0[1] + ev[2] * p0[2];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        c.NI[i] = c.C[3 * i] * nn[0] + c.C[3 * i + 1] * nn[1] + c.C[3 * i + 2] * nn[2];
        c.EU[i] = c.C[3 * i] * eu[0] + c.C[3 * i + 1] * eu[1] + c.C[3 * i + 2] * eu[2];
        c.EV[i] = c.C[3 * i] * ev[0] + c.C[3 * i + 1] * ev[1] + c.C[3 * i + 2] * ev[2];
    }
    c.sx = fabs(p.devx);
    c.sy = fabs(p.devy);
}

// pre-filter window and scan cone shared by both fp32 set-ups.  Window = array + 6.8 sigma_centroid (Box-Muller with a
// 32-bit u1 has |z| <= sqrt(-2 ln 2^-33) = 6.77) + slack for fp32 rounding of the projection (~1e-3 px at f = 3500).
// Cone: every point of the (tilted, shifted, noise-padded) array lies within atan(reach / depth) of the boresight.
__device__ __forceinline__ void window_and_cone(const TrialParams& p, float half_u, float half_v, Ctx32& c) {
    const float ff = fabsf((float)p.f);
    const float noise = 6.8f * fmaxf(fabsf((float)p.devx), fabsf((float)p.devy)) + 0.05f + 1e-5f * ff;
    c.limu = opaque(half_u + noise);
    c.limv = opaque(half_v + noise);
    c.sxf = opaque(fabsf((float)p.devx));
    c.syf = opaque(fabsf((float)p.devy));
    c.guard_hi = __double2hiint(fma(1e-4, fabs(p.devx) + fabs(p.devy), 1e-7));
    asm volatile("" : "+r"(c.guard_hi));
    // (double)mag <= maxmag  <=>  mag <= round-down-to-float(maxmag); clamped to a finite value so that the +inf
    // magnitude of the candidate lists' terminator entry never passes
    c.magf = opaque(fminf(__double2float_rd(p.maxmag), 3.0e38f));
#if defined(STMPM_CONE_ATAN)
    const float reach = 1.41421357f * (fmaxf(half_u, half_v) + noise) + hypotf((float)p.ex, (float)p.ey) + 1.0f;
#else
    const float reach = 1.41421357f * (fmaxf(half_u, half_v) + noise) + (fabsf((float)p.ex) + fabsf((float)p.ey)) + 1.0f;   // >= the hypotenuse
#endif
    const float tilt = fminf((fabsf((float)p.phi) + fabsf((float)p.theta)) * 0.0174532925f, 0.5f);
    const float depth = (float)p.f + (float)p.ez - reach * tilt;
    // The hot path only asks "is the cone inside the one the candidate lists were built for": it compares tangents
    // (KArgs::lk_cone_tan, host-computed) and never needs the angle — atanf, a full-precision division and hypotf were ~55
    // instructions per trial.  The fall-back walks derive the angle themselves (cone_angle).
    c.cone_t = (depth > 1.0f) ? __fdividef(reach, depth) * 1.000001f : INFINITY;
#if defined(STMPM_CONE_ATAN)
    c.cone = (depth > 1.0f) ? atanf(reach / depth) + 0.002f : 3.2f;
#else
    c.cone = 3.2f;
#endif
}
__device__ __forceinline__ float cone_angle(const Ctx32& c) { return c.cone_t < 1e30f ? atanf(c.cone_t) + 0.002f : 3.2f; }

// fp32 copies of the plane geometry (what prefilter_geom reads).  Cheap enough to redo for the rare re-scan, so the twelve
// floats need not stay live across the fp64 star loop.
__device__ __forceinline__ void geom32(const Ctx64& d, Ctx32& c) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        c.NI[i] = opaque((float)d.NI[i]);
        c.EU[i] = opaque((float)d.EU[i]);
        c.EV[i] = opaque((float)d.EV[i]);
    }
    c.np0 = opaque((float)d.np0);
    c.cu = opaque((float)d.cu);
    c.cv = opaque((float)d.cv);
}
// boresight = third column of C; its (ra, dec) key the lookup grid
__device__ __forceinline__ void boresight32(const Ctx64& d, Ctx32& c) {
    c.dec0 = asinf(fminf(fmaxf((float)d.C[8], -1.0f), 1.0f));
    const float a0 = atan2f((float)d.C[5], (float)d.C[2]);
    c.ra0 = a0 < 0.0f ? a0 + 6.28318531f : a0;
}
// fp32 context derived from the fp64 one
__device__ __forceinline__ void derive32(const TrialParams& p, const Ctx64& d, float half_u, float half_v, Ctx32& c) {
    geom32(d, c);
    window_and_cone(p, half_u, half_v, c);
    boresight32(d, c);
}

// fp32 context straight from the parameters (pass 1): fp32 sincosf of the fp32-rounded angles.  Angle errors ~|x| 6e-8
// move a star by < 0.03 px for |x| <= 64 rad, inside the window slack; larger arguments are refused (caller takes the
// fp64-derived path).  Returns false when the trial must take that path.
__device__ __forceinline__ bool setup32(const TrialParams& p, float half_u, float half_v, Ctx32& c) {
    const float ra = (float)p.ra, dec = (float)p.dec, roll = (float)p.roll;
    if (!(fabsf(ra) <= 64.0f && fabsf(dec) <= 64.0f && fabsf(roll) <= 64.0f)) return false;
    float sa, ca, sd, cd, sr, cr, sp, cp, st, ct, ss, cs;
    sincosf(ra, &sa, &ca);
    sincosf(dec, &sd, &cd);
    sincosf(roll, &sr, &cr);
    sincosf((float)p.phi * 0.0174532925f, &sp, &cp);
    sincosf((float)p.theta * 0.0174532925f, &st, &ct);
    sincosf((float)p.psi * 0.0174532925f, &ss, &cs);
    const float C[9] = {ca * sd * cr - sa * sr, -ca * sd * sr - sa * cr, ca * cd,
                        sa * sd * cr + ca * sr, -sa * sd * sr + ca * cr, sa * cd,
                        -cd * cr, cd * sr, sd};
    const float eu[3] = {ct * cs, cp * ss + sp * st * cs, sp * ss - cp * st * cs};
    const float ev[3] = {-ct * ss, cp * cs - sp * st * ss, sp * cs + cp * st * ss};
    const float nn[3] = {st, -sp * ct, cp * ct};
    const float p0[3] = {(float)p.ex, (float)p.ey, (float)p.f + (float)p.ez};
    c.np0 = nn[0] * p0[0] + nn[1] * p0[1] + nn[2] * p0[2];
    c.cu = eu[0] * p0[0] + eu[1] * p0[1] + eu[2] * p0[2];
    c.cv = ev[0] * p0[0] + ev[1] * p0[1] + ev[2] * p0[2];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        c.NI[i] = C[3 * i] * nn[0] + C[3 * i + 1] * nn[1] + C[3 * i + 2] * nn[2];
        c.EU[i] = C[3 * i] * eu[0] + C[3 * i + 1] * eu[1] + C[3 * i + 2] * eu[2];
        c.EV[i] = C[3 * i] * ev[0] + C[3 * i + 1] * ev[1] + C[3 * i + 2] * ev[2];
    }
    window_and_cone(p, half_u, half_v, c);
    // boresight = third column of C; its (ra, dec) key the lookup grid
    c.dec0 = asinf(fminf(fmaxf(C[8], -1.0f), 1.0f));
    const float a0 = atan2f(C[5], C[2]);
    c.ra0 = a0 < 0.0f ? a0 + 6.28318531f : a0;
    return true;
}

__device__ __forceinline__ float4 lds_f4(uint32_t addr) {            // read-only shared data (catalog)
    float4 v;
    asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {         // survivor list (written by this thread)
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u16(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// fp32 conservative geometric pre-filter on one shared-memory catalog entry (x, y, z, mag); magnitude tested by caller
__device__ __forceinline__ bool prefilter_geom(const Ctx32& c, const float4 s) {
    const float dn = c.NI[0] * s.x + c.NI[1] * s.y + c.NI[2] * s.z;
    float rcp_dn;                                   // bare MUFU.RCP: dn is in (0.02, 1] when it matters (no range fix-ups)
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rcp_dn) : "f"(dn));
    const float rl = c.np0 * rcp_dn;
    const float u = fmaf(rl, c.EU[0] * s.x + c.EU[1] * s.y + c.EU[2] * s.z, -c.cu);
    const float v = fmaf(rl, c.EV[0] * s.x + c.EV[1] * s.y + c.EV[2] * s.z, -c.cv);
    return dn > 0.02f && fabsf(u) <= c.limu && fabsf(v) <= c.limv;
}

// b = (um, vm, f) / |(um, vm, f)|: seed y ~ 1/|.|, one Newton factor w (2^-44 relative, a common scale of one star's b)
__device__ __forceinline__ void body_vector(double um, double vm, double f_ideal, double f_ideal2, double& bx, double& by, double& bz) {
    const double n2 = fma(vm, vm, fma(um, um, f_ideal2));
    const double y = seed_rsqrt(n2);
#if !defined(STMPM_NO_TRIM64)
    const double yw = y * fma(-0.5 * n2 * y, y, 1.5);     // exp X: 9 instead of 11 FP64-pipe instructions, one deeper
    bx = um * yw; by = vm * yw; bz = f_ideal * yw;
#else
    const double w = fma(-0.5 * n2 * y, y, 1.5);           // (um y), (vm y), (f y) run beside the Newton chain
    bx = (um * y) * w; by = (vm * y) * w; bz = (f_ideal * y) * w;
#endif
}
// n.r > 0 (MODEL.md §3).  exp X: on the high word, integer pipe — exact for every dn but a positive denormal below 2^-1022 * 2^-20
// (survivors of the pre-filter have dn > 0.02; in the band walk and the fall-backs dn is a dot product of unit vectors)
__device__ __forceinline__ bool dn_positive(double dn) {
#if !defined(STMPM_NO_TRIM64)
    return __double2hiint(dn) > 0;
#else
    return dn > 0.0;
#endif
}

// Exact (fp64) star evaluation: MODEL.md §2-§4. Returns detected?; on detection adds b s^T into B.
// (x0, x1) = the star's Philox2x32-10 words (MODEL.md §5), computed by the caller.
template <class Tab>
__device__ __forceinline__ bool star_exact(const Ctx64& c, const Tab T, const StarRec& s,
                                           uint32_t x0, uint32_t x1, double f_ideal, double f_ideal2, double half_u, double half_v,
                                           double* __restrict__ B) {
    const double dn = c.NI[0] * s.x + c.NI[1] * s.y + c.NI[2] * s.z;
    const double du = c.EU[0] * s.x + c.EU[1] * s.y + c.EU[2] * s.z;
    const double dv = c.EV[0] * s.x + c.EV[1] * s.y + c.EV[2] * s.z;
    // pre-filter survivors have dn > 0.02 (anything else is not detectable: dn <= 0 fails the test below, and a star
    // with 0 < dn <= 0.02 lands > 50 f from the array centre), so the MUFU-seeded reciprocal is in range without a clamp
#if defined(STMPM_DN_CLAMP)
    const double lam = fast_div(c.np0, fmax(dn, 1e-3));
#else
    const double lam = fast_div(c.np0, dn);      // dn <= 0 gives a negative / infinite / NaN scale: every comparison below is then false
#endif
    double z1, z2;
    box_muller_t<true>(T, x0, x1, z1, z2);
    const double um = fma(c.sx, z1, fma(lam, du, -c.cu));
    const double vm = fma(c.sy, z2, fma(lam, dv, -c.cv));
    const bool det = dn_positive(dn) && (fabs(um) <= half_u) && (fabs(vm) <= half_v);
    if (det) {
        double bx, by, bz;
        body_vector(um, vm, f_ideal, f_ideal2, bx, by, bz);
        B[0] = fma(bx, s.x, B[0]); B[1] = fma(bx, s.y, B[1]); B[2] = fma(bx, s.z, B[2]);
        B[3] = fma(by, s.x, B[3]); B[4] = fma(by, s.y, B[4]); B[5] = fma(by, s.z, B[5]);
        B[6] = fma(bz, s.x, B[6]); B[7] = fma(bz, s.y, B[7]); B[8] = fma(bz, s.z, B[8]);
    }
    return det;
}

// Two stars per star-loop iteration (round 2, exp U; -DSTMPM_ONE_STAR restores the single-star loop): the same arithmetic as
// star_exact, split into an evaluation without side effects (the body vector is computed whether or not the star is detected,
// so that two evaluations interleave as straight-line code) and the accumulation, which the caller applies in star order —
// B's summation order, and with it every record, is unchanged.
template <class Tab>
__device__ __forceinline__ bool star_eval64(const Ctx64& c, const Tab T, const StarRec& s, uint32_t x0, uint32_t x1, double f_ideal,
                                            double f_ideal2, double half_u, double half_v, double& bx, double& by, double& bz) {
    const double dn = c.NI[0] * s.x + c.NI[1] * s.y + c.NI[2] * s.z;
    const double du = c.EU[0] * s.x + c.EU[1] * s.y + c.EU[2] * s.z;
    const double dv = c.EV[0] * s.x + c.EV[1] * s.y + c.EV[2] * s.z;
    const double lam = fast_div(c.np0, dn);
    double z1, z2;
    box_muller_t<true>(T, x0, x1, z1, z2);
    const double um = fma(c.sx, z1, fma(lam, du, -c.cu));
    const double vm = fma(c.sy, z2, fma(lam, dv, -c.cv));
    body_vector(um, vm, f_ideal, f_ideal2, bx, by, bz);
    return dn_positive(dn) && (fabs(um) <= half_u) && (fabs(vm) <= half_v);
}
__device__ __forceinline__ void star_accumulate(double* __restrict__ B, const StarRec& s, double bx, double by, double bz) {
    B[0] = fma(bx, s.x, B[0]); B[1] = fma(bx, s.y, B[1]); B[2] = fma(bx, s.z, B[2]);
    B[3] = fma(by, s.x, B[3]); B[4] = fma(by, s.y, B[4]); B[5] = fma(by, s.z, B[5]);
    B[6] = fma(bz, s.x, B[6]); B[7] = fma(bz, s.y, B[7]); B[8] = fma(bz, s.z, B[8]);
}

// the exact evaluation as an out-of-line call, for fp32 mode's rare guard-band fall-back (keeps that loop's code small)
template <class Tab>
__device__ STMPM_LIBM_INLINE bool star_exact_rare(const Ctx64* c, const Tab T, const StarRec* s, uint32_t x0, uint32_t x1, double f_ideal,
                                                  double f_ideal2, double half_u, double half_v, double* B) {
    double Bl[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) Bl[k] = B[k];
    const bool det = star_exact(*c, T, *s, x0, x1, f_ideal, f_ideal2, half_u, half_v, Bl);
#pragma unroll
    for (int k = 0; k < 9; ++k) B[k] = Bl[k];
    return det;
}

// fp32 star evaluation (precision = 32, bar 1e-6 rad).  What stays fp64 and why: the three
// plane-triad dot products and the projection scale (an fp32 dot product's 6e-8 absolute error times f = 3500 px is 2e-4 px,
// which 3-4-star geometries amplify past the bar through the roll lever arm) and B (roll observability).  What drops to fp32:
// the whole noise draw (Philox -> fp32 Box-Muller: no table polynomials on the FP64 pipe, no Newton steps), entering the
// fp64 centroid through an integer-op float -> double.  The back-projection uses the bare MUFU reciprocal square root (a
// common scale error ~5e-7 of one star's b: a Wahba weight change, ~1e-10 rad).  Detection counts stay bit-exact: a measured
// centroid within `guard` of an array edge re-takes the whole evaluation on the exact fp64 path.
template <class Tab>
__device__ __forceinline__ bool star_f32(const Ctx64& c, const Ctx32& cf, const Tab T,
                                         const StarRec& s, uint32_t x0, uint32_t x1, double f_ideal, double f_ideal2, double half_u,
                                         double half_v, double* __restrict__ B) {
    const double dn = c.NI[0] * s.x + c.NI[1] * s.y + c.NI[2] * s.z;
    const double du = c.EU[0] * s.x + c.EU[1] * s.y + c.EU[2] * s.z;
    const double dv = c.EV[0] * s.x + c.EV[1] * s.y + c.EV[2] * s.z;
#if defined(STMPM_DN_CLAMP)
    const double lam = fast_div(c.np0, fmax(dn, 1e-3));
#else
    const double lam = fast_div(c.np0, dn);
#endif
    float z1, z2;
    box_muller_f32(T, x0, x1, z1, z2);
    const double um = fma(lam, du, -c.cu) + f2d_bits(cf.sxf * z1);
    const double vm = fma(lam, dv, -c.cv) + f2d_bits(cf.syf * z2);
    const double eu = fabs(um) - half_u, ev = fabs(vm) - half_v;       // signed distance past the array edge
    // |dz| <~ 3e-6 -> the fp32 centroid is within 3e-6 sigma_c of the exact one; guard ~ 1e-4 sigma_c (+ 1e-7 px).  The
    // comparisons run on the high words (integer pipe, not the 16-lane FP64 pipe): |x| <= g  <=  hi(|x|) <= hi(g), which
    // over-covers by less than one unit of g's leading word — the guard only has to be wide enough
    const int hu = __double2hiint(eu), hv = __double2hiint(ev);
    if ((hu & 0x7fffffff) <= cf.guard_hi || (hv & 0x7fffffff) <= cf.guard_hi) {      // rare (~1e-7 of the stars): exact decision + exact b
#if defined(STMPM_SMALL_CODE)
        return star_exact_rare(&c, T, &s, x0, x1, f_ideal, f_ideal2, half_u, half_v, B);
#else
        return star_exact(c, T, s, x0, x1, f_ideal, f_ideal2, half_u, half_v, B);
#endif
    }
    const bool det = (__double2hiint(dn) > 0) && ((hu & hv) < 0);      // n.r > 0 and both distances negative (zero is in the guard)
    if (det) {
        const double y = seed_rsqrt(fma(vm, vm, fma(um, um, f_ideal2)));
        const double bx = um * y, by = vm * y, bz = f_ideal * y;
        B[0] = fma(bx, s.x, B[0]); B[1] = fma(bx, s.y, B[1]); B[2] = fma(bx, s.z, B[2]);
        B[3] = fma(by, s.x, B[3]); B[4] = fma(by, s.y, B[4]); B[5] = fma(by, s.z, B[5]);
        B[6] = fma(bz, s.x, B[6]); B[7] = fma(bz, s.y, B[7]); B[8] = fma(bz, s.z, B[8]);
    }
    return det;
}

// star_f32 as a side-effect-free evaluation for the two-star loop: `guard` reports a centroid inside the guard band (the
// caller then takes star_eval64's decision and body vector for that star instead — rare, ~1e-7 of the stars)
template <class Tab>
__device__ __forceinline__ bool star_eval32(const Ctx64& c, const Ctx32& cf, const Tab T, const StarRec& s, uint32_t x0, uint32_t x1,
                                            double f_ideal, double f_ideal2, double half_u, double half_v, double& bx, double& by,
                                            double& bz, bool& guard) {
    const double dn = c.NI[0] * s.x + c.NI[1] * s.y + c.NI[2] * s.z;
    const double du = c.EU[0] * s.x + c.EU[1] * s.y + c.EU[2] * s.z;
    const double dv = c.EV[0] * s.x + c.EV[1] * s.y + c.EV[2] * s.z;
    const double lam = fast_div(c.np0, dn);
    float z1, z2;
    box_muller_f32(T, x0, x1, z1, z2);
    const double um = fma(lam, du, -c.cu) + f2d_bits(cf.sxf * z1);
    const double vm = fma(lam, dv, -c.cv) + f2d_bits(cf.syf * z2);
    const double eu = fabs(um) - half_u, ev = fabs(vm) - half_v;
    const int hu = __double2hiint(eu), hv = __double2hiint(ev);
    guard = (hu & 0x7fffffff) <= cf.guard_hi || (hv & 0x7fffffff) <= cf.guard_hi;
    const double y = seed_rsqrt(fma(vm, vm, fma(um, um, f_ideal2)));
    bx = um * y; by = vm * y; bz = f_ideal * y;
    return (__double2hiint(dn) > 0) && ((hu & hv) < 0);
}

// Wahba solve in the body frame: B' = (B / n) C, QUEST Newton from lambda0 = 1 (Shuster & Oh 1981), error angle
// 2 atan2(|q_v|, |q_w|) of the optimal rotation A_est C (MODEL.md §4).  Returns arcsec.
__device__ __forceinline__ double quest_error_arcsec(const double* __restrict__ Bi, const double* __restrict__ C,
                                                     int n_det) {
    const double w = fast_div(1.0, (double)n_det);      // a common scale of B (6e-14 relative): the quaternion does not depend on it
    double B[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            B[3 * i + j] = w * (Bi[3 * i] * C[j] + Bi[3 * i + 1] * C[3 + j] + Bi[3 * i + 2] * C[6 + j]);
    const double S00 = 2.0 * B[0], S11 = 2.0 * B[4], S22 = 2.0 * B[8];
    const double S01 = B[1] + B[3], S02 = B[2] + B[6], S12 = B[5] + B[7];
    const double sig = B[0] + B[4] + B[8];
    const double Z0 = B[5] - B[7], Z1 = B[6] - B[2], Z2 = B[1] - B[3];
    const double kap = (S11 * S22 - S12 * S12) + (S00 * S22 - S02 * S02) + (S00 * S11 - S01 * S01);
    const double del = S00 * (S11 * S22 - S12 * S12) - S01 * (S01 * S22 - S12 * S02) + S02 * (S01 * S12 - S11 * S02);
    const double SZ0 = S00 * Z0 + S01 * Z1 + S02 * Z2;
    const double SZ1 = S01 * Z0 + S11 * Z1 + S12 * Z2;
    const double SZ2 = S02 * Z0 + S12 * Z1 + S22 * Z2;
    const double a = sig * sig - kap;
    const double b = sig * sig + Z0 * Z0 + Z1 * Z1 + Z2 * Z2;
    const double cc = del + Z0 * SZ0 + Z1 * SZ1 + Z2 * SZ2;
    const double d = SZ0 * SZ0 + SZ1 * SZ1 + SZ2 * SZ2;
    const double apb = a + b, k0 = a * b + cc * sig - d;
    double lam = 1.0;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const double l2 = lam * lam;
        const double f = l2 * l2 - apb * l2 - cc * lam + k0;
        const double fp = 4.0 * l2 * lam - 2.0 * apb * lam - cc;
        lam -= fast_div(f, fp);      // |fp| ~ product of the eigen-gaps: normal and finite whenever a solve makes sense
    }
    const double alpha = lam * lam - sig * sig + kap;
    const double beta = lam - sig;
    const double gamma = (lam + sig) * alpha - del;
    const double X0 = alpha * Z0 + beta * SZ0 + (S00 * SZ0 + S01 * SZ1 + S02 * SZ2);
    const double X1 = alpha * Z1 + beta * SZ1 + (S01 * SZ0 + S11 * SZ1 + S12 * SZ2);
    const double X2 = alpha * Z2 + beta * SZ2 + (S02 * SZ0 + S12 * SZ1 + S22 * SZ2);
    // |X| by a MUFU-seeded reciprocal square root + one Newton step (2^-44 relative: 1e-17 rad on an arcsecond-class error)
    // instead of libm's correctly rounded sqrt with its special-case branches; an exact zero (or a denormal) stays zero
    const double n2 = X0 * X0 + X1 * X1 + X2 * X2;
    const double yr = seed_rsqrt(fmax(n2, 1e-290));
    const double vn = n2 >= 1e-290 ? n2 * (yr * fma(-0.5 * n2 * yr, yr, 1.5)) : 0.0;
    // theta = 2 atan2(|q_v|, |q_w|).  Attitude errors are arcseconds to a few degrees: for t = |q_v| / |q_w| <= 0.03 the
    // odd series to t^9 is exact to 5e-17 relative (next term t^10 / 11) and replaces libm's ~90-instruction atan2; anything
    // larger (or a zero / non-finite gamma: the comparison is then false) takes libm.
    const double ag = fabs(gamma);
    const double t = fast_div(vn, ag);
    if (t <= 0.03) {
        const double t2 = t * t;
        const double p = fma(t2, fma(t2, fma(t2, fma(t2, 1.0 / 9.0, -1.0 / 7.0), 0.2), -1.0 / 3.0), 1.0);
        return (2.0 * RAD2ARCSEC) * (t * p);
    }
    return 2.0 * atan2_libm(vn, ag) * RAD2ARCSEC;
}

// MODEL.md §6: native-RNG parameter draw for one trial.
struct DistDev {
    double mean[11], sigma[11], thermal_coef;
};
__device__ __forceinline__ void sample_params(const DistDev& d, const MathTables* __restrict__ T, uint32_t t_lo,
                                              uint32_t t_hi, uint32_t k0, uint32_t k1, TrialParams& p) {
    uint32_t w[16];
#pragma unroll
    for (int j = 0; j < 4; ++j)
        philox4x32_10(t_lo, t_hi, (uint32_t)j, TAG_PARA, k0, k1, w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
    const double fov = 0.17453292519943295;   // deg2rad(10), monte_carlo.py:87
    const double PI = 3.141592653589793;
    p.ra = fov + (2.0 * PI - 2.0 * fov) * u01(w[0]);
    p.dec = (-0.5 * PI + fov) + (PI - 2.0 * fov) * u01(w[1]);
    p.roll = -PI + 2.0 * PI * u01(w[2]);
    double z[12];
#pragma unroll
    for (int k = 0; k < 6; ++k) box_muller_tab(T, w[4 + 2 * k], w[5 + 2 * k], z[2 * k], z[2 * k + 1]);
    const double F = d.mean[0] + d.sigma[0] * z[0];
    p.ex = d.mean[1] + d.sigma[1] * z[1];
    p.ey = d.mean[2] + d.sigma[2] * z[2];
    p.ez = d.mean[3] + d.sigma[3] * z[3];
    p.phi = d.mean[4] + d.sigma[4] * z[4];
    p.theta = d.mean[5] + d.sigma[5] * z[5];
    p.psi = d.mean[6] + d.sigma[6] * z[6];
    p.devx = d.mean[7] + d.sigma[7] * z[7];
    p.devy = d.mean[8] + d.sigma[8] * z[8];
    p.maxmag = d.mean[9] + d.sigma[9] * z[9];
    const double dtemp = d.mean[10] + d.sigma[10] * z[10];
    p.f = F * (1.0 + dtemp * d.thermal_coef);   // sensitivity.py:115-117
}
// ra, dec, MAX_MAGNITUDE of a native-RNG trial only (the work key needs nothing else): Philox blocks 0 and 3
__device__ __forceinline__ void sample_key_params(const DistDev& d, const MathTables* __restrict__ T, uint32_t t_lo,
                                                  uint32_t t_hi, uint32_t k0, uint32_t k1, double& ra, double& dec,
                                                  double& roll, double& maxmag) {
    uint32_t w0, w1, w2, w3, v0, v1, v2, v3;
    philox4x32_10(t_lo, t_hi, 0u, TAG_PARA, k0, k1, w0, w1, w2, w3);
    philox4x32_10(t_lo, t_hi, 3u, TAG_PARA, k0, k1, v0, v1, v2, v3);
    const double fov = 0.17453292519943295, PI = 3.141592653589793;
    ra = fov + (2.0 * PI - 2.0 * fov) * u01(w0);
    dec = (-0.5 * PI + fov) + (PI - 2.0 * fov) * u01(w1);
    roll = -PI + 2.0 * PI * u01(w2);
    double z8, z9;
    box_muller_tab(T, v0, v1, z8, z9);      // words 12, 13 -> normals 8, 9
    maxmag = d.mean[9] + d.sigma[9] * z9;
}
#endif  // __CUDACC__

}  // namespace stmpm
