// stmpm_split.cuh — the two-kernel trial pipeline (round 2).  Included by stmpm.cu after the single-kernel path.
//
// Why: the single persistent kernel runs the candidate scan (fp32, ~50 registers of state) and the exact star loop (fp64,
// 128 registers) in the same 512 threads, and its 32 lanes run each loop for the longest of 32 different trip counts
// (lanes 25.8 / 32; the work key only PREDICTS the counts).  Splitting the trial at the hand-over point gives each half the
// launch shape it wants and makes the second half's trip counts EXACTLY uniform:
//   scan kernel   1 024 threads per SM (64 registers): fp32 set-up + candidate scan, trials in input order (coalesced
//                 column reads); writes one 224-byte record per trial — the 13 parameters, the survivor count, the
//                 survivors' catalog indices.
//                 Every block of 1 024 trials is counting-sorted by survivor count in shared memory (a permutation, written
//                 next to the records).
//   solve kernel  512 threads per SM: fp64 set-up, star loop, QUEST solve, records + statistics — over the permutation, so the
//                 32 trials of a warp have (nearly) the SAME number of stars: no idle lanes, no scheduler, no shared-memory
//                 catalog.
// A trial the scan cannot hand over (scan cone wider than the candidate lists, more than REC_CAP survivors, angles beyond
// the fp32 set-up's range) is marked CNT_SLOW and evaluated in full by the solve kernel on a sky-grid walk through global
// memory; when more than 1/64 of a chunk is marked, the chunk is given to the single-kernel path instead (device-side
// decision, no host synchronisation).  Results are the single-kernel path's bit for bit: the same star set reaches the
// same star_exact / star_f32 / quest_error_arcsec, only in another order (B is a sum: the order of its terms is the
// catalog-list order in both paths).
#pragma once

constexpr int TPB_SCAN = 1024;
constexpr int REC_CAP = 56;            // survivor indices a record holds
constexpr uint32_t CNT_SLOW = 255;     // record not handed over: full evaluation in the solve kernel
constexpr int SORT_BINS = 256;         // survivor counts 0 .. 255 (255 = CNT_SLOW sorts last)

struct __align__(32) TrialRec {        // 224 bytes = seven 32-byte sectors
    double p[STMPM_NCOLS];             // the trial's parameters (MODEL.md §1)
    uint32_t cnt;                      // survivors in list[], or CNT_SLOW
    uint32_t index;                    // the trial's index inside the chunk
    uint16_t list[REC_CAP];            // catalog indices of the pre-filter survivors, in candidate-list order
};
static_assert(sizeof(TrialRec) == 224, "TrialRec layout");

struct SplitCtl {                      // device words of one chunk
    unsigned int n_slow;               // trials the scan could not hand over
    unsigned int next_group;           // solve kernel: next group of 32 sorted trials to hand out
    unsigned int pad0[2];
    int skip_split;                    // 1: too many CNT_SLOW trials — the solve kernel returns, the single kernel runs
    int skip_single;                   // 1: the solve kernel runs, the single kernel returns
    int pad[2];
};

// ------------------------------------------------------------------------------------------------ scan kernel
__global__ void __launch_bounds__(TPB_SCAN, 1) split_scan_kernel(const KArgs a, TrialRec* __restrict__ rec, uint32_t* __restrict__ perm,
                                                                long long n, SplitCtl* ctl) {
    extern __shared__ __align__(128) unsigned char smem[];
    float4* s_cat = reinterpret_cast<float4*>(smem + OFF_CAT);
    int2* s_lkband = reinterpret_cast<int2*>(smem + OFF_LKBAND);
    const uint32_t sb_cat = smem_u32(smem) + (uint32_t)OFF_CAT;
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + OFF_MBAR);
    if (threadIdx.x == 0) mbar_init(s_bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {            // the fp32 catalog through the TMA unit, as in trials_kernel
        const uint32_t total = (uint32_t)(a.n_stars + 1) * 16u;
        mbar_expect_tx(s_bar, total);
        const unsigned char* src = reinterpret_cast<const unsigned char*>(a.cat32);
        unsigned char* dst = reinterpret_cast<unsigned char*>(s_cat);
        for (uint32_t off = 0; off < total; off += 32768u) bulk_g2s(dst + off, src + off, min(32768u, total - off), s_bar);
    }
    for (int i = threadIdx.x; i < a.lk_bands; i += TPB_SCAN) s_lkband[i] = __ldg(a.lk_band_tab + i);
    while (!mbar_try_wait(s_bar, 0)) {}
    __syncthreads();

    // block-local counting sort of 1 024 survivor counts: bins, then exclusive offsets, in shared memory
    unsigned int* s_bins = reinterpret_cast<unsigned int*>(smem + OFF_LIST);              // [SORT_BINS] (the list region is unused here)
    const float hu32 = (float)a.half_u, hv32 = (float)a.half_v;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned int n_slow = 0;
    const long long n_blocks = (n + TPB_SCAN - 1) / TPB_SCAN;
    for (long long blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
        const long long base = blk * TPB_SCAN;
        const long long i = base + threadIdx.x;
        const bool valid = i < n;
        if (threadIdx.x < SORT_BINS) s_bins[threadIdx.x] = 0u;
        __syncthreads();
        uint32_t cnt = CNT_SLOW;
        if (valid) {
            TrialParams p;
            p.ra = __ldcs(a.cols[0] + i);     p.dec = __ldcs(a.cols[1] + i);    p.roll = __ldcs(a.cols[2] + i);
            p.f = __ldcs(a.cols[3] + i);      p.ex = __ldcs(a.cols[4] + i);     p.ey = __ldcs(a.cols[5] + i);
            p.ez = __ldcs(a.cols[6] + i);     p.phi = __ldcs(a.cols[7] + i);    p.theta = __ldcs(a.cols[8] + i);
            p.psi = __ldcs(a.cols[9] + i);    p.devx = __ldcs(a.cols[10] + i);  p.devy = __ldcs(a.cols[11] + i);
            p.maxmag = __ldcs(a.cols[12] + i);
            TrialRec* r = rec + i;
            {   // the parameters travel with the record (the solve kernel reads ONE place per trial): 16-byte stores
                double2* rp = reinterpret_cast<double2*>(r->p);
                __stcg(rp + 0, make_double2(p.ra, p.dec));      __stcg(rp + 1, make_double2(p.roll, p.f));
                __stcg(rp + 2, make_double2(p.ex, p.ey));       __stcg(rp + 3, make_double2(p.ez, p.phi));
                __stcg(rp + 4, make_double2(p.theta, p.psi));   __stcg(rp + 5, make_double2(p.devx, p.devy));
            }
            Ctx32 cf;
            const bool ok = setup32(p, hu32, hv32, cf) && cf.cone_t <= a.lk_cone_tan;
            if (ok) {
                const uint16_t* __restrict__ lp =
                    a.lk_list + __ldg(a.lk_off + lookup_cell(s_lkband, a.lk_bands, a.lk_inv_band_h, cf.ra0, cf.dec0));
                uint4* __restrict__ rl = reinterpret_cast<uint4*>(r->list);
                cnt = 0;
                uint32_t buf[4] = {0u, 0u, 0u, 0u};       // eight survivor indices, flushed as one 16-byte store
                bool more = true;
                while (more) {             // eight candidates per 16-byte pack; the +inf terminator star ends every list
                    const uint4 pk = __ldg(reinterpret_cast<const uint4*>(lp));
                    lp += PACK_N;
                    const uint32_t w4[4] = {pk.x, pk.y, pk.z, pk.w};
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const uint32_t idx = (w4[e >> 1] >> (16 * (e & 1))) & 0xffffu;
                        const float4 sc = lds_f4(sb_cat + idx * 16u);
                        more = sc.w <= cf.magf;
                        if (more && prefilter_geom(cf, sc)) {
                            const uint32_t k = cnt & 7u;
                            // slot k of the buffer (k is not a compile-time constant: select, do not index)
                            const uint32_t sh = idx << (16u * (k & 1u));
                            const uint32_t word = k >> 1;
                            buf[0] = word == 0 ? ((k & 1u) ? (buf[0] | sh) : sh) : buf[0];
                            buf[1] = word == 1 ? ((k & 1u) ? (buf[1] | sh) : sh) : buf[1];
                            buf[2] = word == 2 ? ((k & 1u) ? (buf[2] | sh) : sh) : buf[2];
                            buf[3] = word == 3 ? ((k & 1u) ? (buf[3] | sh) : sh) : buf[3];
                            ++cnt;
                            if ((cnt & 7u) == 0u && cnt <= (uint32_t)REC_CAP) __stcg(rl + (cnt >> 3) - 1, make_uint4(buf[0], buf[1], buf[2], buf[3]));
                        }
                    }
                }
                if ((cnt & 7u) != 0u && cnt < (uint32_t)REC_CAP) __stcg(rl + (cnt >> 3), make_uint4(buf[0], buf[1], buf[2], buf[3]));
                if (cnt > (uint32_t)REC_CAP) cnt = CNT_SLOW;
            }
            // last 16 bytes of the parameter block: MAX_MAGNITUDE, the survivor count, the trial's index inside the chunk
            __stcg(reinterpret_cast<uint4*>(r->p + 12),
                   make_uint4((uint32_t)__double2loint(p.maxmag), (uint32_t)__double2hiint(p.maxmag), cnt, (uint32_t)i));
            if (cnt == CNT_SLOW) ++n_slow;
        }
        // ---- counting sort of the block by cnt: rank = offset[cnt] + arrival order inside the bin
        unsigned int slot = 0;
        if (valid) slot = atomicAdd(&s_bins[cnt], 1u);
        __syncthreads();
        if (warp == 0) {                   // exclusive scan of 256 bins by one warp: 8 bins per lane
            unsigned int v[8], tot = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) { v[k] = s_bins[lane * 8 + k]; tot += v[k]; }
            unsigned int incl = tot;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const unsigned int o = __shfl_up_sync(0xffffffffu, incl, off);
                if (lane >= off) incl += o;
            }
            unsigned int run = incl - tot;
#pragma unroll
            for (int k = 0; k < 8; ++k) { s_bins[lane * 8 + k] = run; run += v[k]; }
        }
        __syncthreads();
        if (valid) perm[base + s_bins[cnt] + slot] = (uint32_t)i;
        __syncthreads();                   // s_bins is cleared at the top of the next block
    }
    // trials not handed over, for the fall-back decision
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) n_slow += __shfl_xor_sync(0xffffffffu, n_slow, off);
    if (lane == 0 && n_slow) atomicAdd(&ctl->n_slow, n_slow);
}

// decides, after the scan of a chunk, which kernel evaluates it
__global__ void split_decide_kernel(SplitCtl* ctl, long long n, int* fallbacks_seen) {
    const bool too_many = (long long)ctl->n_slow * 64 > n;    // more than 1/64 of the chunk cannot be handed over
    ctl->skip_split = too_many ? 1 : 0;
    ctl->skip_single = too_many ? 0 : 1;
    if (too_many && fallbacks_seen != nullptr) *reinterpret_cast<volatile int*>(fallbacks_seen) = 1;   // host: stop trying the split here
}

// ------------------------------------------------------------------------------------------------ solve kernel
// full evaluation of a trial the scan could not hand over: conservative fp32 walk over the sky-grid bands that overlap the
// scan cone, reading the fp32 catalog from global memory; every survivor goes straight to the exact evaluation
template <int PREC, class Tab>
__device__ __noinline__ void slow_trial_global(const KArgs* ap, const Ctx64* cp, const Ctx32* cfp, const Tab tabs, uint32_t ka, uint32_t kb,
                                               double* B_out, int* n_det_out) {
    const KArgs& a = *ap;
    const Ctx64 c = *cp;
    Ctx32 cf = *cfp;
    cf.cone = cone_angle(cf);
    double B[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    int n_det = 0;
    const int b_lo = max(0, (int)floorf((cf.dec0 - cf.cone + 1.57079633f) * a.inv_band_h));
    const int b_hi = min(a.n_bands - 1, (int)floorf((cf.dec0 + cf.cone + 1.57079633f) * a.inv_band_h));
    float dalpha = -1.0f;                           // < 0: the cone touches a pole, take whole bands
    if (fabsf(cf.dec0) + cf.cone < 1.5607963f) dalpha = asinf(fminf(sinf(cf.cone) / cosf(cf.dec0), 1.0f)) + 0.002f;
    for (int b = b_lo; b <= b_hi; ++b) {
        const int2 bt = __ldg(a.band_tab + b);
        const int base = bt.x, nra = bt.y;
        int seg_lo[2] = {0, 0}, seg_hi[2] = {0, 0};
        seg_lo[0] = a.cell_start[base]; seg_hi[0] = a.cell_start[base + nra];
        if (dalpha >= 0.0f) {
            const float inv_w = (float)nra * 0.159154943f;
            const int c_lo = (int)floorf((cf.ra0 - dalpha) * inv_w), c_hi = (int)floorf((cf.ra0 + dalpha) * inv_w);
            const int count = c_hi - c_lo + 1;
            if (count < nra) {
                int lo = c_lo % nra;
                if (lo < 0) lo += nra;
                if (lo + count <= nra) { seg_lo[0] = a.cell_start[base + lo]; seg_hi[0] = a.cell_start[base + lo + count]; }
                else {
                    seg_lo[0] = a.cell_start[base + lo]; seg_hi[0] = a.cell_start[base + nra];
                    seg_lo[1] = a.cell_start[base]; seg_hi[1] = a.cell_start[base + lo + count - nra];
                }
            }
        }
        for (int sgm = 0; sgm < 2; ++sgm)
            for (int i = seg_lo[sgm]; i < seg_hi[sgm]; ++i) {
                const float4 sc = __ldg(a.cat32 + i);
                if (sc.w <= cf.magf && prefilter_geom(cf, sc)) {
                    const double4 rr = ld_rec(a.cat64 + i);
                    StarRec s;
                    s.x = rr.x; s.y = rr.y; s.z = rr.z;
                    s.star_id = (uint32_t)(__double_as_longlong(rr.w) & 0xffffffffll);
                    uint32_t x0, x1;
                    philox2x32_10(s.star_id, kb, ka, x0, x1);
                    bool det;
                    if (PREC == 64) det = star_exact(c, tabs, s, x0, x1, a.f_ideal, a.f_ideal2, a.half_u, a.half_v, B);
                    else det = star_f32(c, cf, tabs, s, x0, x1, a.f_ideal, a.f_ideal2, a.half_u, a.half_v, B);
                    n_det += det ? 1 : 0;
                }
            }
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) B_out[k] = B[k];
    *n_det_out = n_det;
}

__device__ __forceinline__ uint32_t ldg_u16(const uint16_t* p) {
    unsigned short v;
    asm("ld.global.nc.u16 %0, [%1];" : "=h"(v) : "l"(p));
    return (uint32_t)v;
}

template <int PREC>
__global__ void __launch_bounds__(TPB, 1) split_solve_kernel(const KArgs a, const TrialRec* __restrict__ rec,
                                                             const uint32_t* __restrict__ perm, long long n, SplitCtl* ctl) {
    __shared__ __align__(16) unsigned char s_tab_raw[sizeof(MathTables)];
    if (*reinterpret_cast<const volatile int*>(&ctl->skip_split) != 0) return;
    {
        const double* gt = reinterpret_cast<const double*>(a.tables);
        double* stb = reinterpret_cast<double*>(s_tab_raw);
        for (int i = threadIdx.x; i < (int)(sizeof(MathTables) / 8); i += TPB) stb[i] = __ldg(gt + i);
    }
    __syncthreads();
    const TabSmem tabs{smem_u32(s_tab_raw)};
    const int lane = threadIdx.x & 31;
    const float hu32 = (float)a.half_u, hv32 = (float)a.half_v;
    const bool want_stats = a.stats != nullptr;
    unsigned int acc_ok = 0, acc_fail = 0;
    unsigned long long acc_ndet = 0;
    double acc_x = 0.0, acc_xx = 0.0, acc_max = 0.0;

    // groups are handed out dynamically: inside every 1 024-trial block they run from few to many stars, so any static
    // stride that is a multiple of 32 would give a warp the same position — the same cost — in every block
    const long long n_groups = (n + 31) >> 5;
    for (;;) {
        unsigned int g32 = 0;
        if (lane == 0) g32 = atomicAdd(&ctl->next_group, 1u);
        const long long g = (long long)__shfl_sync(0xffffffffu, g32, 0);
        if (g >= n_groups) break;
        const long long q = (g << 5) + lane;
        const bool valid = q < n;
        const long long i = valid ? (long long)__ldg(perm + q) : 0;
        const TrialRec* __restrict__ r = rec + i;
        TrialParams p;
        uint32_t cnt = 0;
        if (valid) {
            const double2* rp = reinterpret_cast<const double2*>(r->p);
            const double2 v0 = __ldcs(rp), v1 = __ldcs(rp + 1), v2 = __ldcs(rp + 2), v3 = __ldcs(rp + 3), v4 = __ldcs(rp + 4),
                          v5 = __ldcs(rp + 5);
            p.ra = v0.x; p.dec = v0.y; p.roll = v1.x; p.f = v1.y; p.ex = v2.x; p.ey = v2.y; p.ez = v3.x; p.phi = v3.y;
            p.theta = v4.x; p.psi = v4.y; p.devx = v5.x; p.devy = v5.y;
            const uint4 tail = __ldcs(reinterpret_cast<const uint4*>(r->p + 12));
            p.maxmag = __hiloint2double((int)tail.y, (int)tail.x);
            cnt = tail.z;
        } else {
            p.ra = p.dec = p.roll = 0.0; p.f = 1000.0; p.ex = p.ey = p.ez = p.phi = p.theta = p.psi = 0.0;
            p.devx = p.devy = 0.0; p.maxmag = -100.0;
        }
        const unsigned long long gidx = (unsigned long long)(a.trial0 + i);
        Ctx64 c;
        setup64(p, c);
        Ctx32 cf;
        window_and_cone(p, hu32, hv32, cf);            // fp32 mode reads sxf / syf / guard_hi; the slow path the cone
        uint32_t ka, kb;
        star_key((uint32_t)gidx, (uint32_t)(gidx >> 32), a.k0, a.k1, ka, kb);
#if !defined(STMPM_NO_PARK)
        volatile double c_park[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) c_park[k] = c.C[k];
#endif
        double B[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        int n_det = 0;
        if (cnt != CNT_SLOW) {
            // every lane of the warp has (nearly always exactly) the same cnt: the loop is uniform
            const uint16_t* __restrict__ rl = r->list;
            const int m = (int)cnt;
            uint32_t i_next = m > 1 ? ldg_u16(rl + 1) : 0u;
            double4 rr = make_double4(0.0, 0.0, 0.0, 0.0);
            if (m > 0) rr = ld_rec(a.cat64 + ldg_u16(rl));
            for (int j = 0; j < m; ++j) {
                StarRec s;
                s.x = rr.x; s.y = rr.y; s.z = rr.z;
                s.star_id = (uint32_t)(__double_as_longlong(rr.w) & 0xffffffffll);
                const bool mag_ok = __int_as_float(__double2hiint(rr.w)) <= cf.magf;   // see trials_kernel: also keeps the word alive
                if (j + 1 < m) rr = ld_rec(a.cat64 + i_next);
                if (j + 2 < m) i_next = ldg_u16(rl + j + 2);
                uint32_t x0, x1;
                philox2x32_10(s.star_id, kb, ka, x0, x1);
                bool det;
                if (PREC == 64) det = star_exact(c, tabs, s, x0, x1, a.f_ideal, a.f_ideal2, a.half_u, a.half_v, B);
                else det = star_f32(c, cf, tabs, s, x0, x1, a.f_ideal, a.f_ideal2, a.half_u, a.half_v, B);
                n_det += (det && mag_ok) ? 1 : 0;
            }
        }
#if !defined(STMPM_NO_PARK)
#pragma unroll
        for (int k = 0; k < 9; ++k) c.C[k] = c_park[k];
#endif
        if (cnt == CNT_SLOW && valid) {
            Ctx64 c2 = c;
            Ctx32 cf2 = cf;
            geom32(c, cf2);
            boresight32(c, cf2);
            double B2[9];
            int nd2 = 0;
            slow_trial_global<PREC>(&a, &c2, &cf2, tabs, ka, kb, B2, &nd2);
#pragma unroll
            for (int k = 0; k < 9; ++k) B[k] = B2[k];
            n_det = nd2;
        }
        double err = -1.0;
        if (n_det >= a.n_min) err = quest_error_arcsec(B, c.C, n_det);
        if (valid) {
            if (a.err) a.err[i] = err;
            if (a.ndet) a.ndet[i] = n_det;
            if (a.ndet8) a.ndet8[i] = (uint8_t)min(n_det, 255);
            if (a.err32) a.err32[i] = (float)err;
            if (want_stats) {
                acc_ndet += (unsigned long long)n_det;
                if (err >= 0.0) {
                    ++acc_ok;
                    acc_x += err;
                    acc_xx = fma(err, err, acc_xx);
                    acc_max = fmax(acc_max, err);
                    const long long e = (__double_as_longlong(err) >> a.hist_shift)
                                        - ((long long)(1023 + STMPM_HIST_MIN_EXP) << (52 - a.hist_shift));
                    const int slot = e < 0 ? 6 : (e >= a.n_bins ? 7 : STMPM_STATS_HEAD + (int)e);
                    red_add_f64(a.stats + slot, 1.0);
                } else {
                    ++acc_fail;
                }
            }
        }
    }
    if (want_stats) flush_moments(a.stats, lane, acc_ok, acc_fail, acc_ndet, acc_x, acc_xx, acc_max);
}
