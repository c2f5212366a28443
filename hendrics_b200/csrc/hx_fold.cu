// K1: exact (f, fdot) trial-grid fold search for sm_100a.
//
// Replaces the hot loop of stingray.pulse.search._folding_search (reached from
// hendrics/efsearch.py:1179-1187 through z_n_search / epoch_folding_search):
//   per trial:  _pulse_phase_fast -> _profile_fast -> ef_profile_stat | z_n_binned_events.
//
// Design ("K1"):
//   * lane = trial.  Every thread owns R trials and, for each, a PRIVATE nbin-bin uint32 histogram
//     column in shared memory laid out [bin][lane] (word address = bin*32 + lane), so that every
//     histogram access of a warp hits 32 distinct banks: no atomics, no bank conflicts.
//   * event times (8 B each, nothing else) stream HBM/L2 -> shared memory through a 2-stage
//     cp.async.bulk (TMA engine) + mbarrier pipeline and are broadcast to all threads, two events per
//     LDS.128.  There is no block-wide barrier in the loop: the last warp to finish a stage refills it.
//   * phase arithmetic is the reference's, rounding for rounding: t*f and (0.5*t*t)*fd are separately
//     rounded products, added with one rounding -- never an FMA.  (0.5*t*t)*fd is evaluated as
//     fl(t*t) * (0.5*fd): scaling by 0.5 is exact, so both orders round the same real number.  For
//     power-of-two nbin the trial constants are pre-scaled by nbin*128 (exact), floor() is one
//     round-down add of 1.5*2^52 and the bin byte offset is one AND of the low mantissa word; it is
//     OR-ed into the (aligned) column base.  Per event*trial: 2 DMUL + 2 DADD + LOP3 + ATOMS, plus one
//     DMUL (t*t) per event per thread amortised over its R trials (f only: DMUL + DADD + LOP3 + ATOMS).
//   * the increment is a fire-and-forget `red.shared.add.u32` on the thread's private column
//     (measured on B200: 26 conflict-free lanes/clk/SM vs 15.8 for LDS+IADD+STS, one issue slot
//     instead of three and no load->add->store dependency chain); the next group of events is
//     prefetched into registers before the current group's increments are issued.
//   * nbin >= 128 (power of two): PACKED counters.  Bins j and j + nbin/2 share one 32-bit word: the
//     low 6+log2(nbin) bits count every event of the pair, the high bits count those of the upper
//     bin, so the increment is (low_word & top_bin_bit) | 1 -- one LOP3, no shift -- and the word
//     offset is the usual AND.  This halves the histogram footprint (twice the resident warps for the
//     default nbin = 128); words are drained into the global scratch every 2^(6+log2 nbin) - 1 events
//     at most, before the low field can carry into the high one.
//   * the statistic epilogue (chi^2 or binned Z^2_n) is fused when a CTA sees all events of its
//     trials; otherwise counts are reduced into a [trial/32][bin][32] global scratch with coalesced
//     RED.ADD and a second small kernel computes the statistics.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <mutex>
#include <type_traits>
#include <vector>

#include "hx_common.cuh"

namespace hx {

#ifndef HX_FOLD_SKEW_FD
#define HX_FOLD_SKEW_FD 0  // 1: skewed software pipeline for the (f, fdot) kernels too (measured on B200: no gain, C3 3.21e12 either way)
#endif
#ifndef HX_FOLD_HYB
#define HX_FOLD_HYB 0  // M2T experiment: of every group of UE events, the fdot term of the last HX_FOLD_HYB is recomputed per thread
#endif                 // (2 FP64 ops per event and thread) instead of loaded from the tile (one broadcast wavefront per event and warp)
#ifndef HX_FOLD_UE_FD
#define HX_FOLD_UE_FD 8  // events per register group of the plain (f, fdot) loop with <= 2 trials per thread (4 and 16 measured: see DESIGN)
#endif
#ifndef HX_FOLD_SKEW_UE2
#define HX_FOLD_SKEW_UE2 4  // events per register group of the skewed (f, fdot) loop with 2 trials per thread
#endif

constexpr int TILE_EV_MAX = 1024;  // events per pipeline stage (runtime: 256..1024, multiple of 64)
constexpr int TILE_PAD = 16;       // events of slack after each stage (register prefetch may read past cnt)
constexpr uint32_t SMEM_HDR = 32;  // two mbarriers + two "warps done" counters

struct FoldArgs {
    const double *ev;  // event times
    long long n_ev;
    const double *freqs;
    long long n_f;
    const double *fdots;
    long long n_trials;
    int nbin;
    int nharm;
    uint32_t binmask;  // POW2: (nbin-1) << 7 ; PACKED: (nbin/2-1) << 7
    uint32_t topbit;   // PACKED: 1 << (6 + log2 nbin), the bit of the scaled phase that selects the upper bin
    int pack_shift;    // PACKED: 6 + log2 nbin
    int flush_tiles;   // PACKED: drain the packed words every this many stages
    long long tb0;     // first 32*R-trial block of this launch (trial batching)
    long long n_tb_row;  // ROWFD: 32*R-trial blocks per fdot row (rows padded to a multiple of 32*R)
    int tile_ev;       // events per pipeline stage
    double scale;      // POW2: nbin*128, generic: 1
    double gen_tmin;   // generic: events at t >= gen_tmin take the short exact path (phase >= 512 cycles); +inf: none
    double *stats;
    uint32_t *scratch;  // null => fused epilogue
    const double *tab;  // [2*nbin] cos, sin
    int warps;
    uint32_t hist_off, tile_off;
    long long ev_per_split;
    int *errflag;
};

enum { MODE_GENERIC = 0, MODE_POW2 = 1, MODE_PACKED = 2, MODE_EXACT = 3 };

// floor(x) for |x| < 2^51, exact
__device__ __forceinline__ double floor_magic(double x) { return __dsub_rn(__dadd_rd(x, HX_MAGIC), HX_MAGIC); }

// MODE_EXACT: phase mod 1 accurate to ~1 ulp of 1.0 whatever the magnitude of t*f (the reference's
// fp64 phase loses log2(t*f) bits).  Error-free products (p + e == a*b exactly, via FMA) are
// reduced mod 1 term by term before they are added.  NOT the parity path: it intentionally differs
// from the reference where fp64 rounding moved a photon across a bin edge.
__device__ __forceinline__ double exact_phase_frac(double t, double tt, double tte, double f, double h, bool has_fd) {
    const double p = __dmul_rn(t, f);
    const double e = __fma_rn(t, f, -p);  // t*f == p + e
    double s = __dsub_rn(p, floor_magic(p));  // exact, in [0, 1)
    double small = e;
    if (has_fd) {
        const double q = __dmul_rn(tt, h);    // h = 0.5*fd ; t*t == tt + tte
        const double qe = __fma_rn(tt, h, -q);
        const double q2 = __dmul_rn(tte, h);
        s = __dadd_rn(s, __dsub_rn(q, floor_magic(q)));
        small = __dadd_rn(__dadd_rn(e, qe), q2);
    }
    s = __dadd_rn(s, small);
    return __dsub_rn(s, floor_magic(s));  // back into [0, 1]
}

// tt = fl(t*t); fdH = 0.5*fd*scale (exact scalings), so fl(tt*fdH) == scale * fl(((0.5*t)*t)*fd)
// Returns the shared-memory word address of the bin; `inc` is the value to add (1 unless PACKED).
// `m2` is the already rounded fdot term fl(tt*fdH) (computed once per event and thread when the R trials
// of a thread share their fdot, ROWFD, else per trial).
// GF (MODE_GENERIC only): the event has a phase of at least 512 cycles for every trial of the launch (t >= gen_tmin, decided
// per stage) and takes the short exact path below.
template <bool HAS_FD, int MODE, bool GF = false>
__device__ __forceinline__ uint32_t bin_addr(double t, double tt, double tte, double m2, double fN, double fdH,
                                             uint32_t mask, uint32_t topbit, int nbin, uint32_t cb, uint32_t &inc) {
    if (MODE == MODE_EXACT) {
        inc = 1u;
        const double fr = exact_phase_frac(t, tt, tte, fN, fdH, HAS_FD);
        int b = __double2loint(__dadd_rd(__dmul_rn(fr, (double)nbin), HX_MAGIC));
        b = min(b, nbin - 1);
        return cb + ((uint32_t)b << 7);
    }
    double x = __dmul_rn(t, fN);
    if (HAS_FD) x = __dadd_rn(x, m2);
    inc = 1u;
    if (MODE == MODE_POW2 || MODE == MODE_PACKED) {
        // x = nbin*128*phase exactly; floor -> low word; bits [7, 7+log2 nbin) are the bin
        double y = __dadd_rd(x, HX_MAGIC);
        uint32_t lo = (uint32_t)__double2loint(y);
        if (MODE == MODE_PACKED) inc = (lo & topbit) | 1u;
        return (lo & mask) | cb;
    } else {
        if (GF) {
            // Any nbin, x >= 512 cycles.  The reference computes fr = x - floor(x) (exact) and floor(fl(fr * nbin)).  fr is a
            // multiple of ulp(x) >= 2^-43 below 1, so fr * nbin (nbin < 2^10) has at most 53 significant bits: the product
            // is exact and the reference's bin is floor(fr * nbin) = floor(x * nbin) - floor(x) * nbin in exact arithmetic.
            // floor(x * nbin) is taken from the product rounded DOWN (rounding down never crosses an integer, integers
            // being representable), so: 3 FP64 ops + 2 IMAD instead of 5 FP64 ops, no clamp needed (fr * nbin < nbin).
            const uint32_t lp = (uint32_t)__double2loint(__dadd_rd(__dmul_rd(x, (double)nbin), HX_MAGIC));
            const uint32_t lq = (uint32_t)__double2loint(__dadd_rd(x, HX_MAGIC));
            return cb + ((lp - lq * (uint32_t)nbin) << 7);  // mod 2^32 arithmetic; the true difference is in [0, nbin)
        }
        double fl = __dsub_rn(__dadd_rd(x, HX_MAGIC), HX_MAGIC);  // floor(x), exact
        double fr = __dsub_rn(x, fl);                              // ph -= floor(ph)
        double z = __dmul_rn(fr, (double)nbin);
        int b = __double2loint(__dadd_rd(z, HX_MAGIC));  // floor(z), 0 <= z <= nbin
        b = min(b, nbin - 1);                            // A4 clamp
        return cb + ((uint32_t)b << 7);
    }
}

__device__ __forceinline__ void red_shared_add(uint32_t addr, uint32_t v) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(v));
}

template <bool HAS_FD, int UE, bool EXACT>
struct EvGroup {
    double t[UE], tt[HAS_FD ? UE : 1], tte[(EXACT && HAS_FD) ? UE : 1], m2[HAS_FD ? UE : 1];
    __device__ __forceinline__ void load(const unsigned char *tp, int e) {
#pragma unroll
        for (int u = 0; u < UE; u += 2) {
            double2 v = *reinterpret_cast<const double2 *>(tp + (size_t)(e + u) * 8);
            t[u] = v.x, t[u + 1] = v.y;
        }
    }
    // M2T: the fdot terms fl(fl(t*t) * fdH) of the stage were computed once per CTA into a second tile
    __device__ __forceinline__ void load_m2(const unsigned char *mp, int e) {
        if (HAS_FD) {
#pragma unroll
            for (int u = 0; u < UE; u += 2) {
                double2 v = *reinterpret_cast<const double2 *>(mp + (size_t)(e + u) * 8);
                m2[u] = v.x, m2[u + 1] = v.y;
            }
        }
    }
    // hybrid M2T: load the terms of the first n_load events of the group, recompute the others
    __device__ __forceinline__ void load_m2_n(const unsigned char *mp, int e, int n_load) {
        if (HAS_FD) {
#pragma unroll
            for (int u = 0; u < UE; u += 2)
                if (u < n_load) {
                    double2 v = *reinterpret_cast<const double2 *>(mp + (size_t)(e + u) * 8);
                    m2[u] = v.x, m2[u + 1] = v.y;
                }
        }
    }
    __device__ __forceinline__ void squares_from(double fdH_row, int from) {
        if (HAS_FD) {
#pragma unroll
            for (int u = 0; u < UE; u++)
                if (u >= from) m2[u] = __dmul_rn(__dmul_rn(t[u], t[u]), fdH_row);
        }
    }
    // fdH_row: the thread's fdot constant when all its trials share it (ROWFD), else unused
    __device__ __forceinline__ void squares(double fdH_row = 0.0) {
        if (HAS_FD) {
#pragma unroll
            for (int u = 0; u < UE; u++) {
                tt[u] = __dmul_rn(t[u], t[u]);
                if (EXACT) tte[u] = __fma_rn(t[u], t[u], -tt[u]);
                m2[u] = __dmul_rn(tt[u], fdH_row);
            }
        }
    }
};

// events per register group: enough independent DMUL->DADD->DADD.RM->LOP3->ATOMS chains per warp
// (R*UE) to cover the FP64 latency with the few warps the histograms leave room for
template <int R, bool HAS_FD, bool SKEW>
struct GroupSize {
    static constexpr int value = HAS_FD ? (SKEW ? (R >= 4 ? 2 : R >= 2 ? HX_FOLD_SKEW_UE2 : 8) : (R >= 8 ? 2 : R >= 4 ? 4 : HX_FOLD_UE_FD))
                                        : (R >= 4 ? 2 : 4);
};

// ROWFD: the R trials of a thread lie in the same fdot row (trial blocks are laid out row by row, rows
// padded to a multiple of 32*R), so the fdot term is one DMUL per event and thread instead of per trial.
// M2T (with ROWFD, CTAs aligned on fdot rows): the fdot term of every event of a stage is computed ONCE per CTA into a
// second shared-memory tile and broadcast like the times -- 3 FP64 ops per event*trial (DMUL, DADD, DADD.RM) instead
// of 3 + 2/R, for one more shared-memory wavefront per event and warp.  Every warp computes the (identical) tile
// itself right after the stage lands, so no block barrier is needed: concurrent writers store the same values.
template <int R, bool HAS_FD, int MODE, bool ROWFD, bool M2T = false>
__global__ void __launch_bounds__(512, 1) fold_kernel(const FoldArgs a) {
    static_assert(!ROWFD || HAS_FD, "ROWFD only makes sense with fdot");
    static_assert(!M2T || (ROWFD && MODE != MODE_EXACT), "M2T needs fdot rows and a fp64 parity mode");
    constexpr bool POW2 = MODE == MODE_POW2 || MODE == MODE_PACKED;
    constexpr bool PACKED = MODE == MODE_PACKED;
    constexpr bool EXACT = MODE == MODE_EXACT;
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr bool SKEW = !EXACT && (!HAS_FD || HX_FOLD_SKEW_FD);  // the exact-phase extension keeps the plain loop
    // plain (f, fdot) loop with <= 2 trials per thread: 8 events per register group with packed counters (nbin >= 128), 4 with
    // plain ones (measured on B200, profiles/r02_variants_ue.txt: nbin 64 +4.5 %, nbin 32 +2.5 %, nbin 128 -4 % with 4)
    constexpr int UE = (HAS_FD && !SKEW && R <= 2 && !PACKED && !EXACT) ? 4 : GroupSize<R, HAS_FD, SKEW>::value;
    const uint32_t sbase = smem_u32(smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t blk = (uint32_t)a.nbin << (PACKED ? 6 : 7);  // bytes of one (warp, r) histogram block
    const uint32_t hist = sbase + a.hist_off;
    const uint32_t tileb = sbase + a.tile_off;
    const int TILE_EV = a.tile_ev;
    const uint32_t TILE_BYTES = (uint32_t)(TILE_EV + TILE_PAD) * 8u;
    int *done = reinterpret_cast<int *>(smem + 16);  // [2] warps that finished the stage
    const uint32_t m2_off = a.tile_off + 2u * TILE_BYTES;  // M2T: two stages of fdot terms behind the two stages of times

    if (POW2 && (hist & (blk - 1)) != 0) {  // host/device layout disagreement: refuse to compute
        if (tid == 0) atomicExch(a.errflag, 1);
        return;
    }

    // zero this CTA's histograms
    {
        const uint32_t nbytes = (uint32_t)a.warps * R * blk;
        for (uint32_t o = tid * 4; o < nbytes; o += blockDim.x * 4) sts_u32(hist + o, 0u);
    }

    // trial constants
    const long long tbl = (long long)blockIdx.x * a.warps + warp;  // block of 32*R trials within this launch
    const long long tb = a.tb0 + tbl;
    double fN[R], fdH[R];
    uint32_t cb[R];
    bool valid[R];
    // index of trial r of this thread in stats[] / the (fdot, f) grid (recomputed where needed: keeping
    // R 64-bit values live through the main loop costs occupancy)
    auto trial_index = [&](int r) -> long long {
        if (ROWFD) {
            const long long ifd = tb / a.n_tb_row;
            return ifd * a.n_f + ((tb - ifd * a.n_tb_row) * R + r) * 32 + lane;
        }
        return (tb * R + r) * 32 + lane;
    };
#pragma unroll
    for (int r = 0; r < R; r++) {
        long long ifd, jf;
        if (ROWFD) {
            ifd = tb / a.n_tb_row;
            jf = ((tb - ifd * a.n_tb_row) * R + r) * 32 + lane;
            valid[r] = jf < a.n_f && ifd * a.n_f + jf < a.n_trials;
        } else {
            const long long g = (tb * R + r) * 32 + lane;
            valid[r] = g < a.n_trials;
            const long long gi = valid[r] ? g : 0;
            ifd = gi / a.n_f, jf = gi - ifd * a.n_f;
        }
        if (!valid[r]) ifd = 0, jf = 0;
        fN[r] = valid[r] ? __dmul_rn(a.freqs[jf], a.scale) : 0.0;
        fdH[r] = (HAS_FD && valid[r]) ? __dmul_rn(__dmul_rn(0.5, a.fdots[ifd]), a.scale) : 0.0;
        cb[r] = hist + (uint32_t)(warp * R + r) * blk + (uint32_t)lane * 4u;
    }
    // ROWFD: one fdot constant per thread (0 when none of its trials is valid: those bins are discarded)
    double fdH_row = 0.0;
    if (ROWFD) {
#pragma unroll
        for (int r = 0; r < R; r++)
            if (valid[r]) fdH_row = fdH[r];
    }
    if (M2T) {  // CTA-uniform (rows are padded to whole CTAs): also right for lanes without a valid trial
        const long long ifd_cta = min((a.tb0 + (long long)blockIdx.x * a.warps) / a.n_tb_row, a.n_trials / a.n_f - 1);
        fdH_row = __dmul_rn(__dmul_rn(0.5, a.fdots[ifd_cta]), a.scale);
    }

    // event range of this CTA
    const long long e0 = (long long)blockIdx.y * a.ev_per_split;
    const long long e1 = min(a.n_ev, e0 + a.ev_per_split);
    const int ntiles = (e1 > e0) ? (int)((e1 - e0 + TILE_EV - 1) / TILE_EV) : 0;
    const uint32_t bar[2] = {sbase, sbase + 8};

    auto issue = [&](int it) {
        const long long s = e0 + (long long)it * TILE_EV;
        const int cnt = (int)min((long long)TILE_EV, e1 - s);
        const uint32_t bytes = (uint32_t)cnt * 8u, bulk = bytes & ~15u;
        const uint32_t dst = tileb + (uint32_t)(it & 1) * TILE_BYTES;
        const char *src = reinterpret_cast<const char *>(a.ev) + s * 8;
        if (bytes != bulk) {  // odd trailing fp64: plain copy, ordered before the arrive below
            double v = *reinterpret_cast<const double *>(src + bulk);
            *reinterpret_cast<double *>(smem + a.tile_off + (uint32_t)(it & 1) * TILE_BYTES + bulk) = v;
        }
        if (bulk) {
            mbar_expect_tx(bar[it & 1], bulk);
            bulk_g2s(dst, src, bulk, bar[it & 1]);
        } else {
            mbar_arrive(bar[it & 1]);
        }
    };

    if (tid == 0) {
        mbar_init(bar[0], 1);
        mbar_init(bar[1], 1);
        done[0] = 0, done[1] = 0;
        fence_mbar_init();
        fence_proxy_async();
        if (ntiles > 0) issue(0);
        if (ntiles > 1) issue(1);
    }
    __syncthreads();

    const uint32_t mask = a.binmask, topbit = a.topbit;
    const int nbin = a.nbin;
    const double gen_tmin = a.gen_tmin;
    const int nwarps = a.warps;

    // PACKED: drain this thread's packed words into the global scratch and clear them
    auto drain_packed = [&]() {
        const int half = nbin >> 1;
        const uint32_t lowmask = topbit - 1u;
#pragma unroll
        for (int r = 0; r < R; r++) {
            uint32_t *dst = a.scratch + ((size_t)(tbl * R + r) * nbin) * 32 + lane;
            for (int w = 0; w < half; w++) {
                const uint32_t ad = cb[r] + ((uint32_t)w << 7);
                const uint32_t v = lds_u32(ad);
                if (v) {
                    sts_u32(ad, 0u);
                    const uint32_t up = v >> a.pack_shift, lowc = (v & lowmask) - up;
                    if (lowc) atomicAdd(dst + (size_t)w * 32, lowc);
                    if (up) atomicAdd(dst + (size_t)(w + half) * 32, up);
                }
            }
        }
    };

    for (int it = 0; it < ntiles; it++) {
        const int buf = it & 1;
        mbar_wait(bar[buf], (uint32_t)(it >> 1) & 1u);
        const int cnt = (int)min((long long)TILE_EV, e1 - (e0 + (long long)it * TILE_EV));
        const unsigned char *tp = smem + a.tile_off + (uint32_t)buf * TILE_BYTES;
        unsigned char *mp = smem + m2_off + (uint32_t)buf * TILE_BYTES;
        if (M2T) {
            for (int q = lane; q < cnt; q += 32) {
                const double tq = *reinterpret_cast<const double *>(tp + (size_t)q * 8);
                *reinterpret_cast<double *>(mp + (size_t)q * 8) = __dmul_rn(__dmul_rn(tq, tq), fdH_row);
            }
            __syncwarp();
        }

        // the events of one stage; GF (generic mode only): every event of the stage takes the short exact path
        auto run_stage = [&](auto gf_tag) {
        constexpr bool GF = decltype(gf_tag)::value;
        int e = 0;
        EvGroup<HAS_FD, UE, EXACT> cur, nxt;
        if constexpr (SKEW) {
        // Skewed software pipeline: the bin addresses of group g+1 are computed (FP64 pipe) while the
        // increments of group g are issued (LSU), so the two instruction classes interleave finely
        // instead of alternating in long runs.
        uint32_t ad_cur[UE * R], inc_cur[PACKED ? UE * R : 1], ad_nxt[UE * R], inc_nxt[PACKED ? UE * R : 1];
        auto addr1 = [&](const EvGroup<HAS_FD, UE, EXACT> &g, int u, int r, uint32_t &inc) -> uint32_t {
            const int uu = HAS_FD ? u : 0;
            const double m2 = ROWFD ? g.m2[uu] : (HAS_FD ? __dmul_rn(g.tt[uu], fdH[r]) : 0.0);
            return bin_addr<HAS_FD, MODE, GF>(g.t[u], HAS_FD ? g.tt[uu] : 0.0, 0.0, m2, fN[r], fdH[r], mask, topbit, nbin,
                                          cb[r], inc);
        };
        const int ngroups = cnt / UE;
        if (ngroups > 0) {
            cur.load(tp, 0);
            nxt.load(tp, UE);
            if (M2T) cur.load_m2(mp, 0), nxt.load_m2(mp, UE);
            else cur.squares(fdH_row);
#pragma unroll
            for (int i = 0; i < UE * R; i++) {
                uint32_t inc;
                ad_cur[i] = addr1(cur, i / R, i % R, inc);
                if (PACKED) inc_cur[i] = inc;
            }
            for (int g = 0; g < ngroups; g++) {
                // group g+1 (garbage past the last group: computed, never issued)
#pragma unroll
                for (int u = 0; u < UE; u++) {
                    cur.t[u] = nxt.t[u];
                    if (M2T) cur.m2[u] = nxt.m2[u];
                }
                nxt.load(tp, (g + 2) * UE);  // may read the pad after the stage; never used past cnt
                if (M2T) nxt.load_m2(mp, (g + 2) * UE);
                else cur.squares(fdH_row);
#pragma unroll
                for (int i = 0; i < UE * R; i++) {
                    uint32_t inc;
                    ad_nxt[i] = addr1(cur, i / R, i % R, inc);
                    if (PACKED) inc_nxt[i] = inc;
                    if (PACKED) red_shared_add(ad_cur[i], inc_cur[i]);
                    else red_shared_inc(ad_cur[i]);
                }
#pragma unroll
                for (int i = 0; i < UE * R; i++) {
                    ad_cur[i] = ad_nxt[i];
                    if (PACKED) inc_cur[i] = inc_nxt[i];
                }
            }
            e = ngroups * UE;
        }
        } else {
        cur.load(tp, 0);
        if (M2T) cur.load_m2(mp, 0);
        for (; e + UE <= cnt; e += UE) {
            nxt.load(tp, e + UE);  // may read the pad after the stage; never used past cnt
            if (M2T && HX_FOLD_HYB > 0 && HX_FOLD_HYB < UE) nxt.load_m2_n(mp, e + UE, UE - HX_FOLD_HYB), cur.squares_from(fdH_row, UE - HX_FOLD_HYB);
            else if (M2T) nxt.load_m2(mp, e + UE);
            else cur.squares(fdH_row);
#pragma unroll
            for (int u = 0; u < UE; u++) {
#pragma unroll
                for (int r = 0; r < R; r++) {
                    uint32_t inc;
                    const int uu = HAS_FD ? u : 0;
                    const double m2 = ROWFD ? cur.m2[uu] : (HAS_FD ? __dmul_rn(cur.tt[uu], fdH[r]) : 0.0);
                    const uint32_t ad = bin_addr<HAS_FD, MODE, GF>(cur.t[u], HAS_FD ? cur.tt[uu] : 0.0, (EXACT && HAS_FD) ? cur.tte[uu] : 0.0, m2, fN[r], fdH[r], mask, topbit, nbin, cb[r], inc);
                    if (PACKED) red_shared_add(ad, inc);
                    else red_shared_inc(ad);
                }
            }
#pragma unroll
            for (int u = 0; u < UE; u++) {
                cur.t[u] = nxt.t[u];
                if (M2T && u < UE - HX_FOLD_HYB) cur.m2[u] = nxt.m2[u];
            }
        }
        }
        for (; e < cnt; e++) {  // ragged end of the last stage
            const double t0 = *reinterpret_cast<const double *>(tp + (size_t)e * 8);
            const double tt0 = HAS_FD ? __dmul_rn(t0, t0) : 0.0;
            const double tte0 = (EXACT && HAS_FD) ? __fma_rn(t0, t0, -tt0) : 0.0;
#pragma unroll
            for (int r = 0; r < R; r++) {
                uint32_t inc;
                const double m2 = HAS_FD ? __dmul_rn(tt0, ROWFD ? fdH_row : fdH[r]) : 0.0;
                const uint32_t ad = bin_addr<HAS_FD, MODE, GF>(t0, tt0, tte0, m2, fN[r], fdH[r], mask, topbit, nbin, cb[r], inc);
                if (PACKED) red_shared_add(ad, inc);
                else red_shared_inc(ad);
            }
        }
        };
        if constexpr (MODE == MODE_GENERIC) {
            // warp-uniform choice per stage: all its events late enough (phase >= 512 cycles for every trial)?
            bool late = true;
            for (int q = lane; q < cnt; q += 32) late &= *reinterpret_cast<const double *>(tp + (size_t)q * 8) >= gen_tmin;
            if (__all_sync(0xffffffffu, late)) run_stage(std::true_type{});
            else run_stage(std::false_type{});
        } else {
            run_stage(std::false_type{});
        }
        if (PACKED && ((it + 1) % a.flush_tiles == 0)) drain_packed();
        // stage hand-back without a block barrier: the last warp to finish `buf` refills it
        __syncwarp();
        if (lane == 0) {
            const int prev = atomicAdd(&done[buf], 1);
            if (prev == nwarps - 1) {
                done[buf] = 0;
                if (it + 2 < ntiles) {
                    fence_proxy_async();
                    issue(it + 2);
                }
            }
        }
    }
    if (PACKED) {  // the packed path always finishes through the scratch + fold_stat_kernel
        drain_packed();
        return;
    }
    __syncthreads();  // all histograms complete; tile region idle

    if (a.scratch != nullptr) {
        // partial counts -> global scratch [trial/32][bin][32], coalesced across the warp
#pragma unroll
        for (int r = 0; r < R; r++) {
            uint32_t *dst = a.scratch + ((size_t)(tbl * R + r) * nbin) * 32 + lane;
            for (int j = 0; j < nbin; j++) {
                uint32_t c = lds_u32(cb[r] + ((uint32_t)j << 7));
                if (c) atomicAdd(dst + (size_t)j * 32, c);
            }
        }
        return;
    }

    // fused epilogue: tables into the (now idle) tile region
    double *tabs = reinterpret_cast<double *>(smem + a.tile_off);
    for (int i = tid; i < 2 * nbin; i += blockDim.x) tabs[i] = a.tab[i];
    __syncthreads();
#pragma unroll
    for (int r = 0; r < R; r++) {
        if (!valid[r]) continue;
        const uint32_t base = cb[r];
        double st = profile_stat([&](int j) { return lds_u32(base + ((uint32_t)j << 7)); }, nbin, a.nharm, tabs,
                                 tabs + nbin);
        a.stats[trial_index(r)] = st;
    }
}

// scratch [trial/32][bin][32] -> stats (+ optional [trial][nbin] profiles)
// Scratch slot l of this launch holds the trial with kernel-order index l0 + l.  Kernel order is the grid
// order g itself (n_fp == 0) or row-padded: index = i_fd * n_fp + j_f with j_f < n_f valid (ROWFD).
__global__ void __launch_bounds__(256) fold_stat_kernel(const uint32_t *__restrict__ scratch, long long l0,
                                                        long long n, long long n_f, long long n_fp,
                                                        long long n_trials, int nbin, int nharm,
                                                        const double *__restrict__ tab, double *__restrict__ stats,
                                                        uint32_t *__restrict__ profiles) {
    extern __shared__ double tabs[];
    for (int i = threadIdx.x; i < 2 * nbin; i += blockDim.x) tabs[i] = tab[i];
    __syncthreads();
    long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n) return;
    long long g = l0 + l;
    if (n_fp > 0) {
        const long long ifd = g / n_fp, jf = g - ifd * n_fp;
        if (jf >= n_f) return;
        g = ifd * n_f + jf;
    }
    if (g >= n_trials) return;
    const uint32_t *src = scratch + (size_t)(l >> 5) * nbin * 32 + (l & 31);
    stats[g] = profile_stat([&](int j) { return src[(size_t)j * 32]; }, nbin, nharm, tabs, tabs + nbin);
    if (profiles)
        for (int j = 0; j < nbin; j++) profiles[(size_t)g * nbin + j] = src[(size_t)j * 32];
}

// min/max of an fp64 array (range checks for the fast path); out = {min, max}
__global__ void minmax_kernel(const double *__restrict__ x, long long n, double *out) {
    double lo = INFINITY, hi = -INFINITY;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double v = x[i];
        if (v != v) v = INFINITY, lo = -INFINITY;  // NaN: make the range non-finite so that the host rejects it
        lo = fmin(lo, v);
        hi = fmax(hi, v);
    }
    for (int o = 16; o; o >>= 1) {
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0 && lo <= hi) {
        // fp64 atomic min/max through an order-preserving integer mapping:
        // non-negative doubles keep their bit pattern, negative ones map to -(magnitude bits)
        auto enc = [](double d) {
            unsigned long long b = (unsigned long long)__double_as_longlong(d);
            return (long long)((b >> 63) ? 0x8000000000000000ull - b : b);
        };
        atomicMin(reinterpret_cast<long long *>(out), enc(lo));
        atomicMax(reinterpret_cast<long long *>(out) + 1, enc(hi));
    }
}
static double dec_ordered(long long e) {
    unsigned long long u = (unsigned long long)e;
    unsigned long long b = (e < 0) ? 0x8000000000000000ull - u : u;
    double d;
    memcpy(&d, &b, 8);
    return d;
}

// ---- host side --------------------------------------------------------------------------------

struct FoldPlan {
    int R, W, ctas_per_sm, tile_ev;
    uint32_t hist_off, tile_off, smem_bytes;
};

// two stages of times (+ two stages of fdot terms with M2T); the fused epilogue reuses the region for its tables
static uint32_t tile_region_bytes(int tile_ev, int nbin, bool m2t) {
    return std::max<uint32_t>((m2t ? 4u : 2u) * (uint32_t)(tile_ev + TILE_PAD) * 8u, (uint32_t)nbin * 16u);
}

static bool plan_layout(const DeviceInfo &di, int nbin, int R, int W, int mode, int regs, bool m2t, FoldPlan *p) {
    const bool pow2 = mode == MODE_POW2 || mode == MODE_PACKED;
    const uint32_t blk = (uint32_t)nbin * (mode == MODE_PACKED ? 64u : 128u);
    const uint32_t hist_bytes = (uint32_t)W * R * blk;
    uint32_t hist_off, tile_off;
    int tile_ev = m2t ? TILE_EV_MAX / 2 : TILE_EV_MAX;  // M2T stages hold two arrays; small stages keep 3 CTAs per SM at nbin = 128
    if (m2t)
        if (const char *e = getenv("HX_FOLD_M2T_TILE")) tile_ev = std::max(64, std::min(TILE_EV_MAX, atoi(e) / 64 * 64));  // tuning hook
    if (pow2) {
        uint32_t abs0 = di.dyn_smem_base + SMEM_HDR;
        uint32_t hist_abs = (abs0 + blk - 1) / blk * blk;
        hist_off = hist_abs - di.dyn_smem_base;
        const uint32_t gap = hist_off - SMEM_HDR;
        // largest stage that fits in the alignment gap in front of the histograms, if any does
        int fit = 0;
        for (int t = TILE_EV_MAX; t >= (m2t ? 128 : 256); t -= 64)
            if (tile_region_bytes(t, nbin, m2t) <= gap) {
                fit = t;
                break;
            }
        if (fit) {
            tile_ev = fit;
            tile_off = SMEM_HDR;
        } else {
            tile_off = hist_off + hist_bytes;
        }
    } else {
        hist_off = SMEM_HDR;
        tile_off = hist_off + hist_bytes;
    }
    uint32_t total = std::max(hist_off + hist_bytes, tile_off + tile_region_bytes(tile_ev, nbin, m2t));
    if (total > (uint32_t)di.smem_optin) return false;
    int k_smem = di.smem_per_sm / (int)(total + di.reserved_smem);
    const int regs_alloc = (regs + 7) / 8 * 8;                 // register allocation granularity
    int k_thr = std::min(2048, 65536 / regs_alloc) / (32 * W);  // resident threads allowed by the register file
    int k = std::min(std::min(k_smem, k_thr), 32);
    if (k < 1) return false;
    p->R = R, p->W = W, p->ctas_per_sm = k, p->tile_ev = tile_ev;
    p->hist_off = hist_off, p->tile_off = tile_off, p->smem_bytes = total;
    return true;
}

typedef void (*fold_fn_t)(const FoldArgs);
template <int R, bool HAS_FD, bool ROWFD, bool M2T>
static fold_fn_t pick_fold_M(int mode) {
    if constexpr (M2T) {
        return mode == MODE_PACKED ? fold_kernel<R, HAS_FD, MODE_PACKED, ROWFD, true>
               : mode == MODE_POW2 ? fold_kernel<R, HAS_FD, MODE_POW2, ROWFD, true>
                                   : fold_kernel<R, HAS_FD, MODE_GENERIC, ROWFD, true>;
    } else {
        return mode == MODE_PACKED ? fold_kernel<R, HAS_FD, MODE_PACKED, ROWFD>
               : mode == MODE_POW2 ? fold_kernel<R, HAS_FD, MODE_POW2, ROWFD>
               : mode == MODE_EXACT ? fold_kernel<R, HAS_FD, MODE_EXACT, ROWFD>
                                    : fold_kernel<R, HAS_FD, MODE_GENERIC, ROWFD>;
    }
}
template <int R>
static fold_fn_t pick_fold_R(bool has_fd, int mode, bool rowfd, bool m2t) {
    if (!has_fd) return pick_fold_M<R, false, false, false>(mode);
    if constexpr (R == 2 || R == 4) {
        if (rowfd && m2t && mode != MODE_EXACT) return pick_fold_M<R, true, true, true>(mode);
    }
    return rowfd ? pick_fold_M<R, true, true, false>(mode) : pick_fold_M<R, true, false, false>(mode);
}
static fold_fn_t pick_fold(int R, bool has_fd, int mode, bool rowfd, bool m2t = false) {
    switch (R) {
        case 1: return pick_fold_R<1>(has_fd, mode, false, false);  // one trial per thread: nothing to share
        case 2: return pick_fold_R<2>(has_fd, mode, rowfd, m2t);
        case 4: return pick_fold_R<4>(has_fd, mode, rowfd, m2t);
        default: return pick_fold_R<8>(has_fd, mode, rowfd, false);
    }
}

// CUDA events bracketing the last fold_kernel launch (bench.py's live roofline measurement)
static cudaEvent_t g_ev0[64], g_ev1[64];
static bool g_ev_ok[64] = {false};
static int g_last_timed_device = -1;
static long long g_last_grid = 0, g_last_block = 0, g_last_smem = 0;
static int g_last_R = 0, g_last_mode = 0, g_last_rowfd = 0, g_last_m2t = 0, g_last_has_fd = 0;

static int g_fold_R_override = 0;  // testing/tuning hook (hx_fold_set_tuning)
static int g_fold_W_override = 0;
static int g_fold_pack_override = -1;  // -1 auto, 0 never, 1 whenever possible
static int g_fold_m2t = 1;             // per-CTA fdot-term tile (M2T) when applicable; HX_FOLD_M2T=0 disables it
static int g_fold_gen_fast = 1;        // short exact path of the generic mode (hx_fold_set_generic_fast)

// min/max of the event times, trial frequencies and fdots (device arrays): out = [tmin,tmax, fmin,fmax, fdmin,fdmax];
// fails with HX_ERR_ARG when any of them is not finite.  Synchronizes `st`.
int device_ranges(const double *d_times, int64_t n_ev, const double *d_freqs, int64_t n_f, const double *d_fdots,
                  int64_t n_fd, double out[6], cudaStream_t st) {
    AsyncBuf rng(st);
    HX_CUDA_TRY(rng.alloc(6 * sizeof(long long)));
    long long init[6];
    for (int i = 0; i < 3; i++) init[2 * i] = INT64_MAX, init[2 * i + 1] = INT64_MIN;
    HX_CUDA_TRY(cudaMemcpyAsync(rng.p, init, sizeof(init), cudaMemcpyHostToDevice, st));
    if (n_ev > 0) {
        minmax_kernel<<<(int)std::min<long long>((n_ev + 255) / 256, 1184), 256, 0, st>>>(d_times, n_ev,
                                                                                          rng.as<double>());
        HX_CHECK_LAUNCH("minmax_kernel");
    }
    minmax_kernel<<<(int)std::min<long long>((n_f + 255) / 256, 1184), 256, 0, st>>>(d_freqs, n_f,
                                                                                     rng.as<double>() + 2);
    HX_CHECK_LAUNCH("minmax_kernel");
    minmax_kernel<<<1, 256, 0, st>>>(d_fdots, n_fd, rng.as<double>() + 4);
    HX_CHECK_LAUNCH("minmax_kernel");
    long long henc[6];
    HX_CUDA_TRY(cudaMemcpyAsync(henc, rng.p, sizeof(henc), cudaMemcpyDeviceToHost, st));
    HX_CUDA_TRY(cudaStreamSynchronize(st));
    out[0] = out[1] = 0.0;
    if (n_ev > 0) out[0] = dec_ordered(henc[0]), out[1] = dec_ordered(henc[1]);
    for (int i = 2; i < 6; i++) out[i] = dec_ordered(henc[i]);
    for (int i = 0; i < 6; i++)
        if (!std::isfinite(out[i])) {
            set_error("hx_fold_search: non-finite times / frequencies / fdots");
            return HX_ERR_ARG;
        }
    return HX_OK;
}

int fold_search_dev_impl(const double *d_times, int64_t n_ev, const double *d_freqs, int64_t n_f,
                         const double *d_fdots, int64_t n_fd, int nbin, int nharm, unsigned flags, double *d_stats,
                         uint32_t *d_profiles, cudaStream_t st) {
    std::lock_guard<std::recursive_mutex> api_lock(api_mutex());
    if (!d_times || !d_freqs || !d_fdots || !d_stats) {
        set_error("hx_fold_search: null pointer argument");
        return HX_ERR_ARG;
    }
    if (n_ev < 0 || n_f < 0 || n_fd < 0 || nbin < 2 || nbin > HX_MAX_NBIN || nharm < 0 || nharm > HX_MAX_NHARM) {
        set_error("hx_fold_search: bad sizes (n_ev=%lld n_f=%lld n_fd=%lld nbin=%d nharm=%d)", (long long)n_ev,
                  (long long)n_f, (long long)n_fd, nbin, nharm);
        return HX_ERR_ARG;
    }
    if ((reinterpret_cast<uintptr_t>(d_times) & 15) != 0) {
        set_error("hx_fold_search_dev: d_times must be 16-byte aligned");
        return HX_ERR_ARG;
    }
    const long long n_trials = (long long)n_f * n_fd;
    if (n_trials == 0) return HX_OK;
    int device;
    HX_CUDA_TRY(cudaGetDevice(&device));
    DeviceInfo di;
    int rc = get_device_info(device, &di);
    if (rc) return rc;

    // ranges: [tmin,tmax, fmin,fmax, fdmin,fdmax]
    double r6[6];
    if ((rc = device_ranges(d_times, n_ev, d_freqs, n_f, d_fdots, n_fd, r6, st)) != HX_OK) return rc;
    const double tmin = r6[0], tmax = r6[1], fmin = r6[2], fmax = r6[3], fdmin = r6[4], fdmax = r6[5];
    AsyncBuf rng(st);
    HX_CUDA_TRY(rng.alloc(sizeof(int)));
    HX_CUDA_TRY(cudaMemsetAsync(rng.p, 0, sizeof(int), st));
    int *d_err = rng.as<int>();
    const bool has_fd = !(fdmin == 0.0 && fdmax == 0.0);
    const double tabs_ = std::max(std::fabs(tmin), std::fabs(tmax));
    const double fabs_ = std::max(std::fabs(fmin), std::fabs(fmax));
    const double fdabs_ = std::max(std::fabs(fdmin), std::fabs(fdmax));
    const double phmax = tabs_ * fabs_ + 0.5 * tabs_ * tabs_ * fdabs_;
    if (!(phmax < 1125899906842624.0 /* 2^50 */)) {
        set_error("hx_fold_search: |phase| up to %.3g cycles exceeds the fp64 folding range (2^50)", phmax);
        return HX_ERR_RANGE;
    }
    const bool nbin_p2 = (nbin & (nbin - 1)) == 0;
    // fast path needs: power-of-two nbin, scaled phase inside the magic-add range, and provably
    // non-negative phases (a negative phase in (-0.5, 0) is rounded by the reference's
    // `ph - floor(ph)` before binning; the generic path reproduces that rounding).
    bool nonneg = tmin >= 0 && fmin >= 0 && (fdmin >= 0 || fmin + 0.5 * tmax * fdmin * (1 + 1e-12) > 0);
    const bool fast_ok = nbin_p2 && nonneg && phmax * nbin * 128.0 < 1125899906842624.0 &&
                         !(flags & HX_FOLD_FORCE_GENERIC);
    int mode = MODE_GENERIC;
    if (fast_ok && nbin <= 256) mode = MODE_POW2;
    // packed counters (two bins per word) for the large profiles, where the plain layout leaves
    // too few resident warps
    const bool can_pack = fast_ok && nbin >= 32 && nbin <= 512;
    if (can_pack && (g_fold_pack_override == 1 || (g_fold_pack_override < 0 && nbin >= 128))) mode = MODE_PACKED;
    if (g_fold_pack_override == 0 && mode == MODE_PACKED) mode = nbin <= 256 ? MODE_POW2 : MODE_GENERIC;
    if (flags & HX_FOLD_EXACT_PHASE) mode = MODE_EXACT;

    // ---- plan -------------------------------------------------------------------------------
    // (trials per thread, warps per CTA) measured best on B200 for each (footprint, fdot) class
    // (tools/sweep_table.py); anything that does not fit falls back to "most resident warps".
    const int foot = mode == MODE_PACKED ? nbin / 2 : nbin;  // histogram words per trial
    int R, W_pref;
    if (foot <= 16) R = has_fd ? 8 : 4, W_pref = has_fd ? 2 : 8;
    else if (foot <= 32) R = has_fd ? 2 : 4, W_pref = 4;  // with fdot: R = 2 takes the fdot-term tile (M2T), +3 % (profiles/r01_sweep_r3.txt)
    else if (foot <= 64) R = 2, W_pref = 4;
    else if (foot <= 128) R = 1, W_pref = 12;
    else R = 1, W_pref = 4;
    if (g_fold_R_override) R = g_fold_R_override;
    if (g_fold_W_override) W_pref = g_fold_W_override;
    bool m2t_on = g_fold_m2t != 0;
    if (const char *e = getenv("HX_FOLD_M2T")) m2t_on = atoi(e) != 0;  // tuning hook
    // M2T pays when the fdot term is shared by few trials per thread (it removes 2/R FP64 ops per unit)
    // (measured on B200: +10 % at R = 2 -- C3 3.21e12 -> 3.54e12 --, -2 % at R = 4 where only 0.5 op per unit is saved)
    auto use_m2t = [&](int r) { return m2t_on && has_fd && r == 2 && n_f >= 32LL * r && mode != MODE_EXACT; };
    auto kernel_regs = [&](int r, int *out) -> int {
        cudaFuncAttributes fa;
        const bool rf = has_fd && r >= 2 && n_f >= 32LL * r && mode != MODE_EXACT;
        HX_CUDA_TRY(cudaFuncGetAttributes(&fa, reinterpret_cast<const void *>(pick_fold(r, has_fd, mode, rf, use_m2t(r)))));
        *out = fa.numRegs;
        return HX_OK;
    };
    int regs = 64;
    if ((rc = kernel_regs(R, &regs)) != HX_OK) return rc;
    FoldPlan best{};
    int best_res = -1;
    if (plan_layout(di, nbin, R, W_pref, mode, regs, use_m2t(R), &best)) best_res = best.ctas_per_sm * W_pref;
    for (int pass = 0; pass < 2 && best_res < 0; pass++) {
        if (pass == 1) {  // very large nbin: one trial per thread
            if (R == 1) break;
            R = 1;
            if ((rc = kernel_regs(R, &regs)) != HX_OK) return rc;
        }
        for (int W = 1; W <= 16; W++) {
            FoldPlan p;
            if (W == 3 || (W > 4 && W % 4 != 0)) continue;  // keep the SM's four schedulers evenly loaded
            if (!plan_layout(di, nbin, R, W, mode, regs, use_m2t(R), &p)) continue;
            int res = p.ctas_per_sm * W;
            if (res > best_res) best = p, best_res = res;
        }
    }
    if (best_res < 0) {
        set_error("hx_fold_search: nbin=%d does not fit in shared memory", nbin);
        return HX_ERR_ARG;
    }
    // ROWFD: threads with R >= 2 trials take them from one fdot row (rows padded to 32*R frequencies)
    const bool rowfd = has_fd && R >= 2 && n_f >= 32LL * R && mode != MODE_EXACT;
    const bool m2t = rowfd && use_m2t(R);
    long long n_tb_row = (n_f + 32LL * R - 1) / (32LL * R);
    long long n_tb = rowfd ? n_tb_row * n_fd : (n_trials + 32LL * R - 1) / (32LL * R);  // 32*R-trial blocks
    int W = best.W;
    // Few trial blocks: either shrink the CTAs so that every SM gets one, or keep the preferred shape and let the event
    // splitting below multiply the CTAs.  Small CTAs cost resident warps (the stage buffers are per CTA), splitting needs
    // enough events: score both (wave efficiency after the best split x resident warps relative to the preferred plan)
    // and keep the better one.  (C5 strong scaling, 12 500 f per GPU: W = 2 ran 10 warps per SM, W = 4 + splits runs 12.)
    {
        long long want = (n_tb + di.sm_count - 1) / di.sm_count;
        if (want < W) {
            // warps per CTA that keep the four schedulers of an SM evenly loaded: 1, 2 or a multiple of 4 (3- and 6-warp
            // CTAs measured 15-30 % slower in every sweep: profiles/r01_sweep_m2t.txt, r02_sweep_qlane_fused_skew.txt)
            int Wn = (int)std::max<long long>(1, want);
            if (Wn == 3) Wn = 4;
            else if (Wn > 4) Wn = (Wn + 3) / 4 * 4;
            FoldPlan pn;
            if (Wn < W && plan_layout(di, nbin, R, Wn, mode, regs, m2t, &pn)) {
                const double warps_pref = (double)best.ctas_per_sm * W;
                auto score_of = [&](const FoldPlan &pl, int Wc) {
                    const long long ncx = (n_tb + Wc - 1) / Wc, sl = (long long)di.sm_count * pl.ctas_per_sm;
                    const long long s_max = n_ev > 0 ? std::min<long long>(std::max<long long>(1, n_ev / std::max(4 * pl.tile_ev, 8192)), 4096) : 1;
                    double b = 0.0;
                    for (long long sp = 1; sp <= s_max; sp++) {
                        const double waves = (double)(ncx * sp) / (double)sl;
                        const double sc = waves / std::ceil(waves) * (1.0 - 0.04 / std::max(waves, 1.0)) /
                                          (1.0 + 4.0 * nbin * (double)sp / (double)std::max<int64_t>(n_ev, 1));
                        b = std::max(b, sc);
                        if (waves >= 16.0) break;
                    }
                    return b * std::min(1.0, (double)pl.ctas_per_sm * Wc / warps_pref);
                };
                if (score_of(pn, Wn) > score_of(best, W)) best = pn, W = Wn;
            }
        }
    }
    if (m2t) {  // whole CTAs per fdot row: the fdot-term tile of a CTA is for ONE fdot
        n_tb_row = (n_tb_row + W - 1) / W * W;
        n_tb = n_tb_row * n_fd;
    }
    const long long n_fp = rowfd ? n_tb_row * 32LL * R : 0;  // padded row length in kernel order
    // trial batching bounds the global scratch of the non-fused path (1 GiB)
    const size_t scratch_per_tb = (size_t)R * nbin * 32 * sizeof(uint32_t);
    size_t scratch_cap = (size_t)1 << 30;
    if (const char *e = getenv("HX_FOLD_SCRATCH_CAP_KB")) {  // testing hook: force several trial batches
        const long long kb = atoll(e);
        if (kb > 0) scratch_cap = (size_t)kb << 10;
    }
    long long tb_per_batch = std::max<long long>(W, (long long)(scratch_cap / scratch_per_tb) / W * W);
    const bool must_scratch = mode == MODE_PACKED || d_profiles != nullptr || (flags & HX_FOLD_FORCE_SCRATCH);
    const long long n_tb_first = must_scratch ? std::min(n_tb, tb_per_batch) : n_tb;
    const long long n_cta_x = (n_tb_first + W - 1) / W;
    const long long slots = (long long)di.sm_count * best.ctas_per_sm;
    // Event splitting: every CTA runs for the whole event list, so with few CTA-waves the last,
    // partially filled wave costs as much as a full one.  Splitting the events s ways multiplies the
    // number of (shorter) CTAs; pick the s that fills the waves best, charging the extra flush of
    // the partial histograms (nbin*32 global atomics per warp-slot per split).
    long long ev_splits = 1;
    if (n_ev > 0 && (double)n_cta_x / (double)slots < 8.0) {
        const long long s_max = std::min<long long>(std::max<long long>(1, n_ev / std::max(4 * best.tile_ev, 8192)), 4096);
        double best_score = -1.0;
        for (long long sp = 1; sp <= s_max; sp++) {
            const double waves = (double)(n_cta_x * sp) / (double)slots;
            // a single wave cannot overlap the CTAs' prologues / drains with other CTAs' folding (measured on the C5 shape:
            // 1 wave 0.953, 2 waves 0.975, 4 waves 0.977 of the many-wave rate): mild preference for more waves
            const double eff = waves / std::ceil(waves) * (1.0 - 0.04 / std::max(waves, 1.0));
            const double penalty = 1.0 + ((sp > 1 || must_scratch) ? 4.0 * nbin * (double)sp / (double)n_ev : 0.0);
            const double score = eff / penalty;
            if (score > best_score * 1.01) best_score = score, ev_splits = sp;
            if (waves >= 16.0) break;
        }
    }
    long long ev_per_split = (n_ev + ev_splits - 1) / std::max<long long>(ev_splits, 1);
    ev_per_split = (ev_per_split + best.tile_ev - 1) / best.tile_ev * best.tile_ev;
    if (ev_per_split == 0) ev_per_split = best.tile_ev;
    ev_splits = std::max<long long>(1, (n_ev + ev_per_split - 1) / ev_per_split);
    const bool fused = ev_splits == 1 && !must_scratch;
    if (fused) tb_per_batch = n_tb;
    if (n_tb / std::max<long long>(tb_per_batch, 1) > 1000000) {
        set_error("hx_fold_search: trial grid too large");
        return HX_ERR_ARG;
    }

    // ---- device temporaries -----------------------------------------------------------------
    AsyncBuf b_tab(st), b_scratch(st);
    std::vector<double> tab;
    host_tables(nbin, tab);
    HX_CUDA_TRY(b_tab.alloc(tab.size() * sizeof(double)));
    HX_CUDA_TRY(cudaMemcpyAsync(b_tab.p, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    // the grid of a batch overhangs it by < W trial blocks (whole CTAs): the scratch is W-padded
    const size_t sbytes = fused ? 0 : (size_t)((std::min(n_tb, tb_per_batch) + W - 1) / W * W) * scratch_per_tb;
    if (!fused) HX_CUDA_TRY(b_scratch.alloc(sbytes));
    void *d_tab = b_tab.p, *d_scratch = fused ? nullptr : b_scratch.p;

    int log2nbin = 0;
    while ((1 << log2nbin) < nbin) log2nbin++;
    FoldArgs a{};
    a.ev = d_times, a.n_ev = n_ev, a.freqs = d_freqs, a.n_f = n_f, a.fdots = d_fdots, a.n_trials = n_trials;
    a.nbin = nbin, a.nharm = nharm;
    a.binmask = mode == MODE_PACKED ? (uint32_t)(nbin / 2 - 1) << 7 : mode == MODE_POW2 ? (uint32_t)(nbin - 1) << 7 : 0u;
    a.pack_shift = 6 + log2nbin;
    a.topbit = 1u << a.pack_shift;
    a.tile_ev = best.tile_ev;
    // the low field of a packed word holds 2^pack_shift - 1 events before it carries
    a.flush_tiles = std::max(1, (int)(((1u << a.pack_shift) - 1u) / (uint32_t)best.tile_ev));
    a.scale = (mode == MODE_POW2 || mode == MODE_PACKED) ? (double)nbin * 128.0 : 1.0;
    // generic mode, short exact path: needs non-negative phases, phase * nbin inside the magic-add range, and a phase of
    // at least 512 cycles, which holds for every trial once t * (fmin + t * min(fdmin, 0) / 2) >= 512; the bound below is
    // evaluated with a safety margin (events it excludes just take the reference's operation sequence)
    a.gen_tmin = INFINITY;
    if (mode == MODE_GENERIC && nonneg && !(flags & HX_FOLD_FORCE_GENERIC) && phmax * nbin < 1125899906842624.0 /* 2^50 */ &&
        fmin > 0) {
        const double feff = fmin + 0.5 * tmax * std::min(fdmin, 0.0);  // smallest f + t*fd/2 over the data span
        if (feff > 0) a.gen_tmin = 512.0 / feff * (1.0 + 1e-9) + 1e-9;
    }
    if (g_fold_gen_fast == 0) a.gen_tmin = INFINITY;
    a.stats = d_stats, a.scratch = reinterpret_cast<uint32_t *>(d_scratch), a.tab = reinterpret_cast<double *>(d_tab);
    a.warps = W, a.hist_off = best.hist_off, a.tile_off = best.tile_off, a.ev_per_split = ev_per_split;
    a.errflag = d_err;

    a.n_tb_row = n_tb_row;
    fold_fn_t fn = pick_fold(R, has_fd, mode, rowfd, m2t);
    HX_CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)best.smem_bytes));
    if (!g_ev_ok[device]) {
        HX_CUDA_TRY(cudaEventCreate(&g_ev0[device]));
        HX_CUDA_TRY(cudaEventCreate(&g_ev1[device]));
        g_ev_ok[device] = true;
    }
    long long total_ctas = 0;
    for (long long tb0 = 0; tb0 < n_tb; tb0 += tb_per_batch) {
        const long long n_tb_b = std::min(tb_per_batch, n_tb - tb0);
        const long long grid_x = (n_tb_b + W - 1) / W;
        if (!fused) HX_CUDA_TRY(cudaMemsetAsync(d_scratch, 0, (size_t)(grid_x * W) * scratch_per_tb, st));
        a.tb0 = tb0;
        if (tb0 == 0) HX_CUDA_TRY(cudaEventRecord(g_ev0[device], st));
        dim3 grid((unsigned)grid_x, (unsigned)ev_splits);
        fn<<<grid, W * 32, best.smem_bytes, st>>>(a);
        HX_CHECK_LAUNCH("fold_kernel");
        if (tb0 + tb_per_batch >= n_tb) HX_CUDA_TRY(cudaEventRecord(g_ev1[device], st));
        total_ctas += grid_x * ev_splits;
        if (!fused) {
            const long long l0 = tb0 * 32 * R;
            const long long ng = n_tb_b * 32 * R;
            fold_stat_kernel<<<(unsigned)((ng + 255) / 256), 256, 2 * nbin * sizeof(double), st>>>(
                reinterpret_cast<uint32_t *>(d_scratch), l0, ng, n_f, n_fp, n_trials, nbin, nharm,
                reinterpret_cast<double *>(d_tab), d_stats, d_profiles);
            HX_CHECK_LAUNCH("fold_stat_kernel");
        }
    }
    g_last_timed_device = device;
    g_last_grid = total_ctas, g_last_block = W * 32, g_last_smem = best.smem_bytes;
    g_last_R = R, g_last_mode = mode, g_last_rowfd = rowfd, g_last_m2t = m2t, g_last_has_fd = has_fd;
    int herr = 0;
    HX_CUDA_TRY(cudaMemcpyAsync(&herr, d_err, sizeof(int), cudaMemcpyDeviceToHost, st));
    HX_CUDA_TRY(cudaStreamSynchronize(st));
    if (herr) {
        set_error("fold_kernel: shared-memory layout mismatch (dyn_smem_base probe wrong)");
        return HX_ERR_INTERNAL;
    }
    return HX_OK;
}

}  // namespace hx

extern "C" {

int hx_fold_last_kernel_ms(double *ms, int64_t *grid, int64_t *block, int64_t *smem_bytes) {
    if (!ms || hx::g_last_timed_device < 0) {
        hx::set_error("hx_fold_last_kernel_ms: no fold kernel has been launched");
        return HX_ERR_ARG;
    }
    const int d = hx::g_last_timed_device;
    float f = 0.f;
    HX_CUDA_TRY(cudaEventSynchronize(hx::g_ev1[d]));
    HX_CUDA_TRY(cudaEventElapsedTime(&f, hx::g_ev0[d], hx::g_ev1[d]));
    *ms = f;
    if (grid) *grid = hx::g_last_grid;
    if (block) *block = hx::g_last_block;
    if (smem_bytes) *smem_bytes = hx::g_last_smem;
    return HX_OK;
}

int hx_fold_last_kernel_shape(int *R, int *mode, int *rowfd, int *m2t, double *fp64_per_unit,
                              double *smem_wavefronts_per_unit) {
    if (hx::g_last_timed_device < 0) {
        hx::set_error("hx_fold_last_kernel_shape: no fold kernel has been launched");
        return HX_ERR_ARG;
    }
    const int r = hx::g_last_R;
    if (R) *R = r;
    if (mode) *mode = hx::g_last_mode;
    if (rowfd) *rowfd = hx::g_last_rowfd;
    if (m2t) *m2t = hx::g_last_m2t;
    const bool fast = hx::g_last_mode == hx::MODE_POW2 || hx::g_last_mode == hx::MODE_PACKED;
    if (fp64_per_unit) {
        // f only: DMUL + DADD.RM.  With fdot: + DADD, and the fdot term fl(fl(t*t)*fdH): from the per-CTA tile (m2t, 0 per
        // unit), 2 DMUL per event and thread (rowfd) or 1 DMUL per event and thread + 1 per unit.
        double n = 2.0;
        if (hx::g_last_has_fd) n = hx::g_last_m2t ? 3.0 : hx::g_last_rowfd ? 3.0 + 2.0 / r : 4.0 + 1.0 / r;
        *fp64_per_unit = fast ? n : -1.0;
    }
    if (smem_wavefronts_per_unit)  // one ATOMS per unit; a broadcast fp64 costs one wavefront per event and warp (LDS.128 = 2 events)
        *smem_wavefronts_per_unit = 1.0 + (hx::g_last_m2t ? 2.0 : 1.0) / r;
    return HX_OK;
}

int hx_fold_set_generic_fast(int on) {
    hx::g_fold_gen_fast = on ? 1 : 0;
    return HX_OK;
}

int hx_fold_set_packing(int mode) {
    hx::g_fold_pack_override = (mode == 0 || mode == 1) ? mode : -1;
    return HX_OK;
}

int hx_fold_set_tuning(int R, int W) {
    hx::g_fold_R_override = (R == 1 || R == 2 || R == 4 || R == 8) ? R : 0;
    hx::g_fold_W_override = (W >= 1 && W <= 16) ? W : 0;
    return HX_OK;
}

int hx_fold_search_dev(const double *d_times, int64_t n_ev, const double *d_freqs, int64_t n_f,
                       const double *d_fdots, int64_t n_fd, int nbin, int nharm, unsigned flags, double *d_stats,
                       uint32_t *d_profiles, void *stream) {
    return hx::fold_search_dev_impl(d_times, n_ev, d_freqs, n_f, d_fdots, n_fd, nbin, nharm, flags, d_stats,
                                    d_profiles, reinterpret_cast<cudaStream_t>(stream));
}

int hx_fold_search(const double *times, int64_t n_ev, const double *freqs, int64_t n_f, const double *fdots,
                   int64_t n_fd, int nbin, int nharm, unsigned flags, double *stats, uint32_t *profiles, int device) {
    if (!times || !freqs || !fdots || !stats) {
        hx::set_error("hx_fold_search: null pointer argument");
        return HX_ERR_ARG;
    }
    if (n_ev < 0 || n_f < 0 || n_fd < 0) {
        hx::set_error("hx_fold_search: negative size");
        return HX_ERR_ARG;
    }
    hx::DeviceGuard dev_guard;  // declared before the device buffers: they are freed before the device is restored
    HX_CUDA_TRY(dev_guard.enter(device));
    const size_t n_tr = (size_t)n_f * n_fd;
    if (n_tr == 0) return HX_OK;
    hx::DevBuf dt, df, dfd, ds, dp;
    HX_CUDA_TRY(dt.alloc((size_t)n_ev * 8));
    HX_CUDA_TRY(df.alloc((size_t)n_f * 8));
    HX_CUDA_TRY(dfd.alloc((size_t)n_fd * 8));
    HX_CUDA_TRY(ds.alloc(n_tr * 8));
    if (profiles) HX_CUDA_TRY(dp.alloc(n_tr * nbin * 4));
    HX_CUDA_TRY(cudaMemcpy(dt.p, times, (size_t)n_ev * 8, cudaMemcpyHostToDevice));
    HX_CUDA_TRY(cudaMemcpy(df.p, freqs, (size_t)n_f * 8, cudaMemcpyHostToDevice));
    HX_CUDA_TRY(cudaMemcpy(dfd.p, fdots, (size_t)n_fd * 8, cudaMemcpyHostToDevice));
    int rc = hx::fold_search_dev_impl(dt.as<double>(), n_ev, df.as<double>(), n_f, dfd.as<double>(), n_fd, nbin, nharm,
                                      flags, ds.as<double>(), profiles ? dp.as<uint32_t>() : nullptr, 0);
    if (rc) return rc;
    HX_CUDA_TRY(cudaMemcpy(stats, ds.p, n_tr * 8, cudaMemcpyDeviceToHost));
    if (profiles) HX_CUDA_TRY(cudaMemcpy(profiles, dp.p, n_tr * nbin * 4, cudaMemcpyDeviceToHost));
    return HX_OK;
}

}  // extern "C"
