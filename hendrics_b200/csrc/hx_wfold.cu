// K1w: weighted / exposure-corrected (f, fdot) trial-grid fold search for sm_100a.
//
// Replaces the SLOW branch of stingray.pulse.search._folding_search (use_times=True), which
// z_n_search / epoch_folding_search take when `expocorr` is set or `weights` is an array -- i.e. what
// hendrics/efsearch.py:1155-1157 (light-curve input: weights = events.counts) and :1164-1167
// (`--expocorr`: gti passed down) select.  Per trial upstream runs
//     fold_events(ts, f, fd, nbin=, weights=, gti=, expocorr=)   [stingray.pulse.pulsar]
//         phases  = pulse_phase(ts, f, fd)            t*f + (0.5*t^2)*fd, separately rounded, minus floor
//         profile = np.histogram(phases, bins=linspace(0, 1, nbin + 1), weights=weights)
//         profile /= phase_exposure(start_phase, stop_phase, 1, nbin, gti=pulse_phase(gti, f, fd, to_1=False))
//     -> ef_profile_stat(profile) | z_n_binned_events(profile, nharm)
//
// Design: K1's layout (hx_fold.cu).  lane = trial; every thread owns R trials and, for each, a PRIVATE
// nbin-bin histogram column in shared memory laid out [bin][lane], so a warp-wide access never has a bank
// conflict and no other thread ever touches the column.  Two cell types:
//   * CELL_U32 / CELL_UNIT -- the weights are non-negative integers (light-curve counts; checked on the device,
//     sum < 2^32) or every event weighs the same: 4-byte cells, the update is one fire-and-forget
//     `red.shared.add.u32 [cell], w` (1 shared-memory wavefront per warp), exact;
//   * CELL_F64 -- arbitrary fp64 weights: 8-byte cells (2 wavefronts per access), the update is a plain
//     LDS.64 + DADD + STS.64 (the hardware has no fp64 shared atomic: atomicAdd compiles to a CAS loop).  Events
//     are processed in register groups of U: all U*R bin addresses first (independent FP64 chains), then per
//     trial U cell loads in flight; values are forwarded in registers when two events of a group hit the same
//     cell, so the result is exactly the sequential sum in event order.
// Event times and weights stream through a 2-stage cp.async.bulk (TMA engine) + mbarrier pipeline and are
// broadcast to all threads.  Bin index: searchsorted semantics of np.histogram on the explicit edges; for
// power-of-two nbin the edges j/nbin are exact and the bin is floor(phase*nbin) (phase == 1.0 -> last bin, whose
// right edge is closed); otherwise floor(phase*nbin) is corrected by at most one against the edges
// fl(j*fl(1/nbin)) (np.linspace's arithmetic, recomputed in registers).  Partial profiles always go to a global
// fp64 scratch [trial/32][bin][32] (RED.ADD.F64, coalesced); wfold_expo_kernel divides it by the per-trial
// exposure (double-double arithmetic standing in for upstream's long double), wfold_stat_kernel reduces it to
// chi^2 / Z^2_n.
//
// np.histogram sums weights by block-wise sort + cumsum + diff, so fp64-weighted profiles agree to rounding
// (exactly when the weights are integers and the sums stay below 2^53); statistics within the 1e-9 tolerance.
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include <cub/device/device_select.cuh>

#include "hx_common.cuh"

namespace hx {

constexpr int WTILE_EV = 512;        // events per pipeline stage
constexpr int WTILE_PAD = 8;         // events of slack behind a stage (the last register group reads past cnt)
constexpr uint32_t WSMEM_HDR = 32;   // two mbarriers + two "warps done" counters

enum { CELL_F64 = 0, CELL_U32 = 1, CELL_UNIT = 2 };
enum { WMODE_EDGES = 0, WMODE_FLOOR = 1, WMODE_FAST = 2 };

struct WFoldArgs {
    const double *ev;
    const void *wt;  // CELL_F64: double[n_ev]; CELL_U32: uint32[n_ev]; CELL_UNIT: unused
    long long n_ev;
    const double *freqs;
    long long n_f;
    const double *fdots;
    long long n_trials;
    int nbin;
    int warps;
    long long tb0;  // first 32*R-trial block of this launch
    long long ev_per_split;
    double *scratch;     // [n_tb*R][nbin][32]
    double unit_weight;  // CELL_UNIT: weight of every event
};

__device__ __forceinline__ double lds_f64(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_f64(uint32_t addr, double v) {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v));
}

// Byte offset of the bin inside a [bin][lane] column of (1 << SHIFT)/32-byte cells (bin << SHIFT).
//   WMODE_EDGES  any nbin: upstream pulse_phase(..., to_1=True), then np.histogram's bin search on the edges
//   WMODE_FLOOR  power-of-two nbin: the edges j/nbin are exact, the bin is floor(phase*nbin), phase == 1.0 -> last
//   WMODE_FAST   power-of-two nbin and provably non-negative phases: f and 0.5*fd arrive pre-scaled by
//                nbin << SHIFT (exact), floor() is one round-down add of 1.5*2^52 and the offset is one AND of the
//                low mantissa word -- the same argument as K1's MODE_POW2 (hx_fold.cu)
template <bool HAS_FD, int MODE, int SHIFT>
__device__ __forceinline__ uint32_t wbin_off(double t, double tt, double f, double fdh, int nbin, uint32_t mask,
                                             double step) {
    double x = __dmul_rn(t, f);
    if (HAS_FD) x = __dadd_rn(x, __dmul_rn(tt, fdh));  // fl(fl(0.5*fl(t*t))*fd) == fl(fl(t*t)*(0.5*fd))
    if (MODE == WMODE_FAST) return (uint32_t)__double2loint(__dadd_rd(x, HX_MAGIC)) & mask;
    const double fl = __dsub_rn(__dadd_rd(x, HX_MAGIC), HX_MAGIC);
    const double fr = __dsub_rn(x, fl);  // in [0, 1]
    const double y = __dadd_rd(__dmul_rn(fr, (double)nbin), HX_MAGIC);
    int b = min(__double2loint(y), nbin - 1);
    if (MODE == WMODE_EDGES) {
        // np.linspace(0, 1, nbin + 1)[j] = fl(j * step) for j < nbin, exactly 1.0 for j == nbin.  fl(fr * nbin) is
        // within one ulp of the real product and the edges within one ulp of j / nbin, so the floor is off by at
        // most one bin.
        const double bf = (double)b;
        const double lo = __dmul_rn(bf, step);
        const double hi = (b + 1 == nbin) ? 2.0 : __dmul_rn(__dadd_rn(bf, 1.0), step);
        b += (fr >= hi) - (fr < lo);
    }
    return (uint32_t)b << SHIFT;
}

template <int R, int U, bool HAS_FD, int MODE, int CELL>
__global__ void __launch_bounds__(512, 1) wfold_kernel(const WFoldArgs a) {
    constexpr bool F64 = CELL == CELL_F64;
    constexpr int SHIFT = F64 ? 8 : 7;                        // log2 of the bytes of one [lane] row
    constexpr uint32_t WBYTES = F64 ? 8u : CELL == CELL_U32 ? 4u : 0u;  // bytes per event in the weight stage
    extern __shared__ __align__(16) unsigned char smem[];
    const uint32_t sbase = smem_u32(smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nbin = a.nbin;
    const uint32_t blk = (uint32_t)nbin << SHIFT;  // bytes of one (warp, r) histogram block
    constexpr uint32_t TILE_T = (WTILE_EV + WTILE_PAD) * 8u, TILE_W = (WTILE_EV + WTILE_PAD) * WBYTES;
    const uint32_t tile_off = WSMEM_HDR, wt_off = tile_off + 2u * TILE_T;
    const uint32_t hist_off = wt_off + 2u * TILE_W;
    const uint32_t hist = sbase + hist_off;

    for (uint32_t o = tid * 4; o < (uint32_t)a.warps * R * blk; o += blockDim.x * 4) sts_u32(hist + o, 0u);

    const long long tbl = (long long)blockIdx.x * a.warps + warp;
    const long long tb = a.tb0 + tbl;
    const double scale = MODE == WMODE_FAST ? (double)(nbin << SHIFT) : 1.0;
    const uint32_t mask = (uint32_t)(nbin - 1) << SHIFT;
    const double step = 1.0 / (double)nbin;
    double f[R], fdh[R];
    uint32_t cb[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
        const long long g = (tb * R + r) * 32 + lane;
        const bool valid = g < a.n_trials;
        const long long gi = valid ? g : 0;
        const long long ifd = gi / a.n_f, jf = gi - ifd * a.n_f;
        f[r] = __dmul_rn(a.freqs[jf], scale);
        fdh[r] = HAS_FD ? __dmul_rn(__dmul_rn(0.5, a.fdots[ifd]), scale) : 0.0;
        cb[r] = hist + (uint32_t)(warp * R + r) * blk + (uint32_t)lane * (F64 ? 8u : 4u);
    }

    const long long e0 = (long long)blockIdx.y * a.ev_per_split;
    const long long e1 = min(a.n_ev, e0 + a.ev_per_split);
    const int ntiles = (e1 > e0) ? (int)((e1 - e0 + WTILE_EV - 1) / WTILE_EV) : 0;
    const uint32_t bar[2] = {sbase, sbase + 8};

    // one stream (times or weights) of stage `it & 1`: bulk copy of the 16-byte multiple, plain copy of the rest
    auto issue_one = [&](const char *src, uint32_t dst_off, uint32_t bytes, uint32_t b) -> uint32_t {
        const uint32_t bulk = bytes & ~15u;
        for (uint32_t o = bulk; o < bytes; o += 4)  // ordered before the arrive of this thread below
            *reinterpret_cast<uint32_t *>(smem + dst_off + o) = *reinterpret_cast<const uint32_t *>(src + o);
        if (bulk) bulk_g2s(sbase + dst_off, src, bulk, b);
        return bulk;
    };
    auto issue = [&](int it) {
        const long long s = e0 + (long long)it * WTILE_EV;
        const uint32_t cnt = (uint32_t)min((long long)WTILE_EV, e1 - s);
        const uint32_t b = bar[it & 1];
        const uint32_t tx = ((cnt * 8u) & ~15u) + ((cnt * WBYTES) & ~15u);
        if (tx) mbar_expect_tx(b, tx);
        issue_one(reinterpret_cast<const char *>(a.ev) + s * 8, tile_off + (uint32_t)(it & 1) * TILE_T, cnt * 8u, b);
        if (WBYTES)
            issue_one(reinterpret_cast<const char *>(a.wt) + s * WBYTES, wt_off + (uint32_t)(it & 1) * TILE_W,
                      cnt * WBYTES, b);
        if (!tx) mbar_arrive(b);
    };

    int *done = reinterpret_cast<int *>(smem + 16);  // [2] warps that finished the stage
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
    const int nwarps = a.warps;

    for (int it = 0; it < ntiles; it++) {
        const int buf = it & 1;
        mbar_wait(bar[buf], (uint32_t)(it >> 1) & 1u);
        const int cnt = (int)min((long long)WTILE_EV, e1 - (e0 + (long long)it * WTILE_EV));
        const unsigned char *tp = smem + tile_off + (uint32_t)buf * TILE_T;
        const unsigned char *wp = smem + wt_off + (uint32_t)buf * TILE_W;
        // one group of U events; RAGGED: the last, partial group of a stage (events past cnt add zero to a valid cell)
        auto group = [&](int e, auto ragged_tag) {
            constexpr bool RAGGED = decltype(ragged_tag)::value;
            double t[U], tt[HAS_FD ? U : 1];
            double w[F64 ? U : 1];
            uint32_t wi[F64 ? 1 : U];
#pragma unroll
            for (int u = 0; u < U; u += 2) {
                const double2 v = *reinterpret_cast<const double2 *>(tp + (size_t)(e + u) * 8);
                t[u] = v.x, t[u + 1] = v.y;
                if (F64) {
                    const double2 q = *reinterpret_cast<const double2 *>(wp + (size_t)(e + u) * 8);
                    w[u] = q.x, w[u + 1] = q.y;
                } else if (CELL == CELL_U32) {
                    const uint2 q = *reinterpret_cast<const uint2 *>(wp + (size_t)(e + u) * 4);
                    wi[u] = q.x, wi[u + 1] = q.y;
                } else {
                    wi[u] = 1u, wi[u + 1] = 1u;
                }
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                if (RAGGED && e + u >= cnt) {
                    t[u] = 0.0;
                    if (F64) w[u] = 0.0;
                    else wi[u] = 0u;
                }
                if (HAS_FD) tt[u] = __dmul_rn(t[u], t[u]);
            }
            uint32_t ad[R][U];
#pragma unroll
            for (int u = 0; u < U; u++)
#pragma unroll
                for (int r = 0; r < R; r++)
                    ad[r][u] = cb[r] + wbin_off<HAS_FD, MODE, SHIFT>(t[u], HAS_FD ? tt[u] : 0.0, f[r], fdh[r], nbin,
                                                                     mask, step);
            if constexpr (!F64) {
#pragma unroll
                for (int u = 0; u < U; u++)
#pragma unroll
                    for (int r = 0; r < R; r++) red_shared_add_u32(ad[r][u], wi[u]);
            } else {
#pragma unroll
                for (int r = 0; r < R; r++) {
                    double h[U];
#pragma unroll
                    for (int u = 0; u < U; u++) h[u] = lds_f64(ad[r][u]);
#pragma unroll
                    for (int u = 0; u < U; u++) {
                        // the cell's current value: what a previous event of this group just stored, if any
#pragma unroll
                        for (int v = 0; v < u; v++)
                            if (ad[r][v] == ad[r][u]) h[u] = h[v];
                        h[u] = __dadd_rn(h[u], w[u]);
                        sts_f64(ad[r][u], h[u]);
                    }
                }
            }
        };
        int e = 0;
        for (; e + U <= cnt; e += U) group(e, std::false_type{});  // straight-line main loop: no per-event bounds checks
        if (e < cnt) group(e, std::true_type{});
        // stage hand-back without a block barrier (as in K1): the last warp to finish `buf` refills it
        __syncwarp();
        if (lane == 0) {
            __threadfence_block();
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

    // partial profiles -> global scratch [trial/32][bin][32], coalesced across the warp
#pragma unroll
    for (int r = 0; r < R; r++) {
        double *dst = a.scratch + ((size_t)(tbl * R + r) * nbin) * 32 + lane;
        for (int j = 0; j < nbin; j++) {
            double c;
            if (F64) {
                c = lds_f64(cb[r] + ((uint32_t)j << SHIFT));
            } else {
                c = (double)lds_u32(cb[r] + ((uint32_t)j << SHIFT));
                if (CELL == CELL_UNIT) c = __dmul_rn(c, a.unit_weight);
            }
            if (c != 0.0) atomicAdd(dst + (size_t)j * 32, c);
        }
    }
}

// One pass over the weights: uint32 copies when every weight is a non-negative integer (flags[0] = 1 otherwise),
// flags[1] = the 64-bit sum of the integer weights (the u32 cells need it below 2^32), keep[i] = (w[i] != 0) and
// flags[2] = the number of non-zero weights (a light curve at fine time resolution is mostly empty bins).
__global__ void wfold_intweights_kernel(const double *__restrict__ w, long long n, uint32_t *__restrict__ out,
                                        unsigned char *__restrict__ keep, unsigned long long *flags) {
    unsigned long long sum = 0, nnz = 0;
    bool bad = false;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double v = w[i];
        const bool ok = v >= 0.0 && v < 4294967296.0 && v == floor(v);
        bad |= !ok;
        const uint32_t u = ok ? (uint32_t)v : 0u;
        out[i] = u;
        sum += u;
        const bool k = !(v == 0.0);  // NaN weights are kept (and poison the profile, as upstream)
        keep[i] = k;
        nnz += k;
    }
    for (int o = 16; o; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        nnz += __shfl_xor_sync(0xffffffffu, nnz, o);
    }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicExch(flags, 1ull);
    if ((threadIdx.x & 31) == 0 && sum) atomicAdd(flags + 1, sum);
    if ((threadIdx.x & 31) == 0 && nnz) atomicAdd(flags + 2, nnz);
}

// stable stream compaction of `in[n]` by `keep` into `out` (cub::DeviceSelect::Flagged); *d_count receives the count
template <typename T>
static int compact_flagged(const T *in, const unsigned char *keep, T *out, int *d_count, int n, cudaStream_t st) {
    size_t tmp_bytes = 0;
    HX_CUDA_TRY(cub::DeviceSelect::Flagged(nullptr, tmp_bytes, in, keep, out, d_count, n, st));
    AsyncBuf tmp(st);
    HX_CUDA_TRY(tmp.alloc(tmp_bytes));
    HX_CUDA_TRY(cub::DeviceSelect::Flagged(tmp.p, tmp_bytes, in, keep, out, d_count, n, st));
    count_launch(2);
    return HX_OK;
}

// ---- exposure correction -------------------------------------------------------------------------
// double-double (hi + lo, |lo| <= ulp(hi)/2) stands in for upstream's np.longdouble: every quantity
// below is a sum/difference of a handful of fp64 numbers, which double-double represents exactly.
struct dd {
    double hi, lo;
};
__device__ __forceinline__ dd two_sum(double a, double b) {
    const double s = __dadd_rn(a, b), bb = __dsub_rn(s, a);
    return {s, __dadd_rn(__dsub_rn(a, __dsub_rn(s, bb)), __dsub_rn(b, bb))};
}
__device__ __forceinline__ dd dd_add(dd a, dd b) {
    dd s = two_sum(a.hi, b.hi);
    const dd t = two_sum(a.lo, b.lo);
    s.lo = __dadd_rn(s.lo, t.hi);
    s = two_sum(s.hi, s.lo);
    s.lo = __dadd_rn(s.lo, t.lo);
    return two_sum(s.hi, s.lo);
}
__device__ __forceinline__ dd dd_neg(dd a) { return {-a.hi, -a.lo}; }
__device__ __forceinline__ bool dd_lt(dd a, dd b) { return a.hi < b.hi || (a.hi == b.hi && a.lo < b.lo); }
__device__ __forceinline__ bool dd_le(dd a, dd b) { return a.hi < b.hi || (a.hi == b.hi && a.lo <= b.lo); }
__device__ __forceinline__ dd dd_max(dd a, dd b) { return dd_lt(a, b) ? b : a; }
__device__ __forceinline__ dd dd_min(dd a, dd b) { return dd_lt(b, a) ? b : a; }

// pulse_phase(x, f, fd, to_1=False) in fp64, as upstream evaluates it on the GTI edges
__device__ __forceinline__ double phase_raw(double t, double f, double fdh, bool has_fd) {
    double x = __dmul_rn(t, f);
    if (has_fd) x = __dadd_rn(x, __dmul_rn(__dmul_rn(t, t), fdh));
    return x;
}

// stingray.pulse.pulsar.phase_exposure(start_phase, stop_phase, 1, nbin, gti=gti_phases)[j0 .. j0 + EB) before the final
// normalisation, for one trial: the per-bin sequence of fp64 accumulations is upstream's (per GTI: + nper / nbin, then
// + the overlap with each of the one or two fractional limits).  EB bins are kept in registers per pass over the GTIs,
// so the per-GTI quantities are set up nbin / EB times per trial instead of nbin times.
constexpr int EB = 8;
__device__ __forceinline__ void expo_bins(double (&e)[EB], int j0, int nbin, const double *edges, const double *gti, int n_gti,
                                          double f, double fdh, bool has_fd, double start_ph, double stop_ph) {
    dd p0[EB], p1[EB];
#pragma unroll
    for (int k = 0; k < EB; k++) {
        e[k] = 0.0;
        const int j = min(j0 + k, nbin - 1);
        p0[k] = dd{edges[j], 0.0}, p1[k] = dd{edges[j + 1], 0.0};
    }
    for (int g = 0; g < n_gti; g++) {
        double g0 = phase_raw(gti[2 * g], f, fdh, has_fd), g1 = phase_raw(gti[2 * g + 1], f, fdh, has_fd);
        if (!(g0 < stop_ph && g1 > start_ph)) continue;
        if (g0 < start_ph) g0 = start_ph;
        if (g1 > stop_ph) g1 = stop_ph;
        const dd length = two_sum(g1, -g0);  // exact
        // nper = int(nraw): truncation toward zero of the exact value
        double nper = trunc(length.hi);
        if (nper == length.hi) {
            if (length.hi > 0.0 && length.lo < 0.0) nper -= 1.0;
            if (length.hi < 0.0 && length.lo > 0.0) nper += 1.0;
        }
        const double whole = __ddiv_rn(nper, (double)nbin);
        const double sp = __dsub_rn(g0, trunc(g0));                      // fmod(g0, 1), exact
        const dd frac = two_sum(__dsub_rn(length.hi, nper), length.lo);  // nraw - nper, exact
        const dd endp = dd_add(frac, dd{sp, 0.0});
        const dd one = {1.0, 0.0};
        dd l0[2], l1[2];
        int nl = 1;
        if (dd_lt(one, endp)) {
            l0[0] = dd{0.0, 0.0}, l1[0] = dd_add(endp, dd{-1.0, 0.0});
            l0[1] = dd{sp, 0.0}, l1[1] = one;
            nl = 2;
        } else {
            l0[0] = dd{sp, 0.0}, l1[0] = endp;
        }
#pragma unroll
        for (int k = 0; k < EB; k++) {
            e[k] = __dadd_rn(e[k], whole);
            for (int q = 0; q < nl; q++) {
                if (!(dd_le(p0[k], l1[q]) && dd_le(l0[q], p1[k]))) continue;
                const dd w = dd_add(dd_min(p1[k], l1[q]), dd_neg(dd_max(p0[k], l0[q])));
                const dd s = two_sum(e[k], w.hi);
                e[k] = __dadd_rn(s.hi, __dadd_rn(s.lo, w.lo));
            }
        }
    }
}

// scratch[slot][j] /= expo[j] / max(expo), one thread per trial slot (lane-minor layout => coalesced)
__global__ void __launch_bounds__(128) wfold_expo_kernel(double *__restrict__ scratch, double *__restrict__ expo,
                                                         long long l0, long long n, long long n_f,
                                                         long long n_trials, int nbin,
                                                         const double *__restrict__ edges,
                                                         const double *__restrict__ gti, int n_gti,
                                                         const double *__restrict__ freqs,
                                                         const double *__restrict__ fdots, int has_fd,
                                                         double start_time, double stop_time) {
    const long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n) return;
    const long long g = l0 + l;
    if (g >= n_trials) return;
    const long long ifd = g / n_f, jf = g - ifd * n_f;
    const double f = freqs[jf], fdh = has_fd ? __dmul_rn(0.5, fdots[ifd]) : 0.0;
    const double start_ph = phase_raw(start_time, f, fdh, has_fd), stop_ph = phase_raw(stop_time, f, fdh, has_fd);
    const size_t base = (size_t)(l >> 5) * nbin * 32 + (l & 31);
    double emax = -INFINITY;
    for (int j0 = 0; j0 < nbin; j0 += EB) {
        double e[EB];
        expo_bins(e, j0, nbin, edges, gti, n_gti, f, fdh, has_fd != 0, start_ph, stop_ph);
#pragma unroll
        for (int k = 0; k < EB; k++)
            if (j0 + k < nbin) {
                expo[base + (size_t)(j0 + k) * 32] = e[k];
                emax = fmax(emax, e[k]);
            }
    }
    for (int j = 0; j < nbin; j++) {
        const double en = __ddiv_rn(expo[base + (size_t)j * 32], emax);
        scratch[base + (size_t)j * 32] = __ddiv_rn(scratch[base + (size_t)j * 32], en);
    }
}

// fp64 scratch [trial/32][bin][32] -> stats (+ optional [trial][nbin] profiles)
__global__ void __launch_bounds__(256) wfold_stat_kernel(const double *__restrict__ scratch, long long l0, long long n,
                                                         long long n_trials, int nbin, int nharm,
                                                         const double *__restrict__ tab, double *__restrict__ stats,
                                                         double *__restrict__ profiles) {
    extern __shared__ double tabs[];
    for (int i = threadIdx.x; i < 2 * nbin; i += blockDim.x) tabs[i] = tab[i];
    __syncthreads();
    const long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n) return;
    const long long g = l0 + l;
    if (g >= n_trials) return;
    const double *src = scratch + (size_t)(l >> 5) * nbin * 32 + (l & 31);
    stats[g] = profile_stat([&](int j) { return src[(size_t)j * 32]; }, nbin, nharm, tabs, tabs + nbin);
    if (profiles)
        for (int j = 0; j < nbin; j++) profiles[(size_t)g * nbin + j] = src[(size_t)j * 32];
}

typedef void (*wfold_fn_t)(const WFoldArgs);
template <int R, int U, int CELL>
static wfold_fn_t pick_wfold_M(bool has_fd, int mode) {
    if (has_fd)
        return mode == WMODE_FAST    ? wfold_kernel<R, U, true, WMODE_FAST, CELL>
               : mode == WMODE_FLOOR ? wfold_kernel<R, U, true, WMODE_FLOOR, CELL>
                                     : wfold_kernel<R, U, true, WMODE_EDGES, CELL>;
    return mode == WMODE_FAST    ? wfold_kernel<R, U, false, WMODE_FAST, CELL>
           : mode == WMODE_FLOOR ? wfold_kernel<R, U, false, WMODE_FLOOR, CELL>
                                 : wfold_kernel<R, U, false, WMODE_EDGES, CELL>;
}
// fp64 cells: (R, U) in {1, 2} x {2, 4}; u32 cells: (1, 4), (2, 2), (4, 2)
template <int CELL>
static wfold_fn_t pick_wfold_u32(int R, bool has_fd, int mode) {
    // events per register group: 8 independent chains per thread (R * U) now that the main loop is straight-line
    return R == 4 ? pick_wfold_M<4, 2, CELL>(has_fd, mode)
           : R == 2 ? pick_wfold_M<2, 4, CELL>(has_fd, mode)
                    : pick_wfold_M<1, 8, CELL>(has_fd, mode);
}
static wfold_fn_t pick_wfold(int cell, int R, int U, bool has_fd, int mode) {
    if (cell == CELL_U32) return pick_wfold_u32<CELL_U32>(R, has_fd, mode);
    if (cell == CELL_UNIT) return pick_wfold_u32<CELL_UNIT>(R, has_fd, mode);
    if (R == 2) return U == 2 ? pick_wfold_M<2, 2, CELL_F64>(has_fd, mode) : pick_wfold_M<2, 4, CELL_F64>(has_fd, mode);
    return U == 2 ? pick_wfold_M<1, 2, CELL_F64>(has_fd, mode) : pick_wfold_M<1, 4, CELL_F64>(has_fd, mode);
}

static int g_wfold_R = 0, g_wfold_U = 0, g_wfold_W = 0;  // tuning hooks (hx_wfold_set_tuning)
static int g_wfold_force_generic = 0;                    // bit 0: no WMODE_FAST; bit 1: fp64 cells even for integers

static cudaEvent_t g_wev0[64], g_wev1[64];
static bool g_wev_ok[64] = {false};
static int g_wlast_device = -1;
static int g_wlast_cell = -1;
static long long g_wlast_n_used = 0;  // events actually folded by the last call (after dropping zero weights)

int wfold_search_dev_impl(const double *d_times, const double *d_weights, double unit_weight, int64_t n_ev,
                          const double *d_freqs, int64_t n_f, const double *d_fdots, int64_t n_fd, int nbin,
                          int nharm, const double *h_gti, int64_t n_gti, double start_time, double stop_time,
                          double *d_stats, double *d_profiles, cudaStream_t st) {
    std::lock_guard<std::recursive_mutex> api_lock(api_mutex());
    if (!d_times || !d_freqs || !d_fdots || !d_stats) {
        set_error("hx_wfold_search: null pointer argument");
        return HX_ERR_ARG;
    }
    if (n_ev < 0 || n_f < 0 || n_fd < 0 || nbin < 2 || nbin > HX_MAX_NBIN || nharm < 0 || nharm > HX_MAX_NHARM ||
        n_gti < 0 || (n_gti > 0 && !h_gti)) {
        set_error("hx_wfold_search: bad sizes (n_ev=%lld n_f=%lld n_fd=%lld nbin=%d nharm=%d n_gti=%lld)",
                  (long long)n_ev, (long long)n_f, (long long)n_fd, nbin, nharm, (long long)n_gti);
        return HX_ERR_ARG;
    }
    if ((reinterpret_cast<uintptr_t>(d_times) & 15) != 0 || (reinterpret_cast<uintptr_t>(d_weights) & 15) != 0) {
        set_error("hx_wfold_search_dev: d_times / d_weights must be 16-byte aligned");
        return HX_ERR_ARG;
    }
    const long long n_trials = (long long)n_f * n_fd;
    if (n_trials == 0) return HX_OK;
    int device;
    HX_CUDA_TRY(cudaGetDevice(&device));
    DeviceInfo di;
    int rc = get_device_info(device, &di);
    if (rc) return rc;

    double r6[6];
    if ((rc = device_ranges(d_times, n_ev, d_freqs, n_f, d_fdots, n_fd, r6, st)) != HX_OK) return rc;
    const double tmin = r6[0], tmax = r6[1], fmin = r6[2], fmax = r6[3], fdmin = r6[4], fdmax = r6[5];
    // fdot == 0 everywhere: the fdot term adds +0.0 upstream, skip it
    const bool has_fd = !(fdmin == 0.0 && fdmax == 0.0);
    const double tabs_ = std::max(std::fabs(tmin), std::fabs(tmax));
    const double phmax = tabs_ * std::max(std::fabs(fmin), std::fabs(fmax)) +
                         0.5 * tabs_ * tabs_ * std::max(std::fabs(fdmin), std::fabs(fdmax));
    if (!(phmax < 1125899906842624.0 /* 2^50 */)) {
        set_error("hx_wfold_search: |phase| up to %.3g cycles exceeds the fp64 folding range (2^50)", phmax);
        return HX_ERR_RANGE;
    }
    // ---- cell type: u32 counters when the weights are integers (or absent), fp64 cells otherwise
    AsyncBuf b_wi(st), b_flags(st), b_keep(st), b_tc(st), b_wc(st);
    int cell = d_weights ? CELL_F64 : CELL_UNIT;
    const void *d_wt = d_weights;
    const double *d_ev = d_times;  // the event stream the fold kernel reads (compacted below when most weights are 0)
    int64_t n_use = n_ev;
    if (d_weights && n_ev > 0) {
        HX_CUDA_TRY(b_wi.alloc((size_t)n_ev * sizeof(uint32_t)));
        HX_CUDA_TRY(b_keep.alloc((size_t)n_ev));
        HX_CUDA_TRY(b_flags.alloc(4 * sizeof(unsigned long long)));
        HX_CUDA_TRY(cudaMemsetAsync(b_flags.p, 0, 4 * sizeof(unsigned long long), st));
        wfold_intweights_kernel<<<(int)std::min<long long>((n_ev + 255) / 256, 1184), 256, 0, st>>>(
            d_weights, n_ev, b_wi.as<uint32_t>(), b_keep.as<unsigned char>(), b_flags.as<unsigned long long>());
        HX_CHECK_LAUNCH("wfold_intweights_kernel");
        unsigned long long hf[3];
        HX_CUDA_TRY(cudaMemcpyAsync(hf, b_flags.p, sizeof(hf), cudaMemcpyDeviceToHost, st));
        HX_CUDA_TRY(cudaStreamSynchronize(st));
        if (hf[0] == 0 && hf[1] < 4294967296ull && !(g_wfold_force_generic & 2)) cell = CELL_U32, d_wt = b_wi.p;
        // Adding a zero weight changes nothing (upstream's np.histogram adds it all the same): when at least a
        // quarter of the weights are zero -- empty light-curve bins -- fold only the others.  The compaction is
        // stable, so fp64 sums keep their order.
        const long long nnz = (long long)hf[2];
        if (nnz < n_ev - n_ev / 4 && n_ev < 2147483647LL && !(g_wfold_force_generic & 4)) {
            int *d_count = reinterpret_cast<int *>(b_flags.as<unsigned long long>() + 3);
            HX_CUDA_TRY(b_tc.alloc((size_t)std::max<long long>(nnz, 2) * 8));
            HX_CUDA_TRY(b_wc.alloc((size_t)std::max<long long>(nnz, 2) * 8));
            if ((rc = compact_flagged(d_times, b_keep.as<unsigned char>(), b_tc.as<double>(), d_count, (int)n_ev, st))) return rc;
            if (cell == CELL_U32)
                rc = compact_flagged(b_wi.as<uint32_t>(), b_keep.as<unsigned char>(), b_wc.as<uint32_t>(), d_count, (int)n_ev, st);
            else
                rc = compact_flagged(d_weights, b_keep.as<unsigned char>(), b_wc.as<double>(), d_count, (int)n_ev, st);
            if (rc) return rc;
            d_ev = b_tc.as<double>(), d_wt = b_wc.p, n_use = nnz;
        }
    }
    g_wlast_n_used = n_use;
    const bool f64 = cell == CELL_F64;
    const int shift = f64 ? 8 : 7;
    const bool pow2 = (nbin & (nbin - 1)) == 0;
    // same conditions as K1's fast path: non-negative phases (ph - floor(ph) is then exact) inside the magic-add range
    const bool nonneg = tmin >= 0 && fmin >= 0 && (fdmin >= 0 || fmin + 0.5 * tmax * fdmin * (1 + 1e-12) > 0);
    int mode = !pow2 ? WMODE_EDGES
               : (nonneg && phmax * (double)(nbin << shift) < 1125899906842624.0) ? WMODE_FAST
                                                                                  : WMODE_FLOOR;
    if ((g_wfold_force_generic & 1) && mode == WMODE_FAST) mode = WMODE_FLOOR;
    // ---- plan.  A column costs (1 << shift) bytes per bin and warp; the fewer warps fit, the more independent
    // work per thread is needed: R trials, and for fp64 cells register groups of U events (U cell loads in flight).
    int R, U;
    if (f64) {
        // measured on B200 after the straight-line loop (profiles/r02_sweep_wfold.txt): one trial per thread above 32 bins
        R = nbin <= 32 ? 2 : 1, U = nbin <= 32 ? 2 : 4;
        if (g_wfold_R == 1 || g_wfold_R == 2) R = g_wfold_R;
        if (g_wfold_U == 2 || g_wfold_U == 4) U = g_wfold_U;
    } else {
        R = nbin <= 32 ? 4 : nbin <= 64 ? 2 : 1;  // measured on B200 (tools/sweep_wfold.py)
        if (g_wfold_R) R = g_wfold_R;
        U = R == 1 ? 4 : 2;
    }
    const uint32_t wbytes = f64 ? 8u : cell == CELL_U32 ? 4u : 0u;
    const uint32_t fixed = WSMEM_HDR + 2u * (WTILE_EV + WTILE_PAD) * (8u + wbytes);
    const uint32_t blk = (uint32_t)nbin << shift;
    int W = 0;
    for (; R >= 1; R >>= 1) {
        W = std::min<int>(16, (int)(((uint32_t)di.smem_optin - fixed) / ((uint32_t)R * blk)));
        if (W >= 1) break;
    }
    if (W < 1) {
        set_error("hx_wfold_search: nbin=%d does not fit in shared memory (%s cells: nbin <= %d)", nbin,
                  f64 ? "fp64" : "u32", (int)(((uint32_t)di.smem_optin - fixed) >> shift));
        return HX_ERR_ARG;
    }
    if (!f64) U = R == 1 ? 4 : 2;
    {  // k CTAs of W warps per SM: most resident warps, then fewest CTAs
        int best_warps = 0;
        for (int k = 1; k <= 8; k++) {
            const long long room = (long long)di.smem_per_sm / k - (long long)fixed - di.reserved_smem;
            const int w = (int)std::min<long long>(std::min<long long>(16, room / ((long long)R * blk)),
                                                   ((long long)di.smem_optin - fixed) / ((long long)R * blk));
            if (w >= 1 && k * w > best_warps) best_warps = k * w, W = w;
        }
    }
    if (g_wfold_W) W = std::min(W, g_wfold_W);
    // warps per CTA: 1, 2 or a multiple of 4, so that the SM's four schedulers stay evenly loaded (measured on K1 / K2: 3-, 6-
    // and 11-warp CTAs lose 8-30 %)
    auto balanced = [](int w) { return w >= 4 ? w / 4 * 4 : w == 3 ? 2 : w; };
    if (!g_wfold_W) W = balanced(W);
    const long long n_tb = (n_trials + 32LL * R - 1) / (32LL * R);
    {  // very few trials: smaller CTAs so that every SM gets one (the event splitting below multiplies the CTAs otherwise)
        const long long want = (n_tb + di.sm_count - 1) / di.sm_count;
        if (want < W && want <= 2) W = (int)std::max<long long>(1, want);
    }
    const uint32_t smem_bytes = fixed + (uint32_t)W * R * blk;
    const int ctas_per_sm = std::max(1, di.smem_per_sm / (int)(smem_bytes + di.reserved_smem));

    const size_t scratch_per_tb = (size_t)R * nbin * 32 * sizeof(double);
    size_t scratch_cap = (size_t)1 << 30;
    if (const char *e = getenv("HX_FOLD_SCRATCH_CAP_KB")) {
        const long long kb = atoll(e);
        if (kb > 0) scratch_cap = (size_t)kb << 10;
    }
    const long long tb_per_batch = std::max<long long>(W, (long long)(scratch_cap / scratch_per_tb) / W * W);
    const long long n_tb_first = std::min(n_tb, tb_per_batch);
    const long long n_cta_x = (n_tb_first + W - 1) / W;
    const long long slots = (long long)di.sm_count * ctas_per_sm;
    // split the events so that the grid fills at least ~4 waves of CTA slots
    long long ev_splits = 1;
    if (n_use > 0 && n_cta_x < 4 * slots) {
        ev_splits = std::min<long long>((4 * slots + n_cta_x - 1) / n_cta_x, std::max<long long>(1, n_use / (4 * WTILE_EV)));
        ev_splits = std::min<long long>(std::max<long long>(ev_splits, 1), 65535);
    }
    long long ev_per_split = (n_use + ev_splits - 1) / ev_splits;
    ev_per_split = std::max<long long>(WTILE_EV, (ev_per_split + WTILE_EV - 1) / WTILE_EV * WTILE_EV);
    ev_splits = std::max<long long>(1, (n_use + ev_per_split - 1) / ev_per_split);

    // ---- device temporaries
    AsyncBuf b_tab(st), b_scratch(st), b_expo(st), b_edges(st), b_gti(st);
    std::vector<double> tab, edges((size_t)nbin + 1);
    host_tables(nbin, tab);
    {  // np.linspace(0, 1, nbin + 1): arange * step + start, last element forced to stop
        const double step = 1.0 / nbin;
        for (int i = 0; i <= nbin; i++) edges[i] = i * step + 0.0;
        edges[nbin] = 1.0;
    }
    HX_CUDA_TRY(b_tab.alloc(tab.size() * sizeof(double)));
    HX_CUDA_TRY(cudaMemcpyAsync(b_tab.p, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    HX_CUDA_TRY(b_edges.alloc(edges.size() * sizeof(double)));
    HX_CUDA_TRY(cudaMemcpyAsync(b_edges.p, edges.data(), edges.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    const size_t sbytes = (size_t)((std::min(n_tb, tb_per_batch) + W - 1) / W * W) * scratch_per_tb;
    HX_CUDA_TRY(b_scratch.alloc(sbytes));
    if (n_gti > 0) {
        HX_CUDA_TRY(b_expo.alloc(sbytes));
        HX_CUDA_TRY(b_gti.alloc((size_t)n_gti * 16));
        HX_CUDA_TRY(cudaMemcpyAsync(b_gti.p, h_gti, (size_t)n_gti * 16, cudaMemcpyHostToDevice, st));
    }

    WFoldArgs a{};
    a.ev = d_ev, a.wt = d_wt, a.n_ev = n_use, a.freqs = d_freqs, a.n_f = n_f, a.fdots = d_fdots;
    a.n_trials = n_trials, a.nbin = nbin, a.warps = W, a.ev_per_split = ev_per_split;
    a.scratch = b_scratch.as<double>(), a.unit_weight = unit_weight;
    wfold_fn_t fn = pick_wfold(cell, R, U, has_fd, mode);
    HX_CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    if (!g_wev_ok[device]) {
        HX_CUDA_TRY(cudaEventCreate(&g_wev0[device]));
        HX_CUDA_TRY(cudaEventCreate(&g_wev1[device]));
        g_wev_ok[device] = true;
    }
    for (long long tb0 = 0; tb0 < n_tb; tb0 += tb_per_batch) {
        const long long n_tb_b = std::min(tb_per_batch, n_tb - tb0);
        const long long grid_x = (n_tb_b + W - 1) / W;
        // the grid may overhang the batch by < W trial blocks: the scratch allocation is W-padded via tb_per_batch
        HX_CUDA_TRY(cudaMemsetAsync(b_scratch.p, 0, (size_t)(grid_x * W) * scratch_per_tb, st));
        a.tb0 = tb0;
        if (tb0 == 0) HX_CUDA_TRY(cudaEventRecord(g_wev0[device], st));
        dim3 grid((unsigned)grid_x, (unsigned)ev_splits);
        fn<<<grid, W * 32, smem_bytes, st>>>(a);
        HX_CHECK_LAUNCH("wfold_kernel");
        if (tb0 + tb_per_batch >= n_tb) HX_CUDA_TRY(cudaEventRecord(g_wev1[device], st));
        const long long l0 = tb0 * 32 * R, ng = n_tb_b * 32 * R;
        if (n_gti > 0) {
            wfold_expo_kernel<<<(unsigned)((ng + 127) / 128), 128, 0, st>>>(
                b_scratch.as<double>(), b_expo.as<double>(), l0, ng, n_f, n_trials, nbin, b_edges.as<double>(),
                b_gti.as<double>(), (int)n_gti, d_freqs, d_fdots, has_fd ? 1 : 0, start_time, stop_time);
            HX_CHECK_LAUNCH("wfold_expo_kernel");
        }
        wfold_stat_kernel<<<(unsigned)((ng + 255) / 256), 256, 2 * nbin * sizeof(double), st>>>(
            b_scratch.as<double>(), l0, ng, n_trials, nbin, nharm, b_tab.as<double>(), d_stats, d_profiles);
        HX_CHECK_LAUNCH("wfold_stat_kernel");
    }
    g_wlast_device = device, g_wlast_cell = cell;
    HX_CUDA_TRY(cudaStreamSynchronize(st));
    return HX_OK;
}

}  // namespace hx

extern "C" {

int hx_wfold_last_kernel_ms(double *ms, int *cell_kind, int64_t *n_events_folded) {
    if (!ms || hx::g_wlast_device < 0) {
        hx::set_error("hx_wfold_last_kernel_ms: no weighted fold kernel has been launched");
        return HX_ERR_ARG;
    }
    const int d = hx::g_wlast_device;
    float f = 0.f;
    HX_CUDA_TRY(cudaEventSynchronize(hx::g_wev1[d]));
    HX_CUDA_TRY(cudaEventElapsedTime(&f, hx::g_wev0[d], hx::g_wev1[d]));
    *ms = f;
    if (cell_kind) *cell_kind = hx::g_wlast_cell;
    if (n_events_folded) *n_events_folded = hx::g_wlast_n_used;
    return HX_OK;
}

int hx_wfold_set_tuning(int trials_per_thread, int events_per_group, int warps_per_cta, int force_generic) {
    hx::g_wfold_R = (trials_per_thread == 1 || trials_per_thread == 2 || trials_per_thread == 4) ? trials_per_thread : 0;
    hx::g_wfold_U = (events_per_group == 2 || events_per_group == 4) ? events_per_group : 0;
    hx::g_wfold_W = (warps_per_cta >= 1 && warps_per_cta <= 16) ? warps_per_cta : 0;
    hx::g_wfold_force_generic = force_generic & 7;  // bit 0: never WMODE_FAST; bit 1: fp64 cells even for integers; bit 2: keep zero weights
    return HX_OK;
}

int hx_wfold_search_dev(const double *d_times, const double *d_weights, double unit_weight, int64_t n_ev,
                        const double *d_freqs, int64_t n_f, const double *d_fdots, int64_t n_fd, int nbin, int nharm,
                        const double *gti, int64_t n_gti, double start_time, double stop_time, double *d_stats,
                        double *d_profiles, void *stream) {
    return hx::wfold_search_dev_impl(d_times, d_weights, unit_weight, n_ev, d_freqs, n_f, d_fdots, n_fd, nbin, nharm,
                                     gti, n_gti, start_time, stop_time, d_stats, d_profiles,
                                     reinterpret_cast<cudaStream_t>(stream));
}

int hx_wfold_search(const double *times, const double *weights, double unit_weight, int64_t n_ev, const double *freqs,
                    int64_t n_f, const double *fdots, int64_t n_fd, int nbin, int nharm, const double *gti,
                    int64_t n_gti, double start_time, double stop_time, double *stats, double *profiles, int device) {
    if (!times || !freqs || !fdots || !stats) {
        hx::set_error("hx_wfold_search: null pointer argument");
        return HX_ERR_ARG;
    }
    if (n_ev < 0 || n_f < 0 || n_fd < 0 || nbin < 2 || nbin > HX_MAX_NBIN) {
        hx::set_error("hx_wfold_search: bad sizes");
        return HX_ERR_ARG;
    }
    hx::DeviceGuard dev_guard;  // declared before the device buffers: they are freed before the device is restored
    HX_CUDA_TRY(dev_guard.enter(device));
    const size_t n_tr = (size_t)n_f * n_fd;
    if (n_tr == 0) return HX_OK;
    hx::DevBuf dt, dw, df, dfd, ds, dp;
    HX_CUDA_TRY(dt.alloc((size_t)n_ev * 8));
    if (weights) HX_CUDA_TRY(dw.alloc((size_t)n_ev * 8));
    HX_CUDA_TRY(df.alloc((size_t)n_f * 8));
    HX_CUDA_TRY(dfd.alloc((size_t)n_fd * 8));
    HX_CUDA_TRY(ds.alloc(n_tr * 8));
    if (profiles) HX_CUDA_TRY(dp.alloc(n_tr * nbin * 8));
    HX_CUDA_TRY(cudaMemcpy(dt.p, times, (size_t)n_ev * 8, cudaMemcpyHostToDevice));
    if (weights) HX_CUDA_TRY(cudaMemcpy(dw.p, weights, (size_t)n_ev * 8, cudaMemcpyHostToDevice));
    HX_CUDA_TRY(cudaMemcpy(df.p, freqs, (size_t)n_f * 8, cudaMemcpyHostToDevice));
    HX_CUDA_TRY(cudaMemcpy(dfd.p, fdots, (size_t)n_fd * 8, cudaMemcpyHostToDevice));
    int rc = hx::wfold_search_dev_impl(dt.as<double>(), weights ? dw.as<double>() : nullptr, unit_weight, n_ev,
                                       df.as<double>(), n_f, dfd.as<double>(), n_fd, nbin, nharm, gti, n_gti,
                                       start_time, stop_time, ds.as<double>(), profiles ? dp.as<double>() : nullptr, 0);
    if (rc) return rc;
    HX_CUDA_TRY(cudaMemcpy(stats, ds.p, n_tr * 8, cudaMemcpyDeviceToHost));
    if (profiles) HX_CUDA_TRY(cudaMemcpy(profiles, dp.p, n_tr * nbin * 8, cudaMemcpyDeviceToHost));
    return HX_OK;
}

}  // extern "C"
