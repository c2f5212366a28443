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
// Design: K1's layout with fp64 cells.  lane = trial; every thread owns R trials and, for each, a PRIVATE
// nbin-bin fp64 histogram column in shared memory laid out [bin][lane] (byte address = bin*256 + lane*8): a
// warp-wide 64-bit access touches 32 distinct 8-byte slots = 2 conflict-free wavefronts, and because a
// column belongs to one thread the update is a plain LDS.64 + DADD + STS.64 -- no atomics.  Event times and
// weights stream through a 2-stage cp.async.bulk (TMA engine) + mbarrier pipeline and are broadcast to all
// threads.  Bin index: searchsorted semantics of np.histogram on the explicit edges; for power-of-two nbin
// the edges j/nbin are exact and the bin is floor(phase*nbin) (phase == 1.0 -> last bin, whose right edge is
// closed); otherwise floor(phase*nbin) is corrected against the edge table (fl(j*fl(1/nbin)), as
// np.linspace computes it).  Partial profiles always go to a global fp64 scratch [trial/32][bin][32]
// (RED.ADD.F64, coalesced); wfold_expo_kernel divides it by the per-trial exposure (double-double
// arithmetic standing in for upstream's long double), wfold_stat_kernel reduces it to chi^2 / Z^2_n.
//
// Accumulation order differs from np.histogram's (block-wise sort + cumsum + diff), so weighted profiles
// agree to fp64 rounding (exactly, when the weights are integers -- light-curve counts -- and the sums
// stay below 2^53); statistics within the 1e-9 tolerance.
#include <algorithm>
#include <cstdlib>

#include "hx_common.cuh"

namespace hx {

constexpr int WTILE_EV = 1024;       // events per pipeline stage
constexpr uint32_t WSMEM_HDR = 16;   // two mbarriers

struct WFoldArgs {
    const double *ev;
    const double *wt;  // null: unit weights
    long long n_ev;
    const double *freqs;
    long long n_f;
    const double *fdots;
    long long n_trials;
    int nbin;
    int warps;
    long long tb0;  // first 32*R-trial block of this launch
    long long ev_per_split;
    double *scratch;      // [n_tb*R][nbin][32]
    const double *edges;  // [nbin + 1]
    double unit_weight;   // weight of every event when wt == null
};

__device__ __forceinline__ double lds_f64(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_f64(uint32_t addr, double v) {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v));
}

// upstream pulse_phase(..., to_1=True) followed by np.histogram's bin search
template <bool HAS_FD, bool POW2>
__device__ __forceinline__ int wbin(double t, double tt, double f, double fdh, int nbin, const double *edges) {
    double x = __dmul_rn(t, f);
    if (HAS_FD) x = __dadd_rn(x, __dmul_rn(tt, fdh));  // fl(fl(0.5*fl(t*t))*fd) == fl(fl(t*t)*(0.5*fd))
    const double fl = __dsub_rn(__dadd_rd(x, HX_MAGIC), HX_MAGIC);
    const double fr = __dsub_rn(x, fl);  // in [0, 1]
    int b = __double2loint(__dadd_rd(__dmul_rn(fr, (double)nbin), HX_MAGIC));
    b = min(b, nbin - 1);
    if (!POW2) {
        while (b > 0 && fr < edges[b]) b--;
        while (b < nbin - 1 && fr >= edges[b + 1]) b++;
    }
    return b;
}

template <int R, bool HAS_FD, bool POW2, bool HAS_W>
__global__ void __launch_bounds__(512, 1) wfold_kernel(const WFoldArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const uint32_t sbase = smem_u32(smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nbin = a.nbin;
    const uint32_t blk = (uint32_t)nbin * 256u;  // bytes of one (warp, r) histogram block
    const uint32_t edges_off = WSMEM_HDR;
    const uint32_t edges_bytes = ((uint32_t)(nbin + 1) * 8u + 15u) & ~15u;
    const uint32_t tile_off = edges_off + edges_bytes;
    constexpr uint32_t TILE_BYTES = WTILE_EV * 8u;
    const uint32_t hist_off = tile_off + (HAS_W ? 4u : 2u) * TILE_BYTES;
    const uint32_t hist = sbase + hist_off;
    double *s_edges = reinterpret_cast<double *>(smem + edges_off);

    for (uint32_t o = tid * 8; o < (uint32_t)a.warps * R * blk; o += blockDim.x * 8) sts_f64(hist + o, 0.0);
    for (int i = tid; i <= nbin; i += blockDim.x) s_edges[i] = a.edges[i];

    const long long tbl = (long long)blockIdx.x * a.warps + warp;
    const long long tb = a.tb0 + tbl;
    double f[R], fdh[R];
    uint32_t cb[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
        const long long g = (tb * R + r) * 32 + lane;
        const bool valid = g < a.n_trials;
        const long long gi = valid ? g : 0;
        const long long ifd = gi / a.n_f, jf = gi - ifd * a.n_f;
        f[r] = a.freqs[jf];
        fdh[r] = HAS_FD ? __dmul_rn(0.5, a.fdots[ifd]) : 0.0;
        cb[r] = hist + (uint32_t)(warp * R + r) * blk + (uint32_t)lane * 8u;
    }

    const long long e0 = (long long)blockIdx.y * a.ev_per_split;
    const long long e1 = min(a.n_ev, e0 + a.ev_per_split);
    const int ntiles = (e1 > e0) ? (int)((e1 - e0 + WTILE_EV - 1) / WTILE_EV) : 0;
    const uint32_t bar[2] = {sbase, sbase + 8};

    // stage `it & 1`: times at tile_off + buf*TILE_BYTES, weights two stages further
    auto issue = [&](int it) {
        const long long s = e0 + (long long)it * WTILE_EV;
        const int cnt = (int)min((long long)WTILE_EV, e1 - s);
        const uint32_t bytes = (uint32_t)cnt * 8u, bulk = bytes & ~15u;
        const uint32_t off_t = tile_off + (uint32_t)(it & 1) * TILE_BYTES;
        const uint32_t off_w = off_t + 2u * TILE_BYTES;
        const char *src_t = reinterpret_cast<const char *>(a.ev) + s * 8;
        const char *src_w = reinterpret_cast<const char *>(a.wt) + s * 8;
        if (bytes != bulk) {  // odd trailing fp64: plain copy, ordered before the arrive below
            *reinterpret_cast<double *>(smem + off_t + bulk) = *reinterpret_cast<const double *>(src_t + bulk);
            if (HAS_W) *reinterpret_cast<double *>(smem + off_w + bulk) = *reinterpret_cast<const double *>(src_w + bulk);
        }
        if (bulk) {
            mbar_expect_tx(bar[it & 1], HAS_W ? 2u * bulk : bulk);
            bulk_g2s(sbase + off_t, src_t, bulk, bar[it & 1]);
            if (HAS_W) bulk_g2s(sbase + off_w, src_w, bulk, bar[it & 1]);
        } else {
            mbar_arrive(bar[it & 1]);
        }
    };

    if (tid == 0) {
        mbar_init(bar[0], 1);
        mbar_init(bar[1], 1);
        fence_mbar_init();
        fence_proxy_async();
        if (ntiles > 0) issue(0);
        if (ntiles > 1) issue(1);
    }
    __syncthreads();

    const double w1 = a.unit_weight;
    for (int it = 0; it < ntiles; it++) {
        const int buf = it & 1;
        mbar_wait(bar[buf], (uint32_t)(it >> 1) & 1u);
        const int cnt = (int)min((long long)WTILE_EV, e1 - (e0 + (long long)it * WTILE_EV));
        const double *tp = reinterpret_cast<const double *>(smem + tile_off + (uint32_t)buf * TILE_BYTES);
        const double *wp = tp + 2 * WTILE_EV;
#pragma unroll 2
        for (int e = 0; e < cnt; e++) {
            const double t = tp[e];
            const double w = HAS_W ? wp[e] : w1;
            const double tt = HAS_FD ? __dmul_rn(t, t) : 0.0;
#pragma unroll
            for (int r = 0; r < R; r++) {
                const int b = wbin<HAS_FD, POW2>(t, tt, f[r], fdh[r], nbin, s_edges);
                const uint32_t ad = cb[r] + ((uint32_t)b << 8);
                sts_f64(ad, __dadd_rn(lds_f64(ad), w));
            }
        }
        __syncthreads();  // every warp is done with stage `buf`
        if (tid == 0 && it + 2 < ntiles) {
            fence_proxy_async();
            issue(it + 2);
        }
    }

    // partial profiles -> global scratch [trial/32][bin][32], coalesced across the warp
#pragma unroll
    for (int r = 0; r < R; r++) {
        double *dst = a.scratch + ((size_t)(tbl * R + r) * nbin) * 32 + lane;
        for (int j = 0; j < nbin; j++) {
            const double c = lds_f64(cb[r] + ((uint32_t)j << 8));
            if (c != 0.0) atomicAdd(dst + (size_t)j * 32, c);
        }
    }
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

// stingray.pulse.pulsar.phase_exposure(start_phase, stop_phase, 1, nbin, gti=gti_phases)[j] before the
// final normalisation, for one trial and one bin: the per-bin sequence of fp64 accumulations is upstream's.
__device__ double expo_bin(int j, int nbin, const double *edges, const double *gti, int n_gti, double f, double fdh,
                           bool has_fd, double start_ph, double stop_ph) {
    double e = 0.0;
    const dd p0 = {edges[j], 0.0}, p1 = {edges[j + 1], 0.0};
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
        e = __dadd_rn(e, __ddiv_rn(nper, (double)nbin));
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
        for (int k = 0; k < nl; k++) {
            if (!(dd_le(p0, l1[k]) && dd_le(l0[k], p1))) continue;
            const dd w = dd_add(dd_min(p1, l1[k]), dd_neg(dd_max(p0, l0[k])));
            const dd s = two_sum(e, w.hi);
            e = __dadd_rn(s.hi, __dadd_rn(s.lo, w.lo));
        }
    }
    return e;
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
    for (int j = 0; j < nbin; j++) {
        const double e = expo_bin(j, nbin, edges, gti, n_gti, f, fdh, has_fd != 0, start_ph, stop_ph);
        expo[base + (size_t)j * 32] = e;
        emax = fmax(emax, e);
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
template <int R, bool HAS_FD, bool POW2>
static wfold_fn_t pick_wfold_W(bool has_w) {
    return has_w ? wfold_kernel<R, HAS_FD, POW2, true> : wfold_kernel<R, HAS_FD, POW2, false>;
}
template <int R>
static wfold_fn_t pick_wfold_R(bool has_fd, bool pow2, bool has_w) {
    if (has_fd) return pow2 ? pick_wfold_W<R, true, true>(has_w) : pick_wfold_W<R, true, false>(has_w);
    return pow2 ? pick_wfold_W<R, false, true>(has_w) : pick_wfold_W<R, false, false>(has_w);
}
static wfold_fn_t pick_wfold(int R, bool has_fd, bool pow2, bool has_w) {
    switch (R) {
        case 4: return pick_wfold_R<4>(has_fd, pow2, has_w);
        case 2: return pick_wfold_R<2>(has_fd, pow2, has_w);
        default: return pick_wfold_R<1>(has_fd, pow2, has_w);
    }
}

static cudaEvent_t g_wev0[64], g_wev1[64];
static bool g_wev_ok[64] = {false};
static int g_wlast_device = -1;

int wfold_search_dev_impl(const double *d_times, const double *d_weights, double unit_weight, int64_t n_ev,
                          const double *d_freqs, int64_t n_f, const double *d_fdots, int64_t n_fd, int nbin,
                          int nharm, const double *h_gti, int64_t n_gti, double start_time, double stop_time,
                          double *d_stats, double *d_profiles, cudaStream_t st) {
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

    // fdot == 0 everywhere: the fdot term adds +0.0 upstream, skip it
    std::vector<double> h_fd((size_t)n_fd);
    HX_CUDA_TRY(cudaMemcpyAsync(h_fd.data(), d_fdots, (size_t)n_fd * 8, cudaMemcpyDeviceToHost, st));
    HX_CUDA_TRY(cudaStreamSynchronize(st));
    bool has_fd = false;
    for (double v : h_fd) has_fd |= (v != 0.0);

    const bool pow2 = (nbin & (nbin - 1)) == 0;
    const bool has_w = d_weights != nullptr;
    // plan: trials per thread and warps per CTA from the fp64 histogram footprint (256 B per bin and warp)
    int R = nbin <= 16 ? 4 : nbin <= 64 ? 2 : 1;
    const uint32_t edges_bytes = ((uint32_t)(nbin + 1) * 8u + 15u) & ~15u;
    const uint32_t fixed = WSMEM_HDR + edges_bytes + (has_w ? 4u : 2u) * WTILE_EV * 8u;
    const uint32_t blk = (uint32_t)nbin * 256u;
    int W = 0;
    for (; R >= 1; R >>= 1) {
        W = std::min<int>(8, (int)(((uint32_t)di.smem_optin - fixed) / ((uint32_t)R * blk)));
        if (W >= 1) break;
    }
    if (W < 1) {
        set_error("hx_wfold_search: nbin=%d does not fit in shared memory (fp64 profiles: nbin <= %d)", nbin,
                  (int)(((uint32_t)di.smem_optin - fixed) / 256u));
        return HX_ERR_ARG;
    }
    const long long n_tb = (n_trials + 32LL * R - 1) / (32LL * R);
    {  // few trials: smaller CTAs so that every SM gets one
        const long long want = (n_tb + di.sm_count - 1) / di.sm_count;
        if (want < W) W = (int)std::max<long long>(1, want);
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
    if (n_ev > 0 && n_cta_x < 4 * slots) {
        ev_splits = std::min<long long>((4 * slots + n_cta_x - 1) / n_cta_x, std::max<long long>(1, n_ev / (4 * WTILE_EV)));
        ev_splits = std::min<long long>(std::max<long long>(ev_splits, 1), 65535);
    }
    long long ev_per_split = (n_ev + ev_splits - 1) / ev_splits;
    ev_per_split = std::max<long long>(WTILE_EV, (ev_per_split + WTILE_EV - 1) / WTILE_EV * WTILE_EV);
    ev_splits = std::max<long long>(1, (n_ev + ev_per_split - 1) / ev_per_split);

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
    a.ev = d_times, a.wt = d_weights, a.n_ev = n_ev, a.freqs = d_freqs, a.n_f = n_f, a.fdots = d_fdots;
    a.n_trials = n_trials, a.nbin = nbin, a.warps = W, a.ev_per_split = ev_per_split;
    a.scratch = b_scratch.as<double>(), a.edges = b_edges.as<double>(), a.unit_weight = unit_weight;
    wfold_fn_t fn = pick_wfold(R, has_fd, pow2, has_w);
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
    g_wlast_device = device;
    HX_CUDA_TRY(cudaStreamSynchronize(st));
    return HX_OK;
}

}  // namespace hx

extern "C" {

int hx_wfold_last_kernel_ms(double *ms) {
    if (!ms || hx::g_wlast_device < 0) {
        hx::set_error("hx_wfold_last_kernel_ms: no weighted fold kernel has been launched");
        return HX_ERR_ARG;
    }
    const int d = hx::g_wlast_device;
    float f = 0.f;
    HX_CUDA_TRY(cudaEventSynchronize(hx::g_wev1[d]));
    HX_CUDA_TRY(cudaEventElapsedTime(&f, hx::g_wev0[d], hx::g_wev1[d]));
    *ms = f;
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
    HX_CUDA_TRY(cudaSetDevice(device));
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
