// K1win: the exact fold search repeated on MANY small time windows of one sorted event list, in one launch.
//
// Two callers of the reference run the same per-trial fold + statistic over windows of the data:
//   * `dyn_folding_search` (hendrics/efsearch.py:1196-1243; the CLI runs it by default for observations longer
//     than --dynstep = 128 s, :1912-1922): every window [start, stop) of the RAW event times is an independent
//     upstream search -- re-zeroed to its own first event (A1), its single segment half-open so that the events at
//     the window's last time are dropped (A2), skipped when what is left spans < 0.2 of the window's length;
//   * the segment loop of stingray's `_folding_search` (`--segment-size`, hendrics/efsearch.py:1658-1663,
//     1901-1903): windows [s, s + segment_size) of the times re-zeroed ONCE to the first
//     event; statistics of the valid segments are summed (in segment order) and divided by their number.
// A 100 ks observation has 782 windows of ~25 trials x ~130 events each: one K1 launch per window is pure launch
// and synchronisation overhead (0.4 s, profiles/r01_cli_timing.txt).  Here grid.y = window: a planning kernel finds
// every window's event range by binary search with the reference's own comparisons (on fl(t - origin), which is
// monotone in t), then one launch folds all windows.  lane = trial with a private [bin][lane] histogram column like
// K1; the arithmetic replays the reference operation by operation (no pre-scaled fast path: nothing to prove about
// the inputs, any nbin, negative phases included), events are read straight from global memory (a warp-uniform
// address: one L1 broadcast per event).  This kernel is for SMALL windows; the host keeps one K1 launch per window
// when a window is large enough to fill the GPU by itself (see hendrics_b200/search.py).
#include <algorithm>
#include <cmath>
#include <vector>

#include "hx_common.cuh"

namespace hx {

struct WinPlan {
    long long lo, hi;  // events [lo, hi) of the sorted list are folded
    double origin;     // subtracted from every event time (one rounded fp64 subtraction, like numpy's)
    int status;        // 1: folded; 0: skipped by upstream (no events / span < 0.2 segment): NaN (0/0) when it is the
                       // only segment; 2: upstream raises for this window (empty, or zero length): dyn search stores zeros
};

enum { WIN_SEGMENTS = 0, WIN_SEARCHES = 1 };

__global__ void windows_plan_kernel(const double *__restrict__ t, long long n_ev, const double *__restrict__ wstart,
                                    const double *__restrict__ wstop, int n_win, int mode, double origin,
                                    double segment_size, WinPlan *__restrict__ plan) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_win) return;
    // first i in [a, b) with fl(t[i] - o) >= x  (t sorted, the rounded subtraction is monotone)
    auto lower = [&](long long a, long long b, double x, double o) {
        while (a < b) {
            const long long m = (a + b) >> 1;
            if (__dsub_rn(t[m], o) >= x) b = m;
            else a = m + 1;
        }
        return a;
    };
    WinPlan p;
    p.status = 1;
    if (mode == WIN_SEGMENTS) {
        // good = (times >= s) & (times < s + segment_size) on the re-zeroed times
        p.origin = origin;
        p.lo = lower(0, n_ev, wstart[w], origin);
        p.hi = lower(p.lo, n_ev, wstop[w], origin);
        if (p.hi <= p.lo) p.status = 0;  // len(ts) < 1
        else if (__dsub_rn(__dsub_rn(t[p.hi - 1], origin), __dsub_rn(t[p.lo], origin)) < 0.2 * segment_size) p.status = 0;
    } else {
        // times_filt = time[(time >= st) & (time < sp)], then a whole upstream search of it
        const long long lo = lower(0, n_ev, wstart[w], 0.0), hi = lower(lo, n_ev, wstop[w], 0.0);
        p.lo = lo, p.hi = hi, p.origin = 0.0;
        if (hi <= lo) p.status = 2;  // times[0] of an empty array: IndexError
        else {
            p.origin = t[lo];                                     // A1
            const double length = __dsub_rn(t[hi - 1], p.origin);  // times[-1] after re-zeroing = the one segment's size
            if (!(length > 0.0)) p.status = 2;                    // np.arange(0, 0, 0): ZeroDivisionError
            else {
                p.hi = lower(lo, hi, length, p.origin);  // A2: times < 0 + length
                // ts[-1] - ts[0] < 0.2 * segment_size (ts[0] is 0 after re-zeroing)
                if (__dsub_rn(__dsub_rn(t[p.hi - 1], p.origin), __dsub_rn(t[lo], p.origin)) < 0.2 * length) p.status = 0;
            }
        }
    }
    plan[w] = p;
}

constexpr int WIN_MAX_WARPS = 4;

template <bool HAS_FD>
__global__ void __launch_bounds__(32 * WIN_MAX_WARPS) fold_windows_kernel(
    const double *__restrict__ t, const WinPlan *__restrict__ plan, const double *__restrict__ freqs, long long n_f,
    const double *__restrict__ fdots, long long n_trials, int nbin, int nharm, const double *__restrict__ tab,
    double *__restrict__ stats) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    double *tabs = reinterpret_cast<double *>(smem);  // [2 * nbin]
    const uint32_t hist = smem_u32(smem) + (uint32_t)nbin * 16u + (uint32_t)warp * (uint32_t)nbin * 128u + (uint32_t)lane * 4u;
    const WinPlan p = plan[blockIdx.y];
    const long long g = ((long long)blockIdx.x * nwarps + warp) * 32 + lane;
    const bool valid = g < n_trials;
    double *out = stats + (size_t)blockIdx.y * n_trials;
    if (p.status != 1) {
        if (valid) out[g] = p.status == 0 ? __longlong_as_double(0x7ff8000000000000ll) : 0.0;
        return;
    }
    const long long gi = valid ? g : 0;
    const long long ifd = gi / n_f, jf = gi - ifd * n_f;
    const double f = freqs[jf], fd = HAS_FD ? fdots[ifd] : 0.0;
    for (int j = 0; j < nbin; j++) sts_u32(hist + ((uint32_t)j << 7), 0u);
    for (int i = threadIdx.x; i < 2 * nbin; i += blockDim.x) tabs[i] = tab[i];
    __syncthreads();
    const double dn = (double)nbin;
    for (long long i = p.lo; i < p.hi; i++) {
        const double tt = __dsub_rn(__ldg(t + i), p.origin);
        // _pulse_phase_fast: t*f + 0.5*t*t*fd, left to right, every product and the sum separately rounded
        double ph = __dmul_rn(tt, f);
        if (HAS_FD) ph = __dadd_rn(ph, __dmul_rn(__dmul_rn(__dmul_rn(0.5, tt), tt), fd));
        ph = __dsub_rn(ph, floor(ph));
        // _profile_fast: int64(floor(phase * nbin)), overflow into bin nbin - 1 (A4)
        int b = (int)floor(__dmul_rn(ph, dn));
        b = min(max(b, 0), nbin - 1);
        red_shared_inc(hist + ((uint32_t)b << 7));
    }
    __syncwarp();
    if (valid) out[g] = profile_stat([&](int j) { return lds_u32(hist + ((uint32_t)j << 7)); }, nbin, nharm, tabs, tabs + nbin);
}

// in-order sum over the folded windows, per trial (upstream: `stats += stat` segment after segment)
__global__ void windows_sum_kernel(const double *__restrict__ stats, const WinPlan *__restrict__ plan, int n_win,
                                   long long n_trials, double *__restrict__ sum) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_trials) return;
    double acc = 0.0;
    for (int w = 0; w < n_win; w++)
        if (plan[w].status == 1) acc = __dadd_rn(acc, stats[(size_t)w * n_trials + g]);
    sum[g] = acc;
}

}  // namespace hx

extern "C" int hx_fold_windows_dev(const double *d_times, int64_t n_ev, const double *win_start, const double *win_stop,
                                   int64_t n_win, int mode, double origin, double segment_size, const double *d_freqs,
                                   int64_t n_f, const double *d_fdots, int64_t n_fd, int nbin, int nharm, double *d_stats,
                                   double *d_sum, int *h_status, void *stream) {
    using namespace hx;
    std::lock_guard<std::recursive_mutex> api_lock(api_mutex());
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (!d_times || !win_start || !win_stop || !d_freqs || !d_fdots || !d_stats || !h_status) {
        set_error("hx_fold_windows_dev: null pointer argument");
        return HX_ERR_ARG;
    }
    if (n_ev < 0 || n_win < 0 || n_win > 65535 || n_f < 0 || n_fd < 0 || nbin < 2 || nbin > HX_MAX_NBIN || nharm < 0 ||
        nharm > HX_MAX_NHARM || (mode != WIN_SEGMENTS && mode != WIN_SEARCHES)) {
        set_error("hx_fold_windows_dev: bad sizes (n_ev=%lld n_win=%lld n_f=%lld n_fd=%lld nbin=%d nharm=%d mode=%d)",
                  (long long)n_ev, (long long)n_win, (long long)n_f, (long long)n_fd, nbin, nharm, mode);
        return HX_ERR_ARG;
    }
    const long long n_trials = (long long)n_f * n_fd;
    if (n_win == 0 || n_trials == 0) return HX_OK;
    int device;
    HX_CUDA_TRY(cudaGetDevice(&device));
    DeviceInfo di;
    int rc = get_device_info(device, &di);
    if (rc) return rc;
    double r6[6];
    if ((rc = device_ranges(d_times, n_ev, d_freqs, n_f, d_fdots, n_fd, r6, st)) != HX_OK) return rc;
    const bool has_fd = !(r6[4] == 0.0 && r6[5] == 0.0);
    {
        const double tspan = std::max(std::fabs(r6[0] - origin), std::fabs(r6[1] - origin));
        const double tmax = mode == WIN_SEGMENTS ? tspan : std::fabs(r6[1] - r6[0]);
        const double fa = std::max(std::fabs(r6[2]), std::fabs(r6[3])), fda = std::max(std::fabs(r6[4]), std::fabs(r6[5]));
        const double phmax = tmax * fa + 0.5 * tmax * tmax * fda;
        if (!(phmax < 1125899906842624.0)) {
            set_error("hx_fold_windows_dev: |phase| up to %.3g cycles exceeds the fp64 folding range (2^50)", phmax);
            return HX_ERR_RANGE;
        }
    }

    AsyncBuf b_w(st), b_plan(st), b_tab(st);
    HX_CUDA_TRY(b_w.alloc((size_t)n_win * 2 * sizeof(double)));
    HX_CUDA_TRY(b_plan.alloc((size_t)n_win * sizeof(WinPlan)));
    HX_CUDA_TRY(cudaMemcpyAsync(b_w.p, win_start, (size_t)n_win * 8, cudaMemcpyHostToDevice, st));
    HX_CUDA_TRY(cudaMemcpyAsync(b_w.as<double>() + n_win, win_stop, (size_t)n_win * 8, cudaMemcpyHostToDevice, st));
    std::vector<double> tab;
    host_tables(nbin, tab);
    HX_CUDA_TRY(b_tab.alloc(tab.size() * sizeof(double)));
    HX_CUDA_TRY(cudaMemcpyAsync(b_tab.p, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    windows_plan_kernel<<<(unsigned)((n_win + 127) / 128), 128, 0, st>>>(d_times, n_ev, b_w.as<double>(),
                                                                       b_w.as<double>() + n_win, (int)n_win, mode, origin,
                                                                       segment_size, b_plan.as<WinPlan>());
    HX_CHECK_LAUNCH("windows_plan_kernel");

    // warps per CTA: as many 32-trial blocks as fit next to the tables, at most 4, no more than the trials need
    int W = (int)std::min<long long>(WIN_MAX_WARPS, (n_trials + 31) / 32);
    if (W == 3) W = 4;  // 1, 2 or 4 warps: 3-warp CTAs load the SM's four schedulers unevenly (the 4th warp's lanes are masked)
    while (W > 1 && (size_t)nbin * 16 + (size_t)W * nbin * 128 > (size_t)di.smem_optin) W--;
    const size_t smem = (size_t)nbin * 16 + (size_t)W * nbin * 128;
    if (smem > (size_t)di.smem_optin) {
        set_error("hx_fold_windows_dev: nbin=%d does not fit in shared memory", nbin);
        return HX_ERR_ARG;
    }
    auto fn = has_fd ? fold_windows_kernel<true> : fold_windows_kernel<false>;
    HX_CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long gx = (n_trials + 32LL * W - 1) / (32LL * W);
    if (gx > 2147483647LL) {
        set_error("hx_fold_windows_dev: trial grid too large");
        return HX_ERR_ARG;
    }
    dim3 grid((unsigned)gx, (unsigned)n_win);
    fn<<<grid, 32 * W, smem, st>>>(d_times, b_plan.as<WinPlan>(), d_freqs, n_f, d_fdots, n_trials, nbin, nharm,
                                   b_tab.as<double>(), d_stats);
    HX_CHECK_LAUNCH("fold_windows_kernel");
    if (d_sum) {
        windows_sum_kernel<<<(unsigned)((n_trials + 255) / 256), 256, 0, st>>>(d_stats, b_plan.as<WinPlan>(), (int)n_win,
                                                                             n_trials, d_sum);
        HX_CHECK_LAUNCH("windows_sum_kernel");
    }
    std::vector<WinPlan> hp((size_t)n_win);
    HX_CUDA_TRY(cudaMemcpyAsync(hp.data(), b_plan.p, (size_t)n_win * sizeof(WinPlan), cudaMemcpyDeviceToHost, st));
    HX_CUDA_TRY(cudaStreamSynchronize(st));
    for (int64_t w = 0; w < n_win; w++) h_status[w] = hp[(size_t)w].status;
    return HX_OK;
}
