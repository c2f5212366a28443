// Orbital-demodulation fold search for sm_100a (widening).
//
// Replaces the inner loop of hendrics/efsearch.py:263-310 (folding_orbital_search): for ONE spin / orbit parameter set
// (freq, fdot, X = a sin i / c, Porb) and a list of trial epochs of ascending node T0, the reference does per T0
//     new_values = times - X * np.sin(2 * np.pi * (times - T0) / Porb)                      (efsearch.py:295)
//     new_values = new_values - X * np.sin(2 * np.pi * (new_values - T0) / Porb)            (efsearch.py:296)
//     fgrid, stats = fun(new_values, np.array([freq]), fdots=fdot, **fun_kwargs)            (efsearch.py:297)
// and keeps max / min / arg-max of stats[0] over the T0s.  `fun` is epoch_folding_search or z_n_search, i.e. the
// upstream fast branch on the demodulated times: re-zero to the first event (A1), keep 0 <= u < u[-1] (A2), fold.
//
// Design: lane = T0 trial (one per thread), private [bin][lane] u32 histogram column (conflict-free red.shared), events
// staged through the same 2-stage cp.async.bulk + mbarrier pipeline as K1w and broadcast to all lanes.  The work per
// event*trial is two fp64 sin() evaluations (the arithmetic sequence of the reference, operation by operation: the
// argument is fl(fl(2 pi * fl(t - T0)) / Porb), IEEE division) plus one generic fold, so the kernel is FP64-bound and
// needs no more staging finesse than that.  CUDA's sin() and numpy's agree to an ulp or two, not bit for bit: a
// demodulated time can differ in its last bit and -- with probability ~ f * ulp(t) * nbin per event -- land in the
// neighbouring bin.  Parity for this row is therefore "profiles equal up to isolated +-1 moves between adjacent bins,
// statistics within 1e-6" (tests/test_orbital_gpu.py), not bit-exact; the reference itself draws its T0s from an
// unseeded RNG.
#include <algorithm>
#include <cstdlib>

#include "hx_common.cuh"

namespace hx {

constexpr int OTILE_EV = 512;
constexpr uint32_t OSMEM_HDR = 16;

struct OrbArgs {
    const double *ev;
    long long n_ev;
    const double *T0;
    long long n_t0;
    double freq, fdot, X, Porb;
    int nbin, warps;
    long long ev_per_split;
    uint32_t *scratch;  // [n_t0/32][nbin][32]
};

// efsearch.py:295-296, operation by operation
__device__ __forceinline__ double demodulate(double t, double T0, double X, double Porb) {
    const double TWOPI = 6.283185307179586;  // 2 * np.pi
    double a = __ddiv_rn(__dmul_rn(TWOPI, __dsub_rn(t, T0)), Porb);
    const double t1 = __dsub_rn(t, __dmul_rn(X, sin(a)));
    a = __ddiv_rn(__dmul_rn(TWOPI, __dsub_rn(t1, T0)), Porb);
    return __dsub_rn(t1, __dmul_rn(X, sin(a)));
}

template <bool HAS_FD>
__global__ void __launch_bounds__(256) orbital_fold_kernel(const OrbArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const uint32_t sbase = smem_u32(smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nbin = a.nbin;
    const uint32_t blk = (uint32_t)nbin << 7;
    constexpr uint32_t TILE_BYTES = OTILE_EV * 8u;
    const uint32_t tile_off = OSMEM_HDR, hist_off = tile_off + 2u * TILE_BYTES;
    const uint32_t cb = sbase + hist_off + (uint32_t)warp * blk + (uint32_t)lane * 4u;
    for (uint32_t o = tid * 4; o < (uint32_t)a.warps * blk; o += blockDim.x * 4) sts_u32(sbase + hist_off + o, 0u);

    const long long slot = ((long long)blockIdx.x * a.warps + warp) * 32 + lane;
    const double T0 = a.T0[min(slot, a.n_t0 - 1)];
    // A1 / A2 of the search called on the demodulated times: origin = first array element, length = last - first
    const double u0 = demodulate(a.ev[0], T0, a.X, a.Porb);
    const double length = __dsub_rn(demodulate(a.ev[a.n_ev - 1], T0, a.X, a.Porb), u0);
    const double fdh = __dmul_rn(0.5, a.fdot);

    const long long e0 = (long long)blockIdx.y * a.ev_per_split;
    const long long e1 = min(a.n_ev, e0 + a.ev_per_split);
    const int ntiles = (e1 > e0) ? (int)((e1 - e0 + OTILE_EV - 1) / OTILE_EV) : 0;
    const uint32_t bar[2] = {sbase, sbase + 8};
    auto issue = [&](int it) {
        const long long s = e0 + (long long)it * OTILE_EV;
        const uint32_t cnt = (uint32_t)min((long long)OTILE_EV, e1 - s);
        const uint32_t bytes = cnt * 8u, bulk = bytes & ~15u;
        const uint32_t dst = tile_off + (uint32_t)(it & 1) * TILE_BYTES;
        const char *src = reinterpret_cast<const char *>(a.ev) + s * 8;
        if (bytes != bulk) *reinterpret_cast<double *>(smem + dst + bulk) = *reinterpret_cast<const double *>(src + bulk);
        if (bulk) {
            mbar_expect_tx(bar[it & 1], bulk);
            bulk_g2s(sbase + dst, src, bulk, bar[it & 1]);
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

    for (int it = 0; it < ntiles; it++) {
        const int buf = it & 1;
        mbar_wait(bar[buf], (uint32_t)(it >> 1) & 1u);
        const int cnt = (int)min((long long)OTILE_EV, e1 - (e0 + (long long)it * OTILE_EV));
        const double *tp = reinterpret_cast<const double *>(smem + tile_off + (uint32_t)buf * TILE_BYTES);
#pragma unroll 4
        for (int e = 0; e < cnt; e++) {
            const double u = __dsub_rn(demodulate(tp[e], T0, a.X, a.Porb), u0);
            // the fast branch of the search on `u`: phase = u*f + ((0.5*u)*u)*fd, minus its floor, times nbin, floored
            double x = __dmul_rn(u, a.freq);
            if (HAS_FD) x = __dadd_rn(x, __dmul_rn(__dmul_rn(u, u), fdh));
            const double fl = __dsub_rn(__dadd_rd(x, HX_MAGIC), HX_MAGIC);
            const double fr = __dsub_rn(x, fl);
            int b = __double2loint(__dadd_rd(__dmul_rn(fr, (double)nbin), HX_MAGIC));
            b = min(b, nbin - 1);
            if (u >= 0.0 && u < length) red_shared_inc(cb + ((uint32_t)b << 7));
        }
        __syncthreads();
        if (tid == 0 && it + 2 < ntiles) {
            fence_proxy_async();
            issue(it + 2);
        }
    }
    uint32_t *dst = a.scratch + ((size_t)((long long)blockIdx.x * a.warps + warp) * nbin) * 32 + lane;
    for (int j = 0; j < nbin; j++) {
        const uint32_t c = lds_u32(cb + ((uint32_t)j << 7));
        if (c) atomicAdd(dst + (size_t)j * 32, c);
    }
}

__global__ void __launch_bounds__(256) orbital_stat_kernel(const uint32_t *__restrict__ scratch, long long n, int nbin,
                                                           int nharm, const double *__restrict__ tab,
                                                           double *__restrict__ stats, uint32_t *__restrict__ profiles) {
    extern __shared__ double tabs[];
    for (int i = threadIdx.x; i < 2 * nbin; i += blockDim.x) tabs[i] = tab[i];
    __syncthreads();
    const long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n) return;
    const uint32_t *src = scratch + (size_t)(l >> 5) * nbin * 32 + (l & 31);
    stats[l] = profile_stat([&](int j) { return src[(size_t)j * 32]; }, nbin, nharm, tabs, tabs + nbin);
    if (profiles)
        for (int j = 0; j < nbin; j++) profiles[(size_t)l * nbin + j] = src[(size_t)j * 32];
}

int orbital_search_dev_impl(const double *d_times, int64_t n_ev, double freq, double fdot, double X, double Porb,
                            const double *d_T0, int64_t n_t0, int nbin, int nharm, double *d_stats,
                            uint32_t *d_profiles, cudaStream_t st) {
    std::lock_guard<std::recursive_mutex> api_lock(api_mutex());
    if (!d_times || !d_T0 || !d_stats) {
        set_error("hx_orbital_search: null pointer argument");
        return HX_ERR_ARG;
    }
    if (n_ev < 2 || n_t0 < 0 || nbin < 2 || nbin > HX_MAX_NBIN || nharm < 0 || nharm > HX_MAX_NHARM ||
        !std::isfinite(freq) || !std::isfinite(fdot) || !std::isfinite(X) || !(Porb > 0) || !std::isfinite(Porb)) {
        set_error("hx_orbital_search: bad arguments (n_ev=%lld n_t0=%lld nbin=%d nharm=%d Porb=%g)", (long long)n_ev,
                  (long long)n_t0, nbin, nharm, Porb);
        return HX_ERR_ARG;
    }
    if ((reinterpret_cast<uintptr_t>(d_times) & 15) != 0) {
        set_error("hx_orbital_search_dev: d_times must be 16-byte aligned");
        return HX_ERR_ARG;
    }
    if (n_t0 == 0) return HX_OK;
    int device;
    HX_CUDA_TRY(cudaGetDevice(&device));
    DeviceInfo di;
    int rc = get_device_info(device, &di);
    if (rc) return rc;
    const uint32_t fixed = OSMEM_HDR + 2u * OTILE_EV * 8u, blk = (uint32_t)nbin * 128u;
    int W = std::min<int>(8, (int)(((uint32_t)di.smem_optin - fixed) / blk));
    if (W < 1) {
        set_error("hx_orbital_search: nbin=%d does not fit in shared memory", nbin);
        return HX_ERR_ARG;
    }
    const long long n_tb = (n_t0 + 31) / 32;
    // Warps per CTA: the kernel is a long dependent chain per event (two sin(), two divisions), so it lives on resident
    // warps.  The stage buffers are per CTA: larger CTAs leave room for more warps per SM (W = 1: 12 warps per SM at
    // nbin = 64, W = 8: 24), and the event splitting below provides the CTAs when there are few T0 blocks (measured:
    // 1e6 events x 4736 T0, 42.5 -> see profiles/r02_orbital_throughput.txt).  1, 2 or a multiple of 4 warps.
    {
        int best_w = 1, best_res = 0;
        for (int w : {1, 2, 4, 8}) {
            if (w > W || w > n_tb) continue;
            const int k = std::min(16, di.smem_per_sm / (int)(fixed + (uint32_t)w * blk + di.reserved_smem));
            if (k * w > best_res) best_res = k * w, best_w = w;
        }
        W = best_w;
    }
    const long long grid_x = (n_tb + W - 1) / W;
    const uint32_t smem_bytes = fixed + (uint32_t)W * blk;
    const int ctas_per_sm = std::max(1, std::min(16, di.smem_per_sm / (int)(smem_bytes + di.reserved_smem)));
    long long splits = std::max<long long>(1, (4LL * di.sm_count * ctas_per_sm + grid_x - 1) / grid_x);
    splits = std::min<long long>(std::min<long long>(splits, std::max<long long>(1, n_ev / (2 * OTILE_EV))), 65535);
    long long per = (n_ev + splits - 1) / splits;
    per = std::max<long long>(OTILE_EV, (per + OTILE_EV - 1) / OTILE_EV * OTILE_EV);
    splits = (n_ev + per - 1) / per;

    AsyncBuf b_tab(st), b_scratch(st);
    std::vector<double> tab;
    host_tables(nbin, tab);
    HX_CUDA_TRY(b_tab.alloc(tab.size() * sizeof(double)));
    HX_CUDA_TRY(cudaMemcpyAsync(b_tab.p, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    const size_t sbytes = (size_t)(grid_x * W) * nbin * 32 * sizeof(uint32_t);
    HX_CUDA_TRY(b_scratch.alloc(sbytes));
    HX_CUDA_TRY(cudaMemsetAsync(b_scratch.p, 0, sbytes, st));
    OrbArgs a{};
    a.ev = d_times, a.n_ev = n_ev, a.T0 = d_T0, a.n_t0 = n_t0, a.freq = freq, a.fdot = fdot, a.X = X, a.Porb = Porb;
    a.nbin = nbin, a.warps = W, a.ev_per_split = per, a.scratch = b_scratch.as<uint32_t>();
    auto fn = fdot != 0.0 ? orbital_fold_kernel<true> : orbital_fold_kernel<false>;
    HX_CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    fn<<<dim3((unsigned)grid_x, (unsigned)splits), W * 32, smem_bytes, st>>>(a);
    HX_CHECK_LAUNCH("orbital_fold_kernel");
    orbital_stat_kernel<<<(unsigned)((n_t0 + 255) / 256), 256, 2 * nbin * sizeof(double), st>>>(
        b_scratch.as<uint32_t>(), n_t0, nbin, nharm, b_tab.as<double>(), d_stats, d_profiles);
    HX_CHECK_LAUNCH("orbital_stat_kernel");
    HX_CUDA_TRY(cudaStreamSynchronize(st));
    return HX_OK;
}

}  // namespace hx

extern "C" {

int hx_orbital_search_dev(const double *d_times, int64_t n_ev, double freq, double fdot, double X, double Porb,
                          const double *d_T0, int64_t n_t0, int nbin, int nharm, double *d_stats, uint32_t *d_profiles,
                          void *stream) {
    return hx::orbital_search_dev_impl(d_times, n_ev, freq, fdot, X, Porb, d_T0, n_t0, nbin, nharm, d_stats, d_profiles,
                                       reinterpret_cast<cudaStream_t>(stream));
}

int hx_orbital_search(const double *times, int64_t n_ev, double freq, double fdot, double X, double Porb,
                      const double *T0, int64_t n_t0, int nbin, int nharm, double *stats, uint32_t *profiles,
                      int device) {
    if (!times || !T0 || !stats || n_ev < 0 || n_t0 < 0 || nbin < 2 || nbin > HX_MAX_NBIN) {
        hx::set_error("hx_orbital_search: bad arguments");
        return HX_ERR_ARG;
    }
    hx::DeviceGuard dev_guard;  // declared before the device buffers: they are freed before the device is restored
    HX_CUDA_TRY(dev_guard.enter(device));
    if (n_t0 == 0) return HX_OK;
    hx::DevBuf dt, d0, ds, dp;
    HX_CUDA_TRY(dt.alloc((size_t)n_ev * 8));
    HX_CUDA_TRY(d0.alloc((size_t)n_t0 * 8));
    HX_CUDA_TRY(ds.alloc((size_t)n_t0 * 8));
    if (profiles) HX_CUDA_TRY(dp.alloc((size_t)n_t0 * nbin * 4));
    HX_CUDA_TRY(cudaMemcpy(dt.p, times, (size_t)n_ev * 8, cudaMemcpyHostToDevice));
    HX_CUDA_TRY(cudaMemcpy(d0.p, T0, (size_t)n_t0 * 8, cudaMemcpyHostToDevice));
    int rc = hx::orbital_search_dev_impl(dt.as<double>(), n_ev, freq, fdot, X, Porb, d0.as<double>(), n_t0, nbin, nharm,
                                         ds.as<double>(), profiles ? dp.as<uint32_t>() : nullptr, 0);
    if (rc) return rc;
    HX_CUDA_TRY(cudaMemcpy(stats, ds.p, (size_t)n_t0 * 8, cudaMemcpyDeviceToHost));
    if (profiles) HX_CUDA_TRY(cudaMemcpy(profiles, dp.p, (size_t)n_t0 * nbin * 4, cudaMemcpyDeviceToHost));
    return HX_OK;
}

}  // extern "C"
