// K2 + K3: the `--fast` ("quasi-fast folding") search step for sm_100a.
//
// Replaces hendrics/efsearch.py:876-925 (search_with_qffa_step):
//   K2  _fast_phase / _fast_phase_fdot / _fast_phase_fddot (efsearch.py:854-873) fused with
//       histogram2d -> _hist2d_numba_seq (hendrics/base.py:667-693): one streaming pass over the
//       events builds the (time slice x phase bin) profiles of up to QF_CMAX centre frequencies.
//       HBM-bound: 8 B/event per pass; phases and per-event bins never leave the SM.
//   K3  _fast_step / shift_and_sum (efsearch.py:817-851, 350-363) + _z_n_fast_cached
//       (hendrics/ffa.py:59-104): for every (L, Q) bin-shift pair, roll each time slice by
//       rint(b*L + b^2*Q) mod nbin, sum, and reduce the summed profile to Z^2_n.
//
// Arithmetic follows the reference operation by operation (separately rounded fp64 products and
// sums, half-to-even rint, truncating bin indices, right edges excluded), so the 2-D profiles are
// bit-exact and the statistics differ only by summation order.
#include <algorithm>
#include <cmath>
#include <vector>

#include "hx_common.cuh"

namespace hx {

constexpr int QF_CMAX = 8;      // centres fused per pass over the events
constexpr int QF_WROWS = 4;     // time slices kept in the shared-memory window of a CTA
constexpr int QF_THREADS = 256;
constexpr int QF_EV_PER_THREAD = 16;

struct QffaCentres {
    double f[QF_CMAX], fd[QF_CMAX], fdd[QF_CMAX];
};

struct QffaHistArgs {
    const double *t;
    long long n_ev;
    double t_first;        // ranges[1][0]
    double delta0, delta1;  // 1/((1-0)/nbin), 1/((t_last-t_first)/nprof)  (host, reference formula)
    int nbin, nprof;
    uint32_t *prof;  // [C][nprof][nbin]
    QffaCentres c;
};

// MODE 0: t*f ; 1: + 0.5*t*t*fd ; 2: + (1/6)*t*t*t*fdd   (operation order of efsearch.py:854-873)
template <int C, int MODE>
__global__ void __launch_bounds__(QF_THREADS) qffa_hist_kernel(const QffaHistArgs a) {
    extern __shared__ uint32_t sh[];  // [C][QF_WROWS][nbin]
    const int nbin = a.nbin, nprof = a.nprof;
    const int wsize = QF_WROWS * nbin;
    for (int i = threadIdx.x; i < C * wsize; i += blockDim.x) sh[i] = 0u;

    const long long chunk = (long long)QF_THREADS * QF_EV_PER_THREAD;
    const long long base = (long long)blockIdx.x * chunk;
    const long long end = min(a.n_ev, base + chunk);
    // first time slice of this CTA's window (events are sorted, so rows only grow)
    int j_base;
    {
        double jf = __dmul_rn(__dsub_rn(a.t[base], a.t_first), a.delta1);
        j_base = (jf >= 0.0 && jf < (double)nprof) ? __double2loint(__dadd_rd(jf, HX_MAGIC)) : 0;
    }
    __syncthreads();

    for (long long idx = base + threadIdx.x; idx < end; idx += QF_THREADS) {
        const double t = a.t[idx];
        const double jf = __dmul_rn(__dsub_rn(t, a.t_first), a.delta1);
        if (!(jf >= 0.0 && jf < (double)nprof)) continue;  // right edge excluded (base.py:674)
        const int j = __double2loint(__dadd_rd(jf, HX_MAGIC));
        const int jr = j - j_base;
        double tssq = 0.0, h = 0.0, s6 = 0.0;
        if (MODE >= 1) {
            tssq = __dmul_rn(t, t);
            h = __dmul_rn(0.5, tssq);  // == (0.5*t)*t bit for bit
        }
        if (MODE >= 2) s6 = __dmul_rn(__dmul_rn(1.0 / 6, tssq), t);
#pragma unroll
        for (int c = 0; c < C; c++) {
            double ph = __dmul_rn(t, a.c.f[c]);
            if (MODE >= 1) ph = __dadd_rn(ph, __dmul_rn(h, a.c.fd[c]));
            if (MODE >= 2) ph = __dadd_rn(ph, __dmul_rn(s6, a.c.fdd[c]));
            const double fl = __dsub_rn(__dadd_rd(ph, HX_MAGIC), HX_MAGIC);  // floor(ph), exact
            const double x = __dsub_rn(ph, fl);                              // phases - floor(phases)
            const double fi = __dmul_rn(x, a.delta0);                        // (x - 0) * delta0
            const int i = __double2loint(__dadd_rd(fi, HX_MAGIC));           // int(i), i >= 0
            if (i >= nbin) continue;                                         // i < bins[0] fails (x rounded to 1)
            if (jr >= 0 && jr < QF_WROWS)
                atomicAdd(&sh[c * wsize + jr * nbin + i], 1u);
            else
                atomicAdd(&a.prof[((size_t)c * nprof + j) * nbin + i], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C * wsize; i += blockDim.x) {
        uint32_t v = sh[i];
        if (v) {
            int c = i / wsize, rem = i - c * wsize;
            int j = j_base + rem / nbin;
            if (j < nprof) atomicAdd(&a.prof[((size_t)c * nprof + j) * nbin + (rem % nbin)], v);
        }
    }
}

// One CTA per (centre, LQ pair).  shifts computed cooperatively, then every thread accumulates one
// phase bin over a subset of time slices; integer accumulation is exact (the reference adds the same
// integers in fp64, also exactly).
__global__ void __launch_bounds__(256) qffa_step_kernel(const uint32_t *__restrict__ prof, int nprof, int nbin,
                                                        const double *__restrict__ L, const double *__restrict__ Q,
                                                        int n_lq, const double *__restrict__ base_shift,
                                                        const double *__restrict__ tab, int nharm,
                                                        double *__restrict__ stats) {
    extern __shared__ unsigned char qsm[];
    int *shift = reinterpret_cast<int *>(qsm);                                // [nprof]
    unsigned long long *splat = reinterpret_cast<unsigned long long *>(qsm + (((size_t)nprof * 4 + 15) & ~15ull));  // [nbin]
    double *red = reinterpret_cast<double *>(splat + nbin);                   // [2*32]
    const int c = blockIdx.y, lq = blockIdx.x;
    const double l = L[lq], q = Q[lq];
    const bool p2 = (nbin & (nbin - 1)) == 0;
    for (int k = threadIdx.x; k < nprof; k += blockDim.x) {
        const double b = base_shift[k];
        const double tot = __dadd_rn(__dmul_rn(b, l), __dmul_rn(__dmul_rn(b, b), q));
        const double r = __dsub_rn(__dadd_rn(tot, HX_MAGIC), HX_MAGIC);  // np.rint: half to even
        int s = (int)r;
        s = p2 ? (s & (nbin - 1)) : ((s % nbin) + nbin) % nbin;  // np.mod -> [0, nbin)
        shift[k] = s;
    }
    for (int b = threadIdx.x; b < nbin; b += blockDim.x) splat[b] = 0ull;
    __syncthreads();
    const uint32_t *pc = prof + (size_t)c * nprof * nbin;
    // thread -> (bin, slice group)
    const int groups = max(1, (int)blockDim.x / nbin);
    for (int b0 = 0; b0 < nbin; b0 += blockDim.x) {
        const int b = b0 + (int)threadIdx.x % max(nbin, 1);
        const int g = (int)threadIdx.x / nbin;
        if (b < nbin && g < groups) {
            unsigned long long acc = 0ull;
            for (int k = g; k < nprof; k += groups) {
                int src = b - shift[k];
                if (src < 0) src += nbin;
                acc += pc[(size_t)k * nbin + src];
            }
            atomicAdd(&splat[b], acc);
        }
        if (nbin >= (int)blockDim.x) continue;
        break;
    }
    __syncthreads();
    // Z^2_n of splat (hendrics/ffa.py:59-104)
    double tot = 0.0;
    for (int b = threadIdx.x; b < nbin; b += blockDim.x) tot += (double)splat[b];
    double res = 0.0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    auto block_sum2 = [&](double &x, double &y) {
        for (int o = 16; o; o >>= 1) {
            x += __shfl_xor_sync(0xffffffffu, x, o);
            y += __shfl_xor_sync(0xffffffffu, y, o);
        }
        __syncthreads();
        if (lane == 0) red[warp] = x, red[32 + warp] = y;
        __syncthreads();
        x = 0.0, y = 0.0;
        for (int w = 0; w < nw; w++) x += red[w], y += red[32 + w];
    };
    double dummy = 0.0;
    block_sum2(tot, dummy);
    for (int k = 1; k <= nharm; k++) {
        double cs = 0.0, sn = 0.0;
        for (int b = threadIdx.x; b < nbin; b += blockDim.x) {
            int idx = (int)(((long long)b * k) % nbin);
            double p = (double)splat[b];
            cs += tab[idx] * p;
            sn += tab[nbin + idx] * p;
        }
        block_sum2(cs, sn);
        res += cs * cs + sn * sn;
    }
    if (threadIdx.x == 0) stats[(size_t)c * n_lq + lq] = 2.0 / tot * res;
}

typedef void (*qhist_fn_t)(const QffaHistArgs);
template <int C>
static qhist_fn_t pick_qhist_c(int mode) {
    return mode == 0 ? qffa_hist_kernel<C, 0> : mode == 1 ? qffa_hist_kernel<C, 1> : qffa_hist_kernel<C, 2>;
}
static qhist_fn_t pick_qhist(int C, int mode) {
    switch (C) {
        case 1: return pick_qhist_c<1>(mode);
        case 2: return pick_qhist_c<2>(mode);
        case 3: return pick_qhist_c<3>(mode);
        case 4: return pick_qhist_c<4>(mode);
        case 5: return pick_qhist_c<5>(mode);
        case 6: return pick_qhist_c<6>(mode);
        case 7: return pick_qhist_c<7>(mode);
        default: return pick_qhist_c<8>(mode);
    }
}

static int g_qffa_cmax = QF_CMAX;  // tuning hook

int qffa_search_dev_impl(const double *d_t, int64_t n_ev, double t_first, double t_last, const double *mean_f,
                         const double *mean_fd, const double *mean_fdd, int64_t n_centres, int nbin, int nprof,
                         const double *L, const double *Q, int64_t n_lq, int nharm, double *d_stats,
                         uint32_t *d_profiles, cudaStream_t st) {
    if (!d_t || !mean_f || !mean_fd || !mean_fdd || !L || !Q || !d_stats) {
        set_error("hx_qffa_search: null pointer argument");
        return HX_ERR_ARG;
    }
    if (n_ev < 1 || n_centres < 0 || n_lq < 1 || nbin < 2 || nbin > HX_MAX_NBIN || nprof < 1 || nprof > (1 << 20) ||
        nharm < 1 || nharm > HX_MAX_NHARM || n_lq > 65535) {
        set_error("hx_qffa_search: bad sizes (n_ev=%lld n_centres=%lld n_lq=%lld nbin=%d nprof=%d nharm=%d)",
                  (long long)n_ev, (long long)n_centres, (long long)n_lq, nbin, nprof, nharm);
        return HX_ERR_ARG;
    }
    if (n_centres == 0) return HX_OK;
    const int cmax = std::max(1, std::min(g_qffa_cmax, QF_CMAX));
    // host-side constants, reference formulas
    const double delta0 = 1 / ((1.0 - 0.0) / nbin);
    const double delta1 = 1 / ((t_last - t_first) / nprof);
    std::vector<double> host((size_t)nprof + 2 * (size_t)nbin + 2 * (size_t)n_lq);
    double *h_base = host.data(), *h_tab = h_base + nprof, *h_L = h_tab + 2 * nbin, *h_Q = h_L + n_lq;
    {
        // np.linspace(-1, 1, nprof)
        const double step = nprof > 1 ? 2.0 / (nprof - 1) : 0.0;
        for (int k = 0; k < nprof; k++) h_base[k] = -1.0 + k * step;
        if (nprof > 1) h_base[nprof - 1] = 1.0;
        // 2*pi*arange(0, 1, 1/nbin)   (efsearch.py:819)
        const double ps = 1.0 / nbin;
        for (int j = 0; j < nbin; j++) {
            double tw = (2 * M_PI) * (j * ps);
            h_tab[j] = cos(tw);
            h_tab[nbin + j] = sin(tw);
        }
        for (int64_t i = 0; i < n_lq; i++) h_L[i] = L[i], h_Q[i] = Q[i];
    }
    void *d_const = nullptr, *d_scratch = nullptr;
    HX_CUDA_TRY(cudaMallocAsync(&d_const, host.size() * sizeof(double), st));
    struct Free {
        void *&p;
        cudaStream_t s;
        ~Free() {
            if (p) cudaFreeAsync(p, s);
        }
    } f1{d_const, st}, f2{d_scratch, st};
    HX_CUDA_TRY(cudaMemcpyAsync(d_const, host.data(), host.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    const double *d_base = reinterpret_cast<double *>(d_const), *d_tab = d_base + nprof, *d_L = d_tab + 2 * nbin,
                 *d_Q = d_L + n_lq;
    const size_t prof_elems = (size_t)nprof * nbin;
    if (!d_profiles) HX_CUDA_TRY(cudaMallocAsync(&d_scratch, prof_elems * cmax * sizeof(uint32_t), st));

    const size_t step_smem = (((size_t)nprof * 4 + 15) & ~15ull) + (size_t)nbin * 8 + 64 * 8;
    if (step_smem > 200 * 1024) {
        set_error("hx_qffa_search: nprof=%d too large for the shift table in shared memory", nprof);
        return HX_ERR_ARG;
    }
    HX_CUDA_TRY(cudaFuncSetAttribute(qffa_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)step_smem));

    for (int64_t c0 = 0; c0 < n_centres; c0 += cmax) {
        const int C = (int)std::min<int64_t>(cmax, n_centres - c0);
        uint32_t *prof = d_profiles ? d_profiles + (size_t)c0 * prof_elems : reinterpret_cast<uint32_t *>(d_scratch);
        HX_CUDA_TRY(cudaMemsetAsync(prof, 0, prof_elems * C * sizeof(uint32_t), st));
        QffaHistArgs a{};
        a.t = d_t, a.n_ev = n_ev, a.t_first = t_first, a.delta0 = delta0, a.delta1 = delta1;
        a.nbin = nbin, a.nprof = nprof, a.prof = prof;
        int mode = 0;
        for (int c = 0; c < C; c++) {
            a.c.f[c] = mean_f[c0 + c], a.c.fd[c] = mean_fd[c0 + c], a.c.fdd[c] = mean_fdd[c0 + c];
            // efsearch.py:892-897 picks the formula per step; adding exact-zero terms is value-neutral
            if (a.c.fdd[c] != 0) mode = 2;
            else if (a.c.fd[c] != 0) mode = std::max(mode, 1);
        }
        const long long chunk = (long long)QF_THREADS * QF_EV_PER_THREAD;
        const unsigned grid = (unsigned)((n_ev + chunk - 1) / chunk);
        const size_t hsmem = (size_t)C * QF_WROWS * nbin * sizeof(uint32_t);
        qhist_fn_t fn = pick_qhist(C, mode);
        HX_CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hsmem));
        fn<<<grid, QF_THREADS, hsmem, st>>>(a);
        HX_CHECK_LAUNCH("qffa_hist_kernel");
        dim3 g2((unsigned)n_lq, (unsigned)C);
        qffa_step_kernel<<<g2, 256, step_smem, st>>>(prof, nprof, nbin, d_L, d_Q, (int)n_lq, d_base, d_tab, nharm,
                                                      d_stats + (size_t)c0 * n_lq);
        HX_CHECK_LAUNCH("qffa_step_kernel");
    }
    HX_CUDA_TRY(cudaStreamSynchronize(st));
    return HX_OK;
}

}  // namespace hx

extern "C" {

int hx_qffa_set_tuning(int cmax) {
    hx::g_qffa_cmax = (cmax >= 1 && cmax <= hx::QF_CMAX) ? cmax : hx::QF_CMAX;
    return HX_OK;
}

int hx_qffa_search_dev(const double *d_times_c, int64_t n_ev, double t_first, double t_last, const double *mean_f,
                       const double *mean_fd, const double *mean_fdd, int64_t n_centres, int nbin, int nprof,
                       const double *L, const double *Q, int64_t n_lq, int nharm, double *d_stats,
                       uint32_t *d_profiles, void *stream) {
    return hx::qffa_search_dev_impl(d_times_c, n_ev, t_first, t_last, mean_f, mean_fd, mean_fdd, n_centres, nbin,
                                    nprof, L, Q, n_lq, nharm, d_stats, d_profiles,
                                    reinterpret_cast<cudaStream_t>(stream));
}

int hx_qffa_search(const double *times_c, int64_t n_ev, const double *mean_f, const double *mean_fd,
                   const double *mean_fdd, int64_t n_centres, int nbin, int nprof, const double *L, const double *Q,
                   int64_t n_lq, int nharm, double *stats, uint32_t *profiles, int device) {
    if (!times_c || !stats || n_ev < 1 || n_centres < 0 || n_lq < 1) {
        hx::set_error("hx_qffa_search: bad arguments");
        return HX_ERR_ARG;
    }
    HX_CUDA_TRY(cudaSetDevice(device));
    hx::DevBuf dt, ds, dp;
    HX_CUDA_TRY(dt.alloc((size_t)n_ev * 8));
    HX_CUDA_TRY(ds.alloc((size_t)n_centres * n_lq * 8));
    const size_t pe = (size_t)n_centres * nprof * nbin;
    if (profiles) HX_CUDA_TRY(dp.alloc(pe * 4));
    HX_CUDA_TRY(cudaMemcpy(dt.p, times_c, (size_t)n_ev * 8, cudaMemcpyHostToDevice));
    int rc = hx::qffa_search_dev_impl(dt.as<double>(), n_ev, times_c[0], times_c[n_ev - 1], mean_f, mean_fd, mean_fdd,
                                      n_centres, nbin, nprof, L, Q, n_lq, nharm, ds.as<double>(),
                                      profiles ? dp.as<uint32_t>() : nullptr, 0);
    if (rc) return rc;
    HX_CUDA_TRY(cudaMemcpy(stats, ds.p, (size_t)n_centres * n_lq * 8, cudaMemcpyDeviceToHost));
    if (profiles) HX_CUDA_TRY(cudaMemcpy(profiles, dp.p, pe * 4, cudaMemcpyDeviceToHost));
    return HX_OK;
}

}  // extern "C"
