// K4: unbinned Z^2_n over an (f, fdot) trial grid (extension; `binned=False`).
//
// Oracle formula: hendrics/efsearch.py:366-404 (z_n_fast) with unit weights:
//     Z^2_n = 2/N * sum_{k=1..n} [ (sum_i cos k*phi_i)^2 + (sum_i sin k*phi_i)^2 ],  phi_i = 2 pi frac(t_i f + 0.5 t_i^2 fd)
// The phase is the reference's fp64 phase (separately rounded products, one rounded sum); the
// harmonics k >= 2 come from the angle-addition recurrence on (cos phi, sin phi) instead of n
// sincos calls, as BASELINE.json's north_star sketches.  lane = trial, events are staged through
// shared memory and broadcast; per-trial sums live in registers and are combined across event
// splits with fp64 atomics.
#include <algorithm>
#include <cmath>

#include "hx_common.cuh"

namespace hx {

constexpr int ZN_TILE = 1024;
constexpr int ZN_THREADS = 128;
constexpr int ZN_MAX_NH = 8;

// sin/cos(2 pi x) for x in [0, 1]: quadrant by rint(4x) (the remainder r = x - k/4 is exact), then the
// Taylor polynomials of sin/cos on |theta| <= pi/4 (truncation < 5e-17).  ~22 FP64 instructions instead
// of a library sincospi call; this kernel is FP64-pipe-bound, so it is most of its cost.
__device__ __forceinline__ void sincos2pi(double x, double &s, double &c) {
    const double kd = __dsub_rn(__dadd_rn(__dmul_rn(4.0, x), HX_MAGIC), HX_MAGIC);  // rint(4x) in {0..4}
    const int q = (int)kd;
    const double r = __fma_rn(-0.25, kd, x);                     // exact, |r| <= 1/8
    const double th = __dmul_rn(r, 6.283185307179586476925287);  // |theta| <= pi/4
    const double t2 = __dmul_rn(th, th);
    double ps = -1.0 / 1307674368000.0;
    ps = __fma_rn(ps, t2, 1.0 / 6227020800.0);
    ps = __fma_rn(ps, t2, -1.0 / 39916800.0);
    ps = __fma_rn(ps, t2, 1.0 / 362880.0);
    ps = __fma_rn(ps, t2, -1.0 / 5040.0);
    ps = __fma_rn(ps, t2, 1.0 / 120.0);
    ps = __fma_rn(ps, t2, -1.0 / 6.0);
    const double s0 = __fma_rn(__dmul_rn(th, t2), ps, th);
    double pc = 1.0 / 20922789888000.0;
    pc = __fma_rn(pc, t2, -1.0 / 87178291200.0);
    pc = __fma_rn(pc, t2, 1.0 / 479001600.0);
    pc = __fma_rn(pc, t2, -1.0 / 3628800.0);
    pc = __fma_rn(pc, t2, 1.0 / 40320.0);
    pc = __fma_rn(pc, t2, -1.0 / 720.0);
    pc = __fma_rn(pc, t2, 1.0 / 24.0);
    pc = __fma_rn(pc, t2, -0.5);
    const double c0 = __fma_rn(t2, pc, 1.0);
    // angle = k pi/2 + theta
    double ss = (q & 1) ? c0 : s0, cc = (q & 1) ? -s0 : c0;
    if (q & 2) ss = -ss, cc = -cc;
    s = ss, c = cc;
}

template <int NH, bool HAS_FD>
__global__ void __launch_bounds__(ZN_THREADS) zn_unbinned_kernel(const double *__restrict__ t, long long n_ev,
                                                                 const double *__restrict__ freqs, long long n_f,
                                                                 const double *__restrict__ fdots, long long n_trials,
                                                                 long long ev_per_split, double *__restrict__ sums) {
    __shared__ double st[ZN_TILE], sq[HAS_FD ? ZN_TILE : 1];
    const long long g = (long long)blockIdx.x * ZN_THREADS + threadIdx.x;
    const bool valid = g < n_trials;
    const long long gi = valid ? g : 0;
    const long long ifd = gi / n_f, jf = gi - ifd * n_f;
    const double f = freqs[jf], fd = HAS_FD ? fdots[ifd] : 0.0;
    double cs[NH], sn[NH];
#pragma unroll
    for (int k = 0; k < NH; k++) cs[k] = 0.0, sn[k] = 0.0;
    const long long e0 = (long long)blockIdx.y * ev_per_split, e1 = min(n_ev, e0 + ev_per_split);
    for (long long base = e0; base < e1; base += ZN_TILE) {
        const int cnt = (int)min((long long)ZN_TILE, e1 - base);
        __syncthreads();
        for (int i = threadIdx.x; i < cnt; i += ZN_THREADS) {
            double v = t[base + i];
            st[i] = v;
            if (HAS_FD) sq[i] = __dmul_rn(__dmul_rn(0.5, v), v);
        }
        __syncthreads();
#pragma unroll 2
        for (int i = 0; i < cnt; i++) {
            double ph = __dmul_rn(st[i], f);
            if (HAS_FD) ph = __dadd_rn(ph, __dmul_rn(sq[i], fd));
            ph = __dsub_rn(ph, __dsub_rn(__dadd_rd(ph, HX_MAGIC), HX_MAGIC));  // ph -= floor(ph)
            double s1, c1;
            sincos2pi(ph, s1, c1);
            double ck = c1, sk = s1;
            cs[0] += ck, sn[0] += sk;
#pragma unroll
            for (int k = 1; k < NH; k++) {  // angle addition: (ck + i sk) *= (c1 + i s1)
                const double cn = __fma_rn(ck, c1, -__dmul_rn(sk, s1));
                sk = __fma_rn(sk, c1, __dmul_rn(ck, s1));
                ck = cn;
                cs[k] += ck, sn[k] += sk;
            }
        }
    }
    if (valid) {
#pragma unroll
        for (int k = 0; k < NH; k++) {
            atomicAdd(&sums[(size_t)g * 2 * NH + 2 * k], cs[k]);
            atomicAdd(&sums[(size_t)g * 2 * NH + 2 * k + 1], sn[k]);
        }
    }
}

__global__ void zn_finish_kernel(const double *__restrict__ sums, long long n_trials, int nh, double n_ev,
                                 double *__restrict__ stats) {
    long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_trials) return;
    double res = 0.0;
    for (int k = 0; k < nh; k++) {
        double c = sums[(size_t)g * 2 * nh + 2 * k], s = sums[(size_t)g * 2 * nh + 2 * k + 1];
        res += c * c + s * s;
    }
    stats[g] = 2.0 / n_ev * res;
}

typedef void (*zn_fn_t)(const double *, long long, const double *, long long, const double *, long long, long long,
                        double *);
template <int NH>
static zn_fn_t pick_zn_nh(bool fd) {
    return fd ? zn_unbinned_kernel<NH, true> : zn_unbinned_kernel<NH, false>;
}
static zn_fn_t pick_zn(int nh, bool fd) {
    switch (nh) {
        case 1: return pick_zn_nh<1>(fd);
        case 2: return pick_zn_nh<2>(fd);
        case 3: return pick_zn_nh<3>(fd);
        case 4: return pick_zn_nh<4>(fd);
        case 5: return pick_zn_nh<5>(fd);
        case 6: return pick_zn_nh<6>(fd);
        case 7: return pick_zn_nh<7>(fd);
        default: return pick_zn_nh<8>(fd);
    }
}

int zn_unbinned_dev_impl(const double *d_t, int64_t n_ev, const double *d_f, int64_t n_f, const double *d_fd,
                         int64_t n_fd, int nharm, double *d_stats, bool has_fd, cudaStream_t st) {
    std::lock_guard<std::recursive_mutex> api_lock(api_mutex());
    if (!d_t || !d_f || !d_fd || !d_stats || n_ev < 0 || n_f < 0 || n_fd < 0) {
        set_error("hx_zn_unbinned: bad arguments");
        return HX_ERR_ARG;
    }
    if (nharm < 1 || nharm > ZN_MAX_NH) {
        set_error("hx_zn_unbinned: nharm must be in [1, %d] (got %d)", ZN_MAX_NH, nharm);
        return HX_ERR_ARG;
    }
    const long long n_trials = (long long)n_f * n_fd;
    if (n_trials == 0) return HX_OK;
    int device;
    HX_CUDA_TRY(cudaGetDevice(&device));
    DeviceInfo di;
    int rc = get_device_info(device, &di);
    if (rc) return rc;
    const long long n_cta_x = (n_trials + ZN_THREADS - 1) / ZN_THREADS;
    const long long slots = (long long)di.sm_count * 8;
    long long splits = 1;
    if (n_cta_x < slots && n_ev > 0)
        splits = std::min<long long>(std::max<long long>(1, slots / n_cta_x), std::max<long long>(1, n_ev / (2 * ZN_TILE)));
    splits = std::min<long long>(splits, 65535);
    long long per = n_ev > 0 ? (n_ev + splits - 1) / splits : 1;
    per = (per + ZN_TILE - 1) / ZN_TILE * ZN_TILE;
    splits = n_ev > 0 ? (n_ev + per - 1) / per : 1;
    void *d_sums = nullptr;
    const size_t sbytes = (size_t)n_trials * 2 * nharm * sizeof(double);
    HX_CUDA_TRY(cudaMallocAsync(&d_sums, sbytes, st));
    cudaError_t e = cudaMemsetAsync(d_sums, 0, sbytes, st);
    if (e == cudaSuccess) {
        dim3 grid((unsigned)n_cta_x, (unsigned)splits);
        pick_zn(nharm, has_fd)<<<grid, ZN_THREADS, 0, st>>>(d_t, n_ev, d_f, n_f, d_fd, n_trials, per,
                                                              reinterpret_cast<double *>(d_sums));
        e = cudaGetLastError();
        count_launch();
    }
    if (e == cudaSuccess) {
        zn_finish_kernel<<<(unsigned)((n_trials + 255) / 256), 256, 0, st>>>(reinterpret_cast<double *>(d_sums),
                                                                              n_trials, nharm, (double)n_ev, d_stats);
        e = cudaGetLastError();
        count_launch();
    }
    cudaFreeAsync(d_sums, st);
    if (e != cudaSuccess) {
        set_error("hx_zn_unbinned launch failed: %s", cudaGetErrorString(e));
        return HX_ERR_CUDA;
    }
    return HX_OK;
}

}  // namespace hx

extern "C" {

int hx_zn_unbinned_dev(const double *d_times, int64_t n_ev, const double *d_freqs, int64_t n_f,
                       const double *d_fdots, int64_t n_fd, int nharm, double *d_stats, void *stream) {
    // fdot values live on the device: always take the general (f, fdot) kernel
    return hx::zn_unbinned_dev_impl(d_times, n_ev, d_freqs, n_f, d_fdots, n_fd, nharm, d_stats, true,
                                    reinterpret_cast<cudaStream_t>(stream));
}

int hx_zn_unbinned(const double *times, int64_t n_ev, const double *freqs, int64_t n_f, const double *fdots,
                   int64_t n_fd, int nharm, double *stats, int device) {
    if (!times || !freqs || !fdots || !stats || n_ev < 0 || n_f < 0 || n_fd < 0) {
        hx::set_error("hx_zn_unbinned: bad arguments");
        return HX_ERR_ARG;
    }
    hx::DeviceGuard dev_guard;  // declared before the device buffers: they are freed before the device is restored
    HX_CUDA_TRY(dev_guard.enter(device));
    const size_t n_tr = (size_t)n_f * n_fd;
    if (n_tr == 0) return HX_OK;
    bool has_fd = false;
    for (int64_t i = 0; i < n_fd; i++) has_fd |= fdots[i] != 0.0;
    hx::DevBuf dt, df, dfd, ds;
    HX_CUDA_TRY(dt.alloc((size_t)n_ev * 8));
    HX_CUDA_TRY(df.alloc((size_t)n_f * 8));
    HX_CUDA_TRY(dfd.alloc((size_t)n_fd * 8));
    HX_CUDA_TRY(ds.alloc(n_tr * 8));
    HX_CUDA_TRY(cudaMemcpy(dt.p, times, (size_t)n_ev * 8, cudaMemcpyHostToDevice));
    HX_CUDA_TRY(cudaMemcpy(df.p, freqs, (size_t)n_f * 8, cudaMemcpyHostToDevice));
    HX_CUDA_TRY(cudaMemcpy(dfd.p, fdots, (size_t)n_fd * 8, cudaMemcpyHostToDevice));
    int rc = hx::zn_unbinned_dev_impl(dt.as<double>(), n_ev, df.as<double>(), n_f, dfd.as<double>(), n_fd, nharm,
                                      ds.as<double>(), has_fd, 0);
    if (rc) return rc;
    HX_CUDA_TRY(cudaStreamSynchronize(0));
    HX_CUDA_TRY(cudaMemcpy(stats, ds.p, n_tr * 8, cudaMemcpyDeviceToHost));
    return HX_OK;
}

}  // extern "C"
