// K4: unbinned Z^2_n over an (f, fdot) trial grid (extension; `binned=False`).
//
// Oracle formula: hendrics/efsearch.py:366-404 (z_n_fast) with unit weights:
//     Z^2_n = 2/N * sum_{k=1..n} [ (sum_i cos k*phi_i)^2 + (sum_i sin k*phi_i)^2 ],  phi_i = 2 pi frac(t_i f + 0.5 t_i^2 fd)
// This is the kernel BASELINE.json's north_star sketches: lane = trial, event times streamed through shared memory by
// the TMA engine and broadcast to the threads, the phase in fp64 exactly as the reference rounds it, ONE sin/cos
// evaluation per event*trial and the harmonics k >= 2 from the angle-addition recurrence, per-trial sums in registers.
//
// Design:
//   * same staging skeleton as K1 (hx_fold.cu): 2-stage cp.async.bulk + mbarrier pipeline of event tiles, no block-wide
//     barrier in the loop (the last warp to finish a stage refills it).  With fdot, the per-event term fl(fl(0.5 t) t)
//     is computed once per warp into a second tile instead of twice per event*trial.
//   * the reference's phase, rounding for rounding, but scaled by 4 (exact): y = fl(fl(t*4f) + fl(q*4fd)) = 4*phase.
//     n = rint(y) (one magic add + subtract), r = y - n exactly, |r| <= 1/2: the angle is n*pi/2 + r*pi/2, so the
//     quadrant is the low two bits of n and theta = r*pi/2 lies in [-pi/4, pi/4] -- no floor, no separate reduction
//     of the fractional phase: 8 FP64 instructions up to theta^2 (6 without fdot).
//   * sin / cos(theta) from their Taylor polynomials (degree 15 / 16, truncation < 5e-17): 16 FP64 instructions.
//     Per harmonic k >= 2: 4 (complex multiply) and per harmonic 2 accumulations.  Z^2_2 with fdot: 32 FP64
//     instructions per event*trial; the kernel is FP64-pipe-bound by construction (bench.py --workload zn).
//   * four events per loop iteration: independent dependency chains, and the polynomial constants (constant bank ->
//     uniform registers) are fetched once per iteration.
//   * the event list is split across CTAs when the trial grid alone cannot fill the GPU; every (split, trial) writes
//     its 2n partial sums to a buffer and zn_finish_kernel adds the splits IN ORDER: results are bit-reproducible
//     from run to run (no floating-point atomics).
#include <algorithm>
#include <cmath>

#include "hx_common.cuh"

namespace hx {

constexpr int ZN_TILE = 512;       // events per stage
constexpr int ZN_PAD = 2;          // slack read by the two-event loop
constexpr int ZN_WARPS = 4;
constexpr int ZN_MAX_NH = 8;
constexpr uint32_t ZN_HDR = 32;    // two mbarriers + two "warps done" counters

struct ZnArgs {
    const double *ev;
    long long n_ev;
    const double *freqs;
    long long n_f;
    const double *fdots;
    long long n_trials;
    long long ev_per_split;
    double *partial;  // [split][2*NH][n_trials]
};

// Taylor coefficients in CONSTANT memory: FP64 instructions take a constant-bank operand directly, whereas 64-bit
// immediates are materialised with two UMOVs each, every iteration (measured: 15 of 61 issue slots per event*trial).
__constant__ double ZN_SC[7] = {-1.0 / 1307674368000.0, 1.0 / 6227020800.0, -1.0 / 39916800.0, 1.0 / 362880.0,
                                -1.0 / 5040.0,          1.0 / 120.0,        -1.0 / 6.0};
__constant__ double ZN_CC[8] = {1.0 / 20922789888000.0, -1.0 / 87178291200.0, 1.0 / 479001600.0, -1.0 / 3628800.0,
                                1.0 / 40320.0,          -1.0 / 720.0,         1.0 / 24.0,        -0.5};
__constant__ double ZN_HALF_PI = 1.5707963267948966192313;

// sin / cos of n*pi/2 + r*pi/2 for |r| <= 1/2, given the quadrant bits of n
__device__ __forceinline__ void sincos_quadrant(double r, uint32_t q, double &s, double &c) {
    const double th = __dmul_rn(r, ZN_HALF_PI);  // |theta| <= pi/4
    const double t2 = __dmul_rn(th, th);
    double ps = ZN_SC[0];
#pragma unroll
    for (int i = 1; i < 7; i++) ps = __fma_rn(ps, t2, ZN_SC[i]);
    const double s0 = __fma_rn(__dmul_rn(th, t2), ps, th);
    double pc = ZN_CC[0];
#pragma unroll
    for (int i = 1; i < 8; i++) pc = __fma_rn(pc, t2, ZN_CC[i]);
    const double c0 = __fma_rn(t2, pc, 1.0);
    // quadrant 0: (s0, c0); 1: (c0, -s0); 2: (-s0, -c0); 3: (-c0, s0) -- integer selects and sign flips, no FP64 work
    const bool odd = q & 1u;
    unsigned long long sb = (unsigned long long)__double_as_longlong(odd ? c0 : s0);
    unsigned long long cb = (unsigned long long)__double_as_longlong(odd ? s0 : c0);
    sb ^= (unsigned long long)(q & 2u) << 62;
    cb ^= (unsigned long long)((q + 1u) & 2u) << 62;
    s = __longlong_as_double((long long)sb), c = __longlong_as_double((long long)cb);
}

template <int NH, bool HAS_FD>
__global__ void __launch_bounds__(32 * ZN_WARPS) zn_unbinned_kernel(const ZnArgs a) {
    __shared__ __align__(16) unsigned char smem[ZN_HDR + (HAS_FD ? 4 : 2) * (ZN_TILE + ZN_PAD) * 8];
    constexpr uint32_t TILE_BYTES = (ZN_TILE + ZN_PAD) * 8u;
    const uint32_t sbase = smem_u32(smem);
    const int tid = threadIdx.x, lane = tid & 31;
    int *done = reinterpret_cast<int *>(smem + 16);
    const uint32_t bar[2] = {sbase, sbase + 8};

    const long long g = (long long)blockIdx.x * blockDim.x + tid;
    const bool valid = g < a.n_trials;
    const long long gi = valid ? g : 0;
    const long long ifd = gi / a.n_f, jf = gi - ifd * a.n_f;
    const double f4 = __dmul_rn(4.0, a.freqs[jf]);                  // exact scalings: y below is 4 * (the reference's phase)
    const double fd4 = HAS_FD ? __dmul_rn(4.0, a.fdots[ifd]) : 0.0;

    const long long e0 = (long long)blockIdx.y * a.ev_per_split, e1 = min(a.n_ev, e0 + a.ev_per_split);
    const int ntiles = e1 > e0 ? (int)((e1 - e0 + ZN_TILE - 1) / ZN_TILE) : 0;
    auto issue = [&](int it) {
        const long long s = e0 + (long long)it * ZN_TILE;
        const uint32_t cnt = (uint32_t)min((long long)ZN_TILE, e1 - s);
        const uint32_t bytes = cnt * 8u, bulk = bytes & ~15u;
        const uint32_t off = ZN_HDR + (uint32_t)(it & 1) * TILE_BYTES;
        const char *src = reinterpret_cast<const char *>(a.ev) + s * 8;
        if (bytes != bulk) *reinterpret_cast<double *>(smem + off + bulk) = *reinterpret_cast<const double *>(src + bulk);
        if (bulk) {
            mbar_expect_tx(bar[it & 1], bulk);
            bulk_g2s(sbase + off, src, bulk, bar[it & 1]);
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

    double cs[NH], sn[NH];
#pragma unroll
    for (int k = 0; k < NH; k++) cs[k] = 0.0, sn[k] = 0.0;

    auto one_event = [&](double t, double q) {
        double y = __dmul_rn(t, f4);
        if (HAS_FD) y = __dadd_rn(y, __dmul_rn(q, fd4));
        const double yn = __dadd_rn(y, HX_MAGIC);          // rint(y) in the low mantissa bits
        const double r = __dsub_rn(y, __dsub_rn(yn, HX_MAGIC));  // exact, |r| <= 1/2
        double s1, c1;
        sincos_quadrant(r, (uint32_t)__double2loint(yn), s1, c1);
        double ck = c1, sk = s1;
        cs[0] = __dadd_rn(cs[0], ck), sn[0] = __dadd_rn(sn[0], sk);
#pragma unroll
        for (int k = 1; k < NH; k++) {  // (ck + i sk) *= (c1 + i s1)
            const double cn = __fma_rn(ck, c1, -__dmul_rn(sk, s1));
            sk = __fma_rn(sk, c1, __dmul_rn(ck, s1));
            ck = cn;
            cs[k] = __dadd_rn(cs[k], ck), sn[k] = __dadd_rn(sn[k], sk);
        }
    };

    for (int it = 0; it < ntiles; it++) {
        const int buf = it & 1;
        mbar_wait(bar[buf], (uint32_t)(it >> 1) & 1u);
        const int cnt = (int)min((long long)ZN_TILE, e1 - (e0 + (long long)it * ZN_TILE));
        const double *tp = reinterpret_cast<const double *>(smem + ZN_HDR + (uint32_t)buf * TILE_BYTES);
        double *qp = reinterpret_cast<double *>(smem + ZN_HDR + (uint32_t)(2 + buf) * TILE_BYTES);
        if (HAS_FD) {  // every warp writes the same values: no block barrier needed
            for (int i = lane; i < cnt; i += 32) qp[i] = __dmul_rn(__dmul_rn(0.5, tp[i]), tp[i]);
            __syncwarp();
        }
        int e = 0;
        for (; e + 4 <= cnt; e += 4) {  // four events per iteration: the polynomial constants are fetched once for all
            const double2 ta = *reinterpret_cast<const double2 *>(tp + e), tb = *reinterpret_cast<const double2 *>(tp + e + 2);
            double2 qa = make_double2(0.0, 0.0), qb = qa;
            if (HAS_FD) qa = *reinterpret_cast<const double2 *>(qp + e), qb = *reinterpret_cast<const double2 *>(qp + e + 2);
            one_event(ta.x, qa.x);
            one_event(ta.y, qa.y);
            one_event(tb.x, qb.x);
            one_event(tb.y, qb.y);
        }
        for (; e + 1 < cnt; e += 2) {
            const double2 tv = *reinterpret_cast<const double2 *>(tp + e);
            double2 qv = make_double2(0.0, 0.0);
            if (HAS_FD) qv = *reinterpret_cast<const double2 *>(qp + e);
            one_event(tv.x, qv.x);
            one_event(tv.y, qv.y);
        }
        if (e < cnt) one_event(tp[e], HAS_FD ? qp[e] : 0.0);
        __syncwarp();
        if (lane == 0) {
            __threadfence_block();  // this warp's reads of the stage are ordered before its hand-back
            const int prev = atomicAdd(&done[buf], 1);
            if (prev == ZN_WARPS - 1) {
                done[buf] = 0;
                if (it + 2 < ntiles) {
                    fence_proxy_async();
                    issue(it + 2);
                }
            }
        }
    }
    if (valid) {
        double *dst = a.partial + (size_t)blockIdx.y * 2 * NH * a.n_trials + g;
#pragma unroll
        for (int k = 0; k < NH; k++) {
            dst[(size_t)(2 * k) * a.n_trials] = cs[k];
            dst[(size_t)(2 * k + 1) * a.n_trials] = sn[k];
        }
    }
}

// adds the event splits in order and forms 2/N * sum_k (C_k^2 + S_k^2)
__global__ void zn_finish_kernel(const double *__restrict__ partial, long long n_trials, int nh, int splits, double n_ev,
                                 double *__restrict__ stats) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_trials) return;
    double res = 0.0;
    for (int k = 0; k < nh; k++) {
        double c = 0.0, s = 0.0;
        for (int sp = 0; sp < splits; sp++) {
            const double *p = partial + (size_t)sp * 2 * nh * n_trials + g;
            c += p[(size_t)(2 * k) * n_trials];
            s += p[(size_t)(2 * k + 1) * n_trials];
        }
        res += c * c + s * s;
    }
    stats[g] = 2.0 / n_ev * res;
}

typedef void (*zn_fn_t)(const ZnArgs);
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

static cudaEvent_t g_zn_ev0[64], g_zn_ev1[64];
static bool g_zn_ev_ok[64] = {false};
static int g_zn_last_device = -1;
static int g_zn_last_fp64_ops = 0;

int zn_unbinned_dev_impl(const double *d_t, int64_t n_ev, const double *d_f, int64_t n_f, const double *d_fd,
                         int64_t n_fd, int nharm, double *d_stats, cudaStream_t st) {
    std::lock_guard<std::recursive_mutex> api_lock(api_mutex());
    if (!d_t || !d_f || !d_fd || !d_stats || n_ev < 0 || n_f < 0 || n_fd < 0) {
        set_error("hx_zn_unbinned: bad arguments");
        return HX_ERR_ARG;
    }
    if (nharm < 1 || nharm > ZN_MAX_NH) {
        set_error("hx_zn_unbinned: nharm must be in [1, %d] (got %d)", ZN_MAX_NH, nharm);
        return HX_ERR_ARG;
    }
    if ((reinterpret_cast<uintptr_t>(d_t) & 15) != 0) {
        set_error("hx_zn_unbinned_dev: d_times must be 16-byte aligned");
        return HX_ERR_ARG;
    }
    const long long n_trials = (long long)n_f * n_fd;
    if (n_trials == 0) return HX_OK;
    int device;
    HX_CUDA_TRY(cudaGetDevice(&device));
    DeviceInfo di;
    int rc = get_device_info(device, &di);
    if (rc) return rc;
    // finite inputs, and 4 * |phase| inside the range of the magic-add rounding (2^51)
    double r6[6];
    if ((rc = device_ranges(d_t, n_ev, d_f, n_f, d_fd, n_fd, r6, st)) != HX_OK) return rc;
    const double ta = std::max(std::fabs(r6[0]), std::fabs(r6[1])), fa = std::max(std::fabs(r6[2]), std::fabs(r6[3]));
    const double fda = std::max(std::fabs(r6[4]), std::fabs(r6[5]));
    const bool has_fd = fda != 0.0;
    const double phmax = ta * fa + 0.5 * ta * ta * fda;
    if (!(phmax < 281474976710656.0 /* 2^48 */)) {
        set_error("hx_zn_unbinned: |phase| up to %.3g cycles exceeds the fp64 range of this kernel (2^48)", phmax);
        return HX_ERR_RANGE;
    }

    const int threads = 32 * ZN_WARPS;
    const long long n_cta_x = (n_trials + threads - 1) / threads;
    if (n_cta_x > 2147483647LL) {
        set_error("hx_zn_unbinned: trial grid too large");
        return HX_ERR_ARG;
    }
    // fill the GPU: split the events when there are few trial blocks (each split costs 2*nharm doubles per trial)
    const long long slots = (long long)di.sm_count * 8;
    long long splits = 1;
    if (n_cta_x < 2 * slots && n_ev > 0)
        splits = std::min<long long>(std::max<long long>(1, (2 * slots + n_cta_x - 1) / n_cta_x),
                                     std::max<long long>(1, n_ev / (4 * ZN_TILE)));
    splits = std::min<long long>(splits, 4096);
    long long per = n_ev > 0 ? (n_ev + splits - 1) / splits : ZN_TILE;
    per = (per + ZN_TILE - 1) / ZN_TILE * ZN_TILE;
    splits = n_ev > 0 ? (n_ev + per - 1) / per : 1;

    AsyncBuf b_part(st);
    HX_CUDA_TRY(b_part.alloc((size_t)splits * 2 * nharm * n_trials * sizeof(double)));
    ZnArgs a{d_t, n_ev, d_f, n_f, d_fd, n_trials, per, b_part.as<double>()};
    if (device >= 0 && device < 64) {
        if (!g_zn_ev_ok[device]) {
            HX_CUDA_TRY(cudaEventCreate(&g_zn_ev0[device]));
            HX_CUDA_TRY(cudaEventCreate(&g_zn_ev1[device]));
            g_zn_ev_ok[device] = true;
        }
        HX_CUDA_TRY(cudaEventRecord(g_zn_ev0[device], st));
    }
    dim3 grid((unsigned)n_cta_x, (unsigned)splits);
    pick_zn(nharm, has_fd)<<<grid, threads, 0, st>>>(a);
    HX_CHECK_LAUNCH("zn_unbinned_kernel");
    if (device >= 0 && device < 64) {
        HX_CUDA_TRY(cudaEventRecord(g_zn_ev1[device], st));
        g_zn_last_device = device;
        g_zn_last_fp64_ops = (has_fd ? 8 : 6) + 16 + 2 * nharm + 4 * (nharm - 1);
    }
    zn_finish_kernel<<<(unsigned)((n_trials + 255) / 256), 256, 0, st>>>(b_part.as<double>(), n_trials, nharm, (int)splits,
                                                                          (double)n_ev, d_stats);
    HX_CHECK_LAUNCH("zn_finish_kernel");
    return HX_OK;
}

}  // namespace hx

extern "C" {

int hx_zn_unbinned_dev(const double *d_times, int64_t n_ev, const double *d_freqs, int64_t n_f,
                       const double *d_fdots, int64_t n_fd, int nharm, double *d_stats, void *stream) {
    return hx::zn_unbinned_dev_impl(d_times, n_ev, d_freqs, n_f, d_fdots, n_fd, nharm, d_stats,
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
    hx::DevBuf dt, df, dfd, ds;
    HX_CUDA_TRY(dt.alloc((size_t)n_ev * 8));
    HX_CUDA_TRY(df.alloc((size_t)n_f * 8));
    HX_CUDA_TRY(dfd.alloc((size_t)n_fd * 8));
    HX_CUDA_TRY(ds.alloc(n_tr * 8));
    HX_CUDA_TRY(cudaMemcpy(dt.p, times, (size_t)n_ev * 8, cudaMemcpyHostToDevice));
    HX_CUDA_TRY(cudaMemcpy(df.p, freqs, (size_t)n_f * 8, cudaMemcpyHostToDevice));
    HX_CUDA_TRY(cudaMemcpy(dfd.p, fdots, (size_t)n_fd * 8, cudaMemcpyHostToDevice));
    int rc = hx::zn_unbinned_dev_impl(dt.as<double>(), n_ev, df.as<double>(), n_f, dfd.as<double>(), n_fd, nharm,
                                      ds.as<double>(), 0);
    if (rc) return rc;
    HX_CUDA_TRY(cudaStreamSynchronize(0));
    HX_CUDA_TRY(cudaMemcpy(stats, ds.p, n_tr * 8, cudaMemcpyDeviceToHost));
    return HX_OK;
}

// Duration (ms) of the last zn_unbinned_kernel launch on the current device and the FP64 instructions it issues per
// event*trial (bench.py's live roofline); synchronizes on the launch's end event.
int hx_zn_last_kernel_ms(double *ms, int *fp64_ops_per_unit) {
    std::lock_guard<std::recursive_mutex> api_lock(hx::api_mutex());
    const int d = hx::g_zn_last_device;
    if (!ms || d < 0 || !hx::g_zn_ev_ok[d]) {
        hx::set_error("hx_zn_last_kernel_ms: no timed launch");
        return HX_ERR_ARG;
    }
    HX_CUDA_TRY(cudaEventSynchronize(hx::g_zn_ev1[d]));
    float f = 0.f;
    HX_CUDA_TRY(cudaEventElapsedTime(&f, hx::g_zn_ev0[d], hx::g_zn_ev1[d]));
    *ms = (double)f;
    if (fp64_ops_per_unit) *fp64_ops_per_unit = hx::g_zn_last_fp64_ops;
    return HX_OK;
}

}  // extern "C"
