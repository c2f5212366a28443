// Shared helpers for libhxsearch.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <cmath>
#include <mutex>
#include <vector>

#include "../../include/hx_search.h"

namespace hx {

// ---- error plumbing ------------------------------------------------------------------------
void set_error(const char *fmt, ...);
void count_launch(int n = 1);
// The library keeps per-device planning / timing state (last-kernel events, tuning hooks): compute entry points
// serialise on this lock, so concurrent calls from several host threads are safe (they run one after the other).
std::recursive_mutex &api_mutex();

#define HX_CUDA_TRY(expr)                                                                   \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            hx::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                          __LINE__);                                                        \
            return (_e == cudaErrorMemoryAllocation) ? HX_ERR_NOMEM : HX_ERR_CUDA;          \
        }                                                                                   \
    } while (0)

#define HX_CHECK_LAUNCH(name)                                                          \
    do {                                                                               \
        cudaError_t _e = cudaGetLastError();                                           \
        if (_e != cudaSuccess) {                                                       \
            hx::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));   \
            return HX_ERR_CUDA;                                                        \
        }                                                                              \
        hx::count_launch();                                                            \
    } while (0)

// Host-pointer entry points run on the device they are given and put the calling thread back on the device it
// was on, on every exit path (a multi-GPU process keeps one current device per rank: it must not be moved to cuda:0
// behind torch's / NCCL's back).
struct DeviceGuard {
    int prev = -1;
    cudaError_t enter(int device) {
        cudaError_t e = cudaGetDevice(&prev);
        if (e != cudaSuccess) {
            prev = -1;
            return e;
        }
        return cudaSetDevice(device);
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// RAII device buffer for the host-pointer entry points.
struct DevBuf {
    void *p = nullptr;
    ~DevBuf() {
        if (p) cudaFree(p);
    }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 16); }
    template <typename T>
    T *as() {
        return reinterpret_cast<T *>(p);
    }
};

struct DeviceInfo {
    int sm_count;
    int smem_optin;        // max dynamic shared memory per block (opt-in)
    int smem_per_sm;       // shared memory per SM
    int reserved_smem;     // reserved shared memory per block
    uint32_t dyn_smem_base;  // shared-window address at which dynamic smem starts (no static smem)
};
// Cached per device; probes dyn_smem_base with a tiny kernel on first use.
int get_device_info(int device, DeviceInfo *out);
// [tmin,tmax, fmin,fmax, fdmin,fdmax] of device arrays (hx_fold.cu); HX_ERR_ARG when not finite; synchronizes
int device_ranges(const double *d_times, int64_t n_ev, const double *d_freqs, int64_t n_f, const double *d_fdots,
                  int64_t n_fd, double out[6], cudaStream_t st);

// stream-ordered temporary, released on scope exit
struct AsyncBuf {
    void *p = nullptr;
    cudaStream_t s;
    explicit AsyncBuf(cudaStream_t st) : s(st) {}
    ~AsyncBuf() {
        if (p) cudaFreeAsync(p, s);
    }
    cudaError_t alloc(size_t bytes) { return cudaMallocAsync(&p, bytes ? bytes : 16, s); }
    template <typename T>
    T *as() {
        return reinterpret_cast<T *>(p);
    }
};

// cos/sin tables: 2*pi*arange(0, 1, 1/nbin) as hendrics/ffa.py:40,51
inline void host_tables(int nbin, std::vector<double> &tab) {
    tab.resize(2 * (size_t)nbin);
    const double step = 1.0 / nbin;
    for (int j = 0; j < nbin; j++) {
        double tw = (M_PI * 2) * (j * step);
        tab[j] = cos(tw);
        tab[nbin + j] = sin(tw);
    }
}

// ---- device-side PTX helpers -----------------------------------------------------------------
#ifdef __CUDACC__

// 1.5 * 2^52: adding it with round-toward-minus-infinity leaves floor(x) in the low mantissa bits
// (two's complement) for |x| < 2^51.
#define HX_MAGIC 6755399441055744.0

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v));
}
__device__ __forceinline__ void red_shared_add_u32(uint32_t addr, uint32_t v) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(v));
}
__device__ __forceinline__ void red_shared_inc(uint32_t addr) {
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr));
}

// mbarrier + 1-D bulk async copy (TMA engine, SASS UBLKCP)
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy; dst/src 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
// chi^2 / Z^2_n of one profile; `cnt(j)` returns the count of bin j
template <typename F>
__device__ __forceinline__ double profile_stat(F cnt, int nbin, int nharm, const double *ctab, const double *stab) {
    double tot = 0.0;
    for (int j = 0; j < nbin; j++) tot += (double)cnt(j);
    if (nharm == 0) {
        double mean = tot / (double)nbin;
        double chi = 0.0;
        for (int j = 0; j < nbin; j++) {
            double d = (double)cnt(j) - mean;
            chi += d * d / mean;
        }
        return chi;
    }
    double res = 0.0;
    const bool p2 = (nbin & (nbin - 1)) == 0;
    for (int k = 1; k <= nharm; k++) {
        double c = 0.0, s = 0.0;
        int idx = 0;
        for (int j = 0; j < nbin; j++) {
            double p = (double)cnt(j);
            c += ctab[idx] * p;
            s += stab[idx] * p;
            idx += k;
            if (p2)
                idx &= nbin - 1;
            else
                idx %= nbin;
        }
        res += c * c + s * s;
    }
    return 2.0 / tot * res;
}

#endif  // __CUDACC__

}  // namespace hx
