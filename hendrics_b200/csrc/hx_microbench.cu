// Issue-rate microbenchmarks: the measured denominators of the K1 roofline (SURVEY.md 8d asks for
// the FP64 pipe peak to be measured on the box, MEASURED_PEAKS.json has no FP64 entry).
#include <algorithm>

#include "hx_common.cuh"

namespace hx {

constexpr int MB_ITERS = 4096;
constexpr int MB_ILP = 8;

// kind 0: DFMA, kind 1: DMUL + DADD (what the parity kernels must issue: no contraction)
template <int KIND>
__global__ void __launch_bounds__(256) mb_fp64_kernel(double *out, double a, double b) {
    double x[MB_ILP];
#pragma unroll
    for (int i = 0; i < MB_ILP; i++) x[i] = a + (double)(threadIdx.x + i);
    for (int it = 0; it < MB_ITERS; it++) {
#pragma unroll
        for (int i = 0; i < MB_ILP; i++) {
            if (KIND == 0) {
                x[i] = __fma_rn(x[i], a, b);
            } else {
                x[i] = __dadd_rn(__dmul_rn(x[i], a), b);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < MB_ILP; i++) s += x[i];
    if (s == 1.2345e-300) out[0] = s;  // never true; keeps the chain alive
}

// kind 2: LDS + IADD + STS on private [bin][lane] columns (conflict-free); kind 3: red.shared.add
template <int KIND>
__global__ void __launch_bounds__(256) mb_smem_kernel(uint32_t *out, uint32_t seed) {
    extern __shared__ __align__(16) unsigned char dsm[];
    const uint32_t base = smem_u32(dsm);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // 4 columns of 16 bins per thread: block of 16*128 B per (warp, column)
    uint32_t cb[4];
    for (int c = 0; c < 4; c++) cb[c] = base + (uint32_t)(warp * 4 + c) * 2048u + lane * 4u;
    for (int c = 0; c < 4; c++)
        for (int j = 0; j < 16; j++) sts_u32(cb[c] + j * 128, 0u);
    uint32_t r = seed * 747796405u + threadIdx.x * 2891336453u + 1u;
    for (int it = 0; it < MB_ITERS; it++) {
        r = r * 1664525u + 1013904223u;
        uint32_t a0 = cb[0] + ((r >> 4) & 0x780u), a1 = cb[1] + ((r >> 9) & 0x780u);
        uint32_t a2 = cb[2] + ((r >> 14) & 0x780u), a3 = cb[3] + ((r >> 19) & 0x780u);
        if (KIND == 2) {
            uint32_t c0 = lds_u32(a0), c1 = lds_u32(a1), c2 = lds_u32(a2), c3 = lds_u32(a3);
            sts_u32(a0, c0 + 1), sts_u32(a1, c1 + 1), sts_u32(a2, c2 + 1), sts_u32(a3, c3 + 1);
        } else {
            red_shared_inc(a0), red_shared_inc(a1), red_shared_inc(a2), red_shared_inc(a3);
        }
    }
    __syncthreads();
    uint32_t s = 0;
    for (int c = 0; c < 4; c++)
        for (int j = 0; j < 16; j++) s += lds_u32(cb[c] + j * 128);
    if (s != 4u * MB_ITERS) atomicAdd(out, 1u);  // counts are conserved: also a correctness check
}

// kinds 4/5: the instruction mix of the fold kernel's inner loop, without any memory traffic other
// than the conflict-free shared atomics: NDP separately rounded FP64 ops + 1 LOP3 + 1 ATOMS per unit
// (kind 4: NDP = 4, the (f, fdot) path; kind 5: NDP = 2, the f-only path).  Gives the attainable
// event*trial rate of this mix on the machine, i.e. an empirical ceiling for K1.
template <int NDP>
__global__ void __launch_bounds__(256) mb_mix_kernel(uint32_t *out, double a, double b) {
    extern __shared__ __align__(16) unsigned char dsm[];
    const uint32_t base = smem_u32(dsm);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t cb = base + (uint32_t)warp * 2048u + lane * 4u;  // 16 bins x 128 B, private column
    for (int j = 0; j < 16; j++) sts_u32(cb + j * 128, 0u);
    double x[MB_ILP];
#pragma unroll
    for (int i = 0; i < MB_ILP; i++) x[i] = a * (double)(threadIdx.x + 3 * i + 1);
    for (int it = 0; it < MB_ITERS; it++) {
#pragma unroll
        for (int i = 0; i < MB_ILP; i++) {
            double p = __dmul_rn(x[i], a);
            if (NDP == 4) p = __dadd_rn(p, __dmul_rn(x[i], b));
            const double y = __dadd_rd(p, HX_MAGIC);
            red_shared_inc(((uint32_t)__double2loint(y) & 0x780u) | cb);
            x[i] = p;
        }
    }
    __syncthreads();
    uint32_t s = 0;
    for (int j = 0; j < 16; j++) s += lds_u32(cb + j * 128);
    if (s != (uint32_t)MB_ITERS * MB_ILP) atomicAdd(out, 1u);
}

// kind 6: F2I.S64.F64.FLOOR issue rate (the XU pipe's fp64 -> int64 conversion)
__global__ void __launch_bounds__(256) mb_f2i_kernel(double *out, double a) {
    double x[MB_ILP];
#pragma unroll
    for (int i = 0; i < MB_ILP; i++) x[i] = a * (double)(threadIdx.x + i + 1) * 1048576.0;
    long long acc = 0;
    for (int it = 0; it < MB_ITERS; it++) {
#pragma unroll
        for (int i = 0; i < MB_ILP; i++) {
            const long long v = __double2ll_rd(x[i]);
            acc ^= v;
            x[i] = __longlong_as_double(__double_as_longlong(x[i]) + (v & 1));  // integer dependency, no FP64/XU op
        }
    }
    if (acc == 0x123456789abcdefLL) out[0] = (double)acc;
}

// kinds 7..10: the (f, fdot) mix of kind 4 with NF2I of every 8 units taking their floor through F2I.S64.F64.FLOOR
// (XU pipe) instead of the round-down add of 1.5*2^52 (FP64 pipe): does splitting the floor over two pipes pay?
template <int NF2I>
__global__ void __launch_bounds__(256) mb_mix_f2i_kernel(uint32_t *out, double a, double b) {
    extern __shared__ __align__(16) unsigned char dsm[];
    const uint32_t base = smem_u32(dsm);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t cb = base + (uint32_t)warp * 2048u + lane * 4u;
    for (int j = 0; j < 16; j++) sts_u32(cb + j * 128, 0u);
    double x[MB_ILP];
#pragma unroll
    for (int i = 0; i < MB_ILP; i++) x[i] = a * (double)(threadIdx.x + 3 * i + 1);
    for (int it = 0; it < MB_ITERS; it++) {
#pragma unroll
        for (int i = 0; i < MB_ILP; i++) {
            double p = __dmul_rn(x[i], a);
            p = __dadd_rn(p, __dmul_rn(x[i], b));
            uint32_t lo;
            if ((i * NF2I) / 8 != ((i + 1) * NF2I) / 8) {  // NF2I of the 8 chains, evenly spread
                lo = (uint32_t)__double2ll_rd(p);
            } else {
                lo = (uint32_t)__double2loint(__dadd_rd(p, HX_MAGIC));
            }
            red_shared_inc((lo & 0x780u) | cb);
            x[i] = p;
        }
    }
    __syncthreads();
    uint32_t s = 0;
    for (int j = 0; j < 16; j++) s += lds_u32(cb + j * 128);
    if (s != (uint32_t)MB_ITERS * MB_ILP) atomicAdd(out, 1u);
}

// kinds 11/12: K1's (f, fdot) loop at K1's occupancy (4 warps per CTA, ~72 KB of shared memory per CTA => 3 CTAs =
// 12 warps per SM), 2 trials per thread, events broadcast from a shared-memory tile with LDS.128:
//   kind 11 (PRE = false): tt = t*t and m2 = tt*fdh computed per event and thread (4 FP64 ops per unit; today's K1)
//   kind 12 (PRE = true):  m2 read from a second tile, precomputed once per CTA (3 FP64 ops per unit, one more
//                          shared-memory wavefront per event and warp)
template <bool PRE, int R>
__global__ void __launch_bounds__(128) mb_rowcta_kernel(uint32_t *out, double a, double b) {
    extern __shared__ __align__(16) unsigned char dsm[];
    const uint32_t base = smem_u32(dsm);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NEV = 256;
    double *tile = reinterpret_cast<double *>(dsm);        // [NEV] times
    double *mtile = tile + NEV;                             // [NEV] fdot terms
    const uint32_t hist = base + 2 * NEV * 8;
    uint32_t cb[R];
    for (int r = 0; r < R; r++) {
        cb[r] = hist + (uint32_t)(warp * R + r) * 8192u + lane * 4u;  // 64 words x 128 B per column (packed nbin = 128)
        for (int j = 0; j < 64; j++) sts_u32(cb[r] + j * 128, 0u);
    }
    for (int i = threadIdx.x; i < NEV; i += blockDim.x) {
        tile[i] = 1000.0 + 0.37 * i;
        mtile[i] = (1000.0 + 0.37 * i) * (1000.0 + 0.37 * i) * b;
    }
    __syncthreads();
    double f[R];
    for (int r = 0; r < R; r++) f[r] = a * (1.0 + 1e-6 * (threadIdx.x + 1 + 128 * r)) * 16384.0;
    const double fdh = b * 16384.0;
    constexpr int UE = R >= 4 ? 2 : 4;
    for (int it = 0; it < MB_ITERS / 32; it++) {
#pragma unroll 2
        for (int e = 0; e < NEV; e += UE) {
            double t[UE], m[UE];
#pragma unroll
            for (int u = 0; u < UE; u += 2) {
                const double2 tv = *reinterpret_cast<const double2 *>(tile + e + u);
                t[u] = tv.x, t[u + 1] = tv.y;
                if (PRE) {
                    const double2 mv = *reinterpret_cast<const double2 *>(mtile + e + u);
                    m[u] = mv.x, m[u + 1] = mv.y;
                }
            }
            if (!PRE) {
#pragma unroll
                for (int u = 0; u < UE; u++) m[u] = __dmul_rn(__dmul_rn(t[u], t[u]), fdh);
            }
#pragma unroll
            for (int u = 0; u < UE; u++)
#pragma unroll
                for (int r = 0; r < R; r++) {
                    const double y = __dadd_rd(__dadd_rn(__dmul_rn(t[u], f[r]), m[u]), HX_MAGIC);
                    const uint32_t lo = (uint32_t)__double2loint(y);
                    red_shared_add_u32((lo & 0x1f80u) | cb[r], (lo & 0x2000u) | 1u);
                }
        }
    }
    __syncthreads();
    uint32_t s = 0;
    for (int r = 0; r < R; r++)
        for (int j = 0; j < 64; j++) {
            const uint32_t v = lds_u32(cb[r] + j * 128);
            s += (v & 0x1fffu);
        }
    if (s == 0xffffffffu) atomicAdd(out, 1u);
}

// kinds 13/14: where should the broadcast event stream come from?  K1's f-only loop at K1's occupancy for nbin = 128
// (R = 2, 4 warps per CTA, ~72 KB of shared memory per CTA), events read with broadcast LDS.128 from a shared-memory
// tile (kind 13, today's K1) or with indexed LDC.64 from a __constant__ array (kind 14): constant loads do not go through
// the shared-memory data pipe that bounds the f-only kernels.
__constant__ double mb_const_events[4096];

template <bool FROM_CONST>
__global__ void __launch_bounds__(128) mb_evsrc_kernel(uint32_t *out, double a) {
    extern __shared__ __align__(16) unsigned char dsm[];
    const uint32_t base = smem_u32(dsm);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NEV = 512;
    double *tile = reinterpret_cast<double *>(dsm);
    const uint32_t hist = base + NEV * 8;
    uint32_t cb[2];
    for (int r = 0; r < 2; r++) {
        cb[r] = hist + (uint32_t)(warp * 2 + r) * 8192u + lane * 4u;
        for (int j = 0; j < 64; j++) sts_u32(cb[r] + j * 128, 0u);
    }
    for (int i = threadIdx.x; i < NEV; i += blockDim.x) tile[i] = 1000.0 + 0.37 * i;
    __syncthreads();
    const double f[2] = {a * (1.0 + 1e-6 * (threadIdx.x + 1)) * 16384.0, a * (1.0 + 1e-6 * (threadIdx.x + 129)) * 16384.0};
    for (int it = 0; it < MB_ITERS / 32; it++) {
#pragma unroll 2
        for (int e = 0; e < NEV; e += 4) {
            double t[4];
            if (FROM_CONST) {
                const int o = (e + it * 4) & 4092;
#pragma unroll
                for (int u = 0; u < 4; u++) t[u] = mb_const_events[o + u];
            } else {
                const double2 t01 = *reinterpret_cast<const double2 *>(tile + e), t23 = *reinterpret_cast<const double2 *>(tile + e + 2);
                t[0] = t01.x, t[1] = t01.y, t[2] = t23.x, t[3] = t23.y;
            }
#pragma unroll
            for (int u = 0; u < 4; u++)
#pragma unroll
                for (int r = 0; r < 2; r++) {
                    const uint32_t lo = (uint32_t)__double2loint(__dadd_rd(__dmul_rn(t[u], f[r]), HX_MAGIC));
                    red_shared_add_u32((lo & 0x1f80u) | cb[r], (lo & 0x2000u) | 1u);
                }
        }
    }
    __syncthreads();
    uint32_t s = 0;
    for (int r = 0; r < 2; r++)
        for (int j = 0; j < 64; j++) s += lds_u32(cb[r] + j * 128) & 0x1fffu;
    if (s == 0xffffffffu) atomicAdd(out, 1u);
}

}  // namespace hx

extern "C" int hx_microbench(int kind, int reps, double *ops_per_s, int device) {
    using namespace hx;
    if (!ops_per_s || kind < 0 || kind > 16 || reps < 1) {
        set_error("hx_microbench: bad arguments");
        return HX_ERR_ARG;
    }
    hx::DeviceGuard dev_guard;  // declared before the device buffers: they are freed before the device is restored
    HX_CUDA_TRY(dev_guard.enter(device));
    DeviceInfo di;
    int rc = get_device_info(device, &di);
    if (rc) return rc;
    DevBuf out;
    HX_CUDA_TRY(out.alloc(64));
    HX_CUDA_TRY(cudaMemset(out.p, 0, 64));
    cudaEvent_t e0, e1;
    HX_CUDA_TRY(cudaEventCreate(&e0));
    HX_CUDA_TRY(cudaEventCreate(&e1));
    const int threads = 256;
    const int grid = di.sm_count * 8 * 4;  // 4 waves of 8 CTAs/SM
    const size_t smem = (kind >= 4) ? (size_t)(threads / 32) * 2048 : (kind >= 2) ? (size_t)(threads / 32) * 4 * 2048 : 0;
    if (kind == 2) HX_CUDA_TRY(cudaFuncSetAttribute(mb_smem_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (kind == 3) HX_CUDA_TRY(cudaFuncSetAttribute(mb_smem_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (kind == 11) HX_CUDA_TRY(cudaFuncSetAttribute(mb_rowcta_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
    if (kind == 12) HX_CUDA_TRY(cudaFuncSetAttribute(mb_rowcta_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
    if (kind == 15) HX_CUDA_TRY(cudaFuncSetAttribute(mb_rowcta_kernel<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
    if (kind == 16) HX_CUDA_TRY(cudaFuncSetAttribute(mb_rowcta_kernel<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
    if (kind == 13) HX_CUDA_TRY(cudaFuncSetAttribute(mb_evsrc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
    if (kind == 14) {
        HX_CUDA_TRY(cudaFuncSetAttribute(mb_evsrc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
        std::vector<double> ev(4096);
        for (int i = 0; i < 4096; i++) ev[i] = 1000.0 + 0.37 * i;
        HX_CUDA_TRY(cudaMemcpyToSymbol(mb_const_events, ev.data(), ev.size() * sizeof(double)));
    }
    double best_ms = 1e30;
    for (int r = 0; r < reps + 1; r++) {  // first rep is a warm-up
        HX_CUDA_TRY(cudaEventRecord(e0));
        switch (kind) {
            case 0: mb_fp64_kernel<0><<<grid, threads>>>(out.as<double>() + 1, 1.0000001, 1e-9); break;
            case 1: mb_fp64_kernel<1><<<grid, threads>>>(out.as<double>() + 1, 1.0000001, 1e-9); break;
            case 2: mb_smem_kernel<2><<<grid, threads, smem>>>(out.as<uint32_t>(), (uint32_t)r); break;
            case 3: mb_smem_kernel<3><<<grid, threads, smem>>>(out.as<uint32_t>(), (uint32_t)r); break;
            case 4: mb_mix_kernel<4><<<grid, threads, smem>>>(out.as<uint32_t>(), 1.0000001, 1e-9); break;
            case 5: mb_mix_kernel<2><<<grid, threads, smem>>>(out.as<uint32_t>(), 1.0000001, 1e-9); break;
            case 6: mb_f2i_kernel<<<grid, threads>>>(out.as<double>() + 1, 1.0000001); break;
            case 7: mb_mix_f2i_kernel<2><<<grid, threads, smem>>>(out.as<uint32_t>(), 1.0000001, 1e-9); break;
            case 8: mb_mix_f2i_kernel<4><<<grid, threads, smem>>>(out.as<uint32_t>(), 1.0000001, 1e-9); break;
            case 9: mb_mix_f2i_kernel<6><<<grid, threads, smem>>>(out.as<uint32_t>(), 1.0000001, 1e-9); break;
            case 10: mb_mix_f2i_kernel<8><<<grid, threads, smem>>>(out.as<uint32_t>(), 1.0000001, 1e-9); break;
            case 11: mb_rowcta_kernel<false, 2><<<di.sm_count * 3 * 2, 128, 72 * 1024>>>(out.as<uint32_t>(), 9.9, 1e-10); break;
            case 12: mb_rowcta_kernel<true, 2><<<di.sm_count * 3 * 2, 128, 72 * 1024>>>(out.as<uint32_t>(), 9.9, 1e-10); break;
            case 13: mb_evsrc_kernel<false><<<di.sm_count * 3 * 2, 128, 72 * 1024>>>(out.as<uint32_t>(), 9.9); break;
            case 14: mb_evsrc_kernel<true><<<di.sm_count * 3 * 2, 128, 72 * 1024>>>(out.as<uint32_t>(), 9.9); break;
            case 15: mb_rowcta_kernel<false, 4><<<di.sm_count * 3 * 2, 64, 72 * 1024>>>(out.as<uint32_t>(), 9.9, 1e-10); break;
            default: mb_rowcta_kernel<true, 4><<<di.sm_count * 3 * 2, 64, 72 * 1024>>>(out.as<uint32_t>(), 9.9, 1e-10); break;
        }
        HX_CHECK_LAUNCH("microbench kernel");
        HX_CUDA_TRY(cudaEventRecord(e1));
        HX_CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0;
        HX_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        if (r > 0) best_ms = std::min(best_ms, (double)ms);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    uint32_t bad = 0;
    HX_CUDA_TRY(cudaMemcpy(&bad, out.p, 4, cudaMemcpyDeviceToHost));
    if (kind >= 2 && kind <= 10 && bad) {
        set_error("hx_microbench: shared-memory counters not conserved (%u warps)", bad);
        return HX_ERR_INTERNAL;
    }
    double per_thread = (kind == 0) ? (double)MB_ITERS * MB_ILP
                      : (kind == 1) ? (double)MB_ITERS * MB_ILP * 2
                      : (kind >= 4) ? (double)MB_ITERS * MB_ILP  // units (event*trial equivalents; kind 6: F2I)
                                    : (double)MB_ITERS * 4;
    if (kind >= 15) {  // 64 threads x 4 trials
        *ops_per_s = (double)(MB_ITERS / 32) * 256 * 4 * 64 * (double)(di.sm_count * 6) / (best_ms * 1e-3);
        return HX_OK;
    }
    if (kind >= 13) {
        *ops_per_s = (double)(MB_ITERS / 32) * 512 * 2 * 128 * (double)(di.sm_count * 6) / (best_ms * 1e-3);
        return HX_OK;
    }
    if (kind >= 11) {
        *ops_per_s = (double)(MB_ITERS / 32) * 256 * 2 * 128 * (double)(di.sm_count * 6) / (best_ms * 1e-3);
        return HX_OK;
    }
    *ops_per_s = per_thread * threads * (double)grid / (best_ms * 1e-3);
    return HX_OK;
}
