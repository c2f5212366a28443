// Top-k candidates of a statistic array (value descending, ties by ascending flat index).
// Replaces the argpartition branch of hendrics/base.py:1140-1146 (find_peaks_in_image) for
// device-resident statistics; k is small (5 in _analyze_qffa_results, 10 in EFPeriodogram.find_peaks),
// so k streaming arg-max passes (each one HBM read of the array) are cheaper than a sort.
#include <algorithm>
#include <cmath>
#include <vector>

#include "hx_common.cuh"

namespace hx {

struct Cand {
    double val;
    long long idx;
};

// a strictly before b in the output order
__device__ __forceinline__ bool cand_before(double av, long long ai, double bv, long long bi) {
    return av > bv || (av == bv && ai < bi);
}

__device__ __forceinline__ void cand_reduce(double &v, long long &i) {
    for (int o = 16; o; o >>= 1) {
        double ov = __shfl_xor_sync(0xffffffffu, v, o);
        long long oi = __shfl_xor_sync(0xffffffffu, i, o);
        if (oi >= 0 && (i < 0 || cand_before(ov, oi, v, i))) v = ov, i = oi;
    }
}

// best element strictly after `prev` (prev->idx < 0: no constraint); NaN is never selected
__global__ void __launch_bounds__(256) topk_pass_kernel(const double *__restrict__ x, long long n, const Cand *prev,
                                                        Cand *block_out) {
    const double pv = prev->val;
    const long long pi = prev->idx;
    double bv = 0.0;
    long long bi = -1;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double v = x[i];
        if (v != v) continue;
        if (pi >= 0 && !cand_before(pv, pi, v, i)) continue;
        if (bi < 0 || cand_before(v, i, bv, bi)) bv = v, bi = i;
    }
    cand_reduce(bv, bi);
    __shared__ Cand sh[8];
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = Cand{bv, bi};
    __syncthreads();
    if (threadIdx.x < 32) {
        Cand c = threadIdx.x < 8 ? sh[threadIdx.x] : Cand{0.0, -1};
        cand_reduce(c.val, c.idx);
        if (threadIdx.x == 0) block_out[blockIdx.x] = c;
    }
}

__global__ void __launch_bounds__(1024) topk_final_kernel(const Cand *block_in, int nblocks, Cand *out_slot, Cand *prev) {
    double bv = 0.0;
    long long bi = -1;
    for (int i = threadIdx.x; i < nblocks; i += blockDim.x) {
        Cand c = block_in[i];
        if (c.idx >= 0 && (bi < 0 || cand_before(c.val, c.idx, bv, bi))) bv = c.val, bi = c.idx;
    }
    cand_reduce(bv, bi);
    __shared__ Cand sh[32];
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = Cand{bv, bi};
    __syncthreads();
    if (threadIdx.x < 32) {
        Cand c = sh[threadIdx.x];
        cand_reduce(c.val, c.idx);
        if (threadIdx.x == 0) {
            *out_slot = c;
            *prev = c;
        }
    }
}

int topk_dev_impl(const double *d_stats, int64_t n, int k, int64_t *h_idx, double *h_val, int64_t *n_out,
                  cudaStream_t st) {
    std::lock_guard<std::recursive_mutex> api_lock(api_mutex());
    if (!d_stats || !h_idx || !h_val || !n_out || n < 0 || k < 0 || k > 4096) {
        set_error("hx_topk: bad arguments");
        return HX_ERR_ARG;
    }
    const int kk = (int)std::min<int64_t>(k, n);
    *n_out = 0;
    if (kk == 0) return HX_OK;
    int device;
    HX_CUDA_TRY(cudaGetDevice(&device));
    DeviceInfo di;
    int rc = get_device_info(device, &di);
    if (rc) return rc;
    const int nblocks = (int)std::min<long long>((n + 255) / 256, (long long)di.sm_count * 8);
    void *buf = nullptr;
    HX_CUDA_TRY(cudaMallocAsync(&buf, sizeof(Cand) * (size_t)(nblocks + kk + 1), st));
    Cand *d_block = reinterpret_cast<Cand *>(buf), *d_out = d_block + nblocks, *d_prev = d_out + kk;
    Cand none{0.0, -1};
    cudaError_t e = cudaMemcpyAsync(d_prev, &none, sizeof(Cand), cudaMemcpyHostToDevice, st);
    for (int i = 0; i < kk && e == cudaSuccess; i++) {
        topk_pass_kernel<<<nblocks, 256, 0, st>>>(d_stats, n, d_prev, d_block);
        topk_final_kernel<<<1, 1024, 0, st>>>(d_block, nblocks, d_out + i, d_prev);
        count_launch(2);
        e = cudaGetLastError();
    }
    std::vector<Cand> host(kk);
    if (e == cudaSuccess) e = cudaMemcpyAsync(host.data(), d_out, sizeof(Cand) * kk, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFreeAsync(buf, st);
    if (e != cudaSuccess) {
        set_error("hx_topk failed: %s", cudaGetErrorString(e));
        return HX_ERR_CUDA;
    }
    int64_t m = 0;
    for (int i = 0; i < kk; i++) {
        if (host[i].idx < 0) break;  // fewer than k non-NaN entries
        h_idx[m] = host[i].idx;
        h_val[m] = host[i].val;
        m++;
    }
    *n_out = m;
    return HX_OK;
}

}  // namespace hx

extern "C" {

int hx_topk_dev(const double *d_stats, int64_t n, int k, int64_t *h_idx, double *h_val, int64_t *n_out, void *stream) {
    return hx::topk_dev_impl(d_stats, n, k, h_idx, h_val, n_out, reinterpret_cast<cudaStream_t>(stream));
}

int hx_topk(const double *stats, int64_t n, int k, int64_t *idx, double *val, int64_t *n_out, int device) {
    if (!stats && n > 0) {
        hx::set_error("hx_topk: null pointer");
        return HX_ERR_ARG;
    }
    hx::DeviceGuard dev_guard;  // declared before the device buffers: they are freed before the device is restored
    HX_CUDA_TRY(dev_guard.enter(device));
    hx::DevBuf d;
    HX_CUDA_TRY(d.alloc((size_t)std::max<int64_t>(n, 0) * 8));
    if (n > 0) HX_CUDA_TRY(cudaMemcpy(d.p, stats, (size_t)n * 8, cudaMemcpyHostToDevice));
    return hx::topk_dev_impl(d.as<double>(), n, k, idx, val, n_out, 0);
}

}  // extern "C"
