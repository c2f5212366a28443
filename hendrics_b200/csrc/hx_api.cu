// Library-level plumbing of libhxsearch.so: errors, device info, launch counter.
#include <atomic>
#include <cstdarg>
#include <cstring>
#include <mutex>

#include "hx_common.cuh"

namespace hx {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

std::recursive_mutex &api_mutex() {
    static std::recursive_mutex m;
    return m;
}

// Shared-window address at which dynamic shared memory starts for a kernel without static
// shared memory (the fold kernel aligns its histogram blocks on absolute shared addresses).
__global__ void probe_smem_base_kernel(uint32_t *out) {
    extern __shared__ __align__(16) unsigned char dsm[];
    if (threadIdx.x == 0) *out = smem_u32(dsm);
}

static std::mutex g_info_mu;
static DeviceInfo g_info[64];
static bool g_info_ok[64] = {false};

int get_device_info(int device, DeviceInfo *out) {
    if (device < 0 || device >= 64) {
        set_error("device index %d out of range", device);
        return HX_ERR_ARG;
    }
    std::lock_guard<std::mutex> lk(g_info_mu);
    if (!g_info_ok[device]) {
        DeviceInfo di{};
        int major = 0, minor = 0;
        HX_CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
        HX_CUDA_TRY(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
        if (major != 10) {  // the library holds one sm_100a cubin: arch-specific code runs on compute capability 10.x only
            set_error("device %d has compute capability %d.%d; libhxsearch is built for sm_100a (B200) only", device, major,
                      minor);
            return HX_ERR_CUDA;
        }
        HX_CUDA_TRY(cudaDeviceGetAttribute(&di.sm_count, cudaDevAttrMultiProcessorCount, device));
        HX_CUDA_TRY(cudaDeviceGetAttribute(&di.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
        HX_CUDA_TRY(cudaDeviceGetAttribute(&di.smem_per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, device));
        HX_CUDA_TRY(cudaDeviceGetAttribute(&di.reserved_smem, cudaDevAttrReservedSharedMemoryPerBlock, device));
        DeviceGuard guard;  // back on the caller's device on every exit path below
        HX_CUDA_TRY(guard.enter(device));
        {
            // keep stream-ordered temporaries (cudaMallocAsync) cached across calls: by default the
            // pool hands memory back to the OS at every synchronisation point
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
                uint64_t thr = UINT64_MAX;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
            }
        }
        uint32_t *d = nullptr;
        HX_CUDA_TRY(cudaMalloc(&d, sizeof(uint32_t)));
        probe_smem_base_kernel<<<1, 32, 64>>>(d);
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpy(&di.dyn_smem_base, d, sizeof(uint32_t), cudaMemcpyDeviceToHost);
        cudaFree(d);
        if (e != cudaSuccess) {
            if (e == cudaErrorNoKernelImageForDevice)
                set_error("device %d (compute capability %d.%d) cannot run this library's sm_100a kernels", device, major, minor);
            else
                set_error("probe_smem_base_kernel failed: %s", cudaGetErrorString(e));
            return HX_ERR_CUDA;
        }
        count_launch();
        g_info[device] = di;
        g_info_ok[device] = true;
    }
    *out = g_info[device];
    return HX_OK;
}

}  // namespace hx

extern "C" {

int hx_version(void) { return 1000; }

const char *hx_last_error(void) { return hx::g_err; }

int hx_device_count(int *count) {
    if (!count) {
        hx::set_error("hx_device_count: null pointer");
        return HX_ERR_ARG;
    }
    *count = 0;
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess || *count == 0) {
        *count = 0;
        hx::set_error("no CUDA device: %s", cudaGetErrorString(e));
        return HX_ERR_CUDA;
    }
    return HX_OK;
}

int64_t hx_launch_count(void) { return hx::g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
