// Error plumbing and device checks shared by every C-ABI entry point.
#include <stdarg.h>
#include <string.h>

#include <atomic>
#include <map>
#include <mutex>
#include <utility>

#include "common.cuh"

namespace dfb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

long long launches(bool reset);

int check_cuda(cudaError_t e, const char* what, const char* file, int line) {
    if (e == cudaSuccess) return 0;
    const char* base = strrchr(file, '/');
    set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), base ? base + 1 : file, line, what);
    return (int)e;
}

static std::atomic<long long> g_launches{0};
void note_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
long long launches(bool reset) { return reset ? g_launches.exchange(0) : g_launches.load(); }

int sm_count() {
    // cached per device: a process may drive several GPUs (one estimator per --gpu_idx)
    static std::atomic<int> cached[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    int n = cached[dev].load(std::memory_order_relaxed);
    if (n == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
        cached[dev].store(n, std::memory_order_relaxed);
    }
    return n;
}

int ensure_dyn_smem(const void* func, int bytes) {
    // cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device property of the function
    static std::mutex mu;
    static std::map<std::pair<const void*, int>, int> done;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return check_cuda(e, "cudaGetDevice", __FILE__, __LINE__);
    std::lock_guard<std::mutex> lk(mu);
    auto key = std::make_pair(func, dev);
    auto it = done.find(key);
    if (it != done.end() && it->second >= bytes) return 0;
    e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)", __FILE__, __LINE__);
    done[key] = bytes;
    return 0;
}

}  // namespace dfb

extern "C" int dfb_abi_version(void) { return DFB_ABI_VERSION; }

extern "C" const char* dfb_last_error_string(void) { return dfb::g_err; }

extern "C" int dfb_device_check(void) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        cudaGetLastError();
        dfb::set_error("no CUDA device available (%s); deepfit_b200 has no CPU fallback", cudaGetErrorString(e));
        return DFB_E_NODEVICE;
    }
    int major = 0;
    e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e != cudaSuccess || major != 10) {
        cudaGetLastError();
        dfb::set_error("device %d has compute capability major %d; this library is built for sm_100a only", dev, major);
        return DFB_E_NODEVICE;
    }
    return DFB_OK;
}

extern "C" int64_t dfb_launch_count(int32_t reset) { return dfb::launches(reset != 0); }
