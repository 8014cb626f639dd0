// Error plumbing + device queries behind the C ABI.
#include "common.cuh"
#include <string.h>
#include <atomic>

namespace riders {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void count_launches(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) return RIDERS_OK;
    set_error("%s: %s", what, cudaGetErrorString(e));
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ||
        e == cudaErrorNoKernelImageForDevice || e == cudaErrorInvalidDeviceFunction)
        return RIDERS_ENODEVICE;
    return RIDERS_ECUDA;
}

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (!cached[dev]) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

}  // namespace riders

extern "C" const char* riders_last_error_string(void) { return riders::g_err; }
extern "C" int riders_abi_version(void) { return 2; }
extern "C" unsigned long long riders_launch_count(void) { return riders::g_launches.load(std::memory_order_relaxed); }
extern "C" int riders_debug_set(const char* key, int value) {
    if (key && !strcmp(key, "gather_smem_kb")) { riders::set_gather_smem_budget(value); return RIDERS_OK; }
    if (key && !strcmp(key, "tail_tc_dbg")) { riders::set_tail_tc_dbg(value); return RIDERS_OK; }
    if (key && !strcmp(key, "lsq_fused")) { riders::set_lsq_mode(value); return RIDERS_OK; }
    if (key && !strcmp(key, "reduce_tile_w")) { riders::set_reduce_tile_w(value); return RIDERS_OK; }
    if (key && !strcmp(key, "reduce_rows")) { riders::set_reduce_rows(value); return RIDERS_OK; }
    riders::set_error("debug_set: unknown key '%s'", key ? key : "(null)");
    return RIDERS_EINVAL;
}
extern "C" int riders_device_arch(void) {
    int dev = 0, major = 0, minor = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { riders::set_error("no CUDA device"); cudaGetLastError(); return RIDERS_ENODEVICE; }
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    return major * 10 + minor;
}
