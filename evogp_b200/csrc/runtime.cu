// runtime.cu — library-wide state of libevogp_b200.so: last-error text, launch
// counter, device sanity check.  No CPU fallback exists anywhere in this library:
// without a usable sm_100 device every entry point returns EVOGP_ERR_CUDA.
#include <atomic>
#include <cstdarg>
#include <cstring>
#include "common.cuh"

namespace evogp {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

int check_launch(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
        return EVOGP_ERR_CUDA;
    }
    return EVOGP_OK;
}

int ensure_device_ok() {
    static thread_local int checked_dev = -1;
    int dev = -1;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        set_error("no CUDA device: %s (this library has no CPU path)", cudaGetErrorString(e));
        return EVOGP_ERR_CUDA;
    }
    if (dev == checked_dev) return EVOGP_OK;
    int major = 0;
    e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e != cudaSuccess || major != 10) {
        set_error("device %d has compute capability major %d; this library ships sm_100a code only", dev, major);
        return EVOGP_ERR_CUDA;
    }
    checked_dev = dev;
    return EVOGP_OK;
}

}  // namespace evogp

extern "C" int evogp_version(void) { return 100; }
extern "C" const char *evogp_last_error(void) { return evogp::g_err; }
extern "C" unsigned long long evogp_launch_count(void) { return evogp::g_launches.load(); }
