// abi.cu -- error reporting, version and device-property cache of libdssmd_b200.so.
#include "common.cuh"

#include <atomic>

namespace dssmd {

namespace {
thread_local char g_err[512] = "";
std::atomic<int> g_sm_count[64];
}  // namespace

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int check_launch(const char *what) {
    const cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) return DSSMD_OK;
    return fail(DSSMD_ERR_CUDA, "%s: launch failed: %s", what, cudaGetErrorString(e));
}

int sm_count() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    int n = g_sm_count[dev].load(std::memory_order_relaxed);
    if (n == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        g_sm_count[dev].store(n, std::memory_order_relaxed);
    }
    return n;
}

}  // namespace dssmd

extern "C" int dssmd_version(void) { return DSSMD_ABI_VERSION; }
extern "C" const char *dssmd_last_error(void) { return dssmd::g_err; }
