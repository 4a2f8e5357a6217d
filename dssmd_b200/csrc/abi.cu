// abi.cu -- error reporting, version and device-property cache of libdssmd_b200.so.
#include "common.cuh"

#include <atomic>
#include <cstring>
#include <mutex>
#include <vector>

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

namespace {
struct KernelEntry {
    const void *func;
    int dev;
    size_t smem_granted;
    int occ_threads;
    size_t occ_smem;
    int occ;
};
std::mutex g_kmutex;
std::vector<KernelEntry> g_kernels;

KernelEntry &kernel_entry(const void *func) {   // g_kmutex held
    int dev = 0;
    cudaGetDevice(&dev);
    for (auto &e : g_kernels)
        if (e.func == func && e.dev == dev) return e;
    g_kernels.push_back(KernelEntry{func, dev, 48 * 1024, 0, 0, 0});
    return g_kernels.back();
}

}  // namespace

cudaError_t kernel_smem_optin(const void *func, size_t smem) {
    std::lock_guard<std::mutex> lock(g_kmutex);
    KernelEntry &e = kernel_entry(func);
    if (smem <= e.smem_granted) return cudaSuccess;
    const cudaError_t rc = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (rc == cudaSuccess) e.smem_granted = smem;
    return rc;
}

int kernel_occupancy(const void *func, int threads, size_t smem) {
    std::lock_guard<std::mutex> lock(g_kmutex);
    KernelEntry &e = kernel_entry(func);
    if (e.occ > 0 && e.occ_threads == threads && e.occ_smem == smem) return e.occ;
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, func, threads, smem) != cudaSuccess || occ < 1) occ = 1;
    e.occ = occ; e.occ_threads = threads; e.occ_smem = smem;
    return occ;
}

bool tuning_enabled() {
    static const bool on = [] {
        const char *v = std::getenv("DSSMD_TUNING");
        return v && *v && *v != '0';
    }();
    return on;
}

}  // namespace dssmd

extern "C" int dssmd_version(void) { return DSSMD_ABI_VERSION; }
extern "C" const char *dssmd_last_error(void) { return dssmd::g_err; }
