// common.cuh -- shared device/host helpers for libdssmd_b200.so (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <utility>

#include "../../include/dssmd_b200.h"

namespace dssmd {

// ---- thread-local error string (dssmd_last_error) ---------------------------
void set_error(const char *fmt, ...);
int fail(int code, const char *fmt, ...);
int check_launch(const char *what);  // cudaGetLastError -> DSSMD_ERR_CUDA

#define DSSMD_REQUIRE(cond, ...)                                             \
    do {                                                                     \
        if (!(cond)) return ::dssmd::fail(DSSMD_ERR_INVALID_ARGUMENT, __VA_ARGS__); \
    } while (0)

// ---- device properties, cached per device (read-only after first use) ----------
int sm_count();

// ---- per (kernel, device) launch set-up, asked once and then looked up (abi.cu) ------------------------------------------
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) and cudaOccupancyMaxActiveBlocksPerMultiprocessor are properties of a
// kernel on a device; they used to be re-asked on every launch (a few microseconds each, more than some of the kernels).
cudaError_t kernel_smem_optin(const void *func, size_t smem);           // no-op once the kernel has been granted >= smem
int kernel_occupancy(const void *func, int threads, size_t smem);       // resident blocks per SM (>= 1)
template <typename K> inline cudaError_t kernel_smem_optin(K kern, size_t smem) { return kernel_smem_optin(reinterpret_cast<const void *>(kern), smem); }
template <typename K> inline int kernel_occupancy(K kern, int threads, size_t smem) { return kernel_occupancy(reinterpret_cast<const void *>(kern), threads, smem); }
// The DSSMD_* tuning / debug variables are honoured only in processes started with
// DSSMD_TUNING=1 (read once): a production process pays one cached flag test per knob instead of ~30 getenv() scans of the
// environment per call.  The test-suite and the sweep scripts set it.
bool tuning_enabled();

// ---- element conversion --------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// 16-byte async global->shared copy; src_bytes == 0 writes zeros (src must still
// be a valid address).
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void *src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
// 4-byte variant for sources that are only element-aligned (skewed rows)
__device__ __forceinline__ void cp_async4(uint32_t dst_smem, const void *src, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// streaming 128-bit global store (outputs are written once, never re-read here)
__device__ __forceinline__ void st_global_cs(float4 *p, float4 v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// ---- programmatic dependent launch (PDL) ----------------------------------------------------
// The hot-path kernels are short (10-40 us) and follow each other on one stream, so the drain of one kernel and the
// ramp of the next (block scheduling, shared-memory carve-out, barrier init, tensor-map fetch) are a measurable
// share of the step.  Launched with programmatic stream serialization, a kernel's blocks may become resident while
// the previous kernel on the stream is still draining; pdl_wait() (griddepcontrol.wait) then blocks until that
// kernel has completed and its writes are visible, so stream order is preserved as long as no global memory is
// touched before it.  pdl_launch_dependents() lets the NEXT kernel's blocks start their own ramp.
// Both are no-ops for a kernel launched without the attribute.  DSSMD_PDL=0 disables the attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Loads in a kernel that may FOLLOW one of our own kernels on the stream: launched with programmatic stream serialization it
// can become resident while that predecessor is still writing the tensors it reads after griddepcontrol.wait, so they are
// not read-only for the kernel's lifetime and the non-coherent path (ld.global.nc -- __ldg, and what nvcc picks by itself
// for const __restrict__ pointers) is off limits: measured, a consumer on some boxes read stale lines (the attention
// backward reading `delta` from the launch before it).  ld.global.cg caches in L2 only and is always coherent.
// ld.global.ca: coherent as well (ordered by griddepcontrol.wait under the memory model) and cached in L1, for the kernels
// that re-read rows (the warp kernels)
__device__ __forceinline__ float ld_ca(const float *p) { return __ldca(p); }
__device__ __forceinline__ float4 ld_ca(const float4 *p) { return __ldca(p); }
__device__ __forceinline__ float ld_cg(const float *p) { return __ldcg(p); }
__device__ __forceinline__ double ld_cg(const double *p) { return __ldcg(p); }
__device__ __forceinline__ float2 ld_cg(const float2 *p) { return __ldcg(p); }
__device__ __forceinline__ float4 ld_cg(const float4 *p) { return __ldcg(p); }

// L2 prefetch of the lines a block reads FIRST, issued BEFORE griddepcontrol.wait: while the predecessor on the stream is
// still draining, the blocks that are already resident pull their first tiles from HBM into the L2, so the loads after the
// wait start from an L2 hit instead of a cold DRAM access.  A prefetch returns nothing to the thread and the L2 is the
// point of coherence of the device: if the predecessor still writes one of those lines the line is simply updated in
// place, so unlike a load this is legal before the wait.  DSSMD_PDL_PREFETCH=0 (tuning) disables it, =n sets the amount.
// Measured at the 576x960 model shapes (profiles/r04c_pdl_prefetch.log): ONE item per block is the optimum -- corr_fwd 12.35 ->
// 11.7 us, corr_bwd 20.3 -> 18.5 us, warp_fwd 18.6 -> 18.0 us, warp_bwd 29.25 -> 28.9 us; two items give half of that and
// four or more are slower than none (the extra traffic slows the predecessor's last blocks, which everybody waits for).
// Touching one line per later item (address translation only) changed nothing.
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

static inline int pdl_prefetch(int dflt) {   // how much a kernel prefetches (its own unit: groups, items, rows); 0 = nothing
    if (!tuning_enabled()) return dflt;
    const char *v = std::getenv("DSSMD_PDL_PREFETCH");
    return (v && *v) ? std::atoi(v) : dflt;
}

static inline bool pdl_enabled() {
    if (!tuning_enabled()) return true;
    const char *v = std::getenv("DSSMD_PDL");
    return !(v && v[0] == '0');
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(std::forward<Args>(args))...);
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

}  // namespace dssmd
