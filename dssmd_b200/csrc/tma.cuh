// tma.cuh -- mbarrier + TMA (cp.async.bulk.tensor) helpers shared by the kernels that stage
// their tiles with the tensor memory accelerator (sm_100a).
#pragma once

#include <cuda.h>

#include "common.cuh"

namespace dssmd {

__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// bounded wait: a protocol bug must trap, not hang the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (int spin = 0; spin < (1 << 26); ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
    }
    __trap();
}
// make mbarrier.init / generic shared-memory writes visible to the async proxy (TMA, tensor core)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 4-D tiled TMA load: box at (c0, c1, c2, c3) -> shared memory, completes `bytes` on the mbarrier
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap *map, int c0, int c1, int c2, int c3, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
}

// the same box towards the L2 only (no shared-memory destination, no barrier): see prefetch_l2 in common.cuh
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap *map, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// fetch the descriptor itself ahead of the first use (kernel parameter: __grid_constant__)
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// ---- host: tensor maps (driver entry point resolved through the runtime, no -lcuda) ------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = [] {
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            sym = nullptr;
        return reinterpret_cast<EncodeTiledFn>(sym);
    }();
    return fn;
}

// tensor map of `dtype` elements; dims/box fastest first; strides in bytes for dims 1..rank-1; out-of-bounds reads give 0
inline bool make_tmap(CUtensorMap *m, CUtensorMapDataType dtype, const void *base, int rank, const cuuint64_t *dims,
                      const cuuint64_t *strides, const cuuint32_t *box, CUtensorMapSwizzle swizzle) {
    EncodeTiledFn fn = encode_tiled();
    if (!fn) return false;
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    auto encode = [&] {
        return fn(m, dtype, (cuuint32_t)rank, const_cast<void *>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    };
    if (encode()) return true;
    // The encoder is a driver entry point and wants the calling thread to have a current context.  A thread that has
    // only set its device (an autograd worker on its first backward, a fresh nn.DataParallel replica thread --
    // trainfiles/dssmd_traner_new.py:131-135) may not have one bound yet: measured, the first call from such a thread
    // fails and every later one succeeds.  cudaFree(0) makes the runtime bind the primary context; then try again.
    cudaFree(nullptr);
    return encode();
}
inline bool make_tmap_f32(CUtensorMap *m, const void *base, int rank, const cuuint64_t *dims, const cuuint64_t *strides,
                          const cuuint32_t *box, CUtensorMapSwizzle swizzle) {
    return make_tmap(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, base, rank, dims, strides, box, swizzle);
}

}  // namespace dssmd
