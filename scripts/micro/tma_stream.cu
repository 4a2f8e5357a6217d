// Load-skeleton microbenchmark for the correlation kernels' staging pattern on sm_100a: how fast can one SM-resident ring of TMA
// boxes pull the feature rows of a [B, C, H, W] fp32 tensor pair (the model's 1/8-resolution shape B4 C128 72x120) when nothing
// is computed?  Compares box shapes (channels per box, rows per box, halo / out-of-bounds columns), ring depths, blocks per SM,
// and a plain coalesced LDG.128 sweep of the same bytes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_stream tma_stream.cu -lcuda
//   ./tma_stream            (prints one line per configuration: us per pass, GB/s on the bytes actually requested)
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma4(uint32_t dst, const CUtensorMap *m, int c0, int c1, int c2, int c3, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(dst), "l"((uint64_t)m), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
}

struct P {
    int B, C, H, W;
    int BW, X0;        // box width (elements) and start column
    int CB, NY;        // channels and rows per box
    int S;             // ring depth
    int nitems, ipb;   // items = B * (H / NY) * (C / CB); items per block (contiguous)
    int two;           // 1: also the second tensor (box start 0)
    uint32_t box_bytes;
};

constexpr int kMaxS = 16;

__global__ void __launch_bounds__(256) ring_kernel(const P p, const __grid_constant__ CUtensorMap mA, const __grid_constant__ CUtensorMap mB, float *sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long full[kMaxS], empty[kMaxS];
    const int tid = threadIdx.x, lane = tid & 31, nw = blockDim.x >> 5;
    const int i0 = blockIdx.x * p.ipb, i1 = min(i0 + p.ipb, p.nitems);
    if (i0 >= i1) return;
    if (tid == 0) {
        for (int s = 0; s < p.S; ++s) { mbar_init(smem_u32(&full[s]), 1); mbar_init(smem_u32(&empty[s]), nw); }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    const uint32_t stage_bytes = ((p.box_bytes + 127u) & ~127u) * (p.two ? 2u : 1u);
    const int ncb = p.C / p.CB, nyg = p.H / p.NY;
    auto issue = [&](int it, int s) {
        const int cb = it % ncb, r = it / ncb;
        const int yg = r % nyg, b = r / nyg;
        const uint32_t bar = smem_u32(&full[s]);
        const uint32_t dst = smem_u32(smem) + (uint32_t)s * stage_bytes;
        mbar_expect(bar, p.box_bytes * (p.two ? 2u : 1u));
        tma4(dst, &mA, p.X0, yg * p.NY, cb * p.CB, b, bar);
        if (p.two) tma4(dst + ((p.box_bytes + 127u) & ~127u), &mB, 0, yg * p.NY, cb * p.CB, b, bar);
    };
    const int n = i1 - i0;
    if (tid == 0) for (int k = 0; k < p.S - 1 && k < n; ++k) issue(i0 + k, k);
    float acc = 0.f;
    int s = 0; uint32_t ph = 0;
    for (int it = 0; it < n; ++it) {
        if (tid == 0 && it + p.S - 1 < n) {
            if (it > 0) {
                const int sp = s == 0 ? p.S - 1 : s - 1;
                mbar_wait(smem_u32(&empty[sp]), s == 0 ? ph ^ 1u : ph);
            }
            issue(i0 + it + p.S - 1, s == 0 ? p.S - 1 : s - 1);
        }
        __syncwarp();
        mbar_wait(smem_u32(&full[s]), ph);
        acc += reinterpret_cast<const float *>(smem + (size_t)s * stage_bytes)[tid];
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&empty[s]));
        if (++s == p.S) { s = 0; ph ^= 1u; }
    }
    if (acc == 123.456f) sink[0] = acc;
}

// plain coalesced sweep of the same tensors: thread = float4, grid-stride
__global__ void __launch_bounds__(256) ldg_kernel(const float4 *a, const float4 *b, size_t n4, float *sink, int unroll8) {
    float acc = 0.f;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n4; i += 4 * stride) {
        const float4 v0 = __ldcg(a + i), v1 = __ldcg(a + i + stride), v2 = __ldcg(a + i + 2 * stride), v3 = __ldcg(a + i + 3 * stride);
        const float4 w0 = __ldcg(b + i), w1 = __ldcg(b + i + stride), w2 = __ldcg(b + i + 2 * stride), w3 = __ldcg(b + i + 3 * stride);
        acc += v0.x + v1.y + v2.z + v3.w + w0.x + w1.y + w2.z + w3.w;
    }
    for (; i < n4; i += stride) acc += __ldcg(a + i).x + __ldcg(b + i).x;
    if (acc == 123.456f) sink[0] = acc;
}

__global__ void fill_kernel(float *p, size_t n, unsigned seed) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        unsigned h = (unsigned)i * 2654435761u + seed; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        p[i] = (h & 1u) ? 0.f : (float)(h >> 8) * (1.0f / 16777216.0f);
    }
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                             const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char **argv) {
    const int B = 4, C = 128, H = 72, W = 120;
    const int NSETS = argc > 1 ? atoi(argv[1]) : 8;
    const int fill = argc > 2 ? atoi(argv[2]) : 0;      // 1: pseudo-random data (half zeros, like post-ReLU features)
    const size_t pad_mb = argc > 3 ? atoi(argv[3]) : 0; // untouched allocation between the sets (MB)
    const size_t n = (size_t)B * C * H * W;
    std::vector<float *> A(NSETS), Bt(NSETS);
    for (int i = 0; i < NSETS; ++i) {
        CK(cudaMalloc(&A[i], n * 4)); CK(cudaMalloc(&Bt[i], n * 4)); CK(cudaMemset(A[i], 0, n * 4)); CK(cudaMemset(Bt[i], 0, n * 4));
        if (fill) { fill_kernel<<<1024, 256>>>(A[i], n, 17u * i + 1u); fill_kernel<<<1024, 256>>>(Bt[i], n, 17u * i + 9u); }
        if (pad_mb) { void *padp; CK(cudaMalloc(&padp, pad_mb << 20)); }
    }
    CK(cudaDeviceSynchronize());
    float *sink; CK(cudaMalloc(&sink, 4));
    void *sym = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q));
    EncodeFn enc = (EncodeFn)sym;
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaFuncSetAttribute(ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));

    auto run_ring = [&](int BW, int X0, int CB, int NY, int S, int bps, int two, int threads, int l2promo) {
        P p{}; p.B = B; p.C = C; p.H = H; p.W = W; p.BW = BW; p.X0 = X0; p.CB = CB; p.NY = NY; p.S = S; p.two = two;
        p.box_bytes = (uint32_t)BW * NY * CB * 4;
        p.nitems = B * (H / NY) * (C / CB);
        const size_t smem = (size_t)S * ((p.box_bytes + 127u) & ~127u) * (two ? 2 : 1) + 128;
        if (smem * bps > 220 * 1024 || smem > 200 * 1024) { printf("  (skipped: smem %zu x %d)\n", smem, bps); return; }
        int grid = sms * bps; if (grid > p.nitems) grid = p.nitems;
        p.ipb = (p.nitems + grid - 1) / grid; grid = (p.nitems + p.ipb - 1) / p.ipb;
        std::vector<CUtensorMap> mA(NSETS), mB(NSETS);
        const cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)B};
        const cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)C * H * W * 4};
        const cuuint32_t box[4] = {(cuuint32_t)BW, (cuuint32_t)NY, (cuuint32_t)CB, 1}, es[4] = {1, 1, 1, 1};
        const CUtensorMapL2promotion promo = l2promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : l2promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
        for (int i = 0; i < NSETS; ++i) {
            if (enc(&mA[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, A[i], dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS ||
                enc(&mB[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, Bt[i], dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("  (encode failed)\n"); return; }
        }
        for (int i = 0; i < NSETS; ++i) ring_kernel<<<grid, threads, smem>>>(p, mA[i], mB[i], sink);
        CK(cudaDeviceSynchronize());
        const int reps = 5;
        CK(cudaEventRecord(e0));
        for (int r = 0; r < reps; ++r) for (int i = 0; i < NSETS; ++i) ring_kernel<<<grid, threads, smem>>>(p, mA[i], mB[i], sink);
        CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        const double us = ms * 1e3 / (reps * NSETS), bytes = (double)n * 4 * (two ? 2 : 1);
        printf("ring BW=%3d X0=%3d CB=%3d NY=%d S=%2d bps=%d two=%d thr=%3d promo=%d grid=%3d ipb=%2d smem=%6zu : %7.2f us  %7.1f GB/s (valid bytes)\n",
               BW, X0, CB, NY, S, bps, two, threads, l2promo, grid, p.ipb, smem, us, bytes / us / 1e3);
    };
    auto run_ldg = [&](int bps) {
        const int grid = sms * bps;
        for (int i = 0; i < NSETS; ++i) ldg_kernel<<<grid, 256>>>((const float4 *)A[i], (const float4 *)Bt[i], n / 4, sink, 0);
        CK(cudaDeviceSynchronize());
        const int reps = 5;
        CK(cudaEventRecord(e0));
        for (int r = 0; r < reps; ++r) for (int i = 0; i < NSETS; ++i) ldg_kernel<<<grid, 256>>>((const float4 *)A[i], (const float4 *)Bt[i], n / 4, sink, 0);
        CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        const double us = ms * 1e3 / (reps * NSETS), bytes = (double)n * 4 * 2;
        printf("ldg  bps=%d grid=%d : %7.2f us  %7.1f GB/s\n", bps, grid, us, bytes / us / 1e3);
    };

    printf("tensor pair 2 x %.1f MB, %d SMs, back-to-back launches (no PDL), %d rotating sets, fill=%d pad=%zu MB\n", n * 4 / 1e6, sms, NSETS, fill, pad_mb);
    for (int bps : {2, 4, 8}) run_ldg(bps);
    printf("-- the kernels' boxes: one image row, CB channels\n");
    run_ring(144, -24, 32, 1, 4, 1, 1, 256, 1);     // corr_bwd_wide
    run_ring(144, -24, 32, 1, 2, 2, 1, 256, 1);     // two blocks per SM
    run_ring(144, -24, 16, 1, 4, 2, 1, 256, 1);
    run_ring(144, -24, 16, 1, 8, 1, 1, 256, 1);
    run_ring(120, 0, 32, 1, 4, 1, 1, 256, 1);       // no halo, no out-of-bounds columns
    run_ring(120, 0, 32, 1, 2, 2, 1, 256, 1);
    run_ring(128, 0, 32, 1, 4, 1, 1, 256, 1);       // 8 columns out of bounds on the right
    run_ring(120, 0, 32, 1, 4, 1, 1, 256, 0);       // L2 promotion none / 256 B
    run_ring(120, 0, 32, 1, 4, 1, 1, 256, 2);
    run_ring(120, 0, 64, 1, 2, 1, 1, 256, 1);
    run_ring(120, 0, 128, 1, 2, 1, 0, 256, 1);      // one tensor, all channels of a row in one box
    run_ring(120, 0, 32, 1, 4, 1, 0, 256, 1);
    printf("-- several rows per box (contiguous per channel)\n");
    run_ring(120, 0, 32, 2, 2, 1, 1, 256, 1);
    run_ring(120, 0, 16, 2, 4, 1, 1, 256, 1);
    run_ring(120, 0, 16, 4, 2, 1, 1, 256, 1);
    run_ring(120, 0, 8, 4, 4, 1, 1, 256, 1);
    run_ring(120, 0, 8, 8, 2, 1, 1, 256, 1);
    run_ring(120, 0, 4, 8, 4, 1, 1, 256, 1);
    run_ring(120, 0, 2, 24, 4, 1, 1, 256, 1);
    run_ring(120, 0, 1, 72, 2, 1, 1, 256, 1);       // a whole channel plane per box (34.5 KB contiguous)
    run_ring(120, 0, 1, 72, 2, 2, 1, 256, 1);
    run_ring(144, -24, 8, 4, 4, 1, 1, 256, 1);
    run_ring(144, -24, 8, 8, 2, 1, 1, 256, 1);
    return 0;
}
