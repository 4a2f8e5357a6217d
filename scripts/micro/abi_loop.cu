// Times dssmd_corr_bwd / dssmd_corr_fwd through the C ABI with plain back-to-back stream launches (no CUDA graph, no torch) over
// rotating input sets: a cross-check of the Python harness (bench.py / scripts/kbench.py replay CUDA graphs).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o abi_loop abi_loop.cu -I../../include -L../../dssmd_b200 -ldssmd_b200 -Xlinker -rpath -Xlinker '$ORIGIN/../../dssmd_b200'
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "dssmd_b200.h"
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)
__global__ void fill_kernel(float *p, size_t n, unsigned seed) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        unsigned h = (unsigned)i * 2654435761u + seed; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        p[i] = (h & 1u) ? 0.f : (float)(h >> 8) * (1.0f / 16777216.0f);
    }
}
int main(int argc, char **argv) {
    const int B = 4, C = 128, H = 72, W = 120, D = 24;
    const int NSETS = argc > 1 ? atoi(argv[1]) : 8;
    const size_t nf = (size_t)B * C * H * W, nv = (size_t)B * D * H * W;
    std::vector<float *> L(NSETS), R(NSETS), g(NSETS), gL(NSETS), gR(NSETS), vol(NSETS);
    for (int i = 0; i < NSETS; ++i) {
        CK(cudaMalloc(&L[i], nf * 4)); CK(cudaMalloc(&R[i], nf * 4)); CK(cudaMalloc(&gL[i], nf * 4)); CK(cudaMalloc(&gR[i], nf * 4));
        CK(cudaMalloc(&g[i], nv * 4)); CK(cudaMalloc(&vol[i], nv * 4));
        fill_kernel<<<1024, 256>>>(L[i], nf, 3u + i); fill_kernel<<<1024, 256>>>(R[i], nf, 103u + i); fill_kernel<<<1024, 256>>>(g[i], nv, 203u + i);
    }
    CK(cudaDeviceSynchronize());
    cudaStream_t st; CK(cudaStreamCreate(&st));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int which = 0; which < 2; ++which) {
        auto call = [&](int i) {
            const int rc = which == 0 ? dssmd_corr_bwd(g[i], L[i], R[i], gL[i], gR[i], B, C, H, W, D, 0, st)
                                      : dssmd_corr_fwd(L[i], R[i], vol[i], B, C, H, W, D, 0, st);
            if (rc) { printf("error %d: %s\n", rc, dssmd_last_error()); exit(1); }
        };
        for (int i = 0; i < NSETS; ++i) call(i);
        CK(cudaStreamSynchronize(st));
        const int reps = 10;
        CK(cudaEventRecord(e0, st));
        for (int r = 0; r < reps; ++r) for (int i = 0; i < NSETS; ++i) call(i);
        CK(cudaEventRecord(e1, st)); CK(cudaStreamSynchronize(st));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("%s: %.2f us per call (plain launches, %d sets)\n", which == 0 ? "corr_bwd" : "corr_fwd", ms * 1e3 / (reps * NSETS), NSETS);
    }
    return 0;
}
