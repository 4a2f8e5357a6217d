// Issue-rate microbenchmark: FFMA vs FFMA2 (fma.rn.f32x2) on sm_100a, 4 warps per scheduler, 8 independent chains per lane.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o ffma2_rate ffma2_rate.cu && ./ffma2_rate
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(512) k(float *out, float a, float b, int iters) {
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = (float)(threadIdx.x + i);
    unsigned long long p[8], A, B;
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(A) : "f"(a), "f"(a));
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(B) : "f"(b), "f"(b));
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(acc[2 * i]), "f"(acc[2 * i + 1]));
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], a, b);
        } else {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(A), "l"(B));
        }
    }
    float s = 0.f;
    if (MODE == 0) {
#pragma unroll
        for (int i = 0; i < 16; ++i) s += acc[i];
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) { float x, y; asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(p[i])); s += x + y; }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    int sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    float *out;
    cudaMalloc(&out, sizeof(float) * sms * 512);
    const int iters = 20000;
    for (int mode = 0; mode < 2; ++mode) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<sms, 512>>>(out, 1.0001f, 0.5f, iters);
            else k<1><<<sms, 512>>>(out, 1.0001f, 0.5f, iters);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            const double fma = (double)sms * 512 * iters * 64.0;
            if (rep == 2)
                printf("%s: %.3f ms, %.1f TFMA/s, %.1f FMA/clk/SM at %d MHz nominal\n", mode ? "FFMA2" : "FFMA ", ms, fma / ms / 1e9,
                       fma / (ms * 1e-3) / sms / (khz * 1e3), khz / 1000);
        }
    }
    return 0;
}
