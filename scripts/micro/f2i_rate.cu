// Issue-rate microbenchmark: float -> int conversions on sm_100a (F2I.NTZ on the XU pipe) against the magic-number FFMA + IADD
// that replaces them in the warp backward, and int -> float (I2FP).  4 warps per scheduler, 16 independent values per lane.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o f2i_rate f2i_rate.cu && ./f2i_rate
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(512) k(int *out, float a, float b, int iters) {
    float v[16];
    int q[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { v[i] = (float)(threadIdx.x + i) * 0.37f; q[i] = 0; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == 0) q[i] += __float2int_rn(v[i] * a);                                             // FMUL + F2I + IADD
            else if (MODE == 1) q[i] += __float_as_int(fmaf(v[i], a, 12582912.0f)) - 0x4B400000;         // FFMA + IADD (x2, may fuse)
            else if (MODE == 2) { v[i] = (float)q[i] * a + b; q[i] += 3; }                               // I2FP + FFMA + IADD
            else q[i] += __float_as_int(v[i] * a);                                                       // FMUL + IADD (baseline)
        }
        a += 1e-7f;
    }
    int s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += q[i] + (int)v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    int sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    int *out;
    cudaMalloc(&out, sizeof(int) * sms * 512);
    const int iters = 20000;
    const char *names[4] = {"FMUL + F2I.NTZ + IADD   ", "FFMA(magic) + IADD      ", "I2FP + FFMA + IADD      ", "FMUL + IADD (no convert)"};
    for (int mode = 0; mode < 4; ++mode) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<sms, 512>>>(out, 1.0001f, 0.5f, iters);
            else if (mode == 1) k<1><<<sms, 512>>>(out, 1.0001f, 0.5f, iters);
            else if (mode == 2) k<2><<<sms, 512>>>(out, 1.0001f, 0.5f, iters);
            else k<3><<<sms, 512>>>(out, 1.0001f, 0.5f, iters);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            const double n = (double)sms * 512 * iters * 16.0;
            if (rep == 2)
                printf("%s: %.3f ms, %.2f conversions/clk/SM at %d MHz nominal\n", names[mode], ms, n / (ms * 1e-3) / sms / (khz * 1e3), khz / 1000);
        }
    }
    return 0;
}
