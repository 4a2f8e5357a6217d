// photometric.cu -- C ABI of the fused L1 photometric term (see warp_lean.cu for the kernel) and the division self test.
#include "warp_lean.cuh"

namespace dssmd {
int photometric_l1_lean_f32(const float *imgL, const float *imgR, const float *disp, float *g_unit, float *partial, int *neg_flag,
                            int B, int C, int H, int W, cudaStream_t s);
}  // namespace dssmd

using namespace dssmd;

extern "C" int dssmd_photometric_l1_fwd(const void *imgL, const void *imgR, const void *disp, void *g_disp_unit, void *row_sums,
                                        int *neg_flag, int B, int C, int H, int W, int dtype, void *stream) {
    DSSMD_REQUIRE(imgL && imgR && disp && row_sums, "photometric_l1: null pointer");
    DSSMD_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, "photometric_l1: sizes must be positive (B=%d C=%d H=%d W=%d)", B, C, H, W);
    if (dtype != DSSMD_F32) return fail(DSSMD_ERR_UNSUPPORTED, "photometric_l1: only fp32 is implemented (dtype %d)", dtype);
    const int rc = photometric_l1_lean_f32((const float *)imgL, (const float *)imgR, (const float *)disp, (float *)g_disp_unit,
                                           (float *)row_sums, neg_flag, B, C, H, W, static_cast<cudaStream_t>(stream));
    if (rc < 0)
        return fail(DSSMD_ERR_UNSUPPORTED,
                    "photometric_l1: shape not supported by the fused kernel (needs C <= 4, W %% 4 == 0, H >= 2, 16-byte aligned "
                    "tensors; got C=%d H=%d W=%d) -- compose LRwarp_error(...).abs() instead", C, H, W);
    return rc;
}

namespace {
__global__ void selftest_division_kernel(int W_lo, int W_hi, unsigned long long *mismatches) {
    const int W = W_lo + (int)blockIdx.y;
    const float b = (float)(W - 1);
    const float y = div_rn_recip(b);
    const int k_hi = 64 * W_hi;
    unsigned long long bad = 0;
    for (int k = -16384 + (int)(blockIdx.x * blockDim.x + threadIdx.x); k <= k_hi; k += (int)(gridDim.x * blockDim.x)) {
        const float a = (float)k * 0.015625f;
        const float cand[3] = {a, __uint_as_float(__float_as_uint(a) + 1u), __uint_as_float(__float_as_uint(a) - 1u)};
#pragma unroll
        for (int i = 0; i < 3; ++i)
            if (__float_as_uint(div_rn_by(cand[i], b, y)) != __float_as_uint(__fdiv_rn(cand[i], b)) && cand[i] == cand[i] && cand[i] != 0.0f) ++bad;
    }
    if (bad) atomicAdd(mismatches, bad);
}
}  // namespace

extern "C" int dssmd_selftest_division(int W_lo, int W_hi, unsigned long long *mismatches, void *stream) {
    DSSMD_REQUIRE(mismatches, "selftest_division: null pointer");
    DSSMD_REQUIRE(W_lo >= 2 && W_hi >= W_lo && W_hi <= 65536, "selftest_division: need 2 <= W_lo <= W_hi <= 65536 (got %d, %d)", W_lo, W_hi);
    selftest_division_kernel<<<dim3(64, (unsigned)(W_hi - W_lo + 1)), 256, 0, static_cast<cudaStream_t>(stream)>>>(W_lo, W_hi, mismatches);
    return check_launch("selftest_division");
}
