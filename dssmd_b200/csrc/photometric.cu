// photometric.cu -- C ABI of the fused L1 photometric term (see warp_lean.cu for the kernel).
#include "common.cuh"

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
