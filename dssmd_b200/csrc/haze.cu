// haze.cu -- on-GPU haze synthesis of the training loop's input side, one kernel for
// what upstream runs as ~13 elementwise ATen kernels per batch (trainfiles/dssmd_traner_new.py:250-256):
//   convert_disp_to_depth   DeBug/inference.py:57-64     depth = (focal * baseline) / disp
//   depth2trans             DeBug/inference.py:126-134   t     = exp((-depth) * beta)
//   recover_haze_images     DeBug/inference.py:95-112    haze  = (clear * t) + airlight * (1 - t)
// focal, baseline, beta, airlight are per-sample scalars ([B]).  Every operation is individually rounded in upstream's
// order (no FMA contraction), in the dtype upstream's type promotion computes in: the data loaders hand over the
// scalars as float64 (np.array(float), dataloader/ComplexSceneflowLoader.py:112-119), so with float32 images upstream
// computes and returns float64 -- template <I = image / disparity element, S = compute and output element>.
// Inputs are read with ld.global.cg and are not __restrict__ (see ld_cg in common.cuh: a predecessor launched with
// programmatic serialization may still be writing them when this kernel becomes resident).
// HBM-bound: (C + 1) * sizeof(I) read, (C + 2) * sizeof(S) written per pixel when everything is requested.
#include "common.cuh"

namespace dssmd {

namespace {

template <typename S> struct HNum;
template <> struct HNum<float> {
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
    static __device__ __forceinline__ float exp(float a) { return expf(a); }
};
template <> struct HNum<double> {
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
    static __device__ __forceinline__ double exp(double a) { return ::exp(a); }
};

template <typename T, int V> struct Vec { T v[V]; };
template <typename T, int V>
__device__ __forceinline__ Vec<T, V> ldv(const T *p) {
    Vec<T, V> r;
    if constexpr (V == 1) r.v[0] = ld_cg(p);
    else if constexpr (sizeof(T) * V == 16) {
        const float4 q = ld_cg(reinterpret_cast<const float4 *>(p));
        r = *reinterpret_cast<const Vec<T, V> *>(&q);
    } else {  // 32 bytes: 4 doubles
        const float4 q0 = ld_cg(reinterpret_cast<const float4 *>(p)), q1 = ld_cg(reinterpret_cast<const float4 *>(p) + 1);
        float4 *d = reinterpret_cast<float4 *>(&r);
        d[0] = q0; d[1] = q1;
    }
    return r;
}
template <typename T, int V>
__device__ __forceinline__ void stv(T *p, const Vec<T, V> &r) {
    if constexpr (V == 1) *p = r.v[0];
    else if constexpr (sizeof(T) * V == 16) st_global_cs(reinterpret_cast<float4 *>(p), *reinterpret_cast<const float4 *>(&r));
    else {
        const float4 *s = reinterpret_cast<const float4 *>(&r);
        st_global_cs(reinterpret_cast<float4 *>(p), s[0]);
        st_global_cs(reinterpret_cast<float4 *>(p) + 1, s[1]);
    }
}

// grid = (chunks of the H*W plane, B); thread = V consecutive pixels, all channels
template <typename I, typename S, int V>
__global__ void __launch_bounds__(256) haze_kernel(const I *clear, const I *src, int src_is_disp,
                                                   const S *baseline, const S *focal,
                                                   const S *beta, const S *airlight,
                                                   S *__restrict__ haze, S *__restrict__ trans, S *__restrict__ depth,
                                                   int C, long long HW) {
    using N = HNum<S>;
    pdl_launch_dependents();
    pdl_wait();
    const int b = blockIdx.y;
    const S fb = src_is_disp ? N::mul(ld_cg(focal + b), ld_cg(baseline + b)) : (S)0;
    const bool need_t = haze || trans;
    const S be = need_t ? ld_cg(beta + b) : (S)0;
    const S A = haze ? ld_cg(airlight + b) : (S)0;
    const long long nvec = HW / V;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
        const long long o = i * V;
        const Vec<I, V> sv = ldv<I, V>(src + (size_t)b * HW + o);
        Vec<S, V> dv, tv, av;
#pragma unroll
        for (int k = 0; k < V; ++k) {
            dv.v[k] = src_is_disp ? N::div(fb, (S)sv.v[k]) : (S)sv.v[k];
            if (need_t) {
                tv.v[k] = N::exp(N::mul(-dv.v[k], be));
                av.v[k] = N::mul(A, N::sub((S)1, tv.v[k]));
            }
        }
        if (depth) stv<S, V>(depth + (size_t)b * HW + o, dv);
        if (trans) stv<S, V>(trans + (size_t)b * HW + o, tv);
        if (haze) {
            for (int c = 0; c < C; ++c) {
                const size_t pc = ((size_t)b * C + c) * HW + o;
                const Vec<I, V> cv = ldv<I, V>(clear + pc);
                Vec<S, V> hv;
#pragma unroll
                for (int k = 0; k < V; ++k) hv.v[k] = N::add(N::mul((S)cv.v[k], tv.v[k]), av.v[k]);
                stv<S, V>(haze + pc, hv);
            }
        }
    }
}

template <typename I, typename S>
int haze_typed(const void *clear, const void *src, int src_is_disp, const void *baseline, const void *focal, const void *beta,
               const void *airlight, void *haze, void *trans, void *depth, int B, int C, int H, int W, cudaStream_t s) {
    const long long HW = (long long)H * W;
    const uintptr_t all = reinterpret_cast<uintptr_t>(clear) | reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(haze) |
                          reinterpret_cast<uintptr_t>(trans) | reinterpret_cast<uintptr_t>(depth);
    const bool vec = HW % 4 == 0 && (all & 15) == 0;   // 4 pixels per thread as 16-byte loads / stores
    auto launch = [&](auto kern, int V) {
        const long long nvec = HW / V;
        long long gx = (nvec + 255) / 256;
        const long long cap = (long long)sm_count() * 8 / (B < 8 ? B : 8) + 1;   // a few waves over the whole batch
        if (gx > cap) gx = cap;
        if (gx < 1) gx = 1;
        const cudaError_t e = launch_pdl(kern, dim3((unsigned)gx, (unsigned)B), dim3(256), 0, s, (const I *)clear, (const I *)src, src_is_disp,
                                         (const S *)baseline, (const S *)focal, (const S *)beta, (const S *)airlight,
                                         (S *)haze, (S *)trans, (S *)depth, C, HW);
        if (e != cudaSuccess) return fail(DSSMD_ERR_CUDA, "haze: launch failed: %s", cudaGetErrorString(e));
        return check_launch("haze");
    };
    if (vec) return launch(haze_kernel<I, S, 4>, 4);
    return launch(haze_kernel<I, S, 1>, 1);
}

}  // namespace

}  // namespace dssmd

using namespace dssmd;

extern "C" int dssmd_haze_fwd(const void *clear, const void *src, int src_is_disparity, const void *baseline, const void *focal_length,
                              const void *beta, const void *airlight, void *haze, void *trans, void *depth,
                              int B, int C, int H, int W, int image_dtype, int compute_dtype, void *stream) {
    DSSMD_REQUIRE(src, "haze: null source pointer");
    DSSMD_REQUIRE(B > 0 && H > 0 && W > 0 && C >= 0, "haze: sizes must be positive (B=%d C=%d H=%d W=%d)", B, C, H, W);
    DSSMD_REQUIRE(B <= 65535, "haze: B too large");
    DSSMD_REQUIRE(!haze || (clear && airlight && C > 0), "haze: the hazy image needs clear, airlight and C > 0");
    DSSMD_REQUIRE(!(haze || trans) || beta, "haze: transmission needs beta");
    DSSMD_REQUIRE(!src_is_disparity || (baseline && focal_length), "haze: disparity -> depth needs baseline and focal_length");
    if (!haze && !trans && !depth) return DSSMD_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (image_dtype == DSSMD_F32 && compute_dtype == DSSMD_F32)
        return haze_typed<float, float>(clear, src, src_is_disparity, baseline, focal_length, beta, airlight, haze, trans, depth, B, C, H, W, s);
    if (image_dtype == DSSMD_F32 && compute_dtype == DSSMD_F64)
        return haze_typed<float, double>(clear, src, src_is_disparity, baseline, focal_length, beta, airlight, haze, trans, depth, B, C, H, W, s);
    if (image_dtype == DSSMD_F64 && compute_dtype == DSSMD_F64)
        return haze_typed<double, double>(clear, src, src_is_disparity, baseline, focal_length, beta, airlight, haze, trans, depth, B, C, H, W, s);
    return fail(DSSMD_ERR_UNSUPPORTED, "haze: image dtype %d with compute dtype %d is not implemented (f32/f32, f32/f64, f64/f64)", image_dtype, compute_dtype);
}
