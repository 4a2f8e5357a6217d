// asm_loss.cu -- the atmospheric-scattering-model reconstruction loss of the training step,
// value and both gradients in ONE launch per pyramid level.  Upstream, losses/DSSMDLoss.py:81-111
// (RecoveredCleanImagesLoss; V2 :115-145 differs only in the shape it takes the airlight in), called once per level of the
// transmission pyramid by trainfiles/dssmd_traner_new.py:298-302:
//     haze_s, clean_s = F.interpolate(., scale_factor=1/scale, mode='bilinear', align_corners=False)   if scale != 1
//     rec  = (haze_s - A * (1 - t)) / t             A: [B,1,1,1] predicted airlight, t: [B,1,h,w] predicted transmission
//     loss = mean | clamp(rec, 0, 1) - clean_s |
// (upstream's `if transmission_map.min==0` compares a bound method with 0 and is always False, so the `+1e-4` branch
// never runs.)  As ~14 ATen kernels forward and ~20 backward per level this moves ~60 floats per pixel; here a thread
// reads t and the 2 x C image values of a pixel (through the 4-tap downsample when scale != 1), and writes
//     g_t_unit[b,0,y,x] = sum_c sgn_c * 1[0 <= rec_c <= 1] * (A - rec_c) / t        (d rec/dt = (A - rec)/t)
// while the block reduces  sum |diff|  and  sum_c sgn_c * 1[..] * (-(1 - t)/t)  (d rec/dA) into one pair of partial
// sums per block (fixed order: deterministic).  The host side sums the partials and scales by 1/n.
// Algorithmic bytes at scale 1: (1 + 2C) reads + 1 write = 8 floats per pixel for C = 3.
#include "warp_common.cuh"

namespace dssmd {

namespace {

constexpr int kThreads = 256;

struct PixelOut {
    float abs_sum, ga, gt;
};

__device__ __forceinline__ void asm_pixel(float t, float A, const float *h, const float *c, int C, PixelOut &o) {
    using N = Num<float>;
    const float om = N::sub(1.0f, t);
    const float am = N::mul(A, om);
    const float dA = N::div(-om, t);
    float s = 0.f, ga = 0.f, gt = 0.f;
    for (int k = 0; k < C; ++k) {
        const float rec = N::div(N::sub(h[k], am), t);
        const float rc = fminf(fmaxf(rec, 0.0f), 1.0f);
        const float diff = N::sub(rc, c[k]);
        s += fabsf(diff);
        const float sg = (diff > 0.f) ? 1.0f : (diff < 0.f ? -1.0f : 0.0f);
        const bool pass = (rec >= 0.0f) && (rec <= 1.0f);   // torch.clamp passes the gradient on the closed interval
        if (pass && sg != 0.f) {
            ga = fmaf(sg, dA, ga);
            gt = fmaf(sg, N::div(N::sub(A, rec), t), gt);
        }
    }
    o.abs_sum = s; o.ga = ga; o.gt = gt;
}

// grid = (blocks over the h*w plane, B); block partial sums -> partial[(b * gridDim.x + blockIdx.x) * 2 + {0, 1}]
template <int C, bool FULL>
__global__ void __launch_bounds__(kThreads) asm_recovered_l1_kernel(const float *trans, const float *airlight,
                                                                    const float *haze, const float *clean,
                                                                    float *__restrict__ g_t_unit, float *__restrict__ partial,
                                                                    int Cdyn, int H0, int W0, int h, int w) {
    pdl_launch_dependents();
    pdl_wait();
    const int b = blockIdx.y;
    const int CC = C > 0 ? C : Cdyn;
    const long long hw = (long long)h * w, HW0 = (long long)H0 * W0;
    const float A = ld_cg(airlight + b);
    const float *tb = trans + (size_t)b * hw;
    const float *hb = haze + (size_t)b * CC * HW0, *cb = clean + (size_t)b * CC * HW0;
    float s_abs = 0.f, s_ga = 0.f;
    if constexpr (FULL && C > 0) {
        // scale 1, w % 4 == 0, 16-byte aligned tensors: 4 pixels per thread, 128-bit loads
        const long long nv = hw / 4;
        for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < nv; i += (long long)gridDim.x * kThreads) {
            const float4 tv = ld_cg(reinterpret_cast<const float4 *>(tb) + i);
            float4 hv[C], cv[C];
#pragma unroll
            for (int k = 0; k < C; ++k) {
                hv[k] = ld_cg(reinterpret_cast<const float4 *>(hb + (size_t)k * HW0) + i);
                cv[k] = ld_cg(reinterpret_cast<const float4 *>(cb + (size_t)k * HW0) + i);
            }
            const float tt[4] = {tv.x, tv.y, tv.z, tv.w};
            float gt[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float hh[C], cc[C];
#pragma unroll
                for (int k = 0; k < C; ++k) {
                    hh[k] = reinterpret_cast<const float *>(&hv[k])[j];
                    cc[k] = reinterpret_cast<const float *>(&cv[k])[j];
                }
                PixelOut o;
                asm_pixel(tt[j], A, hh, cc, C, o);
                s_abs += o.abs_sum; s_ga += o.ga; gt[j] = o.gt;
            }
            if (g_t_unit) st_global_cs(reinterpret_cast<float4 *>(g_t_unit + (size_t)b * hw) + i, make_float4(gt[0], gt[1], gt[2], gt[3]));
        }
    } else {
        for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < hw; i += (long long)gridDim.x * kThreads) {
            const int y = (int)(i / w), x = (int)(i - (long long)y * w);
            const Src<float> ys = src_index<float>(y, H0, h), xs = src_index<float>(x, W0, w);
            float hh[8], cc[8];
            PixelOut acc{0.f, 0.f, 0.f};
            const float t = ld_cg(tb + i);
            for (int k0 = 0; k0 < CC; k0 += 8) {   // channels in groups of 8 (images have 3)
                const int kc = CC - k0 < 8 ? CC - k0 : 8;
                for (int k = 0; k < kc; ++k) {
                    hh[k] = resize_at<float>(hb + (size_t)(k0 + k) * HW0, H0, W0, ys, xs);
                    cc[k] = resize_at<float>(cb + (size_t)(k0 + k) * HW0, H0, W0, ys, xs);
                }
                PixelOut o;
                asm_pixel(t, A, hh, cc, kc, o);
                acc.abs_sum += o.abs_sum; acc.ga += o.ga; acc.gt += o.gt;
            }
            s_abs += acc.abs_sum; s_ga += acc.ga;
            if (g_t_unit) g_t_unit[(size_t)b * hw + i] = acc.gt;
        }
    }
    // block reduction in a fixed order: lanes by butterfly, warps serially
    __shared__ float red[2][kThreads / 32];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        s_abs += __shfl_xor_sync(0xffffffffu, s_abs, off);
        s_ga += __shfl_xor_sync(0xffffffffu, s_ga, off);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s_abs; red[1][threadIdx.x >> 5] = s_ga; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, g = 0.f;
        for (int k = 0; k < kThreads / 32; ++k) { a += red[0][k]; g += red[1][k]; }
        float *p = partial + ((size_t)b * gridDim.x + blockIdx.x) * 2;
        p[0] = a; p[1] = g;
    }
}

// Even integer scale S (the pyramid levels 1/2, 1/4, 1/8): with align_corners=False the source coordinate of output pixel x
// is S*x + S/2 - 1/2, i.e. exactly between source pixels S*x + S/2 - 1 and S*x + S/2, both weights 1/2 -- no coordinate
// arithmetic, and the taps come as vector loads: S == 2: a thread owns 2 output pixels = one float4 per source row;
// S == 4: one output pixel = the middle of one float4; S == 8: two scalars.  Same operations, in the same order, as
// resize_at() performs for these weights.
template <int S>
__global__ void __launch_bounds__(kThreads) asm_recovered_l1_even_kernel(const float *trans, const float *airlight,
                                                                         const float *haze, const float *clean,
                                                                         float *__restrict__ g_t_unit, float *__restrict__ partial,
                                                                         int C, int H0, int W0, int h, int w) {
    using N = Num<float>;
    pdl_launch_dependents();
    pdl_wait();
    constexpr int OPX = S == 2 ? 2 : 1;        // output pixels per thread
    const int b = blockIdx.y;
    const long long hw = (long long)h * w, HW0 = (long long)H0 * W0;
    const float A = ld_cg(airlight + b);
    const float *tb = trans + (size_t)b * hw;
    const float *hb = haze + (size_t)b * C * HW0, *cb = clean + (size_t)b * C * HW0;
    const int wq = w / OPX;
    const long long nq = (long long)h * wq;
    float s_abs = 0.f, s_ga = 0.f;
    auto down = [&](float a, float b2, float c2, float d2) {   // rows (a b) over (c d), all weights 1/2
        const float top = N::add(N::mul(0.5f, a), N::mul(0.5f, b2));
        const float bot = N::add(N::mul(0.5f, c2), N::mul(0.5f, d2));
        return N::add(N::mul(0.5f, top), N::mul(0.5f, bot));
    };
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < nq; i += (long long)gridDim.x * kThreads) {
        const int y = (int)(i / wq), xq = (int)(i - (long long)y * wq);
        const int x = xq * OPX;
        const size_t r0 = (size_t)(S * y + S / 2 - 1) * W0, r1 = r0 + W0;
        float t[OPX];
        if constexpr (OPX == 2) {
            const float2 tv = ld_cg(reinterpret_cast<const float2 *>(tb + (size_t)y * w + x));
            t[0] = tv.x; t[1] = tv.y;
        } else {
            t[0] = ld_cg(tb + (size_t)y * w + x);
        }
        PixelOut acc[OPX];
#pragma unroll
        for (int j = 0; j < OPX; ++j) acc[j] = PixelOut{0.f, 0.f, 0.f};
        for (int k = 0; k < C; ++k) {
            const float *hp = hb + (size_t)k * HW0, *cp = cb + (size_t)k * HW0;
            float hh[OPX], cc[OPX];
            if constexpr (S == 2) {
                const float4 h0 = ld_cg(reinterpret_cast<const float4 *>(hp + r0 + 2 * x)), h1 = ld_cg(reinterpret_cast<const float4 *>(hp + r1 + 2 * x));
                const float4 c0 = ld_cg(reinterpret_cast<const float4 *>(cp + r0 + 2 * x)), c1 = ld_cg(reinterpret_cast<const float4 *>(cp + r1 + 2 * x));
                hh[0] = down(h0.x, h0.y, h1.x, h1.y); hh[1] = down(h0.z, h0.w, h1.z, h1.w);
                cc[0] = down(c0.x, c0.y, c1.x, c1.y); cc[1] = down(c0.z, c0.w, c1.z, c1.w);
            } else if constexpr (S == 4) {
                const float4 h0 = ld_cg(reinterpret_cast<const float4 *>(hp + r0 + 4 * x)), h1 = ld_cg(reinterpret_cast<const float4 *>(hp + r1 + 4 * x));
                const float4 c0 = ld_cg(reinterpret_cast<const float4 *>(cp + r0 + 4 * x)), c1 = ld_cg(reinterpret_cast<const float4 *>(cp + r1 + 4 * x));
                hh[0] = down(h0.y, h0.z, h1.y, h1.z);
                cc[0] = down(c0.y, c0.z, c1.y, c1.z);
            } else {
                const size_t o = (size_t)S * x + S / 2 - 1;
                hh[0] = down(ld_cg(hp + r0 + o), ld_cg(hp + r0 + o + 1), ld_cg(hp + r1 + o), ld_cg(hp + r1 + o + 1));
                cc[0] = down(ld_cg(cp + r0 + o), ld_cg(cp + r0 + o + 1), ld_cg(cp + r1 + o), ld_cg(cp + r1 + o + 1));
            }
#pragma unroll
            for (int j = 0; j < OPX; ++j) {
                PixelOut o1;
                asm_pixel(t[j], A, &hh[j], &cc[j], 1, o1);
                acc[j].abs_sum += o1.abs_sum; acc[j].ga += o1.ga; acc[j].gt += o1.gt;
            }
        }
#pragma unroll
        for (int j = 0; j < OPX; ++j) { s_abs += acc[j].abs_sum; s_ga += acc[j].ga; }
        if (g_t_unit) {
            float *gp = g_t_unit + (size_t)b * hw + (size_t)y * w + x;
            if constexpr (OPX == 2) *reinterpret_cast<float2 *>(gp) = make_float2(acc[0].gt, acc[1].gt);
            else *gp = acc[0].gt;
        }
    }
    __shared__ float red[2][kThreads / 32];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        s_abs += __shfl_xor_sync(0xffffffffu, s_abs, off);
        s_ga += __shfl_xor_sync(0xffffffffu, s_ga, off);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s_abs; red[1][threadIdx.x >> 5] = s_ga; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, g = 0.f;
        for (int k = 0; k < kThreads / 32; ++k) { a += red[0][k]; g += red[1][k]; }
        float *p = partial + ((size_t)b * gridDim.x + blockIdx.x) * 2;
        p[0] = a; p[1] = g;
    }
}

// out[0] = sum(partial[:, :, 0]) * inv_n (the loss), out[1 + b] = sum(partial[b, :, 1]): one block, fixed order (deterministic)
__global__ void __launch_bounds__(256) asm_reduce_partials_kernel(const float *partial, float *__restrict__ out, int B, int nblk, float inv_n) {
    __shared__ float s_a[8], s_g[8];
    pdl_launch_dependents();
    pdl_wait();
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    float total = 0.f;
    for (int b = 0; b < B; ++b) {
        const float *p = partial + (size_t)b * nblk * 2;
        float a = 0.f, g = 0.f;
        for (int i = tid; i < nblk; i += 256) { a += ld_cg(p + 2 * i); g += ld_cg(p + 2 * i + 1); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); g += __shfl_xor_sync(0xffffffffu, g, o); }
        if (lane == 0) { s_a[wid] = a; s_g[wid] = g; }
        __syncthreads();
        if (tid == 0) {
            float sa = 0.f, sg = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) { sa += s_a[k]; sg += s_g[k]; }
            total += sa;
            out[1 + b] = sg;
        }
        __syncthreads();
    }
    if (tid == 0) out[0] = total * inv_n;
}

int asm_blocks(int h, int w, int B) {
    const long long hw = (long long)h * w;
    long long nb = (hw + kThreads - 1) / kThreads;               // at most one pixel per thread; the cap makes the big levels loop
    const long long cap = ((long long)sm_count() * 8 + B - 1) / B;   // ~8 resident blocks per SM over the whole batch
    if (nb > cap) nb = cap;
    if (nb < 1) nb = 1;
    return (int)nb;
}

}  // namespace

}  // namespace dssmd

using namespace dssmd;

extern "C" int dssmd_asm_recovered_l1_blocks(int B, int h, int w) {
    if (B <= 0 || h <= 0 || w <= 0) return 0;
    return asm_blocks(h, w, B);
}

static int asm_fwd_impl(const void *trans, const void *airlight, const void *haze, const void *clean, void *g_trans_unit,
                        void *partial, int B, int C, int H0, int W0, int h, int w, int dtype, void *stream);

extern "C" int dssmd_asm_recovered_l1_fwd(const void *trans, const void *airlight, const void *haze, const void *clean, void *g_trans_unit,
                                          void *partial, int B, int C, int H0, int W0, int h, int w, int dtype, void *stream) {
    return asm_fwd_impl(trans, airlight, haze, clean, g_trans_unit, partial, B, C, H0, W0, h, w, dtype, stream);
}

extern "C" int dssmd_asm_recovered_l1_loss(const void *trans, const void *airlight, const void *haze, const void *clean, void *g_trans_unit,
                                           void *partial, void *out, int B, int C, int H0, int W0, int h, int w, int dtype, void *stream) {
    DSSMD_REQUIRE(out, "asm_recovered_l1_loss: null output pointer");
    if (const int rc = asm_fwd_impl(trans, airlight, haze, clean, g_trans_unit, partial, B, C, H0, W0, h, w, dtype, stream)) return rc;
    const float inv_n = (float)(1.0 / ((double)B * C * h * w));
    const cudaError_t e = launch_pdl(asm_reduce_partials_kernel, dim3(1), dim3(256), 0, static_cast<cudaStream_t>(stream),
                                     (const float *)partial, (float *)out, B, asm_blocks(h, w, B), inv_n);
    if (e != cudaSuccess) return fail(DSSMD_ERR_CUDA, "asm_recovered_l1_loss: launch failed: %s", cudaGetErrorString(e));
    return check_launch("asm_recovered_l1_loss");
}

static int asm_fwd_impl(const void *trans, const void *airlight, const void *haze, const void *clean, void *g_trans_unit,
                                          void *partial, int B, int C, int H0, int W0, int h, int w, int dtype, void *stream) {
    DSSMD_REQUIRE(trans && airlight && haze && clean && partial, "asm_recovered_l1: null pointer");
    DSSMD_REQUIRE(B > 0 && C > 0 && H0 > 0 && W0 > 0 && h > 0 && w > 0, "asm_recovered_l1: sizes must be positive");
    DSSMD_REQUIRE(B <= 65535, "asm_recovered_l1: B too large");
    DSSMD_REQUIRE(w <= W0 && W0 % w == 0 && H0 == (long long)h * (W0 / w),
                  "asm_recovered_l1: images %dx%d are not an integer multiple of the transmission map %dx%d", H0, W0, h, w);
    if (dtype != DSSMD_F32) return fail(DSSMD_ERR_UNSUPPORTED, "asm_recovered_l1: only fp32 is implemented (dtype %d)", dtype);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int nb = asm_blocks(h, w, B);
    const uintptr_t all = reinterpret_cast<uintptr_t>(trans) | reinterpret_cast<uintptr_t>(haze) | reinterpret_cast<uintptr_t>(clean) |
                          reinterpret_cast<uintptr_t>(g_trans_unit);
    const bool full = (W0 == w) && (((long long)h * w) % 4 == 0) && (all & 15) == 0;
    const dim3 grid((unsigned)nb, (unsigned)B);
    cudaError_t e;
#define DSSMD_ASM_ARGS (const float *)trans, (const float *)airlight, (const float *)haze, (const float *)clean, (float *)g_trans_unit, \
                       (float *)partial, C, H0, W0, h, w
    const int scale = W0 / w;
    const bool even_ok = (W0 % 4 == 0) && (w % 2 == 0) && (all & 15) == 0;   // float4 rows of the images, float2 pairs of t / g_t
    if (scale == 2 && even_ok) e = launch_pdl(asm_recovered_l1_even_kernel<2>, grid, dim3(kThreads), 0, s, DSSMD_ASM_ARGS);
    else if (scale == 4 && even_ok) e = launch_pdl(asm_recovered_l1_even_kernel<4>, grid, dim3(kThreads), 0, s, DSSMD_ASM_ARGS);
    else if (scale == 8) e = launch_pdl(asm_recovered_l1_even_kernel<8>, grid, dim3(kThreads), 0, s, DSSMD_ASM_ARGS);
    else if (full && C == 3) e = launch_pdl(asm_recovered_l1_kernel<3, true>, grid, dim3(kThreads), 0, s, DSSMD_ASM_ARGS);
    else if (full && C == 1) e = launch_pdl(asm_recovered_l1_kernel<1, true>, grid, dim3(kThreads), 0, s, DSSMD_ASM_ARGS);
    else e = launch_pdl(asm_recovered_l1_kernel<0, false>, grid, dim3(kThreads), 0, s, DSSMD_ASM_ARGS);
#undef DSSMD_ASM_ARGS
    if (e != cudaSuccess) return fail(DSSMD_ERR_CUDA, "asm_recovered_l1: launch failed: %s", cudaGetErrorString(e));
    return check_launch("asm_recovered_l1");
}
