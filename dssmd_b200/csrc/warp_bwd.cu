// warp_bwd.cu -- backward of the disparity warp (replaces what autograd derives for
// upstream utils/disparity_warper.py:83-106: aten::grid_sampler_2d_backward composed
// with the coordinate arithmetic of :95-99).
//
// g_img is a scatter-add of g * tap-weight into 4 taps (ATen: global float atomics).
// Here every block owns one image row (b, yy) of g_img, accumulates it in shared
// memory from the few output rows whose vertical taps reach it, and writes it once:
// no global atomics, every g_img element written exactly once, deterministic.
// g_disp is a gather produced by the block with yy == y.
// Algorithmic HBM bytes: (C g + C img + 1 disp + C g_img + 1 g_disp) * s per pixel.
#include "warp_common.cuh"

#include <cstdlib>

namespace dssmd {

template <typename S>
int warp_bwd_v1(const void *g, const void *img, const void *disp, void *g_img, void *g_disp,
                int B, int C, int H, int W, int pad, cudaStream_t s);

namespace {

constexpr int kCC = 4;       // channels accumulated per pass over a row (covers RGB)
constexpr int kMaxPass = 6;  // (output row, vertical weight) pairs that can reach one image row

// ---------------------------------------------------------------------------------
// Fast path (fp32).  Block = one image row (b, yy) of g_img, accumulated in shared
// memory as EXACT integers:
//   * shared-memory float atomics do not exist on sm_100a (atomicAdd(float) compiles
//     to a CAS spin loop); native ATOMS.ADD works on 32-bit integers, so every
//     contribution  g * wy * wx  is quantised to q = rn(v * 2^e) with e chosen per block
//     such that |q| < 2^30, split into q = hi * 65536 + lo and added with two integer
//     atomics (lo: unsigned, hi: signed).  The sums cannot overflow for rows shorter
//     than 65536 pixels, integer addition is associative, so the result is
//     deterministic and exact up to the 2^-30 (relative to the row's largest |g|)
//     quantisation of each term -- tighter than fp32 accumulation.
//   * only output rows y whose vertical taps {y0, y0+1} include yy contribute; the
//     (y, wy) pairs come from the per-row constant iy.
// g_disp of output row yy is a gather done by the same block.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 4) warp_bwd_fixed_kernel(const float *__restrict__ g, const float *__restrict__ img,
                                                                 const float *__restrict__ disp, float *__restrict__ g_img,
                                                                 float *__restrict__ g_disp,
                                                                 int B, int C, int H, int W, int border) {
    using S = float;
    using N = Num<S>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned int *acc_lo = reinterpret_cast<unsigned int *>(smem_raw);       // [kCC][W]
    int *acc_hi = reinterpret_cast<int *>(smem_raw) + (size_t)kCC * W;       // [kCC][W]
    __shared__ int s_pass_y[kMaxPass];
    __shared__ float s_pass_w[kMaxPass];
    __shared__ int s_npass;
    __shared__ unsigned int s_maxbits;

    const int yy = blockIdx.x % H;
    const int b = blockIdx.x / H;
    const long long HW = (long long)H * W;
    const S xscale = N::div((S)W, (S)2);        // d ix / d gx
    const S tscale = N::div((S)2, (S)(W - 1));  // d gx / d t
    const int tid = threadIdx.x, nt = blockDim.x;

    if (tid == 0) {
        int np = 0;
        const int ylo = yy - 2 < 0 ? 0 : yy - 2;
        const int yhi = yy + 2 > H - 1 ? H - 1 : yy + 2;
        for (int y = ylo; y <= yhi; ++y) {
            const S iy_raw = unnorm<S>((S)y, H);
            const Axis<S> ya = make_axis<S>(border ? clamp_border<S>(iy_raw, H) : iy_raw, H);
            if (ya.in0 && ya.i0 == yy && np < kMaxPass) { s_pass_y[np] = y; s_pass_w[np] = ya.w0; ++np; }
            if (ya.in1 && ya.i0 + 1 == yy && np < kMaxPass) { s_pass_y[np] = y; s_pass_w[np] = ya.w1; ++np; }
        }
        s_npass = np;
        s_maxbits = 0u;
    }

    // ---- g_disp of output row yy (a gather) ------------------------------------------------
    if (g_disp) {
        const S iy_raw = unnorm<S>((S)yy, H);
        const Axis<S> ya = make_axis<S>(border ? clamp_border<S>(iy_raw, H) : iy_raw, H);
        for (int x = tid; x < W; x += nt) {
            const S dv = disp[(long long)b * HW + (long long)yy * W + x];
            S ix = unnorm<S>(N::sub((S)x, dv), W);
            S gmult = xscale;
            if (border) {
                // clip_coordinates_set_grad: borders count as out of bounds for the gradient
                if (ix <= (S)0) { ix = (S)0; gmult = (S)0; }
                else if (ix >= (S)(W - 1)) { ix = (S)(W - 1); gmult = (S)0; }
            }
            const Axis<S> xa = make_axis<S>(ix, W);
            S gix = (S)0;
            for (int c = 0; c < C; ++c) {
                const long long plane = ((long long)b * C + c) * HW;
                const S go = g[plane + (long long)yy * W + x];
                const S *p0 = img + plane + (long long)ya.i0 * W;
                const S *p1 = p0 + W;
                // d(sample)/d(ix) over in-bounds taps, ATen order nw, ne, sw, se
                if (xa.in0 && ya.in0) gix = N::sub(gix, N::mul(N::mul(p0[xa.i0], ya.w0), go));
                if (xa.in1 && ya.in0) gix = N::add(gix, N::mul(N::mul(p0[xa.i0 + 1], ya.w0), go));
                if (xa.in0 && ya.in1) gix = N::sub(gix, N::mul(N::mul(p1[xa.i0], ya.w1), go));
                if (xa.in1 && ya.in1) gix = N::add(gix, N::mul(N::mul(p1[xa.i0 + 1], ya.w1), go));
            }
            // chain rule: ix <- gx (W/2, 0 where clipped) <- t (2/(W-1)) <- disp (-1)
            g_disp[(long long)b * HW + (long long)yy * W + x] = -N::mul(tscale, N::mul(gmult, gix));
        }
    }
    if (!g_img) return;
    __syncthreads();
    const int npass = s_npass;

    for (int c0 = 0; c0 < C; c0 += kCC) {
        const int nc = (C - c0 < kCC) ? C - c0 : kCC;
        // ---- scale: largest |g * wy| over everything this block will add -------------------
        unsigned int mbits = 0u;
        for (int ps = 0; ps < npass; ++ps) {
            const S wy = s_pass_w[ps];
            const S *grow = g + ((long long)b * C + c0) * HW + (long long)s_pass_y[ps] * W;
            for (int x = tid; x < W; x += nt)
                for (int c = 0; c < nc; ++c) {
                    const unsigned int bits = __float_as_uint(N::mul(grow[(long long)c * HW + x], wy)) & 0x7fffffffu;
                    mbits = bits > mbits ? bits : mbits;  // NaN/inf bit patterns compare largest
                }
        }
        for (int i = tid; i < nc * W; i += nt) { acc_lo[i] = 0u; acc_hi[i] = 0; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned int other = __shfl_xor_sync(0xffffffffu, mbits, o);
            mbits = other > mbits ? other : mbits;
        }
        if ((tid & 31) == 0) atomicMax(&s_maxbits, mbits);
        __syncthreads();
        const unsigned int maxbits = s_maxbits;
        const int mexp = (int)(maxbits >> 23);          // biased exponent of the largest magnitude
        const bool finite = mexp != 0xff;
        // |v| < 2^(mexp-126)  =>  |v * 2^e| < 2^30 for e = 156 - mexp; e is kept within what a
        // float power of two can express (only gradients below ~1e-29 lose bits to that clamp)
        int e = 156 - (mexp > 0 ? mexp : 1);
        if (e > 127) e = 127;
        const S up = __uint_as_float((unsigned)(e + 127) << 23);                                   // 2^e
        const S down = (e < 127) ? __uint_as_float((unsigned)(127 - e) << 23) : __uint_as_float(0x00400000u);  // 2^-e

        if (finite && maxbits != 0u) {
            for (int ps = 0; ps < npass; ++ps) {
                const int y = s_pass_y[ps];
                const S wy = s_pass_w[ps];
                const S *grow = g + ((long long)b * C + c0) * HW + (long long)y * W;
                for (int x = tid; x < W; x += nt) {
                    const S dv = disp[(long long)b * HW + (long long)y * W + x];
                    S ix = unnorm<S>(N::sub((S)x, dv), W);
                    if (border) ix = clamp_border<S>(ix, W);
                    const Axis<S> xa = make_axis<S>(ix, W);
                    if (!(xa.in0 || xa.in1)) continue;
                    const S wl = N::mul(wy, xa.w0), wr = N::mul(wy, xa.w1);
                    for (int c = 0; c < nc; ++c) {
                        const S go = grow[(long long)c * HW + x];
                        if (xa.in0) {
                            const int q = __float2int_rn(N::mul(N::mul(go, wl), up));
                            atomicAdd(&acc_lo[c * W + xa.i0], (unsigned int)q & 0xffffu);
                            atomicAdd(&acc_hi[c * W + xa.i0], q >> 16);
                        }
                        if (xa.in1) {
                            const int q = __float2int_rn(N::mul(N::mul(go, wr), up));
                            atomicAdd(&acc_lo[c * W + xa.i0 + 1], (unsigned int)q & 0xffffu);
                            atomicAdd(&acc_hi[c * W + xa.i0 + 1], q >> 16);
                        }
                    }
                }
            }
        }
        __syncthreads();
        if (tid == 0) s_maxbits = 0u;  // everyone has read it; visible to the next chunk after the sync below
        // ---- write the row once -----------------------------------------------------------------
        for (int i = tid; i < nc * W; i += nt) {
            const int c = i / W, x = i - c * W;
            S v;
            if (!finite) {
                v = __uint_as_float(0x7fc00000u);  // a non-finite gradient poisons the row chunk
            } else {
                const long long tot = ((long long)acc_hi[i] << 16) + (long long)acc_lo[i];
                v = N::mul((S)tot, down);
            }
            g_img[((long long)b * C + c0 + c) * HW + (long long)yy * W + x] = v;
        }
        __syncthreads();
    }
}

int env_int(const char *name, int dflt) {
    const char *v = std::getenv(name);
    return (v && *v) ? std::atoi(v) : dflt;
}

template <typename S>
int warp_bwd_typed(const void *g, const void *img, const void *disp, void *g_img, void *g_disp,
                   int B, int C, int H, int W, int pad, cudaStream_t s) {
    const int variant = env_int("DSSMD_WARP_BWD_VARIANT", 1);
    if constexpr (sizeof(S) == 4) {
        const size_t smem = (size_t)2 * kCC * W * sizeof(int);
        if (variant != 0 && W <= 65535 && smem <= 160 * 1024) {
            if (smem > 48 * 1024) {
                cudaError_t e = cudaFuncSetAttribute(warp_bwd_fixed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (e != cudaSuccess) return fail(DSSMD_ERR_CUDA, "warp_bwd: cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(e));
            }
            warp_bwd_fixed_kernel<<<(unsigned)(B * H), 256, smem, s>>>((const float *)g, (const float *)img, (const float *)disp,
                                                                    (float *)g_img, (float *)g_disp, B, C, H, W,
                                                                    pad == DSSMD_PAD_BORDER);
            return check_launch("warp_bwd");
        }
    }
    return warp_bwd_v1<S>(g, img, disp, g_img, g_disp, B, C, H, W, pad, s);
}

}  // namespace

}  // namespace dssmd

using namespace dssmd;

extern "C" int dssmd_warp_bwd(const void *g, const void *img, const void *disp, void *g_img, void *g_disp,
                              int B, int C, int H, int W, int padding_mode, int dtype, void *stream) {
    DSSMD_REQUIRE(g && img && disp, "warp_bwd: null input pointer");
    DSSMD_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, "warp_bwd: sizes must be positive (B=%d C=%d H=%d W=%d)", B, C, H, W);
    DSSMD_REQUIRE(padding_mode == DSSMD_PAD_ZEROS || padding_mode == DSSMD_PAD_BORDER, "warp_bwd: unknown padding_mode %d", padding_mode);
    DSSMD_REQUIRE((long long)B * H < (1LL << 31), "warp_bwd: B*H too large");
    if (!g_img && !g_disp) return DSSMD_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    switch (dtype) {
        case DSSMD_F32: return warp_bwd_typed<float>(g, img, disp, g_img, g_disp, B, C, H, W, padding_mode, s);
        case DSSMD_F64: return warp_bwd_typed<double>(g, img, disp, g_img, g_disp, B, C, H, W, padding_mode, s);
        case DSSMD_BF16:
        case DSSMD_F16:
            return fail(DSSMD_ERR_UNSUPPORTED, "warp_bwd: 16-bit dtypes are not supported (coordinates need fp32; see DESIGN.md)");
        default: return fail(DSSMD_ERR_INVALID_ARGUMENT, "warp_bwd: unknown dtype %d", dtype);
    }
}
