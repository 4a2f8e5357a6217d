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

// warp_lean.cu: instruction-lean horizontal-scatter kernel (K1) for image-like shapes; -1 = not eligible
int warp_bwd_band_f32(const float *g, const float *img, const float *disp, float *g_img, float *g_disp,
                      int B, int C, int H, int W, int border, cudaStream_t s);
// warp_bwd_band2.cu: the band kernel rewritten around its instruction count (default); -1 = not eligible
int warp_bwd_band2_f32(const float *g, const float *img, const float *disp, float *g_img, float *g_disp,
                       int B, int C, int H, int W, int border, cudaStream_t s);
int warp_bwd_lean_f32(const float *g, const float *img, const float *disp, float *gV, float *g_disp,
                      int B, int C, int H, int W, int border, cudaStream_t s);

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

// ---------------------------------------------------------------------------------
// Same algorithm, shaped for latency: the block covers the whole row with PXT pixels
// per thread (no strided loops), the <= 6 (output row, weight) candidates sit in fixed
// shared slots, and all address arithmetic is hoisted.  Used for W <= 4096.
// ---------------------------------------------------------------------------------
template <int PXT>
__global__ void __launch_bounds__(1024, 1) warp_bwd_row_kernel(const float *__restrict__ g, const float *__restrict__ img,
                                                                const float *__restrict__ disp, float *__restrict__ g_img,
                                                                float *__restrict__ g_disp,
                                                                int B, int C, int H, int W, int border, int flags) {
    using S = float;
    using N = Num<S>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned int *acc_lo = reinterpret_cast<unsigned int *>(smem_raw);   // [kCC][W]
    int *acc_hi = reinterpret_cast<int *>(smem_raw) + (size_t)kCC * W;   // [kCC][W]
    __shared__ int s_y[6];
    __shared__ float s_w[6];
    __shared__ unsigned int s_maxbits[2];

    const int yy = blockIdx.x % H;
    const int b = blockIdx.x / H;
    const int HW = H * W;  // < 2^31 checked on the host
    const int tid = threadIdx.x, nt = blockDim.x;
    const float *dispb = disp + (long long)b * HW;
    const S halfW = N::div((S)W, (S)2);
    const S wm1 = (S)(W - 1);

    if (tid < 3) {
        const int y = yy - 1 + tid;
        int y0s = -1, y1s = -1;
        float w0 = 0.f, w1 = 0.f;
        if (y >= 0 && y < H) {
            const S iy_raw = unnorm<S>((S)y, H);
            const Axis<S> ya = make_axis<S>(border ? clamp_border<S>(iy_raw, H) : iy_raw, H);
            if (ya.in0 && ya.i0 == yy) { y0s = y; w0 = ya.w0; }
            if (ya.in1 && ya.i0 + 1 == yy) { y1s = y; w1 = ya.w1; }
        }
        s_y[2 * tid] = y0s; s_w[2 * tid] = w0;
        s_y[2 * tid + 1] = y1s; s_w[2 * tid + 1] = w1;
    }
    if (tid < 2) s_maxbits[tid] = 0u;

    // ---- g_disp of output row yy (a gather) ------------------------------------------------
    if (g_disp) {
        const S iy_raw = unnorm<S>((S)yy, H);
        const Axis<S> ya = make_axis<S>(border ? clamp_border<S>(iy_raw, H) : iy_raw, H);
        const S tscale = N::div((S)2, wm1);  // d gx / d t
#pragma unroll
        for (int j = 0; j < PXT; ++j) {
            const int x = tid + j * nt;
            if (x >= W) break;
            const S dv = dispb[yy * W + x];
            S ix = unnorm<S>(N::sub((S)x, dv), W);
            S gmult = halfW;
            if (border) {
                // clip_coordinates_set_grad: borders count as out of bounds for the gradient
                if (ix <= (S)0) { ix = (S)0; gmult = (S)0; }
                else if (ix >= wm1) { ix = wm1; gmult = (S)0; }
            }
            const Axis<S> xa = make_axis<S>(ix, W);
            S gix = (S)0;
            const float *gp = g + (long long)b * C * HW + yy * W + x;
            const float *p0 = img + (long long)b * C * HW + ya.i0 * W + xa.i0;
            for (int c = 0; c < C; ++c, gp += HW, p0 += HW) {
                const S go = *gp;
                // d(sample)/d(ix) over in-bounds taps, ATen order nw, ne, sw, se
                if (xa.in0 && ya.in0) gix = N::sub(gix, N::mul(N::mul(p0[0], ya.w0), go));
                if (xa.in1 && ya.in0) gix = N::add(gix, N::mul(N::mul(p0[1], ya.w0), go));
                if (xa.in0 && ya.in1) gix = N::sub(gix, N::mul(N::mul(p0[W], ya.w1), go));
                if (xa.in1 && ya.in1) gix = N::add(gix, N::mul(N::mul(p0[W + 1], ya.w1), go));
            }
            // chain rule: ix <- gx (W/2, 0 where clipped) <- t (2/(W-1)) <- disp (-1)
            g_disp[(long long)b * HW + yy * W + x] = -N::mul(tscale, N::mul(gmult, gix));
        }
    }
    if (!g_img) return;
    __syncthreads();

    int parity = 0;
    for (int c0 = 0; c0 < C; c0 += kCC) {
        const int nc = (C - c0 < kCC) ? C - c0 : kCC;
        const float *gb = g + ((long long)b * C + c0) * HW;
        // ---- scale: largest |g * wy| this block will add; zero the accumulators ----------
        unsigned int mbits = 0u;
        for (int sl = 0; sl < 6; ++sl) {
            const int y = s_y[sl];
            if (y < 0) continue;
            const S wy = s_w[sl];
#pragma unroll
            for (int j = 0; j < PXT; ++j) {
                const int x = tid + j * nt;
                if (x < W)
                    for (int c = 0; c < nc; ++c) {
                        const unsigned int bits = __float_as_uint(N::mul(gb[c * HW + y * W + x], wy)) & 0x7fffffffu;
                        mbits = bits > mbits ? bits : mbits;  // NaN/inf bit patterns compare largest
                    }
            }
        }
#pragma unroll
        for (int j = 0; j < PXT; ++j) {
            const int x = tid + j * nt;
            if (x < W)
                for (int c = 0; c < nc; ++c) { acc_lo[c * W + x] = 0u; acc_hi[c * W + x] = 0; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned int other = __shfl_xor_sync(0xffffffffu, mbits, o);
            mbits = other > mbits ? other : mbits;
        }
        if ((tid & 31) == 0) atomicMax(&s_maxbits[parity], mbits);
        __syncthreads();
        const unsigned int maxbits = s_maxbits[parity];
        if (tid == 0) s_maxbits[parity ^ 1] = 0u;  // idle slot, next used after the barriers below
        parity ^= 1;
        const int mexp = (int)(maxbits >> 23);
        const bool finite = mexp != 0xff;
        int e = 156 - (mexp > 0 ? mexp : 1);  // |v| < 2^(mexp-126)  =>  |v * 2^e| < 2^30
        if (e > 127) e = 127;
        const S up = __uint_as_float((unsigned)(e + 127) << 23);
        const S down = (e < 127) ? __uint_as_float((unsigned)(127 - e) << 23) : __uint_as_float(0x00400000u);

        if (finite && maxbits != 0u) {
            for (int sl = 0; sl < 6; ++sl) {
                const int y = s_y[sl];
                if (y < 0) continue;
                const S wy = s_w[sl];
#pragma unroll
                for (int j = 0; j < PXT; ++j) {
                    const int x = tid + j * nt;
                    const bool inrow = x < W;
                    const int xc = inrow ? x : W - 1;
                    S ix = unnorm<S>(N::sub((S)xc, dispb[y * W + xc]), W);
                    if (border) ix = clamp_border<S>(ix, W);
                    const Axis<S> xa = make_axis<S>(ix, W);
                    // cells of the two taps (negative = none)
                    const int kL = (inrow && xa.in0) ? xa.i0 : -1000;
                    const int kR = (inrow && xa.in1) ? xa.i0 + 1 : -2000;
                    // x' is close to monotone in x, so the right tap of lane l usually lands on the
                    // left-tap cell of lane l+1: hand it over through a shuffle and halve the atomics
                    // (exact: the merge happens on the quantised integers)
                    const int lane = tid & 31;
                    const int nextL = __shfl_down_sync(0xffffffffu, kL, 1);
                    const int prevR = __shfl_up_sync(0xffffffffu, kR, 1);
                    const bool hand_out = (flags & 1) && lane < 31 && kR >= 0 && nextL == kR;
                    const bool take_in = (flags & 1) && lane > 0 && kL >= 0 && prevR == kL;
                    const S wl = N::mul(wy, xa.w0), wr = N::mul(wy, xa.w1);
                    const float *gp = gb + y * W + xc;
                    unsigned int *lo = acc_lo + xa.i0;
                    int *hi = acc_hi + xa.i0;
                    for (int c = 0; c < nc; ++c, gp += HW, lo += W, hi += W) {
                        const S go = *gp;
                        int q0 = kL >= 0 ? __float2int_rn(N::mul(N::mul(go, wl), up)) : 0;
                        const int q1 = kR >= 0 ? __float2int_rn(N::mul(N::mul(go, wr), up)) : 0;
                        const int qin = __shfl_up_sync(0xffffffffu, q1, 1);
                        if (take_in) q0 += qin;
                        if (!(flags & 2)) {
                            if (kL >= 0) {
                                atomicAdd(lo, (unsigned int)q0 & 0xffffu);
                                atomicAdd(hi, q0 >> 16);
                            }
                            if (kR >= 0 && !hand_out) {
                                atomicAdd(lo + 1, (unsigned int)q1 & 0xffffu);
                                atomicAdd(hi + 1, q1 >> 16);
                            }
                        }
                    }
                }
            }
        }
        __syncthreads();
        // ---- write the row once -----------------------------------------------------------------
        float *ob = g_img + ((long long)b * C + c0) * HW + yy * W;
#pragma unroll
        for (int j = 0; j < PXT; ++j) {
            const int x = tid + j * nt;
            if (x >= W) break;
            for (int c = 0; c < nc; ++c) {
                S v;
                if (!finite) v = __uint_as_float(0x7fc00000u);  // a non-finite gradient poisons the row chunk
                else v = N::mul((S)(((long long)acc_hi[c * W + x] << 16) + (long long)acc_lo[c * W + x]), down);
                ob[c * HW + x] = v;
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------
// Two-kernel path (used when the caller provides a B*C*H*W workspace): the warp is
// separable -- the vertical blend weights are per-row constants -- so
//   K1 (warp_bwd_hscatter_kernel): per OUTPUT row y, gV[c][y][k] = sum_x g[c][y][x] * wx(x->k)
//       (horizontal 2-tap scatter with the exact fixed-point shared atomics; every row is
//       scattered once instead of once per image row it touches) and g_disp[y] (gather);
//   K2 (warp_bwd_vmix_kernel): g_img[c][yy] = sum over the <= 3 rows y whose vertical taps
//       reach yy of wy * gV[c][y]  -- dense, coalesced, read from L2.
// ---------------------------------------------------------------------------------
template <int PXT, int NCT>  // NCT > 0: exactly NCT channels (C == NCT); NCT == 0: any C in chunks of kCC
__global__ void __launch_bounds__(1024, 1) warp_bwd_hscatter_kernel(const float *__restrict__ g, const float *__restrict__ img,
                                                                     const float *__restrict__ disp, float *__restrict__ gV,
                                                                     float *__restrict__ g_disp,
                                                                     int B, int C, int H, int W, int border, int flags) {
    using S = float;
    using N = Num<S>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned int *acc_lo = reinterpret_cast<unsigned int *>(smem_raw);   // [kCC][W]
    int *acc_hi = reinterpret_cast<int *>(smem_raw) + (size_t)kCC * W;   // [kCC][W]
    __shared__ unsigned int s_maxbits[2];

    const int y = blockIdx.x % H;
    const int b = blockIdx.x / H;
    const int HW = H * W;  // C*H*W < 2^31 checked on the host
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
    // per-row constants: computed by one thread (they cost several IEEE divisions), shared by all
    __shared__ float s_rowc[6];  // halfW, tscale, ya.w0, ya.w1 | s_rowi: ya.i0, in-flags
    __shared__ int s_rowi[2];
    if (tid == 0) {
        const S iy_raw = unnorm<S>((S)y, H);
        const Axis<S> a = make_axis<S>(border ? clamp_border<S>(iy_raw, H) : iy_raw, H);
        s_rowc[0] = N::div((S)W, (S)2);
        s_rowc[1] = N::div((S)2, (S)(W - 1));
        s_rowc[2] = a.w0; s_rowc[3] = a.w1;
        s_rowi[0] = a.i0; s_rowi[1] = (a.in0 ? 1 : 0) | (a.in1 ? 2 : 0);
        s_maxbits[0] = 0u; s_maxbits[1] = 0u;
    }
    __syncthreads();
    const S halfW = s_rowc[0];
    const S tscale = s_rowc[1];  // d gx / d t
    const S wm1 = (S)(W - 1);
    Axis<S> ya;
    ya.w0 = s_rowc[2]; ya.w1 = s_rowc[3]; ya.i0 = s_rowi[0];
    ya.in0 = (s_rowi[1] & 1) != 0; ya.in1 = (s_rowi[1] & 2) != 0;
    const S inv_wm1_unused = (S)0; (void)inv_wm1_unused;

    // horizontal taps of this thread's pixels: computed once, reused by every channel chunk
    int kL[PXT], kR[PXT];
    S w0[PXT], w1[PXT], gmult[PXT];
    bool hand_out[PXT], take_in[PXT];
#pragma unroll
    for (int j = 0; j < PXT; ++j) {
        const int x = tid + j * nt;
        const bool inrow = x < W;
        const int xc = inrow ? x : W - 1;
        // unnorm() with the shared W/2: gx = 2*(t/(W-1)) - 1; ix = fma(gx+1, W/2, -0.5)
        const S t = N::sub((S)xc, disp[(long long)b * HW + y * W + xc]);
        const S gxn = N::sub(N::mul((S)2, N::div(t, wm1)), (S)1);
        S ix = N::fma(N::add(gxn, (S)1), halfW, (S)-0.5);
        gmult[j] = halfW;
        if (border) {
            // clip_coordinates_set_grad: borders count as out of bounds for the gradient
            if (ix <= (S)0) { ix = (S)0; gmult[j] = (S)0; }
            else if (ix >= wm1) { ix = wm1; gmult[j] = (S)0; }
            else if (!(ix == ix)) ix = clamp_border<S>(ix, W);
        }
        const Axis<S> xa = make_axis<S>(ix, W);
        kL[j] = (inrow && xa.in0) ? xa.i0 : -1000;
        kR[j] = (inrow && xa.in1) ? xa.i0 + 1 : -2000;
        w0[j] = xa.w0; w1[j] = xa.w1;
        // x' is close to monotone in x: the right tap of lane l usually lands on the left-tap
        // cell of lane l+1 -> hand it over through a shuffle (exact: merged as integers)
        const int nextL = __shfl_down_sync(0xffffffffu, kL[j], 1);
        const int prevR = __shfl_up_sync(0xffffffffu, kR[j], 1);
        hand_out[j] = (flags & 1) && lane < 31 && kR[j] >= 0 && nextL == kR[j];
        take_in[j] = (flags & 1) && lane > 0 && kL[j] >= 0 && prevR == kL[j];
    }

    S gix[PXT];
#pragma unroll
    for (int j = 0; j < PXT; ++j) gix[j] = (S)0;

    int parity = 0;
    constexpr int NCMAX = NCT > 0 ? NCT : kCC;
    for (int c0 = 0; c0 < C; c0 += NCMAX) {
        const int nc = NCT > 0 ? NCT : ((C - c0 < kCC) ? C - c0 : kCC);
        const float *gb = g + ((long long)b * C + c0) * HW + y * W;
        const float *ib = img + ((long long)b * C + c0) * HW;
        S go[PXT][NCMAX];
        unsigned int mbits = 0u;
#pragma unroll
        for (int j = 0; j < PXT; ++j) {
            const int x = tid + j * nt;
            const bool inrow = x < W;
#pragma unroll
            for (int c = 0; c < NCMAX; ++c) {
                go[j][c] = (inrow && c < nc) ? gb[c * HW + x] : (S)0;
                const unsigned int bits = __float_as_uint(go[j][c]) & 0x7fffffffu;
                mbits = bits > mbits ? bits : mbits;  // NaN/inf bit patterns compare largest
                if (inrow && c < nc) { acc_lo[c * W + x] = 0u; acc_hi[c * W + x] = 0; }
            }
            if (g_disp && inrow) {
                // taps of the un-merged axis: left cell i0 = kL or kR-1
                const bool in0 = kL[j] >= 0, in1 = kR[j] >= 0;
                const int i0 = in0 ? kL[j] : kR[j] - 1;
                const float *p0 = ib + ya.i0 * W + i0;
#pragma unroll
                for (int c = 0; c < NCMAX; ++c) {
                    if (c < nc) {
                        const float *pc = p0 + c * HW;
                        // d(sample)/d(ix) over in-bounds taps, ATen order nw, ne, sw, se
                        if (in0 && ya.in0) gix[j] = N::sub(gix[j], N::mul(N::mul(pc[0], ya.w0), go[j][c]));
                        if (in1 && ya.in0) gix[j] = N::add(gix[j], N::mul(N::mul(pc[1], ya.w0), go[j][c]));
                        if (in0 && ya.in1) gix[j] = N::sub(gix[j], N::mul(N::mul(pc[W], ya.w1), go[j][c]));
                        if (in1 && ya.in1) gix[j] = N::add(gix[j], N::mul(N::mul(pc[W + 1], ya.w1), go[j][c]));
                    }
                }
            }
        }
        if (!gV) continue;  // only g_disp requested (block-uniform)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned int other = __shfl_xor_sync(0xffffffffu, mbits, o);
            mbits = other > mbits ? other : mbits;
        }
        if (lane == 0) atomicMax(&s_maxbits[parity], mbits);
        __syncthreads();
        const unsigned int maxbits = s_maxbits[parity];
        if (tid == 0) s_maxbits[parity ^ 1] = 0u;  // idle slot until after the barriers below
        parity ^= 1;
        const int mexp = (int)(maxbits >> 23);
        const bool finite = mexp != 0xff;
        int e = 156 - (mexp > 0 ? mexp : 1);  // |g * wx| <= |g| < 2^(mexp-126)  =>  |v * 2^e| < 2^30
        if (e > 127) e = 127;
        const S up = __uint_as_float((unsigned)(e + 127) << 23);
        const S down = (e < 127) ? __uint_as_float((unsigned)(127 - e) << 23) : __uint_as_float(0x00400000u);

        if (finite && maxbits != 0u) {
#pragma unroll
            for (int j = 0; j < PXT; ++j) {
#pragma unroll
                for (int c = 0; c < NCMAX; ++c) {
                    if (c >= nc) break;
                    int q0 = kL[j] >= 0 ? __float2int_rn(N::mul(N::mul(go[j][c], w0[j]), up)) : 0;
                    const int q1 = kR[j] >= 0 ? __float2int_rn(N::mul(N::mul(go[j][c], w1[j]), up)) : 0;
                    const int qin = __shfl_up_sync(0xffffffffu, q1, 1);
                    if (take_in[j]) q0 += qin;
                    if (kL[j] >= 0) {
                        atomicAdd(&acc_lo[c * W + kL[j]], (unsigned int)q0 & 0xffffu);
                        atomicAdd(&acc_hi[c * W + kL[j]], q0 >> 16);
                    }
                    if (kR[j] >= 0 && !hand_out[j]) {
                        atomicAdd(&acc_lo[c * W + kR[j]], (unsigned int)q1 & 0xffffu);
                        atomicAdd(&acc_hi[c * W + kR[j]], q1 >> 16);
                    }
                }
            }
        }
        __syncthreads();
        float *ob = gV + ((long long)b * C + c0) * HW + y * W;
#pragma unroll
        for (int j = 0; j < PXT; ++j) {
            const int x = tid + j * nt;
            if (x < W) {
#pragma unroll
                for (int c = 0; c < NCMAX; ++c) {
                    if (c < nc) {
                        S v;
                        if (!finite) v = __uint_as_float(0x7fc00000u);  // a non-finite gradient poisons the row
                        else v = N::mul((S)(((long long)acc_hi[c * W + x] << 16) + (long long)acc_lo[c * W + x]), down);
                        ob[c * HW + x] = v;
                    }
                }
            }
        }
        __syncthreads();
    }
    if (g_disp) {
#pragma unroll
        for (int j = 0; j < PXT; ++j) {
            const int x = tid + j * nt;
            // chain rule: ix <- gx (W/2, 0 where clipped) <- t (2/(W-1)) <- disp (-1)
            if (x < W) g_disp[(long long)b * HW + y * W + x] = -N::mul(tscale, N::mul(gmult[j], gix[j]));
        }
    }
}

// ---------------------------------------------------------------------------------
// K1, vectorised (fp32, C <= 4, W % 4 == 0, 16-byte aligned rows): same contract as
// warp_bwd_hscatter_kernel, about a third of its instructions per pixel.
//   * a thread owns 4 CONSECUTIVE pixels: disp / g / gV / g_disp move as float4, and the
//     right tap of pixel j is merged into the left tap of pixel j+1 in registers when they
//     land on the same cell (x' = x - disp is close to monotone), across lanes by shuffle;
//   * the accumulator of a cell is ONE 32-bit integer and every merged tap ONE fire-and-forget
//     shared atomic: contributions are quantised to q = rn(v * 2^e) with e chosen per row from
//     S = sum_x |g[c]| (per channel) such that S * 2^e <= 2^30.  Every cell sum is bounded by S (the tap
//     weights of a pixel add up to <= 1), so no cell can overflow however many taps collect
//     in it (cells at a clamped border collect hundreds), integer adds are associative, and
//     the result is deterministic.  Quantum relative to the row's mean |g|: <= W * 2^-29;
//   * cells live in 4 planes (cell k -> plane k & 3, slot k >> 2) so that lanes, whose
//     cells are 4 apart, hit consecutive banks;
//   * the y-axis taps are recomputed per thread (one division) instead of shared through a barrier.
// ---------------------------------------------------------------------------------
// predicated fire-and-forget 32-bit shared-memory add (native ATOMS/RED on sm_100a, no return value)
__device__ __forceinline__ void red_shared_add(uint32_t addr, int q, bool pred) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\t@p red.shared.add.s32 [%0], %1;\n\t}" ::"r"(addr), "r"(q), "r"((int)pred) : "memory");
}

template <int NCT, int PS>  // PS: slots per shared-memory plane, a compile-time constant >= W/4 + 2 so that every
                            // channel / row / plane offset folds into the LDS / RED immediates
__global__ void __launch_bounds__(PS - 2, 1024 / (PS - 2)) warp_bwd_hs4_kernel(const float *__restrict__ g, const float *__restrict__ img,
                                                            const float *__restrict__ disp, float *__restrict__ gV,
                                                            float *__restrict__ g_disp, int H, int W, int border,
                                                            float halfW, float tscale, float halfH) {
    using S_t = float;
    using N = Num<S_t>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int W4 = W >> 2;
    constexpr int P = PS - 2;                                             // accumulator slots per plane (>= W/4)
    int *acc = reinterpret_cast<int *>(smem_raw);                        // [NCT][4][P]
    float *s_rows = reinterpret_cast<float *>(smem_raw) + NCT * 4 * P;   // [NCT][2][4][PS], only with g_disp
    __shared__ float s_part[NCT][32];                                     // per-warp, per-channel sums of |g|

    const int y = blockIdx.x % H;
    const int b = blockIdx.x / H;
    const int HW = H * W;  // C*H*W < 2^31 checked on the host
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
    const bool active = tid < W4;
    const int x0 = tid * 4;
    // vertical taps of this output row (per thread: one division, no barrier)
    const S_t wm1 = (S_t)(W - 1);
    Axis<S_t> ya;
    {
        const S_t gy = N::sub(N::mul((S_t)2, N::div((S_t)y, (S_t)(H - 1))), (S_t)1);
        const S_t iy_raw = N::fma(N::add(gy, (S_t)1), halfH, (S_t)-0.5);
        ya = make_axis<S_t>(border ? clamp_border<S_t>(iy_raw, H) : iy_raw, H);
    }

    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (g_disp) {
        // the two image rows the gather of d(sample)/d(ix) reads: zero-padded planar copy in shared
        // memory (same layout as warp_fwd_staged_kernel), so every tap is an unpredicated LDS
        float4 st[NCT][2];
        const float *ib = img + (size_t)b * NCT * HW + x0;
#pragma unroll
        for (int c = 0; c < NCT; ++c) {
            st[c][0] = (active && ya.in0) ? *reinterpret_cast<const float4 *>(ib + c * HW + ya.i0 * W) : z4;
            st[c][1] = (active && ya.in1) ? *reinterpret_cast<const float4 *>(ib + c * HW + (ya.i0 + 1) * W) : z4;
        }
#pragma unroll
        for (int c = 0; c < NCT; ++c)
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                float *row = s_rows + (c * 2 + r) * 4 * PS;
                if (active) {
                    row[0 * PS + tid + 1] = st[c][r].x;
                    row[1 * PS + tid + 1] = st[c][r].y;
                    row[2 * PS + tid + 1] = st[c][r].z;
                    row[3 * PS + tid + 1] = st[c][r].w;
                }
            }
        for (int i = tid; i < NCT * 2 * 8; i += nt) {  // the 4 + 4 pad cells of every staged row
            const int rowi = i >> 3, k = i & 7;
            s_rows[rowi * 4 * PS + (k & 3) * PS + ((k >> 2) ? W4 + 1 : 0)] = 0.f;
        }
    }
    if (gV) {
        int4 *a4 = reinterpret_cast<int4 *>(acc);
        for (int i = tid; i < NCT * P; i += nt) a4[i] = make_int4(0, 0, 0, 0);
    }

    const float *grow = g + (size_t)b * NCT * HW + y * W;
    float go[NCT][4];
    float dd[4];
    {
        const float4 d4 = active ? *reinterpret_cast<const float4 *>(disp + (size_t)b * HW + y * W + x0) : z4;
        dd[0] = d4.x; dd[1] = d4.y; dd[2] = d4.z; dd[3] = d4.w;
#pragma unroll
        for (int c = 0; c < NCT; ++c) {
            const float4 v = active ? *reinterpret_cast<const float4 *>(grow + c * HW + x0) : z4;
            go[c][0] = v.x; go[c][1] = v.y; go[c][2] = v.z; go[c][3] = v.w;
        }
    }
    if (gV) {
#pragma unroll
        for (int c = 0; c < NCT; ++c) {
            float asum = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) asum += fabsf(go[c][j]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) asum += __shfl_xor_sync(0xffffffffu, asum, o);
            if (lane == 0) s_part[c][tid >> 5] = asum;
        }
    }

    // horizontal taps
    int kL[4], kR[4];   // cell of the left / right tap, negative when out of bounds
    int xi0[4];         // low tap index as make_axis returns it, always in [-2, W]
    S_t w0[4], w1[4], gmult[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const S_t t = N::sub((S_t)(x0 + j), dd[j]);
        const S_t gxn = N::sub(N::mul((S_t)2, N::div(t, wm1)), (S_t)1);
        S_t ix = N::fma(N::add(gxn, (S_t)1), halfW, (S_t)-0.5);
        gmult[j] = halfW;
        if (border) {
            // clip_coordinates_set_grad: borders count as out of bounds for the gradient
            if (ix <= (S_t)0) { ix = (S_t)0; gmult[j] = (S_t)0; }
            else if (ix >= wm1) { ix = wm1; gmult[j] = (S_t)0; }
            else if (!(ix == ix)) ix = clamp_border<S_t>(ix, W);
        }
        const Axis<S_t> xa = make_axis<S_t>(ix, W);
        xi0[j] = xa.i0;
        kL[j] = (active && xa.in0) ? xa.i0 : -1000;
        kR[j] = (active && xa.in1) ? xa.i0 + 1 : -2000;
        w0[j] = xa.w0; w1[j] = xa.w1;
    }

    // links: right tap of pixel j lands on the left cell of pixel j+1 (j = 3: next lane's pixel 0)
    bool lk[4];
#pragma unroll
    for (int j = 0; j < 3; ++j) lk[j] = kR[j] >= 0 && kR[j] == kL[j + 1];
    {
        const int nextL = __shfl_down_sync(0xffffffffu, kL[0], 1);
        lk[3] = lane < 31 && kR[3] >= 0 && nextL == kR[3];
    }
    const bool tin = __shfl_up_sync(0xffffffffu, (int)lk[3], 1) != 0 && lane > 0;

    __syncthreads();  // accumulators zeroed, image rows staged, per-warp |g| sums written

    if (g_disp && active) {
        float gd[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int p0 = xi0[j] + 4, p1 = xi0[j] + 5;
            const float *t0 = s_rows + (p0 & 3) * PS + (p0 >> 2), *t1 = s_rows + (p1 & 3) * PS + (p1 >> 2);
            S_t gix = (S_t)0;
#pragma unroll
            for (int c = 0; c < NCT; ++c) {
                constexpr int R = 4 * PS;  // one staged row
                // d(sample)/d(ix), ATen order nw, ne, sw, se; out-of-bounds taps read the zero padding
                gix = N::sub(gix, N::mul(N::mul(t0[(2 * c) * R], ya.w0), go[c][j]));
                gix = N::add(gix, N::mul(N::mul(t1[(2 * c) * R], ya.w0), go[c][j]));
                gix = N::sub(gix, N::mul(N::mul(t0[(2 * c + 1) * R], ya.w1), go[c][j]));
                gix = N::add(gix, N::mul(N::mul(t1[(2 * c + 1) * R], ya.w1), go[c][j]));
            }
            // chain rule: ix <- gx (W/2, 0 where clipped) <- t (2/(W-1)) <- disp (-1)
            gd[j] = -N::mul(tscale, N::mul(gmult[j], gix));
        }
        *reinterpret_cast<float4 *>(g_disp + (size_t)b * HW + y * W + x0) = make_float4(gd[0], gd[1], gd[2], gd[3]);
    }
    if (!gV) return;  // only g_disp requested (block-uniform)

    // S[c] = sum of |g[c]| over the row: every warp reduces the per-warp partials with the same
    // butterfly, so all threads hold the same values.  One scale per channel and row.
    S_t up[NCT], down[NCT];
    bool finite[NCT], any = false;
#pragma unroll
    for (int c = 0; c < NCT; ++c) {
        float sc = lane < (nt >> 5) ? s_part[c][lane] : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sc += __shfl_xor_sync(0xffffffffu, sc, o);
        const int sexp = (int)(__float_as_uint(sc) >> 23);  // sc >= 0 or NaN: no sign bit to mask
        finite[c] = sexp < 0xff;                             // inf / NaN gradients poison the row
        int e = 156 - (sexp > 0 ? sexp : 1);                 // sc < 2^(sexp-126)  =>  sc * 2^e < 2^30
        if (e > 127) e = 127;
        up[c] = (finite[c] && sc > 0.f) ? __uint_as_float((unsigned)(e + 127) << 23) : 0.f;  // 0: nothing to add
        down[c] = (e < 127) ? __uint_as_float((unsigned)(127 - e) << 23) : __uint_as_float(0x00400000u);
        any = any || up[c] != 0.f;
    }

    if (any) {
        // byte address of each tap's cell inside channel 0; out-of-bounds taps are never issued
        uint32_t aL[4], aR[4];
        const uint32_t base = smem_u32(acc);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            aL[j] = base + 4u * (uint32_t)((kL[j] & 3) * P + (kL[j] >> 2));
            aR[j] = base + 4u * (uint32_t)((kR[j] & 3) * P + (kR[j] >> 2));
        }
        constexpr uint32_t cstep = 16u * (uint32_t)P;  // bytes between channels
#pragma unroll
        for (int c = 0; c < NCT; ++c) {
            int qL[4], qR[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                // scaling by a power of two is exact, so this is rn(g * w) * 2^e rounded to an integer
                const S_t gs = N::mul(go[c][j], up[c]);
                qL[j] = kL[j] >= 0 ? __float2int_rn(N::mul(gs, w0[j])) : 0;
                qR[j] = kR[j] >= 0 ? __float2int_rn(N::mul(gs, w1[j])) : 0;
            }
            const int qin = __shfl_up_sync(0xffffffffu, qR[3], 1);
            if (tin) qL[0] += qin;
#pragma unroll
            for (int j = 1; j < 4; ++j)
                if (lk[j - 1]) qL[j] += qR[j - 1];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                red_shared_add(aL[j] + c * cstep, qL[j], kL[j] >= 0);
                red_shared_add(aR[j] + c * cstep, qR[j], kR[j] >= 0 && !lk[j]);
            }
        }
    }
    __syncthreads();
    if (!active) return;
    float *ob = gV + (size_t)b * NCT * HW + y * W + x0;
#pragma unroll
    for (int c = 0; c < NCT; ++c) {
        float v[4];
#pragma unroll
        for (int pl = 0; pl < 4; ++pl)
            v[pl] = finite[c] ? N::mul((S_t)acc[c * 4 * P + pl * P + tid], down[c]) : __uint_as_float(0x7fc00000u);
        *reinterpret_cast<float4 *>(ob + c * HW) = make_float4(v[0], v[1], v[2], v[3]);
    }
}

__global__ void __launch_bounds__(256) warp_bwd_vmix_kernel(const float *__restrict__ gV, float *__restrict__ g_img,
                                                             int B, int C, int H, int W, int border, int vec) {
    using S = float;
    using N = Num<S>;
    __shared__ int s_y[6];
    __shared__ float s_w[6];
    const int yy = blockIdx.x % H;
    const int b = blockIdx.x / H;
    const long long HW = (long long)H * W;
    const int tid = threadIdx.x;
    __shared__ int s_n;
    if (tid == 0) {
        int n = 0;
        for (int y = (yy > 0 ? yy - 1 : 0); y <= yy + 1 && y < H; ++y) {
            const S iy_raw = unnorm<S>((S)y, H);
            const Axis<S> ya = make_axis<S>(border ? clamp_border<S>(iy_raw, H) : iy_raw, H);
            if (ya.in0 && ya.i0 == yy) { s_y[n] = y; s_w[n] = ya.w0; ++n; }
            if (ya.in1 && ya.i0 + 1 == yy) { s_y[n] = y; s_w[n] = ya.w1; ++n; }
        }
        s_n = n;
    }
    __syncthreads();
    const int n = s_n;
    const unsigned int HWu = (unsigned)(H * W);
    const float *src0 = gV + (long long)b * C * HW;
    float *dst0 = g_img + (long long)b * C * HW + (long long)yy * W;
    if (vec) {
        const int W4 = W >> 2;
        for (int i = tid; i < C * W4; i += blockDim.x) {
            const int c = i / W4, x4 = i - c * W4;
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int sl = 0; sl < n; ++sl) {
                const float wy = s_w[sl];
                const float4 v = *reinterpret_cast<const float4 *>(src0 + c * HWu + (unsigned)(s_y[sl] * W + 4 * x4));
                a.x = N::fma(wy, v.x, a.x); a.y = N::fma(wy, v.y, a.y);
                a.z = N::fma(wy, v.z, a.z); a.w = N::fma(wy, v.w, a.w);
            }
            st_global_cs(reinterpret_cast<float4 *>(dst0 + c * HWu) + x4, a);
        }
    } else {
        for (int i = tid; i < C * W; i += blockDim.x) {
            const int c = i / W, x = i - c * W;
            float a = 0.f;
            for (int sl = 0; sl < n; ++sl) a = N::fma(s_w[sl], src0[c * HWu + (unsigned)(s_y[sl] * W + x)], a);
            dst0[c * HWu + x] = a;
        }
    }
}

// K2, banded (16-byte aligned rows): block = (image, band of VR consecutive g_img rows).  The VR + 2 candidate
// gV rows of a column are loaded once for the whole band (each gV row is read (VR + 2) / VR times instead of 3),
// and the loads are in flight while VR + 2 lanes work out the vertical taps of the candidate rows in parallel.
// Accumulation order per output (ascending source row, fma chain from 0) is the one of warp_bwd_vmix_kernel.
template <int VR>
__global__ void __launch_bounds__(256) warp_bwd_vmix_band_kernel(const float *__restrict__ gV, float *__restrict__ g_img,
                                                                  int C, int H, int W, int border, float halfH) {
    using S = float;
    using N = Num<S>;
    __shared__ int s_i0[VR + 2], s_in[VR + 2];
    __shared__ float s_w0[VR + 2], s_w1[VR + 2];
    __shared__ float s_wt[VR][3];  // weight of gV row yy - 1 + k in g_img row yy = y_lo + r
    __shared__ int s_on[VR];       // bit k: that row is a tap of yy (a zero weight still propagates NaN, as in ATen)
    const int nbands = (H + VR - 1) / VR;
    const int band = blockIdx.x % nbands, b = blockIdx.x / nbands;
    const int y_lo = band * VR;
    const int tid = threadIdx.x;
    const int W4 = W >> 2, n = C * W4;
    const unsigned HWu = (unsigned)(H * W);
    const float *src0 = gV + (size_t)b * C * HWu;
    float *dst0 = g_img + (size_t)b * C * HWu;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);

    auto load = [&](int i, float4 (&v)[VR + 2]) {
        const int c = i / W4, x4 = i - c * W4;
        const float *p = src0 + c * HWu + 4 * x4;
#pragma unroll
        for (int t = 0; t < VR + 2; ++t) {
            const int y = y_lo - 1 + t;
            v[t] = (y >= 0 && y < H) ? *reinterpret_cast<const float4 *>(p + (unsigned)(y * W)) : z4;
        }
    };
    auto mix = [&](int i, const float4 (&v)[VR + 2]) {
        const int c = i / W4, x4 = i - c * W4;
#pragma unroll
        for (int r = 0; r < VR; ++r) {
            const int yy = y_lo + r;
            if (yy >= H) break;
            const int on = s_on[r];
            float4 a = z4;
#pragma unroll
            for (int k = 0; k < 3; ++k)
                if (on & (1 << k)) {
                    const float wy = s_wt[r][k];
                    a.x = N::fma(wy, v[r + k].x, a.x); a.y = N::fma(wy, v[r + k].y, a.y);
                    a.z = N::fma(wy, v[r + k].z, a.z); a.w = N::fma(wy, v[r + k].w, a.w);
                }
            st_global_cs(reinterpret_cast<float4 *>(dst0 + c * HWu + (unsigned)(yy * W)) + x4, a);
        }
    };

    float4 v[VR + 2];
    if (tid < n) load(tid, v);
    if (tid < VR + 2) {
        const int y = y_lo - 1 + tid;
        int in = 0;
        if (y >= 0 && y < H) {
            const S gy = N::sub(N::mul((S)2, N::div((S)y, (S)(H - 1))), (S)1);
            const S iy_raw = N::fma(N::add(gy, (S)1), halfH, (S)-0.5);
            const Axis<S> ya = make_axis<S>(border ? clamp_border<S>(iy_raw, H) : iy_raw, H);
            s_i0[tid] = ya.i0; s_w0[tid] = ya.w0; s_w1[tid] = ya.w1;
            in = (ya.in0 ? 1 : 0) | (ya.in1 ? 2 : 0);
        }
        s_in[tid] = in;
    }
    if (tid < 32) {
        __syncwarp();
        if (tid < VR) {
            const int yy = y_lo + tid;
            int on = 0;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int t = tid + k;
                const int in = s_in[t];
                float wy = 0.f;
                if ((in & 1) && s_i0[t] == yy) { wy = s_w0[t]; on |= 1 << k; }
                if ((in & 2) && s_i0[t] + 1 == yy) { wy = s_w1[t]; on |= 1 << k; }
                s_wt[tid][k] = wy;
            }
            s_on[tid] = on;
        }
    }
    __syncthreads();
    if (tid < n) mix(tid, v);
    for (int i = tid + 256; i < n; i += 256) {
        load(i, v);
        mix(i, v);
    }
}

int env_int(const char *name, int dflt) {
    if (!tuning_enabled()) return dflt;
    const char *v = std::getenv(name);
    return (v && *v) ? std::atoi(v) : dflt;
}

// K2 launch: banded kernel for 16-byte aligned rows (DSSMD_WARP_BWD_VMIX=0 forces the row kernel)
int launch_vmix(const float *gv, float *gi, int B, int C, int H, int W, int border, int vec, cudaStream_t s) {
    constexpr int VR = 4;
    if (vec && env_int("DSSMD_WARP_BWD_VMIX", 1) != 0 && (long long)B * ceil_div(H, VR) < (1LL << 31)) {
        warp_bwd_vmix_band_kernel<VR><<<(unsigned)(B * ceil_div(H, VR)), 256, 0, s>>>(gv, gi, C, H, W, border, (float)H / 2.0f);
        return check_launch("warp_bwd(vmix band)");
    }
    warp_bwd_vmix_kernel<<<(unsigned)(B * H), 256, 0, s>>>(gv, gi, B, C, H, W, border, vec);
    return check_launch("warp_bwd(vmix)");
}

template <typename S>
int warp_bwd_typed(const void *g, const void *img, const void *disp, void *g_img, void *g_disp,
                   void *workspace, size_t workspace_bytes,
                   int B, int C, int H, int W, int pad, cudaStream_t s) {
    const int variant = env_int("DSSMD_WARP_BWD_VARIANT", 7);   // 7: band2 (default), 6: first band kernel, 5: scatter + vertical mix, ...
    if constexpr (sizeof(S) == 4) {
        const size_t smem = (size_t)2 * kCC * W * sizeof(int);
        const int border = pad == DSSMD_PAD_BORDER;
        const float *gp = (const float *)g, *ip = (const float *)img, *dp = (const float *)disp;
        float *gi = (float *)g_img, *gd = (float *)g_disp;
        if (variant >= 7 && g_img) {   // one-pass band kernel, lean instruction stream: no workspace needed
            const int rc = warp_bwd_band2_f32(gp, ip, dp, gi, gd, B, C, H, W, border, s);
            if (rc >= 0) return rc;
        }
        if (variant >= 6 && g_img) {   // first one-pass band kernel
            const int rc = warp_bwd_band_f32(gp, ip, dp, gi, gd, B, C, H, W, border, s);
            if (rc >= 0) return rc;
        }
        const bool have_ws = workspace && workspace_bytes >= (size_t)B * C * H * W * sizeof(float);
        if (variant >= 4 && C <= 4 && W % 4 == 0 && W <= 4096 && H >= 2 && (long long)C * H * W < (1LL << 31) && (have_ws || !g_img) &&
            ((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(disp) | reinterpret_cast<uintptr_t>(workspace) |
              reinterpret_cast<uintptr_t>(g_disp)) & 15) == 0) {
            void (*kern)(const float *, const float *, const float *, float *, float *, int, int, int, float, float, float) = nullptr;
            const int W4 = W / 4;
            int ps = 0;
#define DSSMD_HS4_PICK(PSV)                                                                                        \
    if (!kern && W4 + 2 <= PSV) {                                                                                  \
        ps = PSV;                                                                                                  \
        kern = C == 1 ? warp_bwd_hs4_kernel<1, PSV> : C == 2 ? warp_bwd_hs4_kernel<2, PSV>                         \
             : C == 3 ? warp_bwd_hs4_kernel<3, PSV> : warp_bwd_hs4_kernel<4, PSV>;                                 \
    }
            DSSMD_HS4_PICK(66) DSSMD_HS4_PICK(130) DSSMD_HS4_PICK(258) DSSMD_HS4_PICK(514) DSSMD_HS4_PICK(1026)
#undef DSSMD_HS4_PICK
            const size_t smem4 = ((size_t)C * 4 * (ps - 2) + (g_disp ? (size_t)C * 2 * 4 * ps : 0)) * sizeof(int);
            if (smem4 > 48 * 1024) {
                cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem4);
                if (e != cudaSuccess) return fail(DSSMD_ERR_CUDA, "warp_bwd: cudaFuncSetAttribute(%zu B): %s", smem4, cudaGetErrorString(e));
            }
            float *gv = g_img ? (float *)workspace : nullptr;
            // W/2, 2/(W-1), H/2: IEEE single divisions, the same values the device would compute
            const float halfW = (float)W / 2.0f, tscale = 2.0f / (float)(W - 1), halfH = (float)H / 2.0f;
            int rc = variant >= 5 ? warp_bwd_lean_f32(gp, ip, dp, gv, gd, B, C, H, W, border, s) : -1;
            if (rc < 0) {
                kern<<<(unsigned)(B * H), round_up(W4, 32), smem4, s>>>(gp, ip, dp, gv, gd, H, W, border, halfW, tscale, halfH);
                rc = check_launch("warp_bwd(hscatter4)");
            }
            if (rc || !g_img) return rc;
            const int vec = (((reinterpret_cast<uintptr_t>(workspace) | reinterpret_cast<uintptr_t>(g_img)) & 15) == 0);
            return launch_vmix(gv, gi, B, C, H, W, border, vec, s);
        }
        if (variant >= 3 && W <= 4096 && (long long)C * H * W < (1LL << 31) && (have_ws || !g_img)) {
            int pxt = W <= 256 ? 1 : (W <= 512 ? 2 : 4);  // ~256-thread blocks: 3-4 resident per SM
            const int o = env_int("DSSMD_WARP_BWD_PXT", 0);
            if ((o == 1 || o == 2 || o == 4) && ceil_div(W, o) <= 1024) pxt = o;
            const int nt = round_up(ceil_div(W, pxt), 32);
            void (*kern)(const float *, const float *, const float *, float *, float *, int, int, int, int, int, int) = nullptr;
#define DSSMD_HS_PICK(P)                                                                            \
    if (pxt == P) kern = C == 1 ? warp_bwd_hscatter_kernel<P, 1> : C == 2 ? warp_bwd_hscatter_kernel<P, 2> \
                       : C == 3 ? warp_bwd_hscatter_kernel<P, 3> : C == 4 ? warp_bwd_hscatter_kernel<P, 4> \
                                                                          : warp_bwd_hscatter_kernel<P, 0>;
            DSSMD_HS_PICK(1) DSSMD_HS_PICK(2) DSSMD_HS_PICK(4)
#undef DSSMD_HS_PICK
            if (smem > 48 * 1024) {
                cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (e != cudaSuccess) return fail(DSSMD_ERR_CUDA, "warp_bwd: cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(e));
            }
            const int flags = env_int("DSSMD_WARP_BWD_FLAGS", 1);
            float *gv = g_img ? (float *)workspace : nullptr;
            kern<<<(unsigned)(B * H), nt, smem, s>>>(gp, ip, dp, gv, gd, B, C, H, W, border, flags);
            int rc = check_launch("warp_bwd(hscatter)");
            if (rc || !g_img) return rc;
            const int vec = (W % 4 == 0) && (((reinterpret_cast<uintptr_t>(workspace) | reinterpret_cast<uintptr_t>(g_img)) & 15) == 0);
            return launch_vmix(gv, gi, B, C, H, W, border, vec, s);
        }
        if (variant >= 2 && W <= 4096 && (long long)C * H * W < (1LL << 31)) {
            int pxt = W <= 512 ? 1 : (W <= 2048 ? 2 : 4);
            const int o = env_int("DSSMD_WARP_BWD_PXT", 0);
            if ((o == 1 || o == 2 || o == 4) && ceil_div(W, o) <= 1024) pxt = o;
            const int nt = round_up(ceil_div(W, pxt), 32);
            auto kern = pxt == 1 ? warp_bwd_row_kernel<1> : (pxt == 2 ? warp_bwd_row_kernel<2> : warp_bwd_row_kernel<4>);
            if (smem > 48 * 1024) {
                cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (e != cudaSuccess) return fail(DSSMD_ERR_CUDA, "warp_bwd: cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(e));
            }
            // flags: bit 0 = merge neighbouring taps through shuffles (default on); bit 1 = debug, skip the atomics
            const int flags = env_int("DSSMD_WARP_BWD_FLAGS", 1);
            kern<<<(unsigned)(B * H), nt, smem, s>>>(gp, ip, dp, gi, gd, B, C, H, W, border, flags);
            return check_launch("warp_bwd");
        }
        if (variant != 0 && W <= 65535 && smem <= 160 * 1024) {
            if (smem > 48 * 1024) {
                cudaError_t e = cudaFuncSetAttribute(warp_bwd_fixed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (e != cudaSuccess) return fail(DSSMD_ERR_CUDA, "warp_bwd: cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(e));
            }
            warp_bwd_fixed_kernel<<<(unsigned)(B * H), 256, smem, s>>>(gp, ip, dp, gi, gd, B, C, H, W, border);
            return check_launch("warp_bwd");
        }
    }
    return warp_bwd_v1<S>(g, img, disp, g_img, g_disp, B, C, H, W, pad, s);
}

}  // namespace

}  // namespace dssmd

using namespace dssmd;

extern "C" size_t dssmd_warp_bwd_workspace_bytes(int B, int C, int H, int W, int dtype) {
    if (dtype != DSSMD_F32 || B <= 0 || C <= 0 || H <= 0 || W <= 0) return 0;
    return (size_t)B * C * H * W * sizeof(float);
}

extern "C" int dssmd_warp_bwd(const void *g, const void *img, const void *disp, void *g_img, void *g_disp,
                              void *workspace, size_t workspace_bytes,
                              int B, int C, int H, int W, int padding_mode, int dtype, void *stream) {
    DSSMD_REQUIRE(g && img && disp, "warp_bwd: null input pointer");
    DSSMD_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, "warp_bwd: sizes must be positive (B=%d C=%d H=%d W=%d)", B, C, H, W);
    DSSMD_REQUIRE(padding_mode == DSSMD_PAD_ZEROS || padding_mode == DSSMD_PAD_BORDER, "warp_bwd: unknown padding_mode %d", padding_mode);
    DSSMD_REQUIRE((long long)B * H < (1LL << 31), "warp_bwd: B*H too large");
    if (!g_img && !g_disp) return DSSMD_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    switch (dtype) {
        case DSSMD_F32: return warp_bwd_typed<float>(g, img, disp, g_img, g_disp, workspace, workspace_bytes, B, C, H, W, padding_mode, s);
        case DSSMD_F64: return warp_bwd_typed<double>(g, img, disp, g_img, g_disp, workspace, workspace_bytes, B, C, H, W, padding_mode, s);
        case DSSMD_BF16:
        case DSSMD_F16:
            return fail(DSSMD_ERR_UNSUPPORTED, "warp_bwd: 16-bit dtypes are not supported (coordinates need fp32; see DESIGN.md)");
        default: return fail(DSSMD_ERR_INVALID_ARGUMENT, "warp_bwd: unknown dtype %d", dtype);
    }
}
