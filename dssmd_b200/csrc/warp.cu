// warp.cu -- disparity-guided bilinear warp, forward and backward.
//
// Replaces disp_warp (upstream utils/disparity_warper.py:83-106, identical copy at
// models/UtilsNet/stereo_op.py:41-63) and the backward autograd derives for it.
// The coordinate arithmetic reproduces the reference's op order exactly
// (SURVEY.md Appendix A.2): the reference normalises with (W-1)/(H-1) but calls
// F.grid_sample with align_corners=False, so
//     t  = x - disp;  gx = 2*(t/(W-1)) - 1;  ix = fma(gx+1, W/2, -0.5)
//     gy = 2*(y/(H-1)) - 1;                   iy = fma(gy+1, H/2, -0.5)
// is a 2-D, 4-tap gather whose vertical blend is constant per row.
// Algorithmic HBM bytes (C image channels, s bytes/elt): forward
// (C img + 1 disp + C out + C mask)*s per pixel, backward (C g + C img + 1 disp +
// C g_img + 1 g_disp)*s per pixel.
#include "warp_common.cuh"

#include <cstdlib>

namespace dssmd {

// warp_lean.cu: instruction-lean kernels for image-like shapes; -1 = not eligible
int warp_fwd_lean_f32(const float *img, const float *disp, float *out, float *mask, int *neg_flag,
                      int B, int C, int H, int W, int border, cudaStream_t s);

namespace {

int env_int(const char *name, int dflt) {
    if (!tuning_enabled()) return dflt;
    const char *v = std::getenv(name);
    return (v && *v) ? std::atoi(v) : dflt;
}

// ---------------------------------------------------------------------------------
// Forward.  One block per image row; a thread owns 4 consecutive pixels and all C
// channels: the coordinate work is done once per pixel (the reference redoes it per
// channel and per grid_sample call), the per-row constants (several IEEE divisions)
// once per block.  The mask and the negative-disparity flag are fused in.
//   NCT > 0: exactly NCT channels (unrolled); NCT == 0: any C.
// ---------------------------------------------------------------------------------
template <typename S, int NCT>
__global__ void __launch_bounds__(512) warp_fwd_kernel(const S *__restrict__ img, const S *__restrict__ disp,
                                                        S *__restrict__ out, S *__restrict__ mask, int *neg_flag,
                                                        int B, int C, int H, int W, int border, int vec) {
    using N = Num<S>;
    __shared__ S s_c[6];    // W/2, ym.w0, ym.w1, yi.w0, yi.w1
    __shared__ int s_i[4];  // ym.i0, ym flags, yi.i0, yi flags
    const int y = blockIdx.x % H;
    const int b = blockIdx.x / H;
    const long long HW = (long long)H * W;
    const int tid = threadIdx.x;
    if (tid == 0) {
        const S iy_raw = unnorm<S>((S)y, H);
        const Axis<S> m = make_axis<S>(iy_raw, H);                                         // mask: un-clamped
        const Axis<S> a = make_axis<S>(border ? clamp_border<S>(iy_raw, H) : iy_raw, H);   // image sample
        s_c[0] = N::div((S)W, (S)2);
        s_c[1] = m.w0; s_c[2] = m.w1; s_c[3] = a.w0; s_c[4] = a.w1;
        s_i[0] = m.i0; s_i[1] = (m.in0 ? 1 : 0) | (m.in1 ? 2 : 0);
        s_i[2] = a.i0; s_i[3] = (a.in0 ? 1 : 0) | (a.in1 ? 2 : 0);
    }
    __syncthreads();
    const S halfW = s_c[0], wm1 = (S)(W - 1);
    const S mw0 = s_c[1], mw1 = s_c[2], yw0 = s_c[3], yw1 = s_c[4];
    const bool min0 = s_i[1] & 1, min1 = s_i[1] & 2, yin0 = s_i[3] & 1, yin1 = s_i[3] & 2;
    const int yi0 = s_i[2];
    const int nc = NCT > 0 ? NCT : C;
    bool bad = false;

    const S *drow = disp + (long long)b * HW + (long long)y * W;
    const S *ibase = img + (long long)b * C * HW + (long long)yi0 * W;
    S *obase = out + (long long)b * C * HW + (long long)y * W;
    S *mbase = mask ? mask + (long long)b * C * HW + (long long)y * W : nullptr;

    for (int x0 = tid * 4; x0 < W; x0 += blockDim.x * 4) {
        S dv[4];
        if (vec) {
            if constexpr (sizeof(S) == 4) {
                const float4 t = *reinterpret_cast<const float4 *>(drow + x0);
                dv[0] = t.x; dv[1] = t.y; dv[2] = t.z; dv[3] = t.w;
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) dv[i] = drow[x0 + i];
            }
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) dv[i] = (x0 + i < W) ? drow[x0 + i] : (S)0;
        }
        int xi0[4], tin[4];             // low x tap; in-bounds bits of taps nw, ne, sw, se
        S wa[4], wb[4], wc[4], wd[4];  // tap weights nw, ne, sw, se
        S mv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (!(dv[i] >= (S)0)) bad = true;
            // reference op order: t = x - d; gx = 2*(t/(W-1)) - 1; ix = fma(gx+1, W/2, -0.5)
            const S t = N::sub((S)(x0 + i), dv[i]);
            const S gxn = N::sub(N::mul((S)2, N::div(t, wm1)), (S)1);
            const S ix_raw = N::fma(N::add(gxn, (S)1), halfW, (S)-0.5);
            const Axis<S> xm = make_axis<S>(ix_raw, W);
            // sum of in-bounds weights of the un-clamped sample, reference order nw, ne, sw, se
            S m = (S)0;
            if (xm.in0 && min0) m = N::add(m, N::mul(mw0, xm.w0));
            if (xm.in1 && min0) m = N::add(m, N::mul(mw0, xm.w1));
            if (xm.in0 && min1) m = N::add(m, N::mul(mw1, xm.w0));
            if (xm.in1 && min1) m = N::add(m, N::mul(mw1, xm.w1));
            mv[i] = (m >= (S)0.9999) ? (S)1 : (S)0;
            Axis<S> xa = xm;
            if (border && !(ix_raw >= (S)0 && ix_raw <= wm1)) xa = make_axis<S>(clamp_border<S>(ix_raw, W), W);
            xi0[i] = xa.i0;
            tin[i] = ((xa.in0 && yin0) ? 1 : 0) | ((xa.in1 && yin0) ? 2 : 0) | ((xa.in0 && yin1) ? 4 : 0) | ((xa.in1 && yin1) ? 8 : 0);
            wa[i] = N::mul(yw0, xa.w0);
            wb[i] = N::mul(yw0, xa.w1);
            wc[i] = N::mul(yw1, xa.w0);
            wd[i] = N::mul(yw1, xa.w1);
        }
        auto channel = [&](int cc) {
            {
                const S *p0 = ibase + (long long)cc * HW;
                S v[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    S acc = (S)0;
                    // only in-bounds taps are read (grid_sample zero-pads the rest)
                    if (tin[i] & 1) acc = N::fma(p0[xi0[i]], wa[i], acc);
                    if (tin[i] & 2) acc = N::fma(p0[xi0[i] + 1], wb[i], acc);
                    if (tin[i] & 4) acc = N::fma(p0[W + xi0[i]], wc[i], acc);
                    if (tin[i] & 8) acc = N::fma(p0[W + xi0[i] + 1], wd[i], acc);
                    v[i] = acc;
                }
                S *o = obase + (long long)cc * HW + x0;
                if (vec && sizeof(S) == 4) {
                    st_global_cs(reinterpret_cast<float4 *>(o), make_float4((float)v[0], (float)v[1], (float)v[2], (float)v[3]));
                    if (mbase)
                        st_global_cs(reinterpret_cast<float4 *>(mbase + (long long)cc * HW + x0),
                                     make_float4((float)mv[0], (float)mv[1], (float)mv[2], (float)mv[3]));
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (x0 + i < W) {
                            o[i] = v[i];
                            if (mbase) mbase[(long long)cc * HW + x0 + i] = mv[i];
                        }
                }
            }
        };
        if constexpr (NCT > 0) {
#pragma unroll
            for (int cc = 0; cc < NCT; ++cc) channel(cc);
        } else {
            for (int cc = 0; cc < nc; ++cc) channel(cc);
        }
    }
    if (bad && neg_flag) *reinterpret_cast<volatile int *>(neg_flag) = 1;  // plain store: the flag may live in mapped host memory
}

// ---------------------------------------------------------------------------------
// Backward.  One block per (b, image row yy).  g_img[b,:,yy,:] only receives
// contributions from output rows y whose vertical taps {y0, y0+1} include yy
// (y0 in {y-1, y}), so the block accumulates that one row for a chunk of channels
// in shared memory and writes it once: no global atomics, every g_img element is
// written exactly once.  g_disp is a gather and is produced by the block with
// yy == y while it walks its own output row.
// ---------------------------------------------------------------------------------
template <typename S>
__global__ void __launch_bounds__(256) warp_bwd_kernel(const S *__restrict__ g, const S *__restrict__ img,
                                                        const S *__restrict__ disp, S *__restrict__ g_img,
                                                        S *__restrict__ g_disp,
                                                        int B, int C, int H, int W, int border, int CC) {
    using N = Num<S>;
    extern __shared__ unsigned char smem_raw[];
    S *acc = reinterpret_cast<S *>(smem_raw);  // [CC][W]

    const int yy = blockIdx.x % H;
    const int b = blockIdx.x / H;
    const long long HW = (long long)H * W;
    const S xscale = N::div((S)W, (S)2);       // d ix / d gx
    const S tscale = N::div((S)2, (S)(W - 1)); // d gx / d t

    for (int c0 = 0; c0 < C; c0 += CC) {
        const int nc = (C - c0 < CC) ? C - c0 : CC;
        if (g_img) {
            for (int i = threadIdx.x; i < nc * W; i += blockDim.x) acc[i] = (S)0;
        }
        __syncthreads();

        const int ylo = yy - 2 < 0 ? 0 : yy - 2;
        const int yhi = yy + 2 > H - 1 ? H - 1 : yy + 2;
        for (int y = ylo; y <= yhi; ++y) {
            const S iy_raw = unnorm<S>((S)y, H);
            const Axis<S> ya = make_axis<S>(border ? clamp_border<S>(iy_raw, H) : iy_raw, H);
            const bool top = ya.in0 && ya.i0 == yy;        // tap row y0   == yy, weight w0
            const bool bot = ya.in1 && ya.i0 + 1 == yy;    // tap row y0+1 == yy, weight w1
            const bool own = (y == yy) && (g_disp != nullptr);
            if (!((g_img && (top || bot)) || own)) continue;
            const S wy = top ? ya.w0 : ya.w1;

            for (int x = threadIdx.x; x < W; x += blockDim.x) {
                const S dv = disp[(long long)b * HW + (long long)y * W + x];
                S ix = unnorm<S>(N::sub((S)x, dv), W);
                S gmult = xscale;
                if (border) {
                    // clip_coordinates_set_grad: borders count as out of bounds for the gradient
                    if (ix <= (S)0) { ix = (S)0; gmult = (S)0; }
                    else if (ix >= (S)(W - 1)) { ix = (S)(W - 1); gmult = (S)0; }
                }
                const Axis<S> xa = make_axis<S>(ix, W);
                S gix = (S)0;
                for (int c = 0; c < nc; ++c) {
                    const long long plane = ((long long)b * C + c0 + c) * HW;
                    const S go = g[plane + (long long)y * W + x];
                    if (g_img && (top || bot)) {
                        S *row = acc + c * W;
                        if (xa.in0) atomicAdd(row + xa.i0, N::mul(go, N::mul(wy, xa.w0)));
                        if (xa.in1) atomicAdd(row + xa.i0 + 1, N::mul(go, N::mul(wy, xa.w1)));
                    }
                    if (own) {
                        const S *p0 = img + plane + (long long)ya.i0 * W;
                        const S *p1 = p0 + W;
                        // d(sample)/d(ix) over in-bounds taps, ATen order nw, ne, sw, se
                        if (xa.in0 && ya.in0) gix = N::sub(gix, N::mul(N::mul(p0[xa.i0], ya.w0), go));
                        if (xa.in1 && ya.in0) gix = N::add(gix, N::mul(N::mul(p0[xa.i0 + 1], ya.w0), go));
                        if (xa.in0 && ya.in1) gix = N::sub(gix, N::mul(N::mul(p1[xa.i0], ya.w1), go));
                        if (xa.in1 && ya.in1) gix = N::add(gix, N::mul(N::mul(p1[xa.i0 + 1], ya.w1), go));
                    }
                }
                if (own) {
                    // chain rule: ix <- gx (W/2, 0 where clipped) <- t (2/(W-1)) <- disp (-1)
                    const S gd = -N::mul(tscale, N::mul(gmult, gix));
                    S *o = g_disp + (long long)b * HW + (long long)y * W + x;
                    *o = (c0 == 0) ? gd : N::add(*o, gd);
                }
            }
        }
        __syncthreads();
        if (g_img) {
            for (int i = threadIdx.x; i < nc * W; i += blockDim.x) {
                const int c = i / W, x = i - c * W;
                g_img[((long long)b * C + c0 + c) * HW + (long long)yy * W + x] = acc[i];
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------
// Bilinear resize, align_corners=False (F.interpolate as called by LRwarp_error,
// upstream utils/disparity_warper.py:110-113).  Index math follows
// ATen/native/UpSample.h area_pixel_compute_source_index / guard_index_and_lambda.
// ---------------------------------------------------------------------------------
template <typename S>
__global__ void __launch_bounds__(256) resize_fwd_kernel(const S *__restrict__ in, S *__restrict__ out,
                                                          int BC, int Hi, int Wi, int Ho, int Wo) {
    const long long n = (long long)BC * Ho * Wo;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % Wo);
        const int y = (int)((i / Wo) % Ho);
        const long long pc = i / ((long long)Wo * Ho);
        const Src<S> ys = src_index<S>(y, Hi, Ho), xs = src_index<S>(x, Wi, Wo);
        out[i] = resize_at<S>(in + pc * Hi * Wi, Hi, Wi, ys, xs);
    }
}

// Backward of the resize as a gather: each source pixel scans the (few) output
// pixels whose two taps can reach it.  For a down-scale by s >= 1 an output pixel
// o reads sources floor(s*(o+.5)-.5) and +1, so source k is touched only by
// outputs in [ (k-1+.5)/s - .5 , (k+1+.5)/s - .5 ]  -> at most a handful.
template <typename S>
__global__ void __launch_bounds__(256) resize_bwd_kernel(const S *__restrict__ g_out, S *__restrict__ g_in,
                                                          int BC, int Hi, int Wi, int Ho, int Wo) {
    using N = Num<S>;
    const long long n = (long long)BC * Hi * Wi;
    const double sy = (double)Hi / Ho, sx = (double)Wi / Wo;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const int kx = (int)(i % Wi);
        const int ky = (int)((i / Wi) % Hi);
        const long long pc = i / ((long long)Wi * Hi);
        int oy0 = (int)::floor((ky - 0.5) / sy - 0.5) - 1, oy1 = (int)::ceil((ky + 1.5) / sy - 0.5) + 1;
        int ox0 = (int)::floor((kx - 0.5) / sx - 0.5) - 1, ox1 = (int)::ceil((kx + 1.5) / sx - 0.5) + 1;
        if (oy0 < 0) oy0 = 0;
        if (ox0 < 0) ox0 = 0;
        if (oy1 > Ho - 1) oy1 = Ho - 1;
        if (ox1 > Wo - 1) ox1 = Wo - 1;
        const S *gp = g_out + pc * Ho * Wo;
        S acc = (S)0;
        for (int oy = oy0; oy <= oy1; ++oy) {
            const Src<S> ys = src_index<S>(oy, Hi, Ho);
            S wy = (S)0;
            if (ys.i0 == ky) wy = N::add(wy, ys.l0);
            if (ys.i1 == ky) wy = N::add(wy, ys.l1);
            if (wy == (S)0) continue;
            for (int ox = ox0; ox <= ox1; ++ox) {
                const Src<S> xs = src_index<S>(ox, Wi, Wo);
                S wx = (S)0;
                if (xs.i0 == kx) wx = N::add(wx, xs.l0);
                if (xs.i1 == kx) wx = N::add(wx, xs.l1);
                if (wx == (S)0) continue;
                acc = N::fma(N::mul(wy, wx), gp[(long long)oy * Wo + ox], acc);
            }
        }
        g_in[i] = acc;
    }
}

// ---------------------------------------------------------------------------------
// Fused LRwarp_error forward: out = resize(imgR) - warp(resize(imgL), disp), with
// the resize folded into the tap fetch (each warped tap interpolates its 4 source
// pixels on the fly; nothing is materialised).  One thread per output pixel.
// ---------------------------------------------------------------------------------
template <typename S>
__global__ void __launch_bounds__(256) lrwarp_fwd_kernel(const S *__restrict__ imgL, const S *__restrict__ disp,
                                                          const S *__restrict__ imgR, S *__restrict__ out,
                                                          int *neg_flag, int B, int C, int H0, int W0, int H, int W) {
    using N = Num<S>;
    const long long n = (long long)B * H * W;
    const long long HW = (long long)H * W, HW0 = (long long)H0 * W0;
    bool bad = false;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % W);
        const int y = (int)((i / W) % H);
        const int b = (int)(i / HW);
        const S dv = disp[i];
        if (!(dv >= (S)0)) bad = true;
        const Axis<S> ya = make_axis<S>(clamp_border<S>(unnorm<S>((S)y, H), H), H);
        const Axis<S> xa = make_axis<S>(clamp_border<S>(unnorm<S>(N::sub((S)x, dv), W), W), W);
        // source taps of the (up to) 2x2 warped taps and of the imgR pixel
        const Src<S> sy0 = src_index<S>(ya.in0 ? ya.i0 : 0, H0, H), sy1 = src_index<S>(ya.in1 ? ya.i0 + 1 : 0, H0, H);
        const Src<S> sx0 = src_index<S>(xa.in0 ? xa.i0 : 0, W0, W), sx1 = src_index<S>(xa.in1 ? xa.i0 + 1 : 0, W0, W);
        const Src<S> ry = src_index<S>(y, H0, H), rx = src_index<S>(x, W0, W);
        for (int c = 0; c < C; ++c) {
            const S *pl = imgL + ((long long)b * C + c) * HW0;
            const S *pr = imgR + ((long long)b * C + c) * HW0;
            S acc = (S)0;
            if (xa.in0 && ya.in0) acc = N::fma(resize_at<S>(pl, H0, W0, sy0, sx0), N::mul(ya.w0, xa.w0), acc);
            if (xa.in1 && ya.in0) acc = N::fma(resize_at<S>(pl, H0, W0, sy0, sx1), N::mul(ya.w0, xa.w1), acc);
            if (xa.in0 && ya.in1) acc = N::fma(resize_at<S>(pl, H0, W0, sy1, sx0), N::mul(ya.w1, xa.w0), acc);
            if (xa.in1 && ya.in1) acc = N::fma(resize_at<S>(pl, H0, W0, sy1, sx1), N::mul(ya.w1, xa.w1), acc);
            out[((long long)b * C + c) * HW + (long long)y * W + x] = N::sub(resize_at<S>(pr, H0, W0, ry, rx), acc);
        }
    }
    if (bad && neg_flag) *reinterpret_cast<volatile int *>(neg_flag) = 1;  // plain store: the flag may live in mapped host memory
}

// ---------------------------------------------------------------------------------
// Forward, staged (fp32, C <= 4, W % 4 == 0, 16-byte aligned rows).  The row kernel
// above spends most of its instructions on 64-bit tap addresses and bounds predicates;
// here the block first copies the two source rows its output row samples (all channels,
// coalesced float4 loads issued before the coordinate arithmetic) into shared memory,
// zero-padded on both sides and zero-filled when a row is out of bounds, so a tap is an
// unpredicated LDS: an out-of-bounds tap reads 0 and fma(0, w, acc) == acc.
// Row layout: position = x + 4 (4 zero cells left, 4 right), position p lives in plane
// p & 3, slot p >> 2 -- a thread's taps are ~4 positions from its neighbour lane's, so
// the lanes of a warp read one plane at consecutive slots (no bank conflicts).
// ---------------------------------------------------------------------------------
template <int NCT, int PS>  // PS: slots per plane, compile-time (>= W/4 + 2): row / channel offsets become immediates
__global__ void __launch_bounds__(PS - 2, 1024 / (PS - 2)) warp_fwd_staged_kernel(const float *__restrict__ img, const float *__restrict__ disp,
                                                               float *__restrict__ out, float *__restrict__ mask, int *neg_flag,
                                                               int H, int W, int border, float halfW, float halfH) {
    using S = float;
    using N = Num<S>;
    extern __shared__ __align__(16) float s_rows[];  // [NCT][2][4][PS]
    const int W4 = W >> 2;
    const int y = blockIdx.x % H;
    const int b = blockIdx.x / H;
    const int HW = H * W;  // C*H*W < 2^31 checked on the host
    const int tid = threadIdx.x;
    const bool active = tid < W4;
    const int x0 = tid * 4;
    const S wm1 = (S)(W - 1);

    // vertical taps: un-clamped for the mask, clamped (border) for the image sample
    const S gy = N::sub(N::mul((S)2, N::div((S)y, (S)(H - 1))), (S)1);
    const S iy_raw = N::fma(N::add(gy, (S)1), halfH, (S)-0.5);
    const Axis<S> ym = make_axis<S>(iy_raw, H);
    const Axis<S> ya = make_axis<S>(border ? clamp_border<S>(iy_raw, H) : iy_raw, H);

    // source rows -> registers (in flight while the coordinates are computed)
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 st[NCT][2];
    {
        const float *ib = img + (size_t)b * NCT * HW + x0;
#pragma unroll
        for (int c = 0; c < NCT; ++c) {
            st[c][0] = (active && ya.in0) ? *reinterpret_cast<const float4 *>(ib + c * HW + ya.i0 * W) : z4;
            st[c][1] = (active && ya.in1) ? *reinterpret_cast<const float4 *>(ib + c * HW + (ya.i0 + 1) * W) : z4;
        }
    }
    const float4 d4 = active ? *reinterpret_cast<const float4 *>(disp + (size_t)b * HW + y * W + x0) : z4;
    const S dv[4] = {d4.x, d4.y, d4.z, d4.w};

    const float *t0[4], *t1[4];     // low / high x tap inside the first staged row
    S wa[4], wb[4], wc[4], wd[4];  // tap weights nw, ne, sw, se (0 where the tap is out of bounds)
    S mv[4];
    bool bad = false;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (!(dv[i] >= (S)0)) bad = true;
        // reference op order: t = x - d; gx = 2*(t/(W-1)) - 1; ix = fma(gx+1, W/2, -0.5)
        const S t = N::sub((S)(x0 + i), dv[i]);
        const S gxn = N::sub(N::mul((S)2, N::div(t, wm1)), (S)1);
        const S ix_raw = N::fma(N::add(gxn, (S)1), halfW, (S)-0.5);
        const Axis<S> xm = make_axis<S>(ix_raw, W);
        // sum of in-bounds weights of the un-clamped sample, reference order nw, ne, sw, se
        S m = (S)0;
        if (xm.in0 && ym.in0) m = N::add(m, N::mul(ym.w0, xm.w0));
        if (xm.in1 && ym.in0) m = N::add(m, N::mul(ym.w0, xm.w1));
        if (xm.in0 && ym.in1) m = N::add(m, N::mul(ym.w1, xm.w0));
        if (xm.in1 && ym.in1) m = N::add(m, N::mul(ym.w1, xm.w1));
        mv[i] = (m >= (S)0.9999) ? (S)1 : (S)0;
        Axis<S> xa = xm;
        if (border && !(ix_raw >= (S)0 && ix_raw <= wm1)) xa = make_axis<S>(clamp_border<S>(ix_raw, W), W);
        const S xw0 = xa.in0 ? xa.w0 : (S)0, xw1 = xa.in1 ? xa.w1 : (S)0;
        wa[i] = N::mul(ya.w0, xw0);
        wb[i] = N::mul(ya.w0, xw1);
        wc[i] = N::mul(ya.w1, xw0);
        wd[i] = N::mul(ya.w1, xw1);
        const int p0 = xa.i0 + 4, p1 = xa.i0 + 5;  // make_axis keeps i0 in [-2, W]
        t0[i] = s_rows + (p0 & 3) * PS + (p0 >> 2);
        t1[i] = s_rows + (p1 & 3) * PS + (p1 >> 2);
    }
    if (active && bad && neg_flag) *reinterpret_cast<volatile int *>(neg_flag) = 1;  // plain store: may be mapped host memory

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
    for (int i = tid; i < NCT * 2 * 8; i += blockDim.x) {  // the 4 + 4 pad cells of every staged row
        const int rowi = i >> 3, k = i & 7;
        s_rows[rowi * 4 * PS + (k & 3) * PS + ((k >> 2) ? W4 + 1 : 0)] = 0.f;
    }
    __syncthreads();
    if (!active) return;

#pragma unroll
    for (int c = 0; c < NCT; ++c) {
        constexpr int R = 4 * PS;  // one staged row
        float v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            S acc = (S)0;
            acc = N::fma(t0[i][(2 * c) * R], wa[i], acc);
            acc = N::fma(t1[i][(2 * c) * R], wb[i], acc);
            acc = N::fma(t0[i][(2 * c + 1) * R], wc[i], acc);
            acc = N::fma(t1[i][(2 * c + 1) * R], wd[i], acc);
            v[i] = acc;
        }
        const size_t o = ((size_t)b * NCT + c) * HW + y * W + x0;
        st_global_cs(reinterpret_cast<float4 *>(out + o), make_float4(v[0], v[1], v[2], v[3]));
        if (mask) st_global_cs(reinterpret_cast<float4 *>(mask + o), make_float4(mv[0], mv[1], mv[2], mv[3]));
    }
}

template <typename S>
int warp_fwd_typed(const void *img, const void *disp, void *warped, void *mask, int *neg_flag,
                   int B, int C, int H, int W, int pad, cudaStream_t s) {
    const uintptr_t bits = reinterpret_cast<uintptr_t>(img) | reinterpret_cast<uintptr_t>(disp) |
                           reinterpret_cast<uintptr_t>(warped) | reinterpret_cast<uintptr_t>(mask);
    const int vec = (W % 4 == 0) && ((bits & 15) == 0);
    int nt = round_up(ceil_div(W, 4), 32);
    if (nt > 512) nt = 512;
    const int border = pad == DSSMD_PAD_BORDER;
    const unsigned grid = (unsigned)(B * H);
    const S *ip = (const S *)img, *dp = (const S *)disp;
    S *op = (S *)warped, *mp = (S *)mask;
    if constexpr (sizeof(S) == 4) {
        if (env_int("DSSMD_WARP_FWD_VARIANT", 1) >= 1) {
            const int rc = warp_fwd_lean_f32(ip, dp, op, mp, neg_flag, B, C, H, W, border, s);
            if (rc >= 0) return rc;
        }
        if (vec && C <= 4 && W <= 4096 && H >= 2 && (long long)C * H * W < (1LL << 31) && env_int("DSSMD_WARP_FWD_VARIANT", 1) >= 1) {
            void (*kern)(const float *, const float *, float *, float *, int *, int, int, int, float, float) = nullptr;
            int ps = 0;
#define DSSMD_FWD_PICK(PSV)                                                                                        \
    if (!kern && W / 4 + 2 <= PSV) {                                                                               \
        ps = PSV;                                                                                                  \
        kern = C == 1 ? warp_fwd_staged_kernel<1, PSV> : C == 2 ? warp_fwd_staged_kernel<2, PSV>                   \
             : C == 3 ? warp_fwd_staged_kernel<3, PSV> : warp_fwd_staged_kernel<4, PSV>;                           \
    }
            DSSMD_FWD_PICK(66) DSSMD_FWD_PICK(130) DSSMD_FWD_PICK(258) DSSMD_FWD_PICK(514) DSSMD_FWD_PICK(1026)
#undef DSSMD_FWD_PICK
            const size_t smem = (size_t)C * 2 * 4 * ps * sizeof(float);
            if (smem > 48 * 1024) {
                cudaError_t e = kernel_smem_optin(kern, smem);
                if (e != cudaSuccess) return fail(DSSMD_ERR_CUDA, "warp_fwd: cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(e));
            }
            // W/2 and H/2: IEEE single divisions, the same values the device would compute
            kern<<<grid, round_up(W / 4, 32), smem, s>>>(ip, dp, op, mp, neg_flag, H, W, border, (float)W / 2.0f, (float)H / 2.0f);
            return check_launch("warp_fwd(staged)");
        }
    }
    switch (C) {
        case 1: warp_fwd_kernel<S, 1><<<grid, nt, 0, s>>>(ip, dp, op, mp, neg_flag, B, C, H, W, border, vec); break;
        case 2: warp_fwd_kernel<S, 2><<<grid, nt, 0, s>>>(ip, dp, op, mp, neg_flag, B, C, H, W, border, vec); break;
        case 3: warp_fwd_kernel<S, 3><<<grid, nt, 0, s>>>(ip, dp, op, mp, neg_flag, B, C, H, W, border, vec); break;
        case 4: warp_fwd_kernel<S, 4><<<grid, nt, 0, s>>>(ip, dp, op, mp, neg_flag, B, C, H, W, border, vec); break;
        default: warp_fwd_kernel<S, 0><<<grid, nt, 0, s>>>(ip, dp, op, mp, neg_flag, B, C, H, W, border, vec); break;
    }
    return check_launch("warp_fwd");
}

}  // namespace

// v1 backward (shared-memory CAS accumulation); kept as the path for C > 4 and for A/B runs
template <typename S>
int warp_bwd_v1(const void *g, const void *img, const void *disp, void *g_img, void *g_disp,
                int B, int C, int H, int W, int pad, cudaStream_t s) {
    // channels per shared-memory pass: at most 48 KB of accumulators
    int cc = (int)((48 * 1024) / ((size_t)W * sizeof(S)));
    if (cc < 1)
        return fail(DSSMD_ERR_UNSUPPORTED, "warp_bwd: a row of W=%d does not fit the shared-memory accumulator", W);
    if (cc > C) cc = C;
    const size_t smem = (size_t)cc * W * sizeof(S);
    warp_bwd_kernel<S><<<(unsigned)(B * H), 256, smem, s>>>((const S *)g, (const S *)img, (const S *)disp,
                                                          (S *)g_img, (S *)g_disp, B, C, H, W,
                                                          pad == DSSMD_PAD_BORDER, cc);
    return check_launch("warp_bwd");
}
template int warp_bwd_v1<float>(const void *, const void *, const void *, void *, void *, int, int, int, int, int, cudaStream_t);
template int warp_bwd_v1<double>(const void *, const void *, const void *, void *, void *, int, int, int, int, int, cudaStream_t);

}  // namespace dssmd

using namespace dssmd;

#define DSSMD_CHECK_WARP_ARGS(name)                                                                       \
    DSSMD_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, name ": sizes must be positive (B=%d C=%d H=%d W=%d)", B, C, H, W); \
    DSSMD_REQUIRE(padding_mode == DSSMD_PAD_ZEROS || padding_mode == DSSMD_PAD_BORDER, name ": unknown padding_mode %d", padding_mode); \
    DSSMD_REQUIRE((long long)B * H < (1LL << 31), name ": B*H too large");

extern "C" int dssmd_warp_fwd(const void *img, const void *disp, void *warped, void *mask, int *neg_flag,
                              int B, int C, int H, int W, int padding_mode, int dtype, void *stream) {
    DSSMD_REQUIRE(img && disp && warped, "warp_fwd: null pointer");
    DSSMD_CHECK_WARP_ARGS("warp_fwd")
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    switch (dtype) {
        case DSSMD_F32: return warp_fwd_typed<float>(img, disp, warped, mask, neg_flag, B, C, H, W, padding_mode, s);
        case DSSMD_F64: return warp_fwd_typed<double>(img, disp, warped, mask, neg_flag, B, C, H, W, padding_mode, s);
        case DSSMD_BF16:
        case DSSMD_F16:
            return fail(DSSMD_ERR_UNSUPPORTED, "warp_fwd: 16-bit dtypes are not supported (coordinates need fp32; see DESIGN.md)");
        default: return fail(DSSMD_ERR_INVALID_ARGUMENT, "warp_fwd: unknown dtype %d", dtype);
    }
}

extern "C" int dssmd_resize_bilinear_fwd(const void *in, void *out, int B, int C, int Hi, int Wi, int Ho, int Wo,
                                         int dtype, void *stream) {
    DSSMD_REQUIRE(in && out, "resize_bilinear_fwd: null pointer");
    DSSMD_REQUIRE(B > 0 && C > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0, "resize_bilinear_fwd: sizes must be positive");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const long long n = (long long)B * C * Ho * Wo;
    if (dtype == DSSMD_F32)
        resize_fwd_kernel<float><<<grid_for(n, 256, 8), 256, 0, s>>>((const float *)in, (float *)out, B * C, Hi, Wi, Ho, Wo);
    else if (dtype == DSSMD_F64)
        resize_fwd_kernel<double><<<grid_for(n, 256, 8), 256, 0, s>>>((const double *)in, (double *)out, B * C, Hi, Wi, Ho, Wo);
    else
        return fail(DSSMD_ERR_UNSUPPORTED, "resize_bilinear_fwd: dtype %d not supported", dtype);
    return check_launch("resize_bilinear_fwd");
}

extern "C" int dssmd_resize_bilinear_bwd(const void *g_out, void *g_in, int B, int C, int Hi, int Wi, int Ho, int Wo,
                                         int dtype, void *stream) {
    DSSMD_REQUIRE(g_out && g_in, "resize_bilinear_bwd: null pointer");
    DSSMD_REQUIRE(B > 0 && C > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0, "resize_bilinear_bwd: sizes must be positive");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const long long n = (long long)B * C * Hi * Wi;
    if (dtype == DSSMD_F32)
        resize_bwd_kernel<float><<<grid_for(n, 256, 8), 256, 0, s>>>((const float *)g_out, (float *)g_in, B * C, Hi, Wi, Ho, Wo);
    else if (dtype == DSSMD_F64)
        resize_bwd_kernel<double><<<grid_for(n, 256, 8), 256, 0, s>>>((const double *)g_out, (double *)g_in, B * C, Hi, Wi, Ho, Wo);
    else
        return fail(DSSMD_ERR_UNSUPPORTED, "resize_bilinear_bwd: dtype %d not supported", dtype);
    return check_launch("resize_bilinear_bwd");
}

extern "C" int dssmd_lrwarp_error_fwd(const void *imgL, const void *disp, const void *imgR, void *out, int *neg_flag,
                                      int B, int C, int H0, int W0, int H, int W, int dtype, void *stream) {
    DSSMD_REQUIRE(imgL && disp && imgR && out, "lrwarp_error_fwd: null pointer");
    DSSMD_REQUIRE(B > 0 && C > 0 && H0 > 0 && W0 > 0 && H > 0 && W > 0, "lrwarp_error_fwd: sizes must be positive");
    DSSMD_REQUIRE(W0 > W || (W0 == W && H0 == H),
                  "lrwarp_error_fwd: images no wider than disp must already have disp's size (got %dx%d vs %dx%d)", H0, W0, H, W);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const long long n = (long long)B * H * W;
    if (dtype == DSSMD_F32)
        lrwarp_fwd_kernel<float><<<grid_for(n, 256, 8), 256, 0, s>>>((const float *)imgL, (const float *)disp, (const float *)imgR,
                                                                   (float *)out, neg_flag, B, C, H0, W0, H, W);
    else if (dtype == DSSMD_F64)
        lrwarp_fwd_kernel<double><<<grid_for(n, 256, 8), 256, 0, s>>>((const double *)imgL, (const double *)disp, (const double *)imgR,
                                                                    (double *)out, neg_flag, B, C, H0, W0, H, W);
    else
        return fail(DSSMD_ERR_UNSUPPORTED, "lrwarp_error_fwd: dtype %d not supported", dtype);
    return check_launch("lrwarp_error_fwd");
}
