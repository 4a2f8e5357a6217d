// warp_lean.cu -- instruction-lean fp32 kernels of the disparity warp for image-like inputs
// (C <= 4, W % 4 == 0, 16-byte aligned rows): forward (+ mask + negative-disparity flag) and the
// horizontal-scatter half (K1) of the backward.  Same maths, same op order of the coordinate chain
// (upstream utils/disparity_warper.py:48-106) as warp.cu / warp_bwd.cu.
//
// ncu showed the first generation of these kernels issue-bound at ~300 instructions per pixel,
// most of them overhead: 64-bit address chains, integer divisions hidden in block-strided loops and in
// blockIdx % H, predicate materialisation around every tap, branches around every bounds case.  Here
//   * the grid is (H, B): no index division; all loops have compile-time trip counts;
//   * padding_mode is a template parameter: with 'border' a clamped coordinate never leaves the row, so
//     taps need no predicates at all (the one tap that can fall on x = W has weight 0 and reads padding);
//   * out-of-bounds taps of the mask / 'zeros' sample get weight 0 by one select instead of a branch;
//   * a shared-memory tap address is two instructions: cell p of a planar row (plane p & 3, slot p >> 2,
//     PS slots per plane) lives at byte p + (p & 3) * (4 * PS - 1);
//   * backward: d(sample)/d(ix) is read from ONE pre-blended difference row per channel,
//     dv[k] = wy0 * (r0[k+1] - r0[k]) + wy1 * (r1[k+1] - r1[k]), 3 loads + 3 FMAs per pixel instead of
//     12 loads + 36 individually rounded operations.
// Bit-exactness: the coordinate chain and the mask sum use individually rounded intrinsics in the
// reference order (masks are bit-identical); sample and gradient sums use fused multiply-adds, as the
// ATen kernels do (agreement ~1e-7, tested at 1e-5).
#include "warp_lean.cuh"

namespace dssmd {

namespace {

// ---------------------------------------------------------------------------------
// Forward.  Block = one output row (blockIdx.x = y, blockIdx.y = image), thread = 4 consecutive
// pixels x all channels.  The vertical tap weights are constants of the output row, so the block first
// blends its two source rows, v[c][x] = wy0 * img[c][y0][x] + wy1 * img[c][y0+1][x], and stages that
// ONE row in shared memory -- zero-padded, planar (cell p = x + 4 in plane p & 3, slot p >> 2: the
// lanes of a warp, whose taps are ~4 cells apart, read consecutive slots of one plane), a cell holding
// all channels of a pixel.  A pixel then costs two vector LDS and 2 FMAs per channel instead of
// 4 scalar LDS and 4 FMAs per channel.
// ---------------------------------------------------------------------------------
template <int NCT, int PS, bool BORDER>
__global__ void __launch_bounds__(PS - 2, 1024 / (PS - 2)) warp_fwd_lean_kernel(const float *img, const float *disp,
                                                                                float *__restrict__ out, float *__restrict__ mask,
                                                                                int *neg_flag, int H, int W, float halfW, float halfH, int pf) {
    constexpr int VW = NCT == 3 ? 4 : NCT;
    extern __shared__ __align__(16) float s_row[];  // [4][PS] cells of VW floats
    const int W4 = W >> 2;
    const int y = blockIdx.x, b = blockIdx.y;
    const int tid = threadIdx.x;
    const unsigned HW = (unsigned)(H * W);  // C*H*W < 2^31 checked on the host
    const bool active = tid < W4;
    const int x0 = tid * 4;
    const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
    // (x / (W-1) stays __fdiv_rn here: its call-out keeps ptxas from sinking the image loads below the coordinate
    // arithmetic, which costs more than the division saves -- measured 18.5 -> 21.0 us at 576x960 with div_rn_by)

    // vertical taps: un-clamped with zeros padding for the mask, clamped ('border') for the sample
    const float iy_raw = unnorm_f((float)y, hm1, halfH);
    int ymi0, yi0;
    float ymw0, ymw1, yw0, yw1;
    axis_zeros(iy_raw, H, ymi0, ymw0, ymw1);
    if (BORDER) axis_border(iy_raw, hm1, yi0, yw0, yw1);
    else { yi0 = ymi0; yw0 = ymw0; yw1 = ymw1; }

    pdl_launch_dependents();
    const int xl = active ? x0 : W - 4;
    const int ya = min(max(yi0, 0), H - 1), yb = min(max(yi0 + 1, 0), H - 1);
    if (pf > 0 && active && (tid & 7) == 0) {
        // this row's lines towards the L2 while the previous kernel drains (common.cuh: prefetch_l2); one thread per 128 bytes
        const float *ib = img + (size_t)b * NCT * HW + xl;
        prefetch_l2(disp + (size_t)b * HW + (unsigned)(y * W + xl));
#pragma unroll
        for (int c = 0; c < NCT; ++c) {
            prefetch_l2(ib + c * HW + (unsigned)(ya * W));
            prefetch_l2(ib + c * HW + (unsigned)(yb * W));
        }
    }
    pdl_wait();   // first load below

    // source rows -> registers (in flight while the coordinates are computed).  Loads are unconditional (a
    // predicated load costs a branch): a tap row outside the image is read from the nearest row and has
    // weight 0, idle threads of the last warp re-read the last pixels.
    float4 st[NCT][2];
    {
        const float *ib = img + (size_t)b * NCT * HW + xl;
#pragma unroll
        for (int c = 0; c < NCT; ++c) {
            st[c][0] = ld_ca(reinterpret_cast<const float4 *>(ib + c * HW + (unsigned)(ya * W)));
            st[c][1] = ld_ca(reinterpret_cast<const float4 *>(ib + c * HW + (unsigned)(yb * W)));
        }
    }
    const unsigned orow = (unsigned)(y * W + x0);
    const float4 d4 = ld_ca(reinterpret_cast<const float4 *>(disp + (size_t)b * HW + (unsigned)(y * W + xl)));
    const float dv[4] = {d4.x, d4.y, d4.z, d4.w};

    uint32_t a0[4], a1[4];  // byte offset of the low / high x tap
    float w0[4], w1[4];     // their weights
    float mv[4];
    bool bad = false;
    const float xf0 = (float)x0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        bad = bad || !(dv[i] >= 0.0f);
        const float t = __fsub_rn(xf0 + (float)i, dv[i]);   // x0 + i is exact
        const float ix = unnorm_f(t, wm1, halfW);
        int mi0, i0;
        float mw0, mw1;
        axis_zeros(ix, W, mi0, mw0, mw1);
        // sum of the in-bounds weights of the un-clamped sample, reference order nw, ne, sw, se
        float m = __fmul_rn(ymw0, mw0);
        m = __fadd_rn(m, __fmul_rn(ymw0, mw1));
        m = __fadd_rn(m, __fmul_rn(ymw1, mw0));
        m = __fadd_rn(m, __fmul_rn(ymw1, mw1));
        mv[i] = (m >= 0.9999f) ? 1.0f : 0.0f;
        if (BORDER) axis_border(ix, wm1, i0, w0[i], w1[i]);
        else { i0 = mi0; w0[i] = mw0; w1[i] = mw1; }
        const uint32_t p = (uint32_t)(i0 + 4);   // i0 in [-2, W]: cells 2 .. W + 5
        a0[i] = vcell_off<PS, VW>(p);
        a1[i] = vcell_off<PS, VW>(p + 1u);
    }
    if (active && bad && neg_flag) *reinterpret_cast<volatile int *>(neg_flag) = 1;  // plain store: may be mapped host memory

    if (active) {
        const uint32_t sa = (uint32_t)(VW * 4) * ((uint32_t)tid + 1u);  // slot tid + 1 of plane 0
#pragma unroll
        for (int j = 0; j < 4; ++j) {   // cell 4*tid + 4 + j: plane j
            float cell[VW];
#pragma unroll
            for (int c = 0; c < VW; ++c)
                cell[c] = c < NCT ? fmaf(yw1, f4c(st[c < NCT ? c : 0][1], j), yw0 * f4c(st[c < NCT ? c : 0][0], j)) : 0.f;
            sts_cell<VW>(s_row, sa + (uint32_t)(j * PS * VW * 4), cell);
        }
    }
    if (tid < 8) {  // the 4 + 4 pad cells
        float z[VW];
#pragma unroll
        for (int c = 0; c < VW; ++c) z[c] = 0.f;
        sts_cell<VW>(s_row, (uint32_t)(((tid & 3) * PS + ((tid >> 2) ? W4 + 1 : 0)) * VW * 4), z);
    }
    __syncthreads();
    if (!active) return;

    float v[NCT][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float t0[VW], t1[VW];
        lds_cell<VW>(s_row, a0[i], t0);
        lds_cell<VW>(s_row, a1[i], t1);
#pragma unroll
        for (int c = 0; c < NCT; ++c) v[c][i] = fmaf(t1[c], w1[i], t0[c] * w0[i]);
    }
    float *ob = out + (size_t)b * NCT * HW + orow;
    float *mb = mask ? mask + (size_t)b * NCT * HW + orow : nullptr;
#pragma unroll
    for (int c = 0; c < NCT; ++c) {
        st_global_cs(reinterpret_cast<float4 *>(ob + c * HW), make_float4(v[c][0], v[c][1], v[c][2], v[c][3]));
        if (mb) st_global_cs(reinterpret_cast<float4 *>(mb + c * HW), make_float4(mv[0], mv[1], mv[2], mv[3]));
    }
}

// ---------------------------------------------------------------------------------
// Backward K1 (same contract as warp_bwd_hs4_kernel): per output row,
//   gV[c][y][k] = sum_x g[c][y][x] * wx(x -> k)     (horizontal scatter, vertical mix done by K2)
//   g_disp[y][x] = -(2/(W-1)) * gmult * sum_c g[c][y][x] * d(sample_c)/d(ix)
// Scatter: a thread owns 4 consecutive pixels; every tap is one fire-and-forget 32-bit shared-memory
// add of q = rn(g * w * 2^e), with e chosen per (row, channel) from S = sum_x |g| so that
// S * 2^e <= 2^30: no cell can overflow (a cell's sum is bounded by S), integer adds are associative,
// the result is deterministic.  The adds are unconditional -- ptxas turns a predicated shared atomic
// into a branch + reconvergence (4 instructions) -- so a tap outside the row adds 0 (its weight is 0)
// to a clamped cell, and the idle threads of the last warp add to a per-lane sink word.
// ---------------------------------------------------------------------------------

template <int NCT, int PS, bool BORDER>
__global__ void __launch_bounds__(PS - 2, 1024 / (PS - 2)) warp_bwd_lean_kernel(const float *g, const float *img,
                                                                                const float *disp, float *__restrict__ gV,
                                                                                float *__restrict__ g_disp, int H, int W,
                                                                                float halfW, float tscale, float halfH, int dbg) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int P = PS + 6;          // accumulator slots per plane: >= W/4 + 1, and = 8 (mod 32) so that the four
                                       // planes start 8 banks apart (lanes whose taps straddle planes do not collide)
    constexpr int ACH = 4 * P * 4 + 128;   // bytes of one accumulator channel: int acc[4][P]; int sink[32]
    constexpr int VW = NCT == 3 ? 4 : NCT;   // floats per difference cell (all channels of a pixel)
    // layout: { int acc[4][P]; int sink[32]; }[NCT]; float dvr[4][PS][VW] (cell p = k + 4)
    __shared__ float s_part[NCT][32];  // per-warp, per-channel sums of |g|

    const int W4 = W >> 2;
    const int y = blockIdx.x, b = blockIdx.y;
    const unsigned HW = (unsigned)(H * W);
    const int tid = threadIdx.x, lane = tid & 31;
    const bool active = tid < W4;
    const int x0 = tid * 4;
    const int xl = active ? x0 : W - 4;   // idle threads of the last warp re-read the last pixels (loads stay unconditional)
    const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
    const float rwm1 = div_rn_recip(wm1);
    const uint32_t abase = smem_u32(smem_raw);      // accumulators (for the RED instructions)
    unsigned char *const dvr = smem_raw + NCT * ACH;  // difference rows
    const unsigned orow = (unsigned)(y * W + xl);

    // output-row loads first
    const float4 d4 = ld_ca(reinterpret_cast<const float4 *>(disp + (size_t)b * HW + orow));
    float go[NCT][4];
    {
        const float *gb = g + (size_t)b * NCT * HW + orow;
#pragma unroll
        for (int c = 0; c < NCT; ++c) {
            const float4 v = ld_ca(reinterpret_cast<const float4 *>(gb + c * HW));
            go[c][0] = v.x; go[c][1] = v.y; go[c][2] = v.z; go[c][3] = v.w;
        }
    }

    if (g_disp) {
        // vertical taps of this output row, then the blended horizontal differences of the two source rows
        // (a tap row outside the image is read from the nearest row: its weight is 0)
        const float iy_raw = unnorm_f((float)y, hm1, halfH);
        int yi0;
        float yw0, yw1;
        if (BORDER) axis_border(iy_raw, hm1, yi0, yw0, yw1);
        else axis_zeros(iy_raw, H, yi0, yw0, yw1);
        const int ya = min(max(yi0, 0), H - 1), yb = min(max(yi0 + 1, 0), H - 1);
        const float *ib = img + (size_t)b * NCT * HW + xl;
        const bool more = xl + 4 < W;
        const int nx = more ? 4 : 3;   // pixel x+4 belongs to the next thread; past the row end it is 0
        float4 r0[NCT], r1[NCT];
        float n0[NCT], n1[NCT];
#pragma unroll
        for (int c = 0; c < NCT; ++c) {
            r0[c] = ld_ca(reinterpret_cast<const float4 *>(ib + c * HW + (unsigned)(ya * W)));
            r1[c] = ld_ca(reinterpret_cast<const float4 *>(ib + c * HW + (unsigned)(yb * W)));
            n0[c] = ld_ca(ib + c * HW + (unsigned)(ya * W + nx));
            n1[c] = ld_ca(ib + c * HW + (unsigned)(yb * W + nx));
        }
        if (active) {
            const uint32_t sa = (uint32_t)(VW * 4) * ((uint32_t)tid + 1u);  // cells 4*tid+4 .. +7: slot tid + 1 of planes 0..3
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float cell[VW];
#pragma unroll
                for (int c = 0; c < VW; ++c) {
                    const int cc = c < NCT ? c : 0;
                    const float a0v = f4c(r0[cc], j), a1v = f4c(r1[cc], j);
                    const float b0v = j < 3 ? f4c(r0[cc], j + 1) : (more ? n0[cc] : 0.f);
                    const float b1v = j < 3 ? f4c(r1[cc], j + 1) : (more ? n1[cc] : 0.f);
                    cell[c] = c < NCT ? fmaf(yw1, b1v - a1v, yw0 * (b0v - a0v)) : 0.f;
                }
                sts_cell<VW>(dvr, sa + (uint32_t)(j * PS * VW * 4), cell);
            }
            if (tid == 0) {
                // cell 3 = k -1 ('zeros': left tap outside, right tap on pixel 0); cells 0..2 are never differentiated
                float z[VW], e[VW];
#pragma unroll
                for (int c = 0; c < VW; ++c) { z[c] = 0.f; e[c] = c < NCT ? fmaf(yw1, r1[c < NCT ? c : 0].x, yw0 * r0[c < NCT ? c : 0].x) : 0.f; }
                sts_cell<VW>(dvr, 0u * PS * VW * 4, z); sts_cell<VW>(dvr, 1u * PS * VW * 4, z);
                sts_cell<VW>(dvr, 2u * PS * VW * 4, z); sts_cell<VW>(dvr, 3u * PS * VW * 4, e);
            }
            if (tid == W4 - 1) {  // cells W + 4 .. W + 7 (k = W .. W + 3): both taps outside
                float z[VW];
#pragma unroll
                for (int c = 0; c < VW; ++c) z[c] = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) sts_cell<VW>(dvr, sa + (uint32_t)(VW * 4 + j * PS * VW * 4), z);
            }
        }
    }
    if (gV) {
        if (active) {  // this thread's 4 cells per channel (slot tid of every plane)
            const uint32_t za = 4u * (uint32_t)tid;
#pragma unroll
            for (int c = 0; c < NCT; ++c)
#pragma unroll
                for (int pl = 0; pl < 4; ++pl) stsi_at(smem_raw, za + (uint32_t)(c * ACH + pl * P * 4), 0);
        }
#pragma unroll
        for (int c = 0; c < NCT; ++c) {
            float asum = (fabsf(go[c][0]) + fabsf(go[c][1])) + (fabsf(go[c][2]) + fabsf(go[c][3]));
            asum = active ? asum : 0.f;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) asum += __shfl_xor_sync(0xffffffffu, asum, o);
            if (lane == 0) s_part[c][tid >> 5] = asum;
        }
    }

    // horizontal taps
    const float dd[4] = {d4.x, d4.y, d4.z, d4.w};
    int k0[4];   // cell of the left tap ('zeros': in [-2, W])
    float w0[4], w1[4], gmult[4];
    const float xf0 = (float)xl;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float t = __fsub_rn(xf0 + (float)j, dd[j]);
        const float ix = unnorm_f(t, wm1, rwm1, halfW);
        if (BORDER) {
            // clip_coordinates_set_grad: a coordinate on or beyond a border has zero gradient
            gmult[j] = (ix > 0.0f && ix < wm1) ? halfW : 0.0f;
            axis_border(ix, wm1, k0[j], w0[j], w1[j]);
        } else {
            gmult[j] = halfW;
            axis_zeros(ix, W, k0[j], w0[j], w1[j]);
        }
    }

    __syncthreads();  // accumulators zeroed, difference rows staged, per-warp |g| sums written

    if (g_disp && active) {
        float gd[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float dvc[VW];
            lds_cell<VW>(dvr, vcell_off<PS, VW>((uint32_t)(k0[j] + 4)), dvc);  // k0 in [-2, W]
            float gix = 0.f;
#pragma unroll
            for (int c = 0; c < NCT; ++c) gix = fmaf(go[c][j], dvc[c], gix);
            // chain rule: ix <- gx (W/2, 0 where clipped) <- t (2/(W-1)) <- disp (-1)
            gd[j] = -(tscale * (gmult[j] * gix));
        }
        *reinterpret_cast<float4 *>(g_disp + (size_t)b * HW + orow) = make_float4(gd[0], gd[1], gd[2], gd[3]);
    }
    if (!gV) return;  // only g_disp requested (block-uniform)

    // S[c] = sum of |g[c]| over the row; every warp reduces the per-warp partials with the same butterfly
    float up[NCT], down[NCT];
#pragma unroll
    for (int c = 0; c < NCT; ++c) {
        float sc = lane < (int)(blockDim.x >> 5) ? s_part[c][lane] : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sc += __shfl_xor_sync(0xffffffffu, sc, o);
        const int sexp = (int)(__float_as_uint(sc) >> 23);  // sc >= 0 or NaN: no sign bit to mask
        const bool finite = sexp < 0xff;                     // inf / NaN gradients poison the row
        int e = 156 - (sexp > 0 ? sexp : 1);                 // sc < 2^(sexp-126)  =>  sc * 2^e < 2^30
        if (e > 127) e = 127;
        up[c] = (finite && sc > 0.f) ? __uint_as_float((unsigned)(e + 127) << 23) : 0.f;  // 0: nothing to add
        down[c] = !finite ? __uint_as_float(0x7fc00000u)
                          : (e < 127) ? __uint_as_float((unsigned)(127 - e) << 23) : __uint_as_float(0x00400000u);
    }

    {
        // cell addresses inside channel 0.  'border': the left tap is always inside, the right tap is at most cell W
        // (plane 0, one slot past the row: inside the allocation) and then has weight 0.
        const uint32_t sink = abase + (uint32_t)(4 * P * 4) + 4u * (uint32_t)lane;   // + c * ACH: the sink of channel c
        uint32_t aL[4], aR[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int kl = k0[j], kr = k0[j] + 1;
            if (!BORDER) { kl = min(max(kl, 0), W); kr = min(max(kr, 0), W); }   // weight 0 outside [0, W)
            const uint32_t p = (uint32_t)kl, q = (uint32_t)kr;
            aL[j] = active ? abase + p + (p & 3u) * (uint32_t)(4 * P - 1) : sink;
            aR[j] = active ? abase + q + (q & 3u) * (uint32_t)(4 * P - 1) : sink;
        }
        if (!(dbg & 1))
#pragma unroll
        for (int c = 0; c < NCT; ++c) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                // scaling by a power of two is exact, so this is rn(g * w) * 2^e rounded to an integer
                const float gs = go[c][j] * up[c];
                red_add(aL[j] + (uint32_t)(c * ACH), __float2int_rn(gs * w0[j]));
                red_add(aR[j] + (uint32_t)(c * ACH), __float2int_rn(gs * w1[j]));
            }
        }
    }
    __syncthreads();
    if (!active) return;
    float *ob = gV + (size_t)b * NCT * HW + orow;
    const uint32_t ra = 4u * (uint32_t)tid;
#pragma unroll
    for (int c = 0; c < NCT; ++c) {
        float v[4];
#pragma unroll
        for (int pl = 0; pl < 4; ++pl) v[pl] = (float)ldsi_at(smem_raw, ra + (uint32_t)(c * ACH + pl * P * 4));
        // a non-finite row sum makes down[c] NaN: the whole row / channel is poisoned, as the reference's would be
        *reinterpret_cast<float4 *>(ob + c * HW) = make_float4(v[0] * down[c], v[1] * down[c], v[2] * down[c], v[3] * down[c]);
    }
}

// ---------------------------------------------------------------------------------
// Fused L1 photometric term on the warp residual (built from upstream's
// LRwarp_error, utils/disparity_warper.py:109-115, with its 'border' warp):
//   err = imgR - warp(imgL, disp);   partial[b*H + y] = sum_{c,x} |err|;
//   g_unit[b,0,y,x] = d(sum |err|)/d(disp) = (2/(W-1)) * gmult * sum_c sign(err_c) * d(sample_c)/d(ix)
// in ONE pass: the forward kernel's two taps t0, t1 of the pre-blended row already give
// d(sample)/d(ix) = t1 - t0, so the gradient w.r.t. the disparity is a by-product of the forward --
// the residual is never written, no backward kernel and no scatter are needed (images carry no
// gradient here).  Row sums are reduced in a fixed order: deterministic.
// ---------------------------------------------------------------------------------
template <int NCT, int PS>
__global__ void __launch_bounds__(PS - 2, 1024 / (PS - 2)) photo_l1_lean_kernel(const float *imgL, const float *imgR,
                                                                                const float *disp, float *__restrict__ g_unit,
                                                                                float *__restrict__ partial, int *neg_flag, int H, int W,
                                                                                float halfW, float tscale, float halfH) {
    constexpr int VW = NCT == 3 ? 4 : NCT;
    extern __shared__ __align__(16) float s_row[];  // [4][PS] cells of VW floats
    __shared__ float s_sum[32];
    const int W4 = W >> 2;
    const int y = blockIdx.x, b = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31;
    const unsigned HW = (unsigned)(H * W);
    const bool active = tid < W4;
    const int x0 = tid * 4;
    const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);

    const float iy_raw = unnorm_f((float)y, hm1, halfH);
    int yi0;
    float yw0, yw1;
    axis_border(iy_raw, hm1, yi0, yw0, yw1);

    const int xl = active ? x0 : W - 4;
    const int ya = min(max(yi0, 0), H - 1), yb = min(max(yi0 + 1, 0), H - 1);
    float4 st[NCT][2], rr[NCT];
    {
        const float *ib = imgL + (size_t)b * NCT * HW + xl;
        const float *rb = imgR + (size_t)b * NCT * HW + (unsigned)(y * W + xl);
#pragma unroll
        for (int c = 0; c < NCT; ++c) {
            st[c][0] = ld_ca(reinterpret_cast<const float4 *>(ib + c * HW + (unsigned)(ya * W)));
            st[c][1] = ld_ca(reinterpret_cast<const float4 *>(ib + c * HW + (unsigned)(yb * W)));
            rr[c] = ld_ca(reinterpret_cast<const float4 *>(rb + c * HW));
        }
    }
    const float4 d4 = ld_ca(reinterpret_cast<const float4 *>(disp + (size_t)b * HW + (unsigned)(y * W + xl)));
    const float dv[4] = {d4.x, d4.y, d4.z, d4.w};

    uint32_t a0[4], a1[4];
    float w0[4], w1[4], gmult[4];
    bool bad = false;
    const float xf0 = (float)xl;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        bad = bad || !(dv[i] >= 0.0f);
        const float t = __fsub_rn(xf0 + (float)i, dv[i]);
        const float ix = unnorm_f(t, wm1, halfW);
        gmult[i] = (ix > 0.0f && ix < wm1) ? halfW : 0.0f;   // clip_coordinates_set_grad
        int i0;
        axis_border(ix, wm1, i0, w0[i], w1[i]);
        const uint32_t p = (uint32_t)(i0 + 4);
        a0[i] = vcell_off<PS, VW>(p);
        a1[i] = vcell_off<PS, VW>(p + 1u);
    }
    if (active && bad && neg_flag) *reinterpret_cast<volatile int *>(neg_flag) = 1;

    if (active) {
        const uint32_t sa = (uint32_t)(VW * 4) * ((uint32_t)tid + 1u);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float cell[VW];
#pragma unroll
            for (int c = 0; c < VW; ++c)
                cell[c] = c < NCT ? fmaf(yw1, f4c(st[c < NCT ? c : 0][1], j), yw0 * f4c(st[c < NCT ? c : 0][0], j)) : 0.f;
            sts_cell<VW>(s_row, sa + (uint32_t)(j * PS * VW * 4), cell);
        }
    }
    if (tid < 8) {
        float z[VW];
#pragma unroll
        for (int c = 0; c < VW; ++c) z[c] = 0.f;
        sts_cell<VW>(s_row, (uint32_t)(((tid & 3) * PS + ((tid >> 2) ? W4 + 1 : 0)) * VW * 4), z);
    }
    __syncthreads();

    float asum = 0.f;
    float gd[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float t0[VW], t1[VW];
        lds_cell<VW>(s_row, a0[i], t0);
        lds_cell<VW>(s_row, a1[i], t1);
        float gix = 0.f;
#pragma unroll
        for (int c = 0; c < NCT; ++c) {
            const float smp = fmaf(t1[c], w1[i], t0[c] * w0[i]);
            const float err = f4c(rr[c], i) - smp;
            asum += fabsf(err);
            const float sg = err > 0.f ? 1.f : (err < 0.f ? -1.f : 0.f);
            gix = fmaf(sg, t1[c] - t0[c], gix);
        }
        gd[i] = tscale * (gmult[i] * gix);
    }
    if (active && g_unit)
        *reinterpret_cast<float4 *>(g_unit + (size_t)b * HW + (unsigned)(y * W + x0)) = make_float4(gd[0], gd[1], gd[2], gd[3]);
    asum = active ? asum : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) asum += __shfl_xor_sync(0xffffffffu, asum, o);
    if (lane == 0) s_sum[tid >> 5] = asum;
    __syncthreads();
    if (tid == 0) {
        float tot = 0.f;
        const int nw = (int)(blockDim.x >> 5);
        for (int k = 0; k < nw; ++k) tot += s_sum[k];
        partial[(size_t)b * H + y] = tot;
    }
}

template <int PS>
int launch_photo(const float *imgL, const float *imgR, const float *disp, float *g_unit, float *partial, int *neg_flag,
                 int B, int C, int H, int W, cudaStream_t s) {
    void (*kern)(const float *, const float *, const float *, float *, float *, int *, int, int, float, float, float) =
        C == 1 ? photo_l1_lean_kernel<1, PS> : C == 2 ? photo_l1_lean_kernel<2, PS>
      : C == 3 ? photo_l1_lean_kernel<3, PS> : photo_l1_lean_kernel<4, PS>;
    const size_t smem = (size_t)(C == 3 ? 4 : C) * 4 * PS * sizeof(float);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(DSSMD_ERR_CUDA, "photometric_l1: cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(e));
    }
    kern<<<dim3((unsigned)H, (unsigned)B), round_up(W / 4, 32), smem, s>>>(imgL, imgR, disp, g_unit, partial, neg_flag, H, W,
                                                                           (float)W / 2.0f, 2.0f / (float)(W - 1), (float)H / 2.0f);
    return check_launch("photometric_l1");
}

template <int PS, bool BORDER>
int launch_fwd(const float *img, const float *disp, float *out, float *mask, int *neg_flag, int B, int C, int H, int W, cudaStream_t s) {
    void (*kern)(const float *, const float *, float *, float *, int *, int, int, float, float, int) =
        C == 1 ? warp_fwd_lean_kernel<1, PS, BORDER> : C == 2 ? warp_fwd_lean_kernel<2, PS, BORDER>
      : C == 3 ? warp_fwd_lean_kernel<3, PS, BORDER> : warp_fwd_lean_kernel<4, PS, BORDER>;
    const size_t smem = (size_t)(C == 3 ? 4 : C) * 4 * PS * sizeof(float);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(DSSMD_ERR_CUDA, "warp_fwd: cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(e));
    }
    // W/2 and H/2: IEEE single divisions, the same values the device would compute
    const cudaError_t le = launch_pdl(kern, dim3((unsigned)H, (unsigned)B), dim3((unsigned)round_up(W / 4, 32)), smem, s, img, disp, out, mask, neg_flag, H, W,
                                      (float)W / 2.0f, (float)H / 2.0f, pdl_prefetch(1));
    if (le != cudaSuccess) return fail(DSSMD_ERR_CUDA, "warp_fwd(lean): launch failed: %s", cudaGetErrorString(le));
    return check_launch("warp_fwd(lean)");
}

template <int PS, bool BORDER>
int launch_bwd(const float *g, const float *img, const float *disp, float *gV, float *g_disp, int B, int C, int H, int W, cudaStream_t s) {
    void (*kern)(const float *, const float *, const float *, float *, float *, int, int, float, float, float, int) =
        C == 1 ? warp_bwd_lean_kernel<1, PS, BORDER> : C == 2 ? warp_bwd_lean_kernel<2, PS, BORDER>
      : C == 3 ? warp_bwd_lean_kernel<3, PS, BORDER> : warp_bwd_lean_kernel<4, PS, BORDER>;
    const size_t smem = ((size_t)C * (4 * (PS + 6) + 32) + (size_t)(C == 3 ? 4 : C) * 4 * PS) * sizeof(int);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(DSSMD_ERR_CUDA, "warp_bwd: cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(e));
    }
    kern<<<dim3((unsigned)H, (unsigned)B), round_up(W / 4, 32), smem, s>>>(g, img, disp, gV, g_disp, H, W, (float)W / 2.0f,
                                                                           2.0f / (float)(W - 1), (float)H / 2.0f, env_int_lean("DSSMD_WARP_DBG", 0));
    return check_launch("warp_bwd(lean scatter)");
}

bool eligible(int B, int C, int H, int W) { return lean_eligible(B, C, H, W); }

}  // namespace

#define DSSMD_LEAN_PICK(FN, ...)                                                       \
    do {                                                                               \
        const int w4 = W / 4;                                                          \
        if (w4 + 2 <= 66) return border ? FN<66, true>(__VA_ARGS__) : FN<66, false>(__VA_ARGS__);       \
        if (w4 + 2 <= 130) return border ? FN<130, true>(__VA_ARGS__) : FN<130, false>(__VA_ARGS__);    \
        if (w4 + 2 <= 258) return border ? FN<258, true>(__VA_ARGS__) : FN<258, false>(__VA_ARGS__);    \
        if (w4 + 2 <= 514) return border ? FN<514, true>(__VA_ARGS__) : FN<514, false>(__VA_ARGS__);    \
        return border ? FN<1026, true>(__VA_ARGS__) : FN<1026, false>(__VA_ARGS__);                     \
    } while (0)

// Both return -1 when the shape is not eligible (the caller falls back to the generic kernels).
int warp_fwd_lean_f32(const float *img, const float *disp, float *out, float *mask, int *neg_flag,
                      int B, int C, int H, int W, int border, cudaStream_t s) {
    if (!eligible(B, C, H, W) || env_int_lean("DSSMD_WARP_LEAN", 1) == 0) return -1;
    if (((reinterpret_cast<uintptr_t>(img) | reinterpret_cast<uintptr_t>(disp) | reinterpret_cast<uintptr_t>(out) |
          reinterpret_cast<uintptr_t>(mask)) & 15) != 0) return -1;
    if (env_int_lean("DSSMD_CORR_DEBUG", 0)) fprintf(stderr, "warp_fwd lean cfg: C=%d H=%d W=%d border=%d\n", C, H, W, border);
    DSSMD_LEAN_PICK(launch_fwd, img, disp, out, mask, neg_flag, B, C, H, W, s);
}

// K1 of the backward: gV (may be null) and g_disp (may be null)
int warp_bwd_lean_f32(const float *g, const float *img, const float *disp, float *gV, float *g_disp,
                      int B, int C, int H, int W, int border, cudaStream_t s) {
    if (!eligible(B, C, H, W) || env_int_lean("DSSMD_WARP_LEAN", 1) == 0) return -1;
    if (((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(img) | reinterpret_cast<uintptr_t>(disp) |
          reinterpret_cast<uintptr_t>(gV) | reinterpret_cast<uintptr_t>(g_disp)) & 15) != 0) return -1;
    DSSMD_LEAN_PICK(launch_bwd, g, img, disp, gV, g_disp, B, C, H, W, s);
}

// fused L1 photometric row sums + d/d(disp); -1 when the shape is not eligible
int photometric_l1_lean_f32(const float *imgL, const float *imgR, const float *disp, float *g_unit, float *partial, int *neg_flag,
                            int B, int C, int H, int W, cudaStream_t s) {
    if (!eligible(B, C, H, W)) return -1;
    if (((reinterpret_cast<uintptr_t>(imgL) | reinterpret_cast<uintptr_t>(imgR) | reinterpret_cast<uintptr_t>(disp) |
          reinterpret_cast<uintptr_t>(g_unit)) & 15) != 0) return -1;
    const int w4 = W / 4;
    if (w4 + 2 <= 66) return launch_photo<66>(imgL, imgR, disp, g_unit, partial, neg_flag, B, C, H, W, s);
    if (w4 + 2 <= 130) return launch_photo<130>(imgL, imgR, disp, g_unit, partial, neg_flag, B, C, H, W, s);
    if (w4 + 2 <= 258) return launch_photo<258>(imgL, imgR, disp, g_unit, partial, neg_flag, B, C, H, W, s);
    if (w4 + 2 <= 514) return launch_photo<514>(imgL, imgR, disp, g_unit, partial, neg_flag, B, C, H, W, s);
    return launch_photo<1026>(imgL, imgR, disp, g_unit, partial, neg_flag, B, C, H, W, s);
}

#undef DSSMD_LEAN_PICK

}  // namespace dssmd
