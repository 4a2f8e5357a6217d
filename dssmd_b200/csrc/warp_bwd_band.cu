// warp_bwd_band.cu -- one-pass backward of the disparity warp for image-like fp32 inputs (C <= 4, W % 4 == 0,
// W <= 2048): horizontal scatter, vertical mix and the disparity gradient in ONE persistent kernel, no workspace.
// Upstream: autograd of utils/disparity_warper.py:83-106 (aten::grid_sampler_2d_backward + the coordinate chain).
//
// The two-kernel path (warp_lean.cu K1 + warp_bwd.cu K2) spends a third of its time in K2 and in the barriers of
// its short-lived one-row blocks.  Here a block owns a BAND of R consecutive rows of one image and walks down it:
//   * the vertical taps of an output row are constants of that row and always fall on rows y-1, y, y+1, so
//     g_img[r] = sum over y in {r-1, r, r+1} of wy(y -> r) * gV[y], gV[y] = horizontal scatter of g[y].  The
//     block scatters row y into 32-bit fixed-point shared-memory accumulators exactly as K1 does (same scale per
//     row and channel, hence the same bits), converts the finished row to float in registers and folds it into two
//     pending output rows; rows y = band start - 1 and band end are scattered a second time by the neighbouring
//     bands (halo), which replaces the gV round trip through L2/HBM and the second launch;
//   * everything is double-buffered (accumulators, difference rows, |g| partial sums), so ONE __syncthreads per row
//     separates "row y scattered" from "row y read out": between two barriers a warp reads out row y-1, computes
//     g_disp of row y, scatters row y and stages row y+1 -- independent instruction streams that hide each other's
//     latencies -- while the loads of row y+1 are in flight from the top of the iteration.
// Accumulation order per output element: ascending source row, fma chain from 0 (the order of K2).  Deterministic.
#include "warp_lean.cuh"

#include <type_traits>

namespace dssmd {

namespace {

constexpr int kMaxBandRows = 32;  // owned rows per band (R) upper bound; the table holds R + 2 processed rows

template <int NCT, int PS, bool BORDER>
__global__ void __launch_bounds__(PS - 2, (512 / (PS - 2)) > 0 ? (512 / (PS - 2)) : 1)
warp_bwd_band_kernel(const float *g, const float *img, const float *disp,
                     float *__restrict__ g_img, float *__restrict__ g_disp, int H, int W, int R,
                     float halfW, float tscale, float halfH) {
    constexpr int P = PS + 6;                 // accumulator slots per plane (see warp_bwd_lean_kernel)
    constexpr int ACH = 4 * P * 4 + 128;      // bytes of one accumulator channel: int acc[4][P]; int sink[32]
    constexpr int ABUF = NCT * ACH;           // one accumulator buffer
    constexpr int VW = NCT == 3 ? 4 : NCT;    // floats per difference cell
    constexpr int DBUF = 4 * PS * VW * 4;     // one difference row: float dvr[4][PS][VW] (cell p = k + 4)
    constexpr int NW = (PS - 2) / 32;         // warps per block (upper bound)
    extern __shared__ __align__(16) unsigned char smem_raw[];   // acc[2] | dvr[2]
    __shared__ float s_part[2][NCT][32];
    __shared__ int s_yi0[kMaxBandRows + 2];
    __shared__ float s_yw0[kMaxBandRows + 2], s_yw1[kMaxBandRows + 2];

    const int W4 = W >> 2;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int b = blockIdx.y;
    const int yb0 = blockIdx.x * R;             // owned rows [yb0, y_end)
    const int y_end = min(yb0 + R, H);
    // Processed rows ys .. y_end: the owned rows plus the neighbours whose vertical taps reach an owned row.  The taps of
    // row y fall on rows {y-1, y} in the upper half of the image and on {y, y+1} in the lower half, so a band that does
    // not straddle the middle needs only ONE of its two neighbours; the other is skipped (row above) or treated like
    // the virtual row H that contributes nothing (row below).
    const float hm1 = (float)(H - 1);
    auto reaches = [&](int y, int r) {   // does output row y have a vertical tap on row r?  (same classification as the table)
        int i0;
        float a, b2;
        const float iy_raw = unnorm_f((float)y, hm1, halfH);
        if (BORDER) axis_border(iy_raw, hm1, i0, a, b2);
        else axis_zeros(iy_raw, H, i0, a, b2);
        return ((unsigned)i0 < (unsigned)H && i0 == r) || ((unsigned)(i0 + 1) < (unsigned)H && i0 + 1 == r);
    };
    const int ys = (yb0 > 0 && reaches(yb0 - 1, yb0)) ? yb0 - 1 : yb0;
    const int y_real = (y_end < H && reaches(y_end, y_end - 1)) ? y_end + 1 : y_end;   // rows y < y_real are scattered
    const unsigned HW = (unsigned)(H * W);
    const bool active = tid < W4;
    const int xl = active ? tid * 4 : W - 4;    // idle threads of the last warp re-read the last pixels
    const float wm1 = (float)(W - 1);
    const float rwm1 = div_rn_recip(wm1);
    const uint32_t abase = smem_u32(smem_raw);
    unsigned char *const dvr_base = smem_raw + 2 * ABUF;
    const int nwarps = (int)(blockDim.x >> 5);

    // ---- prologue: vertical taps of the processed rows, zeroed accumulators --------------------------------
    if (tid < y_end - ys + 1) {
        const int y = ys + tid;
        int i0 = -2;
        float w0 = 0.f, w1 = 0.f;
        if (y < H) {
            const float iy_raw = unnorm_f((float)y, hm1, halfH);
            if (BORDER) axis_border(iy_raw, hm1, i0, w0, w1);
            else axis_zeros(iy_raw, H, i0, w0, w1);
        }
        s_yi0[tid] = i0; s_yw0[tid] = w0; s_yw1[tid] = w1;
    }
    if (active) {
        const uint32_t za = 4u * (uint32_t)tid;
#pragma unroll
        for (int bf = 0; bf < 2; ++bf)
#pragma unroll
            for (int c = 0; c < NCT; ++c)
#pragma unroll
                for (int pl = 0; pl < 4; ++pl) stsi_at(smem_raw, za + (uint32_t)(bf * ABUF + c * ACH + pl * P * 4), 0);
    }

    const float *const gb = g + (size_t)b * NCT * HW + xl;
    const float *const db = disp + (size_t)b * HW + xl;
    const float *const ib = img + (size_t)b * NCT * HW + xl;
    const bool more = xl + 4 < W;
    const int nx = more ? 4 : 3;   // pixel x+4 belongs to the next thread; past the row end it is 0

    float go[NCT][4];              // g of the row being scattered
    int k0[4];                     // its left-tap cells and weights
    float w0[4], w1[4], gmult[4];
    unsigned cmask = 0u;           // 'border': bit j = pixel j is clamped onto the left border cell
    float pendA[NCT][4], pendB[NCT][4];   // partial g_img rows y-1 and y while row y is being read out
    float down_prev[NCT];          // fixed point -> float scale of the row to be read out
#pragma unroll
    for (int c = 0; c < NCT; ++c) {
        down_prev[c] = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) { pendA[c][j] = 0.f; pendB[c][j] = 0.f; }
    }

    // ---- stage helpers ------------------------------------------------------------------------------------------
    // per-warp sums of |g| of a row -> s_part[buf]
    auto abs_sums = [&](const float (&gr)[NCT][4], int buf) {
        float v[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int cc = c < NCT ? c : 0;
            const float a = (fabsf(gr[cc][0]) + fabsf(gr[cc][1])) + (fabsf(gr[cc][2]) + fabsf(gr[cc][3]));
            v[c] = (c < NCT && active) ? a : 0.f;
        }
        // transposed butterfly: each step halves the number of values a lane carries, so the 32-lane sums of all
        // channels cost 6 shuffles instead of 5 per channel; channel c ends up in lanes 8c .. 8c+7 (fixed order)
        const bool h16 = lane & 16, h8 = lane & 8;
        float k0v, k1v;
        if (NCT > 2) {
            const float s0 = h16 ? v[0] : v[2], s1 = h16 ? v[1] : v[3];
            k0v = (h16 ? v[2] : v[0]) + __shfl_xor_sync(0xffffffffu, s0, 16);
            k1v = (h16 ? v[3] : v[1]) + __shfl_xor_sync(0xffffffffu, s1, 16);
        } else {
            k0v = v[0] + __shfl_xor_sync(0xffffffffu, v[0], 16);
            k1v = v[1] + __shfl_xor_sync(0xffffffffu, v[1], 16);
        }
        float r;
        if (NCT > 1) r = (h8 ? k1v : k0v) + __shfl_xor_sync(0xffffffffu, h8 ? k0v : k1v, 8);
        else r = k0v + __shfl_xor_sync(0xffffffffu, k0v, 8);
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
        // lanes 8c hold channel c: NCT > 2: c = 2 * bit4 + bit3; NCT == 2: c = bit3 (both halves); NCT == 1: c = 0
        const int c_of = NCT > 2 ? (lane >> 3) : NCT > 1 ? ((lane >> 3) & 1) : 0;
        if ((lane & 7) == 0 && lane < 8 * NCT && c_of < NCT) s_part[buf][c_of][wid] = r;
    };
    // horizontal taps of a row from its disparities
    auto taps = [&](const float4 &d4) {
        const float dd[4] = {d4.x, d4.y, d4.z, d4.w};
        const float xf0 = (float)xl;
        cmask = 0u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float t = __fsub_rn(xf0 + (float)j, dd[j]);
            const float ix = unnorm_f(t, wm1, rwm1, halfW);
            if (BORDER) {
                gmult[j] = (ix > 0.0f && ix < wm1) ? halfW : 0.0f;   // clip_coordinates_set_grad
                axis_border(ix, wm1, k0[j], w0[j], w1[j]);
                if (ix <= 0.0f) cmask |= 1u << j;                    // clamped onto cell 0 (weights 1, 0)
            } else {
                gmult[j] = halfW;
                axis_zeros(ix, W, k0[j], w0[j], w1[j]);
            }
        }
    };
    // blended horizontal differences of the two source rows of output row y -> dvr[buf]
    auto stage_diff = [&](int y, int buf) {
        const int t = y - ys;
        const int yi0 = s_yi0[t];
        const float yw0 = s_yw0[t], yw1 = s_yw1[t];
        const int ya = min(max(yi0, 0), H - 1), yb = min(max(yi0 + 1, 0), H - 1);
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
            unsigned char *const dvr = dvr_base + buf * DBUF;
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
    };
    // read out the finished accumulators of row y (zeroing them), fold the row into the pending g_img rows and
    // emit g_img row y - 1
    auto read_out = [&](int y) {
        const int t = y - ys;
        const bool real = y < y_real;
        const int yi0 = s_yi0[t];
        const float yw0 = s_yw0[t], yw1 = s_yw1[t];
        const bool in0 = (unsigned)yi0 < (unsigned)H, in1 = (unsigned)(yi0 + 1) < (unsigned)H;
        // weight of gV[y] in g_img rows y - 1, y, y + 1 (a zero weight still propagates NaN, as in ATen)
        float wr[3];
        bool on[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int r = y - 1 + k;
            wr[k] = 0.f; on[k] = false;
            if (real && in0 && yi0 == r) { wr[k] = yw0; on[k] = true; }
            if (real && in1 && yi0 + 1 == r) { wr[k] = yw1; on[k] = true; }
        }
        float cur[NCT][4];
        if (real) {
            const uint32_t ra = 4u * (uint32_t)tid + (uint32_t)((y & 1) * ABUF);
#pragma unroll
            for (int c = 0; c < NCT; ++c)
#pragma unroll
                for (int pl = 0; pl < 4; ++pl) {
                    const uint32_t a = ra + (uint32_t)(c * ACH + pl * P * 4);
                    // a non-finite row sum makes down NaN: the whole row / channel is poisoned, as the reference's would be
                    cur[c][pl] = (float)ldsi_at(smem_raw, a) * down_prev[c];
                    stsi_at(smem_raw, a, 0);
                }
        } else {
#pragma unroll
            for (int c = 0; c < NCT; ++c)
#pragma unroll
                for (int pl = 0; pl < 4; ++pl) cur[c][pl] = 0.f;
        }
        const int r_out = y - 1;
        const bool emit = r_out >= yb0 && r_out < y_end;
        float *ob = g_img + (size_t)b * NCT * HW + (unsigned)(r_out * W + xl);
#pragma unroll
        for (int c = 0; c < NCT; ++c) {
            float o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                o[j] = on[0] ? fmaf(wr[0], cur[c][j], pendA[c][j]) : pendA[c][j];
                pendA[c][j] = on[1] ? fmaf(wr[1], cur[c][j], pendB[c][j]) : pendB[c][j];
                pendB[c][j] = on[2] ? wr[2] * cur[c][j] : 0.f;
            }
            if (emit && active) st_global_cs(reinterpret_cast<float4 *>(ob + c * HW), make_float4(o[0], o[1], o[2], o[3]));
        }
    };

    // ---- first row ------------------------------------------------------------------------------------------------
    __syncthreads();   // table visible
    pdl_launch_dependents();
    pdl_wait();        // nothing above touches global memory
    {
        const unsigned orow = (unsigned)(ys * W);
        const float4 d4 = ld_ca(reinterpret_cast<const float4 *>(db + orow));
#pragma unroll
        for (int c = 0; c < NCT; ++c) {
            const float4 v = ld_ca(reinterpret_cast<const float4 *>(gb + c * HW + orow));
            go[c][0] = v.x; go[c][1] = v.y; go[c][2] = v.z; go[c][3] = v.w;
        }
        if (g_disp && ys >= yb0) stage_diff(ys, ys & 1);
        abs_sums(go, ys & 1);
        taps(d4);
    }
    __syncthreads();

    for (int y = ys; y <= y_end; ++y) {
        const bool real = y < y_real;
        const bool nreal = y + 1 < y_real;   // a next row exists and is scattered
        const int buf = y & 1;

        // loads of row y + 1, in flight while rows y - 1 and y are worked on
        float gn[NCT][4];
        float4 dn = make_float4(0.f, 0.f, 0.f, 0.f);
        if (nreal) {
            const unsigned nrow = (unsigned)((y + 1) * W);
            dn = ld_ca(reinterpret_cast<const float4 *>(db + nrow));
#pragma unroll
            for (int c = 0; c < NCT; ++c) {
                const float4 v = ld_ca(reinterpret_cast<const float4 *>(gb + c * HW + nrow));
                gn[c][0] = v.x; gn[c][1] = v.y; gn[c][2] = v.z; gn[c][3] = v.w;
            }
        }

        // the two image rows that g_disp of row y + 1 differentiates: into L1 now, read by stage_diff at the end of
        // the iteration (one of them is usually there already -- consecutive output rows share a source row)
        if (nreal && g_disp && y + 1 >= yb0 && y + 1 < y_end && (lane & 7) == 0) {
            const int yi0n = s_yi0[y + 1 - ys];
            const int ya = min(max(yi0n, 0), H - 1), yb = min(max(yi0n + 1, 0), H - 1);
#pragma unroll
            for (int c = 0; c < NCT; ++c) {
                asm volatile("prefetch.global.L1 [%0];" ::"l"(ib + c * HW + (unsigned)(ya * W)));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(ib + c * HW + (unsigned)(yb * W)));
            }
        }

        if (y > ys) read_out(y - 1);

        float up[NCT];
        if (real) {
            // g_disp of row y (owned rows only)
            if (g_disp && y >= yb0 && y < y_end && active) {
                const unsigned char *const dvr = dvr_base + buf * DBUF;
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
                *reinterpret_cast<float4 *>(g_disp + (size_t)b * HW + (unsigned)(y * W + xl)) = make_float4(gd[0], gd[1], gd[2], gd[3]);
            }

            // S[c] = sum of |g[c]| over the row -> power-of-two scales (see warp_bwd_lean_kernel)
#pragma unroll
            for (int c = 0; c < NCT; ++c) {
                float sc = lane < nwarps ? s_part[buf][c][lane] : 0.f;
#pragma unroll
                for (int o = (NW > 16 ? 16 : NW > 8 ? 8 : NW > 4 ? 4 : NW > 2 ? 2 : 1); o > 0; o >>= 1) sc += __shfl_xor_sync(0xffffffffu, sc, o);
                sc = __shfl_sync(0xffffffffu, sc, 0);
                const int sexp = (int)(__float_as_uint(sc) >> 23);  // sc >= 0 or NaN: no sign bit to mask
                const bool finite = sexp < 0xff;                     // inf / NaN gradients poison the row
                int e = 156 - (sexp > 0 ? sexp : 1);                 // sc < 2^(sexp-126)  =>  sc * 2^e < 2^30
                if (e > 127) e = 127;
                up[c] = (finite && sc > 0.f) ? __uint_as_float((unsigned)(e + 127) << 23) : 0.f;  // 0: nothing to add
                down_prev[c] = !finite ? __uint_as_float(0x7fc00000u)
                                       : (e < 127) ? __uint_as_float((unsigned)(127 - e) << 23) : __uint_as_float(0x00400000u);
            }

            // scatter row y into acc[buf].  Taps that would pile onto one cell -- 'border': every pixel whose coordinate is
            // clamped to the left edge adds to cell 0 (disparities up to 192 px clamp ~10 % of a row; same-address shared
            // atomics of a warp serialise, 32 wavefronts per instruction); 'zeros': zero-weight taps outside the row --
            // go to the lane's private sink word instead, and the clamped pixels' contributions are summed across the
            // warp in integers (exact, order-free) and added to cell 0 once.
            const uint32_t ab = abase + (uint32_t)(buf * ABUF);
            const uint32_t sink = ab + (uint32_t)(4 * P * 4) + 4u * (uint32_t)lane;   // + c * ACH: the sink of channel c
            // most warps have no clamped pixel: warp-uniform choice between the plain scatter and the one that diverts
            auto scatter = [&](auto clamped_path) {
                constexpr bool CL = decltype(clamped_path)::value;
                uint32_t aL[4], aR[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t p = (uint32_t)k0[j], q = (uint32_t)(k0[j] + 1);
                    bool okL, okR;
                    if (BORDER) { okL = okR = CL ? (active && !((cmask >> j) & 1u)) : active; }   // right tap at most cell W: allocated, weight 0
                    else { okL = active && p < (uint32_t)W; okR = active && q < (uint32_t)W; }
                    aL[j] = okL ? ab + p + (p & 3u) * (uint32_t)(4 * P - 1) : sink;
                    aR[j] = okR ? ab + q + (q & 3u) * (uint32_t)(4 * P - 1) : sink;
                }
#pragma unroll
                for (int c = 0; c < NCT; ++c) {
                    int qc = 0;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float gs = go[c][j] * up[c];   // exact: power-of-two scale
                        const int qL = __float2int_rn(gs * w0[j]);
                        red_add(aL[j] + (uint32_t)(c * ACH), qL);
                        red_add(aR[j] + (uint32_t)(c * ACH), __float2int_rn(gs * w1[j]));
                        if (CL) qc += ((cmask >> j) & 1u) ? qL : 0;
                    }
                    if (CL) {
                        const int tot = __reduce_add_sync(0xffffffffu, active ? qc : 0);
                        if (lane == 0) red_add(ab + (uint32_t)(c * ACH), tot);   // cell 0
                    }
                }
            };
            if (BORDER && __any_sync(0xffffffffu, active && cmask != 0u)) scatter(std::true_type{});
            else scatter(std::false_type{});
        }

        // stage row y + 1
        if (nreal) {
            if (g_disp && y + 1 >= yb0 && y + 1 < y_end) stage_diff(y + 1, buf ^ 1);
            abs_sums(gn, buf ^ 1);
            taps(dn);
#pragma unroll
            for (int c = 0; c < NCT; ++c)
#pragma unroll
                for (int j = 0; j < 4; ++j) go[c][j] = gn[c][j];
        }
        __syncthreads();
    }
    read_out(y_end);
}

template <int PS, bool BORDER>
int launch_band(const float *g, const float *img, const float *disp, float *g_img, float *g_disp, int B, int C, int H, int W, cudaStream_t s) {
    void (*kern)(const float *, const float *, const float *, float *, float *, int, int, int, float, float, float) =
        C == 1 ? warp_bwd_band_kernel<1, PS, BORDER> : C == 2 ? warp_bwd_band_kernel<2, PS, BORDER>
      : C == 3 ? warp_bwd_band_kernel<3, PS, BORDER> : warp_bwd_band_kernel<4, PS, BORDER>;
    const int vw = C == 3 ? 4 : C;
    const size_t smem = 2 * (size_t)C * (4 * (PS + 6) * 4 + 128) + 2 * ((size_t)4 * PS * vw * 4);
    cudaError_t e = kernel_smem_optin(kern, smem);
    if (e != cudaSuccess) return fail(DSSMD_ERR_CUDA, "warp_bwd(band): cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(e));
    const int threads = round_up(W / 4, 32);
    // rows per band: one wave of co-resident blocks over the whole batch (halo cost 2 / R rows per band)
    int R = env_int_lean("DSSMD_WARP_BWD_BAND_ROWS", 0);
    if (R <= 0) {
        const int occ = kernel_occupancy(kern, threads, smem);
        // A band costs R + 1 row passes (one halo row) and the grid runs in ceil(blocks / slots) waves: take the R that minimises
        // waves * (R + 1).  (The rule "one wave" picked R = 32 at 384x1280 x16 -- 192 blocks on 148 one-block SMs, a second wave 30 %
        // full: 179 us; R = 24 gives 256 blocks in two full-ish waves of 25 passes.)
        const long long slots = (long long)occ * sm_count();
        long long best = -1;
        for (int r = 4; r <= kMaxBandRows && r <= (H > 4 ? H : 4); ++r) {
            const long long blocks = (long long)ceil_div(H, r) * B;
            const long long cost = ((blocks + slots - 1) / slots) * (r + 1);
            if (best < 0 || cost <= best) { best = cost; R = r; }   // ties: the larger R (fewer halo rows)
        }
        if (R < 4) R = 4;
    }
    if (R > kMaxBandRows) R = kMaxBandRows;
    if (R > H) R = H;
    e = launch_pdl(kern, dim3((unsigned)ceil_div(H, R), (unsigned)B), dim3((unsigned)threads), smem, s, g, img, disp, g_img, g_disp, H, W, R,
                   (float)W / 2.0f, 2.0f / (float)(W - 1), (float)H / 2.0f);
    if (e != cudaSuccess) return fail(DSSMD_ERR_CUDA, "warp_bwd(band): launch failed: %s", cudaGetErrorString(e));
    return check_launch("warp_bwd(band)");
}

}  // namespace

// g_img must be non-null (g_disp may be null).  -1 when the shape is not eligible (the caller falls back).
int warp_bwd_band_f32(const float *g, const float *img, const float *disp, float *g_img, float *g_disp,
                      int B, int C, int H, int W, int border, cudaStream_t s) {
    if (!g_img || !lean_eligible(B, C, H, W) || W > 2048 || env_int_lean("DSSMD_WARP_BWD_BAND", 1) == 0) return -1;
    if (((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(img) | reinterpret_cast<uintptr_t>(disp) |
          reinterpret_cast<uintptr_t>(g_img) | reinterpret_cast<uintptr_t>(g_disp)) & 15) != 0) return -1;
    const int w4 = W / 4;
    if (w4 + 2 <= 66) return border ? launch_band<66, true>(g, img, disp, g_img, g_disp, B, C, H, W, s) : launch_band<66, false>(g, img, disp, g_img, g_disp, B, C, H, W, s);
    if (w4 + 2 <= 130) return border ? launch_band<130, true>(g, img, disp, g_img, g_disp, B, C, H, W, s) : launch_band<130, false>(g, img, disp, g_img, g_disp, B, C, H, W, s);
    if (w4 + 2 <= 258) return border ? launch_band<258, true>(g, img, disp, g_img, g_disp, B, C, H, W, s) : launch_band<258, false>(g, img, disp, g_img, g_disp, B, C, H, W, s);
    return border ? launch_band<514, true>(g, img, disp, g_img, g_disp, B, C, H, W, s) : launch_band<514, false>(g, img, disp, g_img, g_disp, B, C, H, W, s);
}

}  // namespace dssmd
