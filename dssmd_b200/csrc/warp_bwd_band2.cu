// warp_bwd_band2.cu -- the one-pass band backward of the disparity warp (see warp_bwd_band.cu for the scheme: a block owns
// a band of rows of one image, scatters row y horizontally into 32-bit fixed-point shared-memory accumulators, folds the
// finished row into two pending g_img rows, one barrier per row), rewritten around its instruction count.  ncu of the first
// band kernel (profiles/r01h_summary.md): issue-active 54 % at 16 warps per SM, ~990 instructions per thread and row (4 pixels x
// 3 channels) -- the kernel is bound by what it issues, not by HBM (22 % of peak) or shared memory.  What changed:
//   * ONE fixed-point scale per row (not one per channel), taken from the per-warp sums of |g| that every thread adds up with
//     two LDS.128 in a fixed order (was: a 5-step shuffle reduction per channel in every thread);
//   * tap weights pre-multiplied by the scale, read-out weights pre-multiplied by its inverse and tabulated per row in the
//     prologue (no per-row classification of the vertical taps, no selects in the read-out);
//   * accumulator cells offset by 4 like the difference cells: every tap of both padding modes has an allocated cell, so a
//     tap that must not contribute carries weight 0 instead of a redirected address; idle lanes of the last warp scatter with
//     scale 0 instead of predicated addresses;
//   * a row's fixed-point quantum is 2^-30 of the row's sum of |g| over all channels -- an ABSOLUTE error scale, fine when the
//     row's elements are of comparable size.  Rows where that would hurt take a second, exact path: rows in which one
//     thread's 4 pixels hold more than half of the row's |g| (an outlier: it would set the quantum for every other element),
//     and rows in which most threads see one channel 128x smaller than their total (one scale for all channels would
//     quantise it away).  There every tap adds a second word: the remainder of its first rounding, scaled by 2^19 (lo
//     accumulators: the other, just emptied accumulator buffer), so the row is resolved to 2^-49 of its sum -- still integers,
//     still deterministic.  After an extra barrier each thread folds (hi, lo) of its own cells into fp32.
//   * rows whose gradient is not finite accumulate in fp32 with compare-and-swap adds, taps outside the image skipped: an inf
//     or NaN reaches exactly the cells ATen's scatter would reach (the finite cells of such a row are summed in arrival order).
// Error bound of the fast path: a cell that receives n taps differs from the exact sum by at most n * 2^-31 * 2 * S, S the
// row's sum of |g| over all channels (the scale is the largest power of two with S * scale <= 2^30); exact path: 2^-19 of that.
// Measured and left out (round 2): the float -> fixed-point conversions as fma(g, w, 1.5 * 2^23) - bits(1.5 * 2^23) instead of FMUL + F2I
// for rows whose largest element allows it.  F2I.NTZ does run at only 15.8 conversions / clk / SM (scripts/micro/f2i_rate.cu; 55 for
// the FFMA + IADD pair) and ncu shows the XU pipe as the busiest one of this kernel, yet the kernel did not get faster (576x960 x4
// 28.9 -> 29.2 us, KITTI 97.5 -> 100.9): the conversions are not what the warps wait for.
// Upstream: autograd of utils/disparity_warper.py:83-106 (aten::grid_sampler_2d_backward + the coordinate chain).
#include "warp_lean.cuh"

#include <type_traits>

namespace dssmd {

namespace {

constexpr int kMaxBandRows2 = 32;

__device__ __forceinline__ void red_add_f32_cas(uint32_t addr, float v) {
    // shared-memory fp32 add (there is no native one): compare-and-swap loop; slow path only
    uint32_t old, assumed;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(old) : "r"(addr) : "memory");
    do {
        assumed = old;
        const uint32_t nv = __float_as_uint(__uint_as_float(assumed) + v);
        asm volatile("atom.shared.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "r"(addr), "r"(assumed), "r"(nv) : "memory");
    } while (old != assumed);
}

// OCC3: ask for 768 resident threads per SM instead of 512 (80 registers per thread; the 3-channel build then spills ~60 words:
// kept as a measured alternative, DSSMD_WARP_BWD_BAND_OCC=3)
// fire-and-forget add without the "memory" clobber of red_add(): the accumulator buffer a row is scattered into is touched by
// nothing else between two barriers (read-out and zeroing work on the other buffer), so the compiler may move the row's
// independent global loads and shared-memory traffic across these
__device__ __forceinline__ void red_add_nc(uint32_t addr, int q) {
    asm volatile("red.shared.add.s32 [%0], %1;" ::"r"(addr), "r"(q));
}

// resident blocks asked for: 512 threads per SM (768 with OCC3).  PS = 322 (rows of 1028 .. 1280 pixels: KITTI) asks for two blocks
// of 320 threads (96 registers, 32 - 48 bytes of spill in the 3-channel build): as a 512-thread instance it ran ONE block of 10 warps
// per SM (111.9 -> 97.5 us at 384x1280 x16).  The same for rows of 640 pixels (PS = 162, four blocks of 5 warps instead of three)
// was slower: 22.5 -> 26.5 us at 320x640 x8 (shorter bands, so more halo rows); three blocks of 256 threads at 960 pixels (OCC3,
// 80 registers): 29.0 -> 39.6 us
template <int PS, bool OCC3>
constexpr int band2_min_blocks() {
    return PS == 322 ? 2 : ((OCC3 ? 768 : 512) / (PS - 2)) > 0 ? ((OCC3 ? 768 : 512) / (PS - 2)) : 1;
}

template <int NCT, int PS, bool BORDER, bool OCC3>
__global__ void __launch_bounds__(PS - 2, band2_min_blocks<PS, OCC3>())
warp_bwd_band2_kernel(const float *g, const float *img, const float *disp,
                      float *__restrict__ g_img, float *__restrict__ g_disp, int H, int W, int R,
                      float halfW, float tscale, float halfH, int pf) {
    constexpr int P = PS + 6;                 // accumulator slots per plane; cell of tap position k: p = k + 4
    constexpr int ACH = 4 * P * 4 + 128;      // bytes of one accumulator channel: int acc[4][P]; int sink[32]
    constexpr int ABUF = NCT * ACH;           // one accumulator buffer
    constexpr int VW = NCT == 3 ? 4 : NCT;    // floats per difference cell
    constexpr int DBUF = 4 * PS * VW * 4;     // one difference row: float dvr[4][PS][VW] (cell p = k + 4)
    constexpr int NW = (PS - 2) / 32;         // warps per block (upper bound): 2, 4, 8, 10, 16
    constexpr int NWP = NW <= 4 ? 4 : NW <= 8 ? 8 : 16;   // padded to a power of two (whole float4s; the butterfly below)
    extern __shared__ __align__(16) unsigned char smem_raw[];   // acc[2] | dvr[2]
    __shared__ __align__(16) float s_sum[2][NWP];       // per-warp sums of |g| of a row (all channels)
    __shared__ __align__(16) unsigned s_max[2][NWP];    // per-warp maximum over threads of the thread's sum (bits)
    __shared__ __align__(16) int s_cnt[2][NWP];         // per-warp number of threads that see one channel 128x below their total
    __shared__ __align__(16) float4 s_vw[kMaxBandRows2 + 2];   // per processed row: weight of gV[y] in g_img rows y-1, y, y+1; .w = taps present (bits)
    __shared__ __align__(16) float4 s_vt[kMaxBandRows2 + 2];   // per processed row: its vertical taps (w0, w1, bits of the low row index, -)

    const int W4 = W >> 2;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int b = blockIdx.y;
    const int yb0 = blockIdx.x * R;             // owned rows [yb0, y_end)
    const int y_end = min(yb0 + R, H);
    const float hm1 = (float)(H - 1);
    auto vtaps = [&](int y, int &i0, float &a, float &b2) {
        const float iy_raw = unnorm_f((float)y, hm1, halfH);
        if (BORDER) axis_border(iy_raw, hm1, i0, a, b2);
        else axis_zeros(iy_raw, H, i0, a, b2);
    };
    auto reaches = [&](int y, int r) {   // does output row y have a vertical tap on row r?
        int i0;
        float a, b2;
        vtaps(y, i0, a, b2);
        return ((unsigned)i0 < (unsigned)H && i0 == r) || ((unsigned)(i0 + 1) < (unsigned)H && i0 + 1 == r);
    };
    // processed rows ys .. y_end: the owned rows plus the neighbours whose vertical taps reach an owned row
    const int ys = (yb0 > 0 && reaches(yb0 - 1, yb0)) ? yb0 - 1 : yb0;
    const int y_real = (y_end < H && reaches(y_end, y_end - 1)) ? y_end + 1 : y_end;   // rows y < y_real are scattered
    const unsigned HW = (unsigned)(H * W);
    const bool active = tid < W4;
    const int xl = active ? tid * 4 : W - 4;    // idle threads of the last warp re-read the last pixels (and scatter with scale 0)
    const float wm1 = (float)(W - 1);
    const float rwm1 = div_rn_recip(wm1);
    const uint32_t abase = smem_u32(smem_raw);
    unsigned char *const dvr_base = smem_raw + 2 * ABUF;

    // ---- prologue: vertical weights of the processed rows, zeroed accumulators and statistics ------------------------
    if (tid < y_end - ys + 1) {
        const int y = ys + tid;
        float wr[3] = {0.f, 0.f, 0.f};
        unsigned on = 0u;
        if (y < y_real) {
            int i0;
            float w0, w1;
            vtaps(y, i0, w0, w1);
            const bool in0 = (unsigned)i0 < (unsigned)H, in1 = (unsigned)(i0 + 1) < (unsigned)H;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int r = y - 1 + k;
                if (in0 && i0 == r) { wr[k] = w0; on |= 1u << k; }
                if (in1 && i0 + 1 == r) { wr[k] = w1; on |= 1u << k; }
            }
        }
        s_vw[tid] = make_float4(wr[0], wr[1], wr[2], __uint_as_float(on));
        int i0 = -2;
        float w0 = 0.f, w1 = 0.f;
        if (y < H) vtaps(y, i0, w0, w1);
        s_vt[tid] = make_float4(w0, w1, __int_as_float(i0), 0.f);
    }
    if (tid < 2 * NWP) { (&s_sum[0][0])[tid] = 0.f; (&s_max[0][0])[tid] = 0u; (&s_cnt[0][0])[tid] = 0; }
    {
        // planes 0..3, slots 0 .. P-1 of every channel and buffer; a thread clears slots tid and tid + blockDim.x (P <= 2 * threads)
        for (int s = tid; s < P; s += (int)blockDim.x) {
#pragma unroll
            for (int bf = 0; bf < 2; ++bf)
#pragma unroll
                for (int c = 0; c < NCT; ++c)
#pragma unroll
                    for (int pl = 0; pl < 4; ++pl) stsi_at(smem_raw, (uint32_t)(bf * ABUF + c * ACH + pl * P * 4 + 4 * s), 0);
        }
        if (tid < 32) {
#pragma unroll
            for (int bf = 0; bf < 2; ++bf)
#pragma unroll
                for (int c = 0; c < NCT; ++c) stsi_at(smem_raw, (uint32_t)(bf * ABUF + c * ACH + 4 * P * 4 + 4 * tid), 0);
        }
        // blended rows: cells left of pixel 0 (slot 0 of every plane) and right of pixel W - 1 (slots W4 + 1 .. PS - 1) stay zero
        for (int i = tid; i < 2 * 4 * PS * VW; i += (int)blockDim.x) reinterpret_cast<float *>(smem_raw + 2 * ABUF)[i] = 0.f;
    }

    // per-image bases are uniform over the block (uniform registers); a thread adds 32-bit element offsets c * HW + y * W + xl
    // (lean_eligible: C * H * W < 2^31), so an address is one IMAD.WIDE and no thread holds 64-bit row pointers
    // (opaque(): without it the compiler keeps the 64-bit image offset apart from the parameter and pays four instructions per address)
    auto opaque = [](const float *p) { const float *r; asm volatile("mov.u64 %0, %1;" : "=l"(r) : "l"(p)); return r; };
    const float *const gB = opaque(g + (size_t)b * NCT * HW);
    const float *const dB = opaque(disp + (size_t)b * HW);
    const float *const iB = opaque(img + (size_t)b * NCT * HW);
    float *const oB = const_cast<float *>(opaque(g_img + (size_t)b * NCT * HW));
    float *const odB = const_cast<float *>(opaque(g_disp ? g_disp + (size_t)b * HW : nullptr));
    const bool more = xl + 4 < W;
    const int nx = more ? 4 : 3;   // pixel x + 4 belongs to the next thread; past the row end it is 0 (loaded from x + 3, not used)

    float go[NCT][4];              // g of the row being scattered
    int k0[4];                     // its left-tap positions and weights
    float w0[4], w1[4], gmult[4];
    unsigned cmask = 0u;           // 'border': bit j = pixel j is clamped onto the left border cell
    float pendA[NCT][4], pendB[NCT][4];   // partial g_img rows y-1 and y while row y is being read out
    float down_prev = 0.f;         // fixed point -> float scale of the row to be read out
    bool slow_prev = false;        // ... or that row was accumulated in fp32
#pragma unroll
    for (int c = 0; c < NCT; ++c)
#pragma unroll
        for (int j = 0; j < 4; ++j) { pendA[c][j] = 0.f; pendB[c][j] = 0.f; }

    // ---- stage helpers ------------------------------------------------------------------------------------------
    // per-warp sum and per-warp maximum over threads of the thread's sum of |g| (all channels) -> s_sum / s_max[buf]
    auto abs_sums = [&](const float (&gr)[NCT][4], int buf) {
        float a = 0.f, amin = 0.f;
#pragma unroll
        for (int c = 0; c < NCT; ++c) {
            const float ac = (fabsf(gr[c][0]) + fabsf(gr[c][1])) + (fabsf(gr[c][2]) + fabsf(gr[c][3]));
            a += ac;
            amin = c == 0 ? ac : fminf(amin, ac);
        }
        a = active ? a : 0.f;
        // a >= 0 or NaN: as unsigned bits the order is the numeric one, inf above every finite value, NaN above inf
        const unsigned m = __reduce_max_sync(0xffffffffu, __float_as_uint(a) & 0x7fffffffu);
        int cnt = 0;
        if (NCT > 1) cnt = __reduce_add_sync(0xffffffffu, (active && amin * 128.0f < a) ? 1 : 0);
        float r = a;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
        if (lane == 0) { s_sum[buf][wid] = r; s_max[buf][wid] = m; if (NCT > 1) s_cnt[buf][wid] = cnt; }
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
            // idle lanes of the last warp re-read the last thread's pixels: their (zero) adds go to their own cells right of the
            // row, not 17-fold onto that thread's cells (same-address shared atomics of a warp serialise)
            if (!active) k0[j] = 4 * tid + j;
        }
        if (!active) cmask = 0u;
    };
    // horizontal differences of the blended source rows of output row y -> dvr[buf]: with B[k] = yw0 * img[ya][k] + yw1 * img[yb][k]
    // (0 outside the row), cell k holds D[k] = B[k + 1] - B[k] for all channels; g_disp reads one cell per pixel.  Cells left of
    // k = -1 and right of k = W - 1 were zeroed once in the prologue and are never written.
    auto stage_blend = [&](int y, int buf) {
        const float4 vt = s_vt[y - ys];
        const float yw0 = vt.x, yw1 = vt.y;
        const int yi0 = __float_as_int(vt.z);
        const int ya = min(max(yi0, 0), H - 1), yb = min(max(yi0 + 1, 0), H - 1);
        const unsigned oa = (unsigned)(ya * W + xl), ob2 = (unsigned)(yb * W + xl);
        float bl[NCT][5];
#pragma unroll
        for (int c = 0; c < NCT; ++c) {
            const float4 r0 = ld_ca(reinterpret_cast<const float4 *>(iB + (oa + (unsigned)c * HW)));
            const float4 r1 = ld_ca(reinterpret_cast<const float4 *>(iB + (ob2 + (unsigned)c * HW)));
            const float n0 = ld_ca(iB + (oa + (unsigned)c * HW + (unsigned)nx));
            const float n1 = ld_ca(iB + (ob2 + (unsigned)c * HW + (unsigned)nx));
            bl[c][0] = fmaf(yw1, r1.x, yw0 * r0.x); bl[c][1] = fmaf(yw1, r1.y, yw0 * r0.y);
            bl[c][2] = fmaf(yw1, r1.z, yw0 * r0.z); bl[c][3] = fmaf(yw1, r1.w, yw0 * r0.w);
            bl[c][4] = more ? fmaf(yw1, n1, yw0 * n0) : 0.f;
        }
        if (active) {
            unsigned char *const dvr = dvr_base + buf * DBUF;
            const uint32_t sa = (uint32_t)(VW * 4) * ((uint32_t)tid + 1u);  // cells 4*tid+4 .. +7: slot tid + 1 of planes 0..3
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float cell[VW];
#pragma unroll
                for (int c = 0; c < VW; ++c) cell[c] = c < NCT ? bl[c < NCT ? c : 0][j + 1] - bl[c < NCT ? c : 0][j] : 0.f;
                sts_cell<VW>(dvr, sa + (uint32_t)(j * PS * VW * 4), cell);
            }
            if (!BORDER && tid == 0) {   // cell 3 = k -1 ('zeros' only: left tap outside, right tap on pixel 0): D = B[0] - 0
                float e[VW];
#pragma unroll
                for (int c = 0; c < VW; ++c) e[c] = c < NCT ? bl[c < NCT ? c : 0][0] : 0.f;
                sts_cell<VW>(dvr, 3u * PS * VW * 4, e);
            }
        }
    };
    // read out the finished accumulators of row y (zeroing them), fold the row into the pending g_img rows and emit g_img
    // row y - 1.  A row that was not scattered (beyond y_real) reads zeros with zero weights.
    auto read_out = [&](int y, auto slow_tag) {
        constexpr bool SLOW = decltype(slow_tag)::value;
        // read and clear a cell with ONE shared-memory operation (atom.exch) instead of a load and a store: the shared-memory pipe is
        // this kernel's busiest unit (ncu: 58 %); 576x960 x4 28.9 -> 28.4 us, 384x1280 x16 97.5 -> 96.4 us
        constexpr bool XCHG = true;
        const float4 vw = s_vw[y - ys];
        const uint32_t ra = 4u * (uint32_t)tid + 4u + (uint32_t)((y & 1) * ABUF);   // slot tid + 1 of planes 0..3
        const int r_out = y - 1;
        const bool emit = active && r_out >= yb0 && r_out < y_end;
        const unsigned oo = (unsigned)(r_out * W + xl);
        float wa, wb, wc;
        bool on0 = true, on1 = true, on2 = true;
        if (SLOW) {
            // weights as they are; a tap that is absent must not touch its row even with a NaN operand
            const unsigned on = __float_as_uint(vw.w);
            wa = vw.x; wb = vw.y; wc = vw.z;
            on0 = on & 1u; on1 = on & 2u; on2 = on & 4u;
        } else {
            wa = vw.x * down_prev; wb = vw.y * down_prev; wc = vw.z * down_prev;   // exact: power of two
        }
#pragma unroll
        for (int c = 0; c < NCT; ++c) {
            float o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t a = ra + (uint32_t)(c * ACH + j * P * 4);
                int v;
                if (XCHG) {
                    asm volatile("atom.shared.exch.b32 %0, [%1], %2;" : "=r"(v) : "r"(abase + a), "r"(0) : "memory");
                } else {
                    v = ldsi_at(smem_raw, a);
                    stsi_at(smem_raw, a, 0);
                }
                if (SLOW) {
                    const float cur = __int_as_float(v);
                    o[j] = on0 ? fmaf(wa, cur, pendA[c][j]) : pendA[c][j];
                    pendA[c][j] = on1 ? fmaf(wb, cur, pendB[c][j]) : pendB[c][j];
                    pendB[c][j] = on2 ? wc * cur : 0.f;
                } else {
                    const float cur = (float)v;
                    o[j] = fmaf(wa, cur, pendA[c][j]);
                    pendA[c][j] = fmaf(wb, cur, pendB[c][j]);
                    pendB[c][j] = wc * cur;
                }
            }
            if (emit) st_global_cs(reinterpret_cast<float4 *>(oB + (oo + (unsigned)c * HW)), make_float4(o[0], o[1], o[2], o[3]));
        }
    };

    // ---- first row ------------------------------------------------------------------------------------------------
    __syncthreads();   // table, zeroed accumulators and statistics visible
    pdl_launch_dependents();
    if (pf > 0 && active && (tid & 7) == 0) {
        // the first rows of the band towards the L2 while the previous kernel drains (common.cuh: prefetch_l2); one thread per
        // 128 bytes of a row
        const int yp1 = min(ys + pf, y_real);
        for (int y = ys; y < yp1; ++y) {
            const unsigned o = (unsigned)(y * W + xl);
            prefetch_l2(dB + o);
#pragma unroll
            for (int c = 0; c < NCT; ++c) prefetch_l2(gB + (o + (unsigned)c * HW));
        }
        if (g_disp) {
            const int ya = max(ys - 1, 0), yb = min(yp1, H - 1);   // the source rows those output rows blend
            for (int y = ya; y <= yb; ++y) {
                const unsigned o = (unsigned)(y * W + xl);
#pragma unroll
                for (int c = 0; c < NCT; ++c) prefetch_l2(iB + (o + (unsigned)c * HW));
            }
        }
    }
    pdl_wait();        // no load or store above
    {
        const unsigned orow = (unsigned)(ys * W + xl);
        const float4 d4 = ld_cg(reinterpret_cast<const float4 *>(dB + orow));
#pragma unroll
        for (int c = 0; c < NCT; ++c) {
            const float4 v = ld_cg(reinterpret_cast<const float4 *>(gB + (orow + (unsigned)c * HW)));
            go[c][0] = v.x; go[c][1] = v.y; go[c][2] = v.z; go[c][3] = v.w;
        }
        if (g_disp && ys >= yb0) stage_blend(ys, ys & 1);
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
            const unsigned nrow = (unsigned)((y + 1) * W + xl);
            // g and disp are read once: L2-only loads (ld.global.cg) keep the L1 for the image rows, which two consecutive
            // output rows share (ncu: 15 % L1 hit rate with allocating loads, the blend below waiting on DRAM)
            dn = ld_cg(reinterpret_cast<const float4 *>(dB + nrow));
#pragma unroll
            for (int c = 0; c < NCT; ++c) {
                const float4 v = ld_cg(reinterpret_cast<const float4 *>(gB + (nrow + (unsigned)c * HW)));
                gn[c][0] = v.x; gn[c][1] = v.y; gn[c][2] = v.z; gn[c][3] = v.w;
            }
        }
        const bool ndiff = nreal && g_disp && y + 1 >= yb0 && y + 1 < y_end;
        // the image row that g_disp of row y + 1 differentiates and that row y did not touch: towards L1 now, read by stage_blend
        // at the end of the iteration (consecutive output rows share a source row).  Measured alternatives, both slower (30.0 -
        // 30.2 us against 28.8 at 576x960 x4): the row loaded into registers here, or after the scatter
        if (ndiff && (lane & 7) == 0) {
            const int yb = min(max(__float_as_int(s_vt[y + 1 - ys].z) + 1, 0), H - 1);
            const unsigned po = (unsigned)(yb * W + xl);
#pragma unroll
            for (int c = 0; c < NCT; ++c) asm volatile("prefetch.global.L1 [%0];" ::"l"(iB + (po + (unsigned)c * HW)));
        }
        // scale of row y from the per-warp statistics (fixed order: every thread computes the same bits)
        float up = 0.f, down = 0.f;
        int kind = 0;
        if (real) {
            // lane l reads warp (l mod NWP)'s entry: one wavefront; the xor butterfly adds commute, so every lane ends with the
            // same bits
            float S = s_sum[buf][lane & (NWP - 1)];
            unsigned M = s_max[buf][lane & (NWP - 1)];
            int F = NCT > 1 ? s_cnt[buf][lane & (NWP - 1)] : 0;
#pragma unroll
            for (int o = NWP / 2; o > 0; o >>= 1) {
                S += __shfl_xor_sync(0xffffffffu, S, o);
                M = max(M, __shfl_xor_sync(0xffffffffu, M, o));
                if (NCT > 1) F += __shfl_xor_sync(0xffffffffu, F, o);
            }
            const int sexp = (int)(__float_as_uint(S) >> 23);   // S >= 0 or NaN: no sign bit to mask (NaN: >= 0xff either way)
            int e = 156 - (sexp > 0 ? sexp : 1);                // S < 2^(sexp-126)  =>  S * 2^e < 2^30
            if (e > 127) e = 127;
            up = (S > 0.f) ? __uint_as_float((unsigned)(e + 127) << 23) : 0.f;  // 0: nothing to add
            down = (e < 127) ? __uint_as_float((unsigned)(127 - e) << 23) : __uint_as_float(0x00400000u);
            // 2: not finite -> fp32 compare-and-swap adds.  1: an outlier thread (more than half of the row's |g|) or a channel far
            // below the others in most threads -> exact (hi, lo) fixed point
            kind = (sexp & 0xff) == 0xff ? 2 : ((W4 >= 32 && __uint_as_float(M) > 0.5f * S) || 2 * F > W4) ? 1 : 0;
#ifdef BAND2_CENSUS   // instruction census of the hot path (scripts/sass_census.py): slow paths and clamped scatter compiled out
            kind = 0;
#endif
        }

        if (y > ys) {
            if (slow_prev) read_out(y - 1, std::true_type{});
            else read_out(y - 1, std::false_type{});
        }

        if (real) {
            // g_disp of row y (owned rows only)
            if (g_disp && y >= yb0 && y < y_end && active) {
                const unsigned char *const dvr = dvr_base + buf * DBUF;
                float gd[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float dvc[VW];
                    lds_cell<VW>(dvr, vcell_off<PS, VW>((uint32_t)(k0[j] + 4)), dvc);   // k0 in [-2, W]: cells 2 .. W + 4 exist
                    float gix = 0.f;
#pragma unroll
                    for (int c = 0; c < NCT; ++c) gix = fmaf(go[c][j], dvc[c], gix);
                    // chain rule: ix <- gx (W/2, 0 where clipped) <- t (2/(W-1)) <- disp (-1)
                    gd[j] = -(tscale * (gmult[j] * gix));
                }
                *reinterpret_cast<float4 *>(odB + (unsigned)(y * W + xl)) = make_float4(gd[0], gd[1], gd[2], gd[3]);
            }

            // scatter row y into acc[buf].  'border': every pixel whose coordinate is clamped to the left edge adds to cell
            // k = 0 (same-address shared atomics of a warp serialise): those go to the lane's private sink word, and their
            // contributions are summed across the warp in integers (exact, order-free) and added to the cell once.
            const uint32_t ab = abase + (uint32_t)(buf * ABUF);
            const float ups = active ? up : 0.f;   // idle lanes add zeros
            if (kind == 0) {
                auto scatter = [&](auto clamped_path) {
                    constexpr bool CL = decltype(clamped_path)::value;
                    const uint32_t sink = ab + (uint32_t)(4 * P * 4) + 4u * (uint32_t)lane;   // + c * ACH: the sink of channel c
                    uint32_t aL[4], aR[4];
                    float wl[4], wr2[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint32_t p = (uint32_t)(k0[j] + 4), q = p + 1u;
                        const uint32_t t = ab + p;
                        aL[j] = t + (p & 3u) * (uint32_t)(4 * P - 1);
                        aR[j] = t + (q & 3u) * (uint32_t)(4 * P - 1) + 1u;
                        if (CL && ((cmask >> j) & 1u)) { aL[j] = sink; aR[j] = sink; }
                        wl[j] = w0[j] * ups; wr2[j] = w1[j] * ups;   // exact: power-of-two scale
                    }
#pragma unroll
                    for (int c = 0; c < NCT; ++c) {
                        int qc = 0;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int qL = __float2int_rn(go[c][j] * wl[j]);
                            red_add_nc(aL[j] + (uint32_t)(c * ACH), qL);
                            red_add_nc(aR[j] + (uint32_t)(c * ACH), __float2int_rn(go[c][j] * wr2[j]));
                            if (CL) qc += ((cmask >> j) & 1u) ? qL : 0;
                        }
                        if (CL) {
                            const int tot = __reduce_add_sync(0xffffffffu, qc);
                            if (lane == 0) red_add_nc(ab + 4u + (uint32_t)(c * ACH), tot);   // cell k = 0: p = 4 -> plane 0, slot 1
                        }
                    }
                };
#ifdef BAND2_CENSUS
                scatter(std::false_type{});
#else
                if (BORDER && __any_sync(0xffffffffu, active && cmask != 0u)) scatter(std::true_type{});
                else scatter(std::false_type{});
#endif
            } else if (kind == 1) {
                // exact path: hi = rn(g * w * 2^e) into acc[buf] as in the fast path, lo = rn((g * w * 2^e - hi) * 2^19) into the lo
                // buffer (|lo| <= 2^18 per tap, at most 2 W <= 4096 taps per cell: no overflow).  fma gives the remainder exactly.
                // The lo buffer is the OTHER accumulator buffer: the read-out above has emptied this thread's cells of it, and
                // after a barrier everybody's (the buffer is scattered into again only after the barrier that ends this iteration)
                __syncthreads();   // kind is uniform over the block
                const uint32_t LO_OFF = (uint32_t)((buf ^ 1) * ABUF);
                const uint32_t lob = abase + LO_OFF;
                const float ups1 = active ? up : 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t p = (uint32_t)(k0[j] + 4), q = p + 1u;
                    const uint32_t oL = p + (p & 3u) * (uint32_t)(4 * P - 1);
                    const uint32_t oR = p + (q & 3u) * (uint32_t)(4 * P - 1) + 1u;
                    const float wl = w0[j] * ups1, wr2 = w1[j] * ups1;
#pragma unroll
                    for (int c = 0; c < NCT; ++c) {
                        const float gv = go[c][j];
                        const int hL = __float2int_rn(gv * wl), hR = __float2int_rn(gv * wr2);
                        const int lL = __float2int_rn(fmaf(gv, wl, -(float)hL) * 524288.0f);
                        const int lR = __float2int_rn(fmaf(gv, wr2, -(float)hR) * 524288.0f);
                        red_add(ab + oL + (uint32_t)(c * ACH), hL); red_add(lob + oL + (uint32_t)(c * ACH), lL);
                        red_add(ab + oR + (uint32_t)(c * ACH), hR); red_add(lob + oR + (uint32_t)(c * ACH), lR);
                    }
                }
                __syncthreads();
                // fold (hi, lo) of the thread's own cells into fp32 (what the read-out of a slow row expects), clear lo
                if (active) {
#pragma unroll
                    for (int c = 0; c < NCT; ++c)
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const uint32_t a = 4u * (uint32_t)tid + 4u + (uint32_t)(c * ACH + j * P * 4);
                            const int hi = ldsi_at(smem_raw, (uint32_t)(buf * ABUF) + a), lo = ldsi_at(smem_raw, LO_OFF + a);
                            stsi_at(smem_raw, LO_OFF + a, 0);
                            const float v = fmaf((float)lo, 1.9073486328125e-06f, (float)hi) * down;   // (hi + lo * 2^-19) * 2^-e
                            stsi_at(smem_raw, (uint32_t)(buf * ABUF) + a, __float_as_int(v));
                        }
                }
            } else if (active) {
                // fp32 accumulation (rows with a non-finite gradient): the accumulator words hold float bits; a tap ATen would skip
                // (outside the image) is skipped, so an inf / NaN reaches exactly the cells ATen's scatter would reach
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t p = (uint32_t)(k0[j] + 4), q = p + 1u;
                    const uint32_t aLj = ab + p + (p & 3u) * (uint32_t)(4 * P - 1);
                    const uint32_t aRj = ab + p + (q & 3u) * (uint32_t)(4 * P - 1) + 1u;
                    const bool okL = BORDER ? true : (unsigned)k0[j] < (unsigned)W;
                    const bool okR = (unsigned)(k0[j] + 1) < (unsigned)W;
#pragma unroll
                    for (int c = 0; c < NCT; ++c) {
                        if (okL) red_add_f32_cas(aLj + (uint32_t)(c * ACH), go[c][j] * w0[j]);
                        if (okR) red_add_f32_cas(aRj + (uint32_t)(c * ACH), go[c][j] * w1[j]);
                    }
                }
            }
        }
        down_prev = down;
        slow_prev = kind != 0;

        // stage row y + 1
        if (nreal) {
            if (ndiff) stage_blend(y + 1, buf ^ 1);
            abs_sums(gn, buf ^ 1);
            taps(dn);
#pragma unroll
            for (int c = 0; c < NCT; ++c)
#pragma unroll
                for (int j = 0; j < 4; ++j) go[c][j] = gn[c][j];
        }
        __syncthreads();
    }
    if (slow_prev) read_out(y_end, std::true_type{});
    else read_out(y_end, std::false_type{});
}

template <int PS, bool BORDER, bool OCC3>
int launch_band2(const float *g, const float *img, const float *disp, float *g_img, float *g_disp, int B, int C, int H, int W, cudaStream_t s) {
    void (*kern)(const float *, const float *, const float *, float *, float *, int, int, int, float, float, float, int) =
        C == 1 ? warp_bwd_band2_kernel<1, PS, BORDER, OCC3> : C == 2 ? warp_bwd_band2_kernel<2, PS, BORDER, OCC3>
      : C == 3 ? warp_bwd_band2_kernel<3, PS, BORDER, OCC3> : warp_bwd_band2_kernel<4, PS, BORDER, OCC3>;
    const int vw = C == 3 ? 4 : C;
    const size_t smem = 2 * (size_t)C * (4 * (PS + 6) * 4 + 128) + 2 * ((size_t)4 * PS * vw * 4);   // acc[2] | dvr[2]
    const int threads = round_up(W / 4, 32);
    cudaError_t e = kernel_smem_optin(kern, smem);   // asked once per (kernel, device), then a table lookup (abi.cu)
    if (e != cudaSuccess) return fail(DSSMD_ERR_CUDA, "warp_bwd(band2): cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(e));
    const int occ = kernel_occupancy(kern, threads, smem);
    int R = env_int_lean("DSSMD_WARP_BWD_BAND_ROWS", 0);
    if (R <= 0) {
        // a band costs R + 1 row passes (one halo row) and the grid runs in ceil(blocks / slots) waves: the R that minimises
        // waves * (R + 1); ties: the larger R (fewer halo rows)
        const long long slots = (long long)occ * sm_count();
        long long best = -1;
        for (int r = 4; r <= kMaxBandRows2 && r <= (H > 4 ? H : 4); ++r) {
            const long long blocks = (long long)ceil_div(H, r) * B;
            const long long cost = ((blocks + slots - 1) / slots) * (r + 1);
            if (best < 0 || cost <= best) { best = cost; R = r; }
        }
        if (R < 4) R = 4;
    }
    if (R > kMaxBandRows2) R = kMaxBandRows2;
    if (R > H) R = H;
    if (env_int_lean("DSSMD_CORR_DEBUG", 0))
        fprintf(stderr, "warp_bwd band2 cfg: PS=%d border=%d occ3=%d R=%d grid=(%d,%d) (x%d/SM) threads=%d smem=%zu\n", PS, (int)BORDER, (int)OCC3, R,
                ceil_div(H, R), B, occ, threads, smem);
    e = launch_pdl(kern, dim3((unsigned)ceil_div(H, R), (unsigned)B), dim3((unsigned)threads), smem, s, g, img, disp, g_img, g_disp, H, W, R,
                   (float)W / 2.0f, 2.0f / (float)(W - 1), (float)H / 2.0f, pdl_prefetch(1));
    if (e != cudaSuccess) return fail(DSSMD_ERR_CUDA, "warp_bwd(band2): launch failed: %s", cudaGetErrorString(e));
    return check_launch("warp_bwd(band2)");
}

}  // namespace

// g_img must be non-null (g_disp may be null).  -1 when the shape is not eligible (the caller falls back).
int warp_bwd_band2_f32(const float *g, const float *img, const float *disp, float *g_img, float *g_disp,
                       int B, int C, int H, int W, int border, cudaStream_t s) {
    if (!g_img || !lean_eligible(B, C, H, W) || W > 2048) return -1;
    if (((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(img) | reinterpret_cast<uintptr_t>(disp) |
          reinterpret_cast<uintptr_t>(g_img) | reinterpret_cast<uintptr_t>(g_disp)) & 15) != 0) return -1;
    const int w4 = W / 4;
    const bool occ3 = env_int_lean("DSSMD_WARP_BWD_BAND_OCC", 2) == 3;
#define DSSMD_BAND2(PSV)                                                                                                   \
    if (w4 + 2 <= PSV) {                                                                                                   \
        if (PSV == 258 && occ3)                                                                                            \
            return border ? launch_band2<PSV, true, PSV == 258>(g, img, disp, g_img, g_disp, B, C, H, W, s)                \
                          : launch_band2<PSV, false, PSV == 258>(g, img, disp, g_img, g_disp, B, C, H, W, s);              \
        return border ? launch_band2<PSV, true, false>(g, img, disp, g_img, g_disp, B, C, H, W, s)                         \
                      : launch_band2<PSV, false, false>(g, img, disp, g_img, g_disp, B, C, H, W, s);                       \
    }
    // the 320-thread instance (two blocks per SM at 96 registers) only pays when there are blocks to fill the second slot: measured at
    // 384x1280 over the batch size (scripts/kitti_warp_b_sweep.py), 512- / 320-thread instance: B = 1 13.1 / 17.2 us, B = 2 18.0 / 21.4,
    // B = 4 29.6 / 30.2, B = 6 41.4 / 39.4, B = 8 54.9 / 51.3, B = 16 109.6 / 96.6
    const int w320 = env_int_lean("DSSMD_WARP_BWD_BAND_W320", -1);
    const bool use320 = w320 >= 0 ? w320 != 0 : (long)B * H >= 2048;
    DSSMD_BAND2(66) DSSMD_BAND2(130) DSSMD_BAND2(258)
    if (use320) { DSSMD_BAND2(322) }
    DSSMD_BAND2(514)
#undef DSSMD_BAND2
    return -1;
}

}  // namespace dssmd
