// corr_bwd_fused.cu -- fp32 backward of the correlation volume, BOTH feature gradients in one launch:
//
//   gL[c,x ] = (1/C) * sum_d g[d,x]    * R[c,x-d]
//   gR[c,x'] = (1/C) * sum_d g[d,x'+d] * L[c,x'+d]
//
// (the autograd graph of build_corr, upstream models/UtilsNet/submoudles.py:301-311).  The two-launch path (corr_bwd_reg.cu /
// corr_bwd_mma32.cu, one kernel per gradient) spends ~6 us of each 14-15 us launch at the model's 1/8-resolution shapes on
// its ramp (block start, first TMA tile, tail), reads g twice, and its generic inner loop issues ~107 instructions per
// channel and warp where ~65 are needed (ncu: 7.9 M warp instructions, issue-bound at 16 warps per SM).  Here:
//   * one persistent block per (up to) two rows walks (row, block of CB channels) items; the warps are split in two halves,
//     one per gradient -- same blocking as corr_bwd_reg.cu: a lane keeps a 4-pixel x 12-disparity slab of g in registers
//     (gR: skewed), reads ONE aligned 16-word window of the operand row per channel (4 LDS.128, conflict-free lane map) and
//     issues 48 FFMAs; the D disparities are split over NH neighbouring lanes and summed by a butterfly (fixed order);
//   * an item's two operand tiles (R with a left halo for gL, L with a right halo for gR) and, on a new row, the g tile
//     arrive as TMA boxes (zero fill outside the row = the x - d < 0 / x' + d >= W limits of the sums), three stages deep;
//   * channels per warp and the lane map are compile-time: the channel loop is fully unrolled, shared-memory offsets are
//     immediates, no per-channel index arithmetic.
// No atomics, deterministic.  fp32, rows that fit one tile (W + halo <= 256), D <= 48; everything else keeps the two-launch path.
#include "corr_reg.cuh"

#include <cstdlib>
#include <type_traits>

namespace dssmd {

namespace {

constexpr int kFStages = 3;

struct FusedParams {
    float *gL, *gR;
    int C, H, W, D;
    int WPR;        // warps side by side along x (per gradient)
    int CGW;        // channel groups of warps (per gradient)
    int CB;         // channels per item = CGW * CPW
    int NCB;        // channel blocks per row
    int HALO;       // halo of the operand boxes in elements (12 * NH)
    int WP;         // box width in elements (TX + HALO), operands and g
    int nrows, rpb; // rows (b, y), rows per block
    uint32_t op_bytes, g_bytes;
    float invC;
};

template <int LOGNH, int CPW, int NW>
__global__ void __launch_bounds__(NW * 32, NW <= 8 ? 2 : 1)
corr_bwd_fused_kernel(const FusedParams p, const __grid_constant__ CUtensorMap mapL, const __grid_constant__ CUtensorMap mapR,
                      const __grid_constant__ CUtensorMap mapG) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long s_bar[kFStages];
    constexpr int PX = 4;
    using LM = LaneMap<LOGNH, PX>;
    // stage s: [R box | L box]; after the stages: two g tiles (rows alternate)
    const uint32_t op_stride = (p.op_bytes + 127u) & ~127u, g_stride = (p.g_bytes + 127u) & ~127u;
    const uint32_t stage_bytes = 2u * op_stride;
    const uint32_t g_base = (uint32_t)kFStages * stage_bytes;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r0 = blockIdx.x * p.rpb;
    const int r1 = (r0 + p.rpb < p.nrows) ? r0 + p.rpb : p.nrows;
    if (r0 >= r1) return;
    const int nitems = (r1 - r0) * p.NCB;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kFStages; ++s) mbar_init(smem_u32(&s_bar[s]), 1);
        fence_proxy_async();
    }
    __syncthreads();
    pdl_launch_dependents();  // the next kernel on the stream may start its ramp
    pdl_wait();               // nothing above touches global memory

    // producer state (thread 0): the next item to request, as counters (no divisions in the loop)
    int is = 0, icb = 0, irr = 0, ib = r0 / p.H, iy = r0 - (r0 / p.H) * p.H;
    auto issue = [&]() {  // one thread; requests item (irr, icb) into stage `is` and advances
        const int s = is, cb = icb, rr = irr, b = ib, y = iy;
        if (++is == kFStages) is = 0;
        if (++icb == p.NCB) { icb = 0; ++irr; if (++iy == p.H) { iy = 0; ++ib; } }
        const uint32_t bar = smem_u32(&s_bar[s]);
        const uint32_t dst = smem_u32(smem_raw) + (uint32_t)s * stage_bytes;
        fence_proxy_async();  // the stage was read through the generic proxy before the block barrier
        mbar_arrive_expect_tx(bar, 2u * p.op_bytes + (cb == 0 ? p.g_bytes : 0u));
        tma_load_4d(dst, &mapR, -p.HALO, y, cb * p.CB, b, bar);          // gL: R[x - d], left halo
        tma_load_4d(dst + op_stride, &mapL, 0, y, cb * p.CB, b, bar);    // gR: L[x' + d], right halo
        if (cb == 0) tma_load_4d(smem_u32(smem_raw) + g_base + (uint32_t)(rr & 1) * g_stride, &mapG, 0, y, 0, b, bar);
    };
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < kFStages - 1; ++k)
            if (k < nitems) issue();
    }

    // this warp's gradient, x segment and channel group; this lane's pixels and disparity part
    const int mode = warp >= NW / 2 ? 1 : 0;
    const int w2 = warp - mode * (NW / 2);
    const int wseg = w2 % p.WPR, cgw = w2 / p.WPR;
    const int h = LM::h(lane), pgl = LM::pgl(lane);
    const int xl = (wseg * (32 >> LOGNH) + pgl) * PX;                       // < TX by construction
    const int woff = mode == 0 ? xl + p.HALO - kDP * (h + 1) : xl + kDP * h;  // window start inside a staged row
    const size_t HW = (size_t)p.H * p.W;
    const bool storer = (h == 0) && (xl < p.W);                             // W % 4 == 0: all four pixels or none
    float *const outp = mode == 0 ? p.gL : p.gR;
    // byte offset of this warp's first channel row inside a stage
    const uint32_t wbase = (uint32_t)mode * op_stride + (uint32_t)((cgw * CPW) * p.WP + woff) * 4u;
    const uint32_t wp4 = (uint32_t)p.WP * 4u;

    float gv[kDP][PX];
    int s = 0, cb = 0, rr = 0, b = r0 / p.H, y = r0 - (r0 / p.H) * p.H;   // consumer state: the item being computed
    uint32_t ph = 0;
    for (int it = 0; it < nitems; ++it) {
        if (tid == 0 && it + kFStages - 1 < nitems) issue();   // its stage was released by the barrier ending item it-1
        mbar_wait(smem_u32(&s_bar[s]), ph);

        if (cb == 0) {  // new row: the g slab, pre-multiplied by 1/C
            const float *gs = reinterpret_cast<const float *>(smem_raw + g_base + (size_t)(rr & 1) * g_stride);
#pragma unroll
            for (int dd = 0; dd < kDP; ++dd) {
                const int d = h * kDP + dd;
                const bool dok = d < p.D;
                const float *gd = gs + (dok ? d : 0) * p.WP + xl;
                if (mode == 0) {
                    const float4 v = *reinterpret_cast<const float4 *>(gd);
                    gv[dd][0] = dok ? v.x * p.invC : 0.f; gv[dd][1] = dok ? v.y * p.invC : 0.f;
                    gv[dd][2] = dok ? v.z * p.invC : 0.f; gv[dd][3] = dok ? v.w * p.invC : 0.f;
                } else {
#pragma unroll
                    for (int i = 0; i < PX; ++i) gv[dd][i] = dok ? gd[i + d] * p.invC : 0.f;
                }
            }
        }

        const unsigned char *wrow = smem_raw + (size_t)s * stage_bytes + wbase;
        const int c0 = cb * p.CB + cgw * CPW;
        float *o = outp + ((size_t)b * p.C + c0) * HW + (size_t)y * p.W + xl;
        auto channels = [&](auto mode_tag) {   // the warp's gradient is uniform: one branch per item, not per channel
            constexpr int MODE = decltype(mode_tag)::value;
#pragma unroll
            for (int k = 0; k < CPW; ++k) {
                float w[PX + kDP];
#pragma unroll
                for (int q = 0; q < (PX + kDP) / 4; ++q) {
                    const float4 v = *reinterpret_cast<const float4 *>(wrow + (size_t)k * wp4 + 16 * q);
                    w[4 * q + 0] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
                }
                float acc[PX];
#pragma unroll
                for (int i = 0; i < PX; ++i) acc[i] = gv[0][i] * w[MODE == 0 ? i + kDP : i];
#pragma unroll
                for (int dd = 1; dd < kDP; ++dd)
#pragma unroll
                    for (int i = 0; i < PX; ++i) acc[i] = fmaf(gv[dd][i], w[MODE == 0 ? i - dd + kDP : i + dd], acc[i]);
                // the NH disparity parts of a pixel group: butterfly over the lane bits that carry h (fixed order, all lanes of
                // a group end with the same bits); the h = 0 lane stores
#pragma unroll
                for (int r = 0; r < LOGNH; ++r) {
                    const unsigned m = 1u << LM::hpos(r);
#pragma unroll
                    for (int i = 0; i < PX; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], m);
                }
                if (storer && c0 + k < p.C) st_global_cs(reinterpret_cast<float4 *>(o + (size_t)k * HW), make_float4(acc[0], acc[1], acc[2], acc[3]));
            }
        };
        if (mode == 0) channels(std::integral_constant<int, 0>{});
        else channels(std::integral_constant<int, 1>{});
        __syncthreads();  // everyone is done with stage s (and, on a new row, has its g slab)
        if (++s == kFStages) { s = 0; ph ^= 1u; }
        if (++cb == p.NCB) { cb = 0; ++rr; if (++y == p.H) { y = 0; ++b; } }
    }
}

int env_int_fused(const char *name, int dflt) {
    if (!tuning_enabled()) return dflt;
    const char *v = std::getenv(name);
    return (v && *v) ? std::atoi(v) : dflt;
}

template <int LOGNH, int CPW, int NW>
int launch_fused(FusedParams p, const CUtensorMap &mapL, const CUtensorMap &mapR, const CUtensorMap &mapG, size_t smem, cudaStream_t stream) {
    auto kern = corr_bwd_fused_kernel<LOGNH, CPW, NW>;
    {
        cudaError_t e = kernel_smem_optin(kern, smem);   // asked once per (kernel, device), then a table lookup (abi.cu)
        if (e != cudaSuccess) return fail(DSSMD_ERR_CUDA, "corr_bwd(fused): cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(e));
    }
    const int bps = kernel_occupancy(kern, NW * 32, smem);
    // persistent grid: the blocks that are resident at once, each with an equal share of rows
    int grid = sm_count() * bps;
    if (grid > p.nrows) grid = p.nrows;
    p.rpb = ceil_div(p.nrows, grid);
    grid = ceil_div(p.nrows, p.rpb);
    if (env_int_fused("DSSMD_CORR_DEBUG", 0))
        fprintf(stderr, "corr_bwd fused cfg: NH=%d CPW=%d warps=%d WPR=%d CGW=%d CB=%d NCB=%d WP=%d rows=%d rpb=%d grid=%d (x%d/SM) smem=%zu\n",
                1 << LOGNH, CPW, NW, p.WPR, p.CGW, p.CB, p.NCB, p.WP, p.nrows, p.rpb, grid, bps, smem);
    const cudaError_t le = launch_pdl(kern, dim3((unsigned)grid), dim3((unsigned)(NW * 32)), smem, stream, p, mapL, mapR, mapG);
    if (le != cudaSuccess) return fail(DSSMD_ERR_CUDA, "corr_bwd(fused): launch failed: %s", cudaGetErrorString(le));
    return check_launch("corr_bwd(fused)");
}

}  // namespace

// Both gradients of an fp32 correlation in one launch.  -1 when the shape is not eligible (the caller launches one kernel per
// gradient), else a DSSMD_* code.  gbs = elements between two images of g (the concat buffer's stride for corr_bwd_cat).
int corr_bwd_fused_f32(const void *g, const void *L, const void *R, void *gL, void *gR, int B, int C, int H, int W, int D, long long gbs,
                       cudaStream_t stream) {
    if (!gL || !gR || env_int_fused("DSSMD_CORR_BWD_FUSED", 1) == 0) return -1;
    if (W % 4 != 0 || D > kDP * 4) return -1;
    if (((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(L) | reinterpret_cast<uintptr_t>(R) | reinterpret_cast<uintptr_t>(gL) |
          reinterpret_cast<uintptr_t>(gR)) & 15) != 0) return -1;
    if ((gbs * 4) % 16 != 0 || (long long)B * H >= (1LL << 30)) return -1;
    FusedParams p{};
    p.gL = (float *)gL; p.gR = (float *)gR;
    p.C = C; p.H = H; p.W = W; p.D = D;
    p.invC = 1.0f / (float)C;
    int nh = 1, lognh = 0;
    while (nh * kDP < D) { nh *= 2; ++lognh; }
    p.HALO = kDP * nh;
    const int pxw = (32 / nh) * 4;            // pixels one warp covers
    const int wpr = ceil_div(W, pxw);
    const int tx = wpr * pxw;
    if (tx + p.HALO > 256) return -1;         // one TMA box per row
    // rows that leave more than a quarter of the lanes idle are better served by the mma.sync kernels (corr_bwd.cu)
    if (env_int_fused("DSSMD_CORR_BWD_FUSED", 1) < 2 && (double)W < 0.75 * (double)tx) return -1;
    p.WPR = wpr;
    p.WP = tx + p.HALO;
    // warps: two gradients x wpr x cgw.  Taken by default only where 8 warps (two resident blocks of 256 threads) divide:
    // measured against the two-launch path -- 576x960 model shape (W = 120, wpr 2) 27.9 -> 23.2 us; KITTI (W = 160, wpr 3: one
    // block of 12 warps per SM) 74.6 -> 84.1 us and C256 80x160 D41 (wpr 5, 10 warps) 158 -> 270 us stay on the mma.sync kernels.
    // DSSMD_CORR_BWD_FUSED=2 takes it wherever it is eligible.
    int nw = 0, cgw = 0;
    for (int cand : {8, 12, 16, 6, 10, 4, 2}) {
        if (cand % (2 * wpr) == 0) { nw = cand; cgw = cand / (2 * wpr); break; }
    }
    if (nw == 0 || (nw != 8 && env_int_fused("DSSMD_CORR_BWD_FUSED", 1) < 2)) return -1;
    constexpr int CPW = 8;
    p.CGW = cgw;
    p.CB = cgw * CPW;
    p.NCB = ceil_div(C, p.CB);
    if (p.NCB < kFStages - 1) return -1;   // the g tile of row r + 1 must not be requested while row r still reads its own
    p.op_bytes = (uint32_t)((size_t)p.CB * p.WP * 4);
    p.g_bytes = (uint32_t)((size_t)D * p.WP * 4);
    const size_t smem = (size_t)kFStages * 2 * ((p.op_bytes + 127u) & ~127u) + 2 * (size_t)((p.g_bytes + 127u) & ~127u) + 128;
    if (smem > 200 * 1024) return -1;
    p.nrows = B * H;
    CUtensorMap mapL, mapR, mapG;
    {
        const cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)B};
        const cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)C * H * W * 4};
        const cuuint32_t box[4] = {(cuuint32_t)p.WP, 1, (cuuint32_t)p.CB, 1};
        if (!make_tmap_f32(&mapL, L, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return -1;
        if (!make_tmap_f32(&mapR, R, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return -1;
    }
    {
        const cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)B};
        const cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)gbs * 4};
        const cuuint32_t box[4] = {(cuuint32_t)p.WP, 1, (cuuint32_t)D, 1};
        if (!make_tmap_f32(&mapG, g, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return -1;
    }
#define DSSMD_FUSED_NW(LN, NWV) if (nw == NWV) return launch_fused<LN, CPW, NWV>(p, mapL, mapR, mapG, smem, stream);
#define DSSMD_FUSED_LN(LN) DSSMD_FUSED_NW(LN, 8) DSSMD_FUSED_NW(LN, 12) DSSMD_FUSED_NW(LN, 16) DSSMD_FUSED_NW(LN, 6) DSSMD_FUSED_NW(LN, 10) \
                           DSSMD_FUSED_NW(LN, 4) DSSMD_FUSED_NW(LN, 2)
    switch (lognh) {
        case 0: DSSMD_FUSED_LN(0) break;
        case 1: DSSMD_FUSED_LN(1) break;
        default: DSSMD_FUSED_LN(2) break;
    }
#undef DSSMD_FUSED_LN
#undef DSSMD_FUSED_NW
    return -1;
}

}  // namespace dssmd
