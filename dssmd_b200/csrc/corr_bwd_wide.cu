// corr_bwd_wide.cu -- fp32 backward of the correlation volume for the models' own disparity range (13 <= D <= 24: max_disp 192
// at 1/8 resolution), both feature gradients in one launch:
//
//   gL[c,x ] = (1/C) * sum_d g[d,x]    * R[c,x-d]
//   gR[c,x'] = (1/C) * sum_d g[d,x'+d] * L[c,x'+d]
//
// (the autograd graph of build_corr, upstream models/UtilsNet/submoudles.py:301-311).  corr_bwd_fused.cu splits the D
// disparities of a pixel group over two lanes (4 pixels x 12 disparities each) and pays for it on every channel: a butterfly
// of shuffles and adds, half of the lanes not storing, two overlapping 16-word windows where one of 28 words would do --
// ncu: 13.3 M warp instructions of which 7.1 M are the FFMAs.  Here a lane owns 4 pixels x ALL 24 disparities:
//   * the g slab of a row lives in 96 registers (gR: skewed), a channel costs 7 LDS.128 (one aligned 28-word window of the
//     operand row; lane l starts at 16-byte chunk l), 96 FFMAs in four independent chains and ONE predicated float4 store --
//     104 instructions per 128 pixels and channel against 2 x 67 (10.0 M warp instructions); no cross-lane traffic;
//   * that needs ~200 registers, so one block of 8 warps per SM (four channel groups per gradient); what keeps the two warps
//     of a scheduler busy is the instruction-level parallelism of a straight-line channel loop: the store is predicated
//     inside the asm (a branch around it made every channel a basic block of its own, and the window loads of channel k + 1
//     could not start under the FFMAs of channel k: 22.7 -> 20.1 us at the 576x960 model shape);
//   * the warps are decoupled: a ring of TMA stages with a full and an empty mbarrier each (no block barrier in the loop),
//     one elected lane re-arms a stage as soon as all eight warps have released it; an item's two operand tiles (R with a
//     left halo for gL, L with a right halo for gR) and, on a new row, the g tile arrive as TMA boxes whose zero fill outside
//     the row *is* the x - d < 0 / x' + d >= W limit of the sums;
//   * persistent blocks own contiguous ranges of (row, x tile, channel block) items, channel blocks fastest, so a row may be
//     shared by two blocks (320 rows on 148 SMs: 9 items per block instead of 3 whole rows on 107 blocks).
// No atomics, deterministic (fixed summation order d = 0 .. D-1 per output).
// Measured and left out (profiles/r02w_corr_bwd_wide.md): packed fma.rn.f32x2 over pairs of disparities (6.9 M instead of
// 10.0 M warp instructions, same 20 us: the FMA pipe takes two cycles per packed instruction and the dependent-issue stalls grow
// by what the issue slots save), explicit double buffering of the window registers (21.0 us), 16 channels per item, ring depths 2..8.
#include "corr_reg.cuh"

#include <cstdlib>
#include <type_traits>

namespace dssmd {

namespace {

constexpr int kWD = 24;        // disparities per lane (D <= 24)
constexpr int kWideMaxStages = 8;

struct WideParams {
    float *gL, *gR;
    int pf;                   // items of a block whose boxes are prefetched towards the L2 before griddepcontrol.wait; 0 = none
    int C, H, W, D;
    int TX, XT;     // x tile in pixels (multiple of 4, <= 128), tiles per row
    int CB;         // channels per item = NWG * CPW
    int NCB;        // channel blocks per row
    int WP;         // box width in elements (TX + 24), operands and g
    int S;          // ring depth
    int nitems, ipb;   // (row, x tile, channel block) items, items per block
    uint32_t op_bytes, g_bytes;
    float invC;
};

__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done != 0;
}
// bounded wait that the compiler cannot unroll into 64 copies (a protocol bug traps instead of hanging the GPU)
__device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity) {
    for (int spin = 0; spin < (1 << 26); ++spin)
        if (mbar_try(bar, parity)) return;
    __trap();
}
__device__ __forceinline__ void mbar_wait_w(uint32_t bar, uint32_t parity) {
    if (!mbar_try(bar, parity)) mbar_wait_slow(bar, parity);
}

// predicated streaming store without control flow (see the header)
__device__ __forceinline__ void st_global_cs_if(float *p, float a, float b, float c, float d, bool on) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.u32 p, %5, 0;\n\t"
        "@p st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};\n\t}"
        ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d), "r"((unsigned)on) : "memory");
}

template <int CPW, int NWG, int UNR>
__global__ void __launch_bounds__(2 * NWG * 32, 1)
corr_bwd_wide_kernel(const WideParams p, const __grid_constant__ CUtensorMap mapL, const __grid_constant__ CUtensorMap mapR,
                     const __grid_constant__ CUtensorMap mapG) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long s_full[kWideMaxStages], s_empty[kWideMaxStages];
    constexpr int PX = 4, NW = 2 * NWG;
    // stage s: [R box | L box]; after the stages: two g tiles (groups alternate)
    const uint32_t op_stride = (p.op_bytes + 127u) & ~127u, g_stride = (p.g_bytes + 127u) & ~127u;
    const uint32_t stage_bytes = 2u * op_stride;
    const uint32_t g_base = (uint32_t)p.S * stage_bytes;
    const uint32_t sbase = smem_u32(smem_raw);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int i0 = blockIdx.x * p.ipb;
    const int i1 = (i0 + p.ipb < p.nitems) ? i0 + p.ipb : p.nitems;
    if (i0 >= i1) return;
    const int nitems = i1 - i0;

    if (tid == 0) {
        for (int s = 0; s < p.S; ++s) {
            mbar_init(smem_u32(&s_full[s]), 1);
            mbar_init(smem_u32(&s_empty[s]), NW);
        }
        fence_proxy_async();
    }
    __syncthreads();
    pdl_launch_dependents();  // the next kernel on the stream may start its ramp

    // everything that does not touch global memory comes BEFORE griddepcontrol.wait: launched behind another kernel the block does it
    // while that kernel drains
    // first item of this block: channel block, x tile, row
    const int cb0 = i0 % p.NCB, grp0 = i0 / p.NCB;
    const int xt0 = grp0 % p.XT, row0 = grp0 / p.XT;
    const int b0 = row0 / p.H, y0 = row0 - b0 * p.H;

    // producer state (warp 0, lane 0): the next item to request, as counters (no divisions in the loop)
    int is = 0, icb = cb0, ign = 0, ixt = xt0, ib = b0, iy = y0;
    bool ifirst = true;
    auto issue = [&]() {  // one thread; requests the next item into stage `is` and advances
        const int s = is, cb = icb, gn = ign, b = ib, y = iy, X0 = ixt * p.TX;
        const bool with_g = cb == 0 || ifirst;   // a block that starts in the middle of a row needs that row's g tile as well
        ifirst = false;
        if (++is == p.S) is = 0;
        if (++icb == p.NCB) {
            icb = 0; ++ign;
            if (++ixt == p.XT) { ixt = 0; if (++iy == p.H) { iy = 0; ++ib; } }
        }
        const uint32_t bar = smem_u32(&s_full[s]);
        const uint32_t dst = sbase + (uint32_t)s * stage_bytes;
        mbar_arrive_expect_tx(bar, 2u * p.op_bytes + (with_g ? p.g_bytes : 0u));
        tma_load_4d(dst, &mapR, X0 - kWD, y, cb * p.CB, b, bar);         // gL: R[x - d], left halo
        tma_load_4d(dst + op_stride, &mapL, X0, y, cb * p.CB, b, bar);   // gR: L[x' + d], right halo
        if (with_g) tma_load_4d(sbase + g_base + (uint32_t)(gn & 1) * g_stride, &mapG, X0, y, 0, b, bar);
    };

    // this warp's gradient and channel group; this lane's pixels
    const int mode = warp >= NWG ? 1 : 0;
    const int cgw = warp - mode * NWG;
    const int xl = lane * PX;
    const bool lane_on = xl < p.TX;
    const size_t HW = (size_t)p.H * p.W;
    float *const outp = mode == 0 ? p.gL : p.gR;
    // byte offset of this lane's window in this warp's first channel row inside a stage: box column xl for both gradients
    // (gL: the box starts 24 pixels left of the tile, the window covers x - 24 .. x + 3; gR: x' .. x' + 27)
    const uint32_t wp4 = (uint32_t)p.WP * 4u;
    const uint32_t wbase = (uint32_t)mode * op_stride + (uint32_t)(cgw * CPW) * wp4 + (uint32_t)(lane_on ? xl : 0) * 4u;

    if (p.pf != 0 && tid == (int)blockDim.x - 1) {
        // the boxes of this block's first item(s) towards the L2 while the previous kernel drains (common.cuh: prefetch_l2)
        tma_prefetch_desc(&mapL); tma_prefetch_desc(&mapR); tma_prefetch_desc(&mapG);
        int kcb = cb0, kxt = xt0, ky = y0, kb = b0;
        const int np = nitems < p.pf ? nitems : p.pf;
        for (int k = 0; k < np; ++k) {
            tma_prefetch_4d(&mapR, kxt * p.TX - kWD, ky, kcb * p.CB, kb);
            tma_prefetch_4d(&mapL, kxt * p.TX, ky, kcb * p.CB, kb);
            if (k == 0 || kcb == 0) tma_prefetch_4d(&mapG, kxt * p.TX, ky, 0, kb);
            if (++kcb == p.NCB) {
                kcb = 0;
                if (++kxt == p.XT) { kxt = 0; if (++ky == p.H) { ky = 0; ++kb; } }
            }
        }
    }
    pdl_wait();               // no load or store above
    if (tid == 0) {
        for (int k = 0; k < p.S - 1; ++k)
            if (k < nitems) issue();
    }

    float gv[kWD][PX];
    int s = 0, cb = cb0, gn = 0, xt = xt0, b = b0, y = y0;   // consumer state: the item being computed
    uint32_t ph = 0;
    for (int it = 0; it < nitems; ++it) {
        if (tid == 0 && it + p.S - 1 < nitems) {
            // the stage of item it-1 is re-armed once all warps have released it (this warp did so at the end of the last pass)
            if (it > 0) {
                const int sp = s == 0 ? p.S - 1 : s - 1;
                const uint32_t php = s == 0 ? ph ^ 1u : ph;   // parity of the pass item it-1 belonged to
                mbar_wait_w(smem_u32(&s_empty[sp]), php);
            }
            issue();
        }
        __syncwarp();
        mbar_wait_w(smem_u32(&s_full[s]), ph);

        if (cb == 0 || it == 0) {  // new (row, x tile): the g slab, pre-multiplied by 1/C
            const float *gs = reinterpret_cast<const float *>(smem_raw + g_base + (size_t)(gn & 1) * g_stride) + (lane_on ? xl : 0);
#pragma unroll
            for (int d = 0; d < kWD; ++d) {
                const bool dok = d < p.D;
                const float *gd = gs + (dok ? d : 0) * p.WP;
                if (mode == 0) {
                    const float4 v = *reinterpret_cast<const float4 *>(gd);
                    gv[d][0] = dok ? v.x * p.invC : 0.f; gv[d][1] = dok ? v.y * p.invC : 0.f;
                    gv[d][2] = dok ? v.z * p.invC : 0.f; gv[d][3] = dok ? v.w * p.invC : 0.f;
                } else {
#pragma unroll
                    for (int i = 0; i < PX; ++i) gv[d][i] = dok ? gd[i + d] * p.invC : 0.f;
                }
            }
        }

        const unsigned char *wrow = smem_raw + (size_t)s * stage_bytes + wbase;
        const int c0 = cb * p.CB + cgw * CPW;
        const int x = xt * p.TX + xl;
        const bool storer = lane_on && x < p.W;   // W % 4 == 0: all four pixels or none
        const int nk = p.C - c0;                  // channels of this warp that exist (TMA zero-fills the others)
        float *o = outp + ((size_t)b * p.C + c0) * HW + (size_t)y * p.W + x;
        auto channels = [&](auto mode_tag) {   // the warp's gradient is uniform: one branch per item, not per channel
            constexpr int MODE = decltype(mode_tag)::value;
#pragma unroll UNR
            for (int k = 0; k < CPW; ++k) {
                float w[kWD + PX];
#pragma unroll
                for (int q = 0; q < (kWD + PX) / 4; ++q) {
                    const float4 v = *reinterpret_cast<const float4 *>(wrow + 16 * q);
                    w[4 * q + 0] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
                }
                wrow += wp4;
                float acc[PX];
#pragma unroll
                for (int i = 0; i < PX; ++i) acc[i] = gv[0][i] * w[MODE == 0 ? i + kWD : i];
#pragma unroll
                for (int d = 1; d < kWD; ++d)
#pragma unroll
                    for (int i = 0; i < PX; ++i) acc[i] = fmaf(gv[d][i], w[MODE == 0 ? i - d + kWD : i + d], acc[i]);
                st_global_cs_if(o, acc[0], acc[1], acc[2], acc[3], storer && k < nk);
                o += HW;
            }
        };
        if (mode == 0) channels(std::integral_constant<int, 0>{});
        else channels(std::integral_constant<int, 1>{});
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&s_empty[s]));   // this warp is done with stage s (and, on a new group, has its g slab)
        if (++s == p.S) { s = 0; ph ^= 1u; }
        if (++cb == p.NCB) {
            cb = 0; ++gn;
            if (++xt == p.XT) { xt = 0; if (++y == p.H) { y = 0; ++b; } }
        }
    }
}

int env_int_wide(const char *name, int dflt) {
    if (!tuning_enabled()) return dflt;
    const char *v = std::getenv(name);
    return (v && *v) ? std::atoi(v) : dflt;
}

template <int CPW, int NWG, int UNR>
int launch_wide(WideParams p, const CUtensorMap &mapL, const CUtensorMap &mapR, const CUtensorMap &mapG, size_t smem, cudaStream_t stream) {
    auto kern = corr_bwd_wide_kernel<CPW, NWG, UNR>;
    {
        cudaError_t e = kernel_smem_optin(kern, smem);   // asked once per (kernel, device), then a table lookup (abi.cu)
        if (e != cudaSuccess) return fail(DSSMD_ERR_CUDA, "corr_bwd(wide): cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(e));
    }
    // persistent grid: one block per SM, each with an equal share of items
    int grid = sm_count();
    if (grid > p.nitems) grid = p.nitems;
    p.ipb = ceil_div(p.nitems, grid);
    grid = ceil_div(p.nitems, p.ipb);
    if (env_int_wide("DSSMD_CORR_DEBUG", 0))
        fprintf(stderr, "corr_bwd wide cfg: CPW=%d warps=%d TX=%d XT=%d CB=%d NCB=%d WP=%d S=%d items=%d ipb=%d grid=%d smem=%zu\n",
                CPW, 2 * NWG, p.TX, p.XT, p.CB, p.NCB, p.WP, p.S, p.nitems, p.ipb, grid, smem);
    const cudaError_t le = launch_pdl(kern, dim3((unsigned)grid), dim3((unsigned)(2 * NWG * 32)), smem, stream, p, mapL, mapR, mapG);
    if (le != cudaSuccess) return fail(DSSMD_ERR_CUDA, "corr_bwd(wide): launch failed: %s", cudaGetErrorString(le));
    return check_launch("corr_bwd(wide)");
}

}  // namespace

// Both gradients of an fp32 correlation with 13 <= D <= 24 in one launch.  -1 when the shape is not eligible (the caller goes on
// to the other kernels), else a DSSMD_* code.  gbs = elements between two images of g (the concat buffer's stride for
// corr_bwd_cat).  Taken by default for rows of 80 .. 128 pixels (one x tile with at least 20 of a warp's 32 lanes busy):
// measured against the kernels it replaces -- 576x960 model shape (W = 120) 23.1 -> 20.1 us, 320x640 (W = 80) 30.0 -> 26.7 us;
// KITTI (W = 160, two tiles of 80 pixels) at batch 16 74.7 -> 90.3 us stays on the mma.sync kernels, at batch 1 .. 4 it is taken
// (batch 1: 16.9 -> 9.9 us).  DSSMD_CORR_BWD_WIDE=2 takes it
// wherever it is eligible, =0 never.
int corr_bwd_wide_f32(const void *g, const void *L, const void *R, void *gL, void *gR, int B, int C, int H, int W, int D, long long gbs,
                      cudaStream_t stream) {
    const int want = env_int_wide("DSSMD_CORR_BWD_WIDE", 1);
    if (!gL || !gR || want == 0) return -1;
    if (W % 4 != 0 || D > kWD || D < 13) return -1;   // D <= 12: half of the slab would be zeros
    if (((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(L) | reinterpret_cast<uintptr_t>(R) | reinterpret_cast<uintptr_t>(gL) |
          reinterpret_cast<uintptr_t>(gR)) & 15) != 0) return -1;
    if ((gbs * 4) % 16 != 0 || (long long)B * H * 64 >= (1LL << 30)) return -1;
    WideParams p{};
    p.gL = (float *)gL; p.gR = (float *)gR;
    p.pf = pdl_prefetch(1);
    p.C = C; p.H = H; p.W = W; p.D = D;
    p.invC = 1.0f / (float)C;
    p.XT = ceil_div(W, 128);
    p.TX = round_up(ceil_div(W, p.XT), 4);
    // several x tiles per row (KITTI, W = 160: two tiles of 80 pixels): only while the problem is small -- measured at C128 48x160 D24
    // over the batch size (scripts/kitti_b_sweep.py), wide vs the two mma.sync launches: B = 1 10.4 / 20.2 us, B = 4 26.1 / 30.6,
    // B = 6 33.5 / 32.8, B = 16 82.9 / 74.5
    if (want < 2 && (p.TX < 80 || (p.XT > 1 && (long)B * H * p.XT > 3L * sm_count()))) return -1;
    p.WP = p.TX + kWD;
    constexpr int NWG = 4;
    int cpw = env_int_wide("DSSMD_CORR_BWD_WIDE_CPW", 8);
    if (cpw != 4 && cpw != 8) cpw = 8;
    if (C < NWG * cpw) cpw = 4;
    if (C < NWG * cpw) return -1;
    p.CB = NWG * cpw;
    p.NCB = ceil_div(C, p.CB);
    p.op_bytes = (uint32_t)((size_t)p.CB * p.WP * 4);
    p.g_bytes = (uint32_t)((size_t)D * p.WP * 4);
    const size_t stage = 2 * (size_t)((p.op_bytes + 127u) & ~127u), gt = 2 * (size_t)((p.g_bytes + 127u) & ~127u);
    int S = env_int_wide("DSSMD_CORR_BWD_WIDE_STAGES", 0);
    if (S < 2) S = (int)((200 * 1024 - gt - 128) / stage);
    if (S > kWideMaxStages) S = kWideMaxStages;
    // the g tile of group r + 2 reuses the buffer of group r: it is requested S - 1 items ahead, after every warp has released the
    // item S items back, which must not precede the first item of group r (where the warps read its slab) -- a block's first group
    // may consist of a single item, so S <= NCB + 1
    if (S > p.NCB + 1) S = p.NCB + 1;
    if (S < 2) return -1;
    p.S = S;
    const size_t smem = (size_t)S * stage + gt + 128;
    if (smem > 200 * 1024) return -1;
    p.nitems = B * H * p.XT * p.NCB;
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
    const int unr = env_int_wide("DSSMD_CORR_BWD_WIDE_UNR", 4);
    if (cpw == 8) {
        if (unr == 2) return launch_wide<8, NWG, 2>(p, mapL, mapR, mapG, smem, stream);
        if (unr == 8) return launch_wide<8, NWG, 8>(p, mapL, mapR, mapG, smem, stream);
        return launch_wide<8, NWG, 4>(p, mapL, mapR, mapG, smem, stream);
    }
    return launch_wide<4, NWG, 4>(p, mapL, mapR, mapG, smem, stream);
}

}  // namespace dssmd
