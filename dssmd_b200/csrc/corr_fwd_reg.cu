// corr_fwd_reg.cu -- forward of the correlation volume with the outputs held in registers and the
// feature tiles staged by TMA (fp32 / bf16 / fp16 tensors, fp32 arithmetic):
//
//   out[b,d,y,x] = (1/C) * sum_c L[b,c,y,x] * R[b,c,y,x-d]      (0 where x < d)
//
// Replaces build_corr, upstream models/UtilsNet/submoudles.py:301-311.  Same blocking as the backward
// (corr_bwd_reg.cu): a lane owns 4 pixels x 12 disparities -- here as 48 ACCUMULATORS -- and walks over
// the channels; per channel it reads its 4 L values and ONE aligned 16-element window of the R row from
// shared memory and issues 48 FFMAs.  The D disparities are split in parts of 12 over NH lanes (different
// outputs: no reduction across lanes), the lane map keeps the window loads free of bank conflicts.
// The channels of a block are split over CGW warps (a row of a 1/8-resolution feature map is too short to
// fill an SM otherwise); their partial sums are combined through shared memory in a fixed order.
//
// Work item = (row (b,y), x tile, block of CB channels), channel blocks fastest; a persistent block walks
// whole (row, x tile) groups.  Both tiles of an item arrive as one TMA box each from [W,H,C,B] views; the R
// box starts HALO pixels left of the tile and TMA's zero fill of x - d < 0 *is* the reference's zero border.
// Two stages: the boxes of item i+1 are in flight while item i is computed.
#include "corr_reg.cuh"

#include <cstdlib>

namespace dssmd {

namespace {

struct FwdRegParams {
    void *out;
    int pf;             // items of a block whose boxes are prefetched towards the L2 before griddepcontrol.wait; 0 = none
    long long OBS;  // elements between two images of `out`
    int C, H, W, D;
    int NH;         // lanes per pixel group (power of two), D <= 12*NH
    int WPR;        // warps side by side along x
    int CGW;        // channel groups of warps
    int CB;         // channels per item
    int NCB;        // channel blocks
    int TX;         // x tile in pixels
    int XT;         // x tiles per row
    int HALO;       // left halo of the R box in elements (12*NH rounded up to a multiple of 16 bytes)
    int WP;         // R box width in elements (TX + HALO)
    int ngroups, gpb;  // (row, x tile) groups, groups per block
    uint32_t l_bytes, r_bytes, stage_bytes;
    float invC;
};

template <typename T, int LOGNH, int CPW>
__global__ void __launch_bounds__(512) corr_fwd_reg_kernel(const FwdRegParams p, const __grid_constant__ CUtensorMap mapL,
                                                           const __grid_constant__ CUtensorMap mapR) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long s_bar[2];
    constexpr int PX = 4, NH = 1 << LOGNH;
    using LM = LaneMap<LOGNH, PX>;
    const uint32_t l_stride = (p.l_bytes + 127u) & ~127u;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g0 = blockIdx.x * p.gpb;
    const int g1 = (g0 + p.gpb < p.ngroups) ? g0 + p.gpb : p.ngroups;
    if (g0 >= g1) return;
    const int i0 = g0 * p.NCB, i1 = g1 * p.NCB;

    if (tid == 0) {
        mbar_init(smem_u32(&s_bar[0]), 1);
        mbar_init(smem_u32(&s_bar[1]), 1);
        fence_proxy_async();
    }
    __syncthreads();
    pdl_launch_dependents();  // the next kernel on the stream may start its ramp

    // Everything that does not touch global memory comes BEFORE griddepcontrol.wait: launched behind another kernel the block
    // does it while that kernel drains (ncu: ~650 instructions per warp of set-up, a quarter of what a warp executes here).
    // First group of this block (a block owns whole groups, so its first item is channel block 0); the loop below keeps
    // (channel block, x tile, row, image) as counters: the divisions of an item index by run-time values cost ~25 instructions
    // each, several per item and warp.
    const int xt0 = g0 % p.XT, row0 = g0 / p.XT;
    const int b0 = row0 / p.H, y0 = row0 - b0 * p.H;
    int icb = 0, ixt = xt0, iy = y0, ib = b0;   // producer: the next item to request
    auto issue = [&](int s) {  // one thread; requests the next item into stage s and advances
        const uint32_t bar = smem_u32(&s_bar[s]);
        const uint32_t dst = smem_u32(smem_raw) + (uint32_t)s * p.stage_bytes;
        fence_proxy_async();  // the stage was read / written through the generic proxy before the block barrier
        mbar_arrive_expect_tx(bar, p.l_bytes + p.r_bytes);
        const int X0 = ixt * p.TX;
        tma_load_4d(dst, &mapL, X0, iy, icb * p.CB, ib, bar);
        tma_load_4d(dst + l_stride, &mapR, X0 - p.HALO, iy, icb * p.CB, ib, bar);
        if (++icb == p.NCB) {
            icb = 0;
            if (++ixt == p.XT) { ixt = 0; if (++iy == p.H) { iy = 0; ++ib; } }
        }
    };

    // this lane's pixels (inside the x tile) and disparity part
    const int wseg = warp % p.WPR, cgw = warp / p.WPR;
    const int h = LM::h(lane), pgl = LM::pgl(lane);
    const int xl = (wseg * (32 >> LOGNH) + pgl) * PX;       // < TX by construction
    const int woff = xl + p.HALO - kDP * (h + 1);           // window start inside a staged R row
    const size_t HW = (size_t)p.H * p.W;

    if (p.pf != 0 && tid == (int)blockDim.x - 1) {
        // the boxes of this block's first item(s) towards the L2 while the previous kernel drains (common.cuh: prefetch_l2)
        tma_prefetch_desc(&mapL); tma_prefetch_desc(&mapR);
        int kcb = 0, kxt = xt0, ky = y0, kb = b0;
        const int np = (i1 - i0 < p.pf) ? i1 - i0 : p.pf;
        for (int k = 0; k < np; ++k) {
            tma_prefetch_4d(&mapL, kxt * p.TX, ky, kcb * p.CB, kb);
            tma_prefetch_4d(&mapR, kxt * p.TX - p.HALO, ky, kcb * p.CB, kb);
            if (++kcb == p.NCB) {
                kcb = 0;
                if (++kxt == p.XT) { kxt = 0; if (++ky == p.H) { ky = 0; ++kb; } }
            }
        }
    }
    pdl_wait();               // no load or store above
    if (tid == 0) issue(0);

    float acc[kDP][PX];
    int cb = 0, xt = xt0, y = y0, b = b0;   // consumer: the item being computed
    for (int it = i0; it < i1; ++it) {
        const int s = (it - i0) & 1;
        const uint32_t ph = ((it - i0) >> 1) & 1u;
        if (tid == 0 && it + 1 < i1) issue(s ^ 1);  // stage s^1 was released by the barrier ending item it-1
        unsigned char *stage = smem_raw + (size_t)s * p.stage_bytes;
        const T *lt = reinterpret_cast<const T *>(stage);
        const T *rt = reinterpret_cast<const T *>(stage + l_stride);
        if (cb == 0) {
#pragma unroll
            for (int dd = 0; dd < kDP; ++dd)
#pragma unroll
                for (int i = 0; i < PX; ++i) acc[dd][i] = 0.f;
        }
        mbar_wait(smem_u32(&s_bar[s]), ph);

        const int c0 = cb * p.CB;
        const int cn = (p.C - c0 < p.CB) ? p.C - c0 : p.CB;
        const T *lrow = lt + cgw * p.TX + xl;
        const T *wrow = rt + cgw * p.WP + woff;
        const int lstep = p.CGW * p.TX, wstep = p.CGW * p.WP;
        auto channel = [&](const T *lr, const T *wr) {
            float l[PX], w[PX + kDP];
            load4<T>(lr, l[0], l[1], l[2], l[3]);
#pragma unroll
            for (int q = 0; q < (PX + kDP) / 4; ++q) load4<T>(wr + 4 * q, w[4 * q + 0], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
#pragma unroll
            for (int dd = 0; dd < kDP; ++dd)
#pragma unroll
                for (int i = 0; i < PX; ++i) acc[dd][i] = fmaf(l[i], w[i - dd + kDP], acc[dd][i]);
        };
        if (CPW > 0 && cn == p.CB) {
            // full item with a compile-time channel count per warp: straight-line code, the window loads of channel k + 1 start under
            // the FFMAs of channel k (with a run-time trip count the loads of every iteration wait at the top of the loop body)
#pragma unroll
            for (int k = 0; k < CPW; ++k) channel(lrow + k * lstep, wrow + k * wstep);
        } else {
#pragma unroll 2
            for (int c = cgw; c < cn; c += p.CGW, lrow += lstep, wrow += wstep) channel(lrow, wrow);
        }

        if (cb == p.NCB - 1) {
            // all channels of this (row, x tile) are in: combine the CGW partial sums and write the 12 x 4 outputs
            const int x = xt * p.TX + xl;
            T *o = reinterpret_cast<T *>(p.out) + (size_t)b * p.OBS + (size_t)y * p.W + x;
            if (p.CGW == 1) {
                if (x < p.W) {
#pragma unroll
                    for (int dd = 0; dd < kDP; ++dd) {
                        const int d = h * kDP + dd;
                        if (d < p.D) {
                            const float v[4] = {acc[dd][0] * p.invC, acc[dd][1] * p.invC, acc[dd][2] * p.invC, acc[dd][3] * p.invC};
                            store_n<T, 4>(o + (size_t)d * HW, v);
                        }
                    }
                }
            } else {
                // partials through shared memory (the current stage is free once everyone has finished reading it):
                // red[wseg][cgw][dd][lane] as float4; warp cgw then sums rows dd = cgw, cgw + CGW, ... over the
                // sources 0 .. CGW-1 in that order (deterministic) and stores them
                __syncthreads();
                float4 *red = reinterpret_cast<float4 *>(stage);
                float4 *mine = red + ((size_t)(wseg * p.CGW + cgw) * kDP) * 32 + lane;
#pragma unroll
                for (int dd = 0; dd < kDP; ++dd) mine[dd * 32] = make_float4(acc[dd][0], acc[dd][1], acc[dd][2], acc[dd][3]);
                __syncthreads();
                const float4 *src = red + ((size_t)(wseg * p.CGW) * kDP) * 32 + lane;
                for (int dd = cgw; dd < kDP; dd += p.CGW) {
                    float4 a = src[dd * 32];
                    for (int k = 1; k < p.CGW; ++k) {
                        const float4 t = src[((size_t)k * kDP + dd) * 32];
                        a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
                    }
                    const int d = h * kDP + dd;
                    if (d < p.D && x < p.W) {
                        const float v[4] = {a.x * p.invC, a.y * p.invC, a.z * p.invC, a.w * p.invC};
                        store_n<T, 4>(o + (size_t)d * HW, v);
                    }
                }
            }
        }
        __syncthreads();  // everyone is done with stage s
        if (++cb == p.NCB) {
            cb = 0;
            if (++xt == p.XT) { xt = 0; if (++y == p.H) { y = 0; ++b; } }
        }
    }
}

int env_int_fwd(const char *name, int dflt) {
    if (!tuning_enabled()) return dflt;
    const char *v = std::getenv(name);
    return (v && *v) ? std::atoi(v) : dflt;
}

template <typename T, int LOGNH, int CPW>
int launch_fwd_reg_cpw(FwdRegParams p, const CUtensorMap &mapL, const CUtensorMap &mapR, int threads, size_t smem, cudaStream_t stream) {
    auto kern = corr_fwd_reg_kernel<T, LOGNH, CPW>;
    {
        cudaError_t e = kernel_smem_optin(kern, smem);
        if (e != cudaSuccess) return fail(DSSMD_ERR_CUDA, "corr_fwd: cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(e));
    }
    // persistent grid: the blocks that are resident at once, each with an equal share of (row, x tile) groups
    const int bps = kernel_occupancy(kern, threads, smem);
    int grid = sm_count() * bps;
    if (grid > p.ngroups) grid = p.ngroups;
    p.gpb = ceil_div(p.ngroups, grid);
    grid = ceil_div(p.ngroups, p.gpb);
    if (env_int_fwd("DSSMD_CORR_DEBUG", 0))
        fprintf(stderr, "corr_fwd reg cfg: es=%d NH=%d WPR=%d TX=%d XT=%d CGW=%d CB=%d NCB=%d groups=%d gpb=%d grid=%d (x%d/SM) threads=%d smem=%zu\n",
                (int)sizeof(T), p.NH, p.WPR, p.TX, p.XT, p.CGW, p.CB, p.NCB, p.ngroups, p.gpb, grid, bps, threads, smem);
    const cudaError_t le = launch_pdl(kern, dim3((unsigned)grid), dim3((unsigned)threads), smem, stream, p, mapL, mapR);
    if (le != cudaSuccess) return fail(DSSMD_ERR_CUDA, "corr_fwd(reg): launch failed: %s", cudaGetErrorString(le));
    return check_launch("corr_fwd(reg)");
}

// channels per warp and item as a compile-time constant where it is 8 or 4 (fp32 model shapes: 32 channels per item over 4
// channel groups); 0 = run-time loop
template <typename T, int LOGNH>
int launch_fwd_reg(FwdRegParams p, const CUtensorMap &mapL, const CUtensorMap &mapR, int threads, size_t smem, cudaStream_t stream) {
    if constexpr (sizeof(T) == 4) {
        if (env_int_fwd("DSSMD_CORR_FWD_UNROLLED", 1) != 0) {
            if (p.CB == 8 * p.CGW) return launch_fwd_reg_cpw<T, LOGNH, 8>(p, mapL, mapR, threads, smem, stream);
            if (p.CB == 4 * p.CGW) return launch_fwd_reg_cpw<T, LOGNH, 4>(p, mapL, mapR, threads, smem, stream);
        }
    }
    return launch_fwd_reg_cpw<T, LOGNH, 0>(p, mapL, mapR, threads, smem, stream);
}

template <typename T>
int corr_fwd_reg_typed(const void *L, const void *R, void *out, int B, int C, int H, int W, int D, long long obs, CUtensorMapDataType tmtype, cudaStream_t stream) {
    constexpr int ES = (int)sizeof(T);
    constexpr int EPV = 16 / ES;  // elements per 16 bytes: rows and boxes are multiples of it
    if (W % EPV != 0 || D > kDP * 16) return -1;
    if (((reinterpret_cast<uintptr_t>(L) | reinterpret_cast<uintptr_t>(R) | reinterpret_cast<uintptr_t>(out)) & 15) != 0) return -1;
    if ((long long)B * H * 64 >= (1LL << 31)) return -1;
    FwdRegParams p{};
    p.out = out; p.OBS = obs;
    p.pf = pdl_prefetch(1);
    p.C = C; p.H = H; p.W = W; p.D = D;
    p.invC = 1.0f / (float)C;
    int nh = 1, lognh = 0;
    while (nh * kDP < D) { nh *= 2; ++lognh; }
    p.NH = nh;
    p.HALO = round_up(kDP * nh, EPV);
    const int pxw = (32 / nh) * 4;  // pixels one warp covers
    int wpr = ceil_div(W, pxw);
    while (wpr > 1 && wpr * pxw + p.HALO > 256) wpr = (wpr + 1) / 2;  // the R box must fit a 256-element TMA box
    if (wpr * pxw + p.HALO > 256 || wpr > 16) return -1;
    p.WPR = wpr;
    p.TX = wpr * pxw;
    p.XT = ceil_div(W, p.TX);
    p.WP = p.TX + p.HALO;
    p.ngroups = B * H * p.XT;
    // fp32 rows that leave more than a quarter of the lanes idle are better served by the narrow-tile kernel
    if (ES == 4 && env_int_fwd("DSSMD_CORR_FWD_REG", 1) < 2 && (double)W < 0.75 * (double)(p.XT * p.TX)) return -1;
    // channel split: ~8 warps per block, more when there are too few (row, x tile) groups to fill the SMs twice
    int cgw = env_int_fwd("DSSMD_CORR_FWD_CGW", 0);
    if (cgw < 1) {
        cgw = ceil_div(8, wpr);
        if ((long)p.ngroups < 2L * sm_count() && cgw * wpr < 16) cgw = 16 / wpr;
    }
    while (cgw > 1 && wpr * cgw > 16) --cgw;
    if (cgw > 12) cgw = 12;
    if (cgw > C) cgw = C;
    // channels per stage: two stages (+ the partial-sum buffer, which aliases a stage) for two resident blocks
    const size_t budget = (size_t)env_int_fwd("DSSMD_CORR_FWD_SMEM_KB2", 100) * 1024 / 2;
    while (cgw > 1 && (size_t)wpr * cgw * kDP * 32 * 16 + 256 > budget) --cgw;
    p.CGW = cgw;
    const size_t red_bytes = cgw > 1 ? (size_t)wpr * cgw * kDP * 32 * 16 : 0;
    if (red_bytes + 256 > budget) return -1;
    const size_t chb = (size_t)(p.TX + p.WP) * ES;   // bytes per channel and stage
    int cb = env_int_fwd("DSSMD_CORR_FWD_CB", 0);
    if (cb < 1) cb = 32;
    cb = round_up(cb, cgw);
    while (cb > cgw && chb * cb + 256 > budget) cb -= cgw;
    if (chb * cb + 256 > budget) return -1;
    if (cb > 256) cb = 256 / cgw * cgw;
    if (cb > round_up(C, cgw)) cb = round_up(C, cgw);
    p.CB = cb;
    p.NCB = ceil_div(C, cb);
    p.l_bytes = (uint32_t)((size_t)cb * p.TX * ES);
    p.r_bytes = (uint32_t)((size_t)cb * p.WP * ES);
    size_t stage = (size_t)((p.l_bytes + 127u) & ~127u) + ((p.r_bytes + 127u) & ~127u);
    if (stage < red_bytes) stage = (red_bytes + 127) & ~(size_t)127;
    p.stage_bytes = (uint32_t)stage;
    const size_t smem = 2 * stage + 128;
    if (smem > 200 * 1024) return -1;
    const int threads = wpr * cgw * 32;
    CUtensorMap mapL, mapR;
    const cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)B};
    const cuuint64_t strides[3] = {(cuuint64_t)W * ES, (cuuint64_t)H * W * ES, (cuuint64_t)C * H * W * ES};
    {
        const cuuint32_t box[4] = {(cuuint32_t)p.TX, 1, (cuuint32_t)cb, 1};
        if (!make_tmap(&mapL, tmtype, L, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return -1;
    }
    {
        const cuuint32_t box[4] = {(cuuint32_t)p.WP, 1, (cuuint32_t)cb, 1};
        if (!make_tmap(&mapR, tmtype, R, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return -1;
    }
    switch (lognh) {
        case 0: return launch_fwd_reg<T, 0>(p, mapL, mapR, threads, smem, stream);
        case 1: return launch_fwd_reg<T, 1>(p, mapL, mapR, threads, smem, stream);
        case 2: return launch_fwd_reg<T, 2>(p, mapL, mapR, threads, smem, stream);
        case 3: return launch_fwd_reg<T, 3>(p, mapL, mapR, threads, smem, stream);
        default: return launch_fwd_reg<T, 4>(p, mapL, mapR, threads, smem, stream);
    }
}

}  // namespace

// dtype: DSSMD_F32 / BF16 / F16 (all tensors).  Returns -1 when the shape is not eligible (the caller falls back to the
// cp.async kernels), else a DSSMD_* code.
int corr_fwd_reg(const void *L, const void *R, void *out, int B, int C, int H, int W, int D, long long obs, int dtype, cudaStream_t stream) {
    switch (dtype) {
        case DSSMD_F32: return corr_fwd_reg_typed<float>(L, R, out, B, C, H, W, D, obs, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, stream);
        case DSSMD_BF16: return corr_fwd_reg_typed<__nv_bfloat16>(L, R, out, B, C, H, W, D, obs, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, stream);
        case DSSMD_F16: return corr_fwd_reg_typed<__half>(L, R, out, B, C, H, W, D, obs, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, stream);
        default: return -1;
    }
}

}  // namespace dssmd
