// corr_bwd_reg.cu -- fp32 backward of the correlation volume with the gradient slab held
// in registers and the tiles staged by TMA (fast path of corr_bwd.cu; same maths, see the
// formulas there).
//
//   gL[c,x ] = (1/C) * sum_d g[d,x]    * R[c,x-d]        (MODE 0)
//   gR[c,x'] = (1/C) * sum_d g[d,x'+d] * L[c,x'+d]       (MODE 1)
//
// For one image row the g values a pixel needs do not depend on the channel, the operand
// values do.  So a lane keeps a PX-pixel x 12-disparity slab of g in registers (MODE 1:
// skewed, gs[d][x'] = g[d][x'+d]) and walks over the channels; per channel it reads ONE
// aligned (PX+12)-word window of the operand row from shared memory and issues PX*12 FFMAs
// -- 12 (PX=4) or 19 (PX=8) FFMAs per LDS.128, against the 4 an LDS.128 (4 shared-memory
// wavefronts) has to feed to keep the FMA pipe busy.  The D disparities are split in NH parts
// of 12 over neighbouring lanes (lane = group*NH + part); the parts are summed with log2(NH)
// shuffles in a fixed order (deterministic, no atomics).
//
// Work item = (row (b,y), x tile, block of CB channels).  A persistent block walks a contiguous
// range of items, channel blocks fastest, so the g slab stays in registers while the row and x
// tile do not change.  Both tiles of an item arrive by ONE TMA box each ([W,H,C,B] views of the
// operand and of g): the box starts 12*NH pixels left of the tile (MODE 0) or extends 12*NH
// pixels to its right (MODE 1) and TMA zero-fills what falls outside the row, which is exactly
// the x-d < 0 / x'+d >= W boundary of the sums -- the inner loop has no bounds tests.  Two
// stages: the boxes of item i+1 are in flight while item i is computed.
#include "corr_reg.cuh"

#include <cstdlib>

namespace dssmd {

namespace {


struct RegParams {
    void *out;
    int C, H, W, D;
    int NH, LOGNH;  // lanes per pixel group (power of two), D <= 12*NH
    int WPR;        // warps side by side along x
    int CGW;        // channel groups of warps
    int CB;         // channels per item
    int NCB;        // channel blocks
    int TX;         // x tile in pixels
    int XT;         // x tiles per row
    int HALO;       // halo of the operand box in elements: 12*NH, rounded up so that the box is a multiple of 16 bytes
    int WP;         // operand box width in elements (TX + HALO)
    int GW;         // g box width in elements (MODE 0: TX, MODE 1: WP)
    int nitems, ipb;  // work items, items per block
    uint32_t op_bytes, g_bytes;
    float invC;
};

template <typename T, int MODE, int LOGNH, int PX>
__global__ void __launch_bounds__(PX == 4 ? 512 : 256) corr_bwd_reg_kernel(const RegParams p, const __grid_constant__ CUtensorMap mapIn,
                                                                            const __grid_constant__ CUtensorMap mapG) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long s_bar[2];
    // stage s: operand tile [CB][WP] then g tile [D][GW], each 128-byte aligned
    const uint32_t op_stride = (p.op_bytes + 127u) & ~127u, g_stride = (p.g_bytes + 127u) & ~127u;
    const uint32_t stage_bytes = op_stride + g_stride;

    constexpr int NH = 1 << LOGNH;
    using LM = LaneMap<LOGNH, PX>;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int i0 = blockIdx.x * p.ipb;
    const int i1 = (i0 + p.ipb < p.nitems) ? i0 + p.ipb : p.nitems;
    if (i0 >= i1) return;

    if (tid == 0) {
        mbar_init(smem_u32(&s_bar[0]), 1);
        mbar_init(smem_u32(&s_bar[1]), 1);
        fence_proxy_async();
    }
    __syncthreads();
    pdl_launch_dependents();  // the next kernel on the stream may start its ramp
    pdl_wait();               // nothing above touches global memory

    auto issue = [&](int it, int s) {  // one thread
        const int cb = it % p.NCB, rx = it / p.NCB;
        const int xt = rx % p.XT, row = rx / p.XT;
        const int b = row / p.H, y = row - b * p.H;
        const bool newg = (it == i0) || (cb == 0);
        const uint32_t bar = smem_u32(&s_bar[s]);
        const uint32_t dst = smem_u32(smem_raw) + (uint32_t)s * stage_bytes;
        fence_proxy_async();  // the stage was read through the generic proxy before the block barrier
        mbar_arrive_expect_tx(bar, p.op_bytes + (newg ? p.g_bytes : 0u));
        const int X0 = xt * p.TX;
        tma_load_4d(dst, &mapIn, MODE == 0 ? X0 - p.HALO : X0, y, cb * p.CB, b, bar);
        if (newg) tma_load_4d(dst + op_stride, &mapG, X0, y, 0, b, bar);
    };
    if (tid == 0) issue(i0, 0);

    // this lane's pixels (inside the x tile) and disparity part
    const int wseg = warp % p.WPR, cgw = warp / p.WPR;
    const int h = LM::h(lane), pgl = LM::pgl(lane);
    const int xl = (wseg * (32 >> LOGNH) + pgl) * PX;  // < TX by construction
    const int woff = MODE == 0 ? xl + p.HALO - kDP * (h + 1) : xl + kDP * h;  // window start inside a staged row
    const size_t HW = (size_t)p.H * p.W;

    float gv[kDP][PX];
    for (int it = i0; it < i1; ++it) {
        const int s = (it - i0) & 1;
        const uint32_t ph = ((it - i0) >> 1) & 1u;
        if (tid == 0 && it + 1 < i1) issue(it + 1, s ^ 1);  // stage s^1 was released by the barrier ending item it-1
        const int cb = it % p.NCB, rx = it / p.NCB;
        const int xt = rx % p.XT, row = rx / p.XT;
        const int b = row / p.H, y = row - b * p.H;
        const int X0 = xt * p.TX;
        const T *ops = reinterpret_cast<const T *>(smem_raw + (size_t)s * stage_bytes);
        mbar_wait(smem_u32(&s_bar[s]), ph);

        if (it == i0 || cb == 0) {  // new row / x tile: reload the g slab
            const T *gs = reinterpret_cast<const T *>(smem_raw + (size_t)s * stage_bytes + op_stride);
#pragma unroll
            for (int dd = 0; dd < kDP; ++dd) {
                const int d = h * kDP + dd;
                const bool dok = d < p.D;
                const T *gd = gs + (dok ? d : 0) * p.GW + xl;
                if (MODE == 0) {
#pragma unroll
                    for (int q = 0; q < PX / 4; ++q) {
                        float a, b2, c2, d2;
                        load4<T>(gd + 4 * q, a, b2, c2, d2);
                        gv[dd][4 * q + 0] = dok ? a * p.invC : 0.f; gv[dd][4 * q + 1] = dok ? b2 * p.invC : 0.f;
                        gv[dd][4 * q + 2] = dok ? c2 * p.invC : 0.f; gv[dd][4 * q + 3] = dok ? d2 * p.invC : 0.f;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < PX; ++i) gv[dd][i] = dok ? to_f32<T>(gd[i + d]) * p.invC : 0.f;
                }
            }
        }

        const int c0 = cb * p.CB;
        const int cn = (p.C - c0 < p.CB) ? p.C - c0 : p.CB;
        const int x = X0 + xl;
        // The NH partial sums of a pixel group are combined by a reduce-scatter: in round r the lanes that
        // differ in bit r of h exchange halves, so after min(LOGNH, log2 PX) rounds a lane owns PX >> rounds
        // finished pixels (at offset `xo` inside the group); further rounds (NH > PX) are a butterfly on the
        // single remaining value and only the lanes whose extra h bits are 0 store.  Fixed order: deterministic.
        constexpr int LOGPX = PX == 8 ? 3 : 2;
        constexpr int RS = LOGNH < LOGPX ? LOGNH : LOGPX;   // reduce-scatter rounds
        constexpr int SPL = PX >> RS;                        // finished pixels per lane
        int xo = 0;
        bool storer = true;
#pragma unroll
        for (int r = 0; r < LOGNH; ++r) {
            const int bit = (h >> r) & 1;
            if (r < RS) xo += bit * (PX >> (r + 1));
            else storer = storer && !bit;
        }
        storer = storer && (x + xo < p.W);   // W % 4 == 0 and x + xo is a multiple of SPL <= 4: all or nothing
        T *o = reinterpret_cast<T *>(p.out) + ((size_t)b * p.C + c0 + cgw) * HW + (size_t)y * p.W + x + xo;
        const size_t ostep = (size_t)p.CGW * HW;
        const T *wrow = ops + cgw * p.WP + woff;
        const int wstep = p.CGW * p.WP;
#pragma unroll 2
        for (int c = cgw; c < cn; c += p.CGW, o += ostep, wrow += wstep) {
            float w[PX + kDP];
#pragma unroll
            for (int q = 0; q < (PX + kDP) / 4; ++q) load4<T>(wrow + 4 * q, w[4 * q + 0], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
            float acc[PX];
            if constexpr (PX == 4) {  // two chains per pixel: a warp then has 8 (16 with the unroll) independent FFMAs in flight
                float acb[PX];
#pragma unroll
                for (int i = 0; i < PX; ++i) {
                    acc[i] = gv[0][i] * w[MODE == 0 ? i + kDP : i];
                    acb[i] = gv[1][i] * w[MODE == 0 ? i - 1 + kDP : i + 1];
                }
#pragma unroll
                for (int dd = 2; dd < kDP; dd += 2)
#pragma unroll
                    for (int i = 0; i < PX; ++i) {
                        acc[i] = fmaf(gv[dd][i], w[MODE == 0 ? i - dd + kDP : i + dd], acc[i]);
                        acb[i] = fmaf(gv[dd + 1][i], w[MODE == 0 ? i - dd - 1 + kDP : i + dd + 1], acb[i]);
                    }
#pragma unroll
                for (int i = 0; i < PX; ++i) acc[i] += acb[i];
            } else {
#pragma unroll
                for (int i = 0; i < PX; ++i) acc[i] = 0.f;
#pragma unroll
                for (int dd = 0; dd < kDP; ++dd)
#pragma unroll
                    for (int i = 0; i < PX; ++i)
                        acc[i] = fmaf(gv[dd][i], w[MODE == 0 ? i - dd + kDP : i + dd], acc[i]);
            }
#pragma unroll
            for (int r = 0; r < LOGNH; ++r) {
                const unsigned m = 1u << LM::hpos(r);
                if (r < RS) {
                    const int n2 = PX >> (r + 1);
                    const bool bit = (h >> r) & 1;
#pragma unroll
                    for (int i = 0; i < n2; ++i) {
                        const float send = bit ? acc[i] : acc[i + n2];
                        const float keep = bit ? acc[i + n2] : acc[i];
                        acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, m);
                    }
                } else {
                    acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], m);
                }
            }
            if (storer) {
                if constexpr (SPL == 8) {
                    store_n<T, 4>(o, acc);
                    store_n<T, 4>(o + 4, acc + 4);
                } else {
                    store_n<T, SPL>(o, acc);
                }
            }
        }
        __syncthreads();  // everyone is done with stage s (and, on a new row, has its g slab)
    }
}

int env_int_reg(const char *name, int dflt) {
    if (!tuning_enabled()) return dflt;
    const char *v = std::getenv(name);
    return (v && *v) ? std::atoi(v) : dflt;
}

template <typename T, int MODE, int LOGNH, int PX>
int launch_reg(RegParams p, const CUtensorMap &mapIn, const CUtensorMap &mapG, int threads, size_t smem, cudaStream_t stream) {
    auto kern = corr_bwd_reg_kernel<T, MODE, LOGNH, PX>;
    {
        cudaError_t e = kernel_smem_optin(kern, smem);
        if (e != cudaSuccess) return fail(DSSMD_ERR_CUDA, "corr_bwd: cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(e));
    }
    // persistent grid: exactly the blocks that are resident at once, each with an equal share of items
    int bps = kernel_occupancy(kern, threads, smem);
    const int o = env_int_reg("DSSMD_CORR_BWD_BPS", 0);
    if (o >= 1 && o < bps) bps = o;
    int grid = sm_count() * bps;
    if (grid > p.nitems) grid = p.nitems;
    p.ipb = ceil_div(p.nitems, grid);
    grid = ceil_div(p.nitems, p.ipb);
    if (env_int_reg("DSSMD_CORR_DEBUG", 0))
        fprintf(stderr, "corr_bwd reg cfg: es=%d mode=%d PX=%d NH=%d WPR=%d TX=%d XT=%d CGW=%d CB=%d items=%d ipb=%d grid=%d (x%d/SM) threads=%d smem=%zu\n",
                (int)sizeof(T), MODE, PX, p.NH, p.WPR, p.TX, p.XT, p.CGW, p.CB, p.nitems, p.ipb, grid, bps, threads, smem);
    const cudaError_t le = launch_pdl(kern, dim3((unsigned)grid), dim3((unsigned)threads), smem, stream, p, mapIn, mapG);
    if (le != cudaSuccess) return fail(DSSMD_ERR_CUDA, "corr_bwd(reg): launch failed: %s", cudaGetErrorString(le));
    return check_launch(MODE == 0 ? "corr_bwd(gL, reg)" : "corr_bwd(gR, reg)");
}

}  // namespace

namespace {

// -1 = not eligible; with DSSMD_CORR_DEBUG set, says which test rejected the shape
int reg_reject(int where) {
    if (env_int_reg("DSSMD_CORR_DEBUG", 0)) fprintf(stderr, "corr_bwd reg: not eligible (test %d), generic kernel instead\n", where);
    return -1;
}

template <typename T>
int corr_bwd_reg_typed(const void *g, const void *in, void *out, int mode, int B, int C, int H, int W, int D, long long gbs, CUtensorMapDataType tmtype,
                       cudaStream_t stream) {
    constexpr int ES = (int)sizeof(T);
    constexpr int EPV = 16 / ES;                    // elements per 16 bytes: rows and boxes are multiples of it
    if (W % EPV != 0 || D > kDP * 16) return reg_reject(1);
    if (((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) != 0) return reg_reject(2);
    if ((long long)B * H * 64 >= (1LL << 31)) return reg_reject(3);
    RegParams p{};
    p.out = out;
    p.C = C; p.H = H; p.W = W; p.D = D;
    p.invC = 1.0f / (float)C;
    int nh = 1, lognh = 0;
    while (nh * kDP < D) { nh *= 2; ++lognh; }
    p.NH = nh; p.LOGNH = lognh;
    p.HALO = round_up(kDP * nh, EPV);
    // pixels per lane: 8 (19 FFMAs per LDS.128 instead of 12, needs the D split so that the window loads stay
    // conflict-free) when that does not leave more lanes idle at the end of the row than 4 does (fp32 only)
    auto lanes_used = [&](int px_) { const int pw = (32 / nh) * px_; return (double)W / (double)(ceil_div(W, pw) * pw); };
    int px = (ES == 4 && nh >= 4 && nh <= 8 && lanes_used(8) >= lanes_used(4)) ? 8 : 4;
    if (ES == 4) {
        const int o = env_int_reg("DSSMD_CORR_BWD_PX", 0);
        if (o == 4 || (o == 8 && nh >= 2 && nh <= 8)) px = o;
    }
    const int pxw = (32 / nh) * px;                 // pixels one warp covers
    // x tile: the operand box (tile + halo) must fit a 256-element TMA box
    int wpr = ceil_div(W, pxw);
    while (wpr > 1 && wpr * pxw + p.HALO > 256) wpr = (wpr + 1) / 2;
    if (wpr * pxw + p.HALO > 256) return reg_reject(4);
    p.WPR = wpr;
    p.TX = wpr * pxw;
    p.XT = ceil_div(W, p.TX);
    p.WP = p.TX + p.HALO;
    p.GW = mode == 0 ? p.TX : p.WP;
    int cgw = env_int_reg("DSSMD_CORR_BWD_CGW", 0);
    if (cgw < 1) cgw = 8 / wpr;
    if (cgw < 1) cgw = 1;
    const int max_warps = px == 4 ? 16 : 8;   // __launch_bounds__ of the kernel
    if (wpr > max_warps) return reg_reject(5);
    while (cgw > 1 && wpr * cgw > max_warps) --cgw;
    if (cgw > C) cgw = C;
    p.CGW = cgw;
    // channels per item: small enough that every SM sees several items (the pipeline needs a
    // few to hide its first load), large enough to amortise the per-item barrier
    int cb = env_int_reg("DSSMD_CORR_BWD_CB", 0);
    if (cb < 1) {
        cb = C;
        while (cb > 8 * cgw && (long)B * H * p.XT * ceil_div(C, cb) < 6L * sm_count()) cb = (cb + 1) / 2;
        cb = round_up(cb, cgw);
    }
    if (cb > 256) cb = 256;
    if (cb > C) cb = C;
    // two stages of (operand tile + g tile) must leave room for two resident blocks
    {
        const size_t gbytes = (size_t)D * p.GW * ES + 128, rowb = (size_t)p.WP * ES;
        size_t budget = (size_t)env_int_reg("DSSMD_CORR_BWD_SMEM_KB2", 100) * 1024 / 2;
        // large D (the max_disp stress sweep: D = 96 on the 1/8 feature map): the g tile alone outgrows half an SM -- one
        // resident block with the whole budget is still far ahead of the generic kernel (D-sweep, profiles/r02j_measure.json:
        // 131 / 156 us per gradient at D = 96 there)
        if (gbytes + rowb * cgw > budget) budget = (size_t)99 * 1024;
        if (gbytes + rowb * cgw > budget) return reg_reject(6);
        while (cb > cgw && gbytes + rowb * cb + 128 > budget) cb -= cgw;
    }
    p.CB = cb;
    p.NCB = ceil_div(C, cb);
    p.op_bytes = (uint32_t)((size_t)cb * p.WP * ES);
    p.g_bytes = (uint32_t)((size_t)D * p.GW * ES);
    const size_t stage = (size_t)((p.op_bytes + 127u) & ~127u) + ((p.g_bytes + 127u) & ~127u);
    const size_t smem = 2 * stage + 128;
    if (smem > 200 * 1024) return reg_reject(7);
    p.nitems = B * H * p.XT * p.NCB;
    const int threads = wpr * cgw * 32;
    CUtensorMap mapIn, mapG;
    {
        const cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)B};
        const cuuint64_t strides[3] = {(cuuint64_t)W * ES, (cuuint64_t)H * W * ES, (cuuint64_t)C * H * W * ES};
        const cuuint32_t box[4] = {(cuuint32_t)p.WP, 1, (cuuint32_t)cb, 1};
        if (!make_tmap(&mapIn, tmtype, in, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return reg_reject(8);
    }
    {
        const cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)B};
        const cuuint64_t strides[3] = {(cuuint64_t)W * ES, (cuuint64_t)H * W * ES, (cuuint64_t)gbs * ES};
        const cuuint32_t box[4] = {(cuuint32_t)p.GW, 1, (cuuint32_t)D, 1};
        if (!make_tmap(&mapG, tmtype, g, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return reg_reject(9);
    }
#define DSSMD_REG_LAUNCH(M)                                                                     \
    if constexpr (ES == 4) {                                                                    \
        if (px == 8) {                                                                          \
            switch (lognh) {                                                                    \
                case 1: return launch_reg<T, M, 1, 8>(p, mapIn, mapG, threads, smem, stream);   \
                case 2: return launch_reg<T, M, 2, 8>(p, mapIn, mapG, threads, smem, stream);   \
                default: return launch_reg<T, M, 3, 8>(p, mapIn, mapG, threads, smem, stream);  \
            }                                                                                   \
        }                                                                                       \
    }                                                                                           \
    switch (lognh) {                                                                            \
        case 0: return launch_reg<T, M, 0, 4>(p, mapIn, mapG, threads, smem, stream);           \
        case 1: return launch_reg<T, M, 1, 4>(p, mapIn, mapG, threads, smem, stream);           \
        case 2: return launch_reg<T, M, 2, 4>(p, mapIn, mapG, threads, smem, stream);           \
        case 3: return launch_reg<T, M, 3, 4>(p, mapIn, mapG, threads, smem, stream);           \
        default: return launch_reg<T, M, 4, 4>(p, mapIn, mapG, threads, smem, stream);          \
    }
    if (mode == 0) { DSSMD_REG_LAUNCH(0) }
    DSSMD_REG_LAUNCH(1)
#undef DSSMD_REG_LAUNCH
}

}  // namespace

// mode 0: out = gL, in = R;  mode 1: out = gR, in = L.  dtype: DSSMD_F32 / BF16 / F16 (all tensors).  Returns -1 when
// the shape is not eligible (the caller falls back to the generic kernel), else a DSSMD_* code.
int corr_bwd_reg(const void *g, const void *in, void *out, int mode, int B, int C, int H, int W, int D, long long gbs, int dtype, cudaStream_t stream) {
    switch (dtype) {
        case DSSMD_F32: return corr_bwd_reg_typed<float>(g, in, out, mode, B, C, H, W, D, gbs, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, stream);
        case DSSMD_BF16: return corr_bwd_reg_typed<__nv_bfloat16>(g, in, out, mode, B, C, H, W, D, gbs, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, stream);
        case DSSMD_F16: return corr_bwd_reg_typed<__half>(g, in, out, mode, B, C, H, W, D, gbs, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, stream);
        default: return -1;
    }
}

}  // namespace dssmd
