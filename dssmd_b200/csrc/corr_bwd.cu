// corr_bwd.cu -- backward of the 1-D correlation cost volume.
//
// Replaces the graph autograd builds from build_corr (upstream
// models/UtilsNet/submoudles.py:304-308):
//   gL[b,c,y,x ] = (1/C) * sum_{d<=min(D-1,x)}      g[b,d,y,x]    * R[b,c,y,x-d]
//   gR[b,c,y,x'] = (1/C) * sum_{d<=min(D-1,W-1-x')} g[b,d,y,x'+d] * L[b,c,y,x'+d]
//
// Both gradients are gathers (no atomics).  One kernel template, two modes:
//   MODE 0 (gL): operand = R, read at x - d   (band extends HALO pixels to the left)
//   MODE 1 (gR): operand = L, read at x' + d  (band extends HALO pixels to the right),
//                g is staged skewed: gs[d][x'] = g[d][x'+d]
// A thread owns 8 consecutive pixels x 8 channels (64 accumulators) and streams
// over d four at a time: per step-of-4 it loads 4x8 g values (8 LDS.128) and,
// per channel, ONE new 16-byte chunk of the operand row -- the other two chunks
// of its 12-wide sliding window stay in registers from the previous step.
// => 16 LDS.128 per 256 FFMA.  Shared memory uses the same [chunk][line] XOR
// swizzled layout as corr_fwd.cu.
// Algorithmic HBM bytes per launch pair: (B*D*H*W + 4*B*C*H*W) * sizeof(T).
#include "common.cuh"

#include <cstdlib>
#include <type_traits>

namespace dssmd {

// corr_bwd_reg.cu: fp32 fast path (g slab in registers); -1 = shape not eligible
int corr_bwd_reg(const void *g, const void *in, void *out, int mode, int B, int C, int H, int W, int D, long long gbs, int dtype, cudaStream_t stream);
// corr_bwd_mma32.cu: fp32 on mma.sync with 3xTF32 splitting; -1 = shape not eligible
int corr_bwd_mma32(const void *g, const void *in, void *out, int mode, int B, int C, int H, int W, int D, long long gbs, cudaStream_t stream);
// ... both gradients in one launch (grid.y = gradient); -1 = not eligible
int corr_bwd_mma32_both(const void *g, const void *L, const void *R, void *gL, void *gR, int B, int C, int H, int W, int D, long long gbs, cudaStream_t stream);
// corr_bwd_fused.cu: both gradients of an fp32 correlation in one launch; -1 = not eligible
int corr_bwd_fused_f32(const void *g, const void *L, const void *R, void *gL, void *gR, int B, int C, int H, int W, int D, long long gbs,
                       cudaStream_t stream);
// corr_bwd_wide.cu: the same for D <= 24 with a lane owning all disparities of its pixels; -1 = not eligible
int corr_bwd_wide_f32(const void *g, const void *L, const void *R, void *gL, void *gR, int B, int C, int H, int W, int D, long long gbs,
                       cudaStream_t stream);
// corr_bwd_mma16.cu: 16-bit tensors on the tensor cores; -1 = shape not eligible
int corr_bwd_mma16(const void *g, const void *in, void *out, int mode, int B, int C, int H, int W, int D, long long gbs, int dtype, cudaStream_t stream);

namespace {

constexpr int kPix = 8;   // pixels per thread
constexpr int kCh = 8;    // channels per thread
constexpr int kMaxRows = 16;
constexpr int kMaxFill = 4;  // 16-byte chunks one lane copies per operand line (<= 128 chunks per line)

struct CorrBwdParams {
    const void *g;
    const void *In;   // R (mode 0) or L (mode 1)
    void *out;        // gL (mode 0) or gR (mode 1)
    long long GBS;    // elements between two images of g (D*H*W, or the channel count of a concat gradient * H*W)
    int B, C, H, W, D;
    int NR;     // B*H rows
    int TX;     // tile width (multiple of 8)
    int NPG;    // TX/8
    int XT;     // x tiles per row
    int ROWS;   // rows per block
    int CG;     // channel groups (of 8) per stage
    int CK;     // channels per stage = 8*CG
    int DP;     // D rounded up to a multiple of 4
    int HALO;   // operand band halo in pixels (multiple of 8, >= DP)
    int TXc;    // chunks per line of the g tile
    int RXc;    // chunks per line of the operand band
    int NLG;    // g lines   (ROWS*DP, padded odd)
    int NL;     // operand lines (ROWS*CK, padded odd)
    int CPB;    // channels per block (grid.y splits C)
    int NKC;    // channel chunks per block
    int c_pow2;
};

__device__ __forceinline__ int swz(int j) { return j ^ ((j >> 3) & 1); }

template <typename T>
__device__ __forceinline__ float4 load4_generic(const T *row, int gx, int W, bool ok) {
    float v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int x = gx + i;
        v[i] = (ok && x >= 0 && x < W) ? to_f32<T>(row[x]) : 0.f;
    }
    return make_float4(v[0], v[1], v[2], v[3]);
}

template <typename T, int MODE, bool ASYNC>
__global__ void __launch_bounds__(256) corr_bwd_kernel(const CorrBwdParams p) {
    extern __shared__ float4 smem[];
    __shared__ long long s_rowbase_in[kMaxRows];  // element offset of (b, c=0, y, 0) in [B,C,H,W]
    __shared__ long long s_rowbase_g[kMaxRows];   // element offset of (b, d=0, y, 0) in [B,D,H,W]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nwarps = blockDim.x >> 5;
    const int xt = blockIdx.x % p.XT;
    const int R0 = (blockIdx.x / p.XT) * p.ROWS;
    const int X0 = xt * p.TX;
    const int CB0 = blockIdx.y * p.CPB;  // first channel of this block
    const int CEND = (CB0 + p.CPB < p.C) ? CB0 + p.CPB : p.C;
    const long long HW = (long long)p.H * p.W;

    if (tid < p.ROWS) {
        const int row = R0 + tid;
        long long bi = -1, bg = -1;
        if (row < p.NR) {
            const int b = row / p.H, y = row - b * p.H;
            bi = (long long)b * p.C * HW + (long long)y * p.W;
            bg = (long long)b * p.GBS + (long long)y * p.W;
        }
        s_rowbase_in[tid] = bi;
        s_rowbase_g[tid] = bg;
    }
    __syncthreads();

    float4 *gs = smem;                     // g tile:   [TXc][NLG]
    float4 *is = smem + p.TXc * p.NLG;     // operand:  [RXc][NL]

    const T *Gg = static_cast<const T *>(p.g);
    const T *Ig = static_cast<const T *>(p.In);

    // ---- stage the g tile once per block ---------------------------------------
    {
        const int nlines = p.ROWS * p.DP;
        for (int line = warp; line < nlines; line += nwarps) {
            const int rr = line / p.DP, d = line - rr * p.DP;
            const long long rb = s_rowbase_g[rr];
            const bool lv = (rb >= 0) && (d < p.D);
            const T *row = Gg + (lv ? rb + (long long)d * HW : 0);
            for (int j = lane; j < p.TXc; j += 32) {
                float4 *dst = gs + swz(j) * p.NLG + line;
                const int gx = X0 + 4 * j + (MODE == 1 ? d : 0);
                if constexpr (ASYNC && MODE == 0) {
                    const bool ok = lv && gx < p.W;
                    cp_async16(smem_u32(dst), ok ? (const void *)(row + gx) : (const void *)Gg, ok ? 16 : 0);
                } else if constexpr (ASYNC) {
                    // skewed row: source start x0+4j+d is only 4-byte aligned -> four 4-byte async copies
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const bool ok = lv && gx + i < p.W;
                        cp_async4(smem_u32(dst) + 4 * i, ok ? (const void *)(row + gx + i) : (const void *)Gg, ok ? 4 : 0);
                    }
                } else {
                    *dst = load4_generic<T>(row, gx, p.W, lv);
                }
            }
        }
    }

    // ---- this thread's work item --------------------------------------------------
    const int pg = tid % p.NPG;
    const int t2 = tid / p.NPG;
    const int cg = t2 % p.CG;
    const int r = t2 / p.CG;
    const int xs = pg * kPix;
    const bool row_ok = (r < p.ROWS) && (R0 + r < p.NR) && (X0 + xs < p.W);
    const int rsafe = row_ok ? r : 0;

    const int gline0 = rsafe * p.DP;
    const int og0 = swz(2 * pg) * p.NLG + gline0;
    const int og1 = swz(2 * pg + 1) * p.NLG + gline0;
    const int iline0 = rsafe * p.CK + cg * kCh;
    const int jbase = 2 * pg + (MODE == 0 ? p.HALO / 4 : 0);  // chunk of the thread's own 8 pixels in the band
    const int IX0 = MODE == 0 ? X0 - p.HALO : X0;             // global x of band-local index 0
    const int nsteps = p.DP / 4;
    const int nlines_in = p.ROWS * p.CK;

    const float Cf = (float)p.C, invC = 1.0f / Cf;

    // per-lane fill descriptors (line-invariant): operand chunks this lane copies
    int f_dst[kMaxFill], f_gx[kMaxFill];
#pragma unroll
    for (int k = 0; k < kMaxFill; ++k) {
        const int j = lane + 32 * k;
        f_dst[k] = (j < p.RXc) ? swz(j) * p.NL : -1;
        f_gx[k] = IX0 + 4 * j;
    }

    for (int kc = 0; kc < p.NKC; ++kc) {
        __syncthreads();  // previous chunk fully consumed
        // ---- stage CK channels of the operand band ------------------------------------
        for (int line = warp; line < nlines_in; line += nwarps) {
            const int rr = line / p.CK, cc = line - rr * p.CK;
            const int c = CB0 + kc * p.CK + cc;
            const long long rb = s_rowbase_in[rr];
            const bool lv = (rb >= 0) && (c < CEND);
            const T *row = Ig + (lv ? rb + (long long)c * HW : 0);
#pragma unroll
            for (int k = 0; k < kMaxFill; ++k) {
                if (f_dst[k] < 0) break;
                float4 *dst = is + f_dst[k] + line;
                const int gx = f_gx[k];
                if constexpr (ASYNC) {
                    const bool ok = lv && gx >= 0 && gx < p.W;
                    cp_async16(smem_u32(dst), ok ? (const void *)(row + gx) : (const void *)Ig, ok ? 16 : 0);
                } else {
                    *dst = load4_generic<T>(row, gx, p.W, lv);
                }
            }
        }
        if constexpr (ASYNC) {
            cp_async_commit();
            cp_async_wait<0>();
        }
        __syncthreads();

        const int c0 = CB0 + kc * p.CK + cg * kCh;
        if (!row_ok || c0 >= CEND) continue;

        float acc[kCh][kPix];
        float kept[kCh][8];
#pragma unroll
        for (int i = 0; i < kCh; ++i) {
            const float4 a = is[swz(jbase) * p.NL + iline0 + i];
            const float4 b = is[swz(jbase + 1) * p.NL + iline0 + i];
            kept[i][0] = a.x; kept[i][1] = a.y; kept[i][2] = a.z; kept[i][3] = a.w;
            kept[i][4] = b.x; kept[i][5] = b.y; kept[i][6] = b.z; kept[i][7] = b.w;
#pragma unroll
            for (int px = 0; px < kPix; ++px) acc[i][px] = 0.f;
        }

#pragma unroll 2
        for (int m = 0; m < nsteps; ++m) {
            float gv[4][kPix];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 a = gs[og0 + 4 * m + j], b = gs[og1 + 4 * m + j];
                gv[j][0] = a.x; gv[j][1] = a.y; gv[j][2] = a.z; gv[j][3] = a.w;
                gv[j][4] = b.x; gv[j][5] = b.y; gv[j][6] = b.z; gv[j][7] = b.w;
            }
            const int jn = MODE == 0 ? jbase - m - 1 : jbase + m + 2;  // the one new chunk
            const int on = swz(jn) * p.NL + iline0;
#pragma unroll
            for (int i = 0; i < kCh; ++i) {
                const float4 nw = is[on + i];
                float w[12];
                if (MODE == 0) {
                    w[0] = nw.x; w[1] = nw.y; w[2] = nw.z; w[3] = nw.w;
#pragma unroll
                    for (int q = 0; q < 8; ++q) w[4 + q] = kept[i][q];
                } else {
#pragma unroll
                    for (int q = 0; q < 8; ++q) w[q] = kept[i][q];
                    w[8] = nw.x; w[9] = nw.y; w[10] = nw.z; w[11] = nw.w;
                }
                // mode 0: In[x - d]  = w[px - j + 4];   mode 1: In[x' + d] = w[px + j]
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int px = 0; px < kPix; ++px)
                        acc[i][px] = fmaf(gv[j][px], w[MODE == 0 ? px - j + 4 : px + j], acc[i][px]);
                // slide the window by 4
#pragma unroll
                for (int q = 0; q < 8; ++q) kept[i][q] = w[MODE == 0 ? q : q + 4];
            }
        }

        // ---- epilogue: scale by 1/C and store 8 px x 8 ch -------------------------------
        const int row = R0 + r;
        const int b = row / p.H, y = row - b * p.H;
        const int xg = X0 + xs;
        T *ob = static_cast<T *>(p.out) + ((long long)b * p.C * p.H + y) * p.W + xg;
        const bool vec = (sizeof(T) == 4) && ((p.W & 3) == 0) && (xg + kPix <= p.W) &&
                         ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0);
#pragma unroll
        for (int i = 0; i < kCh; ++i) {
            const int c = c0 + i;
            if (c >= CEND) break;
            float v[kPix];
#pragma unroll
            for (int px = 0; px < kPix; ++px)
                v[px] = p.c_pow2 ? acc[i][px] * invC : __fdiv_rn(acc[i][px], Cf);
            T *o = ob + (long long)c * HW;
            if (vec) {
                st_global_cs(reinterpret_cast<float4 *>(o), make_float4(v[0], v[1], v[2], v[3]));
                st_global_cs(reinterpret_cast<float4 *>(o) + 1, make_float4(v[4], v[5], v[6], v[7]));
            } else {
#pragma unroll
                for (int px = 0; px < kPix; ++px)
                    if (xg + px < p.W) o[px] = from_f32<T>(v[px]);
            }
        }
    }
}

// ---- fp64: plain gather kernels (arbiter precision, not a hot path) ---------------
__global__ void corr_bwd_f64_kernel(const double *g, const double *L, const double *R, double *gL, double *gR,
                                    int B, int C, int H, int W, int D, long long gbs) {
    const long long n = (long long)B * C * H * W;
    const long long HW = (long long)H * W;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % W);
        const int y = (int)((i / W) % H);
        const int b = (int)(i / (HW * C));
        const double *gr = g + (long long)b * gbs + (long long)y * W;
        const long long rowoff = i - x;
        double a = 0.0, r = 0.0;
        for (int d = 0; d < D && d <= x; ++d) a += (gr[d * HW + x] / (double)C) * R[rowoff + x - d];
        for (int d = 0; d < D && x + d < W; ++d) r += (gr[d * HW + x + d] / (double)C) * L[rowoff + x + d];
        if (gL) gL[i] = a;
        if (gR) gR[i] = r;
    }
}

int env_int(const char *name, int dflt) {
    if (!tuning_enabled()) return dflt;
    const char *v = std::getenv(name);
    return (v && *v) ? std::atoi(v) : dflt;
}

template <typename T, int MODE>
int launch_mode(const void *g, const void *In, void *out, int B, int C, int H, int W, int D, long long gbs, cudaStream_t stream) {
    if constexpr (sizeof(T) == 2) {
        if (env_int("DSSMD_CORR_BWD_VARIANT", -1) < 0) {   // no variant forced: mma.sync on the banded row GEMM
            const int rc = corr_bwd_mma16(g, In, out, MODE, B, C, H, W, D, gbs, std::is_same<T, __nv_bfloat16>::value ? DSSMD_BF16 : DSSMD_F16, stream);
            if (rc >= 0) return rc;
        }
    }
    if constexpr (sizeof(T) == 4) {
        // 3 = mma.sync with 3xTF32 splitting.  Unset: taken where it measured faster than the register-blocked FFMA kernel -- problems of
        // several thousand 16-pixel warp tiles (cfg2 117 / 123 -> 81 / 77 us, KITTI 45 / 46 -> 37 / 38 us), rows that fill the FFMA
        // kernel's tiles badly (W = 80: 16.7 / 17.8 -> 15.4 / 14.9 us), and gL everywhere (cfg5 14.9 -> 13.9 us; gR there 15.4 vs 17.0)
        const int forced = env_int("DSSMD_CORR_BWD_VARIANT", -1);
        const long warp_tiles = (long)B * H * ceil_div(W, 16);
        const bool fills_reg_tiles = (double)W >= 0.75 * (double)(ceil_div(W, 128) * 128);
        if (forced == 3 || (forced < 0 && (MODE == 0 || !fills_reg_tiles || warp_tiles >= 4000))) {
            const int rc = corr_bwd_mma32(g, In, out, MODE, B, C, H, W, D, gbs, stream);
            if (rc >= 0) return rc;
        }
    }
    if (env_int("DSSMD_CORR_BWD_VARIANT", 1) >= 1) {
        constexpr int dtype = sizeof(T) == 4 ? DSSMD_F32 : std::is_same<T, __nv_bfloat16>::value ? DSSMD_BF16 : DSSMD_F16;
        const int rc = corr_bwd_reg(g, In, out, MODE, B, C, H, W, D, gbs, dtype, stream);
        if (rc >= 0) return rc;
    }
    CorrBwdParams p{};
    p.g = g; p.In = In; p.out = out; p.GBS = gbs;
    p.B = B; p.C = C; p.H = H; p.W = W; p.D = D;
    p.NR = B * H;
    p.c_pow2 = (C & (C - 1)) == 0;
    p.DP = round_up(D, 4);
    p.HALO = round_up(p.DP, 8);

    const int NPGF = ceil_div(W, kPix);
    const int cgroups = ceil_div(C, kCh);
    const size_t budget = (size_t)env_int("DSSMD_CORR_BWD_SMEM_KB", 40) * 1024;

    // tile width: whole row when its band fits 128 chunks, and the g tile at most half the budget
    int npg_cap = (128 * 4 - p.HALO) / kPix;
    if (npg_cap > 32) npg_cap = 32;
    while (npg_cap > 4 && (size_t)(npg_cap * 2) * (p.DP | 1) * 16 > budget / 2) npg_cap = (npg_cap + 1) / 2;
    if (npg_cap < 1) npg_cap = 1;
    p.XT = ceil_div(NPGF, npg_cap);
    p.NPG = ceil_div(NPGF, p.XT);
    p.TX = p.NPG * kPix;
    p.TXc = p.TX / 4;
    p.RXc = (p.TX + p.HALO) / 4;

    // threads: ROWS x CG x NPG; channel groups per stage CG (power of two), ~64-128 threads per block
    int cg = 1;
    while (cg * 2 <= cgroups && cg * 2 * p.NPG <= 64 && cg < 8) cg *= 2;
    cg = env_int("DSSMD_CORR_BWD_CG", cg);
    if (cg < 1 || cg > 16 || cg * p.NPG > 256) cg = 1;
    int rows = 64 / (cg * p.NPG);
    rows = env_int("DSSMD_CORR_BWD_ROWS", rows);
    if (rows < 1) rows = 1;
    if (rows > kMaxRows) rows = kMaxRows;
    if (rows > p.NR) rows = p.NR;
    while (rows > 1 && rows * cg * p.NPG > 256) --rows;
    size_t smem_bytes = 0;
    for (;;) {
        p.ROWS = rows; p.CG = cg; p.CK = cg * kCh;
        p.NLG = (p.ROWS * p.DP) | 1;
        p.NL = (p.ROWS * p.CK) | 1;
        smem_bytes = ((size_t)p.TXc * p.NLG + (size_t)p.RXc * p.NL) * 16;
        if (smem_bytes <= budget) break;
        if (rows > 1) { rows = (rows + 1) / 2; continue; }
        if (cg > 1) { cg /= 2; continue; }
        break;
    }
    if (smem_bytes > 200 * 1024)
        return fail(DSSMD_ERR_UNSUPPORTED, "corr_bwd: tile needs %zu B of shared memory (W=%d D=%d)", smem_bytes, W, D);

    // channel ranges per block (grid.y): enough blocks for >= ~10 warps per SM, each block >= 1 chunk
    const int nchunks = ceil_div(C, p.CK);
    const long blocks_xy = (long)p.XT * ceil_div(p.NR, p.ROWS);
    const int warps_blk = ceil_div(p.ROWS * p.CG * p.NPG, 32);
    int cz = 1;
    while (cz < nchunks && blocks_xy * cz * warps_blk < (long)sm_count() * 12) cz *= 2;
    cz = env_int("DSSMD_CORR_BWD_CZ", cz);
    if (cz < 1) cz = 1;
    if (cz > nchunks) cz = nchunks;
    p.NKC = ceil_div(nchunks, cz);
    p.CPB = p.NKC * p.CK;
    cz = ceil_div(C, p.CPB);

    const int threads = round_up(p.ROWS * p.CG * p.NPG, 32);
    const bool async = (sizeof(T) == 4) && (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(g) & 15) == 0) &&
                       ((reinterpret_cast<uintptr_t>(In) & 15) == 0);
    dim3 grid((unsigned)(p.XT * ceil_div(p.NR, p.ROWS)), (unsigned)cz);
    if (env_int("DSSMD_CORR_DEBUG", 0))
        fprintf(stderr, "corr_bwd cfg: mode=%d NPG=%d XT=%d ROWS=%d CG=%d CK=%d CZ=%d NKC=%d threads=%d grid=(%u,%u) smem=%zu async=%d\n",
                MODE, p.NPG, p.XT, p.ROWS, p.CG, p.CK, cz, p.NKC, threads, grid.x, grid.y, smem_bytes, (int)async);
    cudaError_t e;
    if (async) {
        if constexpr (sizeof(T) == 4) {
            auto kern = corr_bwd_kernel<T, MODE, true>;
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
            if (e != cudaSuccess) return fail(DSSMD_ERR_CUDA, "corr_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            kern<<<grid, threads, smem_bytes, stream>>>(p);
        }
    } else {
        auto kern = corr_bwd_kernel<T, MODE, false>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
        if (e != cudaSuccess) return fail(DSSMD_ERR_CUDA, "corr_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        kern<<<grid, threads, smem_bytes, stream>>>(p);
    }
    return check_launch(MODE == 0 ? "corr_bwd(gL)" : "corr_bwd(gR)");
}

template <typename T>
int launch_typed(const void *g, const void *L, const void *R, void *gL, void *gR,
                 int B, int C, int H, int W, int D, long long gbs, cudaStream_t stream) {
    if constexpr (sizeof(T) == 4) {
        // both gradients wanted (the autograd case): one fused launch where the shape is eligible (corr_bwd_fused.cu).  A forced
        // per-gradient variant (DSSMD_CORR_BWD_VARIANT) keeps the two-launch path; DSSMD_CORR_BWD_FUSED=0 disables the fusion.
        if (gL && gR && env_int("DSSMD_CORR_BWD_VARIANT", -1) < 0) {
            // Where both gradients would take the mma.sync kernel (launch_mode's rule) they are issued as ONE launch.  Rows that fill
            // the FFMA kernels' 128-pixel lane tiles badly (W = 80, 160) go there FIRST: measured over the batch size
            // (scripts/corr_bwd_b_sweep.py), wide FFMA kernel / one-launch mma.sync: 48x160 B = 1 9.4 / 8.3 us, B = 4 24.4 / 17.6,
            // B = 16 58.6 (two mma.sync launches) / 53.2; 40x80 B = 1 6.2 / 5.9, B = 8 20.0 / 15.8, B = 16 36.3 / 25.2; 72x120 (fills
            // its tile) B = 4 18.3 / 19.0, B = 8 33.2 / 35.9: the wide kernel keeps it.  DSSMD_CORR_BWD_WIDE=2 forces the wide kernel.
            const long warp_tiles = (long)B * H * ceil_div(W, 16);
            const bool fills_reg_tiles = (double)W >= 0.75 * (double)(ceil_div(W, 128) * 128);
            const int both = env_int("DSSMD_MMA32_BOTH", 1);
            int rc = -1;
            if (!fills_reg_tiles && both != 0 && env_int("DSSMD_CORR_BWD_WIDE", 1) < 2) {
                rc = corr_bwd_mma32_both(g, L, R, gL, gR, B, C, H, W, D, gbs, stream);
                if (rc >= 0) return rc;
            }
            rc = corr_bwd_wide_f32(g, L, R, gL, gR, B, C, H, W, D, gbs, stream);
            if (rc >= 0) return rc;
            rc = corr_bwd_fused_f32(g, L, R, gL, gR, B, C, H, W, D, gbs, stream);
            if (rc >= 0) return rc;
            if (!fills_reg_tiles || warp_tiles >= 4000 || both == 2) {
                rc = corr_bwd_mma32_both(g, L, R, gL, gR, B, C, H, W, D, gbs, stream);
                if (rc >= 0) return rc;
            }
        }
    }
    if (gL) {
        const int rc = launch_mode<T, 0>(g, R, gL, B, C, H, W, D, gbs, stream);
        if (rc) return rc;
    }
    if (gR) {
        const int rc = launch_mode<T, 1>(g, L, gR, B, C, H, W, D, gbs, stream);
        if (rc) return rc;
    }
    return DSSMD_OK;
}

}  // namespace

}  // namespace dssmd

namespace dssmd {
namespace {
int corr_bwd_entry(const void *g, const void *L, const void *R, void *gL, void *gR, int B, int C, int H, int W, int D, long long gbs,
                   int dtype, cudaStream_t s) {
    switch (dtype) {
        case DSSMD_F32: return launch_typed<float>(g, L, R, gL, gR, B, C, H, W, D, gbs, s);
        case DSSMD_BF16: return launch_typed<__nv_bfloat16>(g, L, R, gL, gR, B, C, H, W, D, gbs, s);
        case DSSMD_F16: return launch_typed<__half>(g, L, R, gL, gR, B, C, H, W, D, gbs, s);
        case DSSMD_F64: {
            const long long n = (long long)B * C * H * W;
            const int nt = 256;
            long long nb = (n + nt - 1) / nt;
            if (nb > 148LL * 32) nb = 148LL * 32;
            corr_bwd_f64_kernel<<<(unsigned)nb, nt, 0, s>>>((const double *)g, (const double *)L, (const double *)R,
                                                            (double *)gL, (double *)gR, B, C, H, W, D, gbs);
            return check_launch("corr_bwd_f64");
        }
        default: return fail(DSSMD_ERR_INVALID_ARGUMENT, "corr_bwd: unknown dtype %d", dtype);
    }
}
}  // namespace
}  // namespace dssmd

extern "C" int dssmd_corr_bwd(const void *g, const void *L, const void *R, void *gL, void *gR,
                              int B, int C, int H, int W, int D, int dtype, void *stream) {
    using namespace dssmd;
    DSSMD_REQUIRE(g && L && R, "corr_bwd: null input pointer");
    DSSMD_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && D > 0, "corr_bwd: sizes must be positive (B=%d C=%d H=%d W=%d D=%d)", B, C, H, W, D);
    DSSMD_REQUIRE((long long)B * H < (1LL << 31), "corr_bwd: B*H too large");
    if (!gL && !gR) return DSSMD_OK;
    return corr_bwd_entry(g, L, R, gL, gR, B, C, H, W, D, (long long)D * H * W, dtype, static_cast<cudaStream_t>(stream));
}

extern "C" int dssmd_corr_bwd_cat(const void *g_cat, const void *L, const void *R, void *gL, void *gR,
                                  int B, int C, int H, int W, int D, int cat_channels, int channel_offset, int dtype, void *stream) {
    using namespace dssmd;
    DSSMD_REQUIRE(g_cat && L && R, "corr_bwd_cat: null input pointer");
    DSSMD_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && D > 0, "corr_bwd_cat: sizes must be positive (B=%d C=%d H=%d W=%d D=%d)", B, C, H, W, D);
    DSSMD_REQUIRE((long long)B * H < (1LL << 31), "corr_bwd_cat: B*H too large");
    DSSMD_REQUIRE(channel_offset >= 0 && cat_channels >= channel_offset + D,
                  "corr_bwd_cat: channels [%d, %d) do not fit a gradient of %d channels", channel_offset, channel_offset + D, cat_channels);
    DSSMD_REQUIRE(dtype == DSSMD_F32 || dtype == DSSMD_BF16 || dtype == DSSMD_F16 || dtype == DSSMD_F64, "corr_bwd_cat: unknown dtype %d", dtype);
    if (!gL && !gR) return DSSMD_OK;
    const int es = dtype == DSSMD_F64 ? 8 : dtype == DSSMD_F32 ? 4 : 2;
    const char *gp = static_cast<const char *>(g_cat) + (size_t)channel_offset * H * W * es;
    return corr_bwd_entry(gp, L, R, gL, gR, B, C, H, W, D, (long long)cat_channels * H * W, dtype, static_cast<cudaStream_t>(stream));
}
