// corr_fwd.cu -- 1-D horizontal correlation cost volume, forward.
//
// Replaces build_corr (upstream models/UtilsNet/submoudles.py:301-311):
//   out[b,d,y,x] = (1/C) * sum_c L[b,c,y,x] * R[b,c,y,x-d]   (x >= d, else 0)
//
// Kernel shape (fp32 FFMA, all D disparities of a pixel in one pass over C):
//   * a block owns ROWS image rows x one TX-pixel tile x NDGB disparity groups;
//   * per pipeline stage it stages CK channels of the L tile and of the
//     (TX + HALO)-wide R band in shared memory (cp.async 16 B, zero-filled left
//     of x = 0, which is exactly the reference's zero border);
//   * a thread owns 8 consecutive pixels x DT consecutive disparities: per
//     channel it issues 2 + (DT+8)/4 LDS.128 and 8*DT FFMA, accumulators stay
//     in registers over the whole channel loop;
//   * small problems (fewer work items than the chip has warp slots) split the
//     channels of each stage over CS thread groups and reduce the partial sums
//     through shared memory at the end;
//   * shared memory is laid out [16-byte chunk][line] with an XOR swizzle on
//     the chunk index so that 8 lanes reading chunks 2*lane+k hit 8 distinct
//     bank groups, and with an odd line count so the fill is conflict-free.
// Algorithmic HBM bytes per launch: (2*B*C*H*W + B*D*H*W) * sizeof(T).
#include "common.cuh"

#include <cstdlib>
#include <type_traits>

namespace dssmd {

// tensor-core path (corr_fwd_umma.cu); returns -1 when the shape is outside its limits
int corr_fwd_umma_f32(const void *L, const void *R, void *out, int B, int C, int H, int W, int D, long long obs, cudaStream_t stream);
// register-blocked, TMA-staged FFMA path (corr_fwd_reg.cu), fp32 / bf16 / fp16; -1 when the shape is not eligible
int corr_fwd_reg(const void *L, const void *R, void *out, int B, int C, int H, int W, int D, long long obs, int dtype, cudaStream_t stream);
// second tcgen05 design (corr_fwd_umma2.cu): row-aligned tiles, deep raw ring; -1 when the shape is outside its limits
int corr_fwd_umma2_f32(const void *L, const void *R, void *out, int B, int C, int H, int W, int D, long long obs, cudaStream_t stream);
// fp32 on mma.sync with 3xTF32 splitting (corr_fwd_mma32.cu); -1 when the shape is outside its limits
int corr_fwd_mma32(const void *L, const void *R, void *out, int B, int C, int H, int W, int D, long long obs, cudaStream_t stream);
// tensor-core path for 16-bit tensors (corr_fwd_mma16.cu); -1 when the shape is outside its limits
int corr_fwd_mma16(const void *L, const void *R, void *out, int B, int C, int H, int W, int D, long long obs, int dtype, cudaStream_t stream);

namespace {

constexpr int kPix = 8;        // pixels per thread
constexpr int kMaxRows = 16;   // rows per block upper bound
constexpr int kStages = 3;     // cp.async pipeline depth
constexpr int kMaxFill = 4;    // 16-byte chunks one lane copies per line (=> <= 128 chunks per line)

struct CorrFwdParams {
    const void *L;
    const void *R;
    void *out;
    long long OBS;  // elements between two images of `out` (D*H*W, or the channel count of a concat buffer * H*W)
    int B, C, H, W, D;
    int NR;     // B*H image rows
    int TX;     // tile width in pixels (multiple of 8)
    int NPG;    // pixel groups per tile = TX/8
    int NDGB;   // disparity groups per block
    int ROWS;   // rows per block
    int XT;     // x tiles per row
    int CS;     // channel-split groups (power of two, divides CK)
    int LOGCS;
    int CK;     // channels per stage (power of two)
    int LOGCK;
    int NL;     // lines (= ROWS*CK) padded to an odd count
    int TXc;    // 16-byte chunks of the L tile per line
    int RXc;    // 16-byte chunks of the R band per line
    int HALO;   // left halo of the R band in pixels (multiple of 8)
    int NKC;    // channel chunks = ceil(C/CK)
    int NITEMS; // ROWS*NDGB*NPG work items per block
    int c_pow2; // C is a power of two -> multiply by exact 1/C
};

__device__ __forceinline__ int swz(int j) { return j ^ ((j >> 3) & 1); }

// four consecutive elements starting at gx of one row, zero outside [0,W)
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

template <typename T, int DT, bool ASYNC>
__global__ void __launch_bounds__(256) corr_fwd_kernel(const CorrFwdParams p) {
    extern __shared__ float4 smem[];
    __shared__ long long s_rowbase[kMaxRows];  // element offset of (b, c=0, y, x=0); -1 = no such row

    constexpr int NCH = (DT + 8) / 4;  // R chunks per thread per channel
    constexpr int STAGES = ASYNC ? kStages : 1;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nwarps = blockDim.x >> 5;
    const int xt = blockIdx.x % p.XT;
    const int R0 = (blockIdx.x / p.XT) * p.ROWS;
    const int X0 = xt * p.TX;
    const int DB0 = blockIdx.y * p.NDGB * DT;  // first disparity of this block
    const long long HW = (long long)p.H * p.W;

    if (tid < p.ROWS) {
        const int row = R0 + tid;
        long long base = -1;
        if (row < p.NR) {
            const int b = row / p.H, y = row - b * p.H;
            base = (long long)b * p.C * HW + (long long)y * p.W;
        }
        s_rowbase[tid] = base;
    }
    __syncthreads();

    // ---- this thread's work item: (row r, disparity group dgl, pixel group pg) x channel split cs ----
    const int item = tid % p.NITEMS;
    const int cs = tid / p.NITEMS;
    const int pg = item % p.NPG;
    const int t2 = item / p.NPG;
    const int dgl = t2 % p.NDGB;
    const int r = t2 / p.NDGB;
    const int d0 = DB0 + dgl * DT;
    const int xs = pg * kPix;  // tile-local first pixel
    const bool active = (cs < p.CS) && (R0 + r < p.NR) && (d0 < p.D) && (X0 + xs < p.W);

    const int stage_cells = (p.TXc + p.RXc) * p.NL;
    const int line0 = (r << p.LOGCK) + (cs < p.CS ? cs : 0);
    int oL[2], oR[NCH];
    {
        const int jl = 2 * pg;
        oL[0] = swz(jl) * p.NL + line0;
        oL[1] = swz(jl + 1) * p.NL + line0;
        const int jr = (xs + p.HALO - (dgl + 1) * DT) >> 2;
#pragma unroll
        for (int k = 0; k < NCH; ++k) oR[k] = (p.TXc + swz(jr + k)) * p.NL + line0;
    }

    float acc[kPix][DT];
#pragma unroll
    for (int i = 0; i < kPix; ++i)
#pragma unroll
        for (int j = 0; j < DT; ++j) acc[i][j] = 0.f;

    const T *Lg = static_cast<const T *>(p.L);
    const T *Rg = static_cast<const T *>(p.R);
    const int nlines = p.ROWS << p.LOGCK;
    const int ncells_line = p.TXc + p.RXc;
    const int RX0 = X0 - p.HALO - DB0;  // global x of R-band local index 0

    // ---- per-lane fill descriptors: which chunks of a line this lane copies (line-invariant) ----
    int f_dst[kMaxFill];   // cell offset (without line), -1 = none
    int f_gx[kMaxFill];    // global x of the chunk's first element
    bool f_isR[kMaxFill];  // chunk belongs to the R band (else the L tile)
#pragma unroll
    for (int k = 0; k < kMaxFill; ++k) {
        const int j = lane + 32 * k;
        const bool isR = j >= p.TXc;
        const int jj = isR ? j - p.TXc : j;
        const int gx = (isR ? RX0 : X0) + 4 * jj;
        f_dst[k] = (j < ncells_line) ? ((isR ? p.TXc : 0) + swz(jj)) * p.NL : -1;
        f_gx[k] = gx;
        f_isR[k] = isR;
    }

    auto fill = [&](int stage, int kc) {
        float4 *st = smem + stage * stage_cells;
        for (int line = warp; line < nlines; line += nwarps) {
            const int rr = line >> p.LOGCK;
            const int c = (kc << p.LOGCK) + (line & (p.CK - 1));
            const long long rb = s_rowbase[rr];
            const bool lv = (rb >= 0) && (c < p.C);
            const long long off = lv ? rb + (long long)c * HW : 0;
            const T *lrow = Lg + off, *rrow = Rg + off;
#pragma unroll
            for (int k = 0; k < kMaxFill; ++k) {
                if (f_dst[k] < 0) break;
                float4 *dst = st + f_dst[k] + line;
                const T *row = f_isR[k] ? rrow : lrow;
                const int gx = f_gx[k];
                if constexpr (ASYNC) {
                    const bool ok = lv && gx >= 0 && gx < p.W;  // W % 4 == 0: chunk all-in or all-out
                    cp_async16(smem_u32(dst), ok ? (const void *)(row + gx) : (const void *)Lg, ok ? 16 : 0);
                } else {
                    *dst = load4_generic<T>(row, gx, p.W, lv);
                }
            }
        }
    };

    auto compute = [&](int stage) {
        const float4 *st = smem + stage * stage_cells;
#pragma unroll 2
        for (int cc = 0; cc < p.CK; cc += p.CS) {
            const float4 a0 = st[oL[0] + cc], a1 = st[oL[1] + cc];
            const float l[kPix] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float w[NCH * 4];
#pragma unroll
            for (int k = 0; k < NCH; ++k) {
                const float4 t = st[oR[k] + cc];
                w[4 * k + 0] = t.x; w[4 * k + 1] = t.y; w[4 * k + 2] = t.z; w[4 * k + 3] = t.w;
            }
            // w[i] = R[xs_global - d0 - DT + i]  =>  R[x - d] = w[px - dd + DT]
#pragma unroll
            for (int px = 0; px < kPix; ++px)
#pragma unroll
                for (int dd = 0; dd < DT; ++dd)
                    acc[px][dd] = fmaf(l[px], w[px - dd + DT], acc[px][dd]);
        }
    };

    if constexpr (ASYNC) {
#pragma unroll
        for (int s = 0; s < STAGES - 1; ++s) {
            if (s < p.NKC) fill(s, s);
            cp_async_commit();
        }
        for (int kc = 0; kc < p.NKC; ++kc) {
            cp_async_wait<STAGES - 2>();
            __syncthreads();
            const int nk = kc + STAGES - 1;
            if (nk < p.NKC) fill(nk % STAGES, nk);
            cp_async_commit();
            if (active) compute(kc % STAGES);
        }
        cp_async_wait<0>();
    } else {
        for (int kc = 0; kc < p.NKC; ++kc) {
            __syncthreads();
            fill(0, kc);
            __syncthreads();
            if (active) compute(0);
        }
    }

    // ---- reduce the channel-split partial sums through shared memory -----------------
    if (p.CS > 1) {
        constexpr int NV = kPix * DT / 4;  // float4s per thread
        __syncthreads();                   // pipeline buffers are free now
        float4 *red = smem;                // [cs-1][NV][NITEMS]
        if (active && cs > 0) {
            float4 *dst = red + (size_t)(cs - 1) * NV * p.NITEMS + item;
#pragma unroll
            for (int q = 0; q < NV; ++q) {
                const int e = 4 * q;
                dst[q * p.NITEMS] = make_float4(acc[e / DT][e % DT], acc[(e + 1) / DT][(e + 1) % DT],
                                               acc[(e + 2) / DT][(e + 2) % DT], acc[(e + 3) / DT][(e + 3) % DT]);
            }
        }
        __syncthreads();
        if (active && cs == 0) {
            for (int s = 1; s < p.CS; ++s) {
                const float4 *src = red + (size_t)(s - 1) * NV * p.NITEMS + item;
#pragma unroll
                for (int q = 0; q < NV; ++q) {
                    const float4 v = src[q * p.NITEMS];
                    const int e = 4 * q;
                    acc[e / DT][e % DT] += v.x;
                    acc[(e + 1) / DT][(e + 1) % DT] += v.y;
                    acc[(e + 2) / DT][(e + 2) % DT] += v.z;
                    acc[(e + 3) / DT][(e + 3) % DT] += v.w;
                }
            }
        }
    }
    if (!active || cs != 0) return;

    // ---- epilogue: mean over C, zero left border, store ----------------------------
    const int row = R0 + r;
    const int b = row / p.H, y = row - b * p.H;
    const int xg = X0 + xs;
    const float Cf = (float)p.C;
    const float invC = 1.0f / Cf;
    T *ob = static_cast<T *>(p.out) + (long long)b * p.OBS + (long long)y * p.W + xg;
    const bool vec = (sizeof(T) == 4) && ((p.W & 3) == 0) && (xg + kPix <= p.W) &&
                     ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0);
#pragma unroll
    for (int dd = 0; dd < DT; ++dd) {
        const int d = d0 + dd;
        if (d >= p.D) break;
        float v[kPix];
#pragma unroll
        for (int px = 0; px < kPix; ++px) {
            const float m = p.c_pow2 ? acc[px][dd] * invC : __fdiv_rn(acc[px][dd], Cf);
            v[px] = (xg + px >= d) ? m : 0.f;
        }
        T *o = ob + (long long)d * HW;
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

// ---------------------------------------------------------------------------------
// Narrow-tile variant (the default for D <= 64): a thread owns 8 pixels x 4
// disparities (32 accumulators, ~64 registers => 4x the resident warps of the wide
// tile), and the lanes of a warp run over the *disparity groups* of a pixel group
// first: the L chunks are then a warp-wide broadcast and the R chunks of neighbouring
// lanes are neighbouring 16-byte chunks, so no swizzle is needed and, with the line
// count NL a compile-time constant, every LDS.128 address is base + immediate.
// One image row per block; CS thread groups split the channels of a stage.
// ---------------------------------------------------------------------------------
template <typename T, int CK, int CS, bool ASYNC>
__global__ void __maxnreg__(96) corr_fwd_d4_kernel(const CorrFwdParams p) {
    extern __shared__ float4 smem[];
    constexpr int DT = 4;
    constexpr int NL = CK + 1;           // odd => conflict-free [chunk][line] layout
    constexpr int STAGES = ASYNC ? kStages : 1;
    constexpr int CPT = CK / CS;         // channels per thread per stage

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nwarps = blockDim.x >> 5;
    const int xt = blockIdx.x % p.XT;
    const int row = blockIdx.x / p.XT;
    const int X0 = xt * p.TX;
    const long long HW = (long long)p.H * p.W;
    const int b = row / p.H, y = row - b * p.H;
    const long long rowbase = (long long)b * p.C * HW + (long long)y * p.W;

    // work item: pixel group pg, disparity group dg (fastest), channel split cs
    const int item = tid % p.NITEMS;
    const int cs = tid / p.NITEMS;
    const int dg = item % p.NDGB;
    const int pg = item / p.NDGB;
    const int d0 = dg * DT;
    const int xs = pg * kPix;
    const bool active = (cs < CS) && (X0 + xs < p.W);

    const int stage_cells = (p.TXc + p.RXc) * NL;
    const int csafe = cs < CS ? cs : 0;
    const int oL = (2 * pg) * NL + csafe;                                        // + k*NL + i*CS
    const int oR = (p.TXc + 2 * pg + p.HALO / 4 - dg - 1) * NL + csafe;          // + k*NL + i*CS

    float acc[kPix][DT];
#pragma unroll
    for (int i = 0; i < kPix; ++i)
#pragma unroll
        for (int j = 0; j < DT; ++j) acc[i][j] = 0.f;

    const T *Lg = static_cast<const T *>(p.L) + rowbase;
    const T *Rg = static_cast<const T *>(p.R) + rowbase;
    const int ncells_line = p.TXc + p.RXc;
    const int RX0 = X0 - p.HALO;  // global x of R-band local index 0

    int f_dst[kMaxFill], f_gx[kMaxFill];
    bool f_isR[kMaxFill];
#pragma unroll
    for (int k = 0; k < kMaxFill; ++k) {
        const int j = lane + 32 * k;
        const bool isR = j >= p.TXc;
        const int jj = isR ? j - p.TXc : j;
        f_dst[k] = (j < ncells_line) ? j * NL : -1;
        f_gx[k] = (isR ? RX0 : X0) + 4 * jj;
        f_isR[k] = isR;
    }

    auto fill = [&](int stage, int kc) {
        float4 *st = smem + stage * stage_cells;
        for (int line = warp; line < CK; line += nwarps) {
            const int c = kc * CK + line;
            const bool lv = c < p.C;
            const long long off = lv ? (long long)c * HW : 0;
            const T *lrow = Lg + off, *rrow = Rg + off;
#pragma unroll
            for (int k = 0; k < kMaxFill; ++k) {
                if (f_dst[k] < 0) break;
                float4 *dst = st + f_dst[k] + line;
                const T *src = f_isR[k] ? rrow : lrow;
                const int gx = f_gx[k];
                if constexpr (ASYNC) {
                    const bool ok = lv && gx >= 0 && gx < p.W;
                    cp_async16(smem_u32(dst), ok ? (const void *)(src + gx) : (const void *)Lg, ok ? 16 : 0);
                } else {
                    *dst = load4_generic<T>(src, gx, p.W, lv);
                }
            }
        }
    };

    auto compute = [&](int stage) {
        const float4 *sl = smem + stage * stage_cells + oL;
        const float4 *sr = smem + stage * stage_cells + oR;
#pragma unroll 4
        for (int i = 0; i < CPT; ++i) {
            const float4 a0 = sl[i * CS], a1 = sl[NL + i * CS];
            const float4 w0 = sr[i * CS], w1 = sr[NL + i * CS], w2 = sr[2 * NL + i * CS];
            const float l[kPix] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float w[12] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x, w2.y, w2.z, w2.w};
            // w[i] = R[xs_global - d0 - 4 + i]  =>  R[x - d] = w[px - dd + 4]
#pragma unroll
            for (int px = 0; px < kPix; ++px)
#pragma unroll
                for (int dd = 0; dd < DT; ++dd)
                    acc[px][dd] = fmaf(l[px], w[px - dd + DT], acc[px][dd]);
        }
    };

    if constexpr (ASYNC) {
#pragma unroll
        for (int s = 0; s < STAGES - 1; ++s) {
            if (s < p.NKC) fill(s, s);
            cp_async_commit();
        }
        for (int kc = 0; kc < p.NKC; ++kc) {
            cp_async_wait<STAGES - 2>();
            __syncthreads();
            const int nk = kc + STAGES - 1;
            if (nk < p.NKC) fill(nk % STAGES, nk);
            cp_async_commit();
            if (active) compute(kc % STAGES);
        }
        cp_async_wait<0>();
    } else {
        for (int kc = 0; kc < p.NKC; ++kc) {
            __syncthreads();
            fill(0, kc);
            __syncthreads();
            if (active) compute(0);
        }
    }

    if constexpr (CS > 1) {
        constexpr int NV = kPix * DT / 4;
        __syncthreads();
        float4 *red = smem;  // [cs-1][NV][NITEMS]
        if (active && cs > 0) {
            float4 *dst = red + (size_t)(cs - 1) * NV * p.NITEMS + item;
#pragma unroll
            for (int q = 0; q < NV; ++q) dst[q * p.NITEMS] = make_float4(acc[q][0], acc[q][1], acc[q][2], acc[q][3]);
        }
        __syncthreads();
        if (active && cs == 0) {
#pragma unroll
            for (int s = 1; s < CS; ++s) {
                const float4 *src = red + (size_t)(s - 1) * NV * p.NITEMS + item;
#pragma unroll
                for (int q = 0; q < NV; ++q) {
                    const float4 v = src[q * p.NITEMS];
                    acc[q][0] += v.x; acc[q][1] += v.y; acc[q][2] += v.z; acc[q][3] += v.w;
                }
            }
        }
    }
    if (!active || cs != 0) return;

    const int xg = X0 + xs;
    const float Cf = (float)p.C;
    const float invC = 1.0f / Cf;
    T *ob = static_cast<T *>(p.out) + (long long)b * p.OBS + (long long)y * p.W + xg;
    const bool vec = (sizeof(T) == 4) && ((p.W & 3) == 0) && (xg + kPix <= p.W) &&
                     ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0);
#pragma unroll
    for (int dd = 0; dd < DT; ++dd) {
        const int d = d0 + dd;
        if (d >= p.D) break;
        float v[kPix];
#pragma unroll
        for (int px = 0; px < kPix; ++px) {
            const float m = p.c_pow2 ? acc[px][dd] * invC : __fdiv_rn(acc[px][dd], Cf);
            v[px] = (xg + px >= d) ? m : 0.f;
        }
        T *o = ob + (long long)d * HW;
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

// ---- fp64: plain one-thread-per-output kernel (arbiter precision, not a hot path) ----
__global__ void corr_fwd_f64_kernel(const double *L, const double *R, double *out,
                                    int B, int C, int H, int W, int D, long long obs) {
    const long long n = (long long)B * D * H * W;
    const long long HW = (long long)H * W;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % W);
        const int y = (int)((i / W) % H);
        const int d = (int)((i / HW) % D);
        const int b = (int)(i / (HW * D));
        double s = 0.0;
        if (x >= d) {
            const double *l = L + (long long)b * C * HW + (long long)y * W + x;
            const double *r = R + (long long)b * C * HW + (long long)y * W + x - d;
            for (int c = 0; c < C; ++c) s += l[c * HW] * r[c * HW];
            s /= (double)C;
        }
        out[(long long)b * obs + (i - (long long)b * HW * D)] = s;
    }
}

int env_int(const char *name, int dflt) {
    if (!tuning_enabled()) return dflt;
    const char *v = std::getenv(name);
    return (v && *v) ? std::atoi(v) : dflt;
}

template <typename T, int DT, bool ASYNC>
int launch_cfg(CorrFwdParams &p, size_t smem_bytes, dim3 grid, int threads, cudaStream_t stream) {
    auto kern = corr_fwd_kernel<T, DT, ASYNC>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) return fail(DSSMD_ERR_CUDA, "corr_fwd: cudaFuncSetAttribute(%zu B): %s", smem_bytes, cudaGetErrorString(e));
    kern<<<grid, threads, smem_bytes, stream>>>(p);
    return check_launch("corr_fwd");
}

template <typename T, int CK, int CS, bool ASYNC>
int launch_d4_cfg(CorrFwdParams &p, size_t smem_bytes, dim3 grid, int threads, cudaStream_t stream) {
    auto kern = corr_fwd_d4_kernel<T, CK, CS, ASYNC>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) return fail(DSSMD_ERR_CUDA, "corr_fwd: cudaFuncSetAttribute(%zu B): %s", smem_bytes, cudaGetErrorString(e));
    kern<<<grid, threads, smem_bytes, stream>>>(p);
    return check_launch("corr_fwd");
}

template <typename T, int CK, bool ASYNC>
int launch_d4_cs(CorrFwdParams &p, size_t smem_bytes, dim3 grid, int threads, cudaStream_t stream) {
    switch (p.CS) {
        case 1: return launch_d4_cfg<T, CK, 1, ASYNC>(p, smem_bytes, grid, threads, stream);
        case 2: return launch_d4_cfg<T, CK, 2, ASYNC>(p, smem_bytes, grid, threads, stream);
        case 4: return launch_d4_cfg<T, CK, 4, ASYNC>(p, smem_bytes, grid, threads, stream);
        default: return fail(DSSMD_ERR_UNSUPPORTED, "corr_fwd: channel split %d", p.CS);
    }
}

// narrow-tile path; returns -1 when the shape is outside its limits (caller falls back)
template <typename T>
int launch_d4(const void *L, const void *R, void *out, int B, int C, int H, int W, int D, long long obs, cudaStream_t stream) {
    CorrFwdParams p{};
    p.L = L; p.R = R; p.out = out; p.OBS = obs;
    p.B = B; p.C = C; p.H = H; p.W = W; p.D = D;
    p.NR = B * H;
    p.c_pow2 = (C & (C - 1)) == 0;
    const int max_threads = 224;  // __launch_bounds__(224, 3): 96 registers, 3 blocks per SM
    const int NDG = ceil_div(D, 4);
    if (NDG > 16) return -1;
    p.NDGB = NDG;
    p.HALO = 4 * NDG;
    const int NPGF = ceil_div(W, kPix);
    int npg_cap = max_threads / NDG;
    const int cap2 = (128 * 4 - p.HALO) / (2 * kPix);  // <= 128 chunks per line
    if (npg_cap > cap2) npg_cap = cap2;
    if (npg_cap < 1) return -1;
    // x tiles: split a row only when its work items exceed one 128-thread block (measured on
    // B200: 80x160/D=41 runs 1.6x faster as two 110-item tiles, model shapes are best unsplit)
    int xt = ceil_div(NPGF, npg_cap);
    while (ceil_div(NPGF, xt) * NDG > 128 && ceil_div(NPGF, xt + 1) >= 5) ++xt;
    xt = env_int("DSSMD_CORR_FWD_XT", xt);
    if (xt < ceil_div(NPGF, npg_cap)) xt = ceil_div(NPGF, npg_cap);
    if (xt > NPGF) xt = NPGF;
    p.XT = xt;
    p.NPG = ceil_div(NPGF, p.XT);
    p.TX = p.NPG * kPix;
    p.TXc = p.TX / 4;
    p.RXc = (p.TX + p.HALO) / 4;
    p.ROWS = 1;
    p.NITEMS = p.NPG * NDG;

    const long base_warps = ((long)p.NR * p.XT * p.NITEMS + 31) / 32;
    const long want_warps = (long)sm_count() * 16;
    int cs = 1;
    while (cs < 4 && base_warps * cs < want_warps && p.NITEMS * cs * 2 <= max_threads) cs *= 2;
    cs = env_int("DSSMD_CORR_FWD_CS", cs);
    if ((cs != 1 && cs != 2 && cs != 4) || p.NITEMS * cs > max_threads) cs = 1;
    p.CS = cs;
    const int threads = round_up(p.NITEMS * cs, 32);

    const bool async = (sizeof(T) == 4) && (W % 4 == 0) &&
                       ((reinterpret_cast<uintptr_t>(L) & 15) == 0) && ((reinterpret_cast<uintptr_t>(R) & 15) == 0);
    const int stages = async ? kStages : 1;
    int ck = env_int("DSSMD_CORR_FWD_CK", 0);
    if (ck != 8 && ck != 16) ck = ((size_t)stages * (p.TXc + p.RXc) * 17 * 16 <= 36 * 1024) ? 16 : 8;
    p.CK = ck;
    p.NL = ck + 1;
    size_t smem_bytes = (size_t)stages * (p.TXc + p.RXc) * p.NL * 16;
    const size_t red_bytes = cs > 1 ? (size_t)(cs - 1) * 8 * p.NITEMS * 16 : 0;
    if (red_bytes > smem_bytes) smem_bytes = red_bytes;
    p.NKC = ceil_div(C, ck);
    dim3 grid((unsigned)(p.XT * p.NR));
    if (env_int("DSSMD_CORR_DEBUG", 0))
        fprintf(stderr, "corr_fwd d4 cfg: NDG=%d NPG=%d XT=%d CS=%d CK=%d threads=%d grid=%u smem=%zu async=%d\n",
                NDG, p.NPG, p.XT, p.CS, p.CK, threads, grid.x, smem_bytes, (int)async);
    if (async) {
        if constexpr (sizeof(T) == 4) {
            return ck == 16 ? launch_d4_cs<T, 16, true>(p, smem_bytes, grid, threads, stream)
                            : launch_d4_cs<T, 8, true>(p, smem_bytes, grid, threads, stream);
        }
    }
    return ck == 16 ? launch_d4_cs<T, 16, false>(p, smem_bytes, grid, threads, stream)
                    : launch_d4_cs<T, 8, false>(p, smem_bytes, grid, threads, stream);
}

template <typename T>
int launch_typed(const void *L, const void *R, void *out, int B, int C, int H, int W, int D, long long obs, cudaStream_t stream) {
    // 0 wide FFMA tile, 1 register-blocked / narrow FFMA tile, 2 tcgen05 (3xTF32); unset = by size: the persistent
    // tensor-core kernel needs several 128-pixel tiles per SM to amortise its pipeline fill, and its work grows with
    // 128 + D per tile where the FFMA kernels' grows with D (measured, C256 D41 800 tiles: 87 vs 108 us;
    // C128 D24 960 tiles: 50 vs 42 us; 270 tiles: 21 vs 13 us)
    int variant = env_int("DSSMD_CORR_FWD_VARIANT", -1);
    if constexpr (sizeof(T) == 2) {
        // 16-bit tensors: mma.sync on the banded row GEMM unless a variant is forced
        if (variant < 0) {
            const int rc = corr_fwd_mma16(L, R, out, B, C, H, W, D, obs, std::is_same<T, __nv_bfloat16>::value ? DSSMD_BF16 : DSSMD_F16, stream);
            if (rc >= 0) return rc;
        }
    }
    if constexpr (sizeof(T) == 4) {
        if (variant == 4) {   // second tcgen05 design, forced
            const int rc = corr_fwd_umma2_f32(L, R, out, B, C, H, W, D, obs, stream);
            if (rc >= 0) return rc;
        }
    }
    if constexpr (sizeof(T) == 4) {
        // 3 = mma.sync with 3xTF32 splitting.  Unset: taken where it measured faster than the other fp32 kernels -- rows that fill the
        // register-blocked kernel's 128-pixel tiles badly (W = 80: 12.9 vs 18.5 us at 320x640 / 8) and problems of several thousand
        // 16-pixel warp tiles (cfg2 73.8 vs 86.2 us on tcgen05, KITTI 35.6 vs 41.6 us); the register-blocked kernel keeps the small
        // well-fitting shapes with few disparities.  Since the cross terms of the split run on bf16 MMAs and a stage holds 32 channels
        // (corr_fwd_mma32.cu) it also wins at the 576x960 model shape with D = 24 (10.3 vs 10.7 - 10.9 us; D = 12: 9.2 vs 8.9,
        // D = 5: 10.0 vs 8.9 -- scripts/cfg5_fwd_variants.py): D >= 20 goes there as well
        const long warp_tiles = (long)B * H * ceil_div(W, 16);
        const bool fills_reg_tiles = (double)W >= 0.75 * (double)(ceil_div(W, 128) * 128);
        if (variant == 3 || (variant < 0 && (!fills_reg_tiles || warp_tiles >= 3500 || (D >= 20 && C >= 32)))) {
            const int rc = corr_fwd_mma32(L, R, out, B, C, H, W, D, obs, stream);
            if (rc >= 0) return rc;
        }
    }
    // the tcgen05 kernels (2, 4) are opt-in: the mma.sync 3xTF32 kernel is faster on every shape measured (profiles/r02s_tcgen05_vs_mma32.md);
    // until round 2 a narrow window of problem sizes still selected variant 2 by itself
    if (variant < 0) variant = 1;
    if constexpr (sizeof(T) == 4) {
        if (variant == 2) {
            const int rc = corr_fwd_umma_f32(L, R, out, B, C, H, W, D, obs, stream);
            if (rc >= 0) return rc;
        }
    }
    if (variant >= 1 && env_int("DSSMD_CORR_FWD_REG", 1) != 0) {
        constexpr int dtype = sizeof(T) == 4 ? DSSMD_F32 : std::is_same<T, __nv_bfloat16>::value ? DSSMD_BF16 : DSSMD_F16;
        const int rc = corr_fwd_reg(L, R, out, B, C, H, W, D, obs, dtype, stream);
        if (rc >= 0) return rc;
    }
    if (variant >= 1 && D <= 64) {
        const int rc = launch_d4<T>(L, R, out, B, C, H, W, D, obs, stream);
        if (rc >= 0) return rc;
    }
    CorrFwdParams p{};
    p.L = L; p.R = R; p.out = out; p.OBS = obs;
    p.B = B; p.C = C; p.H = H; p.W = W; p.D = D;
    p.NR = B * H;
    p.c_pow2 = (C & (C - 1)) == 0;

    // disparities per thread: minimise shared-memory words read per pixel group
    int DT = 16;
    {
        long best = -1;
        for (int cand : {16, 12, 8}) {
            const long cost = (long)ceil_div(D, cand) * (cand + 16);
            if (best < 0 || cost < best) { best = cost; DT = cand; }
        }
        const int o = env_int("DSSMD_CORR_FWD_DT", 0);  // tuning override
        if (o == 8 || o == 12 || o == 16) DT = o;
    }
    const int max_threads = 256;
    const int NDG = ceil_div(D, DT);
    const int NPGF = ceil_div(W, kPix);

    // tile: all disparity groups of a pixel group live in one block (they share the R band)
    p.NDGB = NDG < 16 ? NDG : 16;  // HALO <= 256 keeps a line within 128 chunks
    p.HALO = round_up(p.NDGB * DT, 8);
    int npg_cap = (128 * 4 - p.HALO) / (2 * kPix);  // <= 128 chunks per line (fill descriptors)
    if (npg_cap > max_threads / p.NDGB) npg_cap = max_threads / p.NDGB;
    if (npg_cap < 1) npg_cap = 1;
    p.XT = ceil_div(NPGF, npg_cap);
    p.NPG = ceil_div(NPGF, p.XT);  // balance the tiles
    p.TX = p.NPG * kPix;
    p.TXc = p.TX / 4;
    p.RXc = (p.TX + p.HALO) / 4;
    const int items_row = p.NPG * p.NDGB;

    // parallelism: want >= 8 warps per SM; small problems split the channels over CS groups
    const long base_warps = ((long)p.NR * p.XT * ceil_div(NDG, p.NDGB) * items_row + 31) / 32;
    const long want_warps = (long)sm_count() * 8;
    int cs = 1;
    while (cs < 8 && base_warps * cs < want_warps && items_row * cs * 2 <= max_threads && cs * 2 <= C) cs *= 2;
    cs = env_int("DSSMD_CORR_FWD_CS", cs);
    if (cs < 1 || (cs & (cs - 1)) || items_row * cs > max_threads) cs = 1;
    p.CS = cs;
    p.LOGCS = 0;
    while ((1 << p.LOGCS) < cs) ++p.LOGCS;

    // rows per block: fill at least 64 threads, never more than one wave needs
    int rows = 64 / (items_row * cs);
    rows = env_int("DSSMD_CORR_FWD_ROWS", rows);
    if (rows < 1) rows = 1;
    if (rows > kMaxRows) rows = kMaxRows;
    if (rows > p.NR) rows = p.NR;
    while (rows > 1 && rows * items_row * cs > max_threads) --rows;
    p.ROWS = rows;
    p.NITEMS = rows * items_row;
    const int threads = round_up(p.NITEMS * cs, 32);

    const bool async = (sizeof(T) == 4) && (W % 4 == 0) &&
                       ((reinterpret_cast<uintptr_t>(L) & 15) == 0) && ((reinterpret_cast<uintptr_t>(R) & 15) == 0);
    const int stages = async ? kStages : 1;
    const size_t budget = (size_t)env_int("DSSMD_CORR_FWD_SMEM_KB", 40) * 1024;
    int ck = env_int("DSSMD_CORR_FWD_CK", 16);
    if (ck < 1 || (ck & (ck - 1))) ck = 16;
    while (ck > 1 && ck / 2 >= C && ck / 2 >= cs) ck /= 2;  // no point staging more channels than exist
    if (ck < cs) ck = cs;
    size_t smem_bytes = 0;
    for (;; ck /= 2) {
        const int nl = (p.ROWS * ck) | 1;
        smem_bytes = (size_t)stages * (p.TXc + p.RXc) * nl * 16;
        if (smem_bytes <= budget || ck <= cs) { p.CK = ck; p.NL = nl; break; }
    }
    const size_t red_bytes = cs > 1 ? (size_t)(cs - 1) * (kPix * DT / 4) * p.NITEMS * 16 : 0;
    if (red_bytes > smem_bytes) smem_bytes = red_bytes;
    if (smem_bytes > 200 * 1024)
        return fail(DSSMD_ERR_UNSUPPORTED, "corr_fwd: tile needs %zu B of shared memory (W=%d D=%d)", smem_bytes, W, D);
    p.LOGCK = 0;
    while ((1 << p.LOGCK) < p.CK) ++p.LOGCK;
    p.NKC = ceil_div(C, p.CK);

    dim3 grid((unsigned)(p.XT * ceil_div(p.NR, p.ROWS)), (unsigned)ceil_div(NDG, p.NDGB));
    if (env_int("DSSMD_CORR_DEBUG", 0))
        fprintf(stderr, "corr_fwd cfg: DT=%d NDGB=%d NPG=%d XT=%d ROWS=%d CS=%d CK=%d NL=%d threads=%d grid=(%u,%u) smem=%zu async=%d\n",
                DT, p.NDGB, p.NPG, p.XT, p.ROWS, p.CS, p.CK, p.NL, threads, grid.x, grid.y, smem_bytes, (int)async);
#define DSSMD_CORR_FWD_CASE(dt)                                                              \
    case dt:                                                                                 \
        return async ? launch_cfg<T, dt, true>(p, smem_bytes, grid, threads, stream)         \
                     : launch_cfg<T, dt, false>(p, smem_bytes, grid, threads, stream);
    if constexpr (sizeof(T) == 4) {
        switch (DT) {
            DSSMD_CORR_FWD_CASE(16)
            DSSMD_CORR_FWD_CASE(12)
            DSSMD_CORR_FWD_CASE(8)
        }
    } else {
        switch (DT) {
            case 16: return launch_cfg<T, 16, false>(p, smem_bytes, grid, threads, stream);
            case 12: return launch_cfg<T, 12, false>(p, smem_bytes, grid, threads, stream);
            case 8: return launch_cfg<T, 8, false>(p, smem_bytes, grid, threads, stream);
        }
    }
#undef DSSMD_CORR_FWD_CASE
    return fail(DSSMD_ERR_UNSUPPORTED, "corr_fwd: no kernel for DT=%d", DT);
}

}  // namespace

}  // namespace dssmd

namespace dssmd {
namespace {
int corr_fwd_entry(const void *L, const void *R, void *out, int B, int C, int H, int W, int D, long long obs, int dtype, cudaStream_t s) {
    switch (dtype) {
        case DSSMD_F32: return launch_typed<float>(L, R, out, B, C, H, W, D, obs, s);
        case DSSMD_BF16: return launch_typed<__nv_bfloat16>(L, R, out, B, C, H, W, D, obs, s);
        case DSSMD_F16: return launch_typed<__half>(L, R, out, B, C, H, W, D, obs, s);
        case DSSMD_F64: {
            const long long n = (long long)B * D * H * W;
            const int nt = 256;
            long long nb = (n + nt - 1) / nt;
            if (nb > 148LL * 32) nb = 148LL * 32;
            corr_fwd_f64_kernel<<<(unsigned)nb, nt, 0, s>>>((const double *)L, (const double *)R, (double *)out, B, C, H, W, D, obs);
            return check_launch("corr_fwd_f64");
        }
        default: return fail(DSSMD_ERR_INVALID_ARGUMENT, "corr_fwd: unknown dtype %d", dtype);
    }
}
int dtype_size(int dtype) { return dtype == DSSMD_F64 ? 8 : dtype == DSSMD_F32 ? 4 : 2; }
}  // namespace
}  // namespace dssmd

extern "C" int dssmd_corr_fwd(const void *L, const void *R, void *out, int B, int C, int H, int W, int D,
                              int dtype, void *stream) {
    using namespace dssmd;
    DSSMD_REQUIRE(L && R && out, "corr_fwd: null pointer");
    DSSMD_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && D > 0, "corr_fwd: sizes must be positive (B=%d C=%d H=%d W=%d D=%d)", B, C, H, W, D);
    DSSMD_REQUIRE((long long)B * H < (1LL << 31), "corr_fwd: B*H too large");
    return corr_fwd_entry(L, R, out, B, C, H, W, D, (long long)D * H * W, dtype, static_cast<cudaStream_t>(stream));
}

extern "C" int dssmd_corr_fwd_cat(const void *L, const void *R, void *cat, int B, int C, int H, int W, int D,
                                  int cat_channels, int channel_offset, int dtype, void *stream) {
    using namespace dssmd;
    DSSMD_REQUIRE(L && R && cat, "corr_fwd_cat: null pointer");
    DSSMD_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && D > 0, "corr_fwd_cat: sizes must be positive (B=%d C=%d H=%d W=%d D=%d)", B, C, H, W, D);
    DSSMD_REQUIRE((long long)B * H < (1LL << 31), "corr_fwd_cat: B*H too large");
    DSSMD_REQUIRE(channel_offset >= 0 && cat_channels >= channel_offset + D,
                  "corr_fwd_cat: channels [%d, %d) do not fit a buffer of %d channels", channel_offset, channel_offset + D, cat_channels);
    DSSMD_REQUIRE(dtype == DSSMD_F32 || dtype == DSSMD_BF16 || dtype == DSSMD_F16 || dtype == DSSMD_F64, "corr_fwd_cat: unknown dtype %d", dtype);
    char *o = static_cast<char *>(cat) + (size_t)channel_offset * H * W * dtype_size(dtype);
    return corr_fwd_entry(L, R, o, B, C, H, W, D, (long long)cat_channels * H * W, dtype, static_cast<cudaStream_t>(stream));
}
