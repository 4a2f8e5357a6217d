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
//   * shared memory is laid out [16-byte chunk][line] with an XOR swizzle on
//     the chunk index so that 8 lanes reading chunks 2*lane+k hit 8 distinct
//     bank groups, and with an odd line count so the fill is conflict-free.
// Algorithmic HBM bytes per launch: (2*B*C*H*W + B*D*H*W) * sizeof(T).
#include "common.cuh"

namespace dssmd {

namespace {

constexpr int kPix = 8;        // pixels per thread
constexpr int kMaxRows = 16;   // rows per block upper bound
constexpr int kStages = 3;     // cp.async pipeline depth

struct CorrFwdParams {
    const void *L;
    const void *R;
    void *out;
    int B, C, H, W, D;
    int NR;     // B*H image rows
    int TX;     // tile width in pixels (multiple of 8)
    int NPG;    // pixel groups per tile = TX/8
    int NDGB;   // disparity groups per block
    int ROWS;   // rows per block
    int XT;     // x tiles per row
    int CK;     // channels per stage (power of two)
    int LOGCK;
    int NL;     // lines (= ROWS*CK) padded to an odd count
    int TXc;    // 16-byte chunks of the L tile per line
    int RXc;    // 16-byte chunks of the R band per line
    int HALO;   // left halo of the R band in pixels (multiple of 8)
    int NKC;    // channel chunks = ceil(C/CK)
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

template <typename T, int DT, int NT, bool ASYNC>
__global__ void __launch_bounds__(NT) corr_fwd_kernel(const CorrFwdParams p) {
    extern __shared__ float4 smem[];
    __shared__ long long s_rowbase[kMaxRows];  // element offset of (b, c=0, y, x=0); -1 = no such row

    constexpr int NCH = (DT + 8) / 4;  // R chunks per thread per channel
    constexpr int NWARPS = NT / 32;
    constexpr int STAGES = ASYNC ? kStages : 1;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
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

    // ---- this thread's work item ------------------------------------------------
    const int pg = tid % p.NPG;
    const int t2 = tid / p.NPG;
    const int dgl = t2 % p.NDGB;
    const int r = t2 / p.NDGB;
    const int d0 = DB0 + dgl * DT;
    const int xs = pg * kPix;  // tile-local first pixel
    const bool active = (r < p.ROWS) && (R0 + r < p.NR) && (d0 < p.D) && (X0 + xs < p.W);

    const int stage_cells = (p.TXc + p.RXc) * p.NL;
    const int line0 = (active ? r : 0) << p.LOGCK;
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

    auto fill = [&](int stage, int kc) {
        float4 *st = smem + stage * stage_cells;
        for (int line = warp; line < nlines; line += NWARPS) {
            const int rr = line >> p.LOGCK;
            const int c = (kc << p.LOGCK) + (line & (p.CK - 1));
            const long long rb = s_rowbase[rr];
            const bool lv = (rb >= 0) && (c < p.C);
            const long long off = lv ? rb + (long long)c * HW : 0;
            const T *lrow = Lg + off, *rrow = Rg + off;
            for (int j = lane; j < ncells_line; j += 32) {
                const bool isR = j >= p.TXc;
                const int jj = isR ? j - p.TXc : j;
                const int gx = (isR ? RX0 : X0) + 4 * jj;
                float4 *dst = st + ((isR ? p.TXc : 0) + swz(jj)) * p.NL + line;
                const T *row = isR ? rrow : lrow;
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
        for (int cc = 0; cc < p.CK; ++cc) {
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

    if (!active) return;

    // ---- epilogue: mean over C, zero left border, store ----------------------------
    const int row = R0 + r;
    const int b = row / p.H, y = row - b * p.H;
    const int xg = X0 + xs;
    const float Cf = (float)p.C;
    const float invC = 1.0f / Cf;
    T *ob = static_cast<T *>(p.out) + ((long long)b * p.D * p.H + y) * p.W + xg;
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
                                    int B, int C, int H, int W, int D) {
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
        out[i] = s;
    }
}

template <typename T, int DT, bool ASYNC>
int launch_cfg(CorrFwdParams &p, size_t smem_bytes, dim3 grid, cudaStream_t stream) {
    constexpr int NT = 128;
    auto kern = corr_fwd_kernel<T, DT, NT, ASYNC>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) return fail(DSSMD_ERR_CUDA, "corr_fwd: cudaFuncSetAttribute(%zu B): %s", smem_bytes, cudaGetErrorString(e));
    kern<<<grid, NT, smem_bytes, stream>>>(p);
    return check_launch("corr_fwd");
}

template <typename T>
int launch_typed(const void *L, const void *R, void *out, int B, int C, int H, int W, int D, cudaStream_t stream) {
    constexpr int NT = 128;
    CorrFwdParams p{};
    p.L = L; p.R = R; p.out = out;
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
    }
    const int NDG = ceil_div(D, DT);
    const int NPGF = ceil_div(W, kPix);
    p.NDGB = NDG < NT / 4 ? NDG : NT / 4;
    p.NPG = NPGF < NT / p.NDGB ? NPGF : NT / p.NDGB;
    p.TX = p.NPG * kPix;
    p.XT = ceil_div(NPGF, p.NPG);
    int rows = NT / (p.NPG * p.NDGB);
    if (rows < 1) rows = 1;
    if (rows > kMaxRows) rows = kMaxRows;
    if (rows > p.NR) rows = p.NR;
    p.ROWS = rows;
    p.HALO = round_up(p.NDGB * DT, 8);
    p.TXc = p.TX / 4;
    p.RXc = (p.TX + p.HALO) / 4;

    const bool async = (sizeof(T) == 4) && (W % 4 == 0) &&
                       ((reinterpret_cast<uintptr_t>(L) & 15) == 0) && ((reinterpret_cast<uintptr_t>(R) & 15) == 0);
    const int stages = async ? kStages : 1;
    const size_t budget = 100 * 1024;
    int ck = 16;
    while (ck > 1 && ck / 2 >= C) ck /= 2;  // no point staging more channels than exist
    size_t smem_bytes = 0;
    for (;; ck /= 2) {
        const int nl = (p.ROWS * ck) | 1;
        smem_bytes = (size_t)stages * (p.TXc + p.RXc) * nl * 16;
        if (smem_bytes <= budget || ck == 1) { p.CK = ck; p.NL = nl; break; }
    }
    if (smem_bytes > 200 * 1024)
        return fail(DSSMD_ERR_UNSUPPORTED, "corr_fwd: tile needs %zu B of shared memory (W=%d D=%d)", smem_bytes, W, D);
    p.LOGCK = 0;
    while ((1 << p.LOGCK) < p.CK) ++p.LOGCK;
    p.NKC = ceil_div(C, p.CK);

    dim3 grid((unsigned)(p.XT * ceil_div(p.NR, p.ROWS)), (unsigned)ceil_div(NDG, p.NDGB));
#define DSSMD_CORR_FWD_CASE(dt)                                                              \
    case dt:                                                                                 \
        return async ? launch_cfg<T, dt, true>(p, smem_bytes, grid, stream)                  \
                     : launch_cfg<T, dt, false>(p, smem_bytes, grid, stream);
    if constexpr (sizeof(T) == 4) {
        switch (DT) {
            DSSMD_CORR_FWD_CASE(16)
            DSSMD_CORR_FWD_CASE(12)
            DSSMD_CORR_FWD_CASE(8)
        }
    } else {
        switch (DT) {
            case 16: return launch_cfg<T, 16, false>(p, smem_bytes, grid, stream);
            case 12: return launch_cfg<T, 12, false>(p, smem_bytes, grid, stream);
            case 8: return launch_cfg<T, 8, false>(p, smem_bytes, grid, stream);
        }
    }
#undef DSSMD_CORR_FWD_CASE
    return fail(DSSMD_ERR_UNSUPPORTED, "corr_fwd: no kernel for DT=%d", DT);
}

}  // namespace

}  // namespace dssmd

extern "C" int dssmd_corr_fwd(const void *L, const void *R, void *out, int B, int C, int H, int W, int D,
                              int dtype, void *stream) {
    using namespace dssmd;
    DSSMD_REQUIRE(L && R && out, "corr_fwd: null pointer");
    DSSMD_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && D > 0, "corr_fwd: sizes must be positive (B=%d C=%d H=%d W=%d D=%d)", B, C, H, W, D);
    DSSMD_REQUIRE((long long)B * H < (1LL << 31), "corr_fwd: B*H too large");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    switch (dtype) {
        case DSSMD_F32: return launch_typed<float>(L, R, out, B, C, H, W, D, s);
        case DSSMD_BF16: return launch_typed<__nv_bfloat16>(L, R, out, B, C, H, W, D, s);
        case DSSMD_F16: return launch_typed<__half>(L, R, out, B, C, H, W, D, s);
        case DSSMD_F64: {
            const long long n = (long long)B * D * H * W;
            const int nt = 256;
            long long nb = (n + nt - 1) / nt;
            if (nb > 148LL * 32) nb = 148LL * 32;
            corr_fwd_f64_kernel<<<(unsigned)nb, nt, 0, s>>>((const double *)L, (const double *)R, (double *)out, B, C, H, W, D);
            return check_launch("corr_fwd_f64");
        }
        default: return fail(DSSMD_ERR_INVALID_ARGUMENT, "corr_fwd: unknown dtype %d", dtype);
    }
}
