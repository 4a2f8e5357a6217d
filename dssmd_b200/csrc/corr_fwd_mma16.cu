// corr_fwd_mma16.cu -- correlation forward for 16-bit tensors (bf16 / fp16) on the tensor cores.
//
//   out[b,d,y,x] = (1/C) * sum_c L[b,c,y,x] * R[b,c,y,x-d]        (upstream: models/UtilsNet/submoudles.py:301-311)
//
// With 16-bit inputs the products are exact in fp32 and no operand splitting is needed, so the banded row GEMM maps onto
// mma.sync.m16n8k16 directly: for 16 consecutive pixels x0 .. x0+15 of one image row,
//   S[m][n] = sum_c L[c][x0 + m] * R[c][xr0 + n],   xr0 = x0 - HALO  (HALO = D - 1 rounded up to 8),  n < HALO + 16
//   out[d][x0 + m] = S[m][m + HALO - d] / C
// i.e. a 16 x (HALO + 16) x C GEMM per warp whose diagonal band is the result (24 of 40 columns at D = 24, 41 of 56 at
// D = 41).  The FFMA kernels spend ~1 instruction per multiply-add and are bound by the FMA pipe at a fraction of the
// HBM roofline for 16-bit tensors (cfg2: 119 us for 113 MB); here a warp instruction does 2048 of them.
//
// Block = one image row x a tile of up to 256 pixels, one warp per 16 pixels.  The channel rows of L and R are staged
// [channel][pixel] in shared memory by cp.async (16-byte chunks; chunks left of x = 0 or past the row end are zero-filled:
// the reference's zero border), two stages of 32 channels.  Both operands are K-major in memory (a channel row holds
// consecutive pixels), which is the transpose of what the fragments want: ldmatrix.trans delivers them, and a row
// stride of an odd number of 16-byte chunks keeps its 8 row addresses on distinct banks.  Epilogue: the accumulator
// fragments scatter their band elements into a per-warp [D][16] staging tile, which goes out as 32-byte row segments.
#include "common.cuh"
#include "tma.cuh"

#include <type_traits>

namespace dssmd {

namespace {

constexpr int kCK = 32;          // channels per stage (two k16 steps)
constexpr int kMaxWarps = 16;    // 256 pixels per block

struct Mma16Params {
    const void *L;
    const void *R;
    void *out;
    long long OBS;   // elements between two images of out
    int C, H, W, D;
    int HALO;        // D - 1 rounded up to a multiple of 8
    int TX;          // pixels per block tile (multiple of 16)
    int XT;          // tiles per row
    int RW;          // staged R columns per tile: TX + 16 * NP - 16
    int LS, RS;      // row strides of the staged tiles in bytes (odd multiples of 16)
    float invC;
};

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}

template <typename T>
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    if constexpr (std::is_same<T, __nv_bfloat16>::value) {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    } else {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    }
}

// NP = pairs of 8-column tiles per warp (16 * NP >= HALO + 16)
// TMA = true: the two tiles of a stage arrive as one tensor-map box each (the boxes are an odd number of 16-byte chunks wide, so the
// dense box rows are the conflict-free stride; out-of-range columns / channels are zero-filled by the TMA unit); false: cp.async.
template <typename T, int NP, bool TMA>
__global__ void __launch_bounds__(kMaxWarps * 32) corr_fwd_mma16_kernel(const Mma16Params p, const __grid_constant__ CUtensorMap mapL,
                                                                        const __grid_constant__ CUtensorMap mapR) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long s_bar[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int row = blockIdx.x / p.XT, xt = blockIdx.x - row * p.XT;
    const int b = row / p.H, y = row - b * p.H;
    const int X0 = xt * p.TX;                 // first pixel of the block tile
    const size_t HW = (size_t)p.H * p.W;
    const T *Lg = static_cast<const T *>(p.L) + (size_t)b * p.C * HW + (size_t)y * p.W;
    const T *Rg = static_cast<const T *>(p.R) + (size_t)b * p.C * HW + (size_t)y * p.W;
    const uint32_t stage_bytes = (uint32_t)kCK * (uint32_t)(p.LS + p.RS);
    const uint32_t sbase = smem_u32(smem);
    const int lchunks = p.TX / 8, rchunks = p.RW / 8;     // 16-byte chunks per staged row
    const int nck = (p.C + kCK - 1) / kCK;

    // stage `s` <- channels [c0, c0 + kCK): L columns X0 .. X0 + TX, R columns X0 - HALO .. X0 - HALO + RW
    // stage `s` <- channels [c0, c0 + kCK): L columns X0 .. X0 + TX, R columns X0 - HALO .. X0 - HALO + RW.  Flat chunk index over
    // the block: every LDGSTS warp instruction carries 32 chunks (measured: a warp-per-row mapping with 20-27 of 32 lanes busy is
    // 1.7x slower -- the copy path is bound by LDGSTS issue, ~20 cycles per warp instruction)
    auto stage_in = [&](int c0, int s) {
        const uint32_t ls = sbase + (uint32_t)s * stage_bytes, rs = ls + (uint32_t)kCK * (uint32_t)p.LS;
        for (int i = tid; i < kCK * lchunks; i += blockDim.x) {
            const int c = i / lchunks, ch = i - c * lchunks;
            const int x = X0 + 8 * ch;
            const bool ok = (c0 + c < p.C) && (x < p.W);
            cp_async16(ls + (uint32_t)(c * p.LS + 16 * ch), ok ? (const void *)(Lg + (size_t)(c0 + c) * HW + x) : (const void *)Lg, ok ? 16 : 0);
        }
        for (int i = tid; i < kCK * rchunks; i += blockDim.x) {
            const int c = i / rchunks, ch = i - c * rchunks;
            const int x = X0 - p.HALO + 8 * ch;
            const bool ok = (c0 + c < p.C) && (x >= 0) && (x < p.W);
            cp_async16(rs + (uint32_t)(c * p.RS + 16 * ch), ok ? (const void *)(Rg + (size_t)(c0 + c) * HW + x) : (const void *)Rg, ok ? 16 : 0);
        }
        cp_async_commit();
    };

    float acc[2 * NP][4];
#pragma unroll
    for (int n = 0; n < 2 * NP; ++n)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[n][k] = 0.f;

    auto issue = [&](int c0, int s) {   // one thread
        const uint32_t ls = sbase + (uint32_t)s * stage_bytes, rs = ls + (uint32_t)kCK * (uint32_t)p.LS;
        const uint32_t bar = smem_u32(&s_bar[s]);
        fence_proxy_async();   // the stage was read through the generic proxy before the block barrier
        mbar_arrive_expect_tx(bar, stage_bytes);
        tma_load_4d(ls, &mapL, X0, y, c0, b, bar);
        tma_load_4d(rs, &mapR, X0 - p.HALO, y, c0, b, bar);
    };
    if constexpr (TMA) {
        if (tid == 0) {
            mbar_init(smem_u32(&s_bar[0]), 1);
            mbar_init(smem_u32(&s_bar[1]), 1);
            fence_proxy_async();
        }
        __syncthreads();
        if (tid == 0) issue(0, 0);
    } else {
        stage_in(0, 0);
    }
    // ldmatrix row addresses of this lane: matrix j = lane / 8, row r = lane % 8
    const int j = lane >> 3, r8 = lane & 7;
    const int wx = warp * 16;                              // this warp's first pixel inside the block tile
    // A (16 px x 16 ch): matrices (k 0-7, m 0-7), (k 0-7, m 8-15), (k 8-15, m 0-7), (k 8-15, m 8-15)
    const uint32_t a_off = (uint32_t)(((j >> 1) * 8 + r8) * p.LS + (wx + (j & 1) * 8) * 2);
    // B pair q (16 ch x 16 cols): matrices (k 0-7, n 0-7), (k 8-15, n 0-7), (k 0-7, n 8-15), (k 8-15, n 8-15) of columns wx + 16 q ..
    const uint32_t b_off = (uint32_t)(((j & 1) * 8 + r8) * p.RS + (wx + (j >> 1) * 8) * 2);

    for (int ck = 0; ck < nck; ++ck) {
        const int s = ck & 1;
        if constexpr (TMA) {
            if (tid == 0 && ck + 1 < nck) issue((ck + 1) * kCK, s ^ 1);   // stage s^1 was released by the barrier ending iteration ck-1
            mbar_wait(smem_u32(&s_bar[s]), (uint32_t)((ck >> 1) & 1));
        } else {
            if (ck + 1 < nck) { stage_in((ck + 1) * kCK, s ^ 1); cp_async_wait<1>(); }
            else cp_async_wait<0>();
            __syncthreads();
        }
        const uint32_t ls = sbase + (uint32_t)s * stage_bytes, rs = ls + (uint32_t)kCK * (uint32_t)p.LS;
#pragma unroll
        for (int k16 = 0; k16 < kCK / 16; ++k16) {
            uint32_t a[4];
            ldmatrix_x4_trans(ls + (uint32_t)(k16 * 16 * p.LS) + a_off, a[0], a[1], a[2], a[3]);
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                uint32_t b0, b1, b2, b3;
                ldmatrix_x4_trans(rs + (uint32_t)(k16 * 16 * p.RS) + b_off + (uint32_t)(q * 32), b0, b1, b2, b3);
                mma16816<T>(acc[2 * q], a, b0, b1);
                mma16816<T>(acc[2 * q + 1], a, b2, b3);
            }
        }
        __syncthreads();   // stage s may be refilled by the next iteration's stage_in
    }

    // ---- epilogue: band of the accumulator -> per-warp staging tile [D][17] (floats) -> 32-byte row segments -------------
    float *so = reinterpret_cast<float *>(smem) + (size_t)warp * p.D * 17;
    const int g = lane >> 2, tq = lane & 3;
#pragma unroll
    for (int n = 0; n < 2 * NP; ++n)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int m = g + ((k & 2) ? 8 : 0);
            const int col = n * 8 + 2 * tq + (k & 1);
            const int d = m + p.HALO - col;
            if (d >= 0 && d < p.D) so[d * 17 + m] = acc[n][k] * p.invC;
        }
    __syncwarp();
    const int x = X0 + wx + (lane & 15);
    if (warp < nwarps && x < p.W) {
        T *o = static_cast<T *>(p.out) + (size_t)b * p.OBS + (size_t)y * p.W + x;
        for (int d = lane >> 4; d < p.D; d += 2) o[(size_t)d * HW] = from_f32<T>(so[d * 17 + (lane & 15)]);
    }
}

int env_int_m16(const char *name, int dflt) {
    if (!tuning_enabled()) return dflt;
    const char *v = std::getenv(name);
    return (v && *v) ? std::atoi(v) : dflt;
}

template <typename T, int NP, bool TMA>
int launch_np2(const Mma16Params &p, const CUtensorMap &mapL, const CUtensorMap &mapR, int grid, int threads, size_t smem, cudaStream_t stream) {
    auto kern = corr_fwd_mma16_kernel<T, NP, TMA>;
    if (smem > 48 * 1024) {
        const cudaError_t e = kernel_smem_optin(kern, smem);
        if (e != cudaSuccess) return fail(DSSMD_ERR_CUDA, "corr_fwd(mma16): cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(e));
    }
    kern<<<grid, threads, smem, stream>>>(p, mapL, mapR);
    return check_launch("corr_fwd(mma16)");
}

template <typename T, int NP>
int launch_np(const Mma16Params &p, bool tma, const CUtensorMap &mapL, const CUtensorMap &mapR, int grid, int threads, size_t smem, cudaStream_t stream) {
    return tma ? launch_np2<T, NP, true>(p, mapL, mapR, grid, threads, smem, stream) : launch_np2<T, NP, false>(p, mapL, mapR, grid, threads, smem, stream);
}

template <typename T>
int corr_fwd_mma16_typed(const void *L, const void *R, void *out, int B, int C, int H, int W, int D, long long obs, cudaStream_t stream) {
    if (W % 8 != 0 || D > 81 || C < 16) return -1;
    if (((reinterpret_cast<uintptr_t>(L) | reinterpret_cast<uintptr_t>(R)) & 15) != 0) return -1;
    Mma16Params p{};
    p.L = L; p.R = R; p.out = out; p.OBS = obs;
    p.C = C; p.H = H; p.W = W; p.D = D;
    p.invC = 1.0f / (float)C;
    p.HALO = round_up(D - 1, 8);
    const int np = ceil_div(p.HALO + 16, 16);            // pairs of 8-column tiles per warp
    if (np < 1 || np > 6) return -1;
    int warps = ceil_div(W, 16);
    int maxw = env_int_m16("DSSMD_MMA16_MAXW", 8);   // measured: two 80-pixel tiles per 160-pixel row beat one (39.8 vs 43.7 us at cfg2)
    if (maxw < 1 || maxw > kMaxWarps) maxw = kMaxWarps;
    p.XT = ceil_div(warps, maxw);
    warps = ceil_div(warps, p.XT);                       // balance the tiles of a row
    p.TX = warps * 16;
    p.RW = p.TX + 16 * np - 16;
    auto odd16 = [](int bytes) { int c = (bytes + 15) / 16; if ((c & 1) == 0) ++c; return c * 16; };
    p.LS = odd16(p.TX * 2);
    p.RS = odd16(p.RW * 2);
    // TMA staging: boxes as wide as the padded rows (<= 256 elements), rows of W*2 bytes a multiple of 16 (W % 8 == 0)
    CUtensorMap mapL{}, mapR{};
    bool tma = env_int_m16("DSSMD_MMA16_TMA", 1) != 0 && p.LS / 2 <= 256 && p.RS / 2 <= 256 && (reinterpret_cast<uintptr_t>(out) & 1) == 0;
    if (tma) {
        const CUtensorMapDataType tmt = std::is_same<T, __nv_bfloat16>::value ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
        const cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)B};
        const cuuint64_t strides[3] = {(cuuint64_t)W * 2, (cuuint64_t)H * W * 2, (cuuint64_t)C * H * W * 2};
        const cuuint32_t boxl[4] = {(cuuint32_t)(p.LS / 2), 1, (cuuint32_t)kCK, 1}, boxr[4] = {(cuuint32_t)(p.RS / 2), 1, (cuuint32_t)kCK, 1};
        tma = make_tmap(&mapL, tmt, L, 4, dims, strides, boxl, CU_TENSOR_MAP_SWIZZLE_NONE) &&
              make_tmap(&mapR, tmt, R, 4, dims, strides, boxr, CU_TENSOR_MAP_SWIZZLE_NONE);
    }
    size_t smem = 2 * (size_t)kCK * (size_t)(p.LS + p.RS);
    const size_t epi = (size_t)warps * D * 17 * sizeof(float);
    if (epi > smem) smem = epi;
    if (smem > 200 * 1024) return -1;
    const long long blocks = (long long)B * H * p.XT;
    if (blocks >= (1LL << 31)) return -1;
    if (env_int_m16("DSSMD_CORR_DEBUG", 0))
        fprintf(stderr, "corr_fwd mma16 cfg: NP=%d HALO=%d TX=%d XT=%d RW=%d LS=%d RS=%d warps=%d grid=%lld smem=%zu tma=%d\n", np, p.HALO, p.TX, p.XT,
                p.RW, p.LS, p.RS, warps, blocks, smem, (int)tma);
    switch (np) {
        case 1: return launch_np<T, 1>(p, tma, mapL, mapR, (int)blocks, warps * 32, smem, stream);
        case 2: return launch_np<T, 2>(p, tma, mapL, mapR, (int)blocks, warps * 32, smem, stream);
        case 3: return launch_np<T, 3>(p, tma, mapL, mapR, (int)blocks, warps * 32, smem, stream);
        case 4: return launch_np<T, 4>(p, tma, mapL, mapR, (int)blocks, warps * 32, smem, stream);
        case 5: return launch_np<T, 5>(p, tma, mapL, mapR, (int)blocks, warps * 32, smem, stream);
        default: return launch_np<T, 6>(p, tma, mapL, mapR, (int)blocks, warps * 32, smem, stream);
    }
}

}  // namespace

// dtype: DSSMD_BF16 / DSSMD_F16.  -1 = shape not eligible (W % 8 != 0, D > 81, C < 16, unaligned): the caller falls back.
int corr_fwd_mma16(const void *L, const void *R, void *out, int B, int C, int H, int W, int D, long long obs, int dtype, cudaStream_t stream) {
    if (env_int_m16("DSSMD_CORR_FWD_MMA16", 1) == 0) return -1;
    if (dtype == DSSMD_BF16) return corr_fwd_mma16_typed<__nv_bfloat16>(L, R, out, B, C, H, W, D, obs, stream);
    if (dtype == DSSMD_F16) return corr_fwd_mma16_typed<__half>(L, R, out, B, C, H, W, D, obs, stream);
    return -1;
}

}  // namespace dssmd
