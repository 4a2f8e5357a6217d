// corr_bwd_mma16.cu -- backward of the correlation volume for 16-bit tensors (bf16 / fp16) on the tensor cores.
//
//   gL[c,x ] = (1/C) * sum_d g[d,x]    * R[c,x-d]        (MODE 0)
//   gR[c,x'] = (1/C) * sum_d g[d,x'+d] * L[c,x'+d]       (MODE 1)       (the graph autograd builds from submoudles.py:304-308)
//
// For 16 consecutive pixels x0 .. x0+15 of one image row both are a 16 x K x C GEMM with a BANDED left operand built
// from g (K = HALO + 16 operand positions, HALO = D - 1 rounded up to 8, K rounded up to 16):
//   MODE 0:  gL[x0+m][c] = sum_k G[m][k] * R[c][x0 - HALO + k],   G[m][k] = g[m + HALO - k][x0 + m]   (0 where d is outside [0, D))
//   MODE 1:  gR[x0+m][c] = sum_k G[m][k] * L[c][x0 + k],          G[m][k] = g[k - m][x0 + k]
// G does not depend on the channel: a warp builds its A fragments ONCE from the staged g tile and keeps them in registers
// (the role the g slab plays in corr_bwd_reg.cu), then walks over the channels 16 at a time: the operand rows are
// [channel][position] in shared memory, which is exactly the "col" B operand of mma.sync.m16n8k16 (two consecutive
// positions of one channel per register): plain ldmatrix, no transpose.  Products of 16-bit values are exact in fp32 and the
// accumulation is fp32.  Block = one image row x a tile of up to 256 pixels, a warp per 16 pixels; operand rows staged by
// cp.async (zero fill outside the row = the x-d < 0 / x'+d >= W limits of the sums), two stages of 32 channels.
#include "common.cuh"
#include "tma.cuh"

#include <type_traits>

namespace dssmd {

namespace {

constexpr int kCK = 32;          // channels per stage
constexpr int kMaxWarps = 16;

struct BwdMma16Params {
    const void *g;
    const void *In;      // R (MODE 0) or L (MODE 1)
    void *out;           // gL or gR
    long long GBS;       // elements between two images of g
    int C, H, W, D;
    int HALO;            // D - 1 rounded up to a multiple of 8
    int TX, XT;          // pixels per block tile (multiple of 16), tiles per row
    int RW;              // staged operand positions per tile: TX + 16 * NP - 16
    int GWD;             // staged g columns: TX (MODE 0) or RW (MODE 1)
    int RS, GS;          // row strides in bytes of the operand tile (odd multiple of 16) and of the g tile
    float invC;
};

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}

template <typename T>
__device__ __forceinline__ void mma16816b(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    if constexpr (std::is_same<T, __nv_bfloat16>::value) {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    } else {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    }
}

__device__ __forceinline__ void store_pair(__nv_bfloat16 *o, float a, float b) { *reinterpret_cast<__nv_bfloat162 *>(o) = __floats2bfloat162_rn(a, b); }
__device__ __forceinline__ void store_pair(__half *o, float a, float b) { *reinterpret_cast<__half2 *>(o) = __floats2half2_rn(a, b); }

// TMA = true: the g tile and the operand rows of a stage arrive as tensor-map boxes (dense rows of an odd number of 16-byte chunks,
// zero fill outside the row); false: cp.async
template <typename T, int MODE, int NP, bool TMA>
__global__ void __launch_bounds__(kMaxWarps * 32) corr_bwd_mma16_kernel(const BwdMma16Params p, const __grid_constant__ CUtensorMap mapIn,
                                                                        const __grid_constant__ CUtensorMap mapG) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long s_bar[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row = blockIdx.x / p.XT, xt = blockIdx.x - row * p.XT;
    const int b = row / p.H, y = row - b * p.H;
    const int X0 = xt * p.TX;
    const size_t HW = (size_t)p.H * p.W;
    const T *Gg = static_cast<const T *>(p.g) + (size_t)b * p.GBS + (size_t)y * p.W;
    const T *Ig = static_cast<const T *>(p.In) + (size_t)b * p.C * HW + (size_t)y * p.W;
    const uint32_t g_bytes = (uint32_t)p.D * (uint32_t)p.GS;
    const uint32_t g_pad = (g_bytes + 127u) & ~127u;      // the operand stages behind the g tile stay 128-byte aligned (TMA)
    const uint32_t stage_bytes = (uint32_t)kCK * (uint32_t)p.RS;
    const uint32_t sbase = smem_u32(smem);                 // g tile | stage 0 | stage 1
    const int rchunks = p.RW / 8, gchunks = p.GWD / 8;
    const int xin0 = MODE == 0 ? X0 - p.HALO : X0;        // position of operand column 0
    const int nck = (p.C + kCK - 1) / kCK;

    // flat chunk index over the block: every LDGSTS warp instruction carries 32 chunks (see corr_fwd_mma16.cu)
    auto stage_in = [&](int c0, int s) {
        const uint32_t rs = sbase + g_pad + (uint32_t)s * stage_bytes;
        for (int i = tid; i < kCK * rchunks; i += blockDim.x) {
            const int c = i / rchunks, ch = i - c * rchunks;
            const int x = xin0 + 8 * ch;
            const bool ok = (c0 + c < p.C) && (x >= 0) && (x < p.W);
            cp_async16(rs + (uint32_t)(c * p.RS + 16 * ch), ok ? (const void *)(Ig + (size_t)(c0 + c) * HW + x) : (const void *)Ig, ok ? 16 : 0);
        }
        cp_async_commit();
    };
    auto issue = [&](int c0, int s, bool with_g) {   // one thread
        const uint32_t rs = sbase + g_pad + (uint32_t)s * stage_bytes;
        const uint32_t bar = smem_u32(&s_bar[s]);
        fence_proxy_async();
        mbar_arrive_expect_tx(bar, stage_bytes + (with_g ? g_bytes : 0u));
        tma_load_4d(rs, &mapIn, xin0, y, c0, b, bar);
        if (with_g) tma_load_4d(sbase, &mapG, X0, y, 0, b, bar);
    };
    if constexpr (TMA) {
        if (tid == 0) {
            mbar_init(smem_u32(&s_bar[0]), 1);
            mbar_init(smem_u32(&s_bar[1]), 1);
            fence_proxy_async();
        }
        __syncthreads();
        if (tid == 0) issue(0, 0, true);
    } else {
        // g tile: D rows of columns X0 .. X0 + GWD (zero past the row end), in the first cp.async group together with stage 0
        for (int i = tid; i < p.D * gchunks; i += blockDim.x) {
            const int d = i / gchunks, ch = i - d * gchunks;
            const int x = X0 + 8 * ch;
            const bool ok = x < p.W;
            cp_async16(sbase + (uint32_t)(d * p.GS + 16 * ch), ok ? (const void *)(Gg + (size_t)d * HW + x) : (const void *)Gg, ok ? 16 : 0);
        }
        stage_in(0, 0);
    }

    const int g8 = lane >> 2, tq = lane & 3;
    const int wx = warp * 16;
    uint32_t afr[NP][4];          // the banded operand G of this warp's 16 pixels, all K steps
    // The product is issued transposed -- channels as the M dimension, pixels as N -- so that an accumulator register pair holds two
    // NEIGHBOURING pixels of one channel (one packed 4-byte store, four lanes = 16 contiguous bytes of a channel row):
    //   A (16 channels x 16 positions) = the staged rows as they lie: matrices (m 0-7, k 0-7), (m 8-15, k 0-7), (m 0-7, k 8-15), (m 8-15, k 8-15)
    //   B (16 positions x 8 pixels)    = G^T: the "col" fragment (k = 2 tq .., n = g8) is G[pixel g8][k ..], i.e. afr[q][0], afr[q][2] for
    //                                    pixels 0-7 and afr[q][1], afr[q][3] for pixels 8-15
    const int j = lane >> 3, r8 = lane & 7;
    const uint32_t a_off = (uint32_t)(((j & 1) * 8 + r8) * p.RS + (wx + (j >> 1) * 8) * 2);
    T *const ob = static_cast<T *>(p.out) + (size_t)b * p.C * HW + (size_t)y * p.W + X0 + wx;
    const bool st0 = X0 + wx + 2 * tq < p.W, st1 = X0 + wx + 8 + 2 * tq < p.W;   // W % 8 == 0: a pixel pair is in or out together

    for (int ck = 0; ck < nck; ++ck) {
        const int s = ck & 1;
        if constexpr (TMA) {
            if (tid == 0 && ck + 1 < nck) issue((ck + 1) * kCK, s ^ 1, false);
            mbar_wait(smem_u32(&s_bar[s]), (uint32_t)((ck >> 1) & 1));
        } else {
            if (ck + 1 < nck) { stage_in((ck + 1) * kCK, s ^ 1); cp_async_wait<1>(); }
            else cp_async_wait<0>();
            __syncthreads();
        }
        if (ck == 0) {
            // A fragments from the g tile: a0 (m = g8, k = 2 tq ..), a1 (m = g8 + 8), a2 (k + 8), a3 (m + 8, k + 8)
            const unsigned short *gt = reinterpret_cast<const unsigned short *>(smem);
            const int gs2 = p.GS / 2;
#pragma unroll
            for (int q = 0; q < NP; ++q)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int m = g8 + ((i & 1) ? 8 : 0);
                    uint32_t v = 0;
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int k = 16 * q + ((i & 2) ? 8 : 0) + 2 * tq + e;
                        const int d = MODE == 0 ? m + p.HALO - k : k - m;
                        const int xc = MODE == 0 ? wx + m : wx + k;            // column of the g tile
                        const unsigned short h = (d >= 0 && d < p.D) ? gt[d * gs2 + xc] : (unsigned short)0;
                        v |= (uint32_t)h << (16 * e);
                    }
                    afr[q][i] = v;
                }
        }
        const uint32_t rs = sbase + g_pad + (uint32_t)s * stage_bytes;
        const int c0 = ck * kCK;
#pragma unroll
        for (int n16 = 0; n16 < kCK / 16; ++n16) {
            float acc[2][4];
#pragma unroll
            for (int t = 0; t < 2; ++t)
#pragma unroll
                for (int k = 0; k < 4; ++k) acc[t][k] = 0.f;
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                uint32_t a[4];
                ldmatrix_x4(rs + (uint32_t)(n16 * 16 * p.RS) + a_off + (uint32_t)(q * 32), a[0], a[1], a[2], a[3]);
                mma16816b<T>(acc[0], a, afr[q][0], afr[q][2]);
                mma16816b<T>(acc[1], a, afr[q][1], afr[q][3]);
            }
            // acc[t]: c0/c1 = (channel g8, pixels 8 t + 2 tq, + 1), c2/c3 = (channel g8 + 8, same pixels)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int c = c0 + n16 * 16 + hh * 8 + g8;
                if (c < p.C) {
                    T *o = ob + (size_t)c * HW + 2 * tq;
                    if (st0) store_pair(o, acc[0][2 * hh] * p.invC, acc[0][2 * hh + 1] * p.invC);
                    if (st1) store_pair(o + 8, acc[1][2 * hh] * p.invC, acc[1][2 * hh + 1] * p.invC);
                }
            }
        }
        __syncthreads();
    }
}

int env_int_bm16(const char *name, int dflt) {
    if (!tuning_enabled()) return dflt;
    const char *v = std::getenv(name);
    return (v && *v) ? std::atoi(v) : dflt;
}

template <typename T, int MODE, int NP, bool TMA>
int launch_bnp2(const BwdMma16Params &p, const CUtensorMap &mapIn, const CUtensorMap &mapG, int grid, int threads, size_t smem, cudaStream_t stream) {
    auto kern = corr_bwd_mma16_kernel<T, MODE, NP, TMA>;
    if (smem > 48 * 1024) {
        const cudaError_t e = kernel_smem_optin(kern, smem);
        if (e != cudaSuccess) return fail(DSSMD_ERR_CUDA, "corr_bwd(mma16): cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(e));
    }
    kern<<<grid, threads, smem, stream>>>(p, mapIn, mapG);
    return check_launch(MODE == 0 ? "corr_bwd(gL, mma16)" : "corr_bwd(gR, mma16)");
}

template <typename T, int MODE, int NP>
int launch_bnp(const BwdMma16Params &p, bool tma, const CUtensorMap &mapIn, const CUtensorMap &mapG, int grid, int threads, size_t smem, cudaStream_t stream) {
    return tma ? launch_bnp2<T, MODE, NP, true>(p, mapIn, mapG, grid, threads, smem, stream)
               : launch_bnp2<T, MODE, NP, false>(p, mapIn, mapG, grid, threads, smem, stream);
}

template <typename T, int MODE>
int corr_bwd_mma16_typed(const void *g, const void *in, void *out, int B, int C, int H, int W, int D, long long gbs, cudaStream_t stream) {
    if (W % 8 != 0 || D > 81 || C < 16) return -1;
    if (((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(in)) & 15) != 0 || (gbs % 8) != 0) return -1;
    if ((reinterpret_cast<uintptr_t>(out) & 3) != 0) return -1;   // packed 4-byte stores
    BwdMma16Params p{};
    p.g = g; p.In = in; p.out = out; p.GBS = gbs;
    p.C = C; p.H = H; p.W = W; p.D = D;
    p.invC = 1.0f / (float)C;
    p.HALO = round_up(D - 1, 8);
    const int np = ceil_div(p.HALO + 16, 16);
    if (np < 1 || np > 6) return -1;
    int warps = ceil_div(W, 16);
    int maxw = env_int_bm16("DSSMD_MMA16_MAXW", 8);
    if (maxw < 1 || maxw > kMaxWarps) maxw = kMaxWarps;
    p.XT = ceil_div(warps, maxw);
    warps = ceil_div(warps, p.XT);
    p.TX = warps * 16;
    p.RW = p.TX + 16 * np - 16;
    p.GWD = MODE == 0 ? p.TX : p.RW;
    auto odd16 = [](int bytes) { int c = (bytes + 15) / 16; if ((c & 1) == 0) ++c; return c * 16; };
    p.RS = odd16(p.RW * 2);
    p.GS = odd16(p.GWD * 2);
    CUtensorMap mapIn{}, mapG{};
    bool tma = env_int_bm16("DSSMD_MMA16_TMA", 1) != 0 && p.RS / 2 <= 256 && p.GS / 2 <= 256;
    if (tma) {
        const CUtensorMapDataType tmt = std::is_same<T, __nv_bfloat16>::value ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
        const cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)B};
        const cuuint64_t strides[3] = {(cuuint64_t)W * 2, (cuuint64_t)H * W * 2, (cuuint64_t)C * H * W * 2};
        const cuuint32_t box[4] = {(cuuint32_t)(p.RS / 2), 1, (cuuint32_t)kCK, 1};
        const cuuint64_t gdims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)B};
        const cuuint64_t gstrides[3] = {(cuuint64_t)W * 2, (cuuint64_t)H * W * 2, (cuuint64_t)gbs * 2};
        const cuuint32_t gbox[4] = {(cuuint32_t)(p.GS / 2), 1, (cuuint32_t)D, 1};
        tma = make_tmap(&mapIn, tmt, in, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE) &&
              make_tmap(&mapG, tmt, g, 4, gdims, gstrides, gbox, CU_TENSOR_MAP_SWIZZLE_NONE);
    }
    const size_t smem = (((size_t)D * p.GS + 127) & ~(size_t)127) + 2 * (size_t)kCK * p.RS;
    if (smem > 200 * 1024) return -1;
    const long long blocks = (long long)B * H * p.XT;
    if (blocks >= (1LL << 31)) return -1;
    if (env_int_bm16("DSSMD_CORR_DEBUG", 0))
        fprintf(stderr, "corr_bwd mma16 cfg: mode=%d NP=%d HALO=%d TX=%d XT=%d RW=%d RS=%d GS=%d warps=%d grid=%lld smem=%zu tma=%d\n", MODE, np, p.HALO,
                p.TX, p.XT, p.RW, p.RS, p.GS, warps, blocks, smem, (int)tma);
    switch (np) {
        case 1: return launch_bnp<T, MODE, 1>(p, tma, mapIn, mapG, (int)blocks, warps * 32, smem, stream);
        case 2: return launch_bnp<T, MODE, 2>(p, tma, mapIn, mapG, (int)blocks, warps * 32, smem, stream);
        case 3: return launch_bnp<T, MODE, 3>(p, tma, mapIn, mapG, (int)blocks, warps * 32, smem, stream);
        case 4: return launch_bnp<T, MODE, 4>(p, tma, mapIn, mapG, (int)blocks, warps * 32, smem, stream);
        case 5: return launch_bnp<T, MODE, 5>(p, tma, mapIn, mapG, (int)blocks, warps * 32, smem, stream);
        default: return launch_bnp<T, MODE, 6>(p, tma, mapIn, mapG, (int)blocks, warps * 32, smem, stream);
    }
}

}  // namespace

// mode 0: out = gL, in = R;  mode 1: out = gR, in = L.  dtype: DSSMD_BF16 / DSSMD_F16.  -1 = not eligible (caller falls back).
int corr_bwd_mma16(const void *g, const void *in, void *out, int mode, int B, int C, int H, int W, int D, long long gbs, int dtype,
                   cudaStream_t stream) {
    if (env_int_bm16("DSSMD_CORR_BWD_MMA16", 1) == 0) return -1;
    if (dtype == DSSMD_BF16)
        return mode == 0 ? corr_bwd_mma16_typed<__nv_bfloat16, 0>(g, in, out, B, C, H, W, D, gbs, stream)
                         : corr_bwd_mma16_typed<__nv_bfloat16, 1>(g, in, out, B, C, H, W, D, gbs, stream);
    if (dtype == DSSMD_F16)
        return mode == 0 ? corr_bwd_mma16_typed<__half, 0>(g, in, out, B, C, H, W, D, gbs, stream)
                         : corr_bwd_mma16_typed<__half, 1>(g, in, out, B, C, H, W, D, gbs, stream);
    return -1;
}

}  // namespace dssmd
