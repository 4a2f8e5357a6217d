// corr_bwd_mma32.cu -- fp32 backward of the correlation volume on the warp-level tensor instructions with 3xTF32 splitting.
//
//   gL[c,x ] = (1/C) * sum_d g[d,x]    * R[c,x-d]        (MODE 0)
//   gR[c,x'] = (1/C) * sum_d g[d,x'+d] * L[c,x'+d]       (MODE 1)       (the graph autograd builds from submoudles.py:304-308)
//
// The formulation of corr_bwd_mma16.cu (a banded operand G built once per warp from the staged g tile, the product issued
// transposed: channels as M, the 16 pixels of the warp as two N tiles) on mma.sync.m16n8k8.tf32, with both operands split in
// registers into hi = tf32(v) and lo = tf32(v - hi) and three products per step (lo*hi, hi*lo, hi*hi; fp32 accumulate):
//   MODE 0:  gL[c][x0+m] = sum_k R[c][x0 - HALO + k] * G[m][k],   G[m][k] = g[m + HALO - k][x0 + m]    (0 outside 0 <= d < D)
//   MODE 1:  gR[c][x0+m] = sum_k L[c][x0 + k]        * G[m][k],   G[m][k] = g[k - m][x0 + k]
// HALO = D - 1 rounded up to 4, K = HALO + 16 rounded up to 8.  The split G fragments (2 x 4 x K/8 registers) stay in
// registers for the whole block, like the g slab of corr_bwd_reg.cu.  Operand rows and the g tile arrive by TMA; the operand
// box is 4 (mod 32) floats wide so that the scalar A-fragment loads (row = channel g / g + 8, column = position tq / tq + 4) hit
// 32 distinct banks.  An accumulator pair is two neighbouring pixels of one channel: 8-byte stores, 32 contiguous bytes per
// four lanes.
#include "common.cuh"
#include "tma.cuh"

namespace dssmd {

namespace {

constexpr int kCK = 32;          // channels per stage (two 16-channel M tiles); 16 / 64 measured slower with the one-launch kernel
                                 // (cfg2 112 / 133 us against 105.6, 320x640 18.5 / 19.6 against 16.2: profiles/r04as_bwd_mma32_ck.log)
constexpr int kMaxWarps = 8;     // 128 pixels per block

struct BwdMma32Params {
    float *out;
    int C, H, W, D;
    int HALO;        // D - 1 rounded up to a multiple of 4
    int TX, XT;      // pixels per block tile (multiple of 16), tiles per row
    int RW;          // staged operand row width in floats (= box width, 4 mod 32)
    int GW;          // staged g row width in floats (= box width, multiple of 4)
    float invC;
    int pf;          // first stage's boxes prefetched towards the L2 before griddepcontrol.wait (common.cuh: prefetch_l2)
};

static inline bool bwd_mma32_bf16x() {   // DSSMD_MMA32_BF16X=0 (tuning): all three products on tf32
    if (!tuning_enabled()) return true;
    const char *v = std::getenv("DSSMD_MMA32_BF16X");
    return !(v && v[0] == '0');
}

__device__ __forceinline__ void split_tf32b(float v, uint32_t &hi, uint32_t &lo) {
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(v));
    const float r = v - __uint_as_float(hi);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}

__device__ __forceinline__ void mma1688b(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// hi = v rounded to tf32 (nearest, ties away), lo = v - hi: the two integer operations of cvt.rna.tf32.f32 without its compare and
// select for inf / NaN (see corr_fwd_mma32.cu: a NaN still reaches the accumulator through lo)
__device__ __forceinline__ void split_rnb(float v, uint32_t &hi, float &lo) {
    hi = (__float_as_uint(v) + 0x1000u) & 0xffffe000u;
    lo = v - __uint_as_float(hi);
}

// two bf16 in one register, `first` in the low half (the lower k index of a fragment pair)
__device__ __forceinline__ uint32_t pack_bf16b(float first, float second) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(second), "f"(first));
    return r;
}
__device__ __forceinline__ void mma16816_bf16b(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// NK = k8 steps (8 * NK >= HALO + 16).
// BX: the two cross terms (hi x lo, lo x hi) of a PAIR of k8 steps on one bf16 m16n8k16 MMA each, as in corr_fwd_mma32.cu (the
// operands of a term that is 2^-11 of the product need 8 significant bits; the contraction index -- here the band position --
// is permuted the same way for A and G^T): 2 tf32 + 2 bf16 MMAs per pair and pixel tile instead of 6 tf32; an odd last step
// keeps its three tf32 products.
// TB: the most threads a block may have.  Blocks of up to five warps (rows tiled in 80 pixels: cfg2, KITTI, the 320x640 crops) ask for
// four resident blocks (96 registers instead of ~125 and three blocks: cfg2 backward 116.0 -> 114.2 us, KITTI 60.5 -> 59.1 us)
template <int MODE, int NK, bool BX>
__device__ __forceinline__ void corr_bwd_mma32_body(const BwdMma32Params &p, const CUtensorMap &mapIn, const CUtensorMap &mapG) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long s_bar[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row = blockIdx.x / p.XT, xt = blockIdx.x - row * p.XT;
    const int b = row / p.H, y = row - b * p.H;
    const int X0 = xt * p.TX;
    const size_t HW = (size_t)p.H * p.W;
    const uint32_t g_bytes = (uint32_t)p.D * (uint32_t)p.GW * 4u;
    const uint32_t g_pad = (g_bytes + 127u) & ~127u;
    const uint32_t stage_bytes = (uint32_t)kCK * (uint32_t)p.RW * 4u;
    const uint32_t sbase = smem_u32(smem);                 // g tile | stage 0 | stage 1
    const int xin0 = MODE == 0 ? X0 - p.HALO : X0;        // position of operand column 0
    const int nck = (p.C + kCK - 1) / kCK;

    auto issue = [&](int c0, int s, bool with_g) {   // one thread
        const uint32_t rs = sbase + g_pad + (uint32_t)s * stage_bytes;
        const uint32_t bar = smem_u32(&s_bar[s]);
        fence_proxy_async();
        mbar_arrive_expect_tx(bar, stage_bytes + (with_g ? g_bytes : 0u));
        tma_load_4d(rs, &mapIn, xin0, y, c0, b, bar);
        if (with_g) tma_load_4d(sbase, &mapG, X0, y, 0, b, bar);
    };
    if (tid == 0) {
        mbar_init(smem_u32(&s_bar[0]), 1);
        mbar_init(smem_u32(&s_bar[1]), 1);
        fence_proxy_async();
    }
    __syncthreads();
    pdl_launch_dependents();
    if (p.pf && tid == (int)blockDim.x - 1) {   // the first stage towards the L2 while the previous kernel drains
        tma_prefetch_desc(&mapIn); tma_prefetch_desc(&mapG);
        tma_prefetch_4d(&mapIn, xin0, y, 0, b);
        tma_prefetch_4d(&mapG, X0, y, 0, b);
    }
    pdl_wait();            // no load or store above
    if (tid == 0) issue(0, 0, true);

    const int g8 = lane >> 2, tq = lane & 3;
    const int wx = warp * 16;
    uint32_t gh[NK][2][2], gl[NK][2][2];     // G^T "col" fragments (k = tq / tq + 4, pixel 8 t + g8) of every k8 step, hi and lo
    constexpr int NP = BX ? NK / 2 : 0;      // pairs of k8 steps whose cross terms go through bf16
    uint32_t ghb[NP > 0 ? NP : 1][2][2], glb[NP > 0 ? NP : 1][2][2];   // ... packed: [pair][pixel tile][step of the pair]
    const int a_idx = g8 * p.RW + wx + tq;   // A fragment: (channel g8 / g8 + 8, position tq / tq + 4) of each k8 step
    float *const ob = p.out + (size_t)b * p.C * HW + (size_t)y * p.W + X0 + wx + 2 * tq;
    const bool st0 = X0 + wx + 2 * tq < p.W, st1 = X0 + wx + 8 + 2 * tq < p.W;   // W % 4 == 0 and even offsets: a pixel pair is in or out together

    for (int ck = 0; ck < nck; ++ck) {
        const int s = ck & 1;
        if (tid == 0 && ck + 1 < nck) issue((ck + 1) * kCK, s ^ 1, false);   // stage s^1 was released by the barrier ending iteration ck-1
        mbar_wait(smem_u32(&s_bar[s]), (uint32_t)((ck >> 1) & 1));
        if (ck == 0) {
            const float *gt = reinterpret_cast<const float *>(smem);
#pragma unroll
            for (int q = 0; q < NK; ++q)
#pragma unroll
                for (int t = 0; t < 2; ++t)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int m = 8 * t + g8, k = 8 * q + tq + 4 * e;
                        const int d = MODE == 0 ? m + p.HALO - k : k - m;
                        const int xc = MODE == 0 ? wx + m : wx + k;
                        const float v = (d >= 0 && d < p.D) ? gt[d * p.GW + xc] : 0.f;
                        split_tf32b(v, gh[q][t][e], gl[q][t][e]);
                    }
            if constexpr (BX) {
#pragma unroll
                for (int pq = 0; pq < NP; ++pq)
#pragma unroll
                    for (int t = 0; t < 2; ++t)
#pragma unroll
                        for (int st = 0; st < 2; ++st) {
                            const int q = 2 * pq + st;
                            ghb[pq][t][st] = pack_bf16b(__uint_as_float(gh[q][t][0]), __uint_as_float(gh[q][t][1]));
                            glb[pq][t][st] = pack_bf16b(__uint_as_float(gl[q][t][0]), __uint_as_float(gl[q][t][1]));
                        }
            }
        }
        const float *Rs = reinterpret_cast<const float *>(smem + g_pad + (size_t)s * stage_bytes);
        const int c0 = ck * kCK;
#pragma unroll
        for (int mt = 0; mt < kCK / 16; ++mt) {
            float acc[2][4];
#pragma unroll
            for (int t = 0; t < 2; ++t)
#pragma unroll
                for (int k = 0; k < 4; ++k) acc[t][k] = 0.f;
            const float *ra = Rs + mt * 16 * p.RW + a_idx;
            if constexpr (BX) {
#pragma unroll
                for (int pq = 0; pq < NP; ++pq) {
                    uint32_t ah[2][4], ahb[4], alb[4];
                    float lo[2][4];
#pragma unroll
                    for (int st = 0; st < 2; ++st) {
                        const int q = 2 * pq + st;
                        const float av[4] = {ra[8 * q], ra[8 * p.RW + 8 * q], ra[8 * q + 4], ra[8 * p.RW + 8 * q + 4]};
#pragma unroll
                        for (int i = 0; i < 4; ++i) split_rnb(av[i], ah[st][i], lo[st][i]);
                    }
                    ahb[0] = pack_bf16b(__uint_as_float(ah[0][0]), __uint_as_float(ah[0][2]));
                    ahb[1] = pack_bf16b(__uint_as_float(ah[0][1]), __uint_as_float(ah[0][3]));
                    ahb[2] = pack_bf16b(__uint_as_float(ah[1][0]), __uint_as_float(ah[1][2]));
                    ahb[3] = pack_bf16b(__uint_as_float(ah[1][1]), __uint_as_float(ah[1][3]));
                    alb[0] = pack_bf16b(lo[0][0], lo[0][2]);
                    alb[1] = pack_bf16b(lo[0][1], lo[0][3]);
                    alb[2] = pack_bf16b(lo[1][0], lo[1][2]);
                    alb[3] = pack_bf16b(lo[1][1], lo[1][3]);
#pragma unroll
                    for (int t = 0; t < 2; ++t) {
                        mma16816_bf16b(acc[t], alb, ghb[pq][t][0], ghb[pq][t][1]);     // small terms first
                        mma16816_bf16b(acc[t], ahb, glb[pq][t][0], glb[pq][t][1]);
                        mma1688b(acc[t], ah[0], gh[2 * pq][t][0], gh[2 * pq][t][1]);
                        mma1688b(acc[t], ah[1], gh[2 * pq + 1][t][0], gh[2 * pq + 1][t][1]);
                    }
                }
            }
#pragma unroll
            for (int q = BX ? 2 * NP : 0; q < NK; ++q) {
                uint32_t ah[4], al[4];
                if constexpr (BX) {   // the cheap split; lo truncated to tf32 (2^-21 of the product)
                    const float av[4] = {ra[8 * q], ra[8 * p.RW + 8 * q], ra[8 * q + 4], ra[8 * p.RW + 8 * q + 4]};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float lo;
                        split_rnb(av[i], ah[i], lo);
                        al[i] = __float_as_uint(lo) & 0xffffe000u;
                    }
                } else {
                    split_tf32b(ra[8 * q], ah[0], al[0]);
                    split_tf32b(ra[8 * p.RW + 8 * q], ah[1], al[1]);
                    split_tf32b(ra[8 * q + 4], ah[2], al[2]);
                    split_tf32b(ra[8 * p.RW + 8 * q + 4], ah[3], al[3]);
                }
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    mma1688b(acc[t], al, gh[q][t][0], gh[q][t][1]);     // small terms first
                    mma1688b(acc[t], ah, gl[q][t][0], gl[q][t][1]);
                    mma1688b(acc[t], ah, gh[q][t][0], gh[q][t][1]);
                }
            }
            // acc[t]: c0/c1 = (channel g8, pixels 8 t + 2 tq, + 1), c2/c3 = (channel g8 + 8, same pixels)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int c = c0 + mt * 16 + hh * 8 + g8;
                if (c < p.C) {
                    float *o = ob + (size_t)c * HW;
                    if (st0) *reinterpret_cast<float2 *>(o) = make_float2(acc[0][2 * hh] * p.invC, acc[0][2 * hh + 1] * p.invC);
                    if (st1) *reinterpret_cast<float2 *>(o + 8) = make_float2(acc[1][2 * hh] * p.invC, acc[1][2 * hh + 1] * p.invC);
                }
            }
        }
        __syncthreads();   // stage s may be refilled by the next iteration's TMA
    }
}

template <int MODE, int NK, bool BX, int TB>
__global__ void __launch_bounds__(TB, TB <= 160 ? 4 : 2) corr_bwd_mma32_kernel(const BwdMma32Params p, const __grid_constant__ CUtensorMap mapIn,
                                                                        const __grid_constant__ CUtensorMap mapG) {
    corr_bwd_mma32_body<MODE, NK, BX>(p, mapIn, mapG);
}

// both gradients in ONE launch: blockIdx.y = 0 computes gL (operand R), 1 computes gR (operand L); the two halves of the grid share the
// launch ramp and drain, and the second gradient's blocks fill the SMs the first one's last wave leaves idle
template <int NK, int TB>
__global__ void __launch_bounds__(TB, TB <= 160 ? 4 : 2) corr_bwd_mma32_both_kernel(const BwdMma32Params p0, const BwdMma32Params p1,
                                                                             const __grid_constant__ CUtensorMap mapIn0, const __grid_constant__ CUtensorMap mapG0,
                                                                             const __grid_constant__ CUtensorMap mapIn1, const __grid_constant__ CUtensorMap mapG1) {
    if (blockIdx.y == 0) corr_bwd_mma32_body<0, NK, true>(p0, mapIn0, mapG0);
    else corr_bwd_mma32_body<1, NK, true>(p1, mapIn1, mapG1);
}

int env_int_bm32(const char *name, int dflt) {
    if (!tuning_enabled()) return dflt;
    const char *v = std::getenv(name);
    return (v && *v) ? std::atoi(v) : dflt;
}

template <int MODE, int NK, bool BX, int TB>
int launch_bnk_tb(const BwdMma32Params &p, const CUtensorMap &mapIn, const CUtensorMap &mapG, int grid, int threads, size_t smem, cudaStream_t stream) {
    auto kern = corr_bwd_mma32_kernel<MODE, NK, BX, TB>;
    {
        const cudaError_t e = kernel_smem_optin(kern, smem);
        if (e != cudaSuccess) return fail(DSSMD_ERR_CUDA, "corr_bwd(mma32): cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(e));
    }
    const cudaError_t le = launch_pdl(kern, dim3((unsigned)grid), dim3((unsigned)threads), smem, stream, p, mapIn, mapG);
    if (le != cudaSuccess) return fail(DSSMD_ERR_CUDA, "corr_bwd(mma32): launch failed: %s", cudaGetErrorString(le));
    return check_launch(MODE == 0 ? "corr_bwd(gL, mma32)" : "corr_bwd(gR, mma32)");
}

template <int MODE, int NK>
int launch_bnk(const BwdMma32Params &p, const CUtensorMap &mapIn, const CUtensorMap &mapG, int grid, int threads, size_t smem, cudaStream_t stream) {
    if (bwd_mma32_bf16x()) {
        if constexpr (NK <= 8) {   // beyond that the 96-register build spills
            if (threads <= 160) return launch_bnk_tb<MODE, NK, true, 160>(p, mapIn, mapG, grid, threads, smem, stream);
        }
        return launch_bnk_tb<MODE, NK, true, kMaxWarps * 32>(p, mapIn, mapG, grid, threads, smem, stream);
    }
    return launch_bnk_tb<MODE, NK, false, kMaxWarps * 32>(p, mapIn, mapG, grid, threads, smem, stream);
}

struct BwdPlan {
    BwdMma32Params p;
    CUtensorMap mapIn, mapG;
    int nk, warps;
    long long blocks;
    size_t smem;
};

// parameters, tensor maps and launch geometry of one gradient; false = shape not eligible
template <int MODE>
bool plan_bwd_mma32(BwdPlan &pl, const void *g, const void *in, void *out, int B, int C, int H, int W, int D, long long gbs) {
    if (W % 4 != 0 || D > 81 || C < 8) return false;
    if (((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(in)) & 15) != 0 || (gbs % 4) != 0) return false;
    if ((reinterpret_cast<uintptr_t>(out) & 7) != 0) return false;   // 8-byte stores
    BwdMma32Params &p = pl.p;
    p = BwdMma32Params{};
    p.out = static_cast<float *>(out);
    p.pf = pdl_prefetch(1);
    p.C = C; p.H = H; p.W = W; p.D = D;
    p.invC = 1.0f / (float)C;
    p.HALO = round_up(D - 1, 4);
    const int nk = ceil_div(p.HALO + 16, 8);
    if (nk < 2 || nk > 12) return false;
    int warps = ceil_div(W, 16);
    int maxw = env_int_bm32("DSSMD_MMA32_MAXW", kMaxWarps);
    if (maxw < 1 || maxw > kMaxWarps) maxw = kMaxWarps;
    p.XT = ceil_div(warps, maxw);
    warps = ceil_div(warps, p.XT);
    p.TX = warps * 16;
    const int need = p.TX + 8 * nk - 16;                  // operand positions a block touches
    p.RW = ((need + 27) / 32) * 32 + 4;                   // smallest value >= need that is 4 (mod 32)
    p.GW = MODE == 0 ? p.TX : round_up(need, 4);
    if (p.RW > 256 || p.GW > 256) return false;
    pl.smem = (((size_t)D * p.GW * 4 + 127) & ~(size_t)127) + 2 * (size_t)kCK * p.RW * 4;
    if (pl.smem > 200 * 1024) return false;
    pl.blocks = (long long)B * H * p.XT;
    if (pl.blocks >= (1LL << 31)) return false;
    const cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)B};
    const cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)C * H * W * 4};
    const cuuint32_t box[4] = {(cuuint32_t)p.RW, 1, (cuuint32_t)kCK, 1};
    const cuuint64_t gdims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)B};
    const cuuint64_t gstrides[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)gbs * 4};
    const cuuint32_t gbox[4] = {(cuuint32_t)p.GW, 1, (cuuint32_t)D, 1};
    if (!make_tmap_f32(&pl.mapIn, in, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE) ||
        !make_tmap_f32(&pl.mapG, g, 4, gdims, gstrides, gbox, CU_TENSOR_MAP_SWIZZLE_NONE)) return false;
    pl.nk = nk; pl.warps = warps;
    return true;
}

void debug_plan(const BwdPlan &pl, int mode) {
    if (env_int_bm32("DSSMD_CORR_DEBUG", 0))
        fprintf(stderr, "corr_bwd mma32 cfg: mode=%d NK=%d HALO=%d TX=%d XT=%d RW=%d GW=%d warps=%d grid=%lld smem=%zu\n", mode, pl.nk, pl.p.HALO, pl.p.TX,
                pl.p.XT, pl.p.RW, pl.p.GW, pl.warps, pl.blocks, pl.smem);
}

template <int MODE>
int corr_bwd_mma32_mode(const void *g, const void *in, void *out, int B, int C, int H, int W, int D, long long gbs, cudaStream_t stream) {
    BwdPlan pl;
    if (!plan_bwd_mma32<MODE>(pl, g, in, out, B, C, H, W, D, gbs)) return -1;
    debug_plan(pl, MODE);
#define DSSMD_BNK_CASE(n) case n: return launch_bnk<MODE, n>(pl.p, pl.mapIn, pl.mapG, (int)pl.blocks, pl.warps * 32, pl.smem, stream);
    switch (pl.nk) {
        DSSMD_BNK_CASE(2) DSSMD_BNK_CASE(3) DSSMD_BNK_CASE(4) DSSMD_BNK_CASE(5) DSSMD_BNK_CASE(6) DSSMD_BNK_CASE(7)
        DSSMD_BNK_CASE(8) DSSMD_BNK_CASE(9) DSSMD_BNK_CASE(10) DSSMD_BNK_CASE(11)
        default: return launch_bnk<MODE, 12>(pl.p, pl.mapIn, pl.mapG, (int)pl.blocks, pl.warps * 32, pl.smem, stream);
    }
#undef DSSMD_BNK_CASE
}

template <int NK, int TB>
int launch_both(const BwdPlan &a, const BwdPlan &b, cudaStream_t stream) {
    auto kern = corr_bwd_mma32_both_kernel<NK, TB>;
    const size_t smem = a.smem > b.smem ? a.smem : b.smem;
    {
        const cudaError_t e = kernel_smem_optin(kern, smem);
        if (e != cudaSuccess) return fail(DSSMD_ERR_CUDA, "corr_bwd(mma32, both): cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(e));
    }
    const cudaError_t le = launch_pdl(kern, dim3((unsigned)a.blocks, 2u), dim3((unsigned)(a.warps * 32)), smem, stream, a.p, b.p, a.mapIn, a.mapG, b.mapIn, b.mapG);
    if (le != cudaSuccess) return fail(DSSMD_ERR_CUDA, "corr_bwd(mma32, both): launch failed: %s", cudaGetErrorString(le));
    return check_launch("corr_bwd(gL + gR, mma32)");
}

}  // namespace

// Both gradients in one launch (grid.y = gradient).  Blocks of up to five warps and NK <= 8 (the 96-register build): the shapes the
// per-gradient kernels are taken for by default.  -1 = not eligible (the caller issues one launch per gradient).
int corr_bwd_mma32_both(const void *g, const void *L, const void *R, void *gL, void *gR, int B, int C, int H, int W, int D, long long gbs, cudaStream_t stream) {
    if (!bwd_mma32_bf16x() || env_int_bm32("DSSMD_MMA32_BOTH", 1) == 0) return -1;
    BwdPlan a, b;
    if (!plan_bwd_mma32<0>(a, g, R, gL, B, C, H, W, D, gbs) || !plan_bwd_mma32<1>(b, g, L, gR, B, C, H, W, D, gbs)) return -1;
    if (a.nk != b.nk || a.warps != b.warps || a.blocks != b.blocks || a.nk > 8) return -1;
    debug_plan(a, 0); debug_plan(b, 1);
    if (env_int_bm32("DSSMD_CORR_DEBUG", 0)) fprintf(stderr, "corr_bwd mma32 both cfg: one launch, grid=(%lld,2)\n", a.blocks);
#define DSSMD_BOTH_CASE(n) case n: return a.warps * 32 <= 160 ? launch_both<n, 160>(a, b, stream) : launch_both<n, kMaxWarps * 32>(a, b, stream);
    switch (a.nk) {
        DSSMD_BOTH_CASE(2) DSSMD_BOTH_CASE(3) DSSMD_BOTH_CASE(4) DSSMD_BOTH_CASE(5) DSSMD_BOTH_CASE(6) DSSMD_BOTH_CASE(7)
        default: return a.warps * 32 <= 160 ? launch_both<8, 160>(a, b, stream) : launch_both<8, kMaxWarps * 32>(a, b, stream);
    }
#undef DSSMD_BOTH_CASE
}

// fp32 tensors.  mode 0: out = gL, in = R;  mode 1: out = gR, in = L.  -1 = not eligible (caller falls back).
int corr_bwd_mma32(const void *g, const void *in, void *out, int mode, int B, int C, int H, int W, int D, long long gbs, cudaStream_t stream) {
    return mode == 0 ? corr_bwd_mma32_mode<0>(g, in, out, B, C, H, W, D, gbs, stream) : corr_bwd_mma32_mode<1>(g, in, out, B, C, H, W, D, gbs, stream);
}

}  // namespace dssmd
