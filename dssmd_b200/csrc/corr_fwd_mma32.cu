// corr_fwd_mma32.cu -- fp32 correlation forward on the warp-level tensor instructions with 3xTF32 operand splitting.
//
//   out[b,d,y,x] = (1/C) * sum_c L[b,c,y,x] * R[b,c,y,x-d]        (upstream: models/UtilsNet/submoudles.py:301-311)
//
// Same banded row GEMM as corr_fwd_mma16.cu -- per warp 16 pixels x (HALO + 16) R positions x C channels, the diagonal band of
// the result is the output -- on mma.sync.m16n8k8.tf32 with every fp32 operand split in registers into hi = tf32(v) and
// lo = tf32(v - hi): S ~= Lhi*Rhi + Llo*Rhi + Lhi*Rlo (fp32 accumulate; the dropped lo*lo term is 2^-22 relative, measured
// error ~1e-6 against the 1e-5 bar -- the scheme corr_fwd_umma.cu uses on tcgen05).  The register-blocked FFMA kernel needs ~1.2
// warp instructions per channel and pixel and is latency-bound at 16 warps per SM; here the fragment loads, the splits and
// the three MMAs of a k8 step come to ~0.55, at 476 tf32 MAC/clk/SM (scripts/micro/mma_rate.cu) against 128 FMA/clk/SM.
//
// Tiles arrive by TMA ([W,H,C,B] maps, 16 channels per stage, two stages): boxes whose width is 8 (mod 32) floats, so that the
// dense box rows put the scalar fragment loads (row = channel tq / tq + 4, column = pixel g / g + 8) on 32 distinct banks; the
// zero fill left of x = 0, past the row end and past the last channel is the TMA unit's = the reference's zero border.
#include "common.cuh"
#include "tma.cuh"

namespace dssmd {

namespace {

// channels per stage: a kernel template parameter CK, 16 or 32 (chosen by corr_fwd_mma32 below)
constexpr int kMaxWarps = 8;     // 128 pixels per block

struct Mma32Params {
    float *out;
    long long OBS;   // elements between two images of out
    int C, H, W, D;
    int HALO;        // D - 1 rounded up to a multiple of 4
    int TX, XT;      // pixels per block tile (multiple of 16), tiles per row
    int LW, RW;      // staged row widths in floats (= box widths, 8 mod 32)
    float invC;
    int CK;          // channels per stage (16 or 32)
    int pf;          // first stage's boxes prefetched towards the L2 before griddepcontrol.wait (common.cuh: prefetch_l2)
};

__device__ __forceinline__ void split_tf32(float v, uint32_t &hi, uint32_t &lo) {
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(v));
    const float r = v - __uint_as_float(hi);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}

__device__ __forceinline__ void mma1688(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// hi = v rounded to tf32 (nearest, ties away: the bit pattern plus half an ulp of tf32, low 13 bits cleared), lo = v - hi (exact).
// cvt.rna.tf32.f32 compiles to these two integer operations PLUS a compare and a select that keep inf / NaN patterns unchanged
// (4 ALU instructions per element; ncu: the ALU pipe was the busiest unit of this kernel at 53 %).  Not needed here: an inf
// keeps its pattern anyway, and a NaN whose payload overflows still gives lo = NaN, which reaches the accumulator through the
// cross terms like it did before.
__device__ __forceinline__ void split_rn(float v, uint32_t &hi, float &lo) {
    hi = (__float_as_uint(v) + 0x1000u) & 0xffffe000u;
    lo = v - __uint_as_float(hi);
}

// two bf16 in one register, `first` in the low half (the lower k index of a fragment pair)
__device__ __forceinline__ uint32_t pack_bf16(float first, float second) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(second), "f"(first));
    return r;
}
__device__ __forceinline__ void mma16816_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// NT = 8-column tiles per warp (8 * NT >= HALO + 16).
// BX: the two CROSS terms on the bf16 instruction.  hi * lo and lo * hi are 2^-11 of the product, so their operands need only 8
// significant bits for the same 2^-20 relative error that dropping lo * lo already allows: bf16(hi) x bf16(lo) + bf16(lo) x bf16(hi) on
// mma.m16n8k16.bf16 -- ONE instruction per 16 channels and term where tf32 needs two k8 steps.  Per 16 channels and 8-column tile
// 2 tf32 + 2 bf16 MMAs instead of 6 tf32 (the kernel is bound by the tensor pipe of this instruction family).  The contraction index
// may be permuted freely as long as A and B agree: bf16 slot pair (2 tq, 2 tq + 1) takes the lane's tf32 elements (tq, tq + 4) of the first
// k8 step, pair (2 tq + 8, 2 tq + 9) those of the second, for both operands.
template <int NT, bool BX, int kCK>
__global__ void __launch_bounds__(kMaxWarps * 32) corr_fwd_mma32_kernel(const Mma32Params p, const __grid_constant__ CUtensorMap mapL,
                                                                        const __grid_constant__ CUtensorMap mapR) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long s_bar[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row = blockIdx.x / p.XT, xt = blockIdx.x - row * p.XT;
    const int b = row / p.H, y = row - b * p.H;
    const int X0 = xt * p.TX;
    const size_t HW = (size_t)p.H * p.W;
    const uint32_t l_bytes = (uint32_t)kCK * (uint32_t)p.LW * 4u, r_bytes = (uint32_t)kCK * (uint32_t)p.RW * 4u;
    const uint32_t stage_bytes = l_bytes + r_bytes;
    const uint32_t sbase = smem_u32(smem);
    const int nck = (p.C + kCK - 1) / kCK;

    auto issue = [&](int c0, int s) {   // one thread
        const uint32_t ls = sbase + (uint32_t)s * stage_bytes, rs = ls + l_bytes;
        const uint32_t bar = smem_u32(&s_bar[s]);
        fence_proxy_async();   // the stage was read through the generic proxy before the block barrier
        mbar_arrive_expect_tx(bar, stage_bytes);
        tma_load_4d(ls, &mapL, X0, y, c0, b, bar);
        tma_load_4d(rs, &mapR, X0 - p.HALO, y, c0, b, bar);
    };
    if (tid == 0) {
        mbar_init(smem_u32(&s_bar[0]), 1);
        mbar_init(smem_u32(&s_bar[1]), 1);
        fence_proxy_async();
    }
    __syncthreads();
    pdl_launch_dependents();
    if (p.pf && tid == (int)blockDim.x - 1) {   // the first stage towards the L2 while the previous kernel drains
        tma_prefetch_desc(&mapL); tma_prefetch_desc(&mapR);
        tma_prefetch_4d(&mapL, X0, y, 0, b);
        tma_prefetch_4d(&mapR, X0 - p.HALO, y, 0, b);
    }
    pdl_wait();            // no load or store above
    if (tid == 0) issue(0, 0);

    float acc[NT][4];
#pragma unroll
    for (int n = 0; n < NT; ++n)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[n][k] = 0.f;

    const int g = lane >> 2, tq = lane & 3;
    const int wx = warp * 16;
    // fragment elements of this lane: A (pixel g / g + 8, channel tq / tq + 4), B (channel tq / tq + 4, column g of each tile)
    const int a_idx = tq * p.LW + wx + g, b_idx = tq * p.RW + wx + g;

    for (int ck = 0; ck < nck; ++ck) {
        const int s = ck & 1;
        if (tid == 0 && ck + 1 < nck) issue((ck + 1) * kCK, s ^ 1);   // stage s^1 was released by the barrier ending iteration ck-1
        mbar_wait(smem_u32(&s_bar[s]), (uint32_t)((ck >> 1) & 1));
        const float *Ls = reinterpret_cast<const float *>(smem + (size_t)s * stage_bytes);
        const float *Rs = reinterpret_cast<const float *>(smem + (size_t)s * stage_bytes + l_bytes);
        if constexpr (BX) {
#pragma unroll
            for (int k16 = 0; k16 < kCK / 16; ++k16) {
                uint32_t ah[2][4], ahb[4], alb[4];
                {
                    float lo[2][4];
#pragma unroll
                    for (int k8 = 0; k8 < 2; ++k8) {
                        const float *la = Ls + (k16 * 16 + k8 * 8) * p.LW + a_idx;
                        split_rn(la[0], ah[k8][0], lo[k8][0]);
                        split_rn(la[8], ah[k8][1], lo[k8][1]);
                        split_rn(la[4 * p.LW], ah[k8][2], lo[k8][2]);
                        split_rn(la[4 * p.LW + 8], ah[k8][3], lo[k8][3]);
                    }
                    // rows g (elements 0, 2) and g + 8 (elements 1, 3); k pairs (tq, tq + 4) of each k8 step
                    ahb[0] = pack_bf16(__uint_as_float(ah[0][0]), __uint_as_float(ah[0][2]));
                    ahb[1] = pack_bf16(__uint_as_float(ah[0][1]), __uint_as_float(ah[0][3]));
                    ahb[2] = pack_bf16(__uint_as_float(ah[1][0]), __uint_as_float(ah[1][2]));
                    ahb[3] = pack_bf16(__uint_as_float(ah[1][1]), __uint_as_float(ah[1][3]));
                    alb[0] = pack_bf16(lo[0][0], lo[0][2]);
                    alb[1] = pack_bf16(lo[0][1], lo[0][3]);
                    alb[2] = pack_bf16(lo[1][0], lo[1][2]);
                    alb[3] = pack_bf16(lo[1][1], lo[1][3]);
                }
                const float *rb = Rs + k16 * 16 * p.RW + b_idx;
#pragma unroll
                for (int n = 0; n < NT; ++n) {
                    uint32_t h00, h01, h10, h11;
                    float l00, l01, l10, l11;
                    split_rn(rb[8 * n], h00, l00);
                    split_rn(rb[4 * p.RW + 8 * n], h01, l01);
                    split_rn(rb[8 * p.RW + 8 * n], h10, l10);
                    split_rn(rb[12 * p.RW + 8 * n], h11, l11);
                    const uint32_t bhb0 = pack_bf16(__uint_as_float(h00), __uint_as_float(h01)), bhb1 = pack_bf16(__uint_as_float(h10), __uint_as_float(h11));
                    const uint32_t blb0 = pack_bf16(l00, l01), blb1 = pack_bf16(l10, l11);
                    mma16816_bf16(acc[n], alb, bhb0, bhb1);   // small terms first
                    mma16816_bf16(acc[n], ahb, blb0, blb1);
                    mma1688(acc[n], ah[0], h00, h01);
                    mma1688(acc[n], ah[1], h10, h11);
                }
            }
        } else {
#pragma unroll
        for (int k8 = 0; k8 < kCK / 8; ++k8) {
            const float *la = Ls + k8 * 8 * p.LW + a_idx;
            uint32_t ah[4], al[4];
            split_tf32(la[0], ah[0], al[0]);
            split_tf32(la[8], ah[1], al[1]);
            split_tf32(la[4 * p.LW], ah[2], al[2]);
            split_tf32(la[4 * p.LW + 8], ah[3], al[3]);
            const float *rb = Rs + k8 * 8 * p.RW + b_idx;
#pragma unroll
            for (int n = 0; n < NT; ++n) {
                uint32_t bh0, bl0, bh1, bl1;
                split_tf32(rb[8 * n], bh0, bl0);
                split_tf32(rb[4 * p.RW + 8 * n], bh1, bl1);
                mma1688(acc[n], al, bh0, bh1);     // small terms first
                mma1688(acc[n], ah, bl0, bl1);
                mma1688(acc[n], ah, bh0, bh1);
            }
        }
        }
        __syncthreads();   // stage s may be refilled by the next iteration's TMA
    }

    // ---- epilogue: band of the accumulator -> per-warp staging tile [D][17] -> 64-byte row segments ------------------------
    float *so = reinterpret_cast<float *>(smem) + (size_t)warp * p.D * 17;
#pragma unroll
    for (int n = 0; n < NT; ++n)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int m = g + ((k & 2) ? 8 : 0);
            const int col = n * 8 + 2 * tq + (k & 1);
            const int d = m + p.HALO - col;
            if (d >= 0 && d < p.D) so[d * 17 + m] = acc[n][k] * p.invC;
        }
    __syncwarp();
    const int x = X0 + wx + (lane & 15);
    if (x < p.W) {
        float *o = p.out + (size_t)b * p.OBS + (size_t)y * p.W + x;
        for (int d = lane >> 4; d < p.D; d += 2) o[(size_t)d * HW] = so[d * 17 + (lane & 15)];
    }
}

int env_int_m32(const char *name, int dflt) {
    if (!tuning_enabled()) return dflt;
    const char *v = std::getenv(name);
    return (v && *v) ? std::atoi(v) : dflt;
}

template <int NT, bool BX, int CK>
int launch_nt_bx(const Mma32Params &p, const CUtensorMap &mapL, const CUtensorMap &mapR, int grid, int threads, size_t smem, cudaStream_t stream) {
    auto kern = corr_fwd_mma32_kernel<NT, BX, CK>;
    {
        const cudaError_t e = kernel_smem_optin(kern, smem);
        if (e != cudaSuccess) return fail(DSSMD_ERR_CUDA, "corr_fwd(mma32): cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(e));
    }
    const cudaError_t le = launch_pdl(kern, dim3((unsigned)grid), dim3((unsigned)threads), smem, stream, p, mapL, mapR);
    if (le != cudaSuccess) return fail(DSSMD_ERR_CUDA, "corr_fwd(mma32): launch failed: %s", cudaGetErrorString(le));
    return check_launch("corr_fwd(mma32)");
}

template <int NT>
int launch_nt(const Mma32Params &p, const CUtensorMap &mapL, const CUtensorMap &mapR, int grid, int threads, size_t smem, cudaStream_t stream) {
    if (env_int_m32("DSSMD_MMA32_BF16X", 1) != 0) {
        if (p.CK == 32) return launch_nt_bx<NT, true, 32>(p, mapL, mapR, grid, threads, smem, stream);
        return launch_nt_bx<NT, true, 16>(p, mapL, mapR, grid, threads, smem, stream);
    }
    return launch_nt_bx<NT, false, 16>(p, mapL, mapR, grid, threads, smem, stream);
}

}  // namespace

// fp32 tensors.  -1 = shape not eligible (W % 4 != 0, D > 81, unaligned, tensor maps unavailable): the caller falls back.
int corr_fwd_mma32(const void *L, const void *R, void *out, int B, int C, int H, int W, int D, long long obs, cudaStream_t stream) {
    if (W % 4 != 0 || D > 81 || C < 8) return -1;
    if (((reinterpret_cast<uintptr_t>(L) | reinterpret_cast<uintptr_t>(R)) & 15) != 0) return -1;
    Mma32Params p{};
    p.out = static_cast<float *>(out); p.OBS = obs;
    p.pf = pdl_prefetch(1);
    p.C = C; p.H = H; p.W = W; p.D = D;
    p.invC = 1.0f / (float)C;
    p.HALO = round_up(D - 1, 4);
    const int nt = ceil_div(p.HALO + 16, 8);
    if (nt < 2 || nt > 12) return -1;
    int warps = ceil_div(W, 16);
    int maxw = env_int_m32("DSSMD_MMA32_MAXW", kMaxWarps);
    if (maxw < 1 || maxw > kMaxWarps) maxw = kMaxWarps;
    p.XT = ceil_div(warps, maxw);
    warps = ceil_div(warps, p.XT);
    p.TX = warps * 16;
    auto w8mod32 = [](int w) { int r = ((w + 23) / 32) * 32 + 8; return r; };   // smallest value >= w that is 8 (mod 32)
    p.LW = w8mod32(p.TX);
    p.RW = w8mod32(p.TX + 8 * nt - 16);
    if (p.LW > 256 || p.RW > 256) return -1;
    // channels per stage: 32 (half the TMA operations and block barriers) while two such stages leave room for four blocks per SM,
    // else 16.  Measured, 16 / 32: cfg5 shape 12.4 / 10.8 us, KITTI x16 29.6 / 29.0, 320x640 12.1 / 11.7, cfg2 (61 KB of stages with 32)
    // 54.6 / 57.8.  DSSMD_MMA32_CK forces one.
    int kCK = env_int_m32("DSSMD_MMA32_CK", 0);
    if (kCK != 16 && kCK != 32) {
        const size_t smem32 = 2 * (size_t)32 * (size_t)(p.LW + p.RW) * 4;
        const long long slots = smem32 <= 200 * 1024 ? (long long)sm_count() * (long long)((220 * 1024) / smem32) : 0;
        // ... or while every block of the problem is resident anyway (the 576x960 model shape: 288 blocks of 78 KB)
        kCK = (C >= 32 && (smem32 <= 56 * 1024 || (long long)B * H * p.XT <= slots)) ? 32 : 16;
    }
    if (env_int_m32("DSSMD_MMA32_BF16X", 1) == 0) kCK = 16;
    p.CK = kCK;
    size_t smem = 2 * (size_t)kCK * (size_t)(p.LW + p.RW) * 4;
    const size_t epi = (size_t)warps * D * 17 * sizeof(float);
    if (epi > smem) smem = epi;
    if (smem > 200 * 1024) return -1;
    const long long blocks = (long long)B * H * p.XT;
    if (blocks >= (1LL << 31)) return -1;
    CUtensorMap mapL, mapR;
    const cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)B};
    const cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)C * H * W * 4};
    const cuuint32_t boxl[4] = {(cuuint32_t)p.LW, 1, (cuuint32_t)kCK, 1}, boxr[4] = {(cuuint32_t)p.RW, 1, (cuuint32_t)kCK, 1};
    if (!make_tmap_f32(&mapL, L, 4, dims, strides, boxl, CU_TENSOR_MAP_SWIZZLE_NONE) ||
        !make_tmap_f32(&mapR, R, 4, dims, strides, boxr, CU_TENSOR_MAP_SWIZZLE_NONE)) return -1;
    if (env_int_m32("DSSMD_CORR_DEBUG", 0))
        fprintf(stderr, "corr_fwd mma32 cfg: NT=%d HALO=%d TX=%d XT=%d LW=%d RW=%d warps=%d grid=%lld smem=%zu\n", nt, p.HALO, p.TX, p.XT, p.LW, p.RW,
                warps, blocks, smem);
    switch (nt) {
        case 2: return launch_nt<2>(p, mapL, mapR, (int)blocks, warps * 32, smem, stream);
        case 3: return launch_nt<3>(p, mapL, mapR, (int)blocks, warps * 32, smem, stream);
        case 4: return launch_nt<4>(p, mapL, mapR, (int)blocks, warps * 32, smem, stream);
        case 5: return launch_nt<5>(p, mapL, mapR, (int)blocks, warps * 32, smem, stream);
        case 6: return launch_nt<6>(p, mapL, mapR, (int)blocks, warps * 32, smem, stream);
        case 7: return launch_nt<7>(p, mapL, mapR, (int)blocks, warps * 32, smem, stream);
        case 8: return launch_nt<8>(p, mapL, mapR, (int)blocks, warps * 32, smem, stream);
        case 9: return launch_nt<9>(p, mapL, mapR, (int)blocks, warps * 32, smem, stream);
        case 10: return launch_nt<10>(p, mapL, mapR, (int)blocks, warps * 32, smem, stream);
        case 11: return launch_nt<11>(p, mapL, mapR, (int)blocks, warps * 32, smem, stream);
        default: return launch_nt<12>(p, mapL, mapR, (int)blocks, warps * 32, smem, stream);
    }
}

}  // namespace dssmd
