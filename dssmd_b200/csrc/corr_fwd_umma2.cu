// corr_fwd_umma2.cu -- correlation forward on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM), fp32 in /
// fp32 out with 3xTF32 splitting; second design.  Replaces build_corr, upstream models/UtilsNet/submoudles.py:301-311.
//
// The first tcgen05 kernel (corr_fwd_umma.cu) lost to the mma.sync kernel at the shapes it was built for (C256 80x160 D41:
// 85 vs 74 us).  ncu of both (profiles/r02p_tcgen05_vs_mma32.md): its tensor pipe is busy 41 % of the time and the split warps
// spend 18 % of their samples waiting for TMA data -- the operand ring is too shallow for the latency it has to cover (4 stages,
// each holding the raw tile AND its split copy, 2 KB boxes issued by one thread), and tiles of 128 FLATTENED pixels touch two
// image rows, so the MMA's N is always the 256-column maximum.  Here:
//   * a tile is 128 pixels of ONE image row: N = 128 + D + alignment, rounded to 16 (176 instead of 256 at D = 41; the row's tail
//     tile of 32 pixels needs N = 80), so less tensor time and fewer operand bytes per pixel;
//   * truncation split: the tensor core reads the top 19 bits of an fp32 word, so the RAW tile as TMA delivered it already is
//     `hi`; only lo = v - trunc(v) is written, into a SEPARATE, shallow ring (3 slots).  The raw ring holds nothing else and is
//     as deep as shared memory allows (6-8 stages of 16 channels): TMA runs far ahead of the split and the MMAs;
//   * the TMA unit needs ~150 cycles per tensor-box operation whatever its size (measured: both tcgen05 kernels with 2 KB boxes ran
//     at one box per 140-160 cycles per SM = 14 B/clk, 30 % of the HBM rate, with one or with two producer threads), so a box now
//     carries GC = 64 channels (8 KB): the raw ring is organised per box position -- [group][box][GC rows of 128 B] -- and one
//     operation fills the rows of four 16-channel sub-stages; the MMA descriptors step through a box's rows (LBO = GC x 128 B);
//   * two producer threads (A boxes / B boxes);
//   * per K step of 8 channels three MMAs, hi x lo + lo x hi + hi x hi (the dropped lo x lo is 2^-20 relative), fp32 accumulate
//     in TMEM, two accumulators so the epilogue of tile t (tcgen05.ld of the band window, barrel shift to align the diagonal,
//     x 1/C, coalesced stores) overlaps the MMAs of tile t + 1.
// One persistent, warp-specialised CTA per SM (480 threads): warps 0-3 epilogue (the four TMEM lane quarters), 4-11 split,
// 12 / 13 TMA producers, 14 MMA issue.
#include "common.cuh"
#include "tma.cuh"

#include <cuda.h>

#include <cstdlib>

namespace dssmd {

namespace {

constexpr int kTileM = 128;
constexpr int kKC = 16;            // channels per stage (= 2 MMA K steps of 8)
constexpr int kBoxX = 32;          // pixels per TMA box = one 128-byte swizzle row
constexpr int kBoxBytes = kBoxX * kKC * 4;   // 2048
constexpr int kABoxes = kTileM / kBoxX;      // 4
constexpr int kLoStages = 3;
constexpr int kMaxRawStages = 8;   // raw groups
constexpr int kEpiWarps = 4, kXfWarps = 8;
constexpr int kXfThreads = kXfWarps * 32;
constexpr int kLoadWarpA = kEpiWarps + kXfWarps;   // 12
constexpr int kLoadWarpB = kLoadWarpA + 1;         // 13
constexpr int kMmaWarp = kLoadWarpB + 1;           // 14
constexpr int kThreads = (kMmaWarp + 1) * 32;      // 480

struct Umma2Params {
    float *out;
    long long OBS;  // elements between two images of `out`
    int B, C, H, W, D;
    int XT;       // tiles per row
    int ntiles;   // B * H * XT
    int HALO;     // D - 1
    int NB;       // B boxes reserved per stage (max over the tiles of a row)
    int n32;      // round the MMA's N to whole 32-column boxes instead of 16 (development switch)
    int GC;       // channels per TMA box / raw group (16, 32 or 64)
    int NSUB;     // 16-channel sub-stages per group
    int NG;       // groups per tile = C / GC
    int RG;       // raw groups in the ring
    int c_pow2;
    int dbg;      // development: 1 no epilogue work, 2 no split work, 4 no MMAs, 8 no TMA (timings of the remaining stages; results wrong)
};

struct Tile {
    int b, y, X0;
    int len;     // valid pixels
    int xp0;     // x' of B column 0 (multiple of 4, may be negative)
    int nbA, nbB;
    int N;       // MMA N (multiple of 16)
};

__device__ __forceinline__ Tile tile_of(const Umma2Params &p, int t) {
    Tile tl;
    const int xt = t % p.XT, row = t / p.XT;
    tl.b = row / p.H; tl.y = row - tl.b * p.H;
    tl.X0 = xt * kTileM;
    tl.len = p.W - tl.X0 < kTileM ? p.W - tl.X0 : kTileM;
    tl.xp0 = (tl.X0 - p.HALO) & ~3;                       // TMA box starts must be 16-byte aligned
    const int ncol = tl.X0 + tl.len - tl.xp0;
    tl.nbA = (tl.len + kBoxX - 1) / kBoxX;
    tl.nbB = (ncol + kBoxX - 1) / kBoxX;
    tl.N = p.n32 ? tl.nbB * kBoxX : ((ncol + 15) & ~15);
    return tl;
}

// barriers (shared): indices into one array
enum {
    kRawFull = 0,                              // [kMaxRawStages]  TMA complete_tx (two producers)  -> SPLIT
    kRawEmpty = kRawFull + kMaxRawStages,      // [kMaxRawStages]  MMA (tcgen05.commit)             -> TMA producers
    kLoFull = kRawEmpty + kMaxRawStages,       // [kLoStages]      SPLIT                            -> MMA
    kLoEmpty = kLoFull + kLoStages,            // [kLoStages]      MMA (tcgen05.commit)             -> SPLIT
    kAccFull = kLoEmpty + kLoStages,           // [2]              MMA (tcgen05.commit)             -> EPILOGUE
    kAccEmpty = kAccFull + 2,                  // [2]              EPILOGUE                         -> MMA
    kNumBars = kAccEmpty + 2
};

__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
    // MN-major tf32 operands only exist in the SWIZZLE_128B_BASE32B layout (32-byte swizzle atoms, 4-row
    // K groups): start>>4 [0,14) | LBO>>4 [16,30) = stride between 32-pixel boxes | SBO>>4 [32,46) =
    // stride between groups of 4 channel rows | version=1 [46,48) | layout=1 [61,64)
    const uint64_t lbo = lbo_bytes >> 4, sbo = 512 >> 4;
    return (uint64_t)((smem_addr >> 4) & 0x3fff) | (lbo << 16) | (sbo << 32) | (1ull << 46) | (1ull << 61);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
        : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float *v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}


template <int NW16>  // epilogue window = 16 * NW16 accumulator columns per thread
__global__ void __launch_bounds__(kThreads, 1) corr_fwd_umma2_kernel(const Umma2Params p, const __grid_constant__ CUtensorMap mapA,
                                                                      const __grid_constant__ CUtensorMap mapB) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t s_bars[kNumBars];
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long HW = (long long)p.H * p.W;
    // raw ring: RG groups of [A boxes (4) | B boxes (NB)], a box = GC channel rows x 128 B (one TMA operation);
    // lo ring: kLoStages slots of [A boxes | B boxes] x 16 channel rows (one 16-channel sub-stage)
    const int nbox = kABoxes + p.NB;
    const uint32_t rbox = (uint32_t)p.GC * 128u;            // bytes of a raw box
    const uint32_t group_bytes = (uint32_t)nbox * rbox;
    const int a_bytes = kABoxes * kBoxBytes;
    const int stage_bytes = nbox * kBoxBytes;                // one lo slot
    const uint32_t raw0 = smem_u32(smem), lo0 = raw0 + (uint32_t)p.RG * group_bytes;
    const uint32_t bars = smem_u32(&s_bars[0]);
#define BAR(which, idx) (bars + 8u * (uint32_t)((which) + (idx)))

    if (tid == 0) {
        for (int i = 0; i < kMaxRawStages; ++i) { mbar_init(BAR(kRawFull, i), 2); mbar_init(BAR(kRawEmpty, i), 1); }
        for (int i = 0; i < kLoStages; ++i) { mbar_init(BAR(kLoFull, i), kXfWarps); mbar_init(BAR(kLoEmpty, i), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(BAR(kAccFull, i), 1); mbar_init(BAR(kAccEmpty, i), kEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = s_tmem;

    if (warp == kLoadWarpA || warp == kLoadWarpB) {
        // =============================== TMA producers ========================================
        if (lane == 0) {
            const bool isA = warp == kLoadWarpA;
            uint32_t gs = 0;
            for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
                const Tile tl = tile_of(p, tile);
                const int nb = isA ? tl.nbA : tl.nbB;
                const uint32_t bytes = (uint32_t)nb * rbox;
                for (int g = 0; g < p.NG; ++g, ++gs) {
                    const uint32_t slot = gs % (uint32_t)p.RG, ph = (gs / (uint32_t)p.RG) & 1u;
                    mbar_wait(BAR(kRawEmpty, slot), ph ^ 1u);
                    const uint32_t full = BAR(kRawFull, slot);
                    if (p.dbg & 8) { mbar_arrive(full); continue; }
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full), "r"(bytes) : "memory");
                    const uint32_t dst = raw0 + slot * group_bytes + (isA ? 0u : (uint32_t)kABoxes * rbox);
                    const CUtensorMap *map = isA ? &mapA : &mapB;
                    const int x0 = isA ? tl.X0 : tl.xp0;
                    for (int i = 0; i < nb; ++i)
                        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                                     ::"r"(dst + (uint32_t)i * rbox), "l"(reinterpret_cast<uint64_t>(map)),
                                       "r"(x0 + i * kBoxX), "r"(tl.y), "r"(g * p.GC), "r"(tl.b), "r"(full) : "memory");
                }
            }
        }
    } else if (warp >= kEpiWarps && warp < kLoadWarpA) {
        // =============================== SPLIT: lo = v - trunc(v) ===============================
        const int t = tid - kEpiWarps * 32;  // 0..255
        uint32_t gs = 0, ls = 0;
        for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
            const Tile tl = tile_of(p, tile);
            // only the boxes this tile loads: A boxes 0 .. nbA-1 and B boxes 4 .. 4+nbB-1 (a row's tail tile has far fewer)
            const int nA = tl.nbA << 7, nAB = (tl.nbA + tl.nbB) << 7;
            for (int g = 0; g < p.NG; ++g, ++gs) {
                const uint32_t rslot = gs % (uint32_t)p.RG, rph = (gs / (uint32_t)p.RG) & 1u;
                mbar_wait(BAR(kRawFull, rslot), rph);
                const unsigned char *grp = smem + (size_t)rslot * group_bytes;
                for (int sub = 0; sub < p.NSUB; ++sub, ++ls) {
                    const uint32_t lslot = ls % (uint32_t)kLoStages, lph = (ls / (uint32_t)kLoStages) & 1u;
                    mbar_wait(BAR(kLoEmpty, lslot), lph ^ 1u);   // the MMAs that read this lo slot have completed
                    float4 *lo = reinterpret_cast<float4 *>(smem + (size_t)p.RG * group_bytes + (size_t)lslot * stage_bytes);
#pragma unroll 4
                    for (int i0 = t; i0 < ((p.dbg & 2) ? 0 : nAB); i0 += kXfThreads) {
                        const int i = i0 < nA ? i0 : i0 - nA + (kABoxes << 7);
                        const int box = i >> 7, w = i & 127;   // 128 float4 = 16 rows of 128 B per box
                        const float4 v = *reinterpret_cast<const float4 *>(grp + (size_t)box * rbox + (size_t)sub * kBoxBytes + (size_t)w * 16);
                        float4 l;
                        l.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
                        l.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
                        l.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
                        l.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
                        lo[i] = l;
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> visible to the tensor core
                    __syncwarp();
                    if (lane == 0) mbar_arrive(BAR(kLoFull, lslot));
                }
            }
        }
    } else if (warp == kMmaWarp) {
        // =============================== MMA ===================================================
        if (lane == 0) {
            uint32_t gs = 0, ls = 0, tcount = 0;
            for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++tcount) {
                const Tile tl = tile_of(p, tile);
                // instruction descriptor: D=F32, A=B=TF32, both MN-major (bits 15, 16), N, M=128
                const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(tl.N >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
                const uint32_t ab = tcount & 1u, aph = (tcount >> 1) & 1u;
                mbar_wait(BAR(kAccEmpty, ab), aph ^ 1u);  // epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t acc = tmem_base + ab * 256u;
                for (int g = 0; g < p.NG; ++g, ++gs) {
                    const uint32_t rslot = gs % (uint32_t)p.RG;
                    for (int sub = 0; sub < p.NSUB; ++sub, ++ls) {
                        const uint32_t lslot = ls % (uint32_t)kLoStages, lph = (ls / (uint32_t)kLoStages) & 1u;
                        mbar_wait(BAR(kLoFull, lslot), lph);   // the split warps have seen RawFull of this group: raw and lo are in place
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t ah = raw0 + rslot * group_bytes + (uint32_t)sub * kBoxBytes, bh = ah + (uint32_t)kABoxes * rbox;
                        const uint32_t al = lo0 + lslot * (uint32_t)stage_bytes, bl = al + a_bytes;
#pragma unroll
                        for (int ks = 0; ks < ((p.dbg & 4) ? 0 : kKC / 8); ++ks) {
                            const uint32_t o = ks * 1024;  // 8 channel rows of 128 B per K step
                            umma_tf32(acc, make_desc(al + o, kBoxBytes), make_desc(bh + o, rbox), idesc, (g > 0 || sub > 0 || ks > 0) ? 1u : 0u);
                            umma_tf32(acc, make_desc(ah + o, rbox), make_desc(bl + o, kBoxBytes), idesc, 1u);
                            umma_tf32(acc, make_desc(ah + o, rbox), make_desc(bh + o, rbox), idesc, 1u);
                        }
                        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(BAR(kLoEmpty, lslot)) : "memory");
                    }
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(BAR(kRawEmpty, rslot)) : "memory");
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(BAR(kAccFull, ab)) : "memory");
            }
        }
    } else {
        // =============================== EPILOGUE (warps 0-3) ===================================
        const int m = warp * 32 + lane;
        const float Cf = (float)p.C, invC = 1.0f / Cf;
        uint32_t tcount = 0;
        for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++tcount) {
            const Tile tl = tile_of(p, tile);
            const uint32_t ab = tcount & 1u, aph = (tcount >> 1) & 1u;
            mbar_wait(BAR(kAccFull, ab), aph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t acc = tmem_base + ab * 256u + ((uint32_t)(warp * 32) << 16);
            if (warp * 32 < tl.len && !(p.dbg & 1)) {   // warp-uniform: this quarter holds valid pixels
                // column of (pixel m, disparity d) = K + m - d
                const int K = tl.X0 - tl.xp0;
                const int c0 = (K + warp * 32 - p.HALO) & ~7;
                float v[16 * NW16];
#pragma unroll
                for (int w = 0; w < NW16; ++w) {
                    if (c0 + 16 * w < 256) {
                        tmem_ld16(acc + (uint32_t)(c0 + 16 * w), v + 16 * w);
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[16 * w + i] = 0.f;
                    }
                }
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                const bool mine = m < tl.len;
                const int sh = mine ? (K + m - p.HALO) - c0 : 0;  // 0 .. 38
                // barrel shift left by sh: after the six stages v[i] holds the original v[i + sh]
#define DSSMD_SHIFT_STAGE(BIT)                                                     \
    {                                                                              \
        const bool on = (sh & BIT) != 0;                                           \
        _Pragma("unroll") for (int i = 0; i + BIT < 16 * NW16; ++i) v[i] = on ? v[i + BIT] : v[i]; \
    }
                DSSMD_SHIFT_STAGE(32) DSSMD_SHIFT_STAGE(16) DSSMD_SHIFT_STAGE(8)
                DSSMD_SHIFT_STAGE(4) DSSMD_SHIFT_STAGE(2) DSSMD_SHIFT_STAGE(1)
#undef DSSMD_SHIFT_STAGE
                // v[k] = S[m][K + m - HALO + k]  <=>  d = HALO - k = D - 1 - k
                if (mine) {
                    const int x = tl.X0 + m;
                    float *o = p.out + (long long)tl.b * p.OBS + (long long)tl.y * p.W + x;
#pragma unroll
                    for (int k = 0; k < 16 * NW16 - 38; ++k) {
                        const int d = p.D - 1 - k;
                        if (d < 0) break;
                        const float val = p.c_pow2 ? v[k] * invC : __fdiv_rn(v[k], Cf);
                        o[(long long)d * HW] = (x >= d) ? val : 0.f;
                    }
                }
            }
            // all of this warp's TMEM reads are complete (tcgen05.wait::ld above): release the accumulator
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(kAccEmpty, ab));
        }
    }
#undef BAR

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == kMmaWarp) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

int env_int_u2(const char *name, int dflt) {
    if (!tuning_enabled()) return dflt;
    const char *v = std::getenv(name);
    return (v && *v) ? std::atoi(v) : dflt;
}

}  // namespace

// Returns -1 when the shape is outside this kernel's limits (the caller takes another kernel).
int corr_fwd_umma2_f32(const void *L, const void *R, void *out, int B, int C, int H, int W, int D, long long obs, cudaStream_t stream) {
    if (W % 4 != 0 || W < 32 || D < 2 || D > 58) return -1;
    if (((reinterpret_cast<uintptr_t>(L) | reinterpret_cast<uintptr_t>(R)) & 15) != 0) return -1;
    if (C % kKC != 0) return -1;
    if ((long long)B * H * ceil_div(W, kTileM) >= (1LL << 30)) return -1;
    Umma2Params p{};
    p.out = (float *)out; p.OBS = obs;
    p.B = B; p.C = C; p.H = H; p.W = W; p.D = D;
    p.HALO = D - 1;
    p.XT = ceil_div(W, kTileM);
    p.ntiles = B * H * p.XT;
    p.c_pow2 = (C & (C - 1)) == 0;
    p.dbg = env_int_u2("DSSMD_UMMA2_DBG", 0);
    // B boxes per stage: the widest tile of a row
    int nb = 0;
    for (int xt = 0; xt < p.XT; ++xt) {
        const int X0 = xt * kTileM, len = W - X0 < kTileM ? W - X0 : kTileM;
        const int xp0 = (X0 - p.HALO) & ~3, ncol = X0 + len - xp0;
        if (((ncol + 15) & ~15) > 256) return -1;
        nb = std::max(nb, ceil_div(ncol, kBoxX));
    }
    p.NB = nb;
    p.n32 = env_int_u2("DSSMD_UMMA2_N32", 0);
    if (nb * kBoxX > 256 && p.n32) return -1;
    // channels per TMA box: as many as divide C and leave room for two raw groups next to the lo ring
    const size_t lo_bytes = (size_t)kLoStages * (kABoxes + nb) * kBoxBytes;
    int gc = env_int_u2("DSSMD_UMMA2_GC", 64);
    if (gc != 16 && gc != 32 && gc != 64) gc = 64;
    while (gc > 16 && (C % gc != 0 || 2 * (size_t)(kABoxes + nb) * gc * 128 + lo_bytes > 224 * 1024)) gc /= 2;
    if (C % gc != 0) return -1;
    p.GC = gc; p.NSUB = gc / kKC; p.NG = C / gc;
    const size_t group = (size_t)(kABoxes + nb) * gc * 128;
    int rg = (int)((224 * 1024 - lo_bytes) / group);
    rg = env_int_u2("DSSMD_UMMA2_STAGES", rg);
    if (rg > kMaxRawStages) rg = kMaxRawStages;
    if (rg < 2) return -1;
    p.RG = rg;
    const size_t smem = (size_t)rg * group + lo_bytes;

    alignas(64) CUtensorMap mapA, mapB;
    const cuuint64_t HWl = (cuuint64_t)H * W;
    const cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)B};
    const cuuint64_t strides[3] = {(cuuint64_t)W * 4, HWl * 4, HWl * 4 * (cuuint64_t)C};
    const cuuint32_t box[4] = {(cuuint32_t)kBoxX, 1, (cuuint32_t)gc, 1};
    if (!make_tmap_f32(&mapA, L, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return -1;
    if (!make_tmap_f32(&mapB, R, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return -1;
    const int cap = sm_count();  // one persistent CTA per SM
    const unsigned grid = (unsigned)(p.ntiles < cap ? p.ntiles : cap);
    const int nw16 = (D + 38 <= 80) ? 5 : 6;
    cudaError_t e;
    if (nw16 == 5) {
        e = kernel_smem_optin(corr_fwd_umma2_kernel<5>, smem);
        if (e != cudaSuccess) return fail(DSSMD_ERR_CUDA, "corr_fwd(umma2): cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e));
        corr_fwd_umma2_kernel<5><<<grid, kThreads, smem, stream>>>(p, mapA, mapB);
    } else {
        e = kernel_smem_optin(corr_fwd_umma2_kernel<6>, smem);
        if (e != cudaSuccess) return fail(DSSMD_ERR_CUDA, "corr_fwd(umma2): cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e));
        corr_fwd_umma2_kernel<6><<<grid, kThreads, smem, stream>>>(p, mapA, mapB);
    }
    if (env_int_u2("DSSMD_CORR_DEBUG", 0))
        fprintf(stderr, "corr_fwd umma2 cfg: XT=%d NB=%d GC=%d NG=%d RG=%d smem=%zu grid=%u nw16=%d tiles=%d\n", p.XT, nb, gc, p.NG, rg, smem, grid, nw16, p.ntiles);
    return check_launch("corr_fwd(umma2)");
}

}  // namespace dssmd
