// corr_fwd_umma.cu -- correlation forward on the 5th-generation tensor cores
// (tcgen05.mma, accumulator in TMEM), fp32 in / fp32 out with 3xTF32 splitting.
//
// Why: the FFMA kernels (corr_fwd.cu) are bound by shared-memory wavefronts -- every
// operand word a thread multiplies has to come through the 128 B/clk LDS pipe, and ncu
// shows that pipe ~85 % busy at ~35 % FMA utilisation.  The tensor core reads its
// operands from shared memory itself, with full reuse across the 128 x N tile.
//
// Formulation (per tile of 128 consecutive pixels of the flattened [B*H, W] image):
//   S[m][n] = sum_c L[c][pixel m] * R[c][column n]        one 128 x N x C GEMM
//   out[d][pixel m] = S[m][n(m) - d] / C                  a diagonal band of S
// where the N columns hold, for every image row the tile touches, the R positions
// x_first - (D-1) ... x_last of that row (left of x = 0 zero-filled = the reference's
// zero border).  Only D of the N columns of a row of S are used (D/N ~ 12-20 %), which
// the tensor core can afford.
//
// fp32 accuracy: a = a_hi + a_lo with a_hi = a rounded to TF32 (10 mantissa bits),
// a_lo = a - a_hi (exact); S ~= Lhi*Rhi + Llo*Rhi + Lhi*Rlo, three kind::tf32 MMAs per
// K step into the same fp32 accumulator.  Dropped Llo*Rlo is 2^-22 relative.
//
// One persistent, warp-specialised CTA per SM (448 threads), mbarrier hand-offs only:
//   warp  12     LOAD       TMA tensor copies (cp.async.bulk.tensor, 128B swizzle with 32B atoms): per stage of
//                           16 channels, boxes {32 pixels, 16 channels} land in shared memory
//                           exactly in the MN-major SWIZZLE_128B_BASE32B layout tcgen05.mma reads
//                           (row = one channel, 128 B = 32 consecutive pixels): 4 boxes of L for
//                           the tile, and for every image row it touches the boxes of R covering
//                           x - (D-1) .. x_last (negative x' and pixels past the image are zero
//                           filled by the TMA unit = the reference's zero border).  No transpose.
//                           (Per-thread cp.async/LDGSTS measured ~25 B/clk/SM on B200: too slow.)
//   warps  4-11  SPLIT      element-wise, in place: hi = tf32(v) overwrites v, lo = v - hi goes to
//                           a second buffer with the same layout (one LDS.128 + 2 STS.128 per 4)
//   warp  13     MMA        one thread issues 6 tcgen05.mma per stage (3 products x 2 K steps of
//                           8 channels) and tcgen05.commit -> frees the stage
//   warps  0-3   EPILOGUE   (= the four TMEM lane quarters) tcgen05.ld the 32+D+7 wide window, a
//                           6-stage barrel shift by (lane - first lane) puts each thread's
//                           diagonal band in registers 0..D-1, then D coalesced 128-byte stores.
//                           The accumulator is double-buffered in TMEM (2 x 256 columns) so the
//                           epilogue of tile t overlaps the main loop of tile t+1.
#include "common.cuh"
#include "tma.cuh"

#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)

#include <cstdlib>

namespace dssmd {

namespace {

constexpr int kTileM = 128;
constexpr int kKC = 16;            // channels per pipeline stage (= 2 MMA K steps of 8)
constexpr int kMaxSeg = 4;         // image rows one 128-pixel tile may touch
constexpr int kStages = 4;        // ring of operand stages: [hi (TMA lands raw here) | lo]
constexpr int kBoxX = 32;         // pixels per TMA box = one 128-byte swizzle row
constexpr int kBoxBytes = kBoxX * kKC * 4;  // 2048
constexpr int kEpiWarps = 4, kXfWarps = 8;
constexpr int kXfThreads = kXfWarps * 32;
constexpr int kLoadWarp = kEpiWarps + kXfWarps;               // 12
constexpr int kMmaWarp = kLoadWarp + 1;                       // 13
constexpr int kThreads = (kEpiWarps + kXfWarps + 2) * 32;     // 448

struct UmmaParams {
    const float *L;
    const float *R;
    float *out;
    long long OBS;  // elements between two images of `out`
    int B, C, H, W, D;
    int P;        // B*H*W pixels (< 2^31)
    int tpi;      // tiles per batch image = ceil(H*W / 128)
    int ntiles;   // B * tpi
    int SMAX;     // max image rows a tile touches
    int HALO;     // D - 1
    int NMMA;     // MMA N (multiple of 16, <= 256)
    int NS;       // pipeline stages per tile = ceil(C / kKC)
    int c_pow2;
    int dbg;      // debug bits: 1 no MMA, 2 no transform math, 4 no cp.async, 8 no epilogue stores/loads
};

struct Segment {
    int row;    // flattened image row b*H + y
    int xa;     // first pixel x of the segment
    int len;    // pixels
    int m0;     // tile-local index of its first pixel
    int col0;   // first B column of the segment (multiple of 4)
    int xp0;    // x' of column col0 (multiple of 4, may be negative)
    int ncol;   // B columns reserved (multiple of 4)
};

// barriers (shared): indices into one array
enum {
    kRawFull = 0,                        // [kStages]  LOAD (TMA complete_tx) -> SPLIT
    kOpFull = kRawFull + kStages,        // [kStages]  SPLIT -> MMA
    kOpEmpty = kOpFull + kStages,        // [kStages]  MMA (tcgen05.commit) -> LOAD
    kAccFull = kOpEmpty + kStages,       // [2]        MMA (tcgen05.commit) -> EPILOGUE
    kAccEmpty = kAccFull + 2,            // [2]           EPILOGUE -> MMA
    kNumBars = kAccEmpty + 2
};

__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    // MN-major tf32 operands only exist in the SWIZZLE_128B_BASE32B layout (32-byte swizzle atoms, 4-row
    // K groups): start>>4 [0,14) | LBO>>4 [16,30) = stride between 32-pixel boxes | SBO>>4 [32,46) =
    // stride between groups of 4 channel rows | version=1 [46,48) | layout=1 [61,64)
    const uint64_t lbo = kBoxBytes >> 4, sbo = 512 >> 4;
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

// rows of the image a tile touches and where their R columns live in the B operand
__device__ __forceinline__ int tile_segments(const UmmaParams &p, int tile, Segment *seg) {
    // tiles never cross a batch image: tile = (b, t) with t < tiles_per_img; rows are flattened b*H + y
    const int b = tile / p.tpi, t = tile - b * p.tpi;
    const int img0 = b * p.H * p.W;
    const int P0 = img0 + t * kTileM;
    const int pend = img0 + p.H * p.W;
    const int qend = (P0 + kTileM < pend) ? P0 + kTileM : pend;
    int n = 0, col = 0, q = P0;
#pragma unroll
    for (int it = 0; it < kMaxSeg; ++it) {
        if (q < qend) {
            const int row = q / p.W;
            const int xa = q - row * p.W;
            const int len = (p.W - xa < qend - q) ? p.W - xa : qend - q;
            seg[it].row = row; seg[it].xa = xa; seg[it].len = len; seg[it].m0 = q - P0;
            seg[it].xp0 = (xa - p.HALO) & ~3;  // TMA box starts must be 16-byte aligned: floor to a multiple of 4
            seg[it].col0 = col;
            seg[it].ncol = ((xa + len - seg[it].xp0 + kBoxX - 1) / kBoxX) * kBoxX;  // whole 32-column boxes
            col += seg[it].ncol;
            q += len;
            n = it + 1;
        } else {
            seg[it].row = 0; seg[it].xa = 0; seg[it].len = 0; seg[it].m0 = kTileM; seg[it].xp0 = 0; seg[it].col0 = 1 << 20; seg[it].ncol = 0;
        }
    }
    return n;
}

template <int NW16>  // epilogue window = 16*NW16 accumulator columns per thread
__global__ void __launch_bounds__(kThreads, 1) corr_fwd_umma_kernel(const UmmaParams p, const __grid_constant__ CUtensorMap mapA,
                                                                     const __grid_constant__ CUtensorMap mapB) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t s_bars[kNumBars];
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long HW = (long long)p.H * p.W;
    // one stage = [hi: A boxes (4) | B boxes (NMMA/32)] [lo: same]; every box is 2048 B = 16 channel rows x 128 B
    const int a_bytes = (kTileM / kBoxX) * kBoxBytes, b_bytes = (p.NMMA / kBoxX) * kBoxBytes;
    const int half_bytes = a_bytes + b_bytes, stage_bytes = 2 * half_bytes;
    const uint32_t bars = smem_u32(&s_bars[0]);
#define BAR(which, idx) (bars + 8u * (uint32_t)((which) + (idx)))

    if (tid == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(BAR(kRawFull, i), 1); mbar_init(BAR(kOpFull, i), kXfWarps); mbar_init(BAR(kOpEmpty, i), 1); }
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


    if (warp == kLoadWarp) {
        // =============================== LOAD ==================================================
        if (lane == 0) {
            uint32_t gs = 0;
            for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
                Segment seg[kMaxSeg];
                const int nseg = tile_segments(p, tile, seg);
                const int b = tile / p.tpi, t = tile - b * p.tpi;
                int nboxB = 0;
#pragma unroll
                for (int k = 0; k < kMaxSeg; ++k) nboxB += seg[k].ncol / kBoxX;
                const uint32_t bytes = (uint32_t)((kTileM / kBoxX + nboxB) * kBoxBytes);
                for (int s = 0; s < p.NS; ++s, ++gs) {
                    const uint32_t slot = gs % kStages, ph = (gs / kStages) & 1u;
                    mbar_wait(BAR(kOpEmpty, slot), ph ^ 1u);
                    const uint32_t full = BAR(kRawFull, slot);
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full), "r"(bytes) : "memory");
                    if (p.dbg & 4) {
                        asm volatile("mbarrier.complete_tx.shared::cta.b64 [%0], %1;" ::"r"(full), "r"(bytes) : "memory");
                        continue;
                    }
                    const uint32_t dst = smem_u32(smem + (size_t)slot * stage_bytes);
#pragma unroll
                    for (int i = 0; i < kTileM / kBoxX; ++i)
                        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                                     ::"r"(dst + (uint32_t)(i * kBoxBytes)), "l"(reinterpret_cast<uint64_t>(&mapA)),
                                       "r"(t * kTileM + i * kBoxX), "r"(s * kKC), "r"(b), "r"(full) : "memory");
#pragma unroll
                    for (int k = 0; k < kMaxSeg; ++k) {
                        if (k >= nseg) break;
                        const int y = seg[k].row - b * p.H;
                        for (int i = 0; i < seg[k].ncol / kBoxX; ++i)
                            asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                                         ::"r"(dst + (uint32_t)(a_bytes + (seg[k].col0 / kBoxX + i) * kBoxBytes)), "l"(reinterpret_cast<uint64_t>(&mapB)),
                                           "r"(seg[k].xp0 + i * kBoxX), "r"(y), "r"(s * kKC), "r"(b), "r"(full) : "memory");
                    }
                }
            }
        }
    } else if (warp >= kEpiWarps && warp < kLoadWarp) {
        // =============================== SPLIT =================================================
        const int t = tid - kEpiWarps * 32;  // 0..255
        const int nvec = half_bytes >> 4;    // float4s per half stage
        uint32_t gs = 0;
        for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
            for (int s = 0; s < p.NS; ++s, ++gs) {
                const uint32_t slot = gs % kStages, ph = (gs / kStages) & 1u;
                mbar_wait(BAR(kRawFull, slot), ph);
                float4 *hi = reinterpret_cast<float4 *>(smem + (size_t)slot * stage_bytes);
                float4 *lo = reinterpret_cast<float4 *>(smem + (size_t)slot * stage_bytes + half_bytes);
                if (p.dbg & 16) {
                    // truncation split: the tensor core reads only the top 19 bits of an fp32 word, so
                    // the raw value already is hi = trunc(v); only lo = v - trunc(v) has to be written
#pragma unroll 2
                    for (int i = t; i < nvec; i += kXfThreads) {
                        const float4 v = hi[i];
                        float4 l;
                        l.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
                        l.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
                        l.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
                        l.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
                        lo[i] = l;
                    }
                } else if (!(p.dbg & 2)) {
#pragma unroll 2
                    for (int i = t; i < nvec; i += kXfThreads) {
                        const float4 v = hi[i];
                        float4 h, l;
                        // round to nearest TF32 (10 mantissa bits): |lo| <= 2^-11 |v|, dropped lo*lo is 2^-22
                        h.x = __uint_as_float((__float_as_uint(v.x) + 0x1000u) & 0xffffe000u); l.x = v.x - h.x;
                        h.y = __uint_as_float((__float_as_uint(v.y) + 0x1000u) & 0xffffe000u); l.y = v.y - h.y;
                        h.z = __uint_as_float((__float_as_uint(v.z) + 0x1000u) & 0xffffe000u); l.z = v.z - h.z;
                        h.w = __uint_as_float((__float_as_uint(v.w) + 0x1000u) & 0xffffe000u); l.w = v.w - h.w;
                        hi[i] = h;
                        lo[i] = l;
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> visible to the tensor core
                __syncwarp();
                if (lane == 0) mbar_arrive(BAR(kOpFull, slot));
            }
        }
    } else if (warp == kMmaWarp) {
        // =============================== MMA ===================================================
        if (lane == 0) {
            // instruction descriptor: D=F32, A=B=TF32, K-major both, N, M=128
            // (A and B are MN-major: bits 15 and 16)
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(p.NMMA >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
            uint32_t gs = 0, tl = 0;
            for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++tl) {
                const uint32_t ab = tl & 1u, aph = (tl >> 1) & 1u;
                mbar_wait(BAR(kAccEmpty, ab), aph ^ 1u);  // epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t acc = tmem_base + ab * 256u;
                for (int s = 0; s < p.NS; ++s, ++gs) {
                    const uint32_t ob = gs % kStages, oph = (gs / kStages) & 1u;
                    mbar_wait(BAR(kOpFull, ob), oph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t ah = smem_u32(smem + (size_t)ob * stage_bytes), bh = ah + a_bytes, al = ah + half_bytes, bl = al + a_bytes;
                    if (p.dbg & 1) { mbar_arrive(BAR(kOpEmpty, ob)); continue; }
#pragma unroll
                    for (int ks = 0; ks < kKC / 8; ++ks) {
                        const uint32_t o = ks * 1024;  // 8 channel rows of 128 B per K step
                        umma_tf32(acc, make_desc(al + o), make_desc(bh + o), idesc, (s > 0 || ks > 0) ? 1u : 0u);
                        umma_tf32(acc, make_desc(ah + o), make_desc(bl + o), idesc, 1u);
                        umma_tf32(acc, make_desc(ah + o), make_desc(bh + o), idesc, 1u);
                    }
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(BAR(kOpEmpty, ob)) : "memory");
                }
                if (p.dbg & 1) mbar_arrive(BAR(kAccFull, ab));
                else asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(BAR(kAccFull, ab)) : "memory");
            }
        }
    } else {
        // =============================== EPILOGUE (warps 0-3) ===================================
        const int m = warp * 32 + lane;
        const float Cf = (float)p.C, invC = 1.0f / Cf;
        uint32_t tl = 0;
        for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++tl) {
            Segment seg[kMaxSeg];
            tile_segments(p, tile, seg);
            const uint32_t ab = tl & 1u, aph = (tl >> 1) & 1u;
            mbar_wait(BAR(kAccFull, ab), aph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t acc = tmem_base + ab * 256u + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
            for (int s = 0; s < kMaxSeg; ++s) {
                const Segment sg = seg[s];
                const int mlo = sg.m0 > warp * 32 ? sg.m0 : warp * 32;
                const int mhi = (sg.m0 + sg.len < warp * 32 + 32) ? sg.m0 + sg.len : warp * 32 + 32;
                if (mlo >= mhi || (p.dbg & 8)) continue;  // warp-uniform
                // column of (pixel m, disparity d) = K + m - d
                const int K = sg.col0 + sg.xa - sg.xp0 - sg.m0;
                const int c0 = (K + mlo - p.HALO) & ~7;
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
                const bool mine = m >= mlo && m < mhi;
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
                    const int x = sg.xa + (m - sg.m0);
                    const int b = sg.row / p.H, y = sg.row - b * p.H;
                    float *o = p.out + (long long)b * p.OBS + (long long)y * p.W + x;
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

int env_int(const char *name, int dflt) {
    if (!tuning_enabled()) return dflt;
    const char *v = std::getenv(name);
    return (v && *v) ? std::atoi(v) : dflt;
}

}  // namespace

namespace {

}  // namespace

// Returns -1 when the shape is outside this kernel's limits (caller falls back to FFMA).
int corr_fwd_umma_f32(const void *L, const void *R, void *out, int B, int C, int H, int W, int D, long long obs, cudaStream_t stream) {
    if (W % 4 != 0 || W < 43 || D < 2 || D > 58) return -1;
    if (((reinterpret_cast<uintptr_t>(L) | reinterpret_cast<uintptr_t>(R)) & 15) != 0) return -1;
    if (C % kKC != 0) return -1;
    if ((long long)B * H * W >= (1LL << 31) - 256) return -1;
    const int smax = 2 + 126 / W;
    if (smax > kMaxSeg) return -1;
    // B columns come in whole 32-pixel boxes per image row the tile touches: exact maximum over the tiles of an image
    int maxbox = 0;
    {
        const int HWi = H * W, tpi = (HWi + kTileM - 1) / kTileM;
        for (int t = 0; t < tpi; ++t) {
            int q = t * kTileM, nb = 0;
            const int qend = q + kTileM < HWi ? q + kTileM : HWi;
            while (q < qend) {
                const int xa = q % W;
                const int len = (W - xa < qend - q) ? W - xa : qend - q;
                nb += (xa + len - ((xa - (D - 1)) & ~3) + kBoxX - 1) / kBoxX;
                q += len;
            }
            if (nb > maxbox) maxbox = nb;
        }
    }
    const int nmma = kBoxX * maxbox;
    if (nmma > 256 || nmma < 16) return -1;
    UmmaParams p{};
    p.L = (const float *)L; p.R = (const float *)R; p.out = (float *)out; p.OBS = obs;
    p.B = B; p.C = C; p.H = H; p.W = W; p.D = D;
    p.P = B * H * W;
    p.tpi = (H * W + kTileM - 1) / kTileM;
    p.ntiles = B * p.tpi;
    p.SMAX = smax;
    p.HALO = D - 1;
    p.NMMA = nmma;
    p.NS = C / kKC;
    p.c_pow2 = (C & (C - 1)) == 0;
    p.dbg = env_int("DSSMD_UMMA_SKIP", 0);
    const size_t smem = (size_t)kStages * 2 * (kTileM + nmma) * kKC * 4;
    if (smem > 225 * 1024) return -1;

    alignas(64) CUtensorMap mapA, mapB;
    const cuuint64_t HWl = (cuuint64_t)H * W;
    {
        const cuuint64_t dims[3] = {HWl, (cuuint64_t)C, (cuuint64_t)B};
        const cuuint64_t strides[2] = {HWl * 4, HWl * 4 * (cuuint64_t)C};
        const cuuint32_t box[3] = {(cuuint32_t)kBoxX, (cuuint32_t)kKC, 1};
        if (!make_tmap_f32(&mapA, L, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return -1;
    }
    {
        const cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)B};
        const cuuint64_t strides[3] = {(cuuint64_t)W * 4, HWl * 4, HWl * 4 * (cuuint64_t)C};
        const cuuint32_t box[4] = {(cuuint32_t)kBoxX, 1, (cuuint32_t)kKC, 1};
        if (!make_tmap_f32(&mapB, R, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return -1;
    }
    const int cap = sm_count();  // one persistent CTA per SM
    const unsigned grid = (unsigned)(p.ntiles < cap ? p.ntiles : cap);
    const int nw16 = (D + 38 <= 80) ? 5 : 6;
    cudaError_t e;
    if (nw16 == 5) {
        e = cudaFuncSetAttribute(corr_fwd_umma_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(DSSMD_ERR_CUDA, "corr_fwd(umma): cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e));
        corr_fwd_umma_kernel<5><<<grid, kThreads, smem, stream>>>(p, mapA, mapB);
    } else {
        e = cudaFuncSetAttribute(corr_fwd_umma_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(DSSMD_ERR_CUDA, "corr_fwd(umma): cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e));
        corr_fwd_umma_kernel<6><<<grid, kThreads, smem, stream>>>(p, mapA, mapB);
    }
    if (env_int("DSSMD_CORR_DEBUG", 0))
        fprintf(stderr, "corr_fwd umma cfg: NMMA=%d NS=%d smem=%zu grid=%u nw16=%d tiles=%d\n", nmma, p.NS, smem, grid, nw16, p.ntiles);
    return check_launch("corr_fwd(umma)");
}

}  // namespace dssmd
