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
//   warp  12     LOAD       bulk async copies (cp.async.bulk = TMA engine, byte-counted on an
//                           mbarrier) of raw fp32 [channel][operand row] stages, ring of 4.
//                           (Per-thread cp.async/LDGSTS was measured at ~25 B/clk/SM issue
//                           throughput on B200 -- not enough to stream at HBM rate.)
//   warps  4-11  TRANSFORM  transpose + hi/lo split of a landed stage into the K-major,
//                           un-swizzled UMMA core-matrix layout (8 rows x 16 B; SBO = 512 B,
//                           LBO = 128 B); two operand buffers
//   warp  13     MMA        one thread issues 6 tcgen05.mma per stage (3 products x 2 K
//                           steps of 8) and tcgen05.commit -> frees the operand buffer
//   warps  0-3   EPILOGUE   (= the four TMEM lane quarters) tcgen05.ld the 32+D+7 wide
//                           window, a 6-stage barrel shift by (lane - first lane) puts each
//                           thread's diagonal band in registers 0..D-1, then D coalesced
//                           128-byte stores.  The accumulator is double-buffered in TMEM
//                           (2 x 256 columns) so the epilogue of tile t overlaps the
//                           main loop of tile t+1.
#include "common.cuh"

#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)

#include <cstdlib>

namespace dssmd {

namespace {

constexpr int kTileM = 128;
constexpr int kKC = 16;            // channels per pipeline stage (= 2 MMA K steps of 8)
constexpr int kChunksK = kKC / 4;  // 16-byte chunks along K per operand row
constexpr int kMaxSeg = 4;         // image rows one 128-pixel tile may touch
constexpr int kOpStages = 2;      // operand (hi/lo, K-major) buffers
constexpr int kRawStages = 3;     // raw fp32 ring filled by TMA tensor copies
constexpr int kEpiWarps = 4, kXfWarps = 8;
constexpr int kXfThreads = kXfWarps * 32;
constexpr int kLoadWarp = kEpiWarps + kXfWarps;               // 12
constexpr int kMmaWarp = kLoadWarp + 1;                       // 13
constexpr int kThreads = (kEpiWarps + kXfWarps + 2) * 32;     // 448
constexpr int kMaxXfItems = 6;     // (128 + 256) rows * 4 chunks / 256 threads

struct UmmaParams {
    const float *L;
    const float *R;
    float *out;
    int B, C, H, W, D;
    int P;        // B*H*W pixels (< 2^31)
    int tpi;      // tiles per batch image = ceil(H*W / 128)
    int ntiles;   // B * tpi
    int SMAX;     // max image rows a tile touches
    int raw_stage_floats;  // 16 * (128 + SMAX * W)
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
    kRawFull = 0,                        // [kRawStages]  LOAD (bulk-copy complete_tx) -> TRANSFORM
    kRawEmpty = kRawFull + kRawStages,   // [kRawStages]  TRANSFORM -> LOAD
    kOpFull = kRawEmpty + kRawStages,    // [kOpStages]   TRANSFORM -> MMA
    kOpEmpty = kOpFull + kOpStages,      // [kOpStages]   MMA (tcgen05.commit) -> TRANSFORM
    kAccFull = kOpEmpty + kOpStages,     // [2]           MMA (tcgen05.commit) -> EPILOGUE
    kAccEmpty = kAccFull + 2,            // [2]           EPILOGUE -> MMA
    kNumBars = kAccEmpty + 2
};

__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// bounded wait: a protocol bug must trap, not hang the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (int spin = 0; spin < (1 << 26); ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
    }
    __trap();
}

__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    // K-major, no swizzle: start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48)
    const uint64_t lbo = 128 >> 4, sbo = (kChunksK * 128) >> 4;
    return (uint64_t)((smem_addr >> 4) & 0x3fff) | (lbo << 16) | (sbo << 32) | (1ull << 46);
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

// operand byte offset of (row r, 16-byte K chunk j) in the K-major core-matrix layout
__device__ __forceinline__ int op_off(int r, int j) { return ((r >> 3) * kChunksK + j) * 128 + (r & 7) * 16; }

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
            seg[it].xp0 = xa - p.HALO;
            seg[it].col0 = col;
            seg[it].ncol = len + p.HALO;
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
    const int a_bytes = kTileM * kKC * 4, b_bytes = p.NMMA * kKC * 4;
    const int op_bytes = 2 * (a_bytes + b_bytes);  // one operand buffer: A_hi | A_lo | B_hi | B_lo
    unsigned char *ops = smem;                     // [kOpStages][op_bytes]
    const uint32_t bars = smem_u32(&s_bars[0]);
#define BAR(which, idx) (bars + 8u * (uint32_t)((which) + (idx)))

    if (tid == 0) {
        for (int i = 0; i < kRawStages; ++i) { mbar_init(BAR(kRawFull, i), 1); mbar_init(BAR(kRawEmpty, i), kXfWarps); }
        for (int i = 0; i < kOpStages; ++i) { mbar_init(BAR(kOpFull, i), kXfWarps); mbar_init(BAR(kOpEmpty, i), 1); }
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

    const int NR = kTileM + p.NMMA;      // operand rows: [A pixels | B columns]
    // raw stage: A box [kKC][128] then, per image row the tile touches, a full R row box [kKC][W]
    float *raw = reinterpret_cast<float *>(smem + (size_t)kOpStages * op_bytes);  // [kRawStages][raw_stage_floats]

    if (warp == kLoadWarp) {
        // =============================== LOAD ==================================================
        // TMA tensor copies, one thread: per stage one box {128 pixels, 16 channels} of L from
        // the [B][C][H*W] view (a tile is contiguous in H*W; pixels past the image are zero
        // filled by the TMA unit) and, per image row the tile touches, one box {W, 1 row, 16
        // channels} of R from the [B][C][H][W] view.  Completion is byte-counted on the mbarrier.
        if (lane == 0) {
            uint32_t gs = 0;
            for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
                Segment seg[kMaxSeg];
                const int nseg = tile_segments(p, tile, seg);
                const int b = tile / p.tpi, t = tile - b * p.tpi;
                const uint32_t bytes = (uint32_t)(kKC * (kTileM + nseg * p.W) * 4);
                for (int s = 0; s < p.NS; ++s, ++gs) {
                    const uint32_t slot = gs % kRawStages, ph = (gs / kRawStages) & 1u;
                    mbar_wait(BAR(kRawEmpty, slot), ph ^ 1u);
                    const uint32_t full = BAR(kRawFull, slot);
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full), "r"(bytes) : "memory");
                    if (p.dbg & 4) {
                        asm volatile("mbarrier.complete_tx.shared::cta.b64 [%0], %1;" ::"r"(full), "r"(bytes) : "memory");
                        continue;
                    }
                    const uint32_t dst = smem_u32(raw + (size_t)slot * p.raw_stage_floats);
                    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(&mapA)), "r"(t * kTileM), "r"(s * kKC), "r"(b), "r"(full) : "memory");
#pragma unroll
                    for (int k = 0; k < kMaxSeg; ++k) {
                        if (k >= nseg) break;
                        const int y = seg[k].row - b * p.H;
                        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                                     ::"r"(dst + 4u * (uint32_t)(kKC * (kTileM + k * p.W))), "l"(reinterpret_cast<uint64_t>(&mapB)),
                                       "r"(0), "r"(y), "r"(s * kKC), "r"(b), "r"(full) : "memory");
                    }
                }
            }
        }
    } else if (warp >= kEpiWarps && warp < kLoadWarp) {
        // =============================== TRANSFORM =============================================
        // transpose + hi/lo split of a landed raw stage into an operand buffer
        const int t = tid - kEpiWarps * 32;  // 0..255
        // item k of this thread = (K chunk j, operand row r); lanes run over rows (conflict-free both ways)
        int t_j[kMaxXfItems], t_dst[kMaxXfItems], t_lo[kMaxXfItems], t_row[kMaxXfItems];
#pragma unroll
        for (int k = 0; k < kMaxXfItems; ++k) {
            const int i = t + k * kXfThreads;
            t_j[k] = 0; t_dst[k] = 0; t_lo[k] = 0; t_row[k] = -1;
            if (i < NR * kChunksK) {
                const int j = i / NR, r = i - j * NR;
                t_j[k] = j;
                t_row[k] = r;
                const bool isB = r >= kTileM;
                t_dst[k] = (isB ? 2 * a_bytes : 0) + op_off(isB ? r - kTileM : r, j);
                t_lo[k] = isB ? b_bytes : a_bytes;
            }
        }
        uint32_t gs = 0;
        for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
            // where each operand row of this tile lives in a raw stage (float offset of its channel 4j)
            // and its channel stride; -1: the row is all zero (x' < 0, unused column, pixel past the image)
            Segment seg[kMaxSeg];
            tile_segments(p, tile, seg);
            int src_off[kMaxXfItems], src_cs[kMaxXfItems];
#pragma unroll
            for (int k = 0; k < kMaxXfItems; ++k) {
                src_off[k] = -1; src_cs[k] = 0;
                const int r = t_row[k];
                if (r < 0) continue;
#pragma unroll
                for (int s = 0; s < kMaxSeg; ++s) {
                    if (seg[s].len == 0) continue;
                    if (r < kTileM) {
                        if (r >= seg[s].m0 && r < seg[s].m0 + seg[s].len) { src_off[k] = 4 * t_j[k] * kTileM + r; src_cs[k] = kTileM; }
                    } else {
                        const int n = r - kTileM;
                        const int xp = seg[s].xp0 + (n - seg[s].col0);
                        if (n >= seg[s].col0 && n < seg[s].col0 + seg[s].ncol && xp >= 0) {
                            src_off[k] = kKC * (kTileM + s * p.W) + 4 * t_j[k] * p.W + xp; src_cs[k] = p.W;
                        }
                    }
                }
            }
            for (int s = 0; s < p.NS; ++s, ++gs) {
                const uint32_t slot = gs % kRawStages, rph = (gs / kRawStages) & 1u;
                const uint32_t ob = gs % kOpStages, oph = (gs / kOpStages) & 1u;
                mbar_wait(BAR(kRawFull, slot), rph);
                mbar_wait(BAR(kOpEmpty, ob), oph ^ 1u);
                const float *rs = raw + (size_t)slot * p.raw_stage_floats;
                unsigned char *opb = ops + (size_t)ob * op_bytes;
#pragma unroll
                for (int k = 0; k < kMaxXfItems; ++k) {
                    if (t_row[k] < 0 || (p.dbg & 2)) continue;
                    float hi[4], lo[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float v = src_off[k] >= 0 ? rs[src_off[k] + q * src_cs[k]] : 0.f;
                        // round to nearest TF32 (10 mantissa bits): |lo| <= 2^-11 |v|, dropped lo*lo is 2^-22
                        hi[q] = __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u);
                        lo[q] = v - hi[q];
                    }
                    *reinterpret_cast<float4 *>(opb + t_dst[k]) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<float4 *>(opb + t_dst[k] + t_lo[k]) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> visible to the tensor core
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(BAR(kOpFull, ob));
                    mbar_arrive(BAR(kRawEmpty, slot));
                }
            }
        }
    } else if (warp == kMmaWarp) {
        // =============================== MMA ===================================================
        if (lane == 0) {
            // instruction descriptor: D=F32, A=B=TF32, K-major both, N, M=128
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.NMMA >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
            uint32_t gs = 0, tl = 0;
            for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++tl) {
                const uint32_t ab = tl & 1u, aph = (tl >> 1) & 1u;
                mbar_wait(BAR(kAccEmpty, ab), aph ^ 1u);  // epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t acc = tmem_base + ab * 256u;
                for (int s = 0; s < p.NS; ++s, ++gs) {
                    const uint32_t ob = gs % kOpStages, oph = (gs / kOpStages) & 1u;
                    mbar_wait(BAR(kOpFull, ob), oph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t ah = smem_u32(ops + (size_t)ob * op_bytes), al = ah + a_bytes, bh = ah + 2 * a_bytes, bl = bh + b_bytes;
                    if (p.dbg & 1) { mbar_arrive(BAR(kOpEmpty, ob)); continue; }
#pragma unroll
                    for (int ks = 0; ks < kKC / 8; ++ks) {
                        const uint32_t o = ks * 2 * 128;  // two 16-byte K chunks per step
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
                    float *o = p.out + ((long long)b * p.D * p.H + y) * p.W + x;
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
    const char *v = std::getenv(name);
    return (v && *v) ? std::atoi(v) : dflt;
}

}  // namespace

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = [] {
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            sym = nullptr;
        return reinterpret_cast<EncodeTiledFn>(sym);
    }();
    return fn;
}

// fp32 tensor map; dims/box fastest first; strides in bytes for dims 1..rank-1
bool make_tmap_f32(CUtensorMap *m, const void *base, int rank, const cuuint64_t *dims, const cuuint64_t *strides, const cuuint32_t *box) {
    EncodeTiledFn fn = encode_tiled();
    if (!fn) return false;
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void *>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

// Returns -1 when the shape is outside this kernel's limits (caller falls back to FFMA).
int corr_fwd_umma_f32(const void *L, const void *R, void *out, int B, int C, int H, int W, int D, cudaStream_t stream) {
    if (W % 4 != 0 || W < 43 || W > 256 || D < 2 || D > 58) return -1;
    if (((reinterpret_cast<uintptr_t>(L) | reinterpret_cast<uintptr_t>(R)) & 15) != 0) return -1;
    if (C % kKC != 0) return -1;
    if ((long long)B * H * W >= (1LL << 31) - 256) return -1;
    const int smax = 2 + 126 / W;
    if (smax > kMaxSeg) return -1;
    const int nmma = round_up(kTileM + smax * (D - 1), 16);
    if (nmma > 256) return -1;
    UmmaParams p{};
    p.L = (const float *)L; p.R = (const float *)R; p.out = (float *)out;
    p.B = B; p.C = C; p.H = H; p.W = W; p.D = D;
    p.P = B * H * W;
    p.tpi = (H * W + kTileM - 1) / kTileM;
    p.ntiles = B * p.tpi;
    p.SMAX = smax;
    p.raw_stage_floats = kKC * (kTileM + smax * W);
    p.HALO = D - 1;
    p.NMMA = nmma;
    p.NS = C / kKC;
    p.c_pow2 = (C & (C - 1)) == 0;
    p.dbg = env_int("DSSMD_UMMA_SKIP", 0);
    const size_t op_bytes = (size_t)2 * (kTileM + nmma) * kKC * 4;
    const size_t smem = kOpStages * op_bytes + (size_t)kRawStages * p.raw_stage_floats * 4;
    if (smem > 225 * 1024) return -1;

    alignas(64) CUtensorMap mapA, mapB;
    const cuuint64_t HWl = (cuuint64_t)H * W;
    {
        const cuuint64_t dims[3] = {HWl, (cuuint64_t)C, (cuuint64_t)B};
        const cuuint64_t strides[2] = {HWl * 4, HWl * 4 * (cuuint64_t)C};
        const cuuint32_t box[3] = {(cuuint32_t)kTileM, (cuuint32_t)kKC, 1};
        if (!make_tmap_f32(&mapA, L, 3, dims, strides, box)) return -1;
    }
    {
        const cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)B};
        const cuuint64_t strides[3] = {(cuuint64_t)W * 4, HWl * 4, HWl * 4 * (cuuint64_t)C};
        const cuuint32_t box[4] = {(cuuint32_t)W, 1, (cuuint32_t)kKC, 1};
        if (!make_tmap_f32(&mapB, R, 4, dims, strides, box)) return -1;
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
