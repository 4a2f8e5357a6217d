// epipolar_attn_umma.cu -- forward of the attention along the epipolar line on the 5th-generation tensor cores (tcgen05 + TMEM),
// fp32 in / fp32 out at the fp32 parity bar.  Problem and upstream expression: epipolar_attn.cu (models/BidNet/STTM_Simple.py:43-57);
// numerics: epipolar_attn_mma.cu -- every operand is divided by a power of two >= its tile's largest magnitude (exact), stored
// times 2^10 as an fp16 hi / lo pair (22 significant bits), and every product is hi*hi + lo*hi + hi*lo with fp32 accumulation.
// The mma.sync kernel is bound by the legacy tensor path (50 % of 949 MAC/clk/SM); kind::f16 on tcgen05 runs at ~4x that rate
// and takes its operands from shared memory, so the warps only stage tiles and do the softmax.
//
// One CTA per SM owns 128 queries of one image row at a time (M = 128 = the TMEM lanes) and walks over key tiles of 128:
//   S  = Q~^T K~   tcgen05.mma M128 N128 K16 x 2 channel steps x 3 products  -> TMEM [128 lanes x 128 columns], two buffers
//   P  = exp2(S fs - m)  one softmax THREAD per query row (tcgen05.ld 32x32b: no shuffles), split into fp16 hi / lo in registers
//   O += P V~^T    tcgen05.mma M128 N32 K16 x 8 key steps x 3 products, A = P from TMEM  -> TMEM [128 x 32] per query tile
// Operand tiles are fp16 planes in the no-swizzle canonical layouts (8 x 16-byte core matrices): the staging threads read 8
// consecutive pixels of one channel from the NCHW tensor and write ONE 16-byte unit -- for Q and K that is the MN-major
// layout of a [pixels x channels] operand, for V the K-major layout of a [channels x keys] operand: the same address
// arithmetic.  P never touches shared memory: its threads write the packed halves over the S buffer they came from (tcgen05.st)
// and the second GEMM takes its A operand from tensor memory.
// Warps: 0-7 softmax -- TMEM lane quarter = warp % 4, and the two warps of a quarter split a row's 128 score columns (and its 32
// output channels) in halves: TWO threads per query row, their row maxima exchanged through shared memory once per tile --,
// 8-10 staging (global fp32 -> scale -> hi / lo planes), 11 MMA issue (one lane).  A thread pulls its 64 scores into registers
// with one tcgen05.ld pair and releases the S buffer at once; P is double-buffered, so the exponentials of tile t + 1 run under
// the P V of tile t and only the rescaling of the O row (online softmax: when a row's maximum or the V scale unit changes,
// tcgen05.ld / st of 16 columns) waits for it.  Deterministic.
#include "common.cuh"
#include "tma.cuh"

#include <cstdlib>

namespace dssmd {

namespace {

constexpr int kUC = 32;            // channels
constexpr int kUM = 128;           // queries per tile
constexpr int kUN = 128;           // keys per tile
constexpr int kUG = 5;             // query tiles per group: K / V tiles are staged once per group (5 x 32 O columns + 2 x 128 S columns of TMEM)
constexpr int kUThreads = 384;     // 8 softmax + 3 staging + 1 MMA warp (13 warps would be allocated as 16: 128 registers per thread)
constexpr float kUUp = 1024.0f;    // up-shift of stored operands (see epipolar_attn_mma.cu)
constexpr float kUUpP = 1024.0f;   // probabilities

// shared memory (bytes): Q hi/lo 2 x 8192 per query tile of the group | K hi/lo, V hi/lo per buffer 4 x 8192, two buffers (P lives in TMEM)
constexpr int kPlane = kUM * kUC * 2;                 // 8192
constexpr int kOffQ = 0, kOffKV = kUG * 2 * kPlane, kUSmem = kOffKV + 2 * 4 * kPlane;   // 147456

enum { kKvFull = 0, kKvEmpty = 2, kSFull = 4, kSEmpty = 6, kPFull = 8, kODone = 9, kUBars = 10 };

__device__ __forceinline__ void u_split(float x, float y, uint32_t &hi, uint32_t &lo) {
    const __half2 h = __floats2half2_rn(x, y);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(x - hf.x, y - hf.y);
    hi = *reinterpret_cast<const uint32_t *>(&h);
    lo = *reinterpret_cast<const uint32_t *>(&l);
}
__device__ __forceinline__ float u_pow2_ceil(float m) {
    const uint32_t u = __float_as_uint(m);
    const uint32_t e = u >> 23;
    if (e == 0u || e >= 254u) return 1.0f;
    return __uint_as_float(((u & 0x007fffffu) ? e + 1u : e) << 23);
}
__device__ __forceinline__ float u_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// shared-memory matrix descriptor, no swizzle (layout type 0), version 1: start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32
__device__ __forceinline__ uint64_t u_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void u_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 (or 16) consecutive accumulator columns of this thread's TMEM lane; the caller issues tcgen05.wait::ld
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, float *v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
          "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
          "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
          "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st32u_nowait(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
          "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]),
          "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// A operand from tensor memory (K-major: lane = row, two 16-bit K elements per 32-bit column), B from shared memory
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void u_wait(uint32_t bar, uint32_t parity) {   // bounded: a protocol bug traps instead of hanging the GPU
    for (int spin = 0; spin < (1 << 27); ++spin) {
        uint32_t done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
    }
    __trap();
}

// [32 channels][128 pixels] fp32 tile of a row of an NCHW tensor (pixels from x0, zero past W) -> fp16 hi / lo planes of 16-byte
// units: unit (channel c, pixel group xg) at xg * 512 + (c / 8) * 128 + (c % 8) * 16.  NT threads, 512 units.
// Two phases around `sync` (a barrier over the NT threads): the tile's largest magnitude first.
template <int NT, typename Sync>
__device__ __forceinline__ float stage_tile(const float *rowbase, size_t HW, int x0, int W, unsigned char *hi, unsigned char *lo, unsigned *slot,
                                            int t, int lane, Sync sync) {
    constexpr int NU = (512 + NT - 1) / NT;   // units per thread (the last round may be partial)
    float4 v[NU][2];
    float m = 0.f;
#pragma unroll
    for (int i = 0; i < NU; ++i) {
        const int u = t + i * NT, c = (u >> 4) & 31, xg = u & 15;
        const int x = x0 + 8 * xg;
        const bool on = (512 % NT == 0) || u < 512;
        const float *src = rowbase + (size_t)c * HW + x;
        v[i][0] = on && x < W ? ld_cg(reinterpret_cast<const float4 *>(src)) : make_float4(0.f, 0.f, 0.f, 0.f);
        v[i][1] = on && x + 4 < W ? ld_cg(reinterpret_cast<const float4 *>(src + 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int h = 0; h < 2; ++h) m = fmaxf(m, fmaxf(fmaxf(fabsf(v[i][h].x), fabsf(v[i][h].y)), fmaxf(fabsf(v[i][h].z), fabsf(v[i][h].w))));
    }
    const unsigned wm = __reduce_max_sync(0xffffffffu, __float_as_uint(m));
    if (lane == 0) atomicMax(slot, wm);
    sync();
    const float s = u_pow2_ceil(__uint_as_float(*slot));
    const float f = kUUp / s;
#pragma unroll
    for (int i = 0; i < NU; ++i) {
        const int u = t + i * NT, c = u >> 4, xg = u & 15;
        if ((512 % NT != 0) && u >= 512) break;
        uint32_t h[4], l[4];
        u_split(v[i][0].x * f, v[i][0].y * f, h[0], l[0]);
        u_split(v[i][0].z * f, v[i][0].w * f, h[1], l[1]);
        u_split(v[i][1].x * f, v[i][1].y * f, h[2], l[2]);
        u_split(v[i][1].z * f, v[i][1].w * f, h[3], l[3]);
        const int off = xg * 512 + (c >> 3) * 128 + (c & 7) * 16;
        *reinterpret_cast<uint4 *>(hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4 *>(lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
    }
    return s;
}

__global__ void __launch_bounds__(kUThreads, 1)
epipolar_attn_fwd_umma_kernel(const float *q, const float *k, const float *v, float *__restrict__ out, float *__restrict__ lse,
                              int H, int W, float scale, int groups_per_row, int ngroups_total) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t s_bars[kUBars];
    __shared__ uint32_t s_tmem;
    __shared__ unsigned s_qmax[kUG], s_kmax[2], s_vmax[2];
    __shared__ float s_sk[2], s_sv[2];
    __shared__ float s_rmax[2][2][kUM];   // [pair parity][column half][row]: the halves' row maxima of a (key tile, query tile) pair
    __shared__ float s_lsum[2][kUM];      // [column half][row]: the halves' row sums at the end of a query tile

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t HW = (size_t)H * W;
    const uint32_t sb = smem_u32(smem), bars = smem_u32(&s_bars[0]);
#define UBAR(i) (bars + 8u * (uint32_t)(i))
    const int nkt = (W + kUN - 1) / kUN;   // key tiles per row
    const int nqt = (W + kUM - 1) / kUM;   // query tiles per row

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(UBAR(kKvFull + i), 3); mbar_init(UBAR(kKvEmpty + i), 1);
            mbar_init(UBAR(kSFull + i), 1); mbar_init(UBAR(kSEmpty + i), 1);
        }
        mbar_init(UBAR(kPFull), 8);
        mbar_init(UBAR(kODone), 1);
        for (int g = 0; g < kUG; ++g) s_qmax[g] = 0u;
        s_kmax[0] = s_kmax[1] = 0u; s_vmax[0] = s_vmax[1] = 0u;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 11) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    pdl_launch_dependents();
    pdl_wait();

    // running counters of barrier uses: every role walks the same sequence of key tiles (kv) and (key tile, query tile) pairs (pc)
    uint32_t kv_count = 0, pair_count = 0;
    for (int grp = blockIdx.x; grp < ngroups_total; grp += gridDim.x) {
        const int r = grp / groups_per_row, qt0 = (grp - r * groups_per_row) * kUG;
        const int G = nqt - qt0 < kUG ? nqt - qt0 : kUG;     // query tiles of this group
        const int b = r / H, y = r - b * H;
        const size_t rowoff = (size_t)b * kUC * HW + (size_t)y * W;

        // ---- the group's Q tiles: the eight softmax warps stage them, then a block barrier (it also separates consecutive groups: the
        // MMAs that read the previous Q planes completed before the softmax threads left that group) ------------------------------
        if (warp < 8) {
            for (int g = 0; g < G; ++g)
                stage_tile<256>(q + rowoff, HW, (qt0 + g) * kUM, W, smem + kOffQ + g * 2 * kPlane, smem + kOffQ + g * 2 * kPlane + kPlane, &s_qmax[g], tid, lane,
                                [] { asm volatile("bar.sync 1, 256;" ::: "memory"); });
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        __syncthreads();
        float sq[kUG];
#pragma unroll
        for (int g = 0; g < kUG; ++g) sq[g] = u_pow2_ceil(__uint_as_float(s_qmax[g]));
        __syncthreads();
        if (tid < kUG) s_qmax[tid] = 0u;

        if (warp >= 8 && warp < 11) {
            // =============================== staging: K and V tiles (once per group) ===============
            const int t = tid - 256;
            for (int kt = 0; kt < nkt; ++kt) {
                const uint32_t c = kv_count + (uint32_t)kt, buf = c & 1u, ph = (c >> 1) & 1u;
                u_wait(UBAR(kKvEmpty + buf), ph ^ 1u);   // the P V products of the key tile that used this buffer have completed
                unsigned char *base = smem + kOffKV + buf * 4 * kPlane;
                auto sync = [] { asm volatile("bar.sync 2, 96;" ::: "memory"); };
                const float sk = stage_tile<96>(k + rowoff, HW, kt * kUN, W, base, base + kPlane, &s_kmax[buf], t, lane, sync);
                const float sv = stage_tile<96>(v + rowoff, HW, kt * kUN, W, base + 2 * kPlane, base + 3 * kPlane, &s_vmax[buf], t, lane, sync);
                if (t == 0) { s_sk[buf] = sk; s_sv[buf] = sv; }
                sync();                                   // everybody has read the maxima; the slots may be cleared for their next use
                if (t == 0) { s_kmax[buf] = 0u; s_vmax[buf] = 0u; }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> visible to the tensor core
                __syncwarp();
                if (lane == 0) mbar_arrive(UBAR(kKvFull + buf));
            }
        } else if (warp == 11) {
            // =============================== MMA issue ============================================
            // ONE thread issues ~30 MMAs per (key tile, query tile) pair: its instruction stream is on the critical path, so the
            // descriptors are built once per group / key tile and advanced by adding to their address field
            if (lane == 0) {
                const uint32_t idesc1 = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(kUN >> 3) << 17) | ((uint32_t)(kUM >> 4) << 24);   // F16 x F16 -> F32, A, B MN-major
                const uint32_t idesc2 = (1u << 4) | ((uint32_t)(kUC >> 3) << 17) | ((uint32_t)(kUM >> 4) << 24);                            // A, B K-major, N = 32
                const uint64_t dq0 = u_desc(sb + kOffQ, 128, 512);            // Q hi of query tile 0; lo: + kPlane; tile g: + g * 2 * kPlane
                const uint64_t dk0 = u_desc(sb + kOffKV, 128, 512);           // K hi of buffer 0; lo: + kPlane; buffer 1: + 4 * kPlane
                const uint64_t dv0 = u_desc(sb + kOffKV + 2 * kPlane, 512, 128);   // V hi of buffer 0
                constexpr uint64_t PL = kPlane >> 4;                          // descriptor address units are 16 bytes
                const int npairs = nkt * G;
                int kt1 = 0, g1 = 0;          // the pair gemm1 takes next
                auto gemm1 = [&]() {          // S[c & 1] = Q_g~^T K_kt~ for pair (kt1, g1), then advance
                    const uint32_t c = pair_count + (uint32_t)(kt1 * G + g1), sbuf = c & 1u, sph = (c >> 1) & 1u;
                    const uint32_t ck = kv_count + (uint32_t)kt1, kbuf = ck & 1u;
                    if (g1 == 0) u_wait(UBAR(kKvFull + kbuf), (ck >> 1) & 1u);
                    u_wait(UBAR(kSEmpty + sbuf), sph ^ 1u);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint64_t qh = dq0 + (uint64_t)g1 * 2 * PL, ql = qh + PL;
                    const uint64_t kh = dk0 + (uint64_t)kbuf * 4 * PL, kl = kh + PL;
                    const uint32_t sd = tmem + sbuf * kUN;
                    umma_f16(sd, ql, kh, idesc1, 0u);
                    umma_f16(sd, qh, kl, idesc1, 1u);
                    umma_f16(sd, qh, kh, idesc1, 1u);
                    umma_f16(sd, ql + 16, kh + 16, idesc1, 1u);   // second 16 channels: two channel groups of 128 B further
                    umma_f16(sd, qh + 16, kl + 16, idesc1, 1u);
                    umma_f16(sd, qh + 16, kh + 16, idesc1, 1u);
                    u_commit(UBAR(kSFull + sbuf));
                    if (++g1 == G) { g1 = 0; ++kt1; }
                };
                gemm1();
                int kt = 0, g = 0;
                for (int pi = 0; pi < npairs; ++pi) {
                    const uint32_t c = pair_count + (uint32_t)pi;
                    const uint32_t kbuf = (kv_count + (uint32_t)kt) & 1u;
                    if (pi + 1 < npairs) gemm1();         // the next S while the softmax threads work on this one
                    u_wait(UBAR(kPFull), c & 1u);         // P written, O rows rescaled
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint64_t vh = dv0 + (uint64_t)kbuf * 4 * PL, vl = vh + PL;
                    const uint32_t od = tmem + 2 * kUN + (uint32_t)g * kUC;
                    const uint32_t pth = tmem + (c & 1u) * kUN, ptl = pth + 64;   // P hi / lo over the S buffer they were computed from
                    umma_f16_ts(od, ptl, vh, idesc2, kt > 0 ? 1u : 0u);
                    umma_f16_ts(od, pth, vl, idesc2, 1u);
                    umma_f16_ts(od, pth, vh, idesc2, 1u);
#pragma unroll
                    for (int ks = 1; ks < kUN / 16; ++ks) {   // 16 keys = 8 columns of packed halves | two key chunks of V (1024 B)
                        umma_f16_ts(od, ptl + ks * 8, vh + ks * 64, idesc2, 1u);
                        umma_f16_ts(od, pth + ks * 8, vl + ks * 64, idesc2, 1u);
                        umma_f16_ts(od, pth + ks * 8, vh + ks * 64, idesc2, 1u);
                    }
                    u_commit(UBAR(kODone));                              // O_g updated
                    u_commit(UBAR(kSEmpty + (c & 1u)));                  // P read: the S buffer may take the pair after next
                    if (g == G - 1) u_commit(UBAR(kKvEmpty + kbuf));     // K / V buffer free
                    if (++g == G) { g = 0; ++kt; }
                }
            }
        } else if (warp < 8) {
            // =============================== softmax: two threads per query row ===================
            const int quarter = warp & 3, half = warp >> 2;
            const int m = quarter * 32 + lane;                    // query row inside a tile = TMEM lane
            const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
            const float qs0 = scale * 1.4426950408889634f * (1.0f / (kUUp * kUUp));
            float mrun[kUG], lrun[kUG], svu[kUG];                 // per query tile: running maximum (exp2 domain), this half's sum, V unit of O
#pragma unroll
            for (int g = 0; g < kUG; ++g) { mrun[g] = -INFINITY; lrun[g] = 0.f; svu[g] = 0.f; }
            auto rowsync = [] { asm volatile("bar.sync 3, 256;" ::: "memory"); };
            uint32_t c = pair_count;
            for (int kt = 0; kt < nkt; ++kt) {
                const uint32_t kbuf = (kv_count + (uint32_t)kt) & 1u;
#pragma unroll
                for (int g = 0; g < kUG; ++g) {
                    if (g >= G) break;
                    const uint32_t sbuf = c & 1u, sph = (c >> 1) & 1u;
                    u_wait(UBAR(kSFull + sbuf), sph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const float fs = qs0 * sq[g] * s_sk[kbuf];
                    const float sv_t = s_sv[kbuf];
                    float sc[64];                                     // this thread's 64 scores of the pair
                    tmem_ld32_nowait(tmem + lane_addr + sbuf * kUN + half * 64, sc);
                    tmem_ld32_nowait(tmem + lane_addr + sbuf * kUN + half * 64 + 32, sc + 32);
                    tmem_wait_ld();
                    const int j0 = kt * kUN + half * 64;              // first key of this thread's columns
                    float mx = -INFINITY;
                    if (j0 + 64 <= W) {                               // uniform over the warp
#pragma unroll
                        for (int i = 0; i < 64; ++i) mx = fmaxf(mx, sc[i]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 64; ++i) { sc[i] = (j0 + i < W) ? sc[i] : -INFINITY; mx = fmaxf(mx, sc[i]); }
                    }
                    mx *= fs;                                         // fs > 0: the maximum of the scaled scores (-inf stays -inf)
                    s_rmax[c & 1u][half][m] = mx;
                    rowsync();
                    const float mn = fmaxf(mrun[g], fmaxf(mx, s_rmax[c & 1u][half ^ 1][m]));   // finite: every key tile holds a key
                    const float corr = u_exp2(mrun[g] - mn);          // 0 on the first key tile
                    const float sv_new = fmaxf(svu[g], sv_t);
                    const float unit = svu[g] / sv_new, pvu = sv_t / sv_new * kUUpP;           // exact powers of two
                    const float off = mn - (float)((int)(__float_as_uint(pvu) >> 23) - 127);   // exp2(s fs - off) = p * pvu
                    mrun[g] = mn; svu[g] = sv_new;
                    // probabilities -> packed fp16 hi / lo pairs, written over this thread's half of the S buffer (every thread of the
                    // row pair has its scores in registers: the row barrier above); all of it runs under the P V of the previous pair
                    float psum = 0.f;
                    uint32_t phi[32], plo[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float p0 = u_exp2(fmaf(sc[2 * i], fs, -off)), p1 = u_exp2(fmaf(sc[2 * i + 1], fs, -off));   // exp2(-inf) = 0 for masked keys
                        psum += p0 + p1;
                        u_split(p0, p1, phi[i], plo[i]);
                    }
                    lrun[g] = fmaf(lrun[g], corr, psum * (1.0f / pvu));   // exact rescale back to probabilities
                    tmem_st32u_nowait(tmem + lane_addr + sbuf * kUN + half * 32, phi);
                    tmem_st32u_nowait(tmem + lane_addr + sbuf * kUN + 64 + half * 32, plo);
                    tmem_wait_st();
                    if (c != pair_count) {   // the previous P V may still be updating an O tile
                        u_wait(UBAR(kODone), (c - 1u) & 1u);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    }
                    if (kt > 0) {   // rescale this half's 16 channels of O_g
                        const float f = corr * unit;
                        if (__any_sync(0xffffffffu, f != 1.0f)) {
                            float o16[16];
                            tmem_ld16(tmem + lane_addr + 2 * kUN + g * kUC + half * 16, o16);
#pragma unroll
                            for (int i = 0; i < 16; ++i) o16[i] *= f;
                            tmem_st16(tmem + lane_addr + 2 * kUN + g * kUC + half * 16, o16);
                        }
                    }
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(UBAR(kPFull));
                    ++c;
                }
            }
            // ---- epilogue of the group: O_g * sv / l -> out[b, c, y, x] (this half: 16 channels); lse in natural units -----------
            u_wait(UBAR(kODone), (c - 1u) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int g = 0; g < kUG; ++g) {
                if (g >= G) break;
                s_lsum[half][m] = lrun[g];
                float o16[16];
                tmem_ld16(tmem + lane_addr + 2 * kUN + g * kUC + half * 16, o16);
                rowsync();
                const float ltot = s_lsum[0][m] + s_lsum[1][m];   // fixed order: both halves compute the same bits
                rowsync();                                        // s_lsum is reused by the next query tile
                const int x = (qt0 + g) * kUM + m;
                if (x < W) {
                    const float inv = svu[g] * (1.0f / (kUUp * kUUpP)) / ltot;
                    float *orow = out + rowoff + (size_t)(half * 16) * HW + x;
#pragma unroll
                    for (int ch = 0; ch < 16; ++ch) orow[(size_t)ch * HW] = o16[ch] * inv;
                    if (lse && half == 0) lse[(size_t)r * W + x] = (mrun[g] + log2f(ltot)) * 0.6931471805599453f;
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        }
        kv_count += (uint32_t)nkt;
        pair_count += (uint32_t)(nkt * G);
        __syncthreads();   // the group's TMEM / shared-memory state is dead: next group
    }
#undef UBAR
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 11) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

int env_int_au(const char *name, int dflt) {
    if (!tuning_enabled()) return dflt;
    const char *v = std::getenv(name);
    return (v && *v) ? std::atoi(v) : dflt;
}

}  // namespace

// fp32 tensors, C = 32, W % 8 == 0: -1 = not eligible (the caller takes the mma.sync kernel), else a DSSMD_* code.
int epipolar_attn_fwd_umma(const float *q, const float *k, const float *v, float *out, float *lse, int B, int C, int H, int W, float scale,
                           cudaStream_t s) {
    if (env_int_au("DSSMD_ATTN_FWD_UMMA", 1) == 0) return -1;
    if (C != kUC || W % 8 != 0) return -1;
    auto kern = epipolar_attn_fwd_umma_kernel;
    const size_t smem = kUSmem;
    const cudaError_t ea = kernel_smem_optin(kern, smem);
    if (ea != cudaSuccess) return fail(DSSMD_ERR_CUDA, "epipolar_attn_fwd(umma): cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(ea));
    const int qtiles = ceil_div(ceil_div(W, kUM), kUG);   // groups of up to kUG query tiles per image row
    const long long total = (long long)B * H * qtiles;
    if (total >= (1LL << 31)) return -1;
    int grid = sm_count();
    if (grid > total) grid = (int)total;
    if (env_int_au("DSSMD_CORR_DEBUG", 0))
        fprintf(stderr, "epipolar_attn_fwd umma cfg: grid=%d groups=%lld (x%d per row, <= 5 query tiles each) key tiles=%d threads=%d smem=%zu\n", grid, total, qtiles,
                ceil_div(W, kUN), kUThreads, smem);
    const cudaError_t e = launch_pdl(kern, dim3((unsigned)grid), dim3(kUThreads), smem, s, q, k, v, out, lse, H, W, scale, qtiles, (int)total);
    if (e != cudaSuccess) return fail(DSSMD_ERR_CUDA, "epipolar_attn_fwd(umma): launch failed: %s", cudaGetErrorString(e));
    return check_launch("epipolar_attn_fwd(umma)");
}

}  // namespace dssmd
