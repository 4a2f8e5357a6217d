// epipolar_attn_mma.cu -- forward of the attention along the epipolar line (see epipolar_attn.cu for the problem and upstream's
// expression, models/BidNet/STTM_Simple.py:43-57) on the warp-level tensor instructions, fp32 in / fp32 out at the fp32 parity bar.
//
// The FFMA kernel runs at 31 TFLOP/s -- it is a pair of GEMMs (S = Q^T K over the C channels, O = P V over the W keys) on CUDA
// cores.  fp32 accuracy on tensor cores needs operand splitting; with tf32 (3xTF32, the correlation kernels' scheme) the legacy
// mma.sync path tops out at 476 MAC/clk/SM / 3 products = 92 TFLOP/s at a full pipe.  The fp16 instruction runs at twice that rate
// (949 MAC/clk/SM, scripts/micro/mma_rate.cu) and a pair of halves carries the same 22 significand bits as a tf32 pair, PROVIDED the
// values sit in fp16's range -- so every staged tile is divided by a power of two >= its largest magnitude first (exact), and
//     x / s = hi + lo,   hi = fp16(x / s),   lo = fp16(x / s - hi)        (stored times 2^10, see kUp: 22 significant bits)
//     a . b ~= hi_a hi_b + lo_a hi_b + hi_a lo_b                          (fp32 accumulation; the dropped lo lo term is 2^-22)
// with the scales folded back in fp32: S = (sq sk scale) (Q~^T K~); the V scale rides on the online-softmax correction (the
// accumulator is kept in units of the largest V scale seen so far, a change of units is an exact power-of-two multiply).
// Probabilities are <= 1 by construction and are split the same way in registers.
//
// Block = 128 queries of one image row (8 warps x 16 queries), key tiles of 64.  Q / K / V tiles are read as fp32 rows of the
// NCHW tensors ([c][x], pixels contiguous), split once while they are staged, and kept in shared memory as fp16 hi / lo planes
// in the same [c][x] layout (row stride an odd number of 16-byte chunks: conflict-free ldmatrix).  That layout is the transpose
// of what the Q^T and K fragments want -> ldmatrix.trans; V^T is the B operand of the second GEMM as it lies (plain ldmatrix).
// The S accumulators of two neighbouring 8-key tiles ARE the A fragment of a 16-key step of the second GEMM (the usual
// flash-attention register reuse), so P never touches shared memory.  Online softmax in the exp2 domain, one running maximum
// and sum per query row (two rows per thread, reduced over the four lanes of a quad).  Deterministic (fixed order).
#include "common.cuh"

#include <cstdlib>

namespace dssmd {

namespace {

constexpr int kQT = 128;          // queries per block
constexpr int kKT = 64;           // keys per tile
constexpr int kQS = kQT + 8;      // row strides in halves: 272 B = 17 chunks, 144 B = 9 chunks of 16 bytes
constexpr int kKS = kKT + 8;
constexpr int kAThreads = 256;
// Every split operand is stored multiplied by a power of two on top of its normalisation to [-1, 1]: the `lo` halves of values
// down to 2^-13 of the tile's scale then stay in fp16's NORMAL range (>= 2^-14) and keep their 11 bits -- without it the lo of a
// typical element (|x / s| ~ 0.1, lo ~ 2^-15) is subnormal, the pair carries an absolute 2^-25 instead of a relative 2^-22, and
// the backward's recomputed probabilities land at 1.0e-5 of the reference instead of ~1e-6.  All factors are folded back in fp32.
constexpr float kUp = 1024.0f;     // q, k, v, dO tiles: |stored| <= 1024
constexpr float kUpP = 1024.0f;    // probabilities (<= 1)
constexpr float kUpD = 256.0f;     // D~ = P (dP~ - delta~), |D~| <= 2 C = 64

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// (x, y) -> packed fp16 hi pair and lo pair (the remainders of the first rounding)
__device__ __forceinline__ void split_h2(float x, float y, uint32_t &hi, uint32_t &lo) {
    const __half2 h = __floats2half2_rn(x, y);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(x - hf.x, y - hf.y);
    hi = *reinterpret_cast<const uint32_t *>(&h);
    lo = *reinterpret_cast<const uint32_t *>(&l);
}
// smallest power of two >= m (m >= 0 finite), 1 for m == 0 or a non-finite m (NaN / inf then propagate through the products)
__device__ __forceinline__ float pow2_ceil(float m) {
    const uint32_t u = __float_as_uint(m);
    const uint32_t e = u >> 23;
    if (e == 0u || e >= 254u) return 1.0f;                        // zero / subnormal / huge / inf / NaN
    return __uint_as_float(((u & 0x007fffffu) ? e + 1u : e) << 23);
}
__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ---- staging: a [C][T] tile of a row of an NCHW tensor, fp32 -> registers -> (scale) -> fp16 hi / lo planes -------------------
// NV = float4 vectors per thread: C * T / 4 / 256
template <int C, int T>
struct TileRegs {
    static constexpr int NV = C * T / 4 / kAThreads;
    float4 v[NV];
    __device__ __forceinline__ void load(const float *rowbase, size_t HW, int x0, int W, int tid) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int idx = tid + i * kAThreads;
            const int c = idx / (T / 4), f = idx % (T / 4);
            const int x = x0 + 4 * f;
            v[i] = (x < W) ? ld_cg(reinterpret_cast<const float4 *>(rowbase + (size_t)c * HW + x)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    __device__ __forceinline__ float absmax() const {
        float m = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) m = fmaxf(m, fmaxf(fmaxf(fabsf(v[i].x), fabsf(v[i].y)), fmaxf(fabsf(v[i].z), fabsf(v[i].w))));
        return m;   // fmaxf drops NaNs: a NaN element is scaled by whatever the others give and stays NaN
    }
    template <int STRIDE>
    __device__ __forceinline__ void store(__half *hi, __half *lo, float inv, int tid) const {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int idx = tid + i * kAThreads;
            const int c = idx / (T / 4), f = idx % (T / 4);
            uint32_t h0, l0, h1, l1;
            split_h2(v[i].x * inv, v[i].y * inv, h0, l0);
            split_h2(v[i].z * inv, v[i].w * inv, h1, l1);
            *reinterpret_cast<uint2 *>(hi + c * STRIDE + 4 * f) = make_uint2(h0, h1);
            *reinterpret_cast<uint2 *>(lo + c * STRIDE + 4 * f) = make_uint2(l0, l1);
        }
    }
};

// block-wide maximum of a non-negative value into *slot (unsigned bits order like the floats); the caller synchronises
__device__ __forceinline__ void block_max_into(unsigned *slot, float m, int lane) {
    const unsigned w = __reduce_max_sync(0xffffffffu, __float_as_uint(m));
    if (lane == 0) atomicMax(slot, w);
}

template <int CH>   // C = 16 * CH channels
__global__ void __launch_bounds__(kAThreads, 2)
epipolar_attn_fwd_mma_kernel(const float *q, const float *k, const float *v, float *__restrict__ out, float *__restrict__ lse,
                             int H, int W, float scale) {
    constexpr int C = 16 * CH;
    constexpr int NO = C / 8;             // 8-channel output tiles per warp
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned s_qmax, s_kmax[2], s_vmax[2];
    __half *Qh = reinterpret_cast<__half *>(smem_raw), *Ql = Qh + C * kQS;
    __half *Kh = Ql + C * kQS;            // [2][C][kKS] each
    __half *Kl = Kh + 2 * C * kKS, *Vh = Kl + 2 * C * kKS, *Vl = Vh + 2 * C * kKS;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, tq = lane & 3;
    const int r = blockIdx.y, b = r / H, y = r - b * H;
    const int x0 = blockIdx.x * kQT;
    const size_t HW = (size_t)H * W;
    const size_t rowoff = (size_t)b * C * HW + (size_t)y * W;
    const float *qrow = q + rowoff, *krow = k + rowoff, *vrow = v + rowoff;

    if (tid == 0) { s_qmax = 0u; s_kmax[0] = s_kmax[1] = 0u; s_vmax[0] = s_vmax[1] = 0u; }
    pdl_launch_dependents();
    pdl_wait();
    __syncthreads();

    // ---- Q tile and the first K / V tile --------------------------------------------------------------------------------------
    TileRegs<C, kQT> qr;
    TileRegs<C, kKT> kr, vr;
    qr.load(qrow, HW, x0, W, tid);
    kr.load(krow, HW, 0, W, tid);
    vr.load(vrow, HW, 0, W, tid);
    block_max_into(&s_qmax, qr.absmax(), lane);
    block_max_into(&s_kmax[0], kr.absmax(), lane);
    block_max_into(&s_vmax[0], vr.absmax(), lane);
    __syncthreads();
    const float sq = pow2_ceil(__uint_as_float(s_qmax));
    float sk = pow2_ceil(__uint_as_float(s_kmax[0]));
    float sv = pow2_ceil(__uint_as_float(s_vmax[0]));   // units of the O accumulator: the largest V scale so far
    qr.template store<kQS>(Qh, Ql, kUp / sq, tid);
    kr.template store<kKS>(Kh, Kl, kUp / sk, tid);
    vr.template store<kKS>(Vh, Vl, kUp / sv, tid);
    __syncthreads();

    // this warp's Q^T fragments (16 queries x C channels), hi and lo, for the whole kernel
    uint32_t qh[CH][4], ql[CH][4];
    {
        const int xq = warp * 16 + ((lane >> 3) & 1) * 8;
        const int cq = (lane & 7) + ((lane >> 4) & 1) * 8;
#pragma unroll
        for (int ks = 0; ks < CH; ++ks) {
            ldsm_x4_t(smem_u32(Qh + (ks * 16 + cq) * kQS + xq), qh[ks]);
            ldsm_x4_t(smem_u32(Ql + (ks * 16 + cq) * kQS + xq), ql[ks]);
        }
    }

    float o[NO][4];
#pragma unroll
    for (int n = 0; n < NO; ++n)
#pragma unroll
        for (int i = 0; i < 4; ++i) o[n][i] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;   // rows g and g + 8 of this warp's 16 queries (exp2 domain)
    const float qscale = scale * 1.4426950408889634f * sq * (1.0f / (kUp * kUp));

    const int ntiles = (W + kKT - 1) / kKT;
    for (int t = 0; t < ntiles; ++t) {
        const int buf = t & 1;
        const bool more = t + 1 < ntiles;
        if (more) {   // the next tile's rows towards registers while this one is computed
            kr.load(krow, HW, (t + 1) * kKT, W, tid);
            vr.load(vrow, HW, (t + 1) * kKT, W, tid);
        }
        const __half *kh = Kh + buf * C * kKS, *kl = Kl + buf * C * kKS, *vh = Vh + buf * C * kKS, *vl = Vl + buf * C * kKS;

        // ---- S = Q~^T K~: 16 queries x 64 keys, three products per 16-channel step (small terms first) ------------------------
        float s[8][4];
#pragma unroll
        for (int n = 0; n < 8; ++n)
#pragma unroll
            for (int i = 0; i < 4; ++i) s[n][i] = 0.f;
        {
            // B fragments of two neighbouring 8-key tiles per ldmatrix.x4.trans: matrices (c 0-7, n), (c 8-15, n), (c 0-7, n+1), (c 8-15, n+1)
            const int cb = (lane & 7) + ((lane >> 3) & 1) * 8;
            const int nb = ((lane >> 4) & 1) * 8;
#pragma unroll
            for (int ks = 0; ks < CH; ++ks) {
#pragma unroll
                for (int npp = 0; npp < 2; ++npp) {   // four accumulators at a time: the three products of one accumulator are 4 MMAs apart
                    uint32_t bh[2][4], bl[2][4];
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const int off = (ks * 16 + cb) * kKS + (2 * npp + u) * 16 + nb;
                        ldsm_x4_t(smem_u32(kh + off), bh[u]);
                        ldsm_x4_t(smem_u32(kl + off), bl[u]);
                    }
#pragma unroll
                    for (int u = 0; u < 2; ++u) { mma16816(s[4 * npp + 2 * u], ql[ks], bh[u][0], bh[u][1]); mma16816(s[4 * npp + 2 * u + 1], ql[ks], bh[u][2], bh[u][3]); }
#pragma unroll
                    for (int u = 0; u < 2; ++u) { mma16816(s[4 * npp + 2 * u], qh[ks], bl[u][0], bl[u][1]); mma16816(s[4 * npp + 2 * u + 1], qh[ks], bl[u][2], bl[u][3]); }
#pragma unroll
                    for (int u = 0; u < 2; ++u) { mma16816(s[4 * npp + 2 * u], qh[ks], bh[u][0], bh[u][1]); mma16816(s[4 * npp + 2 * u + 1], qh[ks], bh[u][2], bh[u][3]); }
                }
            }
        }

        // ---- online softmax (exp2 domain): thread holds rows g (s[n][0..1]) and g + 8 (s[n][2..3]), keys 8 n + 2 tq + {0, 1} ----
        const float fs = qscale * sk;
        const int j0 = t * kKT + 2 * tq;
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            const bool in0 = j0 + 8 * n < W, in1 = j0 + 8 * n + 1 < W;
            s[n][0] = in0 ? s[n][0] * fs : -INFINITY; s[n][1] = in1 ? s[n][1] * fs : -INFINITY;
            s[n][2] = in0 ? s[n][2] * fs : -INFINITY; s[n][3] = in1 ? s[n][3] * fs : -INFINITY;
            mx0 = fmaxf(mx0, fmaxf(s[n][0], s[n][1]));
            mx1 = fmaxf(mx1, fmaxf(s[n][2], s[n][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);   // finite: every tile holds at least one key
        // V units: the accumulator moves to the larger scale (exact power-of-two factor), this tile was stored in units of sv_t
        const float sv_t = pow2_ceil(__uint_as_float(s_vmax[buf]));
        const float sv_new = fmaxf(sv, sv_t);
        const float unit = sv / sv_new;            // <= 1, exact
        const float pv_unit = sv_t / sv_new * kUpP;   // this tile's V~ values are in units of sv_t: fold the change (and P's up-shift) into P
        const float c0 = fast_exp2(m0 - mn0), c1 = fast_exp2(m1 - mn1);   // exp2(-inf) = 0 on the first tile
        m0 = mn0; m1 = mn1; sv = sv_new;
        float ps0 = 0.f, ps1 = 0.f;
        uint32_t ph[4][4], pl[4][4];   // A fragments of the second GEMM: 16-key step ks = 8-key tiles 2 ks, 2 ks + 1
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            const float p0 = fast_exp2(s[n][0] - mn0), p1 = fast_exp2(s[n][1] - mn0);
            const float p2 = fast_exp2(s[n][2] - mn1), p3 = fast_exp2(s[n][3] - mn1);
            ps0 += p0 + p1; ps1 += p2 + p3;
            split_h2(p0 * pv_unit, p1 * pv_unit, ph[n >> 1][(n & 1) * 2 + 0], pl[n >> 1][(n & 1) * 2 + 0]);
            split_h2(p2 * pv_unit, p3 * pv_unit, ph[n >> 1][(n & 1) * 2 + 1], pl[n >> 1][(n & 1) * 2 + 1]);
        }
        l0 = fmaf(l0, c0, ps0); l1 = fmaf(l1, c1, ps1);   // per-thread partial sums: the quad is reduced once at the end
        const float f0 = c0 * unit, f1 = c1 * unit;
#pragma unroll
        for (int n = 0; n < NO; ++n) { o[n][0] *= f0; o[n][1] *= f0; o[n][2] *= f1; o[n][3] *= f1; }

        // ---- O += P V~^T: keys in 4 steps of 16, channels in tiles of 8; V [c][j] is the col-major B operand as it lies ---------
        {
            // two 8-channel tiles per ldmatrix.x4: matrices (c-tile n, j 0-7), (c-tile n, j 8-15), (c-tile n+1, j 0-7), (c-tile n+1, j 8-15)
            const int cv = (lane & 7) + ((lane >> 4) & 1) * 8;
            const int jv = ((lane >> 3) & 1) * 8;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                uint32_t bh[NO / 2][4], bl[NO / 2][4];
#pragma unroll
                for (int np = 0; np < NO / 2; ++np) {
                    const int off = (np * 16 + cv) * kKS + ks * 16 + jv;
                    ldsm_x4(smem_u32(vh + off), bh[np]);
                    ldsm_x4(smem_u32(vl + off), bl[np]);
                }
#pragma unroll
                for (int np = 0; np < NO / 2; ++np) { mma16816(o[2 * np], pl[ks], bh[np][0], bh[np][1]); mma16816(o[2 * np + 1], pl[ks], bh[np][2], bh[np][3]); }
#pragma unroll
                for (int np = 0; np < NO / 2; ++np) { mma16816(o[2 * np], ph[ks], bl[np][0], bl[np][1]); mma16816(o[2 * np + 1], ph[ks], bl[np][2], bl[np][3]); }
#pragma unroll
                for (int np = 0; np < NO / 2; ++np) { mma16816(o[2 * np], ph[ks], bh[np][0], bh[np][1]); mma16816(o[2 * np + 1], ph[ks], bh[np][2], bh[np][3]); }
            }
        }

        // ---- stage the next tile into the other buffer --------------------------------------------------------------------------
        if (more) {
            block_max_into(&s_kmax[buf ^ 1], kr.absmax(), lane);
            block_max_into(&s_vmax[buf ^ 1], vr.absmax(), lane);
            __syncthreads();   // both maxima complete; everybody has finished computing on buffer buf ^ 1 (two tiles ago)
            // slot buf was last read in this tile's softmax, which every thread has left; its next atomicMax comes after the second barrier
            if (tid == 0) { s_kmax[buf] = 0u; s_vmax[buf] = 0u; }
            sk = pow2_ceil(__uint_as_float(s_kmax[buf ^ 1]));
            const float svn = pow2_ceil(__uint_as_float(s_vmax[buf ^ 1]));
            kr.template store<kKS>(Kh + (buf ^ 1) * C * kKS, Kl + (buf ^ 1) * C * kKS, kUp / sk, tid);
            vr.template store<kKS>(Vh + (buf ^ 1) * C * kKS, Vl + (buf ^ 1) * C * kKS, kUp / svn, tid);
            __syncthreads();   // tile t + 1 staged
        }
    }

    // ---- epilogue: O * sv / l -> out[b, c, y, x]; lse in natural units --------------------------------------------------------------
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = sv * (1.0f / (kUp * kUpP)) / l0, i1 = sv * (1.0f / (kUp * kUpP)) / l1;
    const int xa = x0 + warp * 16 + g, xb = xa + 8;
    float *orow = out + rowoff;
#pragma unroll
    for (int n = 0; n < NO; ++n) {
        const int c = 8 * n + 2 * tq;
        if (xa < W) { orow[(size_t)c * HW + xa] = o[n][0] * i0; orow[(size_t)(c + 1) * HW + xa] = o[n][1] * i0; }
        if (xb < W) { orow[(size_t)c * HW + xb] = o[n][2] * i1; orow[(size_t)(c + 1) * HW + xb] = o[n][3] * i1; }
    }
    if (lse && tq == 0) {
        const float ln2 = 0.6931471805599453f;
        if (xa < W) lse[(size_t)r * W + xa] = (m0 + log2f(l0)) * ln2;
        if (xb < W) lse[(size_t)r * W + xb] = (m1 + log2f(l1)) * ln2;
    }
}

// ================================================================================================================================
// Backward on the same machinery.  With lse and delta[x] = sum_c dO[c,x] O[c,x] (FlashAttention's recomputation):
//     P = exp(scale Q^T K - lse[x]);  dP = dO^T V;  dS = P o (dP - delta[x])
//     dQ[c,x] = scale sum_j dS[x,j] K[c,j];   dK[c,j] = scale sum_x dS[x,j] Q[c,x];   dV[c,j] = sum_x P[x,j] dO[c,x]
// Seven GEMMs (S and dP are recomputed by both kernels), all with fp16 hi / lo split operands.  Here the scales are one power
// of two per IMAGE ROW and tensor (the largest |q|, |k|, |v|, |dO| of the row, found by the launch that also writes delta):
// a per-tile scale of V would leave delta / (s_dO s_V) unbounded for fp16, with per-row scales
//     D~ = P (dP~ - delta~),  dP~ = dO~^T V~,  delta~ = delta / (s_dO s_V),  |D~| <= 2 C
// and the three results are D~ K~, D~^T Q~, P^T dO~ times constants of the row.  Both kernels walk 16 columns at a time: the S and dP
// accumulators of a 16-column step become the A fragments of the gradient GEMMs of that step at once (few live registers).

// one block per image row: delta[r, x] and stats[r] = (s_q, s_k, s_v, s_dO)
__global__ void __launch_bounds__(256) attn_bwd_prep_kernel(const float *q, const float *k, const float *v, const float *g_out, const float *out,
                                                            float *__restrict__ delta, float *__restrict__ stats, int C, int H, int W) {
    __shared__ unsigned s_m[4];
    const int tid = threadIdx.x, lane = tid & 31;
    const int r = blockIdx.x, b = r / H, y = r - b * H;
    const size_t HW = (size_t)H * W;
    const size_t rowoff = (size_t)b * C * HW + (size_t)y * W;
    if (tid < 4) s_m[tid] = 0u;
    pdl_launch_dependents();
    pdl_wait();
    __syncthreads();
    float mq = 0.f, mk = 0.f, mv = 0.f, mg = 0.f;
    auto amax4 = [](float m, const float4 &a) { return fmaxf(m, fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w)))); };
    for (int x = 4 * tid; x < W; x += 4 * 256) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int c = 0; c < C; ++c) {
            const size_t o = rowoff + (size_t)c * HW + x;
            const float4 a = ld_cg(reinterpret_cast<const float4 *>(g_out + o)), p = ld_cg(reinterpret_cast<const float4 *>(out + o));
            const float4 fq = ld_cg(reinterpret_cast<const float4 *>(q + o)), fk = ld_cg(reinterpret_cast<const float4 *>(k + o));
            const float4 fv = ld_cg(reinterpret_cast<const float4 *>(v + o));
            acc.x = fmaf(a.x, p.x, acc.x); acc.y = fmaf(a.y, p.y, acc.y); acc.z = fmaf(a.z, p.z, acc.z); acc.w = fmaf(a.w, p.w, acc.w);
            mq = amax4(mq, fq); mk = amax4(mk, fk); mv = amax4(mv, fv); mg = amax4(mg, a);
        }
        *reinterpret_cast<float4 *>(delta + (size_t)r * W + x) = acc;
    }
    block_max_into(&s_m[0], mq, lane); block_max_into(&s_m[1], mk, lane);
    block_max_into(&s_m[2], mv, lane); block_max_into(&s_m[3], mg, lane);
    __syncthreads();
    if (tid < 4) stats[(size_t)r * 4 + tid] = pow2_ceil(__uint_as_float(s_m[tid]));
}

// A fragments of a 16-column step from the values of its two 8-column accumulator tiles (rows g / g + 8)
__device__ __forceinline__ void frag_from_acc(const float (&t0)[4], const float (&t1)[4], uint32_t (&hi)[4], uint32_t (&lo)[4]) {
    split_h2(t0[0], t0[1], hi[0], lo[0]);
    split_h2(t0[2], t0[3], hi[1], lo[1]);
    split_h2(t1[0], t1[1], hi[2], lo[2]);
    split_h2(t1[2], t1[3], hi[3], lo[3]);
}
// KEYMAJOR = false: block rows = 128 queries, columns = keys, writes dQ.  true: rows = 128 keys, columns = queries, writes dK and dV.
// Row side (X1, X2) = (Q, dO) | (K, V); column side (Y1, Y2) = (K, V) | (Q, dO).  S (or S^T) = X1^T Y1, dP (or dP^T) = X2^T Y2.
template <int CH, bool KEYMAJOR>
__global__ void __launch_bounds__(kAThreads, 2)
epipolar_attn_bwd_mma_kernel(const float *q, const float *k, const float *v, const float *g_out, const float *lse, const float *delta,
                             const float *stats, float *__restrict__ g1, float *__restrict__ g2, int H, int W, float scale) {
    constexpr int C = 16 * CH;
    constexpr int NO = C / 8;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ float s_l2[2][kKT], s_dl[2][kKT];   // KEYMAJOR: lse (exp2 domain) and delta~ of the column tile's queries
    __half *X1h = reinterpret_cast<__half *>(smem_raw), *X1l = X1h + C * kQS, *X2h = X1l + C * kQS, *X2l = X2h + C * kQS;
    __half *Y1h = X2l + C * kQS;          // [2][C][kKS] each
    __half *Y1l = Y1h + 2 * C * kKS, *Y2h = Y1l + 2 * C * kKS, *Y2l = Y2h + 2 * C * kKS;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, tq = lane & 3;
    const int r = blockIdx.y, b = r / H, y = r - b * H;
    const int x0 = blockIdx.x * kQT;
    const size_t HW = (size_t)H * W;
    const size_t rowoff = (size_t)b * C * HW + (size_t)y * W;
    const float *x1 = (KEYMAJOR ? k : q) + rowoff, *x2 = (KEYMAJOR ? v : g_out) + rowoff;
    const float *y1 = (KEYMAJOR ? q : k) + rowoff, *y2 = (KEYMAJOR ? g_out : v) + rowoff;
    const float *lrow = lse + (size_t)r * W, *drow = delta + (size_t)r * W;
    pdl_launch_dependents();
    pdl_wait();
    const float sq = ld_cg(stats + (size_t)r * 4 + 0), sk = ld_cg(stats + (size_t)r * 4 + 1);
    const float sv = ld_cg(stats + (size_t)r * 4 + 2), sg = ld_cg(stats + (size_t)r * 4 + 3);
    const float ix1 = kUp / (KEYMAJOR ? sk : sq), ix2 = kUp / (KEYMAJOR ? sv : sg);
    const float iy1 = kUp / (KEYMAJOR ? sq : sk), iy2 = kUp / (KEYMAJOR ? sg : sv);
    const float kLog2e = 1.4426950408889634f;
    const float fs = scale * kLog2e * sq * sk * (1.0f / (kUp * kUp));
    const float dn = 1.0f / (kUp * kUp);   // dP~ back to products of [-1, 1] operands
    const float idl = 1.0f / (sg * sv);

    TileRegs<C, kQT> xr;
    TileRegs<C, kKT> yr1, yr2;
    float cl = 0.f, cd = 0.f;   // KEYMAJOR: thread tid < 64 carries one column's lse / delta towards shared memory
    xr.load(x1, HW, x0, W, tid);
    xr.template store<kQS>(X1h, X1l, ix1, tid);
    xr.load(x2, HW, x0, W, tid);
    xr.template store<kQS>(X2h, X2l, ix2, tid);
    yr1.load(y1, HW, 0, W, tid);
    yr2.load(y2, HW, 0, W, tid);
    yr1.template store<kKS>(Y1h, Y1l, iy1, tid);
    yr2.template store<kKS>(Y2h, Y2l, iy2, tid);
    if (KEYMAJOR && tid < kKT) {
        s_l2[0][tid] = tid < W ? ld_cg(lrow + tid) * kLog2e : 0.f;
        s_dl[0][tid] = tid < W ? ld_cg(drow + tid) * idl : 0.f;
    }
    __syncthreads();

    uint32_t a1h[CH][4], a1l[CH][4], a2h[CH][4], a2l[CH][4];   // this warp's 16 rows of X1^T and X2^T
    {
        const int xq = warp * 16 + ((lane >> 3) & 1) * 8;
        const int cq = (lane & 7) + ((lane >> 4) & 1) * 8;
#pragma unroll
        for (int ks = 0; ks < CH; ++ks) {
            ldsm_x4_t(smem_u32(X1h + (ks * 16 + cq) * kQS + xq), a1h[ks]);
            ldsm_x4_t(smem_u32(X1l + (ks * 16 + cq) * kQS + xq), a1l[ks]);
            ldsm_x4_t(smem_u32(X2h + (ks * 16 + cq) * kQS + xq), a2h[ks]);
            ldsm_x4_t(smem_u32(X2l + (ks * 16 + cq) * kQS + xq), a2l[ks]);
        }
    }
    // softmax statistics belong to the QUERY: the row here (queries as rows) or the column (keys as rows)
    const int xa = x0 + warp * 16 + g, xb = xa + 8;
    float rl0 = 0.f, rl1 = 0.f, rd0 = 0.f, rd1 = 0.f;
    if (!KEYMAJOR) {
        rl0 = xa < W ? ld_cg(lrow + xa) * kLog2e : 0.f; rl1 = xb < W ? ld_cg(lrow + xb) * kLog2e : 0.f;
        rd0 = xa < W ? ld_cg(drow + xa) * idl : 0.f;    rd1 = xb < W ? ld_cg(drow + xb) * idl : 0.f;
    }

    float acc1[NO][4], acc2[KEYMAJOR ? NO : 1][4];
#pragma unroll
    for (int n = 0; n < NO; ++n)
#pragma unroll
        for (int i = 0; i < 4; ++i) { acc1[n][i] = 0.f; if (KEYMAJOR) acc2[n][i] = 0.f; }

    const int ntiles = (W + kKT - 1) / kKT;
    for (int t = 0; t < ntiles; ++t) {
        const int buf = t & 1;
        const bool more = t + 1 < ntiles;
        if (more) {
            yr1.load(y1, HW, (t + 1) * kKT, W, tid);
            yr2.load(y2, HW, (t + 1) * kKT, W, tid);
            if (KEYMAJOR && tid < kKT) {
                const int xc = (t + 1) * kKT + tid;
                cl = xc < W ? ld_cg(lrow + xc) * kLog2e : 0.f;
                cd = xc < W ? ld_cg(drow + xc) * idl : 0.f;
            }
        }
        const __half *y1h = Y1h + buf * C * kKS, *y1l = Y1l + buf * C * kKS, *y2h = Y2h + buf * C * kKS, *y2l = Y2l + buf * C * kKS;
        const int cb = (lane & 7) + ((lane >> 3) & 1) * 8, nb = ((lane >> 4) & 1) * 8;   // .trans B fragments (k = channel, n = column)
        const int cv = (lane & 7) + ((lane >> 4) & 1) * 8, jv = ((lane >> 3) & 1) * 8;   // plain B fragments (k = column, n = channel)
#pragma unroll
        for (int ks = 0; ks < kKT / 16; ++ks) {
            float s2[2][4], dp2[2][4];
#pragma unroll
            for (int n = 0; n < 2; ++n)
#pragma unroll
                for (int i = 0; i < 4; ++i) { s2[n][i] = 0.f; dp2[n][i] = 0.f; }
#pragma unroll
            for (int kc = 0; kc < CH; ++kc) {
                uint32_t bh[4], bl[4], ch[4], cl2[4];   // S and dP together: four independent accumulators per product
                const int off = (kc * 16 + cb) * kKS + ks * 16 + nb;
                ldsm_x4_t(smem_u32(y1h + off), bh);
                ldsm_x4_t(smem_u32(y1l + off), bl);
                ldsm_x4_t(smem_u32(y2h + off), ch);
                ldsm_x4_t(smem_u32(y2l + off), cl2);
                mma16816(s2[0], a1l[kc], bh[0], bh[1]); mma16816(s2[1], a1l[kc], bh[2], bh[3]);
                mma16816(dp2[0], a2l[kc], ch[0], ch[1]); mma16816(dp2[1], a2l[kc], ch[2], ch[3]);
                mma16816(s2[0], a1h[kc], bl[0], bl[1]); mma16816(s2[1], a1h[kc], bl[2], bl[3]);
                mma16816(dp2[0], a2h[kc], cl2[0], cl2[1]); mma16816(dp2[1], a2h[kc], cl2[2], cl2[3]);
                mma16816(s2[0], a1h[kc], bh[0], bh[1]); mma16816(s2[1], a1h[kc], bh[2], bh[3]);
                mma16816(dp2[0], a2h[kc], ch[0], ch[1]); mma16816(dp2[1], a2h[kc], ch[2], ch[3]);
            }
            // P = exp2(S fs - lse2), D~ = P (dP~ - delta~); columns past the row end carry nothing
            float pr[2][4], ds[2][4];
#pragma unroll
            for (int n = 0; n < 2; ++n) {
                const int jl = ks * 16 + n * 8 + 2 * tq;        // column inside the tile
                const int jg = t * kKT + jl;
                const bool in0 = jg < W, in1 = jg + 1 < W;
                float l0c, l1c, d0c, d1c;
                if (KEYMAJOR) { l0c = s_l2[buf][jl]; l1c = s_l2[buf][jl + 1]; d0c = s_dl[buf][jl]; d1c = s_dl[buf][jl + 1]; }
                pr[n][0] = in0 ? fast_exp2(fmaf(s2[n][0], fs, -(KEYMAJOR ? l0c : rl0))) : 0.f;
                pr[n][1] = in1 ? fast_exp2(fmaf(s2[n][1], fs, -(KEYMAJOR ? l1c : rl0))) : 0.f;
                pr[n][2] = in0 ? fast_exp2(fmaf(s2[n][2], fs, -(KEYMAJOR ? l0c : rl1))) : 0.f;
                pr[n][3] = in1 ? fast_exp2(fmaf(s2[n][3], fs, -(KEYMAJOR ? l1c : rl1))) : 0.f;
                ds[n][0] = pr[n][0] * kUpD * fmaf(dp2[n][0], dn, -(KEYMAJOR ? d0c : rd0));
                ds[n][1] = pr[n][1] * kUpD * fmaf(dp2[n][1], dn, -(KEYMAJOR ? d1c : rd0));
                ds[n][2] = pr[n][2] * kUpD * fmaf(dp2[n][2], dn, -(KEYMAJOR ? d0c : rd1));
                ds[n][3] = pr[n][3] * kUpD * fmaf(dp2[n][3], dn, -(KEYMAJOR ? d1c : rd1));
            }
            uint32_t dh[4], dl[4];
            frag_from_acc(ds[0], ds[1], dh, dl);
            // acc1 += D~ Y1~^T (dQ += dS K | dK += dS^T Q); KEYMAJOR: acc2 += P^T dO~ (dV)
            {
                uint32_t bh[NO / 2][4], bl[NO / 2][4];
#pragma unroll
                for (int np = 0; np < NO / 2; ++np) {
                    const int off = (np * 16 + cv) * kKS + ks * 16 + jv;
                    ldsm_x4(smem_u32(y1h + off), bh[np]);
                    ldsm_x4(smem_u32(y1l + off), bl[np]);
                }
#pragma unroll
                for (int np = 0; np < NO / 2; ++np) { mma16816(acc1[2 * np], dl, bh[np][0], bh[np][1]); mma16816(acc1[2 * np + 1], dl, bh[np][2], bh[np][3]); }
#pragma unroll
                for (int np = 0; np < NO / 2; ++np) { mma16816(acc1[2 * np], dh, bl[np][0], bl[np][1]); mma16816(acc1[2 * np + 1], dh, bl[np][2], bl[np][3]); }
#pragma unroll
                for (int np = 0; np < NO / 2; ++np) { mma16816(acc1[2 * np], dh, bh[np][0], bh[np][1]); mma16816(acc1[2 * np + 1], dh, bh[np][2], bh[np][3]); }
            }
            if constexpr (KEYMAJOR) {
                uint32_t ph[4], pl[4];
#pragma unroll
                for (int n = 0; n < 2; ++n)
#pragma unroll
                    for (int i = 0; i < 4; ++i) pr[n][i] *= kUpP;
                frag_from_acc(pr[0], pr[1], ph, pl);
                uint32_t bh[NO / 2][4], bl[NO / 2][4];
#pragma unroll
                for (int np = 0; np < NO / 2; ++np) {
                    const int off = (np * 16 + cv) * kKS + ks * 16 + jv;
                    ldsm_x4(smem_u32(y2h + off), bh[np]);
                    ldsm_x4(smem_u32(y2l + off), bl[np]);
                }
#pragma unroll
                for (int np = 0; np < NO / 2; ++np) { mma16816(acc2[2 * np], pl, bh[np][0], bh[np][1]); mma16816(acc2[2 * np + 1], pl, bh[np][2], bh[np][3]); }
#pragma unroll
                for (int np = 0; np < NO / 2; ++np) { mma16816(acc2[2 * np], ph, bl[np][0], bl[np][1]); mma16816(acc2[2 * np + 1], ph, bl[np][2], bl[np][3]); }
#pragma unroll
                for (int np = 0; np < NO / 2; ++np) { mma16816(acc2[2 * np], ph, bh[np][0], bh[np][1]); mma16816(acc2[2 * np + 1], ph, bh[np][2], bh[np][3]); }
            }
        }
        if (more) {   // buffer buf ^ 1 was last read two tiles ago, before the barrier that ended the previous iteration
            yr1.template store<kKS>(Y1h + (buf ^ 1) * C * kKS, Y1l + (buf ^ 1) * C * kKS, iy1, tid);
            yr2.template store<kKS>(Y2h + (buf ^ 1) * C * kKS, Y2l + (buf ^ 1) * C * kKS, iy2, tid);
            if (KEYMAJOR && tid < kKT) { s_l2[buf ^ 1][tid] = cl; s_dl[buf ^ 1][tid] = cd; }
        }
        __syncthreads();
    }

    // ---- epilogue: constants of the row ----------------------------------------------------------------------------------------
    const float f1 = scale * sg * sv * (KEYMAJOR ? sq : sk) * (1.0f / (kUpD * kUp)), f2 = sg * (1.0f / (kUpP * kUp));
    float *o1 = g1 + rowoff, *o2 = KEYMAJOR ? g2 + rowoff : nullptr;
#pragma unroll
    for (int n = 0; n < NO; ++n) {
        const int c = 8 * n + 2 * tq;
        if (xa < W) { o1[(size_t)c * HW + xa] = acc1[n][0] * f1; o1[(size_t)(c + 1) * HW + xa] = acc1[n][1] * f1; }
        if (xb < W) { o1[(size_t)c * HW + xb] = acc1[n][2] * f1; o1[(size_t)(c + 1) * HW + xb] = acc1[n][3] * f1; }
        if constexpr (KEYMAJOR) {
            if (xa < W) { o2[(size_t)c * HW + xa] = acc2[n][0] * f2; o2[(size_t)(c + 1) * HW + xa] = acc2[n][1] * f2; }
            if (xb < W) { o2[(size_t)c * HW + xb] = acc2[n][2] * f2; o2[(size_t)(c + 1) * HW + xb] = acc2[n][3] * f2; }
        }
    }
}

int env_int_attn(const char *name, int dflt) {
    if (!tuning_enabled()) return dflt;
    const char *v = std::getenv(name);
    return (v && *v) ? std::atoi(v) : dflt;
}

template <int CH>
int launch_attn_mma(const float *q, const float *k, const float *v, float *out, float *lse, int B, int H, int W, float scale, cudaStream_t s) {
    constexpr int C = 16 * CH;
    auto kern = epipolar_attn_fwd_mma_kernel<CH>;
    const size_t smem = (size_t)(2 * C * kQS + 8 * C * kKS) * sizeof(__half);
    const cudaError_t ea = kernel_smem_optin(kern, smem);
    if (ea != cudaSuccess) return fail(DSSMD_ERR_CUDA, "epipolar_attn_fwd(mma): cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(ea));
    if (env_int_attn("DSSMD_CORR_DEBUG", 0))
        fprintf(stderr, "epipolar_attn_fwd mma cfg: C=%d grid=(%d,%d) threads=%d smem=%zu\n", C, ceil_div(W, kQT), B * H, kAThreads, smem);
    const cudaError_t e = launch_pdl(kern, dim3((unsigned)ceil_div(W, kQT), (unsigned)(B * H)), dim3(kAThreads), smem, s, q, k, v, out, lse, H, W, scale);
    if (e != cudaSuccess) return fail(DSSMD_ERR_CUDA, "epipolar_attn_fwd(mma): launch failed: %s", cudaGetErrorString(e));
    return check_launch("epipolar_attn_fwd(mma)");
}

template <int CH>
int launch_attn_bwd_mma(const float *q, const float *k, const float *v, const float *out, const float *g_out, const float *lse, float *delta,
                        float *stats, float *gq, float *gk, float *gv, int B, int H, int W, float scale, cudaStream_t s) {
    constexpr int C = 16 * CH;
    {
        const cudaError_t e = launch_pdl(attn_bwd_prep_kernel, dim3((unsigned)(B * H)), dim3(256), 0, s, q, k, v, g_out, out, delta, stats, C, H, W);
        if (e != cudaSuccess) return fail(DSSMD_ERR_CUDA, "epipolar_attn_bwd(mma, prep): launch failed: %s", cudaGetErrorString(e));
    }
    const size_t smem = (size_t)(4 * C * kQS + 8 * C * kKS) * sizeof(__half);
    const dim3 grid((unsigned)ceil_div(W, kQT), (unsigned)(B * H));
    if (env_int_attn("DSSMD_CORR_DEBUG", 0))
        fprintf(stderr, "epipolar_attn_bwd mma cfg: C=%d grid=(%u,%u) threads=%d smem=%zu dq=%d dkv=%d\n", C, grid.x, grid.y, kAThreads, smem, gq != nullptr, gk != nullptr);
    if (gq) {
        auto kern = epipolar_attn_bwd_mma_kernel<CH, false>;
        const cudaError_t ea = kernel_smem_optin(kern, smem);
        if (ea != cudaSuccess) return fail(DSSMD_ERR_CUDA, "epipolar_attn_bwd(mma, dQ): cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(ea));
        const cudaError_t e = launch_pdl(kern, grid, dim3(kAThreads), smem, s, q, k, v, g_out, lse, (const float *)delta, (const float *)stats, gq, (float *)nullptr, H, W, scale);
        if (e != cudaSuccess) return fail(DSSMD_ERR_CUDA, "epipolar_attn_bwd(mma, dQ): launch failed: %s", cudaGetErrorString(e));
    }
    if (gk && gv) {
        auto kern = epipolar_attn_bwd_mma_kernel<CH, true>;
        const cudaError_t ea = kernel_smem_optin(kern, smem);
        if (ea != cudaSuccess) return fail(DSSMD_ERR_CUDA, "epipolar_attn_bwd(mma, dK dV): cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(ea));
        const cudaError_t e = launch_pdl(kern, grid, dim3(kAThreads), smem, s, q, k, v, g_out, lse, (const float *)delta, (const float *)stats, gk, gv, H, W, scale);
        if (e != cudaSuccess) return fail(DSSMD_ERR_CUDA, "epipolar_attn_bwd(mma, dK dV): launch failed: %s", cudaGetErrorString(e));
    }
    return check_launch("epipolar_attn_bwd(mma)");
}

}  // namespace

// fp32 tensors, C in {16, 32}: -1 = not eligible (the caller takes the FFMA kernel), else a DSSMD_* code.  DSSMD_ATTN_FWD_MMA=0 disables it.
int epipolar_attn_fwd_mma(const float *q, const float *k, const float *v, float *out, float *lse, int B, int C, int H, int W, float scale,
                          cudaStream_t s) {
    if (env_int_attn("DSSMD_ATTN_FWD_MMA", 1) == 0) return -1;
    if (C == 16) return launch_attn_mma<1>(q, k, v, out, lse, B, H, W, scale, s);
    if (C == 32) return launch_attn_mma<2>(q, k, v, out, lse, B, H, W, scale, s);
    return -1;
}

// The backward for the same shapes; delta: B*H*W floats followed by 4*B*H floats of per-row scales (both written here).
// DSSMD_ATTN_BWD_MMA=0 disables it.
int epipolar_attn_bwd_mma(const float *q, const float *k, const float *v, const float *out, const float *g_out, const float *lse, float *delta,
                          float *gq, float *gk, float *gv, int B, int C, int H, int W, float scale, cudaStream_t s) {
    if (env_int_attn("DSSMD_ATTN_BWD_MMA", 1) == 0) return -1;
    float *stats = delta + (size_t)B * H * W;
    if (C == 16) return launch_attn_bwd_mma<1>(q, k, v, out, g_out, lse, delta, stats, gq, gk, gv, B, H, W, scale, s);
    if (C == 32) return launch_attn_bwd_mma<2>(q, k, v, out, g_out, lse, delta, stats, gq, gk, gv, B, H, W, scale, s);
    return -1;
}

}  // namespace dssmd
