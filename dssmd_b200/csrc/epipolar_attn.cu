// epipolar_attn.cu -- attention along the epipolar line (one image row) between left and right features, forward and
// backward, without ever materialising the W x W attention matrix.
// Upstream: the core of STTM_Single.forward, models/BidNet/STTM_Simple.py:43-57 (used at full resolution by
// models/BidNet/BidNet.py:115, 239-240):
//     q, k, v: [B,C,H,W] -> permute(0,2,3,1).reshape(B*H, W, C).contiguous()          (three copies)
//     dots = matmul(q, k^T) * scale;  attn = softmax(dots, -1);  out = matmul(attn, v)   ([B*H, W, W] written twice, read twice)
//     out.view(B,H,W,C).permute(0,3,1,2)                                               (a strided view; `cat` copies it)
// At BidNet's shape (B8, C32, 320x640) the two [2560, 640, 640] fp32 intermediates are 4.2 GB each.  Here a block owns
// 64 queries of one row, reads the C x 64 tiles of q / k / v straight out of the NCHW tensors (a channel's 64 pixels are
// contiguous), keeps the running row maximum and sum of the softmax in registers (the usual online rescaling), and writes
// out[B,C,H,W] in NCHW plus the row log-sum-exp the backward needs.  fp32 FFMA throughout (upstream's matmuls run with
// torch.backends.cuda.matmul.allow_tf32 = False, and the parity bar is 1e-5).
//
// Backward: with lse and delta[x] = sum_c dO[c,x] O[c,x] the probabilities are recomputed tile by tile:
//     P = exp(scale * Q^T K - lse[x]);  dP = dO^T V;  dS = P o (dP - delta[x])
//     dQ[c,x] = scale * sum_x' dS[x,x'] K[c,x'];   dK[c,x'] = scale * sum_x dS[x,x'] Q[c,x];   dV[c,x'] = sum_x P[x,x'] dO[c,x]
// One kernel template does both halves: a block owns 64 "row" pixels and walks over the 64-pixel "column" tiles -- rows =
// queries for dQ; rows = keys (the transposed problem: S^T = K^T Q, so the tiles it needs transposed come out that way by
// swapping the operands) for dK and dV.  No atomics, fixed summation order: deterministic.
#include "common.cuh"

namespace dssmd {

namespace {

constexpr int kT = 64;        // tile edge in pixels
constexpr int kLD = 68;       // row stride of a staged tile in floats: 16-byte aligned rows, conflict-free LDS.128 over channels
constexpr int kThreads = 256; // 16 x 16 threads, each a 4 x 4 block of the 64 x 64 score tile

// C x 64 tile of a [B,C,H,W] tensor (row `rowbase` = &t[b,0,y,0], channel stride HW) -> shared [C][kLD], zero past W
template <int CH>
__device__ __forceinline__ void load_tile(float *dst, const float *rowbase, size_t HW, int x0, int W, int tid) {
#pragma unroll 2
    for (int idx = tid; idx < 16 * CH * 16; idx += kThreads) {
        const int c = idx >> 4, f = idx & 15;
        const int x = x0 + 4 * f;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (x < W) v = ld_cg(reinterpret_cast<const float4 *>(rowbase + (size_t)c * HW + x));
        *reinterpret_cast<float4 *>(dst + c * kLD + 4 * f) = v;
    }
}

// s[i][j] = sum_c X[c][4 ty + i] * Y[c][4 tx + j]
template <int CH>
__device__ __forceinline__ void gemm_cc(const float *X, const float *Y, int ty, int tx, float (&s)[4][4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll 8
    for (int c = 0; c < 16 * CH; ++c) {
        const float4 a = *reinterpret_cast<const float4 *>(X + c * kLD + 4 * ty);
        const float4 b = *reinterpret_cast<const float4 *>(Y + c * kLD + 4 * tx);
        const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) s[i][j] = fmaf(av[i], bv[j], s[i][j]);
    }
}

// acc[i][k] += sum_j P[4 ty + i][j] * V[tx + 16 k][j]      (P: [64][kLD], V: [C][kLD])
template <int CH>
__device__ __forceinline__ void gemm_pj(const float *P, const float *V, int ty, int tx, float (&acc)[4][CH]) {
#pragma unroll 4
    for (int j4 = 0; j4 < kT / 4; ++j4) {
        float4 p[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) p[i] = *reinterpret_cast<const float4 *>(P + (4 * ty + i) * kLD + 4 * j4);
#pragma unroll
        for (int k = 0; k < CH; ++k) {
            const float4 v = *reinterpret_cast<const float4 *>(V + (tx + 16 * k) * kLD + 4 * j4);
#pragma unroll
            for (int i = 0; i < 4; ++i)
                acc[i][k] = fmaf(p[i].w, v.w, fmaf(p[i].z, v.z, fmaf(p[i].y, v.y, fmaf(p[i].x, v.x, acc[i][k]))));
        }
    }
}

// acc tile (rows 4 ty + i = pixels, channels tx + 16 k) * mul -> t[b, c, y, x0 + 4 ty ..]
template <int CH>
__device__ __forceinline__ void store_tile(float *rowbase, size_t HW, int x0, int W, int ty, int tx, const float (&acc)[4][CH], const float (&mul)[4]) {
    const int x = x0 + 4 * ty;
    if (x >= W) return;
#pragma unroll
    for (int k = 0; k < CH; ++k)
        *reinterpret_cast<float4 *>(rowbase + (size_t)(tx + 16 * k) * HW + x) =
            make_float4(acc[0][k] * mul[0], acc[1][k] * mul[1], acc[2][k] * mul[2], acc[3][k] * mul[3]);
}

// max / sum over the 16 lanes that share ty (lanes differ in tx = lane & 15)
__device__ __forceinline__ float row_max16(float v) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float row_sum16(float v) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- forward: grid (query tiles, B*H) ---------------------------------------------------------------------------------
template <int CH>
__global__ void __launch_bounds__(kThreads) epipolar_attn_fwd_kernel(const float *q, const float *k,
                                                                     const float *v, float *__restrict__ out,
                                                                     float *__restrict__ lse, int H, int W, float scale) {
    constexpr int C = 16 * CH;
    extern __shared__ __align__(16) float smem[];
    float *Qs = smem, *Ks = Qs + C * kLD, *Vs = Ks + C * kLD, *Ps = Vs + C * kLD;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int r = blockIdx.y, b = r / H, y = r - b * H;
    const int x0 = blockIdx.x * kT;
    const size_t HW = (size_t)H * W;
    const size_t rowoff = (size_t)b * C * HW + (size_t)y * W;
    pdl_launch_dependents();
    pdl_wait();
    load_tile<CH>(Qs, q + rowoff, HW, x0, W, tid);
    float m[4], l[4], o[4][CH];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        m[i] = -INFINITY; l[i] = 0.f;
#pragma unroll
        for (int kk = 0; kk < CH; ++kk) o[i][kk] = 0.f;
    }
    for (int j0 = 0; j0 < W; j0 += kT) {
        __syncthreads();   // the previous tile's K / V / P are no longer read
        load_tile<CH>(Ks, k + rowoff, HW, j0, W, tid);
        load_tile<CH>(Vs, v + rowoff, HW, j0, W, tid);
        __syncthreads();
        float s[4][4];
        gemm_cc<CH>(Qs, Ks, ty, tx, s);
        const int jb = j0 + 4 * tx;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                s[i][j] = (jb + j < W) ? s[i][j] * scale : -INFINITY;
                mx = fmaxf(mx, s[i][j]);
            }
            const float mn = fmaxf(m[i], row_max16(mx));   // finite: every tile holds at least one key
            const float corr = expf(m[i] - mn);            // exp(-inf) = 0 on the first tile
            float ps = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) { s[i][j] = expf(s[i][j] - mn); ps += s[i][j]; }
            l[i] = fmaf(l[i], corr, row_sum16(ps));
            m[i] = mn;
#pragma unroll
            for (int kk = 0; kk < CH; ++kk) o[i][kk] *= corr;
            *reinterpret_cast<float4 *>(Ps + (4 * ty + i) * kLD + 4 * tx) = make_float4(s[i][0], s[i][1], s[i][2], s[i][3]);
        }
        __syncthreads();
        gemm_pj<CH>(Ps, Vs, ty, tx, o);
    }
    float inv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) inv[i] = 1.0f / l[i];
    store_tile<CH>(out + rowoff, HW, x0, W, ty, tx, o, inv);
    if (lse && tx == 0 && x0 + 4 * ty < W)
        *reinterpret_cast<float4 *>(lse + (size_t)r * W + x0 + 4 * ty) =
            make_float4(m[0] + logf(l[0]), m[1] + logf(l[1]), m[2] + logf(l[2]), m[3] + logf(l[3]));
}

// ---- delta[r, x] = sum_c dO[b,c,y,x] * O[b,c,y,x] -------------------------------------------------------------------------
__global__ void __launch_bounds__(256) epipolar_attn_delta_kernel(const float *g_out, const float *out,
                                                                  float *__restrict__ delta, int C, int H, int W, long long n) {
    pdl_launch_dependents();
    pdl_wait();
    const size_t HW = (size_t)H * W;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / (long long)HW, p = i - b * (long long)HW;
        const float *a = g_out + (size_t)b * C * HW + p, *o = out + (size_t)b * C * HW + p;
        float s = 0.f;
        for (int c = 0; c < C; ++c) s = fmaf(ld_cg(a + (size_t)c * HW), ld_cg(o + (size_t)c * HW), s);
        delta[i] = s;   // [B*H, W] is the same linear index as [B, H, W]
    }
}

// ---- backward: grid (row tiles, B*H).  KEYMAJOR = false: rows = queries, writes dQ.  true: rows = keys, writes dK, dV ----
template <int CH, bool KEYMAJOR>
__global__ void __launch_bounds__(kThreads) epipolar_attn_bwd_kernel(const float *q, const float *k,
                                                                     const float *v, const float *g_out,
                                                                     const float *lse, const float *delta,
                                                                     float *__restrict__ g1, float *__restrict__ g2, int H, int W, float scale) {
    constexpr int C = 16 * CH;
    extern __shared__ __align__(16) float smem[];
    float *X1 = smem, *X2 = X1 + C * kLD, *Y1 = X2 + C * kLD, *Y2 = Y1 + C * kLD, *P1 = Y2 + C * kLD, *P2 = P1 + kT * kLD;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int r = blockIdx.y, b = r / H, y = r - b * H;
    const int x0 = blockIdx.x * kT;
    const size_t HW = (size_t)H * W;
    const size_t rowoff = (size_t)b * C * HW + (size_t)y * W;
    // row side: (Q, dO) for dQ; (K, V) for dK / dV.  Column side: the other pair.
    const float *x1 = (KEYMAJOR ? k : q) + rowoff, *x2 = (KEYMAJOR ? v : g_out) + rowoff;
    const float *y1 = (KEYMAJOR ? q : k) + rowoff, *y2 = (KEYMAJOR ? g_out : v) + rowoff;
    const float *lrow = lse + (size_t)r * W, *drow = delta + (size_t)r * W;
    pdl_launch_dependents();
    pdl_wait();
    load_tile<CH>(X1, x1, HW, x0, W, tid);
    load_tile<CH>(X2, x2, HW, x0, W, tid);
    float a1[4][CH], a2[KEYMAJOR ? 4 : 1][KEYMAJOR ? CH : 1];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int kk = 0; kk < CH; ++kk) {
            a1[i][kk] = 0.f;
            if constexpr (KEYMAJOR) a2[i][kk] = 0.f;
        }
    // softmax statistics belong to the QUERY pixel: the row (queries as rows) or the column (keys as rows)
    float rl[4], rd[4];
    if (!KEYMAJOR) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int x = x0 + 4 * ty + i;
            rl[i] = x < W ? ld_cg(lrow + x) : 0.f;
            rd[i] = x < W ? ld_cg(drow + x) : 0.f;
        }
    }
    for (int j0 = 0; j0 < W; j0 += kT) {
        __syncthreads();
        load_tile<CH>(Y1, y1, HW, j0, W, tid);
        load_tile<CH>(Y2, y2, HW, j0, W, tid);
        __syncthreads();
        float s[4][4], dp[4][4];
        gemm_cc<CH>(X1, Y1, ty, tx, s);     // S (or S^T)
        gemm_cc<CH>(X2, Y2, ty, tx, dp);    // dP (or dP^T)
        const int jb = j0 + 4 * tx;
        float cl[4], cd[4];
        if (KEYMAJOR) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                cl[j] = jb + j < W ? ld_cg(lrow + jb + j) : 0.f;
                cd[j] = jb + j < W ? ld_cg(drow + jb + j) : 0.f;
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float pr[4], ds[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float lj = KEYMAJOR ? cl[j] : rl[i], dj = KEYMAJOR ? cd[j] : rd[i];
                pr[j] = (jb + j < W) ? expf(fmaf(s[i][j], scale, -lj)) : 0.f;   // columns past the row end carry nothing
                ds[j] = pr[j] * (dp[i][j] - dj);
            }
            *reinterpret_cast<float4 *>(P1 + (4 * ty + i) * kLD + 4 * tx) = make_float4(ds[0], ds[1], ds[2], ds[3]);
            if (KEYMAJOR) *reinterpret_cast<float4 *>(P2 + (4 * ty + i) * kLD + 4 * tx) = make_float4(pr[0], pr[1], pr[2], pr[3]);
        }
        __syncthreads();
        gemm_pj<CH>(P1, Y1, ty, tx, a1);                  // dQ += dS K        | dK += dS^T Q
        if constexpr (KEYMAJOR) gemm_pj<CH>(P2, Y2, ty, tx, a2);   //         | dV += P^T dO
    }
    const float sc[4] = {scale, scale, scale, scale};
    store_tile<CH>(g1 + rowoff, HW, x0, W, ty, tx, a1, sc);
    if constexpr (KEYMAJOR) {
        const float one[4] = {1.f, 1.f, 1.f, 1.f};
        store_tile<CH>(g2 + rowoff, HW, x0, W, ty, tx, a2, one);
    }
}

template <typename K>
int set_smem(K kern, size_t smem, const char *what) {
    if (smem <= 48 * 1024) return DSSMD_OK;
    const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(DSSMD_ERR_CUDA, "%s: cudaFuncSetAttribute(%zu B): %s", what, smem, cudaGetErrorString(e));
    return DSSMD_OK;
}

template <int CH>
int fwd_ch(const float *q, const float *k, const float *v, float *out, float *lse, int B, int H, int W, float scale, cudaStream_t s) {
    auto kern = epipolar_attn_fwd_kernel<CH>;
    const size_t smem = (size_t)(3 * 16 * CH + kT) * kLD * sizeof(float);
    if (const int rc = set_smem(kern, smem, "epipolar_attn_fwd")) return rc;
    const cudaError_t e = launch_pdl(kern, dim3((unsigned)ceil_div(W, kT), (unsigned)(B * H)), dim3(kThreads), smem, s, q, k, v, out, lse, H, W, scale);
    if (e != cudaSuccess) return fail(DSSMD_ERR_CUDA, "epipolar_attn_fwd: launch failed: %s", cudaGetErrorString(e));
    return check_launch("epipolar_attn_fwd");
}

template <int CH>
int bwd_ch(const float *q, const float *k, const float *v, const float *g_out, const float *lse, const float *delta,
           float *gq, float *gk, float *gv, int B, int H, int W, float scale, cudaStream_t s) {
    const dim3 grid((unsigned)ceil_div(W, kT), (unsigned)(B * H));
    if (gq) {
        auto kern = epipolar_attn_bwd_kernel<CH, false>;
        const size_t smem = (size_t)(4 * 16 * CH + kT) * kLD * sizeof(float);
        if (const int rc = set_smem(kern, smem, "epipolar_attn_bwd(dQ)")) return rc;
        const cudaError_t e = launch_pdl(kern, grid, dim3(kThreads), smem, s, q, k, v, g_out, lse, delta, gq, (float *)nullptr, H, W, scale);
        if (e != cudaSuccess) return fail(DSSMD_ERR_CUDA, "epipolar_attn_bwd(dQ): launch failed: %s", cudaGetErrorString(e));
    }
    if (gk && gv) {
        auto kern = epipolar_attn_bwd_kernel<CH, true>;
        const size_t smem = (size_t)(4 * 16 * CH + 2 * kT) * kLD * sizeof(float);
        if (const int rc = set_smem(kern, smem, "epipolar_attn_bwd(dK,dV)")) return rc;
        const cudaError_t e = launch_pdl(kern, grid, dim3(kThreads), smem, s, q, k, v, g_out, lse, delta, gk, gv, H, W, scale);
        if (e != cudaSuccess) return fail(DSSMD_ERR_CUDA, "epipolar_attn_bwd(dK,dV): launch failed: %s", cudaGetErrorString(e));
    }
    return check_launch("epipolar_attn_bwd");
}

int check_shape(const char *what, int B, int C, int H, int W, int dtype) {
    DSSMD_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, "%s: sizes must be positive (B=%d C=%d H=%d W=%d)", what, B, C, H, W);
    DSSMD_REQUIRE((long long)B * H <= 65535, "%s: B*H = %lld rows exceed the grid limit 65535", what, (long long)B * H);
    if (dtype != DSSMD_F32) return fail(DSSMD_ERR_UNSUPPORTED, "%s: only fp32 is implemented (dtype %d)", what, dtype);
    if (!(C == 16 || C == 32 || C == 64 || C == 128) || W % 4 != 0)
        return fail(DSSMD_ERR_UNSUPPORTED, "%s: needs C in {16, 32, 64, 128} and W %% 4 == 0 (got C=%d W=%d)", what, C, W);
    return DSSMD_OK;
}

}  // namespace

}  // namespace dssmd

using namespace dssmd;

namespace dssmd {
// epipolar_attn_mma.cu: the forward on mma.sync (fp16 hi / lo operand splitting); -1 = not eligible
int epipolar_attn_fwd_mma(const float *q, const float *k, const float *v, float *out, float *lse, int B, int C, int H, int W, float scale,
                          cudaStream_t s);
// epipolar_attn_umma.cu: the forward on tcgen05 + TMEM (C = 32); -1 = not eligible
int epipolar_attn_fwd_umma(const float *q, const float *k, const float *v, float *out, float *lse, int B, int C, int H, int W, float scale,
                           cudaStream_t s);
int epipolar_attn_bwd_mma(const float *q, const float *k, const float *v, const float *out, const float *g_out, const float *lse, float *delta,
                          float *gq, float *gk, float *gv, int B, int C, int H, int W, float scale, cudaStream_t s);
}  // namespace dssmd

extern "C" int dssmd_epipolar_attn_fwd(const void *q, const void *k, const void *v, void *out, void *lse, int B, int C, int H, int W,
                                       float scale, int dtype, void *stream) {
    DSSMD_REQUIRE(q && k && v && out, "epipolar_attn_fwd: null pointer");
    if (const int rc = check_shape("epipolar_attn_fwd", B, C, H, W, dtype)) return rc;
    DSSMD_REQUIRE(((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(out) |
                    reinterpret_cast<uintptr_t>(lse)) & 15) == 0, "epipolar_attn_fwd: tensors must be 16-byte aligned");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const float *Q = (const float *)q, *K = (const float *)k, *V = (const float *)v;
    {
        int rc = epipolar_attn_fwd_umma(Q, K, V, (float *)out, (float *)lse, B, C, H, W, scale, s);
        if (rc >= 0) return rc;
        rc = epipolar_attn_fwd_mma(Q, K, V, (float *)out, (float *)lse, B, C, H, W, scale, s);
        if (rc >= 0) return rc;
    }
    switch (C) {
        case 16: return fwd_ch<1>(Q, K, V, (float *)out, (float *)lse, B, H, W, scale, s);
        case 32: return fwd_ch<2>(Q, K, V, (float *)out, (float *)lse, B, H, W, scale, s);
        case 64: return fwd_ch<4>(Q, K, V, (float *)out, (float *)lse, B, H, W, scale, s);
        default: return fwd_ch<8>(Q, K, V, (float *)out, (float *)lse, B, H, W, scale, s);
    }
}

extern "C" int dssmd_epipolar_attn_bwd(const void *q, const void *k, const void *v, const void *out, const void *g_out, const void *lse,
                                       void *delta, void *g_q, void *g_k, void *g_v, int B, int C, int H, int W, float scale, int dtype,
                                       void *stream) {
    DSSMD_REQUIRE(q && k && v && out && g_out && lse && delta, "epipolar_attn_bwd: null pointer");
    DSSMD_REQUIRE((g_k == nullptr) == (g_v == nullptr), "epipolar_attn_bwd: g_k and g_v come from one kernel: pass both or neither");
    if (const int rc = check_shape("epipolar_attn_bwd", B, C, H, W, dtype)) return rc;
    DSSMD_REQUIRE(((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(g_out) |
                    reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(delta) |
                    reinterpret_cast<uintptr_t>(g_q) | reinterpret_cast<uintptr_t>(g_k) | reinterpret_cast<uintptr_t>(g_v)) & 15) == 0,
                  "epipolar_attn_bwd: tensors must be 16-byte aligned");
    if (!g_q && !g_k) return DSSMD_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    {
        const int rc = epipolar_attn_bwd_mma((const float *)q, (const float *)k, (const float *)v, (const float *)out, (const float *)g_out,
                                             (const float *)lse, (float *)delta, (float *)g_q, (float *)g_k, (float *)g_v, B, C, H, W, scale, s);
        if (rc >= 0) return rc;
    }
    {
        const long long n = (long long)B * H * W;
        long long nb = (n + 255) / 256;
        if (nb > (long long)sm_count() * 16) nb = (long long)sm_count() * 16;
        const cudaError_t e = launch_pdl(epipolar_attn_delta_kernel, dim3((unsigned)nb), dim3(256), 0, s, (const float *)g_out, (const float *)out,
                                         (float *)delta, C, H, W, n);
        if (e != cudaSuccess) return fail(DSSMD_ERR_CUDA, "epipolar_attn_bwd(delta): launch failed: %s", cudaGetErrorString(e));
    }
    const float *Q = (const float *)q, *K = (const float *)k, *V = (const float *)v, *G = (const float *)g_out;
    const float *L = (const float *)lse, *D = (const float *)delta;
    switch (C) {
        case 16: return bwd_ch<1>(Q, K, V, G, L, D, (float *)g_q, (float *)g_k, (float *)g_v, B, H, W, scale, s);
        case 32: return bwd_ch<2>(Q, K, V, G, L, D, (float *)g_q, (float *)g_k, (float *)g_v, B, H, W, scale, s);
        case 64: return bwd_ch<4>(Q, K, V, G, L, D, (float *)g_q, (float *)g_k, (float *)g_v, B, H, W, scale, s);
        default: return bwd_ch<8>(Q, K, V, G, L, D, (float *)g_q, (float *)g_k, (float *)g_v, B, H, W, scale, s);
    }
}
