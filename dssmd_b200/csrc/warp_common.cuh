// warp_common.cuh -- coordinate / bilinear helpers shared by the warp kernels.
// Every arithmetic op is an individually rounded intrinsic (never contracted), so the
// sampling coordinates reproduce the reference's op order bit for bit.
#pragma once

#include "common.cuh"

namespace dssmd {

// ---- scalar helpers: every op individually rounded, never contracted ----------
template <typename S> struct Num;
template <> struct Num<float> {
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
    static __device__ __forceinline__ float fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
    static __device__ __forceinline__ float floor(float a) { return floorf(a); }
};
template <> struct Num<double> {
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
    static __device__ __forceinline__ double fma(double a, double b, double c) { return __fma_rn(a, b, c); }
    static __device__ __forceinline__ double floor(double a) { return ::floor(a); }
};

// pixel coordinate -> grid_sample source index (reference op order)
template <typename S>
__device__ __forceinline__ S unnorm(S pix, int size) {
    using N = Num<S>;
    const S g = N::sub(N::mul((S)2, N::div(pix, (S)(size - 1))), (S)1);
    return N::fma(N::add(g, (S)1), N::div((S)size, (S)2), (S)-0.5);
}

// One axis of a bilinear tap pair: integer index of the low tap, the two
// weights, and whether each tap is inside [0, size).
template <typename S>
struct Axis {
    int i0;
    S w0, w1;      // weights of taps i0 and i0+1
    bool in0, in1;
};

template <typename S>
__device__ __forceinline__ Axis<S> make_axis(S coord, int size) {
    using N = Num<S>;
    Axis<S> a;
    const S f = N::floor(coord);
    a.w1 = N::sub(coord, f);
    a.w0 = N::sub((S)1, a.w1);
    // NaN / huge coordinates: every comparison false -> both taps out of bounds
    const bool finite = (f >= (S)-2) && (f <= (S)size);
    a.i0 = finite ? (int)f : -2;
    a.in0 = finite && a.i0 >= 0 && a.i0 < size;
    a.in1 = finite && a.i0 + 1 >= 0 && a.i0 + 1 < size;
    return a;
}

template <typename S>
__device__ __forceinline__ S clamp_border(S c, int size) {
    // ATen clip_coordinates: min(size-1, max(c, 0)); NaN propagates like std::min/max do there
    return fmin((S)(size - 1), fmax(c, (S)0));
}

template <typename S>
struct Src {
    int i0, i1;
    S l0, l1;
};

template <typename S>
__device__ __forceinline__ Src<S> src_index(int dst, int in_size, int out_size) {
    using N = Num<S>;
    Src<S> r;
    if (in_size == out_size) { r.i0 = r.i1 = dst; r.l0 = (S)1; r.l1 = (S)0; return r; }
    const S scale = N::div((S)in_size, (S)out_size);
    S src = N::sub(N::mul(scale, N::add((S)dst, (S)0.5)), (S)0.5);
    if (src < (S)0) src = (S)0;
    int k = (int)src;
    if (k > in_size - 1) k = in_size - 1;
    S lam = N::sub(src, (S)k);
    lam = fmin(fmax(lam, (S)0), (S)1);
    r.i0 = k;
    r.i1 = k + (k < in_size - 1 ? 1 : 0);
    r.l1 = lam;
    r.l0 = N::sub((S)1, lam);
    return r;
}

template <typename S>
__device__ __forceinline__ S resize_at(const S *plane, int Hi, int Wi, const Src<S> &ys, const Src<S> &xs) {
    using N = Num<S>;
    const S *r0 = plane + (long long)ys.i0 * Wi, *r1 = plane + (long long)ys.i1 * Wi;
    const S top = N::add(N::mul(xs.l0, r0[xs.i0]), N::mul(xs.l1, r0[xs.i1]));
    const S bot = N::add(N::mul(xs.l0, r1[xs.i0]), N::mul(xs.l1, r1[xs.i1]));
    return N::add(N::mul(ys.l0, top), N::mul(ys.l1, bot));
}

inline unsigned grid_for(long long n, int threads, int per_sm) {
    long long nb = (n + threads - 1) / threads;
    const long long cap = (long long)sm_count() * per_sm;
    if (nb > cap) nb = cap;
    if (nb < 1) nb = 1;
    return (unsigned)nb;
}

}  // namespace dssmd
