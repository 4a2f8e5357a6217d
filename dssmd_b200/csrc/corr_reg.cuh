// corr_reg.cuh -- pieces shared by the register-blocked, TMA-staged correlation kernels
// (corr_fwd_reg.cu, corr_bwd_reg.cu): element access for fp32 / 16-bit tensors and the lane map that
// keeps the sliding-window loads free of bank conflicts.
#pragma once

#include "common.cuh"
#include "tma.cuh"

namespace dssmd {

namespace {

constexpr int kDP = 12;  // disparities per lane

// ---- element types: fp32, or 16-bit storage with fp32 arithmetic ------------------------------
__device__ __forceinline__ void unpack2(const __nv_bfloat16 *, uint32_t u, float &lo, float &hi) {
    lo = __uint_as_float(u << 16);
    hi = __uint_as_float(u & 0xffff0000u);
}
__device__ __forceinline__ void unpack2(const __half *, uint32_t u, float &lo, float &hi) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&u));
    lo = f.x; hi = f.y;
}
__device__ __forceinline__ uint32_t pack2(const __nv_bfloat16 *, float lo, float hi) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t *>(&v);
}
__device__ __forceinline__ uint32_t pack2(const __half *, float lo, float hi) {
    const __half2 v = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t *>(&v);
}
// 4 consecutive elements at a 4-element aligned shared-memory address
template <typename T>
__device__ __forceinline__ void load4(const T *p, float &a, float &b, float &c, float &d) {
    if constexpr (sizeof(T) == 4) {
        const float4 v = *reinterpret_cast<const float4 *>(p);
        a = v.x; b = v.y; c = v.z; d = v.w;
    } else {
        const uint2 v = *reinterpret_cast<const uint2 *>(p);
        unpack2(p, v.x, a, b);
        unpack2(p, v.y, c, d);
    }
}
// N (1, 2 or 4) finished pixels -> global memory at an N-element aligned address, streaming
template <typename T, int N>
__device__ __forceinline__ void store_n(T *o, const float *v) {
    if constexpr (sizeof(T) == 4) {
        if constexpr (N == 4) st_global_cs(reinterpret_cast<float4 *>(o), make_float4(v[0], v[1], v[2], v[3]));
        else if constexpr (N == 2) asm volatile("st.global.cs.v2.f32 [%0], {%1, %2};" ::"l"(o), "f"(v[0]), "f"(v[1]) : "memory");
        else asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(o), "f"(v[0]) : "memory");
    } else {
        if constexpr (N == 4) {
            const uint32_t a = pack2(o, v[0], v[1]), b = pack2(o, v[2], v[3]);
            asm volatile("st.global.cs.v2.b32 [%0], {%1, %2};" ::"l"(o), "r"(a), "r"(b) : "memory");
        } else if constexpr (N == 2) {
            const uint32_t a = pack2(o, v[0], v[1]);
            asm volatile("st.global.cs.b32 [%0], %1;" ::"l"(o), "r"(a) : "memory");
        } else {
            *o = from_f32<T>(v[0]);
        }
    }
}

// Lane -> (pixel group pgl, disparity part h).  A lane's window starts at 16-byte chunk
// (PX/4)*pgl -+ 3*h of the staged row, and an LDS.128 is served a quarter-warp (8 consecutive lanes) at a
// time: the map puts the lane bits so that the 8 chunks of every quarter-warp are distinct modulo 8, i.e. the
// window loads are free of bank conflicts.  kHPos[k] is the lane bit that carries bit k of h.
template <int LOGNH, int PX> struct LaneMap;
template <int PX> struct LaneMap<0, PX> {
    static __device__ __forceinline__ int h(int) { return 0; }
    static __device__ __forceinline__ int pgl(int lane) { return lane; }
    static __device__ __forceinline__ constexpr int hpos(int) { return 0; }
};
template <> struct LaneMap<1, 4> {  // chunks pgl - 3h: 8 consecutive pgl per quarter-warp
    static __device__ __forceinline__ int h(int lane) { return lane >> 4; }
    static __device__ __forceinline__ int pgl(int lane) { return lane & 15; }
    static __device__ __forceinline__ constexpr int hpos(int) { return 4; }
};
template <> struct LaneMap<2, 4> {
    static __device__ __forceinline__ int h(int lane) { return lane >> 3; }
    static __device__ __forceinline__ int pgl(int lane) { return lane & 7; }
    static __device__ __forceinline__ constexpr int hpos(int k) { return 3 + k; }
};
template <> struct LaneMap<3, 4> {  // quarter-warp: pgl 0..3 and h2 (-12 chunks = 4 mod 8)
    static __device__ __forceinline__ int h(int lane) { return ((lane >> 3) & 3) | (lane & 4); }
    static __device__ __forceinline__ int pgl(int lane) { return lane & 3; }
    static __device__ __forceinline__ constexpr int hpos(int k) { return k == 2 ? 2 : 3 + k; }
};
template <> struct LaneMap<4, 4> {  // quarter-warp: pgl 0..1, h1 (-6 = 2), h2 (-12 = 4)
    static __device__ __forceinline__ int h(int lane) { return ((lane >> 3) & 1) | (lane & 6) | ((lane >> 1) & 8); }
    static __device__ __forceinline__ int pgl(int lane) { return lane & 1; }
    static __device__ __forceinline__ constexpr int hpos(int k) { return k == 0 ? 3 : k == 3 ? 4 : k; }
};
template <> struct LaneMap<1, 8> {  // chunks 2*pgl - 3h: quarter-warp = pgl 0..3 (even chunks) x h0 (odd shift)
    static __device__ __forceinline__ int h(int lane) { return (lane >> 2) & 1; }
    static __device__ __forceinline__ int pgl(int lane) { return (lane & 3) | ((lane >> 1) & 12); }
    static __device__ __forceinline__ constexpr int hpos(int) { return 2; }
};
template <> struct LaneMap<2, 8> {
    static __device__ __forceinline__ int h(int lane) { return ((lane >> 2) & 1) | ((lane >> 3) & 2); }
    static __device__ __forceinline__ int pgl(int lane) { return (lane & 3) | ((lane >> 1) & 4); }
    static __device__ __forceinline__ constexpr int hpos(int k) { return k == 0 ? 2 : 4; }
};
template <> struct LaneMap<3, 8> {
    static __device__ __forceinline__ int h(int lane) { return lane >> 2; }
    static __device__ __forceinline__ int pgl(int lane) { return lane & 3; }
    static __device__ __forceinline__ constexpr int hpos(int k) { return 2 + k; }
};


}  // namespace

}  // namespace dssmd
