// common.cuh -- shared device helpers for the CNSN kernels (sm_100a).
//
// Layout everywhere: dense NCHW.  An "instance" is one (n,c) plane of M = H*W elements; it is
// contiguous in memory, so every kernel streams instances with 128-bit accesses whenever
// M*sizeof(T) is a multiple of 16 bytes and the base pointer is 16-byte aligned (VEC path),
// and falls back to element accesses otherwise (e.g. 7x7 fp32 planes: 196 B).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/cnsn_b200.h"

namespace cnsn {

constexpr int kWarp = 32;
constexpr int kBlock = 256;              // threads per CTA for the streaming kernels
constexpr int kWarpsPerBlock = kBlock / kWarp;

void note_launch();                      // api.cu: bumps the process-wide launch counter

// ---------------------------------------------------------------------------------------
// element <-> float conversion and 128-bit vector access
// ---------------------------------------------------------------------------------------
template <typename T> struct VecOf;      // elements per 16-byte vector
template <> struct VecOf<float> { static constexpr int n = 4; };
template <> struct VecOf<__nv_bfloat16> { static constexpr int n = 8; };
template <> struct VecOf<__half> { static constexpr int n = 8; };

__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ float to_f(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }

// Streaming 128-bit load: read-only path, do not allocate in L1 (each byte is used once per pass).
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
    uint4 r;
    asm("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
// Same, but mark the line evict-first in L2 as well (last use of the data).  sm_100a only accepts
// the inline .L2::evict_first qualifier on 256-bit loads, so 128-bit loads carry a cache-hint policy.
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_normal() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint4 ldg_hint(const void* p, uint64_t pol) {
    uint4 r;
    asm("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(pol));
    return r;
}
// Streaming 128-bit store (write-once output).
__device__ __forceinline__ void stg_stream(void* p, const uint4& v) {
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w));
}

template <typename T> __device__ __forceinline__ void unpack(const uint4& raw, float (&v)[VecOf<T>::n]);
template <> __device__ __forceinline__ void unpack<float>(const uint4& raw, float (&v)[4]) {
    v[0] = __uint_as_float(raw.x); v[1] = __uint_as_float(raw.y);
    v[2] = __uint_as_float(raw.z); v[3] = __uint_as_float(raw.w);
}
template <> __device__ __forceinline__ void unpack<__nv_bfloat16>(const uint4& raw, float (&v)[8]) {
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {           // bf16 -> fp32 is a 16-bit shift
        v[2 * i] = __uint_as_float(w[i] << 16);
        v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}
template <> __device__ __forceinline__ void unpack<__half>(const uint4& raw, float (&v)[8]) {
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
        v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
}
template <typename T> __device__ __forceinline__ uint4 pack(const float (&v)[VecOf<T>::n]);
template <> __device__ __forceinline__ uint4 pack<float>(const float (&v)[4]) {
    return make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
}
template <> __device__ __forceinline__ uint4 pack<__nv_bfloat16>(const float (&v)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        w[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}
template <> __device__ __forceinline__ uint4 pack<__half>(const float (&v)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __half2 h = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
        w[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// ---------------------------------------------------------------------------------------
// Welford / Chan running moments (count, mean, M2 = sum of squared deviations), fp32
// ---------------------------------------------------------------------------------------
struct Moments {
    float n, mean, m2;
};
__device__ __forceinline__ Moments moments_zero() { return Moments{0.f, 0.f, 0.f}; }

// Chan et al. pairwise merge; exact for empty operands.
__device__ __forceinline__ Moments merge(const Moments& a, const Moments& b) {
    const float n = a.n + b.n;
    if (n == 0.f) return a;
    const float d = b.mean - a.mean;
    const float wb = b.n / n;
    Moments r;
    r.n = n;
    r.mean = a.mean + d * wb;
    r.m2 = a.m2 + b.m2 + d * d * a.n * wb;
    return r;
}
// Fold K register-resident values: exact two-pass on the K values, then one Chan merge.
template <int K> __device__ __forceinline__ void fold(Moments& acc, const float (&v)[K]) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < K; ++i) s += v[i];
    const float m = s * (1.f / K);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < K; ++i) { const float d = v[i] - m; q = fmaf(d, d, q); }
    const float n = acc.n + K;
    const float d = m - acc.mean;
    const float wb = __fdividef((float)K, n);
    acc.mean = fmaf(d, wb, acc.mean);
    acc.m2 = acc.m2 + q + d * d * acc.n * wb;
    acc.n = n;
}
__device__ __forceinline__ void fold1(Moments& acc, float v) {   // classic Welford step
    acc.n += 1.f;
    const float d = v - acc.mean;
    acc.mean += __fdividef(d, acc.n);
    acc.m2 = fmaf(d, v - acc.mean, acc.m2);
}
__device__ __forceinline__ Moments warp_merge(Moments a) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Moments b;
        b.n = __shfl_xor_sync(0xffffffffu, a.n, o);
        b.mean = __shfl_xor_sync(0xffffffffu, a.mean, o);
        b.m2 = __shfl_xor_sync(0xffffffffu, a.m2, o);
        a = merge(a, b);
    }
    return a;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------------------------------
// "team" = the TPI threads that own one instance: a warp (TPI=32) or the whole CTA (TPI=256)
// ---------------------------------------------------------------------------------------
template <int TPI> struct Team {
    static_assert(TPI == kWarp || TPI == kBlock, "team is a warp or the CTA");
    static constexpr int per_block = kBlock / TPI;
    __device__ static int rank() { return TPI == kWarp ? (threadIdx.x & 31) : threadIdx.x; }
    __device__ static long long instance() {
        return TPI == kWarp ? (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5)
                            : (long long)blockIdx.x;
    }
    // All threads of the team get the merged value.  scratch: per-CTA shared, >= 8 entries/kind.
    __device__ static Moments all_merge(Moments a, Moments* scratch) {
        a = warp_merge(a);
        if (TPI == kWarp) return a;
        const int w = threadIdx.x >> 5;
        __syncthreads();                     // scratch may still be read from a previous call
        if ((threadIdx.x & 31) == 0) scratch[w] = a;
        __syncthreads();
        Moments r = scratch[0];
#pragma unroll
        for (int i = 1; i < kWarpsPerBlock; ++i) r = merge(r, scratch[i]);
        return r;
    }
    __device__ static float all_sum(float v, float* scratch) {
        v = warp_sum(v);
        if (TPI == kWarp) return v;
        const int w = threadIdx.x >> 5;
        __syncthreads();
        if ((threadIdx.x & 31) == 0) scratch[w] = v;
        __syncthreads();
        float r = 0.f;
#pragma unroll
        for (int i = 0; i < kWarpsPerBlock; ++i) r += scratch[i];
        return r;
    }
};

struct Window {
    int h0, h1, w0, w1;
    __host__ __device__ int rows() const { return h1 - h0; }
    __host__ __device__ int cols() const { return w1 - w0; }
    __host__ __device__ int area() const { return rows() * cols(); }
    __host__ __device__ bool full(int H, int W) const { return h0 == 0 && w0 == 0 && h1 == H && w1 == W; }
    __device__ bool has(int h, int w) const { return h >= h0 && h < h1 && w >= w0 && w < w1; }
};

// Moments of one instance over a window.  VEC: use 128-bit loads over the flat plane (only legal
// when the window is the full plane and M*sizeof(T) % 16 == 0 and the base is 16-byte aligned).
template <typename T, int TPI, bool VEC>
__device__ __forceinline__ Moments instance_moments(const T* __restrict__ plane, int W, int M,
                                                    const Window& win, bool win_full) {
    Moments acc = moments_zero();
    const int r = Team<TPI>::rank();
    if (VEC) {
        constexpr int V = VecOf<T>::n;
        constexpr int U = 4;                 // independent 128-bit loads in flight per thread
        const uint4* p = reinterpret_cast<const uint4*>(plane);
        const int nv = M / V;
        for (int i = r; i < nv; i += TPI * U) {
            uint4 raw[U];
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (i + u * TPI < nv) raw[u] = ldg_stream(p + i + u * TPI);
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (i + u * TPI < nv) {
                    float v[V];
                    unpack<T>(raw[u], v);
                    fold<V>(acc, v);
                }
        }
    } else if (win_full) {
#pragma unroll 4
        for (int i = r; i < M; i += TPI) fold1(acc, to_f(plane[i]));
    } else {
        const int cols = win.cols(), area = win.area();
        for (int i = r; i < area; i += TPI) {
            const int hh = i / cols, ww = i - hh * cols;
            fold1(acc, to_f(plane[(win.h0 + hh) * W + win.w0 + ww]));
        }
    }
    return acc;
}

__device__ __forceinline__ float std_from(const Moments& m, float eps) {
    // unbiased variance; a 1-element window yields 0/0 = NaN exactly like torch.var
    return sqrtf(m.m2 / (m.n - 1.f) + eps);
}

// ---------------------------------------------------------------------------------------
// host-side launch helpers
// ---------------------------------------------------------------------------------------
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline size_t esize(int dtype) { return dtype == CNSN_F32 ? 4 : 2; }
// Can every instance be streamed with 16-byte vectors?
inline bool vec_ok(const void* p, int dtype, int M) { return aligned16(p) && ((size_t)M * esize(dtype)) % 16 == 0; }
inline bool vec_ok2(const void* a, const void* b, int dtype, int M) { return vec_ok(a, dtype, M) && aligned16(b); }
// Team size: a warp per instance up to 4096 elements, the whole CTA above.
inline int team_for(int M) { return M <= 4096 ? kWarp : kBlock; }
inline unsigned grid_for(long long instances, int tpi) {
    const long long per = kBlock / tpi;
    return (unsigned)((instances + per - 1) / per);
}

#define CNSN_DISPATCH_DTYPE(dtype, T, ...)                                  \
    switch (dtype) {                                                        \
        case CNSN_F32: { using T = float; __VA_ARGS__; break; }             \
        case CNSN_BF16: { using T = __nv_bfloat16; __VA_ARGS__; break; }    \
        case CNSN_F16: { using T = __half; __VA_ARGS__; break; }            \
        default: return CNSN_E_BADARG;                                      \
    }
#define CNSN_DISPATCH_BOOL(flag, NAME, ...)                                 \
    if (flag) { constexpr bool NAME = true; __VA_ARGS__; }                  \
    else { constexpr bool NAME = false; __VA_ARGS__; }
#define CNSN_DISPATCH_TEAM(tpi, TPI, ...)                                   \
    if ((tpi) == ::cnsn::kWarp) { constexpr int TPI = ::cnsn::kWarp; __VA_ARGS__; } \
    else { constexpr int TPI = ::cnsn::kBlock; __VA_ARGS__; }

int check_dims(int N, int C, int H, int W);                     // stats.cu
int check_window(const Window& w, int H, int W);                // stats.cu
int launch_instance_stats(const void* x, int dtype, long long inst, int H, int W, const Window& win,
                          float eps, float* mean, float* sd, cudaStream_t s);   // stats.cu

inline int launch_status() {
    note_launch();
    return (int)cudaGetLastError();
}

}  // namespace cnsn

namespace cnsn {
// Elementwise pass over one instance: out[i] = f(a[i], b[i], i).  TWO: b is read (else ignored).
// LAST: inputs are not needed again after this pass (evict-first in L2).
template <typename T, int TPI, bool VEC, bool TWO, typename F>
__device__ __forceinline__ void plane_map(const T* __restrict__ a, const T* __restrict__ b,
                                          T* __restrict__ out, int M, F f) {
    const int r = Team<TPI>::rank();
    if (VEC) {
        constexpr int V = VecOf<T>::n;
        const uint4* pa = reinterpret_cast<const uint4*>(a);
        const uint4* pb = reinterpret_cast<const uint4*>(b);
        uint4* po = reinterpret_cast<uint4*>(out);
        const int nv = M / V;
        constexpr int U = TWO ? 2 : 4;       // 4 independent 128-bit loads in flight per thread
        const uint64_t pol = l2_policy_evict_first();
        for (int i = r; i < nv; i += TPI * U) {
            uint4 ra[U], rb[U];
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (i + u * TPI < nv) {
                    ra[u] = ldg_hint(pa + i + u * TPI, pol);
                    if (TWO) rb[u] = ldg_hint(pb + i + u * TPI, pol);
                }
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (i + u * TPI < nv) {
                    const int iv = i + u * TPI;
                    float va[V], vb[V], vo[V];
                    unpack<T>(ra[u], va);
                    if (TWO) unpack<T>(rb[u], vb);
#pragma unroll
                    for (int j = 0; j < V; ++j) vo[j] = f(va[j], TWO ? vb[j] : 0.f, iv * V + j);
                    stg_stream(po + iv, pack<T>(vo));
                }
        }
    } else {
#pragma unroll 4
        for (int i = r; i < M; i += TPI)
            out[i] = from_f<T>(f(to_f(a[i]), TWO ? to_f(b[i]) : 0.f, i));
    }
}
}  // namespace cnsn
