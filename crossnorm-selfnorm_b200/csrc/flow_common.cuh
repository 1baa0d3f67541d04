// flow_common.cuh -- building blocks shared by the ticket-ordered dataflow kernels (selfnorm_flow.cu,
// crossnorm_flow.cu): gpu-scope flags, polled 8-byte words, team / CTA reductions, shared-memory vector loads.
#pragma once

#include <stdlib.h>

#include <mutex>
#include <unordered_map>

#include "fused_common.cuh"

namespace cnsn {
namespace flow {

using fused::smem_u32;

constexpr unsigned kSpin = 1u << 25;    // bounded polls (>= 64 ns each, i.e. seconds): trap instead of hanging the GPU

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

// Sum over the TPI threads of a team (TPI <= 32: lanes of a warp; else TPI/32 consecutive warps).
// Every thread of the CTA must call it (it may contain __syncthreads).
template <int TPI>
__device__ __forceinline__ float team_sum(float v, float* sm) {
#pragma unroll
    for (int o = (TPI < 32 ? TPI : 32) >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (TPI <= 32) return v;
    const int w = threadIdx.x >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm[w] = v;
    __syncthreads();
    constexpr int WPT = TPI > 32 ? TPI / 32 : 1;
    const int w0 = (w / WPT) * WPT;
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < WPT; ++i) r += sm[w0 + i];
    return r;
}
template <int TPI>
__device__ __forceinline__ Moments team_merge(Moments a, Moments* sm) {
#pragma unroll
    for (int o = (TPI < 32 ? TPI : 32) >> 1; o > 0; o >>= 1) {
        Moments b;
        b.n = __shfl_xor_sync(0xffffffffu, a.n, o);
        b.mean = __shfl_xor_sync(0xffffffffu, a.mean, o);
        b.m2 = __shfl_xor_sync(0xffffffffu, a.m2, o);
        a = merge(a, b);
    }
    if (TPI <= 32) return a;
    const int w = threadIdx.x >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm[w] = a;
    __syncthreads();
    constexpr int WPT = TPI > 32 ? TPI / 32 : 1;
    const int w0 = (w / WPT) * WPT;
    Moments r = sm[w0];
#pragma unroll
    for (int i = 1; i < WPT; ++i) r = merge(r, sm[w0 + i]);
    return r;
}
// Sums over the whole CTA of TH threads (channel fold by the last R item).
template <int K, int TH>
__device__ __forceinline__ void cta_sums(float (&v)[K], float (*sm)[TH / 32]) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = warp_sum(v[k]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) sm[k][warp] = v[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < TH / 32; ++w) s += sm[k][w];
        v[k] = s;
    }
}

__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
    return r;
}

// one element of a shared-memory-resident plane, as float
template <typename T> __device__ __forceinline__ float lds_elem(uint32_t base, int e);
template <> __device__ __forceinline__ float lds_elem<float>(uint32_t base, int e) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(base + 4u * e));
    return v;
}
template <> __device__ __forceinline__ float lds_elem<__nv_bfloat16>(uint32_t base, int e) {
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(base + 2u * e));
    return __uint_as_float((unsigned)v << 16);
}
template <> __device__ __forceinline__ float lds_elem<__half>(uint32_t base, int e) {
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(base + 2u * e));
    return __half2float(__ushort_as_half(v));
}

__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ float2 poll_word(const float2* p, int sleep_ns) {
    float2 v = fused::ll_peek(p);
    unsigned spins = 0;
    while (!fused::ll_valid(v)) {
        __nanosleep(sleep_ns);
        v = fused::ll_peek(p);
        if (++spins > kSpin) __trap();
    }
    return v;
}

// sum over a window of f(x [, dy]) for one shared-memory-resident instance; the team's threads split the work.
// full: 128-bit reads over the flat plane; else element reads over the window's rows.
template <typename T, bool TWO, typename F>
__device__ __forceinline__ void window_accumulate(uint32_t sx, uint32_t sdy, int W, int M, const Window& win, bool full,
                                                  int r, int tpi, F f) {
    constexpr int V = VecOf<T>::n;
    if (full) {
        const int nv = M / V;
#pragma unroll 4
        for (int i = r; i < nv; i += tpi) {
            float vx[V], vd[V];
            unpack<T>(lds128(sx + 16u * i), vx);
            if (TWO) unpack<T>(lds128(sdy + 16u * i), vd);
#pragma unroll
            for (int e = 0; e < V; ++e) f(vx[e], TWO ? vd[e] : 0.f, e);
        }
    } else {
        const int cols = win.cols(), area = win.area();
        int hh = r / cols, ww = r - hh * cols;               // one division per thread, then incremental
        const int dh = tpi / cols, dw = tpi - dh * cols;
        for (int i = r; i < area; i += tpi) {
            const int o = (win.h0 + hh) * W + win.w0 + ww;
            f(lds_elem<T>(sx, o), TWO ? lds_elem<T>(sdy, o) : 0.f, i);
            hh += dh; ww += dw;
            if (ww >= cols) { ww -= cols; ++hh; }
        }
    }
}

// exact two-pass (mean, sqrt(unbiased var + eps)) of one window; every thread of the CTA must call it
template <typename T, int TPI>
__device__ __forceinline__ float2 window_stats(uint32_t sx, int W, int M, const Window& win, bool full, int r, bool live,
                                               float eps, float* sm0, float* sm1) {
    const float cnt = (float)win.area();
    float s0 = 0.f, s1 = 0.f;
    if (live) window_accumulate<T, false>(sx, 0u, W, M, win, full, r, TPI, [&](float x, float, int e) { if (e & 1) s1 += x; else s0 += x; });
    const float mean = team_sum<TPI>(s0 + s1, sm0) / cnt;
    s0 = s1 = 0.f;
    if (live) window_accumulate<T, false>(sx, 0u, W, M, win, full, r, TPI, [&](float x, float, int e) {
        const float d = x - mean;
        if (e & 1) s1 = fmaf(d, d, s1); else s0 = fmaf(d, d, s0);
    });
    const float m2 = team_sum<TPI>(s0 + s1, sm1);
    // a 1-element window yields 0/0 = NaN exactly like torch.var
    return make_float2(mean, sqrtf(m2 / (cnt - 1.f) + eps));
}

// host: per-(device, kernel, dynamic shared memory) launch preparation, done once: opt in to the shared-memory
// size, prefer the maximum carve-out, ask the occupancy API how many CTAs fit an SM.  (Three driver calls that
// would otherwise be paid on every launch; they dominate the host time of small tensors.)
struct DeviceShape { int sms = 0, smem_optin = 0; };
static inline DeviceShape device_shape() {
    static std::mutex mu;
    static std::unordered_map<int, DeviceShape> cache;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(dev);
    if (it != cache.end()) return it->second;
    DeviceShape d;
    cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&d.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cache[dev] = d;
    return d;
}
template <typename K>
static inline cudaError_t prepare_kernel(K fn, int threads, size_t dsmem, int* ctas_per_sm) {
    static std::mutex mu;
    static std::unordered_map<unsigned long long, int> cache;     // per kernel instantiation (K is its type)
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long key = ((unsigned long long)dev << 56) ^ ((unsigned long long)(uintptr_t)fn * 0x9e3779b97f4a7c15ull) ^ dsmem;
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *ctas_per_sm = it->second; return cudaSuccess; }
    cudaError_t e = cudaSuccess;
    if (dsmem > 48 * 1024) e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsmem);
    if (e == cudaSuccess && dsmem) e = cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, fn, threads, dsmem);
    if (e == cudaSuccess) cache[key] = *ctas_per_sm;
    return e;
}

// host: integer environment knob (A/B measurements only)
static inline int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}


}  // namespace flow
}  // namespace cnsn
