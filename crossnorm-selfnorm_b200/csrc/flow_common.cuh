// flow_common.cuh -- building blocks shared by the ticket-ordered dataflow kernels (selfnorm_flow.cu, crossnorm_flow.cu,
// site_flow.cu, ibn_flow.cu): mbarrier + TMA bulk-copy wrappers, polled 8-byte words ("data is the flag"), the
// asynchronous error word, team / CTA reductions, shared-memory vector loads, and the host-side launch helpers
// (tuning knobs, per-kernel preparation, cooperative persistent launch).
#pragma once

#include <stdlib.h>

#include <mutex>
#include <unordered_map>

#include "common.cuh"

namespace cnsn {

// ---- tuning knobs (A/B measurements and tests only) --------------------------------------------------------
// Process-wide, set through cnsn_tune() (include/cnsn_b200.h) -- NEVER read from the environment, so an
// inherited variable cannot change which kernel a training job runs.  0 / -1 = "default dispatch".
struct Knobs {
    int selfnorm_impl = 0;      // 0 auto, 1 three-kernel path, 3 dataflow kernels only
    int crossnorm_impl = 0;     // 0 auto, 1 two-kernel path
    int flow_mode = 0;          // SelfNorm forward: 0 auto, 1 shared-memory-resident, 2 L2 items
    int flow_bwd = 0;           // SelfNorm backward: 0 auto, 1 resident, 2 x resident + dy through L2, 3 L2 items
    int flow_d = 0;             // L2 items: look-ahead in channels (0 = from lookahead_mb)
    int flow_tpi = 0;           // threads per instance override
    int flow_batches = 6;       // L2 items: batches of kU loads a thread covers its plane in
    int lookahead_mb = 40;      // L2 items: bytes kept between the reduce and the apply stream
    int item_kb = 25;           // resident items: shared memory per CTA aimed at
    int grp_kb = 20;            // channel-group items: shared memory per CTA aimed at
    int keep = 1;               // L2 items: R loads evict-last
    int pf = -1;                // resident items: L2 prefetch distance in tickets (-1 = half the resident CTAs)
    int rpf = 0;                // L2 items: prefetch distance in channels
    int poll_ns = 100;          // sleep between polls of a published word
    int i3 = 1;                 // forward: three planes per resident item (192 threads) where two would be the geometry
    int tm = 1;                 // SelfNorm: the shared-memory + tensor-memory pipeline (selfnorm_tmem.cu) where it applies
    int tm_items = 16;          // ... which needs at least this many items per SM to fill its two stages (tests: 0)
    int cooperative = 1;        // resident items: cooperative launch (co-residency guaranteed by the driver)
    int grid_cap = 0;           // resident items: cap the persistent grid (0 = every CTA the GPU holds)
    int debug = 0;              // print the chosen geometry to stderr
    int trace = 0;              // resident SelfNorm: per-item timestamps to $CNSN_FLOW_TRACE (synchronous, debug)
};
Knobs& knobs();                                  // api.cu

// ---- asynchronous error word ---------------------------------------------------------------------------------
// A wait that exceeds its bound (seconds: only a bug, a debugger or a sanitizer gets there) does NOT trap -- a trap
// destroys the CUDA context of a training job.  The waiting thread records a code in a pinned, device-mapped host
// word and carries on with whatever it read; the host finds the code at its next call (CNSN_E_TIMEOUT) or through
// cnsn_async_error().  The results of that one launch are undefined, the process stays usable.
unsigned* async_error_word();                    // api.cu: device-visible pointer, allocated on first use
int async_error_peek();                          // api.cu: host-side read (0 = none)
enum { kErrPollTimeout = 1, kErrTmaTimeout = 2 };

namespace flow {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
constexpr unsigned long long kWaitBoundNs = 20ull * 1000 * 1000 * 1000;    // 20 s
static __device__ __noinline__ void report_timeout(unsigned* err, unsigned code) {
    if (err) atomicCAS(err, 0u, code);
    __threadfence_system();
}

// ---- mbarrier + TMA 1-D bulk copies (SASS: UBLKCP / UBLKPF / SYNCS) ---------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, unsigned parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Waiting warps sleep explicitly: a try_wait that wakes on every TMA chunk starves the warps doing arithmetic.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity, unsigned* err) {
    unsigned spins = 0;
    unsigned long long t0 = 0;
    while (!mbar_try_wait(bar, parity)) {
        __nanosleep(200);
        if ((++spins & 0xfffu) == 0) {
            const unsigned long long now = gtime();
            if (!t0) t0 = now;
            else if (now - t0 > kWaitBoundNs) { report_timeout(err, kErrTmaTimeout); return; }
        }
    }
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, unsigned bytes, uint64_t* bar,
                                            uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
// L2 prefetch of a future item: the HBM latency (and its tail) is paid before shared memory is tied up.
__device__ __forceinline__ void tma_prefetch_l2(const void* src_gmem, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(src_gmem), "r"(bytes) : "memory");
}
// generic-proxy accesses to shared memory made so far are ordered before later async-proxy (TMA) writes
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- "data is the flag" exchange (as in low-latency collective protocols) ----------------------------------
// A per-instance result is ONE aligned 8-byte store whose second word can never equal the sentinel, into an area
// pre-filled with the sentinel (0xff bytes).  Readers poll the 8-byte word itself: no separate flag, no fence, one
// L2 hop each way.  A NaN payload is canonicalised (0x7fc00000), an all-ones NaN cannot be published.
constexpr unsigned kSentinel = 0xffffffffu;
__device__ __forceinline__ void ll_publish(float2* slot, float a, float b) {
    if (b != b) b = __uint_as_float(0x7fc00000u);
    asm volatile("st.relaxed.gpu.global.v2.f32 [%0], {%1, %2};" :: "l"(slot), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ float2 ll_peek(const float2* slot) {
    float2 v;
    asm volatile("ld.relaxed.gpu.global.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(slot));
    return v;
}
__device__ __forceinline__ bool ll_valid(const float2& v) { return __float_as_uint(v.y) != kSentinel; }

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
// Who synchronises with whom: the whole CTA (default), or one 128-thread group of a fat CTA (named barrier; the
// tensor-memory kernel runs four independent groups per CTA, selfnorm_tmem.cu).
struct CtaSync {
    __device__ __forceinline__ static int tid() { return threadIdx.x; }
    __device__ __forceinline__ static void sync() { __syncthreads(); }
};
struct Group128Sync {
    __device__ __forceinline__ static int tid() { return threadIdx.x & 127; }
    __device__ __forceinline__ static void sync() {
        asm volatile("bar.sync %0, 128;" :: "r"((threadIdx.x >> 7) + 1) : "memory");     // barrier 0 stays the CTA's
    }
};
// Sums over the TH threads that synchronise through S (channel fold by the last item of a channel).
template <int K, int TH, typename S = CtaSync>
__device__ __forceinline__ void cta_sums(float (&v)[K], float (*sm)[TH / 32]) {
    const int warp = S::tid() >> 5, lane = S::tid() & 31;
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = warp_sum(v[k]);
    S::sync();
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) sm[k][warp] = v[k];
    }
    S::sync();
#pragma unroll
    for (int k = 0; k < K; ++k) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < TH / 32; ++w) s += sm[k][w];
        v[k] = s;
    }
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
// Poll one published word until it is there.  Bounded (kWaitBoundNs): on expiry the error word is set and the
// sentinel-valued read is returned -- no trap.
__device__ __forceinline__ float2 poll_word(const float2* p, int sleep_ns, unsigned* err) {
    float2 v = ll_peek(p);
    unsigned spins = 0;
    unsigned long long t0 = 0;
    while (!ll_valid(v)) {
        __nanosleep(sleep_ns);
        v = ll_peek(p);
        if ((++spins & 0xfffu) == 0) {
            const unsigned long long now = gtime();
            if (!t0) t0 = now;
            else if (now - t0 > kWaitBoundNs) { report_timeout(err, kErrPollTimeout); break; }
        }
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

// The persistent ticket loop every shared-memory-resident kernel runs (cooperative launch, grid = the CTAs the GPU
// holds at once): take a ticket, run the item, until the tickets run out.  The item's bulk copies complete on ONE
// mbarrier per CTA whose phase parity alternates with the iteration; the barrier at the end of an iteration orders
// the item's last shared-memory reads before the next item's bulk copies overwrite them.
#define CNSN_TICKET_LOOP(ARGS, ITEM_CALL)                                                         \
    extern __shared__ __align__(128) unsigned char dsm_loop[];                                    \
    __shared__ unsigned s_ticket;                                                                 \
    if (threadIdx.x == 0) {                                                                       \
        mbar_init(reinterpret_cast<uint64_t*>(dsm_loop), 1);                                      \
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");                        \
    }                                                                                             \
    for (unsigned it = 0;; ++it) {                                                                \
        if (threadIdx.x == 0) s_ticket = atomicAdd((ARGS).ticket, 1u) + 1u; /* starts at 0xffffffff */ \
        __syncthreads();                                                                          \
        const unsigned t = s_ticket;                                                              \
        if (t >= (ARGS).items) break;                                                             \
        ITEM_CALL;                                                                                \
        fence_proxy_async_smem();                                                                 \
        __syncthreads();                                                                          \
    }

// host: per-(device, kernel) launch preparation, done once: opt in to the device's maximum dynamic shared memory
// (the attribute is per function and last-write-wins, so it is set ONCE to the maximum, never to one call's size),
// prefer the maximum carve-out; the occupancy for a given dynamic size is cached per (device, kernel, size).
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
    static std::unordered_map<unsigned long long, int> occ;       // (device, kernel, dsmem) -> CTAs per SM
    static std::unordered_map<unsigned long long, bool> ready;    // (device, kernel) -> attributes set
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long fkey = ((unsigned long long)dev << 56) ^ ((unsigned long long)(uintptr_t)fn * 0x9e3779b97f4a7c15ull);
    const unsigned long long key = fkey ^ (dsmem * 0xc2b2ae3d27d4eb4full);
    std::lock_guard<std::mutex> lock(mu);
    auto it = occ.find(key);
    if (it != occ.end()) { *ctas_per_sm = it->second; return cudaSuccess; }
    cudaError_t e = cudaSuccess;
    if (!ready.count(fkey)) {
        const DeviceShape d = device_shape();
        cudaFuncAttributes fa{};
        e = cudaFuncGetAttributes(&fa, fn);
        if (e != cudaSuccess) return e;
        // static + dynamic shared memory together are bounded by the opt-in limit
        e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, d.smem_optin - (int)fa.sharedSizeBytes);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        if (e != cudaSuccess) return e;
        ready[fkey] = true;
    }
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, fn, threads, dsmem);
    if (e == cudaSuccess) occ[key] = *ctas_per_sm;
    return e;
}

// host: launch a persistent ticket kernel.  The grid is capped at the number of CTAs the GPU holds at once
// (occupancy x SMs) and launched COOPERATIVELY: the driver then guarantees that every CTA of the grid is resident
// at the same time whatever else shares the GPU (NCCL kernels, other streams, MPS), which is what the spin-waits
// between the CTAs of a channel rely on.  Each CTA loops over tickets until they run out.  A grid the driver will not
// co-schedule (cudaErrorCooperativeLaunchTooLarge: fewer SMs than the occupancy query assumed, e.g. under an SM
// partition) is not launched: the callers then take the path that needs no residency (L2 items / general kernels).
template <typename K, typename A>
static inline cudaError_t launch_persistent(K fn, const A& args, unsigned items, unsigned channel_items, int per_sm, int sms,
                                            int threads, size_t dsmem, cudaStream_t stream) {
    unsigned long long cap = (unsigned long long)per_sm * sms;
    const Knobs& kn = knobs();
    // test knob: a smaller grid (every CTA then runs many items); never below one channel's items -- a channel's
    // items wait for each other, so all of them must be held by CTAs at the same time
    if (kn.grid_cap > 0 && (unsigned long long)kn.grid_cap < cap) cap = kn.grid_cap < (int)channel_items ? channel_items : kn.grid_cap;
    const unsigned grid = (unsigned)(items < cap ? items : cap);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = dsmem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = kn.cooperative ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, fn, args);
}

}  // namespace flow
}  // namespace cnsn
