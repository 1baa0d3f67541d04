// fused_common.cuh -- building blocks of the persistent fused SelfNorm kernels (sm_100a):
// mbarrier + TMA bulk-copy wrappers, gpu-scope acquire/release flags, named barriers, and the
// static work schedule shared by the forward and backward kernels.
//
// Schedule.  Channels are cut into G groups of kk consecutive channels.  For one sample n the kk
// instances of a group are contiguous in NCHW memory: that run (kk*M elements) is a "unit", moved by
// ONE cp.async.bulk.  Group g consists of N units (one per sample); CTA b (of B co-resident CTAs)
// owns the units n = first(b,g) + j*B.  first() rotates with g so that the ceil/floor imbalance of
// N/B averages out across groups instead of pinning the same CTAs.
#pragma once

#include "common.cuh"

namespace cnsn {
namespace fused {

constexpr int kStatsWarps = 8;
constexpr int kApplyWarps = 16;                  // the apply stream is latency-bound per warp (L2 round trips): more warps,
constexpr int kApplyTeam = 1;                    //   warps per unit (teams of 3 cut the per-unit latency but measured slower overall)
constexpr int kApplyTeams = kApplyWarps / kApplyTeam;
constexpr int kWgThreads = 256;                 // threads in the stats / apply warpgroups
constexpr int kWarpProducer = kStatsWarps + kApplyWarps;      // issues TMA bulk loads
constexpr int kWarpChan = kWarpProducer + 1;    // first of kChanWarps channel warps (apply slot sl -> warp sl % kChanWarps)
constexpr int kChanWarps = 3;
constexpr int kThreads = (kWarpChan + kChanWarps) * 32;
constexpr int kSlots = 36;                      // max apply-stream slots (runtime: Schedule::R, a multiple of kChanWarps)
constexpr int kMaxPairs = 1024;                 // N*kk per group handled by one channel warp (32 slots/lane)
constexpr unsigned kSentinel = 0xffffffffu;     // "not yet published" tag in the second word of a pair
constexpr int kMaxStages = 12;
constexpr int kMaxInst = 64;                    // instances per stage
constexpr int kMaxKK = 32;                      // channels per group
constexpr unsigned kSpinLimit = 1u << 21;       // bounded global polls (~1 us each): trap instead of hanging the GPU
constexpr unsigned kWaitLimit = 1u << 22;       // bounded mbarrier waits (~0.25 us each)
constexpr unsigned kWaitSleepNs = 200;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier -----------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
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
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    unsigned spins = 0;
    // Waiting warps must not burn issue slots: the SM arbiter favours high warp ids, and a try_wait
    // that wakes on every TMA chunk arrival starves the warps doing the arithmetic.  Sleep explicitly.
    while (!mbar_try_wait(bar, parity)) {
        __nanosleep(kWaitSleepNs);
        if (++spins > kWaitLimit) __trap();
    }
}
// ---- TMA 1-D bulk copy global -> shared, completion on an mbarrier -------------------------
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, unsigned bytes, uint64_t* bar,
                                            uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
// Same without a cache hint (default L2 policy: the line stays until the apply stream re-reads it).
__device__ __forceinline__ void tma_load_1d_plain(void* dst_smem, const void* src_gmem, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// L2 prefetch of a future unit: the HBM latency (and its tail) is paid before a ring slot is tied up.
__device__ __forceinline__ void tma_prefetch_l2(const void* src_gmem, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(src_gmem), "r"(bytes) : "memory");
}
// ---- gpu-scope flags ------------------------------------------------------------------------
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_release_add(unsigned* p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void spin_until_ge(const unsigned* p, unsigned target) {
    unsigned spins = 0;
    while (ld_acquire(p) < target) {
        __nanosleep(40);
        if (++spins > kSpinLimit) __trap();
    }
}
__device__ __forceinline__ float ld_cg(const float* p) { return __ldcg(p); }   // L2 only: cross-CTA data

// "Data is the flag" exchange (as in low-latency collective protocols): a per-instance result is ONE
// aligned 8-byte store whose second word can never equal kSentinel, into an area pre-filled with
// kSentinel.  Readers poll the 8-byte word itself: no separate flag, no fence, one L2 hop each way.
__device__ __forceinline__ void ll_publish(float2* slot, float a, float b) {
    if (b != b) b = __uint_as_float(0x7fc00000u);      // canonical NaN: never the sentinel pattern
    asm volatile("st.relaxed.gpu.global.v2.f32 [%0], {%1, %2};" :: "l"(slot), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ float2 ll_peek(const float2* slot) {
    float2 v;
    asm volatile("ld.relaxed.gpu.global.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(slot));
    return v;
}
__device__ __forceinline__ bool ll_valid(const float2& v) { return __float_as_uint(v.y) != kSentinel; }

// ---- schedule ---------------------------------------------------------------------------------
struct Schedule {
    int N, C, M;            // batch, channels, plane elements
    int kk, G;              // channels per group, groups
    int B;                  // CTAs
    int upc;                // max units per CTA per group = ceil(N/B)
    int S;                  // ring stages
    int rot;                // rotation stride (coprime with B)
    int lpi;                // lanes per instance inside a warp (4..32, power of two)
    int R;                  // apply-stream slots = how many groups the reduce stream may run ahead of apply
    unsigned unit_elems;    // kk*M
    __device__ __forceinline__ int first(int b, int g) const { return (b + (int)(((long long)g * rot) % B)) % B; }
    __device__ __forceinline__ int count(int first_n) const { return first_n < N ? (N - first_n + B - 1) / B : 0; }
    // Which of `nw` role warps serves group g.  It is a function of the ring STAGE, so all phases of a
    // stage's mbarriers are awaited by one warp, in order: parity waits can then never alias (a warp
    // that reached a barrier a whole phase early would sail through it).
    __device__ __forceinline__ bool mine(int g, int w, int nw) const { return (g % S) % nw == w; }
    // Finer split for the streaming roles: unit j of stage st belongs to warp (st*upc + j) % nw, so the
    // upc (< nw) units of one group are processed by upc different warps concurrently; the stage's
    // completion barrier then counts upc arrivals (a warp arrives even when its unit is absent).
    __device__ __forceinline__ bool my_unit(int st, int j, int w, int nw) const { return (st * upc + j) % nw == w; }
};

// Division-free walk over the groups: ring stage, mbarrier phase parity and the CTA's first sample of
// the group are advanced incrementally (integer division by runtime values costs ~100 instructions).
struct GroupIter {
    int g, st, ph, first, cnt;
    int sl, sp;                                   // apply-stream slot (g % R) and its phase parity
    int q, rem;                                   // N = q*B + rem
    __device__ __forceinline__ void init(const Schedule& s, int b) {
        g = 0; st = 0; ph = 0; sl = 0; sp = 0; first = b; q = s.N / s.B; rem = s.N - q * s.B;
        cnt = q + (first < rem ? 1 : 0);
    }
    __device__ __forceinline__ void next(const Schedule& s) {
        ++g;
        if (++st == s.S) { st = 0; ph ^= 1; }
        if (++sl == s.R) { sl = 0; sp ^= 1; }
        first += s.rot;
        if (first >= s.B) first -= s.B;
        cnt = q + (first < rem ? 1 : 0);
    }
};

// Mean / sum of squared deviations of smem-resident instances in ONE pass over shared memory (the
// shared-memory pipe, not HBM, is what the stats warps contend for).  A warp handles 32/lpi instances
// at a time, lpi lanes each (rank r); `live` = this lane's instance exists.
//   1. shift K = mean of a strided sample spanning the whole plane (one 128-bit load per lane);
//   2. one pass accumulating sum(x-K) and sum((x-K)^2) with independent FMA chains;
//   3. mean = K + S1/M,  M2 = S2 - S1^2/M.
// Because K is within a fraction of a standard deviation of the true mean, S1^2/M is orders of
// magnitude below S2 and the subtraction loses no precision (this is the shifted-data algorithm, not
// the E[x^2]-mean^2 formula); cross-lane reductions are plain adds.
constexpr int kBatch = 8;

// Batched loops are written as "full batches without any bounds check, then a scalar tail": a bounds
// check per load compiles to a branch region (BSSY/BSYNC) per load and made this loop 3x slower.
template <typename T>
__device__ __forceinline__ float2 smem_mean_m2(const T* inst, int M, int r, int lpi, bool vec, bool live) {
    constexpr int V = VecOf<T>::n;
    const int nv = M / V;
    // 1. shift estimate: one element (vector) per lane, spread over the plane
    float ks = 0.f, kn = 0.f;
    if (live) {
        if (vec) {
            float v[V];
            unpack<T>(reinterpret_cast<const uint4*>(inst)[r * (nv / lpi)], v);
#pragma unroll
            for (int j = 0; j < V; ++j) ks += v[j];
            kn = (float)V;
        } else {
            ks = to_f(inst[r * (M / lpi)]);
            kn = 1.f;
        }
    }
    for (int o = lpi >> 1; o > 0; o >>= 1) {
        ks += __shfl_xor_sync(0xffffffffu, ks, o);
        kn += __shfl_xor_sync(0xffffffffu, kn, o);
    }
    const float K = kn > 0.f ? __fdividef(ks, kn) : 0.f;
    // 2. shifted sums
    float s1 = 0.f, s2 = 0.f;
    if (live) {
        if (vec) {
            const uint4* p = reinterpret_cast<const uint4*>(inst);
            float a1[4] = {0.f, 0.f, 0.f, 0.f}, a2[4] = {0.f, 0.f, 0.f, 0.f};
            const int step = lpi * kBatch;
            const int nfull = (nv / step) * step;
            int i = r;
            for (; i < nfull; i += step) {
                uint4 raw[kBatch];
#pragma unroll
                for (int u = 0; u < kBatch; ++u) raw[u] = p[i + u * lpi];
#pragma unroll
                for (int u = 0; u < kBatch; ++u) {
                    float v[V];
                    unpack<T>(raw[u], v);
#pragma unroll
                    for (int j = 0; j < V; ++j) {
                        const float d = v[j] - K;
                        a1[j & 3] += d;
                        a2[j & 3] = fmaf(d, d, a2[j & 3]);
                    }
                }
            }
            for (; i < nv; i += lpi) {
                float v[V];
                unpack<T>(p[i], v);
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    const float d = v[j] - K;
                    a1[j & 3] += d;
                    a2[j & 3] = fmaf(d, d, a2[j & 3]);
                }
            }
            s1 = (a1[0] + a1[1]) + (a1[2] + a1[3]);
            s2 = (a2[0] + a2[1]) + (a2[2] + a2[3]);
        } else {
            for (int i = r; i < M; i += lpi) {
                const float d = to_f(inst[i]) - K;
                s1 += d;
                s2 = fmaf(d, d, s2);
            }
        }
    }
    for (int o = lpi >> 1; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    const float dm = s1 / (float)M;
    return make_float2(K + dm, fmaxf(s2 - s1 * dm, 0.f));
}

}  // namespace fused
}  // namespace cnsn
