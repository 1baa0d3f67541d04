// selfnorm_flow.cu -- SelfNorm forward / backward as ONE ticket-ordered dataflow kernel (sm_100a).
//
// The gate of channel c couples all N instances of c (BatchNorm1d over the batch, models/cnsn.py:121,
// :138): no element of a channel can be written before the whole channel has been reduced.  The
// three-kernel path (selfnorm.cu) therefore reads x (and dy) twice from HBM; the persistent kernels
// (selfnorm_fused.cu) fix the traffic but bind every unit of work to one CTA, so each channel is gated
// by the slowest of 148 software pipelines.  Here the hardware CTA scheduler does the balancing:
//
//   * the grid is a list of small, non-persistent work items, two kinds per channel:
//       R(c, j)  reduce I instances of channel c   (forward: mean / std; backward: sum dy*x),
//                publish one 8-byte word per instance into a [C][N] area, then count itself done;
//                the LAST R item of a channel (atomic counter) folds the N words into the channel
//                constants (forward: BatchNorm batch mean / rstd + running statistics; backward:
//                dgamma, dbeta, dw and the two scalars of the batch-norm backward) and raises ready[c];
//       A(c, j)  apply: issue the plane loads, wait for ready[c], rebuild the instance's gate (or its
//                backward coefficients) and stream y / dx out.
//   * items are handed out in TICKET order (one atomicAdd per CTA -- or blockIdx order, see `order`):
//       R(0..D-1), then alternately R(k+D, j), A(k, j), ...  so an A item only ever waits for items with
//       smaller tickets, which already run and never wait themselves => no deadlock, no cooperative launch.
//   * D (look-ahead, in channels) is a few microseconds of HBM time: when A(k, *) is dispatched its
//     planes were read D channels ago and are still in L2 (D*N*M*sizeof(T)*tensors bytes, sized to a
//     fraction of the 126 MB L2), so the second read never goes to HBM: traffic 2*S forward, 3*S backward.
//   * R loads are marked L2 evict_last, A re-reads evict_first (last use), results are streaming stores, so
//     the planes between the two streams are what L2 keeps.  Threads keep kU 128-bit loads per tensor in
//     flight; TPI threads share an instance (pick_tpi: the resident window must stay narrow enough for L2).
//
// Planes must be streamable with 16-byte vectors; other shapes use the three-kernel path.
#include <stdio.h>
#include <stdlib.h>

#include "selfnorm_fold.cuh"

namespace cnsn {
namespace flow {

constexpr int kT = 256;                 // threads per CTA
constexpr int kUReg = 4;                // 128-bit loads in flight per thread and tensor


// Fold the N published words of channel c into the channel constants (whole CTA of TH threads; the caller has
// made the words visible: fence + atomic counter).  Forward: BatchNorm batch mean / rstd of s = w0*mu + w1*sd,
// running statistics (models/cnsn.py:121,138); backward: dgamma, dbeta, dw and the two batch-norm backward scalars.
template <bool BWD, int TH>
__device__ __forceinline__ void channel_fold(const FArgs& a, unsigned c, float (*s_f)[TH / 32]) {
    const int N = a.N, C = a.C;
    const float2* pb = a.pub + (size_t)c * N;
    const float invN = 1.f / N;
    if (!BWD) {
        const float w0 = a.w[2 * c], w1 = a.w[2 * c + 1];
        float m, q;
        if (a.training) {
            float v[1] = {0.f};
            for (int k = threadIdx.x; k < N; k += TH) { const float2 p = __ldcg(pb + k); v[0] += fmaf(w0, p.x, w1 * p.y); }
            cta_sums<1, TH>(v, s_f);
            m = v[0] / N;
            v[0] = 0.f;
            for (int k = threadIdx.x; k < N; k += TH) {
                const float2 p = __ldcg(pb + k);
                const float d = fmaf(w0, p.x, w1 * p.y) - m;
                v[0] = fmaf(d, d, v[0]);
            }
            cta_sums<1, TH>(v, s_f);
            q = v[0] / N;                // biased variance normalises (BatchNorm semantics)
            if (threadIdx.x == 0) {
                a.run_mean[c] = (1.f - a.momentum) * a.run_mean[c] + a.momentum * m;
                a.run_var[c] = (1.f - a.momentum) * a.run_var[c] + a.momentum * (q * N / (N - 1.f));
                if (a.nbt && c == 0) *a.nbt += 1;
            }
        } else {
            m = a.run_mean[c]; q = a.run_var[c];
        }
        if (threadIdx.x == 0) {
            const float rstd = 1.f / sqrtf(q + a.bn_eps);
            a.r[c] = rstd;
            a.chan[c] = make_float2(m, rstd);
        }
    } else {
        const float ga = a.gamma[c], rstd = a.r[c];
        float v[2] = {0.f, 0.f};
        for (int k = threadIdx.x; k < N; k += TH) { const float2 p = __ldcg(pb + k); v[0] = fmaf(p.x, p.y, v[0]); v[1] += p.x; }
        cta_sums<2, TH>(v, s_f);
        const float dgam = v[0], dbet = v[1];
        const float k1 = a.training ? ga * dbet * invN : 0.f, k2 = a.training ? ga * dgam * invN : 0.f;
        v[0] = v[1] = 0.f;
        for (int k = threadIdx.x; k < N; k += TH) {
            const float2 p = __ldcg(pb + k);
            const size_t i = (size_t)k * C + c;
            const float ds = rstd * (p.x * ga - k1 - p.y * k2);
            v[0] = fmaf(ds, a.mu[i], v[0]); v[1] = fmaf(ds, a.sd[i], v[1]);
        }
        cta_sums<2, TH>(v, s_f);
        if (threadIdx.x == 0) {
            a.dgamma[c] = dgam; a.dbeta[c] = dbet;
            a.dw[2 * c] = v[0]; a.dw[2 * c + 1] = v[1];
            a.chan[c] = make_float2(k1, k2);
        }
    }
}

#define CNSN_FTRACE(slot)                                                                      \
    do {                                                                                       \
        if (a.trace && threadIdx.x == 0) a.trace[(size_t)t * 8 + (slot)] = gtime();             \
    } while (0)

// 128-bit load of data WRITTEN earlier in this kernel by another CTA (ordered by a release / acquire flag): the
// coherent path, not ld.global.nc.
__device__ __forceinline__ uint4 ldg_hint_coherent(const void* p, uint64_t pol) {
    uint4 r;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(pol) : "memory");
    return r;
}
// Streaming store that keeps the line in L2 (the sum z = x + res is re-read by the A items).
__device__ __forceinline__ void stg_hint(void* p, const uint4& v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol) : "memory");
}

// L2-resident items.  Optional block fusion: forward with a.res != NULL reduces z = x + res,
// writes z (backward needs it) and applies to z; a.relu clamps y at 0 (forward) / masks dy where z <= 0 (backward:
// the gate is a sigmoid, so sign(g*z) = sign(z)).
template <typename T, bool BWD, bool ADD, int TPI>
__global__ void __launch_bounds__(kT, (BWD || ADD) ? 4 : 5) k_sn_flow(const FArgs a) {
    static_assert(!(BWD && ADD), "the fused add is a forward feature");
    constexpr int I = kT / TPI;          // instances per item
    constexpr int V = VecOf<T>::n;
    constexpr int kU = kUReg;
    __shared__ unsigned s_word;
    __shared__ float s_f[2][kT / 32];
    __shared__ Moments s_m[kT / 32];

    // ---- which item am I --------------------------------------------------------------------
    if (threadIdx.x == 0) s_word = atomicAdd(a.ticket, 1u);
    __syncthreads();
    const unsigned t = s_word;
    __syncthreads();                                         // s_word is reused below
    const unsigned nI = (unsigned)a.nI, D = (unsigned)a.D, C = (unsigned)a.C;
    bool isA;
    unsigned c, j;
    if (t < D * nI) {
        isA = false; c = t / nI; j = t - c * nI;
    } else {
        const unsigned u = t - D * nI, both = (C - D) * 2u * nI;
        if (u < both) {
            const unsigned k = u / (2u * nI), i = u - k * 2u * nI;
            isA = (i & 1u) != 0; j = i >> 1; c = isA ? k : k + D;
        } else {
            const unsigned v = u - both, k = v / nI;
            isA = true; c = (C - D) + k; j = v - k * nI;
        }
    }
    const int N = a.N, M = a.M;
    const int n = (int)j * I + (int)(threadIdx.x / TPI);
    const int r = threadIdx.x % TPI;
    const bool live = n < N;
    const size_t nc = (size_t)(live ? n : 0) * C + c;
    const int nv = M / V;
    constexpr bool add = ADD;                             // forward block fusion: x + res
    const bool relu = a.relu != 0;
    constexpr bool two = BWD || ADD;
    // R items read x (and dy / res); A items read what gets applied: x, or the sum z written by the R items
    const uint4* px = reinterpret_cast<const uint4*>(static_cast<const T*>((isA && add) ? a.zout : a.x) + nc * M);
    const uint4* pd = two ? reinterpret_cast<const uint4*>(static_cast<const T*>(BWD ? a.dy : a.res) + nc * M) : nullptr;
    constexpr int kStep = TPI * kU;

    if (!isA) {
        // =============================================================== R item
        const uint64_t pol_keep = l2_policy_evict_last();
        const uint64_t pol_once = l2_policy_evict_first();
        const bool keep = a.keep != 0 && !add;           // fused add: the inputs are not needed again, z is
        uint4* pz = add ? reinterpret_cast<uint4*>(static_cast<T*>(a.zout) + nc * M) : nullptr;
        float sxy = 0.f, pre_g = 0.f, pre_s = 0.f;
        if (BWD && live && r == 0) { pre_g = a.gate[nc]; pre_s = a.shat[nc]; }   // issued ahead of the plane loads
        // L2 prefetch (cp.async.bulk.prefetch) of the planes the R item pf_dist channels ahead will read: its
        // loads then find them in L2 and the item holds its registers for an L2, not an HBM, round trip.
        if (a.pf_dist && c + (unsigned)a.pf_dist < C && r == 0 && live) {
            const size_t off = ((size_t)n * C + c + (unsigned)a.pf_dist) * M;
            const unsigned pbytes = (unsigned)M * (unsigned)sizeof(T);
            tma_prefetch_l2(static_cast<const T*>(a.x) + off, pbytes);
            if (two) tma_prefetch_l2(static_cast<const T*>(BWD ? a.dy : a.res) + off, pbytes);
        }
        Moments acc = moments_zero();
        for (int i0 = r; i0 < nv; i0 += kStep) {
            uint4 rx[kU], rd[kU];
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const int i = i0 + u * TPI;
                if (live && i < nv) {
                    rx[u] = keep ? ldg_hint(px + i, pol_keep) : (add ? ldg_hint(px + i, pol_once) : ldg_stream(px + i));
                    if (two) rd[u] = keep ? ldg_hint(pd + i, pol_keep) : (add ? ldg_hint(pd + i, pol_once) : ldg_stream(pd + i));
                }
            }
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const int i = i0 + u * TPI;
                if (live && i < nv) {
                    float vx[V], vd[V];
                    unpack<T>(rx[u], vx);
                    if (two) unpack<T>(rd[u], vd);
                    if (BWD) {
#pragma unroll
                        for (int e = 0; e < V; ++e) {
                            const float d = (relu && !(vx[e] > 0.f)) ? 0.f : vd[e];
                            sxy = fmaf(d, vx[e], sxy);
                        }
                    } else {
                        if (add) {
#pragma unroll
                            for (int e = 0; e < V; ++e) vx[e] += vd[e];
                            const uint4 z = pack<T>(vx);
                            stg_hint(pz + i, z, pol_keep);    // stays in L2 for the A item
                            unpack<T>(z, vx);                 // statistics of the ROUNDED sum, as an unfused add gives
                        }
                        fold<V>(acc, vx);
                    }
                }
            }
        }
        if (BWD) {
            sxy = team_sum<TPI>(sxy, s_f[0]);
            if (live && r == 0) a.pub[(size_t)c * N + n] = make_float2(sxy * pre_g * (1.f - pre_g), pre_s);
        } else {
            acc = team_merge<TPI>(acc, s_m);
            if (live && r == 0) {
                const float mean = acc.mean, sdev = std_from(acc, a.eps);
                a.mu[nc] = mean; a.sd[nc] = sdev;
                a.pub[(size_t)c * N + n] = make_float2(mean, sdev);
            }
        }
        // ---- count this item; the last one of the channel folds the channel ----------------------
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            s_word = atomicAdd(a.done + c, 1u);
        }
        __syncthreads();
        if (s_word != nI - 1) return;
        __threadfence();
        channel_fold<BWD, kT>(a, c, s_f);
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            st_release_u32(a.ready + c, 1u);
        }
        return;
    }

    // =================================================================== A item
    const uint64_t pol = l2_policy_evict_first();
    uint4* po = reinterpret_cast<uint4*>(static_cast<T*>(a.out) + nc * M);
    uint4 rx[kU], rd[kU];
    auto issue = [&](int i0) {
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const int i = i0 + u * TPI;
            if (live && i < nv) {
                rx[u] = add ? ldg_hint_coherent(px + i, pol) : ldg_hint(px + i, pol);   // L2 hit, last use
                if (BWD) rd[u] = ldg_hint(pd + i, pol);
            }
        }
    };
    // With the fused add the planes are WRITTEN by this kernel's R items: their visibility comes with ready[c]
    // (release / acquire), so the loads wait for the flag; otherwise they are issued under the flag wait.
    if (!add) issue(r);
    // ... nor do the saved per-instance statistics and the parameters: fetch them under the flag wait too
    float p_w0 = 0.f, p_w1 = 0.f, p_ga = 0.f, p_b = 0.f, p_gt = 0.f, p_mu = 0.f, p_sd = 1.f;
    if (live) {
        p_w0 = a.w[2 * c]; p_w1 = a.w[2 * c + 1]; p_ga = a.gamma[c];
        if (BWD) { p_b = a.r[c]; p_gt = a.gate[nc]; p_mu = a.mu[nc]; p_sd = a.sd[nc]; }
        else p_b = a.beta[c];
    }
    if (threadIdx.x == 0) {                                  // the R items of c hold smaller tickets: they run or are done
        unsigned spins = 0;
        unsigned long long t0 = 0;
        while (ld_acquire_u32(a.ready + c) == 0u) {
            __nanosleep(64);
            if ((++spins & 0xfffu) == 0) {
                const unsigned long long now = gtime();
                if (!t0) t0 = now;
                else if (now - t0 > kWaitBoundNs) { report_timeout(a.err, kErrPollTimeout); break; }
            }
        }
    }
    __syncthreads();
    if (add) issue(r);
    float ca = 0.f, cb = 0.f, cc = 0.f;                   // out = ca*dy + cb*x + cc
    if (live) {
        const float2 cm = __ldcg(a.chan + c);
        const float2 own = __ldcg(a.pub + (size_t)c * N + n);
        const float w0 = p_w0, w1 = p_w1, ga = p_ga;
        if (BWD) {
            const float gt = p_gt, mean = p_mu, sdev = p_sd, rstd = p_b;
            const float ds = rstd * (own.x * ga - cm.x - own.y * cm.y);
            ca = gt;
            cb = ds * w1 * (1.f / (M - 1.f)) / sdev;
            cc = ds * w0 * (1.f / M) - cb * mean;
        } else {
            const float sh = (fmaf(w0, own.x, w1 * own.y) - cm.x) * cm.y;
            const float gt = 1.f / (1.f + expf(-fmaf(ga, sh, p_b)));
            if (r == 0) { a.gate[nc] = gt; a.shat[nc] = sh; }
            cb = gt;
        }
    }
    for (int i0 = r;;) {
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const int i = i0 + u * TPI;
            if (live && i < nv) {
                float vx[V], vd[V], vo[V];
                unpack<T>(rx[u], vx);
                if (BWD) unpack<T>(rd[u], vd);
#pragma unroll
                for (int e = 0; e < V; ++e) {
                    if (BWD) {
                        const float d = (relu && !(vx[e] > 0.f)) ? 0.f : vd[e];
                        vo[e] = fmaf(ca, d, fmaf(cb, vx[e], cc));
                    } else {
                        const float y = fmaf(cb, vx[e], 0.f);
                        vo[e] = relu ? fmaxf(y, 0.f) : y;
                    }
                }
                stg_stream(po + i, pack<T>(vo));
            }
        }
        i0 += kStep;
        if (i0 - r >= nv) break;                          // CTA-uniform bound (r < TPI <= kStep)
        issue(i0);
    }
}

// =============================================================================================
// Shared-memory-resident variant: ONE item = reduce AND apply of I instances, the planes stay in shared
// memory in between, so every byte crosses L2 exactly once in each direction (forward 2*S, backward 3*S of
// HBM traffic AND of L2 traffic; the L2-resident variant above pays a second L2 read, and the L2 slices --
// not HBM -- are what saturates first on B200: profiles/README.md).
//
//   CTA(ticket t): channel c = t / nI (channel-major), instances j*I .. j*I+I-1
//     1. cp.async.bulk the planes (x [, dy]) into shared memory          (TMA, whole item in flight at once)
//     2. per-instance reduction out of shared memory (forward: exact two-pass mean / variance), published as
//        ONE aligned 8-byte word per instance into a sentinel-filled [C][N] area ("data is the flag": no
//        fence, no counter)
//     3. the CTA holding the channel's LAST ticket polls the channel's N words, folds them and publishes the
//        channel constants as one 8-byte word; every other CTA of the channel polls that word
//     4. rebuild the gate / backward coefficients, apply out of shared memory, stream the result out
//
// Deadlock freedom: tickets are handed out in increasing order, so the co-resident CTAs always include the
// lowest unfinished tickets; a channel's nI items are consecutive tickets, so as long as the GPU can hold nI
// CTAs at once (checked on the host with the occupancy API) the lowest unfinished channel is always completely
// resident, none of its CTAs waits before it has published, and it completes.  All N planes of a channel (and
// the next few) live in shared memory across the GPU: N*M*sizeof(T)*tensors must fit a fraction of 148 x 227 KB.

// DYG (backward only): dy is NOT staged in shared memory -- it is streamed from global memory twice (reduce:
// L2 hit thanks to the predecessor's prefetch, marked evict-last; apply: L2 hit, evict-first) while x stays
// resident.  Twice the instances fit on chip, at 7 L2 transactions per byte of S instead of 6 (both planes
// resident) or 8 (k_sn_flow).  For channels whose x AND dy do not fit the GPU's shared memory.
template <typename T, bool BWD, bool ADD, bool DYG, int TPI, int TH>
__device__ __forceinline__ void sn_res_item(const FArgs& a, const unsigned t, const unsigned par) {
    static_assert(!(BWD && ADD), "the fused add is a forward feature");
    static_assert(BWD || !DYG, "DYG is a backward variant");
    constexpr bool two = (BWD && !DYG) || ADD;               // a second plane per instance in shared memory: dy / res
    constexpr int kB = 4;                                    // DYG: 128-bit loads of dy in flight per thread
    constexpr int I = TH / TPI;
    constexpr int V = VecOf<T>::n;
    extern __shared__ __align__(128) unsigned char dsm[];    // [mbarrier | I planes of x | I planes of dy]
    __shared__ float2 s_chan;
    __shared__ float s_f[2][TH / 32];
    uint64_t* bar = reinterpret_cast<uint64_t*>(dsm);
    const unsigned nI = (unsigned)a.nI;
    const unsigned c = t / nI, j = t - c * nI;
    const int N = a.N, C = a.C, M = a.M;
    CNSN_FTRACE(0);                                          // 0 ticket taken
    const int n = (int)j * I + (int)(threadIdx.x / TPI);
    const int r = threadIdx.x % TPI;
    const bool live = n < N;
    const size_t nc = (size_t)(live ? n : 0) * C + c;
    const int nv = M / V;
    const unsigned pbytes = (unsigned)M * (unsigned)sizeof(T);
    const uint32_t sx = smem_u32(dsm) + 128u + (threadIdx.x / TPI) * pbytes;
    const uint32_t sdy = sx + (unsigned)I * pbytes;
    if (threadIdx.x < 32) {                                  // lane q fetches instance q of the item
        const int first = (int)j * I;
        const int nlive = min(I, N - first);
        const uint64_t pol = l2_policy_evict_first();        // read once: do not keep it in L2
        const T* second = static_cast<const T*>(BWD ? a.dy : a.res);
        if (threadIdx.x == 0) mbar_arrive_expect_tx(bar, (unsigned)nlive * pbytes * (two ? 2u : 1u));
        __syncwarp();
        for (int q = threadIdx.x; q < nlive; q += 32) {
            const size_t off = ((size_t)(first + q) * C + c) * M;
            unsigned char* dst = dsm + 128 + (size_t)q * pbytes;
            tma_load_1d(dst, static_cast<const T*>(a.x) + off, pbytes, bar, pol);
            if (two) tma_load_1d(dst + (size_t)I * pbytes, second + off, pbytes, bar, pol);   // not for DYG
        }
        // L2 prefetch for the CTA that will take this one's place: its TMA loads then hit L2 instead of paying
        // the HBM latency (and its tail) while its shared memory is already tied up.  Same L2 traffic.
        const unsigned tf = t + (unsigned)a.pf_dist;
        if (a.pf_dist && tf < a.items) {
            const unsigned cf = tf / nI, jf = tf - cf * nI;
            const int ff = (int)jf * I, nf = min(I, N - ff);
            for (int q = threadIdx.x; q < nf; q += 32) {
                const size_t off = ((size_t)(ff + q) * C + cf) * M;
                tma_prefetch_l2(static_cast<const T*>(a.x) + off, pbytes);
                if (two || DYG) tma_prefetch_l2(second + off, pbytes);
            }
        }
    }
    // everything that does not depend on the channel is fetched under the TMA latency
    const bool folder = j == nI - 1;                         // holds the channel's last ticket
    float pre_g = 0.f, pre_s = 0.f, p_w0, p_w1, p_ga, p_b, p_mu = 0.f, p_sd = 1.f, p_rm = 0.f, p_rv = 1.f;
    p_w0 = a.w[2 * c]; p_w1 = a.w[2 * c + 1]; p_ga = a.gamma[c];
    if (BWD) {
        p_b = a.r[c];
        if (live) { pre_g = a.gate[nc]; pre_s = a.shat[nc]; p_mu = a.mu[nc]; p_sd = a.sd[nc]; }
    } else {
        p_b = a.beta[c];
        if (folder && threadIdx.x == 0) { p_rm = a.run_mean[c]; p_rv = a.run_var[c]; }
    }
    // DYG: dy comes from global memory; the first batch is issued before the planes of x have landed
    const uint4* gdy = DYG ? reinterpret_cast<const uint4*>(static_cast<const T*>(a.dy) + nc * M) : nullptr;
    uint4 rdy[kB];
    const uint64_t pol_keep = l2_policy_evict_last(), pol_once = l2_policy_evict_first();
    auto issue_dy = [&](int i0, uint64_t pol) {
#pragma unroll
        for (int u = 0; u < kB; ++u) {
            const int i = i0 + u * TPI;
            if (live && i < nv) rdy[u] = ldg_hint(gdy + i, pol);
        }
    };
    if (DYG) issue_dy(r, pol_keep);
    mbar_wait(bar, par, a.err);
    CNSN_FTRACE(1);                                          // 1 planes landed

    // ---- reduce out of shared memory -----------------------------------------------------------
    const bool relu = a.relu != 0;
    float own_x = 0.f, own_y = 0.f;                          // this instance's published word
    if (BWD && DYG) {
        float s0 = 0.f, s1 = 0.f;
        for (int i0 = r;;) {
#pragma unroll
            for (int u = 0; u < kB; ++u) {
                const int i = i0 + u * TPI;
                if (live && i < nv) {
                    float vx[V], vd[V];
                    unpack<T>(lds128(sx + 16u * i), vx);
                    unpack<T>(rdy[u], vd);
#pragma unroll
                    for (int e = 0; e < V; ++e) {
                        const float d = (relu && !(vx[e] > 0.f)) ? 0.f : vd[e];
                        if (e & 1) s1 = fmaf(d, vx[e], s1); else s0 = fmaf(d, vx[e], s0);
                    }
                }
            }
            i0 += TPI * kB;
            if (i0 - r >= nv) break;
            issue_dy(i0, pol_keep);
        }
        const float sxy = team_sum<TPI>(s0 + s1, s_f[0]);
        own_x = sxy * pre_g * (1.f - pre_g); own_y = pre_s;
    } else if (BWD) {
        float s0 = 0.f, s1 = 0.f;
        if (live) {
#pragma unroll 4
            for (int i = r; i < nv; i += TPI) {
                float vx[V], vd[V];
                unpack<T>(lds128(sx + 16u * i), vx);
                unpack<T>(lds128(sdy + 16u * i), vd);
#pragma unroll
                for (int e = 0; e < V; ++e) {
                    const float d = (relu && !(vx[e] > 0.f)) ? 0.f : vd[e];
                    if (e & 1) s1 = fmaf(d, vx[e], s1); else s0 = fmaf(d, vx[e], s0);
                }
            }
        }
        const float sxy = team_sum<TPI>(s0 + s1, s_f[0]);
        own_x = sxy * pre_g * (1.f - pre_g); own_y = pre_s;
    } else {                                                 // exact two-pass: the plane is on chip
        float s0 = 0.f, s1 = 0.f;
        if (live) {
            uint4* pz = ADD ? reinterpret_cast<uint4*>(static_cast<T*>(a.zout) + nc * M) : nullptr;
#pragma unroll 4
            for (int i = r; i < nv; i += TPI) {
                float vx[V];
                unpack<T>(lds128(sx + 16u * i), vx);
                if (ADD) {                                   // z = x + res: kept in place of x, and written out
                    float vr[V];
                    unpack<T>(lds128(sdy + 16u * i), vr);
#pragma unroll
                    for (int e = 0; e < V; ++e) vx[e] += vr[e];
                    const uint4 z = pack<T>(vx);
                    sts128(sx + 16u * i, z);                 // thread-private slots: no barrier needed
                    stg_stream(pz + i, z);
                    unpack<T>(z, vx);                        // statistics of the rounded sum
                }
#pragma unroll
                for (int e = 0; e < V; ++e) { if (e & 1) s1 += vx[e]; else s0 += vx[e]; }
            }
        }
        const float mean = team_sum<TPI>(s0 + s1, s_f[0]) * (1.f / M);
        s0 = s1 = 0.f;
        if (live) {
#pragma unroll 4
            for (int i = r; i < nv; i += TPI) {
                float vx[V];
                unpack<T>(lds128(sx + 16u * i), vx);
#pragma unroll
                for (int e = 0; e < V; ++e) { const float d = vx[e] - mean; if (e & 1) s1 = fmaf(d, d, s1); else s0 = fmaf(d, d, s0); }
            }
        }
        const float m2 = team_sum<TPI>(s0 + s1, s_f[1]);
        own_x = mean; own_y = sqrtf(m2 / (M - 1.f) + a.eps);
        if (live && r == 0) { a.mu[nc] = own_x; a.sd[nc] = own_y; }
    }
    // eval mode: the channel constants are the running statistics (forward) / vanish (backward, k1 = k2 = 0):
    // nothing to wait for -- a single pass per instance.  Backward still publishes (parameter gradients).
    const bool batch_coupled = a.training != 0;
    if (live && r == 0 && (BWD || batch_coupled)) ll_publish(a.pub + (size_t)c * N + n, own_x, own_y);
    CNSN_FTRACE(2);                                          // 2 reduced + published

    // ---- channel constants ------------------------------------------------------------------------
    float2* flag = a.chan + 4u * c;                          // one 32-byte sector per channel
    if (!batch_coupled && !(BWD && folder)) {
        if (threadIdx.x == 0) {
            if (BWD) {
                s_chan = make_float2(0.f, 0.f);
            } else {
                const float rstd = 1.f / sqrtf(a.run_var[c] + a.bn_eps);
                s_chan = make_float2(a.run_mean[c], rstd);
                if (folder) a.r[c] = rstd;
            }
        }
    } else if (folder) {
        const float2 cst = fold_publish<BWD, TH>(a, c, flag, p_w0, p_w1, p_ga, p_b, p_rm, p_rv, s_f);
        CNSN_FTRACE(3);                                      // 3 (folder) channel folded
        if (threadIdx.x == 0) s_chan = cst;
    } else if (threadIdx.x == 0) {
        s_chan = poll_word(flag, a.poll_ns, a.err);
    }
    __syncthreads();
    CNSN_FTRACE(4);                                          // 4 channel constants known

    // ---- apply out of shared memory -------------------------------------------------------------
    if (!live) return;
    const float2 cm = s_chan;
    float ca = 0.f, cb = 0.f, cc = 0.f;                       // out = ca*dy + cb*x + cc
    if (BWD) {
        const float ds = p_b * (own_x * p_ga - cm.x - own_y * cm.y);
        ca = pre_g;
        cb = ds * p_w1 * (1.f / (M - 1.f)) / p_sd;
        cc = ds * p_w0 * (1.f / M) - cb * p_mu;
    } else {
        const float sh = (fmaf(p_w0, own_x, p_w1 * own_y) - cm.x) * cm.y;
        const float gt = 1.f / (1.f + expf(-fmaf(p_ga, sh, p_b)));
        if (live && r == 0) { a.gate[nc] = gt; a.shat[nc] = sh; }
        cb = gt;
    }
    uint4* po = reinterpret_cast<uint4*>(static_cast<T*>(a.out) + nc * M);
    if (DYG) {                                               // second read of dy: L2 hit, last use
        issue_dy(r, pol_once);
        for (int i0 = r;;) {
#pragma unroll
            for (int u = 0; u < kB; ++u) {
                const int i = i0 + u * TPI;
                if (i < nv) {
                    float vx[V], vd[V], vo[V];
                    unpack<T>(lds128(sx + 16u * i), vx);
                    unpack<T>(rdy[u], vd);
#pragma unroll
                    for (int e = 0; e < V; ++e) {
                        const float d = (relu && !(vx[e] > 0.f)) ? 0.f : vd[e];
                        vo[e] = fmaf(ca, d, fmaf(cb, vx[e], cc));
                    }
                    stg_stream(po + i, pack<T>(vo));
                }
            }
            i0 += TPI * kB;
            if (i0 - r >= nv) break;
            issue_dy(i0, pol_once);
        }
    } else {
#pragma unroll 4
        for (int i = r; i < nv; i += TPI) {
            float vx[V], vd[V], vo[V];
            unpack<T>(lds128(sx + 16u * i), vx);
            if (BWD) unpack<T>(lds128(sdy + 16u * i), vd);
#pragma unroll
            for (int e = 0; e < V; ++e) {
                if (BWD) {
                    const float d = (relu && !(vx[e] > 0.f)) ? 0.f : vd[e];
                    vo[e] = fmaf(ca, d, fmaf(cb, vx[e], cc));
                } else {
                    const float y = fmaf(cb, vx[e], 0.f);
                    vo[e] = relu ? fmaxf(y, 0.f) : y;
                }
            }
            stg_stream(po + i, pack<T>(vo));
        }
    }
    CNSN_FTRACE(5);                                          // 5 applied
}

template <typename T, bool BWD, bool ADD, bool DYG, int TPI, int TH>
__global__ void __launch_bounds__(TH) k_sn_res(const FArgs a) {
    CNSN_TICKET_LOOP(a, (sn_res_item<T, BWD, ADD, DYG, TPI, TH>(a, t, it & 1u)))
}

// =============================================================================================
// Channel-group variant of the shared-memory-resident kernel, for planes that are NOT a multiple of 16 bytes
// (7x7 fp32 = 196 B, 14x14 bf16 = 392 B, 7x7 bf16 = 98 B: the last two stages of ResNet-50).  kk adjacent
// channels of one sample are contiguous in NCHW and kk*M*sizeof(T) IS a multiple of 16 for some kk in {2,4,8}:
// that run (a "super-plane") is what TMA fetches and what the apply phase streams out with 128-bit accesses,
// looking up each element's channel in a per-instance coefficient table; the per-instance reductions read
// shared memory element-wise.  An item = I samples x kk channels = 128 / TPI instances (TPI = 1, 2 or 4 threads
// each, chosen so that an item is ~20-40 KB); tickets are group-major; the group's kk channels are folded by its
// last kk tickets, one channel each (every CTA of the group is co-resident, see above).
constexpr int kGrpT = 128;

template <typename T, bool BWD, bool ADD, int TPI>
__device__ __forceinline__ void sn_grp_item(const FArgs& a, const unsigned t, const unsigned par) {
    static_assert(!(BWD && ADD), "the fused add is a forward feature");
    constexpr int TH = kGrpT, P = kGrpT / TPI;
    constexpr int V = VecOf<T>::n;
    constexpr bool two = BWD || ADD;
    extern __shared__ __align__(128) unsigned char dsm[];    // [mbarrier | I super-planes of x | I of dy / res]
    __shared__ float4 s_coef[P];                             // per instance: out = .x*dy + .y*x + .z
    __shared__ float s_f[2][TH / 32];
    uint64_t* bar = reinterpret_cast<uint64_t*>(dsm);
    const unsigned nI = (unsigned)a.nI;
    const unsigned g = t / nI, j = t - g * nI;
    CNSN_FTRACE(0);                                          // 0 ticket taken
    const int N = a.N, C = a.C, M = a.M, kk = a.kk, I = P / kk;
    const unsigned pbytes = (unsigned)M * (unsigned)sizeof(T), sp = (unsigned)kk * pbytes;
    const int first = (int)j * I, nlive = min(I, N - first);
    const uint32_t sbase = smem_u32(dsm) + 128u;
    const uint32_t soff2 = (unsigned)I * sp;                 // second tensor's region
    if (threadIdx.x < 32) {                                  // lane q fetches sample q's run of kk planes
        const uint64_t pol = l2_policy_evict_first();
        const T* second = static_cast<const T*>(BWD ? a.dy : a.res);
        if (threadIdx.x == 0) mbar_arrive_expect_tx(bar, (unsigned)nlive * sp * (two ? 2u : 1u));
        __syncwarp();
        for (int q = threadIdx.x; q < nlive; q += 32) {
            const size_t off = ((size_t)(first + q) * C + (size_t)g * kk) * M;
            unsigned char* dst = dsm + 128 + (size_t)q * sp;
            tma_load_1d(dst, static_cast<const T*>(a.x) + off, sp, bar, pol);
            if (two) tma_load_1d(dst + soff2, second + off, sp, bar, pol);
        }
        const unsigned tf = t + (unsigned)a.pf_dist;
        if (a.pf_dist && tf < a.items) {
            const unsigned gf = tf / nI, jf = tf - gf * nI;
            const int ff = (int)jf * I, nf = min(I, N - ff);
            for (int q = threadIdx.x; q < nf; q += 32) {
                const size_t off = ((size_t)(ff + q) * C + (size_t)gf * kk) * M;
                tma_prefetch_l2(static_cast<const T*>(a.x) + off, sp);
                if (two) tma_prefetch_l2(second + off, sp);
            }
        }
    }
    // this team's instance
    const int team = threadIdx.x / TPI, r = threadIdx.x % TPI;
    const int q = team / kk, cl = team - q * kk;
    const int n = first + q;
    const bool live = q < nlive;
    const unsigned c = g * (unsigned)kk + (unsigned)cl;
    const size_t nc = (size_t)(live ? n : 0) * C + c;
    const uint32_t sx = sbase + (unsigned)q * sp + (unsigned)cl * pbytes;
    const uint32_t sdy = sx + soff2;
    const bool folder = j == nI - 1;
    const bool relu = a.relu != 0;
    const bool batch_coupled = a.training != 0;
    float pre_g = 0.f, pre_s = 0.f, p_w0, p_w1, p_ga, p_b, p_mu = 0.f, p_sd = 1.f;
    p_w0 = a.w[2 * c]; p_w1 = a.w[2 * c + 1]; p_ga = a.gamma[c];
    if (BWD) {
        p_b = a.r[c];
        if (live) { pre_g = a.gate[nc]; pre_s = a.shat[nc]; p_mu = a.mu[nc]; p_sd = a.sd[nc]; }
    } else {
        p_b = a.beta[c];
    }
    mbar_wait(bar, par, a.err);
    CNSN_FTRACE(1);                                          // 1 planes landed
    const int vps = (int)(sp / 16u), nvec = nlive * vps;     // 128-bit vectors per super-plane / in the item
    // Flat walk over the item's 128-bit vectors, division-free: thread tid takes vectors tid, tid + TH, ...; (sample qv,
    // vector w within its super-plane) advance incrementally, the plane of an element comes from a multiply-high
    // (integer divisions by runtime values cost ~30 instructions each and made this walk the longest phase of an item:
    // 3.8 us of 12.8, gpurun_out/r3d_trace_grp_fwd.log).
    const int vq0 = (int)threadIdx.x / vps, vw0 = (int)threadIdx.x - vq0 * vps;       // one division per thread and item
    const int vdq = TH / vps, vdw = TH - vdq * vps;
    const unsigned magicM = 0xffffffffu / (unsigned)M + 1u;  // floor(e / M) = umulhi(e, magicM) for e * M < 2^32
    const size_t srow = (size_t)C * M;                       // elements between consecutive samples
    const size_t gbase = ((size_t)first * C + (size_t)g * kk) * M;
    if (ADD) {                                               // z = x + res over the item, in place and written out
        T* zb = static_cast<T*>(a.zout) + gbase;
        int qv = vq0, w = vw0;
        for (int vi = threadIdx.x; vi < nvec; vi += TH) {
            float vx[V], vr[V];
            unpack<T>(lds128(sbase + 16u * vi), vx);
            unpack<T>(lds128(sbase + soff2 + 16u * vi), vr);
#pragma unroll
            for (int e = 0; e < V; ++e) vx[e] += vr[e];
            const uint4 z = pack<T>(vx);
            sts128(sbase + 16u * vi, z);
            stg_stream(reinterpret_cast<uint4*>(zb + (size_t)qv * srow) + w, z);
            qv += vdq; w += vdw;
            if (w >= vps) { w -= vps; ++qv; }
        }
        __syncthreads();
    }
    // ---- per-instance reduction, element-wise out of shared memory ---------------------------------
    float own_x = 0.f, own_y = 0.f;
    if (BWD) {
        float s0 = 0.f;
        if (live)
            for (int e = r; e < M; e += TPI) {
                const float x = lds_elem<T>(sx, e);
                const float d = (relu && !(x > 0.f)) ? 0.f : lds_elem<T>(sdy, e);
                s0 = fmaf(d, x, s0);
            }
        const float sxy = team_sum<TPI>(s0, s_f[0]);
        own_x = sxy * pre_g * (1.f - pre_g); own_y = pre_s;
    } else {
        float s0 = 0.f;
        if (live) for (int e = r; e < M; e += TPI) s0 += lds_elem<T>(sx, e);
        const float mean = team_sum<TPI>(s0, s_f[0]) * (1.f / M);
        s0 = 0.f;
        if (live) for (int e = r; e < M; e += TPI) { const float d = lds_elem<T>(sx, e) - mean; s0 = fmaf(d, d, s0); }
        const float m2 = team_sum<TPI>(s0, s_f[1]);
        own_x = mean; own_y = sqrtf(m2 / (M - 1.f) + a.eps);
        if (live && r == 0) { a.mu[nc] = own_x; a.sd[nc] = own_y; }
    }
    if (live && r == 0 && (BWD || batch_coupled)) ll_publish(a.pub + (size_t)c * N + n, own_x, own_y);
    CNSN_FTRACE(2);                                          // 2 reduced + published
    // ---- channel constants ------------------------------------------------------------------------
    if (BWD || batch_coupled) {                              // channel k of the group: folded by ticket nI-1-(k mod nI)
        for (int k = (int)(nI - 1u - j); k < kk; k += (int)nI) {
            const unsigned ch = g * (unsigned)kk + (unsigned)k;
            const float w0 = a.w[2 * ch], w1 = a.w[2 * ch + 1], ga = a.gamma[ch];
            const float pb2 = BWD ? a.r[ch] : 0.f;
            float rm = 0.f, rv = 1.f;
            if (!BWD && threadIdx.x == 0) { rm = a.run_mean[ch]; rv = a.run_var[ch]; }
            fold_publish<BWD, TH>(a, ch, a.chan + 4u * ch, w0, w1, ga, pb2, rm, rv, s_f);
        }
    }
    CNSN_FTRACE(3);                                          // 3 this item's folds (if any) done
    float2 cm;
    if (batch_coupled) {
        cm = poll_word(a.chan + 4u * c, a.poll_ns, a.err);          // 4 lanes per address; the folder's own words are there
    } else if (BWD) {
        cm = make_float2(0.f, 0.f);
    } else {
        const float rstd = 1.f / sqrtf(a.run_var[c] + a.bn_eps);
        cm = make_float2(a.run_mean[c], rstd);
        if (folder && q == 0 && r == 0) a.r[c] = rstd;
    }
    CNSN_FTRACE(4);                                          // 4 (thread 0's) channel constants known
    if (live && r == 0) {
        if (BWD) {
            const float ds = p_b * (own_x * p_ga - cm.x - own_y * cm.y);
            const float cb = ds * p_w1 * (1.f / (M - 1.f)) / p_sd;
            s_coef[team] = make_float4(pre_g, cb, ds * p_w0 * (1.f / M) - cb * p_mu, 0.f);
        } else {
            const float sh = (fmaf(p_w0, own_x, p_w1 * own_y) - cm.x) * cm.y;
            const float gt = 1.f / (1.f + expf(-fmaf(p_ga, sh, p_b)));
            a.gate[nc] = gt; a.shat[nc] = sh;
            s_coef[team] = make_float4(0.f, gt, 0.f, 0.f);
        }
    }
    __syncthreads();
    // ---- apply: the item's super-planes as flat 128-bit vectors ------------------------------------------
    {
        T* ob = static_cast<T*>(a.out) + gbase;
        int qv = vq0, w = vw0;
        for (int vi = threadIdx.x; vi < nvec; vi += TH) {
            const int e0 = w * V, cl0 = (int)__umulhi((unsigned)e0, magicM), bound = (cl0 + 1) * M;
            const float4 k0 = s_coef[qv * kk + cl0];
            const float4 k1 = (cl0 + 1 < kk) ? s_coef[qv * kk + cl0 + 1] : k0;     // a vector spans at most two planes (M >= V)
            float vx[V], vd[V], vo[V];
            unpack<T>(lds128(sbase + 16u * vi), vx);
            if (BWD) unpack<T>(lds128(sbase + soff2 + 16u * vi), vd);
#pragma unroll
            for (int e = 0; e < V; ++e) {
                const float4 k = (e0 + e < bound) ? k0 : k1;
                if (BWD) {
                    const float d = (relu && !(vx[e] > 0.f)) ? 0.f : vd[e];
                    vo[e] = fmaf(k.x, d, fmaf(k.y, vx[e], k.z));
                } else {
                    const float y = fmaf(k.y, vx[e], 0.f);
                    vo[e] = relu ? fmaxf(y, 0.f) : y;
                }
            }
            stg_stream(reinterpret_cast<uint4*>(ob + (size_t)qv * srow) + w, pack<T>(vo));
            qv += vdq; w += vdw;
            if (w >= vps) { w -= vps; ++qv; }
        }
    }
    CNSN_FTRACE(5);                                          // 5 applied
}

template <typename T, bool BWD, bool ADD, int TPI>
__global__ void __launch_bounds__(kGrpT) k_sn_grp(const FArgs a) {
    CNSN_TICKET_LOOP(a, (sn_grp_item<T, BWD, ADD, TPI>(a, t, it & 1u)))
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
// Threads per instance: a power of two that covers the plane in about kBatches batches of kU loads.  Fewer
// threads per instance amortise the per-item latencies (ticket, flag, fence) over more bytes; more threads
// keep the resident window (and with it the L2 footprint between the two reads) narrow.  Measured on
// (256,256,56,56) fp32: TPI 64 (4 batches of 4 loads) is the optimum; 6 batches is the rule that reproduces the
// measured optimum on the other shapes as well (profiles/README.md).
static int pick_tpi(int nv) {
    const int batches = knobs().flow_batches;
    int tpi = 8;
    while (tpi < kT && tpi * kUReg * batches < nv) tpi <<= 1;
    return tpi;
}

template <bool BWD>
static int launch(FArgs& a, int dtype, float* scratch, cudaStream_t stream) {
    const int N = a.N, C = a.C;
    const int esz = (int)esize(dtype);
    if (((size_t)a.M * esz) % 16) return -100;
    if (N < 1 || C < 1) return -100;
    const int nv = a.M * esz / 16;
    int tpi = pick_tpi(nv);
    const Knobs& kn = knobs();
    if (const int v = kn.flow_tpi) { if (v >= 8 && v <= kT && (v & (v - 1)) == 0) tpi = v; }
    const int I = kT / tpi;
    a.nI = (N + I - 1) / I;
    // Look-ahead: enough channels to cover the R items in flight plus the fold latency, bounded by L2.
    const size_t chan_bytes = (size_t)N * a.M * esz * (BWD ? 2 : 1);
    int D = (int)(((size_t)kn.lookahead_mb << 20) / (chan_bytes ? chan_bytes : 1));
    if (D < 2) D = 2;
    if (const int v = kn.flow_d) D = v;
    if (D > C) D = C;
    if (D < 1) D = 1;
    a.D = D;
    a.keep = kn.keep;
    a.pf_dist = kn.rpf;                                    // R items: L2 prefetch distance in channels (0 = off)
    a.err = async_error_word();
    const unsigned long long items = 2ull * C * a.nI;
    if (items > 0x7fffffffull) return -100;
    // scratch: pub [C][N] float2 | chan [C] float2 | done [C] | ready [C] | ticket
    a.pub = reinterpret_cast<float2*>(scratch);
    a.chan = a.pub + (size_t)N * C;
    a.done = reinterpret_cast<unsigned*>(a.chan + C);
    a.ready = a.done + C;
    a.ticket = a.ready + C;
    cudaError_t e = cudaMemsetAsync(a.done, 0, (2 * (size_t)C + 1) * sizeof(unsigned), stream);
    if (e != cudaSuccess) return (int)e;
    const dim3 grid((unsigned)items), block(kT);
#define CNSN_FLOW_CASE(TPI_)                                                                 \
    case TPI_:                                                                               \
        if (!BWD && a.res) k_sn_flow<T, false, true, TPI_><<<grid, block, 0, stream>>>(a);   \
        else k_sn_flow<T, BWD, false, TPI_><<<grid, block, 0, stream>>>(a);                  \
        break;
    CNSN_DISPATCH_DTYPE(dtype, T, switch (tpi) {
        CNSN_FLOW_CASE(8) CNSN_FLOW_CASE(16) CNSN_FLOW_CASE(32) CNSN_FLOW_CASE(64) CNSN_FLOW_CASE(128) CNSN_FLOW_CASE(256)
        default: return -100;
    });
#undef CNSN_FLOW_CASE
    if (kn.debug)
        fprintf(stderr, "[cnsn flow] %s tpi=%d I=%d nI=%d D=%d items=%llu add=%d relu=%d\n", BWD ? "bwd" : "fwd", tpi, I,
                a.nI, D, items, a.res != nullptr, a.relu);
    return launch_status();
}


// Debug only (cnsn_tune("trace", 1) + $CNSN_FLOW_TRACE = output path): per-item globaltimer stamps, written synchronously.
static const char* trace_begin(FArgs& a, unsigned long long items, cudaStream_t stream) {
    a.trace = nullptr;
    const char* path = knobs().trace ? getenv("CNSN_FLOW_TRACE") : nullptr;
    const size_t bytes = (size_t)items * 8 * sizeof(unsigned long long);
    if (path && cudaMalloc(&a.trace, bytes) == cudaSuccess) cudaMemsetAsync(a.trace, 0, bytes, stream);
    return path;
}
static void trace_end(FArgs& a, const char* path, unsigned long long items, int per_sm, bool bwd, cudaStream_t stream) {
    if (!a.trace) return;
    const size_t bytes = (size_t)items * 8 * sizeof(unsigned long long);
    cudaStreamSynchronize(stream);
    unsigned long long* h = (unsigned long long*)malloc(bytes);
    cudaMemcpy(h, a.trace, bytes, cudaMemcpyDeviceToHost);
    if (FILE* f = fopen(path, "wb")) {
        const int hdr[4] = {(int)items, a.nI, per_sm, bwd ? 1 : 0};
        fwrite(hdr, sizeof(int), 4, f);
        fwrite(h, 1, bytes, f);
        fclose(f);
    }
    free(h);
    cudaFree(a.trace);
    a.trace = nullptr;
}

constexpr int kResT = 128;              // threads per CTA of the shared-memory-resident kernel

// Forward items of THREE planes and 192 threads (64 per plane) where the generic geometry would be two planes per
// 128-thread item: at the north-star plane size (12.5 KB) an SM holds 8 items x 2 planes = 16 planes and cannot take a
// ninth item; 6 items x 3 planes = 18 planes fit the same 228 KB -- 12.5 % more bytes on chip per SM, measured 0.322 ->
// 0.312 ms at (256,256,56,56) fp32 (gpurun_out/r2a_sweep.log).  Same kernel template (it is generic in TH and TPI).
// cnsn_tune("i3", 0) selects the two-plane geometry (A/B, tests).
static int launch_res_i3(FArgs& a, int dtype, float* scratch, cudaStream_t stream) {
    constexpr int kT3 = 192, kTpi3 = 64, kI3 = kT3 / kTpi3;
    const int N = a.N, C = a.C;
    const size_t pbytes = (size_t)a.M * esize(dtype);
    if (pbytes % 16 || N < kI3 || pbytes * 2 > (25u << 10) + 512 || pbytes * 4 <= (25u << 10) + 512) return -100;
    const size_t dsmem = 128 + kI3 * pbytes;
    const DeviceShape ds = device_shape();
    a.nI = (N + kI3 - 1) / kI3;
    a.D = 0;
    const Knobs& kn = knobs();
    const unsigned long long items = (unsigned long long)C * a.nI;
    if (items > 0x7fffffffull) return -100;
    a.pub = reinterpret_cast<float2*>(scratch);
    a.chan = a.pub + (size_t)N * C;
    a.ticket = reinterpret_cast<unsigned*>(a.chan + 4 * (size_t)C);
    a.done = nullptr; a.ready = nullptr; a.trace = nullptr;
    a.poll_ns = kn.poll_ns;
    a.items = (unsigned)items;
    a.err = async_error_word();
    const size_t fill_bytes = ((size_t)N * C + 4 * (size_t)C + 1) * sizeof(float2);
    cudaError_t e = cudaSuccess;
    int per_sm = 0;
    CNSN_DISPATCH_DTYPE(dtype, T, {
        auto fn = k_sn_res<T, false, false, false, kTpi3, kT3>;
        e = prepare_kernel(fn, kT3, dsmem, &per_sm);
        if (e != cudaSuccess) return (int)e;
        if ((long long)per_sm * ds.sms < 2ll * a.nI) return -100;
        // only where it really puts more planes on an SM than the two-plane geometry does (12.5 KB planes: 6 x 3 = 18
        // against 8 x 2 = 16; 12.8 KB bf16 80x80 planes: 5 x 3 = 15 against 8 x 2 = 16 -- measured slower, r2b_sweep.log)
        int per_sm2 = 0;
        e = prepare_kernel(k_sn_res<T, false, false, false, 64, 128>, 128, 128 + 2 * pbytes, &per_sm2);
        if (e != cudaSuccess) return (int)e;
        if (3 * per_sm <= 2 * per_sm2) return -100;
        a.pf_dist = kn.pf >= 0 ? kn.pf : per_sm * ds.sms / 2;
        e = cudaMemsetAsync(a.pub, 0xff, fill_bytes, stream);
        if (e != cudaSuccess) return (int)e;
        e = launch_persistent(fn, a, a.items, (unsigned)a.nI, per_sm, ds.sms, kT3, dsmem, stream);
        if (e == cudaErrorCooperativeLaunchTooLarge) { (void)cudaGetLastError(); return -100; }
        if (e != cudaSuccess) return (int)e;
    });
    if (kn.debug)
        fprintf(stderr, "[cnsn flow/res-i3] fwd tpi=%d I=%d nI=%d items=%llu smem=%zu ctas/sm=%d\n", kTpi3, kI3, a.nI, items, dsmem, per_sm);
    return launch_status();
}

// Shared-memory-resident path.  Returns -100 when the shape does not fit (the caller uses the L2 path).
template <bool BWD>
static int launch_res(FArgs& a, int dtype, float* scratch, cudaStream_t stream, bool dy_from_global = false) {
    const int N = a.N, C = a.C;
    const int esz = (int)esize(dtype);
    if (((size_t)a.M * esz) % 16 || N < 1 || C < 1) return -100;
    const bool add = !BWD && a.res != nullptr;
    const bool dyg = BWD && dy_from_global;
    const Knobs& kn = knobs();
    if (!BWD && !add && kn.i3) {
        const int rc3 = launch_res_i3(a, dtype, scratch, stream);
        if (rc3 != -100) return rc3;
    }
    const size_t inst_bytes = (size_t)a.M * esz * (((BWD && !dyg) || add) ? 2 : 1);
    const size_t target = (size_t)kn.item_kb << 10;
    int inst = 1;
    while (inst < 16 && (size_t)(2 * inst) * inst_bytes <= target + 512 && 2 * inst <= N) inst <<= 1;
    if (const int v = kn.flow_tpi) { if (v >= 8 && v <= kResT && (v & (v - 1)) == 0) inst = kResT / v; }
    const int tpi = kResT / inst;
    const size_t dsmem = 128 + (size_t)inst * inst_bytes;
    const DeviceShape ds = device_shape();
    const int sms = ds.sms;
    if (dsmem > (size_t)ds.smem_optin / 2) return -100;      // at least two CTAs per SM
    a.nI = (N + inst - 1) / inst;
    a.D = 0;
    const unsigned long long items = (unsigned long long)C * a.nI;
    if (items > 0x7fffffffull) return -100;
    // scratch: pub [C][N] float2 | channel words [C] x 4 float2 (one 32-byte sector each) | ticket.  Everything is
    // pre-filled with 0xff: the sentinel of the 8-byte words, and a ticket counter that wraps to 0 on first use.
    a.pub = reinterpret_cast<float2*>(scratch);
    a.chan = a.pub + (size_t)N * C;
    a.ticket = reinterpret_cast<unsigned*>(a.chan + 4 * (size_t)C);
    a.done = nullptr;
    a.ready = nullptr;
    a.poll_ns = kn.poll_ns;
    a.items = (unsigned)items;
    a.err = async_error_word();
    const size_t fill_bytes = ((size_t)N * C + 4 * (size_t)C + 1) * sizeof(float2);
    cudaError_t e = cudaSuccess;
    int per_sm = 0;
    const char* trace_path = trace_begin(a, items, stream);
#define CNSN_RES_CASE(TPI_)                                                                              \
    case TPI_: {                                                                                         \
        auto fn = add ? k_sn_res<T, false, !BWD, false, TPI_, kResT>                                     \
                      : (dyg ? k_sn_res<T, BWD, false, BWD, TPI_, kResT> : k_sn_res<T, BWD, false, false, TPI_, kResT>); \
        e = prepare_kernel(fn, kResT, dsmem, &per_sm);                                                   \
        if (e != cudaSuccess) return (int)e;                                                             \
        /* the channel being completed must be resident as a whole (deadlock freedom), with room to spare */ \
        if ((long long)per_sm * sms < 2ll * a.nI) return -100;                                           \
        a.pf_dist = kn.pf >= 0 ? kn.pf : per_sm * sms / 2;                                               \
        e = cudaMemsetAsync(a.pub, 0xff, fill_bytes, stream);                                            \
        if (e != cudaSuccess) return (int)e;                                                             \
        e = launch_persistent(fn, a, a.items, (unsigned)a.nI, per_sm, sms, kResT, dsmem, stream);                        \
        if (e == cudaErrorCooperativeLaunchTooLarge) { (void)cudaGetLastError(); return -100; }          \
        if (e != cudaSuccess) return (int)e;                                                             \
    } break;
    CNSN_DISPATCH_DTYPE(dtype, T, switch (tpi) {
        CNSN_RES_CASE(8) CNSN_RES_CASE(16) CNSN_RES_CASE(32) CNSN_RES_CASE(64) CNSN_RES_CASE(128)
        default: return -100;
    });
#undef CNSN_RES_CASE
    trace_end(a, trace_path, items, per_sm, BWD, stream);
    if (kn.debug)
        fprintf(stderr, "[cnsn flow/res] %s tpi=%d I=%d nI=%d items=%llu smem=%zu ctas/sm=%d dyg=%d\n", BWD ? "bwd" : "fwd",
                tpi, inst, a.nI, items, dsmem, per_sm, (int)dyg);
    return launch_status();
}


// Channel-group path for planes that are not a multiple of 16 bytes.  Returns -100 when it does not apply.
template <bool BWD>
static int launch_grp(FArgs& a, int dtype, float* scratch, cudaStream_t stream) {
    const int N = a.N, C = a.C;
    const int esz = (int)esize(dtype);
    const size_t pb = (size_t)a.M * esz;
    if (pb % 16 == 0 || N < 1 || C < 1 || a.M < 16 / esz * 2) return -100;
    int kk = 0;
    for (int k = 2; k <= 8; k <<= 1) if ((k * pb) % 16 == 0 && C % k == 0) { kk = k; break; }
    if (!kk) return -100;
    const bool add = !BWD && a.res != nullptr;
    const size_t ib = pb * ((BWD || add) ? 2 : 1);               // bytes per instance in shared memory
    const Knobs& kn = knobs();
    const size_t want = (size_t)kn.grp_kb << 10;                  // item size aimed at
    const int tpi = (32 * ib >= want || kk > 32) ? 4 : (64 * ib >= want || kk > 64) ? 2 : 1;
    const int I = (kGrpT / tpi) / kk;
    if (I < 1) return -100;
    const size_t dsmem = 128 + (size_t)I * kk * ib;
    const DeviceShape ds = device_shape();
    if (dsmem > (size_t)ds.smem_optin / 2) return -100;
    a.kk = kk;
    a.nI = (N + I - 1) / I;
    a.D = 0;
    a.poll_ns = kn.poll_ns;
    a.err = async_error_word();
    const unsigned long long items = (unsigned long long)(C / kk) * a.nI;
    if (items > 0x7fffffffull) return -100;
    a.items = (unsigned)items;
    a.pub = reinterpret_cast<float2*>(scratch);
    a.chan = a.pub + (size_t)N * C;
    a.ticket = reinterpret_cast<unsigned*>(a.chan + 4 * (size_t)C);
    a.done = nullptr; a.ready = nullptr;
    const char* trace_path = trace_begin(a, items, stream);
    const size_t fill_bytes = ((size_t)N * C + 4 * (size_t)C + 1) * sizeof(float2);
    cudaError_t e = cudaSuccess;
    int per_sm = 0;
    CNSN_DISPATCH_DTYPE(dtype, T, {
        auto fn = tpi == 4 ? (add ? k_sn_grp<T, false, !BWD, 4> : k_sn_grp<T, BWD, false, 4>)
                : tpi == 2 ? (add ? k_sn_grp<T, false, !BWD, 2> : k_sn_grp<T, BWD, false, 2>)
                           : (add ? k_sn_grp<T, false, !BWD, 1> : k_sn_grp<T, BWD, false, 1>);
        e = prepare_kernel(fn, kGrpT, dsmem, &per_sm);
        if (e != cudaSuccess) return (int)e;
        if ((long long)per_sm * ds.sms < 2ll * a.nI) return -100;
        a.pf_dist = kn.pf >= 0 ? kn.pf : per_sm * ds.sms / 2;
        e = cudaMemsetAsync(a.pub, 0xff, fill_bytes, stream);
        if (e != cudaSuccess) return (int)e;
        e = launch_persistent(fn, a, a.items, (unsigned)a.nI, per_sm, ds.sms, kGrpT, dsmem, stream);
        if (e == cudaErrorCooperativeLaunchTooLarge) { (void)cudaGetLastError(); return -100; }
        if (e != cudaSuccess) return (int)e;
    });
    trace_end(a, trace_path, items, per_sm, BWD, stream);
    if (kn.debug)
        fprintf(stderr, "[cnsn flow/grp] %s kk=%d tpi=%d I=%d nI=%d items=%llu smem=%zu ctas/sm=%d add=%d\n", BWD ? "bwd" : "fwd", kk,
                tpi, I, a.nI, items, dsmem, per_sm, (int)add);
    return launch_status();
}

// Which kernel: the shared-memory-resident one when a channel (all N planes, x [and dy]) is a small enough part
// of the GPU's shared memory -- measured cross-over on B200 (profiles/README.md): 1/8 of 148 x 200 KB forward,
// 1/16 backward (the L2-resident backward is the stronger alternative).  cnsn_tune("flow_mode", 1|2) forces one.
static bool use_resident(size_t chan_bytes, bool bwd) {
    if (const int m = knobs().flow_mode) return m == 1;
    const size_t on_chip = (size_t)device_shape().sms * 200 * 1024;
    return chan_bytes * (bwd ? 16 : 8) <= on_chip;
}

size_t scratch_floats(int N, int C) { return 2 * (size_t)N * C + 36 * (size_t)C + 8; }

// Both return 0 when launched, >0 cuda error, -100 when the path does not apply.
int selfnorm_flow_fwd(const void* x, const void* res, void* z, void* y, int relu, int dtype, int N, int C, int H, int W,
                      const cnsn_gate_params* g, int training, float momentum, float bn_eps, float eps,
                      float* mu, float* sd, float* gate, float* shat, float* r, float* scratch,
                      cudaStream_t stream) {
    if (!aligned16(x) || !aligned16(y) || (res && (!aligned16(res) || !aligned16(z)))) return -100;
    if (async_error_peek()) return CNSN_E_TIMEOUT;
    FArgs a{};
    a.x = x; a.dy = nullptr; a.out = y; a.N = N; a.C = C; a.M = H * W;
    a.res = res; a.zout = res ? z : nullptr; a.relu = relu;
    const bool odd = ((size_t)H * W * esize(dtype)) % 16 != 0;
    a.training = training; a.momentum = momentum; a.bn_eps = bn_eps; a.eps = eps;
    a.w = g->w; a.gamma = g->gamma; a.beta = g->beta; a.run_mean = g->run_mean; a.run_var = g->run_var; a.nbt = g->nbt;
    a.mu = mu; a.sd = sd; a.gate = gate; a.shat = shat; a.r = r;
    if (odd) return launch_grp<false>(a, dtype, scratch, stream);
    if (knobs().flow_mode == 0) {                            // default dispatch: shared + tensor memory pipeline first
        const int rc = selfnorm_tmem_fwd(a, dtype, scratch, stream);
        if (rc != -100) return rc;
    }
    if (use_resident((size_t)N * H * W * esize(dtype) * (res ? 2 : 1), false)) {
        const int rc = launch_res<false>(a, dtype, scratch, stream);
        if (rc != -100) return rc;
    }
    return launch<false>(a, dtype, scratch, stream);
}

int selfnorm_flow_bwd(const void* x, const void* dy, void* dx, int relu, int dtype, int N, int C, int H, int W,
                      const cnsn_gate_params* g, int training,
                      float* mu, float* sd, float* gate, float* shat, float* r,
                      const cnsn_gate_grads* dg, float* scratch, cudaStream_t stream) {
    if (!aligned16(x) || !aligned16(dy) || !aligned16(dx)) return -100;
    if (async_error_peek()) return CNSN_E_TIMEOUT;
    FArgs a{};
    a.x = x; a.dy = dy; a.out = dx; a.N = N; a.C = C; a.M = H * W;
    a.relu = relu;
    a.training = training;
    a.w = g->w; a.gamma = g->gamma;
    a.mu = mu; a.sd = sd; a.gate = gate; a.shat = shat; a.r = r;
    a.dw = dg->dw; a.dgamma = dg->dgamma; a.dbeta = dg->dbeta;
    // backward: both planes resident when the channel is small, L2 items otherwise.  The third variant -- x
    // resident, dy streamed through L2 (DYG) -- measured slower than both on B200 (0.50-0.61 ms against 0.476 /
    // 0.511 at the north-star shape: two more L2 round trips per item outweigh the doubled capacity); it stays
    // selectable.  cnsn_tune("flow_bwd", 1|2|3) forces one (A/B measurements).
    const size_t chan_x = (size_t)N * H * W * esize(dtype);
    if ((((size_t)H * W * esize(dtype)) % 16) != 0) return launch_grp<true>(a, dtype, scratch, stream);
    if (knobs().flow_bwd == 0 && knobs().flow_mode == 0) {
        const int rc = selfnorm_tmem_bwd(a, dtype, scratch, stream);
        if (rc != -100) return rc;
    }
    int mode = use_resident(2 * chan_x, true) ? 0 : 2;
    if (const int b = knobs().flow_bwd) mode = b - 1;          // 1 resident, 2 x resident + dy through L2, 3 L2 items
    if (mode < 2) {
        const int rc = launch_res<true>(a, dtype, scratch, stream, mode == 1);
        if (rc != -100) return rc;
    }
    return launch<true>(a, dtype, scratch, stream);
}

}  // namespace flow
}  // namespace cnsn
