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

#include "common.cuh"

namespace cnsn {
namespace flow {

constexpr int kT = 256;                 // threads per CTA
constexpr int kU = 4;                   // 128-bit loads in flight per thread and tensor
constexpr unsigned kSpin = 1u << 22;    // bounded polls (>= 64 ns each): trap instead of hanging the GPU

struct FArgs {
    const void* x; const void* dy; void* out;      // forward: dy == nullptr, out = y; backward: out = dx
    int N, C, M;
    int nI;                 // items per channel and phase = ceil(N / I)
    int D;                  // look-ahead in channels (1..C)
    int training;
    int order;              // 0: atomic ticket per CTA; 1: blockIdx.x (relies on in-order CTA dispatch)
    int keep;               // L2 policy of the R loads: 0 default, 1 evict_last (default: the A items re-read them)
    float momentum, bn_eps, eps;
    const float* w; const float* gamma; const float* beta;
    float* run_mean; float* run_var; long long* nbt;
    float* mu; float* sd; float* gate; float* shat; float* r;
    float* dw; float* dgamma; float* dbeta;
    float2* pub;            // [C][N] forward: (mu, sd); backward: (dz, shat)
    float2* chan;           // [C]    forward: (m, rstd); backward: (k1, k2)
    unsigned* done;         // [C]    R items of the channel that have published
    unsigned* ready;        // [C]    channel constants are in chan[c]
    unsigned* ticket;       // [1]
};

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
// Sums over the whole CTA (channel fold by the last R item).
template <int K>
__device__ __forceinline__ void cta_sums(float (&v)[K], float (*sm)[kT / 32]) {
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
        for (int w = 0; w < kT / 32; ++w) s += sm[k][w];
        v[k] = s;
    }
}

template <typename T, bool BWD, int TPI>
__global__ void __launch_bounds__(kT) k_sn_flow(const FArgs a) {
    constexpr int I = kT / TPI;          // instances per item
    constexpr int V = VecOf<T>::n;
    __shared__ unsigned s_word;
    __shared__ float s_f[2][kT / 32];
    __shared__ Moments s_m[kT / 32];

    // ---- which item am I --------------------------------------------------------------------
    unsigned t = blockIdx.x;
    if (a.order == 0) {
        if (threadIdx.x == 0) s_word = atomicAdd(a.ticket, 1u);
        __syncthreads();
        t = s_word;
    }
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
    const uint4* px = reinterpret_cast<const uint4*>(static_cast<const T*>(a.x) + nc * M);
    const uint4* pd = BWD ? reinterpret_cast<const uint4*>(static_cast<const T*>(a.dy) + nc * M) : nullptr;
    constexpr int kStep = TPI * kU;

    if (!isA) {
        // =============================================================== R item
        const uint64_t pol = a.keep ? l2_policy_evict_last() : 0;
        float sxy = 0.f, pre_g = 0.f, pre_s = 0.f;
        if (BWD && live && r == 0) { pre_g = a.gate[nc]; pre_s = a.shat[nc]; }   // issued ahead of the plane loads
        Moments acc = moments_zero();
        for (int i0 = r; i0 < nv; i0 += kStep) {
            uint4 rx[kU], rd[kU];
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const int i = i0 + u * TPI;
                if (live && i < nv) {
                    rx[u] = a.keep ? ldg_hint(px + i, pol) : ldg_stream(px + i);
                    if (BWD) rd[u] = a.keep ? ldg_hint(pd + i, pol) : ldg_stream(pd + i);
                }
            }
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const int i = i0 + u * TPI;
                if (live && i < nv) {
                    float vx[V];
                    unpack<T>(rx[u], vx);
                    if (BWD) {
                        float vd[V];
                        unpack<T>(rd[u], vd);
#pragma unroll
                        for (int e = 0; e < V; ++e) sxy = fmaf(vd[e], vx[e], sxy);
                    } else {
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
        const float2* pb = a.pub + (size_t)c * N;
        const float invN = 1.f / N;
        if (!BWD) {
            const float w0 = a.w[2 * c], w1 = a.w[2 * c + 1];
            float m, q;
            if (a.training) {
                float v[1] = {0.f};
                for (int k = threadIdx.x; k < N; k += kT) { const float2 p = __ldcg(pb + k); v[0] += fmaf(w0, p.x, w1 * p.y); }
                cta_sums<1>(v, s_f);
                m = v[0] / N;
                v[0] = 0.f;
                for (int k = threadIdx.x; k < N; k += kT) {
                    const float2 p = __ldcg(pb + k);
                    const float d = fmaf(w0, p.x, w1 * p.y) - m;
                    v[0] = fmaf(d, d, v[0]);
                }
                cta_sums<1>(v, s_f);
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
            for (int k = threadIdx.x; k < N; k += kT) { const float2 p = __ldcg(pb + k); v[0] = fmaf(p.x, p.y, v[0]); v[1] += p.x; }
            cta_sums<2>(v, s_f);
            const float dgam = v[0], dbet = v[1];
            const float k1 = a.training ? ga * dbet * invN : 0.f, k2 = a.training ? ga * dgam * invN : 0.f;
            v[0] = v[1] = 0.f;
            for (int k = threadIdx.x; k < N; k += kT) {
                const float2 p = __ldcg(pb + k);
                const size_t i = (size_t)k * C + c;
                const float ds = rstd * (p.x * ga - k1 - p.y * k2);
                v[0] = fmaf(ds, a.mu[i], v[0]); v[1] = fmaf(ds, a.sd[i], v[1]);
            }
            cta_sums<2>(v, s_f);
            if (threadIdx.x == 0) {
                a.dgamma[c] = dgam; a.dbeta[c] = dbet;
                a.dw[2 * c] = v[0]; a.dw[2 * c + 1] = v[1];
                a.chan[c] = make_float2(k1, k2);
            }
        }
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
                rx[u] = ldg_hint(px + i, pol);           // L2 hit (read D channels ago), last use
                if (BWD) rd[u] = ldg_hint(pd + i, pol);
            }
        }
    };
    issue(r);                                            // the plane loads do not depend on the channel
    // ... nor do the saved per-instance statistics and the parameters: fetch them under the flag wait too
    float p_w0 = 0.f, p_w1 = 0.f, p_ga = 0.f, p_b = 0.f, p_gt = 0.f, p_mu = 0.f, p_sd = 1.f;
    if (live) {
        p_w0 = a.w[2 * c]; p_w1 = a.w[2 * c + 1]; p_ga = a.gamma[c];
        if (BWD) { p_b = a.r[c]; p_gt = a.gate[nc]; p_mu = a.mu[nc]; p_sd = a.sd[nc]; }
        else p_b = a.beta[c];
    }
    if (threadIdx.x == 0) {
        unsigned spins = 0;
        while (ld_acquire_u32(a.ready + c) == 0u) {
            __nanosleep(64);
            if (++spins > kSpin) __trap();
        }
    }
    __syncthreads();
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
                for (int e = 0; e < V; ++e) vo[e] = BWD ? fmaf(ca, vd[e], fmaf(cb, vx[e], cc)) : fmaf(cb, vx[e], 0.f);
                stg_stream(po + i, pack<T>(vo));
            }
        }
        i0 += kStep;
        if (i0 - r >= nv) break;                          // CTA-uniform bound (r < TPI <= kStep)
        issue(i0);
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

// Threads per instance: a power of two that covers the plane in about kBatches batches of kU loads.  Fewer
// threads per instance amortise the per-item latencies (ticket, flag, fence) over more bytes; more threads
// keep the resident window (and with it the L2 footprint between the two reads) narrow.  Measured on
// (256,256,56,56) fp32: TPI 64 (4 batches of 4 loads) is the optimum; 6 batches is the rule that reproduces the
// measured optimum on the other shapes as well (profiles/README.md).
static int pick_tpi(int nv) {
    const int batches = env_int("CNSN_FLOW_BATCHES", 6);
    int tpi = 8;
    while (tpi < kT && tpi * kU * batches < nv) tpi <<= 1;
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
    if (const int v = env_int("CNSN_FLOW_TPI", 0)) { if (v >= 8 && v <= kT && (v & (v - 1)) == 0) tpi = v; }
    const int I = kT / tpi;
    a.nI = (N + I - 1) / I;
    // Look-ahead: enough channels to cover the R items in flight plus the fold latency, bounded by L2.
    const size_t chan_bytes = (size_t)N * a.M * esz * (BWD ? 2 : 1);
    int D = (int)(((size_t)env_int("CNSN_FLOW_LOOKAHEAD_MB", 40) << 20) / (chan_bytes ? chan_bytes : 1));
    if (D < 2) D = 2;
    if (const int v = env_int("CNSN_FLOW_D", 0)) D = v;
    if (D > C) D = C;
    if (D < 1) D = 1;
    a.D = D;
    a.order = env_int("CNSN_FLOW_ORDER", 0);
    a.keep = env_int("CNSN_FLOW_KEEP", 1);
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
    case TPI_: k_sn_flow<T, BWD, TPI_><<<grid, block, 0, stream>>>(a); break;
    CNSN_DISPATCH_DTYPE(dtype, T, switch (tpi) {
        CNSN_FLOW_CASE(8) CNSN_FLOW_CASE(16) CNSN_FLOW_CASE(32) CNSN_FLOW_CASE(64) CNSN_FLOW_CASE(128) CNSN_FLOW_CASE(256)
        default: return -100;
    });
#undef CNSN_FLOW_CASE
    if (getenv("CNSN_FLOW_DEBUG"))
        fprintf(stderr, "[cnsn flow] %s tpi=%d I=%d nI=%d D=%d items=%llu order=%d\n", BWD ? "bwd" : "fwd", tpi, I, a.nI, D,
                items, a.order);
    return launch_status();
}

size_t scratch_floats(int N, int C) { return 2 * (size_t)N * C + 4 * (size_t)C + 8; }

// Both return 0 when launched, >0 cuda error, -100 when the path does not apply.
int selfnorm_flow_fwd(const void* x, void* y, int dtype, int N, int C, int H, int W,
                      const cnsn_gate_params* g, int training, float momentum, float bn_eps, float eps,
                      float* mu, float* sd, float* gate, float* shat, float* r, float* scratch,
                      cudaStream_t stream) {
    if (!aligned16(x) || !aligned16(y)) return -100;
    FArgs a{};
    a.x = x; a.dy = nullptr; a.out = y; a.N = N; a.C = C; a.M = H * W;
    a.training = training; a.momentum = momentum; a.bn_eps = bn_eps; a.eps = eps;
    a.w = g->w; a.gamma = g->gamma; a.beta = g->beta; a.run_mean = g->run_mean; a.run_var = g->run_var; a.nbt = g->nbt;
    a.mu = mu; a.sd = sd; a.gate = gate; a.shat = shat; a.r = r;
    return launch<false>(a, dtype, scratch, stream);
}

int selfnorm_flow_bwd(const void* x, const void* dy, void* dx, int dtype, int N, int C, int H, int W,
                      const cnsn_gate_params* g, int training,
                      float* mu, float* sd, float* gate, float* shat, float* r,
                      const cnsn_gate_grads* dg, float* scratch, cudaStream_t stream) {
    if (!aligned16(x) || !aligned16(dy) || !aligned16(dx)) return -100;
    FArgs a{};
    a.x = x; a.dy = dy; a.out = dx; a.N = N; a.C = C; a.M = H * W;
    a.training = training;
    a.w = g->w; a.gamma = g->gamma;
    a.mu = mu; a.sd = sd; a.gate = gate; a.shat = shat; a.r = r;
    a.dw = dg->dw; a.dgamma = dg->dgamma; a.dbeta = dg->dbeta;
    return launch<true>(a, dtype, scratch, stream);
}

}  // namespace flow
}  // namespace cnsn
