// selfnorm_cluster.cu -- SelfNorm forward / backward with one thread-block CLUSTER per channel
// (sm_100a: 16-CTA clusters, distributed shared memory, cluster-scope mbarriers, TMA bulk loads).
//
// The gate of channel c couples all N instances of c (BatchNorm1d over the batch, models/cnsn.py:121,
// :138).  selfnorm_fused.cu spreads a channel over all 148 CTAs, so every channel is gated by the
// slowest CTA of the whole GPU.  Here a channel belongs to ONE cluster of 16 CTAs:
//
//   * cluster k walks channels c = k, k+K, ...; CTA rank q of the cluster owns samples n = q, q+16, ...
//   * reduce stream (producer warp + 8 reduce warps): TMA 1-D bulk loads of whole planes into a shared-
//     memory ring; per-instance statistics out of shared memory; the 8-byte result is stored into the
//     pair buffer of EVERY CTA of the cluster (st.shared::cluster, one lane per destination) followed by a
//     remote mbarrier.arrive.release.cluster on that CTA's `chan_done` barrier (count = N arrivals).
//   * channel warp: waits on its own CTA's `chan_done` (acquire.cluster) -- all N pairs of the channel are
//     then in local shared memory -- reduces over N, stages the channel constants.
//   * apply stream (12 warps): rebuilds the gate / backward coefficients, re-reads the plane with 128-bit
//     loads (an L2 hit: the plane was loaded by this CTA's own reduce stream one or two channels ago),
//     streams the result out; then tells every CTA of the cluster that its pair buffer slot is free again
//     (remote arrive on `buf_free`, count = N).
//
// No device-wide dependency exists: clusters never talk to each other, so no cooperative launch and no
// global polling; the L2 working set is the 1-2 channels per cluster between the two streams.
// HBM traffic: forward 2*S, backward 3*S.
#include <stdio.h>
#include <stdlib.h>

#include "fused_common.cuh"

namespace cnsn {
namespace cluster {

using namespace fused;

constexpr int kCS = 16;                 // CTAs per cluster (non-portable size; 8 GPCs -> up to 8 clusters resident)
constexpr int kNB = 3;                  // pair-buffer slots = channels the reduce stream may run ahead of apply
constexpr int kCThreads = (kStatsWarps + kApplyWarps + 2) * 32;   // + producer + channel warp
constexpr int kWProducer = kStatsWarps + kApplyWarps;
constexpr int kWChannel = kWProducer + 1;
constexpr int kMaxN = 1024;
constexpr int kCStages = 16;            // max ring stages of this kernel (one plane per stage)

struct CArgs {
    const void* x; const void* dy; void* out;
    int N, C, M, K;                     // K = clusters in the grid
    int S;                              // ring stages (one plane, or one x/dy plane pair, per stage)
    int training;
    float momentum, bn_eps, eps;
    const float* w; const float* gamma; const float* beta;
    float* run_mean; float* run_var; long long* nbt;
    float* mu; float* sd; float* gate; float* shat; float* r;
    float* dw; float* dgamma; float* dbeta;
    unsigned stage_bytes, off_pairs, off_data;
};

struct CMeta { float a, b, c, d, e, f; };   // fwd: m, rstd, gamma, beta, w0, w1 ; bwd: k1, k2, gamma, rstd, w0, w1

__device__ __forceinline__ unsigned cluster_rank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_id() { unsigned r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address -> shared::cluster address of the same variable in CTA `rank`
__device__ __forceinline__ uint32_t remote_addr(const void* local, unsigned rank) {
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(local)), "r"(rank));
    return ra;
}
__device__ __forceinline__ void st_remote_f2(uint32_t ra, float a, float b) {
    asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" :: "r"(ra), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t ra) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" :: "r"(ra) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, unsigned parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, unsigned parity) {
    unsigned spins = 0;
    while (!mbar_try_wait_cluster(bar, parity)) {
        __nanosleep(100);
        if (++spins > kWaitLimit) __trap();
    }
}

// sum dy*x and sum dy of one smem-resident plane pair, whole warp
template <typename T>
__device__ __forceinline__ float2 plane_dot(const T* ix, const T* id, int nv, int lane) {
    constexpr int V = VecOf<T>::n;
    const uint4* px = reinterpret_cast<const uint4*>(ix);
    const uint4* pd = reinterpret_cast<const uint4*>(id);
    float a0[2] = {0.f, 0.f}, a1[2] = {0.f, 0.f};
    const int step = 32 * 4, nfull = (nv / step) * step;
    int i = lane;
    for (; i < nfull; i += step) {
        uint4 rx[4], rd[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { rx[u] = px[i + u * 32]; rd[u] = pd[i + u * 32]; }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float vx[V], vd[V];
            unpack<T>(rx[u], vx);
            unpack<T>(rd[u], vd);
#pragma unroll
            for (int e = 0; e < V; ++e) { a0[e & 1] = fmaf(vd[e], vx[e], a0[e & 1]); a1[e & 1] += vd[e]; }
        }
    }
    for (; i < nv; i += 32) {
        float vx[V], vd[V];
        unpack<T>(px[i], vx);
        unpack<T>(pd[i], vd);
#pragma unroll
        for (int e = 0; e < V; ++e) { a0[e & 1] = fmaf(vd[e], vx[e], a0[e & 1]); a1[e & 1] += vd[e]; }
    }
    return make_float2(warp_sum(a0[0] + a0[1]), warp_sum(a1[0] + a1[1]));
}

template <typename T, bool BWD>
__global__ void __launch_bounds__(kCThreads, 1) k_sn_cluster(const CArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int N = a.N, C = a.C, M = a.M, K = a.K, S = a.S;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);           // [kCStages] plane landed
    uint64_t* empty = full + kCStages;                            // [kCStages] plane reduced
    uint64_t* chan_done = empty + kCStages;                     // [kNB] all N pairs of the channel are here (remote arrivals)
    uint64_t* buf_free = chan_done + kNB;                         // [kNB] all N instances of the channel applied (remote arrivals)
    uint64_t* chan_ready = buf_free + kNB;                        // [kNB] channel constants staged (local)
    CMeta* cmeta = reinterpret_cast<CMeta*>(chan_ready + kNB);    // [kNB]
    float2* pairs = reinterpret_cast<float2*>(smem + a.off_pairs);   // [kNB][N]
    unsigned char* data = smem + a.off_data;                      // [S][stage_bytes]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = (int)cluster_rank(), k = (int)cluster_id();
    if (threadIdx.x == 0) {
        for (int i = 0; i < S; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        // cluster-scope barriers count one arrival per instance (aggregating per CTA behind cluster-scope fences
        // was measured slower: the fences flush L1)
        for (int i = 0; i < kNB; ++i) { mbar_init(&chan_done[i], N); mbar_init(&buf_free[i], N); mbar_init(&chan_ready[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cluster_sync_all();                                          // barriers of every CTA exist before any remote arrive

    const T* x = static_cast<const T*>(a.x);
    const T* dy = static_cast<const T*>(a.dy);
    T* out = static_cast<T*>(a.out);
    constexpr int V = VecOf<T>::n;
    const int nv = M / V;
    const unsigned plane_bytes = (unsigned)(M * sizeof(T));
    const int mine = (N - q + kCS - 1) / kCS;                    // my samples per channel: n = q + 16*j
    const int nchan = (C - k + K - 1) / K;                       // channels of this cluster: c = k + K*t

    if (warp == kWProducer) {
        if (lane == 0) {
            if (!BWD && blockIdx.x == 0 && a.training && a.nbt) *a.nbt += 1;
            long long item = 0;
            for (int t = 0; t < nchan; ++t) {
                const int c = k + K * t;
                for (int j = 0; j < mine; ++j, ++item) {
                    const int st = (int)(item % S), ph = (int)((item / S) & 1);
                    mbar_wait(&empty[st], ph ^ 1);
                    const size_t off = ((size_t)(q + kCS * j) * C + c) * M;
                    unsigned char* dst = data + (size_t)st * a.stage_bytes;
                    mbar_arrive_expect_tx(&full[st], plane_bytes * (BWD ? 2 : 1));
                    tma_load_1d_plain(dst, x + off, plane_bytes, &full[st]);
                    if (BWD) tma_load_1d_plain(dst + plane_bytes, dy + off, plane_bytes, &full[st]);
                }
            }
        }
    } else if (warp < kStatsWarps) {
        // ------------------------------------------------------------ reduce stream: item -> warp item % 8
        long long item = 0;
        for (int t = 0; t < nchan; ++t) {
            const int b = t % kNB, bph = (t / kNB) & 1;
            mbar_wait_cluster(&buf_free[b], bph ^ 1);            // every CTA is done reading slot b (channel t-kNB)
            for (int j = 0; j < mine; ++j, ++item) {
                const int st = (int)(item % S), ph = (int)((item / S) & 1);
                if (st % kStatsWarps != warp) continue;          // by STAGE: one warp sees all phases of a stage's barrier, in order
                mbar_wait(&full[st], ph);
                const T* plane = reinterpret_cast<const T*>(data + (size_t)st * a.stage_bytes);
                float2 word;
                if (BWD) {
                    word = plane_dot<T>(plane, reinterpret_cast<const T*>(data + (size_t)st * a.stage_bytes + plane_bytes), nv, lane);
                } else {
                    const float2 mq = smem_mean_m2<T>(plane, M, lane, 32, true, true);
                    word = make_float2(mq.x, sqrtf(mq.y / (float)(M - 1) + a.eps));
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[st]);          // the plane may be overwritten
                if (lane < kCS) {                                // lane l delivers to CTA l of the cluster
                    const int n = q + kCS * j;
                    st_remote_f2(remote_addr(&pairs[b * N + n], lane), word.x, word.y);
                    mbar_arrive_remote(remote_addr(&chan_done[b], lane));   // release: the store above is visible first
                }
            }
        }
    } else if (warp < kWProducer) {
        // ------------------------------------------------------------ apply stream: item -> warp item % 12
        const int aw = warp - kStatsWarps;
        const uint64_t pol = l2_policy_evict_first();
        const float invM = 1.f / M, invM1 = 1.f / (M - 1.f);
        long long item = 0;
        for (int t = 0; t < nchan; ++t) {
            const int c = k + K * t, b = t % kNB, bph = (t / kNB) & 1;
            mbar_wait(&chan_ready[b], bph);
            const CMeta cm = cmeta[b];
            for (int j = 0; j < mine; ++j, ++item) {
                if ((int)(item % kApplyWarps) != aw) continue;
                const int n = q + kCS * j;
                const size_t nc = (size_t)n * C + c;
                const float2 own = pairs[b * N + n];
                float ca, cb, cc;
                if (BWD) {
                    const float gt = a.gate[nc], sh = a.shat[nc], mean = a.mu[nc], sdev = a.sd[nc];
                    const float ds = cm.d * (own.x * gt * (1.f - gt) * cm.c - cm.a - sh * cm.b);
                    ca = gt;
                    cb = ds * cm.f * invM1 / sdev;
                    cc = ds * cm.e * invM - cb * mean;
                } else {
                    const float sh = (fmaf(cm.e, own.x, cm.f * own.y) - cm.a) * cm.b;
                    const float gt = 1.f / (1.f + expf(-fmaf(cm.c, sh, cm.d)));
                    if (lane == 0) { a.mu[nc] = own.x; a.sd[nc] = own.y; a.gate[nc] = gt; a.shat[nc] = sh; }
                    ca = 0.f; cb = gt; cc = 0.f;
                }
                const uint4* px = reinterpret_cast<const uint4*>(x + nc * M);
                const uint4* pd = reinterpret_cast<const uint4*>(BWD ? dy + nc * M : x);
                uint4* po = reinterpret_cast<uint4*>(out + nc * M);
                // Full-duty register ring: U 128-bit loads per tensor stay in flight per lane; as soon as slot s
                // has been consumed (scaled + stored) its register is refilled with slot s+U.  Out-of-range
                // slots are clamped for the load and predicated for the store (no branches in the loop).
                constexpr int U = BWD ? 4 : 8;
                const int nslots = (nv + 31) >> 5;
                uint4 rx[U], rd[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int v = min(lane + 32 * u, nv - 1);
                    rx[u] = ldg_hint(px + v, pol);                       // L2 hit, last use
                    if (BWD) rd[u] = ldg_hint(pd + v, pol);
                }
                for (int s0 = 0; s0 < nslots; s0 += U) {
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int v = lane + 32 * (s0 + u);
                        float vx[V], vd[V], vo[V];
                        unpack<T>(rx[u], vx);
                        if (BWD) unpack<T>(rd[u], vd);
#pragma unroll
                        for (int e = 0; e < V; ++e) vo[e] = BWD ? fmaf(ca, vd[e], fmaf(cb, vx[e], cc)) : vx[e] * cb;
                        if (v < nv) stg_stream(po + v, pack<T>(vo));
                        if (s0 + u + U < nslots) {                       // warp-uniform: refill this register with slot s+U
                            const int vn = min(v + 32 * U, nv - 1);
                            rx[u] = ldg_hint(px + vn, pol);
                            if (BWD) rd[u] = ldg_hint(pd + vn, pol);
                        }
                    }
                }
                __syncwarp();
                if (lane < kCS) mbar_arrive_remote(remote_addr(&buf_free[b], lane));   // one arrival per instance, to every CTA
            }
        }
    } else if (warp == kWChannel) {
        // ------------------------------------------------------------ channel warp
        const float invN = 1.f / N;
        for (int t = 0; t < nchan; ++t) {
            const int c = k + K * t, b = t % kNB, bph = (t / kNB) & 1;
            const float w0 = a.w[2 * c], w1 = a.w[2 * c + 1], ga = a.gamma[c];
            const float p3 = BWD ? a.r[c] : a.beta[c];
            float prm = 0.f, prv = 1.f;
            if (!BWD && (q == 0 || !a.training)) { prm = a.run_mean[c]; prv = a.run_var[c]; }
            mbar_wait_cluster(&chan_done[b], bph);               // all N pairs of channel c are in my shared memory
            const float2* pb = pairs + b * N;
            CMeta cm;
            cm.c = ga; cm.d = p3; cm.e = w0; cm.f = w1;
            if (!BWD) {
                float m = prm, rstd;
                if (a.training) {
                    float sum = 0.f;
                    for (int n = lane; n < N; n += 32) sum += fmaf(w0, pb[n].x, w1 * pb[n].y);
                    m = warp_sum(sum) * invN;
                    float qv = 0.f;
                    for (int n = lane; n < N; n += 32) { const float d = fmaf(w0, pb[n].x, w1 * pb[n].y) - m; qv = fmaf(d, d, qv); }
                    qv = warp_sum(qv) * invN;
                    rstd = 1.f / sqrtf(qv + a.bn_eps);
                    if (q == 0 && lane == 0) {                   // rank 0 of the cluster keeps the channel's books
                        a.run_mean[c] = (1.f - a.momentum) * prm + a.momentum * m;
                        a.run_var[c] = (1.f - a.momentum) * prv + a.momentum * (qv * N / (N - 1.f));
                    }
                } else {
                    rstd = 1.f / sqrtf(prv + a.bn_eps);
                }
                if (q == 0 && lane == 0) a.r[c] = rstd;
                cm.a = m; cm.b = rstd;
            } else {
                float sg = 0.f, sb = 0.f;
                for (int n = lane; n < N; n += 32) {
                    const size_t nc = (size_t)n * C + c;
                    const float gt = a.gate[nc];
                    const float dz = pb[n].x * gt * (1.f - gt);
                    sg = fmaf(dz, a.shat[nc], sg); sb += dz;
                }
                sg = warp_sum(sg); sb = warp_sum(sb);
                const float k1 = a.training ? ga * sb * invN : 0.f, k2 = a.training ? ga * sg * invN : 0.f;
                if (q == 0) {
                    float t0 = 0.f, t1 = 0.f;
                    for (int n = lane; n < N; n += 32) {
                        const size_t nc = (size_t)n * C + c;
                        const float gt = a.gate[nc];
                        const float ds = p3 * (pb[n].x * gt * (1.f - gt) * ga - k1 - a.shat[nc] * k2);
                        t0 = fmaf(ds, a.mu[nc], t0); t1 = fmaf(ds, a.sd[nc], t1);
                    }
                    t0 = warp_sum(t0); t1 = warp_sum(t1);
                    if (lane == 0) { a.dgamma[c] = sg; a.dbeta[c] = sb; a.dw[2 * c] = t0; a.dw[2 * c + 1] = t1; }
                }
                cm.a = k1; cm.b = k2;
            }
            // cmeta[b] is free: chan_done(t) implies buf_free(t - kNB) completed, i.e. every apply warp of every CTA
            // (ours included) is past channel t - kNB
            if (lane == 0) { cmeta[b] = cm; }
            __syncwarp();
            if (lane == 0) mbar_arrive(&chan_ready[b]);
        }
    }
    __syncwarp();
    cluster_sync_all();                                          // nobody exits while peers may still write into its smem
}

// ------------------------------------------------------------------------------------------------
static bool plan(const void* fn, int N, int C, int M, int dtype, int tensors, CArgs& a, unsigned& smem_total, int& K) {
    int dev = 0, smem_optin = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return false;
    cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    const size_t esz = esize(dtype), plane = (size_t)M * esz;
    if (plane % 16 || plane < 8192 || N < 2 * kCS || N > kMaxN) return false;
    if ((size_t)N * C * plane < ((size_t)8 << 20)) return false;
    const unsigned hdr = (2 * kCStages + 3 * kNB) * 8 + (2 * kNB + 2) * 4 + kNB * (unsigned)sizeof(CMeta) + 64;
    a.off_pairs = (hdr + 15) & ~15u;
    a.off_data = (a.off_pairs + (unsigned)(kNB * N * sizeof(float2)) + 127) & ~127u;
    a.stage_bytes = (unsigned)((plane * tensors + 127) & ~(size_t)127);
    int S = (int)(((size_t)smem_optin - a.off_data) / a.stage_bytes);
    if (S > kCStages) S = kCStages;
    if (S >= kStatsWarps) S = (S / kStatsWarps) * kStatsWarps;        // stages map evenly onto the reduce warps
    if (const char* e = getenv("CNSN_CLUSTER_STAGES")) { const int v = atoi(e); if (v >= 2 && v < S) S = v; }
    if (S < 3) return false;
    a.S = S;
    smem_total = a.off_data + (unsigned)S * a.stage_bytes;
    if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_total) != cudaSuccess) return false;
    if (cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) return false;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(kCS); cfg.blockDim = dim3(kCThreads); cfg.dynamicSmemBytes = smem_total;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = kCS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int maxc = 0;
    if (cudaOccupancyMaxActiveClusters(&maxc, fn, &cfg) != cudaSuccess || maxc < 1) { cudaGetLastError(); return false; }
    K = maxc;
    if (const char* e = getenv("CNSN_CLUSTERS")) { const int v = atoi(e); if (v > 0 && v < K) K = v; }
    if (K > C) K = C;
    return true;
}

template <bool BWD>
static int launch(CArgs& a, int dtype, int tensors, cudaStream_t stream) {
    const void* fn = nullptr;
    switch (dtype) {
        case CNSN_F32: fn = (const void*)k_sn_cluster<float, BWD>; break;
        case CNSN_BF16: fn = (const void*)k_sn_cluster<__nv_bfloat16, BWD>; break;
        case CNSN_F16: fn = (const void*)k_sn_cluster<__half, BWD>; break;
        default: return CNSN_E_BADARG;
    }
    unsigned smem_total = 0;
    int K = 0;
    if (!plan(fn, a.N, a.C, a.M, dtype, tensors, a, smem_total, K)) return -100;
    a.K = K;
    if (getenv("CNSN_CLUSTER_DEBUG")) fprintf(stderr, "[cnsn cluster] K=%d clusters x %d CTAs, S=%d stages of %u B, smem %u B\n", K, kCS, a.S, a.stage_bytes, smem_total);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(K * kCS); cfg.blockDim = dim3(kCThreads); cfg.dynamicSmemBytes = smem_total; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = kCS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    void* args[] = {&a};
    const cudaError_t e = cudaLaunchKernelExC(&cfg, fn, args);
    note_launch();
    return (int)e;
}

int selfnorm_cluster_fwd(const void* x, void* y, int dtype, int N, int C, int H, int W,
                         const cnsn_gate_params* g, int training, float momentum, float bn_eps, float eps,
                         float* mu, float* sd, float* gate, float* shat, float* r, cudaStream_t stream) {
    if (!aligned16(x) || !aligned16(y)) return -100;
    CArgs a{};
    a.x = x; a.dy = nullptr; a.out = y; a.N = N; a.C = C; a.M = H * W; a.training = training;
    a.momentum = momentum; a.bn_eps = bn_eps; a.eps = eps;
    a.w = g->w; a.gamma = g->gamma; a.beta = g->beta; a.run_mean = g->run_mean; a.run_var = g->run_var; a.nbt = g->nbt;
    a.mu = mu; a.sd = sd; a.gate = gate; a.shat = shat; a.r = r;
    return launch<false>(a, dtype, 1, stream);
}

int selfnorm_cluster_bwd(const void* x, const void* dy, void* dx, int dtype, int N, int C, int H, int W,
                         const cnsn_gate_params* g, int training,
                         float* mu, float* sd, float* gate, float* shat, float* r,
                         const cnsn_gate_grads* dg, cudaStream_t stream) {
    if (!aligned16(x) || !aligned16(dy) || !aligned16(dx)) return -100;
    CArgs a{};
    a.x = x; a.dy = dy; a.out = dx; a.N = N; a.C = C; a.M = H * W; a.training = training;
    a.w = g->w; a.gamma = g->gamma; a.beta = nullptr;
    a.mu = mu; a.sd = sd; a.gate = gate; a.shat = shat; a.r = r;
    a.dw = dg->dw; a.dgamma = dg->dgamma; a.dbeta = dg->dbeta;
    return launch<true>(a, dtype, 2, stream);
}

}  // namespace cluster
}  // namespace cnsn
