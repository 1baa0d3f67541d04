// selfnorm_fused.cu -- SelfNorm forward as ONE persistent, warp-specialised kernel (sm_100a).
//
// Why: the gate of channel c needs the statistics of ALL N instances of c (BatchNorm1d over the
// batch, models/cnsn.py:121,:138), so a stats-then-apply design reads x twice (3*S of HBM traffic
// for 2*S algorithmic).  Here every byte of x is read from HBM exactly once: each CTA keeps the
// instances it reduced resident in shared memory until the channel's batch statistics are known
// (a device-wide, per-channel-group dependency -- not a kernel boundary), then scales them out of
// shared memory.  HBM traffic = 2*S.
//
// One CTA per SM (cooperative launch => all co-resident), 17 warps:
//   producer (1 warp)   TMA 1-D bulk loads (cp.async.bulk + mbarrier complete_tx) of whole units
//                       (one sample's run of kk channels) into an S-stage shared-memory ring
//   stats    (8 warps)  unit j of stage st -> warp (st*upc+j)%8: one-pass shifted-data mean / variance per
//                       instance out of shared memory (sub-warp teams for small planes); (mu, sd) is kept in smem and published
//                       as one 8-byte "data is the flag" store
//   apply    (8 warps)  unit j of stage st -> warp (st*upc+j)%8.  The owner of unit 0 first does the
//                       stage's channel duty: polls the N*kk published pairs of the group (every CTA does
//                       this redundantly: 8 B per instance out of L2), reduces s = w0*mu + w1*sd over N to
//                       the BatchNorm batch mean / rstd and stages (m, r, gamma, beta, w) in smem (the
//                       channel's owner CTA also updates the running statistics).  Then every apply warp:
//                       g = sigmoid(gamma*shat+beta); y = x*g from shared memory, 128-bit streaming stores
//   The producer also issues cp.async.bulk.prefetch.L2 a few groups ahead, so the HBM latency (and its
//   tail across 148 SMs) is paid before a ring slot is tied up.
//
// Cross-CTA latency chain per group: one 8-byte store, one polled 8-byte load.  Deadlock freedom:
// group g's pairs only need every CTA's stats of group g, which only need that CTA's ring slot,
// which only needs apply of group g-S, which only needs pairs of group g-S.  Every spin is bounded
// and traps.
#include <stdio.h>
#include <stdlib.h>

#include "fused_common.cuh"

namespace cnsn {
namespace fused {

struct FwdArgs {
    Schedule sch;
    const void* x;
    void* y;
    const float* w;
    const float* gamma;
    const float* beta;
    float* run_mean;
    float* run_var;
    long long* nbt;
    float momentum, bn_eps, eps;
    int training;
    float* mu; float* sd; float* gate; float* shat; float* r;   // save block
    float2* pairs;          // [C][N] (mu, sd) exchange area, pre-filled with the sentinel
    unsigned stage_bytes;   // ring slot size (multiple of 128)
    unsigned off_inst, off_chan, off_pair, off_data;   // smem offsets
    unsigned long long* trace;   // debug only (CNSN_FUSED_TRACE): [B][G][8] globaltimer stamps, else NULL
};

__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;                       // global ns timer: comparable across CTAs
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Trace slots per group: 0 load issued, 1 data landed (stats start), 2 stats done, 3 pairs complete,
// 4 chan_ready, 5 apply start, 6 apply done.
#define CNSN_TRACE(slot, g)                                                                   \
    do {                                                                                      \
        if (a.trace) a.trace[((size_t)b * G + (g)) * 8 + (slot)] = gtime();                    \
    } while (0)

struct ChanMeta { float m, r, gamma, beta, w0, w1; };

template <typename T>
__global__ void __launch_bounds__(kThreads, 1) k_sn_fused_fwd(const FwdArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const Schedule& s = a.sch;
    const int S = s.S, kk = s.kk, B = s.B, C = s.C, M = s.M, G = s.G, N = s.N;
    const int Imax = s.upc * kk;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* stats_done = full + kMaxStages;
    uint64_t* chan_ready = stats_done + kMaxStages;
    uint64_t* empty = chan_ready + kMaxStages;
    float2* inst_meta = reinterpret_cast<float2*>(smem + a.off_inst);           // [S][Imax]
    ChanMeta* chan_meta = reinterpret_cast<ChanMeta*>(smem + a.off_chan);       // [S][kk]
    float2* pair_buf = reinterpret_cast<float2*>(smem + a.off_pair);            // [S][kk*N] gathered pairs
    unsigned char* data = smem + a.off_data;                                    // [S][stage_bytes]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, b = blockIdx.x;
    if (threadIdx.x == 0) {
        for (int i = 0; i < S; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&stats_done[i], s.upc);           // one arrival per unit slot of the stage
            mbar_init(&chan_ready[i], 1);
            mbar_init(&empty[i], s.upc);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const T* x = static_cast<const T*>(a.x);
    T* y = static_cast<T*>(a.y);
    const bool vec = ((size_t)M * sizeof(T)) % 16 == 0;
    const unsigned unit_bytes = s.unit_elems * (unsigned)sizeof(T);
    const int lpi = s.lpi, ipw = 32 / lpi, sub = lane / lpi, r = lane % lpi;   // sub-warp teams
    const int total = N * kk, nslots = (total + 31) >> 5;                      // pairs per group (<= kMaxPairs)

    if (warp == kWarpProducer) {
        if (lane == 0) {
            const uint64_t pol = l2_policy_evict_first();     // x is read exactly once
            if (b == 0 && a.training && a.nbt) *a.nbt += 1;
            GroupIter it, pf;
            pf.init(s, b);
            auto prefetch_group = [&](const GroupIter& q) {
                for (int j = 0; j < q.cnt; ++j) {
                    const size_t n = (size_t)q.first + (size_t)j * B;
                    tma_prefetch_l2(x + (n * C + (size_t)q.g * kk) * M, unit_bytes);
                }
            };
            for (int i = 0; i < S + kPrefetchAhead && pf.g < G; ++i, pf.next(s))
                if (i >= S) prefetch_group(pf);             // the first S groups are loaded straight away
            for (it.init(s, b); it.g < G; it.next(s)) {
                const int g = it.g, st = it.st, ph = it.ph, first = it.first, cnt = it.cnt;
                mbar_wait(&empty[st], ph ^ 1);
                CNSN_TRACE(0, g);
                if (cnt == 0) { mbar_arrive(&full[st]); }
                else {
                    mbar_arrive_expect_tx(&full[st], cnt * unit_bytes);
                    for (int j = 0; j < cnt; ++j) {
                        const size_t n = (size_t)first + (size_t)j * B;
                        tma_load_1d(data + (size_t)st * a.stage_bytes + (size_t)j * unit_bytes,
                                    x + (n * C + (size_t)g * kk) * M, unit_bytes, &full[st], pol);
                    }
                }
                if (pf.g < G) { prefetch_group(pf); pf.next(s); }
            }
        }
    } else if (warp < kStatsWarps) {
        // Stats warps: unit j of ring stage st is reduced by warp (st*upc + j) % 8 -- the units of a
        // group and the groups in the ring all proceed in parallel, nothing is synchronised across warps.
        GroupIter it;
        for (it.init(s, b); it.g < G; it.next(s)) {
            const int g = it.g, st = it.st, ph = it.ph, first = it.first, cnt = it.cnt;
            bool any = false;
            for (int j = 0; j < s.upc; ++j) any |= s.my_unit(st, j, warp, kStatsWarps);
            if (!any) continue;
            mbar_wait(&full[st], ph);
            if (lane == 0 && s.my_unit(st, 0, warp, kStatsWarps)) CNSN_TRACE(1, g);
            const T* base = reinterpret_cast<const T*>(data + (size_t)st * a.stage_bytes);
            for (int j = 0; j < cnt; ++j) {
                if (!s.my_unit(st, j, warp, kStatsWarps)) continue;
                const size_t n = (size_t)first + (size_t)j * B;
                for (int c0 = 0; c0 < kk; c0 += ipw) {
                    const int cl = c0 + sub;
                    const bool live = cl < kk;
                    const int inst = j * kk + cl;
                    const float2 mq = smem_mean_m2<T>(base + (size_t)inst * M, M, r, lpi, vec, live);
                    if (lane == 0 && j == 0 && c0 == 0) CNSN_TRACE(7, g);
                    if (live && r == 0) {
                        const float sdv = sqrtf(mq.y / (float)(M - 1) + a.eps);
                        inst_meta[st * Imax + inst] = make_float2(mq.x, sdv);
                        if (a.training) ll_publish(a.pairs + ((size_t)g * kk + cl) * N + n, mq.x, sdv);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) {
                if (s.my_unit(st, 0, warp, kStatsWarps)) CNSN_TRACE(2, g);
                for (int j = 0; j < s.upc; ++j)
                    if (s.my_unit(st, j, warp, kStatsWarps)) mbar_arrive(&stats_done[st]);
            }
        }
    } else if (warp < kStatsWarps + kApplyWarps) {
        const int aw = warp - kStatsWarps;
        GroupIter it;
        for (it.init(s, b); it.g < G; it.next(s)) {
            const int g = it.g, st = it.st, ph = it.ph, first = it.first, cnt = it.cnt;
            bool any = false;
            for (int j = 0; j < s.upc; ++j) any |= s.my_unit(st, j, aw, kApplyWarps);
            if (!any) continue;
            mbar_wait(&stats_done[st], ph);
            mbar_wait(&full[st], ph);
            if (s.my_unit(st, 0, aw, kApplyWarps)) {
                // This warp owns the stage's channel duty: gather the group's (mu, sd) pairs from all
                // CTAs, reduce over N, stage the channel constants for every apply warp of the stage.
                float pw0 = 0.f, pw1 = 0.f, pga = 0.f, pbe = 0.f, prm = 0.f, prv = 1.f;
                if (lane < kk) {
                    const int c = g * kk + lane;
                    pw0 = a.w[2 * c]; pw1 = a.w[2 * c + 1]; pga = a.gamma[c]; pbe = a.beta[c];
                    if (!a.training || c % B == b) { prm = a.run_mean[c]; prv = a.run_var[c]; }
                }
                float mres = 0.f, rres = 1.f, qres = 0.f;
                if (a.training) {
                    float2* my_buf = pair_buf + (size_t)st * total;
                    const float2* pbase = a.pairs + (size_t)g * kk * N;
                    unsigned pend = 0;
                    for (int j = 0; j < nslots; ++j) if (lane + 32 * j < total) pend |= 1u << j;
                    unsigned spins = 0;
                    while (true) {
                        for (int j0 = 0; j0 < nslots; j0 += 8) {
                            float2 v[8];
#pragma unroll
                            for (int u = 0; u < 8; ++u)
                                if (pend & (1u << (j0 + u))) v[u] = ll_peek(pbase + lane + 32 * (j0 + u));
#pragma unroll
                            for (int u = 0; u < 8; ++u)
                                if ((pend & (1u << (j0 + u))) && ll_valid(v[u])) {
                                    my_buf[lane + 32 * (j0 + u)] = v[u];
                                    pend &= ~(1u << (j0 + u));
                                }
                        }
                        if (!__any_sync(0xffffffffu, pend != 0)) break;
                        __nanosleep(100);
                        if (++spins > kSpinLimit) __trap();
                    }
                    __syncwarp();
                    if (lane == 0) CNSN_TRACE(3, g);
                    for (int cl = 0; cl < kk; ++cl) {
                        const float w0 = __shfl_sync(0xffffffffu, pw0, cl), w1 = __shfl_sync(0xffffffffu, pw1, cl);
                        const float2* pb = my_buf + cl * N;
                        float sum = 0.f;
                        for (int n = lane; n < N; n += 32) sum += fmaf(w0, pb[n].x, w1 * pb[n].y);
                        const float m = warp_sum(sum) / N;
                        float q = 0.f;
                        for (int n = lane; n < N; n += 32) {
                            const float d = fmaf(w0, pb[n].x, w1 * pb[n].y) - m;
                            q = fmaf(d, d, q);
                        }
                        q = warp_sum(q) / N;
                        if (lane == cl) { mres = m; qres = q; rres = 1.f / sqrtf(q + a.bn_eps); }
                    }
                } else if (lane < kk) {
                    mres = prm; rres = 1.f / sqrtf(prv + a.bn_eps);
                }
                if (lane < kk) {
                    const int c = g * kk + lane;
                    ChanMeta cm;
                    cm.m = mres; cm.r = rres; cm.gamma = pga; cm.beta = pbe; cm.w0 = pw0; cm.w1 = pw1;
                    chan_meta[st * kk + lane] = cm;
                    if (c % B == b) {                        // this CTA owns the channel's bookkeeping
                        a.r[c] = rres;
                        if (a.training) {
                            a.run_mean[c] = (1.f - a.momentum) * prm + a.momentum * mres;
                            a.run_var[c] = (1.f - a.momentum) * prv + a.momentum * (qres * N / (N - 1.f));
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) { CNSN_TRACE(4, g); mbar_arrive(&chan_ready[st]); }
            }
            mbar_wait(&chan_ready[st], ph);
            if (lane == 0 && s.my_unit(st, 0, aw, kApplyWarps)) CNSN_TRACE(5, g);
            const T* base = reinterpret_cast<const T*>(data + (size_t)st * a.stage_bytes);
            for (int j = 0; j < cnt; ++j) {
                if (!s.my_unit(st, j, aw, kApplyWarps)) continue;
                const size_t n = (size_t)first + (size_t)j * B;
                for (int cl = sub; cl < kk; cl += ipw) {
                    const int inst = j * kk + cl;
                    const size_t nc = n * C + (size_t)g * kk + cl;
                    const float2 ms = inst_meta[st * Imax + inst];
                    const ChanMeta cm = chan_meta[st * kk + cl];
                    const float sh = (fmaf(cm.w0, ms.x, cm.w1 * ms.y) - cm.m) * cm.r;
                    const float gt = 1.f / (1.f + expf(-fmaf(cm.gamma, sh, cm.beta)));
                    if (r == 0) { a.mu[nc] = ms.x; a.sd[nc] = ms.y; a.gate[nc] = gt; a.shat[nc] = sh; }
                    const T* src = base + (size_t)inst * M;
                    T* dst = y + nc * M;
                    if (vec) {
                        constexpr int V = VecOf<T>::n;
                        const uint4* ps = reinterpret_cast<const uint4*>(src);
                        uint4* pd = reinterpret_cast<uint4*>(dst);
                        const int nv = M / V;
                        for (int i = r; i < nv; i += lpi * kBatch) {
                            uint4 raw[kBatch];
#pragma unroll
                            for (int u = 0; u < kBatch; ++u)
                                if (i + u * lpi < nv) raw[u] = ps[i + u * lpi];
#pragma unroll
                            for (int u = 0; u < kBatch; ++u)
                                if (i + u * lpi < nv) {
                                    float v[V];
                                    unpack<T>(raw[u], v);
#pragma unroll
                                    for (int q = 0; q < V; ++q) v[q] *= gt;
                                    stg_stream(pd + i + u * lpi, pack<T>(v));
                                }
                        }
                    } else {
                        for (int i = r; i < M; i += lpi) dst[i] = from_f<T>(to_f(src[i]) * gt);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) {
                if (s.my_unit(st, 0, aw, kApplyWarps)) CNSN_TRACE(6, g);
                for (int j = 0; j < s.upc; ++j)
                    if (s.my_unit(st, j, aw, kApplyWarps)) mbar_arrive(&empty[st]);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host side: plan + launch
// ---------------------------------------------------------------------------------------------
struct DeviceInfo { int sms; int smem_optin; int coop; };
static DeviceInfo device_info() {
    DeviceInfo d{0, 0, 0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return d;
    cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&d.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaDeviceGetAttribute(&d.coop, cudaDevAttrCooperativeLaunch, dev);
    return d;
}

static int pick_rot(int B) {
    static const int primes[] = {61, 59, 53, 47, 43, 41, 37, 31, 29, 23, 19, 17, 13, 11, 7, 5, 3};
    for (int p : primes) if (p < B && B % p != 0) return p;
    return 1;
}

// Ring geometry for a tensor; per_unit_copies = 1 (forward: x) or 2 (backward: x and dy).
bool make_plan(int N, int C, int M, int dtype, int per_unit_copies, int min_stages, const DeviceInfo& d,
               Schedule& s, unsigned& stage_bytes, unsigned& off_inst, unsigned& off_chan, unsigned& off_pair,
               unsigned& off_data, unsigned& smem_total, size_t chan_meta_bytes, size_t inst_meta_bytes) {
    if (d.sms <= 0 || !d.coop || N < 2 || N > kMaxPairs) return false;
    if ((N + d.sms - 1) / d.sms > kStatsWarps) return false;               // unit slots of a stage map to distinct warps
    const size_t esz = esize(dtype), inst_bytes = (size_t)M * esz;
    if ((size_t)N * C * inst_bytes < ((size_t)8 << 20)) return false;       // small tensors: the 3-kernel path
    const int B = d.sms;
    const int upc = (N + B - 1) / B;
    const size_t budget = (size_t)d.smem_optin - 1024;
    int best = 0;
    for (int kk = 1; kk <= kMaxKK && kk <= C; ++kk) {
        if (C % kk) continue;
        if ((kk * inst_bytes) % 16) continue;
        if (upc * kk > kMaxInst || (size_t)N * kk > (size_t)kMaxPairs) break;
        const size_t sb = ((size_t)upc * kk * inst_bytes * per_unit_copies + 127) & ~(size_t)127;
        const size_t per_stage = sb + (size_t)upc * kk * inst_meta_bytes + (size_t)kk * chan_meta_bytes
                                 + (size_t)N * kk * sizeof(float2);             // + the stage's pair staging buffer
        if (per_stage * min_stages + 1024 > budget) break;
        best = kk;
        if (kk * inst_bytes >= 8192) break;                                  // big enough units
    }
    if (!best) return false;
    const int kk = best;
    s.N = N; s.C = C; s.M = M; s.kk = kk; s.G = C / kk; s.B = B; s.upc = upc; s.rot = pick_rot(B);
    s.unit_elems = (unsigned)(kk * M);
    const int Imax = upc * kk;
    s.lpi = M >= 512 ? 32 : M >= 256 ? 16 : M >= 96 ? 8 : 4;
    (void)Imax;
    stage_bytes = (unsigned)(((size_t)upc * kk * inst_bytes * per_unit_copies + 127) & ~(size_t)127);
    const size_t per_stage = stage_bytes + (size_t)Imax * inst_meta_bytes + (size_t)kk * chan_meta_bytes
                             + (size_t)N * kk * sizeof(float2);
    int S = (int)((budget - 1024) / per_stage);
    if (S > kMaxStages) S = kMaxStages;
    if (S > s.G) S = s.G < 1 ? 1 : s.G;
    if (S < min_stages && S < s.G) return false;
    s.S = S;
    const unsigned hdr = 4 * kMaxStages * 8 + 2 * kStatsWarps * 16 + 64;     // barriers + scratch
    off_inst = (hdr + 15) & ~15u;
    off_chan = (off_inst + (unsigned)(S * Imax * inst_meta_bytes) + 15) & ~15u;
    off_pair = (off_chan + (unsigned)(S * kk * chan_meta_bytes) + 15) & ~15u;
    off_data = (off_pair + (unsigned)((size_t)S * N * kk * sizeof(float2)) + 127) & ~127u;
    smem_total = off_data + (unsigned)S * stage_bytes;
    return smem_total <= (unsigned)d.smem_optin;
}

// Returns 0 when launched, >0 cuda error, -100 when the fused path does not apply (caller falls back
// to the three-kernel path).
int selfnorm_fused_fwd(const void* x, void* y, int dtype, int N, int C, int H, int W,
                       const cnsn_gate_params* g, int training, float momentum, float bn_eps, float eps,
                       float* mu, float* sd, float* gate, float* shat, float* r, float* scratch_floats,
                       cudaStream_t stream) {
    DeviceInfo d = device_info();
    if (const char* e = getenv("CNSN_FUSED_CTAS")) {                 // experiment knob: CTAs (<= SM count)
        const int v = atoi(e);
        if (v > 0 && v <= d.sms) d.sms = v;
    }
    FwdArgs a;
    unsigned smem_total = 0;
    if (!aligned16(x) || !aligned16(y)) return -100;
    if (!make_plan(N, C, H * W, dtype, 1, 4, d, a.sch, a.stage_bytes, a.off_inst, a.off_chan, a.off_pair,
                   a.off_data, smem_total, sizeof(ChanMeta), sizeof(float2)))
        return -100;
    a.x = x; a.y = y; a.w = g->w; a.gamma = g->gamma; a.beta = g->beta;
    a.run_mean = g->run_mean; a.run_var = g->run_var; a.nbt = g->nbt;
    a.momentum = momentum; a.bn_eps = bn_eps; a.eps = eps; a.training = training;
    a.mu = mu; a.sd = sd; a.gate = gate; a.shat = shat; a.r = r;
    a.pairs = reinterpret_cast<float2*>(scratch_floats);             // 2*N*C floats, 8-byte aligned
    cudaError_t e = cudaSuccess;
    if (training) e = cudaMemsetAsync(a.pairs, 0xff, (size_t)N * C * sizeof(float2), stream);   // sentinel fill
    if (e != cudaSuccess) return (int)e;
    a.trace = nullptr;
    const char* trace_path = getenv("CNSN_FUSED_TRACE");            // debug: dump per-group timestamps
    const size_t trace_bytes = (size_t)a.sch.B * a.sch.G * 8 * sizeof(unsigned long long);
    if (trace_path) {
        if (cudaMalloc(&a.trace, trace_bytes) != cudaSuccess) a.trace = nullptr;
        else cudaMemsetAsync(a.trace, 0, trace_bytes, stream);
    }
    void* args[] = {&a};
    const void* fn = nullptr;
    switch (dtype) {
        case CNSN_F32: fn = (const void*)k_sn_fused_fwd<float>; break;
        case CNSN_BF16: fn = (const void*)k_sn_fused_fwd<__nv_bfloat16>; break;
        case CNSN_F16: fn = (const void*)k_sn_fused_fwd<__half>; break;
        default: return CNSN_E_BADARG;
    }
    e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_total);
    if (e != cudaSuccess) return (int)e;
    e = cudaLaunchCooperativeKernel(fn, dim3(a.sch.B), dim3(kThreads), args, smem_total, stream);
    note_launch();
    if (a.trace) {                                                   // debug path only: synchronous
        cudaStreamSynchronize(stream);
        unsigned long long* h = (unsigned long long*)malloc(trace_bytes);
        cudaMemcpy(h, a.trace, trace_bytes, cudaMemcpyDeviceToHost);
        if (FILE* f = fopen(trace_path, "wb")) {
            const int hdr[4] = {a.sch.G, a.sch.S, a.sch.B, a.sch.kk};
            fwrite(hdr, sizeof(int), 4, f);
            fwrite(h, 1, trace_bytes, f);
            fclose(f);
        }
        free(h);
        cudaFree(a.trace);
    }
    return (int)e;
}

}  // namespace fused
}  // namespace cnsn
