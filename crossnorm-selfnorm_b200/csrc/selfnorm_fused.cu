// selfnorm_fused.cu -- SelfNorm forward and backward as ONE persistent, warp-specialised kernel each
// (sm_100a): a TMA-fed reduce stream and an L2-fed apply stream running concurrently in every CTA.
//
// Why: the gate of channel c needs the statistics of ALL N instances of c (BatchNorm1d over the
// batch, models/cnsn.py:121,:138), so no element of an instance can be written before every other
// instance of its channel has been reduced.  A stats-then-apply kernel pair therefore reads x twice
// from HBM (3*S of traffic for 2*S algorithmic; 5*S for 3*S in backward).  Here both phases run in
// one kernel, a few channel groups apart, so the second read is an L2 hit: HBM traffic 2*S / 3*S.
//
// One CTA per SM (cooperative launch => all co-resident), 24 warps:
//   producer (1 warp)   TMA 1-D bulk loads (cp.async.bulk + mbarrier complete_tx) of whole units (one
//                       sample's run of kk channels; backward: of x and of dy) into an S-stage shared-
//                       memory ring.  ~170 KB in flight per SM is what keeps HBM busy with 8 consumer
//                       warps.  Throttled to stay at most kSlots groups ahead of the apply stream so the
//                       planes are still in L2 when they are read again.
//   reduce   (8 warps)  unit j of ring stage st -> warp (st*upc+j)%8: per-instance statistics out of shared
//                       memory (forward: one-pass shifted-data mean / variance; backward: sum dy*x, sum
//                       dy), published as ONE aligned 8-byte word per instance ("data is the flag").  The
//                       ring slot is released as soon as it has been reduced.
//   channel  (3 warps)  apply slot sl -> warp sl%3: polls the N*kk published words of a group (every CTA
//                       does this redundantly: 8 B per instance out of L2), reduces over N (forward: the
//                       BatchNorm batch mean / rstd of s = w0*mu + w1*sd; backward: dgamma, dbeta and the
//                       two scalars of the batch-norm backward), stages the channel constants in shared
//                       memory; the channel's owner CTA also writes running statistics / parameter grads.
//   apply    (12 warps) unit j of apply slot sl -> warp (sl*upc+j)%12; rebuilds the instance's gate (or its
//                       backward coefficients), re-reads the planes with 128-bit loads -- L2 hits, marked
//                       evict-first -- and streams the result out with 128-bit streaming stores.
//
// Cross-CTA latency chain per group: one 8-byte store, one polled 8-byte load.  Deadlock freedom:
// producer(g) waits for reduce(g-S) and apply(g-kSlots); apply(g) waits for channel(g), which waits for
// every CTA's reduce(g), which waits for that CTA's producer(g): every wait points to a smaller group
// index or an earlier role of the same group.  Every wait is bounded and traps instead of hanging.
#include <stdio.h>
#include <stdlib.h>

#include "fused_common.cuh"

namespace cnsn {
namespace fused {

struct Args {
    Schedule sch;
    const void* x; const void* dy; void* out;      // forward: dy == nullptr, out = y; backward: out = dx
    const float* w; const float* gamma; const float* beta;
    float* run_mean; float* run_var; long long* nbt;
    float momentum, bn_eps, eps;
    int training;
    float* mu; float* sd; float* gate; float* shat; float* r;    // save block (written by forward, read by backward)
    float* dw; float* dgamma; float* dbeta;                       // backward outputs
    float2* pairs;          // [C][N] published per-instance words, pre-filled with the sentinel
    unsigned stage_bytes;   // ring slot size (multiple of 128)
    unsigned off_chan, off_pair, off_data;   // smem offsets
    int keep_l2;            // experiment: evict-last policy on the ring loads
    int dbg;                // experiment bits: 1 = apply warps skip the data loop, 2 = channel warps skip polling
    unsigned long long* trace;   // debug only (CNSN_FUSED_TRACE): [B][G][8] globaltimer stamps, else NULL
};

struct ChanMeta { float a, b, c, d, e, f; };   // fwd: m, rstd, gamma, beta, w0, w1 ; bwd: k1, k2, gamma, rstd, w0, w1

__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Trace slots per group: 0 load issued, 1 data landed (reduce start), 2 reduced + published, 3 words complete,
// 4 chan_ready, 5 apply start, 6 apply done.
#define CNSN_TRACE(slot, g)                                                                   \
    do {                                                                                      \
        if (a.trace) a.trace[((size_t)b * G + (g)) * 8 + (slot)] = gtime();                    \
    } while (0)

__device__ __forceinline__ float2 wait_word(const float2* p) {
    float2 v = ll_peek(p);
    unsigned spins = 0;
    while (!ll_valid(v)) {
        __nanosleep(64);
        v = ll_peek(p);
        if (++spins > kSpinLimit) __trap();
    }
    return v;
}

// sum dy*x and sum dy of smem-resident instances (backward reduce), lpi lanes per instance.
template <typename T>
__device__ __forceinline__ float2 smem_dot(const T* ix, const T* id, int M, int r, int lpi, bool vec, bool live) {
    float s0 = 0.f, s1 = 0.f;
    if (live) {
        if (vec) {
            constexpr int V = VecOf<T>::n;
            const uint4* px = reinterpret_cast<const uint4*>(ix);
            const uint4* pd = reinterpret_cast<const uint4*>(id);
            const int nv = M / V;
            float a0[2] = {0.f, 0.f}, a1[2] = {0.f, 0.f};
            const int step = lpi * 4;
            const int nfull = (nv / step) * step;
            int i = r;
            for (; i < nfull; i += step) {                   // full batches: no bounds checks (see fused_common.cuh)
                uint4 rx[4], rd[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) { rx[u] = px[i + u * lpi]; rd[u] = pd[i + u * lpi]; }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    float vx[V], vd[V];
                    unpack<T>(rx[u], vx);
                    unpack<T>(rd[u], vd);
#pragma unroll
                    for (int e = 0; e < V; ++e) { a0[e & 1] = fmaf(vd[e], vx[e], a0[e & 1]); a1[e & 1] += vd[e]; }
                }
            }
            for (; i < nv; i += lpi) {
                float vx[V], vd[V];
                unpack<T>(px[i], vx);
                unpack<T>(pd[i], vd);
#pragma unroll
                for (int e = 0; e < V; ++e) { a0[e & 1] = fmaf(vd[e], vx[e], a0[e & 1]); a1[e & 1] += vd[e]; }
            }
            s0 = a0[0] + a0[1]; s1 = a1[0] + a1[1];
        } else {
            for (int i = r; i < M; i += lpi) { const float d = to_f(id[i]); s0 = fmaf(d, to_f(ix[i]), s0); s1 += d; }
        }
    }
    for (int o = lpi >> 1; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    }
    return make_float2(s0, s1);
}

template <typename T, bool BWD>
__global__ void __launch_bounds__(kThreads, 1) k_sn_fused(const Args a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const Schedule& s = a.sch;
    const int S = s.S, kk = s.kk, B = s.B, C = s.C, M = s.M, G = s.G, N = s.N;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);           // [kMaxStages]  TMA data landed
    uint64_t* empty = full + kMaxStages;                          // [kMaxStages]  stage reduced, may be refilled
    uint64_t* chan_ready = empty + kMaxStages;                    // [kSlots]      channel constants staged
    uint64_t* applied = chan_ready + kSlots;                      // [kSlots]      group applied, slot reusable
    volatile int* issued = reinterpret_cast<volatile int*>(applied + kSlots);   // last group the producer issued
    ChanMeta* chan_meta = reinterpret_cast<ChanMeta*>(smem + a.off_chan);       // [kSlots][kk]
    float2* pair_buf = reinterpret_cast<float2*>(smem + a.off_pair);            // [kChanWarps][kk*N]
    unsigned char* data = smem + a.off_data;                                    // [S][stage_bytes]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, b = blockIdx.x;
    if (threadIdx.x == 0) {
        for (int i = 0; i < S; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], s.upc); }
        for (int i = 0; i < s.R; ++i) { mbar_init(&chan_ready[i], 1); mbar_init(&applied[i], s.upc * kApplyTeam); }
        *issued = -1;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const T* x = static_cast<const T*>(a.x);
    const T* dy = static_cast<const T*>(a.dy);
    T* out = static_cast<T*>(a.out);
    constexpr int V = VecOf<T>::n;
    const bool vec = ((size_t)M * sizeof(T)) % 16 == 0;
    const int nv = M / V;
    const unsigned unit_bytes = s.unit_elems * (unsigned)sizeof(T);
    const unsigned dy_off = s.upc * unit_bytes;                                // backward: dy units follow the x units
    const int lpi = s.lpi, ipw = 32 / lpi, sub = lane / lpi, r = lane % lpi;   // sub-warp teams
    const int total = N * kk, nslots = (total + 31) >> 5;                      // words per group (<= kMaxPairs)
    const bool getenv_keep = a.keep_l2 != 0;

    if (warp == kWarpProducer) {
        if (lane == 0) {
            if (!BWD && b == 0 && a.training && a.nbt) *a.nbt += 1;
            const bool keep = getenv_keep;                   // evict-last on the ring loads (experiment)
            const uint64_t pol_keep = l2_policy_evict_last();
            GroupIter it;
            for (it.init(s, b); it.g < G; it.next(s)) {
                const int g = it.g, st = it.st, first = it.first, cnt = it.cnt;
                mbar_wait(&empty[st], it.ph ^ 1);            // stage reduced (group g-S)
                mbar_wait(&applied[it.sl], it.sp ^ 1);       // group g-kSlots applied: bounds the L2 working set
                CNSN_TRACE(0, g);
                if (cnt == 0) { mbar_arrive(&full[st]); }
                else {
                    mbar_arrive_expect_tx(&full[st], cnt * unit_bytes * (BWD ? 2 : 1));
                    unsigned char* dst = data + (size_t)st * a.stage_bytes;
                    for (int j = 0; j < cnt; ++j) {
                        const size_t n = (size_t)first + (size_t)j * B;
                        const size_t off = (n * C + (size_t)g * kk) * M;
                        if (keep) {
                            tma_load_1d(dst + (size_t)j * unit_bytes, x + off, unit_bytes, &full[st], pol_keep);
                            if (BWD) tma_load_1d(dst + dy_off + (size_t)j * unit_bytes, dy + off, unit_bytes, &full[st], pol_keep);
                        } else {
                            tma_load_1d_plain(dst + (size_t)j * unit_bytes, x + off, unit_bytes, &full[st]);
                            if (BWD) tma_load_1d_plain(dst + dy_off + (size_t)j * unit_bytes, dy + off, unit_bytes, &full[st]);
                        }
                    }
                }
                *issued = g;
            }
        }
    } else if (warp < kStatsWarps) {
        // ---------------------------------------------------------------- reduce stream
        GroupIter it;
        for (it.init(s, b); it.g < G; it.next(s)) {
            const int g = it.g, st = it.st, first = it.first, cnt = it.cnt;
            bool any = false;
            for (int j = 0; j < s.upc; ++j) any |= s.my_unit(st, j, warp, kStatsWarps);
            if (!any) continue;
            mbar_wait(&full[st], it.ph);
            if (lane == 0 && s.my_unit(st, 0, warp, kStatsWarps)) CNSN_TRACE(1, g);
            const T* base = reinterpret_cast<const T*>(data + (size_t)st * a.stage_bytes);
            const T* based = reinterpret_cast<const T*>(data + (size_t)st * a.stage_bytes + dy_off);
            for (int j = 0; j < s.upc; ++j) {
                if (!s.my_unit(st, j, warp, kStatsWarps)) continue;
                if (j < cnt) {
                    const size_t n = (size_t)first + (size_t)j * B;
                    for (int c0 = 0; c0 < kk; c0 += ipw) {
                        const int cl = c0 + sub;
                        const bool live = cl < kk;
                        const size_t inst = (size_t)(j * kk + cl) * M;
                        float2 word;
                        if (BWD) {
                            word = smem_dot<T>(base + inst, based + inst, M, r, lpi, vec, live);
                        } else {
                            const float2 mq = smem_mean_m2<T>(base + inst, M, r, lpi, vec, live);
                            word = make_float2(mq.x, sqrtf(mq.y / (float)(M - 1) + a.eps));
                        }
                        if (live && r == 0) ll_publish(a.pairs + ((size_t)g * kk + cl) * N + n, word.x, word.y);
                    }
                }
                __syncwarp();
                if (lane == 0) {
                    if (j == 0) CNSN_TRACE(2, g);
                    mbar_arrive(&empty[st]);                 // one arrival per unit slot: the stage can be refilled
                }
            }
        }
    } else if (warp < kStatsWarps + kApplyWarps) {
        // ---------------------------------------------------------------- apply stream
        const int aw = warp - kStatsWarps;
        const int team = aw / kApplyTeam, q = aw - team * kApplyTeam;      // unit j of slot sl -> team (sl*upc+j) % teams
        const uint64_t pol = l2_policy_evict_first();
        const float invM = 1.f / M, invM1 = 1.f / (M - 1.f);
        GroupIter it;
        for (it.init(s, b); it.g < G; it.next(s)) {
            const int g = it.g, sl = it.sl, first = it.first, cnt = it.cnt;
            bool any = false;
            for (int j = 0; j < s.upc; ++j) any |= s.my_unit(sl, j, team, kApplyTeams);
            if (!any) continue;
            mbar_wait(&chan_ready[sl], it.sp);
            if (lane == 0 && q == 0 && s.my_unit(sl, 0, team, kApplyTeams)) CNSN_TRACE(5, g);
            // kk == 1: the team's warps each take a contiguous third of the plane; kk > 1: whole instances
            // are dealt to the team's warps (and to sub-warp lane groups inside a warp).
            const bool cut = (kk == 1) && vec;
            const int per = (nv + kApplyTeam - 1) / kApplyTeam;
            const int vlo = cut ? q * per : 0, vhi = cut ? min(nv, vlo + per) : nv;
            for (int j = 0; j < s.upc; ++j) {
                if (!s.my_unit(sl, j, team, kApplyTeams)) continue;
                if (j < cnt) {
                    const size_t n = (size_t)first + (size_t)j * B;
                    for (int cl = (cut ? 0 : q * ipw) + sub; cl < kk; cl += (cut ? 1 : kApplyTeam) * ipw) {
                        const int c = g * kk + cl;
                        const size_t nc = n * C + c;
                        const ChanMeta cm = chan_meta[sl * kk + cl];
                        const float2 own = wait_word(a.pairs + (size_t)c * N + n);     // published long ago
                        float ca, cb, cc;                    // out = ca*dy + cb*x + cc   (forward: ca unused)
                        if (BWD) {
                            const float gt = a.gate[nc], sh = a.shat[nc], mean = a.mu[nc], sdev = a.sd[nc];
                            const float ds = cm.d * (own.x * gt * (1.f - gt) * cm.c - cm.a - sh * cm.b);
                            ca = gt;
                            cb = ds * cm.f * invM1 / sdev;
                            cc = ds * cm.e * invM - cb * mean;
                        } else {
                            const float sh = (fmaf(cm.e, own.x, cm.f * own.y) - cm.a) * cm.b;
                            const float gt = 1.f / (1.f + expf(-fmaf(cm.c, sh, cm.d)));
                            if (r == 0 && (!cut || q == 0)) { a.mu[nc] = own.x; a.sd[nc] = own.y; a.gate[nc] = gt; a.shat[nc] = sh; }
                            ca = 0.f; cb = gt; cc = 0.f;
                        }
                        const T* sx = x + nc * M;
                        const T* sd_ = BWD ? dy + nc * M : nullptr;
                        T* dst = out + nc * M;
                        if (a.dbg & 1) continue;
                        if (vec) {
                            const uint4* px = reinterpret_cast<const uint4*>(sx);
                            const uint4* pd = reinterpret_cast<const uint4*>(sd_);
                            uint4* po = reinterpret_cast<uint4*>(dst);
                            constexpr int U = BWD ? 4 : 8;   // independent 128-bit loads in flight per lane and tensor
                            // Full-duty register ring: U loads per tensor stay in flight per lane; a slot's register is
                            // refilled with slot s+U as soon as it has been consumed.  Out-of-range slots are clamped
                            // for the load and predicated for the store (no branch regions in the loop).
                            const int span = vhi - vlo;
                            const int nsl = (span + lpi - 1) / lpi;                   // slots per lane
                            uint4 rx[U], rd[U];
#pragma unroll
                            for (int u = 0; u < U; ++u) {
                                const int v = vlo + min(r + lpi * u, span - 1);
                                rx[u] = ldg_hint(px + v, pol);                        // L2 hit, last use
                                if (BWD) rd[u] = ldg_hint(pd + v, pol);
                            }
                            for (int s0 = 0; s0 < nsl; s0 += U) {
#pragma unroll
                                for (int u = 0; u < U; ++u) {
                                    const int o = r + lpi * (s0 + u);
                                    float vx[V], vd[V], vo[V];
                                    unpack<T>(rx[u], vx);
                                    if (BWD) unpack<T>(rd[u], vd);
#pragma unroll
                                    for (int e = 0; e < V; ++e)
                                        vo[e] = BWD ? fmaf(ca, vd[e], fmaf(cb, vx[e], cc)) : vx[e] * cb;
                                    if (o < span) stg_stream(po + vlo + o, pack<T>(vo));
                                    if (s0 + u + U < nsl) {                           // warp-uniform
                                        const int vn = vlo + min(o + lpi * U, span - 1);
                                        rx[u] = ldg_hint(px + vn, pol);
                                        if (BWD) rd[u] = ldg_hint(pd + vn, pol);
                                    }
                                }
                            }
                        } else {
                            for (int e = r; e < M; e += lpi)
                                dst[e] = from_f<T>(BWD ? fmaf(ca, to_f(sd_[e]), fmaf(cb, to_f(sx[e]), cc)) : to_f(sx[e]) * cb);
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) {
                    if (j == 0 && q == 0) CNSN_TRACE(6, g);
                    mbar_arrive(&applied[sl]);               // one arrival per (unit slot, warp of the team)
                }
            }
        }
    } else if (warp >= kWarpChan) {
        // ---------------------------------------------------------------- channel warps
        const int cw = warp - kWarpChan;
        float2* my_buf = pair_buf + (size_t)cw * total;
        const float invN = 1.f / N;
        GroupIter it;
        for (it.init(s, b); it.g < G; it.next(s)) {
            const int g = it.g, sl = it.sl;
            if (sl % kChanWarps != cw) continue;
            float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f, prm = 0.f, prv = 1.f;   // lane cl < kk: channel constants
            if (lane < kk) {                                 // issue the loads now, they are consumed after the poll
                const int c = g * kk + lane;
                p0 = a.w[2 * c]; p1 = a.w[2 * c + 1]; p2 = a.gamma[c];
                if (BWD) p3 = a.r[c];
                else {
                    p3 = a.beta[c];
                    if (!a.training || c % B == b) { prm = a.run_mean[c]; prv = a.run_var[c]; }
                }
            }
            mbar_wait(&applied[sl], it.sp ^ 1);              // chan_meta[sl] free (group g-kSlots applied)
            {   // do not poll for groups this CTA has not even started to load
                unsigned spins = 0;
                while (*issued < g) { __nanosleep(200); if (++spins > kWaitLimit) __trap(); }
            }
            float ra = 0.f, rb = 0.f, rq = 0.f, rw0 = 0.f, rw1 = 0.f;   // lane cl: per-channel results
            if ((BWD || a.training) && !(a.dbg & 2)) {
                const float2* pbase = a.pairs + (size_t)g * kk * N;   // [kk][N], contiguous
                unsigned pend = 0;                           // bit j: word lane+32j not yet seen
                for (int j = 0; j < nslots; ++j) if (lane + 32 * j < total) pend |= 1u << j;
                unsigned spins = 0;
                while (true) {
                    // issue every outstanding poll load first, then test: one L2 round trip per sweep
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
                    const float2* pb = my_buf + cl * N;
                    const int c = g * kk + cl;
                    if (!BWD) {                              // batch mean / biased variance of s = w0*mu + w1*sd
                        const float w0 = __shfl_sync(0xffffffffu, p0, cl), w1 = __shfl_sync(0xffffffffu, p1, cl);
                        float sum = 0.f;
                        for (int n = lane; n < N; n += 32) sum += fmaf(w0, pb[n].x, w1 * pb[n].y);
                        const float m = warp_sum(sum) * invN;
                        float q = 0.f;
                        for (int n = lane; n < N; n += 32) { const float d = fmaf(w0, pb[n].x, w1 * pb[n].y) - m; q = fmaf(d, d, q); }
                        q = warp_sum(q) * invN;
                        if (lane == cl) { ra = m; rq = q; rb = 1.f / sqrtf(q + a.bn_eps); }
                    } else {                                 // dgamma, dbeta -> k1, k2; owner also dw
                        const float ga = __shfl_sync(0xffffffffu, p2, cl), rstd = __shfl_sync(0xffffffffu, p3, cl);
                        float sg = 0.f, sb = 0.f;
                        for (int n = lane; n < N; n += 32) {
                            const size_t nc = (size_t)n * C + c;
                            const float gt = a.gate[nc];
                            const float dz = pb[n].x * gt * (1.f - gt);
                            sg = fmaf(dz, a.shat[nc], sg); sb += dz;
                        }
                        sg = warp_sum(sg); sb = warp_sum(sb);
                        const float k1 = a.training ? ga * sb * invN : 0.f, k2 = a.training ? ga * sg * invN : 0.f;
                        float t0 = 0.f, t1 = 0.f;
                        if (c % B == b) {                    // warp-uniform: this CTA owns the channel's parameter grads
                            for (int n = lane; n < N; n += 32) {
                                const size_t nc = (size_t)n * C + c;
                                const float gt = a.gate[nc];
                                const float ds = rstd * (pb[n].x * gt * (1.f - gt) * ga - k1 - a.shat[nc] * k2);
                                t0 = fmaf(ds, a.mu[nc], t0); t1 = fmaf(ds, a.sd[nc], t1);
                            }
                            t0 = warp_sum(t0); t1 = warp_sum(t1);
                        }
                        if (lane == cl) { ra = k1; rb = k2; rq = sg; rw0 = t0; rw1 = t1; prm = sb; }
                    }
                }
                __syncwarp();
            }
            if (lane < kk) {
                const int c = g * kk + lane;
                ChanMeta cm;
                if (BWD) {
                    cm.a = ra; cm.b = rb; cm.c = p2; cm.d = p3; cm.e = p0; cm.f = p1;
                    if (c % B == b) { a.dgamma[c] = rq; a.dbeta[c] = prm; a.dw[2 * c] = rw0; a.dw[2 * c + 1] = rw1; }
                } else {
                    if (!a.training) { ra = prm; rb = 1.f / sqrtf(prv + a.bn_eps); }
                    cm.a = ra; cm.b = rb; cm.c = p2; cm.d = p3; cm.e = p0; cm.f = p1;
                    if (c % B == b) {                        // this CTA owns the channel's bookkeeping
                        a.r[c] = rb;
                        if (a.training) {
                            a.run_mean[c] = (1.f - a.momentum) * prm + a.momentum * ra;
                            a.run_var[c] = (1.f - a.momentum) * prv + a.momentum * (rq * N / (N - 1.f));
                        }
                    }
                }
                chan_meta[sl * kk + lane] = cm;
            }
            __syncwarp();
            if (lane == 0) { CNSN_TRACE(4, g); mbar_arrive(&chan_ready[sl]); }
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
    if (const char* e = getenv("CNSN_FUSED_STAGES")) { (void)e; }
    if (const char* e = getenv("CNSN_FUSED_CTAS")) {                 // experiment knob: CTAs (<= SM count)
        const int v = atoi(e);
        if (v > 0 && v <= d.sms) d.sms = v;
    }
    return d;
}

static int pick_rot(int B) {
    static const int primes[] = {61, 59, 53, 47, 43, 41, 37, 31, 29, 23, 19, 17, 13, 11, 7, 5, 3};
    for (int p : primes) if (p < B && B % p != 0) return p;
    return 1;
}

// Ring geometry; tensors = 1 (forward: x) or 2 (backward: x and dy).
static bool make_plan(int N, int C, int M, int dtype, int tensors, const DeviceInfo& d, Args& a, unsigned& smem_total) {
    Schedule& s = a.sch;
    // Measured (profiles/README.md, r01 A/B table): the fused kernels beat the three-kernel path only for
    // large planes (one instance per unit); elsewhere the three-kernel path is used unless forced.
    if ((size_t)M * esize(dtype) < 8192 && !getenv("CNSN_FUSED_FORCE")) return false;
    if (d.sms <= 0 || !d.coop || N < 2 || N > kMaxPairs) return false;
    const int B = d.sms;
    const int upc = (N + B - 1) / B;
    if (upc > kStatsWarps || upc > kApplyTeams) return false;                                   // unit slots of a stage map to distinct warps
    const size_t esz = esize(dtype), inst_bytes = (size_t)M * esz;
    if ((size_t)N * C * inst_bytes < ((size_t)8 << 20)) return false;       // small tensors: the 3-kernel path
    const size_t budget = (size_t)d.smem_optin - 2048;
    const int min_stages = 3;
    int best = 0;
    for (int kk = 1; kk <= kMaxKK && kk <= C; ++kk) {
        if (C % kk) continue;
        if ((kk * inst_bytes) % 16) continue;
        if ((size_t)N * kk > (size_t)kMaxPairs) break;
        const size_t sb = ((size_t)upc * kk * inst_bytes * tensors + 127) & ~(size_t)127;
        const size_t fixed = (size_t)kSlots * kk * sizeof(ChanMeta) + (size_t)kChanWarps * N * kk * sizeof(float2);
        if (sb * min_stages + fixed > budget) break;
        best = kk;
        if (kk * inst_bytes >= 8192) break;                                  // big enough units
    }
    if (!best) return false;
    const int kk = best;
    s.N = N; s.C = C; s.M = M; s.kk = kk; s.G = C / kk; s.B = B; s.upc = upc; s.rot = pick_rot(B);
    s.unit_elems = (unsigned)(kk * M);
    s.lpi = kk >= 8 ? 4 : kk >= 4 ? 8 : kk >= 2 ? 16 : 32;                   // run up to 8 instances of a unit side by side
    {   // Apply-stream slots: bound the L2 working set (planes loaded but not yet applied) to ~1/3 of L2.
        int dev = 0, l2 = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, dev);
        const size_t group_bytes = (size_t)N * kk * inst_bytes * tensors;
        int R = (int)(((size_t)l2 / 3) / (group_bytes ? group_bytes : 1));
        if (const char* e = getenv("CNSN_FUSED_SLOTS")) { const int v = atoi(e); if (v > 0) R = v; }
        R = (R / kChanWarps) * kChanWarps;
        if (R < 2 * kChanWarps) R = 2 * kChanWarps;
        if (R > kSlots) R = kSlots;
        s.R = R;
    }
    a.stage_bytes = (unsigned)(((size_t)upc * kk * inst_bytes * tensors + 127) & ~(size_t)127);
    const unsigned hdr = (2 * kMaxStages + 2 * kSlots) * 8 + 64;            // barriers + issued
    a.off_chan = (hdr + 15) & ~15u;
    a.off_pair = (a.off_chan + (unsigned)(kSlots * kk * sizeof(ChanMeta)) + 15) & ~15u;
    a.off_data = (a.off_pair + (unsigned)((size_t)kChanWarps * N * kk * sizeof(float2)) + 127) & ~127u;
    int S = (int)(((size_t)d.smem_optin - a.off_data) / a.stage_bytes);
    if (S > kMaxStages) S = kMaxStages;
    if (const char* e = getenv("CNSN_FUSED_STAGES")) { const int v = atoi(e); if (v >= 2 && v < S) S = v; }
    if (S > s.G) S = s.G;
    if (S < min_stages && S < s.G) return false;
    s.S = S;
    smem_total = a.off_data + (unsigned)S * a.stage_bytes;
    return smem_total <= (unsigned)d.smem_optin;
}

template <bool BWD>
static int launch(Args& a, int dtype, unsigned smem_total, cudaStream_t stream) {
    cudaError_t e = cudaMemsetAsync(a.pairs, 0xff, (size_t)a.sch.N * a.sch.C * sizeof(float2), stream);   // sentinel fill
    if (e != cudaSuccess) return (int)e;
    a.trace = nullptr;
    a.keep_l2 = getenv("CNSN_FUSED_KEEP") ? 1 : 0;
    a.dbg = getenv("CNSN_FUSED_DBG") ? atoi(getenv("CNSN_FUSED_DBG")) : 0;
    const char* trace_path = BWD ? getenv("CNSN_FUSED_TRACE_BWD") : getenv("CNSN_FUSED_TRACE");   // debug: per-group timestamps
    const size_t trace_bytes = (size_t)a.sch.B * a.sch.G * 8 * sizeof(unsigned long long);
    if (trace_path) {
        if (cudaMalloc(&a.trace, trace_bytes) != cudaSuccess) a.trace = nullptr;
        else cudaMemsetAsync(a.trace, 0, trace_bytes, stream);
    }
    void* args[] = {&a};
    const void* fn = nullptr;
    switch (dtype) {
        case CNSN_F32: fn = (const void*)k_sn_fused<float, BWD>; break;
        case CNSN_BF16: fn = (const void*)k_sn_fused<__nv_bfloat16, BWD>; break;
        case CNSN_F16: fn = (const void*)k_sn_fused<__half, BWD>; break;
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

// Both return 0 when launched, >0 cuda error, -100 when the fused path does not apply (the caller
// falls back to the three-kernel path).
int selfnorm_fused_fwd(const void* x, void* y, int dtype, int N, int C, int H, int W,
                       const cnsn_gate_params* g, int training, float momentum, float bn_eps, float eps,
                       float* mu, float* sd, float* gate, float* shat, float* r, float* scratch_floats,
                       cudaStream_t stream) {
    if (!aligned16(x) || !aligned16(y)) return -100;
    Args a{};
    unsigned smem_total = 0;
    if (!make_plan(N, C, H * W, dtype, 1, device_info(), a, smem_total)) return -100;
    a.x = x; a.dy = nullptr; a.out = y; a.w = g->w; a.gamma = g->gamma; a.beta = g->beta;
    a.run_mean = g->run_mean; a.run_var = g->run_var; a.nbt = g->nbt;
    a.momentum = momentum; a.bn_eps = bn_eps; a.eps = eps; a.training = training;
    a.mu = mu; a.sd = sd; a.gate = gate; a.shat = shat; a.r = r;
    a.pairs = reinterpret_cast<float2*>(scratch_floats);             // 2*N*C floats, 8-byte aligned
    return launch<false>(a, dtype, smem_total, stream);
}

int selfnorm_fused_bwd(const void* x, const void* dy, void* dx, int dtype, int N, int C, int H, int W,
                       const cnsn_gate_params* g, int training,
                       float* mu, float* sd, float* gate, float* shat, float* r,
                       const cnsn_gate_grads* dg, float* scratch_floats, cudaStream_t stream) {
    if (!aligned16(x) || !aligned16(dy) || !aligned16(dx)) return -100;
    Args a{};
    unsigned smem_total = 0;
    if (!make_plan(N, C, H * W, dtype, 2, device_info(), a, smem_total)) return -100;
    a.x = x; a.dy = dy; a.out = dx; a.w = g->w; a.gamma = g->gamma; a.beta = nullptr;
    a.training = training;
    a.mu = mu; a.sd = sd; a.gate = gate; a.shat = shat; a.r = r;
    a.dw = dg->dw; a.dgamma = dg->dgamma; a.dbeta = dg->dbeta;
    a.pairs = reinterpret_cast<float2*>(scratch_floats);
    return launch<true>(a, dtype, smem_total, stream);
}

}  // namespace fused
}  // namespace cnsn
