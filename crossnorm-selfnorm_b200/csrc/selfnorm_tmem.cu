// selfnorm_tmem.cu -- SelfNorm forward / backward with planes resident in shared memory AND tensor memory (sm_100a).
//
// What bounds the shared-memory-resident SelfNorm kernels (selfnorm_flow.cu) is on-chip capacity x item lifetime: a
// plane has to stay on chip from the moment it is fetched until its channel's constants exist (BatchNorm1d over the
// batch couples the N instances of a channel), ~9 us of which ~5 are the wait, and 148 x 227 KB of shared memory divided
// by that lifetime is less than HBM delivers.  Measured with the persistent grid capped (gpurun_out/r2l_cap.log,
// (256,256,56,56) fp32 forward): 18 / 15 / 12 / 9 / 6 planes per SM -> 0.314 / 0.338 / 0.373 / 0.438 / 0.576 ms, i.e.
// t = 0.16 ms + 0.83 ms / (CTAs per SM): more capacity is worth more time until HBM itself binds (0.25 ms).
//
// Blackwell has a second on-chip memory of the same size that this path leaves idle: TENSOR MEMORY (256 KB per SM,
// 128 lanes x 512 columns x 32 bit, tcgen05.alloc / .st / .ld).  There is no matrix product here to put into it -- it is
// used as what it physically is, 256 KB of fast on-chip storage: with the 32x32b access shape a thread owns one TMEM
// lane, i.e. a private row of up to 128 x 32-bit cells per CTA slice, enough for its share of 4 planes.
//
// ONE CTA per SM (a kernel that allocates tensor memory gets an occupancy of one CTA per SM from the driver, and the
// cooperative launch that guarantees co-residency goes by that number), 512 threads = FOUR independent groups of 128
// (named barriers, own mbarrier, own quarter of the shared memory, own 128 of the 512 TMEM columns, own ticket stream).
// Every group runs a two-stage software pipeline over its tickets:
//
//     NEW item (ticket t)   : cp.async.bulk (TMA) its P planes into shared memory, reduce them there (exact two-pass mean /
//                             variance, or sum dy*x), PUBLISH the per-instance words -- before anything is waited for
//     OLD item (ticket t')  : sits in tensor memory since the previous iteration; by now its channel has usually been
//                             folded: poll the channel word, rebuild gates / coefficients, apply straight out of tensor
//                             memory (tcgen05.ld -> registers -> st.global)
//     then the NEW item moves shared memory -> tensor memory (ld.shared -> tcgen05.st), shared memory is handed to the TMA
//     load of the next ticket, NEW becomes OLD.
//
// A plane therefore spends its load + reduce time in shared memory and its wait + apply time in tensor memory, and the
// chip holds twice the planes: 4 groups x (4 + 4) = 32 planes of 12.5 KB per SM against 18.  Measured at
// (256,256,56,56) fp32: forward 0.314 -> 0.269 ms, backward 0.479 -> 0.384 ms (gpurun_out/r2o_sweep.log).
//
// Deadlock freedom is the argument of selfnorm_flow.cu plus one observation: an item is published as soon as its own
// bulk copy has landed and been reduced -- never behind a wait -- so every ticket that has been taken gets published;
// tickets are taken in order by groups that are all co-resident (cooperative launch), a channel's items are consecutive
// tickets, hence the lowest unfinished channel always completes.  A CTA blocked in tcgen05.alloc (another kernel holds
// columns on its SM) has not taken a ticket yet, so nobody waits for it.
//
// Handled here: training mode, planes of 6..16 KB (4..8 128-bit vectors per thread), no fused residual add; everything
// else stays with selfnorm_flow.cu.  cnsn_tune("tm", 0) disables this path (A/B, tests).
#include <stdio.h>

#include <algorithm>

#include "tmem_common.cuh"

namespace cnsn {
namespace flow {

template <typename T, bool BWD, int P, int SL>
__global__ void __launch_bounds__(kTmCta, 1) k_sn_tm(const FArgs a) {
    constexpr int TH = kTmT, V = VecOf<T>::n;
    constexpr int PL = BWD ? 2 : 1;                          // planes per instance: x [, dy]
    using S = Group128Sync;
    static_assert(P * PL * SL * 4 <= kTmCols, "an item must fit the group's TMEM slice");
    extern __shared__ __align__(128) unsigned char dsm_all[]; // [4 mbarriers | group 0: P planes of x, P of dy | group 1 ...]
    __shared__ unsigned s_tickets[kTmGroups], s_tmem;
    __shared__ float2 s_chans[kTmGroups];
    __shared__ float s_fs[kTmGroups][2][TH / 32];
    __shared__ float s_reds[kTmGroups][P][TH / 32];
    const int grp = threadIdx.x >> 7;
    const int tid = threadIdx.x & 127, warp = tid >> 5;
    const unsigned pbytes = (unsigned)a.M * (unsigned)sizeof(T);
    unsigned char* dsm = dsm_all + 128 + (size_t)grp * P * PL * pbytes - 128;   // dsm + 128 = this group's planes
    uint64_t* bar = reinterpret_cast<uint64_t*>(dsm_all) + grp;
    unsigned& s_ticket = s_tickets[grp];
    float2& s_chan = s_chans[grp];
    float (*s_f)[TH / 32] = s_fs[grp];
    float (*s_red)[TH / 32] = s_reds[grp];
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const uint32_t tbase = tmem_alloc_all(&s_tmem);
    // lane field (bits 31:16) = 32 * (warp % 4): the warp's quadrant; column field: this group's 128 columns
    const uint32_t trow = tbase + ((uint32_t)(warp & 3) << 21) + (uint32_t)(grp * kTmCols);

    const int N = a.N, C = a.C, M = a.M, nv = M / V;
    const unsigned nI = (unsigned)a.nI;
    const uint32_t sbase = smem_u32(dsm) + 128u;
    const bool relu = a.relu != 0;
    const float invM = 1.f / M, invM1 = 1.f / (M - 1.f);

    auto take = [&]() -> unsigned {
        if (tid == 0) s_ticket = atomicAdd(a.ticket, 1u) + 1u;       // the counter starts at 0xffffffff
        S::sync();
        const unsigned t = s_ticket;
        S::sync();
        return t;
    };
    auto issue = [&](unsigned t) {                           // the group's first warp: bulk copies of the item's planes, L2 prefetch ahead
        if (tid < 32) {
            const unsigned c = t / nI, j = t - c * nI;
            const int first = (int)j * P, nlive = min(P, N - first);
            const uint64_t pol = l2_policy_evict_first();
            if (tid == 0) mbar_arrive_expect_tx(bar, (unsigned)nlive * pbytes * PL);
            __syncwarp();
            for (int q = tid; q < nlive; q += 32) {
                const size_t off = ((size_t)(first + q) * C + c) * M;
                tma_load_1d(dsm + 128 + (size_t)q * pbytes, static_cast<const T*>(a.x) + off, pbytes, bar, pol);
                if (BWD) tma_load_1d(dsm + 128 + (size_t)(P + q) * pbytes, static_cast<const T*>(a.dy) + off, pbytes, bar, pol);
            }
            const unsigned tf = t + (unsigned)a.pf_dist;
            if (a.pf_dist && tf < a.items) {
                const unsigned cf = tf / nI, jf = tf - cf * nI;
                const int ff = (int)jf * P, nf = min(P, N - ff);
                for (int q = tid; q < nf; q += 32) {
                    const size_t off = ((size_t)(ff + q) * C + cf) * M;
                    tma_prefetch_l2(static_cast<const T*>(a.x) + off, pbytes);
                    if (BWD) tma_prefetch_l2(static_cast<const T*>(a.dy) + off, pbytes);
                }
            }
        }
    };

    // the OLD item: in tensor memory, published, waiting for its channel
    bool have_old = false;
    unsigned c_o = 0;
    int first_o = 0, nlive_o = 0;
    float w0_o = 0.f, w1_o = 0.f, ga_o = 0.f, b_o = 0.f;     // b: forward beta, backward r (1 / batch-norm std)
    float u_o[P], v_o[P];                                    // forward (mu, sd); backward the published (dz, shat)
    float g_o[P], mu_o[P], sd_o[P];                          // backward only: gate, mean, std of the instance
#pragma unroll
    for (int p = 0; p < P; ++p) { u_o[p] = 0.f; v_o[p] = 0.f; g_o[p] = 0.f; mu_o[p] = 0.f; sd_o[p] = 1.f; }
    unsigned par = 0;

    unsigned t_new = take();
    if (t_new < a.items) issue(t_new);
    for (;;) {
        const bool have_new = t_new < a.items;
        unsigned c = 0;
        int first = 0, nlive = 0;
        float w0 = 0.f, w1 = 0.f, ga = 0.f, pb = 0.f;
        float u_n[P], v_n[P], g_n[P], mu_n[P], sd_n[P];
#pragma unroll
        for (int p = 0; p < P; ++p) { u_n[p] = 0.f; v_n[p] = 0.f; g_n[p] = 0.f; mu_n[p] = 0.f; sd_n[p] = 1.f; }
        if (have_new) {
            // ================================================================ NEW: reduce out of shared memory, publish
            c = t_new / nI;
            const unsigned j = t_new - c * nI;
            first = (int)j * P;
            nlive = min(P, N - first);
            const bool folder = j == nI - 1;                 // holds the channel's last ticket
            w0 = a.w[2 * c]; w1 = a.w[2 * c + 1]; ga = a.gamma[c];
            float p_rm = 0.f, p_rv = 1.f;
            if (BWD) {
                pb = a.r[c];
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    g_n[p] = 0.f; v_n[p] = 0.f; mu_n[p] = 0.f; sd_n[p] = 1.f;
                    if (p < nlive) {
                        const size_t nc = (size_t)(first + p) * C + c;
                        g_n[p] = a.gate[nc]; v_n[p] = a.shat[nc]; mu_n[p] = a.mu[nc]; sd_n[p] = a.sd[nc];
                    }
                }
            } else {
                pb = a.beta[c];
                if (folder && tid == 0) { p_rm = a.run_mean[c]; p_rv = a.run_var[c]; }
            }
            mbar_wait(bar, par, a.err);
            par ^= 1u;
            float s[P];
#pragma unroll
            for (int p = 0; p < P; ++p) {
                float s0 = 0.f, s1 = 0.f;
                if (p < nlive) {
#pragma unroll
                    for (int sl = 0; sl < SL; ++sl) {
                        const int i = sl * TH + tid;
                        if (i < nv) {
                            float vx[V];
                            unpack<T>(lds128(sbase + (unsigned)p * pbytes + 16u * i), vx);
                            if (BWD) {
                                float vd[V];
                                unpack<T>(lds128(sbase + (unsigned)(P + p) * pbytes + 16u * i), vd);
#pragma unroll
                                for (int e = 0; e < V; ++e) {
                                    const float d = (relu && !(vx[e] > 0.f)) ? 0.f : vd[e];
                                    if (e & 1) s1 = fmaf(d, vx[e], s1); else s0 = fmaf(d, vx[e], s0);
                                }
                            } else {
#pragma unroll
                                for (int e = 0; e < V; ++e) { if (e & 1) s1 += vx[e]; else s0 += vx[e]; }
                            }
                        }
                    }
                }
                s[p] = s0 + s1;
            }
            cta_sums<P, TH, S>(s, s_red);
            if (BWD) {
#pragma unroll
                for (int p = 0; p < P; ++p) u_n[p] = s[p] * g_n[p] * (1.f - g_n[p]);
            } else {                                         // exact two-pass: second pass around the mean
#pragma unroll
                for (int p = 0; p < P; ++p) u_n[p] = s[p] * invM;
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    float s0 = 0.f, s1 = 0.f;
                    if (p < nlive) {
#pragma unroll
                        for (int sl = 0; sl < SL; ++sl) {
                            const int i = sl * TH + tid;
                            if (i < nv) {
                                float vx[V];
                                unpack<T>(lds128(sbase + (unsigned)p * pbytes + 16u * i), vx);
#pragma unroll
                                for (int e = 0; e < V; ++e) {
                                    const float d = vx[e] - u_n[p];
                                    if (e & 1) s1 = fmaf(d, d, s1); else s0 = fmaf(d, d, s0);
                                }
                            }
                        }
                    }
                    s[p] = s0 + s1;
                }
                cta_sums<P, TH, S>(s, s_red);
#pragma unroll
                for (int p = 0; p < P; ++p) v_n[p] = sqrtf(s[p] * invM1 + a.eps);
            }
#pragma unroll
            for (int p = 0; p < P; ++p) {
                if (tid == p && p < nlive) {
                    const int n = first + p;
                    if (!BWD) { const size_t nc = (size_t)n * C + c; a.mu[nc] = u_n[p]; a.sd[nc] = v_n[p]; }
                    ll_publish(a.pub + (size_t)c * N + n, u_n[p], v_n[p]);
                }
            }
            if (folder) fold_publish<BWD, TH, S>(a, c, a.chan + 4u * c, w0, w1, ga, pb, p_rm, p_rv, s_f);
        }
        if (have_old) {
            // ================================================================ OLD: channel constants, apply out of tensor memory
            if (tid == 0) s_chan = poll_word(a.chan + 4u * c_o, a.poll_ns, a.err);
            S::sync();
            const float2 cm = s_chan;
#pragma unroll
            for (int p = 0; p < P; ++p) {
                if (p < nlive_o) {
                    const size_t nc = (size_t)(first_o + p) * C + c_o;
                    float ca = 0.f, cb, cc = 0.f;            // out = ca*dy + cb*x + cc
                    if (BWD) {
                        const float ds = b_o * (u_o[p] * ga_o - cm.x - v_o[p] * cm.y);
                        ca = g_o[p];
                        cb = ds * w1_o * invM1 / sd_o[p];
                        cc = ds * w0_o * invM - cb * mu_o[p];
                    } else {
                        const float sh = (fmaf(w0_o, u_o[p], w1_o * v_o[p]) - cm.x) * cm.y;
                        cb = 1.f / (1.f + expf(-fmaf(ga_o, sh, b_o)));
                        if (tid == p) { a.gate[nc] = cb; a.shat[nc] = sh; }
                    }
                    uint4* po = reinterpret_cast<uint4*>(static_cast<T*>(a.out) + nc * M);
                    uint4 rx[SL], rd[SL];
#pragma unroll
                    for (int sl = 0; sl < SL; ++sl) {
                        rx[sl] = tmem_ld4(trow + (unsigned)((p * PL * SL + sl) * 4));
                        if constexpr (BWD) rd[sl] = tmem_ld4(trow + (unsigned)((p * PL * SL + SL + sl) * 4));
                        else rd[sl] = make_uint4(0u, 0u, 0u, 0u);
                    }
                    tmem_wait_ld(rx);
                    if constexpr (BWD) tmem_wait_ld(rd);
#pragma unroll
                    for (int sl = 0; sl < SL; ++sl) {
                        const int i = sl * TH + tid;
                        if (i < nv) {
                            float vx[V], vd[V], vo[V];
                            unpack<T>(rx[sl], vx);
                            if (BWD) unpack<T>(rd[sl], vd);
#pragma unroll
                            for (int e = 0; e < V; ++e) {
                                if (BWD) {
                                    const float d = (relu && !(vx[e] > 0.f)) ? 0.f : vd[e];
                                    vo[e] = fmaf(ca, d, fmaf(cb, vx[e], cc));
                                } else {
                                    const float y = cb * vx[e];
                                    vo[e] = relu ? fmaxf(y, 0.f) : y;
                                }
                            }
                            stg_stream(po + i, pack<T>(vo));
                        }
                    }
                }
            }
        }
        if (!have_new) break;
        // ==================================================================== NEW moves shared memory -> tensor memory
#pragma unroll
        for (int p = 0; p < P; ++p) {
            if (p < nlive) {
#pragma unroll
                for (int sl = 0; sl < SL; ++sl) {
                    const int i = sl * TH + tid;
                    uint4 vx = make_uint4(0u, 0u, 0u, 0u), vd = vx;
                    if (i < nv) {
                        vx = lds128(sbase + (unsigned)p * pbytes + 16u * i);
                        if (BWD) vd = lds128(sbase + (unsigned)(P + p) * pbytes + 16u * i);
                    }
                    tmem_st4(trow + (unsigned)((p * PL * SL + sl) * 4), vx);
                    if (BWD) tmem_st4(trow + (unsigned)((p * PL * SL + SL + sl) * 4), vd);
                }
            }
        }
        tmem_wait_st();
        fence_proxy_async_smem();                            // the reads above are ordered before the next bulk copies
        have_old = true;
        c_o = c; first_o = first; nlive_o = nlive;
        w0_o = w0; w1_o = w1; ga_o = ga; b_o = pb;
#pragma unroll
        for (int p = 0; p < P; ++p) { u_o[p] = u_n[p]; v_o[p] = v_n[p]; g_o[p] = g_n[p]; mu_o[p] = mu_n[p]; sd_o[p] = sd_n[p]; }
        t_new = take();                                      // (its barriers also order the shared-memory reads of every thread)
        if (t_new < a.items) issue(t_new);
    }
    tmem_free_all(tbase);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
// Returns 0 when launched, > 0 a cuda error, -100 when the path does not apply (selfnorm_flow.cu then serves).
template <bool BWD>
static int launch_tm(FArgs& a, int dtype, float* scratch, cudaStream_t stream) {
    const Knobs& kn = knobs();
    if (!kn.tm || !a.training || a.res) return -100;
    const int N = a.N, C = a.C;
    const int esz = (int)esize(dtype);
    // 16-bit forward: twice the elements per byte through two reduction passes on 512 threads per SM is issue-bound
    // (measured (256,256,56,56) bf16: 0.238 ms against 0.166 ms for the shared-memory kernel); backward gains (0.234 / 0.256)
    if (!BWD && esz != 4 && kn.tm != 3) return -100;
    const size_t pbytes = (size_t)a.M * esz;
    if (pbytes % 16 || N < 2 || C < 1 || a.M < 2) return -100;
    const int nv = (int)(pbytes / 16);
    const int SL = (nv + kTmT - 1) / kTmT;
    if (SL < 4 || SL > 8) return -100;                       // planes of 6..16 KB: 4..8 vectors per thread
    const int P = (BWD ? 16 : 32) / SL;                      // instances whose share fits a thread's 128 TMEM cells
    if (N < 2 * P) return -100;
    const DeviceShape ds = device_shape();
    const size_t dsmem = 128 + (size_t)kTmGroups * P * pbytes * (BWD ? 2 : 1);      // 4 mbarriers | 4 groups of planes
    if (dsmem + 4096 > (size_t)ds.smem_optin) return -100;
    a.nI = (N + P - 1) / P;
    a.D = 0;
    const unsigned long long items = (unsigned long long)C * a.nI;
    if (items > 0x7fffffffull) return -100;
    if (items < (unsigned long long)kn.tm_items * ds.sms) return -100;   // too few items to fill the two-stage pipeline
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
#define CNSN_TM_CASE(SL_)                                                                                \
    case SL_: {                                                                                          \
        auto fn = k_sn_tm<T, BWD, (BWD ? 16 : 32) / SL_, SL_>;                                           \
        e = prepare_kernel(fn, kTmCta, dsmem, &per_sm);                                                  \
        if (e != cudaSuccess) return (int)e;                                                             \
        if (per_sm != 1) return -100;              /* one CTA owns the SM's tensor memory */             \
        const int groups = kTmGroups * ds.sms;                                                           \
        if ((long long)groups < 2ll * a.nI) return -100;                                                 \
        a.pf_dist = kn.pf >= 0 ? kn.pf : groups / 2;                                                     \
        e = cudaMemsetAsync(a.pub, 0xff, fill_bytes, stream);                                            \
        if (e != cudaSuccess) return (int)e;                                                             \
        /* grid: one CTA per SM; every CTA's four groups take tickets until they run out */             \
        e = launch_persistent(fn, a, (a.items + kTmGroups - 1) / kTmGroups, ((unsigned)a.nI + kTmGroups - 1) / kTmGroups, \
                              1, ds.sms, kTmCta, dsmem, stream);                                         \
        if (e == cudaErrorCooperativeLaunchTooLarge) { (void)cudaGetLastError(); return -100; }          \
        if (e != cudaSuccess) return (int)e;                                                             \
    } break;
    CNSN_DISPATCH_DTYPE(dtype, T, switch (SL) {
        CNSN_TM_CASE(4) CNSN_TM_CASE(5) CNSN_TM_CASE(6) CNSN_TM_CASE(7) CNSN_TM_CASE(8)
        default: return -100;
    });
#undef CNSN_TM_CASE
    if (kn.debug)
        fprintf(stderr, "[cnsn flow/tmem] %s P=%d SL=%d nI=%d items=%llu smem=%zu (one CTA of 4 groups per SM)\n", BWD ? "bwd" : "fwd", P, SL,
                a.nI, items, dsmem);
    return launch_status();
}

int selfnorm_tmem_fwd(FArgs& a, int dtype, float* scratch, cudaStream_t stream) { return launch_tm<false>(a, dtype, scratch, stream); }
int selfnorm_tmem_bwd(FArgs& a, int dtype, float* scratch, cudaStream_t stream) { return launch_tm<true>(a, dtype, scratch, stream); }

}  // namespace flow
}  // namespace cnsn
