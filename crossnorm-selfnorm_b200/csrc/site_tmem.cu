// site_tmem.cu -- the fused CrossNorm -> SelfNorm site (CNSN.forward with both operators firing, models/cnsn.py:159-164)
// as the shared memory + tensor memory pipeline of selfnorm_tmem.cu, for WHOLE-PLANE content and style windows
// (crop = 'neither': z = ca*x + cb on the whole plane, so every statistic of z follows analytically and neither z nor dz
// is ever materialised).  Cropped windows, planes outside 6..16 KB and small tensors stay with site_flow.cu.
//
// The SelfNorm pipeline (one CTA of four 128-thread groups per SM; the NEW item is fetched into shared memory, reduced
// and published at once; the OLD item waits for its channel in tensor memory and is applied from there) gains one
// exchange level per direction:
//
//   forward  NEW: (mu, sd) of x over the plane (CrossNorm's eps) -> level-1 word cn[c][n] -> poll the style source's
//                 level-1 word cn[c][p(n)] -> (ca, cb) -> statistics of z in closed form -> level-2 word sn[c][n] (SelfNorm)
//            OLD: channel constants -> gate -> y = g*ca*x + g*cb out of tensor memory
//   backward NEW: ONE reduction pass over (x, dy): sum d*(x - mu_c), sum d -> SelfNorm's sum dy*z; CrossNorm's (S1, S2)
//                 are affine in the channel constants SelfNorm's fold is about to produce: their three coefficient
//                 words go to the style source's slots cn[c][p(n)][0..2], sum dy*z to sn[c][n] -- nothing waited for
//            OLD: channel constants -> own (S1, S2) and, from three words that are long there, those of the instance
//                 this one is the style source of -> dx = affine in (dy, x) out of tensor memory
//
// Deadlock freedom: a level-1 word is published as soon as the item's own bulk copy has landed and been reduced; a
// level-2 word waits for level-1 words only; channel folds wait for level-2 words only.  Every taken ticket therefore
// reaches its level-1 publish, hence every level-2 publish happens, hence every fold -- as long as all groups are
// co-resident (cooperative launch) and there are at least as many groups as items in a channel (checked on the host),
// so that a group waiting for a partner's level-1 word is never waiting for a ticket nobody can take.
#include <stdio.h>

#include "site_args.cuh"
#include "tmem_common.cuh"

namespace cnsn {
namespace flow {

template <typename T, bool BWD, int P, int SL>
__global__ void __launch_bounds__(kTmCta, 1) k_site_tm(const SiteArgs s) {
    constexpr int TH = kTmT, V = VecOf<T>::n;
    constexpr int PL = BWD ? 2 : 1;
    using S = Group128Sync;
    static_assert(P * PL * SL * 4 <= kTmCols, "an item must fit the group's TMEM slice");
    const FArgs& a = s.sn;
    extern __shared__ __align__(128) unsigned char dsm_all[];
    __shared__ unsigned s_tickets[kTmGroups], s_tmem;
    __shared__ float2 s_chans[kTmGroups];
    __shared__ float s_fs[kTmGroups][2][TH / 32];
    __shared__ float s_reds[kTmGroups][2 * P][TH / 32];
    __shared__ float4 s_xch[kTmGroups][P];                   // per plane, thread p -> the group
    const int grp = threadIdx.x >> 7;
    const int tid = threadIdx.x & 127, warp = tid >> 5;
    const unsigned pbytes = (unsigned)a.M * (unsigned)sizeof(T);
    unsigned char* dsm = dsm_all + (size_t)grp * P * PL * pbytes;          // dsm + 128 = this group's planes
    uint64_t* bar = reinterpret_cast<uint64_t*>(dsm_all) + grp;
    unsigned& s_ticket = s_tickets[grp];
    float2& s_chan = s_chans[grp];
    float (*s_f)[TH / 32] = s_fs[grp];
    float (*s_red)[TH / 32] = s_reds[grp];
    float4* xch = s_xch[grp];
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const uint32_t tbase = tmem_alloc_all(&s_tmem);
    const uint32_t trow = tbase + ((uint32_t)(warp & 3) << 21) + (uint32_t)(grp * kTmCols);

    const int N = a.N, C = a.C, M = a.M, nv = M / V;
    const unsigned nI = (unsigned)a.nI;
    const uint32_t sbase = smem_u32(dsm) + 128u;
    const bool relu = a.relu != 0;
    const float invM = 1.f / M, invM1 = 1.f / (M - 1.f), Mf = (float)M;
    const float lam = s.lam, l1 = 1.f - lam;

    auto take = [&]() -> unsigned {
        if (tid == 0) s_ticket = atomicAdd(a.ticket, 1u) + 1u;
        S::sync();
        const unsigned t = s_ticket;
        S::sync();
        return t;
    };
    auto issue = [&](unsigned t) {
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
    // the rounded CrossNorm output decides the ReLU mask, exactly as the two-operator sequence stores it
    auto zpos = [&](float x, float ca, float cb) -> bool {
        float z = fmaf(ca, x, cb);
        if (sizeof(T) < 4) z = to_f(from_f<T>(z));
        return z > 0.f;
    };

    // the OLD item (in tensor memory)
    bool have_old = false;
    unsigned c_o = 0;
    int first_o = 0, nlive_o = 0;
    float w0_o = 0.f, w1_o = 0.f, ga_o = 0.f, b_o = 0.f;     // b: forward beta, backward r
    float ca_o[P], cb_o[P], u_o[P], v_o[P];                  // CrossNorm map; forward (mu_z, sd_z), backward published (dz, shat)
    float g_o[P], KB_o[P], KC_o[P], E1_o[P], E2_o[P], b1_o[P], b2_o[P], A_o[P], muc_o[P], sdc_o[P];   // backward only
#pragma unroll
    for (int p = 0; p < P; ++p) {
        ca_o[p] = 1.f; cb_o[p] = 0.f; u_o[p] = 0.f; v_o[p] = 0.f; g_o[p] = 0.f; KB_o[p] = 0.f; KC_o[p] = 0.f; E1_o[p] = 0.f;
        E2_o[p] = 0.f; b1_o[p] = 0.f; b2_o[p] = 0.f; A_o[p] = 1.f; muc_o[p] = 0.f; sdc_o[p] = 1.f;
    }
    unsigned par = 0;

    unsigned t_new = take();
    if (t_new < a.items) issue(t_new);
    for (;;) {
        const bool have_new = t_new < a.items;
        unsigned c = 0;
        int first = 0, nlive = 0;
        float w0 = 0.f, w1 = 0.f, ga = 0.f, pb = 0.f;
        float ca_n[P], cb_n[P], u_n[P], v_n[P], g_n[P], KB_n[P], KC_n[P], E1_n[P], E2_n[P], b1_n[P], b2_n[P], A_n[P], muc_n[P], sdc_n[P];
#pragma unroll
        for (int p = 0; p < P; ++p) {
            ca_n[p] = 1.f; cb_n[p] = 0.f; u_n[p] = 0.f; v_n[p] = 0.f; g_n[p] = 0.f; KB_n[p] = 0.f; KC_n[p] = 0.f; E1_n[p] = 0.f;
            E2_n[p] = 0.f; b1_n[p] = 0.f; b2_n[p] = 0.f; A_n[p] = 1.f; muc_n[p] = 0.f; sdc_n[p] = 1.f;
        }
        if (have_new) {
            // ================================================================ NEW
            c = t_new / nI;
            const unsigned j = t_new - c * nI;
            first = (int)j * P;
            nlive = min(P, N - first);
            const bool folder = j == nI - 1;
            w0 = a.w[2 * c]; w1 = a.w[2 * c + 1]; ga = a.gamma[c];
            float p_rm = 0.f, p_rv = 1.f;
            float shat_n[P], pmu_n[P], psd_n[P];
            if (BWD) {
                pb = a.r[c];
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    shat_n[p] = 0.f; pmu_n[p] = 0.f; psd_n[p] = 1.f;
                    if (p < nlive) {
                        const int n = first + p;
                        const size_t nc = (size_t)n * C + c, sc = (size_t)s.perm[n] * C + c;
                        g_n[p] = a.gate[nc]; shat_n[p] = a.shat[nc]; pmu_n[p] = a.mu[nc]; psd_n[p] = a.sd[nc];
                        muc_n[p] = s.mu_c[nc]; sdc_n[p] = s.sd_c[nc];
                        A_n[p] = s.sd_s[sc] / sdc_n[p];
                        ca_n[p] = lam + l1 * A_n[p];
                        cb_n[p] = l1 * (s.mu_s[sc] - muc_n[p] * A_n[p]);
                    }
                }
            } else {
                pb = a.beta[c];
                if (folder && tid == 0) { p_rm = a.run_mean[c]; p_rv = a.run_var[c]; }
            }
            mbar_wait(bar, par, a.err);
            par ^= 1u;
            float sm[2 * P];
#pragma unroll
            for (int p = 0; p < P; ++p) {
                float s0 = 0.f, s1 = 0.f, t0 = 0.f, t1 = 0.f;
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
                                    const float d = (relu && !zpos(vx[e], ca_n[p], cb_n[p])) ? 0.f : vd[e];
                                    if (e & 1) { s1 = fmaf(d, vx[e] - muc_n[p], s1); t1 += d; } else { s0 = fmaf(d, vx[e] - muc_n[p], s0); t0 += d; }
                                }
                            } else {
#pragma unroll
                                for (int e = 0; e < V; ++e) { if (e & 1) s1 += vx[e]; else s0 += vx[e]; }
                            }
                        }
                    }
                }
                sm[p] = s0 + s1;
                sm[P + p] = t0 + t1;
            }
            cta_sums<2 * P, TH, S>(sm, s_red);
            if (BWD) {
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    if (p < nlive) {
                        const float Ac = sm[p], Tc = sm[P + p];
                        const float sxy = fmaf(ca_n[p], fmaf(muc_n[p], Tc, Ac), cb_n[p] * Tc);
                        u_n[p] = sxy * g_n[p] * (1.f - g_n[p]);
                        v_n[p] = shat_n[p];
                        KB_n[p] = w1 * invM1 / psd_n[p];
                        KC_n[p] = w0 * invM - KB_n[p] * pmu_n[p];
                        E1_n[p] = Mf * fmaf(KB_n[p], fmaf(ca_n[p], muc_n[p], cb_n[p]), KC_n[p]);
                        E2_n[p] = KB_n[p] * ca_n[p] * (Mf - 1.f) * (sdc_n[p] * sdc_n[p] - s.cn_eps);
                        b1_n[p] = g_n[p] * Tc;
                        b2_n[p] = g_n[p] * Ac;
                        if (tid == p) {                          // three coefficient words for the style source, then SelfNorm's word
                            const int n = first + p;
                            const float l2 = l1 / sdc_n[p];
                            const float D0 = pb * u_n[p] * ga, D1 = -pb, D2 = -pb * v_n[p];
                            float2* wp = s.pub_cn + ((size_t)c * N + s.perm[n]) * 3;
                            ll_publish(wp, l1 * fmaf(D0, E1_n[p], b1_n[p]), l2 * fmaf(D0, E2_n[p], b2_n[p]));
                            ll_publish(wp + 1, l1 * D1 * E1_n[p], l2 * D1 * E2_n[p]);
                            ll_publish(wp + 2, l1 * D2 * E1_n[p], l2 * D2 * E2_n[p]);
                            ll_publish(a.pub + (size_t)c * N + n, u_n[p], v_n[p]);
                        }
                    }
                }
            } else {
                float mu[P];
#pragma unroll
                for (int p = 0; p < P; ++p) mu[p] = sm[p] * invM;
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
                                    const float d = vx[e] - mu[p];
                                    if (e & 1) s1 = fmaf(d, d, s1); else s0 = fmaf(d, d, s0);
                                }
                            }
                        }
                    }
                    sm[p] = s0 + s1;
                    sm[P + p] = 0.f;
                }
                cta_sums<2 * P, TH, S>(sm, s_red);
                // level 1: every plane's own pair goes out BEFORE any thread polls (the polling threads share a warp: a
                // thread spinning for a word that a later branch of its own warp would publish must not exist)
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    if (tid == p && p < nlive) {
                        const int n = first + p;
                        const size_t nc = (size_t)n * C + c;
                        const float sd = sqrtf(sm[p] * invM1 + s.cn_eps);
                        s.mu_c[nc] = mu[p]; s.sd_c[nc] = sd; s.mu_s[nc] = mu[p]; s.sd_s[nc] = sd;
                        ll_publish(s.pub_cn + (size_t)c * N + n, mu[p], sd);
                        xch[p] = make_float4(mu[p], sd, 0.f, 0.f);
                    }
                }
                // level 2: the style source's pair -> (ca, cb) -> the statistics of z in closed form -> SelfNorm's word
                if (tid < nlive) {
                    const int n = first + tid;
                    const size_t nc = (size_t)n * C + c;
                    const float4 own = xch[tid];
                    const float2 ps = poll_word(s.pub_cn + (size_t)c * N + s.perm[n], a.poll_ns, a.err);
                    const float A = ps.y / own.y;
                    const float ca = lam + l1 * A, cb = l1 * (ps.x - own.x * A);
                    const float mz = fmaf(ca, own.x, cb);
                    const float sz = sqrtf(fmaf(ca * ca, fmaxf(own.y * own.y - s.cn_eps, 0.f), a.eps));
                    a.mu[nc] = mz; a.sd[nc] = sz;
                    ll_publish(a.pub + (size_t)c * N + n, mz, sz);
                    xch[tid] = make_float4(ca, cb, mz, sz);
                }
                S::sync();
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    if (p < nlive) { const float4 q = xch[p]; ca_n[p] = q.x; cb_n[p] = q.y; u_n[p] = q.z; v_n[p] = q.w; }
                }
            }
            if (folder) fold_publish<BWD, TH, S>(a, c, a.chan + 4u * c, w0, w1, ga, pb, p_rm, p_rv, s_f);
        }
        if (have_old) {
            // ================================================================ OLD
            if (tid == 0) s_chan = poll_word(a.chan + 4u * c_o, a.poll_ns, a.err);
            if (BWD) {
                // this instance as somebody's style source: that instance's three words (published before ITS wait)
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    if (tid == 32 + p && p < nlive_o) {
                        const float2* wq = s.pub_cn + ((size_t)c_o * N + first_o + p) * 3;
                        const float2 q0 = poll_word(wq, a.poll_ns, a.err), q1 = poll_word(wq + 1, a.poll_ns, a.err), q2 = poll_word(wq + 2, a.poll_ns, a.err);
                        xch[p] = make_float4(q0.x, q0.y, q1.x, q1.y);
                        s_red[p][0] = q2.x; s_red[p][1] = q2.y;
                    }
                }
            }
            S::sync();
            const float2 cm = s_chan;
#pragma unroll
            for (int p = 0; p < P; ++p) {
                if (p < nlive_o) {
                    const size_t nc = (size_t)(first_o + p) * C + c_o;
                    float kd = 0.f, kx, kc;                  // out = kd*d + kx*x + kc
                    if (BWD) {
                        const float4 q01 = xch[p];
                        const float q2x = s_red[p][0], q2y = s_red[p][1];
                        const float dsn = b_o * (u_o[p] * ga_o - cm.x - v_o[p] * cm.y);
                        const float kb = dsn * KB_o[p], kcc = dsn * KC_o[p];
                        const float zx = kb * ca_o[p], zc = fmaf(kb, cb_o[p], kcc);
                        const float l2 = l1 / sdc_o[p];
                        const float S1 = l1 * fmaf(dsn, E1_o[p], b1_o[p]);
                        const float S2 = l2 * fmaf(dsn, E2_o[p], b2_o[p]);
                        const float dsx = fmaf(q2x, cm.y, fmaf(q01.z, cm.x, q01.x)), dsy = fmaf(q2y, cm.y, fmaf(q01.w, cm.x, q01.y));
                        const float pp = lam + l1 * A_o[p];
                        const float qq = -A_o[p] * S2 / ((Mf - 1.f) * sdc_o[p]);
                        const float r0 = -A_o[p] * S1 / Mf - qq * muc_o[p];
                        const float uu = dsy / ((Mf - 1.f) * sdc_o[p]);      // whole-plane style window: (mu_s, sd_s) = (mu_c, sd_c)
                        const float vv = dsx / Mf - uu * muc_o[p];
                        kd = pp * g_o[p];
                        kx = fmaf(pp, zx, qq) + uu;
                        kc = fmaf(pp, zc, r0) + vv;
                    } else {
                        const float sh = (fmaf(w0_o, u_o[p], w1_o * v_o[p]) - cm.x) * cm.y;
                        const float gt = 1.f / (1.f + expf(-fmaf(ga_o, sh, b_o)));
                        if (tid == p) { a.gate[nc] = gt; a.shat[nc] = sh; }
                        kx = gt * ca_o[p];
                        kc = gt * cb_o[p];
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
                                    const float d = (relu && !zpos(vx[e], ca_o[p], cb_o[p])) ? 0.f : vd[e];
                                    vo[e] = fmaf(kd, d, fmaf(kx, vx[e], kc));
                                } else {
                                    const float y = fmaf(kx, vx[e], kc);
                                    vo[e] = relu ? fmaxf(y, 0.f) : y;
                                }
                            }
                            stg_stream(po + i, pack<T>(vo));
                        }
                    }
                }
            }
            S::sync();                                       // xch / s_red are free again
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
        fence_proxy_async_smem();
        have_old = true;
        c_o = c; first_o = first; nlive_o = nlive;
        w0_o = w0; w1_o = w1; ga_o = ga; b_o = pb;
#pragma unroll
        for (int p = 0; p < P; ++p) {
            ca_o[p] = ca_n[p]; cb_o[p] = cb_n[p]; u_o[p] = u_n[p]; v_o[p] = v_n[p]; g_o[p] = g_n[p]; KB_o[p] = KB_n[p]; KC_o[p] = KC_n[p];
            E1_o[p] = E1_n[p]; E2_o[p] = E2_n[p]; b1_o[p] = b1_n[p]; b2_o[p] = b2_n[p]; A_o[p] = A_n[p]; muc_o[p] = muc_n[p]; sdc_o[p] = sdc_n[p];
        }
        t_new = take();
        if (t_new < a.items) issue(t_new);
    }
    tmem_free_all(tbase);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
template <bool BWD>
static int launch_site_tm(SiteArgs& s, int dtype, float* scratch, cudaStream_t stream) {
    FArgs& a = s.sn;
    const Knobs& kn = knobs();
    if (!kn.tm) return -100;
    const int N = a.N, C = a.C, H = s.H, W = s.W;
    if (!s.cw.full(H, W) || !s.sw.full(H, W)) return -100;   // whole-plane windows only (crop = 'neither')
    const int esz = (int)esize(dtype);
    if (esz != 4 && kn.tm != 3) return -100;                 // as selfnorm_tmem.cu: the 16-bit passes are issue-bound on 512 threads
    const size_t pbytes = (size_t)a.M * esz;
    if (pbytes % 16 || N < 2 || a.M < 2) return -100;
    const int nv = (int)(pbytes / 16);
    const int SL = (nv + kTmT - 1) / kTmT;
    if (SL < 4 || SL > 8) return -100;
    const int P = (BWD ? 16 : 32) / SL;
    if (N < 2 * P) return -100;
    const DeviceShape ds = device_shape();
    const size_t dsmem = 128 + (size_t)kTmGroups * P * pbytes * (BWD ? 2 : 1);
    if (dsmem + 6144 > (size_t)ds.smem_optin) return -100;
    a.nI = (N + P - 1) / P;
    a.D = 0;
    const unsigned long long items = (unsigned long long)C * a.nI;
    if (items > 0x7fffffffull) return -100;
    if (items < (unsigned long long)kn.tm_items * ds.sms) return -100;
    // scratch exactly as site_flow.cu: sn words [C][N] | channel words [C] x 4 | ticket | cn words [C][N] (x 3 backward)
    a.pub = reinterpret_cast<float2*>(scratch);
    a.chan = a.pub + (size_t)N * C;
    a.ticket = reinterpret_cast<unsigned*>(a.chan + 4 * (size_t)C);
    s.pub_cn = a.chan + 4 * (size_t)C + 1;
    a.done = nullptr; a.ready = nullptr; a.trace = nullptr;
    a.poll_ns = kn.poll_ns;
    a.items = (unsigned)items;
    a.err = async_error_word();
    const size_t fill_bytes = ((BWD ? 4 : 2) * (size_t)N * C + 4 * (size_t)C + 1) * sizeof(float2);
    cudaError_t e = cudaSuccess;
    int per_sm = 0;
#define CNSN_SITE_TM_CASE(SL_)                                                                           \
    case SL_: {                                                                                          \
        auto fn = k_site_tm<T, BWD, (BWD ? 16 : 32) / SL_, SL_>;                                         \
        e = prepare_kernel(fn, kTmCta, dsmem, &per_sm);                                                  \
        if (e != cudaSuccess) return (int)e;                                                             \
        if (per_sm != 1) return -100;                                                                    \
        const int groups = kTmGroups * ds.sms;                                                           \
        if ((long long)groups < 2ll * a.nI) return -100;                                                 \
        a.pf_dist = kn.pf >= 0 ? kn.pf : groups / 2;                                                     \
        e = cudaMemsetAsync(a.pub, 0xff, fill_bytes, stream);                                            \
        if (e != cudaSuccess) return (int)e;                                                             \
        e = launch_persistent(fn, s, (a.items + kTmGroups - 1) / kTmGroups, ((unsigned)a.nI + kTmGroups - 1) / kTmGroups, \
                              1, ds.sms, kTmCta, dsmem, stream);                                         \
        if (e == cudaErrorCooperativeLaunchTooLarge) { (void)cudaGetLastError(); return -100; }          \
        if (e != cudaSuccess) return (int)e;                                                             \
    } break;
    CNSN_DISPATCH_DTYPE(dtype, T, switch (SL) {
        CNSN_SITE_TM_CASE(4) CNSN_SITE_TM_CASE(5) CNSN_SITE_TM_CASE(6) CNSN_SITE_TM_CASE(7) CNSN_SITE_TM_CASE(8)
        default: return -100;
    });
#undef CNSN_SITE_TM_CASE
    if (kn.debug)
        fprintf(stderr, "[cnsn flow/site-tmem] %s P=%d SL=%d nI=%d items=%llu smem=%zu\n", BWD ? "bwd" : "fwd", P, SL, a.nI, items, dsmem);
    return launch_status();
}

int site_tmem_fwd(SiteArgs& s, int dtype, float* scratch, cudaStream_t stream) { return launch_site_tm<false>(s, dtype, scratch, stream); }
int site_tmem_bwd(SiteArgs& s, int dtype, float* scratch, cudaStream_t stream) { return launch_site_tm<true>(s, dtype, scratch, stream); }

}  // namespace flow
}  // namespace cnsn
