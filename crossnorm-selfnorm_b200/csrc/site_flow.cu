// site_flow.cu -- a CNSN site whose CrossNorm AND SelfNorm both fire (CNSN.forward, models/cnsn.py:159-164:
// `x = self.crossnorm(x); x = self.selfnorm(x)`) as ONE shared-memory-resident dataflow kernel per direction
// (SURVEY.md 8f-2).  The unfused sequence moves 4*S forward (x -> z, z -> y) and 6*S backward; here the
// CrossNorm output z never leaves the chip: 2*S forward, 3*S backward, and backward keeps no z at all (it is
// rebuilt from x and the O(N*C) CrossNorm statistics).
//
// It is the composition of crossnorm_flow.cu's and selfnorm_flow.cu's resident items (same tickets, same polled
// 8-byte words, same deadlock argument: every wait is for an item of the same channel, a channel's items are
// consecutive tickets and all co-resident, and nothing waits before it has published what its stage owes):
//
//   forward  CTA(ticket t): channel c = t / nI, instances j*I .. j*I+I-1
//     1. cp.async.bulk the planes of x into shared memory                                   (TMA)
//     2. (mu, sd) over the content and the style window; publish the style pair at cn[c][i]; poll cn[c][p(i)]
//     3. (mu_z, sd_z) of z = CrossNorm(x); publish at sn[c][i].  Whole-plane content window (crop 'neither' /
//        'style'): z = ca*x + cb everywhere, so the statistics follow from the content statistics -- no pass, z is
//        not materialised.  Cropped content window: z IN PLACE in shared memory (rounded to T, exactly what the
//        unfused sequence stores), exact two-pass over the plane
//     4. the channel's last ticket folds the N words (BatchNorm1d batch statistics, running statistics) and
//        publishes the channel constants; everyone else polls them
//     5. y = g * z (max(., 0) when the block's ReLU is fused in) out of shared memory, streamed out
//   backward: x and dy resident, TWO passes over shared memory in all
//     1. one reduction pass over (x, dy): z is affine in x per region, so three sums (over the content window:
//        d*(x - mu_c) and d; outside it: d*x) give SelfNorm's sum dy*z AND -- dz = g*dy + b*(z - mu_z) + a being
//        affine in (dy, x) as well -- CrossNorm's S1, S2 in closed form; neither z nor dz is materialised
//     2. CrossNorm's (S1, S2) are affine in the channel constants (k1, k2) SelfNorm's fold is about to produce:
//        their three coefficient words go to cn[c][p(i)][0..2] and sum dy*z to sn[c][i] BEFORE anybody waits
//     3. channel fold (dgamma, dbeta, dw, the two batch-norm-backward scalars) -- the item's only wait
//     4. (k1, k2) known: S1, S2 of this instance and of the instance it is the style source of; dx = affine in
//        (dy, x) per region out of shared memory, streamed out
//
// Training mode only (CrossNorm never fires in eval mode, models/cnsn.py:104).  Shapes outside the resident path
// (planes that are not 16-byte multiples, channels too large for the GPU's shared memory, channel permutation,
// is_two): cnsn_site_supported() says so and the host runs the two operators one after the other.
#include <stdio.h>

#include "site_args.cuh"

namespace cnsn {
namespace flow {

constexpr int kSiteT = 128;             // threads per CTA


// One 16-byte vector of the CrossNorm output: ca*x + cb inside the content window, x outside, rounded to T.
template <typename T>
__device__ __forceinline__ void cn_vec(const float (&vx)[VecOf<T>::n], float (&vz)[VecOf<T>::n], int i, int W,
                                       const Window& cw, bool cfull, float ca, float cb) {
    constexpr int V = VecOf<T>::n;
    if (cfull) {
#pragma unroll
        for (int e = 0; e < V; ++e) vz[e] = fmaf(ca, vx[e], cb);
    } else {
        int h = (i * V) / W, w = i * V - h * W;              // one division per vector
#pragma unroll
        for (int e = 0; e < V; ++e) {
            vz[e] = cw.has(h, w) ? fmaf(ca, vx[e], cb) : vx[e];
            if (++w == W) { w = 0; ++h; }
        }
    }
    if (sizeof(T) < 4) unpack<T>(pack<T>(vz), vz);
}

template <typename T, bool BWD, int TPI>
__device__ __forceinline__ void site_res_item(const SiteArgs& s, const unsigned t, const unsigned par) {
    constexpr int TH = kSiteT;
    constexpr int I = TH / TPI;
    constexpr int V = VecOf<T>::n;
    const FArgs& a = s.sn;
    extern __shared__ __align__(128) unsigned char dsm[];    // [mbarrier | I planes of x | I planes of dy]
    __shared__ float2 s_chan;
    __shared__ float s_f[4][TH / 32];
    uint64_t* bar = reinterpret_cast<uint64_t*>(dsm);
    const unsigned nI = (unsigned)a.nI;
    const unsigned c = t / nI, j = t - c * nI;
    const int N = a.N, C = a.C, H = s.H, W = s.W, M = a.M;
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
        if (threadIdx.x == 0) mbar_arrive_expect_tx(bar, (unsigned)nlive * pbytes * (BWD ? 2u : 1u));
        __syncwarp();
        for (int q = threadIdx.x; q < nlive; q += 32) {
            const size_t off = ((size_t)(first + q) * C + c) * M;
            unsigned char* dst = dsm + 128 + (size_t)q * pbytes;
            tma_load_1d(dst, static_cast<const T*>(a.x) + off, pbytes, bar, pol);
            if (BWD) tma_load_1d(dst + (size_t)I * pbytes, static_cast<const T*>(a.dy) + off, pbytes, bar, pol);
        }
        // L2 prefetch for the CTA that will take this one's place (see selfnorm_flow.cu)
        const unsigned tf = t + (unsigned)a.pf_dist;
        if (a.pf_dist && tf < a.items) {
            const unsigned cf = tf / nI, jf = tf - cf * nI;
            const int ff = (int)jf * I, nf = min(I, N - ff);
            for (int q = threadIdx.x; q < nf; q += 32) {
                const size_t off = ((size_t)(ff + q) * C + cf) * M;
                tma_prefetch_l2(static_cast<const T*>(a.x) + off, pbytes);
                if (BWD) tma_prefetch_l2(static_cast<const T*>(a.dy) + off, pbytes);
            }
        }
    }
    // everything that does not depend on other instances is fetched under the TMA latency
    const Window cw = s.cw, sw = s.sw;
    const bool cfull = cw.full(H, W), sfull = sw.full(H, W);
    const bool same = cw.h0 == sw.h0 && cw.h1 == sw.h1 && cw.w0 == sw.w0 && cw.w1 == sw.w1;
    const float lam = s.lam;
    const bool relu = a.relu != 0;                           // block tail: y = max(y, 0) / dy masked where z <= 0
    const bool folder = j == nI - 1;                         // holds the channel's last ticket
    const int src_n = live ? s.perm[n] : 0;
    const float p_w0 = a.w[2 * c], p_w1 = a.w[2 * c + 1], p_ga = a.gamma[c];
    float p_b, p_rm = 0.f, p_rv = 1.f;
    float pre_g = 0.f, pre_s = 0.f, p_mu = 0.f, p_sd = 1.f;                   // backward: SelfNorm save block
    float muc = 0.f, sdc = 1.f, mus = 0.f, sds = 1.f, mus_src = 0.f, sds_src = 1.f;   // backward: CrossNorm save block
    if (BWD) {
        p_b = a.r[c];
        if (live) {
            pre_g = a.gate[nc]; pre_s = a.shat[nc]; p_mu = a.mu[nc]; p_sd = a.sd[nc];
            muc = s.mu_c[nc]; sdc = s.sd_c[nc]; mus = s.mu_s[nc]; sds = s.sd_s[nc];
            const size_t sc = (size_t)src_n * C + c;
            mus_src = s.mu_s[sc]; sds_src = s.sd_s[sc];
        }
    } else {
        p_b = a.beta[c];
        if (folder && threadIdx.x == 0) { p_rm = a.run_mean[c]; p_rv = a.run_var[c]; }
    }
    mbar_wait(bar, par, a.err);

    float2* flag = a.chan + 4u * c;                          // one 32-byte sector per channel
    uint4* po = reinterpret_cast<uint4*>(static_cast<T*>(a.out) + nc * M);
    if (!BWD) {
        // ---- CrossNorm statistics, pairwise exchange ---------------------------------------------------
        const float2 stc = window_stats<T, TPI>(sx, W, M, cw, cfull, r, live, s.cn_eps, s_f[0], s_f[1]);
        float2 sts = stc;
        if (!same) sts = window_stats<T, TPI>(sx, W, M, sw, sfull, r, live, s.cn_eps, s_f[2], s_f[3]);
        if (live && r == 0) {
            s.mu_c[nc] = stc.x; s.sd_c[nc] = stc.y; s.mu_s[nc] = sts.x; s.sd_s[nc] = sts.y;
            ll_publish(s.pub_cn + (size_t)c * N + n, sts.x, sts.y);
        }
        float ca = 1.f, cb = 0.f;
        if (live) {
            const float2 ps = poll_word(s.pub_cn + (size_t)c * N + src_n, a.poll_ns, a.err);   // the team polls one address
            const float A = ps.y / stc.y;
            ca = lam + (1.f - lam) * A;
            cb = (1.f - lam) * (ps.x - stc.x * A);
        }
        // ---- SelfNorm statistics of z = CrossNorm(x) ----------------------------------------------------------
        float own_x, own_y;
        if (cfull) {
            // whole-plane content window: z = ca*x + cb everywhere, so its statistics follow from the content
            // statistics (mean ca*mu_c + cb, unbiased variance ca^2 * (sd_c^2 - eps_c)): no pass, z is not materialised
            own_x = fmaf(ca, stc.x, cb);
            own_y = sqrtf(fmaf(ca * ca, fmaxf(stc.y * stc.y - s.cn_eps, 0.f), a.eps));
        } else {
            // z in place (thread-private slots, rounded to T as the two-operator sequence stores it), exact two-pass
            float s0 = 0.f, s1 = 0.f;
            if (live) {
#pragma unroll 4
                for (int i = r; i < nv; i += TPI) {
                    float vx[V], vz[V];
                    unpack<T>(lds128(sx + 16u * i), vx);
                    cn_vec<T>(vx, vz, i, W, cw, false, ca, cb);
                    sts128(sx + 16u * i, pack<T>(vz));
#pragma unroll
                    for (int e = 0; e < V; ++e) { if (e & 1) s1 += vz[e]; else s0 += vz[e]; }
                }
            }
            const float mean = team_sum<TPI>(s0 + s1, s_f[0]) * (1.f / M);
            s0 = s1 = 0.f;
            if (live) {
#pragma unroll 4
                for (int i = r; i < nv; i += TPI) {
                    float vz[V];
                    unpack<T>(lds128(sx + 16u * i), vz);
#pragma unroll
                    for (int e = 0; e < V; ++e) { const float d = vz[e] - mean; if (e & 1) s1 = fmaf(d, d, s1); else s0 = fmaf(d, d, s0); }
                }
            }
            const float m2 = team_sum<TPI>(s0 + s1, s_f[1]);
            own_x = mean; own_y = sqrtf(m2 / (M - 1.f) + a.eps);
        }
        if (live && r == 0) {
            a.mu[nc] = own_x; a.sd[nc] = own_y;
            ll_publish(a.pub + (size_t)c * N + n, own_x, own_y);
        }
        // ---- channel constants ---------------------------------------------------------------------------
        if (folder) {
            const float2 cst = fold_publish<false, TH>(a, c, flag, p_w0, p_w1, p_ga, p_b, p_rm, p_rv, s_f);
            if (threadIdx.x == 0) s_chan = cst;
        } else if (threadIdx.x == 0) {
            s_chan = poll_word(flag, a.poll_ns, a.err);
        }
        __syncthreads();
        if (!live) return;
        const float2 cm = s_chan;
        const float sh = (fmaf(p_w0, own_x, p_w1 * own_y) - cm.x) * cm.y;
        const float gt = 1.f / (1.f + expf(-fmaf(p_ga, sh, p_b)));
        if (r == 0) { a.gate[nc] = gt; a.shat[nc] = sh; }
        const float ya = cfull ? gt * ca : gt, yb = cfull ? gt * cb : 0.f;     // cfull: shared memory still holds x
#pragma unroll 4
        for (int i = r; i < nv; i += TPI) {
            float vz[V], vo[V];
            unpack<T>(lds128(sx + 16u * i), vz);
#pragma unroll
            for (int e = 0; e < V; ++e) {
                const float y = fmaf(ya, vz[e], yb);
                vo[e] = relu ? fmaxf(y, 0.f) : y;
            }
            stg_stream(po + i, pack<T>(vo));
        }
    } else {
        // ---- ONE reduction pass over (x, dy) feeds both operators -----------------------------------------
        // z is affine in x per region (ca*x + cb inside the content window, x outside), so with d = dy (masked where
        // z <= 0 when the ReLU is fused in) three sums give everything the two backward passes need:
        //   Ac = sum_cw d*(x - mu_c), Tc = sum_cw d, Po = sum_outside d*x
        //   SelfNorm:  sum d*z = ca*(Ac + mu_c*Tc) + cb*Tc + Po
        //   CrossNorm: dz = g*d + kb*z + kc is affine in (d, x) too, and over the content window sum (x - mu_c) = 0,
        //              sum (x - mu_c)^2 = (Mc - 1)*(sd_c^2 - eps)  =>  sum_cw dz and sum_cw dz*(x - mu_c) in closed
        //              form: dz is never materialised and there is no second reduction pass.
        const float A = sds_src / sdc;
        const float ca = lam + (1.f - lam) * A;
        const float cb = (1.f - lam) * (mus_src - muc * A);
        float a0 = 0.f, a1 = 0.f, t0 = 0.f, t1 = 0.f, o0 = 0.f, o1 = 0.f;
        if (live) {
#pragma unroll 4
            for (int i = r; i < nv; i += TPI) {
                float vx[V], vd[V];
                unpack<T>(lds128(sx + 16u * i), vx);
                unpack<T>(lds128(sdy + 16u * i), vd);
                if (relu) {
                    float vz[V];
                    cn_vec<T>(vx, vz, i, W, cw, cfull, ca, cb);
#pragma unroll
                    for (int e = 0; e < V; ++e) vd[e] = vz[e] > 0.f ? vd[e] : 0.f;
                }
                if (cfull) {
#pragma unroll
                    for (int e = 0; e < V; ++e) {
                        if (e & 1) { a1 = fmaf(vd[e], vx[e] - muc, a1); t1 += vd[e]; } else { a0 = fmaf(vd[e], vx[e] - muc, a0); t0 += vd[e]; }
                    }
                } else {
                    int h = (i * V) / W, w = i * V - h * W;
#pragma unroll
                    for (int e = 0; e < V; ++e) {
                        if (cw.has(h, w)) {
                            if (e & 1) { a1 = fmaf(vd[e], vx[e] - muc, a1); t1 += vd[e]; } else { a0 = fmaf(vd[e], vx[e] - muc, a0); t0 += vd[e]; }
                        } else {
                            if (e & 1) o1 = fmaf(vd[e], vx[e], o1); else o0 = fmaf(vd[e], vx[e], o0);
                        }
                        if (++w == W) { w = 0; ++h; }
                    }
                }
            }
        }
        const float Ac = team_sum<TPI>(a0 + a1, s_f[0]);
        const float Tc = team_sum<TPI>(t0 + t1, s_f[1]);
        const float Po = cfull ? 0.f : team_sum<TPI>(o0 + o1, s_f[2]);      // cfull is uniform over the grid
        const float sxy = fmaf(ca, fmaf(muc, Tc, Ac), fmaf(cb, Tc, Po));
        const float own_x = sxy * pre_g * (1.f - pre_g), own_y = pre_s;
        // SelfNorm's ds = p_b*(own_x*gamma - k1 - own_y*k2) is affine in the channel constants (k1, k2), hence so are
        // dz = pre_g*d + kb*z + kc (kb = ds*KB, kc = ds*KC) and CrossNorm's sums of it over the content window:
        //   sum dz = pre_g*Tc + ds*E1,  sum dz*(x - mu_c) = pre_g*Ac + ds*E2
        // (sum_cw (x - mu_c) = 0, sum_cw (x - mu_c)^2 = (Mc - 1)*(sd_c^2 - eps): dz is never materialised).  So
        // (S1, S2) = w0 + w1*k1 + w2*k2 with three words known NOW: they are handed to the style source before anybody
        // waits, and the only wait of the item is the channel fold -- no second exchange behind it.
        const float Mc = (float)cw.area(), Ms = (float)sw.area();
        const float KB = p_w1 * (1.f / (M - 1.f)) / p_sd;
        const float KC = p_w0 * (1.f / M) - KB * p_mu;
        const float E1 = Mc * fmaf(KB, fmaf(ca, muc, cb), KC);
        const float E2 = KB * ca * (Mc - 1.f) * (sdc * sdc - s.cn_eps);
        const float l1 = 1.f - lam, l2 = (1.f - lam) / sdc;
        const float base1 = pre_g * Tc, base2 = pre_g * Ac;
        if (live && r == 0) {
            const float D0 = p_b * own_x * p_ga, D1 = -p_b, D2 = -p_b * own_y;
            float2* wp = s.pub_cn + ((size_t)c * N + src_n) * 3;
            ll_publish(wp, l1 * fmaf(D0, E1, base1), l2 * fmaf(D0, E2, base2));
            ll_publish(wp + 1, l1 * D1 * E1, l2 * D1 * E2);
            ll_publish(wp + 2, l1 * D2 * E1, l2 * D2 * E2);
            ll_publish(a.pub + (size_t)c * N + n, own_x, own_y);
        }
        if (folder) {
            const float2 cst = fold_publish<true, TH>(a, c, flag, p_w0, p_w1, p_ga, p_b, p_rm, p_rv, s_f);
            if (threadIdx.x == 0) s_chan = cst;
        } else if (threadIdx.x == 0) {
            s_chan = poll_word(flag, a.poll_ns, a.err);
        }
        __syncthreads();
        if (!live) return;
        const float2 cm = s_chan;
        const float dsn = p_b * (own_x * p_ga - cm.x - own_y * cm.y);
        const float kb = dsn * KB, kc = dsn * KC;
        const float zx = kb * ca, zc = fmaf(kb, cb, kc);      // inside the content window: dz = pre_g*d + zx*x + zc
        const float S1 = l1 * fmaf(dsn, E1, base1);
        const float S2 = l2 * fmaf(dsn, E2, base2);
        // this instance as somebody's style source: that instance's three words (published before its own wait)
        const float2* wq = s.pub_cn + ((size_t)c * N + n) * 3;
        const float2 q0 = poll_word(wq, a.poll_ns, a.err), q1 = poll_word(wq + 1, a.poll_ns, a.err), q2 = poll_word(wq + 2, a.poll_ns, a.err);
        const float2 ds = make_float2(fmaf(q2.x, cm.y, fmaf(q1.x, cm.x, q0.x)), fmaf(q2.y, cm.y, fmaf(q1.y, cm.x, q0.y)));
        // CrossNorm backward inside the content window: dx = p*dz + q*x + r0; as somebody's style source: += u*x + v
        const float p = lam + (1.f - lam) * A;
        const float q = -A * S2 / ((Mc - 1.f) * sdc);
        const float r0 = -A * S1 / Mc - q * muc;
        const float u = ds.y / ((Ms - 1.f) * sds);
        const float v = ds.x / Ms - u * mus;
        const float in_d = p * pre_g, in_x = fmaf(p, zx, q), in_c = fmaf(p, zc, r0);   // dx = in_d*d + in_x*x + in_c
        const bool both_full = cfull && sfull;
        const float bx = in_x + u, bc = in_c + v;
#pragma unroll 4
        for (int i = r; i < nv; i += TPI) {
            float vx[V], vd[V], vo[V];
            unpack<T>(lds128(sx + 16u * i), vx);
            unpack<T>(lds128(sdy + 16u * i), vd);
            if (relu) {
                float vz[V];
                cn_vec<T>(vx, vz, i, W, cw, cfull, ca, cb);
#pragma unroll
                for (int e = 0; e < V; ++e) vd[e] = vz[e] > 0.f ? vd[e] : 0.f;
            }
            if (both_full) {
#pragma unroll
                for (int e = 0; e < V; ++e) vo[e] = fmaf(in_d, vd[e], fmaf(bx, vx[e], bc));
            } else {
                int h = (i * V) / W, w = i * V - h * W;
#pragma unroll
                for (int e = 0; e < V; ++e) {
                    float val = cw.has(h, w) ? fmaf(in_d, vd[e], fmaf(in_x, vx[e], in_c)) : fmaf(pre_g, vd[e], fmaf(kb, vx[e], kc));
                    if (sw.has(h, w)) val += fmaf(u, vx[e], v);
                    vo[e] = val;
                    if (++w == W) { w = 0; ++h; }
                }
            }
            stg_stream(po + i, pack<T>(vo));
        }
    }
}

template <typename T, bool BWD, int TPI>
__global__ void __launch_bounds__(kSiteT, 8) k_site_res(const SiteArgs s) {
    CNSN_TICKET_LOOP(s.sn, (site_res_item<T, BWD, TPI>(s, t, it & 1u)))
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct SiteShape { int inst, tpi, nI; size_t dsmem; };

// Item geometry (as the resident SelfNorm / CrossNorm kernels); false when the resident path does not apply.
static bool site_shape(int dtype, int N, int C, int M, bool bwd, SiteShape& g) {
    const int esz = (int)esize(dtype);
    if (((size_t)M * esz) % 16 || N < 2 || C < 1 || M < 2) return false;
    const size_t inst_bytes = (size_t)M * esz * (bwd ? 2 : 1);
    const size_t target = (size_t)knobs().item_kb << 10;
    int inst = 1;
    while (inst < 16 && (size_t)(2 * inst) * inst_bytes <= target + 512 && 2 * inst <= N) inst <<= 1;
    g.inst = inst;
    g.tpi = kSiteT / inst;
    g.dsmem = 128 + (size_t)inst * inst_bytes;
    g.nI = (N + inst - 1) / inst;
    if (g.dsmem > (size_t)device_shape().smem_optin / 2) return false;       // at least two CTAs per SM
    if ((unsigned long long)C * g.nI > 0x7fffffffull) return false;
    return true;
}

// prepare (once per kernel and shared-memory size) and optionally launch; -100 when the shape does not fit
template <bool BWD>
static int launch_site(SiteArgs& s, int dtype, float* scratch, cudaStream_t stream, bool dry_run) {
    FArgs& a = s.sn;
    const int N = a.N, C = a.C;
    SiteShape g;
    if (!site_shape(dtype, N, C, a.M, BWD, g)) return -100;
    const int sms = device_shape().sms;
    a.nI = g.nI;
    const Knobs& kn = knobs();
    a.poll_ns = kn.poll_ns;
    if (!dry_run) a.err = async_error_word();
    a.items = (unsigned)((unsigned long long)C * g.nI);
    // scratch: sn words [C][N] | channel words [C] x 4 (one 32-byte sector each) | ticket | cn words [C][N] (forward)
    // or [C][N][3] (backward); all 0xff
    a.pub = reinterpret_cast<float2*>(scratch);
    a.chan = a.pub + (size_t)N * C;
    a.ticket = reinterpret_cast<unsigned*>(a.chan + 4 * (size_t)C);
    s.pub_cn = a.chan + 4 * (size_t)C + 1;
    const size_t fill_bytes = ((BWD ? 4 : 2) * (size_t)N * C + 4 * (size_t)C + 1) * sizeof(float2);
    cudaError_t e = cudaSuccess;
    int per_sm = 0;
#define CNSN_SITE_CASE(TPI_)                                                                             \
    case TPI_: {                                                                                         \
        auto fn = k_site_res<T, BWD, TPI_>;                                                              \
        e = prepare_kernel(fn, kSiteT, g.dsmem, &per_sm);                                                \
        if (e != cudaSuccess) return (int)e;                                                             \
        if ((long long)per_sm * sms < 2ll * g.nI) return -100;    /* a whole channel must be co-resident */ \
        if (dry_run) return 0;                                                                           \
        a.pf_dist = kn.pf >= 0 ? kn.pf : per_sm * sms / 2;                                               \
        e = cudaMemsetAsync(a.pub, 0xff, fill_bytes, stream);                                            \
        if (e != cudaSuccess) return (int)e;                                                             \
        e = launch_persistent(fn, s, a.items, (unsigned)a.nI, per_sm, sms, kSiteT, g.dsmem, stream);                     \
        if (e == cudaErrorCooperativeLaunchTooLarge) { (void)cudaGetLastError(); return -100; }          \
        if (e != cudaSuccess) return (int)e;                                                             \
    } break;
    CNSN_DISPATCH_DTYPE(dtype, T, switch (g.tpi) {
        CNSN_SITE_CASE(8) CNSN_SITE_CASE(16) CNSN_SITE_CASE(32) CNSN_SITE_CASE(64) CNSN_SITE_CASE(128)
        default: return -100;
    });
#undef CNSN_SITE_CASE
    if (kn.debug)
        fprintf(stderr, "[cnsn flow/site] %s tpi=%d I=%d nI=%d items=%u smem=%zu ctas/sm=%d\n", BWD ? "bwd" : "fwd", g.tpi, g.inst,
                g.nI, a.items, g.dsmem, per_sm);
    return launch_status();
}

// sn words [C][N] | channel words [C] x 4 | ticket | cn words [C][N] x 3 (backward; forward uses one per instance)
static size_t site_scratch_floats(int N, int C) { return 8 * (size_t)N * C + 8 * (size_t)C + 8; }

struct SiteSave {             // offsets (in floats) into the save block
    size_t mu_c, sd_c, mu_s, sd_s, mu, sd, g, shat, r, scratch, total;
    SiteSave(int N, int C) {
        const size_t nc = (size_t)N * C;
        mu_c = 0; sd_c = nc; mu_s = 2 * nc; sd_s = 3 * nc;
        mu = 4 * nc; sd = 5 * nc; g = 6 * nc; shat = 7 * nc; r = 8 * nc;
        scratch = (r + (size_t)C + 1) & ~(size_t)1;           // 8-byte aligned
        total = scratch + site_scratch_floats(N, C);
    }
};

}  // namespace flow
}  // namespace cnsn

using namespace cnsn;
using flow::SiteArgs;
using flow::SiteSave;

extern "C" size_t cnsn_site_save_floats(int N, int C) { return SiteSave(N, C).total; }
extern "C" size_t cnsn_site_workspace_floats(int N, int C) { return flow::site_scratch_floats(N, C); }

static int site_check(const void* x, const void* out, int dtype, int N, int C, int H, int W, const int* perm,
                      const int* content, const int* style, Window& cw, Window& sw) {
    if (!x || !out || !perm || !content || !style || check_dims(N, C, H, W)) return CNSN_E_BADARG;
    if (dtype < CNSN_F32 || dtype > CNSN_F16) return CNSN_E_BADARG;
    cw = Window{content[0], content[1], content[2], content[3]};
    sw = Window{style[0], style[1], style[2], style[3]};
    if (check_window(cw, H, W) || check_window(sw, H, W)) return CNSN_E_BADARG;
    if (N < 2) return CNSN_E_BATCH1;                 // BatchNorm1d raises ValueError in the reference
    if (!aligned16(x) || !aligned16(out)) return CNSN_E_UNSUPPORTED;
    if (async_error_peek()) return CNSN_E_TIMEOUT;
    return 0;
}

extern "C" int cnsn_site_supported(int dtype, int N, int C, int H, int W) {
    if (check_dims(N, C, H, W) || dtype < CNSN_F32 || dtype > CNSN_F16) return 0;
    SiteArgs s{};
    s.sn.N = N; s.sn.C = C; s.sn.M = H * W; s.H = H; s.W = W;
    if (flow::launch_site<false>(s, dtype, nullptr, nullptr, true) != 0) return 0;
    return flow::launch_site<true>(s, dtype, nullptr, nullptr, true) == 0 ? 1 : 0;
}

extern "C" int cnsn_site_fwd(const void* x, void* y, int dtype, int N, int C, int H, int W,
                             const int* perm, const int* content, const int* style, float lam, float cn_eps,
                             const cnsn_gate_params* g, float momentum, float bn_eps, float sn_eps, int relu,
                             float* save, void* stream) {
    Window cw, sw;
    const int rc = site_check(x, y, dtype, N, C, H, W, perm, content, style, cw, sw);
    if (rc) return rc;
    if (!save || !g || !g->w || !g->gamma || !g->beta || !g->run_mean || !g->run_var) return CNSN_E_BADARG;
    const SiteSave L(N, C);
    SiteArgs s{};
    flow::FArgs& a = s.sn;
    a.x = x; a.dy = nullptr; a.out = y; a.N = N; a.C = C; a.M = H * W;
    a.training = 1; a.momentum = momentum; a.bn_eps = bn_eps; a.eps = sn_eps; a.relu = relu ? 1 : 0;
    a.w = g->w; a.gamma = g->gamma; a.beta = g->beta; a.run_mean = g->run_mean; a.run_var = g->run_var; a.nbt = g->nbt;
    a.mu = save + L.mu; a.sd = save + L.sd; a.gate = save + L.g; a.shat = save + L.shat; a.r = save + L.r;
    s.H = H; s.W = W; s.cw = cw; s.sw = sw; s.lam = lam; s.cn_eps = cn_eps; s.perm = perm;
    s.mu_c = save + L.mu_c; s.sd_c = save + L.sd_c; s.mu_s = save + L.mu_s; s.sd_s = save + L.sd_s;
    const int trc = flow::site_tmem_fwd(s, dtype, save + L.scratch, (cudaStream_t)stream);     // whole-plane windows, 6-16 KB planes
    if (trc != -100) return trc;
    const int frc = flow::launch_site<false>(s, dtype, save + L.scratch, (cudaStream_t)stream, false);
    return frc == -100 ? CNSN_E_UNSUPPORTED : frc;
}

extern "C" int cnsn_site_bwd(const void* x, const void* dy, void* dx, int dtype, int N, int C, int H, int W,
                             const int* perm, const int* content, const int* style, float lam, float cn_eps, int relu,
                             const cnsn_gate_params* g, const float* save, const cnsn_gate_grads* dg,
                             float* workspace, void* stream) {
    Window cw, sw;
    const int rc = site_check(x, dx, dtype, N, C, H, W, perm, content, style, cw, sw);
    if (rc) return rc;
    if (!dy || !save || !workspace || !g || !g->w || !g->gamma || !dg || !dg->dw || !dg->dgamma || !dg->dbeta) return CNSN_E_BADARG;
    if (!aligned16(dy)) return CNSN_E_UNSUPPORTED;
    const SiteSave L(N, C);
    float* sv = const_cast<float*>(save);
    SiteArgs s{};
    flow::FArgs& a = s.sn;
    a.x = x; a.dy = dy; a.out = dx; a.N = N; a.C = C; a.M = H * W;
    a.training = 1; a.relu = relu ? 1 : 0;
    a.w = g->w; a.gamma = g->gamma;
    a.mu = sv + L.mu; a.sd = sv + L.sd; a.gate = sv + L.g; a.shat = sv + L.shat; a.r = sv + L.r;
    a.dw = dg->dw; a.dgamma = dg->dgamma; a.dbeta = dg->dbeta;
    s.H = H; s.W = W; s.cw = cw; s.sw = sw; s.lam = lam; s.cn_eps = cn_eps; s.perm = perm;
    s.mu_c = sv + L.mu_c; s.sd_c = sv + L.sd_c; s.mu_s = sv + L.mu_s; s.sd_s = sv + L.sd_s;
    const int trc = flow::site_tmem_bwd(s, dtype, workspace, (cudaStream_t)stream);
    if (trc != -100) return trc;
    const int frc = flow::launch_site<true>(s, dtype, workspace, (cudaStream_t)stream, false);
    return frc == -100 ? CNSN_E_UNSUPPORTED : frc;
}
