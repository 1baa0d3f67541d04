// crossnorm_flow.cu -- 2-instance CrossNorm (cn_op_2ins_space_chan, models/cnsn.py:58-91, with
// instance_norm_mix :20-29) forward / backward as ONE shared-memory-resident dataflow kernel per direction.
//
// Instance (i,c) needs the style statistics of instance (p(i),c) -- and, backward, the sums of the instance
// whose style source it is.  The two-kernel path (crossnorm.cu) therefore reads x (and dy) twice.  Here, as
// in selfnorm_flow.cu's resident kernel, an item keeps its planes in shared memory between the reduction and
// the apply, and the exchange is a polled 8-byte word per instance ("data is the flag"):
//
//   CTA(ticket t): channel c = t / nI (channel-major: partners share the channel), instances j*I .. j*I+I-1
//     1. cp.async.bulk the planes (x [, dy]) into shared memory                      (TMA)
//     2. forward : exact two-pass (mean, std) over the content window and over the style window, out of shared
//                  memory; publish (mu_s, sd_s) at [c][i]
//        backward: S1 = sum_Wc d, S2 = sum_Wc d*xhat (d = (1-lam)*dy); publish (S1, S2) at [c][p(i)]  (the
//                  permutation is a bijection: every slot is written exactly once, no atomics)
//     3. poll the ONE word this instance needs: forward [c][p(i)], backward [c][i]
//     4. apply out of shared memory (piecewise affine per window), stream the result out
//
// HBM and L2 traffic: 2*S forward, 3*S backward.  No channel-wide barrier: an instance waits for one partner.
// Deadlock freedom: as selfnorm_flow.cu (tickets in increasing order; a channel's items are consecutive
// tickets; the GPU holds at least 2*nI CTAs, checked with the occupancy API) -- both partners are resident
// whenever their channel is the lowest unfinished one, and nothing waits before it has published.
// Not handled here (-> crossnorm.cu): channel permutation (`chan=True`, no caller of the reference enables it),
// planes that are not a multiple of 16 bytes, channels too large for the GPU's shared memory.
#include <stdio.h>

#include "flow_common.cuh"

namespace cnsn {
namespace flow {

constexpr int kCnT = 128;               // threads per CTA

struct CNArgs {
    const void* x; const void* dy; void* out;      // forward: dy == nullptr, out = y; backward: out = dx
    int N, C, H, W;
    int nI;                 // items per channel = ceil(N / I)
    int poll_ns;
    int pf_dist;            // L2-prefetch the item pf_dist tickets ahead (0 = off)
    unsigned items;         // total tickets
    Window cw, sw;          // content / style window
    float lam, eps;
    const int* perm;        // [N] style source of every sample
    float* mu_c; float* sd_c; float* mu_s; float* sd_s;     // save block: written by forward, read by backward
    float2* pub;            // [C][N] polled words, pre-filled with the sentinel
    unsigned* ticket;       // starts at 0xffffffff
    unsigned* err;          // asynchronous error word
};

template <typename T, bool BWD, int TPI>
__device__ __forceinline__ void cn_res_item(const CNArgs& a, const unsigned t, const unsigned par) {
    constexpr int TH = kCnT;
    constexpr int I = TH / TPI;
    constexpr int V = VecOf<T>::n;
    extern __shared__ __align__(128) unsigned char dsm[];    // [mbarrier | I planes of x | I planes of dy]
    __shared__ float s_f[4][TH / 32];
    uint64_t* bar = reinterpret_cast<uint64_t*>(dsm);
    const unsigned nI = (unsigned)a.nI;
    const unsigned c = t / nI, j = t - c * nI;
    const int N = a.N, C = a.C, H = a.H, W = a.W, M = H * W;
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
    const Window cw = a.cw, sw = a.sw;
    const bool cfull = cw.full(H, W), sfull = sw.full(H, W);
    const bool same = cw.h0 == sw.h0 && cw.h1 == sw.h1 && cw.w0 == sw.w0 && cw.w1 == sw.w1;
    const float lam = a.lam;
    // fetched under the TMA latency
    const int src_n = live ? a.perm[n] : 0;
    float muc = 0.f, sdc = 1.f, mus = 0.f, sds = 1.f, sds_src = 1.f;
    if (BWD && live) {
        muc = a.mu_c[nc]; sdc = a.sd_c[nc]; mus = a.mu_s[nc]; sds = a.sd_s[nc];
        sds_src = a.sd_s[(size_t)src_n * C + c];
    }
    mbar_wait(bar, par, a.err);

    uint4* po = reinterpret_cast<uint4*>(static_cast<T*>(a.out) + nc * M);
    if (!BWD) {
        // ---- statistics of both windows, publish the style pair, wait for the partner's ----------------
        const float2 stc = window_stats<T, TPI>(sx, W, M, cw, cfull, r, live, a.eps, s_f[0], s_f[1]);
        float2 sts = stc;
        if (!same) sts = window_stats<T, TPI>(sx, W, M, sw, sfull, r, live, a.eps, s_f[2], s_f[3]);
        if (live && r == 0) {
            a.mu_c[nc] = stc.x; a.sd_c[nc] = stc.y; a.mu_s[nc] = sts.x; a.sd_s[nc] = sts.y;
            ll_publish(a.pub + (size_t)c * N + n, sts.x, sts.y);
        }
        if (!live) return;
        const float2 ps = poll_word(a.pub + (size_t)c * N + src_n, a.poll_ns, a.err);   // the team polls one address
        const float A = ps.y / stc.y;
        const float ca = lam + (1.f - lam) * A;
        const float cb = (1.f - lam) * (ps.x - stc.x * A);
#pragma unroll 4
        for (int i = r; i < nv; i += TPI) {
            float vx[V], vo[V];
            unpack<T>(lds128(sx + 16u * i), vx);
            if (cfull) {
#pragma unroll
                for (int e = 0; e < V; ++e) vo[e] = fmaf(ca, vx[e], cb);
            } else {
                int h = (i * V) / W, w = i * V - h * W;      // one division per vector
#pragma unroll
                for (int e = 0; e < V; ++e) {
                    vo[e] = cw.has(h, w) ? fmaf(ca, vx[e], cb) : vx[e];
                    if (++w == W) { w = 0; ++h; }
                }
            }
            stg_stream(po + i, pack<T>(vo));
        }
    } else {
        // ---- S1, S2 over the content window, scattered to the style source; wait for this instance's own ----
        float t0 = 0.f, t1 = 0.f, a0 = 0.f, a1 = 0.f;
        if (live) window_accumulate<T, true>(sx, sdy, W, M, cw, cfull, r, TPI, [&](float x, float d, int e) {
            if (e & 1) { a1 = fmaf(d, x - muc, a1); t1 += d; } else { a0 = fmaf(d, x - muc, a0); t0 += d; }
        });
        const float asum = team_sum<TPI>(a0 + a1, s_f[0]);
        const float tsum = team_sum<TPI>(t0 + t1, s_f[1]);
        const float S1 = (1.f - lam) * tsum;
        const float S2 = (1.f - lam) * asum / sdc;
        if (live && r == 0) ll_publish(a.pub + (size_t)c * N + src_n, S1, S2);
        if (!live) return;
        const float2 ds = poll_word(a.pub + (size_t)c * N + n, a.poll_ns, a.err);        // (dmu_s, dsd_s) of this instance
        const float Mc = (float)cw.area(), Ms = (float)sw.area();
        const float A = sds_src / sdc;
        // inside the content window: dx = p*dy + q*x + r0
        const float p = lam + (1.f - lam) * A;
        const float q = -A * S2 / ((Mc - 1.f) * sdc);
        const float r0 = -A * S1 / Mc - q * muc;
        // inside the style window (this instance as somebody's style source): dx += u*x + v
        const float u = ds.y / ((Ms - 1.f) * sds);
        const float v = ds.x / Ms - u * mus;
        const bool both_full = cfull && sfull;
        const float qq = q + u, rr = r0 + v;
#pragma unroll 4
        for (int i = r; i < nv; i += TPI) {
            float vx[V], vd[V], vo[V];
            unpack<T>(lds128(sx + 16u * i), vx);
            unpack<T>(lds128(sdy + 16u * i), vd);
            if (both_full) {
#pragma unroll
                for (int e = 0; e < V; ++e) vo[e] = fmaf(p, vd[e], fmaf(qq, vx[e], rr));
            } else {
                int h = (i * V) / W, w = i * V - h * W;
#pragma unroll
                for (int e = 0; e < V; ++e) {
                    float val = cw.has(h, w) ? fmaf(p, vd[e], fmaf(q, vx[e], r0)) : vd[e];
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
__global__ void __launch_bounds__(kCnT) k_cn_res(const CNArgs a) {
    CNSN_TICKET_LOOP(a, (cn_res_item<T, BWD, TPI>(a, t, it & 1u)))
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
template <bool BWD>
static int launch_cn(CNArgs& a, int dtype, float* scratch, cudaStream_t stream) {
    const int N = a.N, C = a.C, M = a.H * a.W;
    const int esz = (int)esize(dtype);
    if (((size_t)M * esz) % 16 || N < 1 || C < 1) return -100;
    const size_t inst_bytes = (size_t)M * esz * (BWD ? 2 : 1);
    const Knobs& kn = knobs();
    const size_t target = (size_t)kn.item_kb << 10;
    int inst = 1;
    while (inst < 16 && (size_t)(2 * inst) * inst_bytes <= target + 512 && 2 * inst <= N) inst <<= 1;
    const int tpi = kCnT / inst;
    const size_t dsmem = 128 + (size_t)inst * inst_bytes;
    const DeviceShape ds = device_shape();
    const int sms = ds.sms;
    if (dsmem > (size_t)ds.smem_optin / 2) return -100;      // at least two CTAs per SM
    a.nI = (N + inst - 1) / inst;
    a.poll_ns = kn.poll_ns;
    a.err = async_error_word();
    const unsigned long long items = (unsigned long long)C * a.nI;
    if (items > 0x7fffffffull) return -100;
    a.items = (unsigned)items;
    a.pub = reinterpret_cast<float2*>(scratch);              // [C][N] float2 | ticket
    a.ticket = reinterpret_cast<unsigned*>(a.pub + (size_t)N * C);
    const size_t fill_bytes = ((size_t)N * C + 1) * sizeof(float2);
    cudaError_t e = cudaSuccess;
    int per_sm = 0;
#define CNSN_CN_CASE(TPI_)                                                                               \
    case TPI_: {                                                                                         \
        auto fn = k_cn_res<T, BWD, TPI_>;                                                                \
        e = prepare_kernel(fn, kCnT, dsmem, &per_sm);                                                    \
        if (e != cudaSuccess) return (int)e;                                                             \
        if ((long long)per_sm * sms < 2ll * a.nI) return -100;    /* a whole channel must be co-resident */ \
        a.pf_dist = kn.pf >= 0 ? kn.pf : per_sm * sms / 2;                                               \
        e = cudaMemsetAsync(a.pub, 0xff, fill_bytes, stream);                                            \
        if (e != cudaSuccess) return (int)e;                                                             \
        e = launch_persistent(fn, a, a.items, (unsigned)a.nI, per_sm, sms, kCnT, dsmem, stream);                         \
        if (e == cudaErrorCooperativeLaunchTooLarge) { (void)cudaGetLastError(); return -100; }          \
        if (e != cudaSuccess) return (int)e;                                                             \
    } break;
    CNSN_DISPATCH_DTYPE(dtype, T, switch (tpi) {
        CNSN_CN_CASE(8) CNSN_CN_CASE(16) CNSN_CN_CASE(32) CNSN_CN_CASE(64) CNSN_CN_CASE(128)
        default: return -100;
    });
#undef CNSN_CN_CASE
    if (kn.debug)
        fprintf(stderr, "[cnsn flow/cn] %s tpi=%d I=%d nI=%d items=%llu smem=%zu ctas/sm=%d\n", BWD ? "bwd" : "fwd", tpi, inst, a.nI,
                items, dsmem, per_sm);
    return launch_status();
}

size_t crossnorm_scratch_floats(int N, int C) { return 2 * (size_t)N * C + 8; }

// Both return 0 when launched, >0 cuda error, -100 when the path does not apply (-> crossnorm.cu).
int crossnorm_flow_fwd(const void* x, void* y, int dtype, int N, int C, int H, int W, const int* perm,
                       const Window& cw, const Window& sw, float lam, float eps,
                       float* mu_c, float* sd_c, float* mu_s, float* sd_s, float* scratch, cudaStream_t stream) {
    if (!aligned16(x) || !aligned16(y)) return -100;
    if (async_error_peek()) return CNSN_E_TIMEOUT;
    CNArgs a{};
    a.x = x; a.dy = nullptr; a.out = y; a.N = N; a.C = C; a.H = H; a.W = W;
    a.cw = cw; a.sw = sw; a.lam = lam; a.eps = eps; a.perm = perm;
    a.mu_c = mu_c; a.sd_c = sd_c; a.mu_s = mu_s; a.sd_s = sd_s;
    return launch_cn<false>(a, dtype, scratch, stream);
}

int crossnorm_flow_bwd(const void* x, const void* dy, void* dx, int dtype, int N, int C, int H, int W, const int* perm,
                       const Window& cw, const Window& sw, float lam,
                       const float* mu_c, const float* sd_c, const float* mu_s, const float* sd_s,
                       float* scratch, cudaStream_t stream) {
    if (!aligned16(x) || !aligned16(dy) || !aligned16(dx)) return -100;
    if (async_error_peek()) return CNSN_E_TIMEOUT;
    CNArgs a{};
    a.x = x; a.dy = dy; a.out = dx; a.N = N; a.C = C; a.H = H; a.W = W;
    a.cw = cw; a.sw = sw; a.lam = lam; a.eps = 0.f; a.perm = perm;
    a.mu_c = const_cast<float*>(mu_c); a.sd_c = const_cast<float*>(sd_c);
    a.mu_s = const_cast<float*>(mu_s); a.sd_s = const_cast<float*>(sd_s);
    return launch_cn<true>(a, dtype, scratch, stream);
}

}  // namespace flow
}  // namespace cnsn
