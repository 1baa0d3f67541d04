// bn_grp.cu -- BatchNorm2d (+ fused ReLU) for planes that are NOT a multiple of 16 bytes (7x7 fp32 = 196 B, 14x14 bf16 =
// 392 B, 7x7 bf16 = 98 B: the last two stages of ResNet-50, all of them under bf16 autocast), as one shared-memory-
// resident channel-group kernel per direction -- the layout of selfnorm_flow.cu's k_sn_grp with the arithmetic of
// ibn_flow.cu's batch-norm channels.
//
// Why it exists: under autocast torch normalises bf16 activations with its native batch-norm kernels, which take 58 ms of
// a 166 ms ResNet-50 3-view step on B200 (29 layers at 14x14 and 7x7, 0.7-1.0 ms forward + backward each;
// gpurun_out/r3i_jsdprof.log), and the resident kernel of ibn_flow.cu moves whole planes with 16-byte bulk copies.
//
//   kk adjacent channels of one sample are contiguous in NCHW and kk*M*sizeof(T) IS a multiple of 16 for some kk in
//   {2,4,8}: that run (a "super-plane") is what TMA fetches and what the apply phase streams out with 128-bit accesses.
//   item (ticket t) = I samples x kk channels of group g = t / nI:
//     1. cp.async.bulk the super-planes (x [, dy]) into shared memory
//     2. per instance (n, c): forward (mean, M2) -- exact two-pass --, backward (A, B) = (sum d, sum d*xhat) with d = dy
//        masked where the forward output was <= 0 when the ReLU is fused; one polled 8-byte word per instance
//     3. the group's kk channels are folded by its last kk tickets, one channel each: Chan merge of the N (mean, M2) pairs
//        -> batch mean / rstd, running statistics; backward: dgamma, dbeta and the two means the normalisation removes;
//        one polled word per channel
//     4. per-channel coefficients, apply over the item's super-planes as flat 128-bit vectors, streamed out
//   Eval mode: running statistics, no exchange in the forward; the backward still folds (parameter gradients).
//
// Persistent ticket loop under a cooperative launch, bounded waits, no trap: flow_common.cuh.
#include <stdio.h>

#include "flow_common.cuh"

namespace cnsn {
namespace flow {

constexpr int kBnGrpT = 128;

struct BGArgs {
    const void* x; const void* dy; void* out;
    int N, C, M, kk;
    int nI, poll_ns, pf_dist, training, relu;
    unsigned items;
    float eps, momentum;
    const float* gamma; const float* beta;      // [C]
    float* run_mean; float* run_var; long long* nbt;
    float* save_mean; float* save_rstd;         // [C]: written by forward, read by backward
    float* dgamma; float* dbeta;                // backward outputs
    float2* pub;                                // [C][N] polled words
    float2* chan;                               // [C] x 4 (one 32-byte sector per channel)
    unsigned* ticket;
    unsigned* err;
};

// Fold channel ch: poll its N published words, merge, publish the channel word, write the per-channel outputs.
template <bool BWD, int TH>
__device__ __forceinline__ void bn_fold_publish(const BGArgs& a, unsigned ch, float (*s_f)[TH / 32]) {
    constexpr int kHold = 4;
    const int N = a.N, M = a.M;
    const float2* pb = a.pub + (size_t)ch * N;
    float2 hold[kHold];
#pragma unroll
    for (int u = 0; u < kHold; ++u) {
        const int k = threadIdx.x + u * TH;
        hold[u] = make_float2(0.f, 0.f);
        if (k < N) hold[u] = poll_word(pb + k, 100, a.err);
    }
    for (int k = threadIdx.x + kHold * TH; k < N; k += TH) poll_word(pb + k, 100, a.err);
    float v[2] = {0.f, 0.f};
    const float cnt = (float)N * (float)M;
    if (!BWD) {                                              // Chan merge of N equal-sized (mean, M2) pairs
#pragma unroll
        for (int u = 0; u < kHold; ++u) if (threadIdx.x + u * TH < N) v[0] += hold[u].x;
        for (int k = threadIdx.x + kHold * TH; k < N; k += TH) v[0] += ll_peek(pb + k).x;
        cta_sums<1, TH>(*reinterpret_cast<float(*)[1]>(&v[0]), reinterpret_cast<float(*)[TH / 32]>(s_f[0]));
        const float cmean = v[0] / N;
#pragma unroll
        for (int u = 0; u < kHold; ++u)
            if (threadIdx.x + u * TH < N) { const float d = hold[u].x - cmean; v[1] += hold[u].y + M * d * d; }
        for (int k = threadIdx.x + kHold * TH; k < N; k += TH) {
            const float2 p = ll_peek(pb + k);
            const float d = p.x - cmean;
            v[1] += p.y + M * d * d;
        }
        cta_sums<1, TH>(*reinterpret_cast<float(*)[1]>(&v[1]), reinterpret_cast<float(*)[TH / 32]>(s_f[1]));
        const float crstd = 1.f / sqrtf(v[1] / cnt + a.eps);
        if (threadIdx.x == 0) {
            ll_publish(a.chan + 4u * ch, cmean, crstd);
            a.save_mean[ch] = cmean; a.save_rstd[ch] = crstd;
            a.run_mean[ch] = (1.f - a.momentum) * a.run_mean[ch] + a.momentum * cmean;
            a.run_var[ch] = (1.f - a.momentum) * a.run_var[ch] + a.momentum * (v[1] / (cnt - 1.f));
            if (a.nbt && ch == 0) *a.nbt += 1;
        }
    } else {
#pragma unroll
        for (int u = 0; u < kHold; ++u) if (threadIdx.x + u * TH < N) { v[0] += hold[u].x; v[1] += hold[u].y; }
        for (int k = threadIdx.x + kHold * TH; k < N; k += TH) { const float2 p = ll_peek(pb + k); v[0] += p.x; v[1] += p.y; }
        cta_sums<2, TH>(v, s_f);
        if (threadIdx.x == 0) {
            const float inv = a.training ? 1.f / cnt : 0.f;  // eval: running statistics are constants, nothing to remove
            ll_publish(a.chan + 4u * ch, v[0] * inv, v[1] * inv);
            a.dbeta[ch] = v[0]; a.dgamma[ch] = v[1];
        }
    }
}

template <typename T, bool BWD, int TPI>
__device__ __forceinline__ void bn_grp_item(const BGArgs& a, const unsigned t, const unsigned par) {
    constexpr int TH = kBnGrpT, P = kBnGrpT / TPI;
    constexpr int V = VecOf<T>::n;
    extern __shared__ __align__(128) unsigned char dsm[];    // [mbarrier | I super-planes of x | I of dy]
    __shared__ float4 s_coef[P];                             // per instance: out = .x*d + .y*x + .z
    __shared__ float2 s_fwd[P];                              // backward with ReLU: forward map y = .x*x + .y (the mask)
    __shared__ float s_f[2][TH / 32];
    uint64_t* bar = reinterpret_cast<uint64_t*>(dsm);
    const unsigned nI = (unsigned)a.nI;
    const unsigned g = t / nI, j = t - g * nI;
    const int N = a.N, C = a.C, M = a.M, kk = a.kk, I = P / kk;
    const unsigned pbytes = (unsigned)M * (unsigned)sizeof(T), sp = (unsigned)kk * pbytes;
    const int first = (int)j * I, nlive = min(I, N - first);
    const uint32_t sbase = smem_u32(dsm) + 128u;
    const uint32_t soff2 = (unsigned)I * sp;
    if (threadIdx.x < 32) {                                  // lane q fetches sample q's run of kk planes
        const uint64_t pol = l2_policy_evict_first();
        if (threadIdx.x == 0) mbar_arrive_expect_tx(bar, (unsigned)nlive * sp * (BWD ? 2u : 1u));
        __syncwarp();
        for (int q = threadIdx.x; q < nlive; q += 32) {
            const size_t off = ((size_t)(first + q) * C + (size_t)g * kk) * M;
            unsigned char* dst = dsm + 128 + (size_t)q * sp;
            tma_load_1d(dst, static_cast<const T*>(a.x) + off, sp, bar, pol);
            if (BWD) tma_load_1d(dst + soff2, static_cast<const T*>(a.dy) + off, sp, bar, pol);
        }
        const unsigned tf = t + (unsigned)a.pf_dist;
        if (a.pf_dist && tf < a.items) {
            const unsigned gf = tf / nI, jf = tf - gf * nI;
            const int ff = (int)jf * I, nf = min(I, N - ff);
            for (int q = threadIdx.x; q < nf; q += 32) {
                const size_t off = ((size_t)(ff + q) * C + (size_t)gf * kk) * M;
                tma_prefetch_l2(static_cast<const T*>(a.x) + off, sp);
                if (BWD) tma_prefetch_l2(static_cast<const T*>(a.dy) + off, sp);
            }
        }
    }
    // this team's instance
    const int team = threadIdx.x / TPI, r = threadIdx.x % TPI;
    const int q = team / kk, cl = team - q * kk;
    const int n = first + q;
    const bool live = q < nlive;
    const unsigned c = g * (unsigned)kk + (unsigned)cl;
    const uint32_t sx = sbase + (unsigned)q * sp + (unsigned)cl * pbytes;
    const uint32_t sdy = sx + soff2;
    const bool relu = a.relu != 0;
    const bool coupled = a.training != 0;
    const float gam = a.gamma[c];
    const float bet = (BWD && !relu) ? 0.f : a.beta[c];
    float mean = 0.f, rstd = 1.f;
    if (BWD) { mean = a.save_mean[c]; rstd = a.save_rstd[c]; }
    const float fy = rstd * gam, fc = bet - mean * fy;       // backward: the forward map (ReLU mask)
    mbar_wait(bar, par, a.err);
    // ---- per-instance reduction, element-wise out of shared memory ---------------------------------
    float own_x = 0.f, own_y = 0.f;
    if (BWD) {
        float s0 = 0.f, s1 = 0.f;
        if (live)
            for (int e = r; e < M; e += TPI) {
                const float x = lds_elem<T>(sx, e);
                const float d = (relu && !(fmaf(fy, x, fc) > 0.f)) ? 0.f : lds_elem<T>(sdy, e);
                s0 += d;
                s1 = fmaf(d, (x - mean) * rstd, s1);
            }
        own_x = team_sum<TPI>(s0, s_f[0]);
        own_y = team_sum<TPI>(s1, s_f[1]);
    } else if (coupled) {
        float s0 = 0.f;
        if (live) for (int e = r; e < M; e += TPI) s0 += lds_elem<T>(sx, e);
        own_x = team_sum<TPI>(s0, s_f[0]) * (1.f / M);
        s0 = 0.f;
        if (live) for (int e = r; e < M; e += TPI) { const float d = lds_elem<T>(sx, e) - own_x; s0 = fmaf(d, d, s0); }
        own_y = team_sum<TPI>(s0, s_f[1]);
    }
    if (live && r == 0 && (BWD || coupled)) ll_publish(a.pub + (size_t)c * N + n, own_x, own_y);
    // ---- channel constants: channel k of the group is folded by ticket nI-1-(k mod nI) -----------------
    if (BWD || coupled) {
        for (int k = (int)(nI - 1u - j); k < kk; k += (int)nI) bn_fold_publish<BWD, TH>(a, g * (unsigned)kk + (unsigned)k, s_f);
    }
    if (live && r == 0) {
        float2 cm = make_float2(0.f, 0.f);                   // forward (mean, rstd); backward the two means removed
        if (coupled) {
            cm = poll_word(a.chan + 4u * c, a.poll_ns, a.err);
        } else if (!BWD) {                                   // eval-mode forward: running statistics
            cm = make_float2(a.run_mean[c], 1.f / sqrtf(a.run_var[c] + a.eps));
            if (j == nI - 1 && q == 0) { a.save_mean[c] = cm.x; a.save_rstd[c] = cm.y; }
        }
        if (BWD) {
            // dx = gam*rstd*(d - ma - xhat*mb), xhat = (x - mean)*rstd
            const float ca = gam * rstd;
            const float cbx = -ca * cm.y * rstd;
            s_coef[team] = make_float4(ca, cbx, -ca * cm.x - cbx * mean, 0.f);
            s_fwd[team] = make_float2(fy, fc);
        } else {
            const float cbx = cm.y * gam;
            s_coef[team] = make_float4(0.f, cbx, bet - cm.x * cbx, 0.f);
        }
    }
    __syncthreads();
    // ---- apply: the item's super-planes as flat 128-bit vectors, division-free walk ------------------------
    const int vps = (int)(sp / 16u), nvec = nlive * vps;
    const unsigned magicM = 0xffffffffu / (unsigned)M + 1u;  // floor(e / M) = umulhi(e, magicM) for e * M < 2^32
    const size_t srow = (size_t)C * M;
    T* ob = static_cast<T*>(a.out) + ((size_t)first * C + (size_t)g * kk) * M;
    int qv = (int)threadIdx.x / vps, w = (int)threadIdx.x - qv * vps;
    const int vdq = TH / vps, vdw = TH - vdq * vps;
    for (int vi = threadIdx.x; vi < nvec; vi += TH) {
        const int e0 = w * V, cl0 = (int)__umulhi((unsigned)e0, magicM), bound = (cl0 + 1) * M;
        const int i0 = qv * kk + cl0, i1 = (cl0 + 1 < kk) ? i0 + 1 : i0;     // a vector spans at most two planes (M >= V)
        const float4 k0 = s_coef[i0], k1 = s_coef[i1];
        float2 m0 = make_float2(0.f, 1.f), m1 = m0;
        if (BWD && relu) { m0 = s_fwd[i0]; m1 = s_fwd[i1]; }
        float vx[V], vd[V], vo[V];
        unpack<T>(lds128(sbase + 16u * vi), vx);
        if (BWD) unpack<T>(lds128(sbase + soff2 + 16u * vi), vd);
#pragma unroll
        for (int e = 0; e < V; ++e) {
            const bool lo = e0 + e < bound;
            const float4 k = lo ? k0 : k1;
            if (BWD) {
                const float2 m = lo ? m0 : m1;
                const float d = (relu && !(fmaf(m.x, vx[e], m.y) > 0.f)) ? 0.f : vd[e];
                vo[e] = fmaf(k.x, d, fmaf(k.y, vx[e], k.z));
            } else {
                const float y = fmaf(k.y, vx[e], k.z);
                vo[e] = relu ? fmaxf(y, 0.f) : y;
            }
        }
        stg_stream(reinterpret_cast<uint4*>(ob + (size_t)qv * srow) + w, pack<T>(vo));
        qv += vdq; w += vdw;
        if (w >= vps) { w -= vps; ++qv; }
    }
}

template <typename T, bool BWD, int TPI>
__global__ void __launch_bounds__(kBnGrpT) k_bn_grp(const BGArgs a) {
    CNSN_TICKET_LOOP(a, (bn_grp_item<T, BWD, TPI>(a, t, it & 1u)))
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
// 0 launched (or, dry_run, would launch); > 0 a cuda error; -100 the path does not apply.
template <bool BWD>
static int launch_bn_grp(BGArgs& a, int dtype, float* scratch, cudaStream_t stream, bool dry_run) {
    const int N = a.N, C = a.C;
    const int esz = (int)esize(dtype);
    const size_t pb = (size_t)a.M * esz;
    if (pb % 16 == 0 || N < 1 || C < 1 || a.M < 16 / esz * 2) return -100;
    if ((long long)N * a.M < 2) return -100;
    int kk = 0;
    for (int k = 2; k <= 8; k <<= 1) if ((k * pb) % 16 == 0 && C % k == 0) { kk = k; break; }
    if (!kk) return -100;
    const size_t ib = pb * (BWD ? 2 : 1);                    // bytes per instance in shared memory
    const Knobs& kn = knobs();
    const size_t want = (size_t)kn.grp_kb << 10;
    const int tpi = (32 * ib >= want) ? 4 : (64 * ib >= want) ? 2 : 1;
    const int I = (kBnGrpT / tpi) / kk;
    if (I < 1) return -100;
    const size_t dsmem = 128 + (size_t)I * kk * ib;
    const DeviceShape ds = device_shape();
    if (dsmem > (size_t)ds.smem_optin / 2) return -100;
    a.kk = kk;
    a.nI = (N + I - 1) / I;
    a.poll_ns = kn.poll_ns;
    const unsigned long long items = (unsigned long long)(C / kk) * a.nI;
    if (items > 0x7fffffffull) return -100;
    a.items = (unsigned)items;
    a.pub = reinterpret_cast<float2*>(scratch);              // [C][N] words | [C] x 4 channel words | ticket; all 0xff
    a.chan = a.pub + (size_t)N * C;
    a.ticket = reinterpret_cast<unsigned*>(a.chan + 4 * (size_t)C);
    const size_t fill_bytes = ((size_t)N * C + 4 * (size_t)C + 1) * sizeof(float2);
    cudaError_t e = cudaSuccess;
    int per_sm = 0;
    CNSN_DISPATCH_DTYPE(dtype, T, {
        auto fn = tpi == 4 ? k_bn_grp<T, BWD, 4> : tpi == 2 ? k_bn_grp<T, BWD, 2> : k_bn_grp<T, BWD, 1>;
        e = prepare_kernel(fn, kBnGrpT, dsmem, &per_sm);
        if (e != cudaSuccess) return (int)e;
        if ((long long)per_sm * ds.sms < (long long)a.nI) return -100;      // a group's items wait for each other
        if (dry_run) return 0;
        if (async_error_peek()) return CNSN_E_TIMEOUT;
        a.err = async_error_word();
        a.pf_dist = kn.pf >= 0 ? kn.pf : per_sm * ds.sms / 2;
        e = cudaMemsetAsync(a.pub, 0xff, fill_bytes, stream);
        if (e != cudaSuccess) return (int)e;
        e = launch_persistent(fn, a, a.items, (unsigned)a.nI, per_sm, ds.sms, kBnGrpT, dsmem, stream);
        if (e == cudaErrorCooperativeLaunchTooLarge) { (void)cudaGetLastError(); return -100; }
        if (e != cudaSuccess) return (int)e;
    });
    if (kn.debug)
        fprintf(stderr, "[cnsn flow/bn-grp] %s kk=%d tpi=%d I=%d nI=%d items=%llu smem=%zu ctas/sm=%d\n", BWD ? "bwd" : "fwd", kk, tpi, I,
                a.nI, items, dsmem, per_sm);
    return launch_status();
}

// Entry points for ibn_flow.cu (half == 0 only).  save: [bn_mean C | bn_rstd C]; scratch as above.
int bn_grp_fwd(const void* x, void* y, int dtype, int N, int C, int M, const float* gamma, const float* beta, float* run_mean,
               float* run_var, long long* nbt, int training, int relu, float momentum, float eps, float* save_mean,
               float* save_rstd, float* scratch, cudaStream_t stream, bool dry_run) {
    BGArgs a{};
    a.x = x; a.dy = nullptr; a.out = y; a.N = N; a.C = C; a.M = M;
    a.training = training; a.relu = relu ? 1 : 0; a.momentum = momentum; a.eps = eps;
    a.gamma = gamma; a.beta = beta; a.run_mean = run_mean; a.run_var = run_var; a.nbt = nbt;
    a.save_mean = save_mean; a.save_rstd = save_rstd;
    return launch_bn_grp<false>(a, dtype, scratch, stream, dry_run);
}
int bn_grp_bwd(const void* x, const void* dy, void* dx, int dtype, int N, int C, int M, const float* gamma, const float* beta,
               int training, int relu, float* save_mean, float* save_rstd, float* dgamma, float* dbeta, float* scratch,
               cudaStream_t stream, bool dry_run) {
    BGArgs a{};
    a.x = x; a.dy = dy; a.out = dx; a.N = N; a.C = C; a.M = M;
    a.training = training; a.relu = relu ? 1 : 0;
    a.gamma = gamma; a.beta = beta;
    a.save_mean = save_mean; a.save_rstd = save_rstd; a.dgamma = dgamma; a.dbeta = dbeta;
    return launch_bn_grp<true>(a, dtype, scratch, stream, dry_run);
}

}  // namespace flow
}  // namespace cnsn
