// ibn_flow.cu -- Instance-Batch Normalization (the IBN layer of models/imagenet/resnet_ibn_cnsn.py:24-44, the
// host of the reference's best ImageNet model) forward / backward as ONE shared-memory-resident dataflow kernel
// per direction.
//
// The reference splits x along channels, makes both halves contiguous (two copies), runs nn.InstanceNorm2d
// (affine) on the first `half` channels and nn.BatchNorm2d on the rest, and concatenates (a third copy).  Here
// every channel is handled in place by the same kernel, built like selfnorm_flow.cu's resident kernel:
//
//   CTA(ticket t): channel c = t / nI, instances j*I .. j*I+I-1
//     1. cp.async.bulk the planes (x [, dy]) into shared memory                           (TMA)
//     2. per-instance reduction out of shared memory: forward exact two-pass (mean, M2); backward A = sum dy,
//        B = sum dy * xhat
//     3. c <  half (instance norm): the instance's own statistics are all it needs -- no waiting at all;
//        c >= half (batch norm, training): publish the pair as one polled 8-byte word; the CTA holding the
//        channel's last ticket merges the N pairs (Chan) into the channel constants and publishes them; the
//        others poll that word.  Eval-mode batch norm uses the running statistics: no waiting either.
//     4. apply out of shared memory, 128-bit streaming stores.
//   The channel's last ticket also sums the per-instance pairs into dgamma / dbeta (both kinds).
//
// HBM and L2 traffic: 2*S forward, 3*S backward, no split / cat copies.  Deadlock freedom as selfnorm_flow.cu.
// Planes must be multiples of 16 bytes (every IBN site of ResNet-50-IBN-a in fp32: 56x56, 28x28, 14x14).
#include <stdio.h>

#include "flow_common.cuh"
#include "ibn_general.cuh"

namespace cnsn {
namespace flow {

constexpr int kIbnT = 128;

struct IArgs {
    const void* x; const void* dy; void* out;
    int N, C, M, half;
    int nI, poll_ns, pf_dist, training;
    int relu;                                   // y = max(y, 0) (forward) / dy masked where y <= 0 (backward, y rebuilt from x)
    unsigned items;
    float eps_in, eps_bn, momentum;
    const float* in_w; const float* in_b;       // [half]
    const float* bn_w; const float* bn_b;       // [C - half]
    float* run_mean; float* run_var; long long* nbt;
    float* in_mean; float* in_rstd;             // save: [N][half]
    float* bn_mean; float* bn_rstd;             // save: [C - half]
    float* d_in_w; float* d_in_b; float* d_bn_w; float* d_bn_b;   // backward outputs
    float2* pub;                                // [C][N] polled words
    float2* chan;                               // [C] x 4 (one sector per channel)
    unsigned* ticket;
    unsigned* err;                              // asynchronous error word
};

template <typename T, bool BWD, int TPI>
__device__ __forceinline__ void ibn_res_item(const IArgs& a, const unsigned t, const unsigned par) {
    constexpr int TH = kIbnT, I = TH / TPI, V = VecOf<T>::n, kHold = 4;
    extern __shared__ __align__(128) unsigned char dsm[];
    __shared__ float2 s_chan;
    __shared__ float s_f[2][TH / 32];
    uint64_t* bar = reinterpret_cast<uint64_t*>(dsm);
    const unsigned nI = (unsigned)a.nI;
    const unsigned c = t / nI, j = t - c * nI;
    const int N = a.N, C = a.C, M = a.M, half = a.half;
    const int n = (int)j * I + (int)(threadIdx.x / TPI);
    const int r = threadIdx.x % TPI;
    const bool live = n < N;
    const size_t nc = (size_t)(live ? n : 0) * C + c;
    const int nv = M / V;
    const unsigned pbytes = (unsigned)M * (unsigned)sizeof(T);
    const uint32_t sx = smem_u32(dsm) + 128u + (threadIdx.x / TPI) * pbytes;
    const uint32_t sdy = sx + (unsigned)I * pbytes;
    if (threadIdx.x < 32) {
        const int first = (int)j * I, nlive = min(I, N - first);
        const uint64_t pol = l2_policy_evict_first();
        if (threadIdx.x == 0) mbar_arrive_expect_tx(bar, (unsigned)nlive * pbytes * (BWD ? 2u : 1u));
        __syncwarp();
        for (int q = threadIdx.x; q < nlive; q += 32) {
            const size_t off = ((size_t)(first + q) * C + c) * M;
            unsigned char* dst = dsm + 128 + (size_t)q * pbytes;
            tma_load_1d(dst, static_cast<const T*>(a.x) + off, pbytes, bar, pol);
            if (BWD) tma_load_1d(dst + (size_t)I * pbytes, static_cast<const T*>(a.dy) + off, pbytes, bar, pol);
        }
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
    const bool is_in = (int)c < half;                        // CTA-uniform: instance norm / batch norm channel
    const int cb = (int)c - half;                            // batch-norm channel index
    const bool folder = j == nI - 1;
    const bool coupled = !is_in && a.training != 0;          // only training-mode batch norm waits for the channel
    const float gam = is_in ? a.in_w[c] : a.bn_w[cb];
    const bool relu = a.relu != 0;
    const float bet = (BWD && !relu) ? 0.f : (is_in ? a.in_b[c] : a.bn_b[cb]);
    // backward: the statistics the forward saved
    float mean = 0.f, rstd = 1.f;
    if (BWD) {
        if (is_in) { if (live) { mean = a.in_mean[(size_t)n * half + c]; rstd = a.in_rstd[(size_t)n * half + c]; } }
        else { mean = a.bn_mean[cb]; rstd = a.bn_rstd[cb]; }
    }
    // backward with a fused ReLU: the forward output y = fy*x + fc is rebuilt from x (same fp32 expressions as the forward:
    // same bits), dy passes where y > 0
    const float fy = rstd * gam, fc = bet - mean * fy;
    mbar_wait(bar, par, a.err);

    // ---- per-instance reduction ------------------------------------------------------------------
    float own_x = 0.f, own_y = 0.f;                          // forward (mean, M2); backward (A, B)
    {
        float s0 = 0.f, s1 = 0.f;
        if (!BWD) {
            if (live) {
#pragma unroll 4
                for (int i = r; i < nv; i += TPI) {
                    float vx[V];
                    unpack<T>(lds128(sx + 16u * i), vx);
#pragma unroll
                    for (int e = 0; e < V; ++e) { if (e & 1) s1 += vx[e]; else s0 += vx[e]; }
                }
            }
            own_x = team_sum<TPI>(s0 + s1, s_f[0]) * (1.f / M);
            s0 = s1 = 0.f;
            if (live) {
#pragma unroll 4
                for (int i = r; i < nv; i += TPI) {
                    float vx[V];
                    unpack<T>(lds128(sx + 16u * i), vx);
#pragma unroll
                    for (int e = 0; e < V; ++e) { const float d = vx[e] - own_x; if (e & 1) s1 = fmaf(d, d, s1); else s0 = fmaf(d, d, s0); }
                }
            }
            own_y = team_sum<TPI>(s0 + s1, s_f[1]);
        } else {
            if (live) {
#pragma unroll 4
                for (int i = r; i < nv; i += TPI) {
                    float vx[V], vd[V];
                    unpack<T>(lds128(sx + 16u * i), vx);
                    unpack<T>(lds128(sdy + 16u * i), vd);
#pragma unroll
                    for (int e = 0; e < V; ++e) {
                        const float d = (relu && !(fmaf(fy, vx[e], fc) > 0.f)) ? 0.f : vd[e];
                        s0 += d; s1 = fmaf(d, (vx[e] - mean) * rstd, s1);
                    }
                }
            }
            own_x = team_sum<TPI>(s0, s_f[0]);
            own_y = team_sum<TPI>(s1, s_f[1]);
        }
    }
    // published when somebody will read it: the channel merge (training batch norm), the parameter gradients
    if (live && r == 0 && (coupled || BWD)) ll_publish(a.pub + (size_t)c * N + n, own_x, own_y);

    // ---- channel constants -------------------------------------------------------------------------
    float2* flag = a.chan + 4u * c;
    if (folder && (coupled || BWD)) {
        const float2* pb = a.pub + (size_t)c * N;
        float2 hold[kHold];
#pragma unroll
        for (int u = 0; u < kHold; ++u) {
            const int k = threadIdx.x + u * TH;
            hold[u] = make_float2(0.f, 0.f);
            if (k < N) hold[u] = poll_word(pb + k, 100, a.err);
        }
        for (int k = threadIdx.x + kHold * TH; k < N; k += TH) poll_word(pb + k, 100, a.err);
        float v[2] = {0.f, 0.f};
        if (!BWD) {                                          // Chan merge of N equal-sized (mean, M2) pairs
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
            const float cnt = (float)N * (float)M;
            const float var_b = v[1] / cnt;
            const float crstd = 1.f / sqrtf(var_b + a.eps_bn);
            if (threadIdx.x == 0) {
                ll_publish(flag, cmean, crstd);
                s_chan = make_float2(cmean, crstd);
                a.bn_mean[cb] = cmean; a.bn_rstd[cb] = crstd;
                a.run_mean[cb] = (1.f - a.momentum) * a.run_mean[cb] + a.momentum * cmean;
                a.run_var[cb] = (1.f - a.momentum) * a.run_var[cb] + a.momentum * (v[1] / (cnt - 1.f));
                if (a.nbt && cb == 0) *a.nbt += 1;
            }
        } else {                                             // channel sums of (A, B): dbeta, dgamma, and for batch norm the means
#pragma unroll
            for (int u = 0; u < kHold; ++u) if (threadIdx.x + u * TH < N) { v[0] += hold[u].x; v[1] += hold[u].y; }
            for (int k = threadIdx.x + kHold * TH; k < N; k += TH) { const float2 p = ll_peek(pb + k); v[0] += p.x; v[1] += p.y; }
            cta_sums<2, TH>(v, s_f);
            if (threadIdx.x == 0) {
                if (coupled) {
                    const float inv = 1.f / ((float)N * (float)M);
                    ll_publish(flag, v[0] * inv, v[1] * inv);
                    s_chan = make_float2(v[0] * inv, v[1] * inv);
                }
                if (is_in) { a.d_in_b[c] = v[0]; a.d_in_w[c] = v[1]; }
                else { a.d_bn_b[cb] = v[0]; a.d_bn_w[cb] = v[1]; }
            }
        }
    } else if (coupled && threadIdx.x == 0) {
        s_chan = poll_word(flag, a.poll_ns, a.err);
    }
    if (coupled) __syncthreads();                            // CTA-uniform
    if (!live) return;

    // ---- coefficients: out = ca*dy + cb*x + cc ------------------------------------------------------
    float ca = 0.f, cbx = 0.f, cc = 0.f;
    if (!BWD) {
        float m, rs;
        if (is_in) {
            m = own_x; rs = 1.f / sqrtf(own_y * (1.f / M) + a.eps_in);     // biased variance (InstanceNorm2d)
            if (r == 0) { a.in_mean[(size_t)n * half + c] = m; a.in_rstd[(size_t)n * half + c] = rs; }
        } else if (coupled) {
            m = s_chan.x; rs = s_chan.y;
        } else {                                             // eval-mode batch norm: running statistics
            m = a.run_mean[cb]; rs = 1.f / sqrtf(a.run_var[cb] + a.eps_bn);
            if (folder && threadIdx.x == 0) { a.bn_mean[cb] = m; a.bn_rstd[cb] = rs; }
        }
        cbx = rs * gam; cc = bet - m * cbx;
    } else {
        float ma, mb;                                        // mean of dy and of dy*xhat that the normalisation removes
        if (is_in) { ma = own_x * (1.f / M); mb = own_y * (1.f / M); }
        else if (coupled) { ma = s_chan.x; mb = s_chan.y; }
        else { ma = 0.f; mb = 0.f; }
        // dx = gam*rstd*(dy - ma - xhat*mb), xhat = (x - mean)*rstd
        ca = gam * rstd;
        cbx = -ca * mb * rstd;
        cc = -ca * ma - cbx * mean;
    }
    uint4* po = reinterpret_cast<uint4*>(static_cast<T*>(a.out) + nc * M);
#pragma unroll 4
    for (int i = r; i < nv; i += TPI) {
        float vx[V], vd[V], vo[V];
        unpack<T>(lds128(sx + 16u * i), vx);
        if (BWD) unpack<T>(lds128(sdy + 16u * i), vd);
#pragma unroll
        for (int e = 0; e < V; ++e) {
            if (BWD) {
                const float d = (relu && !(fmaf(fy, vx[e], fc) > 0.f)) ? 0.f : vd[e];
                vo[e] = fmaf(ca, d, fmaf(cbx, vx[e], cc));
            } else {
                const float y = fmaf(cbx, vx[e], cc);
                vo[e] = relu ? fmaxf(y, 0.f) : y;
            }
        }
        stg_stream(po + i, pack<T>(vo));
    }
}

template <typename T, bool BWD, int TPI>
__global__ void __launch_bounds__(kIbnT) k_ibn_res(const IArgs a) {
    CNSN_TICKET_LOOP(a, (ibn_res_item<T, BWD, TPI>(a, t, it & 1u)))
}

template <bool BWD>
static int launch_ibn(IArgs& a, int dtype, float* scratch, cudaStream_t stream, bool dry_run = false) {
    const int N = a.N, C = a.C;
    const int esz = (int)esize(dtype);
    if (((size_t)a.M * esz) % 16) return CNSN_E_UNSUPPORTED;      // no general path for this operator (documented)
    const size_t inst_bytes = (size_t)a.M * esz * (BWD ? 2 : 1);
    int inst = 1;
    while (inst < 16 && (size_t)(2 * inst) * inst_bytes <= (25u << 10) + 512 && 2 * inst <= N) inst <<= 1;
    const int tpi = kIbnT / inst;
    const size_t dsmem = 128 + (size_t)inst * inst_bytes;
    const DeviceShape ds = device_shape();
    if (dsmem > (size_t)ds.smem_optin / 2) return CNSN_E_UNSUPPORTED;      // at least two CTAs per SM
    a.nI = (N + inst - 1) / inst;
    const Knobs& kn = knobs();
    if (!dry_run) {
        if (async_error_peek()) return CNSN_E_TIMEOUT;
        a.err = async_error_word();
    }
    a.poll_ns = kn.poll_ns;
    const unsigned long long items = (unsigned long long)C * a.nI;
    if (items > 0x7fffffffull) return CNSN_E_UNSUPPORTED;
    a.items = (unsigned)items;
    a.pub = reinterpret_cast<float2*>(scratch);
    a.chan = a.pub + (size_t)N * C;
    a.ticket = reinterpret_cast<unsigned*>(a.chan + 4 * (size_t)C);
    const size_t fill_bytes = ((size_t)N * C + 4 * (size_t)C + 1) * sizeof(float2);
    cudaError_t e = cudaSuccess;
    int per_sm = 0;
#define CNSN_IBN_CASE(TPI_)                                                                              \
    case TPI_: {                                                                                         \
        auto fn = k_ibn_res<T, BWD, TPI_>;                                                               \
        e = prepare_kernel(fn, kIbnT, dsmem, &per_sm);                                                   \
        if (e != cudaSuccess) return (int)e;                                                             \
        /* a channel must be co-resident when its items wait for each other (training-mode batch-norm channels); */ \
        /* instance-norm channels never wait (their folder only collects words of earlier tickets) */          \
        /* (the persistent cooperative grid makes ONE co-resident channel sufficient; large planes -- the 112x112 stem */ \
        /* of ResNet-50, 2 CTAs per SM backward -- then run a channel at a time, still far ahead of the alternative)   */ \
        if (a.half < C && a.training && (long long)per_sm * ds.sms < (long long)a.nI) return CNSN_E_UNSUPPORTED;       \
        if (dry_run) return 0;                                                                           \
        a.pf_dist = kn.pf >= 0 ? kn.pf : per_sm * ds.sms / 2;                                            \
        e = cudaMemsetAsync(a.pub, 0xff, fill_bytes, stream);                                            \
        if (e != cudaSuccess) return (int)e;                                                             \
        e = launch_persistent(fn, a, a.items, (unsigned)a.nI, per_sm, ds.sms, kIbnT, dsmem, stream);                     \
        if (e == cudaErrorCooperativeLaunchTooLarge) { (void)cudaGetLastError(); return CNSN_E_UNSUPPORTED; }\
        if (e != cudaSuccess) return (int)e;                                                             \
    } break;
    CNSN_DISPATCH_DTYPE(dtype, T, switch (tpi) {
        CNSN_IBN_CASE(8) CNSN_IBN_CASE(16) CNSN_IBN_CASE(32) CNSN_IBN_CASE(64) CNSN_IBN_CASE(128)
        default: return CNSN_E_UNSUPPORTED;
    });
#undef CNSN_IBN_CASE
    return launch_status();
}

// bn_grp.cu: batch norm (half == 0) over planes that are not multiples of 16 bytes, as channel groups
int bn_grp_fwd(const void* x, void* y, int dtype, int N, int C, int M, const float* gamma, const float* beta, float* run_mean,
               float* run_var, long long* nbt, int training, int relu, float momentum, float eps, float* save_mean,
               float* save_rstd, float* scratch, cudaStream_t stream, bool dry_run);
int bn_grp_bwd(const void* x, const void* dy, void* dx, int dtype, int N, int C, int M, const float* gamma, const float* beta,
               int training, int relu, float* save_mean, float* save_rstd, float* dgamma, float* dbeta, float* scratch,
               cudaStream_t stream, bool dry_run);

}  // namespace flow
}  // namespace cnsn

using namespace cnsn;

static bool odd_planes(int dtype, int M) { return ((size_t)M * esize(dtype)) % 16 != 0; }
static int from_grp(int rc) { return rc == -100 ? CNSN_E_UNSUPPORTED : rc; }

static ibn_general::GArgs general_args(const flow::IArgs& a) {
    ibn_general::GArgs g{};
    g.x = a.x; g.dy = a.dy; g.out = a.out; g.N = a.N; g.C = a.C; g.M = a.M; g.half = a.half; g.training = a.training;
    g.momentum = a.momentum; g.eps_in = a.eps_in; g.eps_bn = a.eps_bn;
    g.in_w = a.in_w; g.in_b = a.in_b; g.bn_w = a.bn_w; g.bn_b = a.bn_b;
    g.run_mean = a.run_mean; g.run_var = a.run_var; g.nbt = a.nbt;
    g.in_mean = a.in_mean; g.in_rstd = a.in_rstd; g.bn_mean = a.bn_mean; g.bn_rstd = a.bn_rstd;
    g.d_in_w = a.d_in_w; g.d_in_b = a.d_in_b; g.d_bn_w = a.d_bn_w; g.d_bn_b = a.d_bn_b;
    return g;
}

// save: [in_mean N*half | in_rstd N*half | bn_mean C-half | bn_rstd C-half | pad | polled words (2*N*C + 8*C + 2)]
static size_t ibn_stats_floats(int N, int C, int half) { return (2 * (size_t)N * half + 2 * (size_t)(C - half) + 1) & ~(size_t)1; }
extern "C" size_t cnsn_ibn_save_floats(int N, int C, int half) { return ibn_stats_floats(N, C, half) + 2 * (size_t)N * C + 8 * (size_t)C + 8; }
extern "C" size_t cnsn_ibn_workspace_floats(int N, int C) { return 2 * (size_t)N * C + 8 * (size_t)C + 8; }

extern "C" int cnsn_ibn_resident(int dtype, int N, int C, int H, int W, int half, int training) {
    if (check_dims(N, C, H, W) || half < 0 || half > C || dtype < CNSN_F32 || dtype > CNSN_F16) return 0;
    flow::IArgs a{};
    a.N = N; a.C = C; a.M = H * W; a.half = half; a.training = training;
    if (half == 0 && odd_planes(dtype, H * W))
        return flow::bn_grp_fwd(nullptr, nullptr, dtype, N, C, H * W, nullptr, nullptr, nullptr, nullptr, nullptr, training, 0, 0.f, 0.f,
                                nullptr, nullptr, nullptr, nullptr, true) == 0 &&
               flow::bn_grp_bwd(nullptr, nullptr, nullptr, dtype, N, C, H * W, nullptr, nullptr, training, 0, nullptr, nullptr, nullptr,
                                nullptr, nullptr, nullptr, true) == 0;
    if (flow::launch_ibn<false>(a, dtype, nullptr, nullptr, true) != 0) return 0;
    return flow::launch_ibn<true>(a, dtype, nullptr, nullptr, true) == 0 ? 1 : 0;
}

extern "C" int cnsn_ibn_fwd(const void* x, void* y, int dtype, int N, int C, int H, int W, int half,
                            const cnsn_ibn_params* p, int training, int relu, float momentum, float eps_in, float eps_bn,
                            float* save, void* stream) {
    if (!x || !y || !p || !save || check_dims(N, C, H, W) || half < 0 || half > C) return CNSN_E_BADARG;
    if (dtype < CNSN_F32 || dtype > CNSN_F16) return CNSN_E_BADARG;
    if ((half > 0 && (!p->in_w || !p->in_b)) || (half < C && (!p->bn_w || !p->bn_b || !p->run_mean || !p->run_var))) return CNSN_E_BADARG;
    if (reinterpret_cast<uintptr_t>(x) % esize(dtype) || reinterpret_cast<uintptr_t>(y) % esize(dtype)) return CNSN_E_ALIGN;
    const bool vec = aligned16(x) && aligned16(y);   // 16-byte base pointers: resident kernel; else the element-wise general path
    if (training && half < C && (long long)N * H * W < 2) return CNSN_E_BATCH1;
    flow::IArgs a{};
    a.x = x; a.dy = nullptr; a.out = y; a.N = N; a.C = C; a.M = H * W; a.half = half;
    a.training = training; a.relu = relu ? 1 : 0; a.momentum = momentum; a.eps_in = eps_in; a.eps_bn = eps_bn;
    a.in_w = p->in_w; a.in_b = p->in_b; a.bn_w = p->bn_w; a.bn_b = p->bn_b;
    a.run_mean = p->run_mean; a.run_var = p->run_var; a.nbt = p->nbt;
    a.in_mean = save; a.in_rstd = save + (size_t)N * half;
    a.bn_mean = save + 2 * (size_t)N * half; a.bn_rstd = a.bn_mean + (C - half);
    float* scratch = save + ibn_stats_floats(N, C, half);
    int rc = CNSN_E_UNSUPPORTED;
    if (vec && half == 0 && odd_planes(dtype, a.M))
        rc = from_grp(flow::bn_grp_fwd(x, y, dtype, N, C, a.M, a.bn_w, a.bn_b, a.run_mean, a.run_var, a.nbt, training, a.relu, momentum,
                                       eps_bn, a.bn_mean, a.bn_rstd, scratch, (cudaStream_t)stream, false));
    else if (vec)
        rc = flow::launch_ibn<false>(a, dtype, scratch, (cudaStream_t)stream);
    if (rc != CNSN_E_UNSUPPORTED || relu) return rc;  // the fused ReLU exists in the resident kernel only (cnsn_ibn_resident)
    ibn_general::GArgs g = general_args(a);          // odd / oversized planes, misaligned slices: the three-kernel path
    return ibn_general::ibn_general_fwd(g, dtype, scratch, (cudaStream_t)stream);
}

extern "C" int cnsn_ibn_bwd(const void* x, const void* dy, void* dx, int dtype, int N, int C, int H, int W, int half,
                            const cnsn_ibn_params* p, int training, int relu, const float* save,
                            float* d_in_w, float* d_in_b, float* d_bn_w, float* d_bn_b,
                            float* workspace, void* stream) {
    if (!x || !dy || !dx || !p || !save || !workspace || check_dims(N, C, H, W) || half < 0 || half > C) return CNSN_E_BADARG;
    if (dtype < CNSN_F32 || dtype > CNSN_F16) return CNSN_E_BADARG;
    if ((half > 0 && (!p->in_w || !d_in_w || !d_in_b)) || (half < C && (!p->bn_w || !d_bn_w || !d_bn_b))) return CNSN_E_BADARG;
    if (relu && ((half > 0 && !p->in_b) || (half < C && !p->bn_b))) return CNSN_E_BADARG;      // the mask is rebuilt from x: needs the bias
    if (reinterpret_cast<uintptr_t>(x) % esize(dtype) || reinterpret_cast<uintptr_t>(dy) % esize(dtype) ||
        reinterpret_cast<uintptr_t>(dx) % esize(dtype)) return CNSN_E_ALIGN;
    const bool vec = aligned16(x) && aligned16(dy) && aligned16(dx);
    flow::IArgs a{};
    a.x = x; a.dy = dy; a.out = dx; a.N = N; a.C = C; a.M = H * W; a.half = half;
    a.training = training; a.relu = relu ? 1 : 0;
    a.in_w = p->in_w; a.bn_w = p->bn_w; a.in_b = p->in_b; a.bn_b = p->bn_b;
    float* sv = const_cast<float*>(save);
    a.in_mean = sv; a.in_rstd = sv + (size_t)N * half;
    a.bn_mean = sv + 2 * (size_t)N * half; a.bn_rstd = a.bn_mean + (C - half);
    a.d_in_w = d_in_w; a.d_in_b = d_in_b; a.d_bn_w = d_bn_w; a.d_bn_b = d_bn_b;
    int rc = CNSN_E_UNSUPPORTED;
    if (vec && half == 0 && odd_planes(dtype, a.M))
        rc = from_grp(flow::bn_grp_bwd(x, dy, dx, dtype, N, C, a.M, a.bn_w, a.bn_b, training, a.relu, a.bn_mean, a.bn_rstd, d_bn_w, d_bn_b,
                                       workspace, (cudaStream_t)stream, false));
    else if (vec)
        rc = flow::launch_ibn<true>(a, dtype, workspace, (cudaStream_t)stream);
    if (rc != CNSN_E_UNSUPPORTED || relu) return rc;
    ibn_general::GArgs g = general_args(a);
    return ibn_general::ibn_general_bwd(g, dtype, workspace, (cudaStream_t)stream);
}
