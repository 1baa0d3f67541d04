// selfnorm_fold.cuh -- argument block of the SelfNorm dataflow kernels and the channel fold of the
// shared-memory-resident ones; shared by selfnorm_flow.cu and site_flow.cu (the fused CrossNorm -> SelfNorm site).
#pragma once

#include "flow_common.cuh"

namespace cnsn {
namespace flow {

struct FArgs {
    const void* x; const void* dy; void* out;      // forward: dy == nullptr, out = y; backward: out = dx
    const void* res; void* zout;                   // forward block fusion: y = f(x + res), the sum goes to zout
    int relu;                                      // forward: y = max(y, 0); backward: dy masked where x <= 0
    int N, C, M;
    int nI;                 // items per channel and phase = ceil(N / I)
    int D;                  // look-ahead in channels (1..C)
    int training;
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
    int poll_ns;            // sleep between polls of ready[c]
    int kk;                 // channel-group kernel: adjacent channels per group
    int pf_dist;            // resident kernel: L2-prefetch the item pf_dist tickets ahead (0 = off)
    unsigned items;         // resident kernel: total tickets
    unsigned* err;          // asynchronous error word (pinned host memory), see flow_common.cuh
    unsigned long long* trace;   // debug only (CNSN_FLOW_TRACE): [items][8] globaltimer stamps, else NULL
};

// selfnorm_tmem.cu: the shared-memory + tensor-memory pipeline; -100 when it does not apply
int selfnorm_tmem_fwd(FArgs& a, int dtype, float* scratch, cudaStream_t stream);
int selfnorm_tmem_bwd(FArgs& a, int dtype, float* scratch, cudaStream_t stream);

// The CTA holding a channel's last ticket: poll the channel's N published words (they stay in registers), fold
// them, publish the channel constants as one 8-byte word at `flag`, write the per-channel outputs (forward:
// running statistics, r; backward: dgamma, dbeta, dw).  Whole CTA of TH threads (or the TH-thread group that synchronises through S); returns the constants.
template <bool BWD, int TH, typename S = CtaSync>
__device__ __forceinline__ float2 fold_publish(const FArgs& a, unsigned c, float2* flag, float p_w0, float p_w1, float p_ga,
                                               float p_b, float p_rm, float p_rv, float (*s_f)[TH / 32]) {
    constexpr int kHold = 4;                                 // published words a folding thread keeps in registers
    const int N = a.N, C = a.C;
    const float2* pb = a.pub + (size_t)c * N;
    const float invN = 1.f / N;
    float2 hold[kHold];
#pragma unroll
    for (int u = 0; u < kHold; ++u) {
        const int k = S::tid() + u * TH;
        hold[u] = make_float2(0.f, 0.f);
        if (k < N) hold[u] = poll_word(pb + k, 100, a.err);
    }
    for (int k = S::tid() + kHold * TH; k < N; k += TH) poll_word(pb + k, 100, a.err);   // N > kHold*TH: re-read below
    float v[2] = {0.f, 0.f};
    float2 cst;
    if (!BWD) {
        float m = p_rm, q = p_rv;
        if (a.training) {
#pragma unroll
            for (int u = 0; u < kHold; ++u) if (S::tid() + u * TH < N) v[0] += fmaf(p_w0, hold[u].x, p_w1 * hold[u].y);
            for (int k = S::tid() + kHold * TH; k < N; k += TH) { const float2 p = ll_peek(pb + k); v[0] += fmaf(p_w0, p.x, p_w1 * p.y); }
            cta_sums<1, TH, S>(*reinterpret_cast<float(*)[1]>(&v[0]), reinterpret_cast<float(*)[TH / 32]>(s_f[0]));
            m = v[0] / N;
            v[1] = 0.f;
#pragma unroll
            for (int u = 0; u < kHold; ++u)
                if (S::tid() + u * TH < N) { const float d = fmaf(p_w0, hold[u].x, p_w1 * hold[u].y) - m; v[1] = fmaf(d, d, v[1]); }
            for (int k = S::tid() + kHold * TH; k < N; k += TH) {
                const float2 p = ll_peek(pb + k);
                const float d = fmaf(p_w0, p.x, p_w1 * p.y) - m;
                v[1] = fmaf(d, d, v[1]);
            }
            cta_sums<1, TH, S>(*reinterpret_cast<float(*)[1]>(&v[1]), reinterpret_cast<float(*)[TH / 32]>(s_f[1]));
            q = v[1] / N;                                // biased variance normalises (BatchNorm semantics)
        }
        // eval: thread 0 holds the running statistics; the other threads' m, q are unused
        const float rstd = 1.f / sqrtf(q + a.bn_eps);
        cst = make_float2(m, rstd);
        if (S::tid() == 0) {
            ll_publish(flag, cst.x, cst.y);       // the channel is ready: 8 bytes, no fence
                        a.r[c] = rstd;
            if (a.training) {
                a.run_mean[c] = (1.f - a.momentum) * p_rm + a.momentum * m;
                a.run_var[c] = (1.f - a.momentum) * p_rv + a.momentum * (q * N / (N - 1.f));
                if (a.nbt && c == 0) *a.nbt += 1;
            }
        }
    } else {
#pragma unroll
        for (int u = 0; u < kHold; ++u) if (S::tid() + u * TH < N) { v[0] = fmaf(hold[u].x, hold[u].y, v[0]); v[1] += hold[u].x; }
        for (int k = S::tid() + kHold * TH; k < N; k += TH) { const float2 p = ll_peek(pb + k); v[0] = fmaf(p.x, p.y, v[0]); v[1] += p.x; }
        cta_sums<2, TH, S>(v, s_f);
        const float dgam = v[0], dbet = v[1];
        const float k1 = a.training ? p_ga * dbet * invN : 0.f, k2 = a.training ? p_ga * dgam * invN : 0.f;
        cst = make_float2(k1, k2);
        if (S::tid() == 0) {
            ll_publish(flag, cst.x, cst.y);
                        a.dgamma[c] = dgam; a.dbeta[c] = dbet;
        }
        // off the critical path: dw = (sum ds*mu, sum ds*sd)
        v[0] = v[1] = 0.f;
#pragma unroll
        for (int u = 0; u < kHold; ++u) {
            const int k = S::tid() + u * TH;
            if (k < N) {
                const size_t i = (size_t)k * C + c;
                const float ds = p_b * (hold[u].x * p_ga - k1 - hold[u].y * k2);
                v[0] = fmaf(ds, a.mu[i], v[0]); v[1] = fmaf(ds, a.sd[i], v[1]);
            }
        }
        for (int k = S::tid() + kHold * TH; k < N; k += TH) {
            const float2 p = ll_peek(pb + k);
            const size_t i = (size_t)k * C + c;
            const float ds = p_b * (p.x * p_ga - k1 - p.y * k2);
            v[0] = fmaf(ds, a.mu[i], v[0]); v[1] = fmaf(ds, a.sd[i], v[1]);
        }
        cta_sums<2, TH, S>(v, s_f);
        if (S::tid() == 0) { a.dw[2 * c] = v[0]; a.dw[2 * c + 1] = v[1]; }
    }
    return cst;
}

}  // namespace flow
}  // namespace cnsn
