// selfnorm_gate.cuh -- the O(N*C) part of SelfNorm shared by the three-kernel path (selfnorm.cu, NCHW) and the
// channels-last path (selfnorm_nhwc.cu): the gate (fc -> BatchNorm1d over the batch -> sigmoid, models/cnsn.py:113-150)
// forward and backward on the [N][C] statistics, and the layout of the save block.  The kernels are `static`: each
// translation unit that includes this header launches its own copy.
#pragma once

#include <math.h>

#include "common.cuh"

namespace cnsn {

constexpr int kGateThreads = 256;   // one CTA per (channel, gate branch); thread t owns samples t, t+256, ...

struct GateFwd {              // one gate branch (g or f) as seen by the forward gate kernel
    const float* w; const float* gamma; const float* beta;
    float* run_mean; float* run_var; long long* nbt;
    float* gate; float* shat; float* r;     // outputs into the save block
};
struct GateBwd {
    const float* w; const float* gamma;
    const float* gate; const float* shat; const float* r;
    float* dw; float* dgamma; float* dbeta;
};

// Block-wide sums of K values over kGateThreads threads; every thread gets the totals.
template <int K>
__device__ __forceinline__ void gate_block_sums(float (&v)[K], float (*sm)[kGateThreads / 32]) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = warp_sum(v[k]);
    __syncthreads();                         // sm may still be read from the previous call
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) sm[k][warp] = v[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kGateThreads / 32; ++w) s += sm[k][w];
        v[k] = s;
    }
}

__device__ __forceinline__ float sigmoidf_acc(float z) { return 1.f / (1.f + expf(-z)); }

// grid = (C, n_gates); block = 256.  The (N,C) statistics are tiny; what matters is latency, so every
// channel gets its own CTA (the first version used C/32 CTAs and took 15-25 us on its own).
static __global__ void __launch_bounds__(kGateThreads)
k_sn_gate_fwd(const float* __restrict__ mu, const float* __restrict__ sd, GateFwd g0, GateFwd g1,
              int N, int C, int training, float momentum, float bn_eps) {
    __shared__ float sm[1][kGateThreads / 32];
    const GateFwd g = blockIdx.y == 0 ? g0 : g1;
    const int c = blockIdx.x;
    const float w0 = g.w[2 * c], w1 = g.w[2 * c + 1];
    float m, q;
    if (training) {
        float acc[1] = {0.f};
        for (int n = threadIdx.x; n < N; n += kGateThreads)
            acc[0] += fmaf(w0, mu[(size_t)n * C + c], w1 * sd[(size_t)n * C + c]);
        gate_block_sums<1>(acc, sm);
        m = acc[0] / N;
        acc[0] = 0.f;
        for (int n = threadIdx.x; n < N; n += kGateThreads) {
            const float d = fmaf(w0, mu[(size_t)n * C + c], w1 * sd[(size_t)n * C + c]) - m;
            acc[0] = fmaf(d, d, acc[0]);
        }
        gate_block_sums<1>(acc, sm);
        q = acc[0] / N;                      // biased variance normalises (BatchNorm semantics)
        if (threadIdx.x == 0) {
            g.run_mean[c] = (1.f - momentum) * g.run_mean[c] + momentum * m;
            g.run_var[c] = (1.f - momentum) * g.run_var[c] + momentum * (q * N / (N - 1.f));
            if (g.nbt && c == 0) *g.nbt += 1;
        }
    } else {
        m = g.run_mean[c];
        q = g.run_var[c];
    }
    const float r = 1.f / sqrtf(q + bn_eps);
    const float ga = g.gamma[c], be = g.beta[c];
    if (threadIdx.x == 0) g.r[c] = r;
    for (int n = threadIdx.x; n < N; n += kGateThreads) {
        const size_t i = (size_t)n * C + c;
        const float sh = (fmaf(w0, mu[i], w1 * sd[i]) - m) * r;
        g.shat[i] = sh;
        g.gate[i] = sigmoidf_acc(fmaf(ga, sh, be));
    }
}

// Gate backward for one channel per CTA.  Produces the parameter gradients and the two per-instance
// coefficients of  dx = g*dy + cb*x + cc   (cb = b, cc = a - b*mu).
// grid = C; block = 256
static __global__ void __launch_bounds__(kGateThreads)
k_sn_gate_bwd(const float* __restrict__ mu, const float* __restrict__ sd,
              const float* __restrict__ sxy, const float* __restrict__ st,
              GateBwd g, GateBwd f, int two, int N, int C, int M, int training,
              float* __restrict__ cb, float* __restrict__ cc) {
    __shared__ float sm[4][kGateThreads / 32];
    const int c = blockIdx.x;
    const float invN = 1.f / N;
    // pass A: dgamma = sum dz*shat, dbeta = sum dz for each gate
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int n = threadIdx.x; n < N; n += kGateThreads) {
        const size_t i = (size_t)n * C + c;
        const float gg = g.gate[i];
        if (!two) {
            const float dz = sxy[i] * gg * (1.f - gg);
            acc[0] = fmaf(dz, g.shat[i], acc[0]); acc[1] += dz;
        } else {                             // sxy holds sum dy*(x-mu): dgate_g = that, dgate_f = mu*T
            const float ff = f.gate[i];
            const float dzg = sxy[i] * gg * (1.f - gg);
            const float dzf = mu[i] * st[i] * ff * (1.f - ff);
            acc[0] = fmaf(dzg, g.shat[i], acc[0]); acc[1] += dzg;
            acc[2] = fmaf(dzf, f.shat[i], acc[2]); acc[3] += dzf;
        }
    }
    gate_block_sums<4>(acc, sm);
    const float dgam_g = acc[0], dbet_g = acc[1], dgam_f = acc[2], dbet_f = acc[3];
    const float gw0 = g.w[2 * c], gw1 = g.w[2 * c + 1], gga = g.gamma[c], gr = g.r[c];
    float fw0 = 0.f, fw1 = 0.f, fga = 0.f, fr = 0.f;
    if (two) { fw0 = f.w[2 * c]; fw1 = f.w[2 * c + 1]; fga = f.gamma[c]; fr = f.r[c]; }
    if (threadIdx.x == 0) {
        g.dgamma[c] = dgam_g; g.dbeta[c] = dbet_g;
        if (two) { f.dgamma[c] = dgam_f; f.dbeta[c] = dbet_f; }
    }
    // pass B: ds, dw = (sum ds*mu, sum ds*sd), coefficients
    const float k1g = training ? gga * dbet_g * invN : 0.f, k2g = training ? gga * dgam_g * invN : 0.f;
    const float k1f = training ? fga * dbet_f * invN : 0.f, k2f = training ? fga * dgam_f * invN : 0.f;
    const float invM = 1.f / M, invM1 = 1.f / (M - 1.f);
    acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
    for (int n = threadIdx.x; n < N; n += kGateThreads) {
        const size_t i = (size_t)n * C + c;
        const float mean = mu[i], sdev = sd[i];
        const float gg = g.gate[i];
        const float dzg = sxy[i] * gg * (1.f - gg);
        const float dsg = gr * (dzg * gga - k1g - g.shat[i] * k2g);
        acc[0] = fmaf(dsg, mean, acc[0]); acc[1] = fmaf(dsg, sdev, acc[1]);
        float dmu = dsg * gw0, dsd = dsg * gw1;
        if (two) {
            const float ff = f.gate[i], T = st[i];
            const float dzf = mean * T * ff * (1.f - ff);
            const float dsf = fr * (dzf * fga - k1f - f.shat[i] * k2f);
            acc[2] = fmaf(dsf, mean, acc[2]); acc[3] = fmaf(dsf, sdev, acc[3]);
            dmu += dsf * fw0 + (ff - gg) * T;
            dsd += dsf * fw1;
        }
        const float b = dsd * invM1 / sdev;
        cb[i] = b;
        cc[i] = dmu * invM - b * mean;
    }
    gate_block_sums<4>(acc, sm);
    if (threadIdx.x == 0) {
        g.dw[2 * c] = acc[0]; g.dw[2 * c + 1] = acc[1];
        if (two) { f.dw[2 * c] = acc[2]; f.dw[2 * c + 1] = acc[3]; }
    }
}

struct SaveLayout {           // offsets (in floats) into the save block
    size_t mu, sd, g, shat_g, f, shat_f, r_g, r_f, scratch, total;
    SaveLayout(int N, int C, bool two) {
        const size_t nc = (size_t)N * C;
        mu = 0; sd = nc; g = 2 * nc; shat_g = 3 * nc;
        f = 4 * nc; shat_f = 5 * nc;
        r_g = two ? 6 * nc : 4 * nc;
        r_f = r_g + C;
        scratch = (r_g + (two ? 2 : 1) * (size_t)C + 1) & ~(size_t)1;   // 8-byte aligned
        total = scratch + 2 * nc + 36 * (size_t)C + 8;  // [C][N] published (mu, sd) words of the fused / flow kernels,
                                                       // then the flow kernel's [C] constants, counters, ticket
    }
};

}  // namespace cnsn
