// bn_nhwc.cuh -- the channels-last batch-norm kernels (see bn_nhwc.cu for the design): shared by bn_nhwc.cu (the
// cnsn_bn_nhwc_* entry points) and selfnorm_nhwc.cu (the fused bottleneck tail bn3 -> add -> SelfNorm -> ReLU).
#pragma once

#include <algorithm>

#include "flow_common.cuh"

namespace cnsn {
namespace bnl {

constexpr int kT = 256;

struct Geom {
    long long R;       // rows = N * H * W
    int C;
    int CG;            // 16-byte vectors per row
    int CGB;           // vectors per row handled by one CTA (channel block); CG % CGB == 0
    int RL;            // row lanes = kT / CGB
    int G;             // row chunks
    long long rows;    // rows per chunk
};

__device__ __forceinline__ double shfl_xor_d(double v, int o) { return __shfl_xor_sync(0xffffffffu, v, o); }

template <typename T>
__global__ void __launch_bounds__(kT) k_bn_nhwc_stats(const T* __restrict__ x, const Geom g, float2* __restrict__ part) {
    constexpr int V = VecOf<T>::n;
    extern __shared__ float sm[];                            // mean [RL][CGB*V] | m2 [RL][CGB*V]
    const int W = g.CGB * V;                                 // channels of this CTA
    float* s_mean = sm;
    float* s_m2 = sm + g.RL * W;
    const int cg = threadIdx.x % g.CGB, rl = threadIdx.x / g.CGB;
    const long long row0 = (long long)blockIdx.x * g.rows;
    const long long nrows = min(g.rows, g.R - row0);
    const uint4* vx = reinterpret_cast<const uint4*>(x) + (size_t)row0 * g.CG + (size_t)blockIdx.y * g.CGB + cg;
    float k[V], s1[V], s2[V];
#pragma unroll
    for (int e = 0; e < V; ++e) { k[e] = 0.f; s1[e] = 0.f; s2[e] = 0.f; }
    if (rl < nrows) unpack<T>(__ldg(vx + (size_t)rl * g.CG), k);
    int nt = 0;
#pragma unroll 8
    for (long long r = rl; r < nrows; r += g.RL) {
        float a[V];
        unpack<T>(__ldg(vx + (size_t)r * g.CG), a);
#pragma unroll
        for (int e = 0; e < V; ++e) { const float d = a[e] - k[e]; s1[e] += d; s2[e] = fmaf(d, d, s2[e]); }
        ++nt;
    }
    const float inv = nt ? 1.f / nt : 0.f;
#pragma unroll
    for (int e = 0; e < V; ++e) {
        s_mean[rl * W + cg * V + e] = k[e] + s1[e] * inv;
        const float m2 = s2[e] - s1[e] * s1[e] * inv;
        s_m2[rl * W + cg * V + e] = m2 < 0.f ? 0.f : m2;      // (not fmaxf: a NaN must stay a NaN)
    }
    __syncthreads();
    // merge the row lanes (lane l holds ceil((nrows - l) / RL) rows): mean = sum n_l m_l / n, M2 = sum (M2_l + n_l (m_l - mean)^2);
    // double accumulators, no division inside the loops
    const int nbase = (int)(nrows / g.RL), nrem = (int)(nrows - (long long)nbase * g.RL);      // lane l: nbase + (l < nrem) rows
    for (int c = threadIdx.x; c < W; c += kT) {
        double sm1 = 0.0;
        for (int l = 0; l < g.RL; ++l) sm1 += (double)(nbase + (l < nrem)) * (double)s_mean[l * W + c];
        const double m = sm1 / (double)nrows;
        double q = 0.0;
        for (int l = 0; l < g.RL; ++l) {
            const double d = (double)s_mean[l * W + c] - m;
            q += (double)s_m2[l * W + c] + (double)(nbase + (l < nrem)) * d * d;
        }
        part[(size_t)blockIdx.x * g.C + (size_t)blockIdx.y * W + c] = make_float2((float)m, (float)q);
    }
}

constexpr int kFoldT = 128;
// sum of v over the CTA's kFoldT threads, in double; every thread gets the total
__device__ __forceinline__ double fold_sum(double v, double* sm) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += shfl_xor_d(v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < kFoldT / 32; ++w) t += sm[w];
    return t;
}

// one CTA per channel.  training: merge the G chunk pairs -- mean = sum n_j m_j / R, M2 = sum (M2_j + n_j (m_j - mean)^2),
// two passes over the (L2-resident) pairs held in registers; eval: the running statistics.
static __global__ void __launch_bounds__(kFoldT) k_bn_nhwc_fold(const float2* __restrict__ part, const Geom g, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, float* __restrict__ run_mean,
                                                         float* __restrict__ run_var, long long* __restrict__ nbt, int training,
                                                         float momentum, float eps, float* __restrict__ save_mean,
                                                         float* __restrict__ save_rstd, float2* __restrict__ coef) {
    __shared__ double sm[kFoldT / 32];
    const int c = blockIdx.x;
    // the channel's scalars first: their (cold) loads overlap the loads of the chunk pairs instead of trailing the reductions
    const float ga = gamma[c], be = beta[c], rm0 = run_mean[c], rv0 = run_var[c];
    float mean, rstd;
    if (training) {
        constexpr int kHold = 8;                             // G <= 592 in practice: every pair stays in registers
        float2 hold[kHold];
        double s1 = 0.0;
#pragma unroll
        for (int u = 0; u < kHold; ++u) {
            const int j = threadIdx.x + u * kFoldT;
            hold[u] = make_float2(0.f, 0.f);
            if (j < g.G) {
                hold[u] = __ldcg(part + (size_t)j * g.C + c);
                s1 += (double)min(g.rows, g.R - (long long)j * g.rows) * (double)hold[u].x;
            }
        }
        for (int j = threadIdx.x + kHold * kFoldT; j < g.G; j += kFoldT)
            s1 += (double)min(g.rows, g.R - (long long)j * g.rows) * (double)__ldcg(part + (size_t)j * g.C + c).x;
        const double ma = fold_sum(s1, sm) / (double)g.R;
        double qa = 0.0;
#pragma unroll
        for (int u = 0; u < kHold; ++u) {
            const int j = threadIdx.x + u * kFoldT;
            if (j < g.G) {
                const double d = (double)hold[u].x - ma;
                qa += (double)hold[u].y + (double)min(g.rows, g.R - (long long)j * g.rows) * d * d;
            }
        }
        for (int j = threadIdx.x + kHold * kFoldT; j < g.G; j += kFoldT) {
            const float2 p = __ldcg(part + (size_t)j * g.C + c);
            const double d = (double)p.x - ma;
            qa += (double)p.y + (double)min(g.rows, g.R - (long long)j * g.rows) * d * d;
        }
        qa = fold_sum(qa, sm);
        const double cnt = (double)g.R;
        mean = (float)ma;
        rstd = 1.f / sqrtf((float)(qa / cnt) + eps);
        if (threadIdx.x == 0) {
            run_mean[c] = (1.f - momentum) * rm0 + momentum * mean;
            run_var[c] = (1.f - momentum) * rv0 + momentum * (float)(qa / (cnt - 1.0));
            if (nbt && c == 0) *nbt += 1;
        }
    } else {
        mean = rm0;
        rstd = 1.f / sqrtf(rv0 + eps);
    }
    if (threadIdx.x == 0) {
        save_mean[c] = mean; save_rstd[c] = rstd;
        const float sc = rstd * ga;
        coef[c] = make_float2(sc, be - mean * sc);
    }
}

// per chunk and channel: (sum d, sum d * xhat)
template <typename T, bool RELU>
__global__ void __launch_bounds__(kT, 4) k_bn_nhwc_reduce(const T* __restrict__ x, const T* __restrict__ dy, const Geom g,
                                                       const float* __restrict__ save_mean, const float* __restrict__ save_rstd,
                                                       const float2* __restrict__ coef, float2* __restrict__ part) {
    constexpr int V = VecOf<T>::n;
    extern __shared__ float sm[];                            // a [RL][W] | b [RL][W]
    const int W = g.CGB * V;
    float* s_a = sm;
    float* s_b = sm + g.RL * W;
    const int cg = threadIdx.x % g.CGB, rl = threadIdx.x / g.CGB;
    const long long row0 = (long long)blockIdx.x * g.rows;
    const long long nrows = min(g.rows, g.R - row0);
    const size_t vb = (size_t)row0 * g.CG + (size_t)blockIdx.y * g.CGB + cg;
    const uint4* vx = reinterpret_cast<const uint4*>(x) + vb;
    const uint4* vd = reinterpret_cast<const uint4*>(dy) + vb;
    const int c0 = blockIdx.y * W + cg * V;
    float mean[V], fs[V], fb[V], a[V], b[V];                // b accumulates d * (x - mean); rstd multiplies the column total
#pragma unroll
    for (int e = 0; e < V; ++e) {
        mean[e] = save_mean[c0 + e];
        const float2 f = RELU ? coef[c0 + e] : make_float2(0.f, 0.f);
        fs[e] = f.x; fb[e] = f.y; a[e] = 0.f; b[e] = 0.f;
    }
#pragma unroll 4
    for (long long r = rl; r < nrows; r += g.RL) {
        float xv[V], dv[V];
        unpack<T>(__ldg(vx + (size_t)r * g.CG), xv);
        unpack<T>(__ldg(vd + (size_t)r * g.CG), dv);
#pragma unroll
        for (int e = 0; e < V; ++e) {
            const float d = (RELU && !(fmaf(fs[e], xv[e], fb[e]) > 0.f)) ? 0.f : dv[e];
            a[e] += d;
            b[e] = fmaf(d, xv[e] - mean[e], b[e]);
        }
    }
#pragma unroll
    for (int e = 0; e < V; ++e) { s_a[rl * W + cg * V + e] = a[e]; s_b[rl * W + cg * V + e] = b[e]; }
    __syncthreads();
    for (int c = threadIdx.x; c < W; c += kT) {
        double ta = 0.0, tb = 0.0;
        for (int l = 0; l < g.RL; ++l) { ta += (double)s_a[l * W + c]; tb += (double)s_b[l * W + c]; }
        tb *= (double)save_rstd[blockIdx.y * W + c];
        part[(size_t)blockIdx.x * g.C + (size_t)blockIdx.y * W + c] = make_float2((float)ta, (float)tb);
    }
}

// one CTA per channel: dbeta = sum d, dgamma = sum d * xhat; dx = ca * d + cb * x + cc
static __global__ void __launch_bounds__(kFoldT) k_bn_nhwc_fold_bwd(const float2* __restrict__ part, const Geom g, const float* __restrict__ gamma,
                                                             int training, const float* __restrict__ save_mean,
                                                             const float* __restrict__ save_rstd, float* __restrict__ dgamma,
                                                             float* __restrict__ dbeta, float* __restrict__ cdx) {
    __shared__ double sm[kFoldT / 32];
    const int c = blockIdx.x;
    const float ga = gamma[c], mean = save_mean[c], rstd = save_rstd[c];     // early: overlaps the loads of the chunk pairs
    double ta = 0.0, tb = 0.0;
#pragma unroll 4
    for (int j = threadIdx.x; j < g.G; j += kFoldT) { const float2 p = __ldcg(part + (size_t)j * g.C + c); ta += (double)p.x; tb += (double)p.y; }
    ta = fold_sum(ta, sm);
    tb = fold_sum(tb, sm);
    if (threadIdx.x == 0) {
        dbeta[c] = (float)ta; dgamma[c] = (float)tb;
        const float inv = training ? 1.f / (float)g.R : 0.f;     // eval: the statistics are constants, nothing to remove
        const float ma = (float)ta * inv, mb = (float)tb * inv;
        const float ca = ga * rstd;
        const float cb = -ca * mb * rstd;
        cdx[c] = ca; cdx[g.C + c] = cb; cdx[2 * g.C + c] = -ca * ma - cb * mean;
    }
}

// RELU: compile-time, so that the backward without a ReLU (bn3 of a bottleneck, the projection's batch norm: the widest
// tensors) does not carry the forward map's coefficients in registers
template <typename T, bool BWD, bool RELU>
__global__ void __launch_bounds__(kT, 4) k_bn_nhwc_apply(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ out, const Geom g,
                                                      const float2* __restrict__ coef, const float* __restrict__ cdx, int reverse) {
    constexpr bool relu = RELU;
    constexpr int V = VecOf<T>::n;
    const int W = g.CGB * V;
    const int cg = threadIdx.x % g.CGB, rl = threadIdx.x / g.CGB;
    // reverse: chunks in REVERSE launch order -- the kernel in front of this one (statistics / reduction) read the same tensors
    // front to back, so the last chunks are the ones still in L2
    const long long row0 = (long long)(reverse ? gridDim.x - 1 - blockIdx.x : blockIdx.x) * g.rows;
    const long long nrows = min(g.rows, g.R - row0);
    const size_t vb = (size_t)row0 * g.CG + (size_t)blockIdx.y * g.CGB + cg;
    const uint4* vx = reinterpret_cast<const uint4*>(x) + vb;
    const uint4* vd = reinterpret_cast<const uint4*>(dy) + vb;
    uint4* vo = reinterpret_cast<uint4*>(out) + vb;
    const int c0 = blockIdx.y * W + cg * V;
    float fs[V], fb[V], ca[V], cb[V], cc[V];
#pragma unroll
    for (int e = 0; e < V; ++e) {
        const float2 f = (!BWD || RELU) ? coef[c0 + e] : make_float2(0.f, 0.f);
        fs[e] = f.x; fb[e] = f.y;
        ca[e] = BWD ? cdx[c0 + e] : 0.f; cb[e] = BWD ? cdx[g.C + c0 + e] : 0.f; cc[e] = BWD ? cdx[2 * g.C + c0 + e] : 0.f;
    }
#pragma unroll 4
    for (long long r = rl; r < nrows; r += g.RL) {
        float xv[V], dv[V], o[V];
        unpack<T>(ldg_stream(vx + (size_t)r * g.CG), xv);
        if (BWD) unpack<T>(ldg_stream(vd + (size_t)r * g.CG), dv);
#pragma unroll
        for (int e = 0; e < V; ++e) {
            if (BWD) {
                const float d = (relu && !(fmaf(fs[e], xv[e], fb[e]) > 0.f)) ? 0.f : dv[e];
                o[e] = fmaf(ca[e], d, fmaf(cb[e], xv[e], cc[e]));
            } else {
                const float y = fmaf(fs[e], xv[e], fb[e]);
                o[e] = relu ? fmaxf(y, 0.f) : y;
            }
        }
        vo[(size_t)r * g.CG] = pack<T>(o);
    }
}

constexpr int kMaxChunks = 2048;

inline int make_geom(Geom& g, int dtype, int N, int C, int H, int W) {
    const int esz = (int)esize(dtype);
    if (((size_t)C * esz) % 16) return CNSN_E_UNSUPPORTED;
    g.R = (long long)N * H * W;
    g.C = C;
    g.CG = C * esz / 16;
    g.CGB = std::min(g.CG, kT);
    if (kT % g.CGB || g.CG % g.CGB) return CNSN_E_UNSUPPORTED;
    if (g.R < 2) return CNSN_E_UNSUPPORTED;
    g.RL = kT / g.CGB;
    const int cblocks = g.CG / g.CGB;
    // 4 CTAs per SM over all channel blocks whatever the tensor size (measured on the WideResNet / ResNet-50 steps: 8 per SM
    // costs more in the fold than it gains, fewer-but-larger chunks for small tensors lose bandwidth), at least 4 rows per
    // row lane, at most kMaxChunks chunks
    long long G = std::max<long long>(1, (148 * 4) / cblocks);
    G = std::min<long long>(G, std::max<long long>(1, g.R / (4 * g.RL)));
    G = std::min<long long>(G, kMaxChunks);
    g.rows = (g.R + G - 1) / G;
    g.G = (int)((g.R + g.rows - 1) / g.rows);
    return 0;
}
inline size_t smem_bytes(const Geom& g, int dtype) { return 2 * (size_t)g.RL * g.CGB * (16 / esize(dtype)) * sizeof(float); }

}  // namespace bnl
}  // namespace cnsn

