// jsd.cu -- the Jensen-Shannon consistency term of the reference's 3-view training steps
// (imagenet.py:367-376, cifar.py:173-182):
//
//   p_v = softmax(logits_v), v = clean, aug1, aug2
//   lm  = log(clamp((p_clean + p_aug1 + p_aug2) / 3, 1e-7, 1))
//   L   = ( KL(lm, p_clean) + KL(lm, p_aug1) + KL(lm, p_aug2) ) / 3,   KL(lm, p) = sum p * (log p - lm) / B
//
// In eager PyTorch this is ~15 launches on (B, classes) tensors; here: one kernel forward (one CTA per row:
// three log-softmaxes, the mixture, the three sums) plus a one-CTA deterministic sum, one kernel backward.
// Everything is latency; there is nothing to say about bandwidth.
#include "common.cuh"

namespace cnsn {

constexpr int kJsdThreads = 128;

__device__ __forceinline__ float cta_reduce(float v, float* sm, bool is_max) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float u = __shfl_xor_sync(0xffffffffu, v, o);
        v = is_max ? fmaxf(v, u) : v + u;
    }
    __syncthreads();
    if (lane == 0) sm[warp] = v;
    __syncthreads();
    float r = sm[0];
#pragma unroll
    for (int i = 1; i < kJsdThreads / 32; ++i) r = is_max ? fmaxf(r, sm[i]) : r + sm[i];
    return r;
}

struct JsdRow { float mx[3], lse[3]; };      // row maximum and log-sum-exp of every view

template <typename T>
__device__ __forceinline__ JsdRow jsd_row_stats(const T* const (&z)[3], int K, float* sm) {
    JsdRow s;
#pragma unroll
    for (int v = 0; v < 3; ++v) {
        float m = -INFINITY;
        for (int k = threadIdx.x; k < K; k += kJsdThreads) m = fmaxf(m, to_f(z[v][k]));
        m = cta_reduce(m, sm, true);
        float e = 0.f;
        for (int k = threadIdx.x; k < K; k += kJsdThreads) e += expf(to_f(z[v][k]) - m);
        e = cta_reduce(e, sm, false);
        s.mx[v] = m; s.lse[v] = logf(e);
    }
    return s;
}

// grid = B rows.  row_loss[b] = sum_v sum_k p_vk * (log p_vk - lm_k)
template <typename T>
__global__ void __launch_bounds__(kJsdThreads)
k_jsd_fwd(const T* __restrict__ z0, const T* __restrict__ z1, const T* __restrict__ z2, int K, float* __restrict__ row_loss) {
    __shared__ float sm[kJsdThreads / 32];
    const size_t b = blockIdx.x;
    const T* const z[3] = {z0 + b * K, z1 + b * K, z2 + b * K};
    const JsdRow s = jsd_row_stats<T>(z, K, sm);
    float acc = 0.f;
    for (int k = threadIdx.x; k < K; k += kJsdThreads) {
        float lp[3], p[3];
#pragma unroll
        for (int v = 0; v < 3; ++v) { lp[v] = to_f(z[v][k]) - s.mx[v] - s.lse[v]; p[v] = expf(lp[v]); }
        const float lm = logf(fminf(fmaxf((p[0] + p[1] + p[2]) * (1.f / 3.f), 1e-7f), 1.f));
#pragma unroll
        for (int v = 0; v < 3; ++v) acc += p[v] * (lp[v] - lm);       // p = 0 contributes 0 (lp is finite)
    }
    acc = cta_reduce(acc, sm, false);
    if (threadIdx.x == 0) row_loss[b] = acc;
}

// one CTA: loss = sum_b row_loss[b] / (3 B), fixed summation order
__global__ void __launch_bounds__(kJsdThreads) k_jsd_sum(const float* __restrict__ row_loss, int B, float* __restrict__ loss) {
    __shared__ float sm[kJsdThreads / 32];
    float a = 0.f;
    for (int b = threadIdx.x; b < B; b += kJsdThreads) a += row_loss[b];
    a = cta_reduce(a, sm, false);
    if (threadIdx.x == 0) *loss = a / (3.f * B);
}

// d logits_v[j] = gout/(3B) * p_vj * (G_vj - sum_k p_vk G_vk),  G_vk = log p_vk - lm_k + 1 - [1e-7 <= m_k <= 1]
template <typename T>
__global__ void __launch_bounds__(kJsdThreads)
k_jsd_bwd(const T* __restrict__ z0, const T* __restrict__ z1, const T* __restrict__ z2, int B, int K,
          const float* __restrict__ gout, T* __restrict__ d0, T* __restrict__ d1, T* __restrict__ d2) {
    __shared__ float sm[kJsdThreads / 32];
    const size_t b = blockIdx.x;
    const T* const z[3] = {z0 + b * K, z1 + b * K, z2 + b * K};
    T* const d[3] = {d0 + b * K, d1 + b * K, d2 + b * K};
    const JsdRow s = jsd_row_stats<T>(z, K, sm);
    float c[3] = {0.f, 0.f, 0.f};
    for (int k = threadIdx.x; k < K; k += kJsdThreads) {
        float lp[3], p[3];
#pragma unroll
        for (int v = 0; v < 3; ++v) { lp[v] = to_f(z[v][k]) - s.mx[v] - s.lse[v]; p[v] = expf(lp[v]); }
        const float m = (p[0] + p[1] + p[2]) * (1.f / 3.f);
        const float lm = logf(fminf(fmaxf(m, 1e-7f), 1.f));
        const float ind = (m >= 1e-7f && m <= 1.f) ? 1.f : 0.f;
#pragma unroll
        for (int v = 0; v < 3; ++v) c[v] = fmaf(p[v], lp[v] - lm + 1.f - ind, c[v]);
    }
#pragma unroll
    for (int v = 0; v < 3; ++v) c[v] = cta_reduce(c[v], sm, false);
    const float scale = gout[0] / (3.f * B);
    for (int k = threadIdx.x; k < K; k += kJsdThreads) {
        float lp[3], p[3];
#pragma unroll
        for (int v = 0; v < 3; ++v) { lp[v] = to_f(z[v][k]) - s.mx[v] - s.lse[v]; p[v] = expf(lp[v]); }
        const float m = (p[0] + p[1] + p[2]) * (1.f / 3.f);
        const float lm = logf(fminf(fmaxf(m, 1e-7f), 1.f));
        const float ind = (m >= 1e-7f && m <= 1.f) ? 1.f : 0.f;
#pragma unroll
        for (int v = 0; v < 3; ++v) d[v][k] = from_f<T>(scale * p[v] * (lp[v] - lm + 1.f - ind - c[v]));
    }
}

}  // namespace cnsn

using namespace cnsn;

extern "C" int cnsn_jsd_fwd(const void* z0, const void* z1, const void* z2, int dtype, int B, int K,
                            float* row_loss, float* loss, void* stream) {
    if (!z0 || !z1 || !z2 || !row_loss || !loss || B < 1 || K < 1) return CNSN_E_BADARG;
    if (dtype < CNSN_F32 || dtype > CNSN_F16) return CNSN_E_BADARG;
    cudaStream_t s = (cudaStream_t)stream;
    CNSN_DISPATCH_DTYPE(dtype, T, k_jsd_fwd<T><<<B, kJsdThreads, 0, s>>>((const T*)z0, (const T*)z1, (const T*)z2, K, row_loss));
    int rc = launch_status();
    if (rc) return rc;
    k_jsd_sum<<<1, kJsdThreads, 0, s>>>(row_loss, B, loss);
    return launch_status();
}

extern "C" int cnsn_jsd_bwd(const void* z0, const void* z1, const void* z2, int dtype, int B, int K,
                            const float* gout, void* d0, void* d1, void* d2, void* stream) {
    if (!z0 || !z1 || !z2 || !gout || !d0 || !d1 || !d2 || B < 1 || K < 1) return CNSN_E_BADARG;
    if (dtype < CNSN_F32 || dtype > CNSN_F16) return CNSN_E_BADARG;
    cudaStream_t s = (cudaStream_t)stream;
    CNSN_DISPATCH_DTYPE(dtype, T, k_jsd_bwd<T><<<B, kJsdThreads, 0, s>>>((const T*)z0, (const T*)z1, (const T*)z2, B, K, gout,
                                                                       (T*)d0, (T*)d1, (T*)d2));
    return launch_status();
}
