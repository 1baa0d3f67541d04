// ibn_general.cu -- the IBN layer (models/imagenet/resnet_ibn_cnsn.py:24-44) for shapes the resident kernel
// (ibn_flow.cu) does not take: planes that are not a multiple of 16 bytes (14x14 bf16, 7x7), planes too large for
// shared memory (224x224), batches whose batch-norm channels cannot be co-resident.  Three stream-ordered kernels
// per direction, one warp per (n,c) plane, scalar coalesced accesses (any plane size):
//
//   forward : k_g_stats      exact two-pass (mean, M2) per plane                                  -> ws[0], ws[1]
//             k_g_fold_fwd   per channel: instance norm -> (mean, rstd) per plane; batch norm -> Chan merge of the N
//                            pairs (training; running statistics updated) or the running statistics (eval);
//                            ws <- per-plane (scale, shift)
//             k_g_apply<0>   y = scale*x + shift
//   backward: k_g_reduce_bwd A = sum dy, B = sum dy*xhat per plane (xhat from the saved statistics)   -> ws[0], ws[1]
//             k_g_fold_bwd   per channel: d_weight = sum_n B, d_bias = sum_n A; the means the normalisation removes
//                            (per plane for instance norm, per channel for training-mode batch norm, none in eval);
//                            ws <- per-plane (cb, cc) of dx = ca*dy + cb*x + cc
//             k_g_apply<1>   dx
// 3*S forward / 5*S backward of traffic against 2*S / 3*S for the resident kernel: the fallback, not the default.
#include "ibn_general.cuh"

namespace cnsn {
namespace ibn_general {

constexpr int kTh = 256;                // threads per CTA: 8 planes per CTA, one warp each
constexpr int kFold = 128;              // threads of a per-channel fold CTA


template <typename T> __device__ __forceinline__ float ldf(const T* p) { return static_cast<float>(*p); }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <> __device__ __forceinline__ float ldf<__half>(const __half* p) { return __half2float(*p); }
template <typename T> __device__ __forceinline__ void stf(T* p, float v) { *p = static_cast<T>(v); }
template <> __device__ __forceinline__ void stf<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ void stf<__half>(__half* p, float v) { *p = __float2half_rn(v); }

// plane index of this warp, or -1
__device__ __forceinline__ long long warp_plane(long long planes) {
    const long long w = (long long)blockIdx.x * (kTh / 32) + (threadIdx.x >> 5);
    return w < planes ? w : -1;
}

template <typename T>
__global__ void __launch_bounds__(kTh) k_g_stats(const GArgs a) {
    const long long pl = warp_plane((long long)a.N * a.C);
    if (pl < 0) return;
    const int lane = threadIdx.x & 31, M = a.M;
    const T* px = static_cast<const T*>(a.x) + pl * M;
    float s = 0.f;
    for (int i = lane; i < M; i += 32) s += ldf(px + i);
    const float mean = warp_sum(s) / M;
    s = 0.f;
    for (int i = lane; i < M; i += 32) { const float d = ldf(px + i) - mean; s = fmaf(d, d, s); }
    s = warp_sum(s);
    if (lane == 0) { a.w0[pl] = mean; a.w1[pl] = s; }
}

// sums over the CTA of kFold threads
template <int K>
__device__ __forceinline__ void fold_sums(float (&v)[K], float (*sm)[kFold / 32]) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = warp_sum(v[k]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) sm[k][warp] = v[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kFold / 32; ++w) s += sm[k][w];
        v[k] = s;
    }
}

__global__ void __launch_bounds__(kFold) k_g_fold_fwd(const GArgs a) {
    __shared__ float sm[2][kFold / 32];
    const int c = blockIdx.x, N = a.N, C = a.C, M = a.M, half = a.half;
    if (c < half) {                                          // instance norm: biased variance (nn.InstanceNorm2d)
        const float g = a.in_w[c], b = a.in_b[c];
        for (int n = threadIdx.x; n < N; n += kFold) {
            const size_t i = (size_t)n * C + c;
            const float m = a.w0[i], rs = 1.f / sqrtf(a.w1[i] * (1.f / M) + a.eps_in);
            a.in_mean[(size_t)n * half + c] = m; a.in_rstd[(size_t)n * half + c] = rs;
            a.w0[i] = rs * g; a.w1[i] = b - m * rs * g;
        }
        return;
    }
    const int cb = c - half;
    float m, rs;
    if (a.training) {                                        // Chan merge of N equal-sized (mean, M2) pairs
        float v[1] = {0.f};
        for (int n = threadIdx.x; n < N; n += kFold) v[0] += a.w0[(size_t)n * C + c];
        fold_sums<1>(v, sm);
        m = v[0] / N;
        v[0] = 0.f;
        for (int n = threadIdx.x; n < N; n += kFold) {
            const size_t i = (size_t)n * C + c;
            const float d = a.w0[i] - m;
            v[0] += a.w1[i] + M * d * d;
        }
        fold_sums<1>(v, sm);
        const float cnt = (float)N * (float)M;
        rs = 1.f / sqrtf(v[0] / cnt + a.eps_bn);
        if (threadIdx.x == 0) {
            a.run_mean[cb] = (1.f - a.momentum) * a.run_mean[cb] + a.momentum * m;
            a.run_var[cb] = (1.f - a.momentum) * a.run_var[cb] + a.momentum * (v[0] / (cnt - 1.f));
            if (a.nbt && cb == 0) *a.nbt += 1;
        }
    } else {
        m = a.run_mean[cb]; rs = 1.f / sqrtf(a.run_var[cb] + a.eps_bn);
    }
    if (threadIdx.x == 0) { a.bn_mean[cb] = m; a.bn_rstd[cb] = rs; }
    const float sc = rs * a.bn_w[cb], sh = a.bn_b[cb] - m * sc;
    __syncthreads();                                         // every thread has read its (mean, M2) pairs
    for (int n = threadIdx.x; n < N; n += kFold) { a.w0[(size_t)n * C + c] = sc; a.w1[(size_t)n * C + c] = sh; }
}

template <typename T>
__global__ void __launch_bounds__(kTh) k_g_reduce_bwd(const GArgs a) {
    const long long pl = warp_plane((long long)a.N * a.C);
    if (pl < 0) return;
    const int lane = threadIdx.x & 31, M = a.M, C = a.C, half = a.half;
    const int n = (int)(pl / C), c = (int)(pl - (long long)n * C);
    const float mean = c < half ? a.in_mean[(size_t)n * half + c] : a.bn_mean[c - half];
    const float rstd = c < half ? a.in_rstd[(size_t)n * half + c] : a.bn_rstd[c - half];
    const T* px = static_cast<const T*>(a.x) + pl * M;
    const T* pd = static_cast<const T*>(a.dy) + pl * M;
    float sa = 0.f, sb = 0.f;
    for (int i = lane; i < M; i += 32) {
        const float d = ldf(pd + i);
        sa += d;
        sb = fmaf(d, (ldf(px + i) - mean) * rstd, sb);
    }
    sa = warp_sum(sa); sb = warp_sum(sb);
    if (lane == 0) { a.w0[pl] = sa; a.w1[pl] = sb; }
}

__global__ void __launch_bounds__(kFold) k_g_fold_bwd(const GArgs a) {
    __shared__ float sm[2][kFold / 32];
    const int c = blockIdx.x, N = a.N, C = a.C, M = a.M, half = a.half;
    const bool is_in = c < half;
    const int cb = c - half;
    float v[2] = {0.f, 0.f};
    for (int n = threadIdx.x; n < N; n += kFold) { v[0] += a.w0[(size_t)n * C + c]; v[1] += a.w1[(size_t)n * C + c]; }
    fold_sums<2>(v, sm);
    if (threadIdx.x == 0) {
        if (is_in) { a.d_in_b[c] = v[0]; a.d_in_w[c] = v[1]; }
        else { a.d_bn_b[cb] = v[0]; a.d_bn_w[cb] = v[1]; }
    }
    const float gam = is_in ? a.in_w[c] : a.bn_w[cb];
    const float inv = 1.f / ((float)N * (float)M);
    // dx = ca*(dy - ma - xhat*mb), ca = gam*rstd  =>  cb = -ca*mb*rstd, cc = -ca*ma - cb*mean
    for (int n = threadIdx.x; n < N; n += kFold) {           // each thread rewrites only the pairs it read
        const size_t i = (size_t)n * C + c;
        float mean, rstd, ma, mb;
        if (is_in) {
            mean = a.in_mean[(size_t)n * half + c]; rstd = a.in_rstd[(size_t)n * half + c];
            ma = a.w0[i] * (1.f / M); mb = a.w1[i] * (1.f / M);
        } else {
            mean = a.bn_mean[cb]; rstd = a.bn_rstd[cb];
            ma = a.training ? v[0] * inv : 0.f; mb = a.training ? v[1] * inv : 0.f;
        }
        const float ca = gam * rstd;
        const float kb = -ca * mb * rstd;
        a.w0[i] = kb; a.w1[i] = -ca * ma - kb * mean;
    }
}

// BWD = false: out = w0*x + w1;  BWD = true: out = ca*dy + w0*x + w1, ca = weight*rstd of the plane
template <typename T, bool BWD>
__global__ void __launch_bounds__(kTh) k_g_apply(const GArgs a) {
    const long long pl = warp_plane((long long)a.N * a.C);
    if (pl < 0) return;
    const int lane = threadIdx.x & 31, M = a.M, C = a.C, half = a.half;
    const float kx = a.w0[pl], kc = a.w1[pl];
    float ca = 0.f;
    if (BWD) {
        const int n = (int)(pl / C), c = (int)(pl - (long long)n * C);
        ca = c < half ? a.in_w[c] * a.in_rstd[(size_t)n * half + c] : a.bn_w[c - half] * a.bn_rstd[c - half];
    }
    const T* px = static_cast<const T*>(a.x) + pl * M;
    const T* pd = BWD ? static_cast<const T*>(a.dy) + pl * M : nullptr;
    T* po = static_cast<T*>(a.out) + pl * M;
    for (int i = lane; i < M; i += 32) {
        const float lin = fmaf(kx, ldf(px + i), kc);
        stf(po + i, BWD ? fmaf(ca, ldf(pd + i), lin) : lin);
    }
}

// ws: 2*N*C floats.  Returns 0 or a cuda error.
int ibn_general_fwd(GArgs& a, int dtype, float* ws, cudaStream_t s) {
    const long long planes = (long long)a.N * a.C;
    a.w0 = ws; a.w1 = ws + planes;
    const unsigned grid = (unsigned)((planes + kTh / 32 - 1) / (kTh / 32));
    CNSN_DISPATCH_DTYPE(dtype, T, k_g_stats<T><<<grid, kTh, 0, s>>>(a));
    int rc = launch_status();
    if (rc) return rc;
    k_g_fold_fwd<<<a.C, kFold, 0, s>>>(a);
    if ((rc = launch_status())) return rc;
    CNSN_DISPATCH_DTYPE(dtype, T, k_g_apply<T, false><<<grid, kTh, 0, s>>>(a));
    return launch_status();
}

int ibn_general_bwd(GArgs& a, int dtype, float* ws, cudaStream_t s) {
    const long long planes = (long long)a.N * a.C;
    a.w0 = ws; a.w1 = ws + planes;
    const unsigned grid = (unsigned)((planes + kTh / 32 - 1) / (kTh / 32));
    CNSN_DISPATCH_DTYPE(dtype, T, k_g_reduce_bwd<T><<<grid, kTh, 0, s>>>(a));
    int rc = launch_status();
    if (rc) return rc;
    k_g_fold_bwd<<<a.C, kFold, 0, s>>>(a);
    if ((rc = launch_status())) return rc;
    CNSN_DISPATCH_DTYPE(dtype, T, k_g_apply<T, true><<<grid, kTh, 0, s>>>(a));
    return launch_status();
}

}  // namespace ibn_general
}  // namespace cnsn
