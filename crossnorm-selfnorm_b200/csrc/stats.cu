// stats.cu -- per-instance statistics (calc_ins_mean_std, models/cnsn.py:8-17), its backward,
// and the per-instance affine map used by instance_norm_mix (models/cnsn.py:27-29).
//
// Roofline: HBM.  Algorithmic bytes per instance: stats = M*sizeof(T) (one read);
// stats_bwd / affine = 2*M*sizeof(T) (read x, write out).
#include "common.cuh"

namespace cnsn {

template <typename T, int TPI, bool VEC>
__global__ void __launch_bounds__(kBlock)
k_instance_stats(const T* __restrict__ x, long long instances, int W, int M, Window win, bool win_full,
                 float eps, float* __restrict__ mean, float* __restrict__ sd) {
    __shared__ Moments scratch[kWarpsPerBlock];
    const long long inst = Team<TPI>::instance();
    if (inst >= instances) return;
    Moments m = instance_moments<T, TPI, VEC>(x + inst * M, W, M, win, win_full);
    m = Team<TPI>::all_merge(m, scratch);
    if (Team<TPI>::rank() == 0) {
        mean[inst] = m.mean;
        sd[inst] = std_from(m, eps);
    }
}

// dx = dmean/Mw + (x-mean)/sd * dsd/(Mw-1) inside the window, 0 outside.
template <typename T, int TPI, bool VEC>
__global__ void __launch_bounds__(kBlock)
k_instance_stats_bwd(const T* __restrict__ x, T* __restrict__ dx, long long instances, int W, int M,
                     Window win, bool win_full, const float* __restrict__ mean,
                     const float* __restrict__ sd, const float* __restrict__ dmean,
                     const float* __restrict__ dsd) {
    const long long inst = Team<TPI>::instance();
    if (inst >= instances) return;
    const float Mw = (float)win.area();
    const float q = dsd[inst] / ((Mw - 1.f) * sd[inst]);
    const float r = dmean[inst] / Mw - q * mean[inst];
    if (win_full) {
        plane_map<T, TPI, VEC, false>(x + inst * M, nullptr, dx + inst * M, M,
                                      [=](float xv, float, int) { return fmaf(q, xv, r); });
    } else {
        plane_map<T, TPI, VEC, false>(x + inst * M, nullptr, dx + inst * M, M, [=](float xv, float, int i) {
            const int h = i / W, w = i - h * W;
            return win.has(h, w) ? fmaf(q, xv, r) : 0.f;
        });
    }
}

template <typename T, int TPI, bool VEC>
__global__ void __launch_bounds__(kBlock)
k_instance_affine(const T* __restrict__ x, T* __restrict__ out, long long instances, int M,
                  const float* __restrict__ scale, const float* __restrict__ shift) {
    const long long inst = Team<TPI>::instance();
    if (inst >= instances) return;
    const float a = scale[inst], b = shift[inst];
    plane_map<T, TPI, VEC, false>(x + inst * M, nullptr, out + inst * M, M,
                                  [=](float xv, float, int) { return fmaf(a, xv, b); });
}


// ---- strided input (sN, sC, sH, sW) -- no dense copy for sliced / transposed views or channels_last ----
// Generic: one warp per instance walks the window row-major (coalesced when sW == 1: NCHW views, crops).
template <typename T>
__global__ void __launch_bounds__(kBlock)
k_instance_stats_strided(const T* __restrict__ x, int N, int C, long long sN, long long sC, long long sH, long long sW,
                         Window win, float eps, float* __restrict__ mean, float* __restrict__ sd) {
    const long long inst = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (inst >= (long long)N * C) return;
    const int n = (int)(inst / C), c = (int)(inst - (long long)n * C);
    const T* base = x + n * sN + c * sC;
    const int cols = win.cols(), area = win.area(), lane = threadIdx.x & 31;
    Moments acc = moments_zero();
    int hh = lane / cols, ww = lane - hh * cols;
    const int dh = 32 / cols, dw = 32 - dh * cols;
    for (int i = lane; i < area; i += 32) {
        fold1(acc, to_f(base[(win.h0 + hh) * sH + (win.w0 + ww) * sW]));
        hh += dh; ww += dw;
        if (ww >= cols) { ww -= cols; ++hh; }
    }
    acc = warp_merge(acc);
    if (lane == 0) { mean[inst] = acc.mean; sd[inst] = std_from(acc, eps); }
}
// channels_last (sC == 1): a CTA takes one sample and 32 adjacent channels; lane = channel (the contiguous dimension),
// the 8 warps split the window's pixels, partial moments merge through shared memory.
template <typename T>
__global__ void __launch_bounds__(kBlock)
k_instance_stats_cl(const T* __restrict__ x, int N, int C, long long sN, long long sH, long long sW, Window win, float eps,
                    float* __restrict__ mean, float* __restrict__ sd) {
    __shared__ Moments part[kWarpsPerBlock][32];
    const int tiles = (C + 31) / 32;
    const int n = blockIdx.x / tiles, c = (blockIdx.x - n * tiles) * 32 + (threadIdx.x & 31);
    const int warp = threadIdx.x >> 5, cols = win.cols(), area = win.area();
    Moments acc = moments_zero();
    if (c < C) {
        const T* base = x + n * sN + c;
        for (int i = warp; i < area; i += kWarpsPerBlock) {
            const int hh = i / cols, ww = i - hh * cols;
            fold1(acc, to_f(base[(win.h0 + hh) * sH + (win.w0 + ww) * sW]));
        }
    }
    part[warp][threadIdx.x & 31] = acc;
    __syncthreads();
    if (warp == 0 && c < C) {
        Moments m = part[0][threadIdx.x];
#pragma unroll
        for (int w = 1; w < kWarpsPerBlock; ++w) m = merge(m, part[w][threadIdx.x]);
        mean[(size_t)n * C + c] = m.mean;
        sd[(size_t)n * C + c] = std_from(m, eps);
    }
}

int check_dims(int N, int C, int H, int W) { return (N > 0 && C > 0 && H > 0 && W > 0) ? 0 : CNSN_E_BADARG; }
int check_window(const Window& w, int H, int W) {
    return (w.h0 >= 0 && w.w0 >= 0 && w.h1 <= H && w.w1 <= W && w.h0 < w.h1 && w.w0 < w.w1) ? 0 : CNSN_E_BADARG;
}

// Internal launcher shared with selfnorm.cu / crossnorm.cu (arguments already validated).
int launch_instance_stats(const void* x, int dtype, long long inst, int H, int W, const Window& win,
                          float eps, float* mean, float* sd, cudaStream_t s) {
    const int M = H * W;
    const bool full = win.full(H, W);
    const bool vec = full && vec_ok(x, dtype, M);
    const int tpi = team_for(full ? M : win.area());
    CNSN_DISPATCH_DTYPE(dtype, T, CNSN_DISPATCH_TEAM(tpi, TPI, CNSN_DISPATCH_BOOL(vec, VEC,
        k_instance_stats<T, TPI, VEC><<<grid_for(inst, TPI), kBlock, 0, s>>>(
            (const T*)x, inst, W, M, win, full, eps, mean, sd))));
    return launch_status();
}

}  // namespace cnsn

using namespace cnsn;

extern "C" int cnsn_instance_stats(const void* x, int dtype, int N, int C, int H, int W,
                                   int h0, int h1, int w0, int w1, float eps,
                                   float* mean, float* sd, void* stream) {
    if (!x || !mean || !sd || check_dims(N, C, H, W)) return CNSN_E_BADARG;
    const Window win{h0, h1, w0, w1};
    if (check_window(win, H, W)) return CNSN_E_BADARG;
    if (reinterpret_cast<uintptr_t>(x) % esize(dtype)) return CNSN_E_ALIGN;
    return launch_instance_stats(x, dtype, (long long)N * C, H, W, win, eps, mean, sd, (cudaStream_t)stream);
}

extern "C" int cnsn_instance_stats_strided(const void* x, int dtype, int N, int C, int H, int W,
                                           long long sN, long long sC, long long sH, long long sW,
                                           int h0, int h1, int w0, int w1, float eps,
                                           float* mean, float* sd, void* stream) {
    if (!x || !mean || !sd || check_dims(N, C, H, W)) return CNSN_E_BADARG;
    const Window win{h0, h1, w0, w1};
    if (check_window(win, H, W)) return CNSN_E_BADARG;
    if (sN < 0 || sC < 0 || sH < 0 || sW < 0) return CNSN_E_BADARG;
    if (reinterpret_cast<uintptr_t>(x) % esize(dtype)) return CNSN_E_ALIGN;
    cudaStream_t s = (cudaStream_t)stream;
    if (sC == (long long)H * W && sH == W && sW == 1 && sN == (long long)C * H * W)      // dense NCHW: the vectorised kernel
        return launch_instance_stats(x, dtype, (long long)N * C, H, W, win, eps, mean, sd, s);
    if (sC == 1 && C >= 8) {
        const unsigned blocks = (unsigned)N * (unsigned)((C + 31) / 32);
        CNSN_DISPATCH_DTYPE(dtype, T, k_instance_stats_cl<T><<<blocks, kBlock, 0, s>>>((const T*)x, N, C, sN, sH, sW, win, eps, mean, sd));
    } else {
        const long long inst = (long long)N * C;
        const unsigned blocks = (unsigned)((inst + kWarpsPerBlock - 1) / kWarpsPerBlock);
        CNSN_DISPATCH_DTYPE(dtype, T, k_instance_stats_strided<T><<<blocks, kBlock, 0, s>>>((const T*)x, N, C, sN, sC, sH, sW, win, eps, mean, sd));
    }
    return launch_status();
}

extern "C" int cnsn_instance_stats_bwd(const void* x, void* dx, int dtype, int N, int C, int H, int W,
                                       int h0, int h1, int w0, int w1,
                                       const float* mean, const float* sd,
                                       const float* dmean, const float* dsd, void* stream) {
    if (!x || !dx || !mean || !sd || !dmean || !dsd || check_dims(N, C, H, W)) return CNSN_E_BADARG;
    const Window win{h0, h1, w0, w1};
    if (check_window(win, H, W)) return CNSN_E_BADARG;
    const int M = H * W;
    const long long inst = (long long)N * C;
    const bool full = win.full(H, W);
    const bool vec = vec_ok2(x, dx, dtype, M);
    const int tpi = team_for(M);
    cudaStream_t s = (cudaStream_t)stream;
    CNSN_DISPATCH_DTYPE(dtype, T, CNSN_DISPATCH_TEAM(tpi, TPI, CNSN_DISPATCH_BOOL(vec, VEC,
        k_instance_stats_bwd<T, TPI, VEC><<<grid_for(inst, TPI), kBlock, 0, s>>>(
            (const T*)x, (T*)dx, inst, W, M, win, full, mean, sd, dmean, dsd))));
    return launch_status();
}

extern "C" int cnsn_instance_affine(const void* x, void* out, int dtype, int N, int C, int H, int W,
                                    const float* scale, const float* shift, void* stream) {
    if (!x || !out || !scale || !shift || check_dims(N, C, H, W)) return CNSN_E_BADARG;
    const int M = H * W;
    const long long inst = (long long)N * C;
    const bool vec = vec_ok2(x, out, dtype, M);
    const int tpi = team_for(M);
    cudaStream_t s = (cudaStream_t)stream;
    CNSN_DISPATCH_DTYPE(dtype, T, CNSN_DISPATCH_TEAM(tpi, TPI, CNSN_DISPATCH_BOOL(vec, VEC,
        k_instance_affine<T, TPI, VEC><<<grid_for(inst, TPI), kBlock, 0, s>>>(
            (const T*)x, (T*)out, inst, M, scale, shift))));
    return launch_status();
}
