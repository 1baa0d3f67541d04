// crossnorm.cu -- 2-instance CrossNorm (cn_op_2ins_space_chan, models/cnsn.py:58-91, with
// instance_norm_mix :20-29) forward / backward: the C entry points, and the general path (two stream-ordered
// kernels per direction) used for channel permutation, odd plane sizes and channels too large for the
// shared-memory-resident dataflow kernel of crossnorm_flow.cu (the default).
//
//   forward : per-instance (content window, style window) statistics        (1 read of x)
//             y = ca*x + cb inside the content window, x outside             (read x, write y)
//               with A = sd_s[p(i),pi(c)] / sd_c[i,c],  ca = lam + (1-lam)*A,
//                    cb = (1-lam) * (mu_s[p(i),pi(c)] - mu_c[i,c]*A)
//   backward: S1 = sum_Wc d, S2 = sum_Wc d*xhat (d = (1-lam)*dy), scattered through the permutation
//             dx = piecewise affine in (dy, x) per window                     (read x, dy, write dx)
//
// Roofline: HBM; algorithmic 2*S forward, 3*S backward; this form moves 3*S / 5*S (for tensors
// that fit the 126 MB L2, e.g. BASELINE config 2 at 16 MiB, the second read is an L2 hit).
// Instance i needs the statistics of instance p(i), so a statistics phase over ALL instances
// precedes the apply phase: that is the kernel boundary here.
#include <stdlib.h>

#include "flow_common.cuh"

namespace cnsn {

template <typename T, int TPI, bool VEC>
__global__ void __launch_bounds__(kBlock)
k_cn_stats(const T* __restrict__ x, long long instances, int H, int W, Window cw, Window sw, float eps,
           float* __restrict__ mu_c, float* __restrict__ sd_c, float* __restrict__ mu_s,
           float* __restrict__ sd_s) {
    __shared__ Moments scratch[kWarpsPerBlock];
    const long long inst = Team<TPI>::instance();
    if (inst >= instances) return;
    const int M = H * W;
    const T* plane = x + inst * M;
    const bool cfull = cw.full(H, W), sfull = sw.full(H, W);
    Moments mc = instance_moments<T, TPI, VEC>(plane, W, M, cw, cfull);   // VEC implies cfull
    mc = Team<TPI>::all_merge(mc, scratch);
    const bool same = cw.h0 == sw.h0 && cw.h1 == sw.h1 && cw.w0 == sw.w0 && cw.w1 == sw.w1;
    Moments ms = mc;
    if (!same) {
        ms = (VEC && sfull) ? instance_moments<T, TPI, VEC>(plane, W, M, sw, true)
                            : instance_moments<T, TPI, false>(plane, W, M, sw, sfull);
        ms = Team<TPI>::all_merge(ms, scratch);
    }
    if (Team<TPI>::rank() == 0) {
        mu_c[inst] = mc.mean; sd_c[inst] = std_from(mc, eps);
        mu_s[inst] = ms.mean; sd_s[inst] = std_from(ms, eps);
    }
}

__device__ __forceinline__ long long style_source(long long inst, int C, const int* __restrict__ perm,
                                                  const int* __restrict__ cperm) {
    const int i = (int)(inst / C), c = (int)(inst - (long long)i * C);
    return (long long)perm[i] * C + (cperm ? cperm[c] : c);
}

template <typename T, int TPI, bool VEC>
__global__ void __launch_bounds__(kBlock)
k_cn_apply_fwd(const T* __restrict__ x, T* __restrict__ y, long long instances, int C, int H, int W,
               const int* __restrict__ perm, const int* __restrict__ cperm, Window cw, float lam,
               const float* __restrict__ mu_c, const float* __restrict__ sd_c,
               const float* __restrict__ mu_s, const float* __restrict__ sd_s) {
    const long long inst = Team<TPI>::instance();
    if (inst >= instances) return;
    const long long src = style_source(inst, C, perm, cperm);
    const float A = sd_s[src] / sd_c[inst];
    const float ca = lam + (1.f - lam) * A;
    const float cb = (1.f - lam) * (mu_s[src] - mu_c[inst] * A);
    const int M = H * W;
    if (cw.full(H, W)) {
        plane_map<T, TPI, VEC, false>(x + inst * M, nullptr, y + inst * M, M,
                                      [=](float xv, float, int) { return fmaf(ca, xv, cb); });
    } else {
        plane_map<T, TPI, VEC, false>(x + inst * M, nullptr, y + inst * M, M, [=](float xv, float, int i) {
            const int h = i / W, w = i - h * W;
            return cw.has(h, w) ? fmaf(ca, xv, cb) : xv;
        });
    }
}

// S1 = sum_Wc d ; S2 = sum_Wc d * (x - mu_c)/sd_c ; d = (1-lam)*dy.  Scattered to the style source.
template <typename T, int TPI, bool VEC>
__global__ void __launch_bounds__(kBlock)
k_cn_reduce_bwd(const T* __restrict__ x, const T* __restrict__ dy, long long instances, int C, int H,
                int W, const int* __restrict__ perm, const int* __restrict__ cperm, Window cw,
                float lam, const float* __restrict__ mu_c, const float* __restrict__ sd_c,
                float* __restrict__ s1, float* __restrict__ s2, float* __restrict__ dmu_s,
                float* __restrict__ dsd_s) {
    __shared__ float scratch[kWarpsPerBlock];
    const long long inst = Team<TPI>::instance();
    if (inst >= instances) return;
    const int M = H * W;
    const T* px = x + inst * M;
    const T* pd = dy + inst * M;
    const float mean = mu_c[inst];
    const int r = Team<TPI>::rank();
    float a = 0.f, t = 0.f;                  // a = sum dy*(x-mean), t = sum dy over the content window
    if (VEC) {                               // VEC implies the content window is the full plane
        constexpr int V = VecOf<T>::n;
        constexpr int U = 2;
        const uint4* vx = reinterpret_cast<const uint4*>(px);
        const uint4* vd = reinterpret_cast<const uint4*>(pd);
        const int nv = M / V;
        for (int i = r; i < nv; i += TPI * U) {
            uint4 rx[U], rd[U];
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (i + u * TPI < nv) { rx[u] = ldg_stream(vx + i + u * TPI); rd[u] = ldg_stream(vd + i + u * TPI); }
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (i + u * TPI < nv) {
                    float fx[V], fd[V];
                    unpack<T>(rx[u], fx);
                    unpack<T>(rd[u], fd);
#pragma unroll
                    for (int j = 0; j < V; ++j) { a = fmaf(fd[j], fx[j] - mean, a); t += fd[j]; }
                }
        }
    } else {
        const int cols = cw.cols(), area = cw.area();
        for (int i = r; i < area; i += TPI) {
            const int hh = i / cols, ww = i - hh * cols;
            const int o = (cw.h0 + hh) * W + cw.w0 + ww;
            const float d = to_f(pd[o]);
            a = fmaf(d, to_f(px[o]) - mean, a);
            t += d;
        }
    }
    a = Team<TPI>::all_sum(a, scratch);
    t = Team<TPI>::all_sum(t, scratch);
    if (r == 0) {
        const float S1 = (1.f - lam) * t;
        const float S2 = (1.f - lam) * a / sd_c[inst];
        s1[inst] = S1; s2[inst] = S2;
        const long long src = style_source(inst, C, perm, cperm);
        dmu_s[src] = S1; dsd_s[src] = S2;    // perm is a bijection: plain stores, no atomics
    }
}

template <typename T, int TPI, bool VEC>
__global__ void __launch_bounds__(kBlock)
k_cn_apply_bwd(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ dx,
               long long instances, int C, int H, int W, const int* __restrict__ perm,
               const int* __restrict__ cperm, Window cw, Window sw, float lam,
               const float* __restrict__ mu_c, const float* __restrict__ sd_c,
               const float* __restrict__ mu_s, const float* __restrict__ sd_s,
               const float* __restrict__ s1, const float* __restrict__ s2,
               const float* __restrict__ dmu_s, const float* __restrict__ dsd_s) {
    const long long inst = Team<TPI>::instance();
    if (inst >= instances) return;
    const long long src = style_source(inst, C, perm, cperm);
    const float Mc = (float)cw.area(), Ms = (float)sw.area();
    const float sdc = sd_c[inst];
    const float A = sd_s[src] / sdc;
    // inside the content window: dx = p*dy + q*x + r0
    const float p = lam + (1.f - lam) * A;
    const float q = -A * s2[inst] / ((Mc - 1.f) * sdc);
    const float r0 = -A * s1[inst] / Mc - q * mu_c[inst];
    // inside the style window (this instance as somebody's style source): dx += u*x + v
    const float u = dsd_s[inst] / ((Ms - 1.f) * sd_s[inst]);
    const float v = dmu_s[inst] / Ms - u * mu_s[inst];
    const int M = H * W;
    const bool cfull = cw.full(H, W), sfull = sw.full(H, W);
    if (cfull && sfull) {
        const float qq = q + u, rr = r0 + v;
        plane_map<T, TPI, VEC, true>(x + inst * M, dy + inst * M, dx + inst * M, M,
                                     [=](float xv, float dv, int) { return fmaf(p, dv, fmaf(qq, xv, rr)); });
    } else {
        plane_map<T, TPI, VEC, true>(x + inst * M, dy + inst * M, dx + inst * M, M, [=](float xv, float dv, int i) {
            const int h = i / W, w = i - h * W;
            float o = cw.has(h, w) ? fmaf(p, dv, fmaf(q, xv, r0)) : dv;
            if (sw.has(h, w)) o += fmaf(u, xv, v);
            return o;
        });
    }
}

namespace flow {
size_t crossnorm_scratch_floats(int N, int C);
int crossnorm_flow_fwd(const void* x, void* y, int dtype, int N, int C, int H, int W, const int* perm,
                       const Window& cw, const Window& sw, float lam, float eps,
                       float* mu_c, float* sd_c, float* mu_s, float* sd_s, float* scratch, cudaStream_t stream);
int crossnorm_flow_bwd(const void* x, const void* dy, void* dx, int dtype, int N, int C, int H, int W, const int* perm,
                       const Window& cw, const Window& sw, float lam,
                       const float* mu_c, const float* sd_c, const float* mu_s, const float* sd_s,
                       float* scratch, cudaStream_t stream);
}
// cnsn_tune("crossnorm_impl", 1) forces the two-kernel path (A/B measurements, tests).
static bool cn_use_flow() { return knobs().crossnorm_impl != 1; }

}  // namespace cnsn

using namespace cnsn;

// save: [mu_c | sd_c | mu_s | sd_s] (N*C each), then the dataflow kernel's polled words
extern "C" size_t cnsn_crossnorm_save_floats(int N, int C) { return 4 * (size_t)N * C + flow::crossnorm_scratch_floats(N, C); }
extern "C" size_t cnsn_crossnorm_workspace_floats(int N, int C) { return 4 * (size_t)N * C + 8; }

static int cn_check(const void* x, const void* y, int dtype, int N, int C, int H, int W, const int* perm,
                    const int* content, const int* style, Window& cw, Window& sw) {
    if (!x || !y || !perm || !content || !style || check_dims(N, C, H, W)) return CNSN_E_BADARG;
    if (dtype < CNSN_F32 || dtype > CNSN_F16) return CNSN_E_BADARG;
    cw = Window{content[0], content[1], content[2], content[3]};
    sw = Window{style[0], style[1], style[2], style[3]};
    if (check_window(cw, H, W) || check_window(sw, H, W)) return CNSN_E_BADARG;
    if (reinterpret_cast<uintptr_t>(x) % esize(dtype) || reinterpret_cast<uintptr_t>(y) % esize(dtype)) return CNSN_E_ALIGN;
    return 0;
}

extern "C" int cnsn_crossnorm_fwd(const void* x, void* y, int dtype, int N, int C, int H, int W,
                                  const int* perm, const int* chan_perm,
                                  const int* content, const int* style,
                                  float lam, float eps, float* save, void* stream) {
    Window cw, sw;
    int rc = cn_check(x, y, dtype, N, C, H, W, perm, content, style, cw, sw);
    if (rc) return rc;
    if (!save) return CNSN_E_BADARG;
    const int M = H * W;
    const long long inst = (long long)N * C;
    const size_t nc = (size_t)inst;
    float* mu_c = save; float* sd_c = save + nc; float* mu_s = save + 2 * nc; float* sd_s = save + 3 * nc;
    cudaStream_t s = (cudaStream_t)stream;
    if (!chan_perm && cn_use_flow()) {
        const int frc = flow::crossnorm_flow_fwd(x, y, dtype, N, C, H, W, perm, cw, sw, lam, eps, mu_c, sd_c, mu_s, sd_s,
                                                 save + 4 * nc, s);
        if (frc != -100) return frc;         // -100: shape not eligible, use the two-kernel path
    }
    {
        const bool vec = cw.full(H, W) && vec_ok(x, dtype, M);
        const int tpi = team_for(M);
        CNSN_DISPATCH_DTYPE(dtype, T, CNSN_DISPATCH_TEAM(tpi, TPI, CNSN_DISPATCH_BOOL(vec, VEC,
            k_cn_stats<T, TPI, VEC><<<grid_for(inst, TPI), kBlock, 0, s>>>(
                (const T*)x, inst, H, W, cw, sw, eps, mu_c, sd_c, mu_s, sd_s))));
        if ((rc = launch_status())) return rc;
    }
    const bool vec = vec_ok2(x, y, dtype, M);
    const int tpi = team_for(M);
    CNSN_DISPATCH_DTYPE(dtype, T, CNSN_DISPATCH_TEAM(tpi, TPI, CNSN_DISPATCH_BOOL(vec, VEC,
        k_cn_apply_fwd<T, TPI, VEC><<<grid_for(inst, TPI), kBlock, 0, s>>>(
            (const T*)x, (T*)y, inst, C, H, W, perm, chan_perm, cw, lam, mu_c, sd_c, mu_s, sd_s))));
    return launch_status();
}

extern "C" int cnsn_crossnorm_bwd(const void* x, const void* dy, void* dx, int dtype,
                                  int N, int C, int H, int W,
                                  const int* perm, const int* chan_perm,
                                  const int* content, const int* style,
                                  float lam, const float* save, float* workspace, void* stream) {
    Window cw, sw;
    int rc = cn_check(x, dx, dtype, N, C, H, W, perm, content, style, cw, sw);
    if (rc) return rc;
    if (!dy || !save || !workspace) return CNSN_E_BADARG;
    const int M = H * W;
    const long long inst = (long long)N * C;
    const size_t nc = (size_t)inst;
    const float* mu_c = save; const float* sd_c = save + nc; const float* mu_s = save + 2 * nc; const float* sd_s = save + 3 * nc;
    float* s1 = workspace; float* s2 = workspace + nc; float* dmu = workspace + 2 * nc; float* dsd = workspace + 3 * nc;
    cudaStream_t s = (cudaStream_t)stream;
    if (!chan_perm && cn_use_flow()) {
        const int frc = flow::crossnorm_flow_bwd(x, dy, dx, dtype, N, C, H, W, perm, cw, sw, lam, mu_c, sd_c, mu_s, sd_s,
                                                 workspace, s);
        if (frc != -100) return frc;
    }
    {
        const bool vec = cw.full(H, W) && vec_ok2(x, dy, dtype, M);
        const int tpi = team_for(cw.area());
        CNSN_DISPATCH_DTYPE(dtype, T, CNSN_DISPATCH_TEAM(tpi, TPI, CNSN_DISPATCH_BOOL(vec, VEC,
            k_cn_reduce_bwd<T, TPI, VEC><<<grid_for(inst, TPI), kBlock, 0, s>>>(
                (const T*)x, (const T*)dy, inst, C, H, W, perm, chan_perm, cw, lam, mu_c, sd_c, s1, s2, dmu, dsd))));
        if ((rc = launch_status())) return rc;
    }
    const bool vec = vec_ok2(x, dy, dtype, M) && aligned16(dx);
    const int tpi = team_for(M);
    CNSN_DISPATCH_DTYPE(dtype, T, CNSN_DISPATCH_TEAM(tpi, TPI, CNSN_DISPATCH_BOOL(vec, VEC,
        k_cn_apply_bwd<T, TPI, VEC><<<grid_for(inst, TPI), kBlock, 0, s>>>(
            (const T*)x, (const T*)dy, (T*)dx, inst, C, H, W, perm, chan_perm, cw, sw, lam,
            mu_c, sd_c, mu_s, sd_s, s1, s2, dmu, dsd))));
    return launch_status();
}
