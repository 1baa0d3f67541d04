// selfnorm.cu -- SelfNorm forward / backward (models/cnsn.py:113-150 and its autograd), v1:
// three stream-ordered kernels per direction.
//
//   forward : instance stats (1 read of x)  ->  gate (O(N*C))  ->  y = x*g (+ mu*(f-g))   (read x, write y)
//   backward: per-instance sum(dy*x), sum(dy) (read x, dy) -> gate backward (O(N*C)) -> dx = g*dy + b*x + c
//
// Roofline: HBM.  Algorithmic bytes: forward 2*S, backward 3*S (S = N*C*H*W*sizeof(T)); this
// three-kernel form moves 3*S and 5*S, the fused persistent kernels (selfnorm_fused.cu) remove the
// re-reads.  The gate couples all N instances of a channel (BatchNorm1d over the batch, :121,:138),
// which is why a reduction phase must complete for the whole channel before any element of it
// can be written.
#include <math.h>
#include <stdlib.h>

#include <algorithm>

#include "flow_common.cuh"
#include "selfnorm_gate.cuh"

namespace cnsn {

// y = x*g                       (one gate)
// y = x*g + mu*(f-g)            (is_two, models/cnsn.py:148)
template <typename T, int TPI, bool VEC>
__global__ void __launch_bounds__(kBlock)
k_sn_apply_fwd(const T* __restrict__ x, T* __restrict__ y, long long instances, int M,
               const float* __restrict__ gate, const float* __restrict__ fgate,
               const float* __restrict__ mu, int relu) {
    const long long inst = Team<TPI>::instance();
    if (inst >= instances) return;
    const float a = gate[inst];
    const float b = fgate ? mu[inst] * (fgate[inst] - a) : 0.f;
    const float lo = relu ? 0.f : -INFINITY;
    plane_map<T, TPI, VEC, false>(x + inst * M, nullptr, y + inst * M, M,
                                  [=](float xv, float, int) { return fmaxf(fmaf(a, xv, b), lo); });
}

// z = x + res (the residual add in front of a pos='post' SelfNorm site), element type T
template <typename T, bool VEC>
__global__ void __launch_bounds__(kBlock)
k_add(const T* __restrict__ x, const T* __restrict__ res, T* __restrict__ z, long long total) {
    constexpr int V = VecOf<T>::n;
    const long long stride = (long long)gridDim.x * kBlock;
    if (VEC) {
        const uint4* px = reinterpret_cast<const uint4*>(x);
        const uint4* pr = reinterpret_cast<const uint4*>(res);
        uint4* pz = reinterpret_cast<uint4*>(z);
        const long long nv = total / V;
        for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < nv; i += stride) {
            float a[V], b[V];
            unpack<T>(ldg_stream(px + i), a);
            unpack<T>(ldg_stream(pr + i), b);
#pragma unroll
            for (int e = 0; e < V; ++e) a[e] += b[e];
            pz[i] = pack<T>(a);
        }
        for (long long i = nv * V + (long long)blockIdx.x * kBlock + threadIdx.x; i < total; i += stride)
            z[i] = from_f<T>(to_f(x[i]) + to_f(res[i]));
    } else {
        for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < total; i += stride)
            z[i] = from_f<T>(to_f(x[i]) + to_f(res[i]));
    }
}

// Per instance: sxy = sum dy*x  (CENTER: sum dy*(x-mu), for the is_two form) and t = sum dy.
template <typename T, int TPI, bool VEC, bool CENTER>
__global__ void __launch_bounds__(kBlock)
k_sn_reduce_bwd(const T* __restrict__ x, const T* __restrict__ dy, long long instances, int M,
                const float* __restrict__ mu, float* __restrict__ sxy, float* __restrict__ st, int relu = 0) {
    __shared__ float scratch[kWarpsPerBlock];
    const long long inst = Team<TPI>::instance();
    if (inst >= instances) return;
    const T* px = x + inst * M;
    const T* pd = dy + inst * M;
    const float mean = CENTER ? mu[inst] : 0.f;
    const int r = Team<TPI>::rank();
    float a = 0.f, t = 0.f;
    if (VEC) {
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
                    for (int j = 0; j < V; ++j) {
                        const float d = (relu && !(fx[j] > 0.f)) ? 0.f : fd[j];   // relu(g*x): masked where x <= 0
                        a = fmaf(d, fx[j] - mean, a); t += d;
                    }
                }
        }
    } else {
#pragma unroll 4
        for (int i = r; i < M; i += TPI) {
            const float xv = to_f(px[i]);
            const float d = (relu && !(xv > 0.f)) ? 0.f : to_f(pd[i]);
            a = fmaf(d, xv - mean, a);
            t += d;
        }
    }
    a = Team<TPI>::all_sum(a, scratch);
    t = Team<TPI>::all_sum(t, scratch);
    if (r == 0) { sxy[inst] = a; st[inst] = t; }
}

// dx = g*dy + cb*x + cc
template <typename T, int TPI, bool VEC>
__global__ void __launch_bounds__(kBlock)
k_sn_apply_bwd(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ dx,
               long long instances, int M, const float* __restrict__ gate,
               const float* __restrict__ cb, const float* __restrict__ cc, int relu) {
    const long long inst = Team<TPI>::instance();
    if (inst >= instances) return;
    const float g = gate[inst], b = cb[inst], c = cc[inst];
    plane_map<T, TPI, VEC, true>(x + inst * M, dy + inst * M, dx + inst * M, M, [=](float xv, float dv, int) {
        const float d = (relu && !(xv > 0.f)) ? 0.f : dv;
        return fmaf(g, d, fmaf(b, xv, c));
    });
}

namespace flow {
size_t scratch_floats(int N, int C);
int selfnorm_flow_fwd(const void* x, const void* res, void* z, void* y, int relu, int dtype, int N, int C, int H, int W,
                      const cnsn_gate_params* g, int training, float momentum, float bn_eps, float eps,
                      float* mu, float* sd, float* gate, float* shat, float* r, float* scratch,
                      cudaStream_t stream);
int selfnorm_flow_bwd(const void* x, const void* dy, void* dx, int relu, int dtype, int N, int C, int H, int W,
                      const cnsn_gate_params* g, int training,
                      float* mu, float* sd, float* gate, float* shat, float* r,
                      const cnsn_gate_grads* dg, float* scratch, cudaStream_t stream);
}
// cnsn_tune("selfnorm_impl", v) selects the path for A/B measurements and tests: 1 = three kernels, 3 = the dataflow
// kernels only, 0 = the default dispatch (dataflow kernels, three-kernel path for shapes they do not take).
enum { kImplAuto = 0, kImplV1 = 1, kImplFlow = 3 };
static int impl_choice() { return knobs().selfnorm_impl; }

static bool gate_ok(const cnsn_gate_params* p) { return p && p->w && p->gamma && p->beta && p->run_mean && p->run_var; }

}  // namespace cnsn

using namespace cnsn;

extern "C" size_t cnsn_selfnorm_save_floats(int N, int C, int is_two) {
    return SaveLayout(N, C, is_two != 0).total;
}
extern "C" size_t cnsn_selfnorm_workspace_floats(int N, int C, int is_two) {
    (void)is_two;
    return 4 * (size_t)N * C + 36 * (size_t)C + 8;   // sxy | st | cb | cc   (fused / flow paths: [C][N] published words, ...)
}

// y = [relu]( SelfNorm(x [+ res]) ); with res the sum is written to z (what backward needs).
static int selfnorm_fwd_impl(const void* x, const void* res, void* z, void* y, int relu, int dtype, int N, int C, int H, int W,
                             const cnsn_gate_params* g, const cnsn_gate_params* f,
                             int training, float momentum, float bn_eps, float eps,
                             float* save, void* stream) {
    if (!x || !y || !save || check_dims(N, C, H, W) || !gate_ok(g) || (f && !gate_ok(f))) return CNSN_E_BADARG;
    if (res && (!z || f)) return CNSN_E_BADARG;      // the fused add needs somewhere to put the sum; not with is_two
    if (dtype < CNSN_F32 || dtype > CNSN_F16) return CNSN_E_BADARG;
    if (reinterpret_cast<uintptr_t>(x) % esize(dtype) || reinterpret_cast<uintptr_t>(y) % esize(dtype)) return CNSN_E_ALIGN;
    if (res && (reinterpret_cast<uintptr_t>(res) % esize(dtype) || reinterpret_cast<uintptr_t>(z) % esize(dtype))) return CNSN_E_ALIGN;
    if (training && N < 2) return CNSN_E_BATCH1;    // BatchNorm1d raises ValueError in the reference
    const bool two = f != nullptr;
    const SaveLayout L(N, C, two);
    const int M = H * W;
    const long long inst = (long long)N * C;
    cudaStream_t s = (cudaStream_t)stream;
    // Default: the ticket-ordered dataflow kernel (measured faster than the three-kernel path and at least as
    // fast as the persistent kernels on every shape of the r01 sweep, profiles/README.md).
    if (!two && (impl_choice() == kImplFlow || impl_choice() == kImplAuto)) {
        const int frc = flow::selfnorm_flow_fwd(x, res, z, y, relu, dtype, N, C, H, W, g, training, momentum, bn_eps, eps,
                                                save + L.mu, save + L.sd, save + L.g, save + L.shat_g,
                                                save + L.r_g, save + L.scratch, s);
        if (frc != -100) return frc;
    }
    if (res) {                                       // general path: materialise the sum, then proceed on it
        const long long total = (long long)N * C * M;
        const bool avec = aligned16(x) && aligned16(res) && aligned16(z);
        const unsigned blocks = (unsigned)std::min<long long>((total / 8 + kBlock - 1) / kBlock + 1, 148 * 16);
        CNSN_DISPATCH_DTYPE(dtype, T, CNSN_DISPATCH_BOOL(avec, VEC,
            k_add<T, VEC><<<blocks, kBlock, 0, s>>>((const T*)x, (const T*)res, (T*)z, total)));
        const int arc = launch_status();
        if (arc) return arc;
        x = z;
    }
    const Window full{0, H, 0, W};
    int rc = launch_instance_stats(x, dtype, inst, H, W, full, eps, save + L.mu, save + L.sd, s);
    if (rc) return rc;
    GateFwd a{g->w, g->gamma, g->beta, g->run_mean, g->run_var, g->nbt, save + L.g, save + L.shat_g, save + L.r_g};
    GateFwd b = a;
    if (two) b = GateFwd{f->w, f->gamma, f->beta, f->run_mean, f->run_var, f->nbt, save + L.f, save + L.shat_f, save + L.r_f};
    k_sn_gate_fwd<<<dim3(C, two ? 2 : 1), kGateThreads, 0, s>>>(
        save + L.mu, save + L.sd, a, b, N, C, training, momentum, bn_eps);
    if ((rc = launch_status())) return rc;
    const bool vec = vec_ok2(x, y, dtype, M);
    const int tpi = team_for(M);
    CNSN_DISPATCH_DTYPE(dtype, T, CNSN_DISPATCH_TEAM(tpi, TPI, CNSN_DISPATCH_BOOL(vec, VEC,
        k_sn_apply_fwd<T, TPI, VEC><<<grid_for(inst, TPI), kBlock, 0, s>>>(
            (const T*)x, (T*)y, inst, M, save + L.g, two ? save + L.f : nullptr, save + L.mu, relu))));
    return launch_status();
}

extern "C" int cnsn_selfnorm_fwd(const void* x, void* y, int dtype, int N, int C, int H, int W,
                                 const cnsn_gate_params* g, const cnsn_gate_params* f,
                                 int training, float momentum, float bn_eps, float eps,
                                 float* save, void* stream) {
    return selfnorm_fwd_impl(x, nullptr, nullptr, y, 0, dtype, N, C, H, W, g, f, training, momentum, bn_eps, eps, save, stream);
}

extern "C" int cnsn_selfnorm_block_fwd(const void* x, const void* res, void* z, void* y, int relu, int dtype,
                                       int N, int C, int H, int W, const cnsn_gate_params* g,
                                       int training, float momentum, float bn_eps, float eps,
                                       float* save, void* stream) {
    return selfnorm_fwd_impl(x, res, z, y, relu ? 1 : 0, dtype, N, C, H, W, g, nullptr, training, momentum, bn_eps, eps, save, stream);
}

static int selfnorm_bwd_impl(const void* x, const void* dy, void* dx, int relu, int dtype,
                             int N, int C, int H, int W,
                             const cnsn_gate_params* g, const cnsn_gate_params* f,
                             int training, const float* save,
                             const cnsn_gate_grads* dg, const cnsn_gate_grads* df,
                             float* workspace, void* stream) {
    if (!x || !dy || !dx || !save || !workspace || check_dims(N, C, H, W)) return CNSN_E_BADARG;
    if (!g || !g->w || !g->gamma || !dg || !dg->dw || !dg->dgamma || !dg->dbeta) return CNSN_E_BADARG;
    const bool two = f != nullptr;
    if (two && (!f->w || !f->gamma || !df || !df->dw || !df->dgamma || !df->dbeta)) return CNSN_E_BADARG;
    if (dtype < CNSN_F32 || dtype > CNSN_F16) return CNSN_E_BADARG;
    const SaveLayout L(N, C, two);
    const int M = H * W;
    const long long inst = (long long)N * C;
    const size_t nc = (size_t)inst;
    float* sxy = workspace; float* st = workspace + nc; float* cb = workspace + 2 * nc; float* cc = workspace + 3 * nc;
    cudaStream_t s = (cudaStream_t)stream;
    // Default: the dataflow kernels (3*S of HBM traffic).
    if (!two && (impl_choice() == kImplFlow || impl_choice() == kImplAuto)) {
        float* sv = const_cast<float*>(save);
        const int frc = flow::selfnorm_flow_bwd(x, dy, dx, relu, dtype, N, C, H, W, g, training, sv + L.mu, sv + L.sd,
                                                sv + L.g, sv + L.shat_g, sv + L.r_g, dg, workspace, s);
        if (frc != -100) return frc;
    }
    const bool vec = vec_ok2(x, dy, dtype, M) && aligned16(dx);
    const int tpi = team_for(M);
    CNSN_DISPATCH_DTYPE(dtype, T, CNSN_DISPATCH_TEAM(tpi, TPI, CNSN_DISPATCH_BOOL(vec, VEC, CNSN_DISPATCH_BOOL(two, CENTER,
        k_sn_reduce_bwd<T, TPI, VEC, CENTER><<<grid_for(inst, TPI), kBlock, 0, s>>>(
            (const T*)x, (const T*)dy, inst, M, save + L.mu, sxy, st, relu)))));
    int rc = launch_status();
    if (rc) return rc;
    GateBwd a{g->w, g->gamma, save + L.g, save + L.shat_g, save + L.r_g, dg->dw, dg->dgamma, dg->dbeta};
    GateBwd b = a;
    if (two) b = GateBwd{f->w, f->gamma, save + L.f, save + L.shat_f, save + L.r_f, df->dw, df->dgamma, df->dbeta};
    k_sn_gate_bwd<<<C, kGateThreads, 0, s>>>(
        save + L.mu, save + L.sd, sxy, st, a, b, two ? 1 : 0, N, C, M, training, cb, cc);
    if ((rc = launch_status())) return rc;
    CNSN_DISPATCH_DTYPE(dtype, T, CNSN_DISPATCH_TEAM(tpi, TPI, CNSN_DISPATCH_BOOL(vec, VEC,
        k_sn_apply_bwd<T, TPI, VEC><<<grid_for(inst, TPI), kBlock, 0, s>>>(
            (const T*)x, (const T*)dy, (T*)dx, inst, M, save + L.g, cb, cc, relu))));
    return launch_status();
}

extern "C" int cnsn_selfnorm_bwd(const void* x, const void* dy, void* dx, int dtype,
                                 int N, int C, int H, int W,
                                 const cnsn_gate_params* g, const cnsn_gate_params* f,
                                 int training, const float* save,
                                 const cnsn_gate_grads* dg, const cnsn_gate_grads* df,
                                 float* workspace, void* stream) {
    return selfnorm_bwd_impl(x, dy, dx, 0, dtype, N, C, H, W, g, f, training, save, dg, df, workspace, stream);
}

extern "C" int cnsn_selfnorm_block_bwd(const void* z, const void* dy, void* dz, int relu, int dtype,
                                       int N, int C, int H, int W, const cnsn_gate_params* g,
                                       int training, const float* save, const cnsn_gate_grads* dg,
                                       float* workspace, void* stream) {
    return selfnorm_bwd_impl(z, dy, dz, relu ? 1 : 0, dtype, N, C, H, W, g, nullptr, training, save, dg, nullptr, workspace, stream);
}

// Per-instance sums needed by the backward of cnsn_instance_affine: sxy = sum dy*x, st = sum dy.
extern "C" int cnsn_instance_dot(const void* x, const void* dy, int dtype, int N, int C, int H, int W,
                                 float* sxy, float* st, void* stream) {
    if (!x || !dy || !sxy || !st || check_dims(N, C, H, W)) return CNSN_E_BADARG;
    if (dtype < CNSN_F32 || dtype > CNSN_F16) return CNSN_E_BADARG;
    const int M = H * W;
    const long long inst = (long long)N * C;
    const bool vec = vec_ok2(x, dy, dtype, M);
    const int tpi = team_for(M);
    cudaStream_t s = (cudaStream_t)stream;
    CNSN_DISPATCH_DTYPE(dtype, T, CNSN_DISPATCH_TEAM(tpi, TPI, CNSN_DISPATCH_BOOL(vec, VEC,
        k_sn_reduce_bwd<T, TPI, VEC, false><<<grid_for(inst, TPI), kBlock, 0, s>>>(
            (const T*)x, (const T*)dy, inst, M, nullptr, sxy, st))));
    return launch_status();
}
