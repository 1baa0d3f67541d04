// selfnorm_nhwc.cu -- SelfNorm block (optional residual add in front, optional ReLU behind; models/cnsn.py:113-150,
// wideresnet_cnsn.py:93-96, resnet_cnsn.py:117-122) on CHANNELS-LAST tensors: logical shape (N, C, H, W), memory order
// N, H, W, C (torch.channels_last).
//
// Why: cuDNN's tensor-core convolutions work in NHWC.  With NCHW activations it converts around every convolution --
// 35 % of a WideResNet-40-2 step on B200 (nchwToNhwcKernel 26 %, nhwcToNchwKernel 9 %; gpurun_out/r3m_wrnprof.log), and the
// plain network runs 31.8 -> 10.1 ms per step once it is channels_last (gpurun_out/r3n_cl_wrn.log).  A channels_last
// network needs its SelfNorm sites in the same layout, or every site pays the two conversions back.
//
// Layout consequences.  In NHWC a sample is one contiguous [HW][C] tile, an instance (n, c) is a COLUMN of it.  A
// sample tile is 32-128 KB at the WideResNet sites; the batch coupling of the gate (BatchNorm1d over N) still needs
// every sample's statistics before any element can be written, and N = 512 tiles do not fit on chip, so this path is
// the three-step form: statistics -> gate on [N][C] -> apply, the second read served by L2 for the activations of a
// CIFAR-size network (16-64 MB per tensor against 126 MB of L2).  Algorithmic bytes: forward 2 S (4 S with the add:
// x, res in; z, y out), backward 3 S; moved: one more read of z (forward) / z and dy (backward), from L2 when they fit.
//
// Work split.  CTA (n, s) owns slab s of sample n: `rows` consecutive pixels x all C channels, contiguous in memory.
// Thread t owns the 16-byte channel vector cg = t mod CG (CG = C * sizeof(T) / 16 vectors per pixel) of the pixels
// rl, rl + RL, ... (rl = t / CG, RL = 256 / CG): every warp access is a contiguous run of whole pixels, a thread's
// channels never change, so per-channel coefficients live in registers.  Column sums: registers -> shared memory
// [RL][C] -> one thread per channel.  The slabs of a sample meet in its last CTA (one counter per sample): Chan merge
// of the slab-wise exact two-pass (mean, M2) forward, plain ordered sum backward -- deterministic.
#include <stdio.h>

#include <algorithm>

#include "bn_nhwc.cuh"
#include "flow_common.cuh"
#include "selfnorm_gate.cuh"

namespace cnsn {
namespace nhwc {

constexpr int kT = 256;

struct Geom {
    int N, C, HW;
    int CG;            // 16-byte vectors per pixel
    int CGB;           // vectors per pixel handled by one CTA (channel block, blockIdx.y); CG % CGB == 0
    int RL;            // row lanes = kT / CGB
    int S;             // slabs per sample
    int rows;          // pixels per slab (the last slab of a sample may be shorter)
};

// Sum over the row lanes: v[e] of every thread -> column totals in s_col[C].  red: [RL][C] floats.
template <int V>
__device__ __forceinline__ void column_sums(const float (&v)[V], float* red, float* s_col, const Geom& g, int cg, int rl) {
    const int W = g.CGB * V;                                 // channels of this CTA
#pragma unroll
    for (int e = 0; e < V; ++e) red[rl * W + cg * V + e] = v[e];
    __syncthreads();
    for (int c = threadIdx.x; c < W; c += kT) {              // a serial chain of up to 128 terms: in double (W threads, once per pass)
        double s = 0.0;
        for (int k = 0; k < g.RL; ++k) s += (double)red[k * W + c];
        s_col[c] = (float)s;
    }
    __syncthreads();
}

// forward statistics of slab (n, s): exact two-pass per slab, Chan merge across the slabs of the sample.
// ADD: z = x + res is formed on the way (rounded to T like torch.add's output), written out, and the statistics are z's.
// AFF (with ADD; the fused bottleneck tail): x is the raw output of conv3 and is first sent through bn3's batch-norm map
// y3 = scale * x + shift (coef[c], rounded to T as the unfused batch norm would have stored it), z = y3 + res.
template <typename T, bool ADD, bool AFF = false>
__global__ void __launch_bounds__(kT) k_nhwc_stats(const T* __restrict__ x, const T* __restrict__ res, T* __restrict__ z,
                                                   const Geom g, float eps, float2* __restrict__ part,
                                                   unsigned* __restrict__ cnt, float* __restrict__ mu, float* __restrict__ sd,
                                                   const float2* __restrict__ coef = nullptr) {
    static_assert(ADD || !AFF, "the batch-norm map comes with the residual add");
    constexpr int V = VecOf<T>::n;
    extern __shared__ float sm[];                            // red [RL][W] | col [W] | mean [W]
    const int W = g.CGB * V, c0 = blockIdx.y * W;            // this CTA's channels: c0 .. c0 + W - 1
    float* red = sm;
    float* s_col = sm + g.RL * W;
    float* s_mean = s_col + W;
    __shared__ unsigned s_last;
    const int n = blockIdx.x / g.S, s = blockIdx.x - n * g.S;
    const int cg = threadIdx.x % g.CGB, rl = threadIdx.x / g.CGB;
    const int row0 = s * g.rows, nrows = min(g.rows, g.HW - row0);
    const size_t vbase = ((size_t)n * g.HW + row0) * g.CG + (size_t)blockIdx.y * g.CGB + cg;
    const uint4* vx = reinterpret_cast<const uint4*>(x) + vbase;
    const uint4* vr = reinterpret_cast<const uint4*>(res) + vbase;
    uint4* vz = reinterpret_cast<uint4*>(z) + vbase;
    float acc[V], fs[V], fb[V];
#pragma unroll
    for (int e = 0; e < V; ++e) {
        acc[e] = 0.f;
        const float2 f = AFF ? coef[c0 + cg * V + e] : make_float2(1.f, 0.f);
        fs[e] = f.x; fb[e] = f.y;
    }
#pragma unroll 4
    for (int r = rl; r < nrows; r += g.RL) {
        float a[V];
        if (ADD) {
            float b[V];
            unpack<T>(AFF ? __ldg(vx + (size_t)r * g.CG) : ldg_stream(vx + (size_t)r * g.CG), a);   // AFF: the backward reads x again
            unpack<T>(ldg_stream(vr + (size_t)r * g.CG), b);
            if (AFF) {
#pragma unroll
                for (int e = 0; e < V; ++e) a[e] = fmaf(fs[e], a[e], fb[e]);
                unpack<T>(pack<T>(a), a);                    // y3 as the element type holds it
            }
#pragma unroll
            for (int e = 0; e < V; ++e) a[e] += b[e];
            const uint4 q = pack<T>(a);
            vz[(size_t)r * g.CG] = q;                         // default policy: the second pass and the apply kernel re-read it
            unpack<T>(q, a);                                 // the statistics are those of the rounded sum
        } else {
            unpack<T>(__ldg(vx + (size_t)r * g.CG), a);
        }
#pragma unroll
        for (int e = 0; e < V; ++e) acc[e] += a[e];
    }
    column_sums<V>(acc, red, s_col, g, cg, rl);
    for (int c = threadIdx.x; c < W; c += kT) s_mean[c] = s_col[c] / nrows;
    __syncthreads();
    float m[V];
#pragma unroll
    for (int e = 0; e < V; ++e) { m[e] = s_mean[cg * V + e]; acc[e] = 0.f; }
    const uint4* v2 = ADD ? reinterpret_cast<const uint4*>(z) + vbase : vx;
#pragma unroll 4
    for (int r = rl; r < nrows; r += g.RL) {
        float a[V];
        unpack<T>(ADD ? __ldcg(v2 + (size_t)r * g.CG) : __ldg(v2 + (size_t)r * g.CG), a);
#pragma unroll
        for (int e = 0; e < V; ++e) { const float d = a[e] - m[e]; acc[e] = fmaf(d, d, acc[e]); }
    }
    column_sums<V>(acc, red, s_col, g, cg, rl);
    const float M = (float)g.HW;
    if (g.S == 1) {
        for (int c = threadIdx.x; c < W; c += kT) {
            mu[(size_t)n * g.C + c0 + c] = s_mean[c];
            sd[(size_t)n * g.C + c0 + c] = sqrtf(s_col[c] / (M - 1.f) + eps);
        }
        return;
    }
    float2* mine = part + ((size_t)n * g.S + s) * g.C + c0;
    for (int c = threadIdx.x; c < W; c += kT) mine[c] = make_float2(s_mean[c], s_col[c]);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(cnt + (size_t)n * gridDim.y + blockIdx.y, 1u) == (unsigned)(g.S - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const float2* all = part + (size_t)n * g.S * g.C + c0;
    for (int c = threadIdx.x; c < W; c += kT) {              // Chan merge in slab order, in double (O(N C S) work in all)
        double cnt_a = 0.0, mean = 0.0, m2 = 0.0;
        for (int k = 0; k < g.S; ++k) {
            const float2 p = __ldcg(all + (size_t)k * g.C + c);
            const double cnt_b = (double)min(g.rows, g.HW - k * g.rows);
            const double tot = cnt_a + cnt_b, d = (double)p.x - mean;
            mean += d * (cnt_b / tot);
            m2 += (double)p.y + d * d * (cnt_a * cnt_b / tot);
            cnt_a = tot;
        }
        mu[(size_t)n * g.C + c0 + c] = (float)mean;
        sd[(size_t)n * g.C + c0 + c] = sqrtf((float)(m2 / ((double)M - 1.0)) + eps);
    }
}

// backward reduction of slab (n, s): sxy[n][c] = sum d * z with d = dy masked where z <= 0 (ReLU behind the site)
template <typename T>
__global__ void __launch_bounds__(kT) k_nhwc_reduce_bwd(const T* __restrict__ z, const T* __restrict__ dy, const Geom g, int relu,
                                                        float* __restrict__ part, unsigned* __restrict__ cnt,
                                                        float* __restrict__ sxy) {
    constexpr int V = VecOf<T>::n;
    extern __shared__ float sm[];
    const int W = g.CGB * V, c0 = blockIdx.y * W;
    float* red = sm;
    float* s_col = sm + g.RL * W;
    __shared__ unsigned s_last;
    const int n = blockIdx.x / g.S, s = blockIdx.x - n * g.S;
    const int cg = threadIdx.x % g.CGB, rl = threadIdx.x / g.CGB;
    const int row0 = s * g.rows, nrows = min(g.rows, g.HW - row0);
    const size_t vbase = ((size_t)n * g.HW + row0) * g.CG + (size_t)blockIdx.y * g.CGB + cg;
    const uint4* vz = reinterpret_cast<const uint4*>(z) + vbase;
    const uint4* vd = reinterpret_cast<const uint4*>(dy) + vbase;
    float acc[V];
#pragma unroll
    for (int e = 0; e < V; ++e) acc[e] = 0.f;
#pragma unroll 4
    for (int r = rl; r < nrows; r += g.RL) {
        float a[V], d[V];
        unpack<T>(__ldg(vz + (size_t)r * g.CG), a);
        unpack<T>(__ldg(vd + (size_t)r * g.CG), d);
#pragma unroll
        for (int e = 0; e < V; ++e) acc[e] = fmaf((relu && !(a[e] > 0.f)) ? 0.f : d[e], a[e], acc[e]);
    }
    column_sums<V>(acc, red, s_col, g, cg, rl);
    if (g.S == 1) {
        for (int c = threadIdx.x; c < W; c += kT) sxy[(size_t)n * g.C + c0 + c] = s_col[c];
        return;
    }
    float* mine = part + ((size_t)n * g.S + s) * g.C + c0;
    for (int c = threadIdx.x; c < W; c += kT) mine[c] = s_col[c];
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(cnt + (size_t)n * gridDim.y + blockIdx.y, 1u) == (unsigned)(g.S - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const float* all = part + (size_t)n * g.S * g.C + c0;
    for (int c = threadIdx.x; c < W; c += kT) {
        double t = 0.0;
        for (int k = 0; k < g.S; ++k) t += (double)__ldcg(all + (size_t)k * g.C + c);
        sxy[(size_t)n * g.C + c0 + c] = (float)t;
    }
}

// forward: y = relu?(z * gate[n][c]);  backward: dx = gate * d + cb * z + cc, d = dy masked where z <= 0
template <typename T, bool BWD>
__global__ void __launch_bounds__(kT, 4) k_nhwc_apply(const T* __restrict__ z, const T* __restrict__ dy, T* __restrict__ out, const Geom g,
                                                   int relu, const float* __restrict__ gate, const float* __restrict__ cb,
                                                   const float* __restrict__ cc) {
    constexpr int V = VecOf<T>::n;
    const unsigned bx = gridDim.x - 1 - blockIdx.x;          // reverse launch order: the statistics / reduction kernel in front read
    const int n = bx / g.S, s = bx - n * g.S;                // front to back, the last slabs are the ones still in L2
    const int cg = threadIdx.x % g.CGB, rl = threadIdx.x / g.CGB;
    const int row0 = s * g.rows, nrows = min(g.rows, g.HW - row0);
    const size_t vbase = ((size_t)n * g.HW + row0) * g.CG + (size_t)blockIdx.y * g.CGB + cg;
    const uint4* vz = reinterpret_cast<const uint4*>(z) + vbase;
    const uint4* vd = reinterpret_cast<const uint4*>(dy) + vbase;
    uint4* vo = reinterpret_cast<uint4*>(out) + vbase;
    float kg[V], kb[V], kc[V];
#pragma unroll
    for (int e = 0; e < V; ++e) {
        const size_t i = (size_t)n * g.C + (size_t)blockIdx.y * g.CGB * V + cg * V + e;
        kg[e] = gate[i];
        kb[e] = BWD ? cb[i] : 0.f;
        kc[e] = BWD ? cc[i] : 0.f;
    }
#pragma unroll 4
    for (int r = rl; r < nrows; r += g.RL) {
        float a[V], d[V], o[V];
        unpack<T>(ldg_stream(vz + (size_t)r * g.CG), a);
        if (BWD) unpack<T>(ldg_stream(vd + (size_t)r * g.CG), d);
#pragma unroll
        for (int e = 0; e < V; ++e) {
            if (BWD) {
                const float dd = (relu && !(a[e] > 0.f)) ? 0.f : d[e];
                o[e] = fmaf(kg[e], dd, fmaf(kb[e], a[e], kc[e]));
            } else {
                const float y = kg[e] * a[e];
                o[e] = relu ? fmaxf(y, 0.f) : y;
            }
        }
        vo[(size_t)r * g.CG] = pack<T>(o);
    }
}

// The fused bottleneck tail, backward middle step, in the BATCH NORM's geometry (row chunks over the whole tensor): forms
// dz = gate * d + cb * z + cc (SelfNorm block backward, d = dy masked where z <= 0), writes it -- it is also the gradient of
// the residual branch --, and on the way accumulates bn3's reduction (sum dz, sum dz * (c - mean)) * (1, rstd) per chunk and
// column: the unfused sequence would write dz, then read it and c again for that.
template <typename T>
__global__ void __launch_bounds__(bnl::kT, 4) k_tail_mid_bwd(const T* __restrict__ z, const T* __restrict__ dy, const T* __restrict__ c,
                                                             T* __restrict__ dz, const bnl::Geom g, int HW, int relu,
                                                             const float* __restrict__ gate, const float* __restrict__ cb,
                                                             const float* __restrict__ cc, const float* __restrict__ bn_mean,
                                                             const float* __restrict__ bn_rstd, float2* __restrict__ part) {
    constexpr int V = VecOf<T>::n;
    extern __shared__ float sm[];                            // a [RL][W] | b [RL][W]
    const int W = g.CGB * V;
    float* s_a = sm;
    float* s_b = sm + g.RL * W;
    const int cg = threadIdx.x % g.CGB, rl = threadIdx.x / g.CGB;
    const unsigned bx = gridDim.x - 1 - blockIdx.x;          // reverse launch order (see k_nhwc_apply)
    const long long row0 = (long long)bx * g.rows;
    const long long nrows = min(g.rows, g.R - row0);
    const size_t vb = (size_t)row0 * g.CG + (size_t)blockIdx.y * g.CGB + cg;
    const uint4* vz = reinterpret_cast<const uint4*>(z) + vb;
    const uint4* vd = reinterpret_cast<const uint4*>(dy) + vb;
    const uint4* vc = reinterpret_cast<const uint4*>(c) + vb;
    uint4* vo = reinterpret_cast<uint4*>(dz) + vb;
    const int c0 = blockIdx.y * W + cg * V;
    float mean[V], a[V], b[V], kg[V], kb[V], kc[V];
#pragma unroll
    for (int e = 0; e < V; ++e) { mean[e] = bn_mean[c0 + e]; a[e] = 0.f; b[e] = 0.f; kg[e] = kb[e] = kc[e] = 0.f; }
    long long n = (row0 + rl) / HW;                          // sample of this thread's current row
    int rem = (int)((row0 + rl) - n * HW);
    long long loaded = -1;
    for (long long r = rl; r < nrows; r += g.RL) {
        if (n != loaded) {                                   // a new sample: its SelfNorm coefficients (rarely: a chunk spans few samples)
            const size_t i = (size_t)n * g.C + c0;
#pragma unroll
            for (int e = 0; e < V; ++e) { kg[e] = gate[i + e]; kb[e] = cb[i + e]; kc[e] = cc[i + e]; }
            loaded = n;
        }
        float zv[V], dv[V], cv[V], o[V];
        unpack<T>(ldg_stream(vz + (size_t)r * g.CG), zv);
        unpack<T>(ldg_stream(vd + (size_t)r * g.CG), dv);
        unpack<T>(__ldg(vc + (size_t)r * g.CG), cv);
#pragma unroll
        for (int e = 0; e < V; ++e) {
            const float d = (relu && !(zv[e] > 0.f)) ? 0.f : dv[e];
            o[e] = fmaf(kg[e], d, fmaf(kb[e], zv[e], kc[e]));
        }
        const uint4 q = pack<T>(o);
        vo[(size_t)r * g.CG] = q;                            // default policy: bn3's apply reads it next
        unpack<T>(q, o);                                     // the reduction sees dz as the element type holds it
#pragma unroll
        for (int e = 0; e < V; ++e) { a[e] += o[e]; b[e] = fmaf(o[e], cv[e] - mean[e], b[e]); }
        rem += g.RL;
        while (rem >= HW) { rem -= HW; ++n; }
    }
#pragma unroll
    for (int e = 0; e < V; ++e) { s_a[rl * W + cg * V + e] = a[e]; s_b[rl * W + cg * V + e] = b[e]; }
    __syncthreads();
    for (int ch = threadIdx.x; ch < W; ch += bnl::kT) {
        double ta = 0.0, tb = 0.0;
        for (int l = 0; l < g.RL; ++l) { ta += (double)s_a[l * W + ch]; tb += (double)s_b[l * W + ch]; }
        tb *= (double)bn_rstd[blockIdx.y * W + ch];
        part[(size_t)bx * g.C + (size_t)blockIdx.y * W + ch] = make_float2((float)ta, (float)tb);
    }
}

// 0 = this geometry is supported
static int make_geom(Geom& g, int dtype, int N, int C, int H, int W) {
    const int esz = (int)esize(dtype);
    if (((size_t)C * esz) % 16) return CNSN_E_UNSUPPORTED;
    g.N = N; g.C = C; g.HW = H * W;
    g.CG = C * esz / 16;
    g.CGB = std::min(g.CG, kT);
    if (kT % g.CGB || g.CG % g.CGB) return CNSN_E_UNSUPPORTED;
    if (g.HW < 2) return CNSN_E_UNSUPPORTED;
    g.RL = kT / g.CGB;
    // slabs of about 32 KB, at least one pixel per row lane and pass; enough CTAs to fill the GPU
    const size_t tile = (size_t)g.HW * C * esz;
    int S = (int)std::max<size_t>(1, tile / (32u << 10));
    S = std::min(S, std::max(1, g.HW / (4 * g.RL)));
    g.rows = (g.HW + S - 1) / S;
    g.S = (g.HW + g.rows - 1) / g.rows;
    if ((long long)N * g.S > 0x7fffffffll) return CNSN_E_UNSUPPORTED;
    return 0;
}
static int max_slabs(int dtype, int C, int H, int W) {
    Geom g{};
    return make_geom(g, dtype, 1, C, H, W) ? 1 : g.S;
}
static size_t smem_bytes(const Geom& g, int dtype) {
    const size_t W = (size_t)g.CGB * (16 / esize(dtype));
    return ((size_t)g.RL * W + 2 * W) * sizeof(float);
}
static int channel_blocks(int dtype, int C) { return std::max(1, (int)(C * esize(dtype) / 16) / kT); }

}  // namespace nhwc
}  // namespace cnsn

using namespace cnsn;

// save: [SaveLayout(N, C, one gate) | slab partials float2 [N][S][C] | counters [N]]
extern "C" size_t cnsn_selfnorm_nhwc_save_floats(int dtype, int N, int C, int H, int W) {
    const size_t S = (size_t)nhwc::max_slabs(dtype, C, H, W);
    return SaveLayout(N, C, false).total + 2 * (size_t)N * S * C + (size_t)N * nhwc::channel_blocks(dtype, C) + 2;
}
// workspace: [sxy | st (unused) | cb | cc : 4 N C | slab partials [N][S][C] | counters [N]]
extern "C" size_t cnsn_selfnorm_nhwc_workspace_floats(int dtype, int N, int C, int H, int W) {
    const size_t S = (size_t)nhwc::max_slabs(dtype, C, H, W);
    return 4 * (size_t)N * C + (size_t)N * S * C + (size_t)N * nhwc::channel_blocks(dtype, C) + 2;
}
extern "C" int cnsn_selfnorm_nhwc_supported(int dtype, int N, int C, int H, int W) {
    nhwc::Geom g{};
    if (check_dims(N, C, H, W) || dtype < CNSN_F32 || dtype > CNSN_F16) return 0;
    return nhwc::make_geom(g, dtype, N, C, H, W) == 0;
}

static bool gate_ok_nhwc(const cnsn_gate_params* p) { return p && p->w && p->gamma && p->beta && p->run_mean && p->run_var; }

extern "C" int cnsn_selfnorm_block_fwd_nhwc(const void* x, const void* res, void* z, void* y, int relu, int dtype,
                                            int N, int C, int H, int W, const cnsn_gate_params* g,
                                            int training, float momentum, float bn_eps, float eps,
                                            float* save, void* stream) {
    if (!x || !y || !save || check_dims(N, C, H, W) || !gate_ok_nhwc(g)) return CNSN_E_BADARG;
    if (res && !z) return CNSN_E_BADARG;
    if (dtype < CNSN_F32 || dtype > CNSN_F16) return CNSN_E_BADARG;
    if (!aligned16(x) || !aligned16(y) || (res && (!aligned16(res) || !aligned16(z)))) return CNSN_E_ALIGN;
    if (training && N < 2) return CNSN_E_BATCH1;
    nhwc::Geom gm{};
    int rc = nhwc::make_geom(gm, dtype, N, C, H, W);
    if (rc) return rc;
    const SaveLayout L(N, C, false);
    float2* part = reinterpret_cast<float2*>(save + L.total);
    unsigned* cnt = reinterpret_cast<unsigned*>(save + L.total + 2 * (size_t)N * gm.S * C);
    cudaStream_t s = (cudaStream_t)stream;
    if (gm.S > 1) {
        const cudaError_t e = cudaMemsetAsync(cnt, 0, (size_t)N * (gm.CG / gm.CGB) * sizeof(unsigned), s);
        if (e != cudaSuccess) return (int)e;
    }
    const dim3 grid((unsigned)((long long)N * gm.S), (unsigned)(gm.CG / gm.CGB));
    const size_t smem = nhwc::smem_bytes(gm, dtype);
    CNSN_DISPATCH_DTYPE(dtype, T, {
        if (res) nhwc::k_nhwc_stats<T, true><<<grid, nhwc::kT, smem, s>>>((const T*)x, (const T*)res, (T*)z, gm, eps, part, cnt,
                                                                         save + L.mu, save + L.sd);
        else nhwc::k_nhwc_stats<T, false><<<grid, nhwc::kT, smem, s>>>((const T*)x, nullptr, nullptr, gm, eps, part, cnt,
                                                                      save + L.mu, save + L.sd);
    });
    if ((rc = launch_status())) return rc;
    GateFwd a{g->w, g->gamma, g->beta, g->run_mean, g->run_var, g->nbt, save + L.g, save + L.shat_g, save + L.r_g};
    k_sn_gate_fwd<<<dim3(C, 1), kGateThreads, 0, s>>>(save + L.mu, save + L.sd, a, a, N, C, training, momentum, bn_eps);
    if ((rc = launch_status())) return rc;
    const void* zz = res ? z : x;
    CNSN_DISPATCH_DTYPE(dtype, T,
        (nhwc::k_nhwc_apply<T, false><<<grid, nhwc::kT, 0, s>>>((const T*)zz, nullptr, (T*)y, gm, relu ? 1 : 0, save + L.g, nullptr, nullptr)));
    return launch_status();
}

extern "C" int cnsn_selfnorm_block_bwd_nhwc(const void* z, const void* dy, void* dz, int relu, int dtype,
                                            int N, int C, int H, int W, const cnsn_gate_params* g,
                                            int training, const float* save, const cnsn_gate_grads* dg,
                                            float* workspace, void* stream) {
    if (!z || !dy || !dz || !save || !workspace || check_dims(N, C, H, W)) return CNSN_E_BADARG;
    if (!g || !g->w || !g->gamma || !dg || !dg->dw || !dg->dgamma || !dg->dbeta) return CNSN_E_BADARG;
    if (dtype < CNSN_F32 || dtype > CNSN_F16) return CNSN_E_BADARG;
    if (!aligned16(z) || !aligned16(dy) || !aligned16(dz)) return CNSN_E_ALIGN;
    nhwc::Geom gm{};
    int rc = nhwc::make_geom(gm, dtype, N, C, H, W);
    if (rc) return rc;
    const SaveLayout L(N, C, false);
    const size_t nc = (size_t)N * C;
    float* sxy = workspace; float* st = workspace + nc; float* cb = workspace + 2 * nc; float* cc = workspace + 3 * nc;
    float* part = workspace + 4 * nc;
    unsigned* cnt = reinterpret_cast<unsigned*>(part + (size_t)N * gm.S * C);
    cudaStream_t s = (cudaStream_t)stream;
    if (gm.S > 1) {
        const cudaError_t e = cudaMemsetAsync(cnt, 0, (size_t)N * (gm.CG / gm.CGB) * sizeof(unsigned), s);
        if (e != cudaSuccess) return (int)e;
    }
    const dim3 grid((unsigned)((long long)N * gm.S), (unsigned)(gm.CG / gm.CGB));
    const size_t smem = nhwc::smem_bytes(gm, dtype);
    CNSN_DISPATCH_DTYPE(dtype, T,
        (nhwc::k_nhwc_reduce_bwd<T><<<grid, nhwc::kT, smem, s>>>((const T*)z, (const T*)dy, gm, relu ? 1 : 0, part, cnt, sxy)));
    if ((rc = launch_status())) return rc;
    GateBwd a{g->w, g->gamma, save + L.g, save + L.shat_g, save + L.r_g, dg->dw, dg->dgamma, dg->dbeta};
    k_sn_gate_bwd<<<C, kGateThreads, 0, s>>>(save + L.mu, save + L.sd, sxy, st, a, a, 0, N, C, H * W, training, cb, cc);
    if ((rc = launch_status())) return rc;
    CNSN_DISPATCH_DTYPE(dtype, T,
        (nhwc::k_nhwc_apply<T, true><<<grid, nhwc::kT, 0, s>>>((const T*)z, (const T*)dy, (T*)dz, gm, relu ? 1 : 0, save + L.g, cb, cc)));
    return launch_status();
}

// ---------------------------------------------------------------------------------------------------------------------------
// The tail of a pos='post' ResNet bottleneck as ONE operator (models/imagenet/resnet_cnsn.py:113-122):
//     out = self.bn3(out); out += identity; out = self.cnsn(out); out = self.relu(out)
// forward : bn3 statistics of c (the raw conv3 output) -> fold -> [z = bn3(c) + res, statistics of z] -> gate -> y
// backward: [sum d z] -> gate backward -> [dz, bn3's reduction] -> fold -> dc
// Against the two operators called one after the other this never writes bn3's output (forward: -2 S) and reads dz for the
// batch-norm reduction while it is being formed (backward: -1 S); every value is rounded exactly where the sequence rounds
// it, so the results are bit-identical to the sequence's.
extern "C" int cnsn_bn_selfnorm_tail_supported(int dtype, int N, int C, int H, int W) {
    nhwc::Geom g{};
    bnl::Geom b{};
    if (check_dims(N, C, H, W) || dtype < CNSN_F32 || dtype > CNSN_F16) return 0;
    return nhwc::make_geom(g, dtype, N, C, H, W) == 0 && bnl::make_geom(b, dtype, N, C, H, W) == 0;
}

extern "C" int cnsn_bn_selfnorm_tail_fwd_nhwc(const void* c, const void* res, void* z, void* y, int relu, int dtype,
                                              int N, int C, int H, int W,
                                              const float* bn_gamma, const float* bn_beta, float* bn_run_mean, float* bn_run_var,
                                              long long* bn_nbt, int bn_training, float bn_momentum, float bn_eps, float* bn_save,
                                              const cnsn_gate_params* g, int training, float momentum, float sn_bn_eps, float eps,
                                              float* sn_save, void* stream) {
    if (!c || !res || !z || !y || !bn_gamma || !bn_beta || !bn_run_mean || !bn_run_var || !bn_save || !sn_save ||
        check_dims(N, C, H, W) || !gate_ok_nhwc(g)) return CNSN_E_BADARG;
    if (dtype < CNSN_F32 || dtype > CNSN_F16) return CNSN_E_BADARG;
    if (!aligned16(c) || !aligned16(res) || !aligned16(z) || !aligned16(y)) return CNSN_E_ALIGN;
    if ((training && N < 2) || (bn_training && (long long)N * H * W < 2)) return CNSN_E_BATCH1;
    nhwc::Geom gm{};
    bnl::Geom bg{};
    int rc = nhwc::make_geom(gm, dtype, N, C, H, W);
    if (!rc) rc = bnl::make_geom(bg, dtype, N, C, H, W);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    // bn3: statistics of c -> mean / rstd / (scale, shift)
    float* bmean = bn_save; float* brstd = bn_save + C;
    float2* coef = reinterpret_cast<float2*>(bn_save + 2 * (size_t)C);
    float2* bpart = reinterpret_cast<float2*>(bn_save + 4 * (size_t)C);
    const dim3 bgrid((unsigned)bg.G, (unsigned)(bg.CG / bg.CGB));
    if (bn_training) {
        CNSN_DISPATCH_DTYPE(dtype, T, (bnl::k_bn_nhwc_stats<T><<<bgrid, bnl::kT, bnl::smem_bytes(bg, dtype), s>>>((const T*)c, bg, bpart)));
        if ((rc = launch_status())) return rc;
    }
    bnl::k_bn_nhwc_fold<<<C, bnl::kFoldT, 0, s>>>(bpart, bg, bn_gamma, bn_beta, bn_run_mean, bn_run_var, bn_nbt, bn_training, bn_momentum,
                                                  bn_eps, bmean, brstd, coef);
    if ((rc = launch_status())) return rc;
    // z = bn3(c) + res and its per-instance statistics, gate, y
    const SaveLayout L(N, C, false);
    float2* part = reinterpret_cast<float2*>(sn_save + L.total);
    unsigned* cnt = reinterpret_cast<unsigned*>(sn_save + L.total + 2 * (size_t)N * gm.S * C);
    if (gm.S > 1) {
        const cudaError_t e = cudaMemsetAsync(cnt, 0, (size_t)N * (gm.CG / gm.CGB) * sizeof(unsigned), s);
        if (e != cudaSuccess) return (int)e;
    }
    const dim3 grid((unsigned)((long long)N * gm.S), (unsigned)(gm.CG / gm.CGB));
    CNSN_DISPATCH_DTYPE(dtype, T, (nhwc::k_nhwc_stats<T, true, true><<<grid, nhwc::kT, nhwc::smem_bytes(gm, dtype), s>>>(
        (const T*)c, (const T*)res, (T*)z, gm, eps, part, cnt, sn_save + L.mu, sn_save + L.sd, coef)));
    if ((rc = launch_status())) return rc;
    GateFwd a{g->w, g->gamma, g->beta, g->run_mean, g->run_var, g->nbt, sn_save + L.g, sn_save + L.shat_g, sn_save + L.r_g};
    k_sn_gate_fwd<<<dim3(C, 1), kGateThreads, 0, s>>>(sn_save + L.mu, sn_save + L.sd, a, a, N, C, training, momentum, sn_bn_eps);
    if ((rc = launch_status())) return rc;
    CNSN_DISPATCH_DTYPE(dtype, T,
        (nhwc::k_nhwc_apply<T, false><<<grid, nhwc::kT, 0, s>>>((const T*)z, nullptr, (T*)y, gm, relu ? 1 : 0, sn_save + L.g, nullptr, nullptr)));
    return launch_status();
}

extern "C" int cnsn_bn_selfnorm_tail_bwd_nhwc(const void* c, const void* z, const void* dy, void* dz, void* dc, int relu, int dtype,
                                              int N, int C, int H, int W,
                                              const float* bn_gamma, int bn_training, const float* bn_save,
                                              float* d_bn_gamma, float* d_bn_beta, float* bn_workspace,
                                              const cnsn_gate_params* g, int training, const float* sn_save,
                                              const cnsn_gate_grads* dg, float* sn_workspace, void* stream) {
    if (!c || !z || !dy || !dz || !dc || !bn_gamma || !bn_save || !d_bn_gamma || !d_bn_beta || !bn_workspace || !sn_save ||
        !sn_workspace || check_dims(N, C, H, W)) return CNSN_E_BADARG;
    if (!g || !g->w || !g->gamma || !dg || !dg->dw || !dg->dgamma || !dg->dbeta) return CNSN_E_BADARG;
    if (dtype < CNSN_F32 || dtype > CNSN_F16) return CNSN_E_BADARG;
    if (!aligned16(c) || !aligned16(z) || !aligned16(dy) || !aligned16(dz) || !aligned16(dc)) return CNSN_E_ALIGN;
    nhwc::Geom gm{};
    bnl::Geom bg{};
    int rc = nhwc::make_geom(gm, dtype, N, C, H, W);
    if (!rc) rc = bnl::make_geom(bg, dtype, N, C, H, W);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const SaveLayout L(N, C, false);
    const size_t nc = (size_t)N * C;
    float* sxy = sn_workspace; float* st = sn_workspace + nc; float* cb = sn_workspace + 2 * nc; float* cc = sn_workspace + 3 * nc;
    float* part = sn_workspace + 4 * nc;
    unsigned* cnt = reinterpret_cast<unsigned*>(part + (size_t)N * gm.S * C);
    if (gm.S > 1) {
        const cudaError_t e = cudaMemsetAsync(cnt, 0, (size_t)N * (gm.CG / gm.CGB) * sizeof(unsigned), s);
        if (e != cudaSuccess) return (int)e;
    }
    const dim3 grid((unsigned)((long long)N * gm.S), (unsigned)(gm.CG / gm.CGB));
    CNSN_DISPATCH_DTYPE(dtype, T, (nhwc::k_nhwc_reduce_bwd<T><<<grid, nhwc::kT, nhwc::smem_bytes(gm, dtype), s>>>(
        (const T*)z, (const T*)dy, gm, relu ? 1 : 0, part, cnt, sxy)));
    if ((rc = launch_status())) return rc;
    GateBwd a{g->w, g->gamma, sn_save + L.g, sn_save + L.shat_g, sn_save + L.r_g, dg->dw, dg->dgamma, dg->dbeta};
    k_sn_gate_bwd<<<C, kGateThreads, 0, s>>>(sn_save + L.mu, sn_save + L.sd, sxy, st, a, a, 0, N, C, H * W, training, cb, cc);
    if ((rc = launch_status())) return rc;
    const float* bmean = bn_save; const float* brstd = bn_save + C;
    const float2* coef = reinterpret_cast<const float2*>(bn_save + 2 * (size_t)C);
    float2* bpart = reinterpret_cast<float2*>(bn_workspace);
    float* cdx = bn_workspace + 2 * (size_t)bg.G * C;
    const dim3 bgrid((unsigned)bg.G, (unsigned)(bg.CG / bg.CGB));
    CNSN_DISPATCH_DTYPE(dtype, T, (nhwc::k_tail_mid_bwd<T><<<bgrid, bnl::kT, bnl::smem_bytes(bg, dtype), s>>>(
        (const T*)z, (const T*)dy, (const T*)c, (T*)dz, bg, H * W, relu ? 1 : 0, sn_save + L.g, cb, cc, bmean, brstd, bpart)));
    if ((rc = launch_status())) return rc;
    bnl::k_bn_nhwc_fold_bwd<<<C, bnl::kFoldT, 0, s>>>(bpart, bg, bn_gamma, bn_training, bmean, brstd, d_bn_gamma, d_bn_beta, cdx);
    if ((rc = launch_status())) return rc;
    CNSN_DISPATCH_DTYPE(dtype, T,
        (bnl::k_bn_nhwc_apply<T, true, false><<<bgrid, bnl::kT, 0, s>>>((const T*)c, (const T*)dz, (T*)dc, bg, coef, cdx, 0)));      // the kernel in front ran back to front
    return launch_status();
}
