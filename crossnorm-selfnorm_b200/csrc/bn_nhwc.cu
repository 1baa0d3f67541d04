// bn_nhwc.cu -- BatchNorm2d (+ the ReLU that follows it in the host blocks) on CHANNELS-LAST tensors: logical shape
// (N, C, H, W), memory order N, H, W, C.  The batch-norm side of cnsn_ibn_* (half = 0: every channel batch norm,
// models/imagenet/resnet_ibn_cnsn.py:24-44 / nn.BatchNorm2d semantics) for a network that keeps its activations in the
// layout cuDNN's convolutions work in (see selfnorm_nhwc.cu for why).
//
// In NHWC the tensor is a matrix [R = N*H*W rows][C columns] and a channel is a column of it:
//   forward : k_bn_nhwc_stats   CTA (chunk, channel block): one pass over its rows, per-thread shifted sums (shift = the
//                               thread's first element, so the sums stay small), merged (Chan) over the CTA's row lanes
//                               -> one (mean, M2) pair per chunk and channel
//             k_bn_nhwc_fold    one CTA per channel: merge of the chunk pairs in double -> batch mean / rstd,
//                               running statistics, the two coefficients of y = scale * x + shift
//             k_bn_nhwc_apply   y = relu?(scale * x + shift), coefficients in registers (a thread's channels never change)
//   backward: k_bn_nhwc_reduce  per chunk and channel (sum d, sum d * xhat), d = dy masked where the forward output was
//                               <= 0 (rebuilt from x with the forward's own coefficients: same fp32 expressions, same bits)
//             k_bn_nhwc_fold_bwd  dgamma, dbeta, and the coefficients of dx = ca * d + cb * x + cc
//             k_bn_nhwc_apply   dx
// Algorithmic bytes 2 S forward, 3 S backward; moved: one more read of x (forward) / x and dy (backward), served by L2
// when the tensor fits (16-64 MB at the WideResNet sites against 126 MB).  Eval mode: running statistics, fold + apply.
#include <stdio.h>

#include "bn_nhwc.cuh"

using namespace cnsn;

extern "C" int cnsn_bn_nhwc_supported(int dtype, int N, int C, int H, int W) {
    bnl::Geom g{};
    if (check_dims(N, C, H, W) || dtype < CNSN_F32 || dtype > CNSN_F16) return 0;
    return bnl::make_geom(g, dtype, N, C, H, W) == 0;
}
// save: [mean C | rstd C | (scale, shift) 2 C | chunk pairs 2 * kMaxChunks' worth at most: 2 G C]
extern "C" size_t cnsn_bn_nhwc_save_floats(int dtype, int N, int C, int H, int W) {
    bnl::Geom g{};
    const size_t G = bnl::make_geom(g, dtype, N, C, H, W) ? 1 : (size_t)g.G;
    return 4 * (size_t)C + 2 * G * C + 2;
}
// workspace: [chunk pairs 2 G C | (ca, cb, cc) 3 C]
extern "C" size_t cnsn_bn_nhwc_workspace_floats(int dtype, int N, int C, int H, int W) {
    bnl::Geom g{};
    const size_t G = bnl::make_geom(g, dtype, N, C, H, W) ? 1 : (size_t)g.G;
    return 2 * G * C + 3 * (size_t)C + 2;
}

extern "C" int cnsn_bn_nhwc_fwd(const void* x, void* y, int dtype, int N, int C, int H, int W,
                                const float* gamma, const float* beta, float* run_mean, float* run_var, long long* nbt,
                                int training, int relu, float momentum, float eps, float* save, void* stream) {
    if (!x || !y || !gamma || !beta || !run_mean || !run_var || !save || check_dims(N, C, H, W)) return CNSN_E_BADARG;
    if (dtype < CNSN_F32 || dtype > CNSN_F16) return CNSN_E_BADARG;
    if (!aligned16(x) || !aligned16(y)) return CNSN_E_ALIGN;
    if (training && (long long)N * H * W < 2) return CNSN_E_BATCH1;
    bnl::Geom g{};
    int rc = bnl::make_geom(g, dtype, N, C, H, W);
    if (rc) return rc;
    float* mean = save; float* rstd = save + C;
    float2* coef = reinterpret_cast<float2*>(save + 2 * (size_t)C);
    float2* part = reinterpret_cast<float2*>(save + 4 * (size_t)C);
    cudaStream_t s = (cudaStream_t)stream;
    const dim3 grid((unsigned)g.G, (unsigned)(g.CG / g.CGB));
    if (training) {
        CNSN_DISPATCH_DTYPE(dtype, T, (bnl::k_bn_nhwc_stats<T><<<grid, bnl::kT, bnl::smem_bytes(g, dtype), s>>>((const T*)x, g, part)));
        if ((rc = launch_status())) return rc;
    }
    bnl::k_bn_nhwc_fold<<<C, bnl::kFoldT, 0, s>>>(part, g, gamma, beta, run_mean, run_var, nbt, training, momentum, eps, mean, rstd, coef);
    if ((rc = launch_status())) return rc;
    CNSN_DISPATCH_DTYPE(dtype, T, CNSN_DISPATCH_BOOL(relu != 0, RELU,
        (bnl::k_bn_nhwc_apply<T, false, RELU><<<grid, bnl::kT, 0, s>>>((const T*)x, nullptr, (T*)y, g, coef, nullptr, 1))));
    return launch_status();
}

extern "C" int cnsn_bn_nhwc_bwd(const void* x, const void* dy, void* dx, int dtype, int N, int C, int H, int W,
                                const float* gamma, int training, int relu, const float* save,
                                float* dgamma, float* dbeta, float* workspace, void* stream) {
    if (!x || !dy || !dx || !gamma || !save || !dgamma || !dbeta || !workspace || check_dims(N, C, H, W)) return CNSN_E_BADARG;
    if (dtype < CNSN_F32 || dtype > CNSN_F16) return CNSN_E_BADARG;
    if (!aligned16(x) || !aligned16(dy) || !aligned16(dx)) return CNSN_E_ALIGN;
    bnl::Geom g{};
    int rc = bnl::make_geom(g, dtype, N, C, H, W);
    if (rc) return rc;
    const float* mean = save; const float* rstd = save + C;
    const float2* coef = reinterpret_cast<const float2*>(save + 2 * (size_t)C);
    float2* part = reinterpret_cast<float2*>(workspace);
    float* cdx = workspace + 2 * (size_t)g.G * C;
    cudaStream_t s = (cudaStream_t)stream;
    const dim3 grid((unsigned)g.G, (unsigned)(g.CG / g.CGB));
    CNSN_DISPATCH_DTYPE(dtype, T, CNSN_DISPATCH_BOOL(relu != 0, RELU, (bnl::k_bn_nhwc_reduce<T, RELU><<<grid, bnl::kT, bnl::smem_bytes(g, dtype), s>>>(
        (const T*)x, (const T*)dy, g, mean, rstd, coef, part))));
    if ((rc = launch_status())) return rc;
    bnl::k_bn_nhwc_fold_bwd<<<C, bnl::kFoldT, 0, s>>>(part, g, gamma, training, mean, rstd, dgamma, dbeta, cdx);
    if ((rc = launch_status())) return rc;
    CNSN_DISPATCH_DTYPE(dtype, T, CNSN_DISPATCH_BOOL(relu != 0, RELU,
        (bnl::k_bn_nhwc_apply<T, true, RELU><<<grid, bnl::kT, 0, s>>>((const T*)x, (const T*)dy, (T*)dx, g, coef, cdx, 1))));
    return launch_status();
}
