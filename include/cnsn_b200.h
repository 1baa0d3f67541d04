/*
 * cnsn_b200.h -- C ABI of libcnsn_b200.so: the CrossNorm / SelfNorm hot path of
 * amazon-science/crossnorm-selfnorm, rebuilt as hand-written sm_100a CUDA kernels.
 *
 * Every entry point replaces a piece of the reference's models/cnsn.py (paths below are
 * relative to the reference checkout) or of the autograd backward PyTorch derives from it.
 * The reference has no FFI of its own (it is pure Python over ATen); these are the symbols a
 * binding for that file would need.  INTEGRATION.md shows the ctypes stub.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross this boundary;
 *   - all tensor pointers are DEVICE pointers to dense NCHW storage (the reference forces
 *     .contiguous() itself at models/cnsn.py:14,16); element type given by `dtype`;
 *   - per-instance statistics, parameters, gradients of parameters and workspaces are fp32;
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous and stream-ordered;
 *   - the library never allocates device memory: workspaces / save areas are caller-owned, sizes come
 *     from the *_floats() helpers; its only state is the launch counter, the tuning knobs of cnsn_tune()
 *     and one 64-byte pinned host word for asynchronous errors (cnsn_async_error);
 *   - return value: 0 on success, a positive cudaError_t, or a negative CNSN_E_* code;
 *     cnsn_error_string() renders either.
 *   - windows are half-open [h0,h1) x [w0,w1) in NCHW rows/cols.  NOTE the reference names
 *     them bbx (dim 2) / bby (dim 3) with W/H swapped (models/cnsn.py:34-35,66,77).
 */
#ifndef CNSN_B200_H
#define CNSN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CNSN_ABI_VERSION 7

enum { CNSN_F32 = 0, CNSN_BF16 = 1, CNSN_F16 = 2 };

enum {
    CNSN_OK = 0,
    CNSN_E_BADARG = -1,   /* null pointer, non-positive dim, bad dtype, bad window, bad perm ptr */
    CNSN_E_WORKSPACE = -2,/* workspace too small */
    CNSN_E_BATCH1 = -3,   /* SelfNorm training with N == 1 (reference: BatchNorm1d ValueError) */
    CNSN_E_ALIGN = -4,    /* tensor base pointer not aligned to its element size */
    CNSN_E_UNSUPPORTED = -5,/* shape outside what this operator's kernels handle (cnsn_site_*: planes must be 16-byte
                             multiples and a channel's N planes must fit the GPU's shared memory; ask
                             cnsn_site_supported() first) */
    CNSN_E_TIMEOUT = -6   /* an EARLIER kernel of this process gave up a bounded wait (peer CTA / bulk copy); its results
                             are undefined.  Reported by the next call and by cnsn_async_error(); never a device trap */
};

/* Library / ABI identification. */
int cnsn_version(void);
const char* cnsn_error_string(int code);
/* Number of CUDA kernels this library has launched in this process (monotonic, atomic). */
unsigned long long cnsn_launch_count(void);
/*
 * Asynchronous error state.  The dataflow kernels wait for peer CTAs with bounded polls; a wait that expires
 * (seconds -- a debugger, a sanitizer or a bug) records a code in a pinned word instead of trapping, so the CUDA
 * context survives.  Returns CNSN_OK or CNSN_E_TIMEOUT; `clear` != 0 resets the state.  Every dataflow entry point
 * also checks it and refuses to launch (CNSN_E_TIMEOUT) until it has been cleared.
 */
int cnsn_async_error(int clear);
/*
 * Measurement / test hook, not part of the drop-in surface: set one process-wide tuning knob by name ("reset"
 * restores the defaults).  The library never reads environment variables; the names are those of struct Knobs in
 * csrc/flow_common.cuh.  Returns CNSN_E_BADARG for an unknown name.
 */
int cnsn_tune(const char* name, int value);

/*
 * Per-(n,c) mean and std = sqrt(unbiased_var + eps) over the window.
 * Replaces calc_ins_mean_std, models/cnsn.py:8-17 (and its use on crops at :66,:77).
 * mean/std: N*C floats.
 */
int cnsn_instance_stats(const void* x, int dtype, int N, int C, int H, int W,
                        int h0, int h1, int w0, int w1, float eps,
                        float* mean, float* std, void* stream);

/*
 * The same statistics of a STRIDED view (element strides sN, sC, sH, sW >= 0): sliced / transposed
 * views, crops and channels_last tensors are reduced where they lie -- the reference copies them first
 * (.contiguous(), models/cnsn.py:14,16).  Coalesced for W-contiguous views (sW == 1) and for channels_last (sC == 1).
 */
int cnsn_instance_stats_strided(const void* x, int dtype, int N, int C, int H, int W,
                                long long sN, long long sC, long long sH, long long sW,
                                int h0, int h1, int w0, int w1, float eps,
                                float* mean, float* std, void* stream);

/*
 * Backward of cnsn_instance_stats (autograd of models/cnsn.py:14-16):
 *   dx = dmean/M + (x - mean)/std * dstd/(M-1) inside the window, 0 outside.
 */
int cnsn_instance_stats_bwd(const void* x, void* dx, int dtype, int N, int C, int H, int W,
                            int h0, int h1, int w0, int w1,
                            const float* mean, const float* std,
                            const float* dmean, const float* dstd, void* stream);

/*
 * out[n,c,:,:] = x[n,c,:,:] * scale[n,c] + shift[n,c]   (scale/shift: N*C floats).
 * The broadcast normalise-and-restyle of instance_norm_mix, models/cnsn.py:27-29, for callers
 * that bring their own statistics.
 */
int cnsn_instance_affine(const void* x, void* out, int dtype, int N, int C, int H, int W,
                         const float* scale, const float* shift, void* stream);

/*
 * Per-instance sums sxy[n,c] = sum_hw dy*x and st[n,c] = sum_hw dy (N*C floats each): the
 * scale / shift gradients of cnsn_instance_affine, i.e. the reductions autograd performs for the
 * broadcast multiply-add at models/cnsn.py:27-29.
 */
int cnsn_instance_dot(const void* x, const void* dy, int dtype, int N, int C, int H, int W,
                      float* sxy, float* st, void* stream);

/* ---------------------------------------------------------------- SelfNorm ---------------
 * Replaces SelfNorm.forward, models/cnsn.py:130-150, and its autograd backward.
 *
 *   mu, sd   = instance stats (eps, 1e-12 in the reference :133)
 *   s        = w[c,0]*mu + w[c,1]*sd                      (depthwise Conv1d k=2, :137)
 *   train:   m = mean_n s, q = biased var_n s; running_mean/var updated (momentum, unbiased q)
 *   eval:    m, q = running_mean, running_var
 *   g        = sigmoid(gamma * (s-m)/sqrt(q+bn_eps) + beta)            (:138-139)
 *   y        = x*g                         or, is_two (f_* non-NULL): x*g + mu*(f-g)  (:142-150)
 *
 * Gate parameter block (all device fp32): w = g_fc.weight viewed (C,2); gamma/beta = g_bn affine;
 * run_mean/run_var = g_bn buffers (updated in place when training); nbt = num_batches_tracked
 * (int64, incremented when training; may be NULL).
 *
 * `save` (caller-owned, cnsn_selfnorm_save_floats() floats) receives what backward needs:
 *   [mu | sd | g | shat_g | (f | shat_f)] each N*C, then [r_g (C) | (r_f (C))], r = 1/sqrt(q+bn_eps).
 */
typedef struct cnsn_gate_params {
    const float* w;        /* (C,2) */
    const float* gamma;    /* (C)   */
    const float* beta;     /* (C)   */
    float* run_mean;       /* (C)   */
    float* run_var;        /* (C)   */
    long long* nbt;        /* scalar or NULL */
} cnsn_gate_params;

typedef struct cnsn_gate_grads {
    float* dw;             /* (C,2) */
    float* dgamma;         /* (C)   */
    float* dbeta;          /* (C)   */
} cnsn_gate_grads;

size_t cnsn_selfnorm_save_floats(int N, int C, int is_two);
size_t cnsn_selfnorm_workspace_floats(int N, int C, int is_two);

int cnsn_selfnorm_fwd(const void* x, void* y, int dtype, int N, int C, int H, int W,
                      const cnsn_gate_params* g, const cnsn_gate_params* f /* NULL unless is_two */,
                      int training, float momentum, float bn_eps, float eps,
                      float* save, void* stream);

/*
 * dx = dy*g + a + b*(x - mu); a = dmu/M, b = dsd/((M-1)*sd), through the BatchNorm-over-batch
 * backward when training.  Parameter gradients are WRITTEN (not accumulated).
 * workspace: cnsn_selfnorm_workspace_floats() floats.
 */
int cnsn_selfnorm_bwd(const void* x, const void* dy, void* dx, int dtype,
                      int N, int C, int H, int W,
                      const cnsn_gate_params* g, const cnsn_gate_params* f,
                      int training, const float* save,
                      const cnsn_gate_grads* dg, const cnsn_gate_grads* df,
                      float* workspace, void* stream);

/*
 * Block fusion for the reference's pos='post' sites: the residual add in front of the site and
 * the ReLU behind it, in the same kernels.  Replaces, in one forward and one backward call,
 *     out += identity; out = self.cnsn(out); out = self.relu(out)      models/imagenet/resnet_cnsn.py:117-122
 *     out = torch.add(x, out); return self.cnsn(out)                   models/cifar/wideresnet_cnsn.py:93-96
 * (SelfNorm-only sites; a site whose CrossNorm fires takes the unfused sequence.)
 *
 *   forward : z = x + res (written to `z`, element type `dtype`; pass res = z = NULL for "no add"),
 *             y = SelfNorm(z), and y = max(y, 0) when `relu`.
 *   backward: dz = SelfNorm backward at z of dy masked where z <= 0 when `relu` (the gate is a sigmoid, so
 *             relu(g*z) passes exactly where z > 0); dx = dres = dz.
 * save / workspace / g / dg as for cnsn_selfnorm_fwd / _bwd.
 */
int cnsn_selfnorm_block_fwd(const void* x, const void* res, void* z, void* y, int relu, int dtype,
                            int N, int C, int H, int W, const cnsn_gate_params* g,
                            int training, float momentum, float bn_eps, float eps,
                            float* save, void* stream);
int cnsn_selfnorm_block_bwd(const void* z, const void* dy, void* dz, int relu, int dtype,
                            int N, int C, int H, int W, const cnsn_gate_params* g,
                            int training, const float* save, const cnsn_gate_grads* dg,
                            float* workspace, void* stream);

/* Channels-last (NHWC) form of the block: the same operator on tensors whose logical shape is (N, C, H, W) and whose
 * memory order is N, H, W, C (torch.channels_last, dense) -- what a network keeps its activations in when cuDNN's
 * NHWC convolutions are not to convert around every call.  Same arguments and results as cnsn_selfnorm_block_fwd /
 * _bwd (res = z = NULL, relu = 0 is the plain models/cnsn.py:130-150 SelfNorm); x, res, z, y (z, dy, dz) are all NHWC.
 * Base pointers must be 16-byte aligned and C * sizeof(T) a multiple of 16 that divides, or is a multiple of, 4096
 * (cnsn_selfnorm_nhwc_supported), else CNSN_E_ALIGN / CNSN_E_UNSUPPORTED and the caller converts to NCHW.
 * save / workspace sizes are this form's own (they hold the per-slab partial statistics as well). */
int cnsn_selfnorm_nhwc_supported(int dtype, int N, int C, int H, int W);
size_t cnsn_selfnorm_nhwc_save_floats(int dtype, int N, int C, int H, int W);
size_t cnsn_selfnorm_nhwc_workspace_floats(int dtype, int N, int C, int H, int W);
int cnsn_selfnorm_block_fwd_nhwc(const void* x, const void* res, void* z, void* y, int relu, int dtype,
                                 int N, int C, int H, int W, const cnsn_gate_params* g,
                                 int training, float momentum, float bn_eps, float eps,
                                 float* save, void* stream);
int cnsn_selfnorm_block_bwd_nhwc(const void* z, const void* dy, void* dz, int relu, int dtype,
                                 int N, int C, int H, int W, const cnsn_gate_params* g,
                                 int training, const float* save, const cnsn_gate_grads* dg,
                                 float* workspace, void* stream);

/* ---------------------------------------------------------------- CrossNorm --------------
 * Replaces cn_op_2ins_space_chan, models/cnsn.py:58-91 (+ instance_norm_mix :20-29), device
 * side only; the host draws perm / windows (RNG contract) and passes them in.
 *
 *   inside the content window:
 *     y[i,c] = lam*x + (1-lam) * ((x - mu_c[i,c])/sd_c[i,c] * sd_s[p(i),pi(c)] + mu_s[p(i),pi(c)])
 *   outside: y = x.          (mu_c, sd_c) over the content window, (mu_s, sd_s) over the style window.
 *
 * perm: N int32 (device), a permutation of 0..N-1.  chan_perm: C int32 (device) or NULL.
 * lam: blend weight, pass 0 for the reference's lam=None.
 * save: 4*N*C floats, receives [mu_c | sd_c | mu_s | sd_s] for backward.
 */
/* content / style: HOST arrays of 4 ints {h0, h1, w0, w1}; pass {0,H,0,W} for "no crop". */
size_t cnsn_crossnorm_save_floats(int N, int C);
size_t cnsn_crossnorm_workspace_floats(int N, int C);

int cnsn_crossnorm_fwd(const void* x, void* y, int dtype, int N, int C, int H, int W,
                       const int* perm, const int* chan_perm,
                       const int* content, const int* style,
                       float lam, float eps, float* save, void* stream);

/*
 * Backward: nothing is detached, so dx carries the gradient through the content
 * statistics and, scattered through the permutation, through the style statistics.
 * workspace: cnsn_crossnorm_workspace_floats() floats.
 */
int cnsn_crossnorm_bwd(const void* x, const void* dy, void* dx, int dtype,
                       int N, int C, int H, int W,
                       const int* perm, const int* chan_perm,
                       const int* content, const int* style,
                       float lam, const float* save, float* workspace, void* stream);

/* ---------------------------------------------------------------- fused CNSN site --------
 * A site whose CrossNorm AND SelfNorm both fire in one step -- CNSN.forward, models/cnsn.py:159-164:
 *     if self.crossnorm and self.crossnorm.active: x = self.crossnorm(x)
 *     if self.selfnorm:                            x = self.selfnorm(x)
 * as one kernel per direction: the CrossNorm output never leaves the chip (2*S forward and 3*S
 * backward of HBM traffic instead of 4*S and 6*S), and backward keeps x plus O(N*C) statistics only.
 *
 *   forward : y = SelfNorm(CrossNorm(x)); arguments as cnsn_crossnorm_fwd (no channel permutation) and
 *             cnsn_selfnorm_fwd (single gate, TRAINING mode -- CrossNorm never fires in eval mode, :104).
 *             cn_eps: models/cnsn.py:8 (1e-5); sn_eps: :133 (1e-12).  Running statistics / nbt are updated.
 *             relu != 0: y = max(y, 0) -- the ReLU behind a pos='post' site (resnet_cnsn.py:122) in the same kernel.
 *   backward: dx of the composition (dy masked where the CrossNorm output is <= 0 when relu); parameter
 *             gradients are WRITTEN.
 * save: cnsn_site_save_floats() floats = [mu_c | sd_c | mu_s | sd_s | mu_z | sd_z | g | shat] (N*C each), r (C),
 * exchange area.  workspace: cnsn_site_workspace_floats() floats.
 * Shapes: planes must be multiples of 16 bytes and a channel's N planes (x and dy) must fit the GPU's shared
 * memory; cnsn_site_supported() returns 1 when both directions can run for the shape on the current device,
 * else 0 (the caller then runs cnsn_crossnorm_* and cnsn_selfnorm_* one after the other -- identical results);
 * the entry points return CNSN_E_UNSUPPORTED for such shapes and CNSN_E_BATCH1 for N == 1.
 */
size_t cnsn_site_save_floats(int N, int C);
size_t cnsn_site_workspace_floats(int N, int C);
int cnsn_site_supported(int dtype, int N, int C, int H, int W);
int cnsn_site_fwd(const void* x, void* y, int dtype, int N, int C, int H, int W,
                  const int* perm, const int* content, const int* style, float lam, float cn_eps,
                  const cnsn_gate_params* g, float momentum, float bn_eps, float sn_eps, int relu,
                  float* save, void* stream);
int cnsn_site_bwd(const void* x, const void* dy, void* dx, int dtype, int N, int C, int H, int W,
                  const int* perm, const int* content, const int* style, float lam, float cn_eps, int relu,
                  const cnsn_gate_params* g, const float* save, const cnsn_gate_grads* dg,
                  float* workspace, void* stream);

/* ---------------------------------------------------------------- IBN ---------------------
 * Instance-Batch Normalization, the IBN layer of models/imagenet/resnet_ibn_cnsn.py:24-44 (split along channels,
 * nn.InstanceNorm2d(half, affine=True) on the first `half` channels, nn.BatchNorm2d(C - half) on the rest,
 * concatenate), as one kernel per direction with no split / cat copies:
 *   c <  half : y = (x - mean_nc) / sqrt(biased_var_nc + eps_in) * in_w[c] + in_b[c]          (always instance statistics)
 *   c >= half : y = (x - m_c) / sqrt(v_c + eps_bn) * bn_w[c-half] + bn_b[c-half]; training: batch statistics over
 *               (N,H,W), running_mean / running_var (unbiased) updated with `momentum`, nbt incremented; eval: running.
 * half == C is a plain nn.InstanceNorm2d(C, affine=True) (the IBN-b blocks and stem, resnet_ibn_cnsn.py:62,122-123,
 * 143-144; bn_* / run_* may be NULL), half == 0 a plain BatchNorm2d.  One resident kernel per direction when planes are
 * multiples of 16 bytes and fit shared memory; every other shape -- and base pointers that are not 16-byte aligned --
 * takes three stream-ordered kernels per direction (same results).  save: cnsn_ibn_save_floats() floats, written by forward, read by backward; workspace:
 * cnsn_ibn_workspace_floats().  Parameter gradients are WRITTEN.
 */
typedef struct cnsn_ibn_params {
    const float* in_w;     /* (half)     IN.weight */
    const float* in_b;     /* (half)     IN.bias */
    const float* bn_w;     /* (C - half) BN.weight */
    const float* bn_b;     /* (C - half) BN.bias */
    float* run_mean;       /* (C - half) BN.running_mean */
    float* run_var;        /* (C - half) BN.running_var */
    long long* nbt;        /* BN.num_batches_tracked or NULL */
} cnsn_ibn_params;

size_t cnsn_ibn_save_floats(int N, int C, int half);
size_t cnsn_ibn_workspace_floats(int N, int C);
/* 1 when BOTH directions of this shape run as the one-launch shared-memory-resident kernel on the current device (what
 * a host model asks before it routes an nn.BatchNorm2d -- half == 0 -- through these entry points instead of cuDNN), else 0
 * (the entry points then take the three-kernel general path: correct, not tuned). */
int cnsn_ibn_resident(int dtype, int N, int C, int H, int W, int half, int training);
/* relu != 0: the ReLU that follows the norm in the host blocks (relu(bn(x)), wideresnet_cnsn.py:66-72, resnet_cnsn.py:
 * 100-110) in the same kernels -- forward y = max(y, 0); backward dy masked where y <= 0, y rebuilt from x and the saved
 * statistics (p->in_b / p->bn_b must then be given to the backward too).  Resident kernel only: CNSN_E_UNSUPPORTED for a
 * shape cnsn_ibn_resident() declines. */
int cnsn_ibn_fwd(const void* x, void* y, int dtype, int N, int C, int H, int W, int half,
                 const cnsn_ibn_params* p, int training, int relu, float momentum, float eps_in, float eps_bn,
                 float* save, void* stream);
int cnsn_ibn_bwd(const void* x, const void* dy, void* dx, int dtype, int N, int C, int H, int W, int half,
                 const cnsn_ibn_params* p, int training, int relu, const float* save,
                 float* d_in_w, float* d_in_b, float* d_bn_w, float* d_bn_b,
                 float* workspace, void* stream);

/* Channels-last (NHWC) BatchNorm2d [+ ReLU]: cnsn_ibn_fwd / _bwd with half = 0 on tensors whose logical shape is
 * (N, C, H, W) and whose memory order is N, H, W, C (torch.channels_last, dense); nn.BatchNorm2d semantics (batch
 * statistics when training, biased variance to normalise, unbiased into run_var; running statistics in eval mode),
 * `relu`: y = max(y, 0) forward, dy masked where y <= 0 backward (y rebuilt from x).  gamma, beta, run_mean, run_var,
 * dgamma, dbeta: fp32 [C].  C * sizeof(T) must be a multiple of 16 that divides, or is a multiple of, 4096
 * (cnsn_bn_nhwc_supported); base pointers 16-byte aligned.  save is written by forward and read by backward. */
int cnsn_bn_nhwc_supported(int dtype, int N, int C, int H, int W);
size_t cnsn_bn_nhwc_save_floats(int dtype, int N, int C, int H, int W);
size_t cnsn_bn_nhwc_workspace_floats(int dtype, int N, int C, int H, int W);
int cnsn_bn_nhwc_fwd(const void* x, void* y, int dtype, int N, int C, int H, int W,
                     const float* gamma, const float* beta, float* run_mean, float* run_var, long long* nbt,
                     int training, int relu, float momentum, float eps, float* save, void* stream);
int cnsn_bn_nhwc_bwd(const void* x, const void* dy, void* dx, int dtype, int N, int C, int H, int W,
                     const float* gamma, int training, int relu, const float* save,
                     float* dgamma, float* dbeta, float* workspace, void* stream);

/* Channels-last nn.MaxPool2d(k, stride, pad) (dilation 1, ceil_mode off; the stem of models/imagenet/resnet_cnsn.py:183,238):
 * x (N, C, H, W) and y (N, C, OH, OW) in N, H, W, C memory order, `code` one byte per OUTPUT element (the position of the
 * maximum inside its window, rows first; 8-byte aligned; written by forward, read by backward).  torch's tie rule: the first
 * of equal maxima wins, NaN propagates.  C * sizeof(T) must be a multiple of 16, 2 * pad <= k <= 15. */
int cnsn_maxpool_nhwc_out(int H, int W, int k, int stride, int pad, int* OH, int* OW);
int cnsn_maxpool_nhwc_fwd(const void* x, void* y, unsigned char* code, int dtype, int N, int C, int H, int W,
                          int k, int stride, int pad, void* stream);
int cnsn_maxpool_nhwc_bwd(const void* dy, const unsigned char* code, void* dx, int dtype, int N, int C, int H, int W,
                          int k, int stride, int pad, void* stream);

/* The tail of a pos='post' ResNet bottleneck as one operator, channels-last (models/imagenet/resnet_cnsn.py:113-122):
 *     out = self.bn3(out); out += identity; out = self.cnsn(out); out = self.relu(out)          (SelfNorm-only site)
 * c: the raw conv3 output, res: the identity branch, z (written): bn3(c) + res -- what backward needs --, y: the result.
 * Backward: dz (written) is the gradient of the residual branch, dc the gradient of c; d_bn_gamma / d_bn_beta / dg are
 * WRITTEN.  bn_save / bn_workspace as cnsn_bn_nhwc_*, sn_save / sn_workspace as cnsn_selfnorm_block_*_nhwc.  Bit-identical
 * to cnsn_bn_nhwc_fwd followed by cnsn_selfnorm_block_fwd_nhwc (and the two backward calls in reverse order); it moves
 * 2 S less forward (bn3's output is never written) and 1 S less backward. */
int cnsn_bn_selfnorm_tail_supported(int dtype, int N, int C, int H, int W);
int cnsn_bn_selfnorm_tail_fwd_nhwc(const void* c, const void* res, void* z, void* y, int relu, int dtype,
                                   int N, int C, int H, int W,
                                   const float* bn_gamma, const float* bn_beta, float* bn_run_mean, float* bn_run_var,
                                   long long* bn_nbt, int bn_training, float bn_momentum, float bn_eps, float* bn_save,
                                   const cnsn_gate_params* g, int training, float momentum, float sn_bn_eps, float eps,
                                   float* sn_save, void* stream);
int cnsn_bn_selfnorm_tail_bwd_nhwc(const void* c, const void* z, const void* dy, void* dz, void* dc, int relu, int dtype,
                                   int N, int C, int H, int W,
                                   const float* bn_gamma, int bn_training, const float* bn_save,
                                   float* d_bn_gamma, float* d_bn_beta, float* bn_workspace,
                                   const cnsn_gate_params* g, int training, const float* sn_save,
                                   const cnsn_gate_grads* dg, float* sn_workspace, void* stream);

/* ---------------------------------------------------------------- JSD consistency -----------
 * The Jensen-Shannon consistency term of the 3-view steps, imagenet.py:367-376 / cifar.py:173-182:
 *   p_v = softmax(logits_v); lm = log(clamp(mean_v p_v, 1e-7, 1));
 *   loss = (kl_div(lm, p_clean) + kl_div(lm, p_aug1) + kl_div(lm, p_aug2)) / 3      (reduction 'batchmean')
 * z0, z1, z2: (B, K) dense logits of the clean / aug1 / aug2 views (element type `dtype`); row_loss: B floats of
 * scratch; loss: 1 float.  Backward: gout is a DEVICE scalar (d total / d loss); d0..d2 receive d loss / d logits.
 */
int cnsn_jsd_fwd(const void* z0, const void* z1, const void* z2, int dtype, int B, int K,
                 float* row_loss, float* loss, void* stream);
int cnsn_jsd_bwd(const void* z0, const void* z1, const void* z2, int dtype, int B, int K,
                 const float* gout, void* d0, void* d1, void* d2, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CNSN_B200_H */
