// cnsn_torch.cpp -- the low-overhead host binding above the C ABI (include/cnsn_b200.h): one C++ autograd node per
// operator, so that a forward or backward call of SelfNorm / CrossNorm / the fused CNSN site costs one pybind call,
// a few caching-allocator allocations and the library's launches -- no Python autograd.Function, no ctypes marshalling,
// no per-call torch.empty / .to() round trips through the interpreter, and a backward that never takes the GIL.
//
// It is the same boundary as the ctypes binding in _lib.py (which stays: CPU host-logic tests with a stand-in backend,
// is_two, channel permutation, parameters that are not fp32): tensors in, plain pointers and sizes down to
// libcnsn_b200.so, nothing but torch plumbing here.  PyTorch provides device memory, streams and the autograd graph.
//
// Reference semantics kept here:
//   * torch.randperm(N) of cn_op_2ins_space_chan (models/cnsn.py:62) is drawn HERE with at::randperm on the default CPU
//     generator -- the same stream and the same number of draws as the Python call it replaces; the numpy draws of
//     cn_rand_bbox (:36-53) stay in Python (numpy's global state) and arrive as windows.
//   * the permutation is staged through a pinned ring and copied asynchronously (the reference's blocking pageable
//     copy of the index tensor is not reproduced).
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <cuda_runtime_api.h>
#include <torch/extension.h>

#include <mutex>
#include <vector>

#include "../../include/cnsn_b200.h"

namespace {

using torch::autograd::AutogradContext;
using torch::autograd::variable_list;

int dtype_code(const at::Tensor& t) {
    switch (t.scalar_type()) {
        case at::kFloat: return CNSN_F32;
        case at::kBFloat16: return CNSN_BF16;
        case at::kHalf: return CNSN_F16;
        default: TORCH_CHECK(false, "cnsn_b200 supports float32 / bfloat16 / float16 tensors, got ", t.scalar_type());
    }
}

void check(int rc) {
    if (rc == 0) return;
    const char* msg = cnsn_error_string(rc);
    if (rc == CNSN_E_BATCH1) throw pybind11::value_error(msg);      // what nn.BatchNorm1d raises inside the reference SelfNorm
    TORCH_CHECK(false, "cnsn_b200 error ", rc, ": ", msg);
}

void require_cuda4(const at::Tensor& x) {
    TORCH_CHECK(x.is_cuda(), "cnsn_b200 operators run only on CUDA tensors (B200, sm_100a); there is no CPU fallback");
    TORCH_CHECK(x.dim() == 4, "expected an (N,C,H,W) tensor");
}

at::Tensor f32_buffer(const at::Tensor& like, int64_t n) {
    return at::empty({n}, like.options().dtype(at::kFloat));
}

// Gate tensors that autograd must NOT see as inputs (buffers updated in place by the kernels): a plain struct passes
// through Function::apply untouched.
struct GateBufs {
    at::Tensor run_mean, run_var, nbt;
};

cnsn_gate_params gate_params(const at::Tensor& w, const at::Tensor& gamma, const at::Tensor& beta, const GateBufs* b) {
    cnsn_gate_params g{};
    g.w = w.data_ptr<float>();
    g.gamma = gamma.data_ptr<float>();
    g.beta = beta.defined() ? beta.data_ptr<float>() : nullptr;
    if (b) {
        g.run_mean = b->run_mean.data_ptr<float>();
        g.run_var = b->run_var.data_ptr<float>();
        g.nbt = b->nbt.defined() ? reinterpret_cast<long long*>(b->nbt.data_ptr<int64_t>()) : nullptr;
    }
    return g;
}

// ---- pinned staging ring for permutation uploads -----------------------------------------------------------------
// kSlots blocks of kWords int32 in pinned host memory; a block is reused only after the copy that read it has
// completed (one event per block; 64 uploads later it always has).
struct PermRing {
    static constexpr int kSlots = 64, kWords = 8192;
    int* host = nullptr;
    cudaEvent_t ev[kSlots] = {};
    bool used[kSlots] = {};
    int next = 0;
    std::mutex mu;

    at::Tensor upload(const at::Tensor& perm_cpu, const at::Tensor& like, cudaStream_t stream) {
        const int64_t n = perm_cpu.numel();
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(stream, &cap);
        TORCH_CHECK(cap == cudaStreamCaptureStatusNone,
                    "cnsn_b200: CrossNorm draws a fresh permutation on the host every call (models/cnsn.py:62) and cannot be "
                    "captured into a CUDA graph; capture the steps whose CrossNorm sites are inactive");
        at::Tensor dev = at::empty({n}, like.options().dtype(at::kInt));
        if (n > kWords) {                                             // very large batches: plain (blocking) path
            dev.copy_(perm_cpu.to(at::kInt));
            return dev;
        }
        std::lock_guard<std::mutex> lock(mu);
        if (!host) {
            TORCH_CHECK(cudaHostAlloc(reinterpret_cast<void**>(&host), sizeof(int) * kSlots * kWords, cudaHostAllocPortable) == cudaSuccess,
                        "cnsn_b200: pinned staging allocation failed");
        }
        const int s = next;
        next = (next + 1) % kSlots;
        if (used[s]) cudaEventSynchronize(ev[s]);
        else TORCH_CHECK(cudaEventCreateWithFlags(&ev[s], cudaEventDisableTiming) == cudaSuccess, "cudaEventCreate failed");
        int* dst = host + (size_t)s * kWords;
        const int64_t* src = perm_cpu.data_ptr<int64_t>();
        for (int64_t i = 0; i < n; ++i) dst[i] = (int)src[i];
        TORCH_CHECK(cudaMemcpyAsync(dev.data_ptr<int>(), dst, sizeof(int) * n, cudaMemcpyHostToDevice, stream) == cudaSuccess,
                    "cnsn_b200: permutation upload failed");
        cudaEventRecord(ev[s], stream);
        used[s] = true;
        return dev;
    }
};
PermRing& ring() {
    static PermRing r;
    return r;
}

struct Win {
    int v[4];
};
Win window(const std::vector<int64_t>& w) {
    TORCH_CHECK(w.size() == 4, "a window is (h0, h1, w0, w1)");
    return Win{{(int)w[0], (int)w[1], (int)w[2], (int)w[3]}};
}

// A channels_last caller gets channels_last back from the operators whose kernels are NCHW (CrossNorm, the fused site):
// whatever follows (cuDNN's NHWC convolutions, forward and backward) then keeps seeing ONE layout -- no new cuDNN plans,
// no conversions further down.
bool is_channels_last(const at::Tensor& t) {
    return t.dim() == 4 && !t.is_contiguous() && t.is_contiguous(at::MemoryFormat::ChannelsLast);
}
at::Tensor like_input(const at::Tensor& t, bool cl) { return cl ? t.contiguous(at::MemoryFormat::ChannelsLast) : t; }

// ---- SelfNorm (single gate), optionally relu?(SelfNorm(x + res)) ---------------------------------------------------
struct SelfNormNode : public torch::autograd::Function<SelfNormNode> {
    static at::Tensor forward(AutogradContext* ctx, const at::Tensor& x_in, const c10::optional<at::Tensor>& res_in, bool relu,
                              bool block, const at::Tensor& w, const at::Tensor& gamma, const at::Tensor& beta, GateBufs bufs,
                              bool training, double momentum, double bn_eps, double eps) {
        require_cuda4(x_in);
        const int N = (int)x_in.size(0), C = (int)x_in.size(1), H = (int)x_in.size(2), W = (int)x_in.size(3);
        // a dense channels_last tensor stays in its layout when the NHWC kernels take the shape (y, z, dz come out channels_last)
        const bool cl = !x_in.is_contiguous() && x_in.is_contiguous(at::MemoryFormat::ChannelsLast) &&
                        (reinterpret_cast<uintptr_t>(x_in.data_ptr()) & 15u) == 0 &&
                        cnsn_selfnorm_nhwc_supported(dtype_code(x_in), N, C, H, W) != 0;
        const at::MemoryFormat fmt = cl ? at::MemoryFormat::ChannelsLast : at::MemoryFormat::Contiguous;
        const at::Tensor x = x_in.contiguous(fmt);
        const bool has_res = res_in.has_value() && res_in->defined();
        at::Tensor res;
        if (has_res) {
            TORCH_CHECK(res_in->sizes() == x.sizes() && res_in->scalar_type() == x.scalar_type() && res_in->is_cuda(), "residual must match x");
            res = res_in->contiguous(fmt);
        }
        const c10::cuda::CUDAGuard guard(x.device());
        cudaStream_t stream = at::cuda::getCurrentCUDAStream();
        at::Tensor save = f32_buffer(x, (int64_t)(cl ? cnsn_selfnorm_nhwc_save_floats(dtype_code(x), N, C, H, W) : cnsn_selfnorm_save_floats(N, C, 0)));
        at::Tensor y = at::empty_like(x);
        at::Tensor z = has_res ? at::empty_like(x) : x;
        const cnsn_gate_params g = gate_params(w, gamma, beta, &bufs);
        if (cl) {
            check(cnsn_selfnorm_block_fwd_nhwc(x.data_ptr(), has_res ? res.data_ptr() : nullptr, has_res ? z.data_ptr() : nullptr,
                                               y.data_ptr(), relu ? 1 : 0, dtype_code(x), N, C, H, W, &g, training ? 1 : 0,
                                               (float)momentum, (float)bn_eps, (float)eps, save.data_ptr<float>(), stream));
        } else if (block) {
            check(cnsn_selfnorm_block_fwd(x.data_ptr(), has_res ? res.data_ptr() : nullptr, has_res ? z.data_ptr() : nullptr,
                                          y.data_ptr(), relu ? 1 : 0, dtype_code(x), N, C, H, W, &g, training ? 1 : 0,
                                          (float)momentum, (float)bn_eps, (float)eps, save.data_ptr<float>(), stream));
        } else {
            check(cnsn_selfnorm_fwd(x.data_ptr(), y.data_ptr(), dtype_code(x), N, C, H, W, &g, nullptr, training ? 1 : 0,
                                    (float)momentum, (float)bn_eps, (float)eps, save.data_ptr<float>(), stream));
        }
        ctx->save_for_backward({z, w, gamma, beta, save});
        ctx->saved_data["relu"] = relu;
        ctx->saved_data["block"] = block;
        ctx->saved_data["training"] = training;
        ctx->saved_data["has_res"] = has_res;
        ctx->saved_data["cl"] = cl;
        return y;
    }

    static variable_list backward(AutogradContext* ctx, variable_list grads) {
        const auto saved = ctx->get_saved_variables();
        const at::Tensor &z = saved[0], &w = saved[1], &gamma = saved[2], &beta = saved[3], &save = saved[4];
        const bool relu = ctx->saved_data["relu"].toBool(), block = ctx->saved_data["block"].toBool();
        const bool training = ctx->saved_data["training"].toBool(), has_res = ctx->saved_data["has_res"].toBool();
        const bool cl = ctx->saved_data["cl"].toBool();
        const at::Tensor dy = grads[0].contiguous(cl ? at::MemoryFormat::ChannelsLast : at::MemoryFormat::Contiguous);
        const c10::cuda::CUDAGuard guard(z.device());
        cudaStream_t stream = at::cuda::getCurrentCUDAStream();
        const int N = (int)z.size(0), C = (int)z.size(1), H = (int)z.size(2), W = (int)z.size(3);
        at::Tensor pg = f32_buffer(z, 4 * (int64_t)C);                // dw (C,2) | dgamma (C) | dbeta (C)
        at::Tensor ws = f32_buffer(z, (int64_t)(cl ? cnsn_selfnorm_nhwc_workspace_floats(dtype_code(z), N, C, H, W)
                                                   : cnsn_selfnorm_workspace_floats(N, C, 0)));
        at::Tensor dz = at::empty_like(z);
        const cnsn_gate_params g = gate_params(w, gamma, beta, nullptr);
        float* p = pg.data_ptr<float>();
        const cnsn_gate_grads gg{p, p + 2 * C, p + 3 * C};
        if (cl) {
            check(cnsn_selfnorm_block_bwd_nhwc(z.data_ptr(), dy.data_ptr(), dz.data_ptr(), relu ? 1 : 0, dtype_code(z), N, C, H, W, &g,
                                               training ? 1 : 0, save.data_ptr<float>(), &gg, ws.data_ptr<float>(), stream));
        } else if (block) {
            check(cnsn_selfnorm_block_bwd(z.data_ptr(), dy.data_ptr(), dz.data_ptr(), relu ? 1 : 0, dtype_code(z), N, C, H, W, &g,
                                          training ? 1 : 0, save.data_ptr<float>(), &gg, ws.data_ptr<float>(), stream));
        } else {
            check(cnsn_selfnorm_bwd(z.data_ptr(), dy.data_ptr(), dz.data_ptr(), dtype_code(z), N, C, H, W, &g, nullptr,
                                    training ? 1 : 0, save.data_ptr<float>(), &gg, nullptr, ws.data_ptr<float>(), stream));
        }
        at::Tensor none;
        return {dz, has_res ? dz : none, none, none, pg.narrow(0, 0, 2 * C).view_as(w), pg.narrow(0, 2 * C, C), pg.narrow(0, 3 * C, C),
                none, none, none, none, none};
    }
};

// ---- CrossNorm (no channel permutation) ------------------------------------------------------------------------------
struct CrossNormNode : public torch::autograd::Function<CrossNormNode> {
    static at::Tensor forward(AutogradContext* ctx, const at::Tensor& x_in, Win cw, Win sw, double lam, double eps) {
        require_cuda4(x_in);
        const at::Tensor x = x_in.contiguous();
        const c10::cuda::CUDAGuard guard(x.device());
        cudaStream_t stream = at::cuda::getCurrentCUDAStream();
        const int N = (int)x.size(0), C = (int)x.size(1), H = (int)x.size(2), W = (int)x.size(3);
        const at::Tensor perm_cpu = at::randperm(N, at::TensorOptions().dtype(at::kLong));     // models/cnsn.py:62
        at::Tensor perm = ring().upload(perm_cpu, x, stream);
        at::Tensor save = f32_buffer(x, (int64_t)cnsn_crossnorm_save_floats(N, C));
        at::Tensor y = at::empty_like(x);
        check(cnsn_crossnorm_fwd(x.data_ptr(), y.data_ptr(), dtype_code(x), N, C, H, W, perm.data_ptr<int>(), nullptr, cw.v, sw.v,
                                 (float)lam, (float)eps, save.data_ptr<float>(), stream));
        ctx->save_for_backward({x, perm, save});
        ctx->saved_data["cw"] = std::vector<int64_t>(cw.v, cw.v + 4);
        ctx->saved_data["sw"] = std::vector<int64_t>(sw.v, sw.v + 4);
        ctx->saved_data["lam"] = lam;
        ctx->saved_data["cl_in"] = is_channels_last(x_in);
        return like_input(y, is_channels_last(x_in));
    }

    static variable_list backward(AutogradContext* ctx, variable_list grads) {
        const auto saved = ctx->get_saved_variables();
        const at::Tensor &x = saved[0], &perm = saved[1], &save = saved[2];
        const Win cw = window(ctx->saved_data["cw"].toIntVector()), sw = window(ctx->saved_data["sw"].toIntVector());
        const double lam = ctx->saved_data["lam"].toDouble();
        const at::Tensor dy = grads[0].contiguous();
        const c10::cuda::CUDAGuard guard(x.device());
        cudaStream_t stream = at::cuda::getCurrentCUDAStream();
        const int N = (int)x.size(0), C = (int)x.size(1), H = (int)x.size(2), W = (int)x.size(3);
        at::Tensor ws = f32_buffer(x, (int64_t)cnsn_crossnorm_workspace_floats(N, C));
        at::Tensor dx = at::empty_like(x);
        check(cnsn_crossnorm_bwd(x.data_ptr(), dy.data_ptr(), dx.data_ptr(), dtype_code(x), N, C, H, W, perm.data_ptr<int>(), nullptr,
                                 cw.v, sw.v, (float)lam, save.data_ptr<float>(), ws.data_ptr<float>(), stream));
        at::Tensor none;
        return {like_input(dx, ctx->saved_data["cl_in"].toBool()), none, none, none, none};
    }
};

// ---- fused site: relu?(SelfNorm(CrossNorm(x))) -------------------------------------------------------------------------
struct SiteNode : public torch::autograd::Function<SiteNode> {
    static at::Tensor forward(AutogradContext* ctx, const at::Tensor& x_in, Win cw, Win sw, double lam, double cn_eps, bool relu,
                              const at::Tensor& w, const at::Tensor& gamma, const at::Tensor& beta, GateBufs bufs,
                              double momentum, double bn_eps, double sn_eps) {
        require_cuda4(x_in);
        const at::Tensor x = x_in.contiguous();
        const c10::cuda::CUDAGuard guard(x.device());
        cudaStream_t stream = at::cuda::getCurrentCUDAStream();
        const int N = (int)x.size(0), C = (int)x.size(1), H = (int)x.size(2), W = (int)x.size(3);
        const at::Tensor perm_cpu = at::randperm(N, at::TensorOptions().dtype(at::kLong));     // models/cnsn.py:62
        at::Tensor perm = ring().upload(perm_cpu, x, stream);
        at::Tensor save = f32_buffer(x, (int64_t)cnsn_site_save_floats(N, C));
        at::Tensor y = at::empty_like(x);
        const cnsn_gate_params g = gate_params(w, gamma, beta, &bufs);
        check(cnsn_site_fwd(x.data_ptr(), y.data_ptr(), dtype_code(x), N, C, H, W, perm.data_ptr<int>(), cw.v, sw.v, (float)lam,
                            (float)cn_eps, &g, (float)momentum, (float)bn_eps, (float)sn_eps, relu ? 1 : 0, save.data_ptr<float>(),
                            stream));
        ctx->save_for_backward({x, perm, save, w, gamma, beta});
        ctx->saved_data["cw"] = std::vector<int64_t>(cw.v, cw.v + 4);
        ctx->saved_data["sw"] = std::vector<int64_t>(sw.v, sw.v + 4);
        ctx->saved_data["lam"] = lam;
        ctx->saved_data["cn_eps"] = cn_eps;
        ctx->saved_data["relu"] = relu;
        ctx->saved_data["cl_in"] = is_channels_last(x_in);
        return like_input(y, is_channels_last(x_in));
    }

    static variable_list backward(AutogradContext* ctx, variable_list grads) {
        const auto saved = ctx->get_saved_variables();
        const at::Tensor &x = saved[0], &perm = saved[1], &save = saved[2], &w = saved[3], &gamma = saved[4], &beta = saved[5];
        const Win cw = window(ctx->saved_data["cw"].toIntVector()), sw = window(ctx->saved_data["sw"].toIntVector());
        const double lam = ctx->saved_data["lam"].toDouble(), cn_eps = ctx->saved_data["cn_eps"].toDouble();
        const bool relu = ctx->saved_data["relu"].toBool();
        const at::Tensor dy = grads[0].contiguous();
        const c10::cuda::CUDAGuard guard(x.device());
        cudaStream_t stream = at::cuda::getCurrentCUDAStream();
        const int N = (int)x.size(0), C = (int)x.size(1), H = (int)x.size(2), W = (int)x.size(3);
        at::Tensor pg = f32_buffer(x, 4 * (int64_t)C);
        at::Tensor ws = f32_buffer(x, (int64_t)cnsn_site_workspace_floats(N, C));
        at::Tensor dx = at::empty_like(x);
        const cnsn_gate_params g = gate_params(w, gamma, beta, nullptr);
        float* p = pg.data_ptr<float>();
        const cnsn_gate_grads gg{p, p + 2 * C, p + 3 * C};
        check(cnsn_site_bwd(x.data_ptr(), dy.data_ptr(), dx.data_ptr(), dtype_code(x), N, C, H, W, perm.data_ptr<int>(), cw.v, sw.v,
                            (float)lam, (float)cn_eps, relu ? 1 : 0, &g, save.data_ptr<float>(), &gg, ws.data_ptr<float>(), stream));
        at::Tensor none;
        return {like_input(dx, ctx->saved_data["cl_in"].toBool()), none, none, none, none, none, pg.narrow(0, 0, 2 * C).view_as(w),
                pg.narrow(0, 2 * C, C), pg.narrow(0, 3 * C, C), none, none, none, none};
    }
};


// ---- IBN / InstanceNorm2d / BatchNorm2d (cnsn_ibn_fwd/_bwd: half == C instance norm, half == 0 batch norm) --------
struct IbnNode : public torch::autograd::Function<IbnNode> {
    static at::Tensor forward(AutogradContext* ctx, const at::Tensor& x_in, int64_t half, bool training, bool relu, double momentum,
                              double eps_in, double eps_bn, GateBufs bufs, const c10::optional<at::Tensor>& in_w,
                              const c10::optional<at::Tensor>& in_b, const c10::optional<at::Tensor>& bn_w,
                              const c10::optional<at::Tensor>& bn_b) {
        require_cuda4(x_in);
        const at::Tensor x = x_in.contiguous();
        const c10::cuda::CUDAGuard guard(x.device());
        cudaStream_t stream = at::cuda::getCurrentCUDAStream();
        const int N = (int)x.size(0), C = (int)x.size(1), H = (int)x.size(2), W = (int)x.size(3);
        auto f32p = [](const c10::optional<at::Tensor>& t) -> const float* {
            return (t.has_value() && t->defined()) ? t->data_ptr<float>() : nullptr;
        };
        cnsn_ibn_params p{};
        p.in_w = f32p(in_w); p.in_b = f32p(in_b); p.bn_w = f32p(bn_w); p.bn_b = f32p(bn_b);
        p.run_mean = bufs.run_mean.defined() ? bufs.run_mean.data_ptr<float>() : nullptr;
        p.run_var = bufs.run_var.defined() ? bufs.run_var.data_ptr<float>() : nullptr;
        p.nbt = bufs.nbt.defined() ? reinterpret_cast<long long*>(bufs.nbt.data_ptr<int64_t>()) : nullptr;
        at::Tensor save = f32_buffer(x, (int64_t)cnsn_ibn_save_floats(N, C, (int)half));
        at::Tensor y = at::empty_like(x);
        check(cnsn_ibn_fwd(x.data_ptr(), y.data_ptr(), dtype_code(x), N, C, H, W, (int)half, &p, training ? 1 : 0, relu ? 1 : 0,
                           (float)momentum, (float)eps_in, (float)eps_bn, save.data_ptr<float>(), stream));
        at::Tensor none;
        auto opt = [&](const c10::optional<at::Tensor>& t) { return (t.has_value() && t->defined()) ? *t : none; };
        ctx->save_for_backward({x, save, opt(in_w), opt(bn_w), opt(in_b), opt(bn_b)});
        ctx->saved_data["half"] = half;
        ctx->saved_data["training"] = training;
        ctx->saved_data["relu"] = relu;
        return y;
    }

    static variable_list backward(AutogradContext* ctx, variable_list grads) {
        const auto saved = ctx->get_saved_variables();
        const at::Tensor &x = saved[0], &save = saved[1], &in_w = saved[2], &bn_w = saved[3], &in_b = saved[4], &bn_b = saved[5];
        const int half = (int)ctx->saved_data["half"].toInt();
        const bool training = ctx->saved_data["training"].toBool(), relu = ctx->saved_data["relu"].toBool();
        const at::Tensor dy = grads[0].contiguous();
        const c10::cuda::CUDAGuard guard(x.device());
        cudaStream_t stream = at::cuda::getCurrentCUDAStream();
        const int N = (int)x.size(0), C = (int)x.size(1), H = (int)x.size(2), W = (int)x.size(3);
        at::Tensor pg = f32_buffer(x, 2 * (int64_t)C);                // d_in_w (half) | d_in_b (half) | d_bn_w (C-half) | d_bn_b (C-half)
        at::Tensor ws = f32_buffer(x, (int64_t)cnsn_ibn_workspace_floats(N, C));
        at::Tensor dx = at::empty_like(x);
        cnsn_ibn_params p{};
        p.in_w = in_w.defined() ? in_w.data_ptr<float>() : nullptr;
        p.bn_w = bn_w.defined() ? bn_w.data_ptr<float>() : nullptr;
        p.in_b = in_b.defined() ? in_b.data_ptr<float>() : nullptr;
        p.bn_b = bn_b.defined() ? bn_b.data_ptr<float>() : nullptr;
        float* g = pg.data_ptr<float>();
        check(cnsn_ibn_bwd(x.data_ptr(), dy.data_ptr(), dx.data_ptr(), dtype_code(x), N, C, H, W, half, &p, training ? 1 : 0,
                           relu ? 1 : 0, save.data_ptr<float>(), g, g + half, g + 2 * half, g + C + half, ws.data_ptr<float>(), stream));
        at::Tensor none;
        const int nb = C - half;
        return {dx, none, none, none, none, none, none, none,
                half > 0 ? pg.narrow(0, 0, half) : none, half > 0 ? pg.narrow(0, half, half) : none,
                nb > 0 ? pg.narrow(0, 2 * half, nb) : none, nb > 0 ? pg.narrow(0, C + half, nb) : none};
    }
};

// ---- channels-last BatchNorm2d [+ ReLU] (cnsn_bn_nhwc_fwd/_bwd) -----------------------------------------------------------
struct BnNhwcNode : public torch::autograd::Function<BnNhwcNode> {
    static at::Tensor forward(AutogradContext* ctx, const at::Tensor& x, bool training, bool relu, double momentum, double eps,
                              GateBufs bufs, const at::Tensor& weight, const at::Tensor& bias) {
        const c10::cuda::CUDAGuard guard(x.device());
        cudaStream_t stream = at::cuda::getCurrentCUDAStream();
        const int N = (int)x.size(0), C = (int)x.size(1), H = (int)x.size(2), W = (int)x.size(3);
        at::Tensor save = f32_buffer(x, (int64_t)cnsn_bn_nhwc_save_floats(dtype_code(x), N, C, H, W));
        at::Tensor y = at::empty_like(x);
        check(cnsn_bn_nhwc_fwd(x.data_ptr(), y.data_ptr(), dtype_code(x), N, C, H, W, weight.data_ptr<float>(), bias.data_ptr<float>(),
                               bufs.run_mean.data_ptr<float>(), bufs.run_var.data_ptr<float>(),
                               bufs.nbt.defined() ? reinterpret_cast<long long*>(bufs.nbt.data_ptr<int64_t>()) : nullptr,
                               training ? 1 : 0, relu ? 1 : 0, (float)momentum, (float)eps, save.data_ptr<float>(), stream));
        ctx->save_for_backward({x, weight, save});
        ctx->saved_data["training"] = training;
        ctx->saved_data["relu"] = relu;
        return y;
    }

    static variable_list backward(AutogradContext* ctx, variable_list grads) {
        const auto saved = ctx->get_saved_variables();
        const at::Tensor &x = saved[0], &weight = saved[1], &save = saved[2];
        const bool training = ctx->saved_data["training"].toBool(), relu = ctx->saved_data["relu"].toBool();
        const at::Tensor dy = grads[0].contiguous(at::MemoryFormat::ChannelsLast);
        const c10::cuda::CUDAGuard guard(x.device());
        cudaStream_t stream = at::cuda::getCurrentCUDAStream();
        const int N = (int)x.size(0), C = (int)x.size(1), H = (int)x.size(2), W = (int)x.size(3);
        at::Tensor pg = f32_buffer(x, 2 * (int64_t)C);                // dgamma | dbeta
        at::Tensor ws = f32_buffer(x, (int64_t)cnsn_bn_nhwc_workspace_floats(dtype_code(x), N, C, H, W));
        at::Tensor dx = at::empty_like(x);
        float* g = pg.data_ptr<float>();
        check(cnsn_bn_nhwc_bwd(x.data_ptr(), dy.data_ptr(), dx.data_ptr(), dtype_code(x), N, C, H, W, weight.data_ptr<float>(),
                               training ? 1 : 0, relu ? 1 : 0, save.data_ptr<float>(), g, g + C, ws.data_ptr<float>(), stream));
        at::Tensor none;
        return {dx, none, none, none, none, none, pg.narrow(0, 0, C), pg.narrow(0, C, C)};
    }
};

// ---- channels-last MaxPool2d (cnsn_maxpool_nhwc_fwd/_bwd) ---------------------------------------------------------------
struct MaxPoolNhwcNode : public torch::autograd::Function<MaxPoolNhwcNode> {
    static at::Tensor forward(AutogradContext* ctx, const at::Tensor& x, int64_t k, int64_t stride, int64_t pad) {
        const c10::cuda::CUDAGuard guard(x.device());
        cudaStream_t stream = at::cuda::getCurrentCUDAStream();
        const int N = (int)x.size(0), C = (int)x.size(1), H = (int)x.size(2), W = (int)x.size(3);
        int OH = 0, OW = 0;
        check(cnsn_maxpool_nhwc_out(H, W, (int)k, (int)stride, (int)pad, &OH, &OW));
        at::Tensor y = at::empty({N, C, OH, OW}, x.options().memory_format(at::MemoryFormat::ChannelsLast));
        at::Tensor code = at::empty({N, OH, OW, C}, x.options().dtype(at::kByte).memory_format(at::MemoryFormat::Contiguous));
        check(cnsn_maxpool_nhwc_fwd(x.data_ptr(), y.data_ptr(), code.data_ptr<uint8_t>(), dtype_code(x), N, C, H, W, (int)k, (int)stride,
                                    (int)pad, stream));
        ctx->save_for_backward({code});
        ctx->saved_data["shape"] = std::vector<int64_t>{N, C, H, W};
        ctx->saved_data["k"] = k; ctx->saved_data["stride"] = stride; ctx->saved_data["pad"] = pad;
        return y;
    }

    static variable_list backward(AutogradContext* ctx, variable_list grads) {
        const at::Tensor code = ctx->get_saved_variables()[0];
        const auto shape = ctx->saved_data["shape"].toIntVector();
        const int k = (int)ctx->saved_data["k"].toInt(), stride = (int)ctx->saved_data["stride"].toInt(), pad = (int)ctx->saved_data["pad"].toInt();
        at::Tensor dy = grads[0];
        if (!is_channels_last(dy)) dy = dy.contiguous().contiguous(at::MemoryFormat::ChannelsLast);
        const c10::cuda::CUDAGuard guard(dy.device());
        cudaStream_t stream = at::cuda::getCurrentCUDAStream();
        at::Tensor dx = at::empty(shape, dy.options().memory_format(at::MemoryFormat::ChannelsLast));
        check(cnsn_maxpool_nhwc_bwd(dy.data_ptr(), code.data_ptr<uint8_t>(), dx.data_ptr(), dtype_code(dy), (int)shape[0], (int)shape[1],
                                    (int)shape[2], (int)shape[3], k, stride, pad, stream));
        at::Tensor none;
        return {dx, none, none, none};
    }
};

// ---- fused bottleneck tail: bn3 -> + identity -> SelfNorm -> ReLU, channels-last (cnsn_bn_selfnorm_tail_*_nhwc) -----------
struct TailNode : public torch::autograd::Function<TailNode> {
    static at::Tensor forward(AutogradContext* ctx, const at::Tensor& c, const at::Tensor& res_in, bool relu,
                              const at::Tensor& bn_w, const at::Tensor& bn_b, GateBufs bn_bufs, bool bn_training, double bn_momentum,
                              double bn_eps, const at::Tensor& w, const at::Tensor& gamma, const at::Tensor& beta, GateBufs bufs,
                              bool training, double momentum, double sn_bn_eps, double eps) {
        const at::Tensor res = res_in.contiguous(at::MemoryFormat::ChannelsLast);
        const c10::cuda::CUDAGuard guard(c.device());
        cudaStream_t stream = at::cuda::getCurrentCUDAStream();
        const int N = (int)c.size(0), C = (int)c.size(1), H = (int)c.size(2), W = (int)c.size(3);
        at::Tensor bn_save = f32_buffer(c, (int64_t)cnsn_bn_nhwc_save_floats(dtype_code(c), N, C, H, W));
        at::Tensor sn_save = f32_buffer(c, (int64_t)cnsn_selfnorm_nhwc_save_floats(dtype_code(c), N, C, H, W));
        at::Tensor z = at::empty_like(c), y = at::empty_like(c);
        const cnsn_gate_params g = gate_params(w, gamma, beta, &bufs);
        check(cnsn_bn_selfnorm_tail_fwd_nhwc(c.data_ptr(), res.data_ptr(), z.data_ptr(), y.data_ptr(), relu ? 1 : 0, dtype_code(c), N, C, H, W,
                                             bn_w.data_ptr<float>(), bn_b.data_ptr<float>(), bn_bufs.run_mean.data_ptr<float>(),
                                             bn_bufs.run_var.data_ptr<float>(),
                                             bn_bufs.nbt.defined() ? reinterpret_cast<long long*>(bn_bufs.nbt.data_ptr<int64_t>()) : nullptr,
                                             bn_training ? 1 : 0, (float)bn_momentum, (float)bn_eps, bn_save.data_ptr<float>(), &g,
                                             training ? 1 : 0, (float)momentum, (float)sn_bn_eps, (float)eps, sn_save.data_ptr<float>(), stream));
        ctx->save_for_backward({c, z, bn_w, w, gamma, beta, bn_save, sn_save});
        ctx->saved_data["relu"] = relu;
        ctx->saved_data["bn_training"] = bn_training;
        ctx->saved_data["training"] = training;
        return y;
    }

    static variable_list backward(AutogradContext* ctx, variable_list grads) {
        const auto saved = ctx->get_saved_variables();
        const at::Tensor &c = saved[0], &z = saved[1], &bn_w = saved[2], &w = saved[3], &gamma = saved[4], &beta = saved[5];
        const at::Tensor &bn_save = saved[6], &sn_save = saved[7];
        const bool relu = ctx->saved_data["relu"].toBool(), bn_training = ctx->saved_data["bn_training"].toBool();
        const bool training = ctx->saved_data["training"].toBool();
        const at::Tensor dy = grads[0].contiguous(at::MemoryFormat::ChannelsLast);
        const c10::cuda::CUDAGuard guard(c.device());
        cudaStream_t stream = at::cuda::getCurrentCUDAStream();
        const int N = (int)c.size(0), C = (int)c.size(1), H = (int)c.size(2), W = (int)c.size(3);
        at::Tensor pg = f32_buffer(c, 6 * (int64_t)C);                // dw (C,2) | dgamma | dbeta | d_bn_w | d_bn_b
        at::Tensor bn_ws = f32_buffer(c, (int64_t)cnsn_bn_nhwc_workspace_floats(dtype_code(c), N, C, H, W));
        at::Tensor sn_ws = f32_buffer(c, (int64_t)cnsn_selfnorm_nhwc_workspace_floats(dtype_code(c), N, C, H, W));
        at::Tensor dz = at::empty_like(c), dc = at::empty_like(c);
        const cnsn_gate_params g = gate_params(w, gamma, beta, nullptr);
        float* p = pg.data_ptr<float>();
        const cnsn_gate_grads gg{p, p + 2 * C, p + 3 * C};
        check(cnsn_bn_selfnorm_tail_bwd_nhwc(c.data_ptr(), z.data_ptr(), dy.data_ptr(), dz.data_ptr(), dc.data_ptr(), relu ? 1 : 0,
                                             dtype_code(c), N, C, H, W, bn_w.data_ptr<float>(), bn_training ? 1 : 0,
                                             bn_save.data_ptr<float>(), p + 4 * C, p + 5 * C, bn_ws.data_ptr<float>(), &g,
                                             training ? 1 : 0, sn_save.data_ptr<float>(), &gg, sn_ws.data_ptr<float>(), stream));
        at::Tensor none;
        return {dc, dz, none, pg.narrow(0, 4 * C, C), pg.narrow(0, 5 * C, C), none, none, none, none,
                pg.narrow(0, 0, 2 * C).view_as(w), pg.narrow(0, 2 * C, C), pg.narrow(0, 3 * C, C), none, none, none, none, none};
    }
};

void check_gate(const at::Tensor& x, const at::Tensor& w, const at::Tensor& gamma, const at::Tensor& beta, const GateBufs& b) {
    const int64_t C = x.size(1);
    for (const at::Tensor* t : {&w, &gamma, &beta, &b.run_mean, &b.run_var}) {
        TORCH_CHECK(t->defined() && t->is_cuda() && t->scalar_type() == at::kFloat && t->is_contiguous() && t->device() == x.device(),
                    "cnsn_b200 fast path: gate parameters and buffers must be fp32, contiguous and on x's device");
    }
    TORCH_CHECK(w.numel() == 2 * C && gamma.numel() == C && beta.numel() == C && b.run_mean.numel() == C && b.run_var.numel() == C,
                "cnsn_b200: gate parameter shapes do not match the channel count");
    TORCH_CHECK(!b.nbt.defined() || (b.nbt.scalar_type() == at::kLong && b.nbt.is_cuda()), "num_batches_tracked must be an int64 CUDA tensor");
}

// ---- Python entry points ---------------------------------------------------------------------------------------------
at::Tensor selfnorm(const at::Tensor& x, const c10::optional<at::Tensor>& residual, bool relu, const at::Tensor& w,
                    const at::Tensor& gamma, const at::Tensor& beta, const at::Tensor& run_mean, const at::Tensor& run_var,
                    const c10::optional<at::Tensor>& nbt, bool training, double momentum, double bn_eps, double eps) {
    require_cuda4(x);
    GateBufs b{run_mean, run_var, nbt.has_value() ? *nbt : at::Tensor()};
    check_gate(x, w, gamma, beta, b);
    const bool block = relu || (residual.has_value() && residual->defined());
    return SelfNormNode::apply(x, residual, relu, block, w, gamma, beta, b, training, momentum, bn_eps, eps);
}

at::Tensor crossnorm(const at::Tensor& x, const std::vector<int64_t>& cwin, const std::vector<int64_t>& swin, double lam, double eps) {
    return CrossNormNode::apply(x, window(cwin), window(swin), lam, eps);
}

at::Tensor site(const at::Tensor& x, const std::vector<int64_t>& cwin, const std::vector<int64_t>& swin, double lam, double cn_eps,
                bool relu, const at::Tensor& w, const at::Tensor& gamma, const at::Tensor& beta, const at::Tensor& run_mean,
                const at::Tensor& run_var, const c10::optional<at::Tensor>& nbt, double momentum, double bn_eps, double sn_eps) {
    require_cuda4(x);
    GateBufs b{run_mean, run_var, nbt.has_value() ? *nbt : at::Tensor()};
    check_gate(x, w, gamma, beta, b);
    return SiteNode::apply(x, window(cwin), window(swin), lam, cn_eps, relu, w, gamma, beta, b, momentum, bn_eps, sn_eps);
}

at::Tensor bn_nhwc(const at::Tensor& x, bool training, bool relu, double momentum, double eps, const at::Tensor& run_mean,
                   const at::Tensor& run_var, const c10::optional<at::Tensor>& nbt, const at::Tensor& weight, const at::Tensor& bias) {
    require_cuda4(x);
    TORCH_CHECK(is_channels_last(x) && (reinterpret_cast<uintptr_t>(x.data_ptr()) & 15u) == 0, "bn_nhwc: x must be a dense, 16-byte aligned channels_last tensor");
    const int64_t C = x.size(1);
    for (const at::Tensor* t : {&weight, &bias, &run_mean, &run_var}) {
        TORCH_CHECK(t->defined() && t->is_cuda() && t->scalar_type() == at::kFloat && t->is_contiguous() && t->numel() == C &&
                    t->device() == x.device(), "bn_nhwc: weight, bias and running statistics must be fp32 [C] on x's device");
    }
    GateBufs b{run_mean, run_var, nbt.has_value() ? *nbt : at::Tensor()};
    return BnNhwcNode::apply(x, training, relu, momentum, eps, b, weight, bias);
}

at::Tensor maxpool_nhwc(const at::Tensor& x, int64_t k, int64_t stride, int64_t pad) {
    require_cuda4(x);
    TORCH_CHECK(is_channels_last(x) && (reinterpret_cast<uintptr_t>(x.data_ptr()) & 15u) == 0, "maxpool_nhwc: x must be a dense, 16-byte aligned channels_last tensor");
    return MaxPoolNhwcNode::apply(x, k, stride, pad);
}

bool bn_sn_tail_supported(const at::Tensor& c) {
    if (!c.is_cuda() || !is_channels_last(c) || (reinterpret_cast<uintptr_t>(c.data_ptr()) & 15u)) return false;
    const c10::cuda::CUDAGuard guard(c.device());
    return cnsn_bn_selfnorm_tail_supported(dtype_code(c), (int)c.size(0), (int)c.size(1), (int)c.size(2), (int)c.size(3)) != 0;
}

at::Tensor bn_sn_tail(const at::Tensor& c, const at::Tensor& res, bool relu, const at::Tensor& bn_w, const at::Tensor& bn_b,
                      const at::Tensor& bn_rm, const at::Tensor& bn_rv, const c10::optional<at::Tensor>& bn_nbt, bool bn_training,
                      double bn_momentum, double bn_eps, const at::Tensor& w, const at::Tensor& gamma, const at::Tensor& beta,
                      const at::Tensor& run_mean, const at::Tensor& run_var, const c10::optional<at::Tensor>& nbt, bool training,
                      double momentum, double sn_bn_eps, double eps) {
    require_cuda4(c);
    TORCH_CHECK(bn_sn_tail_supported(c), "bn_sn_tail: c must be a dense, 16-byte aligned channels_last tensor of a supported shape");
    TORCH_CHECK(res.sizes() == c.sizes() && res.scalar_type() == c.scalar_type() && res.is_cuda(), "bn_sn_tail: the residual must match c");
    const int64_t C = c.size(1);
    for (const at::Tensor* t : {&bn_w, &bn_b, &bn_rm, &bn_rv}) {
        TORCH_CHECK(t->defined() && t->is_cuda() && t->scalar_type() == at::kFloat && t->is_contiguous() && t->numel() == C &&
                    t->device() == c.device(), "bn_sn_tail: batch-norm weight, bias and running statistics must be fp32 [C] on c's device");
    }
    GateBufs bnb{bn_rm, bn_rv, bn_nbt.has_value() ? *bn_nbt : at::Tensor()};
    GateBufs b{run_mean, run_var, nbt.has_value() ? *nbt : at::Tensor()};
    check_gate(c, w, gamma, beta, b);
    return TailNode::apply(c, res, relu, bn_w, bn_b, bnb, bn_training, bn_momentum, bn_eps, w, gamma, beta, b, training, momentum, sn_bn_eps, eps);
}

bool site_supported(const at::Tensor& x) {
    if (!x.is_cuda() || x.dim() != 4 || (reinterpret_cast<uintptr_t>(x.data_ptr()) & 15u)) return false;
    const c10::cuda::CUDAGuard guard(x.device());
    return cnsn_site_supported(dtype_code(x), (int)x.size(0), (int)x.size(1), (int)x.size(2), (int)x.size(3)) != 0;
}

at::Tensor ibn(const at::Tensor& x, int64_t half, bool training, bool relu, double momentum, double eps_in, double eps_bn,
               const c10::optional<at::Tensor>& run_mean, const c10::optional<at::Tensor>& run_var,
               const c10::optional<at::Tensor>& nbt, const c10::optional<at::Tensor>& in_w, const c10::optional<at::Tensor>& in_b,
               const c10::optional<at::Tensor>& bn_w, const c10::optional<at::Tensor>& bn_b) {
    require_cuda4(x);
    const int64_t C = x.size(1);
    TORCH_CHECK(half >= 0 && half <= C, "half must be in [0, C]");
    auto ok = [&](const c10::optional<at::Tensor>& t, int64_t n, const char* what) {
        TORCH_CHECK(t.has_value() && t->defined() && t->is_cuda() && t->scalar_type() == at::kFloat && t->is_contiguous() &&
                    t->numel() == n && t->device() == x.device(), "cnsn_b200 fast path: ", what, " must be fp32, contiguous, on x's device, with ", n, " elements");
    };
    if (half > 0) { ok(in_w, half, "IN.weight"); ok(in_b, half, "IN.bias"); }
    if (half < C) { ok(bn_w, C - half, "BN.weight"); ok(bn_b, C - half, "BN.bias"); ok(run_mean, C - half, "BN.running_mean"); ok(run_var, C - half, "BN.running_var"); }
    GateBufs b{run_mean.has_value() ? *run_mean : at::Tensor(), run_var.has_value() ? *run_var : at::Tensor(),
               nbt.has_value() ? *nbt : at::Tensor()};
    return IbnNode::apply(x, half, training, relu, momentum, eps_in, eps_bn, b, in_w, in_b, bn_w, bn_b);
}

bool ibn_resident(const at::Tensor& x, int64_t half, bool training) {
    if (!x.is_cuda() || x.dim() != 4 || (reinterpret_cast<uintptr_t>(x.data_ptr()) & 15u)) return false;
    const c10::cuda::CUDAGuard guard(x.device());
    return cnsn_ibn_resident(dtype_code(x), (int)x.size(0), (int)x.size(1), (int)x.size(2), (int)x.size(3), (int)half, training ? 1 : 0) != 0;
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.doc() = "cnsn_b200: C++ autograd nodes above the C ABI of libcnsn_b200.so";
    m.def("abi_version", [] { return cnsn_version(); });
    m.def("selfnorm", &selfnorm, "relu?(SelfNorm(x [+ residual])), single gate (models/cnsn.py:130-150)");
    m.def("crossnorm", &crossnorm, "cn_op_2ins_space_chan without channel permutation; draws torch.randperm(N) itself (models/cnsn.py:58-91)");
    m.def("site", &site, "relu?(SelfNorm(CrossNorm(x))) as one kernel per direction (models/cnsn.py:159-164)");
    m.def("site_supported", &site_supported);
    m.def("ibn", &ibn, "IBN / InstanceNorm2d (half == C) / BatchNorm2d (half == 0), one kernel per direction (resnet_ibn_cnsn.py:24-44)");
    m.def("ibn_resident", &ibn_resident);
    m.def("bn_sn_tail", &bn_sn_tail, "relu?(SelfNorm(bn(c) + residual)) on channels_last tensors: the tail of a pos='post' ResNet bottleneck");
    m.def("bn_sn_tail_supported", &bn_sn_tail_supported);
    m.def("maxpool_nhwc", &maxpool_nhwc, "nn.MaxPool2d(k, stride, pad) on a dense channels_last tensor (csrc/pool_nhwc.cu)");
    m.def("bn_nhwc", &bn_nhwc, "nn.BatchNorm2d [+ ReLU] on a dense channels_last tensor, three kernels per direction (csrc/bn_nhwc.cu)");
}
