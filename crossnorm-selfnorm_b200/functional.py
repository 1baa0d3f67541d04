"""torch.autograd.Function wrappers: one fused forward and one fused backward per operator.

Saved for backward: the input tensor plus O(N*C) fp32 statistics -- not the full-size
intermediates (normalised features, expanded stds, masks, gathered copies) the reference's eager
autograd graph keeps alive.
"""
import torch

from . import _lib


def _dense(x):
    # the reference forces .contiguous() itself (models/cnsn.py:14,16); kernels take dense NCHW
    return x if x.is_contiguous() else x.contiguous()


def _is_cl(x):
    return x.dim() == 4 and not x.is_contiguous() and x.is_contiguous(memory_format=torch.channels_last)


def _like_input(t, cl):
    """A channels_last caller gets channels_last back from the operators whose kernels are NCHW (CrossNorm, the fused
    site), forward and backward: what follows (cuDNN's NHWC convolutions) keeps seeing one layout."""
    return t.contiguous(memory_format=torch.channels_last) if cl else t


def _dense_or_nhwc(x):
    """Dense NCHW, or -- where the backend has channels-last kernels for the shape -- the channels_last tensor itself."""
    if x.is_contiguous():
        return x
    is_nhwc = getattr(_lib.backend(), "is_nhwc", None)
    return x if (is_nhwc is not None and x.is_cuda and is_nhwc(x)) else x.contiguous()


class InstanceStats(torch.autograd.Function):
    """(mean, std) over a window -- calc_ins_mean_std, models/cnsn.py:8-17."""

    @staticmethod
    def forward(ctx, x, window, eps):
        # strided views (crops, transposes, channels_last) are reduced where they lie: no dense copy in the forward
        if not getattr(_lib.backend(), "strided_stats", True) or any(s < 0 for s in x.stride()):
            x = _dense(x)
        mean, std = _lib.backend().instance_stats(x, window, eps)
        ctx.save_for_backward(x, mean, std)
        ctx.window = window
        return mean, std

    @staticmethod
    def backward(ctx, dmean, dstd):
        x, mean, std = ctx.saved_tensors
        dx = _lib.backend().instance_stats_bwd(_dense(x), ctx.window, mean, std,
                                               dmean.contiguous().float(), dstd.contiguous().float())
        return dx, None, None


class InstanceAffine(torch.autograd.Function):
    """out = x*scale[n,c] + shift[n,c] -- the broadcast restyle of models/cnsn.py:27-29."""

    @staticmethod
    def forward(ctx, x, scale, shift):
        x = _dense(x)
        scale = scale.contiguous().float()
        shift = shift.contiguous().float()
        ctx.save_for_backward(x, scale)
        return _lib.backend().instance_affine(x, scale, shift)

    @staticmethod
    def backward(ctx, dy):
        x, scale = ctx.saved_tensors
        dy = _dense(dy)
        b = _lib.backend()
        dscale, dshift = b.instance_dot(x, dy)
        dx = b.instance_affine(dy, scale, torch.zeros_like(scale))
        return dx, dscale, dshift


class SelfNormFn(torch.autograd.Function):
    """SelfNorm.forward (models/cnsn.py:130-150) and its backward."""

    @staticmethod
    def forward(ctx, x, training, momentum, bn_eps, eps, g_bufs, f_bufs,
                g_w, g_gamma, g_beta, f_w=None, f_gamma=None, f_beta=None):
        x = _dense(x)
        g = _lib.GateTensors(g_w, g_gamma, g_beta, *g_bufs)
        f = _lib.GateTensors(f_w, f_gamma, f_beta, *f_bufs) if f_w is not None else None
        y, save = _lib.backend().selfnorm_fwd(x, g, f, training, momentum, bn_eps, eps)
        ctx.save_for_backward(x, g_w, g_gamma, g_beta, f_w, f_gamma, f_beta)
        ctx.sn_save = save          # fp32 statistics block (never a graph input)
        ctx.training = training
        return y

    @staticmethod
    def backward(ctx, dy):
        x, g_w, g_gamma, g_beta, f_w, f_gamma, f_beta = ctx.saved_tensors
        save = ctx.sn_save
        g = _lib.GateTensors(g_w, g_gamma, g_beta, None, None, None)
        f = _lib.GateTensors(f_w, f_gamma, f_beta, None, None, None) if f_w is not None else None
        dx, gg, gf = _lib.backend().selfnorm_bwd(x, _dense(dy), g, f, ctx.training, save)
        out = [dx, None, None, None, None, None, None,
               gg[0].view_as(g_w).to(g_w.dtype), gg[1].to(g_gamma.dtype), gg[2].to(g_beta.dtype)]
        if f is not None:
            out += [gf[0].view_as(f_w).to(f_w.dtype), gf[1].to(f_gamma.dtype), gf[2].to(f_beta.dtype)]
        else:
            out += [None, None, None]
        return tuple(out)


class SelfNormBlockFn(torch.autograd.Function):
    """relu?(SelfNorm(x + res)) in one forward and one backward call: the residual add and
    the ReLU around a pos='post' site of models/imagenet/resnet_cnsn.py:117-122 /
    models/cifar/wideresnet_cnsn.py:93-96.  Saved for backward: the sum z and the fp32 statistics block."""

    @staticmethod
    def forward(ctx, x, res, relu, training, momentum, bn_eps, eps, g_bufs, g_w, g_gamma, g_beta):
        x = _dense_or_nhwc(x)
        if res is not None:
            res = _dense(res) if x.is_contiguous() else res      # a channels_last x: the backend brings res into its layout
        g = _lib.GateTensors(g_w, g_gamma, g_beta, *g_bufs)
        y, z, save = _lib.backend().selfnorm_block_fwd(x, res, relu, g, training, momentum, bn_eps, eps)
        ctx.save_for_backward(z, g_w, g_gamma, g_beta)
        ctx.sn_save = save
        ctx.meta = (bool(relu), training, res is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        z, g_w, g_gamma, g_beta = ctx.saved_tensors
        relu, training, has_res = ctx.meta
        g = _lib.GateTensors(g_w, g_gamma, g_beta, None, None, None)
        dz, gg = _lib.backend().selfnorm_block_bwd(z, dy if not z.is_contiguous() else _dense(dy), relu, g, training, ctx.sn_save)
        return (dz, dz if has_res else None, None, None, None, None, None, None,
                gg[0].view_as(g_w).to(g_w.dtype), gg[1].to(g_gamma.dtype), gg[2].to(g_beta.dtype))


class CrossNormFn(torch.autograd.Function):
    """Device half of cn_op_2ins_space_chan (models/cnsn.py:58-91) and its backward."""

    @staticmethod
    def forward(ctx, x, perm, chan_perm, cwin, swin, lam, eps):
        ctx.cl_in = _is_cl(x)
        x = _dense(x)
        y, save = _lib.backend().crossnorm_fwd(x, perm, chan_perm, cwin, swin, lam, eps)
        ctx.save_for_backward(x)
        ctx.cn_save = (perm, chan_perm, save)
        ctx.meta = (cwin, swin, lam)
        return _like_input(y, ctx.cl_in)

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        perm, chan_perm, save = ctx.cn_save
        cwin, swin, lam = ctx.meta
        dx = _lib.backend().crossnorm_bwd(x, _dense(dy), perm, chan_perm, cwin, swin, lam, save)
        return _like_input(dx, ctx.cl_in), None, None, None, None, None, None


class CnsnSiteFn(torch.autograd.Function):
    """A CNSN site whose CrossNorm and SelfNorm both fire (CNSN.forward, models/cnsn.py:159-164) as one forward
    and one backward kernel.  Saved for backward: x, the permutation and O(N*C) fp32 statistics;
    the CrossNorm output is never materialised."""

    @staticmethod
    def forward(ctx, x, perm, cwin, swin, lam, cn_eps, relu, momentum, bn_eps, sn_eps, g_bufs, g_w, g_gamma, g_beta):
        ctx.cl_in = _is_cl(x)
        x = _dense(x)
        g = _lib.GateTensors(g_w, g_gamma, g_beta, *g_bufs)
        y, save = _lib.backend().site_fwd(x, perm, cwin, swin, lam, cn_eps, g, momentum, bn_eps, sn_eps, relu)
        ctx.save_for_backward(x, g_w, g_gamma, g_beta)
        ctx.site = (perm, cwin, swin, lam, cn_eps, bool(relu), save)
        return _like_input(y, ctx.cl_in)

    @staticmethod
    def backward(ctx, dy):
        x, g_w, g_gamma, g_beta = ctx.saved_tensors
        perm, cwin, swin, lam, cn_eps, relu, save = ctx.site
        g = _lib.GateTensors(g_w, g_gamma, g_beta, None, None, None)
        dx, gg = _lib.backend().site_bwd(x, _dense(dy), perm, cwin, swin, lam, cn_eps, g, save, relu)
        return (_like_input(dx, ctx.cl_in), None, None, None, None, None, None, None, None, None, None,
                gg[0].view_as(g_w).to(g_w.dtype), gg[1].to(g_gamma.dtype), gg[2].to(g_beta.dtype))


class JsdConsistencyFn(torch.autograd.Function):
    """The Jensen-Shannon consistency term of the 3-view steps (imagenet.py:367-376, cifar.py:173-182)."""

    @staticmethod
    def forward(ctx, z0, z1, z2):
        z0, z1, z2 = (_dense(z) for z in (z0, z1, z2))
        ctx.save_for_backward(z0, z1, z2)
        return _lib.backend().jsd_fwd(z0, z1, z2)

    @staticmethod
    def backward(ctx, gout):
        z0, z1, z2 = ctx.saved_tensors
        d = _lib.backend().jsd_bwd(z0, z1, z2, gout.contiguous().float())
        return tuple(d)


class IbnFn(torch.autograd.Function):
    """IBN.forward (models/imagenet/resnet_ibn_cnsn.py:38-44) and its backward, one kernel each; ``relu``: the ReLU that
    follows the norm in the host blocks, in the same kernels."""

    @staticmethod
    def forward(ctx, x, half, training, momentum, eps_in, eps_bn, bufs, in_w, in_b, bn_w, bn_b, relu=False):
        x = _dense(x)
        p = {"in_w": in_w, "in_b": in_b, "bn_w": bn_w, "bn_b": bn_b, "run_mean": bufs[0], "run_var": bufs[1], "nbt": bufs[2]}
        y, save = _lib.backend().ibn_fwd(x, half, p, training, momentum, eps_in, eps_bn, relu)
        ctx.save_for_backward(x, in_w, bn_w, in_b, bn_b)
        ctx.ibn = (half, training, save, bool(relu))
        return y

    @staticmethod
    def backward(ctx, dy):
        x, in_w, bn_w, in_b, bn_b = ctx.saved_tensors
        half, training, save, relu = ctx.ibn
        dx, g = _lib.backend().ibn_bwd(x, _dense(dy), half, {"in_w": in_w, "bn_w": bn_w, "in_b": in_b, "bn_b": bn_b}, training,
                                       save, relu)
        gi = (g[0].to(in_w.dtype), g[1].to(in_w.dtype)) if in_w is not None else (None, None)
        gb = (g[2].to(bn_w.dtype), g[3].to(bn_w.dtype)) if bn_w is not None else (None, None)   # half == C: no BN half
        return (dx, None, None, None, None, None, None) + gi + gb + (None,)


class BnNhwcFn(torch.autograd.Function):
    """nn.BatchNorm2d [+ ReLU] on a dense channels_last tensor (cnsn_bn_nhwc_fwd / _bwd, csrc/bn_nhwc.cu); output and input
    gradient stay channels_last.  Saved for backward: x and O(C) fp32 statistics (not the normalised tensor, not the mask)."""

    @staticmethod
    def forward(ctx, x, training, relu, momentum, eps, bufs, weight, bias):
        y, save = _lib.backend().bn_nhwc_fwd(x, weight, bias, bufs[0], bufs[1], bufs[2], training, relu, momentum, eps)
        ctx.save_for_backward(x, weight)
        ctx.bn = (training, bool(relu), save)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        training, relu, save = ctx.bn
        if not dy.is_contiguous(memory_format=torch.channels_last):
            dy = dy.contiguous(memory_format=torch.channels_last)
        dx, dw, db = _lib.backend().bn_nhwc_bwd(x, dy, weight, training, relu, save)
        return dx, None, None, None, None, None, dw, db


class MaxPoolNhwcFn(torch.autograd.Function):
    """nn.MaxPool2d(k, stride, pad) on a dense channels_last tensor (cnsn_maxpool_nhwc_fwd / _bwd, csrc/pool_nhwc.cu).  Saved for
    backward: one byte per output element (torch keeps an int64)."""

    @staticmethod
    def forward(ctx, x, k, stride, pad):
        y, code = _lib.backend().maxpool_nhwc_fwd(x, k, stride, pad)
        ctx.pool = (code, tuple(x.shape), k, stride, pad)
        return y

    @staticmethod
    def backward(ctx, dy):
        code, shape, k, stride, pad = ctx.pool
        if not dy.is_contiguous(memory_format=torch.channels_last) or dy.is_contiguous():
            dy = dy.contiguous(memory_format=torch.channels_last)
        return _lib.backend().maxpool_nhwc_bwd(dy, code, shape, k, stride, pad), None, None, None
