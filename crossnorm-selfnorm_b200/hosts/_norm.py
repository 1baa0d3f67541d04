"""Which BatchNorm2d class a host model builds its blocks with."""
import torch
import torch.nn as nn


def batchnorm2d_for(ops, fast_bn=True):
    """``nn.BatchNorm2d`` -- or, when the CNSN operators are this package's CUDA-backed modules (``ops`` is
    ``cnsn_b200.cnsn``) and ``fast_bn``, its drop-in subclass ``cnsn_b200.ibn.BatchNorm2d`` (same parameters, buffers,
    state-dict keys and arithmetic; one resident kernel per direction where the shape allows, torch otherwise)."""
    if fast_bn and hasattr(ops, "_lib") and getattr(ops, "__name__", "").endswith("cnsn"):
        from ..ibn import BatchNorm2d
        return BatchNorm2d
    return nn.BatchNorm2d


def bn_relu(bn, relu, x):
    """``relu(bn(x))`` of a host block: ONE fused call when ``bn`` is the package's BatchNorm2d drop-in (the ReLU runs
    inside the batch-norm kernels, forward and backward), the two modules one after the other otherwise."""
    from ..ibn import BatchNorm2d
    if isinstance(bn, BatchNorm2d):
        return bn(x, True)
    return relu(bn(x))


def bn_site_relu(bn, cnsn, c, skip):
    """``relu(cnsn(bn(c) + skip))`` -- the tail of a pos='post' ResNet bottleneck (models/imagenet/resnet_cnsn.py:113-122).
    ONE fused operator (csrc/selfnorm_nhwc.cu, cnsn_bn_selfnorm_tail_*: bn's output is never written, its backward
    reduction rides on the SelfNorm backward) when everything lines up -- this package's BatchNorm2d and CNSN with a
    single-gate SelfNorm whose CrossNorm does not fire at this call, a dense channels_last CUDA tensor of a supported shape,
    fp32 parameters, the C++ binding --, else exactly the two calls it replaces.  Same results bit for bit either way."""
    from .. import _lib
    from ..cnsn import CNSN, SelfNorm, _bn_momentum
    from ..ibn import BatchNorm2d, bn_momentum
    sn = getattr(cnsn, "selfnorm", None)
    cn = getattr(cnsn, "crossnorm", None)
    if (FUSE_TAIL and isinstance(bn, BatchNorm2d) and isinstance(cnsn, CNSN) and isinstance(sn, SelfNorm) and sn.f_fc is None
            and not (cn is not None and cn.active) and c.is_cuda and c.dim() == 4 and not c.is_contiguous()
            and bn.affine and bn.track_running_stats and bn.weight.dtype is torch.float32
            and sn.g_fc.weight.dtype is torch.float32 and sn.g_bn.weight is not None and skip.shape == c.shape
            and skip.dtype == c.dtype and c.dtype in (torch.float32, torch.bfloat16, torch.float16)):
        ext = _lib.fast_binding()
        if ext is not None and ext.bn_sn_tail_supported(c):
            g = sn.g_bn
            return ext.bn_sn_tail(c, skip, True, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.num_batches_tracked,
                                  bn.training, bn_momentum(bn), float(bn.eps), sn.g_fc.weight, g.weight, g.bias, g.running_mean,
                                  g.running_var, g.num_batches_tracked, g.training, _bn_momentum(g), float(g.eps), 1e-12)
    return cnsn(bn(c), skip, True)


FUSE_TAIL = True        # module-wide switch (A/B measurements, tests): False -> always the two calls


class MaxPool2d(nn.MaxPool2d):
    """``nn.MaxPool2d`` whose dense channels_last CUDA inputs run this package's NHWC kernels (csrc/pool_nhwc.cu: one byte of
    saved state per output element, gather backward without atomics; same results as torch including the tie rule);
    everything else -- NCHW, CPU, dilation, ceil_mode, return_indices -- is ``nn.MaxPool2d.forward``."""

    def forward(self, x):
        k, s, p, d = self.kernel_size, self.stride, self.padding, self.dilation
        if (x.is_cuda and x.dim() == 4 and all(isinstance(v, int) for v in (k, s, p, d)) and d == 1 and not self.ceil_mode
                and not self.return_indices):
            from .. import _lib
            be = _lib.backend()
            if getattr(be, "name", "") == "cuda" and be.maxpool_nhwc_ok(x, k, s, p):
                ext = _lib.fast_binding()
                if ext is not None:
                    return ext.maxpool_nhwc(x, k, s, p)
                from ..functional import MaxPoolNhwcFn
                return MaxPoolNhwcFn.apply(x, k, s, p)
        return super().forward(x)


def maxpool2d_for(ops, fast=True):
    """``nn.MaxPool2d`` or, with this package's operators, the drop-in above."""
    if fast and hasattr(ops, "_lib") and getattr(ops, "__name__", "").endswith("cnsn"):
        return MaxPool2d
    return nn.MaxPool2d
