"""Which BatchNorm2d class a host model builds its blocks with."""
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
