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
