"""Instance-Batch Normalization with the reference's module surface.

``IBN`` mirrors the class of ``models/imagenet/resnet_ibn_cnsn.py:24-44``: sub-modules ``IN``
(``nn.InstanceNorm2d(half, affine=True)``) and ``BN`` (``nn.BatchNorm2d(planes - half)``) hold the parameters and
buffers, so ``state_dict()`` keys match (``...bn1.IN.weight``, ``...bn1.BN.running_mean`` ...) and the IBN-Net /
reference checkpoints load; the sub-modules are never called -- one CUDA kernel per direction handles both halves in
place, without the reference's split / contiguous / cat copies.
"""
import torch
import torch.nn as nn

from . import _lib
from .functional import BnNhwcFn, IbnFn


def bn_momentum(bn):
    """The exponential-average factor torch's batch norm uses for this call: ``momentum``, or -- ``momentum=None`` --
    the cumulative average 1 / (num_batches_tracked + 1) (a host read of the counter: rare configuration)."""
    if bn.momentum is not None:
        return float(bn.momentum)
    if not bn.training or bn.num_batches_tracked is None:
        return 0.0
    return 1.0 / (int(bn.num_batches_tracked) + 1)

__all__ = ["IBN", "InstanceNorm2d", "BatchNorm2d"]


class IBN(nn.Module):
    def __init__(self, planes, ratio=0.5):
        super().__init__()
        self.half = int(planes * ratio)
        self.IN = nn.InstanceNorm2d(self.half, affine=True)
        self.BN = nn.BatchNorm2d(planes - self.half)

    def forward(self, x):
        assert x.dim() == 4
        bn = self.BN
        momentum = bn_momentum(bn)
        ext = _lib.fast_binding() if (x.is_cuda and bn.weight.dtype is torch.float32) else None
        if ext is not None:                           # C++ autograd node
            return ext.ibn(x, self.half, bn.training, False, momentum, float(self.IN.eps), float(bn.eps), bn.running_mean,
                           bn.running_var, bn.num_batches_tracked, self.IN.weight, self.IN.bias, bn.weight, bn.bias)
        return IbnFn.apply(x, self.half, bn.training, momentum, float(self.IN.eps), float(bn.eps),
                           (bn.running_mean, bn.running_var, bn.num_batches_tracked),
                           self.IN.weight, self.IN.bias, bn.weight, bn.bias)


class InstanceNorm2d(nn.InstanceNorm2d):
    """``nn.InstanceNorm2d(C, affine=True)`` as the reference uses it for IBN-b -- after the residual add of the last
    block of a stage and in the stem (``models/imagenet/resnet_ibn_cnsn.py:62,122-123,143-144``) -- through the same
    kernels as ``IBN`` with every channel on the instance-norm side (``half = C``): one launch per direction.  Same
    class hierarchy and ``state_dict`` keys (``weight``, ``bias``) as the torch module it replaces."""

    def __init__(self, num_features, eps=1e-5, affine=True):
        super().__init__(num_features, eps=eps, affine=True, track_running_stats=False)
        assert affine, "the reference constructs InstanceNorm2d(C, affine=True)"

    def forward(self, x):
        assert x.dim() == 4 and x.size(1) == self.num_features
        ext = _lib.fast_binding() if (x.is_cuda and self.weight.dtype is torch.float32) else None
        if ext is not None:
            return ext.ibn(x, self.num_features, False, False, 0.1, float(self.eps), 1e-5, None, None, None, self.weight,
                           self.bias, None, None)
        return IbnFn.apply(x, self.num_features, False, 0.1, float(self.eps), 1e-5, (None, None, None),
                           self.weight, self.bias, None, None)


class BatchNorm2d(nn.BatchNorm2d):
    """``nn.BatchNorm2d`` of the host blocks (models/cifar/wideresnet_cnsn.py:51-57, models/imagenet/resnet_cnsn.py:60-75)
    through the IBN kernels with every channel on the batch-norm side (``half = 0``): one shared-memory-resident launch
    per direction -- x crosses HBM once forward, x and dy once backward -- instead of cuDNN's one-CTA-per-channel kernels,
    which take 55 % of a WideResNet-40-2 step on B200 (32-128 channels on 148 SMs; profiles/README.md).  Same class
    hierarchy, parameters, buffers and ``state_dict`` keys as the torch module; same arithmetic (batch statistics when
    training, biased variance to normalise, unbiased into ``running_var``, ``momentum=None`` = cumulative average).
    Anything the resident kernels do not take (CPU tensors, planes that are not 16-byte multiples or do not fit shared
    memory, ``affine=False`` / no running statistics) goes to the torch implementation."""

    def forward(self, x, relu=False):
        """``forward(x)`` is ``nn.BatchNorm2d.forward``; ``forward(x, relu=True)`` is ``relu(bn(x))`` -- the pair the host
        blocks apply -- in the same kernels when the resident path takes the shape, else torch's batch norm + relu."""
        # channels_last activations stay with torch: cuDNN's NHWC batch norm keeps the layout (the library's kernels are NCHW)
        if (x.is_cuda and x.dim() == 4 and x.is_contiguous() and self.affine and self.track_running_stats
                and self.weight.dtype is torch.float32 and x.dtype in (torch.float32, torch.bfloat16, torch.float16)):
            training = self.training
            be = _lib.backend()
            if getattr(be, "name", "") == "cuda" and be.ibn_resident(x, 0, training):
                momentum = bn_momentum(self)
                ext = _lib.fast_binding()
                if ext is not None:
                    return ext.ibn(x, 0, training, bool(relu), momentum, 1e-5, float(self.eps), self.running_mean,
                                   self.running_var, self.num_batches_tracked, None, None, self.weight, self.bias)
                return IbnFn.apply(x, 0, training, momentum, 1e-5, float(self.eps),
                                   (self.running_mean, self.running_var, self.num_batches_tracked),
                                   None, None, self.weight, self.bias, bool(relu))
        if (x.is_cuda and x.dim() == 4 and not x.is_contiguous() and self.affine and self.track_running_stats
                and self.weight.dtype is torch.float32 and x.dtype in (torch.float32, torch.bfloat16, torch.float16)):
            be = _lib.backend()                       # channels_last: the NHWC kernels (csrc/bn_nhwc.cu), same fused ReLU
            if getattr(be, "name", "") == "cuda" and be.bn_nhwc_ok(x):
                momentum = bn_momentum(self)
                ext = _lib.fast_binding()
                if ext is not None:
                    return ext.bn_nhwc(x, self.training, bool(relu), momentum, float(self.eps), self.running_mean, self.running_var,
                                       self.num_batches_tracked, self.weight, self.bias)
                return BnNhwcFn.apply(x, self.training, bool(relu), momentum, float(self.eps),
                                      (self.running_mean, self.running_var, self.num_batches_tracked), self.weight, self.bias)
        y = super().forward(x)
        return torch.relu_(y) if relu else y

