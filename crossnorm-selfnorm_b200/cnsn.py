"""Drop-in replacement for the reference's ``models/cnsn.py`` (amazon-science/crossnorm-selfnorm).

Same public names, constructor signatures, ``state_dict`` keys, ``.active`` protocol and host-side
RNG consumption as the reference, so the reference's unmodified WideResNet / ResNeXt / ResNet-50
files work with it (``sys.modules['models.cnsn'] = cnsn_b200.cnsn`` or copy it over the file).
Everything that touches a feature map runs in hand-written sm_100a CUDA kernels behind the C ABI in
``include/cnsn_b200.h``; there is no PyTorch / CPU fallback.

Reference lines (relative to the reference checkout) are cited per symbol.
"""
import functools

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .functional import CnsnSiteFn, CrossNormFn, InstanceAffine, InstanceStats, SelfNormBlockFn, SelfNormFn

__all__ = ["calc_ins_mean_std", "instance_norm_mix", "cn_rand_bbox", "cn_op_2ins_space_chan",
           "CrossNorm", "SelfNorm", "CNSN"]

_CROPS = ("neither", "style", "content", "both")


def _bn_momentum(bn):
    """BatchNorm's update factor for this call (``momentum=None`` means torch's cumulative average)."""
    if bn.momentum is not None:
        return float(bn.momentum)
    if not bn.training or bn.num_batches_tracked is None:
        return 0.0
    return 1.0 / (int(bn.num_batches_tracked) + 1)


def _full(x):
    return (0, x.size(2), 0, x.size(3))


def calc_ins_mean_std(x, eps=1e-5):
    """Per-instance mean and sqrt(unbiased var + eps), each (N,C,1,1).  models/cnsn.py:8-17."""
    assert x.dim() == 4                                   # :12
    N, C = x.shape[:2]
    mean, std = InstanceStats.apply(x, _full(x), float(eps))
    return mean.to(x.dtype).view(N, C, 1, 1), std.to(x.dtype).view(N, C, 1, 1)


def instance_norm_mix(content_feat, style_feat):
    """Replace content statistics with style statistics.  models/cnsn.py:20-29 (eps 1e-5)."""
    assert content_feat.size()[:2] == style_feat.size()[:2]                    # :22
    s_mean, s_std = InstanceStats.apply(style_feat, _full(style_feat), 1e-5)
    c_mean, c_std = InstanceStats.apply(content_feat, _full(content_feat), 1e-5)
    scale = s_std / c_std                                  # O(N*C) glue; the planes go through the kernels
    return InstanceAffine.apply(content_feat, scale, s_mean - c_mean * scale)


def cn_rand_bbox(size, beta, bbx_thres):
    """Sample a crop box; same draws, same order, same return convention as models/cnsn.py:32-55:
    (bbx1, bby1, bbx2, bby2) with bbx* indexing dim 2 and bby* dim 3.  (The reference's ``np.int`` is spelled ``int``
    here -- NumPy >= 1.24 removed the alias -- and its ``np.clip`` on scalars is plain min / max: same integers.)"""
    d2, d3 = int(size[2]), int(size[3])
    while True:
        ratio = np.random.beta(beta, beta)
        side = np.sqrt(ratio)
        ext2, ext3 = int(d2 * side), int(d3 * side)
        c2 = int(np.random.randint(d2))
        c3 = int(np.random.randint(d3))
        bbx1, bbx2 = min(max(c2 - ext2 // 2, 0), d2), min(max(c2 + ext2 // 2, 0), d2)
        bby1, bby2 = min(max(c3 - ext3 // 2, 0), d3), min(max(c3 + ext3 // 2, 0), d3)
        if float(bbx2 - bbx1) * (bby2 - bby1) / (d2 * d3) > bbx_thres:
            return bbx1, bby1, bbx2, bby2


def _draw_windows(x, crop, beta, bbx_thres):
    """The numpy draws of one cn_op_2ins_space_chan call in the reference's order (style box :64-66, then content box
    :74-77), as (content, style) windows (h0, h1, w0, w1)."""
    assert crop in _CROPS                                   # :61
    assert x.dim() == 4
    H, W = x.size(2), x.size(3)
    swin = cwin = (0, H, 0, W)
    if crop in ('style', 'both'):
        a1, b1, a2, b2 = cn_rand_bbox(x.size(), beta=beta, bbx_thres=bbx_thres)
        swin = (a1, a2, b1, b2)
    if crop in ('content', 'both'):
        a1, b1, a2, b2 = cn_rand_bbox(x.size(), beta=beta, bbx_thres=bbx_thres)
        cwin = (a1, a2, b1, b2)
    return cwin, swin


class _PinnedRing:
    """A few persistent pinned int32 staging blocks per device for permutation uploads: no pinned allocation per
    call, and a block is reused only after the copy that read it has completed (CUDA event)."""
    SLOTS, WORDS = 16, 4096

    def __init__(self):
        self.buf = torch.empty((self.SLOTS, self.WORDS), dtype=torch.int32, pin_memory=True)
        self.events = [None] * self.SLOTS
        self.next = 0

    def upload(self, idx, device):
        n = idx.numel()
        if n > self.WORDS:
            host = torch.empty(n, dtype=torch.int32, pin_memory=True)
            host.copy_(idx)
            return host.to(device, non_blocking=True)
        i = self.next
        self.next = (i + 1) % self.SLOTS
        if self.events[i] is not None:
            self.events[i].synchronize()              # almost always already complete (16 uploads ago)
        host = self.buf[i, :n]
        host.copy_(idx)
        out = host.to(device, non_blocking=True)
        ev = self.events[i] or torch.cuda.Event()
        ev.record(torch.cuda.current_stream(device))
        self.events[i] = ev
        return out


_rings = {}


def _to_device_i32(idx, device):
    """Upload a host permutation without the reference's blocking pageable copy (:62)."""
    if device.type != "cuda":
        return idx.to(torch.int32)
    ring = _rings.get(device)
    if ring is None:
        ring = _rings[device] = _PinnedRing()
    return ring.upload(idx, device)


def cn_op_2ins_space_chan(x, crop='neither', beta=1, bbx_thres=0.1, lam=None, chan=False):
    """2-instance CrossNorm with optional crops, channel permutation and lam blend.

    models/cnsn.py:58-91.  The host draws exactly what the reference draws, in its order
    (SURVEY.md A.3): randperm(N) on the CPU generator, style box, randperm(C), content box; the
    device work (two statistics sets, permuted restyle, copy-through outside the content box) is one
    fused CUDA forward and one fused backward.
    """
    ext = _lib.fast_binding() if (x.is_cuda and not chan) else None
    if ext is not None:
        # C++ node: torch.randperm(N) (:62) is drawn inside it on the same CPU generator; torch and numpy streams are
        # independent, so drawing the boxes first changes nothing
        cwin, swin = _draw_windows(x, crop, beta, bbx_thres)
        return ext.crossnorm(x, cwin, swin, 0.0 if lam is None else float(lam), 1e-5)
    perm_d, cperm_d, cwin, swin = _draw_plan(x, crop, beta, bbx_thres, chan)
    return CrossNormFn.apply(x, perm_d, cperm_d, cwin, swin, 0.0 if lam is None else float(lam), 1e-5)


def _draw_plan(x, crop, beta, bbx_thres, chan):
    """The host-side draws of one cn_op_2ins_space_chan call, in the reference's order (models/cnsn.py:61-77,
    SURVEY.md A.3); the permutations are uploaded as int32.  Returns (perm, chan_perm | None, content, style)
    with windows as (h0, h1, w0, w1)."""
    assert crop in _CROPS                                   # :61
    assert x.dim() == 4
    N, C, H, W = x.shape
    perm = torch.randperm(N)                                # :62
    swin = cwin = (0, H, 0, W)
    if crop in ('style', 'both'):                           # :64-66
        a1, b1, a2, b2 = cn_rand_bbox(x.size(), beta=beta, bbx_thres=bbx_thres)
        swin = (int(a1), int(a2), int(b1), int(b2))
    chan_perm = torch.randperm(C) if chan else None         # :70-72
    if crop in ('content', 'both'):                         # :74-77
        a1, b1, a2, b2 = cn_rand_bbox(x.size(), beta=beta, bbx_thres=bbx_thres)
        cwin = (int(a1), int(a2), int(b1), int(b2))
    perm_d = _to_device_i32(perm, x.device)
    cperm_d = _to_device_i32(chan_perm, x.device) if chan else None
    return perm_d, cperm_d, cwin, swin


class CrossNorm(nn.Module):
    """CrossNorm module.  models/cnsn.py:94-110: runs only when training and ``active``; ``active``
    is a one-shot flag the host model sets and every forward resets.  No parameters or buffers."""

    def __init__(self, crop=None, beta=None):
        super().__init__()
        self.active = False
        self.cn_op = functools.partial(cn_op_2ins_space_chan, crop=crop, beta=beta)

    def forward(self, x):
        if self.training and self.active:
            x = self.cn_op(x)
        self.active = False
        return x


class SelfNorm(nn.Module):
    """SelfNorm module.  models/cnsn.py:113-150.

    The parameters live in sub-modules named and constructed exactly as in the reference
    (``g_fc``: depthwise Conv1d k=2 without bias, ``g_bn``: BatchNorm1d; ``f_*`` twins when
    ``is_two``) so initialisation consumes the same RNG draws and ``state_dict()`` keys / shapes
    match; those sub-modules are never called -- one fused kernel sequence reads their tensors.
    """

    def __init__(self, chan_num, is_two=False):
        super().__init__()
        self.g_fc = nn.Conv1d(chan_num, chan_num, kernel_size=2, bias=False, groups=chan_num)
        self.g_bn = nn.BatchNorm1d(chan_num)
        if is_two is True:
            self.f_fc = nn.Conv1d(chan_num, chan_num, kernel_size=2, bias=False, groups=chan_num)
            self.f_bn = nn.BatchNorm1d(chan_num)
        else:
            self.f_fc = None

    @staticmethod
    def _bufs(bn):
        return (bn.running_mean, bn.running_var, bn.num_batches_tracked)

    def forward(self, x, residual=None, relu=False):
        """``forward(x)`` is the reference call.  ``forward(x, residual, relu)`` is the opt-in block fusion
        relu?(SelfNorm(x + residual)) for pos='post' sites (one gate only)."""
        assert x.dim() == 4
        bn = self.g_bn
        if self.f_fc is None and x.is_cuda and self.g_fc.weight.dtype is torch.float32 and bn.weight is not None:
            ext = _lib.fast_binding()
            if ext is not None:                       # C++ autograd node: one call per direction
                return ext.selfnorm(x, residual, bool(relu), self.g_fc.weight, bn.weight, bn.bias, bn.running_mean,
                                    bn.running_var, bn.num_batches_tracked, bn.training, _bn_momentum(bn), float(bn.eps), 1e-12)
        nhwc = (self.f_fc is None and x.is_cuda and not x.is_contiguous()
                and x.is_contiguous(memory_format=torch.channels_last))    # channels_last: the block entry has NHWC kernels
        if residual is not None or relu or nhwc:
            assert self.f_fc is None, "the fused block supports the single-gate SelfNorm"
            assert residual is None or residual.shape == x.shape
            return SelfNormBlockFn.apply(x, residual, bool(relu), bn.training, _bn_momentum(bn), float(bn.eps), 1e-12,
                                         self._bufs(bn), self.g_fc.weight, bn.weight, bn.bias)
        args = [x, bn.training, _bn_momentum(bn), float(bn.eps), 1e-12,          # eps :133
                self._bufs(bn), self._bufs(self.f_bn) if self.f_fc is not None else None,
                self.g_fc.weight, bn.weight, bn.bias]
        if self.f_fc is not None:
            args += [self.f_fc.weight, self.f_bn.weight, self.f_bn.bias]
        return SelfNormFn.apply(*args)


class CNSN(nn.Module):
    """CrossNorm then SelfNorm; either may be None.  models/cnsn.py:152-164."""

    def __init__(self, crossnorm, selfnorm):
        super().__init__()
        self.crossnorm = crossnorm
        self.selfnorm = selfnorm

    def forward(self, x, residual=None, relu=False):
        """``forward(x)`` is the reference call (models/cnsn.py:159-164).  ``forward(x, residual, relu)`` computes
        relu?(CNSN(x + residual)) -- the pos='post' block tail -- fusing add and ReLU into the SelfNorm kernels
        whenever this step's CrossNorm does not fire at the site.  When both operators fire, the pair runs as ONE
        fused kernel per direction (``cnsn_site_fwd/_bwd``) wherever ``_site_fusable`` says so, else
        one after the other; host RNG consumption, ``.active`` handling and results are the same either way."""
        fire = bool(self.crossnorm) and self.crossnorm.active
        if residual is None and not relu:
            if fire:
                if self._site_fusable(x):
                    return self._site(x, False)
                x = self.crossnorm(x)
            if self.selfnorm:
                x = self.selfnorm(x)
            return x
        if fire or not self.selfnorm:
            if residual is not None:
                x = torch.add(residual, x)
            if fire:
                if self._site_fusable(x):
                    return self._site(x, relu)
                x = self.crossnorm(x)
            if self.selfnorm:
                return self.selfnorm(x, None, relu)
            return torch.relu(x) if relu else x
        return self.selfnorm(x, residual, relu)

    fuse_site = True    # class-wide switch (A/B measurements, tests): False -> always the two-operator sequence

    def _site_fusable(self, x):
        """Both operators fire at this call and one fused kernel per direction can serve it: the CrossNorm is the
        stock partial of cn_op_2ins_space_chan without channel permutation, the SelfNorm is single-gate and in
        training mode, and the shape fits the resident kernels (cnsn_site_supported)."""
        cn, sn = self.crossnorm, self.selfnorm
        if not (CNSN.fuse_site and sn and cn.training and cn.active and x.dim() == 4):
            return False
        op = getattr(cn, "cn_op", None)
        if not (isinstance(sn, SelfNorm) and sn.f_fc is None and sn.g_bn.training and isinstance(op, functools.partial)
                and op.func is cn_op_2ins_space_chan and not op.args and not op.keywords.get("chan", False)):
            return False
        return _lib.backend().site_supported(x)

    def _site(self, x, relu):
        """``relu?(selfnorm(crossnorm(x)))`` in one kernel per direction; the host draws are those of
        ``crossnorm(x)`` (same RNG consumption), and the one-shot ``active`` flag is consumed (models/cnsn.py:108)."""
        cn, sn = self.crossnorm, self.selfnorm
        kw = cn.cn_op.keywords
        bn = sn.g_bn
        ext = _lib.fast_binding() if sn.g_fc.weight.dtype is torch.float32 else None
        if ext is not None:
            cwin, swin = _draw_windows(x, kw.get("crop", "neither"), kw.get("beta", 1), kw.get("bbx_thres", 0.1))
            lam = kw.get("lam")
            cn.active = False
            return ext.site(x, cwin, swin, 0.0 if lam is None else float(lam), 1e-5, bool(relu), sn.g_fc.weight, bn.weight,
                            bn.bias, bn.running_mean, bn.running_var, bn.num_batches_tracked, _bn_momentum(bn), float(bn.eps), 1e-12)
        perm_d, _, cwin, swin = _draw_plan(x, kw.get("crop", "neither"), kw.get("beta", 1), kw.get("bbx_thres", 0.1), False)
        lam = kw.get("lam")
        cn.active = False
        bn = sn.g_bn
        return CnsnSiteFn.apply(x, perm_d, cwin, swin, 0.0 if lam is None else float(lam), 1e-5, bool(relu),
                                _bn_momentum(bn), float(bn.eps), 1e-12, SelfNorm._bufs(bn),
                                sn.g_fc.weight, bn.weight, bn.bias)
