"""Pre-activation WideResNet with CNSN sites -- the caller of the hot path in BASELINE config 3.

Written from the behaviour of the reference's ``models/cifar/wideresnet_cnsn.py`` (block wiring
:66-98, site placement ``pos`` in {pre, residual, identity, post}, CrossNorm discovery and the
``_enable_cross_norm`` activation protocol :178-203, classifier head :219-227) so that
  * ``state_dict()`` keys, shapes and -- for equal seeds -- initial values are identical to the
    reference model (checkpoints interchange; tests/test_hosts.py checks it against the live reference),
  * ``forward(x, aug=False)`` consumes host RNG exactly like the reference
    (``np.random.choice(cn_num, active_num, replace=False)`` when ``aug``).
The CNSN operators come from ``ops`` (default: this package's CUDA-backed ``cnsn`` module); the benchmark's
CPU reference arm passes the eager-PyTorch restatement instead.
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from ._norm import batchnorm2d_for, bn_relu

_POSITIONS = ("residual", "identity", "pre", "post")


def _default_ops():
    from .. import cnsn
    return cnsn


class _PreActBlock(nn.Module):
    """BN-ReLU-conv3x3-BN-ReLU-conv3x3 with an additive shortcut and one CNSN site."""

    def __init__(self, cin, cout, stride, pos, beta, crop, cnsn_type, drop_rate, ops, fuse_post=False, fast_bn=True):
        super().__init__()
        BN = batchnorm2d_for(ops, fast_bn)
        self.fuse_post = bool(fuse_post) and pos == "post"
        assert cnsn_type in ("sn", "cn", "cnsn")
        assert pos in _POSITIONS
        self.pos, self.drop_rate, self.same = pos, drop_rate, cin == cout
        # attribute names and creation order follow the reference block so that parameter names and
        # the RNG draws of default initialisers line up
        self.bn1 = BN(cin)
        self.relu1 = nn.ReLU(inplace=True)
        self.conv1 = nn.Conv2d(cin, cout, 3, stride, 1, bias=False)
        self.bn2 = BN(cout)
        self.relu2 = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False)
        self.conv_shortcut = None if self.same else nn.Conv2d(cin, cout, 1, stride, 0, bias=False)
        cross = ops.CrossNorm(crop=crop, beta=beta) if "cn" in cnsn_type else None
        gate_width = cin if (pos == "pre" and not self.same) else cout
        selfn = ops.SelfNorm(gate_width) if "sn" in cnsn_type else None
        self.cnsn = ops.CNSN(crossnorm=cross, selfnorm=selfn)

    def forward(self, x):
        if self.same:
            h = self.cnsn(x) if self.pos == "pre" else x
            h = bn_relu(self.bn1, self.relu1, h)
            skip = x
        else:                                   # projection block: the pre-activation is shared
            x = bn_relu(self.bn1, self.relu1, x)
            h = self.cnsn(x) if self.pos == "pre" else x
            skip = None
        h = bn_relu(self.bn2, self.relu2, self.conv1(h))
        if self.drop_rate > 0:
            h = F.dropout(h, p=self.drop_rate, training=self.training)
        h = self.conv2(h)
        if skip is None:
            skip = self.conv_shortcut(x)
        if self.pos == "residual":
            h = self.cnsn(h)
        elif self.pos == "identity":
            skip = self.cnsn(skip)
        if self.fuse_post:                      # opt-in: the add runs inside the SelfNorm kernels
            return self.cnsn(h, skip, False)
        h = torch.add(skip, h)
        return self.cnsn(h) if self.pos == "post" else h


class _Stage(nn.Module):
    def __init__(self, count, cin, cout, stride, **kw):
        super().__init__()
        self.layer = nn.Sequential(*[
            _PreActBlock(cin if i == 0 else cout, cout, stride if i == 0 else 1, **kw) for i in range(count)])

    def forward(self, x):
        return self.layer(x)


class WideResNet(nn.Module):
    def __init__(self, depth, num_classes, widen_factor=1, drop_rate=0.0, active_num=None, pos=None, beta=None,
                 crop=None, cnsn_type=None, ops=None, verbose=False, fuse_post=False, fast_bn=True):
        super().__init__()
        ops = ops or _default_ops()
        assert (depth - 4) % 6 == 0
        per_stage = (depth - 4) // 6
        widths = [16, 16 * widen_factor, 32 * widen_factor, 64 * widen_factor]
        kw = dict(pos=pos, beta=beta, crop=crop, cnsn_type=cnsn_type, drop_rate=drop_rate, ops=ops, fuse_post=fuse_post,
                  fast_bn=fast_bn)
        self.conv1 = nn.Conv2d(3, widths[0], 3, 1, 1, bias=False)
        self.block1 = _Stage(per_stage, widths[0], widths[1], 1, **kw)
        self.block2 = _Stage(per_stage, widths[1], widths[2], 2, **kw)
        self.block3 = _Stage(per_stage, widths[2], widths[3], 2, **kw)
        self.bn1 = batchnorm2d_for(ops, fast_bn)(widths[3])
        self.relu = nn.ReLU(inplace=True)
        self.fc = nn.Linear(widths[3], num_classes)
        self.n_channels = widths[3]

        self.cn_modules = []                    # plain list, like the reference: not a registered container
        for m in self.modules():
            if isinstance(m, nn.Conv2d):        # He-normal over fan-out; Conv1d gates keep their default init
                fan_out = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2.0 / fan_out))
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()
            elif isinstance(m, nn.Linear) and m.bias is not None:
                m.bias.data.zero_()
            elif isinstance(m, ops.CrossNorm):
                self.cn_modules.append(m)
        if "cn" in cnsn_type:
            self.cn_num = len(self.cn_modules)
            self.active_num = active_num
            assert self.cn_num > 0 and self.active_num > 0
            if verbose:
                print("cn_num: %d, active_num: %d" % (self.cn_num, self.active_num))

    def _enable_cross_norm(self):
        chosen = np.random.choice(self.cn_num, self.active_num, replace=False).tolist()
        assert len(set(chosen)) == self.active_num
        for i in chosen:
            self.cn_modules[i].active = True

    def forward(self, x, aug=False):
        if aug:
            self._enable_cross_norm()
        h = self.block3(self.block2(self.block1(self.conv1(x))))
        h = F.avg_pool2d(bn_relu(self.bn1, self.relu, h), 8)
        return self.fc(h.view(h.size(0), -1))
