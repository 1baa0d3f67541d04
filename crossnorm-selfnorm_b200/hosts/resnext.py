"""CIFAR ResNeXt (bottleneck type C) with CNSN sites -- the third caller SURVEY.md 8b lists.

Written from the behaviour of the reference's ``models/cifar/resnext_cnsn.py`` (block wiring and the four site
positions :84-113, stage construction :181-211, CrossNorm discovery and activation :162-179, :213-218, stem and head
:220-233) so that ``state_dict()`` keys, shapes and -- for equal seeds -- initial values are those of the reference
model and ``forward(x, aug=False)`` consumes host RNG like the reference (tests/test_hosts.py checks both against the
live reference file).  Quirks kept on purpose: the 'post' site sits AFTER the block's ReLU (:108-111), and at
``pos='identity'`` a projection block overwrites the site's output with ``downsample(x)`` (:103-106).
"""
import math

import numpy as np
import torch.nn as nn
import torch.nn.functional as F

from ._norm import batchnorm2d_for, bn_relu


def _relu_(t):
    return F.relu(t, inplace=True)

_POSITIONS = ("residual", "identity", "pre", "post")
_EXPANSION = 4


def _default_ops():
    from .. import cnsn
    return cnsn


class _Bottleneck(nn.Module):
    def __init__(self, cin, planes, cardinality, base_width, pos, beta, crop, cnsn_type, ops, stride=1, downsample=None):
        super().__init__()
        BN = batchnorm2d_for(ops)
        width = int(math.floor(planes * (base_width / 64.0))) * cardinality
        cout = planes * _EXPANSION
        self.conv_reduce = nn.Conv2d(cin, width, 1, 1, 0, bias=False)
        self.bn_reduce = BN(width)
        self.conv_conv = nn.Conv2d(width, width, 3, stride, 1, groups=cardinality, bias=False)
        self.bn = BN(width)
        self.conv_expand = nn.Conv2d(width, cout, 1, 1, 0, bias=False)
        self.bn_expand = BN(cout)
        self.downsample = downsample
        assert cnsn_type in ("sn", "cn", "cnsn") and pos in _POSITIONS
        cross = ops.CrossNorm(crop=crop, beta=beta) if "cn" in cnsn_type else None
        selfn = ops.SelfNorm(cin if pos in ("pre", "identity") else cout) if "sn" in cnsn_type else None
        self.cnsn = ops.CNSN(crossnorm=cross, selfnorm=selfn)
        self.pos = pos

    def forward(self, x):
        skip = x
        if self.pos == "pre":
            x = self.cnsn(x)
        h = bn_relu(self.bn_reduce, _relu_, self.conv_reduce(x))
        h = bn_relu(self.bn, _relu_, self.conv_conv(h))
        h = self.bn_expand(self.conv_expand(h))
        if self.pos == "residual":
            h = self.cnsn(h)
        if self.pos == "identity":
            skip = self.cnsn(skip)
        if self.downsample is not None:
            skip = self.downsample(x)
        out = F.relu(skip + h, inplace=True)
        return self.cnsn(out) if self.pos == "post" else out


class CifarResNeXt(nn.Module):
    def __init__(self, depth, cardinality, base_width, num_classes, active_num=None, pos=None, beta=None, crop=None,
                 cnsn_type=None, ops=None):
        super().__init__()
        ops = ops or _default_ops()
        assert (depth - 2) % 9 == 0, "depth should be one of 29, 38, 47, 56, 101"
        per_stage = (depth - 2) // 9
        self.cardinality, self.base_width, self.num_classes = cardinality, base_width, num_classes
        self.conv_1_3x3 = nn.Conv2d(3, 64, 3, 1, 1, bias=False)
        self.bn_1 = batchnorm2d_for(ops)(64)
        kw = dict(pos=pos, beta=beta, crop=crop, cnsn_type=cnsn_type, ops=ops)
        width = 64
        stages = []
        for planes, stride in ((64, 1), (128, 2), (256, 2)):
            shortcut = None
            if stride != 1 or width != planes * _EXPANSION:
                shortcut = nn.Sequential(nn.Conv2d(width, planes * _EXPANSION, 1, stride, bias=False),
                                         batchnorm2d_for(ops)(planes * _EXPANSION))
            blocks = [_Bottleneck(width, planes, cardinality, base_width, stride=stride, downsample=shortcut, **kw)]
            width = planes * _EXPANSION
            blocks += [_Bottleneck(width, planes, cardinality, base_width, **kw) for _ in range(1, per_stage)]
            stages.append(nn.Sequential(*blocks))
        self.stage_1, self.stage_2, self.stage_3 = stages
        self.avgpool = nn.AvgPool2d(8)
        self.classifier = nn.Linear(width, num_classes)

        self.cn_modules = []                              # plain list, like the reference
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                m.weight.data.normal_(0, math.sqrt(2.0 / (m.kernel_size[0] * m.kernel_size[1] * m.out_channels)))
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()
            elif isinstance(m, nn.Linear):
                nn.init.kaiming_normal_(m.weight)
                m.bias.data.zero_()
            elif isinstance(m, ops.CrossNorm):
                self.cn_modules.append(m)
        if "cn" in cnsn_type:
            self.cn_num, self.active_num = len(self.cn_modules), active_num
            assert self.cn_num > 0 and self.active_num > 0

    def _enable_cross_norm(self):
        chosen = np.random.choice(self.cn_num, self.active_num, replace=False).tolist()
        assert len(set(chosen)) == self.active_num
        for i in chosen:
            self.cn_modules[i].active = True

    def forward(self, x, aug=False):
        if aug:
            self._enable_cross_norm()
        h = bn_relu(self.bn_1, _relu_, self.conv_1_3x3(x))
        h = self.avgpool(self.stage_3(self.stage_2(self.stage_1(h))))
        return self.classifier(h.view(h.size(0), -1))


def resnext29(num_classes=10, cardinality=4, base_width=32, active_num=2, pos="post", beta=1, crop="both",
              cnsn_type="cnsn", **kw):
    """ResNeXt-29 (4x32d) + CNSN, the model of cifar10-scripts/resnext/run-cnsn.sh."""
    return CifarResNeXt(29, cardinality, base_width, num_classes, active_num=active_num, pos=pos, beta=beta, crop=crop,
                        cnsn_type=cnsn_type, **kw)
