"""ResNet-v1.5 bottleneck network with CNSN sites -- the caller of the hot path in BASELINE configs 4 and 5.

Written from the behaviour of the reference's ``models/imagenet/resnet_cnsn.py`` (bottleneck wiring and the four
site positions :86-124, stage construction :205-236, CrossNorm discovery and activation :170-184, :238-243, stem and
head :247-268) so that
  * ``state_dict()`` keys, shapes and -- for equal seeds -- initial values are those of the reference model
    (``layer2.0.downsample.1.weight``, ``layer1.0.cnsn.selfnorm.g_fc.weight`` ...; the published ResNet-50+SN
    checkpoints load),
  * ``forward(x, aug=False)`` consumes host RNG like the reference.
``fuse_post=True`` (opt-in) routes the tail of a ``pos='post'`` block -- ``out += identity``,
``cnsn(out)``, ``relu`` (:117-122) -- through ``CNSN.forward(out, identity, relu=True)``: one fused kernel pair
instead of add / SelfNorm / ReLU.  Results are identical (tests/test_hosts.py, tests/test_gpu_parity.py).
"""
import numpy as np
import torch
import torch.nn as nn

from ._norm import batchnorm2d_for, bn_relu, bn_site_relu, maxpool2d_for

_POSITIONS = ("residual", "pre", "post", "identity")
_EXPANSION = 4


def _default_ops():
    from .. import cnsn
    return cnsn


class Bottleneck(nn.Module):
    """1x1 reduce, 3x3 (carries the stride), 1x1 expand, additive shortcut, ReLU; one optional CNSN site."""

    def __init__(self, cin, planes, stride, shortcut, pos, beta, crop, cnsn_type, ops, fuse_post, ibn=None, ibn_host=False,
                 fast_bn=True):
        super().__init__()
        BN = batchnorm2d_for(ops, fast_bn)
        assert ibn in (None, "a", "b")
        cout = planes * _EXPANSION
        self.ibn_variant = bool(ibn_host)                # wiring of resnet_ibn_cnsn.py (every block of that host)
        self.conv1 = nn.Conv2d(cin, planes, 1, bias=False)
        if ibn == "a":                                   # models/imagenet/resnet_ibn_cnsn.py:53-56
            from ..ibn import IBN
            self.bn1 = IBN(planes)
        else:
            self.bn1 = BN(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride, 1, bias=False)
        self.bn2 = BN(planes)
        self.conv3 = nn.Conv2d(planes, cout, 1, bias=False)
        self.bn3 = BN(cout)
        if ibn == "b":                                   # IBN-b: instance norm after the residual add (:62, :122-123)
            from ..ibn import InstanceNorm2d
            self.IN = InstanceNorm2d(cout, affine=True)
        else:
            self.IN = None
        self.relu = nn.ReLU(inplace=True)
        self.downsample = shortcut
        self.stride = stride
        self.pos = pos
        self.fuse_post = bool(fuse_post) and pos == "post"
        self.cnsn = None
        if self.IN is not None and pos == "post":         # the instance norm takes the 'post' site's place (:67-68)
            self.fuse_post = False
        elif cnsn_type is not None:                       # None: CrossNorm lives in image space only
            assert cnsn_type in ("sn", "cn", "cnsn")
            assert pos in _POSITIONS
            cross = ops.CrossNorm(crop=crop, beta=beta) if "cn" in cnsn_type else None
            selfn = ops.SelfNorm(cin if pos == "pre" else cout) if "sn" in cnsn_type else None
            self.cnsn = ops.CNSN(crossnorm=cross, selfnorm=selfn)

    def forward(self, x):
        h = self.cnsn(x) if self.pos == "pre" else x
        if self.ibn_variant and self.pos == "pre" and self.downsample is not None:
            x = h                                        # the IBN host feeds the projection from the site's output (:98-99,:112-113)
        h = bn_relu(self.bn1, self.relu, self.conv1(h))
        h = bn_relu(self.bn2, self.relu, self.conv2(h))
        skip = x if self.downsample is None else self.downsample(x)
        if self.fuse_post:                                # relu(cnsn(bn3(conv3(h)) + skip)): one fused operator where possible
            return bn_site_relu(self.bn3, self.cnsn, self.conv3(h), skip)
        h = self.bn3(self.conv3(h))
        if self.pos == "residual":
            h = self.cnsn(h)
        elif self.pos == "identity":
            skip = self.cnsn(skip)
        h = h + skip
        if self.IN is not None:
            h = self.IN(h)
        elif self.pos == "post":
            h = self.cnsn(h)
        return self.relu(h)


class ResNet(nn.Module):
    def __init__(self, layers, num_classes=1000, active_num=1, pos=None, beta=None, crop=None, cnsn_type=None,
                 ops=None, fuse_post=False, zero_init_residual=False, ibn_cfg=(None, None, None, None), fast_bn=True):
        super().__init__()
        ops = ops or _default_ops()
        BN = batchnorm2d_for(ops, fast_bn)
        self.conv1 = nn.Conv2d(3, 64, 7, 2, 3, bias=False)
        if ibn_cfg[0] == "b":                             # IBN-b stem (resnet_ibn_cnsn.py:143-144)
            from ..ibn import InstanceNorm2d
            self.bn1 = InstanceNorm2d(64, affine=True)
        else:
            self.bn1 = BN(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = maxpool2d_for(ops, fast_bn)(3, 2, 1)
        kw = dict(pos=pos, beta=beta, crop=crop, cnsn_type=cnsn_type, ops=ops, fuse_post=fuse_post,
                  ibn_host=any(v is not None for v in ibn_cfg), fast_bn=fast_bn)
        width = 64
        stages = []
        for i, (planes, count) in enumerate(zip((64, 128, 256, 512), layers)):
            stride = 1 if i == 0 else 2
            blocks = []
            for b in range(count):
                s = stride if b == 0 else 1
                shortcut = None
                if b == 0 and (s != 1 or width != planes * _EXPANSION):
                    shortcut = nn.Sequential(nn.Conv2d(width, planes * _EXPANSION, 1, s, bias=False),
                                             BN(planes * _EXPANSION))
                ibn = ibn_cfg[i]
                if ibn == "b" and (b == 0 or b < count - 1):   # IBN-b: the last block of a stage, never its first (:204-214)
                    ibn = None
                blocks.append(Bottleneck(width, planes, s, shortcut, ibn=ibn, **kw))
                width = planes * _EXPANSION
            stages.append(nn.Sequential(*blocks))
        self.layer1, self.layer2, self.layer3, self.layer4 = stages
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Linear(width, num_classes)

        self.cn_modules = []                              # plain list, like the reference: not a registered container
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, (nn.BatchNorm2d, nn.InstanceNorm2d)):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
            elif isinstance(m, ops.CrossNorm):
                self.cn_modules.append(m)
        if cnsn_type is not None and "cn" in cnsn_type:
            self.cn_num, self.active_num = len(self.cn_modules), active_num
            assert self.cn_num > 0 and self.active_num > 0
        if zero_init_residual:
            for m in self.modules():
                if isinstance(m, Bottleneck):
                    nn.init.constant_(m.bn3.weight, 0)

    def _enable_cross_norm(self):
        chosen = np.random.choice(self.cn_num, self.active_num, replace=False).tolist()
        assert len(set(chosen)) == self.active_num
        for i in chosen:
            self.cn_modules[i].active = True

    def forward(self, x, aug=False):
        if aug:
            self._enable_cross_norm()
        h = self.maxpool(bn_relu(self.bn1, self.relu, self.conv1(x)))
        h = self.layer4(self.layer3(self.layer2(self.layer1(h))))
        return self.fc(torch.flatten(self.avgpool(h), 1))


def resnet50(active_num=1, pos="post", beta=1, crop="neither", cnsn_type="sn", **kw):
    """ResNet-50 + CNSN; the defaults are imagenet-scripts/run-cnsn.sh's model flags (SelfNorm at 'post')."""
    return ResNet([3, 4, 6, 3], active_num=active_num, pos=pos, beta=beta, crop=crop, cnsn_type=cnsn_type, **kw)


def resnet50_ibn_a(active_num=1, pos="post", beta=1, crop="neither", cnsn_type="sn", **kw):
    """ResNet-50-IBN-a + CNSN (models/imagenet/resnet_ibn_cnsn.py:262-275): IBN replaces bn1 in stages 1-3."""
    return ResNet([3, 4, 6, 3], active_num=active_num, pos=pos, beta=beta, crop=crop, cnsn_type=cnsn_type,
                  ibn_cfg=("a", "a", "a", None), **kw)


def resnet50_ibn_b(active_num=1, pos="post", beta=1, crop="neither", cnsn_type="sn", **kw):
    """ResNet-50-IBN-b + CNSN (the ibn_cfg=('b','b',None,None) variant of models/imagenet/resnet_ibn_cnsn.py):
    instance norm in the stem and after the residual add of the last block of stages 1-2."""
    return ResNet([3, 4, 6, 3], active_num=active_num, pos=pos, beta=beta, crop=crop, cnsn_type=cnsn_type,
                  ibn_cfg=("b", "b", None, None), **kw)
