"""nn.Module wrappers around oracle/eager_chain.py -- TEST / BASELINE INFRASTRUCTURE ONLY.

They give the CPU reference arm of bench.py (and the host-model tests) a CrossNorm / SelfNorm / CNSN
module set with the reference's surface, executing the reference's eager ATen op chain, so the host
models in the product package can be run on CPU for the baseline without the product ever depending
on this directory.
"""
import torch
import torch.nn as nn

from . import eager_chain as E
from .cnsn_oracle import rand_window


def cn_op_2ins_space_chan(x, crop="neither", beta=1, bbx_thres=0.1, lam=None, chan=False):
    """Function form (models/cnsn.py:58-91) for the image-space CrossNorm of the ImageNet step (imagenet.py:215)."""
    assert crop in ("neither", "style", "content", "both")
    perm = torch.randperm(x.size(0))
    sw = rand_window(x.shape, beta, bbx_thres) if crop in ("style", "both") else None
    cperm = torch.randperm(x.size(1)) if chan else None
    cw = rand_window(x.shape, beta, bbx_thres) if crop in ("content", "both") else None
    return E.crossnorm(x, perm, sw, cw, cperm, lam)


class CrossNorm(nn.Module):
    def __init__(self, crop=None, beta=None):
        super().__init__()
        self.active = False
        self.crop, self.beta = crop, beta

    def forward(self, x):
        if self.training and self.active:
            assert self.crop in ("neither", "style", "content", "both")
            perm = torch.randperm(x.size(0))
            sw = rand_window(x.shape, self.beta, 0.1) if self.crop in ("style", "both") else None
            cw = rand_window(x.shape, self.beta, 0.1) if self.crop in ("content", "both") else None
            x = E.crossnorm(x, perm, sw, cw)
        self.active = False
        return x


class SelfNorm(nn.Module):
    def __init__(self, chan_num, is_two=False):
        super().__init__()
        self.g_fc = nn.Conv1d(chan_num, chan_num, kernel_size=2, bias=False, groups=chan_num)
        self.g_bn = nn.BatchNorm1d(chan_num)
        self.f_fc = None

    def forward(self, x):
        b, c = x.shape[:2]
        mu, sd = E.stats(x, eps=1e-12)
        st = torch.cat((mu.squeeze(3), sd.squeeze(3)), -1)
        g = torch.sigmoid(self.g_bn(self.g_fc(st))).view(b, c, 1, 1)
        return x * g.expand_as(x)


class CNSN(nn.Module):
    def __init__(self, crossnorm, selfnorm):
        super().__init__()
        self.crossnorm, self.selfnorm = crossnorm, selfnorm

    def forward(self, x):
        if self.crossnorm and self.crossnorm.active:
            x = self.crossnorm(x)
        if self.selfnorm:
            x = self.selfnorm(x)
        return x
