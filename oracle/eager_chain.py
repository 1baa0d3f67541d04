"""Eager-PyTorch CPU port of the reference's CNSN op chain -- TEST / BASELINE INFRASTRUCTURE ONLY.

Purpose: the CPU baseline (``bench.py`` ``cpu_baseline`` and ``--impl reference``).  The reference
(amazon-science/crossnorm-selfnorm, models/cnsn.py) is a Python file over ATen; it cannot travel to
the GPU box, so this module restates the SAME sequence of ATen calls (var, mean, sqrt, index gather,
sub/div/mul/add broadcasts, depthwise conv1d, batch_norm, sigmoid, masked blend) in functional form
with explicit parameter tensors, and autograd differentiates it exactly as it does the reference.
It is validated against the live reference in tests/test_oracle_vs_reference.py (bit-identical
outputs on CPU) and is never imported by the product package.
"""
import torch
import torch.nn.functional as F


def stats(x, eps=1e-5):
    """models/cnsn.py:8-17: two separate reductions over the flattened plane."""
    n, c = x.shape[:2]
    flat = x.contiguous().view(n, c, -1)
    sd = (flat.var(dim=2) + eps).sqrt().view(n, c, 1, 1)
    mu = x.contiguous().view(n, c, -1).mean(dim=2).view(n, c, 1, 1)
    return mu, sd


def restyle(content, style):
    """models/cnsn.py:20-29."""
    assert content.shape[:2] == style.shape[:2]
    shp = content.size()
    s_mu, s_sd = stats(style)
    c_mu, c_sd = stats(content)
    return (content - c_mu.expand(shp)) / c_sd.expand(shp) * s_sd.expand(shp) + s_mu.expand(shp)


def crossnorm(x, perm, style_window=None, content_window=None, chan_perm=None, lam=None):
    """models/cnsn.py:58-91 with the random draws passed in (perm / windows as (h0,h1,w0,w1))."""
    idx = perm.to(x.device)
    if style_window is not None:
        a, b, c, d = style_window
        other = x[idx, :, a:b, c:d]
    else:
        other = x[idx]
    if chan_perm is not None:
        other = other[:, chan_perm.to(x.device), :, :]
    if content_window is not None:
        a, b, c, d = content_window
        out = torch.zeros_like(x)
        out[:, :, a:b, c:d] = restyle(x[:, :, a:b, c:d], other)
        keep = torch.ones_like(x, requires_grad=False)
        keep[:, :, a:b, c:d] = 0.
        out = x * keep + out
    else:
        out = restyle(x, other)
    if lam is not None:
        out = x * lam + out * (1 - lam)
    return out


class GateState:
    """Parameters/buffers of one SelfNorm gate (what g_fc / g_bn hold in the reference)."""

    def __init__(self, C, dtype=torch.float32, seed=0):
        g = torch.Generator().manual_seed(seed)
        self.fc_w = ((torch.rand(C, 1, 2, generator=g) * 2 - 1) * (0.5 ** 0.5)).to(dtype).requires_grad_(True)
        self.bn_w = torch.ones(C, dtype=dtype, requires_grad=True)
        self.bn_b = torch.zeros(C, dtype=dtype, requires_grad=True)
        self.run_mean = torch.zeros(C, dtype=dtype)
        self.run_var = torch.ones(C, dtype=dtype)


def selfnorm(x, g, training=True, f=None):
    """models/cnsn.py:130-150 (g: GateState; f: second gate for is_two)."""
    b, c = x.shape[:2]
    mu, sd = stats(x, eps=1e-12)
    st = torch.cat((mu.squeeze(3), sd.squeeze(3)), -1)

    def gate(p):
        v = F.conv1d(st, p.fc_w, None, groups=c)
        v = F.batch_norm(v, p.run_mean, p.run_var, p.bn_w, p.bn_b, training, 0.1, 1e-5)
        return torch.sigmoid(v).view(b, c, 1, 1)

    gy = gate(g)
    if f is not None:
        fy = gate(f)
        return x * gy.expand_as(x) + mu.expand_as(x) * (fy.expand_as(x) - gy.expand_as(x))
    return x * gy.expand_as(x)


def time_selfnorm_fwd_bwd(shape, steps, warmup, threads=None, dtype=torch.float32):
    """Wall-clock the eager chain (fwd then autograd bwd) on CPU.  Returns (seconds per step list)."""
    import time
    if threads:
        torch.set_num_threads(threads)
    N, C, H, W = shape
    gen = torch.Generator().manual_seed(0)
    x = (torch.randn(shape, generator=gen) * (0.5 + 1.5 * torch.rand(N, C, 1, 1, generator=gen))
         + torch.randn(N, C, 1, 1, generator=gen)).to(dtype).requires_grad_(True)
    dy = torch.randn(shape, generator=gen).to(dtype)
    g = GateState(C, dtype)
    times = []
    for i in range(warmup + steps):
        x.grad = None
        t0 = time.perf_counter()
        y = selfnorm(x, g, True)
        y.backward(dy)
        t1 = time.perf_counter()
        if i >= warmup:
            times.append(t1 - t0)
    return times
