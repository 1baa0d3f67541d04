"""Timing of a CNSN site whose CrossNorm and SelfNorm both fire: the fused site kernels (cnsn_site_fwd/_bwd)
against this package's two-operator sequence, through the module API (CUDA events; median of `steps`).

    python tools/perf_site.py [steps] [case indices, e.g. 0,4]

Algorithmic bytes of the fused site: 2*S forward + 3*S backward; the sequence moves 4*S + 6*S."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import cnsn_b200.cnsn as M  # noqa: E402
import cnsn_b200._lib as _L  # noqa: E402

_L.tune_from_env()                 # CNSN_TUNE_<KNOB>=value -> cnsn_tune

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
dev = "cuda:0"
CASES = [((512, 32, 32, 32), torch.float32, "both"), ((512, 64, 16, 16), torch.float32, "both"),
         ((512, 128, 8, 8), torch.float32, "both"), ((512, 32, 32, 32), torch.bfloat16, "both"),
         ((256, 256, 56, 56), torch.float32, "neither"), ((256, 512, 28, 28), torch.float32, "both")]


def run(blk, x, dy):
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
    for _ in range(5):
        blk.crossnorm.active = True
        torch.autograd.grad(blk(x), x, dy)
    torch.cuda.synchronize()
    for e in ev:
        blk.crossnorm.active = True
        e[0].record()
        y = blk(x)
        e[1].record()
        torch.autograd.grad(y, x, dy)
        e[2].record()
    torch.cuda.synchronize()
    f = sorted(e[0].elapsed_time(e[1]) for e in ev)[steps // 2]
    b = sorted(e[1].elapsed_time(e[2]) for e in ev)[steps // 2]
    return f, b


def cabi(x, dy, sn, crop):
    """Kernel-limited: the backend calls directly (no host draws, no autograd), fused site vs CrossNorm + SelfNorm."""
    import cnsn_b200._lib as L
    be = L.backend()
    N, C, H, W = x.shape
    perm = torch.randperm(N).to(torch.int32).to(dev)
    full = (0, H, 0, W)
    cw = (H // 8, H - H // 8, W // 4, W) if crop in ("content", "both") else full
    sw = (0, H - H // 4, W // 8, W - W // 8) if crop in ("style", "both") else full
    bn = sn.g_bn
    g = L.GateTensors(sn.g_fc.weight.detach(), bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var,
                      bn.num_batches_tracked)
    xd = x.detach()

    def t(fn):
        for _ in range(3):
            fn()
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(steps)]
        torch.cuda.synchronize()
        for e in ev:
            e[0].record()
            fn()
            e[1].record()
        torch.cuda.synchronize()
        return sorted(e[0].elapsed_time(e[1]) for e in ev)[steps // 2] * 1e3

    y, save = be.site_fwd(xd, perm, cw, sw, 0.0, 1e-5, g, 0.1, 1e-5, 1e-12)
    f1 = t(lambda: be.site_fwd(xd, perm, cw, sw, 0.0, 1e-5, g, 0.1, 1e-5, 1e-12))
    b1 = t(lambda: be.site_bwd(xd, dy, perm, cw, sw, 0.0, 1e-5, g, save))
    z, csave = be.crossnorm_fwd(xd, perm, None, cw, sw, 0.0, 1e-5)
    _, ssave = be.selfnorm_fwd(z, g, None, True, 0.1, 1e-5, 1e-12)
    f0 = t(lambda: be.selfnorm_fwd(be.crossnorm_fwd(xd, perm, None, cw, sw, 0.0, 1e-5)[0], g, None, True, 0.1, 1e-5, 1e-12))
    b0 = t(lambda: be.crossnorm_bwd(xd, be.selfnorm_bwd(z, dy, g, None, True, ssave)[0], perm, None, cw, sw, 0.0, csave))
    return f1, b1, f0, b0


if len(sys.argv) > 2:                                 # optional: comma-separated case indices
    CASES = [CASES[int(i)] for i in sys.argv[2].split(",")]
for shape, dt, crop in CASES:
    torch.manual_seed(0)
    np.random.seed(0)
    x = (torch.randn(shape, device=dev) * (0.5 + torch.rand(shape[0], shape[1], 1, 1, device=dev))).to(dt).requires_grad_(True)
    dy = torch.randn(shape, device=dev).to(dt)
    S = x.numel() * x.element_size()
    blk = M.CNSN(M.CrossNorm(crop=crop, beta=1), M.SelfNorm(shape[1])).to(dev).train()
    out = {}
    for fused in (True, False):
        M.CNSN.fuse_site = fused
        try:
            out[fused] = run(blk, x, dy)
        finally:
            M.CNSN.fuse_site = True
    (f1, b1), (f0, b0) = out[True], out[False]
    print("site %s %s crop=%s | fused fwd %.1f us bwd %.1f us = %.0f GB/s of 5*S | sequence fwd %.1f us bwd %.1f us | x%.2f" % (
        shape, str(dt).split(".")[-1], crop, f1 * 1e3, b1 * 1e3, 5 * S / (f1 + b1) / 1e6, f0 * 1e3, b0 * 1e3,
        (f0 + b0) / (f1 + b1)), flush=True)
    f1, b1, f0, b0 = cabi(x, dy, blk.selfnorm, crop)
    print("     kernel-limited (backend calls) | fused fwd %.1f us (%.0f GB/s) bwd %.1f us (%.0f GB/s) | sequence fwd %.1f us bwd %.1f us | x%.2f" % (
        f1, 2 * S / f1 / 1e3, b1, 3 * S / b1 / 1e3, f0, b0, (f0 + b0) / (f1 + b1)), flush=True)
