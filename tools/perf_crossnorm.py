"""Timing of CrossNorm forward / backward (CUDA events; median of `steps`).

    python tools/perf_crossnorm.py [N,C,H,W] [f32|bf16] [crop] [steps]
Also times the eager-PyTorch op chain of the reference (oracle/eager_chain.py) on the same GPU tensors
for context (tool only; the product never imports oracle/)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import cnsn_b200.cnsn as M  # noqa: E402
import cnsn_b200._lib as _L  # noqa: E402

_L.tune_from_env()                 # CNSN_TUNE_<KNOB>=value -> cnsn_tune
from oracle import eager_chain as E  # noqa: E402

shape = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "128,64,32,32").split(","))
dt = torch.float32 if len(sys.argv) > 2 and sys.argv[2] == "f32" else torch.bfloat16
crop = sys.argv[3] if len(sys.argv) > 3 else "neither"
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 50
dev = "cuda:0"
x = torch.randn(shape, device=dev).to(dt).requires_grad_(True)
dy = torch.randn(shape, device=dev).to(dt)
S = x.numel() * x.element_size()


def run(fn):
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
    for _ in range(5):
        torch.autograd.grad(fn(x), x, dy)
    torch.cuda.synchronize()
    for e in ev:
        e[0].record()
        y = fn(x)
        e[1].record()
        torch.autograd.grad(y, x, dy)
        e[2].record()
    torch.cuda.synchronize()
    f = sorted(e[0].elapsed_time(e[1]) for e in ev)[steps // 2]
    b = sorted(e[1].elapsed_time(e[2]) for e in ev)[steps // 2]
    return f, b


torch.manual_seed(0)
np.random.seed(0)
f, b = run(lambda t: M.cn_op_2ins_space_chan(t, crop=crop, beta=1))
print("cnsn_b200 CrossNorm %s %s crop=%s | fwd %.1f us  bwd %.1f us | fwd+bwd %.0f GB/s (5*S = %.1f MB)" % (
    shape, str(dt).split(".")[-1], crop, f * 1e3, b * 1e3, 5 * S / (f + b) / 1e6, 5 * S / 1e6))


def eager(t):
    perm = torch.randperm(t.size(0))
    from oracle.cnsn_oracle import rand_window
    sw = rand_window(t.shape, 1, 0.1) if crop in ("style", "both") else None
    cw = rand_window(t.shape, 1, 0.1) if crop in ("content", "both") else None
    return E.crossnorm(t, perm, sw, cw)


f2, b2 = run(eager)
print("eager-PyTorch chain on the same GPU            | fwd %.1f us  bwd %.1f us | fwd+bwd %.0f GB/s | speed-up %.1fx" % (
    f2 * 1e3, b2 * 1e3, 5 * S / (f2 + b2) / 1e6, (f2 + b2) / (f + b)))
