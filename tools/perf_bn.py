"""BatchNorm2d drop-in (cnsn_ibn_* with half = 0) against torch / cuDNN at the host models' shapes (CUDA events)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn as nn  # noqa: E402
from cnsn_b200.ibn import BatchNorm2d  # noqa: E402
import cnsn_b200._lib as _L  # noqa: E402

_L.tune_from_env()
dev = "cuda:0"
steps = 30
SHAPES = [((512, 16, 32, 32), torch.float32), ((512, 32, 32, 32), torch.float32), ((512, 64, 16, 16), torch.float32),
          ((512, 128, 8, 8), torch.float32), ((256, 64, 56, 56), torch.float32), ((256, 256, 56, 56), torch.float32),
          ((256, 512, 28, 28), torch.float32), ((256, 1024, 14, 14), torch.float32), ((256, 64, 112, 112), torch.float32),
          ((768, 256, 56, 56), torch.bfloat16)]
for shape, dt in SHAPES:
    x = torch.randn(shape, device=dev).to(dt).requires_grad_(True)
    dy = torch.randn(shape, device=dev).to(dt)
    S = x.numel() * x.element_size()
    row = []
    for mod in (BatchNorm2d(shape[1]).to(dev).train(), nn.BatchNorm2d(shape[1]).to(dev).train()):
        for _ in range(5):
            torch.autograd.grad(mod(x), x, dy)
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
        torch.cuda.synchronize()
        for e in ev:
            e[0].record()
            y = mod(x)
            e[1].record()
            torch.autograd.grad(y, x, dy)
            e[2].record()
        torch.cuda.synchronize()
        f = sorted(e[0].elapsed_time(e[1]) for e in ev)[steps // 2] * 1e3
        b = sorted(e[1].elapsed_time(e[2]) for e in ev)[steps // 2] * 1e3
        row.append((f, b))
    (f0, b0), (f1, b1) = row
    print("%-22s %-8s ours fwd %7.1f us bwd %7.1f us (%5.0f GB/s of 5*S) | torch fwd %7.1f us bwd %7.1f us | x%.2f" % (
        shape, str(dt).split(".")[-1], f0, b0, 5 * S / (f0 + b0) / 1e3, f1, b1, (f1 + b1) / (f0 + b0)), flush=True)
