"""Probe: cnsn_ibn_* general (three-kernel) path vs torch batch norm for planes that are not 16-byte multiples."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn as nn  # noqa: E402
import cnsn_b200._lib as L  # noqa: E402

ext = L.ext()
dev = "cuda:0"
steps = 20
for shape, dt in (((768, 1024, 14, 14), torch.bfloat16), ((768, 2048, 7, 7), torch.bfloat16), ((768, 256, 14, 14), torch.bfloat16),
                  ((768, 512, 7, 7), torch.bfloat16), ((256, 2048, 7, 7), torch.float32), ((256, 512, 7, 7), torch.float32)):
    x = torch.randn(shape, device=dev).to(dt).requires_grad_(True)
    dy = torch.randn(shape, device=dev).to(dt)
    C = shape[1]
    bn = nn.BatchNorm2d(C).to(dev).train()
    S = x.numel() * x.element_size()

    def ours():
        return ext.ibn(x, 0, True, False, 0.1, 1e-5, 1e-5, bn.running_mean, bn.running_var, bn.num_batches_tracked, None, None, bn.weight, bn.bias)

    rows = []
    for fn in (ours, lambda: bn(x)):
        for _ in range(3):
            torch.autograd.grad(fn(), x, dy)
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
        torch.cuda.synchronize()
        for e in ev:
            e[0].record(); y = fn(); e[1].record(); torch.autograd.grad(y, x, dy); e[2].record()
        torch.cuda.synchronize()
        rows.append((sorted(e[0].elapsed_time(e[1]) for e in ev)[steps // 2] * 1e3, sorted(e[1].elapsed_time(e[2]) for e in ev)[steps // 2] * 1e3))
    print("%-22s %-8s general path fwd %7.1f us bwd %7.1f us (%5.0f GB/s of 5*S) | torch fwd %7.1f us bwd %7.1f us" % (
        shape, str(dt).split(".")[-1], rows[0][0], rows[0][1], 5 * S / (rows[0][0] + rows[0][1]) / 1e3, rows[1][0], rows[1][1]), flush=True)
