"""IBN layer timing (CUDA events, module API): cnsn_b200.ibn.IBN vs the reference's composition (split / contiguous /
nn.InstanceNorm2d / nn.BatchNorm2d / cat, models/imagenet/resnet_ibn_cnsn.py:24-44) on the same GPU.

    python tools/perf_ibn.py [N,C,H,W] [steps]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn as nn  # noqa: E402
from cnsn_b200.ibn import IBN  # noqa: E402
import cnsn_b200._lib as _L  # noqa: E402

_L.tune_from_env()


class ReferenceIBN(nn.Module):
    def __init__(self, planes, ratio=0.5):
        super().__init__()
        self.half = int(planes * ratio)
        self.IN = nn.InstanceNorm2d(self.half, affine=True)
        self.BN = nn.BatchNorm2d(planes - self.half)

    def forward(self, x):
        a, b = torch.split(x, self.half, 1)
        return torch.cat((self.IN(a.contiguous()), self.BN(b.contiguous())), 1)


shape = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "256,64,56,56").split(","))
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
x = torch.randn(shape, device="cuda").requires_grad_(True)
dy = torch.randn(shape, device="cuda")
S = x.numel() * 4
out = []
for name, m in (("cnsn_b200 IBN", IBN(shape[1]).cuda().train()), ("reference composition (torch)", ReferenceIBN(shape[1]).cuda().train())):
    for _ in range(5):
        torch.autograd.grad(m(x), x, dy)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
    torch.cuda.synchronize()
    for e in ev:
        e[0].record(); y = m(x); e[1].record(); torch.autograd.grad(y, x, dy); e[2].record()
    torch.cuda.synchronize()
    f = sorted(e[0].elapsed_time(e[1]) for e in ev)[steps // 2]
    b = sorted(e[1].elapsed_time(e[2]) for e in ev)[steps // 2]
    out.append(f + b)
    print("%-30s %s | fwd %.1f us  bwd %.1f us | fwd+bwd %.0f GB/s (5*S)" % (name, shape, f * 1e3, b * 1e3, 5 * S / (f + b) / 1e6))
print("speed-up x%.2f" % (out[1] / out[0]))
