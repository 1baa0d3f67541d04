"""Channels-last kernels (csrc/selfnorm_nhwc.cu, csrc/bn_nhwc.cu) through the module surface, CUDA events, median of `reps`:
the SelfNorm block tail relu(SelfNorm(x + res)) and BatchNorm2d + ReLU, NHWC against this package's NCHW kernels and (batch
norm) against torch's channels_last batch norm + relu (cuDNN).

    python tools/perf_nhwc.py [N,C,H,W dtype]...        default: the WideResNet-40-2 and ResNet-50 site shapes
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn as nn  # noqa: E402
import cnsn_b200.cnsn as M  # noqa: E402
from cnsn_b200.ibn import BatchNorm2d  # noqa: E402

dev = "cuda:0"
reps = int(os.environ.get("PERF_REPS", 20))
cl = torch.channels_last
specs = [((512, 32, 32, 32), torch.float32), ((512, 64, 16, 16), torch.float32), ((512, 128, 8, 8), torch.float32),
         ((256, 256, 56, 56), torch.float32), ((256, 512, 28, 28), torch.float32), ((256, 2048, 7, 7), torch.float32),
         ((768, 256, 56, 56), torch.bfloat16), ((768, 1024, 14, 14), torch.bfloat16)]
if len(sys.argv) > 2:
    specs = [(tuple(int(v) for v in sys.argv[i].split(",")), torch.float32 if sys.argv[i + 1] == "f32" else torch.bfloat16)
             for i in range(1, len(sys.argv) - 1, 2)]


def timed(fwd, x, dy, extra=()):
    for _ in range(3):
        y = fwd()
        torch.autograd.grad(y, (x,) + tuple(extra), dy)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(reps)]
    torch.cuda.synchronize()
    for e in ev:
        e[0].record()
        y = fwd()
        e[1].record()
        torch.autograd.grad(y, (x,) + tuple(extra), dy)
        e[2].record()
    torch.cuda.synchronize()
    f = sorted(e[0].elapsed_time(e[1]) for e in ev)[reps // 2] * 1e3
    b = sorted(e[1].elapsed_time(e[2]) for e in ev)[reps // 2] * 1e3
    return f, b


for shape, dt in specs:
    N, C, H, W = shape
    S = N * C * H * W * (4 if dt == torch.float32 else 2)
    x0 = torch.randn(shape, device=dev).to(dt)
    r0 = torch.randn(shape, device=dev).to(dt)
    d0 = torch.randn(shape, device=dev).to(dt)
    sn = M.SelfNorm(C).to(dev).train()
    bn = BatchNorm2d(C).to(dev).train()
    tbn = nn.BatchNorm2d(C).to(dev).train()
    row = ["%-20s %-8s S = %7.1f MB" % (shape, str(dt).split(".")[-1], S / 1e6)]
    for fmt, name in ((cl, "NHWC"), (torch.contiguous_format, "NCHW")):
        x = x0.contiguous(memory_format=fmt).requires_grad_(True)
        r = r0.contiguous(memory_format=fmt).requires_grad_(True)
        dy = d0.contiguous(memory_format=fmt)
        f, b = timed(lambda: sn(x, r, True), x, dy, (r,))
        row.append("  SelfNorm block %s: fwd %7.1f us (%4.0f GB/s of 4 S) bwd %7.1f us (%4.0f GB/s of 3 S)" % (name, f, 4 * S / f / 1e3, b, 3 * S / b / 1e3))
        f, b = timed(lambda: bn(x, True), x, dy)
        row.append("  BatchNorm2d + ReLU %s: fwd %7.1f us (%4.0f GB/s of 2 S) bwd %7.1f us (%4.0f GB/s of 3 S)" % (name, f, 2 * S / f / 1e3, b, 3 * S / b / 1e3))
    x = x0.contiguous(memory_format=cl).requires_grad_(True)
    f, b = timed(lambda: torch.relu(tbn(x)), x, d0.contiguous(memory_format=cl))
    row.append("  torch batch_norm + relu, channels_last (cuDNN): fwd %7.1f us bwd %7.1f us" % (f, b))
    print("\n".join(row), flush=True)
