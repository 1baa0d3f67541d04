"""Host cost of one operator call through the module surface: wall-clock per call with the GPU kept idle-free (the
loop is asynchronous; when the host is the bottleneck, wall-clock per call = host time per call), for both bindings.

    python tools/perf_host.py [iters]
BASELINE config 2 (CrossNorm (128,64,32,32) bf16, no crop), a WideResNet site with crops, SelfNorm at a WRN site."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import cnsn_b200.cnsn as M  # noqa: E402
import cnsn_b200._lib as L  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 300
dev = "cuda:0"


def timeit(fwd, bwd_of):
    for _ in range(20):
        bwd_of(fwd())
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ys = [fwd() for _ in range(iters)]
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    for y in ys:
        bwd_of(y)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(iters)]
    for e in ev:
        e[0].record()
        y = fwd()
        e[1].record()
        bwd_of(y)
        e[2].record()
    torch.cuda.synchronize()
    f = sorted(e[0].elapsed_time(e[1]) for e in ev)[iters // 2] * 1e3
    b = sorted(e[1].elapsed_time(e[2]) for e in ev)[iters // 2] * 1e3
    return (t1 - t0) / iters * 1e6, (t2 - t1) / iters * 1e6, f, b


for binding in ("ext", "ctypes"):
    L.set_binding(binding)
    torch.manual_seed(0)
    np.random.seed(0)
    x = torch.randn(128, 64, 32, 32, device=dev).to(torch.bfloat16).requires_grad_(True)
    dy = torch.randn(128, 64, 32, 32, device=dev).to(torch.bfloat16)
    r = timeit(lambda: M.cn_op_2ins_space_chan(x, crop="neither", beta=1), lambda y: torch.autograd.grad(y, x, dy))
    print("[%s] cfg2 CrossNorm (128,64,32,32) bf16 neither : host wall fwd %.1f us bwd %.1f us | CUDA events fwd %.1f us bwd %.1f us" % (binding, *r))
    x = torch.randn(512, 32, 32, 32, device=dev).requires_grad_(True)
    dy = torch.randn(512, 32, 32, 32, device=dev)
    r = timeit(lambda: M.cn_op_2ins_space_chan(x, crop="both", beta=1), lambda y: torch.autograd.grad(y, x, dy))
    print("[%s] CrossNorm (512,32,32,32) f32 both          : host wall fwd %.1f us bwd %.1f us | CUDA events fwd %.1f us bwd %.1f us" % (binding, *r))
    sn = M.SelfNorm(32).to(dev).train()
    r = timeit(lambda: sn(x), lambda y: torch.autograd.grad(y, x, dy))
    print("[%s] SelfNorm (512,32,32,32) f32                : host wall fwd %.1f us bwd %.1f us | CUDA events fwd %.1f us bwd %.1f us" % (binding, *r))
    x2 = torch.randn(512, 128, 8, 8, device=dev).requires_grad_(True)
    dy2 = torch.randn(512, 128, 8, 8, device=dev)
    sn2 = M.SelfNorm(128).to(dev).train()
    r = timeit(lambda: sn2(x2), lambda y: torch.autograd.grad(y, x2, dy2))
    print("[%s] SelfNorm (512,128,8,8) f32                 : host wall fwd %.1f us bwd %.1f us | CUDA events fwd %.1f us bwd %.1f us" % (binding, *r))
    blk = M.CNSN(M.CrossNorm(crop="both", beta=1), M.SelfNorm(32)).to(dev).train()

    def site():
        blk.crossnorm.active = True
        return blk(x)
    r = timeit(site, lambda y: torch.autograd.grad(y, x, dy))
    print("[%s] fused site (512,32,32,32) f32 both         : host wall fwd %.1f us bwd %.1f us | CUDA events fwd %.1f us bwd %.1f us" % (binding, *r))
