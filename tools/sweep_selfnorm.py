"""A/B sweep of SelfNorm implementations and knobs in ONE process (CUDA events, inputs larger than L2).

    python tools/sweep_selfnorm.py N,C,H,W f32|bf16 steps "K1=V1 K2=V2" "K1=V3" ...

Each quoted argument is one configuration: tuning knobs of the library (cnsn_tune; struct Knobs in
csrc/flow_common.cuh), e.g. "flow_mode=res i3=1" or "flow_bwd=dyg".  "-" = defaults.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import cnsn_b200.cnsn as M  # noqa: E402
import cnsn_b200._lib as L  # noqa: E402

shape = tuple(int(v) for v in sys.argv[1].split(","))
dt = torch.bfloat16 if sys.argv[2] == "bf16" else torch.float32
steps = int(sys.argv[3])
configs = sys.argv[4:] or ["-"]
dev = "cuda:0"
N, C, H, W = shape
g = torch.Generator(device=dev).manual_seed(0)
x = (torch.randn(shape, device=dev, generator=g) * (0.5 + 1.5 * torch.rand(N, C, 1, 1, device=dev, generator=g))
     + torch.randn(N, C, 1, 1, device=dev, generator=g)).to(dt).requires_grad_(True)
dy = torch.randn(shape, device=dev, generator=g).to(dt)
sn = M.SelfNorm(C).to(dev).train()
S = x.numel() * x.element_size()
ref = None
for cfg in configs:
    kv = dict(p.split("=", 1) for p in cfg.split()) if cfg != "-" else {}
    L.tune(reset=1)
    L.tune(**kv)
    try:
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
        for i in range(3):
            y = sn(x)
            (dx,) = torch.autograd.grad(y, x, dy)
        torch.cuda.synchronize()
        for i in range(steps):
            ev[i][0].record()
            y = sn(x)
            ev[i][1].record()
            (dx,) = torch.autograd.grad(y, x, dy)
            ev[i][2].record()
        torch.cuda.synchronize()
        f = sorted(e[0].elapsed_time(e[1]) for e in ev)
        b = sorted(e[1].elapsed_time(e[2]) for e in ev)
        fm, bm = f[len(f) // 2], b[len(b) // 2]
        chk = ""
        if ref is None:
            ref = (y.detach().float().clone(), dx.float().clone())
        else:
            chk = " | maxdiff y %.2e dx %.2e" % ((y.detach().float() - ref[0]).abs().max().item(),
                                                (dx.float() - ref[1]).abs().max().item())
        print("%s %s [%s] fwd %.3f ms (min %.3f) %.0f GB/s | bwd %.3f ms (min %.3f) %.0f GB/s | fwd+bwd %.0f GB/s%s" % (
            shape, str(dt).split(".")[-1], cfg, fm, f[0], 2 * S / fm / 1e6, bm, b[0], 3 * S / bm / 1e6,
            5 * S / (fm + bm) / 1e6, chk), flush=True)
    except Exception as e:  # noqa: BLE001
        print("%s [%s] FAILED: %r" % (shape, cfg, e), flush=True)
        break
    finally:
        L.tune(reset=1)
