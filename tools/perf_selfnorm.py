"""Quick A/B timing of SelfNorm forward / backward (CUDA events, inputs larger than L2 by default).

    python tools/perf_selfnorm.py [N,C,H,W] [f32|bf16] [steps]
Tuning knobs: CNSN_TUNE_<KNOB>=value in this tool's environment (forwarded through cnsn_tune), e.g. CNSN_TUNE_FLOW_MODE=res CNSN_TUNE_I3=1.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import cnsn_b200.cnsn as M  # noqa: E402
import cnsn_b200._lib as _L  # noqa: E402

_L.tune_from_env()                 # CNSN_TUNE_<KNOB>=value -> cnsn_tune

shape = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "256,256,56,56").split(","))
dt = torch.bfloat16 if len(sys.argv) > 2 and sys.argv[2] == "bf16" else torch.float32
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 30
dev = "cuda:0"
N, C, H, W = shape
g = torch.Generator(device=dev).manual_seed(0)
x = (torch.randn(shape, device=dev, generator=g) * (0.5 + 1.5 * torch.rand(N, C, 1, 1, device=dev, generator=g))
     + torch.randn(N, C, 1, 1, device=dev, generator=g)).to(dt).requires_grad_(True)
dy = torch.randn(shape, device=dev, generator=g).to(dt)
sn = M.SelfNorm(C).to(dev).train(os.environ.get("PERF_EVAL") is None)
S = x.numel() * x.element_size()
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
for i in range(5):
    y = sn(x)
    torch.autograd.grad(y, x, dy)
torch.cuda.synchronize()
for i in range(steps):
    ev[i][0].record()
    y = sn(x)
    ev[i][1].record()
    torch.autograd.grad(y, x, dy)
    ev[i][2].record()
torch.cuda.synchronize()
f = sorted(e[0].elapsed_time(e[1]) for e in ev)
b = sorted(e[1].elapsed_time(e[2]) for e in ev)
fm, bm = f[len(f) // 2], b[len(b) // 2]
tag = "%s%s" % ("EVAL " if os.environ.get("PERF_EVAL") else "", " ".join("%s=%s" % (k[10:].lower(), v) for k, v in sorted(os.environ.items()) if k.startswith("CNSN_TUNE_")) or "defaults")
print("%s %s %s | fwd %.3f ms (min %.3f) %.0f GB/s | bwd %.3f ms (min %.3f) %.0f GB/s | fwd+bwd %.0f GB/s" % (
    shape, str(dt).split(".")[-1], tag, fm, f[0], 2 * S / fm / 1e6, bm, b[0], 3 * S / bm / 1e6, 5 * S / (fm + bm) / 1e6))
