"""Kernel-limited timing of the C-ABI entry points: buffers allocated once, `reps` calls queued back to back,
CUDA events around the batch (so host time per call is hidden as long as it is below the kernel time).

    python tools/perf_cabi.py selfnorm|block|crossnorm N,C,H,W f32|bf16 [crop] [reps]
`block` = cnsn_selfnorm_block_fwd/_bwd: relu(SelfNorm(x + res)), algorithmic bytes 4*S forward (x, res in; z, y out),
3*S backward; the last line of its output times the unfused sequence (torch add, SelfNorm, torch relu) for context.
Tuning knobs: CNSN_TUNE_<KNOB>=value in the environment of THIS tool (it forwards them through cnsn_tune; the library itself never reads the environment).
"""
import ctypes
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from cnsn_b200 import _lib as L  # noqa: E402

L.tune_from_env()                  # CNSN_TUNE_<KNOB>=value -> cnsn_tune

op = sys.argv[1]
shape = tuple(int(v) for v in sys.argv[2].split(","))
dt = torch.float32 if sys.argv[3] == "f32" else torch.bfloat16
crop = sys.argv[4] if len(sys.argv) > 4 else "neither"
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 50
TRAIN = 0 if os.environ.get("PERF_EVAL") else 1
dev = torch.device("cuda:0")
N, C, H, W = shape
h = L.lib()
code = L._dtype_code(torch.empty(0, dtype=dt))
g = torch.Generator(device=dev).manual_seed(0)
x = (torch.randn(shape, device=dev, generator=g) * (0.5 + 1.5 * torch.rand(N, C, 1, 1, device=dev, generator=g))
     + torch.randn(N, C, 1, 1, device=dev, generator=g)).to(dt)
dy = torch.randn(shape, device=dev, generator=g).to(dt)
y = torch.empty_like(x)
dx = torch.empty_like(x)
S = x.numel() * x.element_size()
stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
P = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
f32 = dict(dtype=torch.float32, device=dev)

if op in ("selfnorm", "block"):
    w = torch.randn(C, 2, **f32) * 0.5
    gamma, beta = torch.rand(C, **f32) + 0.5, torch.randn(C, **f32) * 0.1
    rm, rv = torch.zeros(C, **f32), torch.ones(C, **f32)
    nbt = torch.zeros((), dtype=torch.int64, device=dev)
    gp = L.GateParams(*[P(t).value for t in (w, gamma, beta, rm, rv, nbt)])
    save = torch.empty(h.cnsn_selfnorm_save_floats(N, C, 0), **f32)
    ws = torch.empty(h.cnsn_selfnorm_workspace_floats(N, C, 0), **f32)
    grads = torch.empty(4 * C, **f32)
    gg = L.GateGrads(P(grads[:2 * C]).value, P(grads[2 * C:3 * C]).value, P(grads[3 * C:]).value)

    def fwd():
        L._check(h.cnsn_selfnorm_fwd(P(x), P(y), code, N, C, H, W, ctypes.byref(gp), None, TRAIN, 0.1, 1e-5, 1e-12, P(save), stream))

    def bwd():
        L._check(h.cnsn_selfnorm_bwd(P(x), P(dy), P(dx), code, N, C, H, W, ctypes.byref(gp), None, TRAIN, P(save),
                                     ctypes.byref(gg), None, P(ws), stream))
    if op == "block":
        res = torch.randn(shape, device=dev, generator=g).to(dt)
        z = torch.empty_like(x)

        def fwd():  # noqa: F811
            L._check(h.cnsn_selfnorm_block_fwd(P(x), P(res), P(z), P(y), 1, code, N, C, H, W, ctypes.byref(gp), TRAIN,
                                               0.1, 1e-5, 1e-12, P(save), stream))

        def bwd():  # noqa: F811
            L._check(h.cnsn_selfnorm_block_bwd(P(z), P(dy), P(dx), 1, code, N, C, H, W, ctypes.byref(gp), TRAIN, P(save),
                                               ctypes.byref(gg), P(ws), stream))
else:
    perm = torch.randperm(N, generator=torch.Generator().manual_seed(3)).to(torch.int32).to(dev)
    rs = np.random.RandomState(4)

    def box():
        hh, ww = max(2, int(H * 0.7)), max(2, int(W * 0.7))
        h0, w0 = int(rs.randint(0, H - hh + 1)), int(rs.randint(0, W - ww + 1))
        return (h0, h0 + hh, w0, w0 + ww)
    full = (0, H, 0, W)
    swin = box() if crop in ("style", "both") else full
    cwin = box() if crop in ("content", "both") else full
    save = torch.empty(h.cnsn_crossnorm_save_floats(N, C), **f32)
    ws = torch.empty(h.cnsn_crossnorm_workspace_floats(N, C), **f32)
    cw, sw = L._I4(*cwin), L._I4(*swin)

    def fwd():
        L._check(h.cnsn_crossnorm_fwd(P(x), P(y), code, N, C, H, W, P(perm), None, cw, sw, 0.0, 1e-5, P(save), stream))

    def bwd():
        L._check(h.cnsn_crossnorm_bwd(P(x), P(dy), P(dx), code, N, C, H, W, P(perm), None, cw, sw, 0.0, P(save), P(ws), stream))


def timeit(fn):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    host = (time.perf_counter() - t0) / reps
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3, host * 1e6


fwd()
tf, hf = timeit(fwd)
tb, hb = timeit(bwd)
tag = " ".join("%s=%s" % (k, v) for k, v in sorted(os.environ.items()) if k.startswith("CNSN_TUNE_") or k == "PERF_EVAL")
kf = 4 if op == "block" else 2
print("%s %s %s crop=%s [%s] | fwd %.1f us (host %.1f) %.0f GB/s | bwd %.1f us (host %.1f) %.0f GB/s | fwd+bwd %.0f GB/s" % (
    op, shape, str(dt).split(".")[-1], crop, tag or "-", tf, hf, kf * S / tf / 1e3, tb, hb, 3 * S / tb / 1e3, (kf + 3) * S / (tf + tb) / 1e3))
if op == "block":            # context: the same block tail unfused (torch add / relu around the fused SelfNorm kernels)
    import cnsn_b200.cnsn as M
    sn = M.SelfNorm(C).to(dev).train(bool(TRAIN))
    xr, rr = x.clone().requires_grad_(True), res.clone().requires_grad_(True)

    def seq():
        out = torch.relu(sn(torch.add(rr, xr)))
        torch.autograd.grad(out, (xr, rr), dy)

    def fus():
        out = sn(xr, rr, True)
        torch.autograd.grad(out, (xr, rr), dy)
    t1, _ = timeit(seq)
    t2, _ = timeit(fus)
    print("   module API fwd+bwd: unfused (torch add, SelfNorm, torch relu) %.1f us | fused block %.1f us | x%.2f" % (t1, t2, t1 / t2))
