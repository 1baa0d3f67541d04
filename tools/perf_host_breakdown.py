"""Where the host time of one CrossNorm / SelfNorm call goes (wall clock per call over an asynchronous loop).
    python tools/perf_host_breakdown.py"""
import ctypes
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import cnsn_b200.cnsn as M  # noqa: E402
import cnsn_b200._lib as L  # noqa: E402

dev = "cuda:0"
it = 2000


def wall(fn, n=it):
    for _ in range(50):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    return (t1 - t0) / n * 1e6, (t2 - t0) / n * 1e6


x = torch.randn(128, 64, 32, 32, device=dev).to(torch.bfloat16)
N, C, H, W = x.shape
lib = L.lib()
y = torch.empty_like(x)
save = torch.empty(int(lib.cnsn_crossnorm_save_floats(N, C)), device=dev)
perm = torch.randperm(N).to(torch.int32).to(dev)
I4 = ctypes.c_int * 4
full = I4(0, H, 0, W)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for coop in (1, 0):
    L.tune(cooperative=coop)
    r = wall(lambda: lib.cnsn_crossnorm_fwd(ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(y.data_ptr()), 1, N, C, H, W,
                                            ctypes.c_void_p(perm.data_ptr()), None, full, full, 0.0, 1e-5,
                                            ctypes.c_void_p(save.data_ptr()), st))
    print("C ABI cnsn_crossnorm_fwd via ctypes, cooperative=%d : host %.1f us/call (with drain %.1f)" % (coop, *r))
L.tune(reset=1)
print("torch.randperm(128)            : %.1f us" % wall(lambda: torch.randperm(N))[0])
print("torch.empty_like(x)            : %.1f us" % wall(lambda: torch.empty_like(x))[0])
print("torch.empty(32768 f32)         : %.1f us" % wall(lambda: torch.empty(32768, device=dev))[0])
ext = L.ext()
full_t = (0, H, 0, W)
with torch.no_grad():
    print("ext.crossnorm, no_grad         : host %.1f us/call (with drain %.1f)" % wall(lambda: ext.crossnorm(x, full_t, full_t, 0.0, 1e-5)))
xg = x.clone().requires_grad_(True)
print("ext.crossnorm, grad            : host %.1f us/call (with drain %.1f)" % wall(lambda: ext.crossnorm(xg, full_t, full_t, 0.0, 1e-5)))
print("M.cn_op_2ins_space_chan, grad  : host %.1f us/call (with drain %.1f)" % wall(lambda: M.cn_op_2ins_space_chan(xg, crop="neither", beta=1)))
dy = torch.randn_like(x)
# backward: single-node autograd.grad vs a chain of 8 nodes in one engine run
yy = [M.cn_op_2ins_space_chan(xg, crop="neither", beta=1) for _ in range(200)]
torch.cuda.synchronize()
t0 = time.perf_counter()
for v in yy:
    torch.autograd.grad(v, xg, dy)
t1 = time.perf_counter()
torch.cuda.synchronize()
print("autograd.grad of ONE node      : host %.1f us/call" % ((t1 - t0) / 200 * 1e6))


def chain():
    v = xg
    for _ in range(8):
        v = M.cn_op_2ins_space_chan(v, crop="neither", beta=1)
    return v


yy = [chain() for _ in range(50)]
torch.cuda.synchronize()
t0 = time.perf_counter()
for v in yy:
    torch.autograd.grad(v, xg, dy)
t1 = time.perf_counter()
torch.cuda.synchronize()
print("autograd.grad of an 8-node chain: host %.1f us per node" % ((t1 - t0) / 400 * 1e6))
sn = M.SelfNorm(64).to(dev).train()
xs = torch.randn(128, 64, 32, 32, device=dev, requires_grad=True)
print("SelfNorm module fwd, grad      : host %.1f us/call (with drain %.1f)" % wall(lambda: sn(xs)))
