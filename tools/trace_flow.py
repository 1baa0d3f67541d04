"""Per-item timeline of the shared-memory-resident SelfNorm kernel (debug knob: cnsn_tune("trace", 1) + $CNSN_FLOW_TRACE = output path).

    python tools/trace_flow.py [N,C,H,W] [fwd|bwd] [out.bin]
Prints median / p90 of every phase of an item's life and the per-channel critical path.
"""
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import cnsn_b200.cnsn as M  # noqa: E402
import cnsn_b200._lib as L  # noqa: E402

shape = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "256,256,56,56").split(","))
bwd = len(sys.argv) > 2 and sys.argv[2] == "bwd"
out = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "gpurun_out", "flow_trace.bin")
os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
if (shape[2] * shape[3] * 4) % 16 == 0:
    L.tune(flow_mode="res", flow_bwd="res", tm=0)      # the shared-memory-resident kernel; odd planes: the channel-group kernel
L.tune_from_env()
x = torch.randn(shape, device="cuda:0").requires_grad_(True)
dy = torch.randn(shape, device="cuda:0")
sn = M.SelfNorm(shape[1]).cuda().train()
for _ in range(3):
    torch.autograd.grad(sn(x), x, dy)
torch.cuda.synchronize()
if bwd:
    y = sn(x)
    os.environ["CNSN_FLOW_TRACE"] = out
    L.tune(trace=1)
    torch.autograd.grad(y, x, dy)
else:
    os.environ["CNSN_FLOW_TRACE"] = out
    L.tune(trace=1)
    sn(x)
torch.cuda.synchronize()
L.tune(trace=0)
raw = open(out, "rb").read()
items, nI, per_sm, isb = struct.unpack("4i", raw[:16])
t = np.frombuffer(raw[16:], dtype=np.uint64).reshape(items, 8).astype(np.float64)
t0 = t[:, 0].min()
t = (t - t0) / 1e3                       # us
phases = [("load (ticket -> landed)", 0, 1), ("reduce + publish", 1, 2), ("wait for channel", 2, 4), ("apply", 4, 5)]
print("items %d, %d per channel, %d CTAs/SM, %s; kernel span %.1f us" % (items, nI, per_sm, "bwd" if isb else "fwd", t[:, 5].max()))
for nm, i0, i1 in phases:
    d = t[:, i1] - t[:, i0]
    print("  %-26s median %6.2f  p90 %6.2f  max %6.2f us" % (nm, np.median(d), np.percentile(d, 90), d.max()))
life = t[:, 5] - t[:, 0]
print("  %-26s median %6.2f  p90 %6.2f  max %6.2f us" % ("lifetime", np.median(life), np.percentile(life, 90), life.max()))
C = items // nI
tc = t.reshape(C, nI, 8)
first = tc[:, :, 0].min(axis=1)
lastd = tc[:, :, 0].max(axis=1)
landed = tc[:, :, 1].max(axis=1)
published = tc[:, :, 2].max(axis=1)
words = tc[:, :, 3].max(axis=1)           # folds done (resident kernel: the folder is the channel's last item)
known = tc[:, nI - 1, 4]
seen = tc[:, :, 4].max(axis=1)
done = tc[:, :, 5].max(axis=1)
print("per channel (median over channels, us after the channel's first ticket):")
for nm, v in (("last ticket taken", lastd), ("last plane landed", landed), ("last item published", published),
              ("folder saw all words", words), ("folder published consts", known), ("last item saw consts", seen),
              ("last item applied", done)):
    print("  %-24s %6.2f  (p90 %6.2f)" % (nm, np.median(v - first), np.percentile(v - first, 90)))
print("channel completion rate: %.3f us per channel" % ((done.max() - done.min()) / max(1, C - 1)))
for c in (C // 2, C // 2 + 1):
    print("channel %d: first ticket %.1f last ticket %.1f landed %.1f published %.1f words %.1f consts %.1f applied %.1f" % (
        c, first[c], lastd[c], landed[c], published[c], words[c], known[c], done[c]))
