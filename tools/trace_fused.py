"""Debug: run one fused SelfNorm forward with CNSN_FUSED_TRACE and analyse the per-CTA, per-group
timeline (globaltimer ns): which CTAs are last to publish a group, and where their time goes."""
import os
import struct
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import cnsn_b200.cnsn as M  # noqa: E402

shape = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "256,256,56,56").split(","))
out = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/trace.bin"
os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
x = torch.randn(shape, device="cuda:0")
sn = M.SelfNorm(shape[1]).cuda().train()
os.environ["CNSN_SELFNORM_IMPL"] = "persistent"      # the traced kernels are the persistent ones (selfnorm_fused.cu)
for _ in range(3):
    sn(x)
torch.cuda.synchronize()
bwd = os.environ.get("TRACE_BWD") is not None
if bwd:
    xr = x.clone().requires_grad_(True)
    dy = torch.randn_like(x)
    for _ in range(2):
        torch.autograd.grad(sn(xr), xr, dy)
    torch.cuda.synchronize()
    y = sn(xr)
    os.environ["CNSN_FUSED_TRACE_BWD"] = out
    torch.autograd.grad(y, xr, dy)
    torch.cuda.synchronize()
    del os.environ["CNSN_FUSED_TRACE_BWD"]
else:
    os.environ["CNSN_FUSED_TRACE"] = out
    sn(x)
    torch.cuda.synchronize()
    del os.environ["CNSN_FUSED_TRACE"]
raw = open(out, "rb").read()
G, S, B, kk = struct.unpack("4i", raw[:16])
t = np.frombuffer(raw[16:], dtype=np.uint64).reshape(B, G, 8).astype(np.int64)
t0 = t[t > 0].min()
t = np.where(t > 0, t - t0, -1)
names = ["issue", "landed", "stats_done", "pairs_ok", "chan_ready", "apply_start", "apply_done"]
print("G=%d S=%d B=%d kk=%d total %.1f us" % (G, S, B, kk, t.max() / 1e3))
d = {"load": t[:, :, 1] - t[:, :, 0], "stats": t[:, :, 2] - t[:, :, 1], "wait_pairs": t[:, :, 3] - t[:, :, 2],
     "chan": t[:, :, 4] - t[:, :, 3], "apply": t[:, :, 6] - t[:, :, 5], "cycle": t[:, :, 6] - t[:, :, 0]}
print("mean over all CTAs/groups:", {k: int(v.mean()) for k, v in d.items()})
m7 = t[:, :, 7] > 0
print("stats split: landed->passes done %.0f ns, passes done->stats_done %.0f ns" % (
    (t[:, :, 7] - t[:, :, 1])[m7].mean(), (t[:, :, 2] - t[:, :, 7])[m7].mean()))
sd = t[:, :, 2]                       # stats_done per CTA per group
last = sd.argmax(axis=0)              # which CTA publishes last
spread = sd.max(axis=0) - np.median(sd, axis=0)
print("median->last publish spread per group: mean %.0f ns, p90 %.0f" % (spread.mean(), np.percentile(spread, 90)))
cnt = np.bincount(last, minlength=B)
print("CTAs most often last:", [(int(i), int(cnt[i])) for i in np.argsort(-cnt)[:10]])
w = d["wait_pairs"].mean(axis=1)
print("per-CTA mean wait_pairs: min %.0f  median %.0f  max %.0f" % (w.min(), np.median(w), w.max()))
slow = np.argsort(w)[:5]
for c in list(slow) + [0, B // 2]:
    print("CTA %3d: load %5d stats %5d wait %5d chan %5d apply %5d cycle %6d  (times last: %d)" % (
        c, d["load"][c].mean(), d["stats"][c].mean(), d["wait_pairs"][c].mean(), d["chan"][c].mean(),
        d["apply"][c].mean(), d["cycle"][c].mean(), cnt[c]))
g = G // 2
order = np.argsort(sd[:, g])
print("group %d: stats_done spread: first %d, median %d, last %d (cta %d); issue of last cta %d vs median issue %d; landed last %d vs median %d" % (
    g, sd[order[0], g], np.median(sd[:, g]), sd[order[-1], g], order[-1], t[order[-1], g, 0], np.median(t[:, g, 0]),
    t[order[-1], g, 1], np.median(t[:, g, 1])))
