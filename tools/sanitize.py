"""Small forward + backward launches of every dataflow kernel family, for compute-sanitizer:

    compute-sanitizer --tool racecheck|synccheck|memcheck python tools/sanitize.py

SelfNorm (shared-memory-resident k_sn_res, L2 items k_sn_flow, channel groups k_sn_grp, the shared + tensor memory
pipeline k_sn_tm), CrossNorm (k_cn_res), the fused site (k_site_res), IBN / BatchNorm2d (k_ibn_res, k_bn_grp), the channels-last
SelfNorm block, BatchNorm2d and MaxPool2d kernels (k_nhwc_*, k_bn_nhwc_*, k_maxpool_nhwc_*); each result is
checked against the three-kernel general path of the same library, and the asynchronous error state must stay clear."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import cnsn_b200.cnsn as M  # noqa: E402
import cnsn_b200._lib as L  # noqa: E402
from cnsn_b200.ibn import IBN, BatchNorm2d  # noqa: E402

dev = "cuda:0"
n0 = L.launch_count()


def selfnorm(shape, knobs, tag):
    torch.manual_seed(0)
    x = (torch.randn(shape, device=dev) * 1.3 + 0.2).requires_grad_(True)
    dy = torch.randn(shape, device=dev)
    sn = M.SelfNorm(shape[1]).to(dev).train()
    outs = []
    for kn in (knobs, {"selfnorm_impl": "v1"}):
        with L.tuned(**kn):
            sn.g_bn.running_mean.zero_(); sn.g_bn.running_var.fill_(1)
            y = sn(x)
            (dx,) = torch.autograd.grad(y, x, dy)
            outs.append((y.detach(), dx))
    torch.cuda.synchronize()
    err = max(float((a - b).abs().max()) for a, b in zip(*outs))
    print("%-34s %-18s max |dataflow - three-kernel| = %.2e" % (tag, shape, err), flush=True)
    assert err < 1e-4


selfnorm((8, 4, 16, 16), {"flow_mode": "res", "flow_bwd": "res"}, "SelfNorm k_sn_res")
selfnorm((8, 4, 16, 16), {"flow_mode": "res", "flow_bwd": "res", "grid_cap": 2}, "SelfNorm k_sn_res (2 CTAs loop)")
selfnorm((8, 4, 16, 16), {"flow_mode": "l2", "flow_bwd": "l2"}, "SelfNorm k_sn_flow")
selfnorm((8, 8, 7, 7), {}, "SelfNorm k_sn_grp")
selfnorm((10, 3, 40, 40), {"tm_items": 0, "grid_cap": 2}, "SelfNorm k_sn_tm")

torch.manual_seed(1)
np.random.seed(1)
x = torch.randn(8, 4, 16, 16, device=dev, requires_grad=True)
dy = torch.randn(8, 4, 16, 16, device=dev)
for crop in ("neither", "both"):
    res = []
    for kn in ({}, {"crossnorm_impl": "v1"}):
        with L.tuned(**kn):
            torch.manual_seed(2); np.random.seed(3)
            y = M.cn_op_2ins_space_chan(x, crop=crop, beta=1)
            (dx,) = torch.autograd.grad(y, x, dy)
            res.append((y.detach(), dx))
    torch.cuda.synchronize()
    err = max(float((a - b).abs().max()) for a, b in zip(*res))
    print("%-34s %-18s max |dataflow - two-kernel| = %.2e" % ("CrossNorm k_cn_res crop=" + crop, tuple(x.shape), err), flush=True)
    assert err < 1e-4
blk = M.CNSN(M.CrossNorm(crop="both", beta=1), M.SelfNorm(4)).to(dev).train()
res = []
for fused in (True, False):
    M.CNSN.fuse_site = fused
    blk.selfnorm.g_bn.running_mean.zero_(); blk.selfnorm.g_bn.running_var.fill_(1)
    torch.manual_seed(2); np.random.seed(3)
    blk.crossnorm.active = True
    y = blk(x)
    (dx,) = torch.autograd.grad(y, x, dy)
    res.append((y.detach(), dx))
M.CNSN.fuse_site = True
torch.cuda.synchronize()
err = max(float((a - b).abs().max()) for a, b in zip(*res))
print("%-34s %-18s max |fused - sequence| = %.2e" % ("fused site k_site_res", tuple(x.shape), err), flush=True)
assert err < 1e-4
for name, mod, ref in (("IBN k_ibn_res", IBN(4).to(dev).train(), None), ("BatchNorm2d k_ibn_res (half = 0)", BatchNorm2d(4).to(dev).train(), torch.nn.BatchNorm2d(4).to(dev).train())):
    y = mod(x)
    (dx,) = torch.autograd.grad(y, x, dy)
    if ref is not None:
        yr = ref(x)
        (dxr,) = torch.autograd.grad(yr, x, dy)
        err = max(float((y - yr).abs().max()), float((dx - dxr).abs().max()))
        assert err < 1e-4
        print("%-34s %-18s max |ours - torch| = %.2e" % (name, tuple(x.shape), err), flush=True)
    else:
        print("%-34s %-18s ran" % (name, tuple(x.shape)), flush=True)
# ---- the kernels added late in round 2: channel-group batch norm, channels-last SelfNorm block / BatchNorm2d / MaxPool2d
cl = torch.channels_last
xb = torch.randn(12, 8, 7, 7, device=dev, requires_grad=True)              # 7x7 planes: k_bn_grp
db = torch.randn(12, 8, 7, 7, device=dev)
bn, tbn = BatchNorm2d(8).to(dev).train(), torch.nn.BatchNorm2d(8).to(dev).train()
y = bn(xb, True); (dx,) = torch.autograd.grad(y, xb, db)
yr = torch.relu(tbn(xb)); (dxr,) = torch.autograd.grad(yr, xb, db)
err = max(float((y - yr).abs().max()), float((dx - dxr).abs().max()))
print("%-34s %-18s max |ours - torch| = %.2e" % ("BatchNorm2d + ReLU k_bn_grp", tuple(xb.shape), err), flush=True)
assert err < 1e-4
for shape in ((6, 16, 20, 20), (5, 8, 50, 50)):                           # one slab / several slabs per sample
    xc = torch.randn(shape, device=dev).contiguous(memory_format=cl).requires_grad_(True)
    rc = torch.randn(shape, device=dev).contiguous(memory_format=cl).requires_grad_(True)
    dc = torch.randn(shape, device=dev).contiguous(memory_format=cl)
    sn = M.SelfNorm(shape[1]).to(dev).train()
    outs = []
    for xx, rr, dd in ((xc, rc, dc), (xc.detach().contiguous().requires_grad_(True), rc.detach().contiguous().requires_grad_(True), dc.contiguous())):
        sn.g_bn.running_mean.zero_(); sn.g_bn.running_var.fill_(1)
        y = sn(xx, rr, True)
        dx, dr = torch.autograd.grad(y, (xx, rr), dd)
        outs.append((y.detach(), dx, dr))
    err = max(float((a - b).abs().max()) for a, b in zip(*outs))
    print("%-34s %-18s max |NHWC - NCHW kernels| = %.2e" % ("SelfNorm block k_nhwc_*", shape, err), flush=True)
    assert err < 1e-4
    bn, tbn = BatchNorm2d(shape[1]).to(dev).train(), torch.nn.BatchNorm2d(shape[1]).to(dev).train()
    y = bn(xc, True); (dx,) = torch.autograd.grad(y, xc, dc)
    yr = torch.relu(tbn(xc)); (dxr,) = torch.autograd.grad(yr, xc, dc)
    err = max(float((y - yr).abs().max()), float((dx - dxr).abs().max()))
    print("%-34s %-18s max |ours - torch| = %.2e" % ("BatchNorm2d + ReLU k_bn_nhwc_*", shape, err), flush=True)
    assert err < 1e-4
from cnsn_b200.hosts._norm import MaxPool2d  # noqa: E402
xp = torch.relu(torch.randn(4, 8, 13, 11, device=dev)).contiguous(memory_format=cl).requires_grad_(True)
y = MaxPool2d(3, 2, 1)(xp)
dp = torch.randn_like(y)
(dx,) = torch.autograd.grad(y, xp, dp)
yr = torch.nn.functional.max_pool2d(xp, 3, 2, 1)
(dxr,) = torch.autograd.grad(yr, xp, dp)
assert torch.equal(y, yr) and torch.equal(dx, dxr)
print("%-34s %-18s equal to torch" % ("MaxPool2d k_maxpool_nhwc_*", tuple(xp.shape)), flush=True)
torch.cuda.synchronize()
L.async_error()
print("sanitize.py: %d library kernels launched, asynchronous error state clear" % (L.launch_count() - n0))
