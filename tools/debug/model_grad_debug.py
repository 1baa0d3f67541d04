"""Debug: per-parameter gradient error of a WRN-40-2 step, ours (fp32 kernels) and the reference in fp32, both against
the reference run in fp64 on the same GPU."""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import build_ref  # noqa: E402
import cnsn_b200.cnsn as M  # noqa: E402
from cnsn_b200.hosts import WideResNet  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.deterministic = True
DEV = "cuda:0"
RefWRN = build_ref.load("models.cifar.wideresnet_cnsn").WideResNet
kw = dict(widen_factor=2, active_num=2, pos="post", beta=1, crop="both", cnsn_type=sys.argv[1] if len(sys.argv) > 1 else "cnsn")
aug = len(sys.argv) > 2 and sys.argv[2] == "aug"
torch.manual_seed(0)
a = RefWRN(40, 10, **kw).to(DEV).train()
a64 = RefWRN(40, 10, **kw).to(DEV).double().train()
a64.load_state_dict(a.state_dict())
b = WideResNet(40, 10, fuse_post=False, **kw).to(DEV).train()
b.load_state_dict(a.state_dict())
M.CNSN.fuse_site = False
g = torch.Generator().manual_seed(1)
x = torch.randn(64, 3, 32, 32, generator=g).to(DEV)
y = torch.randint(0, 10, (64,), generator=g).to(DEV)
res = []
for net, xx in ((a64, x.double()), (a, x), (b, x)):
    torch.manual_seed(5)
    np.random.seed(6)
    logits = net(xx, aug=aug)
    loss = F.cross_entropy(logits, y)
    loss.backward()
    res.append({k: p.grad.double() for k, p in net.named_parameters()})
names = list(res[0])


def rel(u, v):
    return float((u - v).abs().max() / v.abs().max().clamp_min(1e-30))


print("%-50s %12s %12s %12s" % ("parameter (reverse order)", "ref32-vs-64", "ours-vs-64", "ours-vs-ref32"))
for k in reversed(names):
    if "conv" in k or "fc" in k:
        print("%-50s %12.3e %12.3e %12.3e" % (k, rel(res[1][k], res[0][k]), rel(res[2][k], res[0][k]), rel(res[2][k], res[1][k])))
r1 = np.median([rel(res[1][k], res[0][k]) for k in names])
r2 = np.median([rel(res[2][k], res[0][k]) for k in names])
print("median ref32-vs-64 %.3e  ours-vs-64 %.3e" % (r1, r2))
