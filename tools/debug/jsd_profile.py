"""Where the GPU time of the 3-view JSD step (ResNet-50 + SN, 768 views, bf16 autocast) goes."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402
from cnsn_b200.hosts.resnet import resnet50  # noqa: E402
from cnsn_b200.losses import jsd_consistency  # noqa: E402

dev = torch.device("cuda", 0)
torch.backends.cudnn.benchmark = True
torch.manual_seed(0)
np.random.seed(0)
net = resnet50(fuse_post=True).to(dev).train()
cl = len(sys.argv) > 1 and sys.argv[1] == "cl"
if cl:
    net = net.to(memory_format=torch.channels_last)
opt = torch.optim.SGD(net.parameters(), 0.1, momentum=0.9, weight_decay=1e-4)
B = 256
x = torch.randn(3 * B, 3, 224, 224, device=dev)
if cl:
    x = x.contiguous(memory_format=torch.channels_last)
y = torch.randint(0, 1000, (B,), device=dev)


def step():
    with torch.autocast("cuda", dtype=torch.bfloat16):
        logits = net(x, aug=False)
    lc, l1, l2 = torch.split(logits, B)
    loss = F.cross_entropy(lc, y) + 12 * jsd_consistency(lc, l1, l2)
    opt.zero_grad()
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=32, max_name_column_width=95))
