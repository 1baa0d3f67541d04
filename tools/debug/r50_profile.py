"""Where the GPU time of a ResNet-50 + SelfNorm step (batch 256, fp32) goes (torch.profiler, CUDA time by kernel)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402
from cnsn_b200.hosts.resnet import resnet50  # noqa: E402

dev = torch.device("cuda", 0)
torch.backends.cudnn.benchmark = True
torch.manual_seed(0)
np.random.seed(0)
net = resnet50(fuse_post=True).to(dev).train()
cl = len(sys.argv) > 1 and sys.argv[1] == "cl"
if cl:
    net = net.to(memory_format=torch.channels_last)
opt = torch.optim.SGD(net.parameters(), 0.1, momentum=0.9, weight_decay=1e-4)
x = torch.randn(256, 3, 224, 224, device=dev)
if cl:
    x = x.contiguous(memory_format=torch.channels_last)
y = torch.randint(0, 1000, (256,), device=dev)


def step():
    loss = F.cross_entropy(net(x, aug=False), y)
    opt.zero_grad()
    loss.backward()
    opt.step()
    return loss


for _ in range(4):
    step()
torch.cuda.synchronize()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(5):
    step()
t1.record()
torch.cuda.synchronize()
print("ResNet-50 + SN, batch 256: %.2f ms/step" % (t0.elapsed_time(t1) / 5))
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=36, max_name_column_width=100))
