"""Where the GPU time of a WideResNet-40-2 + CNSN step goes (torch.profiler, CUDA time by kernel)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402
from cnsn_b200.train import make_optimizer, wrn40_2  # noqa: E402

dev = torch.device("cuda", 0)
bench = len(sys.argv) > 1 and sys.argv[1] == "benchmark"
cl = len(sys.argv) > 2 and sys.argv[2] == "cl"
torch.backends.cudnn.benchmark = bench
torch.manual_seed(0)
np.random.seed(0)
net = wrn40_2(fuse_post=True).to(dev).train()
if cl:
    net = net.to(memory_format=torch.channels_last)
opt, sched = make_optimizer(net, 100)
x = torch.randn(512, 3, 32, 32, device=dev)
if cl:
    x = x.contiguous(memory_format=torch.channels_last)
y = torch.randint(0, 10, (512,), device=dev)


def step(aug=False):
    loss = F.cross_entropy(net(x, aug=aug), y)
    opt.zero_grad()
    loss.backward()
    opt.step()
    return loss


for _ in range(8):
    step()
torch.cuda.synchronize()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(10):
    step()
t1.record()
torch.cuda.synchronize()
print("cudnn.benchmark=%s channels_last=%s : %.2f ms/step (no CrossNorm steps)" % (bench, cl, t0.elapsed_time(t1) / 10))
for _ in range(3):
    step(True)
torch.cuda.synchronize()
t0.record()
for _ in range(10):
    step(True)
t1.record()
torch.cuda.synchronize()
print("cudnn.benchmark=%s channels_last=%s : %.2f ms/step (eager steps WITH CrossNorm at 2 sites)" % (bench, cl, t0.elapsed_time(t1) / 10))
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step(len(sys.argv) > 3 and sys.argv[3] == "aug")
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=90))
