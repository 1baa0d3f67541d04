"""Probe: what a channels_last (NHWC) network would save.  Plain torch (torchvision-free) WideResNet-40-2 / ResNet-50 bodies
with nn.BatchNorm2d only (no CNSN operators), NCHW against channels_last: step time and the top kernels of each."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

dev = torch.device("cuda", 0)
torch.backends.cudnn.benchmark = True
which = sys.argv[1] if len(sys.argv) > 1 else "wrn"


import torch.nn as nn  # noqa: E402


class PreAct(nn.Module):                                     # WideResNet basic block (BN-ReLU-conv x2 + shortcut)
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.bn1, self.conv1 = nn.BatchNorm2d(cin), nn.Conv2d(cin, cout, 3, stride, 1, bias=False)
        self.bn2, self.conv2 = nn.BatchNorm2d(cout), nn.Conv2d(cout, cout, 3, 1, 1, bias=False)
        self.short = None if cin == cout else nn.Conv2d(cin, cout, 1, stride, 0, bias=False)

    def forward(self, x):
        o = F.relu(self.bn1(x))
        y = self.conv1(o)
        y = self.conv2(F.relu(self.bn2(y)))
        return y + (x if self.short is None else self.short(o))


class Bottleneck(nn.Module):
    def __init__(self, cin, mid, stride):
        super().__init__()
        self.c1, self.b1 = nn.Conv2d(cin, mid, 1, bias=False), nn.BatchNorm2d(mid)
        self.c2, self.b2 = nn.Conv2d(mid, mid, 3, stride, 1, bias=False), nn.BatchNorm2d(mid)
        self.c3, self.b3 = nn.Conv2d(mid, 4 * mid, 1, bias=False), nn.BatchNorm2d(4 * mid)
        self.down = None
        if stride != 1 or cin != 4 * mid:
            self.down = nn.Sequential(nn.Conv2d(cin, 4 * mid, 1, stride, bias=False), nn.BatchNorm2d(4 * mid))

    def forward(self, x):
        y = F.relu(self.b1(self.c1(x)))
        y = F.relu(self.b2(self.c2(y)))
        y = self.b3(self.c3(y))
        return F.relu(y + (x if self.down is None else self.down(x)))


def build():
    if which == "wrn":
        layers, cin = [nn.Conv2d(3, 16, 3, 1, 1, bias=False)], 16
        for cout, stride in ((32, 1), (64, 2), (128, 2)):
            for i in range(6):
                layers.append(PreAct(cin, cout, stride if i == 0 else 1))
                cin = cout
        layers += [nn.BatchNorm2d(cin), nn.ReLU(), nn.AdaptiveAvgPool2d(1), nn.Flatten(), nn.Linear(cin, 10)]
        return nn.Sequential(*layers), torch.randn(512, 3, 32, 32, device=dev), torch.randint(0, 10, (512,), device=dev)
    layers = [nn.Conv2d(3, 64, 7, 2, 3, bias=False), nn.BatchNorm2d(64), nn.ReLU(), nn.MaxPool2d(3, 2, 1)]
    cin = 64
    for mid, n, stride in ((64, 3, 1), (128, 4, 2), (256, 6, 2), (512, 3, 2)):
        for i in range(n):
            layers.append(Bottleneck(cin, mid, stride if i == 0 else 1))
            cin = 4 * mid
    layers += [nn.AdaptiveAvgPool2d(1), nn.Flatten(), nn.Linear(cin, 1000)]
    return nn.Sequential(*layers), torch.randn(256, 3, 224, 224, device=dev), torch.randint(0, 1000, (256,), device=dev)


for cl in (False, True):
    torch.manual_seed(0)
    net, x, y = build()
    net = net.to(dev).train()
    if cl:
        net = net.to(memory_format=torch.channels_last)
        x = x.contiguous(memory_format=torch.channels_last)
    opt = torch.optim.SGD(net.parameters(), 0.1, momentum=0.9)

    def step():
        loss = F.cross_entropy(net(x), y)
        opt.zero_grad()
        loss.backward()
        opt.step()

    for _ in range(6):
        step()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(10):
        step()
    t1.record()
    torch.cuda.synchronize()
    print("%s plain torch, channels_last=%s: %.2f ms/step" % (which, cl, t0.elapsed_time(t1) / 10), flush=True)
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=100))
    del net, opt
    torch.cuda.empty_cache()
