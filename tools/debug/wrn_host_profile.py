"""Host-side cost of the eager WideResNet-40-2 + CNSN step (the steps whose coin fires CrossNorm cannot be replayed from the
CUDA graph): step time WITH the per-step float(loss) read of cifar.py:134, NCHW against channels_last, and a cProfile of the
channels_last one."""
import cProfile
import os
import pstats
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402
from cnsn_b200.train import make_optimizer, wrn40_2  # noqa: E402

dev = torch.device("cuda", 0)
torch.backends.cudnn.benchmark = True
for cl in (False, True):
    torch.manual_seed(0)
    np.random.seed(0)
    net = wrn40_2(fuse_post=True).to(dev).train()
    x = torch.randn(512, 3, 32, 32, device=dev)
    if cl:
        net = net.to(memory_format=torch.channels_last)
        x = x.contiguous(memory_format=torch.channels_last)
    opt, sched = make_optimizer(net, 100)
    y = torch.randint(0, 10, (512,), device=dev)

    def step(aug):
        loss = F.cross_entropy(net(x, aug=aug), y)
        opt.zero_grad()
        loss.backward()
        opt.step()
        return float(loss.detach())

    for aug in (False, True):
        for _ in range(5):
            step(aug)
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(10):
            step(aug)
        t1.record()
        torch.cuda.synchronize()
        print("channels_last=%s aug=%s: %.2f ms/step eager with the per-step loss read" % (cl, aug, t0.elapsed_time(t1) / 10), flush=True)
    if cl:
        pr = cProfile.Profile()
        pr.enable()
        for _ in range(5):
            step(True)
        pr.disable()
        pstats.Stats(pr).sort_stats("tottime").print_stats(28)
