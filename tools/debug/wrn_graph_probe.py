"""GraphedStep on WideResNet-40-2 + CNSN: time of the replayed steps and of the eager (CrossNorm) steps separately,
NCHW against channels_last, each step with its float(loss) read."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from cnsn_b200.train import GraphedStep, make_optimizer, wrn40_2  # noqa: E402

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
    opt, sched = make_optimizer(net, 200)
    y = torch.randint(0, 10, (512,), device=dev)
    gs = GraphedStep(net, x, y, 1)
    print("channels_last=%s graph=%s %s" % (cl, gs.graph is not None, gs.capture_error or ""), flush=True)
    for prob in (0.0, 1.0, 0.0, 1.0):
        for _ in range(4):
            gs.step(gs.x, gs.y, opt, sched, prob)
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(10):
            gs.step(gs.x, gs.y, opt, sched, prob)
        t1.record()
        torch.cuda.synchronize()
        print("  cn_prob=%.0f: %.2f ms/step" % (prob, t0.elapsed_time(t1) / 10), flush=True)
    # alternate: replay, eager, replay, eager ...
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(20):
        gs.step(gs.x, gs.y, opt, sched, float(i % 2))
    t1.record()
    torch.cuda.synchronize()
    print("  alternating: %.2f ms/step" % (t0.elapsed_time(t1) / 20), flush=True)
    del gs, net, opt
    torch.cuda.empty_cache()
