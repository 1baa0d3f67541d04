"""Training-step context measurements on one GPU (tool only; may import oracle/ for the eager comparison):

    python tools/train_bench.py wrn|resnet50 [batch] [steps]
Prints images/s with the cnsn_b200 kernels (unfused and with the fused pos='post' block tail) and with the
reference's eager-PyTorch CNSN op chain (oracle/eager_modules.py) on the same GPU.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from cnsn_b200.train import bench_resnet50, bench_wrn  # noqa: E402
from oracle import eager_modules  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "wrn"
dev = torch.device("cuda", 0)
if which == "wrn":
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    runs = [("cnsn_b200 kernels", dict()), ("cnsn_b200 kernels, fused post tail", dict(fuse_post=True)),
            ("eager CNSN on the GPU", dict(ops=eager_modules))]
    fn = lambda **kw: bench_wrn(dev, 1, 0, batch=batch, steps=steps, warmup=5, **kw)  # noqa: E731
else:
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    runs = [("cnsn_b200 kernels", dict(fuse_post=False)), ("cnsn_b200 kernels, fused post tail", dict(fuse_post=True)),
            ("eager CNSN on the GPU", dict(fuse_post=False, ops=eager_modules))]
    fn = lambda **kw: bench_resnet50(dev, 1, 0, batch=batch, steps=steps, warmup=3, **kw)  # noqa: E731
base = None
for name, kw in runs:
    r = fn(**kw)
    torch.cuda.empty_cache()
    base = base or r["ms_per_step"]
    print("%-8s batch %d | %-36s %9.0f images/s  %8.2f ms/step  (x%.3f vs first row)  peak mem %.1f GB" % (
        which, batch, name, r["value"], r["ms_per_step"], base / r["ms_per_step"], torch.cuda.max_memory_allocated() / 1e9), flush=True)
    torch.cuda.reset_peak_memory_stats()
