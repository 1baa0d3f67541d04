"""Context measurement (tool only): the same WRN-40-2 training step on the GPU with the reference's eager
PyTorch CNSN op chain (oracle/eager_modules.py) instead of the CUDA kernels."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from cnsn_b200.train import bench_wrn  # noqa: E402
from oracle import eager_modules  # noqa: E402

dev = torch.device("cuda", 0)
a = bench_wrn(dev, 1, 0, batch=512, steps=20, warmup=5)
b = bench_wrn(dev, 1, 0, batch=512, steps=20, warmup=5, ops=eager_modules)
print("cnsn_b200 kernels : %.0f images/s (%.2f ms/step)" % (a["value"], a["ms_per_step"]))
print("eager CNSN on GPU : %.0f images/s (%.2f ms/step)  -> step speed-up %.2fx" % (b["value"], b["ms_per_step"], b["ms_per_step"] / a["ms_per_step"]))
