#!/bin/bash
# round 2, step 4j: ncu launch list (gpu__time_duration) of WideResNet-40-2 + CNSN training steps, channels_last, eager (no graph)
mkdir -p gpurun_out
cat > /tmp/wrn_steps.py <<'PY'
import sys; sys.path.insert(0, '.')
import numpy as np, torch, torch.nn.functional as F
from cnsn_b200.train import make_optimizer, wrn40_2
dev = torch.device("cuda", 0)
torch.backends.cudnn.benchmark = True
torch.manual_seed(0); np.random.seed(0)
net = wrn40_2(fuse_post=True).to(dev).train().to(memory_format=torch.channels_last)
opt, sched = make_optimizer(net, 100)
x = torch.randn(512, 3, 32, 32, device=dev).contiguous(memory_format=torch.channels_last)
y = torch.randint(0, 10, (512,), device=dev)
for i in range(6):
    loss = F.cross_entropy(net(x, aug=False), y)
    opt.zero_grad(); loss.backward(); opt.step()
torch.cuda.synchronize()
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r4j_wrn_launches_all.csv python /tmp/wrn_steps.py > gpurun_out/r4j_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r4j_wrn_launches_all.csv")) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
ix = {h: i for i, h in enumerate(rows[hdr])}
data = rows[hdr + 1:]
# the last step only: kernels are identical per step, take the final sixth of the launches
n = len(data) // 6
last = data[-n:]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in last:
    k = r[ix["Kernel Name"]][:110]
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    us = v / 1e3 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1e3
    agg[k][0] += 1; agg[k][1] += us
tot = sum(v[1] for v in agg.values())
with open("gpurun_out/r4j_wrn_step_launches.txt", "w") as f:
    f.write("WideResNet-40-2 + CNSN, batch 512, channels_last, one eager step without CrossNorm: %d launches, %.2f ms of kernel time under ncu (cold cache, serialised)\n" % (len(last), tot / 1e3))
    for k, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        f.write("%6.2f %%  %8.1f us  %4d x  %s\n" % (100 * us / tot, us, c, k))
print(open("gpurun_out/r4j_wrn_step_launches.txt").read()[:2500])
PY
rm -f gpurun_out/r4j_wrn_launches_all.csv
