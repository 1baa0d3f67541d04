#!/usr/bin/env python
"""bench.py -- CNSN hot-path benchmark (driver contract).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Headline workload (config.workload): SelfNorm forward + backward, train mode, on a
(256,256,56,56) fp32 NCHW tensor -- the shape BASELINE.json's north_star quotes the metric on.
One "step" = one forward + one backward pass of the hot path over one synthetic batch.

  value      algorithmic GB/s = N_gpus * 5*S / step time, inputs resident in HBM (S = bytes of x;
             forward reads x writes y = 2S, backward reads x, dy writes dx = 3S)
  roofline   the dominant call (backward, 3*S) timed with CUDA events inside the timed region,
             against MEASURED_PEAKS.json's HBM copy bandwidth
  e2e        the same metric through the nn.Module API with HOST (pinned) buffers: per step
             H2D of x and dy, forward, backward, D2H of y and dx inside the timed region
             (double-buffered over three streams, as a data loader would feed it)
  sustained  the same step back to back for >= 1 s with clock sampling (the timed region proper is ~16 ms)
  cpu_baseline / --impl reference
             the reference's OWN models/cnsn.py (oracle/_ref, copied at build time by oracle/build_ref.py; the
             bit-identical eager port oracle/eager_modules.py when that copy is absent) on this box's host
             cores, SAME shape; --impl reference honours --steps / --warmup (same config, same steps)
  crossnorm  secondary: CrossNorm fwd+bwd through cn_op_2ins_space_chan (BASELINE config 2, a WideResNet
             site with crops, a north-star-sized tensor)
  site       secondary: a CNSN site with both operators firing, fused site kernels vs the two-operator sequence
  train      secondary: training-step images/s, DDP over NCCL when N>1: WideResNet-40-2 + CNSN (config 3),
             train.resnet50 = ResNet-50 + SN batch 256/GPU (config 4), train.resnet50_jsd = the 3-view JSD
             step in bf16 (config 5)

Multi-GPU: the CNSN path has no cross-GPU exchange -- every rank runs the same
per-GPU workload on its own shard ("weak" scaling, no data-path collective); only the secondary
training step all-reduces gradients.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

NORTH_STAR = (256, 256, 56, 56)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--shape", default=None, help="N,C,H,W override (parity/debug only)")
    ap.add_argument("--dtype", default="f32", choices=["f32", "bf16"])
    ap.add_argument("--no-train", action="store_true", help="skip the secondary WRN-40-2 measurement")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-sustained", action="store_true", help="skip the >= 1 s back-to-back run")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-crossnorm", action="store_true", help="skip the secondary CrossNorm measurements")
    ap.add_argument("--train-batch", type=int, default=512)
    ap.add_argument("--train-steps", type=int, default=20)
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples SM clock and clock-event reasons through NVML while the timed region runs."""
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index, period=0.005):
        self.samples, self.reasons, self.period, self.ok = [], set(), period, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:      # pragma: no cover
            self.err = repr(e)
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                bits = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for b, name in self.REASONS.items():
                    if bits & b and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def __enter__(self):
        if self.ok:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t:
            self._t.join()

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max, "reasons": sorted(self.reasons),
                "samples": len(s)}


def physical_gpu_index(local):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local])
        except Exception:
            return local
    return local


def bind_to_gpu_numa(index):
    """Opt-in (CNSN_BENCH_NUMA=1): pin this rank to the CPUs NVML reports as local to its GPU BEFORE the pinned
    staging buffers of the e2e leg are allocated (first touch then places them on the GPU's NUMA node).  Off by
    default: unmeasured so far."""
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(index))
        return sorted(os.sched_getaffinity(0))
    except Exception as e:      # pragma: no cover
        return "unavailable: %r" % (e,)


def workload_config(shape, dtype, world):
    """The `config` object of the JSON line -- identical in both arms (the reference arm runs the SAME workload)."""
    N, C, H, W = shape
    S = N * C * H * W * (4 if dtype == "f32" else 2)
    return {"workload": "SelfNorm fwd+bwd train-mode, NCHW %s (%d,%d,%d,%d) per GPU, S=%d bytes; "
                        "independent per-GPU batches, no data-path collective" % (dtype, N, C, H, W, S),
            "l2": "inputs (x, dy: %.0f MB each) larger than the 126 MB L2; no flush needed" % (S / 1e6),
            "parallelism": "replicated shards x%d" % world}


METRIC = "CNSN fwd+bwd GB/s (SelfNorm, algorithmic 5*S bytes per step)"


# ----------------------------------------------------------------------------- CPU reference arm
def reference_ops():
    """(module, kind, what): the reference's OWN models/cnsn.py from oracle/_ref (copied there at build time by
    oracle/build_ref.py; travels to the GPU box) -> kind "reference"; else the bit-identical eager-PyTorch port."""
    from oracle import build_ref
    mod = build_ref.load("models.cnsn")
    if mod is not None:
        return mod, "reference", "the reference's own models/cnsn.py (oracle/_ref, unmodified)"
    from oracle import eager_modules
    return eager_modules, "port", "eager-PyTorch port of models/cnsn.py (oracle/eager_modules.py; oracle/_ref absent)"


def cpu_reference_selfnorm(shape, steps, warmup):
    """The reference SelfNorm module (models/cnsn.py:113-150), forward + autograd backward, on this box's host cores
    with every thread torch can use, on the SAME shape as the GPU arm.  Returns the cpu_baseline object + timings."""
    import torch
    ops, kind, what = reference_ops()
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    N, C, H, W = shape
    gen = torch.Generator().manual_seed(1234)
    x = (torch.randn(shape, generator=gen) * (0.5 + 1.5 * torch.rand(N, C, 1, 1, generator=gen))
         + torch.randn(N, C, 1, 1, generator=gen)).requires_grad_(True)
    dy = torch.randn(shape, generator=gen)
    sn = ops.SelfNorm(C).train()
    times = []
    for i in range(warmup + steps):
        x.grad = None
        sn.zero_grad(set_to_none=True)
        t0 = time.perf_counter()
        y = sn(x)
        y.backward(dy)
        t1 = time.perf_counter()
        if i >= warmup:
            times.append(t1 - t0)
        del y
    S = N * C * H * W * 4
    t = sum(times) / len(times)
    return {"value": 5 * S / t / 1e9, "unit": "GB/s", "cores": cores, "kind": kind,
            "sample": "SelfNorm fwd+bwd fp32 on the full (%d,%d,%d,%d) tensor, %d steps after %d warm-up, mean wall clock; %s, "
                      "torch.set_num_threads(%d)" % (N, C, H, W, steps, warmup, what, cores),
            "ms_per_step": t * 1e3, "best_ms": min(times) * 1e3}


def run_reference_arm(args, shape):
    """`--impl reference`: the reference's CPU implementation of the path on the box's host cores, SAME config, metric,
    unit, --steps and --warmup as the GPU arm.  Rank 0 alone runs; the other ranks exit at once."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_reference_selfnorm(shape, args.steps, args.warmup)
    train = None
    if not args.no_train:
        try:                                     # secondary: the same training steps with the reference CNSN on CPU
            import torch
            from cnsn_b200.train import bench_resnet50_cpu, bench_wrn
            ops, kind, what = reference_ops()
            train = bench_wrn(torch.device("cpu"), 1, 0, batch=64, steps=2, warmup=1, ops=ops)
            train["sample"] = "batch 64 (of 512), 2 steps after 1 warm-up, CPU, CNSN operators: %s" % what
            train["resnet50"] = bench_resnet50_cpu(ops, batch=16, steps=1, warmup=1)
        except Exception as e:
            train = {"error": repr(e)[:300]}
    line = {
        "impl": "reference", "metric": METRIC,
        "value": cb["value"], "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(shape, "f32", max(1, args.gpus)),
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "train_summary": train_summary(train),
        "train": train,
    }
    print(json.dumps(line), flush=True)


def train_summary(train):
    """The training-step numbers (BASELINE.json `metric`: WRN-40-2 images/s at 1/2/4/8 GPUs) in compact form, printed
    right after `value` so that a truncated record still carries them."""
    if not isinstance(train, dict) or "value" not in train:
        return None
    out = {"wrn_img_s": round(train["value"], 1), "wrn_ms": round(train["ms_per_step"], 3), "n_gpus": train.get("n_gpus")}
    for key, short in (("resnet50", "r50"), ("resnet50_jsd", "jsd")):
        t = train.get(key)
        if isinstance(t, dict) and "value" in t:
            out[short + "_img_s"] = round(t["value"], 1)
            out[short + "_ms"] = round(t["ms_per_step"], 2)
    if "graph" in train:
        out["wrn_graph"] = train["graph"]
    if "memory_format" in train:
        out["memory_format"] = train["memory_format"]
    return out


def bench_crossnorm(torch, M, dev, steps=30):
    """CrossNorm forward + backward through cn_op_2ins_space_chan (host RNG, perm upload, autograd included):
    BASELINE config 2 (128,64,32,32) bf16 crop='neither' -- a 16 MiB tensor, launch-latency regime, reported in
    microseconds -- and (256,256,56,56) fp32, reported as algorithmic GB/s (5*S per step)."""
    import numpy as np
    out = {}
    for name, shape, dt, crop in (("cfg2_128x64x32x32_bf16_neither", (128, 64, 32, 32), torch.bfloat16, "neither"),
                                  ("wrn_512x32x32x32_f32_both", (512, 32, 32, 32), torch.float32, "both"),
                                  ("large_256x256x56x56_f32_neither", (256, 256, 56, 56), torch.float32, "neither")):
        x = torch.randn(shape, device=dev).to(dt).requires_grad_(True)
        dy = torch.randn(shape, device=dev).to(dt)
        S = x.numel() * x.element_size()
        torch.manual_seed(0)
        np.random.seed(0)
        for _ in range(5):
            torch.autograd.grad(M.cn_op_2ins_space_chan(x, crop=crop, beta=1), x, dy)
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
        torch.cuda.synchronize()
        for e in ev:
            e[0].record()
            y = M.cn_op_2ins_space_chan(x, crop=crop, beta=1)
            e[1].record()
            torch.autograd.grad(y, x, dy)
            e[2].record()
        torch.cuda.synchronize()
        f = sorted(e[0].elapsed_time(e[1]) for e in ev)[steps // 2]
        b = sorted(e[1].elapsed_time(e[2]) for e in ev)[steps // 2]
        out[name] = {"fwd_us": f * 1e3, "bwd_us": b * 1e3, "bytes_5S": 5 * S, "gbs": 5 * S / ((f + b) * 1e-3) / 1e9}
        if S < (64 << 20):
            # a lone autograd.grad pays the engine's start-up and thread hand-off (~40 us) on top of the node; inside a
            # training graph the per-node cost is what counts: 8 calls chained, one backward, per node
            chain = 8
            ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
            for e in ev:
                e[0].record()
                v = x
                for _ in range(chain):
                    v = M.cn_op_2ins_space_chan(v, crop=crop, beta=1)
                e[1].record()
                torch.autograd.grad(v, x, dy)
                e[2].record()
            torch.cuda.synchronize()
            out[name]["fwd_us_in_graph"] = sorted(e[0].elapsed_time(e[1]) for e in ev)[steps // 2] * 1e3 / chain
            out[name]["bwd_us_in_graph"] = sorted(e[1].elapsed_time(e[2]) for e in ev)[steps // 2] * 1e3 / chain
        del x, dy
    return out


def bench_site(torch, M, dev, steps=20):
    """A CNSN site whose CrossNorm AND SelfNorm fire (models/cnsn.py:159-164) through CNSN.forward: the fused site
    kernels (one launch per direction, 5*S algorithmic bytes) against this package's two-operator sequence, on a
    WideResNet-40-2 site of BASELINE config 3 and on a north-star-sized tensor."""
    import numpy as np
    out = {}
    for name, shape, dt, crop in (("wrn_512x32x32x32_f32_both", (512, 32, 32, 32), torch.float32, "both"),
                                  ("large_256x256x56x56_f32_neither", (256, 256, 56, 56), torch.float32, "neither")):
        N, C = shape[:2]
        x = (torch.randn(shape, device=dev) * (0.5 + torch.rand(N, C, 1, 1, device=dev))).to(dt).requires_grad_(True)
        dy = torch.randn(shape, device=dev).to(dt)
        S = x.numel() * x.element_size()
        blk = M.CNSN(M.CrossNorm(crop=crop, beta=1), M.SelfNorm(C)).to(dev).train()
        torch.manual_seed(0)
        np.random.seed(0)
        res = {}
        for fused in (True, False):
            M.CNSN.fuse_site = fused
            try:
                for _ in range(3):
                    blk.crossnorm.active = True
                    torch.autograd.grad(blk(x), x, dy)
                ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
                torch.cuda.synchronize()
                for e in ev:
                    blk.crossnorm.active = True
                    e[0].record()
                    y = blk(x)
                    e[1].record()
                    torch.autograd.grad(y, x, dy)
                    e[2].record()
                torch.cuda.synchronize()
            finally:
                M.CNSN.fuse_site = True
            f = sorted(e[0].elapsed_time(e[1]) for e in ev)[steps // 2]
            b = sorted(e[1].elapsed_time(e[2]) for e in ev)[steps // 2]
            res["fused" if fused else "sequence"] = {"fwd_us": f * 1e3, "bwd_us": b * 1e3}
        t1 = res["fused"]["fwd_us"] + res["fused"]["bwd_us"]
        t0 = res["sequence"]["fwd_us"] + res["sequence"]["bwd_us"]
        out[name] = dict(res, bytes_5S=5 * S, gbs=5 * S / (t1 * 1e-6) / 1e9, speedup_vs_sequence=t0 / t1)
        del x, dy, blk
    return out


# ----------------------------------------------------------------------------- GPU arm
def main():
    args = parse()
    shape = tuple(int(v) for v in args.shape.split(",")) if args.shape else NORTH_STAR
    if args.impl == "reference":
        return run_reference_arm(args, shape)

    import torch
    import torch.distributed as dist
    import cnsn_b200
    import cnsn_b200.cnsn as M

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py (b200 arm) needs a CUDA device"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa(physical_gpu_index(local)) if os.environ.get("CNSN_BENCH_NUMA") == "1" else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.gpus != world and rank == 0 and world > 1:
        print("warning: --gpus %d but WORLD_SIZE %d" % (args.gpus, world), file=sys.stderr)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    N, C, H, W = shape
    tdt = torch.float32 if args.dtype == "f32" else torch.bfloat16
    esz = 4 if args.dtype == "f32" else 2
    S = N * C * H * W * esz
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)       # each rank its own shard
    x = (torch.randn(shape, device=dev, generator=gen) * (0.5 + 1.5 * torch.rand(N, C, 1, 1, device=dev, generator=gen))
         + torch.randn(N, C, 1, 1, device=dev, generator=gen)).to(tdt).requires_grad_(True)
    dy = torch.randn(shape, device=dev, generator=gen).to(tdt)
    sn = M.SelfNorm(C).to(dev).train()

    def step(ev=None):
        if ev:
            ev[0].record()
        y = sn(x)
        if ev:
            ev[1].record()
        (dx,) = torch.autograd.grad(y, x, dy)     # parameter grads are produced by the same fused backward
        if ev:
            ev[2].record()
        return y, dx

    for _ in range(args.warmup):
        step()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    n0 = cnsn_b200.launch_count()
    with ClockSampler(physical_gpu_index(local)) as clk:
        t0.record()
        for i in range(args.steps):
            step(evs[i])
        t1.record()
        barrier()
    launches = cnsn_b200.launch_count() - n0
    ms_total = max_over_ranks(t0.elapsed_time(t1))
    ms_step = ms_total / args.steps
    fwd_ms = sum(e[0].elapsed_time(e[1]) for e in evs) / args.steps
    bwd_ms = sum(e[1].elapsed_time(e[2]) for e in evs) / args.steps
    value = world * 5 * S / (ms_step * 1e-3) / 1e9
    peak, peak_src = peaks()
    bwd_gbs = 3 * S / (bwd_ms * 1e-3) / 1e9
    fwd_gbs = 2 * S / (fwd_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "dram_traffic.json")      # from the committed ncu --set full capture
    if os.path.isfile(tpath):
        try:
            with open(tpath) as f:
                traffic = json.load(f).get("selfnorm_bwd_%s_bytes_per_launch" % args.dtype)
        except Exception:
            traffic = None

    # ---- the same step back to back for >= 1 s: what the clocks do under sustained load (the 20-step timed region above
    # lasts ~16 ms -- two or three clock samples)
    sustained = None
    if not args.no_sustained:
        n_s = max(args.steps, int(1.2 / (ms_step * 1e-3)))
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        with ClockSampler(physical_gpu_index(local), period=0.02) as clk_s:
            s0.record()
            for _ in range(n_s):
                step()
            s1.record()
            torch.cuda.synchronize()
        s_ms = s0.elapsed_time(s1) / n_s
        sustained = {"value": 5 * S / (s_ms * 1e-3) / 1e9, "unit": "GB/s per GPU", "ms_per_step": s_ms, "steps": n_s,
                     "seconds": s_ms * n_s * 1e-3, "frac_of_peak": 5 * S / (s_ms * 1e-3) / 1e9 / peak, "clocks": clk_s.summary()}

    # ---- e2e: module API with HOST pinned buffers, copies inside the timed region
    e2e = None
    if not args.no_e2e:
        e2e_steps = max(2, min(args.steps, 20))          # every step is timed with its copies; more steps amortise the pipeline fill
        hx = torch.empty(shape, dtype=tdt, pin_memory=True).copy_(x.detach())
        hdy = torch.empty(shape, dtype=tdt, pin_memory=True).copy_(dy)
        hy = torch.empty(shape, dtype=tdt, pin_memory=True)
        hdx = torch.empty(shape, dtype=tdt, pin_memory=True)

        # Double-buffered pipeline, as a data loader would feed the module: H2D of step i+1, compute of step i and
        # D2H of step i-1 overlap on three streams (PCIe is full duplex); every step still copies ITS inputs from
        # pinned host memory and reads ITS y and dx back inside the timed region.
        comp = torch.cuda.current_stream()
        h2d, d2h = torch.cuda.Stream(), torch.cuda.Stream()
        bufs = [(torch.empty(shape, dtype=tdt, device=dev), torch.empty(shape, dtype=tdt, device=dev)) for _ in range(2)]
        free_ev = [None, None]

        def e2e_step(i):
            bx, bdy = bufs[i % 2]
            with torch.cuda.stream(h2d):
                if free_ev[i % 2] is not None:
                    h2d.wait_event(free_ev[i % 2])            # the buffer's previous step has been consumed
                bx.copy_(hx, non_blocking=True)
                bdy.copy_(hdy, non_blocking=True)
                ev_in = torch.cuda.Event()
                ev_in.record(h2d)
            comp.wait_event(ev_in)
            xin = bx.detach().requires_grad_(True)
            y = sn(xin)
            ev_y = torch.cuda.Event()
            ev_y.record(comp)
            (gx,) = torch.autograd.grad(y, xin, bdy)
            ev_dx = torch.cuda.Event()
            ev_dx.record(comp)
            free_ev[i % 2] = ev_dx
            with torch.cuda.stream(d2h):
                d2h.wait_event(ev_y)
                hy.copy_(y.detach(), non_blocking=True)
                y.record_stream(d2h)
                d2h.wait_event(ev_dx)
                hdx.copy_(gx, non_blocking=True)
                gx.record_stream(d2h)

        e2e_step(0)
        torch.cuda.synchronize()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(e2e_steps):
            e2e_step(i + 1)
        comp.wait_stream(d2h)
        comp.wait_stream(h2d)
        b.record()
        barrier()
        e_ms = max_over_ranks(a.elapsed_time(b)) / e2e_steps
        # the e2e leg's own roofline: the SAME pinned copies (2*S host->device and 2*S device->host per step, both
        # directions at once) with no compute in between -- what this box's host side delivers to N ranks at a time
        torch.cuda.synchronize()
        barrier()
        a2, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a2.record()
        for i in range(e2e_steps):
            bx, bdy = bufs[i % 2]
            with torch.cuda.stream(h2d):
                bx.copy_(hx, non_blocking=True)
                bdy.copy_(hdy, non_blocking=True)
            with torch.cuda.stream(d2h):
                hy.copy_(bx, non_blocking=True)
                hdx.copy_(bdy, non_blocking=True)
        comp.wait_stream(d2h)
        comp.wait_stream(h2d)
        b2.record()
        barrier()
        c_ms = max_over_ranks(a2.elapsed_time(b2)) / e2e_steps
        e2e = {"value": world * 5 * S / (e_ms * 1e-3) / 1e9, "unit": "GB/s", "h2d_bytes_per_step": 2 * S,
               "d2h_bytes_per_step": 2 * S, "ms_per_step": e_ms, "steps": e2e_steps,
               "api": "cnsn_b200.cnsn.SelfNorm forward + autograd backward on pinned host tensors; "
                      "double-buffered: H2D / compute / D2H of consecutive steps overlap on three streams",
               "copy_only": {"ms_per_step": c_ms, "gbs_per_direction_per_gpu": 2 * S / (c_ms * 1e-3) / 1e9,
                             "what": "the same pinned copies without the compute, all ranks at once: the host-side ceiling of this leg"},
               "frac_of_copy_only": c_ms / e_ms}
        if numa is not None:
            e2e["numa_cpus"] = numa if isinstance(numa, str) else "%d cpus bound (NVML affinity of the GPU)" % len(numa)
        del bufs
        del hx, hdy, hy, hdx

    del x, dy
    torch.cuda.empty_cache()

    # ---- secondary: CrossNorm (BASELINE config 2 and a north-star-sized tensor), module API, CUDA events
    crossnorm = None
    if not args.no_crossnorm:
        try:
            crossnorm = bench_crossnorm(torch, M, dev)
        except Exception as e:
            crossnorm = {"error": repr(e)[:300]}
        torch.cuda.empty_cache()

    site = None
    if not args.no_crossnorm:
        try:
            site = bench_site(torch, M, dev)
        except Exception as e:
            site = {"error": repr(e)[:300]}
        torch.cuda.empty_cache()

    # ---- secondary: WRN-40-2 + CNSN training step
    train = None
    if not args.no_train:
        try:
            from cnsn_b200.train import bench_resnet50, bench_wrn
            train = bench_wrn(dev, world, rank, batch=args.train_batch, steps=args.train_steps, warmup=5, fuse_post=True)
            torch.cuda.empty_cache()
            # BASELINE config 4: ResNet-50 + SelfNorm ('post'), image-space CrossNorm, 224x224, batch 256 per GPU
            train["resnet50"] = bench_resnet50(dev, world, rank, batch=256, steps=8, warmup=3, fuse_post=True)
            torch.cuda.empty_cache()
            # BASELINE config 5: the same network on 3 x 256 views per GPU with the JSD consistency step, bf16
            from cnsn_b200.train import bench_resnet50_jsd
            train["resnet50_jsd"] = bench_resnet50_jsd(dev, world, rank, batch=256, steps=4, warmup=2, fuse_post=True)
        except Exception as e:      # the headline must still be printed
            train = dict(train or {}, error=repr(e)[:300])

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_reference_selfnorm(shape, 4, 1)           # bounded: 5 passes of the full tensor, a few seconds
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {
            "metric": METRIC,
            "value": value, "unit": "GB/s", "train_summary": train_summary(train),
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "config": workload_config(shape, args.dtype, world),
            "roofline": {"bound": "hbm", "achieved": bwd_gbs, "peak": peak, "unit": "GB/s", "frac": bwd_gbs / peak,
                         "traffic": traffic, "kernel": "cnsn_selfnorm_bwd (3*S algorithmic bytes per call)",
                         "peak_source": peak_src,
                         "fwd": {"achieved": fwd_gbs, "frac": fwd_gbs / peak, "ms": fwd_ms, "bytes": 2 * S},
                         "bwd": {"achieved": bwd_gbs, "frac": bwd_gbs / peak, "ms": bwd_ms, "bytes": 3 * S},
                         "step": {"achieved": value / world, "frac": value / world / peak}},
            "clocks": clk.summary(),
            "e2e": e2e,
            "gpu_launches": launches,
            "cpu_baseline": cpu,
            "sustained": sustained,
            "crossnorm": crossnorm,
            "site": site,
            "train": train,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
