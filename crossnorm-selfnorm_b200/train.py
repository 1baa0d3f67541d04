"""Synthetic-data training harness for the CNSN host models (stands in for the reference's cifar.py).

One process per GPU; gradients are all-reduced by DistributedDataParallel over NCCL.  The CNSN path
itself never crosses GPUs: CrossNorm permutes inside the per-GPU batch and SelfNorm's BatchNorm1d uses
per-GPU batch statistics -- the per-replica behaviour of the reference's DataParallel.

The step mirrors ``train_cn`` of the reference's cifar.py:117-145: a coin ``np.random.rand(1) < cn_prob``
decides whether this step's forward activates CrossNorm sites (``net(x, aug=True)``), then cross-entropy,
``zero_grad``, ``backward``, SGD-Nesterov step, per-step cosine LR, and a ``float(loss)`` read (host sync)
every step, as there.
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


def wrn40_2(ops=None, num_classes=10, active_num=2, pos="post", beta=1, crop="both", cnsn_type="cnsn", fuse_post=False):
    """WideResNet-40-2 + CNSN with the hyper-parameters of cifar10-scripts/wideresnet/run-cnsn.sh."""
    from .hosts.wideresnet import WideResNet
    return WideResNet(40, num_classes, widen_factor=2, drop_rate=0.0, active_num=active_num, pos=pos, beta=beta,
                      crop=crop, cnsn_type=cnsn_type, ops=ops, fuse_post=fuse_post)


def cosine_lr(step, total_steps, lr_max, lr_min):
    return lr_min + (lr_max - lr_min) * 0.5 * (1 + math.cos(step / total_steps * math.pi))


def make_optimizer(net, total_steps, lr=0.1, momentum=0.9, wd=5e-4):
    opt = torch.optim.SGD(net.parameters(), lr, momentum=momentum, weight_decay=wd, nesterov=True)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lr_lambda=lambda s: cosine_lr(s, total_steps, 1, 1e-6 / lr))
    return opt, sched


def train_step(net, images, targets, opt, sched, cn_prob):
    """One ``train_cn`` step; returns the loss as a Python float (host sync, as in the reference)."""
    aug = bool(np.random.rand(1) < cn_prob)
    logits = net(images, aug=aug)
    loss = F.cross_entropy(logits, targets)
    opt.zero_grad()
    loss.backward()
    opt.step()
    sched.step()
    return float(loss.detach())


class GraphedStep:
    """The ``train_cn`` step (cifar.py:117-145) with the launch-bound part taken off the host: forward, loss and
    backward of the step WITHOUT CrossNorm -- 1 - cn_prob = 75 % of the steps of BASELINE config 3 -- are captured once
    into a CUDA graph and replayed; steps whose coin activates CrossNorm sites run eagerly (CrossNorm draws a fresh
    permutation and crop boxes on the host every call, models/cnsn.py:62-77, and which sites fire changes per step,
    wideresnet_cnsn.py:199-203).  Same arithmetic either way: the graph replays the very kernels the eager step
    launches.  Gradients live in ONE flat buffer (every ``p.grad`` is a view of it), so the data-parallel exchange is
    a single NCCL all-reduce (sum, then / world) of that buffer after backward -- what DistributedDataParallel's bucket does,
    without its per-step host work; running statistics stay per replica (the reference's DataParallel behaviour).
    The per-step ``float(loss)`` host read of the reference stays."""

    def __init__(self, net, images, targets, world=1, capture=True, loss_fn=None):
        """loss_fn(net, images, targets, aug) -> scalar loss; default: cross-entropy of ``net(images, aug=aug)``
        (cifar.py:132-133).  Everything loss_fn launches is captured, so it must not read the host RNG or sync."""
        self.net, self.world = net, world
        self.loss_fn = loss_fn or (lambda n, x, y, aug: F.cross_entropy(n(x, aug=aug), y))
        self.params = [p for p in net.parameters() if p.requires_grad]
        self.flat = torch.zeros(sum(p.numel() for p in self.params), dtype=self.params[0].dtype, device=images.device)
        off = 0
        for p in self.params:
            chunk = self.flat[off:off + p.numel()]
            if p.dim() == 4 and not p.is_contiguous() and p.is_contiguous(memory_format=torch.channels_last):
                n, c, h, w = p.shape                      # same strides as the parameter (the gradient layout contract)
                p.grad = chunk.view(n, h, w, c).permute(0, 3, 1, 2)
            else:
                p.grad = chunk.view_as(p)
            off += p.numel()
        self.x, self.y = images.clone(), targets.clone()
        self.graph, self.capture_error = None, None
        if capture and images.is_cuda:
            try:                                          # warm-up on a side stream is intended: no warning per parameter
                torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
            except AttributeError:
                pass
            try:
                self._capture()
            except Exception as e:      # noqa: BLE001 -- a step that cannot be captured still trains, eagerly
                self.graph, self.capture_error = None, repr(e)[:300]
                torch.cuda.synchronize()

    def _forward_backward(self, aug):
        self.flat.zero_()
        loss = self.loss_fn(self.net, self.x, self.y, aug)
        loss.backward()                                   # accumulates into the views of the (zeroed) flat buffer
        return loss

    def _capture(self):
        # warm-up on a side stream (cuDNN plans, caching-allocator blocks, the library's per-kernel preparation), with
        # the buffers the forward updates put back afterwards: capture must not change the training state
        bufs = [b for b in self.net.buffers()]
        saved = [b.clone() for b in bufs]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                self._forward_backward(False)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._forward_backward(False)
        with torch.no_grad():
            for b, v in zip(bufs, saved):
                b.copy_(v)
        torch.cuda.synchronize()

    def _finish(self, loss, opt, sched):
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(self.flat)                    # sum, then the mean DistributedDataParallel would deliver
            self.flat.div_(self.world)
        opt.step()
        if sched is not None:
            sched.step()
        return float(loss.detach())

    def _load(self, images, targets):
        if images is not self.x:
            self.x.copy_(images, non_blocking=True)
        if targets is not self.y:
            self.y.copy_(targets, non_blocking=True)

    def step(self, images, targets, opt, sched, cn_prob):
        """``train_cn`` (cifar.py:117-145): the coin activates CrossNorm sites INSIDE the network -> eager step;
        otherwise the captured step is replayed."""
        aug = bool(np.random.rand(1) < cn_prob)           # cifar.py:127
        self._load(images, targets)
        if aug or self.graph is None:
            loss = self._forward_backward(aug)
        else:
            self.graph.replay()
            loss = self.loss
        return self._finish(loss, opt, sched)

    def step_image_cn(self, images, targets, opt, sched, cn_prob, ops, beta=1, crop="neither"):
        """``train_cn_image`` / ``train_cn_image_consist`` (imagenet.py:205-230, :348-385): the coin sends the BATCH through
        image-space CrossNorm before the network (eager: fresh host draws), the network itself never fires CrossNorm
        (``aug=False``) -> EVERY step replays the captured forward + loss + backward."""
        if np.random.rand(1) < cn_prob:                   # imagenet.py:213-215
            images = ops.cn_op_2ins_space_chan(images, beta=beta, crop=crop)
        self._load(images, targets)
        if self.graph is None:
            loss = self._forward_backward(False)
        else:
            self.graph.replay()
            loss = self.loss
        return self._finish(loss, opt, sched)


def bench_wrn(dev, world, rank, batch=512, steps=20, warmup=5, cn_prob=0.25, ops=None, fuse_post=False, graph=None,
              channels_last=None):
    """images/s of WideResNet-40-2 + CNSN training on synthetic CIFAR-shaped data (fp32, batch per GPU).
    channels_last (default: on CUDA with this package's operators): parameters and activations in torch.channels_last,
    the layout cuDNN's tensor-core convolutions work in -- SelfNorm sites run their NHWC kernels (csrc/selfnorm_nhwc.cu),
    a site whose CrossNorm fires converts to NCHW for that call.
    graph (default: on CUDA with this package's operators): GraphedStep -- CUDA graph for the steps without CrossNorm,
    one flat-buffer gradient all-reduce; otherwise the plain eager step (DistributedDataParallel when world > 1)."""
    import torch.distributed as dist
    if dev.type == "cuda":
        torch.backends.cudnn.benchmark = True  # cifar.py:396
    torch.manual_seed(1 + rank)               # parameters are broadcast below; permutations differ per rank
    # ONE numpy stream for all ranks: the per-step coin (cifar.py:127) is drawn once per step for the whole batch in
    # the reference (a single process drives every replica), so every rank takes the same aug / no-aug decision --
    # and no rank waits in the gradient exchange for another rank's slower CrossNorm step
    np.random.seed(1)
    net = wrn40_2(ops=ops, fuse_post=fuse_post).to(dev).train()
    if graph is None:
        graph = dev.type == "cuda" and ops is None
    if channels_last is None:
        channels_last = dev.type == "cuda" and ops is None
    if channels_last:
        net = net.to(memory_format=torch.channels_last)
    model = net
    if world > 1:
        with torch.no_grad():
            for t in list(net.parameters()) + list(net.buffers()):
                dist.broadcast(t, 0)
    if world > 1 and not graph:
        # broadcast_buffers=False: SelfNorm / BatchNorm running statistics stay per replica, as under the
        # reference's DataParallel; only gradients are exchanged
        model = nn.parallel.DistributedDataParallel(net, device_ids=[dev.index] if dev.type == "cuda" else None,
                                                    broadcast_buffers=False, gradient_as_bucket_view=True)
    opt, sched = make_optimizer(model, total_steps=steps + warmup)
    x = torch.randn(batch, 3, 32, 32, device=dev)
    if channels_last:
        x = x.contiguous(memory_format=torch.channels_last)
    y = torch.randint(0, 10, (batch,), device=dev)
    is_cuda = dev.type == "cuda"
    launches0 = None
    if is_cuda:
        from . import _lib
        launches0 = _lib.launch_count()
    gs = GraphedStep(net, x, y, world) if graph else None

    def one_step():
        if gs is not None:
            return gs.step(gs.x, gs.y, opt, sched, cn_prob)
        return train_step(model, x, y, opt, sched, cn_prob)

    for _ in range(warmup):
        one_step()
    if is_cuda:
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
    else:
        import time
        if world > 1:
            dist.barrier()
        w0 = time.perf_counter()
    loss = 0.0
    for _ in range(steps):
        loss = one_step()
    if is_cuda:
        t1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
    else:
        ms = (time.perf_counter() - w0) * 1e3
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
    out = {"metric": "WideResNet-40-2 + CNSN training images/s", "value": world * batch * steps / (ms * 1e-3),
           "unit": "images/s", "ms_per_step": ms / steps, "steps": steps, "warmup": warmup,
           "batch_per_gpu": batch, "n_gpus": world, "dtype": "f32 (TF32 convolutions: %s)" % torch.backends.cudnn.allow_tf32,
           "config": "depth 40, widen 2, cnsn_type=cnsn, pos=post, crop=both, beta=1, active_num=2, cn_prob=%g, "
                     "SGD nesterov lr 0.1 wd 5e-4, cosine LR, synthetic 32x32, fuse_post=%s" % (cn_prob, bool(fuse_post)),
           "final_loss": loss, "params": sum(p.numel() for p in net.parameters()),
           "graph": ("CUDA graph for the steps without CrossNorm (%d %% of them), eager otherwise; one flat-buffer gradient "
                     "all-reduce" % round(100 * (1 - cn_prob))) if gs is not None and gs.graph is not None else "eager"}
    if launches0 is not None:
        from . import _lib
        out["cnsn_kernel_launches"] = _lib.launch_count() - launches0
    out["param_checksum"] = float(sum(p.detach().double().sum() for p in net.parameters()))
    out["memory_format"] = "channels_last" if channels_last else "contiguous (NCHW)"
    return out


def resnet50_step(net, images, targets, opt, cn_prob, ops, beta=1, crop="neither"):
    """One ``train_cn_image`` step of the reference's imagenet.py:205-230: a coin decides whether this batch goes
    through image-space CrossNorm (``cn_op_2ins_space_chan(input, beta, crop)``) before the network; cross-entropy,
    ``zero_grad``, ``backward``, SGD step, and a ``loss.item()`` read (host sync) every step, as there."""
    if np.random.rand(1) < cn_prob:
        images = ops.cn_op_2ins_space_chan(images, beta=beta, crop=crop)
    loss = F.cross_entropy(net(images, aug=False), targets)
    opt.zero_grad()
    loss.backward()
    opt.step()
    return float(loss.detach())


def _broadcast_model(net, world):
    if world > 1:
        import torch.distributed as dist
        with torch.no_grad():
            for t in list(net.parameters()) + list(net.buffers()):
                dist.broadcast(t, 0)


def bench_resnet50(dev, world, rank, batch=256, steps=10, warmup=3, cn_prob=0.5, fuse_post=True, ops=None, graph=True,
                   channels_last=None):
    """images/s of ResNet-50 + SelfNorm ('post') training with image-space CrossNorm on synthetic 224x224 data
    (BASELINE config 4: batch 256 per GPU, SGD lr 0.1 momentum 0.9 wd 1e-4; imagenet-scripts/run-cnsn.sh).
    graph: GraphedStep.step_image_cn -- the network's forward + loss + backward replayed from a CUDA graph on every step
    (CrossNorm acts on the images, before the network), one flat-buffer gradient all-reduce; else the eager step under
    DistributedDataParallel."""
    import torch.distributed as dist
    from . import _lib
    from .hosts.resnet import resnet50
    if ops is None:
        from . import cnsn as ops
    torch.backends.cudnn.benchmark = True      # imagenet.py:534
    torch.manual_seed(1 + rank)
    np.random.seed(1)                          # one coin per step for all ranks, as in the reference's single process
    if channels_last is None:                  # default: on, with this package's operators (they have the NHWC kernels)
        channels_last = ops.__name__.startswith("cnsn_b200")
    net = resnet50(fuse_post=fuse_post, ops=ops).to(dev).train()
    if channels_last:                          # parameters and activations in the layout cuDNN's convolutions work in
        net = net.to(memory_format=torch.channels_last)
    model = net
    _broadcast_model(net, world)
    if world > 1 and not graph:
        model = nn.parallel.DistributedDataParallel(net, device_ids=[dev.index], broadcast_buffers=False)
    opt = torch.optim.SGD(model.parameters(), 0.1, momentum=0.9, weight_decay=1e-4)
    x = torch.randn(batch, 3, 224, 224, device=dev)
    if channels_last:
        x = x.contiguous(memory_format=torch.channels_last)
    y = torch.randint(0, 1000, (batch,), device=dev)
    launches0 = _lib.launch_count()
    gs = GraphedStep(net, x, y, world, loss_fn=lambda n, xx, yy, aug: F.cross_entropy(n(xx, aug=False), yy)) if graph else None

    def one_step():
        if gs is not None:
            return gs.step_image_cn(x, y, opt, None, cn_prob, ops)
        return resnet50_step(model, x, y, opt, cn_prob, ops)

    for _ in range(warmup):
        one_step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    loss = 0.0
    for _ in range(steps):
        loss = one_step()
    t1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return {"metric": "ResNet-50 + SelfNorm training images/s", "value": world * batch * steps / (ms * 1e-3),
            "unit": "images/s", "ms_per_step": ms / steps, "steps": steps, "warmup": warmup, "batch_per_gpu": batch,
            "n_gpus": world, "dtype": "f32 (TF32 convolutions: %s)" % torch.backends.cudnn.allow_tf32,
            "config": "resnet50 cnsn_type=sn pos=post, image-space CrossNorm crop=neither beta=1 cn_prob=%g, SGD lr 0.1 "
                      "momentum 0.9 wd 1e-4, synthetic 224x224, fuse_post=%s" % (cn_prob, bool(fuse_post)),
            "final_loss": loss, "params": sum(p.numel() for p in net.parameters()),
            "graph": "CUDA graph for the network's forward + loss + backward on every step (image-space CrossNorm eager, in "
                     "front); one flat-buffer gradient all-reduce" if gs is not None and gs.graph is not None else "eager",
            "memory_format": "channels_last" if channels_last else "contiguous (NCHW)",
            "cnsn_kernel_launches": _lib.launch_count() - launches0}


def resnet50_jsd_step(net, images_all, targets, opt, cn_prob, ops, jsd, beta=1, crop="neither", autocast=True):
    """One ``train_cn_image_consist`` step of the reference's imagenet.py:348-385: the three views of the batch are
    concatenated, a coin sends the whole 3B batch through image-space CrossNorm, one forward, cross-entropy on the
    clean third plus 12 x the Jensen-Shannon consistency of the three thirds."""
    if np.random.rand(1) < cn_prob:
        images_all = ops.cn_op_2ins_space_chan(images_all, beta=beta, crop=crop)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        logits_all = net(images_all, aug=False)
    lc, l1, l2 = torch.split(logits_all, targets.size(0))
    loss = F.cross_entropy(lc, targets) + 12 * jsd(lc, l1, l2)
    opt.zero_grad()
    loss.backward()
    opt.step()
    return float(loss.detach())


def bench_resnet50_jsd(dev, world, rank, batch=256, steps=5, warmup=2, cn_prob=0.5, fuse_post=True, graph=True,
                       channels_last=True):
    """images/s (clean images: the step processes 3x as many views) of ResNet-50 + SelfNorm with the 3-view JSD
    consistency step, bf16 autocast (BASELINE config 5: 3 x 256 = 768 views per GPU).  graph: as bench_resnet50."""
    import torch.distributed as dist
    from . import _lib, cnsn as ops
    from .hosts.resnet import resnet50
    from .losses import jsd_consistency
    torch.backends.cudnn.benchmark = True      # imagenet.py:534
    torch.manual_seed(1 + rank)
    np.random.seed(1)                          # one coin per step for all ranks, as in the reference's single process
    net = resnet50(fuse_post=fuse_post).to(dev).train()
    if channels_last:
        net = net.to(memory_format=torch.channels_last)
    model = net
    _broadcast_model(net, world)
    if world > 1 and not graph:
        model = nn.parallel.DistributedDataParallel(net, device_ids=[dev.index], broadcast_buffers=False)
    opt = torch.optim.SGD(model.parameters(), 0.1, momentum=0.9, weight_decay=1e-4)
    x = torch.randn(3 * batch, 3, 224, 224, device=dev)
    if channels_last:
        x = x.contiguous(memory_format=torch.channels_last)
    y = torch.randint(0, 1000, (batch,), device=dev)
    launches0 = _lib.launch_count()

    def jsd_loss(n, xx, yy, aug):              # imagenet.py:352-380: one forward of the 3B batch, CE on the clean third + 12 x JSD
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logits_all = n(xx, aug=False)
        lc, l1, l2 = torch.split(logits_all, yy.size(0))
        return F.cross_entropy(lc, yy) + 12 * jsd_consistency(lc, l1, l2)

    gs = GraphedStep(net, x, y, world, loss_fn=jsd_loss) if graph else None

    def one_step():
        if gs is not None:
            return gs.step_image_cn(x, y, opt, None, cn_prob, ops)
        return resnet50_jsd_step(model, x, y, opt, cn_prob, ops, jsd_consistency)

    for _ in range(warmup):
        one_step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    loss = 0.0
    for _ in range(steps):
        loss = one_step()
    t1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return {"metric": "ResNet-50 + SelfNorm + JSD (3 views) training images/s", "value": world * batch * steps / (ms * 1e-3),
            "unit": "clean images/s", "views_per_s": 3 * world * batch * steps / (ms * 1e-3), "ms_per_step": ms / steps,
            "steps": steps, "warmup": warmup, "batch_per_gpu": batch, "views_per_gpu": 3 * batch, "n_gpus": world,
            "dtype": "bf16 autocast (fp32 parameters, fp32 SelfNorm statistics)",
            "config": "resnet50 cnsn_type=sn pos=post, image-space CrossNorm cn_prob=%g, CE + 12 x JSD (cnsn_jsd kernels), SGD lr "
                      "0.1 momentum 0.9 wd 1e-4, synthetic 224x224, fuse_post=%s" % (cn_prob, bool(fuse_post)),
            "final_loss": loss, "cnsn_kernel_launches": _lib.launch_count() - launches0,
            "memory_format": "channels_last" if channels_last else "contiguous (NCHW)",
            "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 1e9}


def bench_resnet50_cpu(ops, batch=16, steps=1, warmup=1, cn_prob=0.5):
    """The same ResNet-50 step on the host cores with a caller-supplied operator set (the CPU reference arm of
    bench.py passes the eager-PyTorch restatement); a bounded sample of BASELINE config 4 (batch 16 of 256)."""
    import time
    from .hosts.resnet import resnet50
    torch.manual_seed(1)
    np.random.seed(1)
    net = resnet50(fuse_post=False, ops=ops).train()
    opt = torch.optim.SGD(net.parameters(), 0.1, momentum=0.9, weight_decay=1e-4)
    x, y = torch.randn(batch, 3, 224, 224), torch.randint(0, 1000, (batch,))
    for _ in range(warmup):
        resnet50_step(net, x, y, opt, cn_prob, ops)
    t0 = time.perf_counter()
    for _ in range(steps):
        loss = resnet50_step(net, x, y, opt, cn_prob, ops)
    dt = time.perf_counter() - t0
    return {"metric": "ResNet-50 + SelfNorm training images/s", "value": batch * steps / dt, "unit": "images/s",
            "ms_per_step": dt / steps * 1e3, "batch_per_gpu": batch, "final_loss": loss,
            "sample": "batch %d (of 256), %d step(s) after %d warm-up, CPU, eager-PyTorch CNSN" % (batch, steps, warmup)}
