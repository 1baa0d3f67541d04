"""Model-level GPU parity for BASELINE configs 3 / 4 / 5: one whole training step of the host models pushed through the
CUDA kernels, against the SAME step with the reference's eager CNSN operators on the same GPU.

Reference side: the reference's OWN files from oracle/_ref (``models/cnsn.py`` inside the reference's own
``WideResNet`` / ``ResNet`` classes; copied at build time by oracle/build_ref.py, travels to the GPU box); when that
copy is absent, this package's host classes with ``ops=oracle.eager_modules`` (the bit-identical eager port).
Product side: ``cnsn_b200.hosts`` with ``ops=cnsn_b200.cnsn`` -- every feature map goes through libcnsn_b200.so.

Both sides start from the same state dict, see the same batch and the same host RNG state (so the per-step coin, the
active sites, permutations and crop boxes are identical: reference draw order models/cnsn.py:61-77,
wideresnet_cnsn.py:199-203), run forward, cross-entropy, backward and one SGD step (cifar.py:117-145,
imagenet.py:195-250, :337-406).  Compared: logits, loss, EVERY parameter gradient, every BatchNorm / SelfNorm buffer
and every parameter after the step.  TF32 is switched off and cuDNN is deterministic, so everything that is not a
CNSN operator is bit-identical for identical inputs and the differences measure the kernels.

Tolerances.  1e-5 is the per-operator bar (tests/test_gpu_parity.py).  A whole step of an untrained 40-50 layer network
is a different matter: its fp32 gradient is ill-conditioned -- the reference's OWN fp32 run differs from the reference run
in fp64 by 1-5 % of max |grad| per parameter tensor (measured here, every time; first seen in
tools/debug/model_grad_debug.py) -- so "1e-5 against the reference's fp32 numbers" would only compare two roundings of
the same noise.  The truth is the reference step in fp64 on the same GPU, and the bar is: the CUDA path is as close to
that truth as the reference's fp32 run is -- forward quantities (logits, loss, buffers) within 5e-4 (measured: 0.4-2.5e-4, the reference's fp32 run 0.6-1.7e-4) and never more than
3x the reference's own fp32 error; gradients and updated parameters: median error over the parameter tensors <= 1.5x the
reference's, worst tensor <= 3x the reference's worst.  Every measured figure goes to gpurun_out/model_parity.json.
"""
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_RECORD = {}


@pytest.fixture(autouse=True)
def _exact_convs():
    """TF32 off, deterministic cuDNN: the non-CNSN layers are then bit-identical on both sides."""
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.deterministic,
             torch.backends.cudnn.benchmark)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False
    yield
    (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.deterministic,
     torch.backends.cudnn.benchmark) = saved


def _reference_classes():
    """(WideResNet class, ResNet class, cnsn module, label) of the reference side."""
    from oracle import build_ref
    wrn = build_ref.load("models.cifar.wideresnet_cnsn")
    if wrn is not None:
        rn = build_ref.load("models.imagenet.resnet_cnsn")
        return wrn.WideResNet, rn.ResNet, build_ref.load("models.cnsn"), "reference files (oracle/_ref)"
    from functools import partial
    from cnsn_b200.hosts import ResNet, WideResNet
    from oracle import eager_modules
    return partial(WideResNet, ops=eager_modules), partial(ResNet, ops=eager_modules), eager_modules, "eager port"


def _rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-12))


def _errors(net, out, truth_net, truth_out):
    """Relative (to the truth tensor's max magnitude) errors of one fp32 run against the fp64 truth."""
    e = {"logits": _rel(out["logits"], truth_out["logits"]),
         "loss": abs(float(out["loss"]) - float(truth_out["loss"])) / max(abs(float(truth_out["loss"])), 1e-12)}
    assert list(out["grads"]) == list(truth_out["grads"])
    e["grads"] = {k: _rel(out["grads"][k], g) for k, g in truth_out["grads"].items()}
    tb, ob = dict(truth_net.named_buffers()), dict(net.named_buffers())
    assert list(tb) == list(ob)
    e["bufs"] = {k: _rel(ob[k], v) for k, v in tb.items() if v.dtype.is_floating_point}
    e["ints_equal"] = all(torch.equal(ob[k], v) for k, v in tb.items() if not v.dtype.is_floating_point)
    e["params"] = {}
    for (k, p), (k2, q) in zip(truth_net.named_parameters(), net.named_parameters()):
        assert k == k2
        e["params"][k] = _rel(q, p)
    return e


def _compare(tag, truth, ref, ours):
    """truth / ref / ours: (net, out) of the reference in fp64, the reference in fp32, the CUDA path in fp32."""
    er, eo = _errors(ref[0], ref[1], *truth), _errors(ours[0], ours[1], *truth)
    worst = lambda d: max(d.items(), key=lambda kv: kv[1]) if d else ("", 0.0)      # noqa: E731
    med = lambda d: float(np.median(list(d.values())))                              # noqa: E731
    rec = {"logits": {"ours": eo["logits"], "reference_fp32": er["logits"]},
           "loss": {"ours": eo["loss"], "reference_fp32": er["loss"]}}
    for key in ("grads", "bufs", "params"):
        rec[key] = {"ours_median": med(eo[key]), "reference_fp32_median": med(er[key]),
                    "ours_worst": worst(eo[key]), "reference_fp32_worst": worst(er[key]), "tensors": len(eo[key])}
    _RECORD[tag] = rec
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "model_parity.json"), "w") as f:
        json.dump(_RECORD, f, indent=1)
    for key in ("logits", "loss"):
        assert eo[key] <= 5e-4 and eo[key] <= max(3 * er[key], 2e-5), (tag, key, rec[key])
    assert eo["ints_equal"], (tag, "integer buffers (num_batches_tracked)")
    r = rec["bufs"]
    assert r["ours_worst"][1] <= 5e-4 and r["ours_worst"][1] <= max(3 * r["reference_fp32_worst"][1], 2e-5), (tag, "buffers", r)
    for key in ("grads", "params"):
        r = rec[key]
        assert r["ours_median"] <= 1.5 * r["reference_fp32_median"] + 1e-6, (tag, key, r)
        assert r["ours_worst"][1] <= 3 * r["reference_fp32_worst"][1] + 1e-6, (tag, key, r)


def _train_step(net, opt, fwd):
    """forward (callable -> logits, loss), zero_grad, backward, SGD step; returns logits, loss, gradients."""
    logits, loss = fwd(net)
    opt.zero_grad()
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in net.named_parameters() if p.grad is not None}
    opt.step()
    return {"logits": logits.detach(), "loss": loss.detach(), "grads": grads}


@pytest.mark.parametrize("aug", [True, False])
@pytest.mark.parametrize("fuse_post,fuse_site,channels_last", [(True, True, False), (False, True, False), (True, False, False),
                                                               (False, False, False), (True, True, True), (False, True, True)])
def test_wrn40_2_step_matches_reference(aug, fuse_post, fuse_site, channels_last):
    """BASELINE config 3: WideResNet-40-2, cnsn_type='cnsn', pos='post', crop='both', active_num=2 (cifar10-scripts/
    wideresnet/run-cnsn.sh), SGD nesterov lr 0.1 wd 5e-4 (cifar.py:398-402); batch 64 of synthetic 32x32 images.
    channels_last: this package's network and its input in torch.channels_last (the layout train.bench_wrn runs in):
    SelfNorm sites through the NHWC kernels, batch norm through cuDNN's NHWC kernels, a site whose CrossNorm fires
    through the NCHW kernels and back."""
    import cnsn_b200
    import cnsn_b200.cnsn as M
    from cnsn_b200.hosts import WideResNet
    RefWRN, _, _, label = _reference_classes()
    kw = dict(widen_factor=2, active_num=2, pos="post", beta=1, crop="both", cnsn_type="cnsn")
    torch.manual_seed(0)
    a = RefWRN(40, 10, **kw).to(DEV).train()
    t = RefWRN(40, 10, **kw).to(DEV).double().train()
    t.load_state_dict(a.state_dict())
    b = WideResNet(40, 10, fuse_post=fuse_post, **kw).to(DEV).train()
    b.load_state_dict(a.state_dict())
    if channels_last:
        b = b.to(memory_format=torch.channels_last)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(64, 3, 32, 32, generator=g).to(DEV)
    y = torch.randint(0, 10, (64,), generator=g).to(DEV)

    def fwd(net):
        torch.manual_seed(5)
        np.random.seed(6)
        xin = x.to(next(net.parameters()).dtype)
        if channels_last and net is b:
            xin = xin.contiguous(memory_format=torch.channels_last)
        logits = net(xin, aug=aug)
        return logits, F.cross_entropy(logits, y)

    outs = []
    M.CNSN.fuse_site = fuse_site
    try:
        for net in (t, a, b):
            n0 = cnsn_b200.launch_count()
            opt = torch.optim.SGD(net.parameters(), 0.1, momentum=0.9, weight_decay=5e-4, nesterov=True)
            outs.append(_train_step(net, opt, fwd))
            torch.cuda.synchronize()
            launched = cnsn_b200.launch_count() - n0
    finally:
        M.CNSN.fuse_site = True
    # 18 sites, one kernel per direction each (sites where CrossNorm fires unfused: two)
    assert launched >= 36, launched
    _compare("wrn40_2 aug=%s fuse_post=%s fuse_site=%s channels_last=%s vs %s" % (aug, fuse_post, fuse_site, channels_last, label),
             (t, outs[0]), (a, outs[1]), (b, outs[2]))


@pytest.mark.parametrize("cn_image", [True, False])
@pytest.mark.parametrize("fuse_post,channels_last", [(True, False), (False, False), (True, True)])
def test_resnet50_step_matches_reference(cn_image, fuse_post, channels_last):
    """BASELINE config 4: ResNet-50, cnsn_type='sn', pos='post' with image-space CrossNorm (imagenet.py:205-230,
    imagenet-scripts/run-cnsn.sh), SGD lr 0.1 momentum 0.9 wd 1e-4; batch 8 of synthetic 224x224 images (the 7x7
    stage then has N = 8 instances per channel: channel-group kernels)."""
    import cnsn_b200.cnsn as M
    from cnsn_b200.hosts import ResNet
    _, RefResNet, ref_ops, label = _reference_classes()
    kw = dict(num_classes=1000, active_num=1, pos="post", beta=1, crop="neither", cnsn_type="sn")
    torch.manual_seed(0)
    a = RefResNet([3, 4, 6, 3], **kw).to(DEV).train()
    t = RefResNet([3, 4, 6, 3], **kw).to(DEV).double().train()
    t.load_state_dict(a.state_dict())
    b = ResNet([3, 4, 6, 3], fuse_post=fuse_post, **kw).to(DEV).train()
    b.load_state_dict(a.state_dict())
    if channels_last:                        # the layout the benchmark runs in: NHWC SelfNorm and batch-norm kernels
        b = b.to(memory_format=torch.channels_last)
    g = torch.Generator().manual_seed(2)
    x = torch.randn(8, 3, 224, 224, generator=g).to(DEV)
    y = torch.randint(0, 1000, (8,), generator=g).to(DEV)

    def make_fwd(ops):
        def fwd(net):
            torch.manual_seed(5)
            np.random.seed(6)
            xx = x.to(next(net.parameters()).dtype)
            if channels_last and net is b:
                xx = xx.contiguous(memory_format=torch.channels_last)
            images = ops.cn_op_2ins_space_chan(xx, beta=1, crop="both") if cn_image else xx     # imagenet.py:215
            logits = net(images, aug=False)
            return logits, F.cross_entropy(logits, y)
        return fwd

    outs = []
    for net, ops in ((t, ref_ops), (a, ref_ops), (b, M)):
        opt = torch.optim.SGD(net.parameters(), 0.1, momentum=0.9, weight_decay=1e-4)
        outs.append(_train_step(net, opt, make_fwd(ops)))
    _compare("resnet50 cn_image=%s fuse_post=%s channels_last=%s vs %s" % (cn_image, fuse_post, channels_last, label),
             (t, outs[0]), (a, outs[1]), (b, outs[2]))


def _reference_jsd(lc, l1, l2):
    """imagenet.py:367-376: softmax, clamped log mixture, three kl_div 'batchmean', mean."""
    pc, p1, p2 = F.softmax(lc, dim=1), F.softmax(l1, dim=1), F.softmax(l2, dim=1)
    pm = torch.clamp((pc + p1 + p2) / 3., 1e-7, 1).log()
    return (F.kl_div(pm, pc, reduction='batchmean') + F.kl_div(pm, p1, reduction='batchmean') +
            F.kl_div(pm, p2, reduction='batchmean')) / 3.


@pytest.mark.parametrize("autocast,channels_last", [(False, False), (True, False), (False, True), (True, True)])
def test_resnet50_jsd_step_matches_reference(autocast, channels_last):
    """BASELINE config 5: the 3-view consistency step (imagenet.py:337-406): views concatenated, image-space CrossNorm on
    the whole 3B batch, one forward, CE on the clean third + 12 x JSD.  fp32: the whole-step tolerances above.  bf16
    autocast (what the benchmark runs): the reference's SelfNorm computes its statistics in bf16 there (SURVEY.md C.4:
    that alone moves a SelfNorm output by up to 0.37), ours in fp32 -- so the autocast leg checks the plumbing (finite,
    same loss within 5 %, gradients for every parameter), not 1e-2 agreement with the bf16-eager chain."""
    import cnsn_b200.cnsn as M
    from cnsn_b200.hosts import ResNet
    from cnsn_b200.losses import jsd_consistency
    _, RefResNet, ref_ops, label = _reference_classes()
    kw = dict(num_classes=1000, active_num=1, pos="post", beta=1, crop="neither", cnsn_type="sn")
    torch.manual_seed(0)
    a = RefResNet([3, 4, 6, 3], **kw).to(DEV).train()
    b = ResNet([3, 4, 6, 3], fuse_post=True, **kw).to(DEV).train()
    b.load_state_dict(a.state_dict())
    if channels_last:
        b = b.to(memory_format=torch.channels_last)
    nets = [(a, ref_ops, _reference_jsd), (b, M, jsd_consistency)]
    if not autocast:
        t = RefResNet([3, 4, 6, 3], **kw).to(DEV).double().train()
        t.load_state_dict(a.state_dict())
        nets.insert(0, (t, ref_ops, _reference_jsd))
    g = torch.Generator().manual_seed(3)
    B = 4
    x = torch.randn(3 * B, 3, 224, 224, generator=g).to(DEV)
    y = torch.randint(0, 1000, (B,), generator=g).to(DEV)

    def make_fwd(ops, jsd):
        def fwd(net):
            torch.manual_seed(5)
            np.random.seed(6)
            xx = x.to(next(net.parameters()).dtype)
            if channels_last and net is b:
                xx = xx.contiguous(memory_format=torch.channels_last)
            images = ops.cn_op_2ins_space_chan(xx, beta=1, crop="neither")   # imagenet.py:355
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                logits_all = net(images, aug=False)
            lc, l1, l2 = torch.split(logits_all, B)
            loss = F.cross_entropy(lc, y) + 12 * jsd(lc, l1, l2)
            return logits_all, loss
        return fwd

    outs = []
    for net, ops, jsd in nets:
        opt = torch.optim.SGD(net.parameters(), 0.1, momentum=0.9, weight_decay=1e-4)
        outs.append(_train_step(net, opt, make_fwd(ops, jsd)))
    if not autocast:
        _compare("resnet50 jsd fp32 channels_last=%s vs %s" % (channels_last, label), (nets[0][0], outs[0]), (a, outs[1]), (b, outs[2]))
        return
    assert torch.isfinite(outs[1]["logits"]).all() and torch.isfinite(outs[1]["loss"])
    assert abs(float(outs[1]["loss"]) - float(outs[0]["loss"])) <= 0.05 * abs(float(outs[0]["loss"]))
    assert list(outs[0]["grads"]) == list(outs[1]["grads"])
    assert all(torch.isfinite(v).all() for v in outs[1]["grads"].values())
