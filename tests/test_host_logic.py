"""CPU tests of the package's HOST logic with an oracle-backed stand-in for the CUDA backend:
module surface, state_dict compatibility, .active protocol, RNG consumption order, autograd wiring,
and the drop-in of the package into the reference's unmodified model files.
No CUDA compute happens here; numerical parity of the kernels is tests/test_gpu_parity.py (-m gpu).
"""
import sys

import numpy as np
import pytest
import torch

import helpers as H
from _refload import load_reference_cnsn, reference_root
from fake_backend import OracleBackend


@pytest.fixture()
def mod():
    import cnsn_b200._lib as L
    import cnsn_b200.cnsn as m
    fake = OracleBackend()
    old = L.set_backend_for_tests(fake)
    m._fake = fake
    yield m
    L.set_backend_for_tests(old)


def test_product_path_has_no_cpu_fallback():
    import cnsn_b200._lib as L
    import cnsn_b200.cnsn as m
    L.set_backend_for_tests(None)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.SelfNorm(4)(torch.randn(2, 4, 3, 3))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.cn_op_2ins_space_chan(torch.randn(2, 4, 3, 3))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.calc_ins_mean_std(torch.randn(2, 4, 3, 3))


def test_package_never_imports_oracle():
    import os
    import re
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "crossnorm-selfnorm_b200")
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f
                assert "fake_backend" not in src, f


@pytest.mark.parametrize("name", H.golden_names("selfnorm_"))
def test_selfnorm_wiring_golden(mod, name):
    g = H.golden(name)
    params, bufs = H.sn_params_from_golden(g)
    two, training = bool(g["is_two"]), bool(g["training"])
    r = H.run_selfnorm(mod, g["x"], g["dy"], params, bufs, "cpu", two, training, torch.float64)
    assert H.maxabs(r["y"], g["y_f64"]) < 1e-6 and H.maxabs(r["dx"], g["dx_f64"]) < 1e-6
    for tag in ("g", "f") if two else ("g",):
        for k in ("w", "gamma", "beta"):
            assert H.relmax(r[f"d{tag}_{k}"], g[f"d{tag}_{k}_f64"]) < 1e-5
        assert H.maxabs(r[f"{tag}_rm_after"], g[f"{tag}_rm_after_f64"]) < 1e-6
        assert H.maxabs(r[f"{tag}_rv_after"], g[f"{tag}_rv_after_f64"]) < 1e-6
        assert int(r[f"{tag}_nbt_after"]) == int(g[f"{tag}_nbt_after"])


@pytest.mark.parametrize("name", H.golden_names("crossnorm_"))
def test_crossnorm_wiring_and_rng_golden(mod, name):
    """Same seeds as the reference run -> same perm / boxes -> same output (pins the RNG contract)."""
    g = H.golden(name)
    y, dx = H.run_crossnorm(mod, g["x"], g["dy"], "cpu", str(g["crop"]), bool(g["chan"]), H.lam_of(g),
                            g["torch_seed"], g["numpy_seed"], torch.float64)
    assert H.maxabs(y, g["y_f64"]) < 1e-6
    assert H.maxabs(dx, g["dx_f64"]) < 1e-6


def test_cn_rand_bbox_matches_reference_stream():
    import cnsn_b200.cnsn as m
    g = H.golden("rng_stream")
    torch.manual_seed(int(g["seed"]))
    np.random.seed(int(g["seed"]) + 1)
    for i in range(int(g["n"])):
        shape, crop = tuple(int(v) for v in g[f"shape{i}"]), str(g[f"crop{i}"])
        perm = torch.randperm(shape[0]).numpy()
        assert np.array_equal(perm, g[f"perm{i}"])
        for key, on in (("sw", crop in ("style", "both")), ("cw", crop in ("content", "both"))):
            if on:
                b = m.cn_rand_bbox(shape, beta=1, bbx_thres=0.1)
                assert (int(b[0]), int(b[2]), int(b[1]), int(b[3])) == tuple(int(v) for v in g[f"{key}{i}"])


def test_stats_and_mix_autograd(mod):
    torch.manual_seed(0)
    c = torch.randn(3, 4, 6, 5, dtype=torch.float64, requires_grad=True)
    s = torch.randn(3, 4, 4, 7, dtype=torch.float64, requires_grad=True)
    out = mod.instance_norm_mix(c, s)
    w = torch.randn_like(out)
    (out * w).sum().backward()
    # same computation in plain torch (the reference formula, cnsn.py:8-29)
    c2, s2 = c.detach().clone().requires_grad_(True), s.detach().clone().requires_grad_(True)
    def st(x):
        v = x.reshape(3, 4, -1)
        return v.mean(2).view(3, 4, 1, 1), (v.var(2) + 1e-5).sqrt().view(3, 4, 1, 1)
    sm, ss = st(s2)
    cm, cs = st(c2)
    ref = (c2 - cm) / cs * ss + sm
    (ref * w).sum().backward()
    assert torch.allclose(out, ref, atol=1e-10)
    assert torch.allclose(c.grad, c2.grad, atol=1e-9) and torch.allclose(s.grad, s2.grad, atol=1e-9)
    m, sd = mod.calc_ins_mean_std(c)
    assert m.shape == sd.shape == (3, 4, 1, 1)
    with pytest.raises(AssertionError):
        mod.calc_ins_mean_std(torch.randn(3, 4, 5))


def test_crossnorm_active_protocol(mod):
    cn = mod.CrossNorm(crop="neither", beta=1)
    assert list(cn.state_dict().keys()) == [] and bool(cn) and cn.active is False
    x = torch.randn(4, 3, 5, 5)
    cn.train()
    assert cn(x) is x                       # inactive -> identity, same object
    cn.active = True
    y = cn(x)
    assert cn.active is False and y is not x and mod._fake.calls == ["crossnorm_fwd"]
    cn.eval()
    cn.active = True
    assert cn(x) is x and cn.active is False    # reset even in eval (cnsn.py:108)
    with pytest.raises(AssertionError):
        mod.cn_op_2ins_space_chan(x, crop="nope")


def test_cnsn_composition(mod):
    x = torch.randn(4, 3, 5, 5)
    assert mod.CNSN(None, None)(x) is x
    cn, sn = mod.CrossNorm("neither", 1), mod.SelfNorm(3)
    blk = mod.CNSN(cn, sn).train()
    blk(x)
    assert mod._fake.calls == ["selfnorm_fwd"]          # inactive CrossNorm is not even called (cnsn.py:160)
    cn.active = True
    blk(x)
    assert mod._fake.calls[1:] == ["site_fwd"] and cn.active is False       # both fire: one fused call (8f-2)
    cn.active = True
    mod.CNSN.fuse_site = False
    try:
        blk(x)
    finally:
        mod.CNSN.fuse_site = True
    assert mod._fake.calls[2:] == ["crossnorm_fwd", "selfnorm_fwd"] and cn.active is False
    cn.active = True
    blk.eval()
    assert blk(x) is not None and mod._fake.calls[4:] == ["selfnorm_fwd"] and cn.active is False   # eval: CN is a no-op


@pytest.mark.parametrize("crop", ["neither", "style", "content", "both"])
@pytest.mark.parametrize("tail", [None, "relu", "res+relu"])
def test_cnsn_site_fusion_equals_two_operator_sequence(mod, crop, tail):
    """When both operators fire, CNSN.forward takes ONE fused call per direction; values, gradients, parameter
    gradients, running statistics, RNG consumption and the .active protocol equal the two-operator sequence."""
    shape = (6, 4, 6, 4)
    g = torch.Generator().manual_seed(11)
    x0, r0, dy = (torch.randn(shape, generator=g, dtype=torch.float64) for _ in range(3))
    res = []
    for fused in (False, True):
        mod.CNSN.fuse_site = fused
        try:
            torch.manual_seed(5)
            np.random.seed(6)
            blk = mod.CNSN(crossnorm=mod.CrossNorm(crop=crop, beta=1), selfnorm=mod.SelfNorm(4)).double().train()
            blk.crossnorm.active = True
            x, r = x0.clone().requires_grad_(True), r0.clone().requires_grad_(True)
            n0 = len(mod._fake.calls)
            if tail is None:
                y = blk(x)
            else:
                y = blk(x, r if tail == "res+relu" else None, True)
            y.backward(dy)
        finally:
            mod.CNSN.fuse_site = True
        calls = mod._fake.calls[n0:]
        assert ("site_fwd" in calls and "site_bwd" in calls) == fused and ("crossnorm_fwd" in calls) != fused
        assert blk.crossnorm.active is False
        res.append((y.detach(), x.grad, r.grad if tail == "res+relu" else x.grad, blk.selfnorm.g_fc.weight.grad,
                    blk.selfnorm.g_bn.weight.grad, blk.selfnorm.g_bn.bias.grad, blk.selfnorm.g_bn.running_mean.clone(),
                    blk.selfnorm.g_bn.running_var.clone(), blk.selfnorm.g_bn.num_batches_tracked.clone(),
                    torch.rand(1), np.random.rand()))
    for a, b in zip(*res):
        assert np.allclose(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64), atol=1e-12), (a, b)


@pytest.mark.parametrize("name", H.golden_names("site_"))
def test_cnsn_site_wiring_golden(mod, name):
    """The reference's CNSN module with both operators firing (fixture from the unmodified reference): same seeds,
    same draws, same outputs and gradients through the fused-site host path."""
    g = H.golden(name)
    params, bufs = H.sn_params_from_golden(g)
    C = g["x"].shape[1]
    sn = H.make_selfnorm(mod, C, params, bufs, "cpu").double()
    blk = mod.CNSN(mod.CrossNorm(crop=str(g["crop"]), beta=1), sn).train()
    blk.crossnorm.active = True
    torch.manual_seed(int(g["torch_seed"]))
    np.random.seed(int(g["numpy_seed"]))
    x = torch.from_numpy(g["x"]).double().requires_grad_(True)
    y = blk(x)
    y.backward(torch.from_numpy(g["dy"]).double())
    assert "site_fwd" in mod._fake.calls and "site_bwd" in mod._fake.calls
    assert H.maxabs(y.detach().numpy(), g["y_f64"]) < 1e-9 and H.maxabs(x.grad.numpy(), g["dx_f64"]) < 1e-9
    # the backend hands parameter gradients back as fp32 tensors (as the CUDA one does)
    assert H.relmax(sn.g_fc.weight.grad.view(C, 2).numpy(), g["dg_w_f64"]) < 1e-6
    assert H.relmax(sn.g_bn.weight.grad.numpy(), g["dg_gamma_f64"]) < 1e-6
    assert H.maxabs(sn.g_bn.running_var.numpy(), g["g_rv_after_f64"]) < 1e-6


@pytest.mark.parametrize("fire", [False, True])
@pytest.mark.parametrize("relu", [False, True])
def test_cnsn_block_fusion_equals_unfused_sequence(mod, fire, relu):
    """CNSN.forward(x, residual, relu) == relu?(CNSN(x + residual)) -- values, input gradients, parameter
    gradients, running statistics, RNG consumption and the .active protocol -- whether or not CrossNorm fires."""
    shape = (6, 4, 5, 5)
    g = torch.Generator().manual_seed(7)
    x0, r0, dy = (torch.randn(shape, generator=g, dtype=torch.float64) for _ in range(3))
    res = []
    for fused in (False, True):
        torch.manual_seed(3)
        np.random.seed(4)
        blk = mod.CNSN(crossnorm=mod.CrossNorm(crop="both", beta=1), selfnorm=mod.SelfNorm(4)).double().train()
        blk.crossnorm.active = fire
        x, r = x0.clone().requires_grad_(True), r0.clone().requires_grad_(True)
        if fused:
            y = blk(x, r, relu)
        else:
            y = blk(torch.add(r, x))
            y = torch.relu(y) if relu else y
        y.backward(dy)
        assert blk.crossnorm.active is False
        res.append((y.detach(), x.grad, r.grad, blk.selfnorm.g_fc.weight.grad, blk.selfnorm.g_bn.weight.grad,
                    blk.selfnorm.g_bn.running_var.clone(), torch.rand(1), np.random.rand()))
    for a, b in zip(*res):
        assert np.allclose(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64), atol=1e-12), (a, b)
    if not fire:
        assert "selfnorm_block_fwd" in mod._fake.calls and "selfnorm_block_bwd" in mod._fake.calls


def test_selfnorm_state_dict_matches_reference(mod):
    ref = load_reference_cnsn()
    if ref is None:
        pytest.skip("reference checkout not present")
    for two in (False, True):
        torch.manual_seed(3)
        a = ref.SelfNorm(8, is_two=two)
        torch.manual_seed(3)
        b = mod.SelfNorm(8, is_two=two)
        sa, sb = a.state_dict(), b.state_dict()
        assert list(sa.keys()) == list(sb.keys())
        for k in sa:
            assert sa[k].shape == sb[k].shape and torch.equal(sa[k], sb[k]), k   # same init draws
        b.load_state_dict(sa)


def test_selfnorm_batch1_raises_like_reference(mod):
    with pytest.raises(ValueError, match="Expected more than 1 value per channel"):
        mod.SelfNorm(4).train()(torch.randn(1, 4, 3, 3))


def _reference_host(modname, cnsn_module):
    """Import a reference host-model file with `models.cnsn` swapped for ours."""
    root = reference_root()
    if root is None:
        pytest.skip("reference checkout not present")
    if not hasattr(np, "int"):
        np.int = int
    saved = {k: v for k, v in sys.modules.items() if k == "models" or k.startswith("models.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, root)
    try:
        import models
        sys.modules["models.cnsn"] = cnsn_module
        models.cnsn = cnsn_module
        host = __import__(modname, fromlist=["x"])
    finally:
        sys.path.remove(root)
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
            del sys.modules[k]
        sys.modules.update(saved)
    return host


def test_dropin_wideresnet_matches_reference(mod):
    """Unmodified reference WideResNet + our cnsn module == unmodified reference end to end
    (same seeds -> same active sites, perms, boxes; logits and gradients agree)."""
    ref = load_reference_cnsn()
    if ref is None:
        pytest.skip("reference checkout not present")
    kw = dict(depth=10, num_classes=10, widen_factor=1, active_num=2, pos="post", beta=1, crop="both",
              cnsn_type="cnsn")
    ours = _reference_host("models.cifar.wideresnet_cnsn", mod)
    theirs = _reference_host("models.cifar.wideresnet_cnsn", ref)
    torch.manual_seed(0)
    na = theirs.WideResNet(**kw).double().train()
    torch.manual_seed(0)
    nb = ours.WideResNet(**kw).double().train()
    assert type(nb.cn_modules[0]) is mod.CrossNorm and len(nb.cn_modules) == len(na.cn_modules)
    assert list(na.state_dict().keys()) == list(nb.state_dict().keys())
    nb.load_state_dict(na.state_dict())
    x = torch.randn(6, 3, 32, 32, dtype=torch.float64)
    outs = []
    for net in (na, nb):
        torch.manual_seed(5)
        np.random.seed(6)
        out = net(x, aug=True)
        out.square().sum().backward()
        outs.append(out)
    # both active sites take the fused CrossNorm -> SelfNorm call, the other sites the plain SelfNorm
    assert mod._fake.calls.count("site_fwd") == 2 and mod._fake.calls.count("site_bwd") == 2
    assert "selfnorm_bwd" in mod._fake.calls and "crossnorm_fwd" not in mod._fake.calls
    assert torch.allclose(outs[0], outs[1], atol=1e-8)
    for (ka, pa), (kb, pb) in zip(na.named_parameters(), nb.named_parameters()):
        assert ka == kb and torch.allclose(pa.grad, pb.grad, atol=1e-7, rtol=1e-6), ka
    for (ka, ba), (kb, bb) in zip(na.named_buffers(), nb.named_buffers()):
        assert torch.allclose(ba.double(), bb.double(), atol=1e-9), ka


@pytest.mark.parametrize("host,ctor,kw,xshape", [
    ("models.cifar.resnext_cnsn", "CifarResNeXt",
     dict(depth=11, cardinality=2, base_width=8, num_classes=10, active_num=2, pos="post", beta=1, crop="both",
          cnsn_type="cnsn"), (4, 3, 32, 32)),
    ("models.imagenet.resnet_cnsn", "ResNet",
     # crop='neither': at 64x64 input the last stage has 2x2 planes, where the reference's box sampler never accepts
     dict(layers=[1, 1, 1, 1], num_classes=7, active_num=2, pos="post", beta=1, crop="neither", cnsn_type="cnsn"),
     (4, 3, 64, 64)),
    ("models.imagenet.resnet_cnsn", "ResNet",
     dict(layers=[1, 1, 1, 1], num_classes=7, active_num=1, pos="post", beta=1, crop="neither", cnsn_type="sn"),
     (4, 3, 64, 64)),
])
def test_dropin_other_reference_hosts(mod, host, ctor, kw, xshape):
    """The other callers SURVEY.md 8b lists -- the reference's unmodified CIFAR ResNeXt and ImageNet ResNet files --
    with `models.cnsn` swapped for this package: same seeds -> same active sites and draws, logits, gradients and
    buffers equal to the unmodified reference's."""
    ref = load_reference_cnsn()
    if ref is None:
        pytest.skip("reference checkout not present")
    ours = _reference_host(host, mod)
    theirs = _reference_host(host, ref)
    torch.manual_seed(0)
    na = getattr(theirs, ctor)(**kw).double().train()
    torch.manual_seed(0)
    nb = getattr(ours, ctor)(**kw).double().train()
    assert list(na.state_dict().keys()) == list(nb.state_dict().keys())
    nb.load_state_dict(na.state_dict())
    x = torch.randn(*xshape, dtype=torch.float64)
    outs = []
    for net in (na, nb):
        torch.manual_seed(5)
        np.random.seed(6)
        out = net(x, aug=kw["cnsn_type"] == "cnsn")      # 'sn' hosts have no feature-space CrossNorm (imagenet.py:215)
        out.square().sum().backward()
        outs.append(out)
    if kw["cnsn_type"] == "cnsn":
        assert mod._fake.calls.count("site_fwd") == kw["active_num"] == mod._fake.calls.count("site_bwd")
    assert "selfnorm_fwd" in mod._fake.calls and "selfnorm_bwd" in mod._fake.calls
    assert torch.allclose(outs[0], outs[1], atol=1e-8)
    for (ka, pa), (kb, pb) in zip(na.named_parameters(), nb.named_parameters()):
        assert ka == kb and torch.allclose(pa.grad, pb.grad, atol=1e-7, rtol=1e-6), ka
    for (ka, ba), (kb, bb) in zip(na.named_buffers(), nb.named_buffers()):
        assert torch.allclose(ba.double(), bb.double(), atol=1e-9), ka


def test_cnsn_site_fusion_declines_what_the_kernels_do_not_cover(mod):
    """Two-gate SelfNorm, channel permutation, eval-mode BatchNorm inside the gate, a backend that says the shape does
    not fit: CNSN.forward runs the two operators in sequence (same results as ever), never the fused call."""
    import functools
    x = torch.randn(4, 3, 6, 4, dtype=torch.float64)

    def calls_of(blk):
        blk.crossnorm.active = True
        n0 = len(mod._fake.calls)
        blk(x)
        assert blk.crossnorm.active is False
        return mod._fake.calls[n0:]

    two = mod.CNSN(mod.CrossNorm("neither", 1), mod.SelfNorm(3, is_two=True)).double().train()
    assert calls_of(two) == ["crossnorm_fwd", "selfnorm_fwd"]
    chan = mod.CNSN(mod.CrossNorm("neither", 1), mod.SelfNorm(3)).double().train()
    chan.crossnorm.cn_op = functools.partial(mod.cn_op_2ins_space_chan, crop="neither", beta=1, chan=True)
    assert calls_of(chan) == ["crossnorm_fwd", "selfnorm_fwd"]
    frozen = mod.CNSN(mod.CrossNorm("neither", 1), mod.SelfNorm(3)).double().train()
    frozen.selfnorm.g_bn.eval()                      # gate BatchNorm on running statistics
    assert calls_of(frozen) == ["crossnorm_fwd", "selfnorm_fwd"]
    plain = mod.CNSN(mod.CrossNorm("neither", 1), mod.SelfNorm(3)).double().train()
    mod._fake.site_supported = lambda t: False       # e.g. 7x7 planes, 224x224 image planes
    try:
        assert calls_of(plain) == ["crossnorm_fwd", "selfnorm_fwd"]
    finally:
        del mod._fake.site_supported
    assert calls_of(plain) == ["site_fwd"]
    only_cn = mod.CNSN(mod.CrossNorm("neither", 1), None).train()
    assert calls_of(only_cn) == ["crossnorm_fwd"]


def test_graphed_step_flat_gradients_follow_channels_last_parameters():
    """train.GraphedStep keeps every gradient as a view of ONE flat buffer (the single all-reduce of the data-parallel step).
    For channels_last parameters the views carry the parameters' strides (autograd's gradient layout contract), and the
    step trains exactly like the plain eager step (CPU: no graph capture, same arithmetic)."""
    import copy
    import torch.nn as nn
    import torch.nn.functional as F
    from cnsn_b200.train import GraphedStep

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.c1, self.c2, self.fc = nn.Conv2d(3, 8, 3, padding=1), nn.Conv2d(8, 4, 1), nn.Linear(4, 5)

        def forward(self, x, aug=False):
            return self.fc(F.relu(self.c2(F.relu(self.c1(x)))).mean((2, 3)))

    torch.manual_seed(0)
    a = Net().to(memory_format=torch.channels_last)
    b = copy.deepcopy(a)
    x = torch.randn(6, 3, 8, 8).contiguous(memory_format=torch.channels_last)
    y = torch.randint(0, 5, (6,))
    gs = GraphedStep(a, x, y, 1, capture=False)
    for p in a.parameters():
        live = [d for d in range(p.dim()) if p.size(d) > 1]          # the stride of a size-1 dimension is arbitrary
        assert [p.grad.stride(d) for d in live] == [p.stride(d) for d in live]
        assert p.grad.untyped_storage().data_ptr() == gs.flat.untyped_storage().data_ptr()
    assert not a.c1.weight.is_contiguous() and a.c1.weight.grad.is_contiguous(memory_format=torch.channels_last)
    oa, ob = torch.optim.SGD(a.parameters(), 0.1, momentum=0.9), torch.optim.SGD(b.parameters(), 0.1, momentum=0.9)
    for _ in range(3):
        la = gs.step(x, y, oa, None, 0.0)
        ob.zero_grad()
        lb = F.cross_entropy(b(x), y)
        lb.backward()
        ob.step()
        assert abs(la - float(lb)) < 1e-6
    for p, q in zip(a.parameters(), b.parameters()):
        assert torch.allclose(p, q, atol=1e-6)
        assert p.grad.untyped_storage().data_ptr() == gs.flat.untyped_storage().data_ptr()      # still views after the steps
