"""Host model (WideResNet + CNSN) against the live reference model file (build container only) and
its own invariants everywhere."""
import sys

import numpy as np
import pytest
import torch

from _refload import load_reference_cnsn, reference_root
from fake_backend import OracleBackend


@pytest.fixture()
def fake():
    import cnsn_b200._lib as L
    f = OracleBackend()
    old = L.set_backend_for_tests(f)
    yield f
    L.set_backend_for_tests(old)


def _reference_host(module, attr):
    root = reference_root()
    if root is None:
        pytest.skip("reference checkout not present")
    ref = load_reference_cnsn()
    saved = {k: v for k, v in sys.modules.items() if k == "models" or k.startswith("models.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, root)
    try:
        import models
        sys.modules["models.cnsn"] = ref
        models.cnsn = ref
        host = __import__(module, fromlist=["x"])
    finally:
        sys.path.remove(root)
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
            del sys.modules[k]
        sys.modules.update(saved)
    return getattr(host, attr)


def _reference_wrn():
    return _reference_host("models.cifar.wideresnet_cnsn", "WideResNet")


@pytest.mark.parametrize("pos,cnsn_type,fuse", [("post", "sn", False), ("post", "sn", True), ("post", "cnsn", True),
                                                ("pre", "cnsn", False), ("residual", "cnsn", False), ("identity", "sn", False)])
def test_resnet_matches_reference_model(fake, pos, cnsn_type, fuse, capsys):
    """ResNet bottleneck host vs models/imagenet/resnet_cnsn.py: same seed -> identical state dict; same inputs and
    host RNG -> same logits and gradients, with and without the fused 'post' tail."""
    from cnsn_b200.hosts import ResNet
    RefResNet = _reference_host("models.imagenet.resnet_cnsn", "ResNet")
    kw = dict(num_classes=7, active_num=1, pos=pos, beta=1, crop="both", cnsn_type=cnsn_type)
    torch.manual_seed(0)
    a = RefResNet([1, 1, 1, 1], **kw).double().train()
    torch.manual_seed(0)
    b = ResNet([1, 1, 1, 1], fuse_post=fuse, **kw).double().train()
    capsys.readouterr()
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa) == list(sb)
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k
    x = torch.randn(4, 3, 64, 64, dtype=torch.float64)
    outs = []
    for net in (a, b):
        torch.manual_seed(5)
        np.random.seed(6)
        o = net(x, aug="cn" in cnsn_type)
        o.square().sum().backward()
        outs.append(o)
    assert torch.allclose(outs[0], outs[1], atol=1e-8)
    for (ka, pa), (kb, pb) in zip(a.named_parameters(), b.named_parameters()):
        assert ka == kb and torch.allclose(pa.grad, pb.grad, atol=1e-7, rtol=1e-6), ka
    if fuse:
        assert "selfnorm_block_fwd" in fake.calls
    a.eval(), b.eval()
    assert torch.allclose(a(x), b(x), atol=1e-8)


@pytest.mark.parametrize("pos", ["post", "pre"])
def test_resnet_ibn_a_matches_reference_model(fake, pos, capsys):
    """ResNet-IBN-a host vs models/imagenet/resnet_ibn_cnsn.py: identical state dict for equal seeds, identical
    logits and gradients (the IBN layers run through cnsn_ibn_fwd/_bwd, here the oracle-backed stand-in)."""
    from cnsn_b200.hosts import ResNet
    RefResNet = _reference_host("models.imagenet.resnet_ibn_cnsn", "ResNet")
    kw = dict(num_classes=7, active_num=1, pos=pos, beta=1, crop="both", cnsn_type="cnsn")
    torch.manual_seed(0)
    a = RefResNet(layers=[1, 1, 1, 1], ibn_cfg=("a", "a", "a", None), **kw).double().train()
    torch.manual_seed(0)
    b = ResNet([1, 1, 1, 1], ibn_cfg=("a", "a", "a", None), **kw).double().train()
    capsys.readouterr()
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa) == list(sb)
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k
    x = torch.randn(4, 3, 224, 224, dtype=torch.float64)[:, :, :224:1, :224:1]
    outs = []
    for net in (a, b):
        torch.manual_seed(5)
        np.random.seed(6)
        o = net(x, aug=True)
        o.square().sum().backward()
        outs.append(o)
    assert "ibn_fwd" in fake.calls and "ibn_bwd" in fake.calls
    assert torch.allclose(outs[0], outs[1], atol=1e-8)
    for (ka, pa), (kb, pb) in zip(a.named_parameters(), b.named_parameters()):
        assert ka == kb and torch.allclose(pa.grad, pb.grad, atol=1e-7, rtol=1e-6), ka


@pytest.mark.parametrize("pos,fuse", [("post", False), ("post", True), ("residual", False)])
def test_resnet_ibn_b_matches_reference_model(fake, pos, fuse, capsys):
    """ResNet-IBN-b wiring (instance norm in the stem and after the residual add of a stage's last block,
    models/imagenet/resnet_ibn_cnsn.py:62,122-123,143-144,204-214) vs the reference file: identical state dict for
    equal seeds, identical logits and gradients; the instance norms run through cnsn_ibn_* with half = C."""
    from cnsn_b200.hosts import ResNet
    from cnsn_b200.ibn import InstanceNorm2d
    RefResNet = _reference_host("models.imagenet.resnet_ibn_cnsn", "ResNet")
    kw = dict(num_classes=7, active_num=1, pos=pos, beta=1, crop="neither", cnsn_type="cnsn")
    torch.manual_seed(0)
    a = RefResNet(layers=[2, 2, 1, 1], ibn_cfg=("b", "b", None, None), **kw).double().train()
    torch.manual_seed(0)
    b = ResNet([2, 2, 1, 1], ibn_cfg=("b", "b", None, None), fuse_post=fuse, **kw).double().train()
    capsys.readouterr()
    assert isinstance(b.bn1, InstanceNorm2d) and b.layer1[0].IN is None and isinstance(b.layer1[1].IN, InstanceNorm2d)
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa) == list(sb)
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k
    x = torch.randn(2, 3, 224, 224, dtype=torch.float64)      # the reference head is AvgPool2d(7): 224 x 224 input
    outs = []
    for net in (a, b):
        torch.manual_seed(5)
        np.random.seed(6)
        o = net(x, aug=True)
        o.square().sum().backward()
        outs.append(o)
    assert fake.calls.count("ibn_fwd") == 3 and fake.calls.count("ibn_bwd") == 3     # stem + two stage tails
    assert torch.allclose(outs[0], outs[1], atol=1e-8)
    for (ka, pa), (kb, pb) in zip(a.named_parameters(), b.named_parameters()):
        assert ka == kb and torch.allclose(pa.grad, pb.grad, atol=1e-7, rtol=1e-6), ka


@pytest.mark.parametrize("pos", ["post", "pre", "residual", "identity"])
def test_resnext_matches_reference_model(fake, pos, capsys):
    """CIFAR ResNeXt host vs models/cifar/resnext_cnsn.py: identical state dict for equal seeds; same inputs and host
    RNG -> same logits and gradients (including the reference's post-ReLU 'post' site and the 'identity' quirk)."""
    from cnsn_b200.hosts import CifarResNeXt
    Ref = _reference_host("models.cifar.resnext_cnsn", "CifarResNeXt")
    kw = dict(depth=20, cardinality=2, base_width=16, num_classes=10, active_num=2, pos=pos, beta=1, crop="both",
              cnsn_type="cnsn")
    torch.manual_seed(0)
    a = Ref(**kw).double().train()
    torch.manual_seed(0)
    b = CifarResNeXt(**kw).double().train()
    capsys.readouterr()
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa) == list(sb)
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k
    x = torch.randn(4, 3, 32, 32, dtype=torch.float64)
    outs = []
    for net in (a, b):
        torch.manual_seed(5)
        np.random.seed(6)
        o = net(x, aug=True)
        o.square().sum().backward()
        outs.append(o)
    assert fake.calls.count("site_fwd") == 2 and "selfnorm_bwd" in fake.calls
    assert torch.allclose(outs[0], outs[1], atol=1e-8)
    for (ka, pa), (kb, pb) in zip(a.named_parameters(), b.named_parameters()):
        assert ka == kb and (pa.grad is None) == (pb.grad is None), ka     # 'identity': a projection block drops its site
        if pa.grad is not None:
            assert torch.allclose(pa.grad, pb.grad, atol=1e-7, rtol=1e-6), ka


def test_resnet50_census():
    """ResNet-50 + SN ('post'): 16 SelfNorm sites with the channel counts of SURVEY.md 8 (cfg4)."""
    from cnsn_b200.hosts import resnet50
    import cnsn_b200.cnsn as m
    net = resnet50()
    widths = [k.g_bn.num_features for k in net.modules() if isinstance(k, m.SelfNorm)]
    assert widths == [256] * 3 + [512] * 4 + [1024] * 6 + [2048] * 3
    assert "layer1.0.cnsn.selfnorm.g_fc.weight" in net.state_dict() and "layer2.0.downsample.1.weight" in net.state_dict()


@pytest.mark.parametrize("pos", ["post", "pre", "residual", "identity"])
def test_wideresnet_matches_reference_model(fake, pos, capsys):
    """Same seed -> identical parameters; same inputs and host RNG -> same logits and gradients."""
    from cnsn_b200.hosts import WideResNet
    RefWRN = _reference_wrn()
    kw = dict(depth=10, num_classes=10, widen_factor=2, active_num=2, pos=pos, beta=1, crop="both", cnsn_type="cnsn")
    torch.manual_seed(0)
    a = RefWRN(**kw).double().train()
    torch.manual_seed(0)
    b = WideResNet(**kw).double().train()
    capsys.readouterr()
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa) == list(sb)
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k
    assert len(a.cn_modules) == len(b.cn_modules) == 3
    x = torch.randn(6, 3, 32, 32, dtype=torch.float64)
    outs = []
    for net in (a, b):
        torch.manual_seed(5)
        np.random.seed(6)
        o = net(x, aug=True)
        o.square().sum().backward()
        outs.append(o)
    assert torch.allclose(outs[0], outs[1], atol=1e-8)
    for (ka, pa), (kb, pb) in zip(a.named_parameters(), b.named_parameters()):
        assert ka == kb and torch.allclose(pa.grad, pb.grad, atol=1e-7, rtol=1e-6), ka
    a.eval(), b.eval()
    assert torch.allclose(a(x), b(x), atol=1e-8)


def test_wrn40_2_census():
    """WRN-40-2 cnsn/post: 2,248,922 parameters, 18 CrossNorm + 18 SelfNorm sites."""
    from cnsn_b200.train import wrn40_2
    import cnsn_b200.cnsn as m
    net = wrn40_2()
    assert sum(p.numel() for p in net.parameters()) == 2248922
    assert len(net.cn_modules) == 18
    assert sum(isinstance(k, m.SelfNorm) for k in net.modules()) == 18
    assert "block1.layer.0.cnsn.selfnorm.g_fc.weight" in net.state_dict()


def test_train_step_on_cpu_with_eager_ops():
    """The harness runs end to end on CPU when given the eager-PyTorch operator set (the CPU baseline arm)."""
    from cnsn_b200.train import bench_wrn
    from oracle import eager_modules
    r = bench_wrn(torch.device("cpu"), 1, 0, batch=8, steps=2, warmup=1, cn_prob=1.0, ops=eager_modules)
    assert r["value"] > 0 and np.isfinite(r["final_loss"]) and r["params"] == 2248922


def test_reference_checkpoint_round_trip(fake, tmp_path, capsys):
    """A checkpoint written the way the reference writes them (DataParallel 'module.' keys inside a training dict, or a
    bare state dict) loads into the host model; values identical, nothing missing, extras returned."""
    from cnsn_b200.hosts import ResNet
    from cnsn_b200.utils import load_reference_checkpoint
    RefResNet = _reference_host("models.imagenet.resnet_cnsn", "ResNet")
    kw = dict(num_classes=5, active_num=1, pos="post", beta=1, crop="both", cnsn_type="sn")
    torch.manual_seed(3)
    ref = RefResNet([1, 1, 1, 1], **kw)
    capsys.readouterr()
    wrapped = {"module." + k: v for k, v in ref.state_dict().items()}
    path = tmp_path / "ckpt.pth.tar"
    torch.save({"epoch": 7, "best_err1": 23.4, "state_dict": wrapped}, path)
    torch.manual_seed(9)
    net = ResNet([1, 1, 1, 1], **kw)
    missing, unexpected, extras = load_reference_checkpoint(net, str(path))
    assert missing == [] and unexpected == [] and extras["epoch"] == 7
    for k, v in ref.state_dict().items():
        assert torch.equal(net.state_dict()[k], v), k
    # bare state dict with an extra head (strict=False, as imagenet.py:518-521)
    bare = dict(ref.state_dict(), **{"aux.weight": torch.zeros(1)})
    missing, unexpected, _ = load_reference_checkpoint(net, bare)
    assert missing == [] and unexpected == ["aux.weight"]


def test_bottleneck_tail_helper_falls_back_to_the_two_calls_on_cpu(fake):
    """hosts._norm.bn_site_relu is relu(cnsn(bn(c) + skip)); without a CUDA channels_last tensor it IS the two calls it
    replaces (the fused operator is a GPU path; here the operators run on the oracle-backed stand-in), and
    hosts._norm.MaxPool2d is nn.MaxPool2d."""
    import torch
    import torch.nn as nn
    import cnsn_b200.cnsn as M
    from cnsn_b200.hosts._norm import MaxPool2d, bn_site_relu
    from cnsn_b200.ibn import BatchNorm2d
    torch.manual_seed(0)
    bn = BatchNorm2d(8).train()
    site = M.CNSN(None, M.SelfNorm(8)).train()
    c, skip = torch.randn(4, 8, 6, 6), torch.randn(4, 8, 6, 6)
    import copy
    bn2, site2 = copy.deepcopy(bn), copy.deepcopy(site)
    a = bn_site_relu(bn, site, c, skip)
    b = torch.relu(site2(bn2(c) + skip))
    assert torch.allclose(a, b, atol=1e-6) and torch.equal(bn.running_mean, bn2.running_mean)
    x = torch.randn(2, 4, 9, 9)
    assert torch.equal(MaxPool2d(3, 2, 1)(x), nn.MaxPool2d(3, 2, 1)(x))
    assert torch.equal(MaxPool2d(3, 2, 1)(x.contiguous(memory_format=torch.channels_last)), nn.MaxPool2d(3, 2, 1)(x))
