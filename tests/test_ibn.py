"""IBN layer: oracle vs the reference's module composition executed with PyTorch (CPU), kernel vs oracle (GPU),
state-dict compatibility with the reference class."""
import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import ibn_oracle as B


class ReferenceIBN(nn.Module):
    """The composition of models/imagenet/resnet_ibn_cnsn.py:24-44 (behavioural restatement for the pin)."""

    def __init__(self, planes, ratio=0.5):
        super().__init__()
        self.half = int(planes * ratio)
        self.IN = nn.InstanceNorm2d(self.half, affine=True)
        self.BN = nn.BatchNorm2d(planes - self.half)

    def forward(self, x):
        a, b = torch.split(x, self.half, 1)
        return torch.cat((self.IN(a.contiguous()), self.BN(b.contiguous())), 1)


def _case(shape, seed):
    rs = np.random.RandomState(seed)
    N, C, H, W = shape
    half = C // 2
    x = rs.standard_normal(shape) * (0.5 + rs.rand(N, C, 1, 1)) + rs.standard_normal((N, C, 1, 1))
    dy = rs.standard_normal(shape)
    p = {"in_w": rs.uniform(0.5, 1.5, half), "in_b": rs.uniform(-0.5, 0.5, half),
         "bn_w": rs.uniform(0.5, 1.5, C - half), "bn_b": rs.uniform(-0.5, 0.5, C - half)}
    bufs = {"rm": rs.uniform(-1, 1, C - half), "rv": rs.uniform(0.5, 2, C - half)}
    return x, dy, half, p, bufs


def O_varied(shape, seed):
    rs = np.random.RandomState(seed)
    N, C = shape[:2]
    return rs.standard_normal(shape) * (0.5 + rs.rand(N, C, 1, 1)) + rs.standard_normal((N, C, 1, 1))


def _load(m, p, bufs):
    with torch.no_grad():
        m.IN.weight.copy_(torch.from_numpy(p["in_w"])); m.IN.bias.copy_(torch.from_numpy(p["in_b"]))
        m.BN.weight.copy_(torch.from_numpy(p["bn_w"])); m.BN.bias.copy_(torch.from_numpy(p["bn_b"]))
        m.BN.running_mean.copy_(torch.from_numpy(bufs["rm"])); m.BN.running_var.copy_(torch.from_numpy(bufs["rv"]))
    return m


@pytest.mark.parametrize("shape", [(4, 6, 5, 5), (3, 8, 4, 6), (2, 2, 3, 3)])
@pytest.mark.parametrize("training", [True, False])
def test_oracle_matches_reference_composition(shape, training):
    x, dy, half, p, bufs = _case(shape, sum(shape))
    m = _load(ReferenceIBN(shape[1]).double(), p, bufs).train(training)
    xt = torch.from_numpy(x).requires_grad_(True)
    y = m(xt)
    y.backward(torch.from_numpy(dy))
    yo, rm, rv = B.ibn_fwd(x, half, p, bufs, training)
    dxo, giw, gib, gbw, gbb = B.ibn_bwd(x, dy, half, p, bufs, training)
    assert np.allclose(y.detach().numpy(), yo, atol=1e-11)
    assert np.allclose(xt.grad.numpy(), dxo, atol=1e-10)
    for a, b in ((m.IN.weight.grad, giw), (m.IN.bias.grad, gib), (m.BN.weight.grad, gbw), (m.BN.bias.grad, gbb)):
        assert np.allclose(a.numpy(), b, atol=1e-9)
    assert np.allclose(m.BN.running_mean.numpy(), rm, atol=1e-12) and np.allclose(m.BN.running_var.numpy(), rv, atol=1e-12)


def test_ibn_state_dict_matches_reference_class():
    from cnsn_b200.ibn import IBN
    torch.manual_seed(0)
    a = ReferenceIBN(64)
    torch.manual_seed(0)
    b = IBN(64)
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa) == list(sb) and all(torch.equal(sa[k], sb[k]) for k in sa)


@pytest.mark.gpu
@pytest.mark.parametrize("shape,dtype", [((8, 16, 8, 8), torch.float32), ((64, 32, 28, 28), torch.float32),
                                         ((256, 8, 56, 56), torch.float32), ((33, 10, 12, 12), torch.float32),
                                         ((600, 6, 4, 4), torch.float32), ((64, 16, 16, 16), torch.bfloat16)])
@pytest.mark.parametrize("training", [True, False])
def test_ibn_kernel_vs_oracle(shape, dtype, training):
    from cnsn_b200.ibn import IBN
    x, dy, half, p, bufs = _case(shape, sum(shape) + 1)
    if dtype != torch.float32:
        x, dy = (torch.from_numpy(v).to(dtype).double().numpy() for v in (x, dy))
    m = _load(IBN(shape[1]), {k: v.astype(np.float32) for k, v in p.items()}, {k: v.astype(np.float32) for k, v in bufs.items()})
    m = m.cuda().train(training)
    xt = torch.from_numpy(x).to(device="cuda", dtype=dtype).requires_grad_(True)
    y = m(xt)
    y.backward(torch.from_numpy(dy).to(device="cuda", dtype=dtype))
    p32 = {k: v.astype(np.float32).astype(np.float64) for k, v in p.items()}
    b32 = {k: v.astype(np.float32).astype(np.float64) for k, v in bufs.items()}
    yo, rm, rv = B.ibn_fwd(x, half, p32, b32, training)
    dxo, giw, gib, gbw, gbb = B.ibn_bwd(x, dy, half, p32, b32, training)
    atol, rtol = (1e-5, 1e-5) if dtype == torch.float32 else (2e-2, 1e-2)
    assert np.allclose(y.detach().double().cpu().numpy(), yo, atol=atol, rtol=rtol)
    assert np.allclose(xt.grad.double().cpu().numpy(), dxo, atol=atol, rtol=rtol)
    ptol = 1e-5 if dtype == torch.float32 else 1e-3
    for a, b in ((m.IN.weight.grad, giw), (m.IN.bias.grad, gib), (m.BN.weight.grad, gbw), (m.BN.bias.grad, gbb)):
        assert np.abs(a.double().cpu().numpy() - b).max() <= ptol * max(np.abs(b).max(), 1e-6)
    assert np.allclose(m.BN.running_mean.double().cpu().numpy(), rm, atol=1e-5)
    assert np.allclose(m.BN.running_var.double().cpu().numpy(), rv, atol=1e-5, rtol=1e-5)
    assert int(m.BN.num_batches_tracked) == (1 if training else 0)


def test_instance_norm_state_dict_and_cpu_semantics():
    """cnsn_b200.ibn.InstanceNorm2d is an nn.InstanceNorm2d(C, affine=True) (isinstance, state dict); the oracle with
    half = C equals torch's module (what the reference's IBN-b blocks call, resnet_ibn_cnsn.py:62,122-123)."""
    from cnsn_b200.ibn import InstanceNorm2d
    m = InstanceNorm2d(6, affine=True)
    r = torch.nn.InstanceNorm2d(6, affine=True)
    assert isinstance(m, torch.nn.InstanceNorm2d) and list(m.state_dict()) == list(r.state_dict()) == ["weight", "bias"]
    rs = np.random.RandomState(0)
    x, dy = rs.standard_normal((3, 6, 5, 4)), rs.standard_normal((3, 6, 5, 4))
    p = {"in_w": rs.uniform(0.5, 1.5, 6), "in_b": rs.uniform(-0.5, 0.5, 6), "bn_w": np.zeros(0), "bn_b": np.zeros(0)}
    bufs = {"rm": np.zeros(0), "rv": np.zeros(0)}
    r = r.double()
    with torch.no_grad():
        r.weight.copy_(torch.from_numpy(p["in_w"]))
        r.bias.copy_(torch.from_numpy(p["in_b"]))
    xt = torch.from_numpy(x).requires_grad_(True)
    y = r(xt)
    y.backward(torch.from_numpy(dy))
    yo, _, _ = B.ibn_fwd(x, 6, p, bufs, True)
    dxo, giw, gib, _, _ = B.ibn_bwd(x, dy, 6, p, bufs, True)
    assert np.abs(y.detach().numpy() - yo).max() < 1e-12 and np.abs(xt.grad.numpy() - dxo).max() < 1e-12
    assert np.abs(r.weight.grad.numpy() - giw).max() < 1e-10 and np.abs(r.bias.grad.numpy() - gib).max() < 1e-10


@pytest.mark.gpu
@pytest.mark.parametrize("shape,dtype", [((8, 16, 8, 8), torch.float32), ((32, 64, 28, 28), torch.float32),
                                         ((16, 8, 56, 56), torch.float32), ((33, 10, 12, 12), torch.float32),
                                         ((32, 16, 16, 16), torch.bfloat16),
                                         # the IBN-b stem plane (112 x 112) with more samples than CTAs fit the GPU at
                                         # two planes per item: instance-norm channels need no co-residency
                                         ((300, 2, 112, 112), torch.float32)])
@pytest.mark.parametrize("training", [True, False])
def test_instance_norm_kernel_vs_oracle(shape, dtype, training):
    """IBN-b's instance norm: cnsn_ibn_fwd/_bwd with half = C (no batch-norm half) through cnsn_b200.ibn.InstanceNorm2d."""
    import cnsn_b200
    from cnsn_b200.ibn import InstanceNorm2d
    C = shape[1]
    rs = np.random.RandomState(sum(shape))
    x = O_varied(shape, sum(shape) + 2)
    dy = rs.standard_normal(shape)
    if dtype != torch.float32:
        x, dy = (torch.from_numpy(v).to(dtype).double().numpy() for v in (x, dy))
    p = {"in_w": rs.uniform(0.5, 1.5, C).astype(np.float32).astype(np.float64),
         "in_b": rs.uniform(-0.5, 0.5, C).astype(np.float32).astype(np.float64), "bn_w": np.zeros(0), "bn_b": np.zeros(0)}
    bufs = {"rm": np.zeros(0), "rv": np.zeros(0)}
    m = InstanceNorm2d(C, affine=True)
    with torch.no_grad():
        m.weight.copy_(torch.from_numpy(p["in_w"]))
        m.bias.copy_(torch.from_numpy(p["in_b"]))
    m = m.cuda().train(training)
    xt = torch.from_numpy(x).to(device="cuda", dtype=dtype).requires_grad_(True)
    n0 = cnsn_b200.launch_count()
    y = m(xt)
    y.backward(torch.from_numpy(dy).to(device="cuda", dtype=dtype))
    torch.cuda.synchronize()
    assert cnsn_b200.launch_count() - n0 == 2
    yo, _, _ = B.ibn_fwd(x, C, p, bufs, training)
    dxo, giw, gib, _, _ = B.ibn_bwd(x, dy, C, p, bufs, training)
    atol, rtol = (1e-5, 1e-5) if dtype == torch.float32 else (2e-2, 1e-2)
    assert np.allclose(y.detach().double().cpu().numpy(), yo, atol=atol, rtol=rtol)
    assert np.allclose(xt.grad.double().cpu().numpy(), dxo, atol=atol, rtol=rtol)
    ptol = 1e-5 if dtype == torch.float32 else 1e-3
    for a, b in ((m.weight.grad, giw), (m.bias.grad, gib)):
        assert np.abs(a.double().cpu().numpy() - b).max() <= ptol * max(np.abs(b).max(), 1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("shape,dtype,half", [((6, 8, 7, 7), torch.float32, 4), ((16, 8, 14, 14), torch.bfloat16, 4),
                                              ((5, 6, 9, 14), torch.float32, 3), ((3, 4, 224, 224), torch.float32, 2),
                                              ((6, 8, 7, 7), torch.float32, 8), ((33, 5, 7, 7), torch.float16, 2)])
@pytest.mark.parametrize("training", [True, False])
def test_ibn_general_path_vs_oracle(shape, dtype, half, training):
    """Shapes outside the resident kernel -- planes that are not 16-byte multiples (7x7, 14x14 bf16, 9x14), planes
    too large for shared memory (224x224) -- take the three-kernel path (ibn_general.cu): same results, three
    launches per direction, no CNSN_E_UNSUPPORTED."""
    import cnsn_b200
    from cnsn_b200 import _lib as L
    N, C, H, W = shape
    rs = np.random.RandomState(sum(shape) + half)
    x = O_varied(shape, sum(shape) + 3)
    dy = rs.standard_normal(shape)
    if dtype != torch.float32:
        x, dy = (torch.from_numpy(v).to(dtype).double().numpy() for v in (x, dy))

    def f32(v):
        return v.astype(np.float32).astype(np.float64)

    p = {"in_w": f32(rs.uniform(0.5, 1.5, half)), "in_b": f32(rs.uniform(-0.5, 0.5, half)),
         "bn_w": f32(rs.uniform(0.5, 1.5, C - half)), "bn_b": f32(rs.uniform(-0.5, 0.5, C - half))}
    bufs = {"rm": f32(rs.uniform(-1, 1, C - half)), "rv": f32(rs.uniform(0.5, 2, C - half))}
    dev = "cuda:0"

    def t(v, dt=torch.float32):
        return torch.from_numpy(v).to(device=dev, dtype=dt)

    pt = {k: t(v) for k, v in p.items()}
    pt.update(run_mean=t(bufs["rm"]), run_var=t(bufs["rv"]), nbt=torch.zeros((), dtype=torch.int64, device=dev))
    if half == C:
        pt.update(bn_w=None, bn_b=None, run_mean=None, run_var=None, nbt=None)
    xt, dyt = t(x, dtype), t(dy, dtype)
    be = L.backend()
    n0 = cnsn_b200.launch_count()
    y, save = be.ibn_fwd(xt, half, pt, training, 0.1, 1e-5, 1e-5)
    dx, g = be.ibn_bwd(xt, dyt, half, pt, training, save)
    torch.cuda.synchronize()
    assert cnsn_b200.launch_count() - n0 == 6
    yo, rm, rv = B.ibn_fwd(x, half, p, bufs, training)
    dxo, giw, gib, gbw, gbb = B.ibn_bwd(x, dy, half, p, bufs, training)
    atol, rtol = (1e-5, 1e-5) if dtype == torch.float32 else (2e-2, 1e-2)
    assert np.allclose(y.double().cpu().numpy(), yo, atol=atol, rtol=rtol)
    assert np.allclose(dx.double().cpu().numpy(), dxo, atol=atol, rtol=rtol)
    ptol = 1e-5 if dtype == torch.float32 else 1e-3
    for a, b in zip(g, (giw, gib, gbw, gbb)):
        if b.size:
            assert np.abs(a.double().cpu().numpy() - b).max() <= ptol * max(np.abs(b).max(), 1e-6)
    if half < C:
        assert np.allclose(pt["run_mean"].double().cpu().numpy(), rm, atol=1e-5)
        assert np.allclose(pt["run_var"].double().cpu().numpy(), rv, atol=1e-5, rtol=1e-5)
        assert int(pt["nbt"]) == (1 if training else 0)


# ------------------------------------------------------------------ BatchNorm2d drop-in (half == 0)
@pytest.mark.gpu
@pytest.mark.parametrize("shape,dtype", [((64, 16, 32, 32), torch.float32), ((128, 32, 32, 32), torch.float32), ((96, 64, 16, 16), torch.float32),
                                         ((64, 128, 8, 8), torch.float32), ((16, 64, 56, 56), torch.float32), ((32, 8, 32, 32), torch.bfloat16),
                                         ((8, 16, 7, 7), torch.float32), ((96, 64, 14, 14), torch.bfloat16), ((96, 128, 7, 7), torch.bfloat16),
                                         ((300, 8, 7, 7), torch.float32), ((33, 24, 9, 9), torch.float16), ((5, 6, 7, 7), torch.float32),
                                         ((64, 32, 14, 14), torch.float16)])
@pytest.mark.parametrize("training", [True, False])
def test_batchnorm2d_dropin_matches_torch(shape, dtype, training):
    """cnsn_b200.ibn.BatchNorm2d (the host blocks' nn.BatchNorm2d through cnsn_ibn_* with half = 0) against torch's own
    nn.BatchNorm2d in fp64 on the same GPU: y, dx, dweight, dbias, running statistics, num_batches_tracked; identical
    state dict.  Planes that are not 16-byte multiples (7x7 fp32, 14x14 / 7x7 / 9x9 16-bit: the last stages of ResNet-50
    under autocast) run the channel-group kernel of csrc/bn_grp.cu; (5,6,7,7): no channel group fits -- torch.)"""
    import torch.nn as nn
    from cnsn_b200.ibn import BatchNorm2d
    dev = "cuda:0"
    g = torch.Generator().manual_seed(0)
    C = shape[1]
    x0 = (torch.randn(shape, generator=g) * (0.5 + torch.rand(1, C, 1, 1, generator=g)) + torch.randn(1, C, 1, 1, generator=g)).to(dtype)
    dy0 = torch.randn(shape, generator=g).to(dtype)
    ref = nn.BatchNorm2d(C).to(dev).double().train(training)
    ours = BatchNorm2d(C).to(dev).train(training)
    with torch.no_grad():
        ref.weight.copy_(torch.rand(C, generator=g) + 0.5)
        ref.bias.copy_(torch.randn(C, generator=g))
        ref.running_mean.copy_(torch.randn(C, generator=g) * 0.1)
        ref.running_var.copy_(torch.rand(C, generator=g) + 0.5)
    ours.load_state_dict({k: (v.float() if v.dtype.is_floating_point else v) for k, v in ref.state_dict().items()})
    assert list(ours.state_dict()) == list(ref.state_dict())
    res = []
    for m, dt in ((ref, torch.float64), (ours, dtype)):
        x = x0.to(dev).to(dt).requires_grad_(True)
        y = m(x)
        y.backward(dy0.to(dev).to(dt))
        res.append((y.detach().double(), x.grad.double(), m.weight.grad.double(), m.bias.grad.double(),
                    m.running_mean.double(), m.running_var.double()))
        assert int(m.num_batches_tracked) == (1 if training else 0)
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    for name, a, b in zip(("y", "dx", "dweight", "dbias", "running_mean", "running_var"), res[0], res[1]):
        err = float((a - b).abs().max() / a.abs().max().clamp_min(1e-12))
        assert err <= tol, (name, err)


@pytest.mark.gpu
def test_batchnorm2d_odd_planes_take_the_group_kernel():
    """The shapes of ResNet-50's last two stages under autocast are claimed by the library (cnsn_ibn_resident), so the
    drop-in runs csrc/bn_grp.cu for them rather than torch's batch norm; shapes with no 16-byte channel group are not."""
    import cnsn_b200._lib as L
    be = L.backend()
    dev = "cuda:0"
    for shape, dtype, want in (((96, 1024, 14, 14), torch.bfloat16, True), ((96, 2048, 7, 7), torch.bfloat16, True),
                               ((96, 512, 7, 7), torch.bfloat16, True), ((32, 512, 7, 7), torch.float32, True),
                               ((5, 6, 7, 7), torch.float32, False), ((4, 7, 7, 7), torch.bfloat16, False)):
        x = torch.empty(shape, device=dev, dtype=dtype)
        for training in (True, False):
            assert bool(be.ibn_resident(x, 0, training)) == want, (shape, dtype, training)
    n0 = L.launch_count()
    from cnsn_b200.ibn import BatchNorm2d
    m = BatchNorm2d(64).to(dev).train()
    x = torch.randn(48, 64, 7, 7, device=dev, dtype=torch.bfloat16, requires_grad=True)
    m(x, True).float().sum().backward()
    torch.cuda.synchronize()
    assert L.launch_count() - n0 >= 2
    L.async_error()


@pytest.mark.gpu
def test_batchnorm2d_dropin_momentum_none_and_fallbacks():
    """momentum=None (cumulative average) follows torch; affine=False and CPU tensors take the torch implementation."""
    import torch.nn as nn
    from cnsn_b200.ibn import BatchNorm2d
    dev = "cuda:0"
    x = torch.randn(32, 8, 16, 16, device=dev)
    a, b = nn.BatchNorm2d(8, momentum=None).to(dev).train(), BatchNorm2d(8, momentum=None).to(dev).train()
    for _ in range(3):
        ya, yb = a(x), b(x)
        x = x * 1.1 + 0.05
    assert torch.allclose(ya, yb, atol=1e-5) and torch.allclose(a.running_mean, b.running_mean, atol=1e-6)
    assert torch.allclose(a.running_var, b.running_var, atol=1e-5) and int(b.num_batches_tracked) == 3
    c = BatchNorm2d(8, affine=False).to(dev).train()
    assert torch.allclose(c(x), nn.functional.batch_norm(x, None, None, training=True), atol=1e-5)
    d = BatchNorm2d(8).train()
    assert d(torch.randn(4, 8, 5, 5)).shape == (4, 8, 5, 5)


@pytest.mark.gpu
@pytest.mark.parametrize("shape,dtype", [((64, 16, 32, 32), torch.float32), ((96, 64, 16, 16), torch.float32), ((16, 64, 56, 56), torch.float32),
                                         ((32, 8, 32, 32), torch.bfloat16), ((96, 64, 14, 14), torch.bfloat16), ((64, 32, 7, 7), torch.float32),
                                         ((200, 16, 7, 7), torch.bfloat16)])
@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("binding", ["ext", "ctypes"])
def test_batchnorm2d_fused_relu_matches_torch(shape, dtype, training, binding):
    """relu(bn(x)) in one kernel per direction (BatchNorm2d.forward(x, relu=True): the pair the host blocks apply) against
    torch's batch norm + relu in fp64: y, dx, dweight, dbias, running statistics; through both host bindings."""
    import torch.nn as nn
    import cnsn_b200._lib as L
    from cnsn_b200.ibn import BatchNorm2d
    dev = "cuda:0"
    g = torch.Generator().manual_seed(0)
    C = shape[1]
    x0 = (torch.randn(shape, generator=g) * (0.5 + torch.rand(1, C, 1, 1, generator=g)) + 0.3 * torch.randn(1, C, 1, 1, generator=g)).to(dtype)
    dy0 = torch.randn(shape, generator=g).to(dtype)
    ref = nn.BatchNorm2d(C).to(dev).double().train(training)
    ours = BatchNorm2d(C).to(dev).train(training)
    with torch.no_grad():
        ref.weight.copy_(torch.rand(C, generator=g) + 0.5)
        ref.bias.copy_(torch.randn(C, generator=g) * 0.5)
        ref.running_mean.copy_(torch.randn(C, generator=g) * 0.1)
        ref.running_var.copy_(torch.rand(C, generator=g) + 0.5)
    ours.load_state_dict({k: (v.float() if v.dtype.is_floating_point else v) for k, v in ref.state_dict().items()})
    old = L.set_binding(binding)
    try:
        res = []
        for m, dt in ((ref, torch.float64), (ours, dtype)):
            x = x0.to(dev).to(dt).requires_grad_(True)
            y = torch.relu(m(x)) if m is ref else m(x, True)
            y.backward(dy0.to(dev).to(dt))
            res.append((y.detach().double(), x.grad.double(), m.weight.grad.double(), m.bias.grad.double(),
                        m.running_mean.double(), m.running_var.double()))
    finally:
        L.set_binding(old)
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    for name, a, b in zip(("y", "dx", "dweight", "dbias", "running_mean", "running_var"), res[0], res[1]):
        if name == "dx":
            # an output within rounding of 0 may fall on the other side of the ReLU than in fp64 (its dx then differs by
            # dy): compare where the fp64 pre-activation is clear of the edge
            xr = x0.to(dev).double()
            pre = torch.nn.functional.batch_norm(xr, ref.running_mean if not training else None, ref.running_var if not training else None,
                                                 ref.weight, ref.bias, training, 0.0, ref.eps)
            clear = pre.abs() > (1e-4 if dtype == torch.float32 else 0.05)
            a, b = a[clear], b[clear]
        err = float((a - b).abs().max() / a.abs().max().clamp_min(1e-12))
        assert err <= tol, (name, err)


@pytest.mark.gpu
@pytest.mark.parametrize("shape,dtype", [((64, 32, 32, 32), torch.float32), ((32, 64, 16, 16), torch.float32), ((16, 128, 8, 8), torch.float32),
                                         ((8, 16, 20, 20), torch.float32), ((6, 8, 50, 50), torch.float32), ((32, 32, 32, 32), torch.bfloat16),
                                         ((4, 2048, 7, 7), torch.float32), ((3, 64, 13, 11), torch.float16), ((8, 24, 9, 9), torch.float32),
                                         ((16, 256, 56, 56), torch.bfloat16)])
@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("relu", [True, False])
@pytest.mark.parametrize("binding", ["ext", "ctypes"])
def test_batchnorm2d_channels_last_matches_torch(shape, dtype, training, relu, binding):
    """[relu](bn(x)) on torch.channels_last tensors through cnsn_bn_nhwc_fwd / _bwd (csrc/bn_nhwc.cu) against torch's batch
    norm [+ relu] in fp64: y, dx, dweight, dbias, running statistics, num_batches_tracked; outputs stay channels_last.
    (4,2048,7,7): two channel blocks per row; (8,24,9,9): 96-byte pixels -- torch's own kernels take it."""
    import torch.nn as nn
    import cnsn_b200._lib as L
    from cnsn_b200.ibn import BatchNorm2d
    dev = "cuda:0"
    cl = torch.channels_last
    g = torch.Generator().manual_seed(0)
    C = shape[1]
    x0 = (torch.randn(shape, generator=g) * (0.5 + torch.rand(1, C, 1, 1, generator=g)) + 0.3 * torch.randn(1, C, 1, 1, generator=g)).to(dtype)
    dy0 = torch.randn(shape, generator=g).to(dtype)
    ref = nn.BatchNorm2d(C).to(dev).double().train(training)
    ours = BatchNorm2d(C).to(dev).train(training)
    with torch.no_grad():
        ref.weight.copy_(torch.rand(C, generator=g) + 0.5)
        ref.bias.copy_(torch.randn(C, generator=g) * 0.5)
        ref.running_mean.copy_(torch.randn(C, generator=g) * 0.1)
        ref.running_var.copy_(torch.rand(C, generator=g) + 0.5)
    ours.load_state_dict({k: (v.float() if v.dtype.is_floating_point else v) for k, v in ref.state_dict().items()})
    old = L.set_binding(binding)
    try:
        res = []
        for m, dt in ((ref, torch.float64), (ours, dtype)):
            x = x0.to(dev).to(dt).contiguous(memory_format=cl).requires_grad_(True)
            n0 = L.launch_count()
            if m is ref:
                y = torch.relu(m(x)) if relu else m(x)
            else:
                y = m(x, relu)
            y.backward(dy0.to(dev).to(dt).contiguous(memory_format=cl))
            torch.cuda.synchronize()
            launched = L.launch_count() - n0
            res.append((y.detach().double(), x.grad.double(), m.weight.grad.double(), m.bias.grad.double(),
                        m.running_mean.double(), m.running_var.double()))
            assert int(m.num_batches_tracked) == (1 if training else 0)
    finally:
        L.set_binding(old)
    x = x0.to(dev).to(dtype).contiguous(memory_format=cl)
    if shape != (8, 24, 9, 9):
        assert L.backend().bn_nhwc_ok(x)
        assert launched == (6 if training else 5)          # statistics (training only), fold, apply | reduce, fold, apply
        assert y.is_contiguous(memory_format=cl)
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    for name, a, b in zip(("y", "dx", "dweight", "dbias", "running_mean", "running_var"), res[0], res[1]):
        if name == "dx" and relu:
            xr = x0.to(dev).double()
            pre = torch.nn.functional.batch_norm(xr, ref.running_mean if not training else None, ref.running_var if not training else None,
                                                 ref.weight, ref.bias, training, 0.0, ref.eps)
            clear = pre.abs() > (1e-4 if dtype == torch.float32 else 0.05)
            a, b = a[clear], b[clear]
        err = float((a - b).abs().max() / a.abs().max().clamp_min(1e-12))
        assert err <= tol, (name, err)


@pytest.mark.gpu
@pytest.mark.parametrize("shape,dtype,k,stride,pad", [((8, 64, 112, 112), torch.float32, 3, 2, 1), ((6, 16, 17, 23), torch.float32, 3, 2, 1),
                                                      ((4, 8, 9, 9), torch.bfloat16, 3, 2, 1), ((3, 32, 12, 12), torch.float16, 2, 2, 0),
                                                      ((2, 4, 10, 7), torch.float32, 3, 1, 1), ((2, 8, 15, 15), torch.float32, 5, 3, 2),
                                                      ((16, 64, 56, 56), torch.bfloat16, 3, 2, 1)])
@pytest.mark.parametrize("binding", ["ext", "ctypes"])
def test_maxpool_channels_last_matches_torch(shape, dtype, k, stride, pad, binding):
    """hosts._norm.MaxPool2d on torch.channels_last inputs (csrc/pool_nhwc.cu) against torch's max_pool2d: outputs bit-equal;
    input gradients bit-equal in fp32 (the same windows win -- the input is ReLU'd, so whole windows tie at 0 and the tie
    rule decides -- and their contributions are added in the same order), 16-bit: against torch in fp32 within one rounding."""
    import torch.nn as nn
    import cnsn_b200._lib as L
    from cnsn_b200.hosts._norm import MaxPool2d
    dev = "cuda:0"
    cl = torch.channels_last
    g = torch.Generator().manual_seed(sum(shape))
    x0 = torch.relu(torch.randn(shape, generator=g)).to(dtype)
    x0[0, :, 0, 0] = float("nan")                           # NaN propagates like torch's
    ours, ref = MaxPool2d(k, stride, pad), nn.MaxPool2d(k, stride, pad)
    old = L.set_binding(binding)
    try:
        x = x0.to(dev).contiguous(memory_format=cl).requires_grad_(True)
        n0 = L.launch_count()
        y = ours(x)
        dy0 = torch.randn(y.shape, generator=g).to(dtype)
        y.backward(dy0.to(dev).contiguous(memory_format=cl))
        torch.cuda.synchronize()
        assert L.launch_count() - n0 == 2 and y.is_contiguous(memory_format=cl) and x.grad.is_contiguous(memory_format=cl)
    finally:
        L.set_binding(old)
    xr = x0.to(dev).float().requires_grad_(True)
    yr = ref(xr)
    yr.backward(dy0.to(dev).float())
    assert torch.equal(torch.nan_to_num(y.float(), nan=-1.0), torch.nan_to_num(yr, nan=-1.0))
    if dtype == torch.float32:
        assert torch.equal(x.grad, xr.grad)
    else:
        assert torch.allclose(x.grad.float(), xr.grad, rtol=1e-2, atol=1e-2)
    # NCHW input: torch's own kernel
    assert torch.equal(torch.nan_to_num(ours(x0.to(dev)), nan=-1.0), torch.nan_to_num(ref(x0.to(dev)), nan=-1.0))


@pytest.mark.gpu
@pytest.mark.parametrize("shape,dtype", [((8, 64, 14, 14), torch.float32), ((6, 256, 8, 8), torch.float32), ((4, 2048, 7, 7), torch.float32),
                                         ((5, 16, 50, 50), torch.float32), ((16, 32, 20, 20), torch.bfloat16), ((3, 64, 13, 11), torch.float16),
                                         ((32, 256, 56, 56), torch.bfloat16)])
@pytest.mark.parametrize("training", [True, False])
def test_bottleneck_tail_fused_equals_the_two_operators(shape, dtype, training):
    """relu(cnsn(bn3(c) + skip)) as ONE operator (hosts._norm.bn_site_relu -> cnsn_bn_selfnorm_tail_*_nhwc) against the two
    operators it replaces (BatchNorm2d, then the SelfNorm block), both on channels_last tensors: every output, gradient,
    parameter gradient and buffer BIT-EQUAL -- the fused kernels round where the sequence rounds and sum in the same order --
    with 10 launches instead of 12 (11 / 9 in eval mode: no statistics kernel)."""
    import copy
    import cnsn_b200._lib as L
    import cnsn_b200.cnsn as M
    import cnsn_b200.hosts._norm as HN
    from cnsn_b200.ibn import BatchNorm2d
    dev = "cuda:0"
    cl = torch.channels_last
    g = torch.Generator().manual_seed(sum(shape))
    C = shape[1]
    c0 = (torch.randn(shape, generator=g) * (0.5 + torch.rand(1, C, 1, 1, generator=g)) + 0.3 * torch.randn(1, C, 1, 1, generator=g)).to(dtype)
    s0 = torch.randn(shape, generator=g).to(dtype)
    dy0 = torch.randn(shape, generator=g).to(dtype)
    torch.manual_seed(3)
    bn = BatchNorm2d(C).to(dev).train(training)
    site = M.CNSN(None, M.SelfNorm(C)).to(dev).train(training)
    with torch.no_grad():
        bn.weight.copy_(torch.rand(C, generator=g) + 0.5); bn.bias.copy_(torch.randn(C, generator=g) * 0.3)
        bn.running_mean.copy_(torch.randn(C, generator=g) * 0.1); bn.running_var.copy_(torch.rand(C, generator=g) + 0.5)
        site.selfnorm.g_bn.weight.copy_(torch.rand(C, generator=g) + 0.5); site.selfnorm.g_bn.bias.copy_(torch.randn(C, generator=g) * 0.3)
    res = []
    for fused in (True, False):
        b, st = copy.deepcopy(bn), copy.deepcopy(site)
        c = c0.to(dev).contiguous(memory_format=cl).requires_grad_(True)
        sk = s0.to(dev).contiguous(memory_format=cl).requires_grad_(True)
        HN.FUSE_TAIL = fused
        try:
            n0 = L.launch_count()
            y = HN.bn_site_relu(b, st, c, sk)
            y.backward(dy0.to(dev).contiguous(memory_format=cl))
            torch.cuda.synchronize()
            launched = L.launch_count() - n0
        finally:
            HN.FUSE_TAIL = True
        assert launched == ((10 if fused else 12) if training else (9 if fused else 11)), (fused, launched)
        assert y.is_contiguous(memory_format=cl) and c.grad.is_contiguous(memory_format=cl)
        res.append([y.detach(), c.grad, sk.grad, b.weight.grad, b.bias.grad, b.running_mean, b.running_var, b.num_batches_tracked,
                    st.selfnorm.g_fc.weight.grad, st.selfnorm.g_bn.weight.grad, st.selfnorm.g_bn.bias.grad,
                    st.selfnorm.g_bn.running_mean, st.selfnorm.g_bn.running_var])
    names = ("y", "dc", "dskip", "d bn.weight", "d bn.bias", "bn.running_mean", "bn.running_var", "bn.num_batches_tracked",
             "d g_fc.weight", "d g_bn.weight", "d g_bn.bias", "g_bn.running_mean", "g_bn.running_var")
    for name, a, b in zip(names, res[0], res[1]):
        assert torch.equal(a, b), (name, float((a.double() - b.double()).abs().max()))
    # and against torch in fp64 (the sequence itself is covered operator by operator elsewhere)
    ref_bn = torch.nn.BatchNorm2d(C).to(dev).double().train(training)
    ref_bn.load_state_dict({k: (v.double() if v.dtype.is_floating_point else v) for k, v in bn.state_dict().items()})
    from oracle import eager_modules as E
    ref_sn = E.SelfNorm(C).to(dev).double().train(training)
    ref_sn.load_state_dict({k: (v.double() if v.dtype.is_floating_point else v) for k, v in site.selfnorm.state_dict().items()})
    yr = torch.relu(ref_sn(ref_bn(c0.to(dev).double()) + s0.to(dev).double()))
    tol = 2e-5 if dtype == torch.float32 else 3e-2
    assert float((res[0][0].double() - yr).abs().max()) <= tol * max(1.0, float(yr.abs().max()))
