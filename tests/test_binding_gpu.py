"""The two host bindings above the C ABI -- the C++ autograd nodes (_cnsn_torch.so, the default on CUDA tensors) and the
ctypes binding (_lib.CudaBackend behind Python autograd functions) -- launch the same kernels with the same arguments:
results must be bit-identical, and the host RNG streams (torch CPU generator, numpy global state) must be consumed
identically (reference draw order models/cnsn.py:61-77)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def mod():
    import cnsn_b200.cnsn as m
    return m


def _both(fn):
    import cnsn_b200._lib as L
    assert L.ext() is not None, "the C++ binding must be built on the GPU box: %r" % (L._ext_error,)
    out = []
    for name in ("ext", "ctypes"):
        old = L.set_binding(name)
        try:
            torch.manual_seed(123)
            np.random.seed(321)
            res = fn()
            out.append((res, torch.get_rng_state().clone(), np.random.get_state()[1].copy(), np.random.get_state()[2]))
        finally:
            L.set_binding(old)
    (a, ta, na, pa), (b, tb, nb, pb) = out
    assert torch.equal(ta, tb), "torch CPU generator consumed differently by the two bindings"
    assert np.array_equal(na, nb) and pa == pb, "numpy global RNG consumed differently by the two bindings"
    assert len(a) == len(b)
    for u, v in zip(a, b):
        assert u.dtype == v.dtype and u.shape == v.shape and torch.equal(u, v)


@pytest.mark.parametrize("shape,dtype", [((8, 16, 8, 8), torch.float32), ((32, 8, 56, 56), torch.float32), ((16, 8, 7, 7), torch.float32),
                                         ((12, 6, 9, 14), torch.float32), ((16, 8, 14, 14), torch.bfloat16)])
@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("res,relu", [(False, False), (True, False), (True, True), (False, True)])
def test_selfnorm_bindings_agree(mod, shape, dtype, training, res, relu):
    g = torch.Generator().manual_seed(0)
    x0 = (torch.randn(shape, generator=g) * 1.3 + 0.2).to(dtype).to(DEV)
    r0 = torch.randn(shape, generator=g).to(dtype).to(DEV)
    dy = torch.randn(shape, generator=g).to(dtype).to(DEV)
    torch.manual_seed(1)
    sn = mod.SelfNorm(shape[1]).to(DEV).train(training)
    state = {k: v.clone() for k, v in sn.state_dict().items()}

    def run():
        sn.load_state_dict(state)
        sn.zero_grad(set_to_none=True)
        x = x0.clone().requires_grad_(True)
        r = r0.clone().requires_grad_(True) if res else None
        y = sn(x, r, relu) if (res or relu) else sn(x)
        y.backward(dy)
        out = [y.detach(), x.grad, sn.g_fc.weight.grad, sn.g_bn.weight.grad, sn.g_bn.bias.grad, sn.g_bn.running_mean.clone(),
               sn.g_bn.running_var.clone(), sn.g_bn.num_batches_tracked.clone()]
        if res:
            out.append(r.grad)
        return out

    _both(run)


@pytest.mark.parametrize("shape,dtype", [((8, 6, 12, 10), torch.float32), ((128, 64, 32, 32), torch.bfloat16), ((6, 5, 7, 7), torch.float32)])
@pytest.mark.parametrize("crop", ["neither", "style", "content", "both"])
@pytest.mark.parametrize("lam", [None, 0.3])
def test_crossnorm_bindings_agree(mod, shape, dtype, crop, lam):
    g = torch.Generator().manual_seed(0)
    x0 = torch.randn(shape, generator=g).to(dtype).to(DEV)
    dy = torch.randn(shape, generator=g).to(dtype).to(DEV)

    def run():
        x = x0.clone().requires_grad_(True)
        y = mod.cn_op_2ins_space_chan(x, crop=crop, beta=1, lam=lam)
        y.backward(dy)
        return [y.detach(), x.grad]

    _both(run)


@pytest.mark.parametrize("shape", [(8, 16, 8, 8), (64, 32, 32, 32), (16, 8, 56, 56)])
@pytest.mark.parametrize("crop", ["neither", "both"])
@pytest.mark.parametrize("relu", [False, True])
def test_site_bindings_agree(mod, shape, crop, relu):
    g = torch.Generator().manual_seed(0)
    x0 = (torch.randn(shape, generator=g) * 1.2 + 0.1).to(DEV)
    dy = torch.randn(shape, generator=g).to(DEV)
    torch.manual_seed(1)
    blk = mod.CNSN(mod.CrossNorm(crop=crop, beta=1), mod.SelfNorm(shape[1])).to(DEV).train()
    state = {k: v.clone() for k, v in blk.state_dict().items()}

    def run():
        blk.load_state_dict(state)
        blk.zero_grad(set_to_none=True)
        blk.crossnorm.active = True
        x = x0.clone().requires_grad_(True)
        y = blk(x, None, relu) if relu else blk(x)
        assert blk.crossnorm.active is False
        y.backward(dy)
        sn = blk.selfnorm
        return [y.detach(), x.grad, sn.g_fc.weight.grad, sn.g_bn.weight.grad, sn.g_bn.bias.grad, sn.g_bn.running_mean.clone(),
                sn.g_bn.running_var.clone()]

    _both(run)


def test_batch_of_one_raises_value_error_in_both_bindings(mod):
    """SelfNorm in training mode with N == 1: ValueError, as nn.BatchNorm1d raises inside the reference."""
    import cnsn_b200._lib as L
    sn = mod.SelfNorm(4).to(DEV).train()
    x = torch.randn(1, 4, 8, 8, device=DEV)
    for name in ("ext", "ctypes"):
        old = L.set_binding(name)
        try:
            with pytest.raises(ValueError, match="Expected more than 1 value per channel"):
                sn(x)
        finally:
            L.set_binding(old)
