"""Pins the oracles to the LIVE reference (build container only; skipped where /root/reference is
absent).  The committed goldens (tests/test_oracle_golden.py) carry the same pin everywhere else."""
import numpy as np
import pytest
import torch

import helpers as H
from _refload import load_reference_cnsn
from oracle import cnsn_oracle as O
from oracle import eager_chain as E

ref = load_reference_cnsn()
pytestmark = pytest.mark.skipif(ref is None, reason="reference checkout not present")


def t64(a):
    return torch.from_numpy(np.asarray(a, np.float64))


@pytest.mark.parametrize("crop", ["neither", "style", "content", "both"])
@pytest.mark.parametrize("chan", [False, True])
@pytest.mark.parametrize("lam", [None, 0.3])
def test_numpy_oracle_crossnorm_vs_reference_autograd(crop, chan, lam):
    shape = (6, 5, 9, 12)
    x = O.varied_input(shape, seed=3, dtype=np.float64)
    dy = np.random.RandomState(5).standard_normal(shape)
    torch.manual_seed(11)
    np.random.seed(12)
    xt = t64(x).requires_grad_(True)
    y = ref.cn_op_2ins_space_chan(xt, crop=crop, beta=1, lam=lam, chan=chan)
    y.backward(t64(dy))
    torch.manual_seed(11)
    np.random.seed(12)
    plan = O.draw_plan(shape, crop=crop, beta=1, chan=chan)
    assert H.maxabs(O.crossnorm_fwd(x, plan, lam), y.detach().numpy()) < 1e-12
    assert H.maxabs(O.crossnorm_bwd(x, dy, plan, lam), xt.grad.numpy()) < 1e-12
    # RNG state after the call must also agree (same number of draws consumed)
    a = (torch.rand(1).item(), np.random.rand())
    torch.manual_seed(11)
    np.random.seed(12)
    ref.cn_op_2ins_space_chan(t64(x), crop=crop, beta=1, lam=lam, chan=chan)
    assert a == (torch.rand(1).item(), np.random.rand())


@pytest.mark.parametrize("two", [False, True])
@pytest.mark.parametrize("training", [True, False])
def test_numpy_oracle_selfnorm_vs_reference_autograd(two, training):
    shape = (5, 7, 6, 4)
    x = O.varied_input(shape, seed=1, dtype=np.float64)
    dy = np.random.RandomState(2).standard_normal(shape)
    params, bufs = H.random_sn_params(shape[1], seed=4, is_two=two)
    m = H.make_selfnorm(ref, shape[1], params, bufs, "cpu", two, training).double()
    with torch.no_grad():               # make_selfnorm loads fp32 values; keep exactly those in fp64
        pass
    xt = t64(x).requires_grad_(True)
    y = m(xt)
    y.backward(t64(dy))
    o = H.oracle_selfnorm(x, dy, params, bufs, training)
    assert H.maxabs(o["y"], y.detach().numpy()) < 1e-12
    assert H.maxabs(o["dx"], xt.grad.numpy()) < 1e-12
    C = shape[1]
    for tag, fc, bn in [("g", m.g_fc, m.g_bn)] + ([("f", m.f_fc, m.f_bn)] if two else []):
        assert H.relmax(o[f"d{tag}_w"], fc.weight.grad.numpy().reshape(C, 2)) < 1e-11
        assert H.relmax(o[f"d{tag}_gamma"], bn.weight.grad.numpy()) < 1e-11
        assert H.relmax(o[f"d{tag}_beta"], bn.bias.grad.numpy()) < 1e-11
        assert H.maxabs(o[f"{tag}_rm_after"], bn.running_mean.numpy()) < 1e-12
        assert H.maxabs(o[f"{tag}_rv_after"], bn.running_var.numpy()) < 1e-12


def test_eager_chain_is_bit_identical_to_reference():
    """The CPU-baseline port must execute the same ATen ops: identical fp32 bits, fwd and bwd."""
    shape = (8, 6, 10, 7)
    x = torch.from_numpy(O.varied_input(shape, seed=2))
    dy = torch.randn(shape, generator=torch.Generator().manual_seed(1))
    # SelfNorm
    torch.manual_seed(0)
    m = ref.SelfNorm(6).train()
    g = E.GateState(6)
    with torch.no_grad():
        g.fc_w.copy_(m.g_fc.weight)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    ya, yb = m(xa), E.selfnorm(xb, g, True)
    ya.backward(dy)
    yb.backward(dy)
    assert torch.equal(ya, yb) and torch.equal(xa.grad, xb.grad)
    assert torch.equal(m.g_fc.weight.grad, g.fc_w.grad) and torch.equal(m.g_bn.running_var, g.run_var)
    # CrossNorm, every crop mode
    for crop in ("neither", "style", "content", "both"):
        torch.manual_seed(3)
        np.random.seed(4)
        xa = x.clone().requires_grad_(True)
        ya = ref.cn_op_2ins_space_chan(xa, crop=crop, beta=1)
        ya.backward(dy)
        torch.manual_seed(3)
        np.random.seed(4)
        plan = O.draw_plan(shape, crop=crop, beta=1)
        xb = x.clone().requires_grad_(True)
        yb = E.crossnorm(xb, torch.from_numpy(plan["perm"]), plan["style_window"], plan["content_window"])
        yb.backward(dy)
        assert torch.equal(ya, yb) and torch.equal(xa.grad, xb.grad), crop


@pytest.mark.parametrize("crop", ["neither", "style", "content", "both"])
def test_numpy_oracle_site_vs_reference_cnsn_module(crop):
    """The reference's CNSN module with its CrossNorm active (models/cnsn.py:159-164), executed live in fp64, against
    the oracle's composition site_fwd / site_bwd -- the checker of the fused site kernels."""
    shape = (6, 5, 8, 12)
    x = O.varied_input(shape, seed=7, dtype=np.float64)
    dy = np.random.RandomState(8).standard_normal(shape)
    params, bufs = H.random_sn_params(shape[1], seed=9)
    sn = H.make_selfnorm(ref, shape[1], params, bufs, "cpu").double()
    blk = ref.CNSN(ref.CrossNorm(crop=crop, beta=1), sn).train()
    blk.crossnorm.active = True
    torch.manual_seed(21)
    np.random.seed(22)
    xt = t64(x).requires_grad_(True)
    y = blk(xt)
    y.backward(t64(dy))
    assert blk.crossnorm.active is False
    torch.manual_seed(21)
    np.random.seed(22)
    plan = O.draw_plan(shape, crop=crop, beta=1)
    yo, _, nb = O.site_fwd(x, plan, params, bufs)
    dxo, gr = O.site_bwd(x, dy, plan, params, bufs)
    assert H.maxabs(yo, y.detach().numpy()) < 1e-12 and H.maxabs(dxo, xt.grad.numpy()) < 1e-11
    assert H.relmax(gr["g_w"], sn.g_fc.weight.grad.view(-1, 2).numpy()) < 1e-10
    assert H.relmax(gr["g_gamma"], sn.g_bn.weight.grad.numpy()) < 1e-10
    assert H.maxabs(nb["g_rv"], sn.g_bn.running_var.numpy()) < 1e-12
