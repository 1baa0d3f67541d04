"""GPU parity of the fused CNSN site (both operators firing, models/cnsn.py:159-164):
cnsn_site_fwd/_bwd through the module surface -> ctypes -> C ABI against
(1) golden fixtures produced by the unmodified reference's CNSN module,
(2) the numpy oracle's composition on seeded inputs (every item geometry of the kernel: 8..128 threads per
    instance, ragged last items, all crop modes, the fused ReLU, bf16 / fp16), and
(3) the two-operator sequence of this package on identical inputs and draws at training sizes.

Tolerances as tests/test_gpu_parity.py: fp32 1e-5 (abs + rel), parameter gradients 1e-5 relative; half precision
allclose(1e-2, 1e-2) (parameter gradients 1e-2 relative) against the fp32 oracle on the upcast inputs.
"""
import numpy as np
import pytest
import torch

import helpers as H
from oracle import cnsn_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def mod():
    import cnsn_b200
    import cnsn_b200.cnsn as m
    assert cnsn_b200.launch_count() >= 0
    return m


def close32(a, b, what):
    b = np.asarray(b, np.float64)
    err = np.abs(np.asarray(a, np.float64) - b)
    tol = H.F32_ATOL + H.F32_ATOL * np.abs(b)
    assert np.all(err <= tol), f"{what}: max err {err.max():.3e} (tol 1e-5 abs+rel)"


def close16(a, b, what):
    np.testing.assert_allclose(a, b, atol=H.BF16_ATOL, rtol=H.BF16_RTOL, err_msg=what)


def run_site(mod, x, dy, params, bufs, crop, tseed, nseed, dtype=torch.float32, relu=False, fused=True):
    """CNSN(CrossNorm(crop), SelfNorm) with the CrossNorm active, fwd + bwd on the GPU; returns numpy results and
    the number of library launches the step took."""
    import cnsn_b200
    C = x.shape[1]
    sn = H.make_selfnorm(mod, C, params, bufs, DEV)
    blk = mod.CNSN(mod.CrossNorm(crop=crop, beta=1), sn).train()
    blk.crossnorm.active = True
    mod.CNSN.fuse_site = fused
    try:
        torch.manual_seed(tseed)
        np.random.seed(nseed)
        xt = torch.from_numpy(x).to(device=DEV, dtype=dtype).requires_grad_(True)
        n0 = cnsn_b200.launch_count()
        y = blk(xt, None, True) if relu else blk(xt)
        y.backward(torch.from_numpy(dy).to(device=DEV, dtype=dtype))
        torch.cuda.synchronize()
        launches = cnsn_b200.launch_count() - n0
    finally:
        mod.CNSN.fuse_site = True
    assert blk.crossnorm.active is False
    out = {"y": y, "dx": xt.grad, "dg_w": sn.g_fc.weight.grad.view(C, 2), "dg_gamma": sn.g_bn.weight.grad,
           "dg_beta": sn.g_bn.bias.grad, "rm": sn.g_bn.running_mean, "rv": sn.g_bn.running_var,
           "nbt": sn.g_bn.num_batches_tracked}
    return {k: v.detach().double().cpu().numpy() for k, v in out.items()}, launches


def oracle_site(x, dy, params, bufs, plan, relu=False, round_to=None, mask=None):
    """The oracle's composition; round_to: element type the CrossNorm output is stored in by the reference
    sequence (half precision runs); mask: where the ReLU passes (default: where the oracle's own z > 0)."""
    z = O.crossnorm_fwd(x, plan)
    if round_to is not None:
        z = torch.from_numpy(z).to(round_to).double().numpy()
    y, nb = O.selfnorm_fwd(z, params, bufs, True)
    d = np.where((z > 0) if mask is None else mask, dy, 0.0) if relu else dy
    dz, gr = O.selfnorm_bwd(z, d, params, bufs, True)
    if round_to is not None:
        dz = torch.from_numpy(dz).to(round_to).double().numpy()
    dx = O.crossnorm_bwd(x, dz, plan)
    return {"y": np.maximum(y, 0.0) if relu else y, "dx": dx, "dg_w": gr["g_w"], "dg_gamma": gr["g_gamma"],
            "dg_beta": gr["g_beta"], "rm": nb["g_rm"], "rv": nb["g_rv"]}


@pytest.mark.parametrize("name", H.golden_names("site_"))
def test_site_golden(mod, name):
    g = H.golden(name)
    params, bufs = H.sn_params_from_golden(g)
    r, launches = run_site(mod, g["x"], g["dy"], params, bufs, str(g["crop"]), int(g["torch_seed"]), int(g["numpy_seed"]))
    assert launches == 2, launches                       # one kernel per direction
    close32(r["y"], g["y_f64"], "y")
    close32(r["dx"], g["dx_f64"], "dx")
    for k in ("w", "gamma", "beta"):
        assert H.relmax(r["dg_" + k], g[f"dg_{k}_f64"]) <= H.PARAM_RTOL, k
    close32(r["rm"], g["g_rm_after_f64"], "running_mean")
    close32(r["rv"], g["g_rv_after_f64"], "running_var")
    assert int(r["nbt"]) == 1


# item geometries: (16,8,8,8) 8 threads per instance; (32,16,32,32) 32 fwd / 64 bwd; (8,4,56,56) 64 fwd / 128 bwd;
# (5,3,8,8) and (37,6,20,20) ragged last items; (64,8,16,16) 16 threads per instance backward
SITE_SHAPES = [(16, 8, 8, 8), (32, 16, 32, 32), (8, 4, 56, 56), (5, 3, 8, 8), (37, 6, 20, 20), (64, 8, 16, 16)]


@pytest.mark.parametrize("shape", SITE_SHAPES)
@pytest.mark.parametrize("crop", ["neither", "style", "content", "both"])
@pytest.mark.parametrize("relu", [False, True])
def test_site_vs_oracle_f32(mod, shape, crop, relu):
    x = O.varied_input(shape, seed=31, dtype=np.float32)
    dy = np.random.RandomState(32).standard_normal(shape).astype(np.float32)
    params, bufs = H.random_sn_params(shape[1], seed=33)
    r, launches = run_site(mod, x, dy, params, bufs, crop, 41, 42, relu=relu)
    assert launches == 2, launches
    torch.manual_seed(41)
    np.random.seed(42)
    plan = O.draw_plan(shape, crop=crop, beta=1)
    # the ReLU mask is the GPU's own decision: an element within fp32 rounding of the kink may fall on either side
    # (y differs by < 1e-6 there); the gradient must be consistent with the side that was taken
    o = oracle_site(x, dy, params, bufs, plan, relu, mask=(r["y"] > 0) if relu else None)
    close32(r["y"], o["y"], "y")
    close32(r["dx"], o["dx"], "dx")
    for k in ("dg_w", "dg_gamma", "dg_beta"):
        assert H.relmax(r[k], o[k]) <= H.PARAM_RTOL, k
    close32(r["rm"], o["rm"], "running_mean")
    close32(r["rv"], o["rv"], "running_var")


@pytest.mark.parametrize("shape,dtype", [((16, 8, 16, 16), torch.bfloat16), ((32, 8, 32, 32), torch.bfloat16),
                                         ((24, 6, 28, 28), torch.float16), ((128, 16, 8, 8), torch.bfloat16)])
@pytest.mark.parametrize("crop", ["neither", "both"])
def test_site_half_vs_oracle(mod, shape, dtype, crop):
    x = torch.from_numpy(O.varied_input(shape, seed=51, dtype=np.float32)).to(dtype).float().numpy()
    dy = torch.from_numpy(np.random.RandomState(52).standard_normal(shape).astype(np.float32)).to(dtype).float().numpy()
    params, bufs = H.random_sn_params(shape[1], seed=53)
    r, launches = run_site(mod, x, dy, params, bufs, crop, 61, 62, dtype=dtype)
    assert launches == 2, launches
    torch.manual_seed(61)
    np.random.seed(62)
    plan = O.draw_plan(shape, crop=crop, beta=1)
    # the fused kernels keep the CrossNorm output (and its gradient) in fp32 on chip: the oracle is the unrounded
    # composition on the upcast inputs; everything within the half-precision tolerance of BASELINE.json (1e-2)
    o = oracle_site(x, dy, params, bufs, plan)
    close16(r["y"], o["y"], "y")
    close16(r["dx"], o["dx"], "dx")
    for k in ("dg_w", "dg_gamma", "dg_beta"):
        assert H.relmax(r[k], o[k]) <= 1e-2, k


@pytest.mark.parametrize("shape,dtype,crop,relu", [((512, 32, 32, 32), torch.float32, "both", False),
                                                   ((512, 64, 16, 16), torch.float32, "both", False),
                                                   ((512, 128, 8, 8), torch.float32, "style", True),
                                                   ((256, 32, 32, 32), torch.bfloat16, "content", False),
                                                   ((64, 64, 56, 56), torch.float32, "neither", True)])
def test_site_training_sizes_vs_two_operator_sequence(mod, shape, dtype, crop, relu):
    """The WideResNet-40-2 sites of BASELINE config 3 (and a ResNet-50 stage-1 plane size): the fused site against
    this package's CrossNorm followed by SelfNorm (each oracle-checked on its own) on identical inputs and draws."""
    g = torch.Generator().manual_seed(5)
    x = (torch.randn(shape, generator=g) * (0.5 + torch.rand(shape[0], shape[1], 1, 1, generator=g))
         + torch.randn(shape[0], shape[1], 1, 1, generator=g)).to(dtype).float().numpy()
    dy = torch.randn(shape, generator=g).to(dtype).float().numpy()
    params, bufs = H.random_sn_params(shape[1], seed=7)
    a, la = run_site(mod, x, dy, params, bufs, crop, 71, 72, dtype=dtype, relu=relu, fused=True)
    b, lb = run_site(mod, x, dy, params, bufs, crop, 71, 72, dtype=dtype, relu=relu, fused=False)
    assert la == 2 and lb >= 4, (la, lb)
    chk = close32 if dtype == torch.float32 else close16
    chk(a["y"], b["y"], "y")
    chk(a["dx"], b["dx"], "dx")
    for k in ("dg_w", "dg_gamma", "dg_beta"):
        assert H.relmax(a[k], b[k]) <= (1e-5 if dtype == torch.float32 else 1e-2), k
    chk(a["rv"], b["rv"], "running_var")


def test_site_unsupported_shapes_take_the_two_operator_sequence(mod):
    """Planes that are not 16-byte multiples (7x7 fp32) and channels too large for the GPU's shared memory
    (224x224 image planes at batch 64): cnsn_site_supported says no, CNSN.forward runs CrossNorm then SelfNorm."""
    import cnsn_b200._lib as L
    for shape in ((16, 8, 7, 7), (64, 3, 224, 224)):
        x = torch.from_numpy(O.varied_input(shape, seed=3)).to(DEV)
        assert not L.backend().site_supported(x)
        params, bufs = H.random_sn_params(shape[1], seed=4)
        dy = np.random.RandomState(5).standard_normal(shape).astype(np.float32)
        r, launches = run_site(mod, x.cpu().numpy(), dy, params, bufs, "both", 81, 82)
        assert launches >= 4
        torch.manual_seed(81)
        np.random.seed(82)
        plan = O.draw_plan(shape, crop="both", beta=1)
        o = oracle_site(x.cpu().numpy().astype(np.float64), dy, params, bufs, plan)
        close32(r["y"], o["y"], "y")
        close32(r["dx"], o["dx"], "dx")
    assert L.backend().site_supported(torch.empty((512, 32, 32, 32), device=DEV))


def test_site_c_abi_error_codes(mod):
    """Direct C-ABI calls: N == 1 is the reference's BatchNorm1d ValueError; an unsupported plane size is
    CNSN_E_UNSUPPORTED, never a silent fallback."""
    import cnsn_b200._lib as L
    be = L.backend()
    sn = mod.SelfNorm(4).to(DEV).train()
    g = L.GateTensors(sn.g_fc.weight, sn.g_bn.weight, sn.g_bn.bias, sn.g_bn.running_mean, sn.g_bn.running_var,
                      sn.g_bn.num_batches_tracked)
    perm1 = torch.zeros(1, dtype=torch.int32, device=DEV)
    with pytest.raises(ValueError, match="Expected more than 1 value per channel"):
        be.site_fwd(torch.randn(1, 4, 8, 8, device=DEV), perm1, (0, 8, 0, 8), (0, 8, 0, 8), 0.0, 1e-5, g, 0.1, 1e-5, 1e-12)
    perm = torch.randperm(6).to(torch.int32).to(DEV)
    with pytest.raises(RuntimeError, match="not supported"):
        be.site_fwd(torch.randn(6, 4, 7, 7, device=DEV), perm, (0, 7, 0, 7), (0, 7, 0, 7), 0.0, 1e-5, g, 0.1, 1e-5, 1e-12)


@pytest.mark.parametrize("shape,crop", [((16, 8, 8, 8), "both"), ((12, 4, 32, 32), "neither"), ((9, 6, 20, 20), "content"),
                                        ((8, 4, 56, 56), "style")])
@pytest.mark.parametrize("relu", [False, True])
def test_site_lam_blend_vs_oracle(mod, shape, crop, relu):
    """The lam blend of cn_op_2ins_space_chan (models/cnsn.py:86-87; no caller of the reference enables it, the C
    ABI carries it): cnsn_site_fwd/_bwd called through the backend with lam = 0.3 against the oracle's composition."""
    import cnsn_b200._lib as L
    C, hh, ww = shape[1], shape[2], shape[3]
    lam = 0.3
    x = O.varied_input(shape, seed=91, dtype=np.float32)
    dy = np.random.RandomState(92).standard_normal(shape).astype(np.float32)
    params, bufs = H.random_sn_params(C, seed=93)
    torch.manual_seed(95)
    np.random.seed(96)
    plan = O.draw_plan(shape, crop=crop, beta=1)
    cw = tuple(int(v) for v in (plan["content_window"] or (0, hh, 0, ww)))
    sw = tuple(int(v) for v in (plan["style_window"] or (0, hh, 0, ww)))
    sn = H.make_selfnorm(mod, C, params, bufs, DEV)
    bn = sn.g_bn
    g = L.GateTensors(sn.g_fc.weight.detach(), bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var,
                      bn.num_batches_tracked)
    be = L.backend()
    xt = torch.from_numpy(x).to(DEV)
    perm = torch.from_numpy(plan["perm"]).to(torch.int32).to(DEV)
    y, save = be.site_fwd(xt, perm, cw, sw, lam, 1e-5, g, 0.1, 1e-5, 1e-12, relu)
    dx, gg = be.site_bwd(xt, torch.from_numpy(dy).to(DEV), perm, cw, sw, lam, 1e-5, g, save, relu)
    y, dx = y.double().cpu().numpy(), dx.double().cpu().numpy()
    z = O.crossnorm_fwd(x, plan, lam)
    yo, nb = O.selfnorm_fwd(z, params, bufs, True)
    d = np.where(y > 0, dy, 0.0) if relu else dy
    dzo, gr = O.selfnorm_bwd(z, d, params, bufs, True)
    dxo = O.crossnorm_bwd(x, dzo, plan, lam)
    close32(y, np.maximum(yo, 0.0) if relu else yo, "y")
    close32(dx, dxo, "dx")
    for a, k in zip(gg, ("g_w", "g_gamma", "g_beta")):
        assert H.relmax(a.double().cpu().numpy(), gr[k]) <= H.PARAM_RTOL, k
    close32(bn.running_var.double().cpu().numpy(), nb["g_rv"], "running_var")


# ------------------------------------------------------------------ shared + tensor memory pipeline (site_tmem.cu)
TM_SITE_SHAPES = [((19, 3, 40, 40), torch.float32), ((16, 2, 48, 48), torch.float32), ((13, 3, 52, 52), torch.float32),
                  ((11, 4, 56, 56), torch.float32), ((9, 2, 64, 64), torch.float32), ((33, 3, 56, 56), torch.bfloat16),
                  ((16, 2, 80, 80), torch.float16)]


@pytest.mark.parametrize("shape,dtype", TM_SITE_SHAPES)
@pytest.mark.parametrize("relu,lam", [(False, None), (True, None), (False, 0.3)])
def test_site_tensor_memory_pipeline_vs_oracle(mod, shape, dtype, relu, lam):
    """Whole-plane windows (crop = 'neither') through the shared + tensor memory pipeline, forced onto small tensors
    (tm_items = 0; tm = 3 also admits the 16-bit types) in every geometry it instantiates: 4..8 vectors per thread,
    ragged last items, persistent grids smaller than the item count (two CTAs: every group runs many two-stage
    iterations and instances regularly find their style source in the same item, the same group or the other CTA),
    the fused ReLU and the lam blend -- against the numpy oracle's composition, and the launch count says one kernel
    per direction."""
    import functools
    import cnsn_b200
    import cnsn_b200._lib as L
    x = O.varied_input(shape, seed=71, dtype=np.float32)
    dy = np.random.RandomState(72).standard_normal(shape).astype(np.float32)
    if dtype != torch.float32:
        x = torch.from_numpy(x).to(dtype).float().numpy()
        dy = torch.from_numpy(dy).to(dtype).float().numpy()
    params, bufs = H.random_sn_params(shape[1], seed=73)
    C = shape[1]
    for knobs in ({"tm_items": 0, "tm": 3}, {"tm_items": 0, "tm": 3, "grid_cap": 2}, {"tm": 0}):
        with L.tuned(**knobs):
            sn = H.make_selfnorm(mod, C, params, bufs, DEV)
            blk = mod.CNSN(mod.CrossNorm(crop="neither", beta=1), sn).train()
            if lam is not None:
                blk.crossnorm.cn_op = functools.partial(mod.cn_op_2ins_space_chan, crop="neither", beta=1, lam=lam)
            blk.crossnorm.active = True
            torch.manual_seed(81)
            np.random.seed(82)
            xt = torch.from_numpy(x).to(device=DEV, dtype=dtype).requires_grad_(True)
            n0 = cnsn_b200.launch_count()
            y = blk(xt, None, True) if relu else blk(xt)
            y.backward(torch.from_numpy(dy).to(device=DEV, dtype=dtype))
            torch.cuda.synchronize()
            assert cnsn_b200.launch_count() - n0 == 2
        r = {"y": y.detach().double().cpu().numpy(), "dx": xt.grad.double().cpu().numpy(),
             "dg_w": sn.g_fc.weight.grad.view(C, 2).double().cpu().numpy(), "dg_gamma": sn.g_bn.weight.grad.double().cpu().numpy(),
             "dg_beta": sn.g_bn.bias.grad.double().cpu().numpy(), "rm": sn.g_bn.running_mean.double().cpu().numpy(),
             "rv": sn.g_bn.running_var.double().cpu().numpy()}
        torch.manual_seed(81)
        np.random.seed(82)
        plan = O.draw_plan(shape, crop="neither", beta=1)
        z = O.crossnorm_fwd(x, plan, lam)
        yo, nb = O.selfnorm_fwd(z, params, bufs, True)
        mask = (r["y"] > 0) if relu else None
        d = np.where(mask, dy, 0.0) if relu else dy
        dz, gr = O.selfnorm_bwd(z, d, params, bufs, True)
        dxo = O.crossnorm_bwd(x, dz, plan, lam)
        if relu:
            yo = np.maximum(yo, 0.0)
        chk = close32 if dtype == torch.float32 else close16
        chk(r["y"], yo, "y")
        chk(r["dx"], dxo, "dx")
        # parameter gradients: 2-4 channels and batches of 9-33 -- the sums over the batch cancel to a few units and the
        # only scale to judge them against is themselves: 3e-5 here (1e-5 at the sizes of test_site_vs_oracle_f32)
        tol = 3e-5 if dtype == torch.float32 else 1e-2
        for k, v in (("dg_w", gr["g_w"]), ("dg_gamma", gr["g_gamma"]), ("dg_beta", gr["g_beta"])):
            assert H.relmax(r[k], v) <= tol, (k, H.relmax(r[k], v), knobs)
        close32(r["rm"], nb["g_rm"], "running_mean")
        close32(r["rv"], nb["g_rv"], "running_var")
    L.async_error()
