"""Robustness of the dataflow kernels on a real GPU: shared-memory attribute across plane sizes, concurrent streams,
non-finite inputs, non-contiguous / channels_last inputs, misaligned slices, the asynchronous error state."""
import numpy as np
import pytest
import torch

import helpers as H
from oracle import cnsn_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def mod():
    import cnsn_b200.cnsn as m
    return m


def _sn_case(mod, shape, seed, dtype=torch.float32):
    x = O.varied_input(shape, seed=seed, dtype=np.float32)
    dy = np.random.RandomState(seed + 1).standard_normal(shape).astype(np.float32)
    params, bufs = H.random_sn_params(shape[1], seed=seed + 2)
    r = H.run_selfnorm(mod, x, dy, params, bufs, DEV, False, True, dtype)
    o = H.oracle_selfnorm(x, dy, params, bufs, True)
    return r, o


def test_alternating_plane_sizes_above_48k_share_one_kernel(mod):
    """Two plane sizes above 48 KB through the SAME kernel instantiation, A -> B -> A (the dynamic
    shared-memory attribute is per function and last-write-wins; it is now set once, to the device maximum): 128x128 and
    112x112 fp32 planes are both one-plane items of k_sn_res<float, ., 128 threads per instance>."""
    for shape in ((6, 3, 128, 128), (6, 3, 112, 112), (6, 3, 128, 128), (5, 2, 120, 120), (6, 3, 112, 112)):
        r, o = _sn_case(mod, shape, seed=shape[2])
        assert H.maxabs(r["y"], o["y"]) < 2e-5 and H.maxabs(r["dx"], o["dx"]) < 2e-5, shape
    # the same through CrossNorm and the fused site
    for hw in (128, 112, 128):
        shape = (6, 3, hw, hw)
        x = O.varied_input(shape, seed=hw)
        dy = np.random.RandomState(1).standard_normal(shape).astype(np.float32)
        y, dx = H.run_crossnorm(mod, x, dy, DEV, "both", False, None, 3, 4)
        torch.manual_seed(3)
        np.random.seed(4)
        plan = O.draw_plan(shape, crop="both")
        assert H.maxabs(y, O.crossnorm_fwd(x, plan)) < 2e-5 and H.maxabs(dx, O.crossnorm_bwd(x, dy, plan)) < 2e-5, shape


def test_concurrent_streams_and_a_busy_gpu(mod):
    """Several dataflow kernels in flight at once -- three streams running SelfNorm (shared-memory-resident and the
    tensor-memory pipeline, whose CTAs each want all 512 TMEM columns of an SM) / CrossNorm forward + backward while a
    fourth keeps the SMs -- and their tensor memory: cuBLAS runs tcgen05 GEMMs -- busy with matmuls: the cooperative persistent launch guarantees each kernel's CTAs are
    co-resident whatever else runs, so nothing stalls or times out and the results are bit-identical to the
    single-stream results (the kernels are deterministic)."""
    import cnsn_b200._lib as L
    L.tune(tm_items=0)                                 # the 56x56 shapes below then take the shared + tensor memory pipeline
    shapes = [(64, 32, 32, 32), (48, 16, 56, 56), (40, 12, 56, 56)]
    work = []
    for i, shape in enumerate(shapes):
        g = torch.Generator(device=DEV).manual_seed(i)
        x = (torch.randn(shape, device=DEV, generator=g) * 1.5 + 0.3).requires_grad_(True)
        dy = torch.randn(shape, device=DEV, generator=g)
        sn = mod.SelfNorm(shape[1]).to(DEV).train()
        work.append((x, dy, sn))

    def one(x, dy, sn):
        y = sn(x)
        (dx,) = torch.autograd.grad(y, x, dy)
        torch.manual_seed(11)
        z = mod.cn_op_2ins_space_chan(x, crop="neither", beta=1)
        (dz,) = torch.autograd.grad(z, x, dy)
        return y.detach(), dx, z.detach(), dz

    ref = [one(*w) for w in work]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream() for _ in work]
    busy = torch.cuda.Stream()
    a = torch.randn(4096, 4096, device=DEV)
    outs = [None] * len(work)
    for it in range(40):
        with torch.cuda.stream(busy):
            for _ in range(4):
                a = (a @ a).clamp_(-1, 1)
        for i, w in enumerate(work):
            with torch.cuda.stream(streams[i]):
                outs[i] = one(*w)
    torch.cuda.synchronize()
    L.async_error()                                   # raises if any kernel gave up a bounded wait
    for r, o in zip(ref, outs):
        for tr, to in zip(r, o):
            assert torch.equal(tr, to)


def _eager(shape_c):
    from oracle import eager_modules
    return eager_modules


@pytest.mark.parametrize("shape,channels_last", [((8, 6, 16, 16), False), ((16, 4, 56, 56), False), ((8, 8, 7, 7), False),
                                                 ((8, 8, 16, 16), True), ((6, 16, 40, 40), True)])
@pytest.mark.parametrize("bad", [float("nan"), float("inf")])
def test_selfnorm_propagates_non_finite_like_the_reference(mod, shape, bad, channels_last):
    """A NaN / Inf element makes its instance's statistics non-finite, BatchNorm1d's batch statistics carry that to the
    whole channel (models/cnsn.py:133-150): the kernels must produce the SAME non-finite pattern as the eager chain on
    the same GPU, finite values elsewhere, and must not stall on the polled words (NaN payloads are canonicalised,
    flow_common.cuh ll_publish) -- forward and backward."""
    import cnsn_b200._lib as L
    E = _eager(shape)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(shape, generator=g) * 1.3 + 0.2
    x[1, 2, 3, 3] = bad                              # one element of instance (1, channel 2)
    dy = torch.randn(shape, generator=g)
    torch.manual_seed(1)
    ours = mod.SelfNorm(shape[1]).to(DEV).train()
    ref = E.SelfNorm(shape[1]).to(DEV).train()
    ref.load_state_dict(ours.state_dict())
    res = []
    for m in (ref, ours):
        xt = x.to(DEV)
        if channels_last and m is ours:              # the NHWC kernels (csrc/selfnorm_nhwc.cu): same pattern, same values
            xt = xt.contiguous(memory_format=torch.channels_last)
        xt = xt.requires_grad_(True)
        y = m(xt)
        y.backward(dy.to(DEV))
        res.append((y.detach(), xt.grad, m.g_bn.running_mean.clone(), m.g_fc.weight.grad.clone()))
    torch.cuda.synchronize()
    L.async_error()
    for i, (tr, to) in enumerate(zip(res[0], res[1])):
        assert torch.equal(torch.isfinite(tr), torch.isfinite(to))
        fin = torch.isfinite(tr)
        if i < 3:
            assert torch.allclose(tr[fin], to[fin], atol=2e-5, rtol=1e-4)
        else:       # dW: sums over the batch that cancel -- the eager fp32 chain itself is only good to ~1e-4 of max |dW|
            assert float((tr[fin] - to[fin]).abs().max()) <= 5e-4 * float(tr[fin].abs().max())
    # channel 2 is non-finite everywhere, every other channel is untouched
    assert not torch.isfinite(res[1][0][:, 2]).any() and torch.isfinite(res[1][0][:, [0, 1, 3]]).all()


@pytest.mark.parametrize("crop", ["neither", "both"])
def test_crossnorm_propagates_non_finite_like_the_reference(mod, crop):
    """CrossNorm: a NaN instance poisons itself and the instance that takes its statistics, nobody else."""
    import cnsn_b200._lib as L
    E = _eager(None)
    shape = (8, 4, 16, 16)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(shape, generator=g)
    x[3, 1, 8, 8] = float("nan")                     # centre pixel: inside every accepted crop box? not guaranteed -> compare to eager
    dy = torch.randn(shape, generator=g)
    res = []
    for ops in (E, mod):
        torch.manual_seed(7)
        np.random.seed(8)
        xt = x.to(DEV).requires_grad_(True)
        y = ops.cn_op_2ins_space_chan(xt, crop=crop, beta=1)
        y.backward(dy.to(DEV))
        res.append((y.detach(), xt.grad))
    torch.cuda.synchronize()
    L.async_error()
    # forward: identical non-finite pattern (the reference's mask path multiplies by 0 inside the content box, so a
    # NaN there stays NaN on both sides); finite values agree
    assert torch.equal(torch.isfinite(res[0][0]), torch.isfinite(res[1][0]))
    fin = torch.isfinite(res[0][0])
    assert torch.allclose(res[0][0][fin], res[1][0][fin], atol=2e-5, rtol=1e-4)
    # backward: wherever the reference is finite, so are we, with the same values (0 * NaN products of the eager
    # graph can make the reference NaN in MORE places than the closed form)
    fin = torch.isfinite(res[0][1])
    assert torch.isfinite(res[1][1][fin]).all()
    assert torch.allclose(res[0][1][fin], res[1][1][fin], atol=2e-5, rtol=1e-4)
    assert torch.isfinite(res[1][0][:, [0, 2, 3]]).all()          # other channels untouched


@pytest.mark.parametrize("how", ["channels_last", "transposed", "sliced", "misaligned"])
def test_non_contiguous_inputs_give_the_dense_result(mod, how):
    """The reference forces .contiguous() (models/cnsn.py:14,16); the kernels take dense NCHW, the binding densifies
    anything else (functional._dense) and falls back to the general kernels for base pointers that are not 16-byte
    aligned: SelfNorm, CrossNorm, the fused site and calc_ins_mean_std give the dense tensor's result for
    channels_last tensors, transposed / sliced views and a misaligned contiguous slice."""
    shape = (6, 8, 12, 12)
    g = torch.Generator().manual_seed(0)
    base = torch.randn(7, 8, 12, 13, generator=g).to(DEV) * 1.2 + 0.1
    if how == "channels_last":
        x = base[:6, :, :, :12].contiguous().to(memory_format=torch.channels_last)
    elif how == "transposed":
        x = base[:6, :, :, :12].transpose(2, 3)
    elif how == "sliced":
        x = base[1:7, :, :, 1:13]
    else:                                             # contiguous, but the data pointer is 4 mod 16
        flat = torch.randn(6 * 8 * 12 * 12 + 1, generator=g).to(DEV)
        x = flat[1:].view(shape)
        assert x.is_contiguous() and x.data_ptr() % 16 != 0
    assert how == "misaligned" or not x.is_contiguous()
    xd = x.contiguous().clone()
    dy = torch.randn(shape, generator=g).to(DEV)
    torch.manual_seed(3)
    sn = mod.SelfNorm(8).to(DEV).train()
    blk = mod.CNSN(mod.CrossNorm(crop="both", beta=1), mod.SelfNorm(8).to(DEV)).train()
    outs = []
    for inp in (xd, x):
        sn.g_bn.running_mean.zero_(), sn.g_bn.running_var.fill_(1)
        t = inp.detach().requires_grad_(True)
        y = sn(t)
        (dx,) = torch.autograd.grad(y, t, dy)
        torch.manual_seed(4)
        np.random.seed(5)
        z = mod.cn_op_2ins_space_chan(t, crop="style", beta=1)
        (dz,) = torch.autograd.grad(z, t, dy)
        torch.manual_seed(6)
        np.random.seed(7)
        blk.crossnorm.active = True
        blk.selfnorm.g_bn.running_mean.zero_(), blk.selfnorm.g_bn.running_var.fill_(1)
        s = blk(t)
        (ds,) = torch.autograd.grad(s, t, dy)
        mu, sd = mod.calc_ins_mean_std(inp)
        outs.append((y.detach(), dx, z.detach(), dz, s.detach(), ds, mu, sd))
    for a, b in zip(outs[0], outs[1]):
        assert a.shape == b.shape and torch.allclose(a, b, atol=2e-6, rtol=1e-5)


def test_async_error_state_is_clear_and_clearable():
    import cnsn_b200._lib as L
    assert L.lib().cnsn_async_error(0) == 0
    L.async_error()                                   # no-op when nothing happened
    assert L.lib().cnsn_tune(b"no_such_knob", 1) != 0


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("how", ["channels_last", "crop_view", "transposed", "channels_last_crop", "few_channels_last"])
def test_strided_statistics_in_place(mod, how, dtype):
    """calc_ins_mean_std (models/cnsn.py:8-17) on views the reference would copy first (.contiguous(), :14,:16): the
    strided entry point cnsn_instance_stats_strided reduces them where they lie -- channels_last (lane = channel),
    W-contiguous crops / slices, transposes -- and agrees with torch on the densified tensor, gradients included."""
    import cnsn_b200._lib as L
    g = torch.Generator().manual_seed(0)
    base = (torch.randn(5, 40, 13, 11, generator=g) * 1.4 + 0.3).to(dtype).to(DEV)
    if how == "channels_last":
        x = base.contiguous(memory_format=torch.channels_last)
    elif how == "crop_view":
        x = base[:, :, 2:11, 1:9]                        # what cn_op_2ins_space_chan crops (:66, :77)
    elif how == "transposed":
        x = base.transpose(2, 3)
    elif how == "channels_last_crop":
        x = base.contiguous(memory_format=torch.channels_last)[1:, 3:37, 1:12, 2:10]
    else:
        x = base[:, :5].contiguous(memory_format=torch.channels_last)
    assert not x.is_contiguous()
    n0 = L.launch_count()
    xt = x.detach().requires_grad_(True)
    mean, std = mod.calc_ins_mean_std(xt, eps=1e-5)
    assert L.launch_count() - n0 == 1                    # one statistics kernel, no copy kernel of ours in front
    xr = x.detach().float().contiguous().requires_grad_(True)
    N, C = xr.shape[:2]
    rm = xr.view(N, C, -1).mean(2).view(N, C, 1, 1)
    rs = (xr.view(N, C, -1).var(2) + 1e-5).sqrt().view(N, C, 1, 1)
    tol = dict(atol=2e-5, rtol=1e-5) if dtype == torch.float32 else dict(atol=2e-2, rtol=1e-2)
    assert mean.shape == (N, C, 1, 1) and torch.allclose(mean.float(), rm, **tol) and torch.allclose(std.float(), rs, **tol)
    gm, gs = torch.randn(N, C, 1, 1, device=DEV), torch.randn(N, C, 1, 1, device=DEV)
    (mean.float() * gm + std.float() * gs).sum().backward()
    (rm * gm + rs * gs).sum().backward()
    assert torch.allclose(xt.grad.float(), xr.grad, **tol)


def test_tensor_memory_pipeline_fuzz_against_the_general_path(mod):
    """Seeded fuzz of the shared + tensor memory pipeline (SelfNorm and the whole-plane site): random batch sizes (ragged
    last items, N barely above two items), channel counts, plane sizes across all five vector-per-thread geometries,
    dtypes, the fused ReLU, persistent grids of 1..5 CTAs -- every run against the three-kernel general path / the
    two-operator sequence of the same library on identical inputs and host draws."""
    import cnsn_b200._lib as L
    rs = np.random.RandomState(2024)
    sides = [40, 44, 48, 52, 56, 60, 64]
    for it in range(36):
        hw = sides[rs.randint(len(sides))]
        dtype = [torch.float32, torch.float32, torch.bfloat16][rs.randint(3)]
        if dtype != torch.float32:
            hw = [56, 64, 72, 80, 88][rs.randint(5)]
        N, C = int(rs.randint(16, 45)), int(rs.randint(1, 5))
        relu, site = bool(rs.randint(2)), bool(rs.randint(2))
        cap = int(rs.randint(1, 6))
        g = torch.Generator().manual_seed(it)
        x0 = (torch.randn(N, C, hw, hw, generator=g) * (0.5 + torch.rand(N, C, 1, 1, generator=g)) + torch.randn(N, C, 1, 1, generator=g)).to(dtype).to(DEV)
        dy = torch.randn(N, C, hw, hw, generator=g).to(dtype).to(DEV)
        torch.manual_seed(it)
        sn = mod.SelfNorm(C).to(DEV).train()
        blk = mod.CNSN(mod.CrossNorm(crop="neither", beta=1), sn).train()
        state = {k: v.clone() for k, v in sn.state_dict().items()}
        outs = []
        for knobs, fused in (({"tm": 3, "tm_items": 0, "grid_cap": cap}, True), ({"tm": 0, "selfnorm_impl": "v1", "crossnorm_impl": "v1"}, False)):
            sn.load_state_dict(state)
            sn.zero_grad(set_to_none=True)
            mod.CNSN.fuse_site = fused
            try:
                with L.tuned(**knobs):
                    torch.manual_seed(100 + it)
                    x = x0.clone().requires_grad_(True)
                    if site:
                        blk.crossnorm.active = True
                        y = blk(x, None, True) if relu else blk(x)
                    else:
                        y = sn(x, None, True) if relu else sn(x)
                    y.backward(dy)
            finally:
                mod.CNSN.fuse_site = True
            outs.append([t.detach().float() for t in (y, x.grad, sn.g_fc.weight.grad, sn.g_bn.weight.grad, sn.g_bn.bias.grad,
                                                      sn.g_bn.running_mean, sn.g_bn.running_var)])
        torch.cuda.synchronize()
        tol = 3e-5 if dtype == torch.float32 else 3e-2
        for name, a, b in zip(("y", "dx", "dW", "dgamma", "dbeta", "running_mean", "running_var"), outs[0], outs[1]):
            if relu and name == "dx":
                keep = outs[1][0] != 0                   # clear of the ReLU edge on the reference side ...
                keep &= outs[0][0] != 0                  # ... and on ours
                a, b = a[keep], b[keep]
            err = float((a - b).abs().max() / b.abs().max().clamp_min(1e-6))
            assert err <= tol, (it, name, err, dict(N=N, C=C, hw=hw, dtype=dtype, relu=relu, site=site, cap=cap))
    L.async_error()
