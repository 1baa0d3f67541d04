"""Pins the numpy oracle (oracle/cnsn_oracle.py) against the committed golden fixtures, which are
outputs of the UNMODIFIED reference run in the build container (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

import helpers as H
from oracle import cnsn_oracle as O

TOL64 = 1e-9          # oracle is float64; fixtures hold the reference's float64 run


@pytest.mark.parametrize("name", H.golden_names("selfnorm_"))
def test_selfnorm_oracle_matches_reference_golden(name):
    g = H.golden(name)
    params, bufs = H.sn_params_from_golden(g)
    o = H.oracle_selfnorm(g["x"], g["dy"], params, bufs, bool(g["training"]))
    assert H.maxabs(o["y"], g["y_f64"]) < TOL64
    assert H.maxabs(o["dx"], g["dx_f64"]) < TOL64
    for tag in ("g", "f") if bool(g["is_two"]) else ("g",):
        for k in ("w", "gamma", "beta"):
            assert H.relmax(o[f"d{tag}_{k}"], g[f"d{tag}_{k}_f64"]) < 1e-8, k
        assert H.maxabs(o[f"{tag}_rm_after"], g[f"{tag}_rm_after_f64"]) < TOL64
        assert H.maxabs(o[f"{tag}_rv_after"], g[f"{tag}_rv_after_f64"]) < TOL64
    # and the reference's own fp32 run sits within the stated fp32 tolerance of it (except plain randn,
    # which is ill-conditioned: SURVEY.md fact 10)
    if "randn" not in name:
        assert H.maxabs(g["y_f32"], o["y"]) < 1e-5 and H.maxabs(g["dx_f32"], o["dx"]) < 1e-5


@pytest.mark.parametrize("name", H.golden_names("crossnorm_"))
def test_crossnorm_oracle_matches_reference_golden(name):
    g = H.golden(name)
    plan = H.plan_from_golden(g)
    lam = H.lam_of(g)
    assert H.maxabs(O.crossnorm_fwd(g["x"], plan, lam), g["y_f64"]) < TOL64
    assert H.maxabs(O.crossnorm_bwd(g["x"], g["dy"], plan, lam), g["dx_f64"]) < TOL64
    # replaying the recorded seeds through the oracle's sampler reproduces the recorded plan
    torch.manual_seed(int(g["torch_seed"]))
    np.random.seed(int(g["numpy_seed"]))
    p2 = O.draw_plan(g["x"].shape, crop=str(g["crop"]), beta=1, chan=bool(g["chan"]))
    assert np.array_equal(p2["perm"], plan["perm"])
    assert p2["style_window"] == plan["style_window"] and p2["content_window"] == plan["content_window"]
    if plan["chan_perm"] is not None:
        assert np.array_equal(p2["chan_perm"], plan["chan_perm"])


@pytest.mark.parametrize("name", H.golden_names("site_"))
def test_site_oracle_matches_reference_golden(name):
    """The reference's CNSN module with both operators firing (models/cnsn.py:159-164) pins the composed oracle
    (the checker of the fused CrossNorm -> SelfNorm site kernels)."""
    g = H.golden(name)
    plan = H.plan_from_golden(g)
    params, bufs = H.sn_params_from_golden(g)
    y, _, nb = O.site_fwd(g["x"], plan, params, bufs)
    dx, gr = O.site_bwd(g["x"], g["dy"], plan, params, bufs)
    assert H.maxabs(y, g["y_f64"]) < TOL64 and H.maxabs(dx, g["dx_f64"]) < TOL64
    for k in ("w", "gamma", "beta"):
        assert H.relmax(gr["g_" + k], g[f"dg_{k}_f64"]) < 1e-8, k
    assert H.maxabs(nb["g_rm"], g["g_rm_after_f64"]) < TOL64 and H.maxabs(nb["g_rv"], g["g_rv_after_f64"]) < TOL64
    assert H.maxabs(g["y_f32"], y) < 1e-5 and H.maxabs(g["dx_f32"], dx) < 1e-5


@pytest.mark.parametrize("name", H.golden_names("stats_"))
def test_stats_oracle_matches_reference_golden(name):
    g = H.golden(name)
    m, s = O.instance_stats(g["x"], float(g["eps"]))
    assert H.maxabs(m, g["mean_f64"]) < TOL64 and H.maxabs(s, g["std_f64"]) < TOL64


def test_rng_stream_golden():
    g = H.golden("rng_stream")
    torch.manual_seed(int(g["seed"]))
    np.random.seed(int(g["seed"]) + 1)
    for i in range(int(g["n"])):
        shape, crop = tuple(int(v) for v in g[f"shape{i}"]), str(g[f"crop{i}"])
        plan = O.draw_plan(shape, crop=crop, beta=1)
        assert np.array_equal(plan["perm"], g[f"perm{i}"])
        assert plan["style_window"] == H.win_or_none(g[f"sw{i}"])
        assert plan["content_window"] == H.win_or_none(g[f"cw{i}"])


def test_oracle_edge_cases():
    # batch of one in training is the reference's ValueError (BatchNorm1d)
    x = np.zeros((1, 2, 3, 3))
    p, b = H.random_sn_params(2, 0)
    with pytest.raises(ValueError, match="Expected more than 1 value per channel"):
        O.selfnorm_fwd(x, p, b, True)
    # identity permutation, full windows: CrossNorm is the identity up to eps
    x = O.varied_input((4, 3, 6, 6), 1, np.float64)
    y = O.crossnorm_fwd(x, {"perm": np.arange(4), "chan_perm": None, "style_window": None, "content_window": None})
    assert H.maxabs(y, x) < 1e-12
