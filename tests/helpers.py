"""Shared helpers for the parity tests (test infrastructure; may import oracle/)."""
import glob
import os

import numpy as np
import torch

from oracle import cnsn_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# Tolerances (BASELINE.json north_star): fp32 1e-5, bf16 1e-2.
F32_ATOL = 1e-5
BF16_ATOL = 1e-2
BF16_RTOL = 1e-2
PARAM_RTOL = 1e-5     # parameter gradients: relative to max|grad| (they reach 1e3-1e4)


def golden(name):
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: d[k] for k in d.files}


def golden_names(prefix):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def win_or_none(a):
    a = tuple(int(v) for v in a)
    return None if a[0] < 0 else a


def plan_from_golden(g):
    cp = g["chan_perm"]
    return {"perm": g["perm"].astype(np.int64), "chan_perm": cp.astype(np.int64) if cp.size else None,
            "style_window": win_or_none(g["style_window"]), "content_window": win_or_none(g["content_window"])}


def lam_of(g):
    v = float(g["lam"])
    return None if np.isnan(v) else v


def sn_params_from_golden(g):
    params = {k: g[k] for k in ("g_w", "g_gamma", "g_beta", "f_w", "f_gamma", "f_beta") if k in g}
    bufs = {k: g[k] for k in ("g_rm", "g_rv", "f_rm", "f_rv") if k in g}
    return params, bufs


def maxabs(a, b):
    return float(np.max(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64))))


def relmax(a, b):
    b = np.asarray(b, np.float64)
    return maxabs(a, b) / max(float(np.max(np.abs(b))), 1e-30)


def make_selfnorm(mod, C, params, bufs, device, is_two=False, training=True, dtype=torch.float32):
    """Instantiate <mod>.SelfNorm and load oracle-style params/buffers into it."""
    m = mod.SelfNorm(C, is_two=is_two)
    with torch.no_grad():
        for tag, fc, bn in [("g", m.g_fc, m.g_bn)] + ([("f", m.f_fc, m.f_bn)] if is_two else []):
            fc.weight.copy_(torch.from_numpy(np.asarray(params[tag + "_w"], np.float32)).view(C, 1, 2))
            bn.weight.copy_(torch.from_numpy(np.asarray(params[tag + "_gamma"], np.float32)))
            bn.bias.copy_(torch.from_numpy(np.asarray(params[tag + "_beta"], np.float32)))
            bn.running_mean.copy_(torch.from_numpy(np.asarray(bufs[tag + "_rm"], np.float32)))
            bn.running_var.copy_(torch.from_numpy(np.asarray(bufs[tag + "_rv"], np.float32)))
    return m.to(device).train(training)


def random_sn_params(C, seed, is_two=False):
    rs = np.random.RandomState(seed)
    params, bufs = {}, {}
    for tag in ("g", "f") if is_two else ("g",):
        params[tag + "_w"] = rs.uniform(-0.7, 0.7, (C, 2)).astype(np.float32)
        params[tag + "_gamma"] = rs.uniform(0.5, 1.5, C).astype(np.float32)
        params[tag + "_beta"] = rs.uniform(-0.5, 0.5, C).astype(np.float32)
        bufs[tag + "_rm"] = rs.uniform(-1, 1, C).astype(np.float32)
        bufs[tag + "_rv"] = rs.uniform(0.5, 2, C).astype(np.float32)
    return params, bufs


def run_selfnorm(mod, x, dy, params, bufs, device, is_two=False, training=True, dtype=torch.float32):
    """Run <mod>.SelfNorm fwd+bwd on `device`; returns dict of numpy results (float64)."""
    C = x.shape[1]
    m = make_selfnorm(mod, C, params, bufs, device, is_two, training)
    xt = torch.from_numpy(x).to(device=device, dtype=dtype).requires_grad_(True)
    y = m(xt)
    y.backward(torch.from_numpy(dy).to(device=device, dtype=dtype))
    out = {"y": y, "dx": xt.grad}
    for tag, fc, bn in [("g", m.g_fc, m.g_bn)] + ([("f", m.f_fc, m.f_bn)] if is_two else []):
        out[f"d{tag}_w"] = fc.weight.grad.view(C, 2)
        out[f"d{tag}_gamma"] = bn.weight.grad
        out[f"d{tag}_beta"] = bn.bias.grad
        out[f"{tag}_rm_after"] = bn.running_mean
        out[f"{tag}_rv_after"] = bn.running_var
        out[f"{tag}_nbt_after"] = bn.num_batches_tracked
    return {k: v.detach().double().cpu().numpy() for k, v in out.items()}


def oracle_selfnorm(x, dy, params, bufs, training=True):
    y, nb = O.selfnorm_fwd(x, params, bufs, training)
    dx, gr = O.selfnorm_bwd(x, dy, params, bufs, training)
    out = {"y": y, "dx": dx}
    for tag in ("g", "f"):
        if tag + "_w" in params:
            out[f"d{tag}_w"] = gr[tag + "_w"]
            out[f"d{tag}_gamma"] = gr[tag + "_gamma"]
            out[f"d{tag}_beta"] = gr[tag + "_beta"]
            out[f"{tag}_rm_after"] = nb[tag + "_rm"]
            out[f"{tag}_rv_after"] = nb[tag + "_rv"]
    return out


def run_crossnorm(mod, x, dy, device, crop, chan, lam, torch_seed, numpy_seed, dtype=torch.float32, beta=1):
    """Run <mod>.cn_op_2ins_space_chan fwd+bwd from a seeded host RNG state."""
    torch.manual_seed(int(torch_seed))
    np.random.seed(int(numpy_seed))
    xt = torch.from_numpy(x).to(device=device, dtype=dtype).requires_grad_(True)
    y = mod.cn_op_2ins_space_chan(xt, crop=crop, beta=beta, lam=lam, chan=chan)
    y.backward(torch.from_numpy(dy).to(device=device, dtype=dtype))
    return y.detach().double().cpu().numpy(), xt.grad.detach().double().cpu().numpy()


def eager_selfnorm_f32(x, dy, params, bufs, training=True):
    """The eager-PyTorch fp32 chain (oracle/eager_chain.py, bit-identical to the reference on CPU):
    used to calibrate what fp32 arithmetic itself can deliver on an ill-conditioned case."""
    from oracle import eager_chain as E
    C = x.shape[1]
    g = E.GateState(C)
    with torch.no_grad():
        g.fc_w.copy_(torch.from_numpy(params["g_w"]).view(C, 1, 2))
        g.bn_w.copy_(torch.from_numpy(params["g_gamma"]))
        g.bn_b.copy_(torch.from_numpy(params["g_beta"]))
        g.run_mean.copy_(torch.from_numpy(bufs["g_rm"]))
        g.run_var.copy_(torch.from_numpy(bufs["g_rv"]))
    xt = torch.from_numpy(x).requires_grad_(True)
    y = E.selfnorm(xt, g, training)
    y.backward(torch.from_numpy(dy))
    return {"y": y.detach().numpy(), "dx": xt.grad.numpy(), "dg_w": g.fc_w.grad.view(C, 2).numpy(),
            "dg_gamma": g.bn_w.grad.numpy(), "dg_beta": g.bn_b.grad.numpy()}
