"""CPU oracle for the CrossNorm / SelfNorm hot path -- TEST INFRASTRUCTURE ONLY.

This file is the checker, never the product.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  The product
package (``crossnorm-selfnorm_b200/``) never imports anything under ``oracle/`` and raises
when its CUDA library is missing.

What it is: a closed-form numpy float64 restatement of ``models/cnsn.py`` of
amazon-science/crossnorm-selfnorm (forward AND the backward PyTorch autograd derives from
it), written from the maths in SURVEY.md Appendix A.  Each function cites the reference
lines it follows (paths relative to /root/reference).

Parity status: PINNED.  The reference ships no golden vectors or tests,
so the oracle is pinned by executing the reference itself in the build container:
``tests/golden/make_golden.py`` imports ``/root/reference/models/cnsn.py``, runs its
forward and autograd backward on fixed seeds and commits the results under
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this oracle against those
fixtures everywhere, and ``tests/test_oracle_vs_reference.py`` re-checks against the live
reference whenever ``/root/reference`` is present.
"""
from __future__ import annotations

import numpy as np

F64 = np.float64

CN_EPS = 1e-5      # models/cnsn.py:8   (default eps, used by instance_norm_mix at :24-25)
SN_EPS = 1e-12     # models/cnsn.py:133
BN_EPS = 1e-5      # nn.BatchNorm1d default, models/cnsn.py:121
BN_MOMENTUM = 0.1  # nn.BatchNorm1d default


# --------------------------------------------------------------------------------------
# windows
# --------------------------------------------------------------------------------------
def full_window(H, W):
    return (0, H, 0, W)


def _win(x, win):
    h0, h1, w0, w1 = win
    return x[:, :, h0:h1, w0:w1]


# --------------------------------------------------------------------------------------
# a1: calc_ins_mean_std                                            models/cnsn.py:8-17
# --------------------------------------------------------------------------------------
def instance_stats(x, eps=CN_EPS, window=None):
    """Per-(n,c) mean and std=sqrt(unbiased_var+eps) over the window (default: whole plane).

    Follows models/cnsn.py:14-16: ``var(dim=2)`` is Bessel-corrected, eps is added to the
    variance, not the std.  Returns two (N,C) float64 arrays.
    """
    x = np.asarray(x, dtype=F64)
    assert x.ndim == 4                                   # models/cnsn.py:12
    if window is not None:
        x = _win(x, window)
    N, C = x.shape[:2]
    flat = x.reshape(N, C, -1)
    M = flat.shape[2]
    mean = flat.mean(axis=2)
    if M > 1:
        var = ((flat - mean[:, :, None]) ** 2).sum(axis=2) / (M - 1)
    else:                                                # torch gives nan for a 1-element var
        var = np.full((N, C), np.nan)
    return mean, np.sqrt(var + eps)


# --------------------------------------------------------------------------------------
# a3: cn_rand_bbox                                                 models/cnsn.py:32-55
# --------------------------------------------------------------------------------------
def rand_window(size, beta, bbx_thres, rng=np.random):
    """Rejection-sample a crop window.  Returns (h0, h1, w0, w1).

    models/cnsn.py:34-35 calls dim 2 "W" and dim 3 "H"; bbx* then slices dim 2 and bby*
    slices dim 3 (:66,:77), so in NCHW terms bbx == rows (h) and bby == cols (w).  Draw
    order per attempt: beta, randint(dim2), randint(dim3) (:37,:43,:44).
    """
    d2, d3 = int(size[2]), int(size[3])
    while True:
        ratio = rng.beta(beta, beta)
        cut = np.sqrt(ratio)
        c2 = int(d2 * cut)                               # np.int truncation, :39-40
        c3 = int(d3 * cut)
        p2 = rng.randint(d2)
        p3 = rng.randint(d3)
        a0 = int(np.clip(p2 - c2 // 2, 0, d2))
        b0 = int(np.clip(p3 - c3 // 2, 0, d3))
        a1 = int(np.clip(p2 + c2 // 2, 0, d2))
        b1 = int(np.clip(p3 + c3 // 2, 0, d3))
        if float(a1 - a0) * (b1 - b0) / (d2 * d3) > bbx_thres:   # :51-53
            return (a0, a1, b0, b1)


def draw_plan(shape, crop="neither", beta=1, bbx_thres=0.1, chan=False):
    """Consume host RNG exactly like one cn_op_2ins_space_chan call (models/cnsn.py:58-91).

    Order: torch.randperm(N) on the CPU generator (:62); style window if crop in
    {style, both} (:64-65); torch.randperm(C) if chan (:70-71); content window if crop in
    {content, both} (:74-76).
    """
    import torch
    assert crop in ("neither", "style", "content", "both")           # :61
    N, C, H, W = shape
    perm = torch.randperm(N).numpy().astype(np.int64)
    swin = cwin = None
    if crop in ("style", "both"):
        swin = rand_window(shape, beta, bbx_thres)
    cperm = None
    if chan:
        cperm = torch.randperm(C).numpy().astype(np.int64)
    if crop in ("content", "both"):
        cwin = rand_window(shape, beta, bbx_thres)
    return {"perm": perm, "chan_perm": cperm, "style_window": swin, "content_window": cwin}


# --------------------------------------------------------------------------------------
# a2 + a4: instance_norm_mix / cn_op_2ins_space_chan               models/cnsn.py:20-29, 58-91
# --------------------------------------------------------------------------------------
def _cn_setup(x, plan):
    N, C, H, W = x.shape
    perm = np.asarray(plan["perm"], dtype=np.int64)
    cperm = plan.get("chan_perm")
    cperm = np.arange(C) if cperm is None else np.asarray(cperm, dtype=np.int64)
    cw = plan.get("content_window") or full_window(H, W)
    sw = plan.get("style_window") or full_window(H, W)
    return perm, cperm, cw, sw


def crossnorm_fwd(x, plan, lam=None, eps=CN_EPS):
    """y = lam*x + (1-lam)*((x-mu_c)/sd_c * sd_s[p,pi] + mu_s[p,pi]) on the content window, x elsewhere.

    models/cnsn.py:64-68 (style source = x[perm] cropped to the style window), :70-72
    (channel permutation of the style source), :74-84 (content window + copy-through mask),
    :20-29 (the mix), :86-89 (lam blend).
    """
    x = np.asarray(x, dtype=F64)
    perm, cperm, cw, sw = _cn_setup(x, plan)
    mu_c, sd_c = instance_stats(x, eps, cw)
    mu_s, sd_s = instance_stats(x, eps, sw)
    mu_t = mu_s[perm][:, cperm]                          # stats of x2 = x[perm][:, cperm]
    sd_t = sd_s[perm][:, cperm]
    h0, h1, w0, w1 = cw
    xa = x.copy()
    xc = x[:, :, h0:h1, w0:w1]
    xa[:, :, h0:h1, w0:w1] = (xc - mu_c[:, :, None, None]) / sd_c[:, :, None, None] \
        * sd_t[:, :, None, None] + mu_t[:, :, None, None]
    if lam is not None:
        return x * lam + xa * (1.0 - lam)
    return xa


def crossnorm_bwd(x, dy, plan, lam=None, eps=CN_EPS):
    """dx of crossnorm_fwd (what autograd derives from models/cnsn.py:58-91).

    Nothing is detached in the reference: gradient flows through the content statistics and,
    via the permutation, through the style instance's statistics back into the style instance.
    """
    x = np.asarray(x, dtype=F64)
    dy = np.asarray(dy, dtype=F64)
    N, C, H, W = x.shape
    perm, cperm, cw, sw = _cn_setup(x, plan)
    l = 0.0 if lam is None else float(lam)
    mu_c, sd_c = instance_stats(x, eps, cw)
    mu_s, sd_s = instance_stats(x, eps, sw)
    mu_t = mu_s[perm][:, cperm]
    sd_t = sd_s[perm][:, cperm]
    h0, h1, w0, w1 = cw
    Mc = (h1 - h0) * (w1 - w0)
    g0, g1, v0, v1 = sw
    Ms = (g1 - g0) * (v1 - v0)

    dx = dy.copy()                                       # copy-through outside the content window
    d = (1.0 - l) * dy[:, :, h0:h1, w0:w1]
    xhat = (x[:, :, h0:h1, w0:w1] - mu_c[:, :, None, None]) / sd_c[:, :, None, None]
    A = sd_t / sd_c
    S1 = d.sum(axis=(2, 3))
    S2 = (d * xhat).sum(axis=(2, 3))
    dx[:, :, h0:h1, w0:w1] = l * dy[:, :, h0:h1, w0:w1] + A[:, :, None, None] * (
        d - S1[:, :, None, None] / Mc - xhat * S2[:, :, None, None] / (Mc - 1))
    # scatter the style-statistics gradients to the instance/channel they were read from
    dmu_s = np.zeros((N, C))
    dsd_s = np.zeros((N, C))
    src_n = perm[:, None].repeat(C, 1)
    src_c = cperm[None, :].repeat(N, 0)
    dmu_s[src_n, src_c] = S1                             # bijection: plain assignment
    dsd_s[src_n, src_c] = S2
    xs = x[:, :, g0:g1, v0:v1]
    dx[:, :, g0:g1, v0:v1] += dmu_s[:, :, None, None] / Ms + \
        (xs - mu_s[:, :, None, None]) / sd_s[:, :, None, None] * dsd_s[:, :, None, None] / (Ms - 1)
    return dx


# --------------------------------------------------------------------------------------
# a6: SelfNorm                                                     models/cnsn.py:113-150
# --------------------------------------------------------------------------------------
def _gate_fwd(mu, sd, w, gamma, beta, run_mean, run_var, training, bn_eps):
    """sigmoid(BN1d(depthwise k=2 conv over (mu, sd))) -- models/cnsn.py:135-140.

    w is (C,2) = g_fc.weight[:,0,:].  Train mode normalises with the batch mean and the
    BIASED batch variance over N; eval mode with the running buffers.
    """
    N = mu.shape[0]
    s = mu * w[None, :, 0] + sd * w[None, :, 1]
    if training:
        if N < 2:
            raise ValueError("Expected more than 1 value per channel when training")
        m = s.mean(axis=0)
        q = ((s - m[None, :]) ** 2).mean(axis=0)
    else:
        m, q = np.asarray(run_mean, F64), np.asarray(run_var, F64)
    r = 1.0 / np.sqrt(q + bn_eps)
    shat = (s - m[None, :]) * r[None, :]
    z = shat * gamma[None, :] + beta[None, :]
    g = 1.0 / (1.0 + np.exp(-z))
    return s, m, q, r, shat, g


def _gate_bwd(dgate, mu, sd, w, gamma, r, shat, g, training):
    """Backward of _gate_fwd: returns (ds, dw (C,2), dgamma, dbeta).  SURVEY.md A.1."""
    dz = dgate * g * (1.0 - g)
    dgamma = (dz * shat).sum(axis=0)
    dbeta = dz.sum(axis=0)
    dshat = dz * gamma[None, :]
    if training:
        ds = r[None, :] * (dshat - dshat.mean(axis=0)[None, :]
                           - shat * (dshat * shat).mean(axis=0)[None, :])
    else:
        ds = r[None, :] * dshat
    dw = np.stack([(ds * mu).sum(axis=0), (ds * sd).sum(axis=0)], axis=1)
    return ds, dw, dgamma, dbeta


def selfnorm_fwd(x, params, buffers, training=True, eps=SN_EPS, bn_eps=BN_EPS,
                 momentum=BN_MOMENTUM):
    """SelfNorm forward.  params: {'g_w':(C,2),'g_gamma','g_beta'[, 'f_w','f_gamma','f_beta']};
    buffers: {'g_rm','g_rv'[, 'f_rm','f_rv']}.  Returns (y, new_buffers).

    models/cnsn.py:130-150.  is_two (f_* present): y = x*g + mu*(f-g) (:142-148).
    Running stats: rm <- (1-mom)*rm + mom*m ; rv <- (1-mom)*rv + mom*q*N/(N-1) (BatchNorm1d).
    """
    x = np.asarray(x, dtype=F64)
    N, C, H, W = x.shape
    mu, sd = instance_stats(x, eps)
    P = {k: np.asarray(v, F64) for k, v in params.items()}
    new_buf = {k: np.asarray(v, F64).copy() for k, v in buffers.items()}
    out = {}
    for tag in ("g", "f"):
        if tag + "_w" not in P:
            continue
        s, m, q, r, shat, gate = _gate_fwd(mu, sd, P[tag + "_w"], P[tag + "_gamma"], P[tag + "_beta"],
                                           buffers.get(tag + "_rm"), buffers.get(tag + "_rv"),
                                           training, bn_eps)
        out[tag] = gate
        if training:
            new_buf[tag + "_rm"] = (1 - momentum) * new_buf[tag + "_rm"] + momentum * m
            new_buf[tag + "_rv"] = (1 - momentum) * new_buf[tag + "_rv"] + momentum * q * N / (N - 1)
    g = out["g"][:, :, None, None]
    if "f" in out:
        f = out["f"][:, :, None, None]
        y = x * g + mu[:, :, None, None] * (f - g)
    else:
        y = x * g
    return y, new_buf


def selfnorm_bwd(x, dy, params, buffers, training=True, eps=SN_EPS, bn_eps=BN_EPS):
    """Backward of selfnorm_fwd (buffers = the ones the forward SAW, i.e. before its update).

    Returns (dx, grads) with grads keyed like params.  SURVEY.md A.1:
      dg = sum_hw dy*x ; dx = dy*g + a + b*(x-mu), a = ds*w0/M, b = ds*w1/((M-1)*sd).
    is_two: dg = sum dy*x - mu*T, df = mu*T, dmu += (f-g)*T with T = sum_hw dy.
    """
    x = np.asarray(x, dtype=F64)
    dy = np.asarray(dy, dtype=F64)
    N, C, H, W = x.shape
    M = H * W
    mu, sd = instance_stats(x, eps)
    P = {k: np.asarray(v, F64) for k, v in params.items()}
    two = "f_w" in P
    fw = {}
    for tag in ("g", "f"):
        if tag + "_w" in P:
            fw[tag] = _gate_fwd(mu, sd, P[tag + "_w"], P[tag + "_gamma"], P[tag + "_beta"],
                                buffers.get(tag + "_rm"), buffers.get(tag + "_rv"), training, bn_eps)
    g = fw["g"][5]
    Sxy = (dy * x).sum(axis=(2, 3))
    T = dy.sum(axis=(2, 3))
    grads = {}
    dmu = np.zeros((N, C))
    dsd = np.zeros((N, C))
    if two:
        f = fw["f"][5]
        dgate = {"g": Sxy - mu * T, "f": mu * T}
        dmu += (f - g) * T
    else:
        dgate = {"g": Sxy}
    for tag in fw:
        s, m, q, r, shat, gate = fw[tag]
        ds, dw, dgam, dbet = _gate_bwd(dgate[tag], mu, sd, P[tag + "_w"], P[tag + "_gamma"], r, shat,
                                       gate, training)
        grads[tag + "_w"], grads[tag + "_gamma"], grads[tag + "_beta"] = dw, dgam, dbet
        dmu += ds * P[tag + "_w"][None, :, 0]
        dsd += ds * P[tag + "_w"][None, :, 1]
    a = dmu / M
    b = dsd / ((M - 1) * sd)
    dx = dy * g[:, :, None, None] + a[:, :, None, None] + b[:, :, None, None] * (x - mu[:, :, None, None])
    return dx, grads


# --------------------------------------------------------------------------------------
# a7: CNSN with both operators firing (the fused site)             models/cnsn.py:159-164
# --------------------------------------------------------------------------------------
def site_fwd(x, plan, params, buffers, lam=None, training=True, cn_eps=CN_EPS, eps=SN_EPS,
             bn_eps=BN_EPS, momentum=BN_MOMENTUM):
    """CNSN.forward when the site's CrossNorm is active: ``x = crossnorm(x); x = selfnorm(x)``
    (models/cnsn.py:160-163).  Returns (y, z, new_buffers) with z the CrossNorm output."""
    z = crossnorm_fwd(x, plan, lam, cn_eps)
    y, nb = selfnorm_fwd(z, params, buffers, training, eps, bn_eps, momentum)
    return y, z, nb


def site_bwd(x, dy, plan, params, buffers, lam=None, training=True, cn_eps=CN_EPS, eps=SN_EPS,
             bn_eps=BN_EPS):
    """Backward of site_fwd: the chain rule through SelfNorm (at z) and then CrossNorm (at x), as
    autograd composes them for models/cnsn.py:160-163.  Returns (dx, grads)."""
    z = crossnorm_fwd(x, plan, lam, cn_eps)
    dz, grads = selfnorm_bwd(z, dy, params, buffers, training, eps, bn_eps)
    return crossnorm_bwd(x, dz, plan, lam, cn_eps), grads


# --------------------------------------------------------------------------------------
# inputs used by the parity tests and the bench
# --------------------------------------------------------------------------------------
def varied_input(shape, seed=0, dtype=np.float32, relu=False):
    """x = randn*(0.5+1.5*U[n,c]) + randn[n,c]: per-instance varied scale/shift so SelfNorm's
    BN-over-batch is well conditioned (plain randn makes every instance (0,1) and the batch
    variance of the gate input tiny -- SURVEY.md fact 10)."""
    rs = np.random.RandomState(seed)
    N, C, H, W = shape
    x = rs.standard_normal(shape) * (0.5 + 1.5 * rs.random_sample((N, C, 1, 1))) \
        + rs.standard_normal((N, C, 1, 1))
    if relu:
        x = np.maximum(x, 0.0)
    return x.astype(dtype)
