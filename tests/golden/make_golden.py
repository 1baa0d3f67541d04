"""Generate tests/golden/*.npz by RUNNING THE UNMODIFIED REFERENCE (models/cnsn.py) on CPU.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference ships no golden vectors, so these fixtures -- forward
outputs and autograd gradients of the reference's own CrossNorm / SelfNorm on fixed seeds --
are what pins the oracle (oracle/cnsn_oracle.py) and, through it, the CUDA path.
Every case stores float64 results ("*_f64": the reference run in double) and float32 results
("*_f32": the reference run as shipped, in float32) of the same inputs.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from _refload import load_reference_cnsn  # noqa: E402
from oracle.cnsn_oracle import varied_input  # noqa: E402

ref = load_reference_cnsn()
assert ref is not None, "reference not found; run this in the build container"


def np32(t):
    return t.detach().to(torch.float32).numpy().copy()


def np64(t):
    return t.detach().to(torch.float64).numpy().copy()


def selfnorm_case(shape, seed, training, is_two, kind):
    N, C, H, W = shape
    if kind == "randn":
        x = np.random.RandomState(seed).standard_normal(shape).astype(np.float32)
    else:
        x = varied_input(shape, seed=seed, dtype=np.float32, relu=(kind == "relu"))
    dy = np.random.RandomState(seed + 100).standard_normal(shape).astype(np.float32)
    out = {"x": x, "dy": dy, "training": np.array(training), "is_two": np.array(is_two)}
    for prec, dt in (("f32", torch.float32), ("f64", torch.float64)):
        torch.manual_seed(seed)
        m = ref.SelfNorm(C, is_two=is_two)
        with torch.no_grad():           # non-trivial BN affine + running buffers
            gen = torch.Generator().manual_seed(seed + 7)
            for bn in [m.g_bn] + ([m.f_bn] if is_two else []):
                bn.weight.copy_(torch.rand(C, generator=gen) + 0.5)
                bn.bias.copy_(torch.rand(C, generator=gen) - 0.5)
                bn.running_mean.copy_(torch.rand(C, generator=gen) * 2 - 1)
                bn.running_var.copy_(torch.rand(C, generator=gen) * 1.5 + 0.5)
        m = m.to(dt).train(training)
        if prec == "f32":
            for tag, fc, bn in [("g", m.g_fc, m.g_bn)] + ([("f", m.f_fc, m.f_bn)] if is_two else []):
                out[tag + "_w"] = np32(fc.weight)[:, 0, :]
                out[tag + "_gamma"] = np32(bn.weight)
                out[tag + "_beta"] = np32(bn.bias)
                out[tag + "_rm"] = np32(bn.running_mean)
                out[tag + "_rv"] = np32(bn.running_var)
        xt = torch.from_numpy(x).to(dt).requires_grad_(True)
        y = m(xt)
        y.backward(torch.from_numpy(dy).to(dt))
        cv = np32 if prec == "f32" else np64
        out["y_" + prec] = cv(y)
        out["dx_" + prec] = cv(xt.grad)
        for tag, fc, bn in [("g", m.g_fc, m.g_bn)] + ([("f", m.f_fc, m.f_bn)] if is_two else []):
            out[f"d{tag}_w_{prec}"] = cv(fc.weight.grad)[:, 0, :]
            out[f"d{tag}_gamma_{prec}"] = cv(bn.weight.grad)
            out[f"d{tag}_beta_{prec}"] = cv(bn.bias.grad)
            out[f"{tag}_rm_after_{prec}"] = cv(bn.running_mean)
            out[f"{tag}_rv_after_{prec}"] = cv(bn.running_var)
            out[f"{tag}_nbt_after"] = np.array(int(bn.num_batches_tracked))
    return out


def crossnorm_case(shape, seed, crop, chan, lam, dtype_tag="f32"):
    """Runs cn_op_2ins_space_chan; records the RNG seeds so the plan can be replayed, and the plan
    itself (recovered by re-drawing with the reference's own sampler from the same state)."""
    x = varied_input(shape, seed=seed, dtype=np.float32)
    if dtype_tag == "bf16":             # bf16 oracle = fp32 reference on the upcast bf16 input
        x = torch.from_numpy(x).to(torch.bfloat16).to(torch.float32).numpy()
    dy = np.random.RandomState(seed + 100).standard_normal(shape).astype(np.float32)
    if dtype_tag == "bf16":
        dy = torch.from_numpy(dy).to(torch.bfloat16).to(torch.float32).numpy()
    out = {"x": x, "dy": dy, "crop": np.array(crop), "chan": np.array(chan),
           "lam": np.array(np.nan if lam is None else lam), "torch_seed": np.array(seed + 1),
           "numpy_seed": np.array(seed + 2)}
    for prec, dt in (("f32", torch.float32), ("f64", torch.float64)):
        torch.manual_seed(seed + 1)
        np.random.seed(seed + 2)
        xt = torch.from_numpy(x).to(dt).requires_grad_(True)
        y = ref.cn_op_2ins_space_chan(xt, crop=crop, beta=1, lam=lam, chan=chan)
        y.backward(torch.from_numpy(dy).to(dt))
        cv = np32 if prec == "f32" else np64
        out["y_" + prec] = cv(y)
        out["dx_" + prec] = cv(xt.grad)
    # the plan, drawn with the reference's own functions in the reference's order (cnsn.py:62-76)
    torch.manual_seed(seed + 1)
    np.random.seed(seed + 2)
    out["perm"] = torch.randperm(shape[0]).numpy()
    sw = cw = (-1, -1, -1, -1)
    if crop in ("style", "both"):
        b = ref.cn_rand_bbox(x.shape, beta=1, bbx_thres=0.1)      # (bbx1, bby1, bbx2, bby2)
        sw = (int(b[0]), int(b[2]), int(b[1]), int(b[3]))         # -> (h0, h1, w0, w1)
    out["chan_perm"] = torch.randperm(shape[1]).numpy() if chan else np.zeros(0, np.int64)
    if crop in ("content", "both"):
        b = ref.cn_rand_bbox(x.shape, beta=1, bbx_thres=0.1)
        cw = (int(b[0]), int(b[2]), int(b[1]), int(b[3]))
    out["style_window"] = np.array(sw)
    out["content_window"] = np.array(cw)
    return out


def site_case(shape, seed, crop):
    """The reference's CNSN module with both operators firing (models/cnsn.py:152-164): CrossNorm(crop, beta=1)
    with ``active = True`` followed by SelfNorm, training mode; plan recorded as in crossnorm_case."""
    N, C, H, W = shape
    x = varied_input(shape, seed=seed, dtype=np.float32)
    dy = np.random.RandomState(seed + 100).standard_normal(shape).astype(np.float32)
    out = {"x": x, "dy": dy, "crop": np.array(crop), "chan": np.array(False), "lam": np.array(np.nan),
           "torch_seed": np.array(seed + 1), "numpy_seed": np.array(seed + 2)}
    for prec, dt in (("f32", torch.float32), ("f64", torch.float64)):
        torch.manual_seed(seed)
        sn = ref.SelfNorm(C)
        with torch.no_grad():
            gen = torch.Generator().manual_seed(seed + 7)
            sn.g_bn.weight.copy_(torch.rand(C, generator=gen) + 0.5)
            sn.g_bn.bias.copy_(torch.rand(C, generator=gen) - 0.5)
            sn.g_bn.running_mean.copy_(torch.rand(C, generator=gen) * 2 - 1)
            sn.g_bn.running_var.copy_(torch.rand(C, generator=gen) * 1.5 + 0.5)
        m = ref.CNSN(ref.CrossNorm(crop=crop, beta=1), sn).to(dt).train(True)
        if prec == "f32":
            out["g_w"] = np32(sn.g_fc.weight)[:, 0, :]
            out["g_gamma"] = np32(sn.g_bn.weight)
            out["g_beta"] = np32(sn.g_bn.bias)
            out["g_rm"] = np32(sn.g_bn.running_mean)
            out["g_rv"] = np32(sn.g_bn.running_var)
        m.crossnorm.active = True
        torch.manual_seed(seed + 1)
        np.random.seed(seed + 2)
        xt = torch.from_numpy(x).to(dt).requires_grad_(True)
        y = m(xt)
        assert m.crossnorm.active is False                      # one-shot flag consumed (:108)
        y.backward(torch.from_numpy(dy).to(dt))
        cv = np32 if prec == "f32" else np64
        out["y_" + prec] = cv(y)
        out["dx_" + prec] = cv(xt.grad)
        out["dg_w_" + prec] = cv(sn.g_fc.weight.grad)[:, 0, :]
        out["dg_gamma_" + prec] = cv(sn.g_bn.weight.grad)
        out["dg_beta_" + prec] = cv(sn.g_bn.bias.grad)
        out["g_rm_after_" + prec] = cv(sn.g_bn.running_mean)
        out["g_rv_after_" + prec] = cv(sn.g_bn.running_var)
    torch.manual_seed(seed + 1)
    np.random.seed(seed + 2)
    out["perm"] = torch.randperm(N).numpy()
    sw = cw = (-1, -1, -1, -1)
    if crop in ("style", "both"):
        b = ref.cn_rand_bbox(x.shape, beta=1, bbx_thres=0.1)
        sw = (int(b[0]), int(b[2]), int(b[1]), int(b[3]))
    out["chan_perm"] = np.zeros(0, np.int64)
    if crop in ("content", "both"):
        b = ref.cn_rand_bbox(x.shape, beta=1, bbx_thres=0.1)
        cw = (int(b[0]), int(b[2]), int(b[1]), int(b[3]))
    out["style_window"] = np.array(sw)
    out["content_window"] = np.array(cw)
    return out


def stats_case(shape, seed, eps):
    x = varied_input(shape, seed=seed, dtype=np.float32, relu=True)
    out = {"x": x, "eps": np.array(eps)}
    for prec, dt in (("f32", torch.float32), ("f64", torch.float64)):
        m, s = ref.calc_ins_mean_std(torch.from_numpy(x).to(dt), eps=eps)
        out["mean_" + prec] = m.numpy()[:, :, 0, 0].copy()
        out["std_" + prec] = s.numpy()[:, :, 0, 0].copy()
    return out


def rng_stream_case(seed, shapes_crops):
    """A sequence of consecutive cn_op calls from ONE seeded RNG state, as in a training step
    where several CrossNorm sites fire: pins the host-side draw order across calls."""
    torch.manual_seed(seed)
    np.random.seed(seed + 1)
    rows = []
    for shape, crop in shapes_crops:
        perm = torch.randperm(shape[0]).numpy()
        sw = cw = (-1, -1, -1, -1)
        if crop in ("style", "both"):
            b = ref.cn_rand_bbox(shape, beta=1, bbx_thres=0.1)
            sw = (int(b[0]), int(b[2]), int(b[1]), int(b[3]))
        if crop in ("content", "both"):
            b = ref.cn_rand_bbox(shape, beta=1, bbx_thres=0.1)
            cw = (int(b[0]), int(b[2]), int(b[1]), int(b[3]))
        rows.append((perm, np.array(sw), np.array(cw)))
    out = {"seed": np.array(seed), "n": np.array(len(rows))}
    for i, ((shape, crop), (perm, sw, cw)) in enumerate(zip(shapes_crops, rows)):
        out[f"shape{i}"] = np.array(shape)
        out[f"crop{i}"] = np.array(crop)
        out[f"perm{i}"] = perm
        out[f"sw{i}"] = sw
        out[f"cw{i}"] = cw
    return out


def main():
    cases = {}
    # BASELINE config 1: SelfNorm (4,16,8,8) fp32, literal randn and the varied distribution
    cases["selfnorm_cfg1_randn"] = selfnorm_case((4, 16, 8, 8), 0, True, False, "randn")
    cases["selfnorm_cfg1_varied"] = selfnorm_case((4, 16, 8, 8), 0, True, False, "varied")
    cases["selfnorm_eval"] = selfnorm_case((4, 16, 8, 8), 1, False, False, "varied")
    cases["selfnorm_two_train"] = selfnorm_case((6, 10, 7, 7), 2, True, True, "varied")
    cases["selfnorm_two_eval"] = selfnorm_case((6, 10, 7, 7), 3, False, True, "varied")
    cases["selfnorm_odd"] = selfnorm_case((3, 5, 14, 14), 4, True, False, "relu")
    cases["selfnorm_56"] = selfnorm_case((4, 4, 56, 56), 5, True, False, "relu")
    k = 10
    for crop in ("neither", "style", "content", "both"):
        for chan in (False, True):
            for lam in (None, 0.25):
                name = f"crossnorm_{crop}_{'chan' if chan else 'nochan'}_{'lam' if lam else 'nolam'}"
                cases[name] = crossnorm_case((8, 6, 12, 10), k, crop, chan, lam)
                k += 3
    cases["crossnorm_cfg2small_bf16"] = crossnorm_case((8, 4, 32, 32), 70, "neither", False, None, "bf16")
    cases["crossnorm_both_7x7"] = crossnorm_case((8, 16, 7, 7), 80, "both", False, None)
    cases["crossnorm_both_bf16"] = crossnorm_case((8, 8, 16, 16), 90, "both", False, None, "bf16")
    for crop in ("neither", "style", "content", "both"):
        cases[f"site_{crop}"] = site_case((8, 6, 12, 8), 200 + len(crop), crop)
    cases["stats_eps1e-5"] = stats_case((5, 9, 11, 13), 20, 1e-5)
    cases["stats_eps1e-12"] = stats_case((4, 8, 56, 56), 21, 1e-12)
    cases["rng_stream"] = rng_stream_case(1, [((128, 32, 32, 32), "both"), ((128, 64, 16, 16), "both"),
                                              ((128, 128, 8, 8), "style"), ((64, 3, 224, 224), "content"),
                                              ((16, 8, 7, 7), "both"), ((16, 8, 9, 5), "neither")])
    total = 0
    only = sys.argv[1] if len(sys.argv) > 1 else ""       # optional name prefix: regenerate a subset
    for name, d in cases.items():
        if not name.startswith(only):
            continue
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **d)
        total += os.path.getsize(path)
        print(f"{name:45s} {os.path.getsize(path) / 1024:8.1f} KiB")
    print("total", total / 1024, "KiB")


if __name__ == "__main__":
    main()
