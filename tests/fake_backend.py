"""Oracle-backed stand-in for cnsn_b200._lib.CudaBackend (TEST INFRASTRUCTURE ONLY).

Lets the CPU test-suite drive the package's host logic (module surface, RNG order, autograd
wiring, drop-in into the reference's model files) without a GPU.  The package never installs it.
"""
import numpy as np
import torch

from oracle import cnsn_oracle as O
from oracle import ibn_oracle as IB


def _np(t):
    return t.detach().to(torch.float64).cpu().numpy()


def _t(a, like, dtype=None):
    return torch.from_numpy(np.ascontiguousarray(a)).to(device=like.device, dtype=dtype or like.dtype)


class OracleBackend:
    name = "oracle-fake"

    def __init__(self):
        self.calls = []

    def instance_stats(self, x, window, eps):
        self.calls.append("instance_stats")
        m, s = O.instance_stats(_np(x), eps, tuple(window))
        return _t(m, x, torch.float32), _t(s, x, torch.float32)

    def instance_stats_bwd(self, x, window, mean, std, dmean, dstd):
        self.calls.append("instance_stats_bwd")
        xn = _np(x)
        h0, h1, w0, w1 = window
        Mw = (h1 - h0) * (w1 - w0)
        dx = np.zeros_like(xn)
        mu, sd, dm, ds = (_np(v)[:, :, None, None] for v in (mean, std, dmean, dstd))
        dx[:, :, h0:h1, w0:w1] = dm / Mw + (xn[:, :, h0:h1, w0:w1] - mu) / sd * ds / (Mw - 1)
        return _t(dx, x)

    def instance_affine(self, x, scale, shift):
        self.calls.append("instance_affine")
        return _t(_np(x) * _np(scale)[:, :, None, None] + _np(shift)[:, :, None, None], x)

    def instance_dot(self, x, dy):
        self.calls.append("instance_dot")
        return (_t((_np(x) * _np(dy)).sum((2, 3)), x, torch.float32), _t(_np(dy).sum((2, 3)), x, torch.float32))

    @staticmethod
    def _pb(g, f):
        params, bufs = {}, {}
        for tag, gt in (("g", g), ("f", f)):
            if gt is None:
                continue
            C = gt.gamma.numel()
            params[tag + "_w"] = _np(gt.w).reshape(C, 2)
            params[tag + "_gamma"] = _np(gt.gamma)
            params[tag + "_beta"] = _np(gt.beta)
            if gt.run_mean is not None:
                bufs[tag + "_rm"] = _np(gt.run_mean)
                bufs[tag + "_rv"] = _np(gt.run_var)
        return params, bufs

    def selfnorm_fwd(self, x, g, f, training, momentum, bn_eps, eps):
        self.calls.append("selfnorm_fwd")
        params, bufs = self._pb(g, f)
        y, nb = O.selfnorm_fwd(_np(x), params, bufs, training, eps, bn_eps, momentum)
        if training:
            for tag, gt in (("g", g), ("f", f)):
                if gt is not None:
                    gt.run_mean.copy_(_t(nb[tag + "_rm"], gt.run_mean))
                    gt.run_var.copy_(_t(nb[tag + "_rv"], gt.run_var))
                    if gt.nbt is not None:
                        gt.nbt += 1
        return _t(y, x), {"bufs": bufs}          # "save": the buffers the forward saw

    def selfnorm_bwd(self, x, dy, g, f, training, save):
        self.calls.append("selfnorm_bwd")
        params, _ = self._pb(g, f)
        dx, gr = O.selfnorm_bwd(_np(x), _np(dy), params, save["bufs"], training)
        def pack(tag):
            return (_t(gr[tag + "_w"], x, torch.float32), _t(gr[tag + "_gamma"], x, torch.float32),
                    _t(gr[tag + "_beta"], x, torch.float32))
        return _t(dx, x), pack("g"), pack("f") if f is not None else None

    def selfnorm_block_fwd(self, x, res, relu, g, training, momentum, bn_eps, eps):
        self.calls.append("selfnorm_block_fwd")
        z = x if res is None else (x + res)
        y, save = self.selfnorm_fwd(z, g, None, training, momentum, bn_eps, eps)
        return (torch.relu(y) if relu else y), z, save

    def selfnorm_block_bwd(self, z, dy, relu, g, training, save):
        self.calls.append("selfnorm_block_bwd")
        d = torch.where(z > 0, dy, torch.zeros_like(dy)) if relu else dy
        dz, gg, _ = self.selfnorm_bwd(z, d, g, None, training, save)
        return dz, gg

    def ibn_resident(self, x, half, training):
        return False                                        # the drop-in BatchNorm2d then takes torch's implementation

    def ibn_fwd(self, x, half, p, training, momentum, eps_in, eps_bn, relu=False):
        self.calls.append("ibn_fwd")
        e = np.zeros(0)                                      # half == C: no batch-norm half
        pp = {k: (_np(p[k]) if p.get(k) is not None else e) for k in ("in_w", "in_b", "bn_w", "bn_b")}
        bufs = {"rm": _np(p["run_mean"]) if p.get("run_mean") is not None else e,
                "rv": _np(p["run_var"]) if p.get("run_var") is not None else e}
        y, rm, rv = IB.ibn_fwd(_np(x), half, pp, bufs, training, momentum, eps_in, eps_bn)
        if training and p.get("run_mean") is not None:
            p["run_mean"].copy_(_t(rm, p["run_mean"]))
            p["run_var"].copy_(_t(rv, p["run_var"]))
            if p.get("nbt") is not None:
                p["nbt"] += 1
        mask = (y > 0) if relu else None
        if relu:
            y = np.maximum(y, 0.0)
        return _t(y, x), {"bufs": bufs, "eps": (eps_in, eps_bn), "mask": mask}

    def ibn_bwd(self, x, dy, half, p, training, save, relu=False):
        self.calls.append("ibn_bwd")
        pp = {"in_w": _np(p["in_w"]) if p.get("in_w") is not None else np.zeros(0),
              "bn_w": _np(p["bn_w"]) if p.get("bn_w") is not None else np.zeros(0)}
        dyn = _np(dy)
        if relu:
            dyn = np.where(save["mask"], dyn, 0.0)
        dx, a, b, c, d = IB.ibn_bwd(_np(x), dyn, half, pp, save["bufs"], training, *save["eps"])
        return _t(dx, x), tuple(_t(v, x, torch.float32) for v in (a, b, c, d))

    @staticmethod
    def _plan(x, perm, chan_perm, cwin, swin):
        return {"perm": perm.cpu().numpy().astype(np.int64),
                "chan_perm": None if chan_perm is None else chan_perm.cpu().numpy().astype(np.int64),
                "content_window": tuple(cwin), "style_window": tuple(swin)}

    def crossnorm_fwd(self, x, perm, chan_perm, cwin, swin, lam, eps):
        self.calls.append("crossnorm_fwd")
        y = O.crossnorm_fwd(_np(x), self._plan(x, perm, chan_perm, cwin, swin), lam, eps)
        return _t(y, x), torch.zeros(1)

    def crossnorm_bwd(self, x, dy, perm, chan_perm, cwin, swin, lam, save):
        self.calls.append("crossnorm_bwd")
        return _t(O.crossnorm_bwd(_np(x), _np(dy), self._plan(x, perm, chan_perm, cwin, swin), lam), x)

    # fused site: the oracle's composition (models/cnsn.py:159-164)
    def site_supported(self, x):
        return True

    def site_fwd(self, x, perm, cwin, swin, lam, cn_eps, g, momentum, bn_eps, sn_eps, relu=False):
        self.calls.append("site_fwd")
        z, _ = self.crossnorm_fwd(x, perm, None, cwin, swin, lam, cn_eps)
        self.calls.pop()
        y, save = self.selfnorm_fwd(z, g, None, True, momentum, bn_eps, sn_eps)
        self.calls.pop()
        return (torch.relu(y) if relu else y), save

    def site_bwd(self, x, dy, perm, cwin, swin, lam, cn_eps, g, save, relu=False):
        self.calls.append("site_bwd")
        z, _ = self.crossnorm_fwd(x, perm, None, cwin, swin, lam, cn_eps)
        d = torch.where(z > 0, dy, torch.zeros_like(dy)) if relu else dy
        dz, gg, _ = self.selfnorm_bwd(z, d, g, None, True, save)
        dx = self.crossnorm_bwd(x, dz, perm, None, cwin, swin, lam, None)
        del self.calls[-3:]
        return dx, gg
