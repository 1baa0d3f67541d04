"""numpy float64 restatement of the IBN layer (TEST INFRASTRUCTURE ONLY; never imported by the package).
Follows models/imagenet/resnet_ibn_cnsn.py:24-44 with nn.InstanceNorm2d(affine) / nn.BatchNorm2d semantics; pinned
against that very module composition executed with PyTorch in tests/test_ibn.py."""
import numpy as np


def ibn_fwd(x, half, p, bufs, training=True, momentum=0.1, eps_in=1e-5, eps_bn=1e-5):
    """Returns (y, new running_mean, new running_var)."""
    x = np.asarray(x, np.float64)
    N, C, H, W = x.shape
    y = np.empty_like(x)
    xi = x[:, :half]
    m = xi.mean((2, 3), keepdims=True)
    v = xi.var((2, 3), keepdims=True)
    y[:, :half] = (xi - m) / np.sqrt(v + eps_in) * p["in_w"][None, :, None, None] + p["in_b"][None, :, None, None]
    xb = x[:, half:]
    rm, rv = np.asarray(bufs["rm"], np.float64), np.asarray(bufs["rv"], np.float64)
    if training:
        m = xb.mean((0, 2, 3))
        v = xb.var((0, 2, 3))
        cnt = N * H * W
        rm = (1 - momentum) * rm + momentum * m
        rv = (1 - momentum) * rv + momentum * v * cnt / (cnt - 1)
    else:
        m, v = rm, rv
    y[:, half:] = (xb - m[None, :, None, None]) / np.sqrt(v[None, :, None, None] + eps_bn) * p["bn_w"][None, :, None, None] \
        + p["bn_b"][None, :, None, None]
    return y, rm, rv


def ibn_bwd(x, dy, half, p, bufs, training=True, eps_in=1e-5, eps_bn=1e-5):
    """Returns (dx, d_in_w, d_in_b, d_bn_w, d_bn_b)."""
    x, dy = np.asarray(x, np.float64), np.asarray(dy, np.float64)
    dx = np.empty_like(x)
    xi, di = x[:, :half], dy[:, :half]
    m = xi.mean((2, 3), keepdims=True)
    rs = 1.0 / np.sqrt(xi.var((2, 3), keepdims=True) + eps_in)
    xh = (xi - m) * rs
    g = p["in_w"][None, :, None, None]
    dx[:, :half] = g * rs * (di - di.mean((2, 3), keepdims=True) - xh * (di * xh).mean((2, 3), keepdims=True))
    d_in_w, d_in_b = (di * xh).sum((0, 2, 3)), di.sum((0, 2, 3))
    xb, db = x[:, half:], dy[:, half:]
    g = p["bn_w"][None, :, None, None]
    if training:
        m = xb.mean((0, 2, 3), keepdims=True)
        rs = 1.0 / np.sqrt(xb.var((0, 2, 3), keepdims=True) + eps_bn)
        xh = (xb - m) * rs
        dx[:, half:] = g * rs * (db - db.mean((0, 2, 3), keepdims=True) - xh * (db * xh).mean((0, 2, 3), keepdims=True))
    else:
        m = np.asarray(bufs["rm"], np.float64)[None, :, None, None]
        rs = 1.0 / np.sqrt(np.asarray(bufs["rv"], np.float64)[None, :, None, None] + eps_bn)
        xh = (xb - m) * rs
        dx[:, half:] = g * rs * db
    return dx, d_in_w, d_in_b, (db * xh).sum((0, 2, 3)), db.sum((0, 2, 3))
