"""numpy float64 restatement of the JSD consistency term (TEST INFRASTRUCTURE ONLY; never imported by the
package).  Follows imagenet.py:367-376 / cifar.py:173-182 of the reference; pinned against those very lines
executed with PyTorch in tests/test_jsd.py."""
import numpy as np


def _log_softmax(z):
    z = np.asarray(z, np.float64)
    m = z.max(axis=1, keepdims=True)
    return z - m - np.log(np.exp(z - m).sum(axis=1, keepdims=True))


def jsd_fwd(z0, z1, z2):
    lp = [_log_softmax(z) for z in (z0, z1, z2)]
    p = [np.exp(v) for v in lp]
    lm = np.log(np.clip((p[0] + p[1] + p[2]) / 3.0, 1e-7, 1.0))
    B = p[0].shape[0]
    return sum((pv * (lpv - lm)).sum() for pv, lpv in zip(p, lp)) / (3.0 * B)


def jsd_bwd(z0, z1, z2, gout=1.0):
    lp = [_log_softmax(z) for z in (z0, z1, z2)]
    p = [np.exp(v) for v in lp]
    m = (p[0] + p[1] + p[2]) / 3.0
    lm = np.log(np.clip(m, 1e-7, 1.0))
    ind = ((m >= 1e-7) & (m <= 1.0)).astype(np.float64)
    B = p[0].shape[0]
    out = []
    for pv, lpv in zip(p, lp):
        G = lpv - lm + 1.0 - ind
        c = (pv * G).sum(axis=1, keepdims=True)
        out.append(gout / (3.0 * B) * pv * (G - c))
    return out
