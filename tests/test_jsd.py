"""JSD consistency term: oracle vs the reference's own lines executed with PyTorch (CPU), kernel vs oracle (GPU)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import jsd_oracle as J


def reference_lines(lc, l1, l2):
    """imagenet.py:367-376 verbatim in behaviour: softmax, clamped log mixture, three kl_div 'batchmean', mean."""
    pc, p1, p2 = F.softmax(lc, dim=1), F.softmax(l1, dim=1), F.softmax(l2, dim=1)
    pm = torch.clamp((pc + p1 + p2) / 3., 1e-7, 1).log()
    return (F.kl_div(pm, pc, reduction='batchmean') + F.kl_div(pm, p1, reduction='batchmean') +
            F.kl_div(pm, p2, reduction='batchmean')) / 3.


def _logits(B, K, seed, scale):
    rs = np.random.RandomState(seed)
    return [rs.standard_normal((B, K)) * scale for _ in range(3)]


@pytest.mark.parametrize("B,K,scale", [(4, 10, 1.0), (7, 1000, 3.0), (3, 5, 12.0), (2, 33, 0.01)])
def test_oracle_matches_reference_lines(B, K, scale):
    z = _logits(B, K, B * K, scale)
    t = [torch.tensor(v, dtype=torch.float64, requires_grad=True) for v in z]
    loss = reference_lines(*t)
    loss.backward()
    assert abs(float(loss) - J.jsd_fwd(*z)) <= 1e-12 * max(1.0, abs(float(loss)))
    for g, o in zip(t, J.jsd_bwd(*z)):
        assert np.allclose(g.grad.numpy(), o, atol=1e-13, rtol=1e-10)


@pytest.mark.gpu
@pytest.mark.parametrize("B,K,scale", [(4, 10, 1.0), (256, 1000, 3.0), (64, 100, 8.0), (3, 5, 12.0), (130, 257, 0.5)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_jsd_kernel_vs_oracle(B, K, scale, dtype):
    from cnsn_b200.losses import jsd_consistency
    z = [torch.tensor(v).to(dtype) for v in _logits(B, K, B + K, scale)]
    t = [v.cuda().requires_grad_(True) for v in z]
    loss = jsd_consistency(*t)
    (12 * loss).backward()
    zf = [v.double().numpy() for v in z]
    ref = J.jsd_fwd(*zf)
    assert abs(float(loss.detach()) - ref) <= 1e-5 * max(1.0, abs(ref)), (float(loss.detach()), ref)
    tol = 1e-5 if dtype == torch.float32 else 1e-2          # fp32 1e-5 / bf16 1e-2 (BASELINE.json north_star), relative to max|grad|
    for g, o in zip(t, J.jsd_bwd(*zf, gout=12.0)):
        assert g.grad.dtype == dtype
        err = np.abs(g.grad.double().cpu().numpy() - o).max()
        assert err <= tol * max(np.abs(o).max(), 1e-12) + 1e-7, (err, np.abs(o).max())
