"""Losses of the reference's training steps that sit next to the CNSN hot path."""
from .functional import JsdConsistencyFn

__all__ = ["jsd_consistency"]


def jsd_consistency(logits_clean, logits_aug1, logits_aug2):
    """``consist_loss`` of imagenet.py:367-376 / cifar.py:173-182: the mean of the three KL divergences between the
    clamped mixture of the softmaxes and each view, 'batchmean' reduction.  One CUDA kernel forward, one backward;
    the caller adds ``12 * jsd_consistency(...)`` to the clean cross-entropy as the reference does.  Returns a
    float32 scalar tensor."""
    assert logits_clean.dim() == 2 and logits_clean.shape == logits_aug1.shape == logits_aug2.shape
    assert logits_clean.dtype == logits_aug1.dtype == logits_aug2.dtype
    return JsdConsistencyFn.apply(logits_clean, logits_aug1, logits_aug2)
