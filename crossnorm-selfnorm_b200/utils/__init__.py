"""Host utilities next to the CNSN hot path."""
from .checkpoint import load_reference_checkpoint  # noqa: F401
