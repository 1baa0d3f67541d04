"""Loading the reference's checkpoints into the host models of this package.

The reference saves either a bare ``state_dict`` (the published ResNet-50 + SN / IBN weights, loaded with
``strict=False`` at imagenet.py:518-521) or a dict with ``state_dict`` / ``optimizer`` / ``epoch`` (cifar.py:415-430,
utils.py ``save_checkpoint``), and -- because it trains under ``nn.DataParallel`` -- every key carries a
``module.`` prefix.  The host models here have the reference's parameter names, so loading is a matter of
unwrapping, stripping the prefix and reporting what did not match.
"""
import torch

__all__ = ["load_reference_checkpoint"]


def load_reference_checkpoint(model, source, strict=False, map_location="cpu"):
    """Load ``source`` (a path or an already loaded object) into ``model``.

    Returns ``(missing_keys, unexpected_keys, extras)``; ``extras`` holds the non-weight entries of a training
    checkpoint (``epoch``, ``best_err1``, ``optimizer`` ...).  ``strict=False`` is the reference's own setting.
    """
    obj = torch.load(source, map_location=map_location) if isinstance(source, (str, bytes)) or hasattr(source, "read") else source
    extras = {}
    if isinstance(obj, dict) and "state_dict" in obj and isinstance(obj["state_dict"], dict):
        extras = {k: v for k, v in obj.items() if k != "state_dict"}
        obj = obj["state_dict"]
    state = {}
    for k, v in obj.items():
        while k.startswith("module."):                 # nn.DataParallel / DistributedDataParallel wrappers
            k = k[len("module."):]
        state[k] = v
    target = model.module if hasattr(model, "module") and isinstance(model.module, torch.nn.Module) else model
    result = target.load_state_dict(state, strict=strict)
    return list(result.missing_keys), list(result.unexpected_keys), extras
