"""Locate and import the UNMODIFIED reference operators (test infrastructure).

The reference is only present in the build container (/root/reference); it does not travel to
the GPU box, so everything that uses this module must skip when it returns None.
"""
import importlib
import os
import sys

import numpy as np


def reference_root():
    for cand in (os.environ.get("CNSN_REFERENCE"), "/root/reference"):
        if cand and os.path.isfile(os.path.join(cand, "models", "cnsn.py")):
            return cand
    return None


def load_reference_cnsn():
    """Return the reference ``models.cnsn`` module, or None when the reference is absent."""
    root = reference_root()
    if root is None:
        return None
    if not hasattr(np, "int"):          # models/cnsn.py:39-40 uses np.int (removed in NumPy 1.24)
        np.int = int
    name = "_reference_models_cnsn"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(root, "models", "cnsn.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod
