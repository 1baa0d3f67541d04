"""Recipe for ``oracle/_ref/``: an UNMODIFIED copy of the reference's hot-path files, made at build time.

TEST / BASELINE INFRASTRUCTURE ONLY -- the product package never imports anything from here.

The reference (amazon-science/crossnorm-selfnorm) is Python over ATen and has no build system; "building" it means
taking ``models/cnsn.py`` (the hot path) and the two host model files BASELINE configs 3 and 4 name as
they lie under ``/root/reference`` and placing them, byte for byte, under ``oracle/_ref/`` -- git-ignored (so no
reference source enters the history) but NOT gpurun-ignored (so it travels to the GPU box, where ``/root/reference``
does not exist, and ``bench.py --impl reference`` can execute the reference's OWN file on the box's host cores).
A manifest with the SHA-256 of every copied file is written next to them.

    python oracle/build_ref.py            # no-op with a message when /root/reference is absent

``load()`` imports the copy as a namespace package (``models.cnsn`` resolves its relative imports exactly as in the
reference checkout) and applies the one shim NumPy >= 1.24 needs (``np.int``, models/cnsn.py:39-40) to the numpy
module -- not to the file.
"""
import hashlib
import importlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
FILES = ("models/cnsn.py", "models/cifar/wideresnet_cnsn.py", "models/imagenet/resnet_cnsn.py")


def reference_root():
    for cand in (os.environ.get("CNSN_REFERENCE"), "/root/reference"):
        if cand and os.path.isfile(os.path.join(cand, "models", "cnsn.py")):
            return cand
    return None


def build(verbose=True):
    """Copy FILES from the reference checkout into oracle/_ref/.  Returns the directory, or None when the reference
    is not present (the GPU box: the prebuilt copy that travelled with the snapshot is used as is)."""
    root = reference_root()
    if root is None:
        if verbose:
            print("oracle/build_ref: no reference checkout here; keeping %s" % (REF_DIR if available() else "nothing"))
        return REF_DIR if available() else None
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(root, rel), os.path.join(REF_DIR, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(REF_DIR, "MANIFEST.json"), "w") as f:
        json.dump({"source": root, "sha256": manifest}, f, indent=1)
    if verbose:
        print("oracle/build_ref: %d reference files -> %s" % (len(FILES), REF_DIR))
    return REF_DIR


def available():
    return os.path.isfile(os.path.join(REF_DIR, "models", "cnsn.py"))


_LOADED = {}          # plain module name ("models", "models.cnsn", ...) -> module object of the copied tree


def load(name="models.cnsn"):
    """Import a module of the copied reference tree (default: the hot path ``models.cnsn``); None when absent.  The
    tree is imported under its own names (``models`` is a namespace package there, relative imports and all) while
    ``sys.modules`` / ``sys.path`` are switched to it, and taken out again afterwards, so that it never shadows --
    or is shadowed by -- another ``models`` package of the process."""
    if not available():
        return None
    import numpy as np
    if not hasattr(np, "int"):
        np.int = int
    if name in _LOADED:
        return _LOADED[name]

    def ours(k):
        return k == "models" or k.startswith("models.")

    saved_path = list(sys.path)
    saved_mods = {k: v for k, v in sys.modules.items() if ours(k)}
    for k in saved_mods:
        del sys.modules[k]
    sys.modules.update(_LOADED)
    sys.path.insert(0, REF_DIR)
    try:
        mod = importlib.import_module(name)
    finally:
        for k in [k for k in sys.modules if ours(k)]:
            _LOADED[k] = sys.modules.pop(k)
        sys.path[:] = saved_path
        sys.modules.update(saved_mods)
    return mod


if __name__ == "__main__":
    build()
