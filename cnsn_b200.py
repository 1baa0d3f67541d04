"""Import shim: ``import cnsn_b200`` loads the package that lives in ``crossnorm-selfnorm_b200/``.

The directory name is fixed by the project layout and is not a valid Python identifier, so this
module registers that directory as the package ``cnsn_b200`` (sub-modules such as
``cnsn_b200.cnsn`` resolve normally afterwards).
"""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "crossnorm-selfnorm_b200")
_spec = importlib.util.spec_from_file_location(
    "cnsn_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["cnsn_b200"] = _mod
_spec.loader.exec_module(_mod)
