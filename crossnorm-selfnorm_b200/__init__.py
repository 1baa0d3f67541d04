"""cnsn_b200 -- B200-native CrossNorm / SelfNorm (drop-in for the reference's models/cnsn.py).

The package directory is ``crossnorm-selfnorm_b200/`` (not an importable identifier); import it as
``cnsn_b200`` through the shim at the repository root.
"""
from . import _lib
from .cnsn import (CNSN, CrossNorm, SelfNorm, calc_ins_mean_std, cn_op_2ins_space_chan,
                   cn_rand_bbox, instance_norm_mix)

__all__ = ["CNSN", "CrossNorm", "SelfNorm", "calc_ins_mean_std", "cn_op_2ins_space_chan",
           "cn_rand_bbox", "instance_norm_mix", "library_path", "launch_count"]


def library_path():
    return _lib.LIB_PATH


def launch_count():
    """Kernels launched by libcnsn_b200.so in this process so far."""
    return _lib.launch_count()
