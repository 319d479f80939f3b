"""smm_jl_b200 -- B200-native implementation of SMM.jl's parallel-tempered BGP MCMC hot path.

(The task names the package `smm.jl_b200`; a dot cannot appear in a Python package name, hence the
underscore.)  Layout: csrc/ = hand-written sm_100a CUDA kernels + the C ABI (include/smm_b200.h),
_lib.py = ctypes binding, api.py = host-side mirror of the reference's MProb / Eval / MAlgoBGP surface,
configs.py = the BASELINE.json workloads.
"""
from ._abi import (BGPConfig, Trace, SMM_OBJ_FAILS, SMM_OBJ_NORM, SMM_OBJ_NORM_MV, SMM_OBJ_NORM_SLOW,  # noqa: F401
                   SMM_OBJ_PANEL)

__all__ = ["BGPConfig", "Trace"]
