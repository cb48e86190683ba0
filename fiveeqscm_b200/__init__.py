"""fiveeqscm_b200 -- B200-native (sm_100a) ensemble integrator for the 5-equation Universal-FaIR
model, behind the call surface of stujen/fiveEqSCM's ``U_FaIR/concentrations.py``.

Only the hot path lives here: CUDA kernels + C ABI (csrc/, include/ufair.h), the ctypes binding
(_abi), the host-side mirror of the reference interface (concentrations), ensemble statistics and
their cross-GPU reduction (stats, dist) and synthetic inputs (params).  No CPU fallback exists.
"""
__version__ = "0.1.0"
