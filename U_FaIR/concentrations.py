"""``U_FaIR.concentrations`` -- same module path and function name as the reference
(U_FaIR/concentrations.py:4), served by the B200 implementation."""
from fiveeqscm_b200.concentrations import (DevicePlan, EnsembleResult, HistSpec, Workspace,  # noqa: F401
                                           calculate_hfc_conc, pinned_result, run_ensemble)

__all__ = ["calculate_hfc_conc", "run_ensemble", "HistSpec", "EnsembleResult", "Workspace"]
