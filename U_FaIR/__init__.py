"""Drop-in package name of the reference (stujen/fiveEqSCM ships ``U_FaIR/concentrations.py``;
``from U_FaIR.concentrations import calculate_hfc_conc``, reference tests/unit/test_hfcs.py:3).
The implementation lives in fiveeqscm_b200 (CUDA, sm_100a); this package only re-exports it."""
