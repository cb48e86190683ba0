"""Generate tests/golden/hfc_reference.npz by RUNNING THE REFERENCE (only works in the build
container, where /root/reference exists).  TEST INFRASTRUCTURE.

The reference ships exactly one function on this path,
``calculate_hfc_conc`` (/root/reference/U_FaIR/concentrations.py:4-5).  It is loaded by file path
(the package name U_FaIR is also used by this repo's drop-in package) and evaluated on the
reference's own test inputs (tests/unit/test_hfcs.py:6-10) plus the edge cases its TODOs name
(tests/unit/test_hfcs.py:15-16) and a [t, member] broadcast.  Inputs and outputs are stored so
the GPU box (which has no /root/reference) can check against them.

    python oracle/gen_golden.py
"""
import importlib.util
import os

import numpy as np

REF = "/root/reference/U_FaIR/concentrations.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "hfc_reference.npz")


def main():
    spec = importlib.util.spec_from_file_location("ref_concentrations", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    rng = np.random.default_rng(20261018)
    cases = {}

    def add(name, emissions, time, lifetime):
        cases[f"{name}__emissions"] = emissions
        cases[f"{name}__time"] = time
        cases[f"{name}__lifetime"] = np.float64(lifetime)
        cases[f"{name}__expected"] = ref.calculate_hfc_conc(emissions, time, lifetime)

    # the reference's own test (tests/unit/test_hfcs.py:6-10), integer inputs and all
    add("ref_test", np.array([10, 0, 0, 0]), np.array([0, 1, 2, 3]), 1.0)
    # `lifetime` is accepted but ignored (U_FaIR/concentrations.py:4 vs :5)
    add("lifetime_ignored", np.array([10, 0, 0, 0]), np.array([0, 1, 2, 3]), 50.0)
    # TODO in the reference's test: pulse not in year zero -> only emissions[0] is read -> zeros
    add("pulse_year1", np.array([0.0, 10.0, 0.0, 0.0]), np.array([0.0, 1.0, 2.0, 3.0]), 1.0)
    # TODO in the reference's test: constant emissions -> still emissions[0] * exp(-t)
    add("constant", np.full(6, 2.5), np.linspace(0.0, 5.0, 6), 1.0)
    # irregular float times, negative times
    add("irregular", np.array([3.25, 1.0, 7.0]), np.array([-1.5, 0.1, 12.75]), 2.0)
    # annual grid 1765..2500 as offsets
    add("annual736", np.concatenate([[1.0], np.zeros(735)]), np.arange(736, dtype=np.float64) * 0.05, 1.0)
    # [t, member] broadcast: the seed of the [gas][t][member] layout
    add("broadcast_tm", rng.uniform(0.0, 20.0, size=(16, 8)), np.linspace(0.0, 7.5, 16)[:, None], 1.0)
    np.savez_compressed(OUT, **cases)
    print("wrote", os.path.normpath(OUT), "with", len(cases) // 4, "cases")


if __name__ == "__main__":
    main()
