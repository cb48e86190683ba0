"""The CUDA path (through the C ABI) on the seeded random configurations of tests/fuzz.py."""
import pytest

from fiveeqscm_b200 import _abi
from tests import fuzz
from tests.util import to_dev, to_np

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    import torch
    assert torch.cuda.is_available()
    from fiveeqscm_b200 import concentrations as c
    _abi.lib()
    return c


@pytest.mark.parametrize("seed", range(fuzz.N_CASES))
def test_random_configuration_matches_oracle(api, seed):
    import torch

    def run(*a, **kw):
        r = api.run_ensemble(*a, **kw)
        torch.cuda.synchronize()
        return r

    fuzz.check_case(seed, run, api.HistSpec, to_dev, to_np)
