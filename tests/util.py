"""Shared helpers for the GPU parity tests: the oracle is the checker, the CUDA path (through the
C ABI / ctypes binding) is the thing under test."""
import numpy as np

from fiveeqscm_b200 import params as P


def ensemble(M, n_t=736, dt=1.0, seed=20261018, dense=True, gases=P.GASES):
    ens = P.sample_ensemble(M, n_t=n_t, dt=dt, seed=seed, gases=gases, dense_pools=dense)
    ens["E"] = P.member_emissions(ens["scen"], ens["scen_idx"], ens["e_scale"])
    return ens


def field_relerr(got, ref, floor=0.0):
    """max |got - ref| / max |ref|: error relative to the field's own scale (the 1e-10 criterion).
    `floor` bounds the scale from below for fields that can be arbitrarily small (see tests/fuzz.py)."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    scale = max(float(np.max(np.abs(ref))) if ref.size else 0.0, floor)
    return float(np.max(np.abs(got - ref)) / (scale if scale > 0 else 1.0))


def to_dev(x):
    import torch
    return None if x is None else torch.from_numpy(np.ascontiguousarray(x)).cuda()


def to_np(x):
    return None if x is None else x.detach().cpu().numpy()
