"""Parameter tables and synthetic inputs for the Universal-FaIR ensemble path.

The reference ships NO parameter values ("Appropriate tunings and parameter sets will be made
available in due course", reference README.md:10), so everything here is a clearly-labelled
ILLUSTRATIVE literature-style default set (AR5-IR / Millar et al. 2017 pool structure,
reference README.md:15-17) -- good for tests and benchmarks, not a calibrated tuning.  Parity is
always defined on identical inputs, so these numbers are not parity-relevant.

Host side only (numpy); arrays come out in the layouts include/ufair.h documents.
"""
from __future__ import annotations

import math

import numpy as np

from . import _abi

GASES = ("co2", "ch4", "n2o")

# a[4], tau[4] (yr), r0, rU, rT, rA, C0, emis2conc, f1, f2, f3
_DEFAULTS = {
    # CO2: ppm, GtC
    "co2": dict(a=[0.2173, 0.2240, 0.2824, 0.2763], tau=[1.0e6, 394.4, 36.54, 4.304],
                r0=32.4, rU=0.019, rT=4.165, rA=0.0, C0=278.0, c=0.469,
                f=[3.74 / math.log(2.0), 0.0, 0.0]),
    # CH4: ppb, MtCH4 -- one active pool; r0 set so that alpha(pre-industrial) = 1
    "ch4": dict(a=[1.0, 0.0, 0.0, 0.0], tau=[9.15, 1.0, 1.0, 1.0],
                r0=None, rU=0.0, rT=-0.30, rA=3.2e-4, C0=720.0, c=0.352,
                f=[0.0, 0.0, 0.036]),
    # N2O: ppb, MtN2O-N
    "n2o": dict(a=[1.0, 0.0, 0.0, 0.0], tau=[116.0, 1.0, 1.0, 1.0],
                r0=None, rU=0.0, rT=0.0, rA=-1.1e-3, C0=270.0, c=0.201,
                f=[0.0, 0.0, 0.12]),
    # generic one-box HFC-style gas (the reference's only shipped case: unit lifetime,
    # concentration == burden, no forcing; U_FaIR/concentrations.py:4-5)
    "hfc": dict(a=[1.0, 0.0, 0.0, 0.0], tau=[1.0, 1.0, 1.0, 1.0],
                r0=None, rU=0.0, rT=0.0, rA=0.0, C0=0.0, c=1.0, f=[0.0, 1.0, 0.0]),
}
THERMAL_DEFAULT = dict(q=[0.33, 0.41], d=[239.0, 4.1])
F2X = 3.74


def default_gas_row(gas: str) -> np.ndarray:
    """One gas's 17 raw parameters in UFAIR_GP_* order."""
    p = _DEFAULTS[gas]
    a = np.array(p["a"], dtype=np.float64)
    tau = np.array(p["tau"], dtype=np.float64)
    r0 = p["r0"]
    if r0 is None:  # iIRF100 at alpha = 1
        r0 = float(np.sum(a * tau * (-np.expm1(-100.0 / tau))))
    row = np.empty(_abi.GP_COUNT, dtype=np.float64)
    row[_abi.GP_A0:_abi.GP_A0 + 4] = a
    row[_abi.GP_TAU0:_abi.GP_TAU0 + 4] = tau
    row[_abi.GP_R0], row[_abi.GP_RU], row[_abi.GP_RT], row[_abi.GP_RA] = r0, p["rU"], p["rT"], p["rA"]
    row[_abi.GP_C0], row[_abi.GP_EMIS2CONC] = p["C0"], p["c"]
    row[_abi.GP_F1:_abi.GP_F3 + 1] = p["f"]
    return row


def default_params(n_member: int = 1, gases=GASES):
    """(gas_params [G][17][M], thermal_params [4][M]) with every member at the defaults."""
    gp = np.stack([default_gas_row(g) for g in gases])[:, :, None] * np.ones((1, 1, n_member))
    tp = np.array(THERMAL_DEFAULT["q"] + THERMAL_DEFAULT["d"], dtype=np.float64)[:, None] * np.ones((1, n_member))
    return np.ascontiguousarray(gp), np.ascontiguousarray(tp)


def sample_params(n_member: int, rng: np.random.Generator, gases=GASES, *, dense_pools: bool = False):
    """Perturbed-parameter ensemble: tau, r0, q, d, f x lognormal(sigma=0.1);
    a x lognormal(0.1) renormalised to sum 1; rU, rT, rA x normal(1, 0.13).

    dense_pools=True gives every gas four active pools (CH4/N2O get a CO2-like split of their
    single lifetime) and three non-zero forcing coefficients, so the general path is exercised
    with nothing skippable (this is the benchmark's workload).
    """
    G, M = len(gases), n_member
    gp, tp = default_params(M, gases)
    if dense_pools:
        for g, name in enumerate(gases):
            if name != "co2":
                tau1 = _DEFAULTS[name]["tau"][0]
                gp[g, _abi.GP_A0:_abi.GP_A0 + 4] = np.array([0.55, 0.25, 0.15, 0.05])[:, None]
                gp[g, _abi.GP_TAU0:_abi.GP_TAU0 + 4] = (tau1 * np.array([1.0, 0.5, 0.2, 0.05]))[:, None]
                a = gp[g, _abi.GP_A0:_abi.GP_A0 + 4, 0]
                tau = gp[g, _abi.GP_TAU0:_abi.GP_TAU0 + 4, 0]
                gp[g, _abi.GP_R0] = float(np.sum(a * tau * (-np.expm1(-100.0 / tau))))
        # ... and every gas all three forcing terms (log, linear, sqrt), so no term is skippable
        dense_f = {"co2": [3.74 / math.log(2.0), 1.0e-4, 0.05], "ch4": [0.02, 1.0e-5, 0.036],
                   "n2o": [0.02, 1.0e-4, 0.12], "hfc": [0.01, 1.0, 0.01]}
        for g, name in enumerate(gases):
            gp[g, _abi.GP_F1:_abi.GP_F3 + 1] = np.array(dense_f[name])[:, None]
            if gp[g, _abi.GP_C0, 0] == 0.0:
                gp[g, _abi.GP_C0] = 1.0
    ln = lambda shape: np.exp(0.1 * rng.standard_normal(shape))
    nm = lambda shape: 1.0 + 0.13 * rng.standard_normal(shape)
    gp[:, _abi.GP_TAU0:_abi.GP_TAU0 + 4] *= ln((G, 4, M))
    a = gp[:, _abi.GP_A0:_abi.GP_A0 + 4] * ln((G, 4, M))
    gp[:, _abi.GP_A0:_abi.GP_A0 + 4] = a / a.sum(axis=1, keepdims=True)
    gp[:, _abi.GP_R0] *= ln((G, M))
    gp[:, _abi.GP_RU] *= nm((G, M))
    gp[:, _abi.GP_RT] *= nm((G, M))
    gp[:, _abi.GP_RA] *= nm((G, M))
    gp[:, _abi.GP_F1:_abi.GP_F3 + 1] *= ln((G, 3, M))
    tp *= ln((4, M))
    return gp, tp


def scenario_emissions(n_t: int = 736, dt: float = 1.0, start_year: float = 1765.0, gases=GASES):
    """Synthetic emission-rate scenarios [G][n_t][4].

    Logistic historical ramp to ~`peak` by 2020, then: (0) linear to 0 by 2100; (1) constant;
    (2) +1 %/yr to 2100 then linear to 0 by 2200; (3) linear to -0.2*peak by 2100 then constant.
    CO2 peak 10 GtC/yr, CH4 350 Mt/yr, N2O 7 MtN/yr, HFC 1.
    """
    peak = {"co2": 10.0, "ch4": 350.0, "n2o": 7.0, "hfc": 1.0}
    yr = start_year + dt * np.arange(n_t)
    hist = 1.0 / (1.0 + np.exp(-(yr - 1970.0) / 22.0))
    hist = hist / (1.0 / (1.0 + np.exp(-(2020.0 - 1970.0) / 22.0)))
    shape = np.empty((n_t, 4))
    fut = yr > 2020.0
    x = np.clip((yr - 2020.0) / 80.0, 0.0, 1.0)
    shape[:, 0] = np.where(fut, 1.0 - x, hist)
    shape[:, 1] = np.where(fut, 1.0, hist)
    grow = 1.01 ** np.clip(yr - 2020.0, 0.0, 80.0)
    down = np.clip(1.0 - (yr - 2100.0) / 100.0, 0.0, 1.0)
    shape[:, 2] = np.where(fut, grow * down, hist)
    shape[:, 3] = np.where(fut, 1.0 - 1.2 * x, hist)
    return np.ascontiguousarray(np.stack([peak[g] * shape for g in gases]))


def member_emissions(scen: np.ndarray, scen_idx: np.ndarray, e_scale: np.ndarray) -> np.ndarray:
    """Expand scenario emissions [G][n_t][S] to per-member [G][n_t][M]: E = scen[.., idx] * scale."""
    return np.ascontiguousarray(scen[:, :, scen_idx] * e_scale[:, None, :])


def sample_ensemble(n_member: int, n_t: int = 736, dt: float = 1.0, seed: int = 20261018,
                    gases=GASES, dense_pools: bool = False):
    """Everything a run needs, from one seed: dict(gas_params, thermal_params, scen, scen_idx,
    e_scale, f_ext) -- the identical bits go to the oracle and to the device in parity tests."""
    rng = np.random.default_rng(seed)
    gp, tp = sample_params(n_member, rng, gases, dense_pools=dense_pools)
    scen = scenario_emissions(n_t, dt, gases=gases)
    scen_idx = rng.integers(0, scen.shape[2], size=n_member).astype(np.int32)
    e_scale = 1.0 + 0.05 * rng.standard_normal((len(gases), n_member))
    yr = 1765.0 + dt * np.arange(n_t)
    f_ext = 0.1 * np.sin(2.0 * np.pi * (yr - 1765.0) / 11.0) - 0.2 * np.exp(-((yr - 1991.0) / 1.5) ** 2)
    return dict(gas_params=gp, thermal_params=tp, scen=scen, scen_idx=scen_idx, e_scale=e_scale,
                f_ext=np.ascontiguousarray(f_ext))


# ------------------------------------------------------------------------------------------------
# on-device sampler (include/ufair.h "on-device ensemble sampler")
# ------------------------------------------------------------------------------------------------
def sampler_tables(gases=GASES, *, dense_pools: bool = False):
    """The perturbed-parameter recipe of :func:`sample_params` as sampler tables:
    dict(gas_base [G][17], gas_sigma, gas_dist, thermal_base [4], thermal_sigma, thermal_dist,
    e_scale_sigma) -- tau, a, r0, f, q, d lognormal(0.1); rU, rT, rA normal(1, 0.13); C0 and the
    emission-to-concentration factor fixed; emission scale normal(1, 0.05)."""

    class _Zero:  # the unperturbed table: sample_params with every normal draw = 0
        def standard_normal(self, shape):
            return np.zeros(shape)

    gp, tp = sample_params(1, _Zero(), gases, dense_pools=dense_pools)
    G = len(gases)
    dist = np.zeros((G, _abi.GP_COUNT), dtype=np.uint8)
    sigma = np.zeros((G, _abi.GP_COUNT))
    for rows, d, s in (((_abi.GP_A0, _abi.GP_A0 + 8), _abi.DIST_LOGNORMAL, 0.1),      # a_1..4, tau_1..4
                       ((_abi.GP_R0, _abi.GP_R0 + 1), _abi.DIST_LOGNORMAL, 0.1),
                       ((_abi.GP_RU, _abi.GP_RA + 1), _abi.DIST_NORMAL, 0.13),
                       ((_abi.GP_F1, _abi.GP_F3 + 1), _abi.DIST_LOGNORMAL, 0.1)):
        dist[:, rows[0]:rows[1]] = d
        sigma[:, rows[0]:rows[1]] = s
    return dict(gas_base=np.ascontiguousarray(gp[:, :, 0]), gas_sigma=sigma, gas_dist=dist,
                thermal_base=np.ascontiguousarray(tp[:, 0]), thermal_sigma=np.full(_abi.TP_COUNT, 0.1),
                thermal_dist=np.full(_abi.TP_COUNT, _abi.DIST_LOGNORMAL, dtype=np.uint8), e_scale_sigma=0.05)


def sample_on_device(n_member: int, seed: int, *, first_member: int = 0, n_scen: int = 4, tables=None, gases=GASES,
                     dense_pools: bool = False, precision: str = "f64", device=None):
    """Members first_member .. first_member + n_member - 1 of the (seed, tables) ensemble, generated
    on the GPU: (gas_params [G][17][M], thermal_params [4][M], e_scale [G][M], scen_idx [M] int32)
    as CUDA tensors.  A member's values depend only on the seed and its global index, so ranks that
    take consecutive member blocks hold exactly the ensemble a single call would produce."""
    import ctypes as C

    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("fiveeqscm_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    t = tables if tables is not None else sampler_tables(gases, dense_pools=dense_pools)
    G = int(np.asarray(t["gas_base"]).shape[0])
    sp = _abi.UfairSampler(n_gas=G, seed=int(seed) & (2 ** 64 - 1), n_scen=int(n_scen),
                           e_scale_sigma=float(t["e_scale_sigma"]))
    for g in range(G):
        for r in range(_abi.GP_COUNT):
            sp.gas_base[g][r] = float(t["gas_base"][g][r])
            sp.gas_sigma[g][r] = float(t["gas_sigma"][g][r])
            sp.gas_dist[g][r] = int(t["gas_dist"][g][r])
    for k in range(_abi.TP_COUNT):
        sp.thermal_base[k], sp.thermal_sigma[k] = float(t["thermal_base"][k]), float(t["thermal_sigma"][k])
        sp.thermal_dist[k] = int(t["thermal_dist"][k])
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    dtype = torch.float64 if precision == "f64" else torch.float32
    M = int(n_member)
    ld = (M + 3) // 4 * 4                                   # 16-byte rows in either precision
    gp = torch.empty(G, _abi.GP_COUNT, ld, dtype=dtype, device=dev)
    tp = torch.empty(_abi.TP_COUNT, ld, dtype=dtype, device=dev)
    esc = torch.empty(G, ld, dtype=dtype, device=dev)
    scen = torch.empty(max(M, 1), dtype=torch.int32, device=dev)
    if ld != M:
        gp.zero_(); tp.zero_(); esc.zero_()
    fn = _abi.lib().ufair_sample_f64 if precision == "f64" else _abi.lib().ufair_sample_f32
    with torch.cuda.device(dev):
        _abi.check(fn(C.byref(sp), int(first_member), M, ld, gp.data_ptr(), tp.data_ptr(), esc.data_ptr(),
                      scen.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
    return gp[..., :M], tp[..., :M], esc[..., :M], scen[:M]
