"""Both oracles against an INDEPENDENT derivation of the model (oracle/ode_anchor.py): the pool and
thermal ODEs integrated numerically over each step, iIRF_h by quadrature of the impulse response, g_1
by numerical differentiation -- no update formula and no closed form shared with the oracles.

This is what stands in for reference vectors on rows a2-a6 (the reference names the functions,
.coveragerc:12-19, but ships no code for them): the oracles are held to 1e-11 of the
continuous model they claim to solve, for 1-3 gases and every alpha mode.
"""
import numpy as np
import pytest

from fiveeqscm_b200 import params as P
from oracle import c_oracle as co
from oracle import ode_anchor as A
from oracle import ufair_oracle as o

TOL = 1e-11
GASES = {1: ("co2",), 2: ("co2", "ch4"), 3: ("co2", "ch4", "n2o")}


def _inputs(n_gas, n_t, seed, dense):
    ens = P.sample_ensemble(2, n_t=736, seed=seed, gases=GASES[n_gas], dense_pools=dense)
    # a stretch of the scenario with real emissions (the ramp around 2000), not the flat pre-industrial start
    E = P.member_emissions(ens["scen"], ens["scen_idx"], ens["e_scale"])[:, 230:230 + n_t, :]
    return np.ascontiguousarray(E), ens["gas_params"], ens["thermal_params"], ens["f_ext"][230:230 + n_t]


def _check(ref, got, m, what):
    for k in ("C", "RF", "T"):
        a, b = ref[k], got[k][..., m]
        scale = max(np.max(np.abs(a)), 1e-3)
        err = np.max(np.abs(a - b)) / scale
        assert err <= TOL, f"{what}: {k} differs from the ODE anchor by {err:.2e}"


_CASES = [(1, A.ALPHA_EXP, {}), (2, A.ALPHA_EXP, {}), (3, A.ALPHA_EXP, {}), (1, A.ALPHA_SINH, {}), (3, A.ALPHA_SINH, {}),
          (1, A.ALPHA_NEWTON, dict(newton_iters=3)), (3, A.ALPHA_NEWTON, dict(newton_iters=2)), (2, A.ALPHA_ONE, {}),
          (3, A.ALPHA_ONE, {})]


@pytest.mark.parametrize("n_gas,mode,kw", _CASES, ids=lambda v: str(v) if not isinstance(v, dict) else "k%d" % v.get("newton_iters", 0))
def test_oracles_solve_the_continuous_model(n_gas, mode, kw):
    """30-digit Taylor-series ODE integration (mpmath.odefun) of every step, 6 steps."""
    n_t = 6
    E, gp, tp, fx = _inputs(n_gas, n_t, seed=11 + n_gas, dense=(n_gas != 2))
    ref = A.run_member(E[:, :, 0], gp[:, :, 0], tp[:, 0], f_ext=fx, alpha_mode=mode, solver="mpmath", **kw)
    got_np = o.oxfair(E, gp, tp, f_ext=fx, alpha_mode=mode, **kw)
    got_c = co.oxfair(E, gp, tp, f_ext=fx, alpha_mode=mode, **kw)
    _check(ref, got_np, 0, "numpy oracle")
    _check(ref, got_c, 0, "C oracle")
    _check(ref, co.oxfair(E, gp, tp, f_ext=fx, alpha_mode=mode, blocked=True, **kw), 0, "blocked C baseline")


@pytest.mark.parametrize("t_mode", [A.T_MID, A.T_END], ids=["mid", "end"])
def test_oracles_solve_the_continuous_model_longer_run(t_mode):
    """scipy DOP853 (rtol 1e-13) over 40 steps with an iIRF ceiling that binds, dt = 0.5."""
    n_t = 40
    E, gp, tp, fx = _inputs(3, n_t, seed=5, dense=True)
    kw = dict(dt=0.5, f_ext=fx, alpha_mode=A.ALPHA_EXP, iirf_max=36.0, t_mode=t_mode)
    ref = A.run_member(E[:, :, 1], gp[:, :, 1], tp[:, 1], solver="scipy", **kw)
    _check(ref, o.oxfair(E, gp, tp, **kw), 1, "numpy oracle")
    _check(ref, co.oxfair(E, gp, tp, **kw), 1, "C oracle")


def test_g1_g0_closed_forms_are_the_expansion_of_the_quadrature():
    """g_1, g_0 (.coveragerc:15-16) against quadrature + numerical differentiation."""
    gp, _ = P.default_params(1)
    a, tau = gp[0, 0:4, 0], gp[0, 4:8, 0]
    g1, i1 = A.prep([A._mp().mpf(float(x)) for x in a], [A._mp().mpf(float(x)) for x in tau], A._mp().mpf(100))
    assert abs(float(g1) - o.g_1(a[:, None], tau[:, None])[0]) <= 1e-12 * float(g1)
    g0 = float(A._mp().exp(-i1 / g1))
    assert abs(g0 - o.g_0(a[:, None], tau[:, None])[0]) <= 1e-12 * g0
