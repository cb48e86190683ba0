"""Property tests of the oracle (CPU, hypothesis): the behaviours the 5-equation model must have
whatever the parameter values -- they argue the fidelity of the restatement where the reference
pins nothing (SURVEY.md section 4 "what the new repo's test plan must add")."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import c_oracle as co
from oracle import ufair_oracle as o


def _gas(a, tau, r0=30.0, rU=0.02, rT=4.0, rA=0.0, C0=278.0, c=0.47, f=(5.4, 0.0, 0.0), M=1):
    gp = np.zeros((1, o.GP_COUNT, M))
    gp[0, o.GP_A0:o.GP_A0 + 4] = np.asarray(a, dtype=float)[:, None]
    gp[0, o.GP_TAU0:o.GP_TAU0 + 4] = np.asarray(tau, dtype=float)[:, None]
    gp[0, o.GP_R0], gp[0, o.GP_RU], gp[0, o.GP_RT], gp[0, o.GP_RA] = r0, rU, rT, rA
    gp[0, o.GP_C0], gp[0, o.GP_EMIS2CONC] = C0, c
    gp[0, o.GP_F1:o.GP_F3 + 1] = np.asarray(f, dtype=float)[:, None]
    tp = np.array([0.33, 0.41, 239.0, 4.1])[:, None] * np.ones((4, M))
    return gp, tp


fractions = st.lists(st.floats(0.05, 1.0), min_size=4, max_size=4).map(lambda v: np.array(v) / np.sum(v))
lifetimes = st.lists(st.floats(1.0, 2000.0), min_size=4, max_size=4)


@settings(max_examples=40, deadline=None)
@given(a=fractions, tau=lifetimes, e_lo=st.floats(0.0, 8.0), bump=st.floats(0.1, 5.0))
def test_more_emissions_never_cool(a, tau, e_lo, bump):
    """Monotonicity: raising every emission raises C, RF and T at every step (alpha grows with
    cumulative uptake and temperature for positive rU, rT, so the feedback has the same sign)."""
    gp, tp = _gas(a, tau)
    n_t = 120
    lo = co.oxfair(np.full((1, n_t, 1), e_lo), gp, tp)
    hi = co.oxfair(np.full((1, n_t, 1), e_lo + bump), gp, tp)
    for k in ("C", "RF", "T"):
        assert np.all(hi[k] >= lo[k])


@settings(max_examples=40, deadline=None)
@given(a=fractions, tau=lifetimes, scale=st.floats(0.1, 10.0))
def test_alpha_one_is_linear_in_emissions(a, tau, scale):
    """With alpha == 1 the pools are linear: scaling the emissions scales C - C0 (and G_cum)."""
    gp, tp = _gas(a, tau, f=(0.0, 1.0, 0.0))
    rng = np.random.default_rng(1)
    E = rng.uniform(0.0, 6.0, (1, 80, 1))
    x = co.oxfair(E, gp, tp, alpha_mode=o.ALPHA_ONE)
    y = co.oxfair(scale * E, gp, tp, alpha_mode=o.ALPHA_ONE)
    np.testing.assert_allclose(y["C"] - 278.0, scale * (x["C"] - 278.0), rtol=1e-11, atol=1e-11)
    np.testing.assert_allclose(y["state"][4], scale * x["state"][4], rtol=1e-12)


@settings(max_examples=25, deadline=None)
@given(a=fractions, tau=lifetimes)
def test_pools_conserve_mass_without_decay(a, tau):
    """Airborne + taken-up = emitted: sum R_i / c <= G_cum always, and -> G_cum as lifetimes -> inf."""
    gp, tp = _gas(a, tau)
    E = np.full((1, 200, 1), 5.0)
    out = co.oxfair(E, gp, tp)
    Ga = out["state"][0:4].sum(axis=0) / 0.47
    assert np.all(Ga <= out["state"][4] * (1 + 1e-12)) and np.all(Ga > 0)
    gp_inf, _ = _gas(a, [1e12] * 4)
    out = co.oxfair(E, gp_inf, tp, alpha_mode=o.ALPHA_ONE)
    np.testing.assert_allclose(out["state"][0:4].sum(axis=0) / 0.47, out["state"][4], rtol=1e-9)


def test_dt_refinement_converges():
    """Halving dt (emission RATES unchanged) converges: the step is exact for piecewise-constant
    forcing of each sub-system, so the coupled error falls at first order or better."""
    gp, tp = _gas([0.2173, 0.2240, 0.2824, 0.2763], [1e6, 394.4, 36.54, 4.304])
    years = 160
    rate = lambda t: 10.0 / (1.0 + np.exp(-(t - 80.0) / 20.0))
    finals = []
    for dt in (1.0, 0.5, 0.25, 0.125):
        n_t = int(years / dt)
        t_mid = (np.arange(n_t) + 0.5) * dt
        out = co.oxfair(rate(t_mid).reshape(1, n_t, 1), gp, tp, dt=dt, t_mode=o.T_END)
        finals.append((out["C"][0, -1, 0], out["T"][-1, 0]))
    errs = [abs(finals[k][0] - finals[-1][0]) for k in range(3)]
    assert errs[0] > errs[1] > errs[2] and errs[1] < 0.6 * errs[0]
    errsT = [abs(finals[k][1] - finals[-1][1]) for k in range(3)]
    assert errsT[0] > errsT[1] > errsT[2]


@settings(max_examples=25, deadline=None)
@given(q1=st.floats(0.05, 1.0), q2=st.floats(0.05, 1.0), d1=st.floats(50.0, 500.0), d2=st.floats(1.0, 10.0),
       F=st.floats(0.1, 8.0))
def test_thermal_equilibrium_is_ecs_like(q1, q2, d1, d2, F):
    """Constant forcing: T -> F (q1 + q2), monotonically from below."""
    gp, _ = _gas([1, 0, 0, 0], [1, 1, 1, 1], f=(0, 0, 0), C0=1.0)
    tp = np.array([[q1], [q2], [d1], [d2]])
    n_t = 6000
    out = co.oxfair(np.zeros((1, n_t, 1)), gp, tp, alpha_mode=o.ALPHA_ONE, f_ext=np.full(n_t, F), t_mode=o.T_END)
    T = out["T"][:, 0]
    assert np.all(np.diff(T) >= -1e-15)
    np.testing.assert_allclose(T[-1], F * (q1 + q2), rtol=1e-4)
