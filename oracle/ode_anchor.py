"""Independent anchor for the oracles: the 5-equation model integrated as ODEs, not by its update formulas.

TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT.

oracle/ufair_oracle.py and oracle/ufair_oracle.c both evaluate the closed-form one-step updates
(`R_i <- E c a_i alpha tau_i (1 - e^{-dt/(alpha tau_i)}) + R_i e^{-dt/(alpha tau_i)}` and its thermal
twin) and the closed forms of g_1 / g_0.  They were written from the same statement of the model, so
their agreement says nothing about that statement.  This module goes back to what the closed forms
are solutions OF (reference README.md:6 "5-equation ... impulse response" model; Millar et al. 2017,
README.md:15) and shares no formula with the oracles:

  * pools:     dR_i/dt = a_i c E - R_i / (alpha tau_i)      integrated numerically over each step,
               alpha and E frozen over the step             (mpmath Taylor-series ODE solver, 30 digits,
                                                             or scipy DOP853 at rtol 1e-13)
  * thermal:   dS_j/dt = (q_j F - S_j) / d_j                integrated numerically, F frozen
  * iIRF_h(alpha) = int_0^h sum_i a_i exp(-s / (alpha tau_i)) ds      by numerical QUADRATURE of the
               impulse response (its definition), never by the antiderivative
  * g_1 = d iIRF_h / d ln(alpha) at alpha = 1                by numerical DIFFERENTIATION of that quadrature
  * g_0 = exp(-iIRF_h(1) / g_1)  (the first-order expansion  iIRF_h(alpha) ~ iIRF_h(1) + g_1 ln alpha
               solved for alpha: the README's "derived functional form")
  * Newton mode: Newton's iteration on iIRF_h(alpha) = iIRF with the quadrature and its numerical
               derivative, same safeguard alpha <- max(alpha_new, alpha / 2), same fixed count
  * forcing:   f1 ln(C/C0) + f2 (C - C0) + f3 (sqrt C - sqrt C0) in 30-digit arithmetic

tests/test_oracle_ode_anchor.py holds both oracles to <= 1e-11 of this for 1-3 gases and every alpha mode.
"""
from __future__ import annotations

import numpy as np

GP_A0, GP_TAU0, GP_R0, GP_RU, GP_RT, GP_RA, GP_C0, GP_C, GP_F1, GP_F2, GP_F3 = 0, 4, 8, 9, 10, 11, 12, 13, 14, 15, 16
ALPHA_EXP, ALPHA_SINH, ALPHA_NEWTON, ALPHA_ONE = 0, 1, 2, 3
T_MID, T_END = 0, 1


def _mp():
    import mpmath
    mpmath.mp.dps = 30
    return mpmath


def iirf_quad(alpha, a, tau, h):
    """iIRF_h(alpha) by quadrature of the impulse response sum_i a_i exp(-s/(alpha tau_i))."""
    mp = _mp()
    pts = [0, h]
    # the integrand's scales differ by 10^6 between pools: give the quadrature the e-folding times as breakpoints
    for t in tau:
        for k in (1, 4, 16):
            x = k * alpha * t
            if 0 < x < h:
                pts.append(x)
    pts = sorted(set(pts))
    return mp.quad(lambda s: sum(ai * mp.exp(-s / (alpha * ti)) for ai, ti in zip(a, tau)), pts)


def _deriv(f, x):
    """f'(x) by Richardson-extrapolated central differences (O(step^4); step sized so that neither the
    truncation error nor the quadrature's 1e-26 noise reaches 1e-18 of the result)."""
    mp = _mp()
    st = mp.mpf("1e-5") * max(abs(x), mp.mpf(1))
    d1 = (f(x + st) - f(x - st)) / (2 * st)
    d2 = (f(x + st / 2) - f(x - st / 2)) / st
    return (4 * d2 - d1) / 3


def prep(a, tau, h):
    """(g1, iIRF_h(1)) with g1 = d iIRF_h / d ln alpha at alpha = 1, by numerical differentiation."""
    mp = _mp()
    i1 = iirf_quad(mp.mpf(1), a, tau, h)
    g1 = _deriv(lambda x: iirf_quad(mp.exp(x), a, tau, h), mp.mpf(0))
    return g1, i1


def alpha_of(iirf, a, tau, h, g1, i1, mode, newton_iters):
    mp = _mp()
    if mode == ALPHA_ONE:
        return mp.mpf(1)
    if mode == ALPHA_SINH:
        return mp.sinh(iirf / g1) / mp.sinh(i1 / g1)
    alpha = mp.exp((iirf - i1) / g1)
    if mode == ALPHA_NEWTON:
        for _ in range(newton_iters):
            f = iirf_quad(alpha, a, tau, h) - iirf
            fp = _deriv(lambda x: iirf_quad(x, a, tau, h), alpha)
            an = alpha - f / fp
            alpha = an if an > alpha / 2 else alpha / 2
    return alpha


def _integrate(rhs, y0, dt, solver):
    """y(dt) of y' = rhs(y), y(0) = y0, by a numerical ODE solver (never by the exponential formula)."""
    mp = _mp()
    if solver == "mpmath":
        f = mp.odefun(lambda x, y: rhs(y), 0, [mp.mpf(v) for v in y0])
        return [v for v in f(mp.mpf(dt))]
    from scipy.integrate import solve_ivp
    y0f = np.array([float(v) for v in y0])
    sol = solve_ivp(lambda x, y: np.array([float(v) for v in rhs(list(y))]), (0.0, float(dt)), y0f, method="DOP853",
                    rtol=1e-13, atol=1e-13 * max(1e-3, float(np.max(np.abs(y0f)))))
    return [mp.mpf(float(v)) for v in sol.y[:, -1]]


def run_member(E, gp, tp, *, dt=1.0, f_ext=None, alpha_mode=ALPHA_EXP, newton_iters=0, iirf_max=None, h=100.0,
               t_mode=T_MID, solver="mpmath"):
    """One member.  E [G][n_t] emission rates, gp [G][17], tp [4] -> dict(C [G][n_t], RF [G][n_t], T [n_t])
    as float64 arrays (rounded once, from 30-digit values)."""
    mp = _mp()
    E = np.asarray(E, dtype=np.float64)
    G, n_t = E.shape
    M = lambda x: mp.mpf(float(x))
    dtm, hm = M(dt), M(h)
    gas = []
    for g in range(G):
        p = [M(v) for v in gp[g]]
        a, tau = p[GP_A0:GP_A0 + 4], p[GP_TAU0:GP_TAU0 + 4]
        g1, i1 = prep(a, tau, hm) if alpha_mode != ALPHA_ONE else (None, None)
        gas.append(dict(p=p, a=a, tau=tau, g1=g1, i1=i1, R=[mp.mpf(0)] * 4, Gc=mp.mpf(0)))
    q, d = [M(tp[0]), M(tp[1])], [M(tp[2]), M(tp[3])]
    S = [mp.mpf(0), mp.mpf(0)]
    Tprev = mp.mpf(0)
    C_out, F_out, T_out = np.empty((G, n_t)), np.empty((G, n_t)), np.empty(n_t)
    for t in range(n_t):
        Ftot = mp.mpf(0)
        for g, s in enumerate(gas):
            p, a, tau = s["p"], s["a"], s["tau"]
            e, c, C0 = M(E[g, t]), p[GP_C], p[GP_C0]
            Ga = sum(s["R"]) / c
            iirf = p[GP_R0] + p[GP_RU] * (s["Gc"] - Ga) + p[GP_RT] * Tprev + p[GP_RA] * Ga
            if iirf_max is not None and iirf > iirf_max:
                iirf = M(iirf_max)
            alpha = alpha_of(iirf, a, tau, hm, s["g1"], s["i1"], alpha_mode, newton_iters)
            # pools and cumulative emissions: five coupled-in-name-only ODEs over one step
            y = _integrate(lambda y: [a[i] * c * e - y[i] / (alpha * tau[i]) for i in range(4)] + [e],
                           s["R"] + [s["Gc"]], dtm, solver)
            s["R"], s["Gc"] = y[:4], y[4]
            C = C0 + sum(s["R"])
            F = p[GP_F2] * (C - C0)
            if p[GP_F1] != 0:
                F += p[GP_F1] * mp.log(C / C0)
            if p[GP_F3] != 0:
                F += p[GP_F3] * (mp.sqrt(C) - mp.sqrt(C0))
            C_out[g, t], F_out[g, t] = float(C), float(F)
            Ftot += F
        if f_ext is not None:
            Ftot += M(f_ext[t])
        S_new = _integrate(lambda y: [(q[j] * Ftot - y[j]) / d[j] for j in range(2)], S, dtm, solver)
        T = (sum(S) + sum(S_new)) / 2 if t_mode == T_MID else sum(S_new)
        S, Tprev = S_new, T
        T_out[t] = float(T)
    return dict(C=C_out, RF=F_out, T=T_out)
