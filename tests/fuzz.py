"""Seeded random configurations checked against the C oracle: every combination of gas count,
member / step counts (ragged tiles and warps), emission layout, external-forcing layout, alpha mode,
clamp, temperature mode, resume state, concentration-driven gases, sparse or dense parameters
(specialised or general kernel) that the dedicated tests only sample by hand.  Shared by the GPU
test (the CUDA path) and a CPU test (the numpy oracle through the same harness)."""
import os

import numpy as np

from oracle import c_oracle as co
from oracle import ufair_oracle as o
from tests.util import ensemble, field_relerr

TOL64 = 1e-10
GASES = ("co2", "ch4", "n2o", "hfc")
_AM = {"exp": o.ALPHA_EXP, "sinh": o.ALPHA_SINH, "newton": o.ALPHA_NEWTON, "one": o.ALPHA_ONE}
N_CASES = int(os.environ.get("UFAIR_FUZZ_CASES", "40"))   # a one-off wider sweep: UFAIR_FUZZ_CASES=400 pytest ...
# Forcing and temperature of a run that lasts one or two steps from pre-industrial are ~1e-6 W m-2 / K,
# and f1 ln(C / C0) at C = C0 (1 + 1e-7) is ill-conditioned in float64 whoever evaluates it (one
# rounding of the argument is 1e-16 / 1e-7 = 1e-9 of the result; the kernel multiplies by 1/C0 where
# the oracle divides).  The 1e-10 criterion is therefore applied to max(field scale, 0.01 W m-2 or K):
# an absolute 1e-12 W m-2 / K for such runs, the plain relative criterion for every run of normal length.
FLOOR = {"RF": 1e-2, "T": 1e-2}


def _case(seed):
    """The configuration a seed stands for."""
    r = np.random.default_rng(1000 + seed)
    G = int(r.integers(1, 5))
    M = int(r.choice([1, 2, 7, 31, 32, 33, 100, 321, 640, 1000, 2049]))
    n_t = int(r.choice([1, 2, 7, 8, 9, 16, 17, 50, 129]))
    c = dict(G=G, M=M, n_t=n_t, dense=bool(r.integers(0, 2)), dt=float(r.choice([1.0, 0.25, 0.1])),
             alpha_mode=str(r.choice(["exp", "exp", "sinh", "newton", "one"])), newton_iters=int(r.integers(0, 4)),
             iirf_max=[None, 97.0, 40.0][int(r.integers(0, 3))], t_mode=str(r.choice(["mid", "end"])),
             e_layout=str(r.choice(["member", "scenario", "scenario+scale"])),
             fext=str(r.choice(["none", "shared", "scenario", "member"])), resume=bool(r.integers(0, 2)),
             driven=int(r.integers(0, 1 << G)) if r.random() < 0.3 else 0, stats=bool(r.integers(0, 2)))
    return c


def check_case(seed, run_ensemble, HistSpec, to_dev, to_np):
    """Build configuration `seed`, run it through `run_ensemble` (the public API's signature) and
    check every output against the C oracle."""
    c = _case(seed)
    G, M, n_t = c["G"], c["M"], c["n_t"]
    ens = ensemble(M, n_t=n_t, dt=c["dt"], dense=c["dense"], gases=GASES[:G], seed=seed)
    gp, tp = ens["gas_params"], ens["thermal_params"]
    kw, okw = {}, {}
    if c["e_layout"] == "member":
        E = ens["E"]
    else:
        E = ens["scen"]
        kw["scen_idx"], okw["scen_idx"] = to_dev(ens["scen_idx"]), ens["scen_idx"]
        if c["e_layout"] == "scenario+scale" and not c["driven"]:
            kw["e_scale"], okw["e_scale"] = to_dev(ens["e_scale"]), ens["e_scale"]
    S = ens["scen"].shape[2]
    if c["fext"] == "shared":
        fx = ens["f_ext"]
    elif c["fext"] == "scenario" and c["e_layout"] != "member":
        fx = np.stack([ens["f_ext"] * (1 + 0.2 * s) for s in range(S)], axis=1)
    elif c["fext"] == "scenario":                           # per-member emissions: one shared series
        fx = ens["f_ext"]
    elif c["fext"] == "member":
        fx = ens["f_ext"][:, None] * (1 + 0.01 * np.arange(M))[None, :]
        kw["fext_per_member"] = okw["fext_per_member"] = True
    if c["fext"] != "none":
        kw["f_ext"], okw["f_ext"] = to_dev(fx), fx
    modes = dict(alpha_mode=c["alpha_mode"], newton_iters=c["newton_iters"], iirf_max=c["iirf_max"], t_mode=c["t_mode"])
    omodes = dict(modes, alpha_mode=_AM[c["alpha_mode"]], t_mode=o.T_END if c["t_mode"] == "end" else o.T_MID)
    state = None
    if c["resume"]:   # a spun-up state from a short forward run of the same members
        state = co.oxfair(ens["E"][:, : max(1, n_t // 2)], gp, tp, dt=c["dt"], **{k: v for k, v in omodes.items()})["state"]
        kw["state_in"], okw["state_in"] = to_dev(state), state
    if c["driven"]:   # concentration-driven gases: feed them the concentrations of a forward run
        fwd = co.oxfair(E, gp, tp, dt=c["dt"], **okw, **omodes)
        Cf = fwd["C"]
        E = np.array(E)
        for g in range(G):
            if (c["driven"] >> g) & 1:
                if c["e_layout"] == "member":
                    E[g] = Cf[g]
                else:       # scenario-shared pathways: take each scenario's first member as its pathway
                    first = [int(np.argmax(ens["scen_idx"] == s)) if np.any(ens["scen_idx"] == s) else 0 for s in range(S)]
                    E[g] = Cf[g][:, first]
        kw["conc_driven"] = [bool((c["driven"] >> g) & 1) for g in range(G)]
        okw["conc_driven"] = c["driven"]
    spec = HistSpec(lo=-1.0, hi=6.0, bins=64, copies=3) if c["stats"] else None
    res = run_ensemble(to_dev(E), to_dev(gp), to_dev(tp), dt=c["dt"], stats=spec,
                       outputs=("C", "RF", "T", "alpha"), **kw, **modes)
    ref = co.oxfair(E, gp, tp, dt=c["dt"], want_alpha=True, **okw, **omodes)
    for k in ("C", "RF", "T", "alpha", "state"):
        err = field_relerr(to_np(getattr(res, k)), ref[k], FLOOR.get(k, 0.0))
        assert err < TOL64, f"{c}: {k} relative error {err:.3e}"
    if c["driven"]:
        for g in range(G):
            assert field_relerr(to_np(res.E)[g], ref["E"][g]) < 1e-9, f"{c}: E[{g}]"
    if spec is not None:
        T = to_np(res.T)
        hist, mom = co.temperature_stats(T, spec.lo, spec.hi, spec.bins)
        assert np.array_equal(to_np(res.hist), hist.astype(np.int64)), f"{c}: histogram"
        assert np.allclose(to_np(res.moments), mom, rtol=1e-12, atol=1e-11), f"{c}: moments"
