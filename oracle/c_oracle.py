"""ctypes loader for the C oracle (oracle/ufair_oracle.c, ufair_oracle_fast.c).  TEST INFRASTRUCTURE ONLY.

Same call shape as oracle.ufair_oracle.oxfair; used as the fast checker in tests/ and as the
timed CPU baseline in bench.py.  The descriptor is the oracle's OWN (oracle/ufo.h, mirrored below):
nothing is imported from the product package, so a field-order or constant slip in the product's
binding cannot be common to both sides of a parity test.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libufair_oracle.so")
_LIB = None

# ---- mirror of oracle/ufo.h -------------------------------------------------------------------
E_MEMBER, E_SCENARIO = 0, 1
FEXT_NONE, FEXT_SCENARIO, FEXT_MEMBER = 0, 1, 2
OUT_C, OUT_RF, OUT_T, OUT_ALPHA, OUT_E = 1, 2, 4, 8, 16
_pd, _pi = C.POINTER(C.c_double), C.POINTER(C.c_int32)


def state_rows(n_gas: int) -> int:
    return 5 * n_gas + 3


class UfoDesc(C.Structure):
    """`struct ufo_desc` of oracle/ufo.h (field order and types must match THAT header)."""
    _fields_ = [
        ("struct_size", C.c_uint64),
        ("emissions", C.c_void_p), ("gas_params", C.c_void_p), ("thermal_params", C.c_void_p),
        ("f_ext", C.c_void_p), ("e_scale", C.c_void_p), ("state_in", C.c_void_p), ("scen_idx", C.c_void_p),
        ("out_C", C.c_void_p), ("out_RF", C.c_void_p), ("out_T", C.c_void_p), ("out_alpha", C.c_void_p),
        ("out_E", C.c_void_p), ("state_out", C.c_void_p),
        ("n_member", C.c_int64), ("ld_member", C.c_int64),
        ("n_gas", C.c_int32), ("n_t", C.c_int32), ("n_scen", C.c_int32),
        ("e_mode", C.c_int32), ("fext_mode", C.c_int32), ("alpha_mode", C.c_int32), ("newton_iters", C.c_int32),
        ("t_mode", C.c_int32), ("out_mask", C.c_int32), ("conc_driven", C.c_int32),
        ("dt", C.c_double), ("iirf_h", C.c_double), ("iirf_max", C.c_double),
    ]

    def __init__(self, **kw):
        super().__init__()
        self.struct_size = C.sizeof(UfoDesc)
        for k, v in kw.items():
            setattr(self, k, v)


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("ufair_oracle.c", "ufair_oracle_fast.c", "ufo.h")]
    stale = (not os.path.exists(_SO)) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(_SO) for p in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True,
                       stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _LIB
    if _LIB is None:
        build()
        L = C.CDLL(_SO)
        L.ufo_run_f64.restype = C.c_int
        L.ufo_run_f64.argtypes = [C.POINTER(UfoDesc), C.c_int]
        L.ufo_run_blocked_f64.restype = C.c_int
        L.ufo_run_blocked_f64.argtypes = [C.POINTER(UfoDesc), C.c_int]
        L.ufo_hfc_pulse.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
        L.ufo_g1g0.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_double, C.c_int,
                               C.c_void_p, C.c_void_p]
        L.ufo_kq.argtypes = [C.c_void_p] * 4 + [C.c_double, C.c_int64, C.c_void_p, C.c_void_p]
        L.ufo_stats_f64.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_double, C.c_double,
                                    C.c_int, C.c_void_p, C.c_void_p]
        L.ufo_max_threads.restype = C.c_int
        _LIB = L
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def oxfair(emissions, gas_params, thermal_params, *, dt=1.0, scen_idx=None, e_scale=None, f_ext=None,
           fext_per_member=False, e_scenario=None, state_in=None, alpha_mode=0, newton_iters=0,
           iirf_max=None, iirf_h=100.0, t_mode=0, want_alpha=False, outputs=("C", "RF", "T"),
           n_threads=0, conc_driven=0, blocked=False):
    """C-oracle twin of oracle.ufair_oracle.oxfair (same arguments, float64 only).

    blocked=False: the textbook scalar loop (the parity checker).  blocked=True: the tiled, vectorised
    rendering of the same loop (ufair_oracle_fast.c: the timed CPU baseline; a few ulp from the
    scalar one, emission-driven runs only)."""
    E = _f64(emissions)
    gp = _f64(gas_params)
    tp = _f64(thermal_params)
    G, n_t = E.shape[0], E.shape[1]
    M = gp.shape[2]
    shared = (scen_idx is not None or E.shape[2] != M) if e_scenario is None else bool(e_scenario)
    si = None if scen_idx is None else np.ascontiguousarray(scen_idx, dtype=np.int32)
    es = _f64(e_scale)
    fx = _f64(f_ext)
    if fx is not None and fx.ndim == 1:
        fx = fx[:, None].copy()
    if fx is not None and shared and not fext_per_member and fx.shape[1] == 1 and E.shape[2] > 1:
        fx = np.ascontiguousarray(np.broadcast_to(fx, (n_t, E.shape[2])))   # one shared series for every scenario
    st_in = _f64(state_in)
    want = set(outputs) | ({"alpha"} if want_alpha else set())
    out = {}
    if "C" in want:
        out["C"] = np.empty((G, n_t, M))
    if "RF" in want:
        out["RF"] = np.empty((G, n_t, M))
    if "T" in want:
        out["T"] = np.empty((n_t, M))
    if "alpha" in want:
        out["alpha"] = np.empty((G, n_t, M))
    if conc_driven or "E" in want:
        out["E"] = np.empty((G, n_t, M))
    out["state"] = np.empty((state_rows(G), M))
    d = UfoDesc(
        n_gas=G, n_t=n_t, n_member=M, ld_member=M,
        n_scen=(E.shape[2] if shared else (1 if fx is None or fext_per_member else fx.shape[1])),
        e_mode=E_SCENARIO if shared else E_MEMBER,
        fext_mode=(FEXT_NONE if fx is None else (FEXT_MEMBER if fext_per_member else FEXT_SCENARIO)),
        alpha_mode=alpha_mode, newton_iters=newton_iters, t_mode=t_mode,
        out_mask=(OUT_C * ("C" in want) | OUT_RF * ("RF" in want) | OUT_T * ("T" in want)
                  | OUT_ALPHA * ("alpha" in want) | OUT_E * ("E" in out)),
        conc_driven=int(conc_driven), out_E=_p(out.get("E")),
        dt=dt, iirf_h=iirf_h, iirf_max=(0.0 if iirf_max is None else float(iirf_max)),
        emissions=_p(E), scen_idx=_p(si), e_scale=_p(es), f_ext=_p(fx), gas_params=_p(gp),
        thermal_params=_p(tp), state_in=_p(st_in),
        out_C=_p(out.get("C")), out_RF=_p(out.get("RF")), out_T=_p(out.get("T")),
        out_alpha=_p(out.get("alpha")), state_out=_p(out["state"]))
    run = lib().ufo_run_blocked_f64 if blocked else lib().ufo_run_f64
    rc = run(C.byref(d), int(n_threads))
    if rc < 0:
        raise RuntimeError(f"{'ufo_run_blocked_f64' if blocked else 'ufo_run_f64'} failed: {rc}")
    out["threads"] = rc
    return out


def hfc_pulse(e0, time):
    e0 = _f64(e0)
    time = _f64(time)
    out = np.empty_like(e0)
    lib().ufo_hfc_pulse(_p(e0), _p(time), _p(out), e0.size)
    return out


def g1g0(a, tau, h=100.0, alpha_mode=0):
    a = _f64(a)
    tau = _f64(tau)
    n = a.shape[1]
    g1 = np.empty(n)
    g0 = np.empty(n)
    lib().ufo_g1g0(_p(a), _p(tau), n, n, h, alpha_mode, _p(g1), _p(g0))
    return g1, g0


def kq(tcr, ecs, d1, d2, f2x=3.74):
    tcr, ecs, d1, d2 = (_f64(x) for x in (tcr, ecs, d1, d2))
    q1 = np.empty_like(tcr)
    q2 = np.empty_like(tcr)
    lib().ufo_kq(_p(tcr), _p(ecs), _p(d1), _p(d2), f2x, tcr.size, _p(q1), _p(q2))
    return q1, q2


def temperature_stats(T, lo, hi, bins):
    T = _f64(T)
    n_t, M = T.shape
    hist = np.zeros((n_t, bins), dtype=np.uint64)
    mom = np.zeros((n_t, 4))
    lib().ufo_stats_f64(_p(T), n_t, M, M, lo, hi, bins, _p(hist), _p(mom))
    return hist, mom


def max_threads() -> int:
    return lib().ufo_max_threads()
