"""Near-minimax polynomial coefficients for the device math in csrc/ufair_math.cuh.

Chebyshev-node interpolation in 60-digit arithmetic (within a small factor of true minimax),
coefficients rounded to the target precision, error re-measured with the ROUNDED coefficients.
Run:  python tools/gen_poly.py
"""
import mpmath as mp
import numpy as np

mp.mp.dps = 60


def cheb_fit(f, a, b, deg):
    n = deg + 1
    xs = [mp.mpf(a + b) / 2 + mp.mpf(b - a) / 2 * mp.cos(mp.pi * (2 * k + 1) / (2 * n)) for k in range(n)]
    A = mp.matrix(n, n)
    y = mp.matrix(n, 1)
    for i, x in enumerate(xs):
        for j in range(n):
            A[i, j] = x ** j
        y[i] = f(x)
    c = mp.lu_solve(A, y)
    return [c[j] for j in range(n)]


def horner(c, x):
    r = mp.mpf(0)
    for cj in reversed(c):
        r = r * x + cj
    return r


def max_err(f_exact, f_approx, a, b, n=4001, rel=True):
    worst = mp.mpf(0)
    for k in range(n):
        x = mp.mpf(a) + (mp.mpf(b) - a) * k / (n - 1)
        if x == 0:
            continue
        e = abs(f_approx(x) - f_exact(x))
        if rel:
            e = e / abs(f_exact(x))
        worst = max(worst, e)
    return worst


def rnd(c, single):
    if single:
        return [mp.mpf(float(np.float32(float(x)))) for x in c]
    return [mp.mpf(float(x)) for x in c]


def report(name, c, single):
    fmt = (lambda v: f"{float(np.float32(float(v))):.9e}f") if single else (lambda v: float(v).hex())
    print(f"  // {name}")
    print("  " + ", ".join(fmt(v) for v in c))


def expm1_q(single, degs):
    # expm1(r) = r + r^2 Q(r),  |r| <= ln2/2
    L = mp.log(2) / 2
    Q = lambda r: (mp.expm1(r) - r) / (r * r) if abs(r) > mp.mpf(10) ** -20 else mp.mpf(1) / 2 + r / 6
    for deg in degs:
        c = rnd(cheb_fit(Q, -L, L, deg), single)
        ap = lambda r: r + r * r * horner(c, r)
        e = max_err(mp.expm1, ap, -L, L)
        print(f"expm1 Q deg {deg} ({'f32' if single else 'f64'}): max rel err {mp.nstr(e, 3)}  ({mp.nstr(e / mp.mpf(2) ** (-24 if single else -53), 3)} ulp-ish)")
        report(f"expm1 Q deg {deg}", c, single)


def log_poly(single, degs):
    # log(m) = 2 s + s^3 * L(s^2), s = (m-1)/(m+1), m in [sqrt(1/2), sqrt(2)] -> |s| <= 0.1716
    smax = (mp.sqrt(2) - 1) / (mp.sqrt(2) + 1)
    Lf = lambda w: (mp.atanh(mp.sqrt(w)) * 2 - 2 * mp.sqrt(w)) / (mp.sqrt(w) ** 3) if w != 0 else mp.mpf(2) / 3
    for deg in degs:
        c = rnd(cheb_fit(Lf, mp.mpf(10) ** -30, smax ** 2, deg), single)
        ap = lambda s: 2 * s + s ** 3 * horner(c, s * s)
        ex = lambda s: 2 * mp.atanh(s)
        e = max_err(ex, ap, -smax, smax)
        print(f"log L deg {deg} ({'f32' if single else 'f64'}): max rel err {mp.nstr(e, 3)}")
        report(f"log L deg {deg}", c, single)


if __name__ == "__main__":
    expm1_q(False, [8, 9, 10])
    log_poly(False, [6])
    expm1_q(True, [3, 4, 5])
    log_poly(True, [2])
