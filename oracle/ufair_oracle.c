/*
 * ufair_oracle.c -- CPU oracle (plain C, float64) for the Universal-FaIR ensemble path.
 *
 * TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load the library built from this file.
 *
 * It takes its OWN descriptor (oracle/ufo.h: host pointers, the oracle's own field order and
 * constants -- deliberately not the product's include/ufair.h, so that a layout mistake on either
 * side shows up as a parity failure instead of cancelling out) and restates the algorithm in
 * textbook form, one scalar member at a time, OpenMP over members.  ufair_oracle_fast.c holds the
 * blocked / vectorised rendering of the same loop that bench.py times as the CPU baseline.
 *
 * Parity status (same as oracle/ufair_oracle.py, which this file mirrors function by function):
 *   - ufo_hfc_pulse: PINNED to reference U_FaIR/concentrations.py:4-5 and the golden vector in
 *     reference tests/unit/test_hfcs.py:6-13.
 *   - everything else: PARITY UNPINNED -- the reference only names step_conc, step_forc,
 *     step_temp, g_1, g_0, alpha_val, k_q, oxfair (.coveragerc:12-19; prose README.md:6-10) and
 *     ships no code or vectors for them.  Restated from the published equations.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "ufo.h"

/* reference U_FaIR/concentrations.py:5 : emissions[0] * np.exp(-time), already broadcast */
int ufo_hfc_pulse(const double* e0, const double* time, double* out, int64_t n) {
  for (int64_t i = 0; i < n; ++i) out[i] = e0[i] * exp(-time[i]);
  return 0;
}

/* g_1 (.coveragerc:15) */
static double g_1(const double a[4], const double tau[4], double h) {
  double s = 0.0;
  for (int i = 0; i < 4; ++i) s += a[i] * tau[i] * (1.0 - (1.0 + h / tau[i]) * exp(-h / tau[i]));
  return s;
}

/* g_0 (.coveragerc:16) */
static double g_0(const double a[4], const double tau[4], double h, double g1, int alpha_mode) {
  double s = 0.0;
  for (int i = 0; i < 4; ++i) s += a[i] * tau[i] * (1.0 - exp(-h / tau[i]));
  s /= g1;
  return alpha_mode == UFO_ALPHA_SINH ? 1.0 / sinh(s) : exp(-s);
}

int ufo_g1g0(const double* a, const double* tau, int64_t n, int64_t ld, double h, int alpha_mode,
             double* g1, double* g0) {
  for (int64_t m = 0; m < n; ++m) {
    double aa[4], tt[4];
    for (int i = 0; i < 4; ++i) { aa[i] = a[i * ld + m]; tt[i] = tau[i * ld + m]; }
    g1[m] = g_1(aa, tt, h);
    g0[m] = g_0(aa, tt, h, g1[m], alpha_mode);
  }
  return 0;
}

/* k_q (.coveragerc:18) */
int ufo_kq(const double* tcr, const double* ecs, const double* d1, const double* d2, double f2x,
           int64_t n, double* q1, double* q2) {
  for (int64_t m = 0; m < n; ++m) {
    double k1 = 1.0 - (d1[m] / 70.0) * (1.0 - exp(-70.0 / d1[m]));
    double k2 = 1.0 - (d2[m] / 70.0) * (1.0 - exp(-70.0 / d2[m]));
    double den = f2x * (k1 - k2);
    q1[m] = (tcr[m] - ecs[m] * k2) / den;
    q2[m] = (ecs[m] * k1 - tcr[m]) / den;
  }
  return 0;
}

/* alpha_val (.coveragerc:17) */
static double alpha_val(double iirf, double g0, double g1, const double a[4], const double tau[4],
                        int alpha_mode, int newton_iters, double h) {
  if (alpha_mode == UFO_ALPHA_ONE) return 1.0;
  if (alpha_mode == UFO_ALPHA_SINH) return g0 * sinh(iirf / g1);
  double alpha = g0 * exp(iirf / g1);
  if (alpha_mode == UFO_ALPHA_NEWTON) {
    for (int k = 0; k < newton_iters; ++k) {
      double f = -iirf, fp = 0.0;
      for (int i = 0; i < 4; ++i) {
        double z = h / (alpha * tau[i]);
        f += a[i] * alpha * tau[i] * (-expm1(-z));
        fp += a[i] * tau[i] * (1.0 - (1.0 + z) * exp(-z));
      }
      double an = alpha - f / fp;
      alpha = an > 0.5 * alpha ? an : 0.5 * alpha;
    }
  }
  return alpha;
}

/* oxfair (.coveragerc:19) for one member; step_conc / step_forc / step_temp inline and labelled */
static void run_member(const ufo_desc* d, int64_t m) {
  const int G = d->n_gas, n_t = d->n_t;
  const int64_t ld = d->ld_member;
  const double* E = (const double*)d->emissions;
  const double* gp = (const double*)d->gas_params;
  const double* tp = (const double*)d->thermal_params;
  const double* sin_ = (const double*)d->state_in;
  const double* esc = (const double*)d->e_scale;
  const double* fx = (const double*)d->f_ext;
  const double dt = d->dt, h = d->iirf_h;
  const int clamp = (d->iirf_max > 0.0 && isfinite(d->iirf_max));
  const int s = d->scen_idx ? d->scen_idx[m] : 0;

  double a[UFO_MAX_GAS][4], tau[UFO_MAX_GAS][4], g1[UFO_MAX_GAS], g0[UFO_MAX_GAS];
  double R[UFO_MAX_GAS][4], Gc[UFO_MAX_GAS], S[2], Tprev;
#define GP(g, r) gp[((int64_t)(g) * UFO_GP_COUNT + (r)) * ld + m]
  for (int g = 0; g < G; ++g) {
    for (int i = 0; i < 4; ++i) { a[g][i] = GP(g, UFO_GP_A0 + i); tau[g][i] = GP(g, UFO_GP_TAU0 + i); }
    g1[g] = g_1(a[g], tau[g], h);
    g0[g] = g_0(a[g], tau[g], h, g1[g], d->alpha_mode);
    for (int i = 0; i < 4; ++i) R[g][i] = sin_ ? sin_[(5 * g + i) * ld + m] : 0.0;
    Gc[g] = sin_ ? sin_[(5 * g + 4) * ld + m] : 0.0;
  }
  S[0] = sin_ ? sin_[(5 * G + 0) * ld + m] : 0.0;
  S[1] = sin_ ? sin_[(5 * G + 1) * ld + m] : 0.0;
  Tprev = sin_ ? sin_[(5 * G + 2) * ld + m] : 0.0;
  const double q[2] = {tp[UFO_TP_Q1 * ld + m], tp[UFO_TP_Q2 * ld + m]};
  const double dd[2] = {tp[UFO_TP_D1 * ld + m], tp[UFO_TP_D2 * ld + m]};

  double* oC = (d->out_mask & UFO_OUT_C) ? (double*)d->out_C : NULL;
  double* oF = (d->out_mask & UFO_OUT_RF) ? (double*)d->out_RF : NULL;
  double* oT = (d->out_mask & UFO_OUT_T) ? (double*)d->out_T : NULL;
  double* oA = (d->out_mask & UFO_OUT_ALPHA) ? (double*)d->out_alpha : NULL;
  double* oE = (d->out_mask & UFO_OUT_E) ? (double*)d->out_E : NULL;

  for (int t = 0; t < n_t; ++t) {
    double Ftot = 0.0;
    for (int g = 0; g < G; ++g) {
      double e;
      if (d->e_mode == UFO_E_SCENARIO) {
        e = E[((int64_t)g * n_t + t) * d->n_scen + s];
        if (esc) e = e * esc[(int64_t)g * ld + m];
      } else {
        e = E[((int64_t)g * n_t + t) * ld + m];
      }
      const double c = GP(g, UFO_GP_EMIS2CONC), C0 = GP(g, UFO_GP_C0);
      /* alpha from the state at t-1 */
      double Ga = (R[g][0] + R[g][1] + R[g][2] + R[g][3]) / c;
      double iirf = GP(g, UFO_GP_R0) + GP(g, UFO_GP_RU) * (Gc[g] - Ga) + GP(g, UFO_GP_RT) * Tprev +
                    GP(g, UFO_GP_RA) * Ga;
      if (clamp && iirf > d->iirf_max) iirf = d->iirf_max;
      double alpha = alpha_val(iirf, g0[g], g1[g], a[g], tau[g], d->alpha_mode, d->newton_iters, h);
      /* concentration-driven gas: the input row is the target C; step_conc is linear in E, so
       * E = (C_target - C0 - sum R_i (1 - m_i)) / (c sum a_i alpha tau_i m_i) */
      if ((d->conc_driven >> g) & 1) {
        double keep = 0.0, gain = 0.0;
        for (int i = 0; i < 4; ++i) {
          double at = alpha * tau[g][i];
          double mi = -expm1(-dt / at);
          keep = keep + R[g][i] * (1.0 - mi);
          gain = gain + c * a[g][i] * at * mi;
        }
        e = (e - C0 - keep) / gain;
      }
      /* step_conc (.coveragerc:12) */
      double sumR = 0.0;
      for (int i = 0; i < 4; ++i) {
        double at = alpha * tau[g][i];
        double mi = -expm1(-dt / at);
        R[g][i] = e * c * a[g][i] * at * mi + R[g][i] * (1.0 - mi);
      }
      sumR = R[g][0] + R[g][1] + R[g][2] + R[g][3];
      Gc[g] = Gc[g] + e * dt;
      double C = C0 + sumR;
      /* step_forc (.coveragerc:13) */
      /* a term whose coefficient is exactly zero contributes exactly zero (C0 = 0 gases) */
      const double f1 = GP(g, UFO_GP_F1), f2 = GP(g, UFO_GP_F2), f3 = GP(g, UFO_GP_F3);
      double logt = (f1 != 0.0) ? f1 * log(C / C0) : 0.0;
      double sqrtt = (f3 != 0.0) ? f3 * (sqrt(C) - sqrt(C0)) : 0.0;
      double F = logt + f2 * (C - C0) + sqrtt;
      const int64_t o = ((int64_t)g * n_t + t) * ld + m;
      if (oC) oC[o] = C;
      if (oF) oF[o] = F;
      if (oA) oA[o] = alpha;
      if (oE) oE[o] = e;
      Ftot = Ftot + F;
    }
    if (d->fext_mode == UFO_FEXT_SCENARIO) Ftot = Ftot + fx[(int64_t)t * d->n_scen + s];
    else if (d->fext_mode == UFO_FEXT_MEMBER) Ftot = Ftot + fx[(int64_t)t * ld + m];
    /* step_temp (.coveragerc:14) */
    double T = 0.0;
    for (int j = 0; j < 2; ++j) {
      double dec = exp(-dt / dd[j]);
      double Sn = q[j] * Ftot * (1.0 - dec) + S[j] * dec;
      T += (d->t_mode == UFO_T_MID) ? (S[j] + Sn) / 2.0 : Sn;
      S[j] = Sn;
    }
    if (oT) oT[(int64_t)t * ld + m] = T;
    Tprev = T;
  }
  double* so = (double*)d->state_out;
  if (so) {
    for (int g = 0; g < G; ++g) {
      for (int i = 0; i < 4; ++i) so[(5 * g + i) * ld + m] = R[g][i];
      so[(5 * g + 4) * ld + m] = Gc[g];
    }
    so[(5 * G + 0) * ld + m] = S[0];
    so[(5 * G + 1) * ld + m] = S[1];
    so[(5 * G + 2) * ld + m] = Tprev;
  }
#undef GP
}

/* n_threads <= 0: all the OpenMP runtime offers. Returns the thread count used. */
int ufo_run_f64(const ufo_desc* d, int n_threads) {
  if (!d || d->struct_size != sizeof(ufo_desc)) return -1;
  if (d->n_gas < 1 || d->n_gas > UFO_MAX_GAS || d->n_t < 0 || d->n_member < 0) return -1;
  int used = 1;
#ifdef _OPENMP
  if (n_threads <= 0) n_threads = omp_get_max_threads();
  used = n_threads;
#pragma omp parallel for schedule(static) num_threads(n_threads)
#endif
  for (int64_t m = 0; m < d->n_member; ++m) run_member(d, m);
  return used;
}

/* a8: per-step histogram + moments of T[n_t][ld] with the kernel's binning rule
 * (sub, then mul, each rounded; NaN not counted). hist: [n_t][bins] u64; mom: [n_t][4]. */
int ufo_stats_f64(const double* T, int n_t, int64_t n_member, int64_t ld, double lo, double hi, int bins,
                  uint64_t* hist, double* mom) {
  const double inv_w = (double)bins / (hi - lo);
  for (int t = 0; t < n_t; ++t) {
    uint64_t* hrow = hist + (int64_t)t * bins;
    memset(hrow, 0, sizeof(uint64_t) * (size_t)bins);
    double s = 0.0, ss = 0.0, mn = INFINITY, mx = -INFINITY;
    for (int64_t m = 0; m < n_member; ++m) {
      volatile double diff = T[(int64_t)t * ld + m] - lo; /* volatile: forbid FMA contraction */
      double x = diff * inv_w;
      if (x == x) {
        double f = floor(x);
        int b = f < 0.0 ? 0 : (f > (double)(bins - 1) ? bins - 1 : (int)f);
        hrow[b] += 1;
      }
      double v = T[(int64_t)t * ld + m];
      s += v; ss += v * v;
      if (v < mn) mn = v;
      if (v > mx) mx = v;
    }
    mom[t * 4 + 0] = s; mom[t * 4 + 1] = ss; mom[t * 4 + 2] = mn; mom[t * 4 + 3] = mx;
  }
  return 0;
}

int ufo_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
