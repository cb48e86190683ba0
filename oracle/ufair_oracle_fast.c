/*
 * ufair_oracle_fast.c -- the CPU baseline bench.py times: the oracle's loop (ufair_oracle.c,
 * run_member) rendered the way a CPU implementation that cares about speed would be written.
 *
 * TEST / MEASUREMENT INFRASTRUCTURE -- NOT PART OF THE PRODUCT.  The textbook scalar version in
 * ufair_oracle.c stays the parity checker; this file is checked against it (tests/test_oracle.py)
 * and is only ever used as the timed CPU arm (bench.py cpu_baseline / --impl reference).
 *
 * Same equations, same order of the five steps (alpha_val -> step_conc -> step_forc -> step_temp,
 * the names the reference reserves in .coveragerc:12-19), but:
 *   - members are processed in tiles of UFO_TILE; inside a tile time is the OUTER loop and the
 *     member lanes the inner one (unit stride in every [..][member] array: one cache line serves
 *     eight members, where the scalar version touches a new line per element);
 *   - per-member parameters and everything that does not change in time (c a_i tau_i, dt/tau_i,
 *     exp(-dt/d_j), 1/g1, ln g0 ...) are hoisted into tile-local arrays once;
 *   - the lane loops are `omp simd` and this file is built -O3 -ffast-math, so exp / expm1 / log
 *     vectorise through glibc's libmvec; target_clones picks AVX-512 / AVX2 / baseline at load time
 *     (the library is built on one machine and timed on another, so no -march=native);
 *   - OpenMP threads take tiles.
 * Vector libm and re-association move results by a few ulp, so this version is NOT bit-identical to
 * the scalar one (tests hold it to 1e-9 relative); it is the speed baseline, not the checker.
 * Emission-driven runs with C0 > 0 only; anything else returns -2 and the caller uses ufo_run_f64.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "ufo.h"

#define UFO_TILE 64

#if defined(__x86_64__) && defined(__GNUC__) && !defined(UFO_NO_CLONES)
#define UFO_CLONES __attribute__((target_clones("arch=x86-64-v4", "arch=x86-64-v3", "default")))
#else
#define UFO_CLONES
#endif

typedef struct {
  /* per gas, per lane */
  double ka[UFO_MAX_GAS][4][UFO_TILE];  /* c a_i tau_i */
  double k0[UFO_MAX_GAS][4][UFO_TILE];  /* dt / tau_i */
  double at[UFO_MAX_GAS][4][UFO_TILE];  /* a_i tau_i (Newton) */
  double hk[UFO_MAX_GAS][4][UFO_TILE];  /* h / tau_i (Newton) */
  double rho0[UFO_MAX_GAS][UFO_TILE], rhoU[UFO_MAX_GAS][UFO_TILE], rhoT[UFO_MAX_GAS][UFO_TILE],
      rhoA[UFO_MAX_GAS][UFO_TILE];      /* r / g1 */
  double lng0[UFO_MAX_GAS][UFO_TILE], g0s[UFO_MAX_GAS][UFO_TILE], umax[UFO_MAX_GAS][UFO_TILE],
      g1[UFO_MAX_GAS][UFO_TILE];
  double invc[UFO_MAX_GAS][UFO_TILE], C0[UFO_MAX_GAS][UFO_TILE], invC0[UFO_MAX_GAS][UFO_TILE],
      sqrtC0[UFO_MAX_GAS][UFO_TILE];
  double f1[UFO_MAX_GAS][UFO_TILE], f2[UFO_MAX_GAS][UFO_TILE], f3[UFO_MAX_GAS][UFO_TILE], esc[UFO_MAX_GAS][UFO_TILE];
  double R[UFO_MAX_GAS][4][UFO_TILE], Gc[UFO_MAX_GAS][UFO_TILE];
  /* thermal */
  double qm[2][UFO_TILE], dec[2][UFO_TILE], S[2][UFO_TILE], Tprev[UFO_TILE];
  /* scratch rows */
  double e[UFO_TILE], alpha[UFO_TILE], Ftot[UFO_TILE], Cr[UFO_TILE], Fr[UFO_TILE], Tr[UFO_TILE];
  int scen[UFO_TILE];
} tile_t;

/* one tile of w <= UFO_TILE members starting at m0; lanes >= w shadow the last member */
UFO_CLONES
static void run_tile(const ufo_desc* d, int64_t m0, int w, tile_t* T) {
  const int G = d->n_gas, n_t = d->n_t, W = UFO_TILE;
  const int64_t ld = d->ld_member;
  const double dt = d->dt, h = d->iirf_h;
  const int clamp = (d->iirf_max > 0.0 && d->iirf_max < 1e300);
  const int amode = d->alpha_mode, tmid = d->t_mode == UFO_T_MID;
  const double* gp = d->gas_params;
  const double* tp = d->thermal_params;
  const double* sin_ = d->state_in;

  /* ---- hoist: raw parameters -> everything constant in time (g_1, g_0 of .coveragerc:15-16) */
  for (int j = 0; j < W; ++j) {
    const int64_t m = m0 + (j < w ? j : w - 1);
    T->scen[j] = d->scen_idx ? d->scen_idx[m] : 0;
    for (int g = 0; g < G; ++g) {
      const double* p = gp + (int64_t)g * UFO_GP_COUNT * ld + m;
      double g1 = 0.0, s = 0.0;
      for (int i = 0; i < 4; ++i) {
        const double a = p[(UFO_GP_A0 + i) * ld], tau = p[(UFO_GP_TAU0 + i) * ld];
        const double z = h / tau, ez = exp(-z);
        g1 += a * tau * (1.0 - (1.0 + z) * ez);
        s += a * tau * (1.0 - ez);
      }
      s /= g1;
      const double c = p[UFO_GP_EMIS2CONC * ld], C0 = p[UFO_GP_C0 * ld], ig1 = 1.0 / g1;
      for (int i = 0; i < 4; ++i) {
        const double a = p[(UFO_GP_A0 + i) * ld], tau = p[(UFO_GP_TAU0 + i) * ld];
        T->ka[g][i][j] = c * a * tau;
        T->k0[g][i][j] = dt / tau;
        T->at[g][i][j] = a * tau;
        T->hk[g][i][j] = h / tau;
        T->R[g][i][j] = sin_ ? sin_[(5 * g + i) * ld + m] : 0.0;
      }
      T->Gc[g][j] = sin_ ? sin_[(5 * g + 4) * ld + m] : 0.0;
      T->rho0[g][j] = p[UFO_GP_R0 * ld] * ig1;
      T->rhoU[g][j] = p[UFO_GP_RU * ld] * ig1;
      T->rhoT[g][j] = p[UFO_GP_RT * ld] * ig1;
      T->rhoA[g][j] = p[UFO_GP_RA * ld] * ig1;
      T->g1[g][j] = g1;
      T->lng0[g][j] = -s;
      T->g0s[g][j] = 1.0 / sinh(s);
      T->umax[g][j] = clamp ? d->iirf_max * ig1 : 1e300;
      T->invc[g][j] = 1.0 / c;
      T->C0[g][j] = C0;
      T->invC0[g][j] = 1.0 / C0;
      T->sqrtC0[g][j] = sqrt(C0);
      T->f1[g][j] = p[UFO_GP_F1 * ld];
      T->f2[g][j] = p[UFO_GP_F2 * ld];
      T->f3[g][j] = p[UFO_GP_F3 * ld];
      T->esc[g][j] = d->e_scale ? d->e_scale[(int64_t)g * ld + m] : 1.0;
    }
    for (int k = 0; k < 2; ++k) {
      const double q = tp[(UFO_TP_Q1 + k) * ld + m], dd = tp[(UFO_TP_D1 + k) * ld + m];
      const double dec = exp(-dt / dd);
      T->dec[k][j] = dec;
      T->qm[k][j] = q * (1.0 - dec);
      T->S[k][j] = sin_ ? sin_[(5 * G + k) * ld + m] : 0.0;
    }
    T->Tprev[j] = sin_ ? sin_[(5 * G + 2) * ld + m] : 0.0;
  }

  double* oC = (d->out_mask & UFO_OUT_C) ? d->out_C : NULL;
  double* oF = (d->out_mask & UFO_OUT_RF) ? d->out_RF : NULL;
  double* oT = (d->out_mask & UFO_OUT_T) ? d->out_T : NULL;
  double* oA = (d->out_mask & UFO_OUT_ALPHA) ? d->out_alpha : NULL;
  double* oE = (d->out_mask & UFO_OUT_E) ? d->out_E : NULL;
  const size_t wb = (size_t)w * sizeof(double);

  /* ---- oxfair (.coveragerc:19): time outermost, member lanes innermost */
  for (int t = 0; t < n_t; ++t) {
    for (int j = 0; j < W; ++j) T->Ftot[j] = 0.0;
    for (int g = 0; g < G; ++g) {
      /* this step's emissions of the tile's lanes */
      if (d->e_mode == UFO_E_SCENARIO) {
        const double* row = d->emissions + ((int64_t)g * n_t + t) * d->n_scen;
        for (int j = 0; j < W; ++j) T->e[j] = row[T->scen[j]] * T->esc[g][j];
      } else {
        const double* row = d->emissions + ((int64_t)g * n_t + t) * ld + m0;
        memcpy(T->e, row, wb);
        for (int j = w; j < W; ++j) T->e[j] = T->e[w - 1];
      }
      double* restrict R0 = T->R[g][0];
      double* restrict R1 = T->R[g][1];
      double* restrict R2 = T->R[g][2];
      double* restrict R3 = T->R[g][3];
      /* alpha_val (.coveragerc:17): state at t-1 -> alpha */
      if (amode == UFO_ALPHA_ONE) {
        for (int j = 0; j < W; ++j) T->alpha[j] = 1.0;
      } else {
#pragma omp simd
        for (int j = 0; j < W; ++j) {
          const double Ga = ((R0[j] + R1[j]) + (R2[j] + R3[j])) * T->invc[g][j];
          double u = T->rho0[g][j] + T->rhoU[g][j] * (T->Gc[g][j] - Ga) + T->rhoT[g][j] * T->Tprev[j] + T->rhoA[g][j] * Ga;
          u = u > T->umax[g][j] ? T->umax[g][j] : u;
          T->Cr[j] = u; /* iIRF / g1 */
        }
        if (amode == UFO_ALPHA_SINH) {
#pragma omp simd
          for (int j = 0; j < W; ++j) {
            const double ep = exp(T->Cr[j]);
            T->alpha[j] = T->g0s[g][j] * 0.5 * (ep - 1.0 / ep);
          }
        } else {
#pragma omp simd
          for (int j = 0; j < W; ++j) T->alpha[j] = exp(T->Cr[j] + T->lng0[g][j]);
        }
        if (amode == UFO_ALPHA_NEWTON) {
          for (int k = 0; k < d->newton_iters; ++k) {
#pragma omp simd
            for (int j = 0; j < W; ++j) {
              const double al = T->alpha[j], ia = 1.0 / al, iirf = T->Cr[j] * T->g1[g][j];
              double f = -iirf, fp = 0.0;
              for (int i = 0; i < 4; ++i) {
                const double z = T->hk[g][i][j] * ia;
                const double mz = -expm1(-z);
                f += T->at[g][i][j] * al * mz;
                fp += T->at[g][i][j] * (mz - z * (1.0 - mz));
              }
              const double an = al - f / fp;
              T->alpha[j] = an > 0.5 * al ? an : 0.5 * al;
            }
          }
        }
      }
      /* step_conc (.coveragerc:12) + step_forc (.coveragerc:13) */
#pragma omp simd
      for (int j = 0; j < W; ++j) {
        const double al = T->alpha[j], ia = 1.0 / al, ea = T->e[j] * al;
        const double m0_ = -expm1(-T->k0[g][0][j] * ia), m1_ = -expm1(-T->k0[g][1][j] * ia);
        const double m2_ = -expm1(-T->k0[g][2][j] * ia), m3_ = -expm1(-T->k0[g][3][j] * ia);
        R0[j] += m0_ * (ea * T->ka[g][0][j] - R0[j]);
        R1[j] += m1_ * (ea * T->ka[g][1][j] - R1[j]);
        R2[j] += m2_ * (ea * T->ka[g][2][j] - R2[j]);
        R3[j] += m3_ * (ea * T->ka[g][3][j] - R3[j]);
        T->Gc[g][j] += T->e[j] * dt;
        const double sumR = (R0[j] + R1[j]) + (R2[j] + R3[j]);
        const double C = T->C0[g][j] + sumR;
        const double lg = T->f1[g][j] != 0.0 ? log(C * T->invC0[g][j]) : 0.0;
        const double sq = T->f3[g][j] != 0.0 ? sqrt(C) - T->sqrtC0[g][j] : 0.0;
        const double F = T->f1[g][j] * lg + T->f2[g][j] * sumR + T->f3[g][j] * sq;
        T->Cr[j] = C;
        T->Fr[j] = F;
        T->Ftot[j] += F;
      }
      const int64_t o = ((int64_t)g * n_t + t) * ld + m0;
      if (oC) memcpy(oC + o, T->Cr, wb);
      if (oF) memcpy(oF + o, T->Fr, wb);
      if (oA) memcpy(oA + o, T->alpha, wb);
      if (oE) memcpy(oE + o, T->e, wb);
    }
    if (d->fext_mode == UFO_FEXT_SCENARIO) {
      const double* row = d->f_ext + (int64_t)t * d->n_scen;
      for (int j = 0; j < W; ++j) T->Ftot[j] += row[T->scen[j]];
    } else if (d->fext_mode == UFO_FEXT_MEMBER) {
      const double* row = d->f_ext + (int64_t)t * ld + m0;
      for (int j = 0; j < w; ++j) T->Ftot[j] += row[j];
    }
    /* step_temp (.coveragerc:14) */
#pragma omp simd
    for (int j = 0; j < W; ++j) {
      const double s0 = T->qm[0][j] * T->Ftot[j] + T->S[0][j] * T->dec[0][j];
      const double s1 = T->qm[1][j] * T->Ftot[j] + T->S[1][j] * T->dec[1][j];
      const double Tn = tmid ? 0.5 * ((T->S[0][j] + T->S[1][j]) + (s0 + s1)) : s0 + s1;
      T->S[0][j] = s0;
      T->S[1][j] = s1;
      T->Tprev[j] = Tn;
      T->Tr[j] = Tn;
    }
    if (oT) memcpy(oT + (int64_t)t * ld + m0, T->Tr, wb);
  }
  double* so = d->state_out;
  if (so) {
    for (int g = 0; g < G; ++g) {
      for (int i = 0; i < 4; ++i) memcpy(so + (5 * g + i) * ld + m0, T->R[g][i], wb);
      memcpy(so + (5 * g + 4) * ld + m0, T->Gc[g], wb);
    }
    memcpy(so + (5 * G + 0) * ld + m0, T->S[0], wb);
    memcpy(so + (5 * G + 1) * ld + m0, T->S[1], wb);
    memcpy(so + (5 * G + 2) * ld + m0, T->Tprev, wb);
  }
}

int ufo_run_blocked_f64(const ufo_desc* d, int n_threads) {
  if (!d || d->struct_size != sizeof(ufo_desc)) return -1;
  if (d->n_gas < 1 || d->n_gas > UFO_MAX_GAS || d->n_t < 0 || d->n_member < 0) return -1;
  if (d->conc_driven) return -2;
  for (int g = 0; g < d->n_gas; ++g) /* the fast path multiplies by 1/C0 */
    for (int64_t m = 0; m < d->n_member; ++m)
      if (!(d->gas_params[((int64_t)g * UFO_GP_COUNT + UFO_GP_C0) * d->ld_member + m] > 0.0)) return -2;
  const int64_t n_tile = (d->n_member + UFO_TILE - 1) / UFO_TILE;
  int used = 1;
#ifdef _OPENMP
  if (n_threads <= 0) n_threads = omp_get_max_threads();
  used = n_threads;
#pragma omp parallel num_threads(n_threads)
#endif
  {
    tile_t* T = (tile_t*)aligned_alloc(64, (sizeof(tile_t) + 63) / 64 * 64);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 4)
#endif
    for (int64_t k = 0; k < n_tile; ++k) {
      const int64_t m0 = k * UFO_TILE;
      const int w = (int)(d->n_member - m0 < UFO_TILE ? d->n_member - m0 : UFO_TILE);
      if (T) run_tile(d, m0, w, T);
    }
    free(T);
  }
  return used;
}
