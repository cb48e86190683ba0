/*
 * ufo.h -- descriptor and constants of the CPU oracle (oracle/ufair_oracle.c, ufair_oracle_fast.c).
 *
 * TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT.
 *
 * This is the oracle's OWN statement of the run description: host pointers, its own field order,
 * its own names.  It deliberately does not include the product's include/ufair.h, and
 * oracle/c_oracle.py mirrors THIS struct, not fiveeqscm_b200/_abi.py -- the two sides share only
 * the documented array layouts ([gas][t][member], member axis fastest; the row order of the
 * parameter and state arrays below), so a field-order or enum slip on either side is caught by
 * the parity tests instead of being common to both.
 *
 * Row order of the arrays (the data format; reference: time is axis 0 of `emissions`,
 * U_FaIR/concentrations.py:5, members are the axis this path adds below it):
 *   gas_params[n_gas][17][ld]: a_1..a_4, tau_1..tau_4, r0, rU, rT, rA, C0, c (emission ->
 *     concentration), f1 (log term), f2 (linear term), f3 (sqrt term)
 *   thermal_params[4][ld]:     q1, q2, d1, d2
 *   state[5 n_gas + 3][ld]:    per gas R_1..R_4, G_cum; then S_1, S_2, T_prev
 */
#ifndef UFO_H_
#define UFO_H_

#include <stdint.h>

#define UFO_MAX_GAS 4

enum { UFO_GP_A0 = 0, UFO_GP_TAU0 = 4, UFO_GP_R0 = 8, UFO_GP_RU = 9, UFO_GP_RT = 10, UFO_GP_RA = 11,
       UFO_GP_C0 = 12, UFO_GP_EMIS2CONC = 13, UFO_GP_F1 = 14, UFO_GP_F2 = 15, UFO_GP_F3 = 16, UFO_GP_COUNT = 17 };
enum { UFO_TP_Q1 = 0, UFO_TP_Q2 = 1, UFO_TP_D1 = 2, UFO_TP_D2 = 3, UFO_TP_COUNT = 4 };

enum { UFO_E_MEMBER = 0, UFO_E_SCENARIO = 1 };                         /* emissions per member | [..][n_scen] + scen_idx */
enum { UFO_FEXT_NONE = 0, UFO_FEXT_SCENARIO = 1, UFO_FEXT_MEMBER = 2 }; /* external forcing: none | [n_t][n_scen] | [n_t][ld] */
enum { UFO_ALPHA_EXP = 0, UFO_ALPHA_SINH = 1, UFO_ALPHA_NEWTON = 2, UFO_ALPHA_ONE = 3 };
enum { UFO_T_MID = 0, UFO_T_END = 1 };
enum { UFO_OUT_C = 1, UFO_OUT_RF = 2, UFO_OUT_T = 4, UFO_OUT_ALPHA = 8, UFO_OUT_E = 16 };

typedef struct ufo_desc {
  uint64_t struct_size; /* = sizeof(ufo_desc) */
  /* ---- inputs (host) ---- */
  const double* emissions;
  const double* gas_params;
  const double* thermal_params;
  const double* f_ext;
  const double* e_scale;
  const double* state_in;
  const int32_t* scen_idx;
  /* ---- outputs (host; NULL = not wanted, also governed by out_mask) ---- */
  double* out_C;
  double* out_RF;
  double* out_T;
  double* out_alpha;
  double* out_E;
  double* state_out;
  /* ---- sizes ---- */
  int64_t n_member;
  int64_t ld_member;
  int32_t n_gas;
  int32_t n_t;
  int32_t n_scen;
  /* ---- modes ---- */
  int32_t e_mode;
  int32_t fext_mode;
  int32_t alpha_mode;
  int32_t newton_iters;
  int32_t t_mode;
  int32_t out_mask;
  int32_t conc_driven; /* bit g: gas g's input rows are target concentrations */
  double dt;
  double iirf_h;
  double iirf_max; /* <= 0 or inf: no ceiling */
} ufo_desc;

/* textbook scalar loop (the checker); returns the thread count used, < 0 on a bad descriptor */
int ufo_run_f64(const ufo_desc* d, int n_threads);
/* the same arithmetic blocked over 64-member tiles and vectorised (the timed CPU baseline;
 * emission-driven runs only); returns the thread count used, < 0 on a bad / unsupported descriptor */
int ufo_run_blocked_f64(const ufo_desc* d, int n_threads);

#endif /* UFO_H_ */
