// ufair_kernel.cuh -- the fused Universal-FaIR time-stepping kernel (sm_100a).
//
// Work decomposition (v5): one LANE = one (member, gas) pair; one WARP = MW consecutive members x
// all NGAS gases (f64: MW = 32, 16, 10, 8 for 1..4 gases; lanes beyond NGAS*MW idle).  The NGAS
// lanes of a member exchange their radiative forcings with warp shuffles, so a warp never waits
// for another warp: there is NO block-level synchronisation anywhere in the kernel.  Each warp
// runs its own TMA pipeline (its own shared-memory ring and mbarriers).
//   Every lane keeps its gas's four pools, cumulative emissions and (redundantly, bit-identically
//   in the NGAS lanes of a member) the two thermal boxes in REGISTERS across the serial time loop;
//   its 23 derived per-member constants sit in shared memory ([param][lane], conflict-free
//   LDS.64 addressed as base + immediate) so that the registers they would pin are free for
//   instruction-level parallelism across the four independent pool exponentials.
//   Why: the loop is a chain of dependent DFMAs (Horner polynomials) and is bound by the warp
//   schedulers' issue rate and dependent-issue latency, not by HBM.  ncu history (profiles/):
//     v1 thread per member, 255 regs, 2 warps/SMSP:  FP64 pipe 35 % busy, stall_wait 3.3 / issue
//     v2 warp per gas + named barrier, 96 regs:      FP64 42 %, 61 % of issue slots non-FP64
//     v3 + constants in c[3], params in smem:        FP64 52 %, 19 % of samples in barrier stalls
//     v4 lanes = (member, gas), shuffles, no barrier: issue slots 69 % busy, 413 instr / warp-step
//        of which 161 FP64; ~40 / step were per-row TMA issue (R2UR + UBLKCP, 12 rows per tile)
//     v5 (this file): one 3-D tensor-map TMA per warp-tile, unrolled tile body, running pointers.
//   The loop body is alpha_val -> step_conc -> step_forc -> (shuffle) -> step_temp, the names the
//   reference reserves in .coveragerc:12-19; `oxfair` is ONE launch.
//
// Memory system
//   * per-member emissions [gas][t][member] are streamed into the warp's shared-memory ring one
//     tile of TT time steps ahead by ONE tensor-map TMA per tile
//     (cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes, box =
//     MW members x TT steps x NGAS gases; out-of-range rows/columns are zero-filled by the
//     hardware, which is what makes ragged tails free), double-buffered on two mbarriers per warp.
//     Per-member external forcing rides the same way through a 2-D map.
//   * C / RF / T are written straight from registers with streaming (st.global.cs) stores; a warp
//     store covers NGAS contiguous MW-member runs, adjacent warps write adjacent runs -- nothing is
//     re-read.
//   * optional statistics: the gas-0 lanes add their member's T to the privatised per-step
//     histogram (RED.ADD.U32) as they go; moments come from a second, HBM-speed pass over the T
//     rows (ufair_abi.cu), which is cheaper than in-loop cross-lane reductions and deterministic.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ufair.h"
#include "ufair_math.cuh"

#ifndef UFAIR_WARPS
#define UFAIR_WARPS 4  // warps per CTA (a CTA is only a launch / shared-memory grouping)
#endif
#ifndef UFAIR_MINB_F64
#define UFAIR_MINB_F64 5  // resident CTAs per SM the register allocator must allow
#endif
#ifndef UFAIR_MINB_F32
#define UFAIR_MINB_F32 8
#endif
#ifndef UFAIR_TT
#define UFAIR_TT 4  // time steps per shared-memory tile
#endif

namespace ufair {

constexpr int kWarps = UFAIR_WARPS;
constexpr int kTT = UFAIR_TT;
constexpr int kStages = 2;  // tile ring depth

// members per warp: rows of MW elements must be a multiple of 16 bytes for the TMA box
constexpr int members_per_warp(int elem_size, int n_gas) {
  const int q = 16 / elem_size;  // elements per 16 bytes
  return (32 / n_gas) / q * q;
}

template <typename Real> struct KArgs {
  int n_gas, n_t;
  long long n_member, ld;
  int n_scen;
  int e_mode, fext_mode, t_mode, out_mask, stats, newton_iters, clamp;
  double dt, h, iirf_max;
  const Real* E;
  const int* scen_idx;
  const Real* e_scale;
  const Real* fext;
  const Real* gp;
  const Real* tp;
  const Real* state_in;
  Real* oC;
  Real* oRF;
  Real* oT;
  Real* oA;
  Real* state_out;
  int hist_bins, hist_copies, hist_t0, hist_rows;
  Real hist_lo, hist_invw;
  unsigned int* hist;
};

// ---- PTX wrappers: mbarrier, tensor-map TMA, shared-space loads/stores with 32-bit addresses ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
          dst),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          dst),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ double lds(uint32_t a, double) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ float lds(uint32_t a, float) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ void sts(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
// keep a loop-invariant value in a register (stops the compiler from re-deriving it every step)
__device__ __forceinline__ void pin(uint32_t& x) { asm volatile("" : "+r"(x)); }

template <typename Real> __device__ __forceinline__ void st_stream(Real* p, Real v) { __stcs(p, v); }

// per-lane derived constants held in shared memory, [param][lane]
enum {
  P_KA0 = 0,   // c a_i tau_i: equilibrium pool per unit (E alpha)
  P_K0 = 4,    // dt / tau_i   (ALPHA_ONE: m_i = 1 - exp(-dt/tau_i))
  P_RHO0 = 8,  // u = rho0 + rhoU Gcum + wR sumR + rhoT T   (= iIRF/g1 [+ ln g0])
  P_RHOU,
  P_WR,
  P_RHOT,
  P_UMAX,
  P_C0,
  P_INVC0,
  P_SQRTC0,
  P_F1,
  P_F2,
  P_F3,
  P_QM0,  // q_j (1 - exp(-dt/d_j))
  P_QM1,
  P_DEC0,  // exp(-dt/d_j)
  P_DEC1,
  P_X0,  // SINH: g0;  NEWTON: g1
  P_X1,  // NEWTON: ln g0
  P_X2,  // NEWTON: 1/c
  P_COUNT
};

// EXP / ONE need none of the P_X* rows, SINH one, NEWTON three
constexpr int par_count(int amode) {
  return amode == UFAIR_ALPHA_NEWTON ? P_COUNT : (amode == UFAIR_ALPHA_SINH ? P_X0 + 1 : P_X0);
}
constexpr size_t round128(size_t b) { return (b + 127) / 128 * 128; }

// per-WARP shared memory, in bytes (every piece 128-byte aligned: tensor-map TMA destinations)
template <typename Real, int NGAS, int AMODE> struct WarpSmem {
  static constexpr int MW = members_per_warp(sizeof(Real), NGAS);
  static constexpr uint32_t e_box = (uint32_t)(NGAS * kTT * MW * sizeof(Real));  // bytes one E box delivers
  static constexpr uint32_t f_box = (uint32_t)(kTT * MW * sizeof(Real));         // bytes one f_ext box delivers
  static constexpr uint32_t e_stage = (uint32_t)round128(e_box);
  static constexpr uint32_t f_stage = (uint32_t)round128(f_box);
  static constexpr uint32_t off_e = 0;
  static constexpr uint32_t off_f = off_e + kStages * e_stage;
  static constexpr uint32_t off_par = off_f + kStages * f_stage;
  static constexpr uint32_t off_bar = off_par + (uint32_t)round128(par_count(AMODE) * 32 * sizeof(Real));
  static constexpr uint32_t bytes = off_bar + 128;
  static constexpr size_t bytes_per_cta = (size_t)bytes * kWarps;
};

constexpr int min_blocks(int elem_size) { return elem_size == 8 ? UFAIR_MINB_F64 : UFAIR_MINB_F32; }

// EMEM: per-member emissions (TMA-staged tile) vs scenario-shared (read-only path + register prefetch)
template <typename Real, int NGAS, int AMODE, bool EMEM>
__global__ void __launch_bounds__(kWarps * 32, min_blocks(sizeof(Real)))
    ufair_integrate_kernel(const __grid_constant__ KArgs<Real> a, const __grid_constant__ CUtensorMap tmE,
                           const __grid_constant__ CUtensorMap tmF) {
  using M = Math<Real>;
  using WS = WarpSmem<Real, NGAS, AMODE>;
  constexpr int MW = WS::MW;
  constexpr int NACT = MW * NGAS;  // working lanes
  constexpr unsigned FULL = 0xffffffffu;
  constexpr uint32_t ES = sizeof(Real);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long wg = (long long)blockIdx.x * kWarps + warp;  // global warp index
  const long long m0 = wg * MW;                                // first member of this warp
  if (m0 >= a.n_member) return;  // whole warp: nothing to do (no CTA-wide sync exists)

  const int g = min(lane / MW, NGAS - 1);    // this lane's gas     (spare lanes shadow the last pair)
  const int i = min(lane - g * MW, MW - 1);  // this lane's member inside the warp
  const long long m_raw = m0 + i;
  const bool active = (lane < NACT) && (m_raw < a.n_member);
  const long long m = min(m_raw, a.n_member - 1);
  const long long ld = a.ld;
  const int n_t = a.n_t;

  const bool fx_member = (a.fext_mode == UFAIR_FEXT_MEMBER);
  const bool fx_scen = (a.fext_mode == UFAIR_FEXT_SCENARIO);
  const bool use_tma = EMEM || fx_member;

  const uint32_t wbase = smem_u32(smem_raw) + (uint32_t)warp * WS::bytes;
  uint32_t e_addr = wbase + WS::off_e + (uint32_t)(g * kTT * MW + i) * ES;  // this lane's E element, stage 0, tt 0
  uint32_t f_addr = wbase + WS::off_f + (uint32_t)i * ES;
  uint32_t par = wbase + WS::off_par + (uint32_t)lane * ES;                 // this lane's parameter column
  const uint32_t bar0 = wbase + WS::off_bar;
  pin(e_addr);
  pin(f_addr);
  pin(par);
#define PAR(k) lds(par + (uint32_t)(k) * 32u * ES, Real())
#define SETPAR(k, v) sts(par + (uint32_t)(k) * 32u * ES, (Real)(v))

  const int n_tile = (n_t + kTT - 1) / kTT;
  if (lane == 0 && use_tma) {
    for (int s = 0; s < kStages; ++s) mbar_init(bar0 + 8u * s, 1);
    fence_barrier_init();
  }
  __syncwarp();
  auto issue_tile = [&](int k) {  // lane 0 only: one bulk tensor copy per tile (two with per-member f_ext)
    const uint32_t s = (uint32_t)(k % kStages);
    const uint32_t bar = bar0 + 8u * s;
    mbar_expect_tx(bar, (EMEM ? WS::e_box : 0u) + (fx_member ? WS::f_box : 0u));
    if (EMEM) tma_load_3d(wbase + WS::off_e + s * WS::e_stage, &tmE, (int)m0, k * kTT, 0, bar);
    if (fx_member) tma_load_2d(wbase + WS::off_f + s * WS::f_stage, &tmF, (int)m0, k * kTT, bar);
  };
  if (use_tma && lane == 0 && n_tile > 0) issue_tile(0);

  // ---------------- prologue: raw parameters -> derived constants, in double for both precisions
  // (g_1 and g_0 of .coveragerc:15-16 are fused here; in FP32 they would cancel catastrophically
  // for the 10^6-year pool, so the one-off prologue always runs in FP64 and rounds once)
  Real R0, R1, R2, R3, Gcum, sumR;
  {
    const Real* p = a.gp + (long long)g * UFAIR_GP_COUNT * ld + m;
    double av[4], tau[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      av[q] = (double)p[(UFAIR_GP_A0 + q) * ld];
      tau[q] = (double)p[(UFAIR_GP_TAU0 + q) * ld];
    }
    const double r0 = p[UFAIR_GP_R0 * ld], rU = p[UFAIR_GP_RU * ld], rT = p[UFAIR_GP_RT * ld], rA = p[UFAIR_GP_RA * ld];
    const double C0d = p[UFAIR_GP_C0 * ld], c = p[UFAIR_GP_EMIS2CONC * ld];
    double g1 = 0.0, sden = 0.0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const double z = a.h / tau[q];
      const double ez = exp(-z);
      g1 += av[q] * tau[q] * (1.0 - (1.0 + z) * ez);
      sden += av[q] * tau[q] * (1.0 - ez);
    }
    const double sarg = sden / g1;
    const double inv_g1 = 1.0 / g1, invc = 1.0 / c;
    const double lng0 = -sarg;
    const double fold = (AMODE == UFAIR_ALPHA_SINH) ? 0.0 : lng0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      SETPAR(P_KA0 + q, c * av[q] * tau[q]);
      SETPAR(P_K0 + q, (AMODE == UFAIR_ALPHA_ONE) ? -expm1(-a.dt / tau[q]) : a.dt / tau[q]);
    }
    SETPAR(P_RHO0, r0 * inv_g1 + fold);
    SETPAR(P_RHOU, rU * inv_g1);
    SETPAR(P_WR, (rA - rU) * inv_g1 * invc);
    SETPAR(P_RHOT, rT * inv_g1);
    SETPAR(P_UMAX, a.clamp ? (a.iirf_max * inv_g1 + fold) : (double)INFINITY);
    SETPAR(P_C0, C0d);
    SETPAR(P_INVC0, 1.0 / C0d);
    SETPAR(P_SQRTC0, sqrt(C0d));
    SETPAR(P_F1, p[UFAIR_GP_F1 * ld]);
    SETPAR(P_F2, p[UFAIR_GP_F2 * ld]);
    SETPAR(P_F3, p[UFAIR_GP_F3 * ld]);
    if (AMODE == UFAIR_ALPHA_SINH) SETPAR(P_X0, 1.0 / sinh(sarg));
    if (AMODE == UFAIR_ALPHA_NEWTON) {
      SETPAR(P_X0, g1);
      SETPAR(P_X1, lng0);
      SETPAR(P_X2, invc);
    }
    const Real* si = a.state_in;
    R0 = si ? si[(long long)(5 * g + 0) * ld + m] : Real(0);
    R1 = si ? si[(long long)(5 * g + 1) * ld + m] : Real(0);
    R2 = si ? si[(long long)(5 * g + 2) * ld + m] : Real(0);
    R3 = si ? si[(long long)(5 * g + 3) * ld + m] : Real(0);
    Gcum = si ? si[(long long)(5 * g + 4) * ld + m] : Real(0);
    sumR = (R0 + R1) + (R2 + R3);
  }
  Real S0, S1, Tprev;
  {
    const Real* tp = a.tp + m;
    const Real* si = a.state_in;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const double qq = tp[(UFAIR_TP_Q1 + q) * ld], d = tp[(UFAIR_TP_D1 + q) * ld];
      const double mj = -expm1(-a.dt / d);
      SETPAR(P_QM0 + q, qq * mj);
      SETPAR(P_DEC0 + q, 1.0 - mj);
    }
    S0 = si ? si[(long long)(5 * NGAS + 0) * ld + m] : Real(0);
    S1 = si ? si[(long long)(5 * NGAS + 1) * ld + m] : Real(0);
    Tprev = si ? si[(long long)(5 * NGAS + 2) * ld + m] : Real(0);
  }
  __syncwarp();
  // a zero forcing coefficient means a zero term; the log / sqrt is skipped only when no lane of
  // the warp needs it (voted once, outside the loop)
  const Real f1v = PAR(P_F1), f3v = PAR(P_F3);
  const bool need_log = __any_sync(FULL, f1v != Real(0));
  const bool need_sqrt = __any_sync(FULL, f3v != Real(0));
  unsigned mk1 = (f1v != Real(0)) ? 0xffffffffu : 0u, mk3 = (f3v != Real(0)) ? 0xffffffffu : 0u;
  pin(mk1);
  pin(mk3);

  const int scen = (a.scen_idx != nullptr) ? a.scen_idx[m] : 0;
  const Real esc = (!EMEM && a.e_scale) ? a.e_scale[(long long)g * ld + m] : Real(1);
  const Real dt = (Real)a.dt;
  const Real hdt = (Real)(a.h / a.dt);
  const bool t_mid = (a.t_mode == UFAIR_T_MID);

  // output predicates packed in one register; running output pointers
  const bool do_hist = a.stats && active && (g == 0);
  unsigned wm = (active ? (unsigned)(a.out_mask & (UFAIR_OUT_C | UFAIR_OUT_RF | UFAIR_OUT_ALPHA)) : 0u) |
                ((active && g == 0 && ((a.out_mask & UFAIR_OUT_T) || a.stats)) ? (unsigned)UFAIR_OUT_T : 0u) |
                (do_hist ? 16u : 0u);
  pin(wm);
  const long long o_gas = (long long)g * n_t * ld + m_raw;
  Real* pC = a.oC + o_gas;
  Real* pRF = a.oRF + o_gas;
  Real* pA = a.oA + o_gas;
  Real* pT = a.oT + m_raw;
  unsigned int* hrow = a.stats ? a.hist + ((size_t)(wg % a.hist_copies) * a.hist_rows + a.hist_t0) * a.hist_bins : nullptr;
  const int bins_m1 = a.hist_bins - 1;

  // scenario-mode inputs: register prefetch one step ahead through the read-only path
  Real e_next = 0, fx_next = 0;
  const Real* e_scen = a.E + ((long long)g * n_t) * a.n_scen + scen;
  const Real* fx_scen_p = a.fext + scen;
  if (!EMEM && n_t > 0) e_next = __ldg(e_scen);
  if (fx_scen && n_t > 0) fx_next = __ldg(fx_scen_p);

  // ---------------- one time step ------------------------------------------------------------------
  auto step = [&](const int t, const uint32_t tt_off) {
    Real e, fx = 0;
    if (EMEM) {
      e = lds(e_addr + tt_off, Real());
    } else {
      e = e_next * esc;
      e_next = __ldg(e_scen + (long long)min(t + 1, n_t - 1) * a.n_scen);
    }
    if (fx_member) fx = lds(f_addr + tt_off, Real());
    if (fx_scen) {
      fx = fx_next;
      fx_next = __ldg(fx_scen_p + (long long)min(t + 1, n_t - 1) * a.n_scen);
    }
    // ---- alpha_val: state at t-1 -> alpha, 1/alpha
    Real alpha, inva;
    if (AMODE == UFAIR_ALPHA_ONE) {
      alpha = Real(1);
      inva = Real(1);
    } else {
      Real u = fma(PAR(P_RHOU), Gcum, fma(PAR(P_WR), sumR, fma(PAR(P_RHOT), Tprev, PAR(P_RHO0))));
      const Real umax = PAR(P_UMAX);
      u = (u > umax) ? umax : u;
      alpha = (AMODE == UFAIR_ALPHA_SINH) ? PAR(P_X0) * M::sinh_pair(u) : M::exp_(u);
      if (AMODE == UFAIR_ALPHA_NEWTON) {
        const Real iirf = (u - PAR(P_X1)) * PAR(P_X0);
        const Real invc = PAR(P_X2);
        for (int it = 0; it < a.newton_iters; ++it) {
          const Real ia = M::rcp(alpha);
          Real f = -iirf, fp = 0;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const Real z = PAR(P_K0 + q) * hdt * ia;
            const Real mz = M::decay(z);
            const Real at = PAR(P_KA0 + q) * invc;  // a_i tau_i
            f = fma(at * alpha, mz, f);
            fp = fma(at, mz - z * (Real(1) - mz), fp);
          }
          const Real an = alpha - f * M::rcp(fp);
          alpha = M::fmax_(an, Real(0.5) * alpha);
        }
      }
      inva = M::rcp(alpha);
    }
    // ---- step_conc: relax each pool toward its equilibrium  E alpha c a_i tau_i
    const Real ea = e * alpha;
    {
      const Real m0_ = (AMODE == UFAIR_ALPHA_ONE) ? PAR(P_K0 + 0) : M::decay(PAR(P_K0 + 0) * inva);
      const Real m1_ = (AMODE == UFAIR_ALPHA_ONE) ? PAR(P_K0 + 1) : M::decay(PAR(P_K0 + 1) * inva);
      const Real m2_ = (AMODE == UFAIR_ALPHA_ONE) ? PAR(P_K0 + 2) : M::decay(PAR(P_K0 + 2) * inva);
      const Real m3_ = (AMODE == UFAIR_ALPHA_ONE) ? PAR(P_K0 + 3) : M::decay(PAR(P_K0 + 3) * inva);
      R0 = fma(m0_, fma(ea, PAR(P_KA0 + 0), -R0), R0);
      R1 = fma(m1_, fma(ea, PAR(P_KA0 + 1), -R1), R1);
      R2 = fma(m2_, fma(ea, PAR(P_KA0 + 2), -R2), R2);
      R3 = fma(m3_, fma(ea, PAR(P_KA0 + 3), -R3), R3);
    }
    Gcum = fma(e, dt, Gcum);
    sumR = (R0 + R1) + (R2 + R3);
    const Real C = PAR(P_C0) + sumR;
    // ---- step_forc
    Real F = PAR(P_F2) * sumR;
    if (need_log) {
      const Real lt = PAR(P_F1) * M::log_(C * PAR(P_INVC0));
      F += M::mask(lt, mk1);
    }
    if (need_sqrt) {
      const Real st = PAR(P_F3) * (M::sqrt_(C) - PAR(P_SQRTC0));
      F += M::mask(st, mk3);
    }
    if (wm & UFAIR_OUT_C) st_stream(pC, C);
    if (wm & UFAIR_OUT_RF) st_stream(pRF, F);
    if (wm & UFAIR_OUT_ALPHA) st_stream(pA, alpha);
    pC += ld;
    pRF += ld;
    pA += ld;
    // ---- total forcing of the member: fixed order g = 0..NGAS-1, identical in its NGAS lanes
    Real Ftot = fx;
#pragma unroll
    for (int gg = 0; gg < NGAS; ++gg) Ftot += __shfl_sync(FULL, F, gg * MW + i);
    // ---- step_temp (computed redundantly, bit-identically, by the NGAS lanes of a member)
    const Real s0 = fma(PAR(P_QM0), Ftot, S0 * PAR(P_DEC0));
    const Real s1 = fma(PAR(P_QM1), Ftot, S1 * PAR(P_DEC1));
    const Real T = t_mid ? Real(0.5) * ((S0 + s0) + (S1 + s1)) : (s0 + s1);
    S0 = s0;
    S1 = s1;
    Tprev = T;
    if (wm & UFAIR_OUT_T) st_stream(pT, T);
    pT += ld;
    if (wm & 16u) {  // gas-0 lane of a real member: one histogram count
      const Real x = M::bin_x(T, a.hist_lo, a.hist_invw);
      if (x == x) {
        const int b = max(0, min(bins_m1, M::floor_to_int(x)));
        atomicAdd(hrow + b, 1u);
      }
    }
    hrow += a.hist_bins;  // next step's histogram row (dangling but unused when stats are off)
  };

  // ---------------- the time loop --------------------------------------------------------------------
  for (int k = 0; k < n_tile; ++k) {
    const uint32_t s = (uint32_t)(k % kStages);
    const int t0 = k * kTT;
    if (use_tma) {
      if (lane == 0 && k + 1 < n_tile) issue_tile(k + 1);  // that stage was drained before the last __syncwarp
      mbar_wait(bar0 + 8u * s, (uint32_t)((k / kStages) & 1));
    }
    // not unrolled on purpose: the 24 FP64 polynomial / reduction constants stay hoisted in uniform
    // registers only while the body is a single copy (unrolling x4 spilled them back to LDC + moves)
    const int nt = min(kTT, n_t - t0);
    uint32_t tt_off = s * WS::e_stage;
    const uint32_t f_adj = s * WS::f_stage - s * WS::e_stage;  // f_addr + tt_off + f_adj addresses the f_ext stage
    f_addr += f_adj;
#pragma unroll 1
    for (int tt = 0; tt < nt; ++tt) {
      step(t0 + tt, tt_off);
      tt_off += (uint32_t)MW * ES;
    }
    f_addr -= f_adj;
    __syncwarp();  // every lane is done with stage s before lane 0 refills it
  }

  // ---------------- epilogue: final state (checkpoint / resume) ---------------------------------
  if (a.state_out && active) {
    Real* so = a.state_out;
    so[(long long)(5 * g + 0) * ld + m] = R0;
    so[(long long)(5 * g + 1) * ld + m] = R1;
    so[(long long)(5 * g + 2) * ld + m] = R2;
    so[(long long)(5 * g + 3) * ld + m] = R3;
    so[(long long)(5 * g + 4) * ld + m] = Gcum;
    if (g == 0) {
      so[(long long)(5 * NGAS + 0) * ld + m] = S0;
      so[(long long)(5 * NGAS + 1) * ld + m] = S1;
      so[(long long)(5 * NGAS + 2) * ld + m] = Tprev;
    }
  }
#undef PAR
#undef SETPAR
}

// one launcher per (Real, NGAS, AMODE); defined in ufair_inst_*.cu
template <typename Real, int NGAS, int AMODE>
cudaError_t launch_integrate(const KArgs<Real>& a, const CUtensorMap& tmE, const CUtensorMap& tmF, cudaStream_t stream);

#define UFAIR_DEFINE_LAUNCH(Real, NGAS, AMODE)                                                                     \
  template <>                                                                                                      \
  cudaError_t launch_integrate<Real, NGAS, AMODE>(const KArgs<Real>& a, const CUtensorMap& tmE,                    \
                                                  const CUtensorMap& tmF, cudaStream_t stream) {                   \
    using WS = WarpSmem<Real, NGAS, AMODE>;                                                                        \
    const size_t smem = WS::bytes_per_cta;                                                                         \
    auto kern = (a.e_mode == UFAIR_E_MEMBER) ? ufair_integrate_kernel<Real, NGAS, AMODE, true>                     \
                                              : ufair_integrate_kernel<Real, NGAS, AMODE, false>;                  \
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);            \
    if (e != cudaSuccess) return e;                                                                                \
    const long long n_warp = (a.n_member + WS::MW - 1) / WS::MW;                                                   \
    const unsigned grid = (unsigned)((n_warp + kWarps - 1) / kWarps);                                              \
    kern<<<grid, kWarps * 32, smem, stream>>>(a, tmE, tmF);                                                        \
    return cudaGetLastError();                                                                                     \
  }

}  // namespace ufair
