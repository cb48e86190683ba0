// ufair_kernel.cuh -- the fused Universal-FaIR time-stepping kernel (sm_100a).
//
// Work decomposition (v2): one WARP = 32 consecutive ensemble members of ONE gas.
//   A CTA owns MEMB = 32*W members and has NGAS*W warps; the NGAS warps that share a 32-member
//   group meet once per time step at a named barrier to exchange their radiative forcings through
//   shared memory.  Every thread keeps its gas's four pools, cumulative emissions, derived
//   parameters and (redundantly, bit-identically in the NGAS threads of a member) the two thermal
//   boxes in REGISTERS across the serial time loop.  Compared with one thread per member this
//   cuts registers per thread ~3x, which is what buys the 16+ resident warps per SM the FP64 pipe
//   needs to stay fed: the loop is a chain of dependent DFMAs (Horner polynomials), so it is
//   latency-bound unless enough independent warps interleave (ncu on v1: 2 warps/SMSP,
//   stall_wait 3.3 cycles per issue, FP64 pipe 35 % busy -- profiles/r1_v1_summary.md).
//   The loop body is alpha_val -> step_conc -> step_forc -> (exchange) -> step_temp, the names the
//   reference reserves in .coveragerc:12-19; `oxfair` is ONE launch.
//
// Memory system
//   * per-member emissions [gas][t][member] (and per-member external forcing) are streamed into
//     shared memory one tile of TT time steps ahead with TMA bulk copies
//     (cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes, one row of MEMB members
//     per (gas, t)), double-buffered on two mbarriers; each thread reads its own column
//     (conflict-free LDS).
//   * C / RF / T rows are written straight from registers with streaming (st.global.cs) stores;
//     each warp store covers one full, aligned 256-byte (f64) run -- nothing is re-read.
//   * optional statistics: each tile's T values are staged in shared memory and folded by the
//     whole CTA into privatised per-step histograms (RED.ADD.U32) and moments (warp-shuffle
//     reduction, then one RED per warp-row).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ufair.h"
#include "ufair_math.cuh"

#ifndef UFAIR_MEMB
#define UFAIR_MEMB 64  // members per CTA (multiple of 32)
#endif
#ifndef UFAIR_MINB_F64
#define UFAIR_MINB_F64 3  // resident CTAs per SM the register allocator must allow (3-gas shape)
#endif
#ifndef UFAIR_MINB_F32
#define UFAIR_MINB_F32 5
#endif

namespace ufair {

constexpr int kMemb = UFAIR_MEMB;  // members per CTA
constexpr int kW = kMemb / 32;     // 32-member groups per CTA
constexpr int kTT = 4;             // time steps per shared-memory tile
constexpr int kStages = 2;         // tile ring depth
static_assert(kMemb % 32 == 0 && kW >= 1 && kW <= 8, "UFAIR_MEMB must be 32..256 in steps of 32");

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
  double* mom;
};

// ---- small PTX wrappers: mbarrier + TMA bulk copy + named barrier -----------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_row(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

template <typename Real> __device__ __forceinline__ void st_stream(Real* p, Real v) { __stcs(p, v); }

// order-preserving map double -> uint64 (so atomicMin/Max on integers orders doubles)
__device__ __forceinline__ unsigned long long enc_ordered(double x) {
  unsigned long long u = (unsigned long long)__double_as_longlong(x);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__host__ __device__ __forceinline__ double dec_ordered(unsigned long long u) {
  unsigned long long b = (u >> 63) ? (u & 0x7fffffffffffffffull) : ~u;
#ifdef __CUDA_ARCH__
  return __longlong_as_double((long long)b);
#else
  double d;
  memcpy(&d, &b, sizeof(d));
  return d;
#endif
}

template <typename Real, int NGAS> struct SmemLayout {
  static constexpr size_t tile_elems = (size_t)kStages * kTT * (NGAS + 1) * kMemb;  // E (+f_ext) ring
  static constexpr size_t ttile_elems = (size_t)2 * kTT * kMemb;                    // T staging (stats)
  static constexpr size_t fx_elems = (size_t)2 * NGAS * kMemb;                      // forcing exchange
  static constexpr size_t bytes = sizeof(Real) * (tile_elems + ttile_elems + fx_elems) + kStages * sizeof(uint64_t);
};

// resident CTAs per SM requested from the register allocator, scaled so that the resident THREAD
// count stays the same for every gas count (a CTA has NGAS * MEMB threads)
constexpr int min_blocks(int elem_size, int n_gas) {
  const int b = (elem_size == 8 ? UFAIR_MINB_F64 : UFAIR_MINB_F32) * 3 / n_gas;
  return b < 1 ? 1 : b;
}

template <typename Real, int NGAS, int AMODE>
__global__ void __launch_bounds__(NGAS* kMemb, min_blocks(sizeof(Real), NGAS))
    ufair_integrate_kernel(const __grid_constant__ KArgs<Real> a) {
  using M = Math<Real>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int mg = warp % kW;  // 32-member group inside the CTA
  const int g = warp / kW;   // this warp's gas
  const int j = mg * 32 + lane;
  const long long m0 = (long long)blockIdx.x * kMemb;
  const long long m_raw = m0 + j;
  const bool active = m_raw < a.n_member;
  const long long m = active ? m_raw : (a.n_member - 1);  // idle lanes shadow the last member
  const long long ld = a.ld;
  const int n_t = a.n_t;

  const bool e_member = (a.e_mode == UFAIR_E_MEMBER);
  const bool fx_member = (a.fext_mode == UFAIR_FEXT_MEMBER);
  const bool fx_scen = (a.fext_mode == UFAIR_FEXT_SCENARIO);
  const int n_row = (e_member ? NGAS : 0) + (fx_member ? 1 : 0);  // rows per time step in a tile
  const int fx_row = e_member ? NGAS : 0;

  using SL = SmemLayout<Real, NGAS>;
  Real* tile = reinterpret_cast<Real*>(smem_raw);
  Real* ttile = tile + SL::tile_elems;
  Real* fxch = ttile + SL::ttile_elems;
  uint64_t* full = reinterpret_cast<uint64_t*>(fxch + SL::fx_elems);

  const int valid_cols = (int)min((long long)kMemb, ld - m0);      // columns that exist in memory
  const uint32_t row_bytes = (uint32_t)valid_cols * sizeof(Real);  // multiple of 16 (ld % (16/sizeof) == 0)
  const int n_tile = (n_t + kTT - 1) / kTT;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
    fence_barrier_init();
  }
  __syncthreads();

  auto issue_tile = [&](int k) {  // thread 0 only
    const int s = k % kStages;
    const int t0 = k * kTT;
    const int nt = min(kTT, n_t - t0);
    mbar_expect_tx(&full[s], (uint32_t)(nt * n_row) * row_bytes);
    Real* dst = tile + (size_t)s * kTT * (NGAS + 1) * kMemb;
    for (int tt = 0; tt < nt; ++tt) {
      if (e_member) {
#pragma unroll
        for (int gg = 0; gg < NGAS; ++gg)
          tma_load_row(dst + (tt * (NGAS + 1) + gg) * kMemb, a.E + ((long long)gg * n_t + (t0 + tt)) * ld + m0,
                       row_bytes, &full[s]);
      }
      if (fx_member)
        tma_load_row(dst + (tt * (NGAS + 1) + fx_row) * kMemb, a.fext + (long long)(t0 + tt) * ld + m0, row_bytes,
                     &full[s]);
    }
  };
  if (n_row > 0 && tid == 0 && n_tile > 0) issue_tile(0);

  // ---------------- prologue: raw parameters -> derived constants, in double for both precisions
  // (g_1 and g_0 of .coveragerc:15-16 are fused here; in FP32 they would cancel catastrophically
  // for the 10^6-year pool, so the one-off prologue always runs in FP64 and rounds once)
  Real kA[4];  // c a_i tau_i               (equilibrium pool per unit E*alpha)
  Real kk[4];  // dt / tau_i                (ALPHA_ONE: m_i = 1 - exp(-dt/tau_i) instead)
  Real rho0, rhoU, wR, rhoT, umax;  // u = rho0 + rhoU Gcum + wR sumR + rhoT T   (= iIRF/g1 [+ ln g0])
  Real C0, invC0, sqrtC0, f1, f2, f3;
  Real g0s = 0, g1n = 0, lng0n = 0, invcn = 0;  // SINH: g0; NEWTON: g1, ln g0, 1/c
  Real R[4], Gcum, sumR;
  {
    const Real* p = a.gp + (long long)g * UFAIR_GP_COUNT * ld + m;
    double av[4], tau[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      av[i] = (double)p[(UFAIR_GP_A0 + i) * ld];
      tau[i] = (double)p[(UFAIR_GP_TAU0 + i) * ld];
    }
    const double r0 = p[UFAIR_GP_R0 * ld], rU = p[UFAIR_GP_RU * ld], rT = p[UFAIR_GP_RT * ld], rA = p[UFAIR_GP_RA * ld];
    const double C0d = p[UFAIR_GP_C0 * ld], c = p[UFAIR_GP_EMIS2CONC * ld];
    double g1 = 0.0, sden = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const double z = a.h / tau[i];
      const double ez = exp(-z);
      g1 += av[i] * tau[i] * (1.0 - (1.0 + z) * ez);
      sden += av[i] * tau[i] * (1.0 - ez);
    }
    const double sarg = sden / g1;
    const double inv_g1 = 1.0 / g1, invc = 1.0 / c;
    const double lng0 = -sarg;
    const double fold = (AMODE == UFAIR_ALPHA_SINH) ? 0.0 : lng0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      kA[i] = (Real)(c * av[i] * tau[i]);
      kk[i] = (Real)((AMODE == UFAIR_ALPHA_ONE) ? -expm1(-a.dt / tau[i]) : a.dt / tau[i]);
    }
    rho0 = (Real)(r0 * inv_g1 + fold);
    rhoU = (Real)(rU * inv_g1);
    wR = (Real)((rA - rU) * inv_g1 * invc);
    rhoT = (Real)(rT * inv_g1);
    umax = a.clamp ? (Real)(a.iirf_max * inv_g1 + fold) : (Real)INFINITY;
    C0 = (Real)C0d;
    invC0 = (Real)(1.0 / C0d);
    sqrtC0 = (Real)sqrt(C0d);
    f1 = p[UFAIR_GP_F1 * ld];
    f2 = p[UFAIR_GP_F2 * ld];
    f3 = p[UFAIR_GP_F3 * ld];
    if (AMODE == UFAIR_ALPHA_SINH) g0s = (Real)(1.0 / sinh(sarg));
    if (AMODE == UFAIR_ALPHA_NEWTON) {
      g1n = (Real)g1;
      lng0n = (Real)lng0;
      invcn = (Real)invc;
    }
    if (a.state_in) {
#pragma unroll
      for (int i = 0; i < 4; ++i) R[i] = a.state_in[(long long)(5 * g + i) * ld + m];
      Gcum = a.state_in[(long long)(5 * g + 4) * ld + m];
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) R[i] = 0;
      Gcum = 0;
    }
    sumR = (R[0] + R[1]) + (R[2] + R[3]);
  }
  Real qm[2], dec[2], S[2], Tprev;
  {
    const Real* tp = a.tp + m;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const double q = tp[(UFAIR_TP_Q1 + i) * ld], d = tp[(UFAIR_TP_D1 + i) * ld];
      const double mj = -expm1(-a.dt / d);
      qm[i] = (Real)(q * mj);
      dec[i] = (Real)(1.0 - mj);
      S[i] = a.state_in ? a.state_in[(long long)(5 * NGAS + i) * ld + m] : Real(0);
    }
    Tprev = a.state_in ? a.state_in[(long long)(5 * NGAS + 2) * ld + m] : Real(0);
  }
  const int scen = (a.scen_idx != nullptr) ? a.scen_idx[m] : 0;
  const Real esc = (!e_member && a.e_scale) ? a.e_scale[(long long)g * ld + m] : Real(1);
  const Real dt = (Real)a.dt;
  const Real hdt = (Real)(a.h / a.dt);

  const bool wC = (a.out_mask & UFAIR_OUT_C) && active, wRF = (a.out_mask & UFAIR_OUT_RF) && active;
  const bool wT = (a.out_mask & UFAIR_OUT_T) && active && (g == 0), wA = (a.out_mask & UFAIR_OUT_ALPHA) && active;
  const long long gas_off = (long long)g * n_t * ld;
  const int hist_copy = a.stats ? (int)(blockIdx.x % (unsigned)a.hist_copies) : 0;
  const int n_valid = (int)min((long long)kMemb, a.n_member - m0);  // real members in this CTA
  const bool stage_T = a.stats && (g == 0);

  // scenario-mode inputs: register prefetch one step ahead through the read-only path
  Real e_next = 0, fx_next = 0;
  if (!e_member && n_t > 0) e_next = __ldg(a.E + ((long long)g * n_t) * a.n_scen + scen);
  if (fx_scen && n_t > 0) fx_next = __ldg(a.fext + scen);

  // ---------------- the time loop --------------------------------------------------------------
  for (int k = 0; k < n_tile; ++k) {
    const int s = k % kStages;
    const int t0 = k * kTT;
    const int nt = min(kTT, n_t - t0);
    if (n_row > 0) {
      if (tid == 0 && k + 1 < n_tile) issue_tile(k + 1);  // its stage was drained before the last barrier
      mbar_wait(&full[s], (uint32_t)((k / kStages) & 1));
    }
    const Real* trow = tile + (size_t)s * kTT * (NGAS + 1) * kMemb + j;
    Real* tt_buf = ttile + (k & 1) * kTT * kMemb;

    for (int tt = 0; tt < nt; ++tt) {
      const int t = t0 + tt;
      Real e, fx = 0;
      if (e_member) {
        e = trow[(tt * (NGAS + 1) + g) * kMemb];
      } else {
        e = e_next * esc;
        e_next = __ldg(a.E + ((long long)g * n_t + min(t + 1, n_t - 1)) * a.n_scen + scen);
      }
      if (fx_member) fx = trow[(tt * (NGAS + 1) + fx_row) * kMemb];
      if (fx_scen) {
        fx = fx_next;
        fx_next = __ldg(a.fext + (long long)min(t + 1, n_t - 1) * a.n_scen + scen);
      }

      // ---- alpha_val: state at t-1 -> alpha, 1/alpha
      Real alpha, inva;
      if (AMODE == UFAIR_ALPHA_ONE) {
        alpha = Real(1);
        inva = Real(1);
      } else {
        Real u = fma(rhoU, Gcum, fma(wR, sumR, fma(rhoT, Tprev, rho0)));
        u = (u > umax) ? umax : u;
        alpha = (AMODE == UFAIR_ALPHA_SINH) ? g0s * M::sinh_pair(u) : M::exp_(u);
        if (AMODE == UFAIR_ALPHA_NEWTON) {
          const Real iirf = (u - lng0n) * g1n;
          for (int it = 0; it < a.newton_iters; ++it) {
            const Real ia = M::rcp(alpha);
            Real f = -iirf, fp = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const Real z = kk[i] * hdt * ia;
              const Real mz = M::decay(z);
              const Real at = kA[i] * invcn;  // a_i tau_i
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
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const Real mi = (AMODE == UFAIR_ALPHA_ONE) ? kk[i] : M::decay(kk[i] * inva);
        R[i] = fma(mi, fma(ea, kA[i], -R[i]), R[i]);
      }
      Gcum = fma(e, dt, Gcum);
      sumR = (R[0] + R[1]) + (R[2] + R[3]);
      const Real C = C0 + sumR;
      // ---- step_forc; a zero coefficient means a zero term.  A term is skipped only when the whole
      //      warp agrees -- a warp holds ONE gas, so with gas-uniform parameter tables this never
      //      diverges (CH4/N2O warps skip the log, CO2 warps the sqrt)
      Real F = f2 * sumR;
      if (__any_sync(0xffffffffu, f1 != Real(0))) {
        const Real lt = f1 * M::log_(C * invC0);
        F += (f1 != Real(0)) ? lt : Real(0);
      }
      if (__any_sync(0xffffffffu, f3 != Real(0))) {
        const Real st = f3 * (M::sqrt_(C) - sqrtC0);
        F += (f3 != Real(0)) ? st : Real(0);
      }
      const long long orow = (long long)t * ld + m_raw;
      if (wC) st_stream(a.oC + gas_off + orow, C);
      if (wRF) st_stream(a.oRF + gas_off + orow, F);
      if (wA) st_stream(a.oA + gas_off + orow, alpha);

      // ---- exchange the per-gas forcings of this 32-member group (double-buffered by step parity)
      Real* fxb = fxch + (t & 1) * NGAS * kMemb;
      Real Ftot;
      if (NGAS > 1) {
        fxb[g * kMemb + j] = F;
        named_bar_sync(1 + mg, NGAS * 32);
        Ftot = fx;
#pragma unroll
        for (int gg = 0; gg < NGAS; ++gg) Ftot += fxb[gg * kMemb + j];
      } else {
        Ftot = fx + F;
      }
      // ---- step_temp (computed redundantly, bit-identically, by the NGAS threads of a member)
      const Real s0 = fma(qm[0], Ftot, S[0] * dec[0]);
      const Real s1 = fma(qm[1], Ftot, S[1] * dec[1]);
      const Real T = (a.t_mode == UFAIR_T_MID) ? Real(0.5) * ((S[0] + s0) + (S[1] + s1)) : (s0 + s1);
      S[0] = s0;
      S[1] = s1;
      Tprev = T;
      if (wT) st_stream(a.oT + orow, T);
      if (stage_T) tt_buf[tt * kMemb + j] = T;
    }

    __syncthreads();  // tile k fully consumed (E stage reusable) and its T values staged

    if (a.stats) {
      for (int tt = warp; tt < nt; tt += NGAS * kW) {
        const int row = a.hist_t0 + t0 + tt;
        unsigned int* hrow = a.hist + ((size_t)hist_copy * a.hist_rows + row) * a.hist_bins;
        double sm = 0.0, ss = 0.0, mn = INFINITY, mx = -INFINITY;
        for (int c = lane; c < n_valid; c += 32) {
          const Real Tv = tt_buf[tt * kMemb + c];
          const Real x = M::bin_x(Tv, a.hist_lo, a.hist_invw);
          if (x == x) {
            const Real fl = M::floor_(x);
            const int b = fl < Real(0) ? 0 : (fl > Real(a.hist_bins - 1) ? a.hist_bins - 1 : (int)fl);
            atomicAdd(hrow + b, 1u);
          }
          const double v = (double)Tv;
          sm += v;
          ss = fma(v, v, ss);
          mn = fmin(mn, v);
          mx = fmax(mx, v);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          sm += __shfl_xor_sync(0xffffffffu, sm, off);
          ss += __shfl_xor_sync(0xffffffffu, ss, off);
          mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, off));
          mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, off));
        }
        if (lane == 0) {
          double* mr = a.mom + ((size_t)hist_copy * a.hist_rows + row) * UFAIR_MOM_COUNT;
          atomicAdd(mr + UFAIR_MOM_SUM, sm);
          atomicAdd(mr + UFAIR_MOM_SUMSQ, ss);
          atomicMin(reinterpret_cast<unsigned long long*>(mr + UFAIR_MOM_MIN), enc_ordered(mn));
          atomicMax(reinterpret_cast<unsigned long long*>(mr + UFAIR_MOM_MAX), enc_ordered(mx));
        }
      }
    }
  }

  // ---------------- epilogue: final state (checkpoint / resume) ---------------------------------
  if (a.state_out && active) {
#pragma unroll
    for (int i = 0; i < 4; ++i) a.state_out[(long long)(5 * g + i) * ld + m] = R[i];
    a.state_out[(long long)(5 * g + 4) * ld + m] = Gcum;
    if (g == 0) {
      a.state_out[(long long)(5 * NGAS + 0) * ld + m] = S[0];
      a.state_out[(long long)(5 * NGAS + 1) * ld + m] = S[1];
      a.state_out[(long long)(5 * NGAS + 2) * ld + m] = Tprev;
    }
  }
}

// one launcher per (Real, NGAS, AMODE); defined in ufair_inst_*.cu
template <typename Real, int NGAS, int AMODE> cudaError_t launch_integrate(const KArgs<Real>& a, cudaStream_t stream);

#define UFAIR_DEFINE_LAUNCH(Real, NGAS, AMODE)                                                             \
  template <> cudaError_t launch_integrate<Real, NGAS, AMODE>(const KArgs<Real>& a, cudaStream_t stream) { \
    const size_t smem = SmemLayout<Real, NGAS>::bytes;                                                     \
    auto kern = ufair_integrate_kernel<Real, NGAS, AMODE>;                                                 \
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
    if (e != cudaSuccess) return e;                                                                        \
    const unsigned grid = (unsigned)((a.n_member + kMemb - 1) / kMemb);                                    \
    kern<<<grid, NGAS * kMemb, smem, stream>>>(a);                                                         \
    return cudaGetLastError();                                                                             \
  }

}  // namespace ufair
