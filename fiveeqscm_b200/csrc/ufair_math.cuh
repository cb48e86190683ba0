// ufair_math.cuh -- device math for the Universal-FaIR integrator (sm_100a).
//
// The FP64 pipe is the binding unit of the FP64 hot loop (15 exp + 3 log + 3 sqrt + 3 rcp per
// member-step), so the transcendentals are hand-rolled here with exactly the properties the
// model needs, instead of calling the general-purpose CUDA libm versions:
//   decay(x) = 1 - exp(-x)  accurate for tiny x (no cancellation; the oracle uses -expm1(-x)),
//   exp_(u)                 for alpha = exp(u), saturating at 2^+-40,
//   rcp / sqrt              MUFU seed + Newton, no special-case slow paths (operands are
//                           positive normal numbers on this path; edge cases handled explicitly),
//   log                     atanh-series with a reciprocal instead of a division.
// Polynomial coefficients come from tools/gen_poly.py (near-minimax, error re-measured with the
// rounded coefficients): expm1 Q deg 9 max rel err 4.1e-17; log L deg 6 4.6e-18 (f64);
// expm1 Q deg 4 2.3e-8; log via MUFU.LG2 (f32).
// All coefficients live in __constant__ memory so that each DFMA takes its coefficient as a
// constant-bank operand (c[3][..]) instead of two MOVs into a register pair -- ncu on v2 showed
// 17 % of all issued instructions were UMOV / IMAD.MOV materialising 64-bit immediates.
// Measured accuracy vs extended precision on the GPU: tests/test_gpu_math.py.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

// 1: FP32 exponentials on MUFU.EX2 (0: range reduction + polynomial on the FMA pipe)
#ifndef UFAIR_F32_MUFU
#define UFAIR_F32_MUFU 1
#endif
// 1: sqrt(0) = 0 through an integer clamp of the MUFU.RSQ64H seed (one instruction) instead of a test on
// the argument and two selects (four)
#ifndef UFAIR_SQRT_SEED_CLAMP
#define UFAIR_SQRT_SEED_CLAMP 1
#endif

namespace ufair {

constexpr unsigned kExpTableBytes = 0u;  // (a table-driven exponential was measured and dropped: DESIGN.md 4.1)

// expm1(r) = r + r^2 Q(r), |r| <= ln2/2; Q[0] + Q[1] r + ... + Q[9] r^9
static __constant__ double cExpQ[10] = {
    0x1.0000000000001p-1, 0x1.5555555555556p-3, 0x1.5555555553d68p-5, 0x1.11111111109b5p-7, 0x1.6c16c17889ef1p-10,
    0x1.a01a01a7c2efep-13, 0x1.a019b9149a41cp-16, 0x1.71de0db2f6b19p-19, 0x1.28917c89a43a7p-22, 0x1.af389ecfc4b9cp-26};
// log(m) = 2 s + s^3 L(s^2), s = (m-1)/(m+1)
static __constant__ double cLogL[7] = {0x1.5555555555558p-1, 0x1.99999999952e2p-2, 0x1.2492492df148dp-2,
                                       0x1.c71c62e5800a1p-3, 0x1.7462b4ab2ef6bp-3, 0x1.39fe606542ddep-3,
                                       0x1.2b584aae78a57p-3};
// log2(e), -ln2_hi, -ln2_lo, magic (2^52+2^51), ln2_hi, ln2_lo, 2^52+2^31
static __constant__ double cK[7] = {0x1.71547652b82fep+0, -0x1.62e42fefa39efp-1, -0x1.abc9e3b39803fp-56, 0x1.8p52,
                                    0x1.62e42fefa39efp-1,  0x1.abc9e3b39803fp-56, 0x1.0000080000000p52};
static __constant__ float cExpQf[5] = {5.000000000e-01f, 1.666657776e-01f, 4.166655615e-02f, 8.363173343e-03f,
                                       1.392617589e-03f};

template <typename Real> struct Math;

// ------------------------------------------------------------------------------------ FP64
template <> struct Math<double> {
  using real = double;
  // 2^52 + 2^51 as a LITERAL: its low word is zero, so it encodes as a 32-bit immediate operand.
  // Measured on B200 (pad experiments): a DFMA with two uniform-register operands
  // holds the FP64 pipe ~3.7 cycles, with one ~2.3, with three vector registers ~2.5.
  static constexpr double kMagic = 6755399441055744.0;

  // Operand kinds matter on the FP64 pipe (measured on B200, tools/micro/dfma_forms.cu): a DFMA whose three
  // sources are three DIFFERENT vector registers holds the pipe for 3 cycles, every other FP64 instruction
  // (a constant-bank / uniform-register / immediate operand, or a register used twice; DADD; DMUL) for 2.
  // The Horner steps below are (register, register, constant); the closing steps are arranged so that they
  // use a register twice or an immediate instead of three registers.

  // Q(r) of expm1(r) = r + r^2 Q(r), |r| <= ln2/2
  static __device__ __forceinline__ double expm1_q(double r) {
    double q = cExpQ[9];
    q = fma(q, r, cExpQ[8]);
    q = fma(q, r, cExpQ[7]);
    q = fma(q, r, cExpQ[6]);
    q = fma(q, r, cExpQ[5]);
    q = fma(q, r, cExpQ[4]);
    q = fma(q, r, cExpQ[3]);
    q = fma(q, r, cExpQ[2]);
    q = fma(q, r, cExpQ[1]);
    return fma(q, r, cExpQ[0]);
  }
  // expm1(r) = r + r (r Q): the closing DFMA reads r twice (2 pipe cycles; r^2 Q + r would read three registers)
  static __device__ __forceinline__ double expm1_reduced(double r) {
    const double q = expm1_q(r);
    return fma(r, r * q, r);
  }

  // y = n ln2 + r with n = rint(y log2 e); returns r, n through `n`
  static __device__ __forceinline__ double reduce(double y, int& n) {
    double t = fma(y, cK[0], kMagic);
    n = __double2loint(t);
    double nd = t - kMagic;
    double r = fma(nd, cK[1], y);
    return fma(nd, cK[2], r);
  }

  static __device__ __forceinline__ void fill_table(uint32_t, int) {}
  // m = 1 - exp(-x) for 0 <= x < 1.4e9 (NaN propagates).  2^n is clamped at 2^-1000 in the integer
  // domain, so the result saturates at exactly 1 without an FP64 compare.  n is the LOW word of the
  // magic-number sum, so it aliases once x log2(e) reaches 2^31 (x ~ 1.49e9); the integrator never gets
  // there: x = (dt / tau_i) / alpha, and alpha's lower saturation is raised per lane (alpha_floor_exp
  // below) so that max_i(dt / tau_i) / alpha stays below 7.6e8.
  static __device__ __forceinline__ double decay(double x, uint32_t) {
    // single-constant reduction: ln2's low word shifts r by n * 2.3e-17, i.e. the result by
    // <= 2^n * |n| * 2.3e-17 absolute -- below half an ulp of m for every n <= -1 (m >= 0.29)
    double t = fma(-x, cK[0], kMagic);
    int n = __double2loint(t);
    double r = fma(t - kMagic, cK[1], -x);
    double p = expm1_reduced(r);
    n = max(n, -1000);
    double s = __hiloint2double((1023 + n) << 20, 0);  // 2^n (exact)
    return fma(-s, p, 1.0 - s);                         // 1 - s(1 + p); 1 - s is exact for n <= 0
  }

  // exp(u), saturating at 2^40 above and at 2^nmin below (nmin >= -40: alpha is kept inside
  // [9e-13, 1.1e12], or a narrower range whose lower end keeps decay()'s argument in its domain);
  // NaN/inf -> NaN.
  static __device__ __forceinline__ double exp_(double u, uint32_t, int nmin = -40) {
    int n;
    double r = reduce(u, n);
    double v = fma(r, fma(r, expm1_q(r), 1.0), 1.0);  // 1 + r (1 + r Q): two immediate-operand DFMAs
    n = max(min(n, 40), nmin);
    return __hiloint2double(__double2hiint(v) + (n << 20), __double2loint(v));
  }
  // the lowest exponent alpha may take for a lane whose fastest pool has dt / tau = k0max: with
  // alpha >= 2^nmin / sqrt(2), x = k0max / alpha <= k0max 2^(-nmin) sqrt(2) < 2^28 2 sqrt(2) = 7.6e8.
  // `slack` lowers alpha's reach below the seed (Newton mode halves alpha at most once per iteration).
  static __device__ __forceinline__ int alpha_floor_exp(double k0max, int slack) {
    const int e = ((__double2hiint(k0max) >> 20) & 0x7ff) - 1023;  // ilogb for positive normal k0max
    return min(40, max(-40, e - 28 + slack));
  }

  // 1/a for positive normal a: MUFU.RCP64H seed y (relative error e ~ 2^-20) and one third-order
  // step y (1 + e + e^2), error e^3: 3 DFMA, ~1 ulp.
  static __device__ __forceinline__ double rcp(double a) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    double e = fma(-a, y, 1.0);
    return fma(y, fma(e, e, e), y);
  }

  // sqrt(a) for a >= 0: MUFU.RSQ64H seed y, t = a y, e = 1 - t y, sqrt = t (1 + e/2 + 3 e^2 / 8),
  // error 5 e^3 / 16: 5 FP64 ops.  sqrt(0) = 0 (integer clamp of the seed, no FP64 compare); a < 0 -> non-finite
  // (-inf with the seed clamp, NaN without: either way the member's next step is NaN).
  static __device__ __forceinline__ double sqrt_(double a) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
#if UFAIR_SQRT_SEED_CLAMP
    // a = 0: the seed is +inf and a y would be NaN; clamping the seed's high word to the largest finite
    // value (one integer min) gives t = 0 and sqrt(0) = 0 without a test on a and two selects (measured:
    // 29.8 -> 29.4 ms).  The NaN seed of a negative argument is clamped too, so sqrt(a < 0) = -inf.
    y = __hiloint2double(min(__double2hiint(y), 0x7fe00000), 0);
#endif
    double t = a * y;
    double e = fma(-t, y, 1.0);
    double p = fma(e, 0.375, 0.5) * e;
    double g = fma(t, p, t);
#if UFAIR_SQRT_SEED_CLAMP
    return g;
#else
    return ((__double2hiint(a) << 1) | __double2loint(a)) == 0 ? 0.0 : g;
#endif
  }

  // log(y) for a positive normal y: 2 atanh-series with a reciprocal instead of a division.  Branch-free on
  // purpose: a special-case branch per logarithm splits the time step into basic blocks and stops the
  // compiler from interleaving the independent gases of a lane (measured: 33.1 -> 29.6 ms on the dense
  // all-gases-per-lane kernel).  `not_normal(y)` says whether y is outside the fast path's domain; callers
  // test it once per step for all their logarithms and patch the rare cases with log_special().
  static __device__ __forceinline__ double log_fast(double y) {
    int hi = __double2hiint(y), lo = __double2loint(y);
    // y = 2^e m with m in [sqrt(1/2), sqrt(2)): bias the high word so the exponent field rolls over
    // exactly at the mantissa of sqrt(2) (0x6a09e...), the fdlibm normalisation -- 4 integer ops
    const int hx = hi + (0x3ff00000 - 0x3fe6a09e);
    const int e = (hx >> 20) - 1023;
    const int mh = (hx & 0x000fffff) + 0x3fe6a09e;
    double m = __hiloint2double(mh, lo);
    double f = m - 1.0;
    double d = m + 1.0, rc = rcp(d);
    double s = f * rc;  // f/d to ~1.5 ulp; 2 s is the leading term, so log is good to ~2.4 ulp (measured)
    double w = s * s;
    double L = cLogL[6];
    L = fma(L, w, cLogL[5]);
    L = fma(L, w, cLogL[4]);
    L = fma(L, w, cLogL[3]);
    L = fma(L, w, cLogL[2]);
    L = fma(L, w, cLogL[1]);
    L = fma(L, w, cLogL[0]);
    double ed = __hiloint2double(0x43300000, e ^ 0x80000000) - cK[6];  // (double)e
    // (s (2 + w L) would save a DMUL and a DFMA, but rounding 2 + w L costs the logarithm about one more ulp: it failed the 2.5-ulp test)
    double t = fma(s * w, L, ed * cK[5]);
    t = fma(2.0, s, t);
    return fma(ed, cK[4], t);
  }
  // not a positive normal number (never on a physical trajectory)
  static __device__ __forceinline__ bool not_normal(double y) {
    return (unsigned)(__double2hiint(y) - 0x00100000) >= 0x7fe00000u;
  }
  // ... and what the logarithm is there: -inf for +-0 and subnormals (flushed), NaN for negatives, y itself
  // for +inf / NaN
  static __device__ __forceinline__ double log_special(double y) {
    const int hi = __double2hiint(y);
    return (unsigned)(hi & 0x7fffffff) < 0x00100000u ? -INFINITY : (hi < 0 ? __longlong_as_double(0x7ff8000000000000ll) : y);
  }
  static __device__ __forceinline__ double log_(double y) {  // any y (probe, one-off uses)
    return __builtin_expect(not_normal(y), 0) ? log_special(y) : log_fast(y);
  }

  static __device__ __forceinline__ double fmax_(double a, double b) { return fmax(a, b); }
  static __device__ __forceinline__ int floor_to_int(double a) { return __double2int_rd(a); }  // saturating
  // histogram coordinate: subtract, then multiply, each rounded (no FMA contraction) so the
  // binning is reproducible on the host
  static __device__ __forceinline__ double bin_x(double T, double lo, double invw) {
    return __dmul_rn(__dsub_rn(T, lo), invw);
  }
  // v & mask on both words: selects v (mask = ~0) or +0.0 (mask = 0) on the integer ALU
  static __device__ __forceinline__ double mask(double v, unsigned m) {
    return __hiloint2double(__double2hiint(v) & (int)m, __double2loint(v) & (int)m);
  }
  static __device__ __forceinline__ double sinh_pair(double v, uint32_t tb, int nmin = -40) {  // sinh via exp and 1/exp
    double e = exp_(v, tb, nmin);
    return 0.5 * (e - rcp(e));
  }
};

// ------------------------------------------------------------------------------------ FP32
template <> struct Math<float> {
  using real = float;
  static constexpr float kLog2e = 1.4426950408889634f;
  static constexpr float kLn2Hi = 0.693145751953125f;  // 12 trailing zero bits
  static constexpr float kLn2Lo = 1.42860682030941723e-6f;
  static constexpr float kMagic = 12582912.0f;  // 1.5 * 2^23

  static __device__ __forceinline__ float expm1_reduced(float r) {
    float q = cExpQf[4];
    q = fmaf(q, r, cExpQf[3]);
    q = fmaf(q, r, cExpQf[2]);
    q = fmaf(q, r, cExpQf[1]);
    q = fmaf(q, r, cExpQf[0]);
    return fmaf(r * r, q, r);
  }
  static __device__ __forceinline__ float reduce(float y, int& n) {
    float t = fmaf(y, kLog2e, kMagic);
    n = __float_as_int(t) - 0x4b400000;
    float nd = t - kMagic;
    float r = fmaf(nd, -kLn2Hi, y);
    return fmaf(nd, -kLn2Lo, r);
  }
  static __device__ __forceinline__ void fill_table(uint32_t, int) {}
#if UFAIR_F32_MUFU
  static __device__ __forceinline__ float ex2(float t) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t));
    return e;
  }
  // m = 1 - exp(-x), x >= 0.  x < 1/4: x (1 - x/2 + x^2/6 - ... + x^5/720) (truncation < 5e-8, no
  // cancellation); otherwise 1 - MUFU.EX2(-x log2 e) (2 ulp of e <= 0.78: <= 5e-7 of m).  Both are
  // evaluated and one is selected (10 instructions instead of 16 for the reduction + polynomial form);
  // x large -> ex2 underflows to 0 and m = 1; NaN propagates through both arms.
  static __device__ __forceinline__ float decay(float x, uint32_t = 0) {
    float q = -1.0f / 720.0f;
    q = fmaf(q, x, 1.0f / 120.0f);
    q = fmaf(q, x, -1.0f / 24.0f);
    q = fmaf(q, x, 1.0f / 6.0f);
    q = fmaf(q, x, -0.5f);
    q = fmaf(q, x, 1.0f);
    const float big = 1.0f - ex2(x * -kLog2e);
    return x < 0.25f ? x * q : big;
  }
  // exp(u) = 2^n ex2(f), n = rint(u log2 e) clamped to +-40, f = u log2 e - n in two FMAs (so the
  // argument rounding does not grow with |u|): 7 instructions instead of 16
  static __device__ __forceinline__ float exp_(float u, uint32_t = 0, int = -40) {
    const float t = fmaf(u, kLog2e, kMagic);
    int n = __float_as_int(t) - 0x4b400000;
    const float nd = t - kMagic;
    float f = fmaf(u, kLog2e, -nd);
    f = fmaf(u, 1.925963033500011e-8f, f);  // log2(e) - float(log2(e))
    n = max(min(n, 40), -40);
    return __int_as_float(__float_as_int(ex2(f)) + (n << 23));
  }
#else
  static __device__ __forceinline__ float decay(float x, uint32_t = 0) {
    float y = fmaxf(-x, -30.0f);  // FMNMX is a single ALU op in FP32
    int n;
    float r = reduce(y, n);
    float p = expm1_reduced(r);
    float s = __int_as_float((127 + n) << 23);
    return fmaf(-s, p, 1.0f - s);
  }
  static __device__ __forceinline__ float exp_(float u, uint32_t = 0, int = -40) {
    int n;
    float r = reduce(u, n);
    float v = 1.0f + expm1_reduced(r);
    n = max(min(n, 40), -40);
    return __int_as_float(__float_as_int(v) + (n << 23));
  }
#endif
  static __device__ __forceinline__ float rcp(float a) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(a));
    return y;
  }
  static __device__ __forceinline__ float sqrt_(float a) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(a));
    return y;
  }
  static __device__ __forceinline__ float log_(float y) {
    float l;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(y));
    return l * 0.6931471805599453f;
  }
  // MUFU.LG2 handles every argument itself: the FP32 kernels never need the patch-up path
  static __device__ __forceinline__ float log_fast(float y) { return log_(y); }
  static __device__ __forceinline__ bool not_normal(float) { return false; }
  static __device__ __forceinline__ float log_special(float y) { return log_(y); }
  static __device__ __forceinline__ float fmax_(float a, float b) { return fmaxf(a, b); }
  static __device__ __forceinline__ int floor_to_int(float a) { return __float2int_rd(a); }
  static __device__ __forceinline__ float bin_x(float T, float lo, float invw) {
    return __fmul_rn(__fsub_rn(T, lo), invw);
  }
  static __device__ __forceinline__ float mask(float v, unsigned m) { return __int_as_float(__float_as_int(v) & (int)m); }
  static __device__ __forceinline__ int alpha_floor_exp(double, int) { return -40; }  // FP32 decay() saturates by itself
  static __device__ __forceinline__ float sinh_pair(float v, uint32_t = 0, int = -40) {
    float e = exp_(v);
    return 0.5f * (e - rcp(e));
  }
};

}  // namespace ufair
