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

namespace ufair {

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

  // expm1(r) for |r| <= ln2/2:  r + r^2 Q(r)
  static __device__ __forceinline__ double expm1_reduced(double r) {
    double q = cExpQ[9];
    q = fma(q, r, cExpQ[8]);
    q = fma(q, r, cExpQ[7]);
    q = fma(q, r, cExpQ[6]);
    q = fma(q, r, cExpQ[5]);
    q = fma(q, r, cExpQ[4]);
    q = fma(q, r, cExpQ[3]);
    q = fma(q, r, cExpQ[2]);
    q = fma(q, r, cExpQ[1]);
    q = fma(q, r, cExpQ[0]);
    return fma(r * r, q, r);
  }

  // y = n ln2 + r with n = rint(y log2 e); returns r, n through `n`
  static __device__ __forceinline__ double reduce(double y, int& n) {
    double t = fma(y, cK[0], kMagic);
    n = __double2loint(t);
    double nd = t - kMagic;
    double r = fma(nd, cK[1], y);
    return fma(nd, cK[2], r);
  }

  // m = 1 - exp(-x) for 0 <= x <= 1e15 (larger x, inf: NaN; NaN propagates).  2^n is clamped at
  // 2^-1000 in the integer domain, so the result saturates at exactly 1 without an FP64 compare.
  static __device__ __forceinline__ double decay(double x) {
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

  // exp(u), saturating at 2^+-40 (alpha is kept inside [9e-13, 1.1e12]); NaN/inf -> NaN.
  static __device__ __forceinline__ double exp_(double u) {
    int n;
    double r = reduce(u, n);
    double v = 1.0 + expm1_reduced(r);
    n = max(min(n, 40), -40);
    return __hiloint2double(__double2hiint(v) + (n << 20), __double2loint(v));
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
  // error 5 e^3 / 16: 5 FP64 ops.  sqrt(0) = 0 (integer test, no FP64 compare); a < 0 -> NaN.
  static __device__ __forceinline__ double sqrt_(double a) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    double t = a * y;
    double e = fma(-t, y, 1.0);
    double p = fma(e, 0.375, 0.5) * e;
    double g = fma(t, p, t);
    return ((__double2hiint(a) << 1) | __double2loint(a)) == 0 ? 0.0 : g;
  }

  // log(y).  Positive normal y on the fast path; everything else takes the (never hot) libm call.
  static __device__ __forceinline__ double log_(double y) {
    int hi = __double2hiint(y), lo = __double2loint(y);
    // not a positive normal number (never on a physical trajectory): -inf for +-0 and subnormals
    // (flushed), NaN for negatives, y itself for +inf / NaN.  Kept tiny so it stays predicated.
    if (__builtin_expect((unsigned)(hi - 0x00100000) >= 0x7fe00000u, 0))
      return (unsigned)(hi & 0x7fffffff) < 0x00100000u ? -INFINITY : (hi < 0 ? __longlong_as_double(0x7ff8000000000000ll) : y);
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
    double t = fma(s * w, L, ed * cK[5]);
    t = fma(2.0, s, t);
    return fma(ed, cK[4], t);
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
  static __device__ __forceinline__ double sinh_pair(double v) {  // sinh via exp and 1/exp
    double e = exp_(v);
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
  static __device__ __forceinline__ float decay(float x) {
    float y = fmaxf(-x, -30.0f);  // FMNMX is a single ALU op in FP32
    int n;
    float r = reduce(y, n);
    float p = expm1_reduced(r);
    float s = __int_as_float((127 + n) << 23);
    return fmaf(-s, p, 1.0f - s);
  }
  static __device__ __forceinline__ float exp_(float u) {
    int n;
    float r = reduce(u, n);
    float v = 1.0f + expm1_reduced(r);
    n = max(min(n, 40), -40);
    return __int_as_float(__float_as_int(v) + (n << 23));
  }
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
  static __device__ __forceinline__ float fmax_(float a, float b) { return fmaxf(a, b); }
  static __device__ __forceinline__ int floor_to_int(float a) { return __float2int_rd(a); }
  static __device__ __forceinline__ float bin_x(float T, float lo, float invw) {
    return __fmul_rn(__fsub_rn(T, lo), invw);
  }
  static __device__ __forceinline__ float mask(float v, unsigned m) { return __int_as_float(__float_as_int(v) & (int)m); }
  static __device__ __forceinline__ float sinh_pair(float v) {
    float e = exp_(v);
    return 0.5f * (e - rcp(e));
  }
};

}  // namespace ufair
