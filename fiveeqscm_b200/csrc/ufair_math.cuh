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

// 1: table-driven exponentials (32-entry 2^(-j/32) table per warp in shared memory, degree-4
// polynomial: 11 FP64 operations per exponential instead of 17).  Measured on B200:
// FP64 instructions per warp-step 147 -> 122, but every exponential gains a bank-conflicted LDS.128
// and ~10 integer / select instructions on its dependent chain, and the LSU wavefront pipe (63 % busy
// before) becomes the limiter: 38.2 -> 38.8 ms dense, 25.4 -> 29.2 ms on the specialised forms.
// Kept (parity-green, tests/test_gpu_math.py passes with it) as a measured alternative; default off.
#ifndef UFAIR_EXP_TABLE
#define UFAIR_EXP_TABLE 0
#endif
// 1: FP32 exponentials on MUFU.EX2 (0: range reduction + polynomial on the FMA pipe)
#ifndef UFAIR_F32_MUFU
#define UFAIR_F32_MUFU 1
#endif

namespace ufair {

constexpr unsigned kExpTableBytes = UFAIR_EXP_TABLE ? 512u : 0u;  // per warp, FP64 kernels only

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
// Table-driven exponentials (UFAIR_EXP_TABLE=1): x = (32 n + j) ln2/32 + r with
// |r| <= ln2/64, e^-x = 2^-n T_j e^-r... (see Math<double>::decay).  Pairs (T_j, U_j) =
// (2^(-j/32), 1 - 2^(-j/32)), both correctly rounded; copied to shared memory by each warp, because
// a lane-dependent index into __constant__ memory would serialise.
static __constant__ double cExpT[64] = {
    0x1.0000000000000p+0, 0x0.0p+0,
    0x1.f50765b6e4540p-1, 0x1.5f134923757f3p-6,
    0x1.ea4afa2a490dap-1, 0x1.5b505d5b6f268p-5,
    0x1.dfc97337b9b5fp-1, 0x1.01b466423250ap-4,
    0x1.d5818dcfba487p-1, 0x1.53f391822dbc7p-4,
    0x1.cb720dcef9069p-1, 0x1.a46f918837cb7p-4,
    0x1.c199bdd85529cp-1, 0x1.f332113d56b1fp-4,
    0x1.b7f76f2fb5e47p-1, 0x1.20224341286e4p-3,
    0x1.ae89f995ad3adp-1, 0x1.45d819a94b14bp-3,
    0x1.a5503b23e255dp-1, 0x1.6abf137076a8ep-3,
    0x1.9c49182a3f090p-1, 0x1.8edb9f5703dc0p-3,
    0x1.93737b0cdc5e5p-1, 0x1.b23213cc8e86cp-3,
    0x1.8ace5422aa0dbp-1, 0x1.d4c6af7557c93p-3,
    0x1.82589994cce13p-1, 0x1.f69d99accc7b6p-3,
    0x1.7a11473eb0187p-1, 0x1.0bdd71829fcf2p-2,
    0x1.71f75e8ec5f74p-1, 0x1.1c1142e274118p-2,
    0x1.6a09e667f3bcdp-1, 0x1.2bec333018867p-2,
    0x1.6247eb03a5585p-1, 0x1.3b7029f8b54f7p-2,
    0x1.5ab07dd485429p-1, 0x1.4a9f0456f57adp-2,
    0x1.5342b569d4f82p-1, 0x1.597a952c560fcp-2,
    0x1.4bfdad5362a27p-1, 0x1.6804a5593abb2p-2,
    0x1.44e086061892dp-1, 0x1.763ef3f3ceda6p-2,
    0x1.3dea64c123422p-1, 0x1.842b367db97bcp-2,
    0x1.371a7373aa9cbp-1, 0x1.91cb1918aac6bp-2,
    0x1.306fe0a31b715p-1, 0x1.9f203eb9c91d6p-2,
    0x1.29e9df51fdee1p-1, 0x1.ac2c415c0423ep-2,
    0x1.2387a6e756238p-1, 0x1.b8f0b23153b8fp-2,
    0x1.1d4873168b9aap-1, 0x1.c56f19d2e8cabp-2,
    0x1.172b83c7d517bp-1, 0x1.d1a8f87055d0ap-2,
    0x1.11301d0125b51p-1, 0x1.dd9fc5fdb495fp-2,
    0x1.0b5586cf9890fp-1, 0x1.e954f260cede1p-2,
    0x1.059b0d3158574p-1, 0x1.f4c9e59d4f518p-2};
// expm1(r) = r + r^2 (1/2 + r Q4(r)), |r| <= ln2/64: max rel err 2.0e-17 (tools/gen_poly.py)
static __constant__ double cExpQ4[4] = {0x1.555555554dd45p-3, 0x1.555555555194dp-5, 0x1.11114f8a7941cp-7,
                                        0x1.6c16ffe57d9c9p-10};
// 32 log2(e), ln2/32 high and low parts
static __constant__ double cK32[3] = {0x1.71547652b82fep+5, 0x1.62e42fefa39efp-6, 0x1.abc9e3b39803fp-61};
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

#if UFAIR_EXP_TABLE
  // expm1(r) for |r| <= ln2/64: r + r^2 (1/2 + r Q4(r)); the 1/2 is an immediate operand
  static __device__ __forceinline__ double expm1_small(double r) {
    double q = cExpQ4[3];
    q = fma(q, r, cExpQ4[2]);
    q = fma(q, r, cExpQ4[1]);
    q = fma(q, r, cExpQ4[0]);
    q = fma(q, r, 0.5);
    return fma(r * r, q, r);
  }
  static __device__ __forceinline__ void lds_pair(uint32_t a, double& t, double& u) {
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(t), "=d"(u) : "r"(a));
  }
  static __device__ __forceinline__ double lds_one(uint32_t a) {
    double t;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(t) : "r"(a));
    return t;
  }
  // each warp's copy of cExpT in shared memory (tb = its 32-bit shared address); lane j fills entry j
  static __device__ __forceinline__ void fill_table(uint32_t tb, int lane) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(tb + 16u * (uint32_t)lane), "d"(cExpT[2 * lane]),
                 "d"(cExpT[2 * lane + 1])
                 : "memory");
  }

  // m = 1 - exp(-x), x >= 0 (NaN, +inf -> NaN).  k = rint(32 x / ln2) = 32 n + j, r = k ln2/32 - x:
  //   exp(-x) = 2^-n T_j e^r,   m = (1 - 2^-n T_j) - 2^-n T_j expm1(r).
  // For n = 0 the first term is the tabulated U_j = 1 - T_j (correctly rounded: no cancellation error
  // from T_j's own rounding, so tiny x keep ~1 ulp); for n >= 1 it is a plain subtraction (m >= 1/2).
  // 2^-n is applied in the integer domain and clamped at 2^-1000; x >= 4.6e7, where k no longer fits
  // the low word of the magic-number sum, saturates the same way (m = 1).  11 FP64 operations.
  static __device__ __forceinline__ double decay(double x, uint32_t tb) {
    const double t = fma(x, cK32[0], kMagic);
    int k = __double2loint(t);
    double r = fma(t - kMagic, cK32[1], -x);
    const bool big = (unsigned)(__double2hiint(t) - 0x43380001) < 0x3cb7ffffu;  // finite, k >= 2^32
    k = big ? 32000 : k;
    r = __hiloint2double(big ? 0 : __double2hiint(r), big ? 0 : __double2loint(r));
    const int n = min(k >> 5, 1000);
    double T, U;
    lds_pair(tb + 16u * (uint32_t)(k & 31), T, U);
    const double sT = __hiloint2double(__double2hiint(T) - (n << 20), __double2loint(T));
    const double p = expm1_small(r);
    const double d = 1.0 - sT;
    const double X = __hiloint2double(n == 0 ? __double2hiint(U) : __double2hiint(d),
                                      n == 0 ? __double2loint(U) : __double2loint(d));
    return fma(-sT, p, X);
  }

  // exp(u), saturating at 2^+-40 (alpha is kept inside [9e-13, 1.1e12]); |u| < 4.6e7, NaN/inf -> NaN.
  // k = rint(-32 u / ln2) = 32 n + j (floor division), r = u + k ln2/32: exp(u) = 2^-n T_j (1 + expm1(r)).
  static __device__ __forceinline__ double exp_(double u, uint32_t tb) {
    const double t = fma(-u, cK32[0], kMagic);
    const int k = __double2loint(t);
    const double kd = t - kMagic;
    double r = fma(kd, cK32[1], u);
    r = fma(kd, cK32[2], r);
    const int n = max(min(k >> 5, 40), -40);
    const double T = lds_one(tb + 16u * (uint32_t)(k & 31));
    const double sT = __hiloint2double(__double2hiint(T) - (n << 20), __double2loint(T));
    return fma(sT, expm1_small(r), sT);
  }
#else
  static __device__ __forceinline__ void fill_table(uint32_t, int) {}
  // m = 1 - exp(-x) for 0 <= x <= 1e15 (larger x, inf: NaN; NaN propagates).  2^n is clamped at
  // 2^-1000 in the integer domain, so the result saturates at exactly 1 without an FP64 compare.
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

  // exp(u), saturating at 2^+-40 (alpha is kept inside [9e-13, 1.1e12]); NaN/inf -> NaN.
  static __device__ __forceinline__ double exp_(double u, uint32_t) {
    int n;
    double r = reduce(u, n);
    double v = 1.0 + expm1_reduced(r);
    n = max(min(n, 40), -40);
    return __hiloint2double(__double2hiint(v) + (n << 20), __double2loint(v));
  }
#endif

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
#ifndef UFAIR_EXP_LOG_NOSPECIAL  // experiment: what the special-case branch costs (basic-block splitting)
    if (__builtin_expect((unsigned)(hi - 0x00100000) >= 0x7fe00000u, 0))
#else
    if (false)
#endif
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
  static __device__ __forceinline__ double sinh_pair(double v, uint32_t tb) {  // sinh via exp and 1/exp
    double e = exp_(v, tb);
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
  static __device__ __forceinline__ float exp_(float u, uint32_t = 0) {
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
  static __device__ __forceinline__ float exp_(float u, uint32_t = 0) {
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
  static __device__ __forceinline__ float fmax_(float a, float b) { return fmaxf(a, b); }
  static __device__ __forceinline__ int floor_to_int(float a) { return __float2int_rd(a); }
  static __device__ __forceinline__ float bin_x(float T, float lo, float invw) {
    return __fmul_rn(__fsub_rn(T, lo), invw);
  }
  static __device__ __forceinline__ float mask(float v, unsigned m) { return __int_as_float(__float_as_int(v) & (int)m); }
  static __device__ __forceinline__ float sinh_pair(float v, uint32_t = 0) {
    float e = exp_(v);
    return 0.5f * (e - rcp(e));
  }
};

}  // namespace ufair
