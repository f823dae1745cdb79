// zm_math.h -- portable IEEE-754 binary64 transcendentals for the ZM convection path.
//
// Why this exists: the Zhang-McFarlane hot path (reference physics/zm_conv.F90) is
// dominated by log / log10 / 10**x / exp calls inside Brent root finds
// (zm_conv.F90:5280-5300, 5440-5457, 5421-5437 -> external qsat_water).  Integer outputs
// (ideep, jt, maxg, lcl, lel) hang off float comparisons of those results, so a libm
// that differs by one ulp between host and device can flip an index.  Every function
// below is written with nothing but +,-,*,fma() (each correctly rounded on x86-64 and on
// sm_100a), integer bit moves and two 128-entry tables, so the SAME source compiled by gcc
// (-ffp-contract=off) and by nvcc (-fmad=false) returns bit-identical doubles.
// The CUDA kernels always use these; the CPU oracle can be built either with glibc libm
// (closest to what gfortran would link) or with this header (bit-exact checker).
//
// Design for the GPU: no divisions (a double division costs ~140 cycles of latency on B200),
// no data-dependent branches (special cases are resolved by selects at the end so the four
// independent transcendentals of one Goff-Gratch evaluation can be interleaved by the
// scheduler), table-driven range reduction (log: z*invc-1 via one fma, exp2: 2^(j/128)).
//
// Accuracy (tests/test_capi_and_host.py, against mpmath): < 0.6 ulp for every function.
#pragma once
#include <stdint.h>
#include "zm_math_tables.h"
#include "zm_svp_table.h"
#if !defined(__CUDACC__)
#include <math.h>
#include <string.h>
#define ZM_HD static inline
#else
#define ZM_HD __host__ __device__ __forceinline__
#endif

namespace zmm {

#if defined(__CUDACC__)
__device__ const double d_log_tab[384] = ZMM_LOG_TAB_VALUES;
__device__ const double d_exp2_tab[256] = ZMM_EXP2_TAB_VALUES;
__device__ __align__(16) const double d_svp_tab[ZMM_SVP_NCELL * ZMM_SVP_NCOEF] = ZMM_SVP_TAB_VALUES;
#endif
static const double h_log_tab[384] = ZMM_LOG_TAB_VALUES;
static const double h_exp2_tab[256] = ZMM_EXP2_TAB_VALUES;
static const double h_svp_tab[ZMM_SVP_NCELL * ZMM_SVP_NCOEF] = ZMM_SVP_TAB_VALUES;

// static shared memory of a kernel that calls hot_tables_load() and hot_svp_load()
#define ZMM_HOT_SMEM_BYTES (128 * 16 + 128 * 8 + 128 * 16 + ZMM_SVP_NCELL * ZMM_SVP_NCOEF * 8)
#if defined(__CUDACC__)
// Shared-memory copies used by the *_hot variants (one LDS.128 per entry instead of __ldg loads that each
// need a descriptor in uniform registers).  A kernel that calls a *_hot function must call
// hot_tables_load() with all threads of the block before its first use.
__shared__ double2 s_log_a[128];      // (invc, logc_hi)
__shared__ double  s_log_b[128];      // logc_lo
__shared__ double2 s_exp2[128];       // (2^(j/128) hi, lo)
// saturation vapour pressure cells, 80 bytes each: consecutive cells start 20 banks apart, so the eight lanes of an
// LDS.128 phase that read eight different cells spread over all 32 banks
__shared__ double2 s_svp[ZMM_SVP_NCELL * ZMM_SVP_NCOEF / 2];
__device__ __forceinline__ void hot_tables_load() {
  for (int i = threadIdx.x; i < 128; i += blockDim.x) {
    s_log_a[i] = make_double2(d_log_tab[3 * i], d_log_tab[3 * i + 1]);
    s_log_b[i] = d_log_tab[3 * i + 2];
    s_exp2[i] = make_double2(d_exp2_tab[2 * i], d_exp2_tab[2 * i + 1]);
  }
  __syncthreads();
}
// kernels that evaluate svp_water<true> (the state function of the CAPE passes) call this as well
__device__ __forceinline__ void hot_svp_load() {
  for (int i = threadIdx.x; i < ZMM_SVP_NCELL * ZMM_SVP_NCOEF / 2; i += blockDim.x)
    s_svp[i] = make_double2(d_svp_tab[2 * i], d_svp_tab[2 * i + 1]);
  __syncthreads();
}
#endif

template <bool HOT> ZM_HD void log_entry(int i, double& invc, double& lch, double& lcl) {
#if defined(__CUDA_ARCH__)
  if (HOT) { const double2 a = s_log_a[i]; invc = a.x; lch = a.y; lcl = s_log_b[i]; return; }
  invc = __ldg(&d_log_tab[3 * i]); lch = __ldg(&d_log_tab[3 * i + 1]); lcl = __ldg(&d_log_tab[3 * i + 2]);
#else
  invc = h_log_tab[3 * i]; lch = h_log_tab[3 * i + 1]; lcl = h_log_tab[3 * i + 2];
#endif
}
template <bool HOT> ZM_HD void exp2_entry(int j, double& th, double& tl) {
#if defined(__CUDA_ARCH__)
  if (HOT) { const double2 a = s_exp2[j]; th = a.x; tl = a.y; return; }
  th = __ldg(&d_exp2_tab[2 * j]); tl = __ldg(&d_exp2_tab[2 * j + 1]);
#else
  th = h_exp2_tab[2 * j]; tl = h_exp2_tab[2 * j + 1];
#endif
}

ZM_HD uint64_t d2u(double x) {
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(x);
#else
  uint64_t u; memcpy(&u, &x, 8); return u;
#endif
}
ZM_HD double u2d(uint64_t u) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)u);
#else
  double x; memcpy(&x, &u, 8); return x;
#endif
}

// ---- constants (hi/lo pairs generated with mpmath at 200 bits) ------------------------
#define ZMM_LN2_HI32   0.6931471803691238      /* ln2 rounded to 32 bits: k*LN2_HI32 exact */
#define ZMM_LN2_LO32   1.9082149292705877e-10
#define ZMM_LN2_HI     0.6931471805599453
#define ZMM_LN2_LO     2.3190468138462996e-17
#define ZMM_INVLN2_HI  1.4426950408889634
#define ZMM_INVLN2_LO  2.0355273740931033e-17
#define ZMM_INVLN10_HI 0.4342944819032518
#define ZMM_INVLN10_LO 1.098319650216765e-17
#define ZMM_LOG2_10_HI 3.321928094887362
#define ZMM_LOG2_10_LO 1.661617516973592e-16
#define ZMM_INF        (u2d(0x7ff0000000000000ULL))
#define ZMM_NAN        (u2d(0x7ff8000000000000ULL))

// ---- log core: log(x) = hi + lo (unevaluated, ~2^-68 relative), x positive & finite -------------
// x = 2^k * z, z in [0.6855, 1.371); z*invc - 1 = r with |r| < 0.004 (one fma, exact rounding);
// log(x) = k*ln2 + logc + log1p(r), logc = -log(invc) tabulated as hi (multiple of 2^-32) + lo.
// HOT = true: x must be a positive NORMAL finite number (no subnormal scaling); same bits as HOT = false there.
template <bool HOT> ZM_HD void log_core_t(double x, double& hi, double& lo) {
  uint64_t ix = d2u(x);
  bool sub = false;
  if (!HOT) {
    // subnormals: scale by 2^54 (select, no branch)
    sub = ix < 0x0010000000000000ULL;
    const uint64_t ixs = d2u(x * 18014398509481984.0);
    ix = sub ? ixs : ix;
  }
  const uint64_t tmp = ix - 0x3fe5f00000000000ULL;
  const int i = (int)((tmp >> 45) & 127);
  int k = (int)((int64_t)tmp >> 52);
  if (!HOT) k = sub ? k - 54 : k;
  const double z = u2d(ix - (tmp & 0xfff0000000000000ULL));
  double invc, lch, lcl;
  log_entry<HOT>(i, invc, lch, lcl);
  const double r = fma(z, invc, -1.0);
  const double kd = (double)k;
  const double w = fma(kd, ZMM_LN2_HI32, lch);            // exact: both are multiples of 2^-32
  // s1 + e1 = w + r exactly (TwoSum)
  const double s1 = w + r;
  const double bb = s1 - w;
  const double e1 = (w - (s1 - bb)) + (r - bb);
  // r^2 exactly as r2h + r2l; the quadratic term -r^2/2
  const double r2h = r * r;
  const double r2l = fma(r, r, -r2h);
  const double q = -0.5 * r2h;
  // s2 + e2 = s1 + q (|s1| >= |q| whenever s1 != 0 matters; Fast2Sum)
  const double s2 = s1 + q;
  const double e2 = (s1 - s2) + q;
  // cubic and higher: r^3 * (1/3 - r/4 + r^2/5 - r^3/6 + r^4/7 - r^5/8)
  const double r4 = r2h * r2h;
  const double p01 = fma(r, -0.25, 0.3333333333333333);
  const double p23 = fma(r, -0.16666666666666666, 0.2);
  const double p45 = fma(r, -0.125, 0.14285714285714285);
  const double p = fma(r4, p45, fma(r2h, p23, p01));
  const double r3 = r2h * r;
  double l = fma(kd, ZMM_LN2_LO32, lcl);
  l = l + (e1 + e2);
  l = fma(-0.5, r2l, l);
  l = fma(r3, p, l);
  hi = s2 + l;
  lo = l - (hi - s2);
}
ZM_HD void log_core(double x, double& hi, double& lo) { log_core_t<false>(x, hi, lo); }

// resolves x <= 0, inf, nan for the log family with selects
ZM_HD double log_fixup(double x, double res) {
  const bool ok = (x > 0.0) && (x < ZMM_INF);
  double sp = (x == 0.0) ? -ZMM_INF : ((x > 0.0) ? x : ZMM_NAN);   // +inf -> +inf, <0 or nan -> nan
  return ok ? res : sp;
}

// Natural log in one double, < 0.6 ulp: the same reduction as log_core (x = 2^k z, r = z*invc - 1, |r| < 0.004),
//   log x = (k ln2 + logc) + r - r^2/2 + r^3 (1/3 - r/4 + ... - r^5/8)
// with one compensated step (hi + lo = w + r exactly, |w| >= |r| or w = 0) instead of the double-double bookkeeping
// log10 / pow need: 19 floating-point operations against 28.  HOT = true: positive NORMAL finite x only.
template <bool HOT> ZM_HD double log1_t(double x) {
  uint64_t ix = d2u(x);
  bool sub = false;
  if (!HOT) {
    sub = ix < 0x0010000000000000ULL;
    const uint64_t ixs = d2u(x * 18014398509481984.0);
    ix = sub ? ixs : ix;
  }
  const uint64_t tmp = ix - 0x3fe5f00000000000ULL;
  const int i = (int)((tmp >> 45) & 127);
  int k = (int)((int64_t)tmp >> 52);
  if (!HOT) k = sub ? k - 54 : k;
  const double z = u2d(ix - (tmp & 0xfff0000000000000ULL));
  double invc, lch, lcl;
  log_entry<HOT>(i, invc, lch, lcl);
  const double r = fma(z, invc, -1.0);
  const double kd = (double)k;
  const double w = fma(kd, ZMM_LN2_HI32, lch);            // exact: both are multiples of 2^-32
  const double hi = w + r;
  const double lo = ((w - hi) + r) + fma(kd, ZMM_LN2_LO32, lcl);
  const double r2 = r * r;
  const double r4 = r2 * r2;
  const double r3 = r2 * r;
  const double p01 = fma(r, -0.25, 0.3333333333333333);
  const double p23 = fma(r, -0.16666666666666666, 0.2);
  const double p45 = fma(r, -0.125, 0.14285714285714285);
  const double p = fma(r4, p45, fma(r2, p23, p01));
  const double t = fma(-0.5, r2, lo);
  return fma(r3, p, t) + hi;
}
ZM_HD double log_(double x) { return log_fixup(x, log1_t<false>(x)); }

// log10, < 0.6 ulp
ZM_HD double log10_(double x) {
  double hi, lo; log_core(x, hi, lo);
  const double p  = hi * ZMM_INVLN10_HI;
  const double pe = fma(hi, ZMM_INVLN10_HI, -p);
  const double res = p + (pe + fma(hi, ZMM_INVLN10_LO, lo * ZMM_INVLN10_HI));
  return log_fixup(x, res);
}

// Hot variants: identical bits to log_/log10_ for positive normal finite x, no special-case selects.
ZM_HD double log_hot(double x) { return log1_t<true>(x); }
ZM_HD double log10_hot(double x) {
  double hi, lo; log_core_t<true>(x, hi, lo);
  const double p  = hi * ZMM_INVLN10_HI;
  const double pe = fma(hi, ZMM_INVLN10_HI, -p);
  return p + (pe + fma(hi, ZMM_INVLN10_LO, lo * ZMM_INVLN10_HI));
}

// 2^(yh+yl), yh+yl an unevaluated sum with |yl| << |yh|; < 0.6 ulp.  Branch-free:
// k = rint(128*yh); j = k mod 128; 2^y = 2^(k div 128) * T[j] * exp((yh - k/128 + yl) * ln2)
// HOT = true: |yh| < 1000 required (no clamp); same bits as HOT = false there.
template <bool HOT> ZM_HD double exp2_dd_t(double yh, double yl) {
  double yc = yh;
  if (!HOT) {
    yc = (yh < 1100.0) ? yh : 1100.0;                // clamp (nan stays nan: fixed up by callers)
    yc = (yc > -1100.0) ? yc : -1100.0;
  }
  const double kd = rint(yc * 128.0);
  const int k = (int)kd;
  const double f = fma(kd, -0.0078125, yc);          // exact
  const double r = (f + yl) * ZMM_LN2_HI;
  const int j = k & 127;
  const int e = (k - j) >> 7;
  double th, tl;
  exp2_entry<HOT>(j, th, tl);
  // expm1(r) = r + r^2*(1/2 + r/6 + r^2/24 + r^3/120)
  const double r2 = r * r;
  const double c01 = fma(r, 0.16666666666666666, 0.5);
  const double c23 = fma(r, 0.008333333333333333, 0.041666666666666664);
  const double pm1 = fma(r2, fma(r2, c23, c01), r);
  const double v = th + fma(th, pm1, tl);
  // scale by 2^e in two exact-or-single-rounding steps (e in [-1101, 1101])
  const int e1 = e >> 1, e2 = e - e1;
  return (v * u2d((uint64_t)(e1 + 1023) << 52)) * u2d((uint64_t)(e2 + 1023) << 52);
}
ZM_HD double exp2_dd(double yh, double yl) { return exp2_dd_t<false>(yh, yl); }

// exp, < 0.6 ulp
ZM_HD double exp_(double x) {
  const double yh = x * ZMM_INVLN2_HI;
  const double yl = fma(x, ZMM_INVLN2_HI, -yh) + x * ZMM_INVLN2_LO;
  const double res = exp2_dd(yh, yl);
  return (x == x) ? res : x;
}

// 10**x  (Fortran `10._r8**x`, i.e. pow(10.0, x)), < 0.6 ulp
ZM_HD double pow10_(double x) {
  const double yh = x * ZMM_LOG2_10_HI;
  const double yl = fma(x, ZMM_LOG2_10_HI, -yh) + x * ZMM_LOG2_10_LO;
  const double res = exp2_dd(yh, yl);
  return (x == x) ? res : x;
}

// 10**x for |x| < 300 (no clamp, no nan select): identical bits to pow10_ there
ZM_HD double pow10_hot(double x) {
  const double yh = x * ZMM_LOG2_10_HI;
  const double yl = fma(x, ZMM_LOG2_10_HI, -yh) + x * ZMM_LOG2_10_LO;
  return exp2_dd_t<true>(yh, yl);
}

// ---- Goff-Gratch saturation vapour pressure over water, Pa ----------------------------------------------------
// The formula as CAM's wv_sat_methods codes it (GoffGratch_svp_water; reached from zm_conv.F90:5423,5433 through
// wv_saturation::qsat_water).  log10(1013.246) is a compile-time constant in the reference build.
ZM_HD double svp_water_formula(double t) {
  const double tboil = 373.16;
  return pow10_(-7.90298 * (tboil / t - 1.0) + 5.02808 * log10_(tboil / t) -
                1.3816e-7 * (pow10_(11.344 * (1.0 - t / tboil)) - 1.0) +
                8.1328e-3 * (pow10_(-3.49149 * (tboil / t - 1.0)) - 1.0) + 3.0057148979490314) * 100.0;
}
#if defined(__CUDACC__)
__device__ __noinline__ double svp_water_formula_cold(double t) { return svp_water_formula(t); }
#endif
// The same function from the per-kelvin polynomials of zm_svp_table.h (scripts/gen_svp_table.py): for
// ZMM_SVP_T0 - 0.5 <= t < ZMM_SVP_T0 + ZMM_SVP_NCELL - 0.5 one table cell and 12 multiply-adds (the formula costs
// three 10**x, one log10 and two divisions, and in double it carries +-2 ulp from the rounding of its exponent
// alone, tens of ulp below 200 K); < 2 ulp against the formula in 300-bit arithmetic from 139.5 K up; outside the
// table the formula.  The cells below 140 K exist for the parcels of cold, dry columns, which the CAPE sweep lifts
// to 40 hPa along a dry adiabat (100-140 K; one lane there used to send its whole warp through the formula): the
// formula collapses super-exponentially down there (es(120 K) = 9e-17 Pa, es(100 K) = 3e-42 Pa), the degree-9 cells
// follow it to 3e-13 relative at 125 K, 1e-11 at 120 K, 4e-5 at 100 K -- an absolute error below 1e-23 Pa (one ulp at 139 K)
// everywhere, 1e-28 of any pressure the path sees: no term of the state function can resolve it.
// The cell is the integer nearest to t: t + 1.5*2^52 rounds t to that integer and leaves it in the low word of
// the sum (no float<->int conversion on the dependency chain); x = t - centre is exact, |x| <= 0.5.
// HOT = true reads the table from shared memory (hot_svp_load()).
template <bool HOT> ZM_HD double svp_water(double t) {
  if (!(t >= ZMM_SVP_T0 - 0.5 && t < ZMM_SVP_T0 + (ZMM_SVP_NCELL - 0.5))) {
#if defined(__CUDA_ARCH__)
    return svp_water_formula_cold(t);
#else
    return svp_water_formula(t);
#endif
  }
  const double y = t + 6755399441055744.0;                 // 1.5 * 2^52
  const int i = (int)(uint32_t)d2u(y) - (int)ZMM_SVP_T0;   // round(t) - T0
  const double x = t - (y - 6755399441055744.0);
  double c0, c1, c2, c3, c4, c5, c6, c7, c8, c9;
#if defined(__CUDA_ARCH__)
  if (HOT) {
    const double2* cc = &s_svp[i * (ZMM_SVP_NCOEF / 2)];
    const double2 a = cc[0], b = cc[1], c = cc[2], d = cc[3], e = cc[4];
    c0 = a.x; c1 = a.y; c2 = b.x; c3 = b.y; c4 = c.x; c5 = c.y; c6 = d.x; c7 = d.y; c8 = e.x; c9 = e.y;
  } else {
    const double2* cc = reinterpret_cast<const double2*>(d_svp_tab) + i * (ZMM_SVP_NCOEF / 2);
    const double2 a = __ldg(cc), b = __ldg(cc + 1), c = __ldg(cc + 2), d = __ldg(cc + 3), e = __ldg(cc + 4);
    c0 = a.x; c1 = a.y; c2 = b.x; c3 = b.y; c4 = c.x; c5 = c.y; c6 = d.x; c7 = d.y; c8 = e.x; c9 = e.y;
  }
#else
  const double* cc = h_svp_tab + i * ZMM_SVP_NCOEF;
  c0 = cc[0]; c1 = cc[1]; c2 = cc[2]; c3 = cc[3]; c4 = cc[4]; c5 = cc[5]; c6 = cc[6]; c7 = cc[7]; c8 = cc[8]; c9 = cc[9];
#endif
  // es = c0 + x*(c1 + c2 x + ... + c9 x^8): the bracket by Estrin's scheme (depth 4), c0 added last so that
  // the result carries one dominant rounding
  const double x2 = x * x;
  const double x4 = x2 * x2;
  const double x8 = x4 * x4;
  const double p12 = fma(c2, x, c1);
  const double p34 = fma(c4, x, c3);
  const double p56 = fma(c6, x, c5);
  const double p78 = fma(c8, x, c7);
  const double q0 = fma(p34, x2, p12);
  const double q1 = fma(p78, x2, p56);
  const double r0 = fma(q1, x4, q0);
  const double tl = fma(c9, x8, r0);
  return fma(x, tl, c0);
}

// a / b given rb = RN(1/b) (correctly rounded reciprocal of a fixed divisor): first quotient estimate plus
// one exact-residual correction (Markstein); equals the IEEE-754 quotient for normal-range operands.
ZM_HD double div_rcp(double a, double b, double rb) {
  const double q = a * rb;
  const double rem = fma(-b, q, a);
  return fma(rb, rem, q);
}

// x**y for x > 0 (general real power, Fortran `x**y` with real y), < 0.7 ulp.
// x == 0 -> 0 for y > 0; x < 0 -> nan (the physics never raises a negative base).
ZM_HD double pow_(double x, double y) {
  double lh, ll; log_core(x, lh, ll);
  const double ph = y * lh;
  const double pl = fma(y, lh, -ph) + y * ll;
  const double qh = ph * ZMM_INVLN2_HI;
  const double ql = fma(ph, ZMM_INVLN2_HI, -qh) + fma(ph, ZMM_INVLN2_LO, pl * ZMM_INVLN2_HI);
  double res = exp2_dd(qh, ql);
  const bool ok = (x > 0.0) && (x < ZMM_INF) && (y == y);
  double sp;
  if (!(x == x) || !(y == y)) sp = x + y;
  else if (y == 0.0) sp = 1.0;
  else if (x == 0.0) sp = (y > 0.0) ? 0.0 : ZMM_INF;
  else if (x < 0.0) sp = ZMM_NAN;
  else sp = (y > 0.0) ? x : 0.0;                      // x == +inf
  return ok ? res : sp;
}

// kept for callers that need an exact power-of-two scaling
ZM_HD double ldexp_(double v, int k) {
  const int k1 = k >> 1, k2 = k - k1;
  return (v * u2d((uint64_t)(k1 + 1023) << 52)) * u2d((uint64_t)(k2 + 1023) << 52);
}

}  // namespace zmm
