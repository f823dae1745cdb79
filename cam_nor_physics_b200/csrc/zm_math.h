// zm_math.h -- portable IEEE-754 binary64 transcendentals for the ZM convection path.
//
// Why this exists: the Zhang-McFarlane hot path (reference physics/zm_conv.F90) is
// dominated by log / log10 / 10**x / exp calls inside Brent root finds
// (zm_conv.F90:5280-5300, 5440-5457, 5421-5437 -> external qsat_water).  Integer outputs
// (ideep, jt, maxg, lcl, lel) hang off float comparisons of those results, so a libm
// that differs by one ulp between host and device can flip an index.  Every function
// below is written with nothing but +,-,*,/ and explicit fma() (each correctly rounded on
// x86-64 and on sm_100a) plus integer bit moves, so the SAME source compiled by gcc
// (-ffp-contract=off) and by nvcc (-fmad=false) returns bit-identical doubles.
// The CUDA kernels always use these; the CPU oracle can be built either with glibc libm
// (closest to what gfortran would link) or with this header (bit-exact checker).
//
// Accuracy (measured against mpmath in tests/test_zm_math.py): < 1 ulp for every function.
// Domain handling is IEEE-like for the cases the physics can reach (x<=0, inf, nan,
// subnormal inputs, overflow/underflow of the result).
#pragma once
#include <stdint.h>
#if !defined(__CUDACC__)
#include <math.h>
#include <string.h>
#define ZM_HD static inline
#else
#define ZM_HD __host__ __device__ __forceinline__
#endif

namespace zmm {

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
#define ZMM_INVLN2     1.4426950408889634
#define ZMM_INVLN10_HI 0.4342944819032518
#define ZMM_INVLN10_LO 1.098319650216765e-17
#define ZMM_LOG2_10_HI 3.321928094887362
#define ZMM_LOG2_10_LO 1.661617516973592e-16

// v * 2^k, k any int; exact unless the result is subnormal/overflows.
ZM_HD double ldexp_(double v, int k) {
  if (k >= -1021 && k <= 1023) return v * u2d((uint64_t)(k + 1023) << 52);
  if (k > 1023) {
    v *= u2d((uint64_t)2046 << 52); k -= 1023;           // * 2^1023
    if (k > 1023) { v *= u2d((uint64_t)2046 << 52); k -= 1023; if (k > 1023) k = 1023; }
    return v * u2d((uint64_t)(k + 1023) << 52);
  }
  v *= u2d((uint64_t)(-969 + 1023) << 52); k += 969;      // * 2^-969 (keeps 53 bits alive)
  if (k < -1021) { v *= u2d((uint64_t)(-969 + 1023) << 52); k += 969; if (k < -1021) k = -1021; }
  return v * u2d((uint64_t)(k + 1023) << 52);
}

// Split positive finite normal/subnormal x into x = m * 2^e with m in [sqrt(1/2), sqrt(2)).
ZM_HD void frexp_sqrt2(double x, double& m, int& e) {
  uint64_t ix = d2u(x);
  int bias = 1023;
  if (ix < 0x0010000000000000ULL) {            // subnormal: scale by 2^54
    x *= 18014398509481984.0; ix = d2u(x); bias += 54;
  }
  e = (int)(ix >> 52) - bias;
  uint64_t mant = ix & 0x000fffffffffffffULL;
  if (mant > 0x0006a09e667f3bcdULL) { e += 1; m = u2d(mant | 0x3fe0000000000000ULL); }
  else                              {         m = u2d(mant | 0x3ff0000000000000ULL); }
}

// R(z), z = s*s, with log(1+f) = 2s + s*R(z), s = f/(2+f), |s| <= 0.1716.
// Coefficients: the classic 7-term minimax fit of 2*(z/3 + z^2/5 + z^3/7 + ...).
ZM_HD double log_R(double z) {
  const double Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01,
               Lg3 = 2.857142874366239149e-01, Lg4 = 2.222219843214978396e-01,
               Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
               Lg7 = 1.479819860511658591e-01;
  double w  = z * z;
  double t1 = w * fma(w, fma(w, Lg6, Lg4), Lg2);
  double t2 = z * fma(w, fma(w, fma(w, Lg7, Lg5), Lg3), Lg1);
  return t2 + t1;
}

// Non-finite / non-positive argument handling shared by the log family.
// returns true if *out holds the final answer.
ZM_HD bool log_special(double x, double* out) {
  uint64_t ix = d2u(x);
  if (ix - 1ULL < 0x7ff0000000000000ULL - 1ULL) return false;   // positive finite non-zero
  if ((ix << 1) == 0) { *out = -1.0 / 0.0; return true; }        // +-0 -> -inf
  if (ix == 0x7ff0000000000000ULL) { *out = x; return true; }    // +inf
  if ((ix >> 63) && ((ix << 1) <= 0xffe0000000000000ULL)) { *out = 0.0 / 0.0; return true; } // <0
  *out = x + x; return true;                                     // nan
}

// natural log, < 1 ulp.
ZM_HD double log_(double x) {
  double sp; if (log_special(x, &sp)) return sp;
  double m; int e; frexp_sqrt2(x, m, e);
  double f = m - 1.0;
  double s = f / (2.0 + f);
  double R = log_R(s * s);
  double hfsq = 0.5 * f * f;
  double dk = (double)e;
  return dk * ZMM_LN2_HI32 - ((hfsq - (s * (hfsq + R) + dk * ZMM_LN2_LO32)) - f);
}

// natural log as an unevaluated sum hi+lo good to ~2^-68 relative.
ZM_HD void log_dd(double x, double& hi, double& lo) {
  double m; int e; frexp_sqrt2(x, m, e);
  double f  = m - 1.0;                         // exact
  double dh = 2.0 + f;                         // rounded
  double dl = f - (dh - 2.0);                  // exact tail of 2+f
  double sh = f / dh;
  double r  = fma(-sh, dh, f);                 // exact remainder f - sh*dh
  r = fma(-sh, dl, r);
  double sl = r * (0.5 * (1.0 - sh));          // 1/(2+f) = (1-s)/2
  double R  = log_R(sh * sh);
  double th = 2.0 * sh;
  double tl = fma(sh, R, 2.0 * sl);
  double dk = (double)e;
  double a  = dk * ZMM_LN2_HI32;               // exact
  double A  = a + th;                          // |a| >= |th| or a == 0: Fast2Sum is exact
  double ae = (a == 0.0) ? 0.0 : (th - (A - a));
  double L  = ae + fma(dk, ZMM_LN2_LO32, tl);
  hi = A + L;
  lo = L - (hi - A);
}

// log10, < 1 ulp.
ZM_HD double log10_(double x) {
  double sp; if (log_special(x, &sp)) return sp;
  double hi, lo; log_dd(x, hi, lo);
  double p  = hi * ZMM_INVLN10_HI;
  double pe = fma(hi, ZMM_INVLN10_HI, -p);
  return p + (pe + fma(hi, ZMM_INVLN10_LO, lo * ZMM_INVLN10_HI));
}

// exp(rh + rl) for |rh| <= ~0.36, |rl| << |rh|.  Degree-13 Taylor, Horner with fma.
ZM_HD double exp_reduced(double rh, double rl) {
  double p = 1.6059043836821613e-10;                 // 1/13!
  p = fma(p, rh, 2.08767569878681e-09);              // 1/12!
  p = fma(p, rh, 2.505210838544172e-08);             // 1/11!
  p = fma(p, rh, 2.755731922398589e-07);             // 1/10!
  p = fma(p, rh, 2.7557319223985893e-06);            // 1/9!
  p = fma(p, rh, 2.48015873015873e-05);              // 1/8!
  p = fma(p, rh, 0.0001984126984126984);             // 1/7!
  p = fma(p, rh, 0.001388888888888889);              // 1/6!
  p = fma(p, rh, 0.008333333333333333);              // 1/5!
  p = fma(p, rh, 0.041666666666666664);              // 1/4!
  p = fma(p, rh, 0.16666666666666666);               // 1/3!
  p = fma(p, rh, 0.5);
  // exp(rh) = 1 + rh + rh^2 * p ; fold the tail rl in to first order.
  double r2 = rh * rh;
  double t  = fma(r2, p, rl);                        // rh^2*p + rl   (rl*exp ~ rl)
  t = fma(rh, rl, t);                                // second-order cross term
  double s = 1.0 + rh;                               // Fast2Sum: 1 >= |rh|
  double e = rh - (s - 1.0);
  return s + (e + t);
}

// exp, < 1 ulp.
ZM_HD double exp_(double x) {
  if (!(x == x)) return x + x;
  if (x > 709.782712893384) return 1.0 / 0.0;
  if (x < -745.1332191019412) return 0.0;
  double kd = rint(x * ZMM_INVLN2);
  double hi = fma(-kd, ZMM_LN2_HI32, x);
  double lo = kd * ZMM_LN2_LO32;
  double rh = hi - lo;
  double rl = (hi - rh) - lo;
  return ldexp_(exp_reduced(rh, rl), (int)kd);
}

// 2^(yh+yl) given yh+yl as an unevaluated sum (|yl| << |yh|), < 1 ulp overall.
ZM_HD double exp2_dd(double yh, double yl) {
  if (yh > 1024.0) return 1.0 / 0.0;
  if (yh < -1080.0) return 0.0;
  double kd = rint(yh);
  double f0 = yh - kd;                               // exact
  double fh = f0 + yl;
  double fl = yl - (fh - f0);                        // |f0| >= |yl| normally; tail only matters to 2^-106
  double rh = fh * ZMM_LN2_HI;
  double rl = fma(fh, ZMM_LN2_HI, -rh) + fma(fh, ZMM_LN2_LO, fl * ZMM_LN2_HI);
  return ldexp_(exp_reduced(rh, rl), (int)kd);
}

// 10**x  (Fortran `10._r8**x`, i.e. pow(10.0, x)), < 1 ulp.
ZM_HD double pow10_(double x) {
  if (!(x == x)) return x + x;
  double yh = x * ZMM_LOG2_10_HI;
  double yl = fma(x, ZMM_LOG2_10_HI, -yh) + x * ZMM_LOG2_10_LO;
  return exp2_dd(yh, yl);
}

// x**y for x > 0 (general real power, Fortran `x**y` with real y), < 1 ulp.
// x == 0 -> 0 for y > 0; x < 0 -> nan (the physics never raises a negative base).
ZM_HD double pow_(double x, double y) {
  if (!(x == x) || !(y == y)) return x + y;
  if (y == 0.0) return 1.0;
  if (x == 0.0) return (y > 0.0) ? 0.0 : 1.0 / 0.0;
  if (x < 0.0) return 0.0 / 0.0;
  if (x == 1.0 / 0.0) return (y > 0.0) ? x : 0.0;
  double lh, ll; log_dd(x, lh, ll);
  // p = y * (lh + ll) as hi/lo, then convert to base 2.
  double ph = y * lh;
  double pl = fma(y, lh, -ph) + y * ll;
  double qh = ph * ZMM_INVLN2;
  double ql = fma(ph, ZMM_INVLN2, -qh) + fma(ph, 2.0355273740931033e-17, pl * ZMM_INVLN2);
  return exp2_dd(qh, ql);
}

}  // namespace zmm
