// zm_externals.hpp -- TEST INFRASTRUCTURE (CPU oracle).  Not part of the product path.
//
// Arithmetic that the reference reaches through `use` statements whose modules are NOT in
// /root/reference (un-vendored ESCOMP/CAM + CIME share code, no pinned version):
//   physconst / shr_const_mod      zm_conv.F90:19-20
//   wv_saturation::qsat_water      zm_conv.F90:5423,5433   (every qsat_hPa)
//   wv_saturation::qsat            zm_conv.F90:1729,1804   (table version, zm_conv_evap)
//   cloud_fraction::cldfrc_fice    zm_conv.F90:18,1809
// They are restated here from the published CAM algorithms (Goff & Gratch 1946 saturation
// vapour pressure as coded in CAM's wv_sat_methods; CAM's estblf 1-K table 127.16..375.16 K
// with a 20 K water/ice transition; cldfrc_fice linear ramps).  The glibc-libm flavour evaluates the Goff-Gratch
// formula as written; the portable flavour takes it from the product's per-kelvin polynomial table of the same
// formula (zm_math.h svp_water, < 1.3 ulp) so that it stays the bit-exact checker of the kernels.  PARITY UNPINNED: nothing in
// /root/reference holds source, tests or golden values for these call sites, so this header
// *defines* them for the oracle; the CUDA library restates the same formulas independently
// (cam_nor_physics_b200/csrc/zm_device.cuh) and the Fortran stubs in fortran/ must mirror it.
#pragma once
#include <cmath>
#include <vector>

#ifdef ZMO_PORTABLE_MATH
#include "zm_math.h"
namespace zmo {
static inline double m_log(double x)   { return zmm::log_(x); }
static inline double m_log10(double x) { return zmm::log10_(x); }
static inline double m_exp(double x)   { return zmm::exp_(x); }
static inline double m_pow10(double x) { return zmm::pow10_(x); }
static inline double m_pow(double x, double y) { return zmm::pow_(x, y); }
// the saturation vapour pressure as the CUDA library evaluates it (per-kelvin polynomials of the Goff-Gratch
// formula, the formula outside 140..350 K): this flavour is the bit-exact checker of the kernels
static inline double m_svp_water(double t) { return zmm::svp_water<false>(t); }
static const char* const math_backend = "portable(zm_math.h)";
}
#else
namespace zmo {
static inline double m_log(double x)   { return std::log(x); }
static inline double m_log10(double x) { return std::log10(x); }
static inline double m_exp(double x)   { return std::exp(x); }
static inline double m_pow10(double x) { return std::pow(10.0, x); }   // Fortran 10._r8**x
static inline double m_pow(double x, double y) { return std::pow(x, y); }
static const char* const math_backend = "glibc-libm";
}
#define ZMO_SVP_FORMULA 1
#endif

namespace zmo {

// ---- physconst (shr_const_mod values) -------------------------------------------------
struct PhysConst {
  double cpair, epsilo, gravit, latice, latvap, tmelt, rair, cpwv, cpliq, rh2o, cpvir, zvir;
};
static inline PhysConst physconst_default() {
  PhysConst c;
  const double boltz = 1.38065e-23, avogad = 6.02214e26;
  const double rgas = avogad * boltz;
  const double mwdair = 28.966, mwwv = 18.016;
  c.cpair  = 1.00464e3;
  c.epsilo = mwwv / mwdair;
  c.gravit = 9.80616;
  c.latice = 3.337e5;
  c.latvap = 2.501e6;
  c.tmelt  = 273.15;
  c.rair   = rgas / mwdair;
  c.cpwv   = 1.810e3;
  c.cpliq  = 4.188e3;
  c.rh2o   = rgas / mwwv;
  c.cpvir  = c.cpwv / c.cpair - 1.0;
  c.zvir   = c.rh2o / c.rair - 1.0;
  return c;
}

// ---- wv_saturation --------------------------------------------------------------------
// Goff-Gratch saturation vapour pressure over water (Pa), t in K.
static inline double gg_svp_water(double t) {
#ifndef ZMO_SVP_FORMULA
  return m_svp_water(t);
#endif
  const double tboil = 373.16;
  return m_pow10(-7.90298 * (tboil / t - 1.0) +
                 5.02808 * m_log10(tboil / t) -
                 1.3816e-7 * (m_pow10(11.344 * (1.0 - t / tboil)) - 1.0) +
                 8.1328e-3 * (m_pow10(-3.49149 * (tboil / t - 1.0)) - 1.0) +
                 3.0057148979490314 /* log10(1013.246_r8): compile-time constant in Fortran, correctly rounded */) * 100.0;
}
// Goff-Gratch saturation vapour pressure over ice (Pa).
static inline double gg_svp_ice(double t) {
  const double h2otrip = 273.16;
  return m_pow10(-9.09718 * (h2otrip / t - 1.0) - 3.56654 * m_log10(h2otrip / t) +
                 0.876793 * (1.0 - t / h2otrip) + 0.7858350313586662 /* log10(6.1071_r8) */) * 100.0;
}
static inline double svp_to_qsat(double es, double p, double epsilo) {
  const double omeps = 1.0 - epsilo;
  if ((p - es) <= 0.0) return 1.0;
  return epsilo * es / (p - omeps * es);
}
// qsat_water(t, p[Pa]) -> es[Pa], qs
static inline void qsat_water(double t, double p, double epsilo, double& es, double& qs) {
  es = gg_svp_water(t);
  qs = svp_to_qsat(es, p, epsilo);
  es = std::fmin(es, p);
}

// water/ice transition used to fill the table (ttrice = 20 K)
static inline double svp_trans(double t, double tmelt) {
  const double ttrice = 20.0;
  double es;
  if (t >= (tmelt - ttrice)) es = gg_svp_water(t); else es = 0.0;
  if (t < tmelt) {
    double esice = gg_svp_ice(t);
    double weight;
    if ((tmelt - t) > ttrice) weight = 1.0; else weight = (tmelt - t) / ttrice;
    es = weight * esice + (1.0 - weight) * es;
  }
  return es;
}
struct EsTable {
  static constexpr double tmin = 127.16, tmax = 375.16;
  std::vector<double> estbl;       // 1-based in the Fortran; 0-based here
  void build(double tmelt) {
    int plenest = (int)std::ceil(tmax - tmin) + 1;
    estbl.resize(plenest + 1);
    for (int i = 1; i <= plenest; ++i) estbl[i - 1] = svp_trans(tmin + (double)(i - 1), tmelt);
    estbl[plenest] = estbl[plenest - 1];
  }
  double estblf(double t) const {
    double t_tmp = std::fmax(std::fmin(t, tmax) - tmin, 0.0);
    int i = (int)t_tmp + 1;
    double weight = t_tmp - std::trunc(t_tmp);
    return (1.0 - weight) * estbl[i - 1] + weight * estbl[i];
  }
  // wv_saturation::qsat (table version), p in Pa
  void qsat(double t, double p, double epsilo, double& es, double& qs) const {
    es = estblf(t);
    qs = svp_to_qsat(es, p, epsilo);
    es = std::fmin(es, p);
  }
};

// ---- cloud_fraction::cldfrc_fice --------------------------------------------------------
static inline void cldfrc_fice(double t, double tmelt, double& fice, double& fsnow) {
  const double tmax_fice = tmelt - 10.0, tmin_fice = tmax_fice - 30.0;
  const double tmax_fsnow = tmelt, tmin_fsnow = tmelt - 5.0;
  if (t > tmax_fice) fice = 0.0;
  else if (t < tmin_fice) fice = 1.0;
  else fice = (tmax_fice - t) / (tmax_fice - tmin_fice);
  if (t > tmax_fsnow) fsnow = 0.0;
  else if (t < tmin_fsnow) fsnow = 1.0;
  else fsnow = (tmax_fsnow - t) / (tmax_fsnow - tmin_fsnow);
}

}  // namespace zmo
