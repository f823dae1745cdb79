// zm_oracle.cpp -- TEST INFRASTRUCTURE: CPU oracle for the CAM-Nor Zhang-McFarlane path.
//
// A line-faithful C++ restatement of /root/reference/physics/zm_conv.F90: same chunk-shaped
// (pcols,pver) column-major arrays, same k-outer / i-inner loop nests, same statement and
// operation order (gfortran -O2 semantics: left-to-right, parentheses honoured, integer
// powers as multiply chains, no FMA contraction => build with -ffp-contract=off).
// Every routine cites the reference lines it follows.  The zmconv_microp branches are out of
// scope (zm_microphysics is not in the reference tree) and are not restated; zm_org is.
//
// PARITY PINNING: the reference ships no tests/golden vectors and cannot be compiled here (no
// Fortran compiler; 12 absent modules).  It is pinned instead against the reference's OWN SOURCE
// TEXT: tests/golden/fortran_exec.py translates physics/zm_conv.F90 statement by statement into
// Python and executes it (zm_convi, zm_convr with buoyan_dilute / buoyan / parcel_dilute / entropy /
// enthalpy / ientropy / ienthalpy / qsat_hPa / cldprp / closure / q1q2_pjr inside, zm_conv_evap,
// momtran, convtran; also geopotential_t and convect_diagnostics_calc); the committed fixtures
// tests/golden/reftext_*.npz hold its outputs for 18 configurations and a 507-column sweep, and the glibc-libm build of this
// oracle reproduces every one of them BIT FOR BIT (tests/test_oracle.py::
// test_oracle_equals_reference_source_text).  What stays unpinned: the arithmetic of modules that
// are not in the reference tree (zm_externals.hpp: qsat_water, the qsat table, cldfrc_fice,
// physconst values, qneg3) and the few glue statements of zm_conv_intr.F90 restated in
// conv_tend_chunk.  Further checks: property tests (inverse identities, water closure, zero-tendency
// rows) and cross-agreement of two math back-ends (glibc libm vs the portable zm_math.h).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load this library.
#include "zm_oracle.h"
#include "zm_externals.hpp"
#include <cstdio>
#include <cstring>
#include <vector>
#include <algorithm>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

using zmo::m_log; using zmo::m_log10; using zmo::m_exp; using zmo::m_pow10; using zmo::m_pow;

// ---- module-level state of zm_conv (zm_conv.F90:42-108), read-only after zm_convi -------
struct Module {
  int pcols = 0, pver = 0, pverp = 0;
  double rl, cpres, ke, ke_lnd, c0_lnd, c0_ocn;
  int num_cin; bool zm_org;
  double tau;
  double tfreez, eps1, momcu, momcd;
  bool zmconv_microp, no_deep_pbl;
  double rgrav, rgas, grav, cp;
  int limcnv;
  // logical parameters zm_conv.F90:75-78 (all .true.)
  // reals zm_conv.F90:83-97
  double capelmt = 70.0, capelmt_lnd = 70.0, tiedke_add = 0.5, tiedke_lnd = 1.0,
         cape_tau = 3.6e3, entrmn = 2e-4, alfadet = 0.1, tentrm = 1e-3, tentr_lnd = 1e-3,
         plclmin = 6.e2, cin_threshd = 0.33;
  bool lparcel_pbl; double dmpdz_param;
  const double parcel_hscale = 0.5;
  double dcol;
  bool cam3;
  // physconst
  double cpair, epsilo, gravit, latice, latvap, tmelt, rair, cpwv, cpliq, rh2o, cpvir, zvir;
  zmo::EsTable estbl;
};
Module g;

thread_local long long cnt[10];
thread_local int brent_fail;
// FP64 operation count of the calling thread (+, -, *, /, compare, abs-compare = 1 each; log / log10 /
// 10**x / exp / x**y counted as calls), per phase of zm_convr: 0 = everything else, 1 = first buoyan_dilute call (the
// dominant kernel's algorithmic work), 2 = second buoyan_dilute call.  Columns: 0 basic operations, 1 log, 2 log10,
// 3 10**x, 4 exp, 5 x**y, 6 state-function evaluations, 7 Brent iterations.  Only buoyan_dilute / parcel_dilute and
// what they call are instrumented with basic-operation counts (FL(n) next to the statements they count).
thread_local int fl_phase = 0;
thread_local long long fl[3][8];
#define FL(n) (fl[fl_phase][0] += (n))
// optional trace of Brent inversions (rcall, icol, lchnk, state-function evaluations) for divergence studies
// zm_org (organisation tracer): the pointer dummies org/orgt/org2d of zm_convr (zm_conv.F90:421-423).
// tl_* point at the chunk being processed; the batch drivers set them from the batch-wide base pointers.
thread_local const double* tl_org = nullptr; thread_local double* tl_orgt = nullptr; thread_local double* tl_org2d = nullptr;
const double* batch_org = nullptr; double* batch_orgt = nullptr; double* batch_org2d = nullptr;
// convtran1 of zm_conv_tend (zm_conv_intr.F90:865-880), attached with zmo_convtran1_fields for the next
// zmo_conv_tend_batch: state%q / fracis / ptend_loc%q, (pcols,pver,pcnst) per chunk
struct Tran1 { int pcnst = 0; const int* doconv = nullptr; const int* dry = nullptr; const double* q = nullptr;
               const double* fracis = nullptr; double* ptend_q = nullptr; };
Tran1 batch_tran1;
thread_local int* trace_buf = nullptr; thread_local int trace_n = 0, trace_cap = 0;

inline double fmax2(double a, double b) { return (a > b) ? a : b; }
inline double fmin2(double a, double b) { return (a < b) ? a : b; }
inline double c_log(double x)   { ++cnt[5]; ++fl[fl_phase][1]; return m_log(x); }
inline double c_exp(double x)   { ++cnt[8]; ++fl[fl_phase][4]; return m_exp(x); }
inline double c_pow(double x, double y) { ++cnt[9]; ++fl[fl_phase][5]; return m_pow(x, y); }

// 1-based column-major views -------------------------------------------------------------
struct A2 {
  double* p; int ld;
  inline double& operator()(int i, int k) const { return p[(size_t)(i - 1) + (size_t)ld * (k - 1)]; }
};
struct A2;
struct C2 {
  const double* p; int ld;
  C2(const double* p_, int ld_) : p(p_), ld(ld_) {}
  inline C2(const A2& a);
  inline double operator()(int i, int k) const { return p[(size_t)(i - 1) + (size_t)ld * (k - 1)]; }
};
inline C2::C2(const A2& a) : p(a.p), ld(a.ld) {}
struct A3 {
  double* p; int ld; int n2;
  inline double& operator()(int i, int k, int m) const {
    return p[(size_t)(i - 1) + (size_t)ld * ((size_t)(k - 1) + (size_t)n2 * (m - 1))]; }
};
struct C3 {
  const double* p; int ld; int n2;
  inline double operator()(int i, int k, int m) const {
    return p[(size_t)(i - 1) + (size_t)ld * ((size_t)(k - 1) + (size_t)n2 * (m - 1))]; }
};
struct W2 {   // owned work array
  std::vector<double> v; int ld;
  W2(int pcols, int n, double init = 0.0) : v((size_t)pcols * n, init), ld(pcols) {}
  inline double& operator()(int i, int k) { return v[(size_t)(i - 1) + (size_t)ld * (k - 1)]; }
  operator A2() { return A2{v.data(), ld}; }
  operator C2() const { return C2{v.data(), ld}; }
  double* data() { return v.data(); }
};
struct W1 {
  std::vector<double> v;
  explicit W1(int n, double init = 0.0) : v((size_t)n, init) {}
  inline double& operator()(int i) { return v[(size_t)(i - 1)]; }
};
struct I1 {
  std::vector<int> v;
  explicit I1(int n, int init = 0) : v((size_t)n, init) {}
  inline int& operator()(int i) { return v[(size_t)(i - 1)]; }
};
inline int nint_(double x) { return (int)std::lround(x); }

// ---- qsat_hPa  zm_conv.F90:5421-5437 ----------------------------------------------------
inline void qsat_hPa(double t, double p, double& es, double& qm) {
  ++cnt[0]; cnt[6] += 1; cnt[7] += 3;
  fl[fl_phase][2] += 1; fl[fl_phase][3] += 3;
  // p*100 (1); Goff-Gratch (zm_externals.hpp): 3 divisions, 3 subtractions of 1, 6 products with the coefficients,
  // 2 (10**x - 1), 4 additions of the terms, *100 (19); svp_to_qsat: p-es, <=, 2 products, -, / (6); min (1); *0.01 (1)
  FL(28);
  zmo::qsat_water(t, p * 100.0, g.epsilo, es, qm);
  es = es * 0.01;
}

// ---- entropy  zm_conv.F90:5280-5300 -----------------------------------------------------
double entropy(double TK, double p, double qtot) {
  ++cnt[1]; ++fl[fl_phase][6];
  FL(23);        // L (3), min (1), e (3), the four terms: 2 + 2 | 2 + 1 | 2 | 3, their 3 sums (16)
  const double pref = 1000.0;
  double qv, qst, e, est, L;
  L = g.rl - (g.cpliq - g.cpwv) * (TK - g.tfreez);
  qsat_hPa(TK, p, est, qst);
  qv = fmin2(qtot, qst);
  e = qv * p / (g.eps1 + qv);
  return (g.cpres + qtot * g.cpliq) * c_log(TK / g.tfreez) - g.rgas * c_log((p - e) / pref) +
         L * qv / TK - qv * g.rh2o * c_log(qv / qst);
}

// ---- enthalpy  zm_conv.F90:5440-5457 ----------------------------------------------------
double enthalpy(double TK, double p, double qtot, double z) {
  ++cnt[2]; ++fl[fl_phase][6];
  FL(13);        // L (3), min (1), (cpres + qtot*cpliq)*TK (3), L*qv (1), (1+qtot)*grav*z (3), 2 sums
  double qv, qst, est, L;
  L = g.rl - (g.cpliq - g.cpwv) * (TK - g.tfreez);
  qsat_hPa(TK, p, est, qst);
  qv = fmin2(qtot, qst);
  return (g.cpres + qtot * g.cpliq) * TK + L * qv + (1.0 + qtot) * g.grav * z;
}

// ---- ientropy / ienthalpy  zm_conv.F90:5304-5414 / 5460-5570 ------------------------------
// Brent's method, statement for statement.  kind 0: entropy, 1: enthalpy(z).
template <int KIND>
void invert(int rcall, int icol, int lchnk, double s, double p, double z, double qt,
            double& T, double& qst, double Tfg) {
  ++cnt[3 + KIND];
  const long long ev0 = cnt[1] + cnt[2];
  auto F = [&](double x) { return KIND == 0 ? entropy(x, p, qt) : enthalpy(x, p, qt, z); };
  double est;
  double a, b, c, d = 0.0, ebr = 0.0, fa, fb, fc, pbr, qbr, rbr, sbr, tol1, xm, tol;
  const int LOOPMAX = 100;
  const double EPS = 3.e-8;
  bool converged = false;

  T = Tfg;
  a = Tfg - 10;
  b = Tfg + 10;
  fa = F(a) - s;
  fb = F(b) - s;
  FL(4);
  c = b;
  fc = fb;
  tol = 0.001;

  for (int i = 0; i <= LOOPMAX; ++i) {
    ++fl[fl_phase][7];
    FL(4 + 1 + 4 + 2 + 2);     // sign tests, |fc| < |fb|, tol1, xm, convergence test
    if ((fb > 0.0 && fc > 0.0) || (fb < 0.0 && fc < 0.0)) {
      c = a;
      fc = fa;
      d = b - a;
      ebr = d;
      FL(1);
    }
    if (std::fabs(fc) < std::fabs(fb)) {
      a = b;
      b = c;
      c = a;
      fa = fb;
      fb = fc;
      fc = fa;
    }
    tol1 = 2.0 * EPS * std::fabs(b) + 0.5 * tol;
    xm = 0.5 * (c - b);
    converged = (std::fabs(xm) <= tol1 || fb == 0.0);
    if (converged) break;

    FL(2);                     // the two comparisons that admit an interpolation step
    if (std::fabs(ebr) >= tol1 && std::fabs(fa) > std::fabs(fb)) {
      sbr = fb / fa;
      FL(2);                   // fb/fa, a == c
      if (a == c) {
        pbr = 2.0 * xm * sbr;
        qbr = 1.0 - sbr;
        FL(3);
      } else {
        qbr = fa / fc;
        rbr = fb / fc;
        pbr = sbr * (2.0 * xm * qbr * (qbr - rbr) - (b - a) * (rbr - 1.0));
        qbr = (qbr - 1.0) * (rbr - 1.0) * (sbr - 1.0);
        FL(2 + 9 + 5);
      }
      if (pbr > 0.0) qbr = -qbr;
      pbr = std::fabs(pbr);
      FL(1 + 8);               // pbr > 0; 2 pbr < min(3 xm qbr - |tol1 qbr|, |ebr qbr|)
      if (2.0 * pbr < fmin2(3.0 * xm * qbr - std::fabs(tol1 * qbr), std::fabs(ebr * qbr))) {
        ebr = d;
        d = pbr / qbr;
        FL(1);
      } else {
        d = xm;
        ebr = d;
      }
    } else {
      d = xm;
      ebr = d;
    }
    a = b;
    fa = fb;
    b = b + ((std::fabs(d) > tol1) ? d : std::copysign(tol1, xm));
    fb = F(b) - s;
    FL(3);                     // |d| > tol1, b + step, F - s
  }
  T = b;
  if (trace_buf && trace_n + 4 <= trace_cap) {
    trace_buf[trace_n] = rcall; trace_buf[trace_n + 1] = icol; trace_buf[trace_n + 2] = lchnk;
    trace_buf[trace_n + 3] = (int)(cnt[1] + cnt[2] - ev0); trace_n += 4;
  }
  qsat_hPa(T, p, est, qst);
  if (!converged) ++brent_fail;   // reference: endrun (zm_conv.F90:5401-5410, 5557-5566)
}

// ---- parcel_dilute  zm_conv.F90:4824-5277 -------------------------------------------------
void parcel_dilute(int lchnk, int ncol, int msg, I1& klaunch, C2 p, C2 z, C2 t, C2 q,
                   const double* tpert_, A2 tp, A2 tpv, A2 qstp, W1& pl, double* tl_, W1& ql,
                   int* lcl_, const double* landfrac_, C2 dmpdz) {
  const double* org_ = g.zm_org ? tl_org : nullptr;        // pointer dummy `org` (zm_conv.F90:4865)
  const double org2rkm = 10.0, org2Tpert = 0.0;             // zm_conv.F90:4948-4951
  auto org = [&](int i, int k) { return org_[(size_t)(i - 1) + (size_t)g.pcols * (k - 1)]; };
  auto landfrac = [&](int i) { return landfrac_[i - 1]; };
  const int pcols = g.pcols, pver = g.pver;
  auto tpert = [&](int i) { return tpert_[i - 1]; };
  auto tl = [&](int i) -> double& { return tl_[i - 1]; };
  auto lcl = [&](int i) -> int& { return lcl_[i - 1]; };

  W2 tmix(pcols, pver), qtmix(pcols, pver), qsmix(pcols, pver), smix(pcols, pver);
  W2 xsh2o(pcols, pver + 1), ds_xsh2o(pcols, pver + 1), ds_freeze(pcols, pver + 1);
  W1 zl(pcols), mp(pcols), qtp(pcols), sp(pcols), sp0(pcols), qtp0(pcols), mp0(pcols);
  double lwmax, dmpdp, dpdz, dzdp, senv, qtenv, penv, zenv, tenv, new_s, new_q, dp, tfguess,
      tscool, qxsk, qxskp1, dsdp, dqtdp, dqxsdp, slcl, qtlcl, qslcl, est;
  int rcall, nit_lheat;

  nit_lheat = 2;
  lwmax = 1.e-3;
  tscool = 0.0;
  new_q = 0.0; new_s = 0.0;

  for (int k = pver; k >= msg + 1; --k) {
    for (int i = 1; i <= ncol; ++i) {
      if (k == klaunch(i)) {
        if (g.lparcel_pbl) {
          qtp0(i) = ql(i);
          sp0(i) = enthalpy(tl(i), pl(i), qtp0(i), zl(i));
        } else {
          qtp0(i) = q(i, k);
          sp0(i) = enthalpy(t(i, k), p(i, k), qtp0(i), z(i, k));
        }
        mp0(i) = 1.0;
        smix(i, k) = sp0(i);
        qtmix(i, k) = qtp0(i);
        tmix(i, k) = t(i, k);
        qsat_hPa(tmix(i, k), p(i, k), est, qsmix(i, k));
      }
      if (k < klaunch(i)) {
        // layer means (1 + 4*2), dpdz/dzdp (4), dmpdp (1), sp/qtp/mp updates (3 + 3 + 2), smix/qtmix (3 + 3), LCL test (2)
        FL(9 + 4 + 1 + 8 + 6 + 2);
        dp = (p(i, k) - p(i, k + 1));
        qtenv = 0.5 * (q(i, k) + q(i, k + 1));
        tenv = 0.5 * (t(i, k) + t(i, k + 1));
        penv = 0.5 * (p(i, k) + p(i, k + 1));
        zenv = 0.5 * (z(i, k) + z(i, k + 1));
        senv = enthalpy(tenv, penv, qtenv, zenv);

        dpdz = -(penv * g.grav) / (g.rgas * tenv);
        dzdp = 1.0 / dpdz;
        if (g.zm_org) {                       // zm_conv.F90:5066-5074 (tht_tweaks: dmpdz_lnd = dmpdz_mask)
          double dmpdz_mask = dmpdz(i, k);
          const double dmpdz_lnd = dmpdz_mask;
          dmpdz_mask = landfrac(i) * dmpdz_lnd + (1.0 - landfrac(i)) * dmpdz_mask;
          dmpdp = (dmpdz_mask / (1.0 + org(i, k) * org2rkm)) * dzdp;
        } else {
          dmpdp = dmpdz(i, k) * dzdp;
        }

        sp(i) = sp(i) - dmpdp * dp * senv;
        qtp(i) = qtp(i) - dmpdp * dp * qtenv;
        mp(i) = mp(i) - dmpdp * dp;

        smix(i, k) = (sp0(i) + sp(i)) / (mp0(i) + mp(i));
        qtmix(i, k) = (qtp0(i) + qtp(i)) / (mp0(i) + mp(i));

        tfguess = tmix(i, k + 1);
        rcall = 2;
        invert<1>(rcall, i, lchnk, smix(i, k), p(i, k), z(i, k), qtmix(i, k), tmix(i, k),
                  qsmix(i, k), tfguess);

        if (qsmix(i, k) <= qtmix(i, k) && qsmix(i, k + 1) > qtmix(i, k + 1)) {
          lcl(i) = k;
          FL(2 + 2 + 2 + 3 + 2 + 2 + 3 + 3);   // the interpolation to the LCL (zm_conv.F90:5113-5127)
          qxsk = qtmix(i, k) - qsmix(i, k);
          qxskp1 = qtmix(i, k + 1) - qsmix(i, k + 1);
          dqxsdp = (qxsk - qxskp1) / dp;
          pl(i) = p(i, k + 1) - qxskp1 / dqxsdp;
          zl(i) = z(i, k + 1) - qxskp1 / dqxsdp * dzdp;
          dsdp = (smix(i, k) - smix(i, k + 1)) / dp;
          dqtdp = (qtmix(i, k) - qtmix(i, k + 1)) / dp;
          slcl = smix(i, k + 1) + dsdp * (pl(i) - p(i, k + 1));
          qtlcl = qtmix(i, k + 1) + dqtdp * (pl(i) - p(i, k + 1));
          tfguess = tmix(i, k);
          rcall = 3;
          invert<1>(rcall, i, lchnk, slcl, pl(i), zl(i), qtlcl, tl(i), qslcl, tfguess);
        }
      }
    }
  }

  // PRECIPITATION/FREEZING LOOP  zm_conv.F90:5166-5273
  for (int k = pver; k >= msg + 1; --k) {
    for (int i = 1; i <= ncol; ++i) {
      if (k == klaunch(i)) {
        tp(i, k) = tmix(i, k);
        qstp(i, k) = q(i, k);
        if (g.zm_org)                         // zm_conv.F90:5186-5188
          tpv(i, k) = (tp(i, k) + (org2Tpert * org(i, k) + tpert(i))) * (1.0 + qstp(i, k) / g.eps1) / (1.0 + qstp(i, k));
        else
          tpv(i, k) = (tp(i, k) + tpert(i)) * (1.0 + qstp(i, k) / g.eps1) / (1.0 + qstp(i, k));
      }
      if (k < klaunch(i)) {
        smix(i, k) = entropy(tmix(i, k), p(i, k), qtmix(i, k));
        FL(6 + 1);                     // tpv of this level (6), new_q > qsmix
        for (int ii = 0; ii <= nit_lheat - 1; ++ii) {
          // xsh2o (3), ds_xsh2o (1 division, 1 difference, max, 2 products, 1 difference = 6), freezing tests (4) and
          // term (up to 5), new_s (2), new_q (1)
          FL(3 + 6 + 4 + 5 + 2 + 1);
          xsh2o(i, k) = fmax2(0.0, qtmix(i, k) - qsmix(i, k) - lwmax);
          ds_xsh2o(i, k) = ds_xsh2o(i, k + 1) -
                           g.cpliq * c_log(tmix(i, k) / g.tfreez) *
                               fmax2(0.0, (xsh2o(i, k) - xsh2o(i, k + 1)));
          if (tmix(i, k) <= g.tfreez + tscool && ds_freeze(i, k + 1) == 0.0) {
            ds_freeze(i, k) = (g.latice / tmix(i, k)) *
                              fmax2(0.0, qtmix(i, k) - qsmix(i, k) - xsh2o(i, k));
          }
          if (tmix(i, k) <= g.tfreez + tscool && ds_freeze(i, k + 1) != 0.0) {
            ds_freeze(i, k) = ds_freeze(i, k + 1) +
                              (g.latice / tmix(i, k)) * fmax2(0.0, (qsmix(i, k + 1) - qsmix(i, k)));
          }
          new_s = smix(i, k) + ds_xsh2o(i, k) + ds_freeze(i, k);
          new_q = qtmix(i, k) - xsh2o(i, k);
          tfguess = tmix(i, k);
          rcall = 4;
          invert<0>(rcall, i, lchnk, new_s, p(i, k), 0.0, new_q, tmix(i, k), qsmix(i, k), tfguess);
        }
        tp(i, k) = tmix(i, k);
        if (new_q > qsmix(i, k)) {
          qstp(i, k) = qsmix(i, k);
        } else {
          qstp(i, k) = new_q;
        }
        if (g.zm_org)                         // zm_conv.F90:5255-5257
          tpv(i, k) = (tp(i, k) + (org2Tpert * org(i, k) + tpert(i))) * (1.0 + qstp(i, k) / g.eps1) / (1.0 + new_q);
        else
          tpv(i, k) = (tp(i, k) + tpert(i)) * (1.0 + qstp(i, k) / g.eps1) / (1.0 + new_q);
      }
    }
  }
}

// ---- buoyan_dilute  zm_conv.F90:4425-4819 -------------------------------------------------
void buoyan_dilute(int lchnk, int ncol, C2 q, C2 t, C2 p, C2 z, C2 pf, A2 tp, A2 qstp,
                   double* tl_, double* cape_, double* cin_, const double* pblt_, int* lcl_,
                   int* lel_, int* lon_, int* mx_, double rd, double grav, double cp, int msg,
                   C2 zi, const double* zs_, const double* tpert_, const double* landfrac_,
                   C2 dmpdz) {
  (void)cp;
  const int pcols = g.pcols, pver = g.pver;
  auto tl = [&](int i) -> double& { return tl_[i - 1]; };
  auto cape = [&](int i) -> double& { return cape_[i - 1]; };
  auto cin = [&](int i) -> double& { return cin_[i - 1]; };
  auto pblt = [&](int i) { return pblt_[i - 1]; };
  auto zs = [&](int i) { return zs_[i - 1]; };
  auto lcl = [&](int i) -> int& { return lcl_[i - 1]; };
  auto lel = [&](int i) -> int& { return lel_[i - 1]; };
  auto lon = [&](int i) -> int& { return lon_[i - 1]; };
  auto mx = [&](int i) -> int& { return mx_[i - 1]; };

  W2 capeten(pcols, 5), cinten(pcols, 5), tv(pcols, pver), tpv(pcols, pver), buoy(pcols, pver + 1);
  W1 pl(pcols), hmax(pcols), hmn(pcols), ql(pcols);
  std::vector<char> plge600(pcols, 0);
  I1 knt(pcols), klaunch(pcols);
  std::vector<int> lelten((size_t)pcols * 5, 0);
  auto LELTEN = [&](int i, int n) -> int& { return lelten[(size_t)(i - 1) + (size_t)pcols * (n - 1)]; };
  W2 hmn_lev(pcols, pver), dp_lev(pcols, pver), hmn_zdp(pcols, pver), q_zdp(pcols, pver);
  W1 parcel_dz(pcols), parcel_ztop(pcols), parcel_dp(pcols), parcel_hdp(pcols), parcel_qdp(pcols),
      pbl_dz(pcols), hpar(pcols), qpar(pcols);
  double dp_zfrac; int ipar = 0;

  for (int n = 1; n <= 5; ++n)
    for (int i = 1; i <= ncol; ++i) { LELTEN(i, n) = pver; capeten(i, n) = 0.0; cinten(i, n) = 0.0; }

  for (int i = 1; i <= ncol; ++i) {
    lon(i) = std::min(pver, nint_(pblt(i)) + 2);
    knt(i) = 0;
    lel(i) = pver;
    mx(i) = lon(i);
    cape(i) = 0.0;
    hmax(i) = 0.0;
    pbl_dz(i) = z(i, nint_(pblt(i))) - zs(i);
    parcel_dz(i) = fmax2(zi(i, pver), g.parcel_hscale * pbl_dz(i));
    parcel_ztop(i) = parcel_dz(i) + zs(i);
    parcel_hdp(i) = 0.0; parcel_dp(i) = 0.0; parcel_qdp(i) = 0.0; hpar(i) = 0.0; qpar(i) = 0.0;
  }
  FL((long long)ncol * pver * 5);            // tv of every level
  FL((long long)ncol * (pver - msg) * (g.lparcel_pbl ? 22 : 19));   // moist static energy of the launch search
  FL((long long)ncol * (pver - msg) * (5 + 2));                       // tv / buoyancy after the parcel, level tests
  for (int k = 1; k <= pver; ++k)
    for (int i = 1; i <= ncol; ++i) {
      tp(i, k) = t(i, k);
      qstp(i, k) = q(i, k);
      tv(i, k) = t(i, k) * (1.0 + q(i, k) / g.eps1) / (1.0 + q(i, k));
      tpv(i, k) = tv(i, k);
      buoy(i, k) = 0.0;
    }

  if (g.lparcel_pbl) {
    for (int k = 1; k <= pver; ++k)
      for (int i = 1; i <= ncol; ++i) {
        hmn_lev(i, k) = (cp + q(i, k) * g.cpliq) * t(i, k) / (1.0 + q(i, k)) +
                        (1.0 + q(i, k) / g.eps1) / (1.0 + q(i, k)) * grav * z(i, k) +
                        (g.rl - (g.cpliq - g.cpwv) * (t(i, k) - g.tfreez)) * q(i, k);
        dp_lev(i, k) = pf(i, k + 1) - pf(i, k);
        hmn_zdp(i, k) = hmn_lev(i, k) * dp_lev(i, k);
        q_zdp(i, k) = q(i, k) * dp_lev(i, k);
      }
    for (int i = 1; i <= ncol; ++i) {
      for (int k = pver; k >= msg + 1; --k) {
        if (zi(i, k + 1) <= parcel_dz(i)) {
          ipar = k;
          if (k == pver) {
            dp_zfrac = 1.0;
          } else {
            dp_zfrac = fmin2(1.0, (parcel_dz(i) - zi(i, k + 1)) / (zi(i, k) - zi(i, k + 1)));
          }
          parcel_hdp(i) = parcel_hdp(i) + hmn_zdp(i, k) * dp_zfrac;
          parcel_qdp(i) = parcel_qdp(i) + q_zdp(i, k) * dp_zfrac;
          parcel_dp(i) = parcel_dp(i) + dp_lev(i, k) * dp_zfrac;
        }
      }
      hpar(i) = parcel_hdp(i) / parcel_dp(i);
      qpar(i) = parcel_qdp(i) / parcel_dp(i);
      mx(i) = ipar;
    }
  } else {
    for (int k = pver; k >= msg + 1; --k)
      for (int i = 1; i <= ncol; ++i) {
        hmn(i) = (cp + q(i, k) * g.cpliq) * t(i, k) / (1.0 + q(i, k)) +
                 (1.0 + q(i, k) / g.eps1) / (1.0 + q(i, k)) * grav * z(i, k) +
                 (g.rl - (g.cpliq - g.cpwv) * (t(i, k) - g.tfreez)) * q(i, k);
        if (k >= nint_(pblt(i)) && k <= lon(i) && hmn(i) > hmax(i)) {
          hmax(i) = hmn(i);
          mx(i) = k;
        }
      }
  }

  if (g.lparcel_pbl) {
    for (int i = 1; i <= ncol; ++i) {
      lcl(i) = mx(i);
      tl(i) = (hpar(i) - g.rl * qpar(i) - grav * parcel_ztop(i)) / cp;
      ql(i) = qpar(i);
      pl(i) = p(i, mx(i));
    }
  } else {
    for (int i = 1; i <= ncol; ++i) {
      lcl(i) = mx(i);
      tl(i) = t(i, mx(i));
      ql(i) = q(i, mx(i));
      pl(i) = p(i, mx(i));
    }
  }

  for (int i = 1; i <= ncol; ++i) klaunch(i) = mx(i);
  parcel_dilute(lchnk, ncol, msg, klaunch, p, z, t, q, tpert_, tp, tpv, qstp, pl, tl_, ql, lcl_,
                landfrac_, dmpdz);

  for (int i = 1; i <= ncol; ++i) plge600[i - 1] = pl(i) >= g.plclmin;

  for (int k = pver; k >= msg + 1; --k)
    for (int i = 1; i <= ncol; ++i) {
      if (k <= mx(i) && plge600[i - 1]) {
        tv(i, k) = t(i, k) * (1.0 + q(i, k) / g.eps1) / (1.0 + q(i, k));
        buoy(i, k) = tpv(i, k) - tv(i, k) + g.tiedke_add;
      } else {
        qstp(i, k) = q(i, k);
        tp(i, k) = t(i, k);
        tpv(i, k) = tv(i, k);
      }
    }

  for (int k = msg + 2; k <= pver; ++k)
    for (int i = 1; i <= ncol; ++i) {
      if (k < lcl(i) && plge600[i - 1]) {
        if (buoy(i, k + 1) > 0.0 && buoy(i, k) <= 0.0) {
          knt(i) = std::min(g.num_cin, knt(i) + 1);
          LELTEN(i, knt(i)) = k;
        }
      }
    }

  for (int n = 1; n <= g.num_cin; ++n)
    for (int k = msg + 1; k <= pver; ++k)
      for (int i = 1; i <= ncol; ++i) {
        if (plge600[i - 1] && k <= mx(i) && k > LELTEN(i, n)) {
          FL(4 + 5);                   // cape term (division, 2 products, sum), cin term (min, division, 2 products, difference)
          capeten(i, n) = capeten(i, n) + rd * buoy(i, k) * c_log(pf(i, k + 1) / pf(i, k));
          cinten(i, n) = cinten(i, n) - rd * fmin2(buoy(i, k), 0.0) * c_log(pf(i, k + 1) / pf(i, k));
        }
      }

  for (int n = 1; n <= g.num_cin; ++n)
    for (int i = 1; i <= ncol; ++i) {
      if (capeten(i, n) > cape(i)) {
        cape(i) = capeten(i, n);
        cin(i) = cinten(i, n);
        lel(i) = LELTEN(i, n);
      }
    }
  for (int i = 1; i <= ncol; ++i) cape(i) = fmax2(cape(i), 0.0);
}

// ---- buoyan (undilute, cam3 only)  zm_conv.F90:2719-3022 ----------------------------------
void buoyan(int lchnk, int ncol, C2 q, C2 t, C2 p, C2 z, C2 pf, A2 tp, A2 qstp, double* tl_,
            double rl, double* cape_, const double* pblt_, int* lcl_, int* lel_, int* lon_,
            int* mx_, double rd, double grav, double cp, int msg, const double* tpert_) {
  (void)lchnk;
  const int pcols = g.pcols, pver = g.pver;
  const double eps1 = g.eps1;
  auto tl = [&](int i) -> double& { return tl_[i - 1]; };
  auto cape = [&](int i) -> double& { return cape_[i - 1]; };
  auto pblt = [&](int i) { return pblt_[i - 1]; };
  auto tpert = [&](int i) { return tpert_[i - 1]; };
  auto lcl = [&](int i) -> int& { return lcl_[i - 1]; };
  auto lel = [&](int i) -> int& { return lel_[i - 1]; };
  auto lon = [&](int i) -> int& { return lon_[i - 1]; };
  auto mx = [&](int i) -> int& { return mx_[i - 1]; };
  // NB reference declares capeten(pcols,num_cin) but loops n=1,5 (zm_conv.F90:2770 vs 2992):
  // only well defined for num_cin=5; the oracle sizes the arrays 5 wide.
  W2 capeten(pcols, 5), tv(pcols, pver), tpv(pcols, pver), buoy(pcols, pver + 1);
  W1 a1(pcols), a2(pcols), estp(pcols), pl(pcols), plexp(pcols), hmax(pcols), hmn(pcols), y(pcols);
  std::vector<char> plge600(pcols, 0);
  I1 knt(pcols);
  std::vector<int> lelten((size_t)pcols * 5, 0);
  auto LELTEN = [&](int i, int n) -> int& { return lelten[(size_t)(i - 1) + (size_t)pcols * (n - 1)]; };
  double e;

  for (int n = 1; n <= 5; ++n)
    for (int i = 1; i <= ncol; ++i) { LELTEN(i, n) = pver; capeten(i, n) = 0.0; }
  for (int i = 1; i <= ncol; ++i) {
    lon(i) = pver; knt(i) = 0; lel(i) = pver; mx(i) = lon(i); cape(i) = 0.0; hmax(i) = 0.0;
  }
  for (int k = 1; k <= pver; ++k)
    for (int i = 1; i <= ncol; ++i) {
      tp(i, k) = t(i, k);
      qstp(i, k) = q(i, k);
      tv(i, k) = t(i, k) * (1.0 + 1.608 * q(i, k)) / (1.0 + q(i, k));
      tpv(i, k) = tv(i, k);
      buoy(i, k) = 0.0;
    }
  for (int k = pver; k >= msg + 1; --k)
    for (int i = 1; i <= ncol; ++i) {
      hmn(i) = cp * t(i, k) + grav * z(i, k) + rl * q(i, k);
      if (k >= nint_(pblt(i)) && k <= lon(i) && hmn(i) > hmax(i)) { hmax(i) = hmn(i); mx(i) = k; }
    }
  for (int i = 1; i <= ncol; ++i) {
    lcl(i) = mx(i);
    e = p(i, mx(i)) * q(i, mx(i)) / (eps1 + q(i, mx(i)));
    tl(i) = 2840.0 / (3.5 * c_log(t(i, mx(i))) - c_log(e) - 4.805) + 55.0;
    if (tl(i) < t(i, mx(i))) {
      plexp(i) = (1.0 / (0.2854 * (1.0 - 0.28 * q(i, mx(i)))));
      pl(i) = p(i, mx(i)) * c_pow(tl(i) / t(i, mx(i)), plexp(i));
    } else {
      tl(i) = t(i, mx(i));
      pl(i) = p(i, mx(i));
    }
  }
  for (int k = pver; k >= msg + 2; --k)
    for (int i = 1; i <= ncol; ++i)
      if (k <= mx(i) && (p(i, k) > pl(i) && p(i, k - 1) <= pl(i))) lcl(i) = k - 1;
  for (int i = 1; i <= ncol; ++i) plge600[i - 1] = pl(i) >= g.plclmin;

  for (int k = pver; k >= msg + 1; --k)
    for (int i = 1; i <= ncol; ++i)
      if (k > lcl(i) && k <= mx(i) && plge600[i - 1]) {
        tv(i, k) = t(i, k) * (1.0 + 1.608 * q(i, k)) / (1.0 + q(i, k));
        qstp(i, k) = q(i, mx(i));
        tp(i, k) = t(i, mx(i)) * c_pow(p(i, k) / p(i, mx(i)), 0.2854 * (1.0 - 0.28 * q(i, mx(i))));
        tpv(i, k) = (tp(i, k) + tpert(i)) * (1.0 + 1.608 * q(i, mx(i))) / (1.0 + q(i, mx(i)));
        buoy(i, k) = tpv(i, k) - tv(i, k) + g.tiedke_add;
      }

  auto taylor = [&](int i, int k, double yv) {
    double tpk = tp(i, k), qs = qstp(i, k);
    a1(i) = cp / rl + qs * (1.0 + qs / eps1) * rl * eps1 / (rd * (tpk * tpk));
    a2(i) = .5 * (qs * (1.0 + 2.0 / eps1 * qs) * (1.0 + qs / eps1) * (eps1 * eps1) * rl * rl /
                      ((rd * rd) * ((tpk * tpk) * (tpk * tpk))) -
                  qs * (1.0 + qs / eps1) * 2.0 * eps1 * rl / (rd * ((tpk * tpk) * tpk)));
    a1(i) = 1.0 / a1(i);
    a2(i) = -a2(i) * ((a1(i) * a1(i)) * a1(i));
    y(i) = yv;
    tp(i, k) = tp(i, k) + a1(i) * y(i) + a2(i) * (y(i) * y(i));
  };

  for (int k = pver; k >= msg + 1; --k)
    for (int i = 1; i <= ncol; ++i)
      if (k == lcl(i) && plge600[i - 1]) {
        tv(i, k) = t(i, k) * (1.0 + 1.608 * q(i, k)) / (1.0 + q(i, k));
        qstp(i, k) = q(i, mx(i));
        tp(i, k) = tl(i) * c_pow(p(i, k) / pl(i), 0.2854 * (1.0 - 0.28 * qstp(i, k)));
        qsat_hPa(tp(i, k), p(i, k), estp(i), qstp(i, k));
        taylor(i, k, q(i, mx(i)) - qstp(i, k));
        qsat_hPa(tp(i, k), p(i, k), estp(i), qstp(i, k));
        tpv(i, k) = (tp(i, k) + tpert(i)) * (1.0 + 1.608 * qstp(i, k)) / (1.0 + q(i, mx(i)));
        buoy(i, k) = tpv(i, k) - tv(i, k) + g.tiedke_add;
      }
  for (int k = pver - 1; k >= msg + 1; --k)
    for (int i = 1; i <= ncol; ++i)
      if (k < lcl(i) && plge600[i - 1]) {
        tv(i, k) = t(i, k) * (1.0 + 1.608 * q(i, k)) / (1.0 + q(i, k));
        qstp(i, k) = qstp(i, k + 1);
        tp(i, k) = tp(i, k + 1) * c_pow(p(i, k) / p(i, k + 1), 0.2854 * (1.0 - 0.28 * qstp(i, k)));
        qsat_hPa(tp(i, k), p(i, k), estp(i), qstp(i, k));
        taylor(i, k, qstp(i, k + 1) - qstp(i, k));
        qsat_hPa(tp(i, k), p(i, k), estp(i), qstp(i, k));
        tpv(i, k) = (tp(i, k) + tpert(i)) * (1.0 + 1.608 * qstp(i, k)) / (1.0 + q(i, mx(i)));
        buoy(i, k) = tpv(i, k) - tv(i, k) + g.tiedke_add;
      }
  for (int k = msg + 2; k <= pver; ++k)
    for (int i = 1; i <= ncol; ++i)
      if (k < lcl(i) && plge600[i - 1])
        if (buoy(i, k + 1) > 0.0 && buoy(i, k) <= 0.0) {
          knt(i) = std::min(5, knt(i) + 1);
          LELTEN(i, knt(i)) = k;
        }
  for (int n = 1; n <= 5; ++n)
    for (int k = msg + 1; k <= pver; ++k)
      for (int i = 1; i <= ncol; ++i)
        if (plge600[i - 1] && k <= mx(i) && k > LELTEN(i, n))
          capeten(i, n) = capeten(i, n) + rd * buoy(i, k) * c_log(pf(i, k + 1) / pf(i, k));
  for (int n = 1; n <= 5; ++n)
    for (int i = 1; i <= ncol; ++i)
      if (capeten(i, n) > cape(i)) { cape(i) = capeten(i, n); lel(i) = LELTEN(i, n); }
  for (int i = 1; i <= ncol; ++i) cape(i) = fmax2(cape(i), 0.0);
}

// ---- cldprp  zm_conv.F90:3024-4026 (zmconv_microp = .false. branches only) ------------------
void cldprp(int lchnk, C2 q, C2 t, C2 u, C2 v, C2 p, C2 z, C2 s, A2 mu, A2 eu, A2 du, A2 md,
            A2 ed, A2 sd, A2 qd, A2 mc, A2 qu, A2 su, C2 zf, A2 qst, A2 hmn, A2 hsat, C2 shat,
            A2 ql, A2 cmeg, const int* jb_, const int* lel_, int* jt_, int* jlcl_,
            const int* mx_, int* j0_, int* jd_, double rl, int il2g, double rd, double grav,
            double cp, int msg, A2 pflx, A2 evp, A2 cu, A2 rprd, int limcnv,
            const double* landfrac_, A2 qcde, C2 qhat) {
  (void)lchnk; (void)u; (void)v; (void)qhat;
  const int pcols = g.pcols, pver = g.pver, pverp = g.pverp;
  const double eps1 = g.eps1, zvir = g.zvir, cpvir = g.cpvir, dcol = g.dcol, tmelt = g.tmelt;
  auto jb = [&](int i) { return jb_[i - 1]; };
  auto lel = [&](int i) { return lel_[i - 1]; };
  auto mx = [&](int i) { return mx_[i - 1]; };
  auto jt = [&](int i) -> int& { return jt_[i - 1]; };
  auto jlcl = [&](int i) -> int& { return jlcl_[i - 1]; };
  auto j0 = [&](int i) -> int& { return j0_[i - 1]; };
  auto jd = [&](int i) -> int& { return jd_[i - 1]; };
  auto landfrac = [&](int i) { return landfrac_[i - 1]; };

  W2 gamma(pcols, pver), dz(pcols, pver), iprm(pcols, pver), hu(pcols, pver), hd(pcols, pver),
      eps(pcols, pver + 1), f(pcols, pver + 1), k1(pcols, pver + 1), i2(pcols, pver + 1), ihat(pcols, pver),
      i3(pcols, pver + 1), idag(pcols, pver), i4(pcols, pver + 1), qsthat(pcols, pver), hsthat(pcols, pver),
      gamhat(pcols, pver), qds(pcols, pver);
  W2 mcp(pcols, pver), mrd(pcols, pver), mrl(pcols, pver), tu(pcols, pver), td(pcols, pver);
  W1 c0mask(pcols), tiedke_msk(pcols), hmin(pcols), expdif(pcols), expnum(pcols), ftemp(pcols),
      eps0(pcols), rmue(pcols), zuef(pcols), zdef(pcols), epsm(pcols), ratmjb(pcols), est(pcols),
      totpcp(pcols), totevp(pcols), alfa(pcols), totfrz(pcols);
  W2 frz(pcols, pver);
  I1 tmplel(pcols);
  std::vector<char> doit(pcols, 0), done(pcols, 0);
  double ql1, estu, qstu, small, mdt;
  int khighest, klowest, kount;

  for (int i = 1; i <= il2g; ++i) {
    ftemp(i) = 0.0; expnum(i) = 0.0; expdif(i) = 0.0;
    c0mask(i) = g.c0_ocn * (1.0 - landfrac(i)) + g.c0_lnd * landfrac(i);
    tiedke_msk(i) = g.tiedke_add * (1.0 - landfrac(i)) + g.tiedke_lnd * landfrac(i);
  }
  for (int k = 1; k <= pver; ++k)
    for (int i = 1; i <= il2g; ++i) dz(i, k) = zf(i, k) - zf(i, k + 1);

  for (int i = 1; i <= il2g; ++i) pflx(i, 1) = 0;

  for (int k = 1; k <= pver; ++k)
    for (int i = 1; i <= il2g; ++i) {
      k1(i, k) = 0.0; i2(i, k) = 0.0; i3(i, k) = 0.0; i4(i, k) = 0.0;
      mu(i, k) = 0.0; f(i, k) = 0.0; eps(i, k) = 0.0; eu(i, k) = 0.0; du(i, k) = 0.0;
      ql(i, k) = 0.0; cu(i, k) = 0.0; evp(i, k) = 0.0; cmeg(i, k) = 0.0;
      qds(i, k) = q(i, k);
      md(i, k) = 0.0; ed(i, k) = 0.0;
      sd(i, k) = s(i, k);
      qd(i, k) = q(i, k);
      mc(i, k) = 0.0;
      qu(i, k) = q(i, k);
      su(i, k) = s(i, k);
      qsat_hPa(t(i, k), p(i, k), est(i), qst(i, k));
      if (p(i, k) - est(i) <= 0.0) qst(i, k) = 1.0;
      mrd(i, k) = (1.0 + zvir * q(i, k)) * rd;
      mcp(i, k) = (1.0 + cpvir * q(i, k)) * cp;
      mrl(i, k) = (1.0 - dcol * (t(i, k) - tmelt)) * rl;
      gamma(i, k) = qst(i, k) * (1.0 + qst(i, k) / eps1) * eps1 * mrl(i, k) /
                    (mrd(i, k) * (t(i, k) * t(i, k))) * mrl(i, k) / mcp(i, k);
      hmn(i, k) = mcp(i, k) * t(i, k) + grav * z(i, k) + mrl(i, k) * q(i, k);
      hsat(i, k) = mcp(i, k) * t(i, k) + grav * z(i, k) + mrl(i, k) * qst(i, k);
      hu(i, k) = hmn(i, k);
      hd(i, k) = hmn(i, k);
      rprd(i, k) = 0.0;
      qcde(i, k) = 0.0;
      frz(i, k) = 0.0;
      // tvuo/tvu (zm_conv.F90:3305-3307) feed only the microphysics branch: not restated.
      td(i, k) = (hd(i, k) - grav * zf(i, k) - (1.0 + dcol * tmelt) * rl * qds(i, k)) /
                 (cp * (1.0 + (cpvir - dcol * (rl / cp)) * qds(i, k)));
    }
  for (int k = 1; k <= msg; ++k)
    for (int i = 1; i <= il2g; ++i) rprd(i, k) = 0.0;

  for (int k = 1; k <= msg + 1; ++k)
    for (int i = 1; i <= il2g; ++i) {
      hsthat(i, k) = hsat(i, k);
      qsthat(i, k) = qst(i, k);
      gamhat(i, k) = gamma(i, k);
    }
  for (int i = 1; i <= il2g; ++i) { totpcp(i) = 0.0; totevp(i) = 0.0; }
  for (int k = msg + 2; k <= pver; ++k)
    for (int i = 1; i <= il2g; ++i) {
      if (std::fabs(qst(i, k - 1) - qst(i, k)) > 1.E-6) {
        qsthat(i, k) = c_log(qst(i, k - 1) / qst(i, k)) * qst(i, k - 1) * qst(i, k) /
                       (qst(i, k - 1) - qst(i, k));
      } else {
        qsthat(i, k) = qst(i, k);
      }
      hsthat(i, k) = mcp(i, k) * shat(i, k) + mrl(i, k) * qsthat(i, k);
      if (std::fabs(gamma(i, k - 1) - gamma(i, k)) > 1.E-6) {
        gamhat(i, k) = c_log(gamma(i, k - 1) / gamma(i, k)) * gamma(i, k - 1) * gamma(i, k) /
                       (gamma(i, k - 1) - gamma(i, k));
      } else {
        gamhat(i, k) = gamma(i, k);
      }
    }

  for (int i = 1; i <= pcols; ++i) jt(i) = pver;
  for (int i = 1; i <= il2g; ++i) {
    jt(i) = std::max(lel(i), limcnv + 1);
    jt(i) = std::min(jt(i), pver);
    jd(i) = pver;
    jlcl(i) = lel(i);
    hmin(i) = 1.E6;
  }
  for (int k = msg + 1; k <= pver; ++k)
    for (int i = 1; i <= il2g; ++i)
      if (hsat(i, k) <= hmin(i) && k >= jt(i) && k <= jb(i)) {
        hmin(i) = hsat(i, k);
        j0(i) = k;
      }
  for (int i = 1; i <= il2g; ++i) {
    j0(i) = std::min(j0(i), jb(i) - 2);
    j0(i) = std::max(j0(i), jt(i) + 2);
    j0(i) = std::min(j0(i), pver);
  }
  for (int k = msg + 1; k <= pver; ++k)
    for (int i = 1; i <= il2g; ++i)
      if (k >= jt(i) && k <= jb(i)) {
        hu(i, k) = hmn(i, mx(i)) + cp * tiedke_msk(i);
        su(i, k) = s(i, mx(i)) + tiedke_msk(i) / (1.0 + cpvir * qu(i, k));
      }

  for (int k = pver - 1; k >= msg + 1; --k)
    for (int i = 1; i <= il2g; ++i)
      if (k < jb(i) && k >= jt(i)) {
        k1(i, k) = k1(i, k + 1) + (hmn(i, mx(i)) - hmn(i, k)) * dz(i, k);
        ihat(i, k) = 0.5 * (k1(i, k + 1) + k1(i, k));
        i2(i, k) = i2(i, k + 1) + ihat(i, k) * dz(i, k);
        idag(i, k) = 0.5 * (i2(i, k + 1) + i2(i, k));
        i3(i, k) = i3(i, k + 1) + idag(i, k) * dz(i, k);
        iprm(i, k) = 0.5 * (i3(i, k + 1) + i3(i, k));
        i4(i, k) = i4(i, k + 1) + iprm(i, k) * dz(i, k);
      }

  for (int i = 1; i <= il2g; ++i) hmin(i) = 1.E6;
  for (int k = msg + 1; k <= pver; ++k)
    for (int i = 1; i <= il2g; ++i)
      if (k >= j0(i) && k <= jb(i) && hmn(i, k) <= hmin(i)) {
        hmin(i) = hmn(i, k);
        expdif(i) = hmn(i, mx(i)) - hmin(i);
      }

  for (int k = msg + 2; k <= pver; ++k)
    for (int i = 1; i <= il2g; ++i) {
      expnum(i) = 0.0;
      ftemp(i) = 0.0;
      if (k < jt(i) || k >= jb(i)) {
        k1(i, k) = 0.0;
        expnum(i) = 0.0;
      } else {
        expnum(i) = hmn(i, mx(i)) - (hsat(i, k - 1) * (zf(i, k) - z(i, k)) +
                                     hsat(i, k) * (z(i, k - 1) - zf(i, k))) / (z(i, k - 1) - z(i, k));
      }
      if ((expdif(i) > 100.0 && expnum(i) > 0.0) && k1(i, k) > expnum(i) * dz(i, k)) {
        ftemp(i) = expnum(i) / k1(i, k);
        double ft = ftemp(i), K1 = k1(i, k), I2 = i2(i, k), I3 = i3(i, k), I4 = i4(i, k);
        f(i, k) = ft + I2 / K1 * (ft * ft) +
                  (2.0 * (I2 * I2) - K1 * I3) / (K1 * K1) * ((ft * ft) * ft) +
                  (-5.0 * K1 * I2 * I3 + 5.0 * ((I2 * I2) * I2) + (K1 * K1) * I4) /
                      ((K1 * K1) * K1) * ((ft * ft) * (ft * ft));
        f(i, k) = fmax2(f(i, k), 0.0);
        f(i, k) = fmin2(f(i, k), g.entrmn);
      }
    }
  for (int i = 1; i <= il2g; ++i)
    if (j0(i) < jb(i))
      if (f(i, j0(i)) < 1.E-6 && f(i, j0(i) + 1) > f(i, j0(i))) j0(i) = j0(i) + 1;
  for (int k = msg + 2; k <= pver; ++k)
    for (int i = 1; i <= il2g; ++i)
      if (k >= jt(i) && k <= j0(i)) f(i, k) = fmax2(f(i, k), f(i, k - 1));
  for (int i = 1; i <= il2g; ++i) {
    eps0(i) = f(i, j0(i));
    eps(i, jb(i)) = eps0(i);
  }
  for (int k = pver; k >= msg + 1; --k)
    for (int i = 1; i <= il2g; ++i)
      if (k >= j0(i) && k <= jb(i)) eps(i, k) = f(i, j0(i));
  for (int k = pver; k >= msg + 1; --k)
    for (int i = 1; i <= il2g; ++i)
      if (k < j0(i) && k >= jt(i)) eps(i, k) = f(i, k);

  // itnum = 1 (zmconv_microp false): single pass of the iter loop, zm_conv.F90:3526-3874
  for (int i = 1; i <= il2g; ++i) {
    if (eps0(i) > 0.0) {
      mu(i, jb(i)) = 1.0;
      eu(i, jb(i)) = mu(i, jb(i)) / dz(i, jb(i));
    }
    tmplel(i) = jt(i);
  }
  for (int k = pver; k >= msg + 1; --k)
    for (int i = 1; i <= il2g; ++i)
      if (eps0(i) > 0.0 && (k >= tmplel(i) && k < jb(i))) {
        zuef(i) = zf(i, k) - zf(i, jb(i));
        rmue(i) = (1.0 / eps0(i)) * (c_exp(eps(i, k + 1) * zuef(i)) - 1.0) / zuef(i);
        mu(i, k) = (1.0 / eps0(i)) * (c_exp(eps(i, k) * zuef(i)) - 1.0) / zuef(i);
        eu(i, k) = (rmue(i) - mu(i, k + 1)) / dz(i, k);
        du(i, k) = (rmue(i) - mu(i, k)) / dz(i, k);
      }

  khighest = pverp;
  klowest = 1;
  for (int i = 1; i <= il2g; ++i) {
    khighest = std::min(khighest, lel(i));
    klowest = std::max(klowest, jb(i));
  }
  for (int k = klowest - 1; k >= khighest; --k)
    for (int i = 1; i <= il2g; ++i)
      if (k <= jb(i) - 1 && k >= lel(i) && eps0(i) > 0.0) {
        if (mu(i, k) < 0.02) {
          hu(i, k) = hmn(i, k);
          mu(i, k) = 0.0;
          eu(i, k) = 0.0;
          du(i, k) = mu(i, k + 1) / dz(i, k);
        } else {
          hu(i, k) = mu(i, k + 1) / mu(i, k) * hu(i, k + 1) +
                     dz(i, k) / mu(i, k) * (eu(i, k) * hmn(i, k) - du(i, k) * hsat(i, k));
        }
      }

  for (int i = 1; i <= il2g; ++i) {
    doit[i - 1] = 1;
    totfrz(i) = 0.0;
    for (int k = pver; k >= msg + 1; --k) totfrz(i) = totfrz(i) + frz(i, k) * dz(i, k);
  }
  for (int k = klowest - 2; k >= khighest - 1; --k)
    for (int i = 1; i <= il2g; ++i)
      if (doit[i - 1] && k <= jb(i) - 2 && k >= lel(i) - 1) {
        if (hu(i, k) <= hsthat(i, k) && hu(i, k + 1) > hsthat(i, k + 1) && mu(i, k) >= 0.02) {
          if (hu(i, k) - hsthat(i, k) < -2000.0) {
            jt(i) = k + 1;
            doit[i - 1] = 0;
          } else {
            jt(i) = k;
            doit[i - 1] = 0;
          }
        } else if ((hu(i, k) > hu(i, jb(i)) && totfrz(i) <= 0.0) || mu(i, k) < 0.02) {
          jt(i) = k + 1;
          doit[i - 1] = 0;
        }
      }

  for (int k = pver; k >= msg + 1; --k)
    for (int i = 1; i <= il2g; ++i) {
      if (k >= lel(i) && k <= jt(i) && eps0(i) > 0.0) {
        mu(i, k) = 0.0;
        eu(i, k) = 0.0;
        du(i, k) = 0.0;
        hu(i, k) = hmn(i, k);
      }
      if (k == jt(i) && eps0(i) > 0.0) {
        du(i, k) = mu(i, k + 1) / dz(i, k);
        eu(i, k) = 0.0;
        mu(i, k) = 0.0;
      }
    }

  for (int k = pver; k >= msg + 2; --k)
    for (int i = 1; i <= il2g; ++i)
      tu(i, k) = (hu(i, k) - grav * zf(i, k) - (1.0 + dcol * tmelt) * rl * qu(i, k)) /
                 (cp * (1.0 + (cpvir - dcol * (rl / cp)) * qu(i, k)));

  for (int i = 1; i <= il2g; ++i) done[i - 1] = 0;
  kount = 0;
  for (int k = pver; k >= msg + 2; --k) {
    for (int i = 1; i <= il2g; ++i) {
      if (k == jb(i) && eps0(i) > 0.0) {
        qu(i, k) = q(i, mx(i));
        tu(i, k) = (hu(i, k) - grav * zf(i, k) - (1.0 + dcol * tmelt) * rl * qu(i, k)) /
                   (cp * (1.0 + (cpvir - dcol * (rl / cp)) * qu(i, k)));
        su(i, k) = (hu(i, k) - (1.0 - dcol * (tu(i, k) - tmelt)) * rl * qu(i, k)) /
                   ((1.0 + cpvir * qu(i, k)) * cp);
      }
      if ((!done[i - 1] && k > jt(i) && k < jb(i)) && eps0(i) > 0.0) {
        su(i, k) = mu(i, k + 1) / mu(i, k) * su(i, k + 1) +
                   dz(i, k) / mu(i, k) * (eu(i, k) - du(i, k)) * s(i, k);
        qu(i, k) = mu(i, k + 1) / mu(i, k) * qu(i, k + 1) +
                   dz(i, k) / mu(i, k) * (eu(i, k) * q(i, k) - du(i, k) * qst(i, k));
        // `0.85` is a default-real (single precision) literal in the reference, zm_conv.F90:3680
        tu(i, k) = su(i, k) - grav / ((1.0 + (double)0.85f * qu(i, k)) * cp) * zf(i, k);
        qsat_hPa(tu(i, k), (p(i, k) + p(i, k - 1)) / 2.0, estu, qstu);
        if (qu(i, k) >= qstu) {
          jlcl(i) = k;
          kount = kount + 1;
          done[i - 1] = 1;
        }
      }
    }
    if (kount >= il2g) break;
  }
  for (int k = msg + 2; k <= pver; ++k)
    for (int i = 1; i <= il2g; ++i)
      if ((k > jt(i) && k <= jlcl(i)) && eps0(i) > 0.0) {
        qu(i, k) = qsthat(i, k) + gamhat(i, k) * (hu(i, k) - hsthat(i, k)) /
                                      ((1.0 - dcol * (tu(i, k) - tmelt)) * rl * (1.0 + gamhat(i, k)));
        su(i, k) = shat(i, k) + (hu(i, k) - hsthat(i, k)) /
                                    ((1.0 + cpvir * qu(i, k)) * cp * (1.0 + gamhat(i, k)));
        tu(i, k) = su(i, k) - grav / ((1.0 + cpvir * qu(i, k)) * cp) * zf(i, k);
      }

  for (int i = 1; i <= il2g; ++i) tmplel(i) = jb(i);
  for (int k = pver; k >= msg + 2; --k)
    for (int i = 1; i <= il2g; ++i)
      if (k >= jt(i) && k < tmplel(i) && eps0(i) > 0.0) {
        cu(i, k) = ((mu(i, k) * su(i, k) - mu(i, k + 1) * su(i, k + 1)) / dz(i, k) -
                    (eu(i, k) - du(i, k)) * s(i, k)) / (rl / cp) *
                   ((1.0 + cpvir * qu(i, k)) / (1.0 - dcol * (tu(i, k) - tmelt)));
        if (k == jt(i)) cu(i, k) = 0.0;
        cu(i, k) = fmax2(0.0, cu(i, k));
      }

  for (int k = pver; k >= msg + 2; --k)
    for (int i = 1; i <= il2g; ++i) {
      rprd(i, k) = 0.0;
      if (k >= jt(i) && k < jb(i) && eps0(i) > 0.0 && mu(i, k) >= 0.0) {
        if (mu(i, k) > 0.0) {
          ql1 = 1.0 / mu(i, k) * (mu(i, k + 1) * ql(i, k + 1) - dz(i, k) * du(i, k) * ql(i, k + 1) +
                                  dz(i, k) * cu(i, k));
          ql(i, k) = ql1 / (1.0 + dz(i, k) * c0mask(i));
        } else {
          ql(i, k) = 0.0;
        }
        totpcp(i) = totpcp(i) + dz(i, k) * (cu(i, k) - du(i, k) * ql(i, k + 1));
        rprd(i, k) = c0mask(i) * mu(i, k) * ql(i, k);
        qcde(i, k) = ql(i, k);
      }
    }

  // downdraft  zm_conv.F90:3880-3975
  for (int i = 1; i <= il2g; ++i) {
    alfa(i) = g.alfadet;
    jt(i) = std::min(jt(i), jb(i) - 1);
    jd(i) = std::max(j0(i), jt(i) + 1);
    jd(i) = std::min(jd(i), jb(i));
    hd(i, jd(i)) = hmn(i, jd(i) - 1);
    if (jd(i) < jb(i) && eps0(i) > 0.0) {
      epsm(i) = eps0(i);
      md(i, jd(i)) = -alfa(i) * epsm(i) / eps0(i);
    }
  }
  for (int k = msg + 1; k <= pver; ++k)
    for (int i = 1; i <= il2g; ++i)
      if ((k > jd(i) && k <= jb(i)) && eps0(i) > 0.0) {
        zdef(i) = zf(i, jd(i)) - zf(i, k);
        md(i, k) = -alfa(i) / (2.0 * eps0(i)) * (c_exp(2.0 * epsm(i) * zdef(i)) - 1.0) / zdef(i);
      }
  for (int k = msg + 1; k <= pver; ++k)
    for (int i = 1; i <= il2g; ++i)
      if ((k >= jt(i) && k <= jb(i)) && eps0(i) > 0.0 && jd(i) < jb(i)) {
        ratmjb(i) = fmin2(std::fabs(mu(i, jb(i)) / md(i, jb(i))), 1.0);
        md(i, k) = md(i, k) * ratmjb(i);
      }

  small = 1.e-20;
  for (int k = msg + 1; k <= pver; ++k)
    for (int i = 1; i <= il2g; ++i)
      if ((k >= jt(i) && k <= pver) && eps0(i) > 0.0) {
        ed(i, k - 1) = (md(i, k - 1) - md(i, k)) / dz(i, k - 1);
        mdt = fmin2(md(i, k), -small);
        hd(i, k) = (md(i, k - 1) * hd(i, k - 1) - dz(i, k - 1) * ed(i, k - 1) * hmn(i, k - 1)) / mdt;
      }

  for (int k = msg + 2; k <= pver; ++k)
    for (int i = 1; i <= il2g; ++i)
      if ((k >= jd(i) && k <= jb(i)) && eps0(i) > 0.0 && jd(i) < jb(i)) {
        qds(i, k) = qsthat(i, k) + gamhat(i, k) * (hd(i, k) - hsthat(i, k)) / (rl * (1.0 + gamhat(i, k)));
        td(i, k) = (hd(i, k) - grav * zf(i, k) - (1.0 + dcol * tmelt) * rl * qds(i, k)) /
                   (cp * (1.0 + (cpvir - dcol * (rl / cp)) * qds(i, k)));
        qds(i, k) = qsthat(i, k) + gamhat(i, k) * (hd(i, k) - hsthat(i, k)) /
                                       ((1.0 - dcol * (td(i, k) - tmelt)) * rl * (1.0 + gamhat(i, k)));
      }

  for (int i = 1; i <= il2g; ++i) {
    qd(i, jd(i)) = qds(i, jd(i));
    int k = jd(i);
    sd(i, jd(i)) = (hd(i, jd(i)) - (1.0 - dcol * (td(i, k) - tmelt)) * rl * qd(i, jd(i))) /
                   ((1.0 + cpvir * qd(i, k)) * cp);
    td(i, k) = sd(i, k) - grav / ((1.0 + cpvir * qd(i, k)) * cp) * zf(i, k);
  }

  for (int k = msg + 2; k <= pver; ++k)
    for (int i = 1; i <= il2g; ++i)
      if (k >= jd(i) && k < jb(i) && eps0(i) > 0.0) {
        qd(i, k + 1) = qds(i, k + 1);
        evp(i, k) = -ed(i, k) * q(i, k) + (md(i, k) * qd(i, k) - md(i, k + 1) * qd(i, k + 1)) / dz(i, k);
        evp(i, k) = fmax2(evp(i, k), 0.0);
        mdt = fmin2(md(i, k + 1), -small);
        sd(i, k + 1) = (((1.0 - dcol * (td(i, k) - tmelt)) * rl / ((1.0 + cpvir * qd(i, k)) * cp) * evp(i, k) -
                         ed(i, k) * s(i, k)) * dz(i, k) + md(i, k) * sd(i, k)) / mdt;
        totevp(i) = totevp(i) - dz(i, k) * ed(i, k) * q(i, k);
      }
  for (int i = 1; i <= il2g; ++i)
    totevp(i) = totevp(i) + md(i, jd(i)) * qd(i, jd(i)) - md(i, jb(i)) * qd(i, jb(i));

  for (int i = 1; i <= il2g; ++i) {
    totpcp(i) = fmax2(totpcp(i), 0.0);
    totevp(i) = fmax2(totevp(i), 0.0);
  }
  for (int k = msg + 2; k <= pver; ++k)
    for (int i = 1; i <= il2g; ++i) {
      if (totevp(i) > 0.0 && totpcp(i) > 0.0) {
        md(i, k) = md(i, k) * fmin2(1.0, totpcp(i) / (totevp(i) + totpcp(i)));
        ed(i, k) = ed(i, k) * fmin2(1.0, totpcp(i) / (totevp(i) + totpcp(i)));
        evp(i, k) = evp(i, k) * fmin2(1.0, totpcp(i) / (totevp(i) + totpcp(i)));
      } else {
        md(i, k) = 0.0;
        ed(i, k) = 0.0;
        evp(i, k) = 0.0;
      }
      cmeg(i, k) = cu(i, k) - evp(i, k);
      rprd(i, k) = rprd(i, k) - evp(i, k);
    }

  for (int i = 1; i <= il2g; ++i) pflx(i, 1) = 0.0;
  for (int k = 2; k <= pverp; ++k)
    for (int i = 1; i <= il2g; ++i) pflx(i, k) = pflx(i, k - 1) + rprd(i, k - 1) * dz(i, k - 1);

  for (int k = msg + 1; k <= pver; ++k)
    for (int i = 1; i <= il2g; ++i) mc(i, k) = mu(i, k) + md(i, k);
}

// ---- closure  zm_conv.F90:4028-4260 -------------------------------------------------------
void closure(int lchnk, C2 q, C2 t, C2 p, C2 z, C2 s, C2 tp, C2 qs, C2 qu, C2 su, C2 mc, C2 du,
             C2 mu, C2 md, C2 qd, C2 sd, C2 qhat, C2 shat, C2 dp, C2 qstp, C2 zf, C2 ql,
             const double* dsubcld_, double* mb_, const double* cape_, const double* tl_,
             const int* lcl_, const int* lel_, const int* jt_, const int* mx_, int il1g, int il2g,
             double rd, double grav, double cp, double rl, int msg, double capelmt) {
  (void)lchnk; (void)z; (void)qs; (void)rl;
  const int pcols = g.pcols, pver = g.pver;
  const double eps1 = g.eps1;
  auto dsubcld = [&](int i) { return dsubcld_[i - 1]; };
  auto mb = [&](int i) -> double& { return mb_[i - 1]; };
  auto cape = [&](int i) { return cape_[i - 1]; };
  auto tl = [&](int i) { return tl_[i - 1]; };
  auto lcl = [&](int i) { return lcl_[i - 1]; };
  auto lel = [&](int i) { return lel_[i - 1]; };
  auto jt = [&](int i) { return jt_[i - 1]; };
  auto mx = [&](int i) { return mx_[i - 1]; };
  W2 dtpdt(pcols, pver), dqsdtp(pcols, pver), dtmdt(pcols, pver), dqmdt(pcols, pver),
      dboydt(pcols, pver), thetavp(pcols, pver), thetavm(pcols, pver);
  W1 dtbdt(pcols), dqbdt(pcols), dtldt(pcols), dadt(pcols);
  double beta, debdt, dltaa, eb;
  int kmin, kmax;
  rl = g.rl;

  for (int i = il1g; i <= il2g; ++i) {
    mb(i) = 0.0;
    eb = p(i, mx(i)) * q(i, mx(i)) / (eps1 + q(i, mx(i)));
    dtbdt(i) = (1.0 / dsubcld(i)) * (mu(i, mx(i)) * (shat(i, mx(i)) - su(i, mx(i))) +
                                     md(i, mx(i)) * (shat(i, mx(i)) - sd(i, mx(i))));
    dqbdt(i) = (1.0 / dsubcld(i)) * (mu(i, mx(i)) * (qhat(i, mx(i)) - qu(i, mx(i))) +
                                     md(i, mx(i)) * (qhat(i, mx(i)) - qd(i, mx(i))));
    double epq = eps1 + q(i, mx(i));
    debdt = eps1 * p(i, mx(i)) / (epq * epq) * dqbdt(i);
    double den = 3.5 * c_log(t(i, mx(i))) - c_log(eb) - 4.805;
    dtldt(i) = -2840.0 * (3.5 / t(i, mx(i)) * dtbdt(i) - debdt / eb) / (den * den);
  }
  for (int k = msg + 1; k <= pver; ++k)
    for (int i = il1g; i <= il2g; ++i) { dtmdt(i, k) = 0.0; dqmdt(i, k) = 0.0; }

  for (int k = msg + 1; k <= pver - 1; ++k)
    for (int i = il1g; i <= il2g; ++i)
      if (k == jt(i)) {
        dqmdt(i, k) = (1.0 / dp(i, k)) * (mu(i, k + 1) * (qu(i, k + 1) - qhat(i, k + 1) + ql(i, k + 1)) +
                                          md(i, k + 1) * (qd(i, k + 1) - qhat(i, k + 1)));
        dtmdt(i, k) = (1.0 / dp(i, k)) * (mu(i, k + 1) * (su(i, k + 1) - shat(i, k + 1) - rl / cp * ql(i, k + 1)) +
                                          md(i, k + 1) * (sd(i, k + 1) - shat(i, k + 1)));
      }

  beta = 0.0;
  for (int k = msg + 1; k <= pver - 1; ++k)
    for (int i = il1g; i <= il2g; ++i)
      if (k > jt(i) && k < mx(i)) {
        dtmdt(i, k) = (mc(i, k) * (shat(i, k) - s(i, k)) - mc(i, k + 1) * (shat(i, k + 1) - s(i, k))) / dp(i, k) -
                      rl / cp * du(i, k) * (beta * ql(i, k) + (1 - beta) * ql(i, k + 1));
        dqmdt(i, k) = (mu(i, k + 1) * (qu(i, k + 1) - qhat(i, k + 1) + cp / rl * (su(i, k + 1) - s(i, k))) -
                       mu(i, k) * (qu(i, k) - qhat(i, k) + cp / rl * (su(i, k) - s(i, k))) +
                       md(i, k + 1) * (qd(i, k + 1) - qhat(i, k + 1) + cp / rl * (sd(i, k + 1) - s(i, k))) -
                       md(i, k) * (qd(i, k) - qhat(i, k) + cp / rl * (sd(i, k) - s(i, k)))) / dp(i, k) +
                      du(i, k) * (beta * ql(i, k) + (1 - beta) * ql(i, k + 1));
      }

  for (int k = msg + 1; k <= pver; ++k)
    for (int i = il1g; i <= il2g; ++i)
      if (k >= lel(i) && k <= lcl(i)) {
        thetavp(i, k) = tp(i, k) * c_pow(1000.0 / p(i, k), rd / cp) * (1.0 + 1.608 * qstp(i, k) - q(i, mx(i)));
        thetavm(i, k) = t(i, k) * c_pow(1000.0 / p(i, k), rd / cp) * (1.0 + 0.608 * q(i, k));
        dqsdtp(i, k) = qstp(i, k) * (1.0 + qstp(i, k) / eps1) * eps1 * rl / (rd * (tp(i, k) * tp(i, k)));
        dtpdt(i, k) = tp(i, k) / (1.0 + rl / cp * (dqsdtp(i, k) - qstp(i, k) / tp(i, k))) *
                      (dtbdt(i) / t(i, mx(i)) +
                       rl / cp * (dqbdt(i) / tl(i) - q(i, mx(i)) / (tl(i) * tl(i)) * dtldt(i)));
        dboydt(i, k) = ((dtpdt(i, k) / tp(i, k) +
                         1.0 / (1.0 + 1.608 * qstp(i, k) - q(i, mx(i))) *
                             (1.608 * dqsdtp(i, k) * dtpdt(i, k) - dqbdt(i))) -
                        (dtmdt(i, k) / t(i, k) + 0.608 / (1.0 + 0.608 * q(i, k)) * dqmdt(i, k))) *
                       grav * thetavp(i, k) / thetavm(i, k);
      }
  for (int k = msg + 1; k <= pver; ++k)
    for (int i = il1g; i <= il2g; ++i)
      if (k > lcl(i) && k < mx(i)) {
        thetavp(i, k) = tp(i, k) * c_pow(1000.0 / p(i, k), rd / cp) * (1.0 + 0.608 * q(i, mx(i)));
        thetavm(i, k) = t(i, k) * c_pow(1000.0 / p(i, k), rd / cp) * (1.0 + 0.608 * q(i, k));
        dboydt(i, k) = (dtbdt(i) / t(i, mx(i)) + 0.608 / (1.0 + 0.608 * q(i, mx(i))) * dqbdt(i) -
                        dtmdt(i, k) / t(i, k) - 0.608 / (1.0 + 0.608 * q(i, k)) * dqmdt(i, k)) *
                       grav * thetavp(i, k) / thetavm(i, k);
      }

  for (int i = il1g; i <= il2g; ++i) dadt(i) = 0.0;
  kmin = lel(il1g); kmax = mx(il1g);
  for (int i = il1g; i <= il2g; ++i) { kmin = std::min(kmin, lel(i)); kmax = std::max(kmax, mx(i)); }
  kmax = kmax - 1;
  for (int k = kmin; k <= kmax; ++k)
    for (int i = il1g; i <= il2g; ++i)
      if (k >= lel(i) && k <= mx(i) - 1) dadt(i) = dadt(i) + dboydt(i, k) * (zf(i, k) - zf(i, k + 1));
  for (int i = il1g; i <= il2g; ++i) {
    dltaa = -1.0 * (cape(i) - capelmt);
    if (dadt(i) != 0.0) mb(i) = fmax2(dltaa / g.tau / dadt(i), 0.0);
  }
}

// ---- q1q2_pjr  zm_conv.F90:4262-4421 ------------------------------------------------------
void q1q2_pjr(int lchnk, A2 dqdt, A2 dsdt, C2 q, C2 qs, C2 qu, C2 su, C2 du, C2 qhat, C2 shat,
              C2 dp, C2 mu, C2 md, C2 sd, C2 qd, C2 ql, const double* dsubcld_, const int* jt_,
              const int* mx_, int il1g, int il2g, double cp, double rl, int msg, A2 dl, C2 evp,
              C2 cu) {
  (void)lchnk; (void)q; (void)qs;
  const int pver = g.pver;
  auto dsubcld = [&](int i) { return dsubcld_[i - 1]; };
  auto jt = [&](int i) { return jt_[i - 1]; };
  auto mx = [&](int i) { return mx_[i - 1]; };
  int kbm, ktm; double emc;
  for (int k = msg + 1; k <= pver; ++k)
    for (int i = il1g; i <= il2g; ++i) { dsdt(i, k) = 0.0; dqdt(i, k) = 0.0; dl(i, k) = 0.0; }
  ktm = pver; kbm = pver;
  for (int i = il1g; i <= il2g; ++i) { ktm = std::min(ktm, jt(i)); kbm = std::min(kbm, mx(i)); }
  for (int k = ktm; k <= pver - 1; ++k)
    for (int i = il1g; i <= il2g; ++i) {
      emc = -cu(i, k) + evp(i, k);
      dsdt(i, k) = -rl / cp * emc + (mu(i, k + 1) * (su(i, k + 1) - shat(i, k + 1)) -
                                     mu(i, k) * (su(i, k) - shat(i, k)) +
                                     md(i, k + 1) * (sd(i, k + 1) - shat(i, k + 1)) -
                                     md(i, k) * (sd(i, k) - shat(i, k))) / dp(i, k);
      dqdt(i, k) = emc + (mu(i, k + 1) * (qu(i, k + 1) - qhat(i, k + 1)) -
                          mu(i, k) * (qu(i, k) - qhat(i, k)) +
                          md(i, k + 1) * (qd(i, k + 1) - qhat(i, k + 1)) -
                          md(i, k) * (qd(i, k) - qhat(i, k))) / dp(i, k);
      dl(i, k) = du(i, k) * ql(i, k + 1);
    }
  for (int k = kbm; k <= pver; ++k)
    for (int i = il1g; i <= il2g; ++i) {
      if (k == mx(i)) {
        dsdt(i, k) = (1.0 / dsubcld(i)) * (-mu(i, k) * (su(i, k) - shat(i, k)) - md(i, k) * (sd(i, k) - shat(i, k)));
        dqdt(i, k) = (1.0 / dsubcld(i)) * (-mu(i, k) * (qu(i, k) - qhat(i, k)) - md(i, k) * (qd(i, k) - qhat(i, k)));
      } else if (k > mx(i)) {
        dsdt(i, k) = dsdt(i, k - 1);
        dqdt(i, k) = dqdt(i, k - 1);
      }
    }
}

// ---- zm_convr  zm_conv.F90:231-1709 -------------------------------------------------------
int convr(int lchnk, int ncol, const double* t_, const double* qh_, double* prec_, double* jctop_,
          double* jcbot_, const double* pblh_, const double* zm_, const double* geos_,
          const double* zi_, double* qtnd_, double* heat_, const double* pap_, const double* paph_,
          const double* dpp_, double delt, double* mcon_, double* cme_, double* cape_, double* eurt_,
          const double* tpert_, double* dlf_, double* pflx_, double* zdu_, double* rprd_,
          double* mu_, double* md_, double* du_, double* eu_, double* ed_, double* dp_,
          double* dsubcld_, int* jt_, int* maxg_, int* ideep_, int* lengath_, double* ql_,
          double* rliq_, const double* landfrac_, double* dif_, double* dnlf_, double* dnif_,
          double* rice_) {
  const int pcols = g.pcols, pver = g.pver, pverp = g.pverp;
  const double grav = g.grav, cpres = g.cpres, rgas = g.rgas, rl = g.rl, rgrav = g.rgrav,
               zvir = g.zvir, gravit = g.gravit;
  brent_fail = 0;
  C2 t{t_, pcols}, qh{qh_, pcols}, pap{pap_, pcols}, paph{paph_, pcols}, dpp{dpp_, pcols},
      zm{zm_, pcols}, zi{zi_, pcols};
  A2 qtnd{qtnd_, pcols}, heat{heat_, pcols}, mcon{mcon_, pcols}, dlf{dlf_, pcols},
      pflx{pflx_, pcols}, cme{cme_, pcols}, zdu{zdu_, pcols}, rprd{rprd_, pcols}, dif{dif_, pcols},
      dnlf{dnlf_, pcols}, dnif{dnif_, pcols}, mu{mu_, pcols}, eu{eu_, pcols}, eurt{eurt_, pcols},
      du{du_, pcols}, md{md_, pcols}, ed{ed_, pcols}, dp{dp_, pcols}, ql{ql_, pcols};
  auto geos = [&](int i) { return geos_[i - 1]; };
  auto pblh = [&](int i) { return pblh_[i - 1]; };
  auto landfrac = [&](int i) { return landfrac_[i - 1]; };
  auto cape = [&](int i) -> double& { return cape_[i - 1]; };
  auto dsubcld = [&](int i) -> double& { return dsubcld_[i - 1]; };
  auto jctop = [&](int i) -> double& { return jctop_[i - 1]; };
  auto jcbot = [&](int i) -> double& { return jcbot_[i - 1]; };
  auto prec = [&](int i) -> double& { return prec_[i - 1]; };
  auto rliq = [&](int i) -> double& { return rliq_[i - 1]; };
  auto rice = [&](int i) -> double& { return rice_[i - 1]; };
  auto ideep = [&](int i) -> int& { return ideep_[i - 1]; };
  auto jt = [&](int i) -> int& { return jt_[i - 1]; };
  auto maxg = [&](int i) -> int& { return maxg_[i - 1]; };
  int& lengath = *lengath_;

  if (g.zm_org) {
    // orgt(:,:) = 0 (zm_conv.F90:555-556); org2d = pressure-thickness weighted column mean of org over the
    // levels where org > 0 (zm_conv.F90:793-819)
    for (size_t e = 0; e < (size_t)pcols * pver; ++e) tl_orgt[e] = 0.0;
    for (int i = 1; i <= ncol; ++i) {
      double orgavg = 0.0, dptot = 0.0;
      for (int k = 1; k <= pver; ++k) {
        const double o = tl_org[(size_t)(i - 1) + (size_t)pcols * (k - 1)];
        if (o > 0) { orgavg = orgavg + dpp(i, k) * o; dptot = dptot + dpp(i, k); }
      }
      if (dptot > 0) orgavg = orgavg / dptot;
      for (int k = 1; k <= pver; ++k) tl_org2d[(size_t)(i - 1) + (size_t)pcols * (k - 1)] = orgavg;
    }
  }

  W1 cin(pcols), zs(pcols), mumax(pcols), pblt(pcols), tl(pcols), capeg(pcols), tlg(pcols),
      landfracg(pcols), mb(pcols), dmmx(pcols), dmsm(pcols), orgc(pcols);
  W2 dlg(pcols, pver), pflxg(pcols, pverp), cug(pcols, pver), evpg(pcols, pver);
  W2 q(pcols, pver), p(pcols, pver), z(pcols, pver), s(pcols, pver), tp(pcols, pver),
      zf(pcols, pverp), pf(pcols, pverp), qstp(pcols, pver);
  I1 lcl(pcols), lel(pcols), lon(pcols), maxi(pcols), lclg(pcols), lelg(pcols), indxd(pcols),
      jlcl(pcols), j0(pcols), jd(pcols);
  W2 qg(pcols, pver), tg(pcols, pver), pg(pcols, pver), zg(pcols, pver), sg(pcols, pver),
      tpg(pcols, pver), zfg(pcols, pverp), qstpg(pcols, pver), ug(pcols, pver), vg(pcols, pver),
      cmeg(pcols, pver), rprdg(pcols, pver);
  W2 dqdt(pcols, pver), dsdt(pcols, pver), sd(pcols, pver), qd(pcols, pver), mc(pcols, pver),
      qhat(pcols, pver), qu(pcols, pver), su(pcols, pver), qs(pcols, pver), shat(pcols, pver),
      hmn(pcols, pver), hsat(pcols, pver), qlg(pcols, pver);
  W2 dmpdz(pcols, pver), qldeg(pcols, pver);
  int msg; double qdifr, sdifr, hk;

  for (int k = 1; k <= pver; ++k) for (int i = 1; i <= pcols; ++i) dmpdz(i, k) = -g.tentrm;
  msg = g.limcnv - 1;

  for (int k = 1; k <= pver; ++k) for (int i = 1; i <= pcols; ++i) { qtnd(i, k) = 0.0; heat(i, k) = 0.0; }
  for (int k = 1; k <= pverp; ++k) for (int i = 1; i <= pcols; ++i) mcon(i, k) = 0.0;
  for (int i = 1; i <= ncol; ++i) { rliq(i) = 0.0; rice(i) = 0.0; }
  for (int i = 1; i <= ncol; ++i) prec(i) = 0.0;
  for (int k = 1; k <= pver; ++k)
    for (int i = 1; i <= ncol; ++i) {
      dqdt(i, k) = 0.0; dsdt(i, k) = 0.0; pflx(i, k) = 0.0; pflxg(i, k) = 0.0; cme(i, k) = 0.0;
      rprd(i, k) = 0.0; zdu(i, k) = 0.0; ql(i, k) = 0.0; qlg(i, k) = 0.0; dlf(i, k) = 0.0;
      dlg(i, k) = 0.0; qldeg(i, k) = 0.0; eurt(i, k) = 0.0; dif(i, k) = 0.0; dnlf(i, k) = 0.0;
      dnif(i, k) = 0.0;
    }
  for (int i = 1; i <= ncol; ++i) { pflx(i, pverp) = 0; pflxg(i, pverp) = 0; }
  for (int i = 1; i <= ncol; ++i) {
    pblt(i) = pver;
    dsubcld(i) = 0.0;
    jctop(i) = pver;
    jcbot(i) = 1;
  }
  for (int i = 1; i <= ncol; ++i) {
    zs(i) = geos(i) * rgrav;
    pf(i, pver + 1) = paph(i, pver + 1) * 0.01;
    zf(i, pver + 1) = zi(i, pver + 1) + zs(i);
  }
  for (int k = 1; k <= pver; ++k)
    for (int i = 1; i <= ncol; ++i) {
      p(i, k) = pap(i, k) * 0.01;
      pf(i, k) = paph(i, k) * 0.01;
      z(i, k) = zm(i, k) + zs(i);
      zf(i, k) = zi(i, k) + zs(i);
    }
  for (int k = pver - 1; k >= msg + 1; --k)
    for (int i = 1; i <= ncol; ++i)
      if (std::fabs(z(i, k) - zs(i) - pblh(i)) < (zf(i, k) - zf(i, k + 1)) * 0.5) pblt(i) = k;

  for (int k = 1; k <= pver; ++k)
    for (int i = 1; i <= ncol; ++i) {
      q(i, k) = qh(i, k);
      s(i, k) = t(i, k) + (grav / ((1.0 + zvir * q(i, k)) * cpres)) * z(i, k);
      tp(i, k) = 0.0;
      shat(i, k) = s(i, k);
      qhat(i, k) = q(i, k);
    }
  for (int i = 1; i <= ncol; ++i) {
    capeg(i) = 0.0; lclg(i) = 1; lelg(i) = pver; maxg(i) = 1; tlg(i) = 400.0; dsubcld(i) = 0.0;
  }

  if (g.cam3) {
    buoyan(lchnk, ncol, q, t, p, z, pf, tp, qstp, tl.v.data(), rl, cape_, pblt.v.data(),
           lcl.v.data(), lel.v.data(), lon.v.data(), maxi.v.data(), rgas, grav, cpres, msg, tpert_);
  } else {
    fl_phase = 1;
    buoyan_dilute(lchnk, ncol, q, t, p, z, pf, tp, qstp, tl.v.data(), cape_, cin.v.data(),
                  pblt.v.data(), lcl.v.data(), lel.v.data(), lon.v.data(), maxi.v.data(), rgas, grav,
                  cpres, msg, zi, zs.v.data(), tpert_, landfrac_, dmpdz);
    fl_phase = 0;
  }

  lengath = 0;
  for (int i = 1; i <= pcols; ++i) ideep(i) = 0;
  for (int i = 1; i <= ncol; ++i)
    if (cape(i) > g.capelmt)
      if (cin(i) < cape(i) * g.cin_threshd) {
        lengath = lengath + 1;
        ideep(lengath) = i;
        indxd(lengath) = i;
      }
  if (lengath == 0) return brent_fail;

  auto gather = [&]() {
    for (int k = 1; k <= pver; ++k)
      for (int i = 1; i <= lengath; ++i) {
        dp(i, k) = 0.01 * dpp(ideep(i), k);
        qg(i, k) = q(ideep(i), k);
        tg(i, k) = t(ideep(i), k);
        pg(i, k) = p(ideep(i), k);
        zg(i, k) = z(ideep(i), k);
        sg(i, k) = s(ideep(i), k);
        tpg(i, k) = tp(ideep(i), k);
        zfg(i, k) = zf(ideep(i), k);
        qstpg(i, k) = qstp(ideep(i), k);
        ug(i, k) = 0.0;
        vg(i, k) = 0.0;
      }
    for (int i = 1; i <= lengath; ++i) zfg(i, pver + 1) = zf(ideep(i), pver + 1);
  };
  auto hats = [&]() {
    for (int k = msg + 1; k <= pver; ++k)
      for (int i = 1; i <= lengath; ++i)
        if (k >= maxg(i)) dsubcld(i) = dsubcld(i) + dp(i, k);
    for (int k = msg + 2; k <= pver; ++k)
      for (int i = 1; i <= lengath; ++i) {
        sdifr = 0.0;
        qdifr = 0.0;
        if (sg(i, k) > 0.0 || sg(i, k - 1) > 0.0)
          sdifr = std::fabs((sg(i, k) - sg(i, k - 1)) / fmax2(sg(i, k - 1), sg(i, k)));
        if (qg(i, k) > 0.0 || qg(i, k - 1) > 0.0)
          qdifr = std::fabs((qg(i, k) - qg(i, k - 1)) / fmax2(qg(i, k - 1), qg(i, k)));
        if (sdifr > 1.E-6) {
          shat(i, k) = c_log(sg(i, k - 1) / sg(i, k)) * sg(i, k - 1) * sg(i, k) / (sg(i, k - 1) - sg(i, k));
        } else {
          shat(i, k) = 0.5 * (sg(i, k) + sg(i, k - 1));
        }
        if (qdifr > 1.E-6) {
          qhat(i, k) = c_log(qg(i, k - 1) / qg(i, k)) * qg(i, k - 1) * qg(i, k) / (qg(i, k - 1) - qg(i, k));
        } else {
          qhat(i, k) = 0.5 * (qg(i, k) + qg(i, k - 1));
        }
      }
  };

  gather();
  for (int i = 1; i <= lengath; ++i) {
    capeg(i) = cape(ideep(i));
    lclg(i) = lcl(ideep(i));
    lelg(i) = lel(ideep(i));
    maxg(i) = maxi(ideep(i));
    tlg(i) = tl(ideep(i));
    landfracg(i) = landfrac(ideep(i));
  }
  hats();

  cldprp(lchnk, qg, tg, ug, vg, pg, zg, sg, mu, eu, du, md, ed, sd, qd, mc, qu, su, zfg, qs, hmn,
         hsat, shat, qlg, cmeg, maxg_, lelg.v.data(), jt_, jlcl.v.data(), maxg_, j0.v.data(),
         jd.v.data(), rl, lengath, rgas, grav, cpres, msg, pflxg, evpg, cug, rprdg, g.limcnv,
         landfracg.v.data(), qldeg, qhat);

  // second call / retrigger (both .true.)  zm_conv.F90:1046-1228
  {
    for (int i = 1; i <= lengath; ++i) {
      hk = 0.0;
      for (int k = 1; k <= pver; ++k) dmpdz(ideep(i), k) = -1.0;
      dmmx(i) = 0.0;
      dmsm(i) = 0.0;
      orgc(i) = 1.0;
      for (int k = pver; k >= msg + 1; --k) {
        if (eu(i, k) > 0) {
          dmmx(i) = -fmax2(-dmmx(i), eu(i, k));
          dmsm(i) = dmsm(i) - eu(i, k);
          hk = hk + 1.0;
        }
      }
      if (hk > 0) {
        dmsm(i) = dmsm(i) / hk;
        double val = dmsm(i) * orgc(i) + dmmx(i) * (1.0 - orgc(i));
        for (int k = 1; k <= pver; ++k) dmpdz(ideep(i), k) = val;
      }
    }
    fl_phase = 2;
    buoyan_dilute(lchnk, ncol, q, t, p, z, pf, tp, qstp, tl.v.data(), cape_, cin.v.data(),
                  pblt.v.data(), lcl.v.data(), lel.v.data(), lon.v.data(), maxi.v.data(), rgas, grav,
                  cpres, msg, zi, zs.v.data(), tpert_, landfrac_, dmpdz);
    fl_phase = 0;

    lengath = 0;
    for (int i = 1; i <= pcols; ++i) { ideep(i) = 0; indxd(i) = 0; }
    for (int i = 1; i <= ncol; ++i)
      if (cape(i) > g.capelmt)
        if (cin(i) < cape(i) * g.cin_threshd) {
          lengath = lengath + 1;
          indxd(lengath) = i;
        }
    if (lengath == 0) return brent_fail;
    for (int ii = 1; ii <= lengath; ++ii) { int i = indxd(ii); ideep(ii) = i; }
    gather();
    for (int i = 1; i <= lengath; ++i) {
      capeg(i) = cape(ideep(i));
      lclg(i) = lcl(ideep(i));
      lelg(i) = lel(ideep(i));
      maxg(i) = maxi(ideep(i));
      tlg(i) = tl(ideep(i));
      landfracg(i) = landfrac(ideep(i));
      dsubcld(i) = 0.0;
    }
    hats();

    cldprp(lchnk, qg, tg, ug, vg, pg, zg, sg, mu, eu, du, md, ed, sd, qd, mc, qu, su, zfg, qs, hmn,
           hsat, shat, qlg, cmeg, maxg_, lelg.v.data(), jt_, jlcl.v.data(), maxg_, j0.v.data(),
           jd.v.data(), rl, lengath, rgas, grav, cpres, msg, pflxg, evpg, cug, rprdg, g.limcnv,
           landfracg.v.data(), qldeg, qhat);
  }

  for (int k = msg + 1; k <= pver; ++k)
    for (int i = 1; i <= lengath; ++i) eurt(ideep(i), k) = -dmpdz(ideep(i), k);

  for (int k = msg + 1; k <= pver; ++k)
    for (int i = 1; i <= lengath; ++i) {
      du(i, k) = du(i, k) * (zfg(i, k) - zfg(i, k + 1)) / dp(i, k);
      eu(i, k) = eu(i, k) * (zfg(i, k) - zfg(i, k + 1)) / dp(i, k);
      ed(i, k) = ed(i, k) * (zfg(i, k) - zfg(i, k + 1)) / dp(i, k);
      cug(i, k) = cug(i, k) * (zfg(i, k) - zfg(i, k + 1)) / dp(i, k);
      cmeg(i, k) = cmeg(i, k) * (zfg(i, k) - zfg(i, k + 1)) / dp(i, k);
      rprdg(i, k) = rprdg(i, k) * (zfg(i, k) - zfg(i, k + 1)) / dp(i, k);
      evpg(i, k) = evpg(i, k) * (zfg(i, k) - zfg(i, k + 1)) / dp(i, k);
    }

  closure(lchnk, qg, tg, pg, zg, sg, tpg, qs, qu, su, mc, du, mu, md, qd, sd, qhat, shat, dp, qstpg,
          zfg, qlg, dsubcld_, mb.v.data(), capeg.v.data(), tlg.v.data(), lclg.v.data(),
          lelg.v.data(), jt_, maxg_, 1, lengath, rgas, grav, cpres, rl, msg, g.capelmt);

  for (int i = 1; i <= lengath; ++i) mumax(i) = 0;
  for (int k = msg + 2; k <= pver; ++k)
    for (int i = 1; i <= lengath; ++i) mumax(i) = fmax2(mumax(i), mu(i, k) / dp(i, k));
  for (int i = 1; i <= lengath; ++i) {
    if (mumax(i) > 0.0) {
      mb(i) = fmin2(mb(i), 0.5 / (delt * mumax(i)));
    } else {
      mb(i) = 0.0;
    }
  }
  if (g.no_deep_pbl)
    for (int i = 1; i <= lengath; ++i)
      if (zm(ideep(i), jt(i)) < pblh(ideep(i))) mb(i) = 0;

  for (int k = msg + 1; k <= pver; ++k)
    for (int i = 1; i <= lengath; ++i) {
      mu(i, k) = mu(i, k) * mb(i);
      md(i, k) = md(i, k) * mb(i);
      mc(i, k) = mc(i, k) * mb(i);
      du(i, k) = du(i, k) * mb(i);
      eu(i, k) = eu(i, k) * mb(i);
      ed(i, k) = ed(i, k) * mb(i);
      cmeg(i, k) = cmeg(i, k) * mb(i);
      rprdg(i, k) = rprdg(i, k) * mb(i);
      cug(i, k) = cug(i, k) * mb(i);
      evpg(i, k) = evpg(i, k) * mb(i);
      pflxg(i, k + 1) = pflxg(i, k + 1) * mb(i) * 100.0 / grav;
    }

  q1q2_pjr(lchnk, dqdt, dsdt, qg, qs, qu, su, du, qhat, shat, dp, mu, md, sd, qd, qldeg, dsubcld_,
           jt_, maxg_, 1, lengath, cpres, rl, msg, dlg, evpg, cug);

  for (int k = msg + 1; k <= pver; ++k)
    for (int i = 1; i <= lengath; ++i) {
      q(ideep(i), k) = qh(ideep(i), k) + 2.0 * delt * dqdt(i, k);
      qtnd(ideep(i), k) = dqdt(i, k);
      cme(ideep(i), k) = cmeg(i, k);
      rprd(ideep(i), k) = rprdg(i, k);
      zdu(ideep(i), k) = du(i, k);
      mcon(ideep(i), k) = mc(i, k);
      heat(ideep(i), k) = dsdt(i, k) * cpres;
      dlf(ideep(i), k) = dlg(i, k);
      pflx(ideep(i), k) = pflxg(i, k);
      ql(ideep(i), k) = qlg(i, k);
    }
  for (int i = 1; i <= lengath; ++i) {
    jctop(ideep(i)) = jt(i);
    jcbot(ideep(i)) = maxg(i);
    pflx(ideep(i), pverp) = pflxg(i, pverp);
  }

  for (int k = pver; k >= msg + 1; --k)
    for (int i = 1; i <= ncol; ++i)
      prec(i) = prec(i) - dpp(i, k) * (q(i, k) - qh(i, k)) - dpp(i, k) * (dlf(i, k) + dif(i, k)) * 2.0 * delt;
  for (int i = 1; i <= ncol; ++i) prec(i) = rgrav * fmax2(prec(i), 0.0) / (2.0 * delt) / 1000.0;

  for (int k = 1; k <= pver; ++k)
    for (int i = 1; i <= ncol; ++i) {
      rliq(i) = rliq(i) + (dlf(i, k) + dif(i, k)) * dpp(i, k) / gravit;
      rice(i) = rice(i) + dif(i, k) * dpp(i, k) / gravit;
    }
  for (int i = 1; i <= ncol; ++i) { rliq(i) = rliq(i) / 1000.0; rice(i) = rice(i) / 1000.0; }
  return brent_fail;
}

}  // namespace

// ================================ C interface ==============================================
extern "C" {

const char* zmo_math_backend(void) { return zmo::math_backend; }

void zmo_params_default(zmo_params_t* p, int pcols, int pver, int limcnv) {
  std::memset(p, 0, sizeof(*p));
  p->pcols = pcols; p->pver = pver; p->limcnv = limcnv;
  p->num_cin = 1; p->zm_org = 0; p->microp = 0; p->no_deep_pbl = 0; p->lparcel_pbl = 0;
  p->cam3 = 0; p->masterproc = 1;
  p->c0_lnd = 0.0075; p->c0_ocn = 0.03; p->ke = 5.0e-6; p->ke_lnd = 1.0e-5;
  p->momcu = 0.7; p->momcd = 0.7; p->tiedke_add = 0.5; p->capelmt = 70.0; p->dmpdz = -1.0e-3;
  p->tau = 3600.0;
  zmo::PhysConst c = zmo::physconst_default();
  p->cpair = c.cpair; p->epsilo = c.epsilo; p->gravit = c.gravit; p->latice = c.latice;
  p->latvap = c.latvap; p->tmelt = c.tmelt; p->rair = c.rair; p->cpwv = c.cpwv; p->cpliq = c.cpliq;
  p->rh2o = c.rh2o; p->cpvir = c.cpvir; p->zvir = c.zvir;
}

// zm_convi  zm_conv.F90:115-227
int zmo_convi(const zmo_params_t* p) {
  if (p->microp) return 2;                   // zm_microphysics is not part of the reference tree
  g.pcols = p->pcols; g.pver = p->pver; g.pverp = p->pver + 1;
  g.cpair = p->cpair; g.epsilo = p->epsilo; g.gravit = p->gravit; g.latice = p->latice;
  g.latvap = p->latvap; g.tmelt = p->tmelt; g.rair = p->rair; g.cpwv = p->cpwv; g.cpliq = p->cpliq;
  g.rh2o = p->rh2o; g.cpvir = p->cpvir; g.zvir = p->zvir;
  g.limcnv = p->limcnv;
  g.tfreez = g.tmelt;
  g.eps1 = g.epsilo;
  g.rl = g.latvap;
  g.cpres = g.cpair;
  g.rgrav = 1.0 / g.gravit;
  g.rgas = g.rair;
  g.grav = g.gravit;
  g.cp = g.cpres;
  g.dcol = (g.cpliq - g.cpwv) / g.latvap;
  g.c0_lnd = p->c0_lnd; g.c0_ocn = p->c0_ocn; g.num_cin = p->num_cin; g.ke = p->ke;
  g.ke_lnd = p->ke_lnd; g.zm_org = p->zm_org != 0; g.momcu = p->momcu; g.momcd = p->momcd;
  g.zmconv_microp = false;
  g.tiedke_add = p->tiedke_add; g.capelmt = p->capelmt; g.dmpdz_param = p->dmpdz;
  g.no_deep_pbl = p->no_deep_pbl != 0; g.lparcel_pbl = p->lparcel_pbl != 0; g.tau = p->tau;
  g.cam3 = p->cam3 != 0;
  g.tentrm = 1e-3;
  if (p->masterproc) {
    if (g.num_cin > 5) return 1;             // endrun: NUM_CIN must not exceed 5
    g.tentrm = -g.dmpdz_param;               // zm_conv.F90:213 (inside `if (masterproc)`)
  }
  g.estbl.build(g.tmelt);
  return 0;
}

int zmo_convr(int lchnk, int ncol, const double* t, const double* qh, double* prec, double* jctop,
              double* jcbot, const double* pblh, const double* zm, const double* geos,
              const double* zi, double* qtnd, double* heat, const double* pap, const double* paph,
              const double* dpp, double delt, double* mcon, double* cme, double* cape, double* eurt,
              const double* tpert, double* dlf, double* pflx, double* zdu, double* rprd, double* mu,
              double* md, double* du, double* eu, double* ed, double* dp, double* dsubcld, int* jt,
              int* maxg, int* ideep, int* lengath, double* ql, double* rliq, const double* landfrac,
              double* dif, double* dnlf, double* dnif, double* rice) {
  return convr(lchnk, ncol, t, qh, prec, jctop, jcbot, pblh, zm, geos, zi, qtnd, heat, pap, paph, dpp,
               delt, mcon, cme, cape, eurt, tpert, dlf, pflx, zdu, rprd, mu, md, du, eu, ed, dp,
               dsubcld, jt, maxg, ideep, lengath, ql, rliq, landfrac, dif, dnlf, dnif, rice);
}

int zmo_buoyan_dilute(int lchnk, int ncol, const double* q, const double* t, const double* p,
                      const double* z, const double* pf, double* tp, double* qstp, double* tl,
                      double* cape, double* cin, const double* pblt, int* lcl, int* lel, int* lon,
                      int* mx, const double* zi, const double* zs, const double* tpert,
                      const double* landfrac, const double* dmpdz) {
  brent_fail = 0;
  const int pc = g.pcols;
  buoyan_dilute(lchnk, ncol, C2{q, pc}, C2{t, pc}, C2{p, pc}, C2{z, pc}, C2{pf, pc}, A2{tp, pc},
                A2{qstp, pc}, tl, cape, cin, pblt, lcl, lel, lon, mx, g.rgas, g.grav, g.cpres,
                g.limcnv - 1, C2{zi, pc}, zs, tpert, landfrac, C2{dmpdz, pc});
  return brent_fail;
}

// ---- zm_conv_evap  zm_conv.F90:1712-1972 (old_snow = .true.) -------------------------------
void zmo_conv_evap(int ncol, int lchnk, const double* t_, const double* pmid_, const double* pdel_,
                   const double* q_, const double* landfrac_, double* tend_s_, double* tend_s_snwprd_,
                   double* tend_s_snwevmlt_, double* tend_q_, const double* prdprec_,
                   const double* cldfrc_, double deltat, double* prec_, double* snow_,
                   double* ntprprd_, double* ntsnprd_, double* flxprec_, double* flxsnow_) {
  (void)lchnk; (void)landfrac_; (void)deltat;
  const int pcols = g.pcols, pver = g.pver;
  const double tmelt = g.tmelt, gravit = g.gravit, latice = g.latice, latvap = g.latvap;
  C2 t{t_, pcols}, pmid{pmid_, pcols}, pdel{pdel_, pcols}, q{q_, pcols}, prdprec{prdprec_, pcols},
      cldfrc{cldfrc_, pcols};
  A2 tend_s{tend_s_, pcols}, tend_q{tend_q_, pcols}, tend_s_snwprd{tend_s_snwprd_, pcols},
      tend_s_snwevmlt{tend_s_snwevmlt_, pcols}, flxprec{flxprec_, pcols}, flxsnow{flxsnow_, pcols},
      ntprprd{ntprprd_, pcols}, ntsnprd{ntsnprd_, pcols};
  auto prec = [&](int i) -> double& { return prec_[i - 1]; };
  auto snow = [&](int i) -> double& { return snow_[i - 1]; };
  W2 es(pcols, pver), fice(pcols, pver), fsnow_conv(pcols, pver), qs(pcols, pver);
  W1 evpvint(pcols), evpprec(pcols), evpsnow(pcols), snowmlt(pcols), flxsntm(pcols);
  double work1, work2, kemask, evplimit;

  for (int i = 1; i <= ncol; ++i) prec(i) = prec(i) * 1000.0;
  for (int k = 1; k <= pver; ++k)
    for (int i = 1; i <= ncol; ++i) g.estbl.qsat(t(i, k), pmid(i, k), g.epsilo, es(i, k), qs(i, k));
  for (int k = 1; k <= pver; ++k)
    for (int i = 1; i <= ncol; ++i) zmo::cldfrc_fice(t(i, k), tmelt, fice(i, k), fsnow_conv(i, k));
  for (int i = 1; i <= ncol; ++i) { flxprec(i, 1) = 0.0; flxsnow(i, 1) = 0.0; evpvint(i) = 0.0; }

  for (int k = 1; k <= pver; ++k)
    for (int i = 1; i <= ncol; ++i) {
      if (t(i, k) > tmelt) {
        flxsntm(i) = 0.0;
        snowmlt(i) = flxsnow(i, k) * gravit / pdel(i, k);
      } else {
        flxsntm(i) = flxsnow(i, k);
        snowmlt(i) = 0.0;
      }
      evplimit = fmax2(1.0 - q(i, k) / (1.0 + q(i, k)) / qs(i, k), 0.0);
      if (g.zm_org) kemask = g.ke * (1.0 - landfrac_[i - 1]) + g.ke_lnd * landfrac_[i - 1];   // zm_conv.F90:1860-1864
      else kemask = g.ke;
      evpprec(i) = kemask * (1.0 - cldfrc(i, k)) * evplimit * std::sqrt(flxprec(i, k));
      // tht_tweaks: the second evplimit assignment is commented out (zm_conv.F90:1875-1877)
      evplimit = fmin2(evplimit, flxprec(i, k) * gravit / pdel(i, k));
      evplimit = fmin2(evplimit, (prec(i) - evpvint(i)) * gravit / pdel(i, k));
      evpprec(i) = fmin2(evplimit, evpprec(i));
      if (flxprec(i, k) > 0.0) {
        work1 = fmin2(fmax2(0.0, flxsntm(i) / flxprec(i, k)), 1.0);
        evpsnow(i) = evpprec(i) * work1;
      } else {
        evpsnow(i) = 0.0;
      }
      evpvint(i) = evpvint(i) + evpprec(i) * pdel(i, k) / gravit;
      ntprprd(i, k) = prdprec(i, k) - evpprec(i);
      if (flxprec(i, k) > 0.0) {
        work1 = fmin2(fmax2(0.0, flxsnow(i, k) / flxprec(i, k)), 1.0);
      } else {
        work1 = 0.0;
      }
      work2 = fmax2(fsnow_conv(i, k), work1);
      if (snowmlt(i) > 0.0) work2 = 0.0;
      ntsnprd(i, k) = prdprec(i, k) * work2 - evpsnow(i) - snowmlt(i);
      tend_s_snwprd(i, k) = prdprec(i, k) * work2 * latice;
      tend_s_snwevmlt(i, k) = -(evpsnow(i) + snowmlt(i)) * latice;
      flxprec(i, k + 1) = flxprec(i, k) + ntprprd(i, k) * pdel(i, k) / gravit;
      flxsnow(i, k + 1) = flxsnow(i, k) + ntsnprd(i, k) * pdel(i, k) / gravit;
      flxprec(i, k + 1) = fmax2(flxprec(i, k + 1), 0.0);
      flxsnow(i, k + 1) = fmax2(flxsnow(i, k + 1), 0.0);
      tend_s(i, k) = -evpprec(i) * latvap + ntsnprd(i, k) * latice;
      tend_q(i, k) = evpprec(i);
    }
  for (int i = 1; i <= ncol; ++i) {
    prec(i) = flxprec(i, pver + 1) / 1000.0;
    snow(i) = flxsnow(i, pver + 1) / 1000.0;
  }
}

// ---- convtran  zm_conv.F90:1976-2311 -------------------------------------------------------
void zmo_convtran(int lchnk, const int* doconvtran, const double* q_, int ncnst, const double* mu_,
                  const double* md_, const double* du_, const double* eu_, const double* ed_,
                  const double* dp_, const double* dsubcld_, const int* jt_, const int* mx_,
                  const int* ideep_, int il1g, int il2g, int nstep, const double* fracis_,
                  double* dqdt_, const double* dpdry_, double dt, const int* cnst_is_dry) {
  (void)lchnk; (void)dsubcld_; (void)nstep; (void)dt;
  const int pcols = g.pcols, pver = g.pver;
  C3 q{q_, pcols, pver}, fracis{fracis_, pcols, pver};
  A3 dqdt{dqdt_, pcols, pver};
  C2 mu{mu_, pcols}, md{md_, pcols}, du{du_, pcols}, eu{eu_, pcols}, ed{ed_, pcols}, dp{dp_, pcols},
      dpdry{dpdry_, pcols};
  auto jt = [&](int i) { return jt_[i - 1]; };
  auto mx = [&](int i) { return mx_[i - 1]; };
  auto ideep = [&](int i) { return ideep_[i - 1]; };
  int kbm, kk, kkp1, km1, kp1, ktm, k;
  double cabv, cbel, cdifr, small, mbsth, mupdudp, minc, maxc, fluxin, fluxout, netflux;
  W2 chat(pcols, pver), cond(pcols, pver), cnst(pcols, pver), fisg(pcols, pver), conu(pcols, pver),
      dcondt(pcols, pver), dutmp(pcols, pver), eutmp(pcols, pver), edtmp(pcols, pver), dptmp(pcols, pver);

  small = 1.e-36;
  mbsth = 1.e-15;
  ktm = pver; kbm = pver;
  for (int i = il1g; i <= il2g; ++i) { ktm = std::min(ktm, jt(i)); kbm = std::min(kbm, mx(i)); }

  for (int m = 2; m <= ncnst; ++m) {
    if (!doconvtran[m - 1]) continue;
    if (cnst_is_dry[m - 1]) {
      for (k = 1; k <= pver; ++k)
        for (int i = il1g; i <= il2g; ++i) {
          dptmp(i, k) = dpdry(i, k);
          dutmp(i, k) = du(i, k) * dp(i, k) / dpdry(i, k);
          eutmp(i, k) = eu(i, k) * dp(i, k) / dpdry(i, k);
          edtmp(i, k) = ed(i, k) * dp(i, k) / dpdry(i, k);
        }
    } else {
      for (k = 1; k <= pver; ++k)
        for (int i = il1g; i <= il2g; ++i) {
          dptmp(i, k) = dp(i, k); dutmp(i, k) = du(i, k); eutmp(i, k) = eu(i, k); edtmp(i, k) = ed(i, k);
        }
    }
    for (k = 1; k <= pver; ++k)
      for (int i = il1g; i <= il2g; ++i) {
        cnst(i, k) = q(ideep(i), k, m);
        fisg(i, k) = fracis(ideep(i), k, m);
      }
    for (k = 1; k <= pver; ++k) {
      km1 = std::max(1, k - 1);
      for (int i = il1g; i <= il2g; ++i) {
        minc = fmin2(cnst(i, km1), cnst(i, k));
        maxc = fmax2(cnst(i, km1), cnst(i, k));
        if (minc < 0) {
          cdifr = 0.0;
        } else {
          cdifr = std::fabs(cnst(i, k) - cnst(i, km1)) / fmax2(maxc, small);
        }
        if (cdifr > 1.E-6) {
          cabv = fmax2(cnst(i, km1), maxc * 1.e-12);
          cbel = fmax2(cnst(i, k), maxc * 1.e-12);
          chat(i, k) = c_log(cabv / cbel) / (cabv - cbel) * cabv * cbel;
        } else {
          chat(i, k) = 0.5 * (cnst(i, k) + cnst(i, km1));
        }
        conu(i, k) = chat(i, k);
        cond(i, k) = chat(i, k);
        dcondt(i, k) = 0.0;
      }
    }
    k = 2; km1 = 1; kk = pver;
    for (int i = il1g; i <= il2g; ++i) {
      mupdudp = mu(i, kk) + dutmp(i, kk) * dptmp(i, kk);
      if (mupdudp > mbsth) conu(i, kk) = (+eutmp(i, kk) * fisg(i, kk) * cnst(i, kk) * dptmp(i, kk)) / mupdudp;
      if (md(i, k) < -mbsth) cond(i, k) = (-edtmp(i, km1) * fisg(i, km1) * cnst(i, km1) * dptmp(i, km1)) / md(i, k);
    }
    for (kk = pver - 1; kk >= 1; --kk) {
      kkp1 = std::min(pver, kk + 1);
      for (int i = il1g; i <= il2g; ++i) {
        mupdudp = mu(i, kk) + dutmp(i, kk) * dptmp(i, kk);
        if (mupdudp > mbsth)
          conu(i, kk) = (mu(i, kkp1) * conu(i, kkp1) + eutmp(i, kk) * fisg(i, kk) * cnst(i, kk) * dptmp(i, kk)) / mupdudp;
      }
    }
    for (k = 3; k <= pver; ++k) {
      km1 = std::max(1, k - 1);
      for (int i = il1g; i <= il2g; ++i)
        if (md(i, k) < -mbsth)
          cond(i, k) = (md(i, km1) * cond(i, km1) - edtmp(i, km1) * fisg(i, km1) * cnst(i, km1) * dptmp(i, km1)) / md(i, k);
    }
    for (k = ktm; k <= pver; ++k) {
      km1 = std::max(1, k - 1);
      kp1 = std::min(pver, k + 1);
      for (int i = il1g; i <= il2g; ++i) {
        fluxin = mu(i, kp1) * conu(i, kp1) + mu(i, k) * fmin2(chat(i, k), cnst(i, km1)) -
                 (md(i, k) * cond(i, k) + md(i, kp1) * fmin2(chat(i, kp1), cnst(i, kp1)));
        fluxout = mu(i, k) * conu(i, k) + mu(i, kp1) * fmin2(chat(i, kp1), cnst(i, k)) -
                  (md(i, kp1) * cond(i, kp1) + md(i, k) * fmin2(chat(i, k), cnst(i, k)));
        netflux = fluxin - fluxout;
        if (std::fabs(netflux) < fmax2(fluxin, fluxout) * 1.e-12) netflux = 0.0;
        dcondt(i, k) = netflux / dptmp(i, k);
      }
    }
    for (k = kbm; k <= pver; ++k) {
      km1 = std::max(1, k - 1);
      for (int i = il1g; i <= il2g; ++i) {
        if (k == mx(i)) {
          fluxin = mu(i, k) * fmin2(chat(i, k), cnst(i, km1)) - md(i, k) * cond(i, k);
          fluxout = mu(i, k) * conu(i, k) - md(i, k) * fmin2(chat(i, k), cnst(i, k));
          netflux = fluxin - fluxout;
          if (std::fabs(netflux) < fmax2(fluxin, fluxout) * 1.e-12) netflux = 0.0;
          dcondt(i, k) = netflux / dptmp(i, k);
        } else if (k > mx(i)) {
          dcondt(i, k) = 0.0;
        }
      }
    }
    for (k = 1; k <= pver; ++k) for (int i = 1; i <= pcols; ++i) dqdt(i, k, m) = 0.0;
    for (k = 1; k <= pver; ++k)
      for (int i = il1g; i <= il2g; ++i) dqdt(ideep(i), k, m) = dcondt(i, k);
  }
}

// ---- momtran  zm_conv.F90:2315-2715 --------------------------------------------------------
void zmo_momtran(int lchnk, int ncol, const int* domomtran, const double* q_, int ncnst,
                 const double* mu_, const double* md_, const double* du_, const double* eu_,
                 const double* ed_, const double* dp_, const double* dsubcld_, const int* jt_,
                 const int* mx_, const int* ideep_, int il1g, int il2g, int nstep, double* dqdt_,
                 double* pguall_, double* pgdall_, double* icwu_, double* icwd_, double dt,
                 double* seten_) {
  (void)lchnk; (void)dsubcld_; (void)nstep;
  const int pcols = g.pcols, pver = g.pver, pverp = g.pverp;
  C3 q{q_, pcols, pver};
  A3 dqdt{dqdt_, pcols, pver}, pguall{pguall_, pcols, pver}, pgdall{pgdall_, pcols, pver},
      icwu{icwu_, pcols, pver}, icwd{icwd_, pcols, pver};
  A2 seten{seten_, pcols};
  C2 mu{mu_, pcols}, md{md_, pcols}, du{du_, pcols}, eu{eu_, pcols}, ed{ed_, pcols}, dp{dp_, pcols};
  auto jt = [&](int i) { return jt_[i - 1]; };
  auto mx = [&](int i) { return mx_[i - 1]; };
  auto ideep = [&](int i) { return ideep_[i - 1]; };
  int k, kbm, kk, kkp1, km1, kp1, ktm, ii;
  double mbsth, mupdudp;
  W2 chat(pcols, pver), cond(pcols, pver), cnst(pcols, pver), conu(pcols, pver), dcondt(pcols, pver),
      mududp(pcols, pver), mddudp(pcols, pver), pgu(pcols, pver), pgd(pcols, pver), gseten(pcols, pver);
  std::vector<double> mflux_((size_t)pcols * pverp * ncnst, 0.0), wind0_((size_t)pcols * pver * ncnst, 0.0),
      windf_((size_t)pcols * pver * ncnst, 0.0);
  A3 mflux{mflux_.data(), pcols, pverp}, wind0{wind0_.data(), pcols, pver}, windf{windf_.data(), pcols, pver};
  double fkeb, fket, ketend_cons, ketend, utop, ubot, vtop, vbot, gset2;

  for (int m = 1; m <= ncnst; ++m)
    for (k = 1; k <= pver; ++k)
      for (int i = 1; i <= pcols; ++i) { pguall(i, k, m) = 0.0; pgdall(i, k, m) = 0.0; }
  for (int m = 1; m <= ncnst; ++m)
    for (k = 1; k <= pver; ++k)
      for (int i = 1; i <= ncol; ++i) { icwu(i, k, m) = q(i, k, m); icwd(i, k, m) = q(i, k, m); }
  for (k = 1; k <= pver; ++k) for (int i = 1; i <= pcols; ++i) { seten(i, k) = 0.0; gseten(i, k) = 0.0; }
  mbsth = 1.e-15;
  ktm = pver; kbm = pver;
  for (int i = il1g; i <= il2g; ++i) { ktm = std::min(ktm, jt(i)); kbm = std::min(kbm, mx(i)); }

  for (int m = 1; m <= ncnst; ++m) {
    if (!domomtran[m - 1]) continue;
    for (k = 1; k <= pver; ++k)
      for (int i = il1g; i <= il2g; ++i) { cnst(i, k) = q(ideep(i), k, m); wind0(i, k, m) = cnst(i, k); }
    for (k = 1; k <= pver; ++k) {
      km1 = std::max(1, k - 1);
      for (int i = il1g; i <= il2g; ++i) {
        chat(i, k) = 0.5 * (cnst(i, k) + cnst(i, km1));
        conu(i, k) = chat(i, k);
        cond(i, k) = chat(i, k);
        dcondt(i, k) = 0.0;
      }
    }
    k = 1;
    for (int i = 1; i <= il2g; ++i) { pgu(i, k) = 0.0; pgd(i, k) = 0.0; }
    for (k = 2; k <= pver - 1; ++k) {
      km1 = std::max(1, k - 1);
      kp1 = std::min(pver, k + 1);
      for (int i = il1g; i <= il2g; ++i) {
        mududp(i, k) = (mu(i, k) * (cnst(i, k) - cnst(i, km1)) / dp(i, km1) +
                        mu(i, kp1) * (cnst(i, kp1) - cnst(i, k)) / dp(i, k));
        pgu(i, k) = -g.momcu * 0.5 * mududp(i, k);
        mddudp(i, k) = (md(i, k) * (cnst(i, k) - cnst(i, km1)) / dp(i, km1) +
                        md(i, kp1) * (cnst(i, kp1) - cnst(i, k)) / dp(i, k));
        pgd(i, k) = -g.momcd * 0.5 * mddudp(i, k);
      }
    }
    k = pver;
    km1 = std::max(1, k - 1);
    for (int i = il1g; i <= il2g; ++i) {
      mududp(i, k) = mu(i, k) * (cnst(i, k) - cnst(i, km1)) / dp(i, km1);
      pgu(i, k) = -g.momcu * mududp(i, k);
      mddudp(i, k) = md(i, k) * (cnst(i, k) - cnst(i, km1)) / dp(i, km1);
      pgd(i, k) = -g.momcd * mddudp(i, k);
    }
    k = 2; km1 = 1; kk = pver;
    for (int i = il1g; i <= il2g; ++i) {
      mupdudp = mu(i, kk) + du(i, kk) * dp(i, kk);
      if (mupdudp > mbsth)
        conu(i, kk) = (+eu(i, kk) * cnst(i, kk) * dp(i, kk) + pgu(i, kk) * dp(i, kk)) / mupdudp;
      if (md(i, k) < -mbsth)   // precedence as written in the reference, zm_conv.F90:2554
        cond(i, k) = (-ed(i, km1) * cnst(i, km1) * dp(i, km1)) - pgd(i, km1) * dp(i, km1) / md(i, k);
    }
    for (kk = pver - 1; kk >= 1; --kk) {
      kkp1 = std::min(pver, kk + 1);
      for (int i = il1g; i <= il2g; ++i) {
        mupdudp = mu(i, kk) + du(i, kk) * dp(i, kk);
        if (mupdudp > mbsth)
          conu(i, kk) = (mu(i, kkp1) * conu(i, kkp1) + eu(i, kk) * cnst(i, kk) * dp(i, kk) + pgu(i, kk) * dp(i, kk)) / mupdudp;
      }
    }
    for (k = 3; k <= pver; ++k) {
      km1 = std::max(1, k - 1);
      for (int i = il1g; i <= il2g; ++i)
        if (md(i, k) < -mbsth)
          cond(i, k) = (md(i, km1) * cond(i, km1) - ed(i, km1) * cnst(i, km1) * dp(i, km1) - pgd(i, km1) * dp(i, km1)) / md(i, k);
    }
    for (k = ktm; k <= pver; ++k) {
      km1 = std::max(1, k - 1);
      kp1 = std::min(pver, k + 1);
      for (int i = il1g; i <= il2g; ++i)
        dcondt(i, k) = +(mu(i, kp1) * (conu(i, kp1) - chat(i, kp1)) - mu(i, k) * (conu(i, k) - chat(i, k)) +
                         md(i, kp1) * (cond(i, kp1) - chat(i, kp1)) - md(i, k) * (cond(i, k) - chat(i, k))) / dp(i, k);
    }
    for (k = kbm; k <= pver; ++k)
      for (int i = il1g; i <= il2g; ++i)
        if (k == mx(i))
          dcondt(i, k) = (1.0 / dp(i, k)) * (-mu(i, k) * (conu(i, k) - chat(i, k)) - md(i, k) * (cond(i, k) - chat(i, k)));

    for (k = 1; k <= pver; ++k) for (int i = 1; i <= pcols; ++i) dqdt(i, k, m) = 0.0;
    for (k = 1; k <= pver; ++k)
      for (int i = il1g; i <= il2g; ++i) {
        ii = ideep(i);
        dqdt(ii, k, m) = dcondt(i, k);
        pguall(ii, k, m) = -pgu(i, k);
        pgdall(ii, k, m) = -pgd(i, k);
        icwu(ii, k, m) = conu(i, k);
        icwd(ii, k, m) = cond(i, k);
      }
    for (k = ktm; k <= pver; ++k)
      for (int i = il1g; i <= il2g; ++i)
        mflux(i, k, m) = -mu(i, k) * (conu(i, k) - chat(i, k)) - md(i, k) * (cond(i, k) - chat(i, k));
    for (k = ktm; k <= pver; ++k)
      for (int i = il1g; i <= il2g; ++i) {
        kp1 = k + 1;
        windf(i, k, m) = cnst(i, k) - (mflux(i, kp1, m) - mflux(i, k, m)) * dt / dp(i, k);
      }
  }

  // KE dissipation heating hard-codes components 1 and 2 (zm_conv.F90:2684-2695)
  for (k = ktm; k <= pver; ++k) {
    km1 = std::max(1, k - 1);
    kp1 = std::min(pver, k + 1);
    for (int i = il1g; i <= il2g; ++i) {
      utop = (wind0(i, k, 1) + wind0(i, km1, 1)) / 2.0;
      vtop = (wind0(i, k, 2) + wind0(i, km1, 2)) / 2.0;
      ubot = (wind0(i, kp1, 1) + wind0(i, k, 1)) / 2.0;
      vbot = (wind0(i, kp1, 2) + wind0(i, k, 2)) / 2.0;
      fket = utop * mflux(i, k, 1) + vtop * mflux(i, k, 2);
      fkeb = ubot * mflux(i, k + 1, 1) + vbot * mflux(i, k + 1, 2);
      ketend_cons = (fket - fkeb) / dp(i, k);
      ketend = ((windf(i, k, 1) * windf(i, k, 1) + windf(i, k, 2) * windf(i, k, 2)) -
                (wind0(i, k, 1) * wind0(i, k, 1) + wind0(i, k, 2) * wind0(i, k, 2))) * 0.5 / dt;
      gset2 = ketend_cons - ketend;
      gseten(i, k) = gset2;
    }
  }
  for (k = 1; k <= pver; ++k)
    for (int i = il1g; i <= il2g; ++i) { ii = ideep(i); seten(ii, k) = gseten(i, k); }
}

double zmo_entropy(double tk, double p, double qtot) { return entropy(tk, p, qtot); }
double zmo_enthalpy(double tk, double p, double qtot, double z) { return enthalpy(tk, p, qtot, z); }
int zmo_ientropy(double s, double p, double qt, double tfg, double* t, double* qst) {
  brent_fail = 0; invert<0>(0, 1, 0, s, p, 0.0, qt, *t, *qst, tfg); return brent_fail;
}
int zmo_ienthalpy(double s, double p, double z, double qt, double tfg, double* t, double* qst) {
  brent_fail = 0; invert<1>(0, 1, 0, s, p, z, qt, *t, *qst, tfg); return brent_fail;
}
void zmo_qsat_hpa(double t, double p, double* es, double* qm) { qsat_hPa(t, p, *es, *qm); }
void zmo_qsat_table(double t, double p, double* es, double* qs) { g.estbl.qsat(t, p, g.epsilo, *es, *qs); }

// zm_org: attach the org/orgt/org2d fields (pointer dummies of zm_convr, zm_conv.F90:421-423).  For the
// single-chunk entry points they address one chunk; for the *_batch drivers the whole batch [chunk][k][i].
void zmo_org_fields(const double* org, double* orgt, double* org2d) {
  tl_org = org; tl_orgt = orgt; tl_org2d = org2d;
  batch_org = org; batch_orgt = orgt; batch_org2d = org2d;
}
void zmo_trace_set(int* buf, int cap) { trace_buf = buf; trace_cap = cap; trace_n = 0; }
int zmo_trace_count(void) { return trace_n / 4; }
void zmo_counters_reset(void) {
  for (int i = 0; i < 10; ++i) cnt[i] = 0;
  for (auto& ph : fl) for (auto& v : ph) v = 0;
}
void zmo_flops_get(long long* out24) {
  for (int ph = 0; ph < 3; ++ph) for (int j = 0; j < 8; ++j) out24[ph * 8 + j] = fl[ph][j];
}
void zmo_counters_get(long long* out10) { for (int i = 0; i < 10; ++i) out10[i] = cnt[i]; }

int zmo_convr_batch(int nchunks, const int* ncol, const double* t, const double* qh, double* prec,
                    double* jctop, double* jcbot, const double* pblh, const double* zm,
                    const double* geos, const double* zi, double* qtnd, double* heat,
                    const double* pap, const double* paph, const double* dpp, double delt,
                    double* mcon, double* cme, double* cape, double* eurt, const double* tpert,
                    double* dlf, double* pflx, double* zdu, double* rprd, double* mu, double* md,
                    double* du, double* eu, double* ed, double* dp, double* dsubcld, int* jt,
                    int* maxg, int* ideep, int* lengath, double* ql, double* rliq,
                    const double* landfrac, double* dif, double* dnlf, double* dnif, double* rice,
                    int nthreads) {
  const size_t pc = g.pcols, L = (size_t)g.pcols * g.pver, Lp = (size_t)g.pcols * g.pverp;
  int fails = 0;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#else
  (void)nthreads;
#endif
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : fails)
  for (int c = 0; c < nchunks; ++c) {
    if (g.zm_org) { tl_org = batch_org + (size_t)c * L; tl_orgt = batch_orgt + (size_t)c * L; tl_org2d = batch_org2d + (size_t)c * L; }
    fails += convr(c + 1, ncol[c], t + c * L, qh + c * L, prec + c * pc, jctop + c * pc, jcbot + c * pc,
                   pblh + c * pc, zm + c * L, geos + c * pc, zi + c * Lp, qtnd + c * L, heat + c * L,
                   pap + c * L, paph + c * Lp, dpp + c * L, delt, mcon + c * Lp, cme + c * L,
                   cape + c * pc, eurt + c * L, tpert + c * pc, dlf + c * L, pflx + c * Lp, zdu + c * L,
                   rprd + c * L, mu + c * L, md + c * L, du + c * L, eu + c * L, ed + c * L, dp + c * L,
                   dsubcld + c * pc, jt + c * pc, maxg + c * pc, ideep + c * pc, lengath + c,
                   ql + c * L, rliq + c * pc, landfrac + c * pc, dif + c * L, dnlf + c * L,
                   dnif + c * L, rice + c * pc);
  }
  return fails;
}

// ---- zm_conv_tend sequence (zm_conv_intr.F90:662-836), one chunk ------------------------------
// zm_convr(delt=0.5*ztodt) -> mcon unit conversion (:693) -> physics_update(state1) [t, q(:,:,1) with
// the external qneg3 clip at qmin(1)=1e-12; physics_types.F90:322-329,427] -> zm_conv_evap (:764) ->
// momtran (:822) -> ptend_all = physics_ptend_sum of the three ptend_loc (:736,803,833).
static int conv_tend_chunk(int lchnk, int ncol, const double* t, const double* q, const double* u,
                           const double* v, const double* pmid, const double* pint, const double* pdel,
                           const double* zm, const double* zi, const double* phis, const double* pblh,
                           const double* tpert, const double* landfrac, const double* cld, double ztodt,
                           double* ptend_s, double* ptend_q, double* ptend_u, double* ptend_v, double* mcon,
                           double* cme, double* pflx, double* zdu, double* rliq, double* rice, double* jctop,
                           double* jcbot, double* prec, double* snow, double* ql, double* rprd,
                           double* evapcdp, double* flxprec, double* flxsnow, double* dlf, double* mu,
                           double* md, double* du, double* eu, double* ed, double* dp, double* dsubcld,
                           int* jt, int* maxg, int* ideep, int* lengath, double* cape) {
  const int pcols = g.pcols, pver = g.pver, pverp = g.pverp;
  const size_t n2 = (size_t)pcols * pver;
  std::vector<double> heat(n2), qtnd(n2), eurt(n2), dif(n2), dnlf(n2), dnif(n2), t1(n2), q1(n2), ev_s(n2),
      ev_q(n2), snwprd(n2), snwevmlt(n2), ntprprd(n2), ntsnprd(n2), seten(n2), winds(2 * n2), wtend(2 * n2),
      pgu(2 * n2), pgd(2 * n2), icwu(2 * n2), icwd(2 * n2);
  int fails = convr(lchnk, ncol, t, q, prec, jctop, jcbot, pblh, zm, phis, zi, qtnd.data(), heat.data(), pmid,
                    pint, pdel, 0.5 * ztodt, mcon, cme, cape, eurt.data(), tpert, dlf, pflx, zdu, rprd, mu, md,
                    du, eu, ed, dp, dsubcld, jt, maxg, ideep, lengath, ql, rliq, landfrac, dif.data(),
                    dnlf.data(), dnif.data(), rice);
  for (int k = 1; k <= pverp; ++k)
    for (int i = 1; i <= ncol; ++i) {
      size_t e = (size_t)(i - 1) + (size_t)pcols * (k - 1);
      mcon[e] = mcon[e] * 100.0 / g.gravit;
    }
  for (size_t e = 0; e < n2; ++e) {
    t1[e] = t[e] + heat[e] * ztodt / g.cpair;
    double qn = q[e] + qtnd[e] * ztodt;
    q1[e] = (qn < 1.e-12) ? 1.e-12 : qn;
    winds[e] = u[e];
    winds[n2 + e] = v[e];
  }
  zmo_conv_evap(ncol, lchnk, t1.data(), pmid, pdel, q1.data(), landfrac, ev_s.data(), snwprd.data(),
                snwevmlt.data(), ev_q.data(), rprd, cld, ztodt, prec, snow, ntprprd.data(), ntsnprd.data(),
                flxprec, flxsnow);
  if (g.zm_org) {
    // zm_conv_intr.F90:773-777: ptend_loc%q(:ncol,:,ixorg) from |evapcdp| and org; physics_ptend_sum adds it to
    // the (zero) org tendency zm_convr returned, so orgt ends up holding ptend_all%q(:,:,ixorg)
    for (int k = 1; k <= pver; ++k)
      for (int i = 1; i <= ncol; ++i) {
        size_t e = (size_t)(i - 1) + (size_t)pcols * (k - 1);
        double x = std::min(1.0, std::max(0.0, (50.0 * 1000.0 * 1000.0 * std::fabs(ev_q[e])) - (tl_org[e] / 10800.0)));
        x = (x - tl_org[e]) / ztodt;
        tl_orgt[e] = tl_orgt[e] + x;
      }
  }
  // momentum transport is skipped by the cam3 physics package (zm_conv_intr.F90:808-859): ptend_all then carries
  // no wind tendency and no KE-dissipation heating
  if (!g.cam3) {
    const int domom[2] = {1, 1};
    zmo_momtran(lchnk, ncol, domom, winds.data(), 2, mu, md, du, eu, ed, dp, dsubcld, jt, maxg, ideep, 1,
                *lengath, 0, wtend.data(), pgu.data(), pgd.data(), icwu.data(), icwd.data(), ztodt, seten.data());
  }
  for (size_t e = 0; e < n2; ++e) {
    ptend_s[e] = g.cam3 ? (heat[e] + ev_s[e]) : (heat[e] + ev_s[e]) + seten[e];
    ptend_q[e] = qtnd[e] + ev_q[e];
    ptend_u[e] = g.cam3 ? 0.0 : wtend[e];
    ptend_v[e] = g.cam3 ? 0.0 : wtend[n2 + e];
    evapcdp[e] = ev_q[e];
  }
  // convtran1 (zm_conv_intr.F90:865-880): cloud liquid / ice (cnst_is_convtran1) on state1%q -- constituents
  // m >= 2 of state1 are those of state (only q(:,:,1), and the org tracer under zm_org, were updated) --
  // with the mass fluxes of this step and fake_dpdry = 0
  if (batch_tran1.pcnst > 0) {
    const size_t n3 = n2 * batch_tran1.pcnst, off = (size_t)(lchnk - 1) * n3;
    std::vector<double> fake_dpdry(n2, 0.0);
    zmo_convtran(lchnk, batch_tran1.doconv, batch_tran1.q + off, batch_tran1.pcnst, mu, md, du, eu, ed, dp, dsubcld,
                 jt, maxg, ideep, 1, *lengath, 0, batch_tran1.fracis + off, batch_tran1.ptend_q + off,
                 fake_dpdry.data(), ztodt, batch_tran1.dry);
  }
  return fails;
}

// zm_conv_tend_2 (zm_conv_intr.F90:955-1028) chunk by chunk under OpenMP: dpdry(i,:) = pdeldry(ideep(i),:)/100 for
// i <= lengath, 0 beyond (:1014-1017), then convtran (:1020-1024)
int zmo_conv_tend_2_batch(int nchunks, const int* doconvtran, const double* q, int pcnst, const double* pdeldry,
                          const double* fracis, double* ptend_q, double ztodt, const int* cnst_is_dry,
                          const double* mu, const double* md, const double* du, const double* eu, const double* ed,
                          const double* dp, const double* dsubcld, const int* jt, const int* maxg, const int* ideep,
                          const int* lengath, int nthreads) {
  const size_t pc = g.pcols, L = (size_t)g.pcols * g.pver, n3 = L * pcnst;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#else
  (void)nthreads;
#endif
#pragma omp parallel for schedule(dynamic, 4)
  for (int c = 0; c < nchunks; ++c) {
    std::vector<double> dpdry(L, 0.0);
    for (int i = 0; i < lengath[c]; ++i)
      for (int k = 0; k < g.pver; ++k)
        dpdry[(size_t)k * pc + i] = pdeldry[c * L + (size_t)k * pc + (ideep[c * pc + i] - 1)] / 100.0;
    zmo_convtran(c + 1, doconvtran, q + c * n3, pcnst, mu + c * L, md + c * L, du + c * L, eu + c * L, ed + c * L,
                 dp + c * L, dsubcld + c * pc, jt + c * pc, maxg + c * pc, ideep + c * pc, 1, lengath[c], 0,
                 fracis + c * n3, ptend_q + c * n3, dpdry.data(), ztodt, cnst_is_dry);
  }
  return 0;
}

void zmo_convtran1_fields(int pcnst, const int* doconvtran, const int* cnst_is_dry, const double* q,
                          const double* fracis, double* ptend_q) {
  batch_tran1 = Tran1{pcnst, doconvtran, cnst_is_dry, q, fracis, ptend_q};
}

// zm_conv_tend's history diagnostics that involve arithmetic (zm_conv_intr.F90:685-688, 700-706, 721-729), one chunk:
// freqzm, the ungathered mass fluxes mu_out / md_out in kg/m2/s, and the cloud top / base pressures pcont / pconb
void zmo_conv_tend_diag(int ncol, const double* ps, const double* pmid_, const double* mu_, const double* md_,
                        const int* jt, const int* maxg, const int* ideep, int lengath, double* freqzm,
                        double* mu_out_, double* md_out_, double* pcont, double* pconb) {
  const int pcols = g.pcols, pver = g.pver;
  C2 pmid{pmid_, pcols}, mu{mu_, pcols}, md{md_, pcols};
  A2 mu_out{mu_out_, pcols}, md_out{md_out_, pcols};
  for (int i = 1; i <= pcols; ++i) freqzm[i - 1] = 0.0;
  for (int i = 1; i <= lengath; ++i) freqzm[ideep[i - 1] - 1] = 1.0;
  for (int k = 1; k <= pver; ++k) for (int i = 1; i <= pcols; ++i) { mu_out(i, k) = 0.0; md_out(i, k) = 0.0; }  // :575-576
  for (int i = 1; i <= lengath; ++i)
    for (int k = 1; k <= pver; ++k) {
      const int ii = ideep[i - 1];
      mu_out(ii, k) = mu(i, k) * 100.0 / g.gravit;
      md_out(ii, k) = md(i, k) * 100.0 / g.gravit;
    }
  for (int i = 1; i <= ncol; ++i) { pcont[i - 1] = ps[i - 1]; pconb[i - 1] = ps[i - 1]; }
  for (int i = 1; i <= lengath; ++i)
    if (maxg[i - 1] > jt[i - 1]) {
      pcont[ideep[i - 1] - 1] = pmid(ideep[i - 1], jt[i - 1]);
      pconb[ideep[i - 1] - 1] = pmid(ideep[i - 1], maxg[i - 1]);
    }
}

int zmo_conv_tend_batch(int nchunks, const int* ncol, const double* t, const double* q, const double* u,
                        const double* v, const double* pmid, const double* pint, const double* pdel,
                        const double* zm, const double* zi, const double* phis, const double* pblh,
                        const double* tpert, const double* landfrac, const double* cld, double ztodt,
                        double* ptend_s, double* ptend_q, double* ptend_u, double* ptend_v, double* mcon,
                        double* cme, double* pflx, double* zdu, double* rliq, double* rice, double* jctop,
                        double* jcbot, double* prec, double* snow, double* ql, double* rprd, double* evapcdp,
                        double* flxprec, double* flxsnow, double* dlf, double* mu, double* md, double* du,
                        double* eu, double* ed, double* dp, double* dsubcld, int* jt, int* maxg, int* ideep,
                        int* lengath, double* cape, int nthreads) {
  const size_t pc = g.pcols, L = (size_t)g.pcols * g.pver, Lp = (size_t)g.pcols * g.pverp;
  int fails = 0;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#else
  (void)nthreads;
#endif
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : fails)
  for (int c = 0; c < nchunks; ++c) {
    if (g.zm_org) { tl_org = batch_org + (size_t)c * L; tl_orgt = batch_orgt + (size_t)c * L; tl_org2d = batch_org2d + (size_t)c * L; }
    fails += conv_tend_chunk(
        c + 1, ncol[c], t + c * L, q + c * L, u + c * L, v + c * L, pmid + c * L, pint + c * Lp, pdel + c * L,
        zm + c * L, zi + c * Lp, phis + c * pc, pblh + c * pc, tpert + c * pc, landfrac + c * pc, cld + c * L,
        ztodt, ptend_s + c * L, ptend_q + c * L, ptend_u + c * L, ptend_v + c * L, mcon + c * Lp, cme + c * L,
        pflx + c * Lp, zdu + c * L, rliq + c * pc, rice + c * pc, jctop + c * pc, jcbot + c * pc, prec + c * pc,
        snow + c * pc, ql + c * L, rprd + c * L, evapcdp + c * L, flxprec + c * Lp, flxsnow + c * Lp,
        dlf + c * L, mu + c * L, md + c * L, du + c * L, eu + c * L, ed + c * L, dp + c * L, dsubcld + c * pc,
        jt + c * pc, maxg + c * pc, ideep + c * pc, lengath + c, cape + c * pc);
  }
  batch_tran1 = Tran1{};        // one-shot, like the attached org fields
  return fails;
}

// ---- "next" rows N4: neighbours of the path with source in the reference ----
// geopotential_t, FV ('LR') and EUL/SE hydrostatic branches of physics/geopotential.F90:153-247
// (the generalized-virtual-temperature branch :248-310 is zmo_geopotential_t_gen below).
void zmo_geopotential_t(int ncol, int dycore_lr, const double* piln_, const double* pmln_, const double* pint_,
                        const double* pmid_, const double* pdel_, const double* rpdel_, const double* t_,
                        const double* q_, const double* rair_, double gravit, const double* zvir_, double* zi_,
                        double* zm_) {
  (void)pmln_;
  const int pcols = g.pcols, pver = g.pver, pverp = g.pverp;
  C2 piln{piln_, pcols}, pint{pint_, pcols}, pmid{pmid_, pcols}, pdel{pdel_, pcols}, rpdel{rpdel_, pcols},
      t{t_, pcols}, q{q_, pcols}, rair{rair_, pcols}, zvir{zvir_, pcols};
  A2 zi{zi_, pcols}, zm{zm_, pcols};
  std::vector<double> hkk(ncol + 1), hkl(ncol + 1);
  W2 rog(pcols, pver);
  for (int k = 1; k <= pver; ++k) for (int i = 1; i <= ncol; ++i) rog(i, k) = rair(i, k) / gravit;
  for (int i = 1; i <= ncol; ++i) zi(i, pverp) = 0.0;
  for (int k = pver; k >= 1; --k) {
    if (dycore_lr) {
      for (int i = 1; i <= ncol; ++i) {
        hkl[i] = piln(i, k + 1) - piln(i, k);
        hkk[i] = 1.0 - pint(i, k) * hkl[i] * rpdel(i, k);
      }
    } else {
      for (int i = 1; i <= ncol; ++i) {
        hkl[i] = pdel(i, k) / pmid(i, k);
        hkk[i] = 0.5 * hkl[i];
      }
    }
    for (int i = 1; i <= ncol; ++i) {
      double tvfac = 1.0 + zvir(i, k) * q(i, k);
      double tv = t(i, k) * tvfac;
      zm(i, k) = zi(i, k + 1) + rog(i, k) * tv * hkk[i];
      zi(i, k) = zi(i, k + 1) + rog(i, k) * tv * hkl[i];
    }
  }
}

// geopotential_t, generalized-virtual-temperature branch (physics/geopotential.F90:248-310), taken when
// dycore_is('MPAS') or dycore_is('SE'): q3 is q(pcols,pver,ncnst), species_idx the 1-based constituent indices of
// air_composition::thermodynamic_active_species_idx (the module is not in the reference tree: an input here).
void zmo_geopotential_t_gen(int ncol, int dycore_lr, int ncnst, int nspecies, const int* species_idx,
                            const double* piln_, const double* pmln_, const double* pint_, const double* pmid_,
                            const double* pdel_, const double* rpdel_, const double* t_, const double* q3_,
                            const double* rair_, double gravit, const double* zvir_, double* zi_, double* zm_) {
  (void)pmln_; (void)ncnst;
  const int pcols = g.pcols, pver = g.pver, pverp = g.pverp;
  C2 piln{piln_, pcols}, pint{pint_, pcols}, pmid{pmid_, pcols}, pdel{pdel_, pcols}, rpdel{rpdel_, pcols},
      t{t_, pcols}, rair{rair_, pcols}, zvir{zvir_, pcols};
  auto q = [&](int i, int k, int m) { return q3_[((size_t)(m - 1) * pver + (k - 1)) * pcols + (i - 1)]; };
  A2 zi{zi_, pcols}, zm{zm_, pcols};
  std::vector<double> hkk(ncol + 1), hkl(ncol + 1);
  W2 rog(pcols, pver), qfac(pcols, pver), sum_dry_mixing_ratio(pcols, pver);
  for (int k = 1; k <= pver; ++k) for (int i = 1; i <= ncol; ++i) rog(i, k) = rair(i, k) / gravit;
  for (int i = 1; i <= ncol; ++i) zi(i, pverp) = 0.0;
  // factor converting wet to dry mixing ratio (:255-263)
  for (int k = 1; k <= pver; ++k) for (int i = 1; i <= ncol; ++i) qfac(i, k) = 1.0;
  for (int idx = 1; idx <= nspecies; ++idx)
    for (int k = 1; k <= pver; ++k)
      for (int i = 1; i <= ncol; ++i) qfac(i, k) = qfac(i, k) - q(i, k, species_idx[idx - 1]);
  for (int k = 1; k <= pver; ++k) for (int i = 1; i <= ncol; ++i) qfac(i, k) = 1.0 / qfac(i, k);
  // sum of dry water mixing ratios (:265-275)
  for (int k = 1; k <= pver; ++k) for (int i = 1; i <= ncol; ++i) sum_dry_mixing_ratio(i, k) = 1.0;
  for (int idx = 1; idx <= nspecies; ++idx)
    for (int k = 1; k <= pver; ++k)
      for (int i = 1; i <= ncol; ++i)
        sum_dry_mixing_ratio(i, k) = sum_dry_mixing_ratio(i, k) + q(i, k, species_idx[idx - 1]) * qfac(i, k);
  for (int k = 1; k <= pver; ++k)
    for (int i = 1; i <= ncol; ++i) sum_dry_mixing_ratio(i, k) = 1.0 / sum_dry_mixing_ratio(i, k);
  for (int k = pver; k >= 1; --k) {
    if (dycore_lr) {
      for (int i = 1; i <= ncol; ++i) {
        hkl[i] = piln(i, k + 1) - piln(i, k);
        hkk[i] = 1.0 - pint(i, k) * hkl[i] * rpdel(i, k);
      }
    } else {
      for (int i = 1; i <= ncol; ++i) {
        hkl[i] = pdel(i, k) / pmid(i, k);
        hkk[i] = 0.5 * hkl[i];
      }
    }
    for (int i = 1; i <= ncol; ++i) {
      double tvfac = (1.0 + (zvir(i, k) + 1.0) * q(i, k, 1) * qfac(i, k)) * sum_dry_mixing_ratio(i, k);   // :303
      double tv = t(i, k) * tvfac;
      zm(i, k) = zi(i, k + 1) + rog(i, k) * tv * hkk[i];
      zi(i, k) = zi(i, k + 1) + rog(i, k) * tv * hkl[i];
    }
  }
}

// convect_diagnostics_calc (physics/convect_diagnostics.F90:115-249) for shallow_scheme == 'CLUBB_SGS'
// (the only configuration in which the routine's locals cnt2/cnb2 and the intent(out) qc2/rliq2 are
// defined, :187-198): merges the (zeroed) shallow fields into the deep-convection outputs.
void zmo_convect_diagnostics(int ncol, double* cmfmc_, double* qc_, double* qc2_, double* rliq_, double* rliq2_,
                             const double* pmid_, const double* rprddp_, double* cnt_, double* cnb_,
                             double* cmfmc2_, double* rprdsh_, double* rprdtot_, double* pcnt_, double* pcnb_) {
  const int pcols = g.pcols, pver = g.pver, pverp = g.pverp;
  A2 cmfmc{cmfmc_, pcols}, qc{qc_, pcols}, qc2{qc2_, pcols}, cmfmc2{cmfmc2_, pcols}, rprdsh{rprdsh_, pcols},
      rprdtot{rprdtot_, pcols};
  C2 pmid{pmid_, pcols}, rprddp{rprddp_, pcols};
  std::vector<double> cnt2(pcols), cnb2(pcols);
  for (int k = 1; k <= pverp; ++k) for (int i = 1; i <= pcols; ++i) cmfmc2(i, k) = 0.0;
  for (int k = 1; k <= pver; ++k) for (int i = 1; i <= pcols; ++i) { rprdsh(i, k) = 0.0; qc2(i, k) = 0.0; }
  for (int i = 1; i <= pcols; ++i) { rliq2_[i - 1] = 0.0; cnt2[i - 1] = (double)pver; cnb2[i - 1] = 1.0; }
  for (int k = 1; k <= pverp; ++k) for (int i = 1; i <= ncol; ++i) cmfmc(i, k) = cmfmc(i, k) + cmfmc2(i, k);
  for (int i = 1; i <= ncol; ++i) {
    if (cnt2[i - 1] < cnt_[i - 1]) cnt_[i - 1] = cnt2[i - 1];
    if (cnb2[i - 1] > cnb_[i - 1]) cnb_[i - 1] = cnb2[i - 1];
    if (cnb_[i - 1] == 1.0) cnb_[i - 1] = cnt_[i - 1];
    pcnt_[i - 1] = pmid(i, (int)cnt_[i - 1]);
    pcnb_[i - 1] = pmid(i, (int)cnb_[i - 1]);
  }
  for (int k = 1; k <= pver; ++k) for (int i = 1; i <= ncol; ++i) rprdtot(i, k) = rprdsh(i, k) + rprddp(i, k);
  for (int k = 1; k <= pver; ++k) for (int i = 1; i <= ncol; ++i) qc(i, k) = qc(i, k) + qc2(i, k);
  for (int i = 1; i <= ncol; ++i) rliq_[i - 1] = rliq_[i - 1] + rliq2_[i - 1];
}

}  // extern "C"
