// zm_plume.cuh -- per-column cloud model: cldprp + closure + q1q2_pjr + scatter + precipitation.
//
// One thread owns one convective (gathered) column; the ~40 per-level work arrays of the
// reference's cldprp/closure/q1q2_pjr live in thread-local arrays (local memory, interleaved
// per thread by the hardware so a warp's access to level k is coalesced and L1-resident).
// Reference: zm_conv.F90:3024-4026 (cldprp), 4028-4260 (closure), 4262-4421 (q1q2_pjr) and the
// glue of zm_convr 926-1027, 1047-1078, 1233-1331, 1495-1511, 1616-1649.
// The reference's chunk-wide loop bounds (khighest/klowest 3573-3578, kmin/kmax 4245-4246,
// ktm/kbm 4350-4355) only trim loops whose bodies are re-guarded per column, so a per-column
// formulation is numerically identical (SURVEY.md section 8a notes a6-a8).
#pragma once
#include "zm_kernels.cuh"

template <int LMAX>
struct PlumeCol {
  // 1-based: index k in 1..pver (+1 where interfaces are needed)
  double q[LMAX + 2], t[LMAX + 2], p[LMAX + 2], z[LMAX + 2], s[LMAX + 2], zf[LMAX + 2], dz[LMAX + 2];
  double shat[LMAX + 2], qhat[LMAX + 2], dp[LMAX + 2];
  double mu[LMAX + 2], eu[LMAX + 2], du[LMAX + 2], md[LMAX + 2], ed[LMAX + 2], sd[LMAX + 2],
      qd[LMAX + 2], mc[LMAX + 2], qu[LMAX + 2], su[LMAX + 2], qst[LMAX + 2], hmn[LMAX + 2],
      hsat[LMAX + 2], ql[LMAX + 2], cmeg[LMAX + 2], pflx[LMAX + 3], evp[LMAX + 2], cu[LMAX + 2],
      rprd[LMAX + 2], qcde[LMAX + 2];
  int jt, jlcl, j0, jd;
};

// cldprp for one column.  FULL=false stops after the cloud-top reset (zm_conv.F90:3646): that is
// all pass 1 needs (eu in 1/m) to diagnose the pass-2 entrainment rate.
template <int LMAX, bool FULL>
__device__ __noinline__ void cldprp_column(PlumeCol<LMAX>& C, int jb, int lel, double landfrac) {
  const int pver = P.pver, pverp = P.pverp, msg = P.msg, limcnv = P.limcnv;
  const double eps1 = P.eps1, zvir = P.zvir, cpvir = P.cpvir, dcol = P.dcol, tmelt = P.tmelt;
  const double rl = P.rl, rd = P.rgas, grav = P.grav, cp = P.cpres;
  const int mx = jb;
  double gamma[LMAX + 2], hu[LMAX + 2], hd[LMAX + 2], eps[LMAX + 3], f[LMAX + 3], k1[LMAX + 3],
      i2[LMAX + 3], i3[LMAX + 3], i4[LMAX + 3], qsthat[LMAX + 2], hsthat[LMAX + 2], gamhat[LMAX + 2],
      qds[LMAX + 2], mcp[LMAX + 2], mrl[LMAX + 2], tu[LMAX + 2], td[LMAX + 2];

  const double c0mask = P.c0_ocn * (1.0 - landfrac) + P.c0_lnd * landfrac;
  const double tiedke_msk = P.tiedke_add * (1.0 - landfrac) + P.tiedke_lnd * landfrac;
  for (int k = 1; k <= pver; ++k) C.dz[k] = C.zf[k] - C.zf[k + 1];
  C.pflx[1] = 0.0;
  for (int k = 1; k <= pver; ++k) {
    k1[k] = 0.0; i2[k] = 0.0; i3[k] = 0.0; i4[k] = 0.0;
    C.mu[k] = 0.0; f[k] = 0.0; eps[k] = 0.0; C.eu[k] = 0.0; C.du[k] = 0.0; C.ql[k] = 0.0;
    C.cu[k] = 0.0; C.evp[k] = 0.0; C.cmeg[k] = 0.0;
    qds[k] = C.q[k];
    C.md[k] = 0.0; C.ed[k] = 0.0;
    C.sd[k] = C.s[k];
    C.qd[k] = C.q[k];
    C.mc[k] = 0.0;
    C.qu[k] = C.q[k];
    C.su[k] = C.s[k];
    double est, qs;
    qsat_hPa(C.t[k], C.p[k], est, qs);
    if (C.p[k] - est <= 0.0) qs = 1.0;
    C.qst[k] = qs;
    double mrd = (1.0 + zvir * C.q[k]) * rd;
    mcp[k] = (1.0 + cpvir * C.q[k]) * cp;
    mrl[k] = (1.0 - dcol * (C.t[k] - tmelt)) * rl;
    gamma[k] = qs * (1.0 + qs / eps1) * eps1 * mrl[k] / (mrd * (C.t[k] * C.t[k])) * mrl[k] / mcp[k];
    C.hmn[k] = mcp[k] * C.t[k] + grav * C.z[k] + mrl[k] * C.q[k];
    C.hsat[k] = mcp[k] * C.t[k] + grav * C.z[k] + mrl[k] * qs;
    hu[k] = C.hmn[k];
    hd[k] = C.hmn[k];
    C.rprd[k] = 0.0;
    C.qcde[k] = 0.0;
    td[k] = (hd[k] - grav * C.zf[k] - (1.0 + dcol * tmelt) * rl * qds[k]) /
            (cp * (1.0 + (cpvir - dcol * (rl / cp)) * qds[k]));
  }
  k1[pver + 1] = 0.0; i2[pver + 1] = 0.0; i3[pver + 1] = 0.0; i4[pver + 1] = 0.0;
  f[pver + 1] = 0.0; eps[pver + 1] = 0.0; C.mu[pver + 1] = 0.0; C.ql[pver + 1] = 0.0;

  for (int k = 1; k <= msg + 1; ++k) { hsthat[k] = C.hsat[k]; qsthat[k] = C.qst[k]; gamhat[k] = gamma[k]; }
  double totpcp = 0.0, totevp = 0.0;
  for (int k = msg + 2; k <= pver; ++k) {
    if (fabs(C.qst[k - 1] - C.qst[k]) > 1.E-6)
      qsthat[k] = zmm::log_(C.qst[k - 1] / C.qst[k]) * C.qst[k - 1] * C.qst[k] / (C.qst[k - 1] - C.qst[k]);
    else
      qsthat[k] = C.qst[k];
    hsthat[k] = mcp[k] * C.shat[k] + mrl[k] * qsthat[k];
    if (fabs(gamma[k - 1] - gamma[k]) > 1.E-6)
      gamhat[k] = zmm::log_(gamma[k - 1] / gamma[k]) * gamma[k - 1] * gamma[k] / (gamma[k - 1] - gamma[k]);
    else
      gamhat[k] = gamma[k];
  }

  int jt = max(lel, limcnv + 1);
  jt = min(jt, pver);
  int jd = pver;
  int jlcl = lel;
  double hmin = 1.E6;
  int j0 = 0;
  for (int k = msg + 1; k <= pver; ++k)
    if (C.hsat[k] <= hmin && k >= jt && k <= jb) { hmin = C.hsat[k]; j0 = k; }
  j0 = min(j0, jb - 2);
  j0 = max(j0, jt + 2);
  j0 = min(j0, pver);
  for (int k = msg + 1; k <= pver; ++k)
    if (k >= jt && k <= jb) {
      hu[k] = C.hmn[mx] + cp * tiedke_msk;
      C.su[k] = C.s[mx] + tiedke_msk / (1.0 + cpvir * C.qu[k]);
    }
  for (int k = pver - 1; k >= msg + 1; --k)
    if (k < jb && k >= jt) {
      k1[k] = k1[k + 1] + (C.hmn[mx] - C.hmn[k]) * C.dz[k];
      double ihat = 0.5 * (k1[k + 1] + k1[k]);
      i2[k] = i2[k + 1] + ihat * C.dz[k];
      double idag = 0.5 * (i2[k + 1] + i2[k]);
      i3[k] = i3[k + 1] + idag * C.dz[k];
      double iprm = 0.5 * (i3[k + 1] + i3[k]);
      i4[k] = i4[k + 1] + iprm * C.dz[k];
    }
  hmin = 1.E6;
  double expdif = 0.0;
  for (int k = msg + 1; k <= pver; ++k)
    if (k >= j0 && k <= jb && C.hmn[k] <= hmin) { hmin = C.hmn[k]; expdif = C.hmn[mx] - hmin; }

  for (int k = msg + 2; k <= pver; ++k) {
    double expnum = 0.0;
    if (k < jt || k >= jb) {
      k1[k] = 0.0;
      expnum = 0.0;
    } else {
      expnum = C.hmn[mx] - (C.hsat[k - 1] * (C.zf[k] - C.z[k]) + C.hsat[k] * (C.z[k - 1] - C.zf[k])) /
                               (C.z[k - 1] - C.z[k]);
    }
    if ((expdif > 100.0 && expnum > 0.0) && k1[k] > expnum * C.dz[k]) {
      double ft = expnum / k1[k], K1 = k1[k], I2 = i2[k], I3 = i3[k], I4 = i4[k];
      double fk = ft + I2 / K1 * (ft * ft) + (2.0 * (I2 * I2) - K1 * I3) / (K1 * K1) * ((ft * ft) * ft) +
                  (-5.0 * K1 * I2 * I3 + 5.0 * ((I2 * I2) * I2) + (K1 * K1) * I4) / ((K1 * K1) * K1) *
                      ((ft * ft) * (ft * ft));
      fk = fmax2(fk, 0.0);
      fk = fmin2(fk, P.entrmn);
      f[k] = fk;
    }
  }
  if (j0 < jb)
    if (f[j0] < 1.E-6 && f[j0 + 1] > f[j0]) j0 = j0 + 1;
  for (int k = msg + 2; k <= pver; ++k)
    if (k >= jt && k <= j0) f[k] = fmax2(f[k], f[k - 1]);
  const double eps0 = f[j0];
  eps[jb] = eps0;
  for (int k = pver; k >= msg + 1; --k)
    if (k >= j0 && k <= jb) eps[k] = f[j0];
  for (int k = pver; k >= msg + 1; --k)
    if (k < j0 && k >= jt) eps[k] = f[k];

  if (eps0 > 0.0) {
    C.mu[jb] = 1.0;
    C.eu[jb] = C.mu[jb] / C.dz[jb];
  }
  {
    const int tmplel = jt;
    for (int k = pver; k >= msg + 1; --k)
      if (eps0 > 0.0 && (k >= tmplel && k < jb)) {
        double zuef = C.zf[k] - C.zf[jb];
        double rmue = (1.0 / eps0) * (zmm::exp_(eps[k + 1] * zuef) - 1.0) / zuef;
        C.mu[k] = (1.0 / eps0) * (zmm::exp_(eps[k] * zuef) - 1.0) / zuef;
        C.eu[k] = (rmue - C.mu[k + 1]) / C.dz[k];
        C.du[k] = (rmue - C.mu[k]) / C.dz[k];
      }
  }
  for (int k = jb - 1; k >= lel; --k)
    if (eps0 > 0.0) {
      if (C.mu[k] < 0.02) {
        hu[k] = C.hmn[k];
        C.mu[k] = 0.0;
        C.eu[k] = 0.0;
        C.du[k] = C.mu[k + 1] / C.dz[k];
      } else {
        hu[k] = C.mu[k + 1] / C.mu[k] * hu[k + 1] +
                C.dz[k] / C.mu[k] * (C.eu[k] * C.hmn[k] - C.du[k] * C.hsat[k]);
      }
    }
  {
    bool doit = true;
    const double totfrz = 0.0;
    for (int k = jb - 2; k >= lel - 1; --k)
      if (doit) {
        if (hu[k] <= hsthat[k] && hu[k + 1] > hsthat[k + 1] && C.mu[k] >= 0.02) {
          if (hu[k] - hsthat[k] < -2000.0) { jt = k + 1; doit = false; }
          else                             { jt = k;     doit = false; }
        } else if ((hu[k] > hu[jb] && totfrz <= 0.0) || C.mu[k] < 0.02) {
          jt = k + 1;
          doit = false;
        }
      }
  }
  for (int k = pver; k >= msg + 1; --k) {
    if (k >= lel && k <= jt && eps0 > 0.0) {
      C.mu[k] = 0.0; C.eu[k] = 0.0; C.du[k] = 0.0; hu[k] = C.hmn[k];
    }
    if (k == jt && eps0 > 0.0) {
      C.du[k] = C.mu[k + 1] / C.dz[k];
      C.eu[k] = 0.0;
      C.mu[k] = 0.0;
    }
  }
  C.jt = jt; C.j0 = j0; C.jlcl = jlcl; C.jd = jd;
  if (!FULL) return;

  for (int k = pver; k >= msg + 2; --k)
    tu[k] = (hu[k] - grav * C.zf[k] - (1.0 + dcol * tmelt) * rl * C.qu[k]) /
            (cp * (1.0 + (cpvir - dcol * (rl / cp)) * C.qu[k]));
  {
    bool done = false;
    for (int k = pver; k >= msg + 2; --k) {
      if (k == jb && eps0 > 0.0) {
        C.qu[k] = C.q[mx];
        tu[k] = (hu[k] - grav * C.zf[k] - (1.0 + dcol * tmelt) * rl * C.qu[k]) /
                (cp * (1.0 + (cpvir - dcol * (rl / cp)) * C.qu[k]));
        C.su[k] = (hu[k] - (1.0 - dcol * (tu[k] - tmelt)) * rl * C.qu[k]) / ((1.0 + cpvir * C.qu[k]) * cp);
      }
      if ((!done && k > jt && k < jb) && eps0 > 0.0) {
        C.su[k] = C.mu[k + 1] / C.mu[k] * C.su[k + 1] + C.dz[k] / C.mu[k] * (C.eu[k] - C.du[k]) * C.s[k];
        C.qu[k] = C.mu[k + 1] / C.mu[k] * C.qu[k + 1] +
                  C.dz[k] / C.mu[k] * (C.eu[k] * C.q[k] - C.du[k] * C.qst[k]);
        // default-real literal 0.85 in the reference (zm_conv.F90:3680) == (double)0.85f
        tu[k] = C.su[k] - grav / ((1.0 + 0.85000002384185791015625 * C.qu[k]) * cp) * C.zf[k];
        double qstu = qsat_hPa_q(tu[k], (C.p[k] + C.p[k - 1]) / 2.0);
        if (C.qu[k] >= qstu) { jlcl = k; done = true; }
      }
    }
  }
  for (int k = msg + 2; k <= pver; ++k)
    if ((k > jt && k <= jlcl) && eps0 > 0.0) {
      C.qu[k] = qsthat[k] + gamhat[k] * (hu[k] - hsthat[k]) /
                                ((1.0 - dcol * (tu[k] - tmelt)) * rl * (1.0 + gamhat[k]));
      C.su[k] = C.shat[k] + (hu[k] - hsthat[k]) / ((1.0 + cpvir * C.qu[k]) * cp * (1.0 + gamhat[k]));
      tu[k] = C.su[k] - grav / ((1.0 + cpvir * C.qu[k]) * cp) * C.zf[k];
    }
  for (int k = pver; k >= msg + 2; --k)
    if (k >= jt && k < jb && eps0 > 0.0) {
      double cuk = ((C.mu[k] * C.su[k] - C.mu[k + 1] * C.su[k + 1]) / C.dz[k] - (C.eu[k] - C.du[k]) * C.s[k]) /
                   (rl / cp) * ((1.0 + cpvir * C.qu[k]) / (1.0 - dcol * (tu[k] - tmelt)));
      if (k == jt) cuk = 0.0;
      C.cu[k] = fmax2(0.0, cuk);
    }
  for (int k = pver; k >= msg + 2; --k) {
    C.rprd[k] = 0.0;
    if (k >= jt && k < jb && eps0 > 0.0 && C.mu[k] >= 0.0) {
      if (C.mu[k] > 0.0) {
        double ql1 = 1.0 / C.mu[k] * (C.mu[k + 1] * C.ql[k + 1] - C.dz[k] * C.du[k] * C.ql[k + 1] + C.dz[k] * C.cu[k]);
        C.ql[k] = ql1 / (1.0 + C.dz[k] * c0mask);
      } else {
        C.ql[k] = 0.0;
      }
      totpcp = totpcp + C.dz[k] * (C.cu[k] - C.du[k] * C.ql[k + 1]);
      C.rprd[k] = c0mask * C.mu[k] * C.ql[k];
      C.qcde[k] = C.ql[k];
    }
  }

  // downdraft (zm_conv.F90:3880-3975)
  const double alfa = P.alfadet;
  double epsm = 0.0;
  jt = min(jt, jb - 1);
  jd = max(j0, jt + 1);
  jd = min(jd, jb);
  hd[jd] = C.hmn[jd - 1];
  if (jd < jb && eps0 > 0.0) {
    epsm = eps0;
    C.md[jd] = -alfa * epsm / eps0;
  }
  for (int k = msg + 1; k <= pver; ++k)
    if ((k > jd && k <= jb) && eps0 > 0.0) {
      double zdef = C.zf[jd] - C.zf[k];
      C.md[k] = -alfa / (2.0 * eps0) * (zmm::exp_(2.0 * epsm * zdef) - 1.0) / zdef;
    }
  for (int k = msg + 1; k <= pver; ++k)
    if ((k >= jt && k <= jb) && eps0 > 0.0 && jd < jb) {
      double ratmjb = fmin2(fabs(C.mu[jb] / C.md[jb]), 1.0);
      C.md[k] = C.md[k] * ratmjb;
    }
  const double small = 1.e-20;
  for (int k = msg + 1; k <= pver; ++k)
    if ((k >= jt && k <= pver) && eps0 > 0.0) {
      C.ed[k - 1] = (C.md[k - 1] - C.md[k]) / C.dz[k - 1];
      double mdt = fmin2(C.md[k], -small);
      hd[k] = (C.md[k - 1] * hd[k - 1] - C.dz[k - 1] * C.ed[k - 1] * C.hmn[k - 1]) / mdt;
    }
  for (int k = msg + 2; k <= pver; ++k)
    if ((k >= jd && k <= jb) && eps0 > 0.0 && jd < jb) {
      qds[k] = qsthat[k] + gamhat[k] * (hd[k] - hsthat[k]) / (rl * (1.0 + gamhat[k]));
      td[k] = (hd[k] - grav * C.zf[k] - (1.0 + dcol * tmelt) * rl * qds[k]) /
              (cp * (1.0 + (cpvir - dcol * (rl / cp)) * qds[k]));
      qds[k] = qsthat[k] + gamhat[k] * (hd[k] - hsthat[k]) /
                               ((1.0 - dcol * (td[k] - tmelt)) * rl * (1.0 + gamhat[k]));
    }
  {
    C.qd[jd] = qds[jd];
    const int k = jd;
    C.sd[jd] = (hd[jd] - (1.0 - dcol * (td[k] - tmelt)) * rl * C.qd[jd]) / ((1.0 + cpvir * C.qd[k]) * cp);
    td[k] = C.sd[k] - grav / ((1.0 + cpvir * C.qd[k]) * cp) * C.zf[k];
  }
  for (int k = msg + 2; k <= pver; ++k)
    if (k >= jd && k < jb && eps0 > 0.0) {
      C.qd[k + 1] = qds[k + 1];
      double ev = -C.ed[k] * C.q[k] + (C.md[k] * C.qd[k] - C.md[k + 1] * C.qd[k + 1]) / C.dz[k];
      ev = fmax2(ev, 0.0);
      C.evp[k] = ev;
      double mdt = fmin2(C.md[k + 1], -small);
      C.sd[k + 1] = (((1.0 - dcol * (td[k] - tmelt)) * rl / ((1.0 + cpvir * C.qd[k]) * cp) * ev - C.ed[k] * C.s[k]) * C.dz[k] +
                     C.md[k] * C.sd[k]) / mdt;
      totevp = totevp - C.dz[k] * C.ed[k] * C.q[k];
    }
  totevp = totevp + C.md[jd] * C.qd[jd] - C.md[jb] * C.qd[jb];
  totpcp = fmax2(totpcp, 0.0);
  totevp = fmax2(totevp, 0.0);
  for (int k = msg + 2; k <= pver; ++k) {
    if (totevp > 0.0 && totpcp > 0.0) {
      C.md[k] = C.md[k] * fmin2(1.0, totpcp / (totevp + totpcp));
      C.ed[k] = C.ed[k] * fmin2(1.0, totpcp / (totevp + totpcp));
      C.evp[k] = C.evp[k] * fmin2(1.0, totpcp / (totevp + totpcp));
    } else {
      C.md[k] = 0.0; C.ed[k] = 0.0; C.evp[k] = 0.0;
    }
    C.cmeg[k] = C.cu[k] - C.evp[k];
    C.rprd[k] = C.rprd[k] - C.evp[k];
  }
  C.pflx[1] = 0.0;
  for (int k = 2; k <= pverp; ++k) C.pflx[k] = C.pflx[k - 1] + C.rprd[k - 1] * C.dz[k - 1];
  for (int k = msg + 1; k <= pver; ++k) C.mc[k] = C.mu[k] + C.md[k];
  C.jt = jt; C.j0 = j0; C.jlcl = jlcl; C.jd = jd;
}

// gather one column into thread-local arrays (zm_conv.F90:926-940, 980-1027 / 1114-1195)
template <int LMAX>
__device__ __forceinline__ double gather_column(PlumeCol<LMAX>& C, const ConvrIn& in, int c, int i, int maxg) {
  const int pver = P.pver, pcols = P.pcols, msg = P.msg;
  const double zs = in.geos[(size_t)c * pcols + i] * P.rgrav;
  for (int k = 1; k <= pver; ++k) {
    size_t e = cidx(c, k - 1, i, pver);
    C.dp[k] = 0.01 * in.dpp[e];
    C.q[k] = in.qh[e];
    C.t[k] = in.t[e];
    C.p[k] = in.pap[e] * 0.01;
    C.z[k] = in.zm[e] + zs;
    C.s[k] = C.t[k] + (P.grav / ((1.0 + P.zvir * C.q[k]) * P.cpres)) * C.z[k];
    C.zf[k] = in.zi[cidx(c, k - 1, i, pver + 1)] + zs;
  }
  C.zf[pver + 1] = in.zi[cidx(c, pver, i, pver + 1)] + zs;
  double dsubcld = 0.0;
  for (int k = msg + 1; k <= pver; ++k)
    if (k >= maxg) dsubcld = dsubcld + C.dp[k];
  for (int k = 1; k <= msg + 1; ++k) { C.shat[k] = C.s[k]; C.qhat[k] = C.q[k]; }
  for (int k = msg + 2; k <= pver; ++k) {
    double sdifr = 0.0, qdifr = 0.0;
    if (C.s[k] > 0.0 || C.s[k - 1] > 0.0) sdifr = fabs((C.s[k] - C.s[k - 1]) / fmax2(C.s[k - 1], C.s[k]));
    if (C.q[k] > 0.0 || C.q[k - 1] > 0.0) qdifr = fabs((C.q[k] - C.q[k - 1]) / fmax2(C.q[k - 1], C.q[k]));
    if (sdifr > 1.E-6) C.shat[k] = zmm::log_(C.s[k - 1] / C.s[k]) * C.s[k - 1] * C.s[k] / (C.s[k - 1] - C.s[k]);
    else               C.shat[k] = 0.5 * (C.s[k] + C.s[k - 1]);
    if (qdifr > 1.E-6) C.qhat[k] = zmm::log_(C.q[k - 1] / C.q[k]) * C.q[k - 1] * C.q[k] / (C.q[k - 1] - C.q[k]);
    else               C.qhat[k] = 0.5 * (C.q[k] + C.q[k - 1]);
  }
  return dsubcld;
}

// ---- pass-1 plume: diagnose the pass-2 test-parcel entrainment rate (zm_conv.F90:1047-1078) ---
template <int LMAX>
__global__ void __launch_bounds__(64)
k_cldprp_pass1(ConvrIn in, ConvrWork w) {
  int gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= w.count[0]) return;
  const int col = w.wl1[gid];
  const int pcols = P.pcols, pver = P.pver, msg = P.msg;
  const int c = col / pcols, i = col - c * pcols;
  PlumeCol<LMAX> C;
  const int maxg = w.mx[col];
  gather_column<LMAX>(C, in, c, i, maxg);
  cldprp_column<LMAX, false>(C, maxg, w.lel[col], in.landfrac[(size_t)c * pcols + i]);
  double hk = 0.0, dmmx = 0.0, dmsm = 0.0;
  const double orgc = 1.0;
  double dm = -1.0;
  for (int k = pver; k >= msg + 1; --k)
    if (C.eu[k] > 0.0) {
      dmmx = -fmax2(-dmmx, C.eu[k]);
      dmsm = dmsm - C.eu[k];
      hk = hk + 1.0;
    }
  if (hk > 0.0) {
    dmsm = dmsm / hk;
    dm = dmsm * orgc + dmmx * (1.0 - orgc);
  }
  w.dmpdz[col] = dm;
}

// ---- final plume: cldprp #2 + closure + limiter + q1q2 + scatter + prec -----------------------
template <int LMAX>
__global__ void __launch_bounds__(64)
k_plume(ConvrIn in, ConvrOut o, ConvrWork w) {
  int gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= w.count[1]) return;
  const int col = w.wl2[2 * gid], slot = w.wl2[2 * gid + 1];
  const int pcols = P.pcols, pver = P.pver, pverp = P.pverp, msg = P.msg;
  const int ncolpad = in.nchunks * pcols;
  const int c = col / pcols, i = col - c * pcols;      // ungathered position
  const int gi = slot - c * pcols;                     // gathered position (0-based)
  const double eps1 = P.eps1, rl = P.rl, rd = P.rgas, grav = P.grav, cp = P.cpres;
  const double delt = in.delt;
  PlumeCol<LMAX> C;
  double tp[LMAX + 2], qstp[LMAX + 2];
  const int maxg = w.mx[col], lel = w.lel[col], lcl = w.lcl[col];
  const double capeg = w.cape[col], tlg = w.tl[col];
  const double landfrac = in.landfrac[(size_t)c * pcols + i];
  const double dsubcld = gather_column<LMAX>(C, in, c, i, maxg);
  for (int k = 1; k <= pver; ++k) {
    tp[k] = w.tp[(size_t)(k - 1) * ncolpad + col];
    qstp[k] = w.qstp[(size_t)(k - 1) * ncolpad + col];
  }
  cldprp_column<LMAX, true>(C, maxg, lel, landfrac);
  const int jt = C.jt, mx = maxg;
  const double dmpdz = w.dmpdz[col];

  // 1/m -> 1/mb (zm_conv.F90:1252-1262)
  for (int k = msg + 1; k <= pver; ++k) {
    double dzf = C.zf[k] - C.zf[k + 1];
    C.du[k] = C.du[k] * dzf / C.dp[k];
    C.eu[k] = C.eu[k] * dzf / C.dp[k];
    C.ed[k] = C.ed[k] * dzf / C.dp[k];
    C.cu[k] = C.cu[k] * dzf / C.dp[k];
    C.cmeg[k] = C.cmeg[k] * dzf / C.dp[k];
    C.rprd[k] = C.rprd[k] * dzf / C.dp[k];
    C.evp[k] = C.evp[k] * dzf / C.dp[k];
  }

  // ---- closure (zm_conv.F90:4028-4260) ----
  double mb = 0.0;
  {
    double dtmdt[LMAX + 2], dqmdt[LMAX + 2];
    double eb = C.p[mx] * C.q[mx] / (eps1 + C.q[mx]);
    double dtbdt = (1.0 / dsubcld) * (C.mu[mx] * (C.shat[mx] - C.su[mx]) + C.md[mx] * (C.shat[mx] - C.sd[mx]));
    double dqbdt = (1.0 / dsubcld) * (C.mu[mx] * (C.qhat[mx] - C.qu[mx]) + C.md[mx] * (C.qhat[mx] - C.qd[mx]));
    double epq = eps1 + C.q[mx];
    double debdt = eps1 * C.p[mx] / (epq * epq) * dqbdt;
    double den = 3.5 * zmm::log_(C.t[mx]) - zmm::log_(eb) - 4.805;
    double dtldt = -2840.0 * (3.5 / C.t[mx] * dtbdt - debdt / eb) / (den * den);
    for (int k = msg + 1; k <= pver; ++k) { dtmdt[k] = 0.0; dqmdt[k] = 0.0; }
    for (int k = msg + 1; k <= pver - 1; ++k)
      if (k == jt) {
        dqmdt[k] = (1.0 / C.dp[k]) * (C.mu[k + 1] * (C.qu[k + 1] - C.qhat[k + 1] + C.ql[k + 1]) +
                                      C.md[k + 1] * (C.qd[k + 1] - C.qhat[k + 1]));
        dtmdt[k] = (1.0 / C.dp[k]) * (C.mu[k + 1] * (C.su[k + 1] - C.shat[k + 1] - rl / cp * C.ql[k + 1]) +
                                      C.md[k + 1] * (C.sd[k + 1] - C.shat[k + 1]));
      }
    const double beta = 0.0;
    for (int k = msg + 1; k <= pver - 1; ++k)
      if (k > jt && k < mx) {
        dtmdt[k] = (C.mc[k] * (C.shat[k] - C.s[k]) - C.mc[k + 1] * (C.shat[k + 1] - C.s[k])) / C.dp[k] -
                   rl / cp * C.du[k] * (beta * C.ql[k] + (1 - beta) * C.ql[k + 1]);
        dqmdt[k] = (C.mu[k + 1] * (C.qu[k + 1] - C.qhat[k + 1] + cp / rl * (C.su[k + 1] - C.s[k])) -
                    C.mu[k] * (C.qu[k] - C.qhat[k] + cp / rl * (C.su[k] - C.s[k])) +
                    C.md[k + 1] * (C.qd[k + 1] - C.qhat[k + 1] + cp / rl * (C.sd[k + 1] - C.s[k])) -
                    C.md[k] * (C.qd[k] - C.qhat[k] + cp / rl * (C.sd[k] - C.s[k]))) / C.dp[k] +
                   C.du[k] * (beta * C.ql[k] + (1 - beta) * C.ql[k + 1]);
      }
    double dadt = 0.0;
    // dboydt accumulated in ascending k (zm_conv.F90:4247-4253); rows outside both windows hold
    // 0 here (the reference leaves them unset; lel<=lcl<mx in every reachable case)
    for (int k = msg + 1; k <= pver; ++k) {
      double dboydt = 0.0;
      if (k >= lel && k <= lcl) {
        double pw = zmm::pow_(1000.0 / C.p[k], rd / cp);
        double thetavp = tp[k] * pw * (1.0 + 1.608 * qstp[k] - C.q[mx]);
        double thetavm = C.t[k] * pw * (1.0 + 0.608 * C.q[k]);
        double dqsdtp = qstp[k] * (1.0 + qstp[k] / eps1) * eps1 * rl / (rd * (tp[k] * tp[k]));
        double dtpdt = tp[k] / (1.0 + rl / cp * (dqsdtp - qstp[k] / tp[k])) *
                       (dtbdt / C.t[mx] + rl / cp * (dqbdt / tlg - C.q[mx] / (tlg * tlg) * dtldt));
        dboydt = ((dtpdt / tp[k] + 1.0 / (1.0 + 1.608 * qstp[k] - C.q[mx]) * (1.608 * dqsdtp * dtpdt - dqbdt)) -
                  (dtmdt[k] / C.t[k] + 0.608 / (1.0 + 0.608 * C.q[k]) * dqmdt[k])) *
                 grav * thetavp / thetavm;
      }
      if (k > lcl && k < mx) {
        double pw = zmm::pow_(1000.0 / C.p[k], rd / cp);
        double thetavp = tp[k] * pw * (1.0 + 0.608 * C.q[mx]);
        double thetavm = C.t[k] * pw * (1.0 + 0.608 * C.q[k]);
        dboydt = (dtbdt / C.t[mx] + 0.608 / (1.0 + 0.608 * C.q[mx]) * dqbdt - dtmdt[k] / C.t[k] -
                  0.608 / (1.0 + 0.608 * C.q[k]) * dqmdt[k]) *
                 grav * thetavp / thetavm;
      }
      if (k >= lel && k <= mx - 1) dadt = dadt + dboydt * (C.zf[k] - C.zf[k + 1]);
    }
    double dltaa = -1.0 * (capeg - P.capelmt);
    if (dadt != 0.0) mb = fmax2(dltaa / P.tau / dadt, 0.0);
  }

  // mass-flux limiter (zm_conv.F90:1285-1308)
  {
    double mumax = 0.0;
    for (int k = msg + 2; k <= pver; ++k) mumax = fmax2(mumax, C.mu[k] / C.dp[k]);
    if (mumax > 0.0) mb = fmin2(mb, 0.5 / (delt * mumax));
    else mb = 0.0;
    if (P.no_deep_pbl)
      if (in.zm[cidx(c, jt - 1, i, pver)] < in.pblh[(size_t)c * pcols + i]) mb = 0.0;
  }
  for (int k = msg + 1; k <= pver; ++k) {
    C.mu[k] = C.mu[k] * mb; C.md[k] = C.md[k] * mb; C.mc[k] = C.mc[k] * mb; C.du[k] = C.du[k] * mb;
    C.eu[k] = C.eu[k] * mb; C.ed[k] = C.ed[k] * mb; C.cmeg[k] = C.cmeg[k] * mb; C.rprd[k] = C.rprd[k] * mb;
    C.cu[k] = C.cu[k] * mb; C.evp[k] = C.evp[k] * mb;
    C.pflx[k + 1] = C.pflx[k + 1] * mb * 100.0 / grav;
  }

  // ---- q1q2_pjr (zm_conv.F90:4262-4421) + scatter (1495-1511) + prec (1629-1649) ----
  double prec = 0.0, rliq = 0.0;
  {
    double dsdt[LMAX + 2], dqdt[LMAX + 2], dl[LMAX + 2];
    for (int k = msg + 1; k <= pver; ++k) { dsdt[k] = 0.0; dqdt[k] = 0.0; dl[k] = 0.0; }
    for (int k = msg + 1; k <= pver - 1; ++k) {
      double emc = -C.cu[k] + C.evp[k];
      dsdt[k] = -rl / cp * emc + (C.mu[k + 1] * (C.su[k + 1] - C.shat[k + 1]) - C.mu[k] * (C.su[k] - C.shat[k]) +
                                  C.md[k + 1] * (C.sd[k + 1] - C.shat[k + 1]) - C.md[k] * (C.sd[k] - C.shat[k])) / C.dp[k];
      dqdt[k] = emc + (C.mu[k + 1] * (C.qu[k + 1] - C.qhat[k + 1]) - C.mu[k] * (C.qu[k] - C.qhat[k]) +
                       C.md[k + 1] * (C.qd[k + 1] - C.qhat[k + 1]) - C.md[k] * (C.qd[k] - C.qhat[k])) / C.dp[k];
      dl[k] = C.du[k] * C.qcde[k + 1];
    }
    for (int k = msg + 1; k <= pver; ++k) {
      if (k == mx) {
        dsdt[k] = (1.0 / dsubcld) * (-C.mu[k] * (C.su[k] - C.shat[k]) - C.md[k] * (C.sd[k] - C.shat[k]));
        dqdt[k] = (1.0 / dsubcld) * (-C.mu[k] * (C.qu[k] - C.qhat[k]) - C.md[k] * (C.qd[k] - C.qhat[k]));
      } else if (k > mx) {
        dsdt[k] = dsdt[k - 1];
        dqdt[k] = dqdt[k - 1];
      }
    }
    // scatter to the ungathered column i of chunk c
    for (int k = msg + 1; k <= pver; ++k) {
      size_t e = cidx(c, k - 1, i, pver);
      o.qtnd[e] = dqdt[k];
      o.cme[e] = C.cmeg[k];
      o.rprd[e] = C.rprd[k];
      o.zdu[e] = C.du[k];
      o.heat[e] = dsdt[k] * P.cpres;
      o.dlf[e] = dl[k];
      o.ql[e] = C.ql[k];
      o.eurt[e] = -dmpdz;
      size_t ep = cidx(c, k - 1, i, pverp);
      o.mcon[ep] = C.mc[k];
      o.pflx[ep] = C.pflx[k];
    }
    o.pflx[cidx(c, pverp - 1, i, pverp)] = C.pflx[pverp];
    for (int k = pver; k >= msg + 1; --k) {
      size_t e = cidx(c, k - 1, i, pver);
      double dppk = in.dpp[e], qhk = in.qh[e];
      double qnew = qhk + 2.0 * delt * dqdt[k];
      prec = prec - dppk * (qnew - qhk) - dppk * (dl[k] + 0.0) * 2.0 * delt;
    }
    prec = P.rgrav * fmax2(prec, 0.0) / (2.0 * delt) / 1000.0;
    for (int k = 1; k <= pver; ++k) {
      double dlfk = (k >= msg + 1) ? dl[k] : 0.0;
      rliq = rliq + (dlfk + 0.0) * in.dpp[cidx(c, k - 1, i, pver)] / P.gravit;
    }
    rliq = rliq / 1000.0;
  }
  o.prec[(size_t)c * pcols + i] = prec;
  o.rliq[(size_t)c * pcols + i] = rliq;
  o.jctop[(size_t)c * pcols + i] = (double)jt;
  o.jcbot[(size_t)c * pcols + i] = (double)maxg;
  // gathered outputs at gathered position gi of chunk c
  for (int k = 1; k <= pver; ++k) {
    size_t e = cidx(c, k - 1, gi, pver);
    o.mu[e] = C.mu[k]; o.md[e] = C.md[k]; o.du[e] = C.du[k]; o.eu[e] = C.eu[k]; o.ed[e] = C.ed[k];
    o.dp[e] = C.dp[k];
  }
  o.dsubcld[(size_t)c * pcols + gi] = dsubcld;
  o.jt[(size_t)c * pcols + gi] = jt;
  o.maxg[(size_t)c * pcols + gi] = maxg;
}
