// zm_transport.cuh -- zm_conv_evap, momtran, convtran kernels (HBM-bound level scans).
//   zm_conv_evap  zm_conv.F90:1712-1972   thread per column, top-down scan, no per-level storage
//   momtran       zm_conv.F90:2315-2715   thread per gathered column, both wind components
//   convtran      zm_conv.F90:1976-2311   thread per (gathered column, constituent)
// The chunk-wide loop bounds ktm/kbm (zm_conv.F90:2076-2081, 2449-2454) are reproduced exactly
// by a tiny per-chunk reduction kernel so the level ranges match the reference bit for bit.
#pragma once
#include "zm_kernels.cuh"

// ---- zm_conv_evap ------------------------------------------------------------------------------
struct EvapArgs {
  int nchunks;
  const int* ncol;
  const double *t, *pmid, *pdel, *q, *landfrac, *prdprec, *cldfrc;
  double *tend_s, *tend_s_snwprd, *tend_s_snwevmlt, *tend_q, *prec, *snow, *ntprprd, *ntsnprd,
      *flxprec, *flxsnow;
  double deltat;
};

__global__ void __launch_bounds__(128)
k_conv_evap(EvapArgs a) {
  const int pcols = P.pcols, pver = P.pver, pverp = P.pverp;
  int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= a.nchunks * pcols) return;
  const int c = col / pcols, i = col - c * pcols;
  if (i >= a.ncol[c]) return;
  const double tmelt = P.tmelt, gravit = P.gravit, latice = P.latice, latvap = P.latvap;
  double prec = a.prec[col] * 1000.0;
  double flxprec = 0.0, flxsnow = 0.0, evpvint = 0.0;
  a.flxprec[cidx(c, 0, i, pverp)] = 0.0;
  a.flxsnow[cidx(c, 0, i, pverp)] = 0.0;
  for (int k = 1; k <= pver; ++k) {
    const size_t e = cidx(c, k - 1, i, pver);
    const double t = a.t[e], pmid = a.pmid[e], pdel = a.pdel[e], q = a.q[e], prdprec = a.prdprec[e],
                 cldfrc = a.cldfrc[e];
    double es, qs, fice, fsnow_conv;
    qsat_table(t, pmid, es, qs);
    cldfrc_fice(t, fice, fsnow_conv);
    double flxsntm, snowmlt;
    if (t > tmelt) { flxsntm = 0.0; snowmlt = flxsnow * gravit / pdel; }
    else           { flxsntm = flxsnow; snowmlt = 0.0; }
    double evplimit = fmax2(1.0 - q / (1.0 + q) / qs, 0.0);
    const double kemask = P.ke;
    double evpprec = kemask * (1.0 - cldfrc) * evplimit * sqrt(flxprec);
    evplimit = fmin2(evplimit, flxprec * gravit / pdel);
    evplimit = fmin2(evplimit, (prec - evpvint) * gravit / pdel);
    evpprec = fmin2(evplimit, evpprec);
    double evpsnow, work1, work2;
    if (flxprec > 0.0) {
      work1 = fmin2(fmax2(0.0, flxsntm / flxprec), 1.0);
      evpsnow = evpprec * work1;
    } else {
      evpsnow = 0.0;
    }
    evpvint = evpvint + evpprec * pdel / gravit;
    const double ntprprd = prdprec - evpprec;
    if (flxprec > 0.0) work1 = fmin2(fmax2(0.0, flxsnow / flxprec), 1.0);
    else work1 = 0.0;
    work2 = fmax2(fsnow_conv, work1);
    if (snowmlt > 0.0) work2 = 0.0;
    const double ntsnprd = prdprec * work2 - evpsnow - snowmlt;
    a.tend_s_snwprd[e] = prdprec * work2 * latice;
    a.tend_s_snwevmlt[e] = -(evpsnow + snowmlt) * latice;
    a.ntprprd[e] = ntprprd;
    a.ntsnprd[e] = ntsnprd;
    flxprec = flxprec + ntprprd * pdel / gravit;
    flxsnow = flxsnow + ntsnprd * pdel / gravit;
    flxprec = fmax2(flxprec, 0.0);
    flxsnow = fmax2(flxsnow, 0.0);
    a.flxprec[cidx(c, k, i, pverp)] = flxprec;
    a.flxsnow[cidx(c, k, i, pverp)] = flxsnow;
    a.tend_s[e] = -evpprec * latvap + ntsnprd * latice;
    a.tend_q[e] = evpprec;
  }
  a.prec[col] = flxprec / 1000.0;
  a.snow[col] = flxsnow / 1000.0;
}

// ---- chunk-wide ktm / kbm (zm_conv.F90:2076-2081): one warp per chunk ----------------------------
__global__ void k_chunk_bounds(int nchunks, const int* jt, const int* mx, const int* lengath, int* ktm,
                               int* kbm) {
  const int lane = threadIdx.x & 31;
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (c >= nchunks) return;
  const int pcols = P.pcols, n = lengath[c];
  int a = P.pver, b = P.pver;
  for (int i = lane; i < n; i += 32) {
    a = min(a, jt[(size_t)c * pcols + i]);
    b = min(b, mx[(size_t)c * pcols + i]);
  }
  for (int off = 16; off; off >>= 1) {
    a = min(a, __shfl_xor_sync(0xffffffffu, a, off));
    b = min(b, __shfl_xor_sync(0xffffffffu, b, off));
  }
  if (lane == 0) { ktm[c] = a; kbm[c] = b; }
}

// ---- momtran -------------------------------------------------------------------------------------
struct MomArgs {
  int nchunks, ncnst;
  const int *ncol, *jt, *mx, *ideep, *lengath, *ktm, *kbm;
  int domom[2];
  const double *q, *mu, *md, *du, *eu, *ed, *dp;
  double *dqdt, *pguall, *pgdall, *icwu, *icwd, *seten;
  double dt;
};

// initialisation of the outgoing fields (zm_conv.F90:2429-2443, 2630)
__global__ void k_momtran_init(MomArgs a) {
  const int pcols = P.pcols, pver = P.pver;
  const size_t n2 = (size_t)a.nchunks * pcols * pver;
  size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
  for (size_t e = tid; e < n2 * a.ncnst; e += nth) {
    // e -> (c, m, k, i)
    size_t i = e % pcols, r = e / pcols;
    size_t m = (r / pver) % a.ncnst, c = r / ((size_t)pver * a.ncnst);
    a.pguall[e] = 0.0; a.pgdall[e] = 0.0;
    if ((int)i < a.ncol[c]) { a.icwu[e] = a.q[e]; a.icwd[e] = a.q[e]; }
    if (m < 2 && a.domom[m]) a.dqdt[e] = 0.0;
  }
  for (size_t e = tid; e < n2; e += nth) a.seten[e] = 0.0;
}

template <int LMAX>
__global__ void __launch_bounds__(64)
k_momtran(MomArgs a) {
  const int pcols = P.pcols, pver = P.pver;
  int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= a.nchunks * pcols) return;
  const int c = slot / pcols, gi = slot - c * pcols;
  if (gi >= a.lengath[c]) return;
  const int ii = a.ideep[slot] - 1;            // ungathered column (0-based)
  const int jt = a.jt[slot], mx = a.mx[slot];
  (void)jt;
  const int ktm = a.ktm[c], kbm = a.kbm[c];
  const double mbsth = 1.e-15, dt = a.dt;
  double mu[LMAX + 2], md[LMAX + 2], du[LMAX + 2], eu[LMAX + 2], ed[LMAX + 2], dp[LMAX + 2];
  double cnst[LMAX + 2], chat[LMAX + 2], conu[LMAX + 2], cond[LMAX + 2], pgu[LMAX + 2], pgd[LMAX + 2],
      dcondt[LMAX + 2];
  double mflux[2][LMAX + 3], wind0[2][LMAX + 2], windf[2][LMAX + 2];
  for (int k = 1; k <= pver; ++k) {
    size_t e = cidx(c, k - 1, gi, pver);
    mu[k] = a.mu[e]; md[k] = a.md[e]; du[k] = a.du[e]; eu[k] = a.eu[e]; ed[k] = a.ed[e]; dp[k] = a.dp[e];
  }
  for (int m = 0; m < 2; ++m)
    for (int k = 1; k <= pver + 1; ++k) { mflux[m][k] = 0.0; if (k <= pver) { wind0[m][k] = 0.0; windf[m][k] = 0.0; } }

  for (int m = 0; m < a.ncnst && m < 2; ++m) {
    if (!a.domom[m]) continue;
    const size_t mb = ((size_t)c * a.ncnst + m) * pver;       // base level index of constituent m
    for (int k = 1; k <= pver; ++k) {
      cnst[k] = a.q[(mb + k - 1) * pcols + ii];
      wind0[m][k] = cnst[k];
    }
    for (int k = 1; k <= pver; ++k) {
      int km1 = max(1, k - 1);
      chat[k] = 0.5 * (cnst[k] + cnst[km1]);
      conu[k] = chat[k];
      cond[k] = chat[k];
      dcondt[k] = 0.0;
    }
    pgu[1] = 0.0; pgd[1] = 0.0;
    for (int k = 2; k <= pver - 1; ++k) {
      int km1 = max(1, k - 1), kp1 = min(pver, k + 1);
      double mududp = (mu[k] * (cnst[k] - cnst[km1]) / dp[km1] + mu[kp1] * (cnst[kp1] - cnst[k]) / dp[k]);
      pgu[k] = -P.momcu * 0.5 * mududp;
      double mddudp = (md[k] * (cnst[k] - cnst[km1]) / dp[km1] + md[kp1] * (cnst[kp1] - cnst[k]) / dp[k]);
      pgd[k] = -P.momcd * 0.5 * mddudp;
    }
    {
      int k = pver, km1 = max(1, k - 1);
      double mududp = mu[k] * (cnst[k] - cnst[km1]) / dp[km1];
      pgu[k] = -P.momcu * mududp;
      double mddudp = md[k] * (cnst[k] - cnst[km1]) / dp[km1];
      pgd[k] = -P.momcd * mddudp;
    }
    {
      int k = 2, km1 = 1, kk = pver;
      double mupdudp = mu[kk] + du[kk] * dp[kk];
      if (mupdudp > mbsth) conu[kk] = (+eu[kk] * cnst[kk] * dp[kk] + pgu[kk] * dp[kk]) / mupdudp;
      // operator precedence exactly as written in the reference (zm_conv.F90:2554)
      if (md[k] < -mbsth) cond[k] = (-ed[km1] * cnst[km1] * dp[km1]) - pgd[km1] * dp[km1] / md[k];
    }
    for (int kk = pver - 1; kk >= 1; --kk) {
      int kkp1 = min(pver, kk + 1);
      double mupdudp = mu[kk] + du[kk] * dp[kk];
      if (mupdudp > mbsth)
        conu[kk] = (mu[kkp1] * conu[kkp1] + eu[kk] * cnst[kk] * dp[kk] + pgu[kk] * dp[kk]) / mupdudp;
    }
    for (int k = 3; k <= pver; ++k) {
      int km1 = max(1, k - 1);
      if (md[k] < -mbsth)
        cond[k] = (md[km1] * cond[km1] - ed[km1] * cnst[km1] * dp[km1] - pgd[km1] * dp[km1]) / md[k];
    }
    for (int k = ktm; k <= pver; ++k) {
      int kp1 = min(pver, k + 1);
      dcondt[k] = +(mu[kp1] * (conu[kp1] - chat[kp1]) - mu[k] * (conu[k] - chat[k]) +
                    md[kp1] * (cond[kp1] - chat[kp1]) - md[k] * (cond[k] - chat[k])) / dp[k];
    }
    for (int k = kbm; k <= pver; ++k)
      if (k == mx)
        dcondt[k] = (1.0 / dp[k]) * (-mu[k] * (conu[k] - chat[k]) - md[k] * (cond[k] - chat[k]));
    for (int k = 1; k <= pver; ++k) {
      size_t e = (mb + k - 1) * pcols + ii;
      a.dqdt[e] = dcondt[k];
      a.pguall[e] = -pgu[k];
      a.pgdall[e] = -pgd[k];
      a.icwu[e] = conu[k];
      a.icwd[e] = cond[k];
    }
    for (int k = ktm; k <= pver; ++k)
      mflux[m][k] = -mu[k] * (conu[k] - chat[k]) - md[k] * (cond[k] - chat[k]);
    for (int k = ktm; k <= pver; ++k)
      windf[m][k] = cnst[k] - (mflux[m][k + 1] - mflux[m][k]) * dt / dp[k];
  }
  // kinetic-energy dissipation heating (zm_conv.F90:2675-2712)
  for (int k = 1; k <= pver; ++k) {
    double gset2 = 0.0;
    if (k >= ktm) {
      int km1 = max(1, k - 1), kp1 = min(pver, k + 1);
      double utop = (wind0[0][k] + wind0[0][km1]) / 2.0;
      double vtop = (wind0[1][k] + wind0[1][km1]) / 2.0;
      double ubot = (wind0[0][kp1] + wind0[0][k]) / 2.0;
      double vbot = (wind0[1][kp1] + wind0[1][k]) / 2.0;
      double fket = utop * mflux[0][k] + vtop * mflux[1][k];
      double fkeb = ubot * mflux[0][k + 1] + vbot * mflux[1][k + 1];
      double ketend_cons = (fket - fkeb) / dp[k];
      double ketend = ((windf[0][k] * windf[0][k] + windf[1][k] * windf[1][k]) -
                       (wind0[0][k] * wind0[0][k] + wind0[1][k] * wind0[1][k])) * 0.5 / dt;
      gset2 = ketend_cons - ketend;
    }
    a.seten[cidx(c, k - 1, ii, pver)] = gset2;
  }
}

// ---- convtran --------------------------------------------------------------------------------------
struct TranArgs {
  int nchunks, ncnst, nactive;
  const int *jt, *mx, *ideep, *lengath, *ktm, *kbm;
  const int* active;        // [nactive] 0-based constituent indices with doconvtran (m >= 2)
  const int* is_dry;        // [ncnst]
  const double *q, *fracis, *mu, *md, *du, *eu, *ed, *dp, *dpdry;
  double* dqdt;
};

// dqdt(:,:,m) = 0 for every active constituent (zm_conv.F90:2298)
__global__ void k_convtran_zero(TranArgs a) {
  const int pcols = P.pcols, pver = P.pver;
  const size_t per = (size_t)pcols * pver;
  size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
  const size_t total = (size_t)a.nchunks * a.nactive * per;
  for (size_t e = tid; e < total; e += nth) {
    size_t r = e / per, off = e - r * per;
    size_t c = r / a.nactive, j = r - c * a.nactive;
    a.dqdt[((size_t)c * a.ncnst + a.active[j]) * per + off] = 0.0;
  }
}

// blockDim = (TX gathered columns, TY constituents); grid.x over column slots, grid.y over constituents
template <int LMAX>
__global__ void __launch_bounds__(128)
k_convtran(TranArgs a) {
  const int pcols = P.pcols, pver = P.pver;
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  if (slot >= a.nchunks * pcols || j >= a.nactive) return;
  const int c = slot / pcols, gi = slot - c * pcols;
  if (gi >= a.lengath[c]) return;
  const int m = a.active[j];
  const int ii = a.ideep[slot] - 1;
  const int mx = a.mx[slot];
  const int ktm = a.ktm[c], kbm = a.kbm[c];
  const double small = 1.e-36, mbsth = 1.e-15;
  const bool dry = a.is_dry[m] != 0;
  double mu[LMAX + 2], md[LMAX + 2], dutmp[LMAX + 2], eutmp[LMAX + 2], edtmp[LMAX + 2], dptmp[LMAX + 2];
  double cnst[LMAX + 2], fisg[LMAX + 2], chat[LMAX + 2], conu[LMAX + 2], cond[LMAX + 2], dcondt[LMAX + 2];
  const size_t mb = ((size_t)c * a.ncnst + m) * pver;
  for (int k = 1; k <= pver; ++k) {
    size_t e = cidx(c, k - 1, gi, pver);
    mu[k] = a.mu[e]; md[k] = a.md[e];
    double du = a.du[e], eu = a.eu[e], ed = a.ed[e], dp = a.dp[e];
    if (dry) {
      double dpd = a.dpdry[e];
      dptmp[k] = dpd;
      dutmp[k] = du * dp / dpd;
      eutmp[k] = eu * dp / dpd;
      edtmp[k] = ed * dp / dpd;
    } else {
      dptmp[k] = dp; dutmp[k] = du; eutmp[k] = eu; edtmp[k] = ed;
    }
    cnst[k] = a.q[(mb + k - 1) * pcols + ii];
    fisg[k] = a.fracis[(mb + k - 1) * pcols + ii];
  }
  for (int k = 1; k <= pver; ++k) {
    int km1 = max(1, k - 1);
    double minc = fmin2(cnst[km1], cnst[k]);
    double maxc = fmax2(cnst[km1], cnst[k]);
    double cdifr;
    if (minc < 0.0) cdifr = 0.0;
    else cdifr = fabs(cnst[k] - cnst[km1]) / fmax2(maxc, small);
    if (cdifr > 1.E-6) {
      double cabv = fmax2(cnst[km1], maxc * 1.e-12);
      double cbel = fmax2(cnst[k], maxc * 1.e-12);
      chat[k] = zmm::log_(cabv / cbel) / (cabv - cbel) * cabv * cbel;
    } else {
      chat[k] = 0.5 * (cnst[k] + cnst[km1]);
    }
    conu[k] = chat[k];
    cond[k] = chat[k];
    dcondt[k] = 0.0;
  }
  {
    int k = 2, km1 = 1, kk = pver;
    double mupdudp = mu[kk] + dutmp[kk] * dptmp[kk];
    if (mupdudp > mbsth) conu[kk] = (+eutmp[kk] * fisg[kk] * cnst[kk] * dptmp[kk]) / mupdudp;
    if (md[k] < -mbsth) cond[k] = (-edtmp[km1] * fisg[km1] * cnst[km1] * dptmp[km1]) / md[k];
  }
  for (int kk = pver - 1; kk >= 1; --kk) {
    int kkp1 = min(pver, kk + 1);
    double mupdudp = mu[kk] + dutmp[kk] * dptmp[kk];
    if (mupdudp > mbsth)
      conu[kk] = (mu[kkp1] * conu[kkp1] + eutmp[kk] * fisg[kk] * cnst[kk] * dptmp[kk]) / mupdudp;
  }
  for (int k = 3; k <= pver; ++k) {
    int km1 = max(1, k - 1);
    if (md[k] < -mbsth)
      cond[k] = (md[km1] * cond[km1] - edtmp[km1] * fisg[km1] * cnst[km1] * dptmp[km1]) / md[k];
  }
  for (int k = ktm; k <= pver; ++k) {
    int km1 = max(1, k - 1), kp1 = min(pver, k + 1);
    double fluxin = mu[kp1] * conu[kp1] + mu[k] * fmin2(chat[k], cnst[km1]) -
                    (md[k] * cond[k] + md[kp1] * fmin2(chat[kp1], cnst[kp1]));
    double fluxout = mu[k] * conu[k] + mu[kp1] * fmin2(chat[kp1], cnst[k]) -
                     (md[kp1] * cond[kp1] + md[k] * fmin2(chat[k], cnst[k]));
    double netflux = fluxin - fluxout;
    if (fabs(netflux) < fmax2(fluxin, fluxout) * 1.e-12) netflux = 0.0;
    dcondt[k] = netflux / dptmp[k];
  }
  for (int k = kbm; k <= pver; ++k) {
    int km1 = max(1, k - 1);
    if (k == mx) {
      double fluxin = mu[k] * fmin2(chat[k], cnst[km1]) - md[k] * cond[k];
      double fluxout = mu[k] * conu[k] - md[k] * fmin2(chat[k], cnst[k]);
      double netflux = fluxin - fluxout;
      if (fabs(netflux) < fmax2(fluxin, fluxout) * 1.e-12) netflux = 0.0;
      dcondt[k] = netflux / dptmp[k];
    } else if (k > mx) {
      dcondt[k] = 0.0;
    }
  }
  for (int k = 1; k <= pver; ++k) a.dqdt[(mb + k - 1) * pcols + ii] = dcondt[k];
}
