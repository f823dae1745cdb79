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

// One warp per gathered (convective) column, lane == level; the two order-dependent recurrences
// (in-cloud updraft wind bottom-up, downdraft wind top-down; zm_conv.F90:2562-2589) run redundantly
// on every lane in the reference's order, everything else is level parallel.
#define MOM_WARPS 2
enum MomArr { M_MU, M_MD, M_DU, M_EU, M_ED, M_DP, M_C, M_CHAT, M_CONU, M_COND, M_PGU, M_PGD, M_DCONDT,
              M_MFLUX0, M_MFLUX1, M_WIND00, M_WIND01, M_WINDF0, M_WINDF1, M_COUNT };
inline size_t momtran_smem_bytes(int pver) { return (size_t)MOM_WARPS * M_COUNT * (pver + 2) * sizeof(double); }

__global__ void __launch_bounds__(32 * MOM_WARPS)
k_momtran_w(MomArgs a) {
  extern __shared__ double sm_mom[];
  const int pcols = P.pcols, pver = P.pver;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int slot = blockIdx.x * MOM_WARPS + wib;
  if (slot >= a.nchunks * pcols) return;
  const int c = slot / pcols, gi = slot - c * pcols;
  if (gi >= a.lengath[c]) return;
  const int ld = pver + 2;
  double* B = sm_mom + (size_t)wib * M_COUNT * ld;
#define MS(n, k) B[(n) * ld + (k)]
#define MPAR(k, lo, hi) for (int k = (lo) + lane; k <= (hi); k += 32)
  const int ii = a.ideep[slot] - 1;            // ungathered column (0-based)
  const int mx = a.mx[slot];
  const int ktm = a.ktm[c], kbm = a.kbm[c];
  const double mbsth = 1.e-15, dt = a.dt;
  MPAR(k, 1, pver + 1) {
    if (k <= pver) {
      const size_t e = cidx(c, k - 1, gi, pver);
      MS(M_MU, k) = a.mu[e]; MS(M_MD, k) = a.md[e]; MS(M_DU, k) = a.du[e]; MS(M_EU, k) = a.eu[e];
      MS(M_ED, k) = a.ed[e]; MS(M_DP, k) = a.dp[e];
      MS(M_WIND00, k) = 0.0; MS(M_WIND01, k) = 0.0; MS(M_WINDF0, k) = 0.0; MS(M_WINDF1, k) = 0.0;
    }
    MS(M_MFLUX0, k) = 0.0; MS(M_MFLUX1, k) = 0.0;
  }
  __syncwarp();
  for (int m = 0; m < a.ncnst && m < 2; ++m) {
    if (!a.domom[m]) continue;
    const size_t mb = ((size_t)c * a.ncnst + m) * pver;       // base level index of constituent m
    const int MF = m ? M_MFLUX1 : M_MFLUX0, W0 = m ? M_WIND01 : M_WIND00, WF = m ? M_WINDF1 : M_WINDF0;
    MPAR(k, 1, pver) {
      const double v = a.q[(mb + k - 1) * pcols + ii];
      MS(M_C, k) = v; MS(W0, k) = v;
    }
    __syncwarp();
    MPAR(k, 1, pver) {
      const int km1 = max(1, k - 1), kp1 = min(pver, k + 1);
      const double ck = MS(M_C, k), cm = MS(M_C, km1);
      const double chat = 0.5 * (ck + cm);
      MS(M_CHAT, k) = chat; MS(M_CONU, k) = chat; MS(M_COND, k) = chat; MS(M_DCONDT, k) = 0.0;
      double pgu, pgd;
      if (k == 1) {
        pgu = 0.0; pgd = 0.0;
      } else if (k <= pver - 1) {
        const double cp1 = MS(M_C, kp1);
        const double mududp = (MS(M_MU, k) * (ck - cm) / MS(M_DP, km1) + MS(M_MU, kp1) * (cp1 - ck) / MS(M_DP, k));
        pgu = -P.momcu * 0.5 * mududp;
        const double mddudp = (MS(M_MD, k) * (ck - cm) / MS(M_DP, km1) + MS(M_MD, kp1) * (cp1 - ck) / MS(M_DP, k));
        pgd = -P.momcd * 0.5 * mddudp;
      } else {
        const double mududp = MS(M_MU, k) * (ck - cm) / MS(M_DP, km1);
        pgu = -P.momcu * mududp;
        const double mddudp = MS(M_MD, k) * (ck - cm) / MS(M_DP, km1);
        pgd = -P.momcd * mddudp;
      }
      MS(M_PGU, k) = pgu; MS(M_PGD, k) = pgd;
    }
    __syncwarp();
    {
      const int k = 2, km1 = 1, kk = pver;
      const double mupdudp = MS(M_MU, kk) + MS(M_DU, kk) * MS(M_DP, kk);
      if (mupdudp > mbsth)
        MS(M_CONU, kk) = (+MS(M_EU, kk) * MS(M_C, kk) * MS(M_DP, kk) + MS(M_PGU, kk) * MS(M_DP, kk)) / mupdudp;
      // operator precedence exactly as written in the reference (zm_conv.F90:2554)
      if (MS(M_MD, k) < -mbsth)
        MS(M_COND, k) = (-MS(M_ED, km1) * MS(M_C, km1) * MS(M_DP, km1)) - MS(M_PGD, km1) * MS(M_DP, km1) / MS(M_MD, k);
    }
    // the two recurrences are independent of each other: interleaved in one loop
    for (int s = 0; s < pver - 1; ++s) {
      const int kk = pver - 1 - s;                 // updraft: pver-1 .. 1
      if (kk >= 1) {
        const int kkp1 = min(pver, kk + 1);
        const double mupdudp = MS(M_MU, kk) + MS(M_DU, kk) * MS(M_DP, kk);
        if (mupdudp > mbsth)
          MS(M_CONU, kk) = (MS(M_MU, kkp1) * MS(M_CONU, kkp1) + MS(M_EU, kk) * MS(M_C, kk) * MS(M_DP, kk) +
                            MS(M_PGU, kk) * MS(M_DP, kk)) / mupdudp;
      }
      const int k = 3 + s;                         // downdraft: 3 .. pver
      if (k <= pver) {
        const int km1 = k - 1;
        if (MS(M_MD, k) < -mbsth)
          MS(M_COND, k) = (MS(M_MD, km1) * MS(M_COND, km1) - MS(M_ED, km1) * MS(M_C, km1) * MS(M_DP, km1) -
                           MS(M_PGD, km1) * MS(M_DP, km1)) / MS(M_MD, k);
      }
    }
    __syncwarp();
    MPAR(k, 1, pver) {
      double dc = 0.0;
      if (k >= ktm) {
        const int kp1 = min(pver, k + 1);
        dc = +(MS(M_MU, kp1) * (MS(M_CONU, kp1) - MS(M_CHAT, kp1)) - MS(M_MU, k) * (MS(M_CONU, k) - MS(M_CHAT, k)) +
               MS(M_MD, kp1) * (MS(M_COND, kp1) - MS(M_CHAT, kp1)) - MS(M_MD, k) * (MS(M_COND, k) - MS(M_CHAT, k))) / MS(M_DP, k);
      }
      if (k >= kbm && k == mx)
        dc = (1.0 / MS(M_DP, k)) * (-MS(M_MU, k) * (MS(M_CONU, k) - MS(M_CHAT, k)) - MS(M_MD, k) * (MS(M_COND, k) - MS(M_CHAT, k)));
      const size_t e = (mb + k - 1) * pcols + ii;
      a.dqdt[e] = dc;
      a.pguall[e] = -MS(M_PGU, k);
      a.pgdall[e] = -MS(M_PGD, k);
      a.icwu[e] = MS(M_CONU, k);
      a.icwd[e] = MS(M_COND, k);
      if (k >= ktm)
        MS(MF, k) = -MS(M_MU, k) * (MS(M_CONU, k) - MS(M_CHAT, k)) - MS(M_MD, k) * (MS(M_COND, k) - MS(M_CHAT, k));
    }
    __syncwarp();
    MPAR(k, ktm, pver) MS(WF, k) = MS(M_C, k) - (MS(MF, k + 1) - MS(MF, k)) * dt / MS(M_DP, k);
    __syncwarp();
  }
  // kinetic-energy dissipation heating (zm_conv.F90:2675-2712)
  MPAR(k, 1, pver) {
    double gset2 = 0.0;
    if (k >= ktm) {
      const int km1 = max(1, k - 1), kp1 = min(pver, k + 1);
      const double utop = (MS(M_WIND00, k) + MS(M_WIND00, km1)) / 2.0;
      const double vtop = (MS(M_WIND01, k) + MS(M_WIND01, km1)) / 2.0;
      const double ubot = (MS(M_WIND00, kp1) + MS(M_WIND00, k)) / 2.0;
      const double vbot = (MS(M_WIND01, kp1) + MS(M_WIND01, k)) / 2.0;
      const double fket = utop * MS(M_MFLUX0, k) + vtop * MS(M_MFLUX1, k);
      const double fkeb = ubot * MS(M_MFLUX0, k + 1) + vbot * MS(M_MFLUX1, k + 1);
      const double ketend_cons = (fket - fkeb) / MS(M_DP, k);
      const double ketend = ((MS(M_WINDF0, k) * MS(M_WINDF0, k) + MS(M_WINDF1, k) * MS(M_WINDF1, k)) -
                             (MS(M_WIND00, k) * MS(M_WIND00, k) + MS(M_WIND01, k) * MS(M_WIND01, k))) * 0.5 / dt;
      gset2 = ketend_cons - ketend;
    }
    a.seten[cidx(c, k - 1, ii, pver)] = gset2;
  }
#undef MS
#undef MPAR
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
