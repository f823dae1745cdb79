/* zm_oracle.h -- TEST INFRASTRUCTURE.  C interface of the CPU oracle (oracle/zm_oracle.cpp):
 * a line-faithful restatement of /root/reference/physics/zm_conv.F90 (CAM-Nor ZM deep
 * convection).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.  The product library never links or calls this.
 *
 * Array conventions are the Fortran ones: every (pcols,pver[,n]) array is column-major,
 * element (i,k) at (i-1) + pcols*(k-1); indices stored in integer outputs are 1-based.
 */
#ifndef ZM_ORACLE_H
#define ZM_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct zmo_params {
  /* ppgrid */
  int pcols, pver;
  /* zm_convi arguments, zm_conv.F90:115-120 */
  int limcnv, num_cin;
  int zm_org, microp, no_deep_pbl, lparcel_pbl;
  /* cam_physpkg_is('cam3') (zm_conv.F90:871) and masterproc (zm_conv.F90:185,213) */
  int cam3, masterproc;
  double c0_lnd, c0_ocn, ke, ke_lnd, momcu, momcd;
  double tiedke_add, capelmt, dmpdz, tau;
  /* physconst, zm_conv.F90:19-20 */
  double cpair, epsilo, gravit, latice, latvap, tmelt, rair, cpwv, cpliq, rh2o, cpvir, zvir;
} zmo_params_t;

/* fills CAM6-default namelist values + shr_const physconst for the given grid */
void zmo_params_default(zmo_params_t* p, int pcols, int pver, int limcnv);
const char* zmo_math_backend(void);

/* zm_convi (zm_conv.F90:115).  Returns 0, or nonzero for the reference's endrun conditions. */
int zmo_convi(const zmo_params_t* p);

/* zm_convr (zm_conv.F90:231) -- argument order of the Fortran dummy list, minus
 * org/orgt/org2d (attached with zmo_org_fields when zm_org = 1) and conv/aero (zmconv_microp is out of scope).
 * Returns 0, or the number of Brent non-convergence events (reference: endrun). */
int zmo_convr(int lchnk, int ncol,
              const double* t, const double* qh, double* prec, double* jctop, double* jcbot,
              const double* pblh, const double* zm, const double* geos, const double* zi,
              double* qtnd, double* heat, const double* pap, const double* paph, const double* dpp,
              double delt, double* mcon, double* cme, double* cape, double* eurt,
              const double* tpert, double* dlf, double* pflx, double* zdu, double* rprd,
              double* mu, double* md, double* du, double* eu, double* ed,
              double* dp, double* dsubcld, int* jt, int* maxg, int* ideep, int* lengath,
              double* ql, double* rliq, const double* landfrac,
              double* dif, double* dnlf, double* dnif, double* rice);

/* buoyan_dilute (zm_conv.F90:4425) on already-converted inputs (p, pf in hPa; z absolute).
 * dmpdz is (pcols,pver). Used for stage-level parity tests. */
int zmo_buoyan_dilute(int lchnk, int ncol,
                      const double* q, const double* t, const double* p, const double* z,
                      const double* pf, double* tp, double* qstp, double* tl, double* cape,
                      double* cin, const double* pblt, int* lcl, int* lel, int* lon, int* mx,
                      const double* zi, const double* zs, const double* tpert,
                      const double* landfrac, const double* dmpdz);

/* zm_conv_evap (zm_conv.F90:1712); prdsnow absent (old_snow = .true.). */
void zmo_conv_evap(int ncol, int lchnk,
                   const double* t, const double* pmid, const double* pdel, const double* q,
                   const double* landfrac,
                   double* tend_s, double* tend_s_snwprd, double* tend_s_snwevmlt, double* tend_q,
                   const double* prdprec, const double* cldfrc, double deltat,
                   double* prec, double* snow, double* ntprprd, double* ntsnprd,
                   double* flxprec, double* flxsnow);

/* convtran (zm_conv.F90:1976).  doconvtran[ncnst] logical as int; cnst_is_dry[ncnst]
 * replaces the external cnst_get_type_byind(m).eq.'dry'. */
void zmo_convtran(int lchnk, const int* doconvtran, const double* q, int ncnst,
                  const double* mu, const double* md, const double* du, const double* eu,
                  const double* ed, const double* dp, const double* dsubcld,
                  const int* jt, const int* mx, const int* ideep, int il1g, int il2g,
                  int nstep, const double* fracis, double* dqdt, const double* dpdry, double dt,
                  const int* cnst_is_dry);

/* momtran (zm_conv.F90:2315). */
void zmo_momtran(int lchnk, int ncol, const int* domomtran, const double* q, int ncnst,
                 const double* mu, const double* md, const double* du, const double* eu,
                 const double* ed, const double* dp, const double* dsubcld,
                 const int* jt, const int* mx, const int* ideep, int il1g, int il2g,
                 int nstep, double* dqdt, double* pguall, double* pgdall,
                 double* icwu, double* icwd, double dt, double* seten);

/* scalar helpers exposed for property tests (zm_conv.F90:5280,5440,5304,5460,5421) */
double zmo_entropy(double tk, double p, double qtot);
double zmo_enthalpy(double tk, double p, double qtot, double z);
int zmo_ientropy(double s, double p, double qt, double tfg, double* t, double* qst);
int zmo_ienthalpy(double s, double p, double z, double qt, double tfg, double* t, double* qst);
void zmo_qsat_hpa(double t, double p, double* es, double* qm);
/* table qsat used by zm_conv_evap (external wv_saturation::qsat), p in Pa */
void zmo_qsat_table(double t, double p, double* es, double* qs);

/* operation counters for the roofline flop figure (per-thread; reset then read) */
/* zm_org = 1 (organisation tracer): attach org (in), orgt (out: zeroed by zm_convr, zm_conv.F90:555; after
 * zmo_conv_tend_batch it holds ptend_all%q(:,:,ixorg), zm_conv_intr.F90:773-777) and org2d (out, zm_conv.F90:793-819)
 * before calling zmo_convr / zmo_convr_batch / zmo_conv_tend_batch.  Shapes (pcols,pver) per chunk. */
void zmo_org_fields(const double* org, double* orgt, double* org2d);
void zmo_counters_reset(void);
/* optional per-inversion trace: 4 ints per call (rcall, icol, lchnk, state-function evaluations); single thread */
void zmo_trace_set(int* buf, int cap);
int zmo_trace_count(void);
/* out[0]=qsat_hPa calls, [1]=entropy, [2]=enthalpy, [3]=ientropy, [4]=ienthalpy,
 * [5]=log, [6]=log10, [7]=pow10, [8]=exp, [9]=pow  */
void zmo_counters_get(long long* out10);
/* FP64 operation count of the calling thread since zmo_counters_reset, 3 phases x 8 columns:
 * phase 0 everything outside the CAPE passes (transcendental calls only), 1 first buoyan_dilute call, 2 second;
 * columns: basic operations (+ - * / compare = 1 each), log, log10, 10**x, exp, x**y calls, state-function
 * evaluations, Brent iterations.  Run single-threaded (nthreads = 1) to see a whole batch. */
void zmo_flops_get(long long* out24);

/* chunk-loop drivers (OpenMP over chunks, mirroring physpkg.F90:1147) used for timing and
 * for whole-grid parity.  All arrays are [chunk][k][i] i.e. Fortran chunks back-to-back. */
int zmo_convr_batch(int nchunks, const int* ncol,
                    const double* t, const double* qh, double* prec, double* jctop, double* jcbot,
                    const double* pblh, const double* zm, const double* geos, const double* zi,
                    double* qtnd, double* heat, const double* pap, const double* paph,
                    const double* dpp, double delt, double* mcon, double* cme, double* cape,
                    double* eurt, const double* tpert, double* dlf, double* pflx, double* zdu,
                    double* rprd, double* mu, double* md, double* du, double* eu, double* ed,
                    double* dp, double* dsubcld, int* jt, int* maxg, int* ideep, int* lengath,
                    double* ql, double* rliq, const double* landfrac,
                    double* dif, double* dnlf, double* dnif, double* rice, int nthreads);

/* zm_conv_tend sequence (zm_conv_intr.F90:662-836) chunk by chunk under OpenMP: zm_convr ->
 * physics_update(state1) -> zm_conv_evap -> momtran -> ptend_all sums; mcon in kg/m2/s. */
int zmo_conv_tend_batch(int nchunks, const int* ncol, const double* t, const double* q, const double* u,
                        const double* v, const double* pmid, const double* pint, const double* pdel,
                        const double* zm, const double* zi, const double* phis, const double* pblh,
                        const double* tpert, const double* landfrac, const double* cld, double ztodt,
                        double* ptend_s, double* ptend_q, double* ptend_u, double* ptend_v, double* mcon,
                        double* cme, double* pflx, double* zdu, double* rliq, double* rice, double* jctop,
                        double* jcbot, double* prec, double* snow, double* ql, double* rprd, double* evapcdp,
                        double* flxprec, double* flxsnow, double* dlf, double* mu, double* md, double* du,
                        double* eu, double* ed, double* dp, double* dsubcld, int* jt, int* maxg, int* ideep,
                        int* lengath, double* cape, int nthreads);

/* zm_conv_tend_2 (zm_conv_intr.F90:955-1028): dpdry gather + convtran2, OpenMP over chunks */
int zmo_conv_tend_2_batch(int nchunks, const int* doconvtran, const double* q, int pcnst, const double* pdeldry,
                          const double* fracis, double* ptend_q, double ztodt, const int* cnst_is_dry,
                          const double* mu, const double* md, const double* du, const double* eu, const double* ed,
                          const double* dp, const double* dsubcld, const int* jt, const int* maxg, const int* ideep,
                          const int* lengath, int nthreads);
/* convtran1 inside zm_conv_tend (zm_conv_intr.F90:865-880): attach state%q, fracis and ptend_loc%q
 * ((pcols,pver,pcnst) per chunk, chunks back to back) and the constituent flags for the NEXT zmo_conv_tend_batch. */
void zmo_convtran1_fields(int pcnst, const int* doconvtran, const int* cnst_is_dry, const double* q,
                          const double* fracis, double* ptend_q);
/* zm_conv_tend's diagnostics with arithmetic (zm_conv_intr.F90:685-688, 700-706, 721-729), one chunk */
void zmo_conv_tend_diag(int ncol, const double* ps, const double* pmid, const double* mu, const double* md,
                        const int* jt, const int* maxg, const int* ideep, int lengath, double* freqzm,
                        double* mu_out, double* md_out, double* pcont, double* pconb);

/* N4 neighbours: geopotential_t (geopotential.F90:153-247; dycore_lr selects the FV 'LR'
 * branch, else EUL/SE) and convect_diagnostics_calc for shallow_scheme='CLUBB_SGS'
 * (convect_diagnostics.F90:115-249).  Single chunk, Fortran layout. */
void zmo_geopotential_t(int ncol, int dycore_lr, const double* piln, const double* pmln, const double* pint,
                        const double* pmid, const double* pdel, const double* rpdel, const double* t,
                        const double* q, const double* rair, double gravit, const double* zvir, double* zi,
                        double* zm);
/* geopotential_t, generalized-virtual-temperature branch (geopotential.F90:248-310; dycore MPAS / SE): q3 is
 * q(pcols,pver,ncnst), species_idx[nspecies] the 1-based thermodynamic_active_species_idx. */
void zmo_geopotential_t_gen(int ncol, int dycore_lr, int ncnst, int nspecies, const int* species_idx,
                            const double* piln, const double* pmln, const double* pint, const double* pmid,
                            const double* pdel, const double* rpdel, const double* t, const double* q3,
                            const double* rair, double gravit, const double* zvir, double* zi, double* zm);
void zmo_convect_diagnostics(int ncol, double* cmfmc, double* qc, double* qc2, double* rliq, double* rliq2,
                             const double* pmid, const double* rprddp, double* cnt, double* cnb,
                             double* cmfmc2, double* rprdsh, double* rprdtot, double* pcnt, double* pcnb);

#ifdef __cplusplus
}
#endif
#endif
