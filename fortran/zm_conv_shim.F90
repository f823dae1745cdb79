module zm_conv
!---------------------------------------------------------------------------------
! Drop-in replacement for NorESMhub/CAM-Nor-physics physics/zm_conv.F90: the five public
! procedures (zm_conv.F90:33-37) keep their names and dummy-argument lists, so
! zm_conv_intr.F90 (`use zm_conv, only: zm_conv_evap, zm_convr, convtran, momtran`, :13, and
! `zm_convi`, :252) compiles unchanged; the bodies forward to libzmconv_b200.so
! (include/zmconv_b200.h) through ISO_C_BINDING with nchunks = 1.
!
! NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no Fortran compiler (gfortran,
! flang, nvfortran all absent -- probed in the container and on the GPU box).  The file is the
! binding a CAM maintainer would add; see INTEGRATION.md for the batched (multi-chunk) variant
! that a GPU actually needs.
!---------------------------------------------------------------------------------
  use, intrinsic :: iso_c_binding
  use shr_kind_mod,    only: r8 => shr_kind_r8
  use spmd_utils,      only: masterproc
  use ppgrid,          only: pcols, pver, pverp
  use physconst,       only: cpair, epsilo, gravit, latice, latvap, tmelt, rair, &
                             cpwv, cpliq, rh2o, cpvir, zvir
  use cam_abortutils,  only: endrun
  use cam_logfile,     only: iulog
  use zm_microphysics, only: zm_aero_t, zm_conv_t

  implicit none
  save
  private

  public zm_convi, zm_convr, zm_conv_evap, convtran, momtran

  ! mirrors zm_params_t (include/zmconv_b200.h)
  type, bind(C) :: zm_params_t
     integer(c_int) :: pcols, pver, limcnv, num_cin
     integer(c_int) :: zm_org, microp, no_deep_pbl, lparcel_pbl, cam3, masterproc
     real(c_double) :: c0_lnd, c0_ocn, ke, ke_lnd, momcu, momcd
     real(c_double) :: tiedke_add, capelmt, dmpdz, tau
     real(c_double) :: cpair, epsilo, gravit, latice, latvap, tmelt, rair, cpwv, cpliq, rh2o, cpvir, zvir
  end type zm_params_t

  interface
     integer(c_int) function zm_init(p) bind(C, name='zm_init')
       import :: c_int, zm_params_t
       type(zm_params_t), intent(in) :: p
     end function zm_init

     integer(c_int) function zm_org_fields(org, orgt, org2d) bind(C, name='zm_org_fields')
       import :: c_int, c_double
       real(c_double), intent(in)  :: org(*)
       real(c_double), intent(out) :: orgt(*), org2d(*)
     end function zm_org_fields

     integer(c_int) function zm_last_error(buf, buflen) bind(C, name='zm_last_error')
       import :: c_int, c_char
       character(kind=c_char), intent(out) :: buf(*)
       integer(c_int), value :: buflen
     end function zm_last_error

     integer(c_int) function zm_convr_batch(nchunks, ncol, t, qh, prec, jctop, jcbot, pblh, zm, geos, zi, &
          qtnd, heat, pap, paph, dpp, delt, mcon, cme, cape, eurt, tpert, dlf, pflx, zdu, rprd, mu, md, du, &
          eu, ed, dp, dsubcld, jt, maxg, ideep, lengath, ql, rliq, landfrac, dif, dnlf, dnif, rice) &
          bind(C, name='zm_convr_batch')
       import :: c_int, c_double
       integer(c_int), value :: nchunks
       integer(c_int), intent(in) :: ncol(*)
       real(c_double), intent(in) :: t(*), qh(*), pblh(*), zm(*), geos(*), zi(*), pap(*), paph(*), dpp(*), &
                                     tpert(*), landfrac(*)
       real(c_double), value :: delt
       real(c_double), intent(out) :: prec(*), jctop(*), jcbot(*), qtnd(*), heat(*), mcon(*), cme(*), cape(*), &
                                      eurt(*), dlf(*), pflx(*), zdu(*), rprd(*), mu(*), md(*), du(*), eu(*), &
                                      ed(*), dp(*), dsubcld(*), ql(*), rliq(*), dif(*), dnlf(*), dnif(*), rice(*)
       integer(c_int), intent(out) :: jt(*), maxg(*), ideep(*), lengath(*)
     end function zm_convr_batch

     integer(c_int) function zm_conv_evap_batch(nchunks, ncol, t, pmid, pdel, q, landfrac, tend_s, &
          tend_s_snwprd, tend_s_snwevmlt, tend_q, prdprec, cldfrc, deltat, prec, snow, ntprprd, ntsnprd, &
          flxprec, flxsnow) bind(C, name='zm_conv_evap_batch')
       import :: c_int, c_double
       integer(c_int), value :: nchunks
       integer(c_int), intent(in) :: ncol(*)
       real(c_double), intent(in) :: t(*), pmid(*), pdel(*), q(*), landfrac(*), prdprec(*), cldfrc(*)
       real(c_double), value :: deltat
       real(c_double), intent(inout) :: tend_s(*), tend_q(*), prec(*)
       real(c_double), intent(out) :: tend_s_snwprd(*), tend_s_snwevmlt(*), snow(*), ntprprd(*), ntsnprd(*), &
                                      flxprec(*), flxsnow(*)
     end function zm_conv_evap_batch

     integer(c_int) function zm_momtran_batch(nchunks, ncol, domomtran, q, ncnst, mu, md, du, eu, ed, dp, &
          dsubcld, jt, mx, ideep, lengath, dqdt, pguall, pgdall, icwu, icwd, dt, seten) &
          bind(C, name='zm_momtran_batch')
       import :: c_int, c_double
       integer(c_int), value :: nchunks, ncnst
       integer(c_int), intent(in) :: ncol(*), domomtran(*), jt(*), mx(*), ideep(*), lengath(*)
       real(c_double), intent(in) :: q(*), mu(*), md(*), du(*), eu(*), ed(*), dp(*), dsubcld(*)
       real(c_double), value :: dt
       real(c_double), intent(inout) :: dqdt(*), icwu(*), icwd(*)
       real(c_double), intent(out) :: pguall(*), pgdall(*), seten(*)
     end function zm_momtran_batch

     integer(c_int) function zm_convtran_batch(nchunks, doconvtran, q, ncnst, mu, md, du, eu, ed, dp, dsubcld, &
          jt, mx, ideep, lengath, fracis, dqdt, dpdry, dt, cnst_is_dry) bind(C, name='zm_convtran_batch')
       import :: c_int, c_double
       integer(c_int), value :: nchunks, ncnst
       integer(c_int), intent(in) :: doconvtran(*), jt(*), mx(*), ideep(*), lengath(*), cnst_is_dry(*)
       real(c_double), intent(in) :: q(*), mu(*), md(*), du(*), eu(*), ed(*), dp(*), dsubcld(*), fracis(*), dpdry(*)
       real(c_double), value :: dt
       real(c_double), intent(inout) :: dqdt(*)
     end function zm_convtran_batch
  end interface

  logical :: zmconv_microp = .false.

  logical, save :: zm_org_on = .false.    ! zmconv_org as given to zm_convi (read-only afterwards)

contains

  subroutine zm_abort(where, rc)
    character(len=*), intent(in) :: where
    integer(c_int),   intent(in) :: rc
    character(kind=c_char) :: cbuf(1024)
    character(len=1024)    :: msg
    integer :: n, i
    n = zm_last_error(cbuf, 1024_c_int)
    msg = ' '
    do i = 1, min(n, 1023)
       msg(i:i) = cbuf(i)
    end do
    write(iulog,*) trim(where), ': rc=', rc, ' ', trim(msg)
    call endrun('**** ZM_CONV ('//trim(where)//') B200 library error ****')
  end subroutine zm_abort

  ! zm_conv.F90:115-120
  subroutine zm_convi(limcnv_in, zmconv_c0_lnd, zmconv_c0_ocn, zmconv_ke, zmconv_ke_lnd, &
                      zmconv_momcu, zmconv_momcd, zmconv_num_cin, zmconv_org, &
                      zmconv_microp_in, no_deep_pbl_in, zmconv_tiedke_add, &
                      zmconv_capelmt, zmconv_dmpdz, zmconv_parcel_pbl, zmconv_tau)
    use phys_control, only: cam_physpkg_is
    integer,  intent(in) :: limcnv_in, zmconv_num_cin
    real(r8), intent(in) :: zmconv_c0_lnd, zmconv_c0_ocn, zmconv_ke, zmconv_ke_lnd, zmconv_momcu, zmconv_momcd
    logical              :: zmconv_org
    logical,  intent(in) :: zmconv_microp_in, no_deep_pbl_in, zmconv_parcel_pbl
    real(r8), intent(in) :: zmconv_tiedke_add, zmconv_capelmt, zmconv_dmpdz, zmconv_tau
    type(zm_params_t) :: p
    integer(c_int)    :: rc

    p%pcols = pcols; p%pver = pver; p%limcnv = limcnv_in; p%num_cin = zmconv_num_cin
    p%zm_org = merge(1, 0, zmconv_org); p%microp = merge(1, 0, zmconv_microp_in)
    zm_org_on = zmconv_org
    p%no_deep_pbl = merge(1, 0, no_deep_pbl_in); p%lparcel_pbl = merge(1, 0, zmconv_parcel_pbl)
    p%cam3 = merge(1, 0, cam_physpkg_is('cam3')); p%masterproc = merge(1, 0, masterproc)
    p%c0_lnd = zmconv_c0_lnd; p%c0_ocn = zmconv_c0_ocn; p%ke = zmconv_ke; p%ke_lnd = zmconv_ke_lnd
    p%momcu = zmconv_momcu; p%momcd = zmconv_momcd; p%tiedke_add = zmconv_tiedke_add
    p%capelmt = zmconv_capelmt; p%dmpdz = zmconv_dmpdz; p%tau = zmconv_tau
    p%cpair = cpair; p%epsilo = epsilo; p%gravit = gravit; p%latice = latice; p%latvap = latvap
    p%tmelt = tmelt; p%rair = rair; p%cpwv = cpwv; p%cpliq = cpliq; p%rh2o = rh2o; p%cpvir = cpvir; p%zvir = zvir
    zmconv_microp = zmconv_microp_in
    rc = zm_init(p)
    if (rc /= 0) call zm_abort('zm_convi', rc)
  end subroutine zm_convi

  ! zm_conv.F90:231-244.  With zmconv_org the pointer dummies org/orgt/org2d (contiguous (pcols,pver) slices,
  ! zm_conv_intr.F90:656-659) are attached with zm_org_fields before the call; conv/aero are accepted and
  ! ignored (zmconv_microp must be .false., zm_init rejects it otherwise)
  subroutine zm_convr(lchnk, ncol, t, qh, prec, jctop, jcbot, pblh, zm, geos, zi, qtnd, heat, pap, paph, dpp, &
                      delt, mcon, cme, cape, eurt, tpert, dlf, pflx, zdu, rprd, mu, md, du, eu, ed, &
                      dp, dsubcld, jt, maxg, ideep, lengath, ql, rliq, landfrac, org, orgt, org2d, &
                      dif, dnlf, dnif, conv, aero, rice)
    integer,  intent(in)  :: lchnk, ncol
    real(r8), intent(in)  :: t(pcols,pver), qh(pcols,pver), pap(pcols,pver), paph(pcols,pver+1), dpp(pcols,pver), &
                             zm(pcols,pver), geos(pcols), zi(pcols,pver+1), pblh(pcols), tpert(pcols), landfrac(pcols)
    real(r8), intent(in)  :: delt
    type(zm_conv_t), intent(inout) :: conv
    type(zm_aero_t), intent(inout) :: aero
    real(r8), intent(out) :: qtnd(pcols,pver), heat(pcols,pver), mcon(pcols,pverp), dlf(pcols,pver), &
                             pflx(pcols,pverp), cme(pcols,pver), cape(pcols), zdu(pcols,pver), rprd(pcols,pver), &
                             dif(pcols,pver), dnlf(pcols,pver), dnif(pcols,pver), mu(pcols,pver), eu(pcols,pver), &
                             eurt(pcols,pver), du(pcols,pver), md(pcols,pver), ed(pcols,pver), dp(pcols,pver), &
                             dsubcld(pcols), jctop(pcols), jcbot(pcols), prec(pcols), rliq(pcols), rice(pcols), &
                             ql(pcols,pver)
    integer,  intent(out) :: ideep(pcols), lengath, jt(pcols), maxg(pcols)
    real(r8), pointer     :: org(:,:), orgt(:,:), org2d(:,:)
    integer(c_int) :: rc, nc(1), len1(1)

    nc(1) = ncol
    if (zm_org_on) then
       rc = zm_org_fields(org, orgt, org2d)
    end if
    rc = zm_convr_batch(1_c_int, nc, t, qh, prec, jctop, jcbot, pblh, zm, geos, zi, qtnd, heat, pap, paph, dpp, &
                        delt, mcon, cme, cape, eurt, tpert, dlf, pflx, zdu, rprd, mu, md, du, eu, ed, dp, &
                        dsubcld, jt, maxg, ideep, len1, ql, rliq, landfrac, dif, dnlf, dnif, rice)
    lengath = len1(1)
    if (rc /= 0) call zm_abort('zm_convr', rc)      ! rc > 0: Brent non-convergence == reference endrun
  end subroutine zm_convr

  ! zm_conv.F90:1712-1717 (prdsnow is only used with zmconv_microp; ignored)
  subroutine zm_conv_evap(ncol, lchnk, t, pmid, pdel, q, landfrac, tend_s, tend_s_snwprd, tend_s_snwevmlt, tend_q, &
                          prdprec, cldfrc, deltat, prec, snow, ntprprd, ntsnprd, flxprec, flxsnow, prdsnow)
    integer,  intent(in)    :: ncol, lchnk
    real(r8), intent(in)    :: t(pcols,pver), pmid(pcols,pver), pdel(pcols,pver), q(pcols,pver), landfrac(pcols)
    real(r8), intent(inout) :: tend_s(pcols,pver), tend_q(pcols,pver)
    real(r8), intent(out)   :: tend_s_snwprd(pcols,pver), tend_s_snwevmlt(pcols,pver)
    real(r8), intent(in)    :: prdprec(pcols,pver), cldfrc(pcols,pver), deltat
    real(r8), intent(inout) :: prec(pcols)
    real(r8), intent(out)   :: snow(pcols), ntprprd(pcols,pver), ntsnprd(pcols,pver), flxprec(pcols,pverp), &
                               flxsnow(pcols,pverp)
    real(r8), optional, intent(in), allocatable :: prdsnow(:,:)
    integer(c_int) :: rc, nc(1)
    nc(1) = ncol
    rc = zm_conv_evap_batch(1_c_int, nc, t, pmid, pdel, q, landfrac, tend_s, tend_s_snwprd, tend_s_snwevmlt, &
                            tend_q, prdprec, cldfrc, deltat, prec, snow, ntprprd, ntsnprd, flxprec, flxsnow)
    if (rc /= 0) call zm_abort('zm_conv_evap', rc)
  end subroutine zm_conv_evap

  ! zm_conv.F90:1976-1980
  subroutine convtran(lchnk, doconvtran, q, ncnst, mu, md, du, eu, ed, dp, dsubcld, jt, mx, ideep, il1g, il2g, &
                      nstep, fracis, dqdt, dpdry, dt)
    use constituents, only: cnst_get_type_byind
    integer,  intent(in)  :: lchnk, ncnst, il1g, il2g, nstep
    logical,  intent(in)  :: doconvtran(ncnst)
    real(r8), intent(in)  :: q(pcols,pver,ncnst), mu(pcols,pver), md(pcols,pver), du(pcols,pver), eu(pcols,pver), &
                             ed(pcols,pver), dp(pcols,pver), dsubcld(pcols), fracis(pcols,pver,ncnst), &
                             dpdry(pcols,pver), dt
    integer,  intent(in)  :: jt(pcols), mx(pcols), ideep(pcols)
    real(r8), intent(out) :: dqdt(pcols,pver,ncnst)
    integer(c_int) :: rc, len1(1), doit(ncnst), isdry(ncnst)
    integer :: m
    do m = 1, ncnst
       doit(m)  = merge(1, 0, doconvtran(m))
       isdry(m) = merge(1, 0, cnst_get_type_byind(m) .eq. 'dry')      ! zm_conv.F90:2087
    end do
    len1(1) = il2g                                                     ! il1g is always 1 at both call sites
    rc = zm_convtran_batch(1_c_int, doit, q, int(ncnst, c_int), mu, md, du, eu, ed, dp, dsubcld, jt, mx, ideep, &
                           len1, fracis, dqdt, dpdry, dt, isdry)
    if (rc /= 0) call zm_abort('convtran', rc)
  end subroutine convtran

  ! zm_conv.F90:2315-2319
  subroutine momtran(lchnk, ncol, domomtran, q, ncnst, mu, md, du, eu, ed, dp, dsubcld, jt, mx, ideep, il1g, il2g, &
                     nstep, dqdt, pguall, pgdall, icwu, icwd, dt, seten)
    integer,  intent(in)  :: lchnk, ncol, ncnst, il1g, il2g, nstep
    logical,  intent(in)  :: domomtran(ncnst)
    real(r8), intent(in)  :: q(pcols,pver,ncnst), mu(pcols,pver), md(pcols,pver), du(pcols,pver), eu(pcols,pver), &
                             ed(pcols,pver), dp(pcols,pver), dsubcld(pcols), dt
    integer,  intent(in)  :: jt(pcols), mx(pcols), ideep(pcols)
    real(r8), intent(out) :: dqdt(pcols,pver,ncnst), pguall(pcols,pver,ncnst), pgdall(pcols,pver,ncnst), &
                             icwu(pcols,pver,ncnst), icwd(pcols,pver,ncnst), seten(pcols,pver)
    integer(c_int) :: rc, nc(1), len1(1), doit(ncnst)
    integer :: m
    do m = 1, ncnst
       doit(m) = merge(1, 0, domomtran(m))
    end do
    nc(1) = ncol; len1(1) = il2g
    rc = zm_momtran_batch(1_c_int, nc, doit, q, int(ncnst, c_int), mu, md, du, eu, ed, dp, dsubcld, jt, mx, ideep, &
                          len1, dqdt, pguall, pgdall, icwu, icwd, dt, seten)
    if (rc /= 0) call zm_abort('momtran', rc)
  end subroutine momtran

end module zm_conv
