module zm_conv_intr_batched
!---------------------------------------------------------------------------------
! Batched (all chunks of a rank in one call) binding of zm_conv_tend / zm_conv_tend_2
! (NorESMhub/CAM-Nor-physics physics/zm_conv_intr.F90:390-951, 955-1028) to libzmconv_b200.so
! (include/zmconv_b200.h).  This is the entry a GPU needs: the per-chunk shim (zm_conv_shim.F90) feeds it 16
! columns per call.
!
! How it is used (INTEGRATION.md section 2): tphysbc's chunk loop (physpkg.F90:1147-1161, body :2808-2868) is
! split around convect_deep_tend.  Loop 1 runs every chunk up to dadadj_tend and calls zm_batch_gather; one call
! of zm_conv_tend_batched runs zm_convr -> physics_update -> zm_conv_evap -> momtran -> convtran1 for all chunks
! on the device; loop 2 calls zm_batch_scatter for its chunk (ptend_all, the dummy outputs of zm_conv_tend, the
! pbuf fields, the history diagnostics) and carries on with physics_update / check_energy_chng.
! zm_conv_tend_2_batched does the same for zm_conv_tend_2 (called from tphysac).
!
! The five GPTL timers of the reference (t_startf/t_stopf 'zm_convr', 'zm_conv_evap', 'momtran', 'convtran1',
! zm_conv_intr.F90:654-880, and 'convtran2', :1019-1025) cannot bracket host code any more -- the phases run
! back to back on the device -- so zm_batch_report_timers reads the device times of the phases back with
! zm_get_timers under the same names (zm_set_profiling(1) switches the event recording on) and writes them to
! the log; the wall-clock timers 'zm_conv_tend_batch' and 'zm_conv_tend_2_batch' bracket the calls.
!
! NOT COMPILED IN THIS REPOSITORY: the build image has no Fortran compiler.  tests/test_fortran_binding.py
! checks every bind(C) interface of this file and of zm_conv_shim.F90 against include/zmconv_b200.h with the C
! compiler: argument count, order (by name), types, const-ness, pass-by-value.
!---------------------------------------------------------------------------------
  use, intrinsic :: iso_c_binding
  use shr_kind_mod,    only: r8 => shr_kind_r8
  use ppgrid,          only: pcols, pver, pverp
  use physconst,       only: gravit, cpair
  use constituents,    only: pcnst, cnst_is_convtran1, cnst_is_convtran2, cnst_get_type_byind, cnst_get_ind
  use cam_abortutils,  only: endrun
  use cam_logfile,     only: iulog
  use cam_history,     only: outfld
  use perf_mod,        only: t_startf, t_stopf

  implicit none
  save
  private

  public :: zm_batch_init, zm_batch_gather, zm_conv_tend_batched, zm_batch_scatter
  public :: zm_batch_gather_2, zm_conv_tend_2_batched, zm_batch_scatter_2, zm_batch_report_timers

  ! ---- rank-level arrays: the chunks' (pcols,pver) fields back to back, [chunk][k][i] in C terms ----
  integer :: nchunks_b = 0
  integer(c_int),  allocatable :: ncol_b(:)
  real(c_double), allocatable, dimension(:,:,:) :: t_b, q_b, u_b, v_b, pmid_b, pdel_b, zm_b, cld_b        ! (pcols,pver,nchunks)
  real(c_double), allocatable, dimension(:,:,:) :: pint_b, zi_b                                            ! (pcols,pverp,nchunks)
  real(c_double), allocatable, dimension(:,:)   :: phis_b, pblh_b, tpert_b, landfrac_b, ps_b              ! (pcols,nchunks)
  real(c_double), allocatable, dimension(:,:,:) :: ptend_s_b, ptend_q_b, ptend_u_b, ptend_v_b, cme_b, zdu_b, &
                                                    ql_b, rprd_b, evapcdp_b, dlf_b, mu_b, md_b, du_b, eu_b, ed_b, dp_b, &
                                                    mu_out_b, md_out_b
  real(c_double), allocatable, dimension(:,:,:) :: mcon_b, pflx_b, flxprec_b, flxsnow_b
  real(c_double), allocatable, dimension(:,:)   :: rliq_b, rice_b, jctop_b, jcbot_b, prec_b, snow_b, dsubcld_b, cape_b, &
                                                    freqzm_b, pcont_b, pconb_b
  integer(c_int),  allocatable, dimension(:,:)   :: jt_b, maxg_b, ideep_b
  integer(c_int),  allocatable :: lengath_b(:)
  real(c_double), allocatable, dimension(:,:,:,:) :: q3_b, fracis_b, ptend_q3_b                            ! (pcols,pver,pcnst,nchunks)
  real(c_double), allocatable, dimension(:,:,:) :: pdeldry_b
  integer(c_int) :: doconvtran1(pcnst), doconvtran2(pcnst), cnst_is_dry(pcnst)

  interface
     integer(c_int) function zm_conv_tend_batch(nchunks, ncol, t, q, u, v, pmid, pint, pdel, zm, zi, phis, pblh, &
          tpert, landfrac, cld, ztodt, ptend_s, ptend_q, ptend_u, ptend_v, mcon, cme, pflx, zdu, rliq, rice, &
          jctop, jcbot, prec, snow, ql, rprd, evapcdp, flxprec, flxsnow, dlf, mu, md, du, eu, ed, dp, dsubcld, &
          jt, maxg, ideep, lengath, cape) bind(C, name='zm_conv_tend_batch')
       import :: c_int, c_double
       integer(c_int), value :: nchunks
       integer(c_int), intent(in) :: ncol(*)
       real(c_double), intent(in) :: t(*), q(*), u(*), v(*), pmid(*), pint(*), pdel(*), zm(*), zi(*), phis(*), &
                                     pblh(*), tpert(*), landfrac(*), cld(*)
       real(c_double), value :: ztodt
       real(c_double), intent(out) :: ptend_s(*), ptend_q(*), ptend_u(*), ptend_v(*), mcon(*), cme(*), pflx(*), &
                                      zdu(*), rliq(*), rice(*), jctop(*), jcbot(*), prec(*), snow(*), ql(*), rprd(*), &
                                      evapcdp(*), flxprec(*), flxsnow(*), dlf(*), mu(*), md(*), du(*), eu(*), ed(*), &
                                      dp(*), dsubcld(*)
       integer(c_int), intent(out) :: jt(*), maxg(*), ideep(*), lengath(*)
       real(c_double), intent(out) :: cape(*)
     end function zm_conv_tend_batch

     integer(c_int) function zm_conv_tend_2_batch(nchunks, doconvtran, q, pcnst, pdeldry, fracis, ptend_q, ztodt, &
          cnst_is_dry) bind(C, name='zm_conv_tend_2_batch')
       import :: c_int, c_double
       integer(c_int), value :: nchunks
       integer(c_int), intent(in) :: doconvtran(*)
       real(c_double), intent(in) :: q(*)
       integer(c_int), value :: pcnst
       real(c_double), intent(in) :: pdeldry(*), fracis(*)
       real(c_double), intent(inout) :: ptend_q(*)
       real(c_double), value :: ztodt
       integer(c_int), intent(in) :: cnst_is_dry(*)
     end function zm_conv_tend_2_batch

     integer(c_int) function zm_convtran1_fields(pcnst, doconvtran, cnst_is_dry, q, fracis, ptend_q) &
          bind(C, name='zm_convtran1_fields')
       import :: c_int, c_double
       integer(c_int), value :: pcnst
       integer(c_int), intent(in) :: doconvtran(*), cnst_is_dry(*)
       real(c_double), intent(in) :: q(*), fracis(*)
       real(c_double), intent(inout) :: ptend_q(*)
     end function zm_convtran1_fields

     integer(c_int) function zm_org_fields(org, orgt, org2d) bind(C, name='zm_org_fields')
       import :: c_int, c_double
       real(c_double), intent(in)  :: org(*)
       real(c_double), intent(out) :: orgt(*), org2d(*)
     end function zm_org_fields

     integer(c_int) function zm_conv_tend_diag_batch(nchunks, ncol, ps, pmid, mu, md, jt, maxg, ideep, lengath, &
          freqzm, mu_out, md_out, pcont, pconb) bind(C, name='zm_conv_tend_diag_batch')
       import :: c_int, c_double
       integer(c_int), value :: nchunks
       integer(c_int), intent(in) :: ncol(*)
       real(c_double), intent(in) :: ps(*), pmid(*), mu(*), md(*)
       integer(c_int), intent(in) :: jt(*), maxg(*), ideep(*), lengath(*)
       real(c_double), intent(out) :: freqzm(*), mu_out(*), md_out(*), pcont(*), pconb(*)
     end function zm_conv_tend_diag_batch

     integer(c_int) function zm_get_timers(n, names, ms) bind(C, name='zm_get_timers')
       import :: c_int, c_ptr, c_float
       integer(c_int), intent(inout) :: n
       type(c_ptr), intent(out) :: names(*)
       real(c_float), intent(out) :: ms(*)
     end function zm_get_timers

     integer(c_int) function zm_set_profiling(on) bind(C, name='zm_set_profiling')
       import :: c_int
       integer(c_int), value :: on
     end function zm_set_profiling

     integer(c_int) function zm_last_error(buf, buflen) bind(C, name='zm_last_error')
       import :: c_int, c_char
       character(kind=c_char), intent(out) :: buf(*)
       integer(c_int), value :: buflen
     end function zm_last_error
  end interface

contains

  !-------------------------------------------------------------------------------
  subroutine zm_batch_init(nchunks)
    ! once, after zm_conv_init (zm_convi -> zm_init is done by the shim's zm_convi, zm_conv_intr.F90:376-379)
    integer, intent(in) :: nchunks
    integer :: m
    nchunks_b = nchunks
    allocate(ncol_b(nchunks), lengath_b(nchunks))
    allocate(t_b(pcols,pver,nchunks), q_b(pcols,pver,nchunks), u_b(pcols,pver,nchunks), v_b(pcols,pver,nchunks), &
             pmid_b(pcols,pver,nchunks), pdel_b(pcols,pver,nchunks), zm_b(pcols,pver,nchunks), cld_b(pcols,pver,nchunks), &
             pint_b(pcols,pverp,nchunks), zi_b(pcols,pverp,nchunks), phis_b(pcols,nchunks), pblh_b(pcols,nchunks), &
             tpert_b(pcols,nchunks), landfrac_b(pcols,nchunks), ps_b(pcols,nchunks))
    allocate(ptend_s_b(pcols,pver,nchunks), ptend_q_b(pcols,pver,nchunks), ptend_u_b(pcols,pver,nchunks), &
             ptend_v_b(pcols,pver,nchunks), cme_b(pcols,pver,nchunks), zdu_b(pcols,pver,nchunks), ql_b(pcols,pver,nchunks), &
             rprd_b(pcols,pver,nchunks), evapcdp_b(pcols,pver,nchunks), dlf_b(pcols,pver,nchunks), mu_b(pcols,pver,nchunks), &
             md_b(pcols,pver,nchunks), du_b(pcols,pver,nchunks), eu_b(pcols,pver,nchunks), ed_b(pcols,pver,nchunks), &
             dp_b(pcols,pver,nchunks), mu_out_b(pcols,pver,nchunks), md_out_b(pcols,pver,nchunks))
    allocate(mcon_b(pcols,pverp,nchunks), pflx_b(pcols,pverp,nchunks), flxprec_b(pcols,pverp,nchunks), &
             flxsnow_b(pcols,pverp,nchunks))
    allocate(rliq_b(pcols,nchunks), rice_b(pcols,nchunks), jctop_b(pcols,nchunks), jcbot_b(pcols,nchunks), &
             prec_b(pcols,nchunks), snow_b(pcols,nchunks), dsubcld_b(pcols,nchunks), cape_b(pcols,nchunks), &
             freqzm_b(pcols,nchunks), pcont_b(pcols,nchunks), pconb_b(pcols,nchunks))
    allocate(jt_b(pcols,nchunks), maxg_b(pcols,nchunks), ideep_b(pcols,nchunks))
    allocate(q3_b(pcols,pver,pcnst,nchunks), fracis_b(pcols,pver,pcnst,nchunks), ptend_q3_b(pcols,pver,pcnst,nchunks), &
             pdeldry_b(pcols,pver,nchunks))
    ! constituent flags: lq(2:) = cnst_is_convtran1(2:) (zm_conv_intr.F90:868), cnst_is_convtran2 (:1009-1010),
    ! cnst_get_type_byind(m) == 'dry' (zm_conv.F90:2087)
    doconvtran1 = 0; doconvtran2 = 0; cnst_is_dry = 0
    do m = 2, pcnst
       if (cnst_is_convtran1(m)) doconvtran1(m) = 1
       if (cnst_is_convtran2(m)) doconvtran2(m) = 1
       if (cnst_get_type_byind(m) == 'dry') cnst_is_dry(m) = 1
    end do
  end subroutine zm_batch_init

  !-------------------------------------------------------------------------------
  subroutine zm_batch_gather(ic, state, pblh, tpert, landfrac, cld, fracis)
    ! loop 1 of the split chunk loop: chunk ic's inputs of zm_conv_tend (zm_conv_intr.F90:390-394, 591-600)
    use physics_types, only: physics_state
    integer, intent(in) :: ic
    type(physics_state), intent(in) :: state
    real(r8), intent(in) :: pblh(pcols), tpert(pcols), landfrac(pcols), cld(pcols,pver), fracis(pcols,pver,pcnst)
    ncol_b(ic) = state%ncol
    t_b(:,:,ic) = state%t;       q_b(:,:,ic) = state%q(:,:,1)
    u_b(:,:,ic) = state%u;       v_b(:,:,ic) = state%v
    pmid_b(:,:,ic) = state%pmid; pint_b(:,:,ic) = state%pint; pdel_b(:,:,ic) = state%pdel
    zm_b(:,:,ic) = state%zm;     zi_b(:,:,ic) = state%zi
    phis_b(:,ic) = state%phis;   ps_b(:,ic) = state%ps
    pblh_b(:,ic) = pblh; tpert_b(:,ic) = tpert; landfrac_b(:,ic) = landfrac; cld_b(:,:,ic) = cld
    q3_b(:,:,:,ic) = state%q;    fracis_b(:,:,:,ic) = fracis
  end subroutine zm_batch_gather

  !-------------------------------------------------------------------------------
  subroutine zm_conv_tend_batched(ztodt)
    ! zm_conv_tend for every chunk of the rank (zm_conv_intr.F90:662-886) + its history arithmetic (:685-729)
    real(r8), intent(in) :: ztodt
    integer(c_int) :: rc
    call t_startf('zm_conv_tend_batch')
    ptend_q3_b = 0._r8                        ! physics_ptend_init(ptend_loc, ..., lq=lq) zeroes the flagged slices
    rc = zm_convtran1_fields(int(pcnst, c_int), doconvtran1, cnst_is_dry, q3_b, fracis_b, ptend_q3_b)
    rc = zm_conv_tend_batch(int(nchunks_b, c_int), ncol_b, t_b, q_b, u_b, v_b, pmid_b, pint_b, pdel_b, zm_b, zi_b, &
         phis_b, pblh_b, tpert_b, landfrac_b, cld_b, real(ztodt, c_double), ptend_s_b, ptend_q_b, ptend_u_b, ptend_v_b, &
         mcon_b, cme_b, pflx_b, zdu_b, rliq_b, rice_b, jctop_b, jcbot_b, prec_b, snow_b, ql_b, rprd_b, evapcdp_b, &
         flxprec_b, flxsnow_b, dlf_b, mu_b, md_b, du_b, eu_b, ed_b, dp_b, dsubcld_b, jt_b, maxg_b, ideep_b, lengath_b, cape_b)
    if (rc /= 0) call zm_batch_abort('zm_conv_tend_batch', rc)
    rc = zm_conv_tend_diag_batch(int(nchunks_b, c_int), ncol_b, ps_b, pmid_b, mu_b, md_b, jt_b, maxg_b, ideep_b, &
         lengath_b, freqzm_b, mu_out_b, md_out_b, pcont_b, pconb_b)
    if (rc /= 0) call zm_batch_abort('zm_conv_tend_diag_batch', rc)
    call t_stopf('zm_conv_tend_batch')
  end subroutine zm_conv_tend_batched

  !-------------------------------------------------------------------------------
  subroutine zm_batch_scatter(ic, lchnk, ptend_all, mcon, cme, pflx, zdu, rliq, rice, jctop, jcbot, &
                              prec, snow, ql, rprd, evapcdp, flxprec, flxsnow, dlf, mconzm, &
                              mu, md, du, eu, ed, dp, dsubcld, jt, maxg, ideep, lengath)
    ! loop 2 of the split chunk loop: chunk ic's outputs of zm_conv_tend -- ptend_all (lq(1), ls, lu, lv and the
    ! convtran1 constituents), the dummy outputs (zm_conv_intr.F90:390-394), the pbuf fields (:591-621) -- and the
    ! history output of :676-729, 796-857, 882-883
    use physics_types, only: physics_ptend, physics_ptend_init
    integer, intent(in) :: ic, lchnk
    type(physics_ptend), intent(out) :: ptend_all
    real(r8), intent(out) :: mcon(pcols,pverp), cme(pcols,pver), pflx(pcols,pverp), zdu(pcols,pver), rliq(pcols), &
                             rice(pcols), jctop(pcols), jcbot(pcols)
    real(r8), intent(out) :: prec(pcols), snow(pcols), ql(pcols,pver), rprd(pcols,pver), evapcdp(pcols,pver), &
                             flxprec(pcols,pverp), flxsnow(pcols,pverp), dlf(pcols,pver), mconzm(pcols,pverp)
    real(r8), intent(out) :: mu(pcols,pver), md(pcols,pver), du(pcols,pver), eu(pcols,pver), ed(pcols,pver), &
                             dp(pcols,pver), dsubcld(pcols)
    integer,  intent(out) :: jt(pcols), maxg(pcols), ideep(pcols), lengath
    logical  :: lq(pcnst)
    real(r8) :: ftem(pcols,pver)
    integer  :: m, ncol, ixcldliq, ixcldice
    ncol = ncol_b(ic)
    lq(:) = .false.; lq(1) = .true.
    do m = 2, pcnst
       lq(m) = doconvtran1(m) /= 0
    end do
    call physics_ptend_init(ptend_all, pcols, 'zm_conv_tend', ls=.true., lu=.true., lv=.true., lq=lq)
    ptend_all%s(:ncol,:) = ptend_s_b(:ncol,:,ic)
    ptend_all%u(:ncol,:) = ptend_u_b(:ncol,:,ic)
    ptend_all%v(:ncol,:) = ptend_v_b(:ncol,:,ic)
    ptend_all%q(:ncol,:,1) = ptend_q_b(:ncol,:,ic)
    do m = 2, pcnst
       if (lq(m)) ptend_all%q(:ncol,:,m) = ptend_q3_b(:ncol,:,m,ic)
    end do
    mcon = mcon_b(:,:,ic); cme = cme_b(:,:,ic); pflx = pflx_b(:,:,ic); zdu = zdu_b(:,:,ic)
    rliq = rliq_b(:,ic); rice = rice_b(:,ic); jctop = jctop_b(:,ic); jcbot = jcbot_b(:,ic)
    prec = prec_b(:,ic); snow = snow_b(:,ic); ql = ql_b(:,:,ic); rprd = rprd_b(:,:,ic); evapcdp = evapcdp_b(:,:,ic)
    flxprec = flxprec_b(:,:,ic); flxsnow = flxsnow_b(:,:,ic); dlf = dlf_b(:,:,ic); mconzm = mcon_b(:,:,ic)
    mu = mu_b(:,:,ic); md = md_b(:,:,ic); du = du_b(:,:,ic); eu = eu_b(:,:,ic); ed = ed_b(:,:,ic); dp = dp_b(:,:,ic)
    dsubcld = dsubcld_b(:,ic); jt = jt_b(:,ic); maxg = maxg_b(:,ic); ideep = ideep_b(:,ic); lengath = lengath_b(ic)
    ! history (zm_conv_intr.F90:676-729, 796-801, 882-883)
    call outfld('CAPE',     cape_b(:,ic),   pcols, lchnk)
    call outfld('FREQZM  ', freqzm_b(:,ic), pcols, lchnk)
    call outfld('CMFMC_DP', mconzm,         pcols, lchnk)
    call outfld('ZMMU',     mu_out_b(:,:,ic), pcols, lchnk)
    call outfld('ZMMD',     md_out_b(:,:,ic), pcols, lchnk)
    call outfld('DLFZM',    dlf,            pcols, lchnk)
    call outfld('PCONVT  ', pcont_b(:,ic),  pcols, lchnk)
    call outfld('PCONVB  ', pconb_b(:,ic),  pcols, lchnk)
    ftem(:ncol,:) = evapcdp(:ncol,:)
    call outfld('EVAPQZM ', ftem,           pcols, lchnk)
    call outfld('PRECCDZM   ', prec,        pcols, lchnk)
    call outfld('PRECZ   ', prec,           pcols, lchnk)
    call outfld('ZMMTU',    ptend_u_b(:,:,ic), pcols, lchnk)
    call outfld('ZMMTV',    ptend_v_b(:,:,ic), pcols, lchnk)
    call cnst_get_ind('CLDLIQ', ixcldliq)
    call cnst_get_ind('CLDICE', ixcldice)
    call outfld('ZMDICE ', ptend_q3_b(:,:,ixcldice,ic), pcols, lchnk)
    call outfld('ZMDLIQ ', ptend_q3_b(:,:,ixcldliq,ic), pcols, lchnk)
  end subroutine zm_batch_scatter

  !-------------------------------------------------------------------------------
  subroutine zm_batch_gather_2(ic, state, fracis)
    ! inputs of zm_conv_tend_2 (zm_conv_intr.F90:955-1017): state%q, state%pdeldry, fracis; the mass-flux fields
    ! stay in the library's device mirror (pass them as NULL to zm_conv_tend_batch to skip their copy back)
    use physics_types, only: physics_state
    integer, intent(in) :: ic
    type(physics_state), intent(in) :: state
    real(r8), intent(in) :: fracis(pcols,pver,pcnst)
    q3_b(:,:,:,ic) = state%q
    pdeldry_b(:,:,ic) = state%pdeldry
    fracis_b(:,:,:,ic) = fracis
  end subroutine zm_batch_gather_2

  subroutine zm_conv_tend_2_batched(ztodt)
    real(r8), intent(in) :: ztodt
    integer(c_int) :: rc
    call t_startf('zm_conv_tend_2_batch')
    ptend_q3_b = 0._r8
    rc = zm_conv_tend_2_batch(int(nchunks_b, c_int), doconvtran2, q3_b, int(pcnst, c_int), pdeldry_b, fracis_b, &
                              ptend_q3_b, real(ztodt, c_double), cnst_is_dry)
    if (rc /= 0) call zm_batch_abort('zm_conv_tend_2_batch', rc)
    call t_stopf('zm_conv_tend_2_batch')
  end subroutine zm_conv_tend_2_batched

  subroutine zm_batch_scatter_2(ic, ptend)
    use physics_types, only: physics_ptend, physics_ptend_init
    integer, intent(in) :: ic
    type(physics_ptend), intent(out) :: ptend
    logical :: lq(pcnst)
    integer :: m, ncol
    ncol = ncol_b(ic)
    do m = 1, pcnst
       lq(m) = doconvtran2(m) /= 0
    end do
    call physics_ptend_init(ptend, pcols, 'convtran2', lq=lq)
    do m = 2, pcnst
       if (lq(m)) ptend%q(:ncol,:,m) = ptend_q3_b(:ncol,:,m,ic)
    end do
  end subroutine zm_batch_scatter_2

  !-------------------------------------------------------------------------------
  subroutine zm_batch_report_timers()
    ! device time of the last profiled step under the reference's GPTL timer names (zm_conv_intr.F90:654-880,
    ! 1019-1025): zm_convr, zm_conv_evap, momtran, convtran1, convtran2 (+ physics_update for the glue kernels)
    integer(c_int) :: n, rc
    type(c_ptr) :: names(8)
    real(c_float) :: ms(8)
    character(kind=c_char), pointer :: cname(:)
    character(len=32) :: name
    integer :: j, i
    n = 8
    rc = zm_get_timers(n, names, ms)
    do j = 1, n
       call c_f_pointer(names(j), cname, [32])
       name = ' '
       do i = 1, 32
          if (cname(i) == c_null_char) exit
          name(i:i) = cname(i)
       end do
       write(iulog,'(a,a,f10.4,a)') ' ZM device timer ', name, ms(j), ' ms'
    end do
  end subroutine zm_batch_report_timers

  !-------------------------------------------------------------------------------
  subroutine zm_batch_abort(what, rc)
    ! the reference's endrun (Brent non-convergence: zm_conv.F90:5401-5410, 5557-5566) with the library's text
    character(len=*), intent(in) :: what
    integer(c_int), intent(in) :: rc
    character(kind=c_char) :: buf(512)
    character(len=512) :: msg
    integer :: i, n
    n = zm_last_error(buf, 512_c_int)
    msg = ' '
    do i = 1, min(n, 511)
       msg(i:i) = buf(i)
    end do
    write(iulog,*) trim(what), ' rc = ', rc, ': ', trim(msg)
    call endrun('**** ZM_CONV '//trim(what)//': '//trim(msg))
  end subroutine zm_batch_abort

end module zm_conv_intr_batched
