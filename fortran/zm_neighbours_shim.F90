module zm_neighbours_b200
!---------------------------------------------------------------------------------
! Bindings for the two neighbours of the ZM path that SURVEY.md section 8(f) N4 names and that
! libzmconv_b200.so provides (include/zmconv_b200.h):
!
!   geopotential_t            physics/geopotential.F90:153-310   -> geopotential_t_b200 below,
!                             same dummy-argument list; the classic branch (:208-247) forwards to
!                             zm_geopotential_t_batch, the generalized-virtual-temperature branch the
!                             reference takes for dycore MPAS / SE (:248-310) to zm_geopotential_t_gen_batch
!   convect_diagnostics_calc  physics/convect_diagnostics.F90:115-249 -> the array arithmetic of its
!                             CLUBB_SGS configuration, zm_convect_diagnostics_batch (the pbuf / state
!                             unpacking and the outfld calls stay in the caller)
!
! Both forward with nchunks = 1 like fortran/zm_conv_shim.F90; the batched call takes all chunks of a
! rank at once with the arrays of the chunks back to back (INTEGRATION.md section 2).
!
! NOT COMPILED IN THIS REPOSITORY'S CI (no Fortran compiler in the image); the interface blocks are
! checked against the header by tests/test_fortran_binding.py with the C compiler.
!---------------------------------------------------------------------------------
  use, intrinsic :: iso_c_binding
  use shr_kind_mod,    only: r8 => shr_kind_r8
  use ppgrid,          only: pcols, pver, pverp
  use dycore,          only: dycore_is
  use air_composition, only: thermodynamic_active_species_num, thermodynamic_active_species_idx
  use cam_abortutils,  only: endrun

  implicit none
  private
  save

  public geopotential_t_b200, convect_diagnostics_arrays_b200

  interface
     integer(c_int) function zm_last_error(buf, buflen) bind(C, name='zm_last_error')
       import :: c_int, c_char
       character(kind=c_char), intent(out) :: buf(*)
       integer(c_int), value :: buflen
     end function zm_last_error

     integer(c_int) function zm_geopotential_t_batch(nchunks, ncol, dycore_lr, piln, pmln, pint, pmid, pdel, &
          rpdel, t, q, rair, gravit, zvir, zi, zm) bind(C, name='zm_geopotential_t_batch')
       import :: c_int, c_double
       integer(c_int), value :: nchunks, dycore_lr
       integer(c_int), intent(in) :: ncol(*)
       real(c_double), intent(in) :: piln(*), pmln(*), pint(*), pmid(*), pdel(*), rpdel(*), t(*), q(*), rair(*), zvir(*)
       real(c_double), value :: gravit
       real(c_double), intent(out) :: zi(*), zm(*)
     end function zm_geopotential_t_batch

     integer(c_int) function zm_geopotential_t_gen_batch(nchunks, ncol, dycore_lr, ncnst, nspecies, species_idx, &
          piln, pmln, pint, pmid, pdel, rpdel, t, q3, rair, gravit, zvir, zi, zm) &
          bind(C, name='zm_geopotential_t_gen_batch')
       import :: c_int, c_double
       integer(c_int), value :: nchunks, dycore_lr, ncnst, nspecies
       integer(c_int), intent(in) :: ncol(*), species_idx(*)
       real(c_double), intent(in) :: piln(*), pmln(*), pint(*), pmid(*), pdel(*), rpdel(*), t(*), q3(*), rair(*), zvir(*)
       real(c_double), value :: gravit
       real(c_double), intent(out) :: zi(*), zm(*)
     end function zm_geopotential_t_gen_batch

     integer(c_int) function zm_convect_diagnostics_batch(nchunks, ncol, cmfmc, qc, qc2, rliq, rliq2, pmid, &
          rprddp, cnt, cnb, cmfmc2, rprdsh, rprdtot, pcnt, pcnb) bind(C, name='zm_convect_diagnostics_batch')
       import :: c_int, c_double
       integer(c_int), value :: nchunks
       integer(c_int), intent(in) :: ncol(*)
       real(c_double), intent(inout) :: cmfmc(*), qc(*), rliq(*), cnt(*), cnb(*)
       real(c_double), intent(out) :: qc2(*), rliq2(*), cmfmc2(*), rprdsh(*), rprdtot(*), pcnt(*), pcnb(*)
       real(c_double), intent(in) :: pmid(*), rprddp(*)
     end function zm_convect_diagnostics_batch
  end interface

contains

  subroutine nb_abort(where, rc)
    character(len=*), intent(in) :: where
    integer(c_int),   intent(in) :: rc
    character(kind=c_char) :: cbuf(1024)
    character(len=1024)    :: msg
    integer :: n, i
    n = zm_last_error(cbuf, 1024_c_int)
    msg = ' '
    do i = 1, min(n, 1024)
       msg(i:i) = cbuf(i)
    end do
    call endrun('libzmconv_b200 '//trim(where)//': '//trim(msg))
  end subroutine nb_abort

  ! geopotential_t with the reference's dummy-argument list (geopotential.F90:153-156)
  subroutine geopotential_t_b200(                            &
       piln   , pmln   , pint   , pmid   , pdel   , rpdel  , &
       t      , q      , rair   , gravit , zvir   ,          &
       zi     , zm     , ncol   )
    integer,  intent(in)  :: ncol
    real(r8), intent(in)  :: piln (:,:), pmln (:,:), pint (:,:), pmid (:,:), pdel (:,:), rpdel(:,:)
    real(r8), intent(in)  :: t    (:,:)
    real(r8), intent(in)  :: q    (:,:,:)          ! (pcols,pver,:) tracers (moist mixing ratios)
    real(r8), intent(in)  :: rair (:,:), zvir (:,:)
    real(r8), intent(in)  :: gravit
    real(r8), intent(out) :: zi(:,:), zm(:,:)

    integer(c_int) :: rc, nc(1), lr, sp(max(1, thermodynamic_active_species_num))
    integer :: idx

    nc(1) = ncol
    lr = merge(1_c_int, 0_c_int, dycore_is('LR') .or. dycore_is('FV3'))
    if (.not. (dycore_is('MPAS') .or. dycore_is('SE'))) then
       ! geopotential.F90:208-247; q(:,:,1) is the first (pcols,pver) slice of the contiguous array
       rc = zm_geopotential_t_batch(1_c_int, nc, lr, piln, pmln, pint, pmid, pdel, rpdel, t, q, rair, &
                                    real(gravit, c_double), zvir, zi, zm)
    else
       ! geopotential.F90:248-310
       do idx = 1, thermodynamic_active_species_num
          sp(idx) = int(thermodynamic_active_species_idx(idx), c_int)
       end do
       rc = zm_geopotential_t_gen_batch(1_c_int, nc, lr, int(size(q, 3), c_int), &
                                        int(thermodynamic_active_species_num, c_int), sp, piln, pmln, pint, pmid, &
                                        pdel, rpdel, t, q, rair, real(gravit, c_double), zvir, zi, zm)
    end if
    if (rc /= 0) call nb_abort('geopotential_t', rc)
  end subroutine geopotential_t_b200

  ! The array arithmetic of convect_diagnostics_calc for shallow_scheme = 'CLUBB_SGS'
  ! (convect_diagnostics.F90:187-249); the caller keeps the pbuf_get_field / outfld calls around it.
  subroutine convect_diagnostics_arrays_b200(ncol, cmfmc, qc, qc2, rliq, rliq2, pmid, rprddp, cnt, cnb, &
                                             cmfmc2, rprdsh, rprdtot, pcnt, pcnb)
    integer,  intent(in)    :: ncol
    real(r8), intent(inout) :: cmfmc(pcols,pverp), qc(pcols,pver), rliq(pcols), cnt(pcols), cnb(pcols)
    real(r8), intent(out)   :: qc2(pcols,pver), rliq2(pcols), cmfmc2(pcols,pverp), rprdsh(pcols,pver)
    real(r8), intent(out)   :: rprdtot(pcols,pver), pcnt(pcols), pcnb(pcols)
    real(r8), intent(in)    :: pmid(pcols,pver), rprddp(pcols,pver)
    integer(c_int) :: rc, nc(1)
    nc(1) = ncol
    rc = zm_convect_diagnostics_batch(1_c_int, nc, cmfmc, qc, qc2, rliq, rliq2, pmid, rprddp, cnt, cnb, &
                                      cmfmc2, rprdsh, rprdtot, pcnt, pcnb)
    if (rc /= 0) call nb_abort('convect_diagnostics_calc', rc)
  end subroutine convect_diagnostics_arrays_b200

end module zm_neighbours_b200
