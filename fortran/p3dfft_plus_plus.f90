! p3dfft_plus_plus.f90 -- Fortran 2003 interface of the B200-native P3DFFT++ transform path.
! Binds the `_f` entry points of include/Fwrap.h (all arguments by reference, integer handles for grids and plans,
! 1-based idir) and the 44 transform-type IDs.  Counterpart of the reference's build/fp3dfft++mod.f90:1-159; unlike
! the reference module it declares every registered ID (the reference omits DCT2-4/DST2-4 and binds a
! non-existent P3DFFT_EMPTY_TYPE).  Generated from the C prototypes; no Fortran compiler exists in the build
! image, so this file ships untested (see INTEGRATION.md section 3).
module p3dfft_plus_plus
  use iso_c_binding
  implicit none

  integer(C_INT), bind(C, name='P3DFFT_EMPTY_TYPE_SINGLE') :: P3DFFT_EMPTY_TYPE_SINGLE
  integer(C_INT), bind(C, name='P3DFFT_EMPTY_TYPE_DOUBLE') :: P3DFFT_EMPTY_TYPE_DOUBLE
  integer(C_INT), bind(C, name='P3DFFT_EMPTY_TYPE_SINGLE_COMPLEX') :: P3DFFT_EMPTY_TYPE_SINGLE_COMPLEX
  integer(C_INT), bind(C, name='P3DFFT_EMPTY_TYPE_DOUBLE_COMPLEX') :: P3DFFT_EMPTY_TYPE_DOUBLE_COMPLEX
  integer(C_INT), bind(C, name='P3DFFT_R2CFFT_S') :: P3DFFT_R2CFFT_S
  integer(C_INT), bind(C, name='P3DFFT_R2CFFT_D') :: P3DFFT_R2CFFT_D
  integer(C_INT), bind(C, name='P3DFFT_C2RFFT_S') :: P3DFFT_C2RFFT_S
  integer(C_INT), bind(C, name='P3DFFT_C2RFFT_D') :: P3DFFT_C2RFFT_D
  integer(C_INT), bind(C, name='P3DFFT_CFFT_FORWARD_S') :: P3DFFT_CFFT_FORWARD_S
  integer(C_INT), bind(C, name='P3DFFT_CFFT_FORWARD_D') :: P3DFFT_CFFT_FORWARD_D
  integer(C_INT), bind(C, name='P3DFFT_CFFT_BACKWARD_S') :: P3DFFT_CFFT_BACKWARD_S
  integer(C_INT), bind(C, name='P3DFFT_CFFT_BACKWARD_D') :: P3DFFT_CFFT_BACKWARD_D
  integer(C_INT), bind(C, name='P3DFFT_DCT1_REAL_S') :: P3DFFT_DCT1_REAL_S
  integer(C_INT), bind(C, name='P3DFFT_DCT1_REAL_D') :: P3DFFT_DCT1_REAL_D
  integer(C_INT), bind(C, name='P3DFFT_DCT1_COMPLEX_S') :: P3DFFT_DCT1_COMPLEX_S
  integer(C_INT), bind(C, name='P3DFFT_DCT1_COMPLEX_D') :: P3DFFT_DCT1_COMPLEX_D
  integer(C_INT), bind(C, name='P3DFFT_DST1_REAL_S') :: P3DFFT_DST1_REAL_S
  integer(C_INT), bind(C, name='P3DFFT_DST1_REAL_D') :: P3DFFT_DST1_REAL_D
  integer(C_INT), bind(C, name='P3DFFT_DST1_COMPLEX_S') :: P3DFFT_DST1_COMPLEX_S
  integer(C_INT), bind(C, name='P3DFFT_DST1_COMPLEX_D') :: P3DFFT_DST1_COMPLEX_D
  integer(C_INT), bind(C, name='P3DFFT_DCT2_REAL_S') :: P3DFFT_DCT2_REAL_S
  integer(C_INT), bind(C, name='P3DFFT_DCT2_REAL_D') :: P3DFFT_DCT2_REAL_D
  integer(C_INT), bind(C, name='P3DFFT_DCT2_COMPLEX_S') :: P3DFFT_DCT2_COMPLEX_S
  integer(C_INT), bind(C, name='P3DFFT_DCT2_COMPLEX_D') :: P3DFFT_DCT2_COMPLEX_D
  integer(C_INT), bind(C, name='P3DFFT_DST2_REAL_S') :: P3DFFT_DST2_REAL_S
  integer(C_INT), bind(C, name='P3DFFT_DST2_REAL_D') :: P3DFFT_DST2_REAL_D
  integer(C_INT), bind(C, name='P3DFFT_DST2_COMPLEX_S') :: P3DFFT_DST2_COMPLEX_S
  integer(C_INT), bind(C, name='P3DFFT_DST2_COMPLEX_D') :: P3DFFT_DST2_COMPLEX_D
  integer(C_INT), bind(C, name='P3DFFT_DCT3_REAL_S') :: P3DFFT_DCT3_REAL_S
  integer(C_INT), bind(C, name='P3DFFT_DCT3_REAL_D') :: P3DFFT_DCT3_REAL_D
  integer(C_INT), bind(C, name='P3DFFT_DCT3_COMPLEX_S') :: P3DFFT_DCT3_COMPLEX_S
  integer(C_INT), bind(C, name='P3DFFT_DCT3_COMPLEX_D') :: P3DFFT_DCT3_COMPLEX_D
  integer(C_INT), bind(C, name='P3DFFT_DST3_REAL_S') :: P3DFFT_DST3_REAL_S
  integer(C_INT), bind(C, name='P3DFFT_DST3_REAL_D') :: P3DFFT_DST3_REAL_D
  integer(C_INT), bind(C, name='P3DFFT_DST3_COMPLEX_S') :: P3DFFT_DST3_COMPLEX_S
  integer(C_INT), bind(C, name='P3DFFT_DST3_COMPLEX_D') :: P3DFFT_DST3_COMPLEX_D
  integer(C_INT), bind(C, name='P3DFFT_DCT4_REAL_S') :: P3DFFT_DCT4_REAL_S
  integer(C_INT), bind(C, name='P3DFFT_DCT4_REAL_D') :: P3DFFT_DCT4_REAL_D
  integer(C_INT), bind(C, name='P3DFFT_DCT4_COMPLEX_S') :: P3DFFT_DCT4_COMPLEX_S
  integer(C_INT), bind(C, name='P3DFFT_DCT4_COMPLEX_D') :: P3DFFT_DCT4_COMPLEX_D
  integer(C_INT), bind(C, name='P3DFFT_DST4_REAL_S') :: P3DFFT_DST4_REAL_S
  integer(C_INT), bind(C, name='P3DFFT_DST4_REAL_D') :: P3DFFT_DST4_REAL_D
  integer(C_INT), bind(C, name='P3DFFT_DST4_COMPLEX_S') :: P3DFFT_DST4_COMPLEX_S
  integer(C_INT), bind(C, name='P3DFFT_DST4_COMPLEX_D') :: P3DFFT_DST4_COMPLEX_D

  interface
    subroutine p3dfft_setup() bind(C, name='p3dfft_setup')
    end subroutine
    subroutine p3dfft_cleanup() bind(C, name='p3dfft_cleanup')
    end subroutine
    subroutine p3dfft_init_3Dtype(mytype, types) bind(C, name='p3dfft_init_3Dtype_f')
      import
      integer(C_INT) :: mytype, types(3)
    end subroutine
    integer(C_INT) function p3dfft_init_proc_grid(pdims, mpicomm) bind(C, name='p3dfft_init_proc_grid_f')
      import
      integer(C_INT) :: pdims(3), mpicomm
    end function
    subroutine p3dfft_init_data_grid(mygrid, ldims, glob_start, gdims, dim_conj_sym, pgrid, dmap, mem_order) &
        bind(C, name='p3dfft_init_data_grid_f')
      import
      integer(C_INT) :: mygrid, ldims(3), glob_start(3), gdims(3), dim_conj_sym, pgrid, dmap(3), mem_order(3)
    end subroutine
    ! (the reference module spells this interface with three f's, build/fp3dfft++mod.f90:87; kept so that callers compile unchanged)
    subroutine p3dffft_inv_mo(mo, imo) bind(C, name='p3dfft_inv_mo')
      import
      integer(C_INT) :: mo(3), imo(3)
    end subroutine
    subroutine p3dfft_plan_1Dtrans(plan, grid1, grid2, type_id, d) bind(C, name='p3dfft_plan_1Dtrans_f')
      import
      integer(C_INT) :: plan, grid1, grid2, type_id, d
    end subroutine
    subroutine p3dfft_plan_3Dtrans(plan, grid1, grid2, type3d) bind(C, name='p3dfft_plan_3Dtrans_f')
      import
      integer(C_INT) :: plan, grid1, grid2, type3d
    end subroutine
    subroutine p3dfft_exec_1Dtrans_double(plan, a_in, a_out, ow) bind(C, name='p3dfft_exec_1Dtrans_double_f')
      import
      integer(C_INT) :: plan, ow
      real(C_DOUBLE) :: a_in(*), a_out(*)
    end subroutine
    subroutine p3dfft_exec_1Dtrans_single(plan, a_in, a_out, ow) bind(C, name='p3dfft_exec_1Dtrans_single_f')
      import
      integer(C_INT) :: plan, ow
      real(C_FLOAT) :: a_in(*), a_out(*)
    end subroutine
    subroutine p3dfft_exec_3Dtrans_double(plan, a_in, a_out, ow) bind(C, name='p3dfft_exec_3Dtrans_double_f')
      import
      integer(C_INT) :: plan, ow
      real(C_DOUBLE) :: a_in(*), a_out(*)
    end subroutine
    subroutine p3dfft_exec_3Dtrans_single(plan, a_in, a_out, ow) bind(C, name='p3dfft_exec_3Dtrans_single_f')
      import
      integer(C_INT) :: plan, ow
      real(C_FLOAT) :: a_in(*), a_out(*)
    end subroutine
    subroutine p3dfft_exec_3Dderiv_double(plan, a_in, a_out, idir, ow) bind(C, name='p3dfft_exec_3Dderiv_double_f')
      import
      integer(C_INT) :: plan, idir, ow
      real(C_DOUBLE) :: a_in(*), a_out(*)
    end subroutine
    subroutine p3dfft_exec_3Dderiv_single(plan, a_in, a_out, idir, ow) bind(C, name='p3dfft_exec_3Dderiv_single_f')
      import
      integer(C_INT) :: plan, idir, ow
      real(C_FLOAT) :: a_in(*), a_out(*)
    end subroutine
    subroutine p3dfft_compute_deriv_double(a_in, a_out, grid, idir) bind(C, name='p3dfft_compute_deriv_double_f')
      import
      integer(C_INT) :: grid, idir
      real(C_DOUBLE) :: a_in(*), a_out(*)
    end subroutine
    subroutine p3dfft_compute_deriv_single(a_in, a_out, grid, idir) bind(C, name='p3dfft_compute_deriv_single_f')
      import
      integer(C_INT) :: grid, idir
      real(C_FLOAT) :: a_in(*), a_out(*)
    end subroutine
  end interface
end module p3dfft_plus_plus
