! fwrap.f90 -- the six assumed-size entry points Fortran programs call with arrays of any rank and type
! (counterpart of the reference's build/fwrap.f90:81-154): they forward to the bind(C) interfaces of module
! p3dfft_plus_plus.  Arrays are passed by address; complex data is interleaved (re, im) of the base precision.

subroutine p3dfft_1Dtrans_double(plan, a_in, a_out, ow)
  use p3dfft_plus_plus
  implicit none
  integer :: plan, ow
  double precision, target :: a_in(1, 1, *), a_out(1, 1, *)
  call p3dfft_exec_1Dtrans_double(plan, a_in, a_out, ow)
end subroutine

subroutine p3dfft_1Dtrans_single(plan, a_in, a_out, ow)
  use p3dfft_plus_plus
  implicit none
  integer :: plan, ow
  real, target :: a_in(1, 1, *), a_out(1, 1, *)
  call p3dfft_exec_1Dtrans_single(plan, a_in, a_out, ow)
end subroutine

subroutine p3dfft_3Dtrans_double(plan, a_in, a_out, ow)
  use p3dfft_plus_plus
  implicit none
  integer :: plan, ow
  double precision, target :: a_in(1, 1, *), a_out(1, 1, *)
  call p3dfft_exec_3Dtrans_double(plan, a_in, a_out, ow)
end subroutine

subroutine p3dfft_3Dtrans_single(plan, a_in, a_out, ow)
  use p3dfft_plus_plus
  implicit none
  integer :: plan, ow
  real, target :: a_in(1, 1, *), a_out(1, 1, *)
  call p3dfft_exec_3Dtrans_single(plan, a_in, a_out, ow)
end subroutine

subroutine p3dfft_3Dderiv_double(plan, a_in, a_out, idir, ow)
  use p3dfft_plus_plus
  implicit none
  integer :: plan, idir, ow
  double precision, target :: a_in(1, 1, *), a_out(1, 1, *)
  call p3dfft_exec_3Dderiv_double(plan, a_in, a_out, idir, ow)
end subroutine

subroutine p3dfft_3Dderiv_single(plan, a_in, a_out, idir, ow)
  use p3dfft_plus_plus
  implicit none
  integer :: plan, idir, ow
  real, target :: a_in(1, 1, *), a_out(1, 1, *)
  call p3dfft_exec_3Dderiv_single(plan, a_in, a_out, idir, ow)
end subroutine
