! Drop-in replacement of common/ModEwaldFunc.F90 (public procedures, ModEwaldFunc.F90:13-16).  The table versions
! use the context's alpha and rc; the tables themselves (8193 entries, linear interpolation) are built by the
! library exactly as ModEwaldFunc.F90:96-106,150-160 builds them.
module ModEwaldFunc

  use, intrinsic :: iso_c_binding
  use ModDataTypes
  use ModB200

  implicit none
  private
  public :: EwaldCoeff_SL_Exact, EwaldCoeff_DL_Exact, EwaldCoeff_SL, EwaldCoeff_DL

contains

  subroutine EwaldCoeff_SL_Exact(r, alpha, A, B)   ! ModEwaldFunc.F90:25-52
    real(WP) :: r, alpha, A, B
    integer(c_int) :: ierr
    ierr = rbc3d_ewald_coeff_sl_exact(r, alpha, A, B)
  end subroutine EwaldCoeff_SL_Exact

  subroutine EwaldCoeff_DL_Exact(r, alpha, A)      ! ModEwaldFunc.F90:59-79
    real(WP) :: r, alpha, A
    integer(c_int) :: ierr
    ierr = rbc3d_ewald_coeff_dl_exact(r, alpha, A)
  end subroutine EwaldCoeff_DL_Exact

  subroutine EwaldCoeff_SL(r, A, B)                ! ModEwaldFunc.F90:86-131
    real(WP) :: r, A, B
    integer(c_int) :: ierr
    call B200_EnsureInit
    ierr = rbc3d_ewald_coeff_sl(b200_ctx, r, A, B)
  end subroutine EwaldCoeff_SL

  subroutine EwaldCoeff_DL(r, A)                   ! ModEwaldFunc.F90:141-178
    real(WP) :: r, A
    integer(c_int) :: ierr
    call B200_EnsureInit
    ierr = rbc3d_ewald_coeff_dl(b200_ctx, r, A)
  end subroutine EwaldCoeff_DL

end module ModEwaldFunc
