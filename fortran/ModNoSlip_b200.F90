! Drop-in replacement of common/ModNoSlip.F90 (public: NoSlipWall, Compute_Wall_Residual_Vel, WallBuildMat --
! ModNoSlip.F90:32-34).  The reference solves the first-kind equation for the wall tractions with a matrix-free PETSc
! GMRES whose MyMatMult scatters the Krylov vector into wall%f and applies operator #4 through the host path (two
! copies of wall%f and of v per iteration, ~20 kernel launches each).  Here the whole solve -- right-hand side, Krylov
! vectors, both operators, the update of wall%f and the residual wall velocity -- is one library call,
! rbc3d_noslip_solve: the operator of an iteration is replayed from a CUDA graph and only the Hessenberg column
! crosses to the host.  PETSc is no longer needed by this module.
module ModNoSlip

  use, intrinsic :: iso_c_binding
  use ModDataTypes
  use ModDataStruct
  use ModConf
  use ModData
  use ModB200

  implicit none
  private
  public :: NoSlipWall, Compute_Wall_Residual_Vel, WallBuildMat

contains

  ! ModNoSlip.F90:44-149.  GMRES as the reference sets it up (:70-87): no preconditioner, zero initial guess,
  ! rtol = eps_Ewd, at most 60 iterations, PETSc's default restart of 30.
  subroutine NoSlipWall
    integer :: iwall, p, npoint, nindep
    type(t_wall), pointer :: wall
    integer(c_int), allocatable :: indx(:)
    real(WP), allocatable :: f(:, :), slip(:, :), history(:)
    integer(c_int) :: niter, ierr
    integer, parameter :: maxit = 60

    if (nwall == 0) return
    ! The cell geometry and densities that operator #3 of the right-hand side reads are the mirrored ones: the
    ! B200_SyncCells / B200_SyncDensity calls that follow SourceList_UpdateCoord / SourceList_UpdateDensity in
    ! Compute_Rbc_Vel (INTEGRATION.md 3) have run in this time step; the walls were mirrored by PrepareSingIntOnWall.
    call B200_EnsureInit

    npoint = 0
    do iwall = 1, nwall
      npoint = npoint + walls(iwall)%nvert
    end do
    allocate (indx(npoint), f(npoint, 3), slip(npoint, 3), history(maxit + 1))

    ! wall%f of all walls back to back, and the numbering of the unknowns (indxVertGlb, ModData.F90:50-63)
    p = 0
    nindep = 0
    do iwall = 1, nwall
      wall => walls(iwall)
      f(p + 1:p + wall%nvert, :) = wall%f
      indx(p + 1:p + wall%nvert) = wall%indxVertGlb
      nindep = max(nindep, maxval(wall%indxVertGlb))
      p = p + wall%nvert
    end do

    ierr = rbc3d_noslip_solve(b200_ctx, indx, nindep, vBkg, merge(1, 0, nrbc > 0), eps_Ewd, maxit, f, niter, history, slip)
    call B200_Check(ierr, 'NoSlipWall')

    ! wall%f = f0 + df (:131-137)
    p = 0
    do iwall = 1, nwall
      wall => walls(iwall)
      wall%f = f(p + 1:p + wall%nvert, :)
      p = p + wall%nvert
    end do

    if (rootWorld) then
      write (*, '(A)') 'Wall iteration:'
      write (*, '(A,I5,A,ES12.2)') 'niter = ', niter, ' residual = ', history(niter + 1)
      write (*, '(A,ES12.2,A,ES12.2)') 'Residual velocity: vmax = ', maxval(abs(slip)), ' vzmax = ', maxval(abs(slip(:, 3)))
    end if
    deallocate (indx, f, slip, history)
  end subroutine NoSlipWall

  ! ModNoSlip.F90:153-195: operator #3 (c1 = c2 = 1/(4 pi), cells + walls -> wall vertices) summed over the ranks,
  ! plus the background velocity
  subroutine Compute_Wall_Residual_Vel(v)
    real(WP) :: v(:, :)
    real(WP) :: c1
    integer(c_int) :: ierr
    integer :: ii

    call B200_SyncWallTraction
    c1 = 1./(4.*PI)
    v = 0.
    ierr = rbc3d_apply(b200_ctx, c1, c1, merge(1, 0, nrbc > 0), 1, TL_WALLS, v)
    call B200_Check(ierr, 'Compute_Wall_Residual_Vel')
    ierr = rbc3d_collect_array(b200_ctx, TL_WALLS, v)
    call B200_Check(ierr, 'rbc3d_collect_array')
    do ii = 1, 3
      v(:, ii) = v(:, ii) + vBkg(ii)
    end do
  end subroutine Compute_Wall_Residual_Vel

  ! ModNoSlip.F90:197-253 is a testing routine that assembles the dense wall matrix column by column through PETSc;
  ! nothing in the drivers calls it.  The wall self-interaction blocks are available from the library
  ! (rbc3d_wall_matrix); the dense operator is not rebuilt here.
  subroutine WallBuildMat
    stop 'WallBuildMat is not provided by the B200 shim (testing routine of ModNoSlip.F90)'
  end subroutine WallBuildMat

end module ModNoSlip
