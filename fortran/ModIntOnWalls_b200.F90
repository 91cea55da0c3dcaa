! Drop-in replacement of common/ModIntOnWalls.F90 (public: AddIntOnWalls, SingIntOnWall, PrepareSingIntOnWall,
! Tri_Int_Regular, Tri_Int_Duffy, MinDistToTri -- ModIntOnWalls.F90:20-25).  The sparse self-interaction matrix
! t_Wall%lhs (PETSc SeqAIJ in the reference) lives in GPU memory as a block-row matrix; PETSc is no longer needed by
! this module.  Wall geometry is static, so the wall arrays are mirrored once (B200_SyncWalls) and only the
! tractions wall%f travel per wall-GMRES iteration (B200_SyncWallTraction).
module ModIntOnWalls

  use, intrinsic :: iso_c_binding
  use ModDataTypes
  use ModDataStruct
  use ModData
  use ModB200

  implicit none
  private
  public :: AddIntOnWalls, SingIntOnWall, PrepareSingIntOnWall, Tri_Int_Regular, Tri_Int_Duffy, MinDistToTri

contains

  ! ModIntOnWalls.F90:33-130.  wall%f is re-sent before every application: the callers change it between calls
  ! (ModNoSlip.F90:275-283 scatters the Krylov vector into wall%f).
  subroutine AddIntOnWalls(c1, tlist, v)
    real(WP) :: c1
    type(t_TargetList), target :: tlist
    real(WP) :: v(:, :)
    integer(c_int) :: ierr, kind
    if (nwall == 0) return
    call B200_SyncWallTraction
    kind = TlistKind(tlist)
    if (kind == TL_RAW) then
      ierr = rbc3d_targets_set_raw(b200_ctx, tlist%nPoint, tlist%x, merge(1, 0, tlist%active))
      call B200_Check(ierr, 'rbc3d_targets_set_raw')
    end if
    ierr = rbc3d_add_int_on_walls(b200_ctx, c1, kind, v)
    call B200_Check(ierr, 'AddIntOnWalls')
  end subroutine AddIntOnWalls

  ! ModIntOnWalls.F90:181-308.  The reference is called once per wall (ModTimeInt.F90:101-103); the matrices of all
  ! walls are assembled by the first call, later calls for the same geometry are no-ops.
  subroutine PrepareSingIntOnWall(wall)
    type(t_wall) :: wall
    integer(c_int) :: ierr
    if (wall%ID /= walls(1)%ID) return
    call B200_SyncWalls
    ierr = rbc3d_wall_prepare_sing(b200_ctx)
    call B200_Check(ierr, 'PrepareSingIntOnWall')
  end subroutine PrepareSingIntOnWall

  ! ModIntOnWalls.F90:136-172: v = c1 * lhs * wall%f
  subroutine SingIntOnWall(c1, wall, v)
    real(WP) :: c1
    type(t_Wall) :: wall
    real(WP) :: v(:, :)
    integer(c_int) :: ierr
    call B200_SyncWallTraction
    ierr = rbc3d_sing_int_on_wall(b200_ctx, c1, wall%ID - walls(1)%ID, v)
    call B200_Check(ierr, 'SingIntOnWall')
  end subroutine SingIntOnWall

  ! ModIntOnWalls.F90:319-363 (one target, one triangle; the batched entry point is called with n = 1)
  subroutine Tri_Int_Regular(x, f, xtar, rhs, lhs)
    real(WP) :: x(3, 3), f(3, 3), xtar(3), rhs(3)
    real(WP), optional :: lhs(3, 3, 3)
    real(WP) :: xt(9), ft(9), lt(27)
    integer(c_int) :: ierr
    call B200_EnsureInit
    xt = reshape(transpose(x), (/9/)); ft = reshape(transpose(f), (/9/))     ! [corner][component], C order
    if (present(lhs)) then
      ierr = rbc3d_tri_int(b200_ctx, 1, xt, ft, xtar, c_null_ptr, c_null_ptr, rhs, lt)
      lhs = reshape(lt, (/3, 3, 3/), order=(/3, 2, 1/))                       ! C [l][ii][jj] -> lhs(l,ii,jj)
    else
      ierr = rbc3d_tri_int_rhs(b200_ctx, 1, xt, ft, xtar, c_null_ptr, c_null_ptr, rhs, c_null_ptr)
    end if
    call B200_Check(ierr, 'Tri_Int_Regular')
  end subroutine Tri_Int_Regular

  ! ModIntOnWalls.F90:373-465
  subroutine Tri_Int_Duffy(x, f, xtar, s0, t0, rhs, lhs)
    real(WP) :: x(3, 3), f(3, 3), xtar(3), s0, t0, rhs(3)
    real(WP), optional :: lhs(3, 3, 3)
    real(WP), target :: st(2)
    real(WP) :: xt(9), ft(9), lt(27)
    integer(c_int) :: ierr
    call B200_EnsureInit
    xt = reshape(transpose(x), (/9/)); ft = reshape(transpose(f), (/9/))
    st = (/s0, t0/)
    if (present(lhs)) then
      ierr = rbc3d_tri_int(b200_ctx, 1, xt, ft, xtar, c_loc(st(1)), c_loc(st(2)), rhs, lt)
      lhs = reshape(lt, (/3, 3, 3/), order=(/3, 2, 1/))
    else
      ierr = rbc3d_tri_int_rhs(b200_ctx, 1, xt, ft, xtar, c_loc(st(1)), c_loc(st(2)), rhs, c_null_ptr)
    end if
    call B200_Check(ierr, 'Tri_Int_Duffy')
  end subroutine Tri_Int_Duffy

  ! ModIntOnWalls.F90:480-577 (also used by ModRepulsion.F90:12)
  function MinDistToTri(xTar, x, s0, t0, x0)
    real(WP) :: MinDistToTri
    real(WP) :: xTar(3), x(3, 3)
    real(WP), optional :: s0, t0, x0(3)
    real(WP) :: xt(9), d(1), s(1), t(1)
    integer(c_int) :: ierr
    call B200_EnsureInit
    xt = reshape(transpose(x), (/9/))
    ierr = rbc3d_min_dist_to_tri(b200_ctx, 1, xTar, xt, d, s, t)
    call B200_Check(ierr, 'MinDistToTri')
    MinDistToTri = d(1)
    if (present(s0)) s0 = s(1)
    if (present(t0)) t0 = t(1)
    if (present(x0)) x0 = (1 - s(1) - t(1))*x(1, :) + s(1)*x(2, :) + t(1)*x(3, :)
  end function MinDistToTri

end module ModIntOnWalls
