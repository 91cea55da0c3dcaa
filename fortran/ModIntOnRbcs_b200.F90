! Drop-in replacement of common/ModIntOnRbcs.F90 (public: AddIntOnRbcs, ModIntOnRbcs.F90:18): the real-space
! pair sum over the cell list, the singular (RBC_SingInt) and near-singular (RBC_NearSingInt) corrections and
! AddLinearInt all run on the GPU behind one call.
module ModIntOnRbcs

  use, intrinsic :: iso_c_binding
  use ModDataTypes
  use ModDataStruct
  use ModData
  use ModB200

  implicit none
  private
  public :: AddIntOnRbcs

contains

  subroutine AddIntOnRbcs(c1, c2, tlist, v)   ! ModIntOnRbcs.F90:25-158
    real(WP) :: c1, c2
    type(t_TargetList), target :: tlist
    real(WP) :: v(:, :)
    integer(c_int) :: ierr, kind
    if (nrbc == 0) return
    call B200_EnsureInit
    kind = TlistKind(tlist)
    if (kind == TL_RAW) then
      ! TargetList_CreateFromRaw targets (ModPostProcess.F90:40-59): mirror them on first use
      ierr = rbc3d_targets_set_raw(b200_ctx, tlist%nPoint, tlist%x, merge(1, 0, tlist%active))
      call B200_Check(ierr, 'rbc3d_targets_set_raw')
    end if
    ierr = rbc3d_add_int_on_rbcs(b200_ctx, c1, c2, kind, v)
    call B200_Check(ierr, 'AddIntOnRbcs')
  end subroutine AddIntOnRbcs

end module ModIntOnRbcs
