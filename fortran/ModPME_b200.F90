! Drop-in replacement of common/ModPME.F90: same module name, same public procedures (ModPME.F90:44-48), every
! call forwarded to librbc3d_b200.so.  The module-level meshes ff/tt/vv and their Fourier images live in HBM
! inside the library context; ModPFFTW is no longer used (cuFFT).
module ModPME

  use, intrinsic :: iso_c_binding
  use ModDataTypes
  use ModDataStruct
  use ModConf
  use ModData
  use ModB200

  implicit none
  private
  public :: PME_Init, PME_Finalize, PME_Distrib_Source, PME_Transform, PME_Add_Interp_Vel

contains

  subroutine PME_Init                       ! ModPME.F90:252-338 (no-op when an earlier call created the context)
    call B200_Init
  end subroutine PME_Init

  subroutine PME_Finalize                   ! ModPME.F90:342-350
    integer(c_int) :: ierr
    if (.not. c_associated(b200_ctx)) return
    ierr = rbc3d_ctx_destroy(b200_ctx)
    b200_ctx = c_null_ptr
  end subroutine PME_Finalize

  ! ModPME.F90:58-133.  `cells` / `walls` are only tested for presence: the library spreads the source lists that
  ! SourceList_UpdateCoord / SourceList_UpdateDensity mirrored (B200_SyncCells / B200_SyncDensity).
  subroutine PME_Distrib_Source(c1, c2, cells, walls)
    real(WP) :: c1, c2
    type(t_rbc), target, optional :: cells(:)
    type(t_wall), target, optional :: walls(:)
    integer(c_int) :: ierr
    call B200_EnsureInit
    ! wall%f may have changed since the last real-space call, and Fourier-only ranks (PhysEwald = .false.) and
    ! ModPostProcess never go through AddIntOnWalls: send the tractions the spreading uses
    if (present(walls)) then
      if (nwall > 0) call B200_SyncWallTraction
    end if
    ierr = rbc3d_pme_distrib_source(b200_ctx, c1, c2, merge(1, 0, present(cells)), merge(1, 0, present(walls)))
    call B200_Check(ierr, 'PME_Distrib_Source')
  end subroutine PME_Distrib_Source

  subroutine PME_Transform                  ! ModPME.F90:137-222
    integer(c_int) :: ierr
    call B200_EnsureInit
    ierr = rbc3d_pme_transform(b200_ctx)
    call B200_Check(ierr, 'PME_Transform')
  end subroutine PME_Transform

  subroutine PME_Add_Interp_Vel(tlist, v)   ! ModPME.F90:228-248
    type(t_TargetList), target :: tlist
    real(WP) :: v(:, :)
    integer(c_int) :: ierr
    call B200_EnsureInit
    ierr = rbc3d_pme_add_interp_vel(b200_ctx, TlistKind(tlist), v)
    call B200_Check(ierr, 'PME_Add_Interp_Vel')
  end subroutine PME_Add_Interp_Vel

end module ModPME
