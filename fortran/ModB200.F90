! ModB200 -- ISO_C_BINDING interface to librbc3d_b200.so (include/rbc3d.h) and the glue that mirrors the
! reference's module state into the library's device-resident context.
!
! This file and the Mod*_b200.F90 files next to it are the drop-in side of the boundary: they keep the module and
! procedure names of common/ModPME.F90, ModEwaldFunc.F90, ModIntOnRbcs.F90 (and, once the wall kernels land,
! ModIntOnWalls.F90) so that ModVelSolver, ModNoSlip, ModPostProcess and ModTimeInt compile unchanged.  They are
! SOURCE ONLY in this repository: the build image has no Fortran compiler (DESIGN.md "Toolchain"), so they have not
! been compiled here; INTEGRATION.md lists what a maintainer has to do.
!
! Conventions (include/rbc3d.h): real(WP) = c_double, default integer = c_int, arrays x(n,3) are passed whole
! (column major = three planes of length n = the library's SoA(3,n)), logical active(:) is converted to c_int.
module ModB200

  use, intrinsic :: iso_c_binding
  use ModDataTypes
  use ModDataStruct
  use ModConf
  use ModData

  implicit none

  type(c_ptr), save :: b200_ctx = c_null_ptr
  integer(c_int), parameter :: TL_CELLS = 0, TL_RAW = 1, TL_WALLS = 2

  interface
    function rbc3d_ctx_create(ctx, Lb, alpha, eps, P, rc, Nb, device) bind(C, name="rbc3d_ctx_create") result(ierr)
      import
      type(c_ptr) :: ctx
      real(c_double) :: Lb(3)
      real(c_double), value :: alpha, eps, rc
      integer(c_int), value :: P, device
      integer(c_int) :: Nb(3)
      integer(c_int) :: ierr
    end function
    function rbc3d_ctx_destroy(ctx) bind(C, name="rbc3d_ctx_destroy") result(ierr)
      import
      type(c_ptr), value :: ctx
      integer(c_int) :: ierr
    end function
    function rbc3d_comm_unique_id(id128) bind(C, name="rbc3d_comm_unique_id") result(ierr)
      import
      character(kind=c_char) :: id128(128)
      integer(c_int) :: ierr
    end function
    function rbc3d_ctx_attach_comm(ctx, nranks, rank, id128) bind(C, name="rbc3d_ctx_attach_comm") result(ierr)
      import
      type(c_ptr), value :: ctx
      integer(c_int), value :: nranks, rank
      character(kind=c_char) :: id128(128)
      integer(c_int) :: ierr
    end function
    function rbc3d_cells_set_mesh(ctx, ncell, nlat, nlon, th, phi, w) bind(C, name="rbc3d_cells_set_mesh") result(ierr)
      import
      type(c_ptr), value :: ctx
      integer(c_int), value :: ncell, nlat, nlon
      real(c_double) :: th(*), phi(*), w(*)
      integer(c_int) :: ierr
    end function
    function rbc3d_cells_set_geometry(ctx, x, a3, Acoef, Bcoef, area, meshSize, spx, spa3, spdetj, active) &
      bind(C, name="rbc3d_cells_set_geometry") result(ierr)
      import
      type(c_ptr), value :: ctx
      real(c_double) :: x(*), a3(*), Acoef(*), Bcoef(*), area(*), meshSize(*), spx(*), spa3(*), spdetj(*)
      integer(c_int) :: active(*)
      integer(c_int) :: ierr
    end function
    function rbc3d_cells_set_geometry_mesh(ctx, x, a3, detj, Acoef, Bcoef, area, meshSize, active) &
      bind(C, name="rbc3d_cells_set_geometry_mesh") result(ierr)
      import
      type(c_ptr), value :: ctx
      real(c_double) :: x(*), a3(*), detj(*), Acoef(*), Bcoef(*), area(*), meshSize(*)
      integer(c_int) :: active(*)
      integer(c_int) :: ierr
    end function
    function rbc3d_solver_setup(ctx, nlat0, detj) bind(C, name="rbc3d_solver_setup") result(ierr)
      import
      type(c_ptr), value :: ctx
      integer(c_int), value :: nlat0
      real(c_double) :: detj(*)
      integer(c_int) :: ierr
    end function
    function rbc3d_solver_matmult(ctx, u, b) bind(C, name="rbc3d_solver_matmult") result(ierr)
      import
      type(c_ptr), value :: ctx
      real(c_double) :: u(*), b(*)
      integer(c_int) :: ierr
    end function
    function rbc3d_solver_gmres(ctx, rhs, sol, rtol, restart, maxit, niter, history) &
      bind(C, name="rbc3d_solver_gmres") result(ierr)
      import
      type(c_ptr), value :: ctx
      real(c_double) :: rhs(*), sol(*), history(*)
      real(c_double), value :: rtol
      integer(c_int), value :: restart, maxit
      integer(c_int) :: niter, ierr
    end function
    ! several ranks: the cells whose unknowns this rank holds, in vector order (sharded Krylov vectors)
    function rbc3d_solver_cells(ctx, n, cells, cap) bind(C, name="rbc3d_solver_cells") result(ierr)
      import
      type(c_ptr), value :: ctx
      integer(c_int) :: n, cells(*)
      integer(c_int), value :: cap
      integer(c_int) :: ierr
    end function
    ! NoSlipWall's solve resident on the device (ModNoSlip.F90:44-149); f = walls(:)%f back to back, SoA(3,NV)
    function rbc3d_noslip_solve(ctx, indx, nindep, vbkg, use_cells, rtol, maxit, f, niter, history, slip) &
      bind(C, name="rbc3d_noslip_solve") result(ierr)
      import
      type(c_ptr), value :: ctx
      integer(c_int) :: indx(*)
      integer(c_int), value :: nindep, use_cells, maxit
      real(c_double) :: vbkg(3), f(*), history(*), slip(*)
      real(c_double), value :: rtol
      integer(c_int) :: niter, ierr
    end function
    ! Closest_Neighbor_Cell / Closest_Neighbor_Wall (ModRepulsion.F90:480-613) for n points at once
    function rbc3d_closest_neighbors(ctx, n, x, surfId, epsDist, distCell, x0Cell, distWall, x0Wall) &
      bind(C, name="rbc3d_closest_neighbors") result(ierr)
      import
      type(c_ptr), value :: ctx
      integer(c_int), value :: n
      real(c_double) :: x(*), distCell(*), x0Cell(*), distWall(*), x0Wall(*)
      integer(c_int) :: surfId(*)
      real(c_double), value :: epsDist
      integer(c_int) :: ierr
    end function
    function rbc3d_solver_rhs(ctx, vbkg, use_walls, rhs) bind(C, name="rbc3d_solver_rhs") result(ierr)
      import
      type(c_ptr), value :: ctx
      real(c_double) :: vbkg(3), rhs(*)
      integer(c_int), value :: use_walls
      integer(c_int) :: ierr
    end function
    function rbc3d_solver_velocity(ctx, sol, v) bind(C, name="rbc3d_solver_velocity") result(ierr)
      import
      type(c_ptr), value :: ctx
      real(c_double) :: sol(*), v(*)
      integer(c_int) :: ierr
    end function
    function rbc3d_set_replicated_density(ctx, on) bind(C, name="rbc3d_set_replicated_density") result(ierr)
      import
      type(c_ptr), value :: ctx
      integer(c_int), value :: on
      integer(c_int) :: ierr
    end function
    function rbc3d_cells_set_density(ctx, f, g, spF, spG) bind(C, name="rbc3d_cells_set_density") result(ierr)
      import
      type(c_ptr), value :: ctx
      type(c_ptr), value :: f, g, spF, spG   ! c_null_ptr = unchanged
      integer(c_int) :: ierr
    end function
    function rbc3d_targets_set_raw(ctx, n, x, active) bind(C, name="rbc3d_targets_set_raw") result(ierr)
      import
      type(c_ptr), value :: ctx
      integer(c_int), value :: n
      real(c_double) :: x(*)
      integer(c_int) :: active(*)
      integer(c_int) :: ierr
    end function
    function rbc3d_add_int_on_rbcs(ctx, c1, c2, tlist, v) bind(C, name="rbc3d_add_int_on_rbcs") result(ierr)
      import
      type(c_ptr), value :: ctx
      real(c_double), value :: c1, c2
      integer(c_int), value :: tlist
      real(c_double) :: v(*)
      integer(c_int) :: ierr
    end function
    function rbc3d_pme_distrib_source(ctx, c1, c2, use_cells, use_walls) &
      bind(C, name="rbc3d_pme_distrib_source") result(ierr)
      import
      type(c_ptr), value :: ctx
      real(c_double), value :: c1, c2
      integer(c_int), value :: use_cells, use_walls
      integer(c_int) :: ierr
    end function
    function rbc3d_pme_transform(ctx) bind(C, name="rbc3d_pme_transform") result(ierr)
      import
      type(c_ptr), value :: ctx
      integer(c_int) :: ierr
    end function
    function rbc3d_pme_add_interp_vel(ctx, tlist, v) bind(C, name="rbc3d_pme_add_interp_vel") result(ierr)
      import
      type(c_ptr), value :: ctx
      integer(c_int), value :: tlist
      real(c_double) :: v(*)
      integer(c_int) :: ierr
    end function
    function rbc3d_apply(ctx, c1, c2, use_cells, use_walls, tlist, v) bind(C, name="rbc3d_apply") result(ierr)
      import
      type(c_ptr), value :: ctx
      real(c_double), value :: c1, c2
      integer(c_int), value :: use_cells, use_walls, tlist
      real(c_double) :: v(*)
      integer(c_int) :: ierr
    end function
    function rbc3d_collect_array(ctx, tlist, v) bind(C, name="rbc3d_collect_array") result(ierr)
      import
      type(c_ptr), value :: ctx
      integer(c_int), value :: tlist
      real(c_double) :: v(*)
      integer(c_int) :: ierr
    end function
    function rbc3d_ewald_coeff_sl(ctx, r, A, B) bind(C, name="rbc3d_ewald_coeff_sl") result(ierr)
      import
      type(c_ptr), value :: ctx
      real(c_double), value :: r
      real(c_double) :: A, B
      integer(c_int) :: ierr
    end function
    function rbc3d_ewald_coeff_dl(ctx, r, A) bind(C, name="rbc3d_ewald_coeff_dl") result(ierr)
      import
      type(c_ptr), value :: ctx
      real(c_double), value :: r
      real(c_double) :: A
      integer(c_int) :: ierr
    end function
    function rbc3d_ewald_coeff_sl_exact(r, alpha, A, B) bind(C, name="rbc3d_ewald_coeff_sl_exact") result(ierr)
      import
      real(c_double), value :: r, alpha
      real(c_double) :: A, B
      integer(c_int) :: ierr
    end function
    function rbc3d_ewald_coeff_dl_exact(r, alpha, A) bind(C, name="rbc3d_ewald_coeff_dl_exact") result(ierr)
      import
      real(c_double), value :: r, alpha
      real(c_double) :: A
      integer(c_int) :: ierr
    end function
    function rbc3d_walls_set(ctx, nwall, nvert, nele, x, e2v, area, epsDist, active) &
        bind(C, name="rbc3d_walls_set") result(ierr)
      import
      type(c_ptr), value :: ctx
      integer(c_int), value :: nwall
      integer(c_int) :: nvert(*), nele(*), e2v(*), active(*)
      real(c_double) :: x(*), area(*), epsDist(*)
      integer(c_int) :: ierr
    end function
    function rbc3d_walls_set_traction(ctx, f) bind(C, name="rbc3d_walls_set_traction") result(ierr)
      import
      type(c_ptr), value :: ctx
      real(c_double) :: f(*)
      integer(c_int) :: ierr
    end function
    function rbc3d_wall_prepare_sing(ctx) bind(C, name="rbc3d_wall_prepare_sing") result(ierr)
      import
      type(c_ptr), value :: ctx
      integer(c_int) :: ierr
    end function
    function rbc3d_sing_int_on_wall(ctx, c1, iwall, v) bind(C, name="rbc3d_sing_int_on_wall") result(ierr)
      import
      type(c_ptr), value :: ctx
      real(c_double), value :: c1
      integer(c_int), value :: iwall
      real(c_double) :: v(*)
      integer(c_int) :: ierr
    end function
    function rbc3d_add_int_on_walls(ctx, c1, tlist, v) bind(C, name="rbc3d_add_int_on_walls") result(ierr)
      import
      type(c_ptr), value :: ctx
      real(c_double), value :: c1
      integer(c_int), value :: tlist
      real(c_double) :: v(*)
      integer(c_int) :: ierr
    end function
    function rbc3d_min_dist_to_tri(ctx, n, xtar, xtri, dist, s0, t0) bind(C, name="rbc3d_min_dist_to_tri") result(ierr)
      import
      type(c_ptr), value :: ctx
      integer(c_int), value :: n
      real(c_double) :: xtar(*), xtri(*), dist(*), s0(*), t0(*)
      integer(c_int) :: ierr
    end function
    ! two Fortran views of the same C symbol: s0/t0/lhs may be NULL (c_ptr, by value) or arrays
    function rbc3d_tri_int(ctx, n, xtri, ftri, xtar, s0, t0, rhs, lhs) bind(C, name="rbc3d_tri_int") result(ierr)
      import
      type(c_ptr), value :: ctx, s0, t0
      integer(c_int), value :: n
      real(c_double) :: xtri(*), ftri(*), xtar(*), rhs(*), lhs(*)
      integer(c_int) :: ierr
    end function
    function rbc3d_tri_int_rhs(ctx, n, xtri, ftri, xtar, s0, t0, rhs, lhs) bind(C, name="rbc3d_tri_int") result(ierr)
      import
      type(c_ptr), value :: ctx, s0, t0, lhs
      integer(c_int), value :: n
      real(c_double) :: xtri(*), ftri(*), xtar(*), rhs(*)
      integer(c_int) :: ierr
    end function
    function rbc3d_cells_enable_device_splines(ctx, nlat0) bind(C, name="rbc3d_cells_enable_device_splines") result(ierr)
      import
      type(c_ptr), value :: ctx
      integer(c_int), value :: nlat0
      integer(c_int) :: ierr
    end function
    function rbc3d_apply_assign(ctx, c1, c2, use_cells, use_walls, tlist, v) bind(C, name="rbc3d_apply_assign") result(ierr)
      import
      type(c_ptr), value :: ctx
      real(c_double), value :: c1, c2
      integer(c_int), value :: use_cells, use_walls, tlist
      real(c_double) :: v(*)
      integer(c_int) :: ierr
    end function
    function rbc3d_apply_collect(ctx, c1, c2, use_cells, use_walls, tlist, v) bind(C, name="rbc3d_apply_collect") result(ierr)
      import
      type(c_ptr), value :: ctx
      real(c_double), value :: c1, c2
      integer(c_int), value :: use_cells, use_walls, tlist
      real(c_double) :: v(*)
      integer(c_int) :: ierr
    end function
    function rbc3d_last_error() bind(C, name="rbc3d_last_error") result(msg)
      import
      type(c_ptr) :: msg
    end function
    ! Pin host arrays that are passed every matvec (slist_rbc%g, v, ...): call once with c_loc(array) and its size in
    ! bytes, e.g. ierr = rbc3d_host_register(c_loc(slist_rbc%g), int(8*size(slist_rbc%g), c_size_t))
    function rbc3d_host_register(ptr, bytes) bind(C, name="rbc3d_host_register") result(ierr)
      import
      type(c_ptr), value :: ptr
      integer(c_size_t), value :: bytes
      integer(c_int) :: ierr
    end function
    function rbc3d_host_unregister(ptr) bind(C, name="rbc3d_host_unregister") result(ierr)
      import
      type(c_ptr), value :: ptr
      integer(c_int) :: ierr
    end function
    ! SetEwaldPrms (ModConf.F90:348-408) as the library derives it (host arithmetic): rc and Nb for a box
    function rbc3d_set_ewald_prms(Lb, alpha, eps, P, nranks, rc, Nb) bind(C, name="rbc3d_set_ewald_prms") result(ierr)
      import
      real(c_double) :: Lb(3)
      real(c_double), value :: alpha, eps
      integer(c_int), value :: P, nranks
      real(c_double) :: rc
      integer(c_int) :: Nb(3)
      integer(c_int) :: ierr
    end function
  end interface

contains

  ! the reference ignores ierr everywhere and `stop`s on failure; keep `stop`, but say why
  subroutine B200_Check(ierr, where)
    integer(c_int) :: ierr
    character(*) :: where
    if (ierr /= 0) then
      write (*, *) 'librbc3d_b200: ', where, ' failed with code ', ierr
      stop
    end if
  end subroutine B200_Check

  integer(c_int) function TlistKind(tlist)
    type(t_TargetList), target :: tlist
    if (associated(tlist%x, tlist_rbc%x)) then
      TlistKind = TL_CELLS
    else if (associated(tlist%x, tlist_wall%x)) then
      TlistKind = TL_WALLS
    else
      TlistKind = TL_RAW
    end if
  end function TlistKind

  ! Create the context (cuFFT plans, meshes, B-spline moduli, Ewald tables) and join the communicator -- lazily and
  ! idempotently: TimeInt_Init calls PrepareSingIntOnWall (ModTimeInt.F90:87-91) BEFORE PME_Init (:96), and
  ! PostProcess / field_visual programs may touch the lists before either, so every procedure that needs the context
  ! calls B200_EnsureInit first; PME_Init is then a no-op when the context already exists.  The communicator is
  ! attached here, before any geometry reaches the library (rbc3d_ctx_attach_comm requires that order).
  subroutine B200_EnsureInit
    integer(c_int) :: ierr, device, mpierr
    character(kind=c_char) :: id(128)
    if (c_associated(b200_ctx)) return
    device = -1                                   ! the launcher binds one GPU per rank (CUDA_VISIBLE_DEVICES)
    ierr = rbc3d_ctx_create(b200_ctx, Lb, alpha_Ewd, eps_Ewd, PBspln_Ewd, rc_Ewd, Nb_Ewd, device)
    call B200_Check(ierr, 'rbc3d_ctx_create')
    if (numNodes > 1) then
      ierr = 0
      if (nodeNum == 0) ierr = rbc3d_comm_unique_id(id)
      call B200_Check(ierr, 'rbc3d_comm_unique_id')
      call MPI_Bcast(id, 128, MPI_CHARACTER, 0, MPI_COMM_WORLD, mpierr)
      ierr = rbc3d_ctx_attach_comm(b200_ctx, numNodes, nodeNum, id)
      call B200_Check(ierr, 'rbc3d_ctx_attach_comm')
    end if
    if (nrbc > 0) then
      ierr = rbc3d_cells_set_mesh(b200_ctx, nrbc, rbcs(1)%nlat, rbcs(1)%nlon, rbcs(1)%th, rbcs(1)%phi, rbcs(1)%w)
      call B200_Check(ierr, 'rbc3d_cells_set_mesh')
    end if
  end subroutine B200_EnsureInit

  ! PME_Init (ModPME.F90:252-338)
  subroutine B200_Init
    call B200_EnsureInit
  end subroutine B200_Init

  ! SourceList_UpdateCoord(slist_rbc) + TargetList_Update(tlist_rbc): gather the per-cell splines into the ABI
  ! layout [cell][u,u1,u2,u12][var][nlon][2 nlat] and hand geometry + ownership flags to the library
  subroutine B200_SyncCells
    real(WP), allocatable :: spx(:), spa3(:), spdj(:), Acell(:), Bcell(:), area(:), msize(:)
    integer(c_int), allocatable :: act(:)
    integer :: irbc, m, n, o3, o1, ierr
    type(t_rbc), pointer :: rbc
    call B200_EnsureInit
    m = 2*rbcs(1)%nlat; n = rbcs(1)%nlon
    allocate (spx(12*m*n*nrbc), spa3(12*m*n*nrbc), spdj(4*m*n*nrbc))
    allocate (Acell(nrbc), Bcell(nrbc), area(nrbc), msize(nrbc), act(tlist_rbc%nPoint))
    do irbc = 1, nrbc
      rbc => rbcs(irbc)
      o3 = 12*m*n*(irbc - 1); o1 = 4*m*n*(irbc - 1)
      call PackSpline(rbc%spln_x, 3, spx(o3 + 1:o3 + 12*m*n))
      call PackSpline(rbc%spln_a3, 3, spa3(o3 + 1:o3 + 12*m*n))
      call PackSpline(rbc%spln_detJ, 1, spdj(o1 + 1:o1 + 4*m*n))
      Acell(irbc) = Acoef(rbc%celltype); Bcell(irbc) = Bcoef(rbc%celltype)
      area(irbc) = rbc%area; msize(irbc) = rbc%meshSize
    end do
    act = merge(1, 0, tlist_rbc%active)
    ierr = rbc3d_cells_set_geometry(b200_ctx, slist_rbc%x, slist_rbc%a3, Acell, Bcell, area, msize, &
                                    spx, spa3, spdj, act)
    call B200_Check(ierr, 'rbc3d_cells_set_geometry')
  end subroutine B200_SyncCells

  ! SourceList_UpdateDensity(slist_rbc, UpdateF, UpdateG) (+ the density splines of Rbc_BuildSurfaceSource)
  subroutine B200_SyncDensity(updateF, updateG)
    logical, optional :: updateF, updateG
    real(WP), allocatable, target :: spF(:), spG(:)
    type(c_ptr) :: pf, pg, psf, psg
    integer :: irbc, m, n, o3, ierr
    call B200_EnsureInit
    m = 2*rbcs(1)%nlat; n = rbcs(1)%nlon
    pf = c_null_ptr; pg = c_null_ptr; psf = c_null_ptr; psg = c_null_ptr
    if (present(updateF)) then
      if (updateF) then
        allocate (spF(12*m*n*nrbc))
        do irbc = 1, nrbc
          o3 = 12*m*n*(irbc - 1)
          call PackSpline(rbcs(irbc)%spln_FdetJ, 3, spF(o3 + 1:o3 + 12*m*n))
        end do
        pf = c_loc(slist_rbc%f); psf = c_loc(spF)
      end if
    end if
    if (present(updateG)) then
      if (updateG) then
        allocate (spG(12*m*n*nrbc))
        do irbc = 1, nrbc
          o3 = 12*m*n*(irbc - 1)
          call PackSpline(rbcs(irbc)%spln_GdetJ, 3, spG(o3 + 1:o3 + 12*m*n))
        end do
        pg = c_loc(slist_rbc%g); psg = c_loc(spG)
      end if
    end if
    ierr = rbc3d_cells_set_density(b200_ctx, pf, pg, psf, psg)
    call B200_Check(ierr, 'rbc3d_cells_set_density')
  end subroutine B200_SyncDensity

  ! SourceList_UpdateCoord(slist_wall, walls) + TargetList_Update(tlist_wall, walls): wall meshes are static, so this
  ! runs once (from PrepareSingIntOnWall).  tlist_wall%x is the concatenation of wall%x the ABI expects; e2v, area
  ! and epsDist are packed wall after wall.
  subroutine B200_SyncWalls
    integer(c_int), allocatable :: nv(:), ne(:), e2v(:), act(:)
    real(WP), allocatable :: area(:), eps(:)
    integer :: iwall, NE, p, l, ierr
    call B200_EnsureInit
    allocate (nv(nwall), ne(nwall))
    do iwall = 1, nwall
      nv(iwall) = walls(iwall)%nvert; ne(iwall) = walls(iwall)%nele
    end do
    NE = sum(ne)
    allocate (e2v(3*NE), area(NE), eps(NE), act(tlist_wall%nPoint))
    p = 0
    do iwall = 1, nwall
      do l = 1, 3
        e2v((l - 1)*NE + p + 1:(l - 1)*NE + p + ne(iwall)) = walls(iwall)%e2v(:, l)
      end do
      area(p + 1:p + ne(iwall)) = walls(iwall)%area
      eps(p + 1:p + ne(iwall)) = walls(iwall)%epsDist
      p = p + ne(iwall)
    end do
    act = merge(1, 0, tlist_wall%active)
    ierr = rbc3d_walls_set(b200_ctx, nwall, nv, ne, tlist_wall%x, e2v, area, eps, act)
    call B200_Check(ierr, 'rbc3d_walls_set')
  end subroutine B200_SyncWalls

  ! wall%f of every wall, concatenated like tlist_wall%x
  subroutine B200_SyncWallTraction
    real(WP), allocatable :: f(:, :)
    integer :: iwall, p, ierr
    call B200_EnsureInit
    allocate (f(tlist_wall%nPoint, 3))
    p = 0
    do iwall = 1, nwall
      f(p + 1:p + walls(iwall)%nvert, :) = walls(iwall)%f
      p = p + walls(iwall)%nvert
    end do
    ierr = rbc3d_walls_set_traction(b200_ctx, f)
    call B200_Check(ierr, 'rbc3d_walls_set_traction')
  end subroutine B200_SyncWallTraction

  ! t_spline (ModDataStruct.F90:37-43) u, u1, u2, u12 (0:m-1, 0:n-1, nvar) -> [4][nvar][n][m], m fastest
  subroutine PackSpline(spln, nvar, buf)
    type(t_spline) :: spln
    integer :: nvar
    real(WP) :: buf(:)
    integer :: sz
    sz = size(spln%u)
    buf(1:sz) = reshape(spln%u, (/sz/))
    buf(sz + 1:2*sz) = reshape(spln%u1, (/sz/))
    buf(2*sz + 1:3*sz) = reshape(spln%u2, (/sz/))
    buf(3*sz + 1:4*sz) = reshape(spln%u12, (/sz/))
  end subroutine PackSpline

end module ModB200
