/*
 * rbc3d.h -- C ABI of librbc3d_b200.so: the B200-native (sm_100a, FP64) Ewald boundary-integral operator of
 * comp-physics/RBC3D.  This is the drop-in boundary: the reference has no FFI layer, its boundary is the
 * set of Fortran module procedures of ModPME, ModEwaldFunc, ModIntOnRbcs, ModIntOnWalls (+ the list modules
 * ModSourceList / ModTargetList / ModHashTable they read).  Every entry point below names the reference
 * procedure it replaces (paths relative to the reference's common/).  The fortran/ directory holds the ISO_C_BINDING
 * shims that re-export the reference names; INTEGRATION.md shows how they are wired.
 *
 * Conventions
 *  - plain pointers and sizes only; all floating point is double (real(WP)), integers are int32 (default
 *    Fortran integer); host arrays stay owned by the caller, the library keeps device mirrors.
 *  - "SoA(3,N)" = three contiguous planes of length N = Fortran x(N,3).
 *  - cell point index p = cell*nlat*nlon + (ilon-1)*nlat + (ilat-1)      (ModSourceList.F90:110-121)
 *  - spline of a cell with nvar variables: [4 (u,u1,u2,u12)][nvar][nlon][2*nlat], theta index fastest
 *    = the four t_spline component arrays of ModDataStruct.F90:37-43 back to back.
 *  - velocities are ACCUMULATED into v (callers zero it, ModVelSolver.F90:473,571), rows of inactive
 *    targets are left untouched (ModTargetList.F90:196-197).
 *  - every function returns 0 on success, a negative RBC3D_E* code otherwise (the reference `stop`s).
 *  - not re-entrant per context (the reference is single-threaded module state); one context per GPU.
 */
#ifndef RBC3D_H
#define RBC3D_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rbc3d_ctx rbc3d_ctx;

enum {
  RBC3D_OK = 0,
  RBC3D_EINVAL = -1,    /* bad argument / call order */
  RBC3D_ECUDA = -2,     /* CUDA runtime / cuFFT / NCCL failure (see rbc3d_last_error) */
  RBC3D_ENOMEM = -3,
  RBC3D_ESTATE = -4,    /* required state not set (e.g. PME_Transform before PME_Distrib_Source) */
  RBC3D_EOVERFLOW = -5  /* neighbour-cell list exceeded 32 entries (ModRbcSingInt.F90:319) */
};

/* target list kinds (which t_TargetList of ModData.F90:48-49 an operator call refers to) */
enum { RBC3D_TL_CELLS = 0, RBC3D_TL_RAW = 1, RBC3D_TL_WALLS = 2 };

/* stage ids for rbc3d_get_timings (milliseconds of the last operator application, CUDA events) */
enum {
  RBC3D_T_PAIR = 0, RBC3D_T_SING, RBC3D_T_NEARSING, RBC3D_T_LINEAR, RBC3D_T_SPREAD, RBC3D_T_FFT,
  RBC3D_T_KSPACE, RBC3D_T_FFT_INV, RBC3D_T_INTERP, RBC3D_T_COMBINE, RBC3D_T_WALL, RBC3D_T_COMM, RBC3D_T_H2D,
  RBC3D_T_D2H, RBC3D_T_DENSITY, RBC3D_T_TOTAL,
  /* where the PME chain (spread .. interpolate) runs on its own stream beside the real-space chain (pair, singular,
   * near-singular, linear) -- several ranks, or large lists on one rank, see rbc3d_set_overlap --: their two critical
   * paths, each measured on its own stream */
  RBC3D_T_PME_CHAIN, RBC3D_T_REAL_CHAIN, RBC3D_T_COUNT
};

const char *rbc3d_last_error(void);
int rbc3d_version(void);

/* ---- parameters: ModConf.F90:348-408 SetEwaldPrms (host arithmetic, no GPU needed) ---- */
int rbc3d_set_ewald_prms(const double Lb[3], double alpha, double eps, int P, int nranks, double *rc,
                         int Nb[3]);

/* ---- context: replaces the module state of ModConf/ModData/ModPME.  PME_Init / PME_Finalize
 *      (ModPME.F90:252-338, 342-350) happen inside create/destroy. device < 0 = current device. ---- */
int rbc3d_ctx_create(rbc3d_ctx **ctx, const double Lb[3], double alpha, double eps, int P, double rc,
                     const int Nb[3], int device);
int rbc3d_ctx_destroy(rbc3d_ctx *ctx);

/* Multi-GPU: one context per rank; the 128-byte NCCL unique id is created on rank 0 and distributed by the host
 * program (MPI_Bcast in the Fortran driver, torch.distributed in the Python harness).  rbc3d_ctx_attach_comm must come
 * BEFORE any geometry (cells, walls, raw targets) reaches the context: ownership is derived from the rank when a list
 * is built.  Decomposition (ModConf.F90:412-437 DomainDecomp; ModPFFTW.F90:56-89): the PME mesh is split into z-slabs of
 * planes -- a rank spreads the sources whose B-spline support touches its planes (no mesh reduction), transforms its
 * planes in (x, y), the spectra change hands to y-slabs by an all-to-all (grouped ncclSend/ncclRecv) for the transform in
 * z and the k-space multiplier, and back; velocity planes beyond a rank's slab that its targets interpolate from come
 * from their owners.  Targets: the caller's `active` flags (SetActiveFlag), per point or -- as the harness does -- whole
 * cells by the z-slab of their centroid.  Source lists stay replicated, as in the reference. */
int rbc3d_comm_unique_id(void *id128);
int rbc3d_ctx_attach_comm(rbc3d_ctx *ctx, int nranks, int rank, const void *id128);
/* TargetList_CollectArray(tlist, 3, v, MPI_COMM_WORLD) (ModTargetList.F90:172-202): v <- sum over ranks of the
 * per-rank arrays (each rank holds only the rows of its active targets); host SoA(3,n).  No-op on one rank.
 * rbc3d_apply_resident includes this step. */
int rbc3d_collect_array(rbc3d_ctx *ctx, int tlist, double *v);

/* ---- ModEwaldFunc.F90:13-16 (host scalars, the table versions use the context's alpha, rc) ---- */
int rbc3d_ewald_coeff_sl_exact(double r, double alpha, double *A, double *B);
int rbc3d_ewald_coeff_dl_exact(double r, double alpha, double *A);
int rbc3d_ewald_coeff_sl(const rbc3d_ctx *ctx, double r, double *A, double *B);
int rbc3d_ewald_coeff_dl(const rbc3d_ctx *ctx, double r, double *A);

/* ---- cells ----
 * rbc3d_cells_set_mesh: mesh shared by all cells (RBC_Create, ModRbc.F90:44-110) + the shared polar patch
 *   (RbcPolarPatch_Create, ModPolarPatch.F90:27-76; TimeInt_Init, ModTimeInt.F90:62).
 * rbc3d_cells_set_geometry: SourceList_UpdateCoord + TargetList_Update for tlist_rbc (ModSourceList.F90:
 *   92-151, ModTargetList.F90:95-135) incl. the cell list build (HashTable_Build, ModHashTable.F90:23-58), and
 *   the geometry splines written by Rbc_BuildSurfaceSource(xFlag) (ModRbc.F90:736-762).  active may be NULL
 *   (all targets active; SetActiveFlag, ModTargetList.F90:205-233).
 * rbc3d_cells_set_density: SourceList_UpdateDensity (ModSourceList.F90:160-187; f, g already multiplied by
 *   detJ*w) + the density splines of Rbc_BuildSurfaceSource(fFlag/gFlag).  NULL = unchanged.             */
int rbc3d_cells_set_mesh(rbc3d_ctx *ctx, int ncell, int nlat, int nlon, const double *th, const double *phi,
                         const double *w);
int rbc3d_cells_set_geometry(rbc3d_ctx *ctx, const double *x, const double *a3, const double *Acoef_cell,
                             const double *Bcoef_cell, const double *area, const double *meshSize,
                             const double *spx, const double *spa3, const double *spdetj,
                             const int32_t *active);
int rbc3d_cells_set_density(rbc3d_ctx *ctx, const double *f, const double *g, const double *spF,
                            const double *spG);
/* rbc3d_cells_set_geometry with Rbc_BuildSurfaceSource(xFlag) (ModRbc.F90:736-762) done on the device: the caller
 * passes the mesh field detJ(Np) (rbc%detj, point order of x) instead of the splines of x, a3 and detJ, which are built
 * with the same operator as the density splines.  Needs rbc3d_cells_enable_device_splines. */
int rbc3d_cells_set_geometry_mesh(rbc3d_ctx *ctx, const double *x, const double *a3, const double *detj,
                                  const double *Acoef_cell, const double *Bcoef_cell, const double *area,
                                  const double *meshSize, const int32_t *active);
/* Optional: Rbc_BuildSurfaceSource(fFlag/gFlag) on the device (ModRbc.F90:760-802: ShAnalGau + ShFilter(nlat0) +
 * ShSynthEqu + Spline_Build_on_Sphere, ModSpline.F90:121-142, FFT_Diff, ModFFT.F90:25-93).  After this call a
 * density passed to rbc3d_cells_set_density with a NULL spline gets its spline built on the GPU (instead of keeping
 * the previous one), which removes the 2.16 GB spline upload per GMRES matvec at 4096 cells (the matvec only needs
 * g, ModVelSolver.F90:560-565).  nlat0 = the cells' spherical-harmonic order (ModRbc.F90:55). */
int rbc3d_cells_enable_device_splines(rbc3d_ctx *ctx, int nlat0);

/* ---- walls ----
 * rbc3d_walls_set: SourceList_UpdateCoord(slist_wall, walls) (element centroids + cell list, ModSourceList.F90:
 *   127-146) and TargetList_Update(tlist_wall, walls) (all wall vertices, Acoef = 2, ModTargetList.F90:122-131).
 *   x: SoA(3,NV), the vertices of all walls back to back (= tlist_wall%x); e2v: SoA(3,NE) = wall%e2v of every wall
 *   back to back, 1-based vertex numbers local to the wall; area, epsDist [NE] from Wall_ComputeGeometry
 *   (ModWall.F90:118-145); active [NV] or NULL = SetActiveFlag of the wall target list.
 * rbc3d_walls_set_traction: wall%f of all walls, SoA(3,NV).
 * rbc3d_wall_prepare_sing: PrepareSingIntOnWall (ModIntOnWalls.F90:181-308) for every wall: the sparse
 *   self-interaction matrix t_Wall%lhs is assembled in HBM (rows of active vertices only, :199-203).
 * rbc3d_sing_int_on_wall: SingIntOnWall(c1, wall, v) (:136-172): v = c1 * lhs * f, v host SoA(3,nvert(iwall)),
 *   iwall 0-based.
 * rbc3d_add_int_on_walls: AddIntOnWalls(c1, tlist, v) (:33-130): self-interactions through lhs when tlist is the
 *   wall list, the direct Duffy / 7-point loop over all other (target, element) pairs within rc.           */
int rbc3d_walls_set(rbc3d_ctx *ctx, int nwall, const int32_t *nvert, const int32_t *nele, const double *x,
                    const int32_t *e2v, const double *area, const double *epsDist, const int32_t *active);
int rbc3d_walls_set_traction(rbc3d_ctx *ctx, const double *f);
int rbc3d_wall_prepare_sing(rbc3d_ctx *ctx);
int rbc3d_sing_int_on_wall(rbc3d_ctx *ctx, double c1, int iwall, double *v);
int rbc3d_add_int_on_walls(rbc3d_ctx *ctx, double c1, int tlist, double *v);
/* MinDistToTri (:480-577), Tri_Int_Regular (:319-363), Tri_Int_Duffy (:373-465) for n independent
 * (target, triangle) pairs, evaluated on the device.  xtar SoA(3,n); xtri, ftri [n][3 corners][3]; s0 = t0 = NULL
 * selects the regular rule; rhs [n][3] and/or lhs [n][3 corners][3][3] (either may be NULL). */
int rbc3d_min_dist_to_tri(rbc3d_ctx *ctx, int n, const double *xtar, const double *xtri, double *dist, double *s0,
                          double *t0);
int rbc3d_tri_int(rbc3d_ctx *ctx, int n, const double *xtri, const double *ftri, const double *xtar,
                  const double *s0, const double *t0, double *rhs, double *lhs);

/* TargetList_CreateFromRaw (ModTargetList.F90:139-169): arbitrary points, Acoef = 2, indx = -1 */
int rbc3d_targets_set_raw(rbc3d_ctx *ctx, int n, const double *x, const int32_t *active);

/* ---- operator pieces, same names and call protocol as the reference ----
 * AddIntOnRbcs(c1,c2,tlist,v)      ModIntOnRbcs.F90:25-158 (pair sum, singular, near-singular, linear term)
 * PME_Distrib_Source(c1,c2,cells,walls) ModPME.F90:58-133; PME_Transform :137-222; PME_Add_Interp_Vel :228-248
 * v: host SoA(3,n) of the chosen target list, accumulated into.                                          */
int rbc3d_add_int_on_rbcs(rbc3d_ctx *ctx, double c1, double c2, int tlist, double *v);
int rbc3d_pme_distrib_source(rbc3d_ctx *ctx, double c1, double c2, int use_cells, int use_walls);
int rbc3d_pme_transform(rbc3d_ctx *ctx);
int rbc3d_pme_add_interp_vel(rbc3d_ctx *ctx, int tlist, double *v);

/* Fused operator application (what MyMatMult / Compute_Rhs do between zeroing v and CollectArray,
 * ModVelSolver.F90:473-493, 568-584): v += AddIntOnRbcs [+ AddIntOnWalls] + PME.  Host v. */
int rbc3d_apply(rbc3d_ctx *ctx, double c1, double c2, int use_cells, int use_walls, int tlist, double *v);
/* Same, but v is ASSIGNED (rows of inactive targets = 0): the caller's "v = 0" is folded into the call and v is not
 * uploaded (saves 24 B per target of host->device traffic per application). */
int rbc3d_apply_assign(rbc3d_ctx *ctx, double c1, double c2, int use_cells, int use_walls, int tlist, double *v);
/* Same as rbc3d_apply_assign followed by TargetList_CollectArray (ModTargetList.F90:172-202, the call right after the
 * operator in MyMatMult / Compute_Rhs, ModVelSolver.F90:493, 584): with several ranks the rows are summed on the
 * devices (ncclAllReduce) before the single device->host copy, so every rank receives the complete v. */
int rbc3d_apply_collect(rbc3d_ctx *ctx, double c1, double c2, int use_cells, int use_walls, int tlist, double *v);
/* Same with everything resident: result written (not accumulated) to the context's device velocity buffer;
 * rbc3d_get_velocity copies it out.  Used by the benchmark's device-resident timing. */
int rbc3d_apply_resident(rbc3d_ctx *ctx, double c1, double c2, int use_cells, int use_walls, int tlist);
int rbc3d_get_velocity(rbc3d_ctx *ctx, int tlist, double *v);
/* bit mask to skip parts of AddIntOnRbcs (testing / profiling): 1 singular, 2 near-singular, 4 linear, 8 pairs */
/* The PME chain (spread .. interpolate) on a second stream beside the real-space kernels, joined before the combine.
 * mode -1 (default): with several ranks, and on one rank for target lists of 100 000 points or more; 0: never (one
 * stream: the per-stage times of rbc3d_get_timings are then each kernel's own); 1: always.  The environment variable
 * RBC3D_OVERLAP sets the initial mode. */
int rbc3d_set_overlap(rbc3d_ctx *ctx, int mode);
int rbc3d_set_skip_flags(rbc3d_ctx *ctx, int flags);
/* Singular double-layer patch integrals (RBC_SingInt, ModRbcSingInt.F90:29-90): mode 1 (default) caches the
 * density-independent factors of every patch point at geometry time (32 B per point, 8.5 MB per 36x72 cell) when
 * device memory allows, so that GMRES matvecs only interpolate the density; mode 0 always evaluates directly.
 * Takes effect at the next rbc3d_cells_set_geometry. */
int rbc3d_set_sing_cache(rbc3d_ctx *ctx, int mode);
/* state of that cache: 1 if the cached path is active, and the patch points per target it streams (those with a
 * non-zero quadrature weight: the mask table vanishes on its last interval, so the outermost radial node of every ray
 * contributes exactly zero and is left out; 32 B each per matvec) */
int rbc3d_sing_cache_info(rbc3d_ctx *ctx, int32_t *cached, int32_t *points_per_target);
/* Same-surface pairs of the real-space sum (ModIntOnRbcs.F90:81-84): mode 3 (default) symmetric patch-pair kernel
 * (each unordered pair evaluated once) that, for the double-layer operator alone (the GMRES matvec), streams a
 * per-geometry cache of (1 - mask) * EwaldCoeff_DL per unordered pair (8 B per pair slot, ~16 MB per 36x72 cell; built
 * at rbc3d_cells_set_geometry for as many cells as device memory allows, the rest is evaluated directly);
 * mode 1 the same kernel without the cache, mode 2 dense per-cell kernel, mode 0 through the hashed cell list like
 * every other pair (testing). */
int rbc3d_set_pair_self(rbc3d_ctx *ctx, int mode);
/* Several ranks: declare that the host densities passed to rbc3d_cells_set_density are identical on all ranks (the
 * reference replicates rbc%f / rbc%g, ModVelSolver.F90:552-565).  Each rank then uploads only the rows of its own
 * cell block and the blocks are all-gathered over NVLink; rbc3d_cells_set_density becomes a collective call.
 * Needs ncell divisible by the number of ranks (otherwise the full arrays are uploaded as before). */
int rbc3d_set_replicated_density(rbc3d_ctx *ctx, int on);
/* state of that cache (built by the first double-layer-only application after rbc3d_cells_set_geometry, so a time
 * step without a cell solve never builds it): cells cached (of the cells this rank owns targets of) and
 * 256-byte coefficient rows held (= bytes streamed per matvec / 256); either pointer may be NULL */
int rbc3d_pair_cache_info(rbc3d_ctx *ctx, int32_t *cells_cached, int64_t *rows);

/* ---- the cell velocity solve with Krylov vectors and SH transforms resident on the device ----
 * rbc3d_solver_setup: once per geometry, after rbc3d_cells_set_geometry[_mesh] and
 *   rbc3d_cells_enable_device_splines; nlat0 = SH truncation of the unknowns, detj = rbc%detj on the mesh (Np).
 * rbc3d_solver_matmult: b = MyMatMult(u) (ModVelSolver.F90:523-601): Glob_Sph_Trans FOUR_TO_PHYS, operator #2
 *   (c1 = 0, c2 = -1/4pi), + g, Glob_Sph_Trans PHYS_TO_FOUR; u, b host vectors of ncell*3*nlat0^2 doubles packed as
 *   in Glob_Sph_Trans (ModVelSolver.F90:641-719).
 * rbc3d_solver_gmres: KSPSolve of Solve_RBC_Vel (ModVelSolver.F90:74-116) with PETSc's GMRES defaults restated
 *   (restart as given, classical Gram-Schmidt, no preconditioner, residual test rtol*||rhs||); sol: initial guess
 *   in, solution out; history (maxit+1 doubles or NULL): residual norm after every iteration.
 * Several ranks: the unknowns are SHARDED.  A rank holds the coefficients of the cells it owns targets of (whole cells,
 *   every cell active on exactly one rank; rbc3d_solver_cells lists them in vector order), so u, b, rhs and sol are
 *   vectors of rbc3d_solver_dof = owned cells * 3 nlat0^2 doubles; inside a matvec the synthesised densities are
 *   all-gathered over NVLink, the operator rows stay on their rank (no CollectArray, no device->host copy of v) and
 *   the dot products of GMRES are all-reduced.  All solver calls are collective. */
int rbc3d_solver_setup(rbc3d_ctx *ctx, int nlat0, const double *detj);
int rbc3d_solver_dof(rbc3d_ctx *ctx, int64_t *dof);
int rbc3d_solver_cells(rbc3d_ctx *ctx, int32_t *n, int32_t *cells, int cap);
int rbc3d_solver_matmult(rbc3d_ctx *ctx, const double *u, double *b);
int rbc3d_solver_gmres(rbc3d_ctx *ctx, const double *rhs, double *sol, double rtol, int restart, int maxit,
                       int *niter, double *history);
/* rhs = Compute_Rhs (ModVelSolver.F90:455-515): operator #1 (c1 = 1/4pi, c2 = 0; walls' single layer too when
 * use_walls) on the resident single-layer density, + 2 vBkg / Acoef, PHYS_TO_FOUR -- packed coefficients out */
int rbc3d_solver_rhs(rbc3d_ctx *ctx, const double vbkg[3], int use_walls, double *rhs);
/* v = Glob_Sph_Trans(sol, FOUR_TO_PHYS): the surface velocity of a solution (ModVelSolver.F90:124), host SoA(3,Np) */
int rbc3d_solver_velocity(rbc3d_ctx *ctx, const double *sol, double *v);

/* ---- NoSlipWall on the device (ModNoSlip.F90:44-149): the wall-traction solve every time step runs after the cell
 * velocities.  rhs = -Compute_Wall_Residual_Vel (:153-195: operator #3, c1 = c2 = 1/4pi, cells [use_cells] + walls ->
 * wall vertices, + vbkg), MyMatMult (:255-308: operator #4, c1 = 1/4pi, walls -> wall vertices), GMRES with PETSc's
 * defaults restated (restart 30, classical Gram-Schmidt, no preconditioner, zero initial guess, rtol as given -- the
 * reference passes eps_Ewd --, at most maxit iterations), wall%f = wall%f + df.  Tractions, Krylov vectors and both
 * operators stay on the device; only the Hessenberg column crosses to the host per iteration.
 * indx_vert_glb[NV]: indxVertGlb of ModData.F90:50-63 (1-based number of every wall vertex, periodic duplicates found
 * by Wall_Build_V2V sharing their master's number), nindep = number of independent vertices; AssembleArray
 * (:362-384) semantics: towards the unknown vector the last duplicate wins.  f: host SoA(3,NV), wall%f in, updated
 * wall%f out (also left as the context's wall traction); history (maxit + 1 doubles) and slip (host SoA(3,NV): the
 * residual wall velocity after the update, :140-146) may be NULL.  Needs rbc3d_walls_set and rbc3d_wall_prepare_sing;
 * several ranks: collective, every rank passes the same f (the rows of v are summed over the ranks). */
int rbc3d_noslip_solve(rbc3d_ctx *ctx, const int32_t *indx_vert_glb, int nindep, const double vbkg[3], int use_cells,
                       double rtol, int maxit, double *f, int *niter, double *history, double *slip);

/* ---- closest-neighbour queries of ModRepulsion on the GPU cell lists ----
 * Closest_Neighbor_Cell (ModRepulsion.F90:480-546) and Closest_Neighbor_Wall (:556-613) for n points at once (the
 * reference calls them point by point, 4x per time step over all cell points: InterCellRepulsion :270-402,
 * LeukWallRepulsion :405-470).  x: SoA(3,n); surf_id[n]: surface the point lies on (cells 1..ncell, walls ncell+1..;
 * elements / cells of that surface are skipped); eps_dist: the threshold below which the closest mesh point is refined
 * by Spline_FindProjection.  Out: dist_cell[n] / dist_wall[n] (HUGE_VAL where the 27 neighbouring list cells hold
 * no other surface), x0_cell / x0_wall SoA(3,n) the closest points.  Either output pair may be NULL. */
int rbc3d_closest_neighbors(rbc3d_ctx *ctx, int n, const double *x, const int32_t *surf_id, double eps_dist,
                            double *dist_cell, double *x0_cell, double *dist_wall, double *x0_wall);

/* ---- introspection (tests, profiling) ---- */
/* cell list of the cell sources: cid[Np] (0-based, i1 fastest), order[Np] (source indices sorted by cell,
 * ascending index inside a cell), start[Nc1*Nc2*Nc3+1]; any pointer may be NULL */
int rbc3d_cell_list_get(rbc3d_ctx *ctx, int32_t Nc[3], int32_t *cid, int32_t *order, int32_t *start);
/* per cell target: number of sources within rc and an order-independent checksum of their indices */
int rbc3d_neighbor_signature(rbc3d_ctx *ctx, int tlist, int32_t *count, uint64_t *sig);
/* near-singular entries found at geometry time: n entries (target, source cell, active flag, th0, phi0, dist) */
int rbc3d_nearsing_get(rbc3d_ctx *ctx, int tlist, int *n, int32_t *target, int32_t *cell, int32_t *flag,
                       double *th0, double *phi0, double *dist, int cap);
/* wall self-interaction matrix as block rows over the NV vertices: rowptr[NV+1], col[nblk] (global vertex),
 * val[nblk][3][3] (target component, source component) */
int rbc3d_wall_matrix_get(rbc3d_ctx *ctx, int32_t *nblk, int32_t *rowptr, int32_t *col, double *val, int cap);
/* per target of a list: number of wall elements within rc (MinDistToTri), checksum of their indices, how many take
 * the Duffy rule; self_skip = 1 applies the same-surface exclusion of AddIntOnWalls (:92) */
int rbc3d_wall_neighbor_signature(rbc3d_ctx *ctx, int tlist, int self_skip, int32_t *count, uint64_t *sig,
                                  int32_t *nduffy);
/* device copy of a density spline in the ABI layout: which = 0 spline(f detJ), 1 spline(g detJ) */
int rbc3d_cells_get_density_spline(rbc3d_ctx *ctx, int which, double *sp);
/* device copy of the geometry splines (which = 0 x, 1 a3, 2 detJ) in the ABI layout */
int rbc3d_cells_get_geometry_spline(rbc3d_ctx *ctx, int which, double *sp);
int rbc3d_pme_get_grid(rbc3d_ctx *ctx, double *vv /* [3][Nz][Ny][Nx] */);
int rbc3d_get_timings(rbc3d_ctx *ctx, float ms[RBC3D_T_COUNT]);
int rbc3d_get_launch_count(rbc3d_ctx *ctx, long long *launches);
/* Pin caller-owned host arrays that are passed repeatedly (the Fortran driver allocates its arrays once for the
 * whole run), so host<->device copies run at full PCIe rate.  Optional; unregistered memory works, slower. */
int rbc3d_host_register(void *ptr, size_t bytes);
int rbc3d_host_unregister(void *ptr);
/* measured FP64 FMA throughput of this GPU in TFLOP/s (register-resident DFMA chains) */
int rbc3d_measure_fp64_peak(int device, double *tflops);

#ifdef __cplusplus
}
#endif
#endif
