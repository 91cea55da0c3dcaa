/*
 * rbc3d_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C99 + OpenMP) of the Ewald boundary-integral operator of
 * comp-physics/RBC3D (Fortran).  Used only as (1) the parity checker of the CUDA
 * path in tests/ and __graft_entry__.smoke(), and (2) the CPU baseline / reference
 * arm timed by bench.py.  The product (rbc3d_b200/, librbc3d_b200.so) never links,
 * imports or calls anything in this directory.
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures for
 * this path and cannot be compiled in this image (no
 * Fortran compiler, MPI, PETSc, FFTW).  The restatement is validated by physics
 * identities (tests/test_oracle_identities.py) and by a second restatement written
 * separately in NumPy, straight from the Fortran, that agrees with this one to
 * round-off (singular / near-singular integrals, projection and the whole of
 * AddIntOnRbcs per target: tests/test_oracle_singint_numpy.py; Duffy rule and the
 * wall loop: tests/test_oracle_walls_numpy.py; PME: oracle/slabpme.py +
 * tests/test_slab_pme.py).  The one output of the reference that its tree ships,
 * SickleCell.dat, pins the Gauss grid, the point order and the SH truncation
 * (tests/test_reference_golden.py).  Nothing here was checked against an
 * execution of the reference binary.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference/common/).  Index conventions follow the Fortran (1-based
 * ilat/ilon/surface ids in `indx` arrays; -1 = "none").
 *
 * Array layouts (all double unless stated; "SoA(3,N)" = three contiguous planes of
 * length N, i.e. Fortran x(N,3)):
 *   point index of a cell point: p = cell*nlat*nlon + (ilon-1)*nlat + (ilat-1)
 *                                                   (ModSourceList.F90:110-121)
 *   spline of a cell, nvar variables: [4 (u,u1,u2,u12)][nvar][n=nlon][m=2*nlat],
 *                                     m fastest (ModSpline.F90:27-40)
 *   patch tables thG/phiG: [ilon][ilat][iazm][irad], irad fastest
 *                                     (ModPolarPatch.F90:60-61)
 */
#ifndef RBC3D_ORACLE_H
#define RBC3D_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NTAB 8192
#define ORC_NBR_MAX 32

typedef struct {
  /* box + Ewald parameters (ModConf.F90) */
  double Lb[3], iLb[3];
  double alpha, eps, rc;
  int P;
  int Nb[3];
  /* cell list (ModHashTable.F90), single-node semantics */
  int Nc[3];
  double iLbNc[3];
  /* lookup tables (ModEwaldFunc.F90:96-106,150-160; ModBasicMath.F90:360-366) */
  double sl_c1[ORC_NTAB + 1], sl_c2[ORC_NTAB + 1], dl_c1[ORC_NTAB + 1], mask_tab[ORC_NTAB + 1];
  double r_eps;
} orc_params;

/* Cells: everything AddIntOnRbcs / RBC_SingInt / RBC_NearSingInt / PME read. */
typedef struct {
  int ncell, nlat, nlon;
  const double *th, *phi, *w;         /* [nlat], [nlon], [nlat] (w includes 2pi/nlon) */
  const double *x, *a3, *f, *g;       /* SoA(3,Np): rbc%x, rbc%a3, rbc%f, rbc%g (raw densities) */
  const double *detj;                 /* [Np] */
  const double *Acoef, *Bcoef;        /* [ncell]: Acoef(celltype), Bcoef(celltype) */
  const double *area, *meshSize;      /* [ncell] */
  const double *spx, *spa3, *spdetj, *spF, *spG; /* splines, per cell contiguous; nvar 3,3,1,3,3 */
  /* shared polar patch (ModPolarPatch.F90:27-76) */
  double patch_radius;
  int nrad, nazm;
  const double *thG, *phiG, *patch_w; /* [nlon][nlat][nazm][nrad], [nrad] */
} orc_cells;

typedef struct {
  int n;
  const double *x;       /* SoA(3,n) */
  const double *Acoef;   /* [n] */
  const int *indx;       /* SoA(3,n): surfId (1-based cell id, or >ncell for wall, or -1), ilat, ilon */
  const int *active;     /* [n] 0/1 */
} orc_targets;

/* ---- parameters ---- */
void orc_set_ewald_prms(const double Lb[3], double alpha, double eps, int P, int nranks,
                        double *rc, int Nb[3]);
void orc_params_init(orc_params *prm, const double Lb[3], double alpha, double eps, int P,
                     double rc, const int Nb[3]);

/* ---- scalar helpers ---- */
void orc_ewald_coeff_sl_exact(double r, double alpha, double *A, double *B);
void orc_ewald_coeff_dl_exact(double r, double alpha, double *A);
void orc_ewald_coeff_sl(const orc_params *prm, double r, double *A, double *B);
void orc_ewald_coeff_dl(const orc_params *prm, double r, double *A);
double orc_mask_func_exact(double x);
double orc_mask_func(const orc_params *prm, double x);
void orc_bspline_func(double xc, int P, int *imin, double *w);
double orc_dist_on_sphere(double th0, double phi0, double th1, double phi1);
void orc_gauleg(double x1, double x2, int n, double *x, double *w);
void orc_gauleg_sinh(double xmin, double xmax, double a, double b, int n, double *x, double *w);
void orc_polar_patch_build(double th0, double phi0, int nth, const double *thL, int nphi,
                           const double *phiL, double *thG, double *phiG);
void orc_polar_patch_map(double th0, double phi0, double dth, double dphi, double *th, double *phi);
int orc_polar_patch_find_points(double th0, double phi0, double r0, int nth, const double *ths,
                                int nphi, const double *phis, int *ijs);
void orc_rbc_polar_patch_create(const orc_params *prm, int nlat, int nlon, const double *th,
                                const double *phi, double *radius, int *nrad, int *nazm,
                                double *thG, double *phiG, double *w);
void orc_spline_interp(const double *sp, int m, int n, int nvar, double x, double y, double *f);
void orc_spline_find_projection(const double *sp, int m, int n, const double xtar[3], double *th0,
                                double *phi0, double x0[3]);
int orc_quadfit_2d(int npt, const double *xy, const double *f, double c[6]);

/* ---- cell list ---- */
void orc_hash_index(const orc_params *prm, const double x[3], int *i1, int *i2, int *i3);
void orc_hash_build(const orc_params *prm, int n, const double *x, int *hoc, int *next);
/* cell id (0-based, i1 fastest) of every point; bit-exactness target for the GPU cell list */
void orc_cell_ids(const orc_params *prm, int n, const double *x, int *cid);
/* number of sources within rc of every target and an order-independent checksum of their indices */
void orc_neighbor_signature(const orc_params *prm, int ns, const double *xs, int nt,
                            const double *xt, int *count, unsigned long long *sig);

/* ---- real-space operator ---- */
void orc_add_int_on_rbcs(const orc_params *prm, const orc_cells *cells, double c1, double c2,
                         const orc_targets *tl, double *v, int flags);
#define ORC_FLAG_NO_SING 1
#define ORC_FLAG_NO_NEARSING 2
#define ORC_FLAG_NO_LINEAR 4
#define ORC_FLAG_NO_PAIRS 8
void orc_rbc_sing_int(const orc_params *prm, const orc_cells *cells, double c1, double c2, int icell,
                      int ilat0, int ilon0, double dv[3]);
void orc_rbc_nearsing_int(const orc_params *prm, const orc_cells *cells, double c1, double c2,
                          int icell, const double xi[3], const double x0[3], double th0, double phi0,
                          double dv[3]);

/* ---- PME (ModPME.F90 + ModPFFTW.F90), stateful like the reference ---- */
typedef struct orc_pme orc_pme;
orc_pme *orc_pme_init(const orc_params *prm);
void orc_pme_finalize(orc_pme *pme);
/* point sources: x SoA(3,n); f SoA(3,n) or NULL; g,a3 SoA(3,n), Bcoef [n] or NULL */
void orc_pme_distrib_source(orc_pme *pme, double c1, double c2, int n, const double *x,
                            const double *f, const double *g, const double *a3, const double *Bcoef,
                            int accumulate);
void orc_pme_transform(orc_pme *pme);
void orc_pme_add_interp_vel(orc_pme *pme, const orc_targets *tl, double *v);
const double *orc_pme_vv(orc_pme *pme);  /* [3][Nz][Ny][Nx] after transform */
const double *orc_pme_bb(orc_pme *pme);  /* [Nx/2+1][Ny][Nz], z fastest (bb(k,j,i)) */
/* raw FFT access for the tests */
void orc_fft_forward(const orc_params *prm, const double *real_in, double *cplx_out);
void orc_fft_backward(const orc_params *prm, const double *cplx_in, double *real_out);

/* whole cell operator: v += AddIntOnRbcs + PME (ModVelSolver.F90:568-582 pattern) */
void orc_apply_cells(const orc_params *prm, const orc_cells *cells, orc_pme *pme, double c1,
                     double c2, const orc_targets *tl, double *v, int flags);


/* ---- walls (rbc3d_oracle_walls.c): ModIntOnWalls.F90, wall branches of ModSourceList/ModPME ---- */
typedef struct {
  int nwall;
  const int *nvert, *nele; /* [nwall] */
  int id0;                 /* walls(1)%ID: surface id of the first wall (cells are 1..ncell) */
  const double *x;         /* SoA(3,NV): vertices of all walls back to back (= tlist_wall%x) */
  const double *f;         /* SoA(3,NV) tractions wall%f, or NULL */
  const int *e2v;          /* SoA(3,NE): wall%e2v, 1-based vertex numbers local to the wall */
  const double *area, *epsDist; /* [NE] (Wall_ComputeGeometry, ModWall.F90:118-145) */
} orc_walls;
typedef struct orc_wallmat orc_wallmat; /* t_Wall%lhs */

void orc_gq_tri7(double *rs /* [7][2] */, double *w);
void orc_wall_compute_geometry(const orc_walls *W, double *area, double *epsDist);
void orc_wall_centroids(const orc_params *prm, const orc_walls *W, double *xc /* SoA(3,NE) */);
/* x: [l*3+d] = x(l+1,d+1) */
double orc_min_dist_to_tri(const double xTar[3], const double *x, double *s0, double *t0, double *x0);
/* lhs[(l*3+ii)*3+jj] = lhs(l+1,ii+1,jj+1) or NULL */
void orc_tri_int_regular(const orc_params *prm, const double *x, const double *f, const double xtar[3],
                         double rhs[3], double *lhs);
void orc_tri_int_duffy(const orc_params *prm, const double *x, const double *f, const double xtar[3], double s0,
                       double t0, double rhs[3], double *lhs);
orc_wallmat *orc_prepare_sing_int_on_wall(const orc_params *prm, const orc_walls *W, int iwall, const int *active);
void orc_wallmat_free(orc_wallmat *M);
int orc_wallmat_nblk(const orc_wallmat *M);
void orc_wallmat_get(const orc_wallmat *M, int *rowptr, int *col, double *val /* [nblk][3][3] */);
void orc_sing_int_on_wall(const orc_wallmat *M, double c1, const double *f, double *v);
void orc_add_int_on_walls(const orc_params *prm, const orc_walls *W, double c1, const orc_targets *tl, double *v,
                          orc_wallmat *const *mats);
void orc_wall_neighbor_signature(const orc_params *prm, const orc_walls *W, const orc_targets *tl, int self_skip,
                                 int *count, unsigned long long *sig, int *nduffy);
void orc_pme_distrib_walls(orc_pme *pme, const orc_params *prm, double c1, double c2, const orc_walls *W,
                           int accumulate);

/* ---- ModRepulsion.F90: Closest_Neighbor_Cell / _Wall (:480-613), InterCellRepulsion (:270-402) ---- */
void orc_closest_neighbors(const orc_params *prm, const orc_cells *C, const orc_walls *W, int n, const double *x,
                           const int *surfId, double epsDist, double *dist_cell, double *x0_cell, double *dist_wall,
                           double *x0_wall);
int orc_inter_cell_repulsion(const orc_params *prm, const orc_cells *C, const orc_walls *W, const int *active,
                             double epsDist, double *dx, double *dist_min);

int orc_num_threads(void);
void orc_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
