// rbc3d_internal.h -- internal declarations of librbc3d_b200.so (sm_100a, FP64).
// The public surface is include/rbc3d.h.  No CPU fallback exists: every operator entry point launches CUDA
// kernels and fails with RBC3D_ECUDA when no device is present.
#pragma once
#include <cuda_runtime.h>
#include <cufft.h>
#include <stdint.h>

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/rbc3d.h"

#define RBC3D_NTAB 8192
#define RBC3D_NBR_MAX 32

namespace rbc3d {

void set_error(const char *fmt, ...);

#define CUDA_TRY(expr)                                                                          \
  do {                                                                                          \
    cudaError_t e__ = (expr);                                                                   \
    if (e__ != cudaSuccess) {                                                                   \
      rbc3d::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
      return RBC3D_ECUDA;                                                                       \
    }                                                                                           \
  } while (0)
#define CUFFT_TRY(expr)                                                              \
  do {                                                                               \
    cufftResult r__ = (expr);                                                        \
    if (r__ != CUFFT_SUCCESS) {                                                      \
      rbc3d::set_error("%s:%d: %s -> cufft %d", __FILE__, __LINE__, #expr, (int)r__); \
      return RBC3D_ECUDA;                                                            \
    }                                                                                \
  } while (0)
#define RBC_TRY(expr)           \
  do {                          \
    int r__ = (expr);           \
    if (r__ != RBC3D_OK) return r__; \
  } while (0)
#define KERNEL_CHECK() CUDA_TRY(cudaGetLastError())

// Bumped by every device (re)allocation or release: a captured CUDA graph holds raw pointers and is only replayed
// while this has not moved since the capture.
inline long long &alloc_epoch() {
  static long long e = 0;
  return e;
}

template <class T>
struct dbuf {
  T *p = nullptr;
  size_t n = 0;
  int resize(size_t m) {  // grow-only
    if (m <= n && p) return RBC3D_OK;
    alloc_epoch()++;
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
    if (m == 0) return RBC3D_OK;
    cudaError_t e = cudaMalloc((void **)&p, m * sizeof(T));
    if (e != cudaSuccess) {
      set_error("cudaMalloc(%zu bytes) failed: %s", m * sizeof(T), cudaGetErrorString(e));
      return RBC3D_ENOMEM;
    }
    n = m;
    return RBC3D_OK;
  }
  void release() {
    if (p) {
      alloc_epoch()++;
      cudaFree(p);
    }
    p = nullptr;
    n = 0;
  }
};

// Constants every kernel needs (passed by value).
struct Params {
  double Lb[3], iLb[3];
  double alpha, eps, rc;
  double rc2_thr;   // largest double t with sqrt(t) <= rc  (exact replacement of "sqrt(r2) > rc")
  double r_eps;     // 1e-3 * sqrt(alpha/pi)    (ModEwaldFunc.F90:105)
  double tab_scale; // NTAB / rc
  int P;
  int Nb[3];
  int Nc[3];        // cell list blocks (ModHashTable.F90:99-126)
  double iLbNc[3];
  double ih[3];     // Nb / Lb  (ModPME.F90:418)
  // z-slab decomposition (ModConf.F90:412-437); single rank: [0, Nb3)
  int nranks, rank;
};

// Sorted-by-cell view of a point set (cell list = "sort by cell + prefix-scan offsets").
struct CellList {
  int n = 0;            // number of points
  int n_sorted = 0;     // number of points that received a valid cell (active ones)
  dbuf<int> cid;        // [n] cell id (0-based, i1 fastest), -> ncells for inactive points
  dbuf<int> order;      // [n] point indices sorted by cell, ascending index inside a cell
  dbuf<int> start;      // [ncells + 2]
  dbuf<int> keys_tmp, vals_tmp;
  dbuf<char> cub_tmp;
  // PME lists of the P = 8 walk kernels: per SORTED point a 26-double record (B-spline weights wx[8] wy[8] wz[8],
  // mesh cell x in the low half of slot 24, pad) -- geometry-only data, computed once per list
  dbuf<double> w;
  // spreading kernel of this source list, chosen per list by pme_spread_mode: pencil walks (long lists) or 8 x 4 x 4
  // source blocks (short lists, where a pencil is a long serial chain and most of the GPU would idle)
  bool swalk = false;
  int sblk[3] = {4, 4, 4}, nsblk[3] = {0, 0, 0};
};
constexpr int PME_WREC = 26;

struct NearSing {      // geometry-time products of the neighbour scan + Spline_FindProjection
  int n = 0;           // entries (target, other cell within rc)
  int n_active = 0;    // entries that pass the distance check of ModRbcSingInt.F90:122-125
  dbuf<int> cnt, off;  // per sorted target
  dbuf<int> target, cell, pt;  // [n] original target index, source cell, closest mesh point (ilon*nlat+ilat)
  dbuf<double> th0, phi0, dist, x0, a30, xi;  // x0,a30,xi SoA(3,n)
  dbuf<int> flag;      // 1 = active
  dbuf<double> dv;     // SoA(3,n) per-entry corrections of the current application
  dbuf<int> overflow;  // device flag
};

// (target, wall element) pairs within rc of a target list (AddIntOnWalls direct loop), built lazily per geometry
struct WallPairs {
  bool valid = false;
  int npair = 0;
  long long geom_version = -1, tl_version = -1;
  dbuf<int> off, ele, kind, ptarget;  // off[n+1]; per pair: element, 1 = Duffy, target
  dbuf<double> s0, t0, dv;            // closest-point coordinates; dv SoA(3,npair) of the current application
  dbuf<char> tmp;
  void release() {
    off.release(), ele.release(), kind.release(), ptarget.release();
    s0.release(), t0.release(), dv.release(), tmp.release();
    valid = false;
  }
};

struct TargetList {
  int kind = RBC3D_TL_RAW;
  int n = 0;
  bool valid = false;
  dbuf<double> x;       // SoA(3,n)
  dbuf<double> Acoef;   // [n]
  dbuf<int> active;     // [n]
  dbuf<int> surf;       // [n] cell index (0-based) of a cell target, -1 otherwise
  CellList cl;          // sorted by real-space cell
  CellList pl;          // sorted by PME block (interpolation)
  dbuf<int2> tiles;     // (cell, first sorted position) per warp tile
  int ntiles = 0;
  NearSing ns;
  dbuf<double> acc;     // SoA(3,n) un-normalised sums of the current application
  dbuf<double> acc2;    // same for the PME chain when it runs on the second stream (kept apart: deterministic sums)
  dbuf<double> v;       // SoA(3,n) result of rbc3d_apply_resident
  dbuf<double> host_io; // staging for host v
  WallPairs wp;
  long long version = 0; // bumped whenever the list is rebuilt
  // several ranks: velocity-mesh planes this rank's active targets interpolate from, of every rank (lo, count)
  long long halo_version = -1;
  std::vector<int> halo_need;
  // per-geometry (tile, source) pair list of the real-space sum over other surfaces (pairsum.cu)
  bool plist_valid = false;
  int plist_excl = -1, plist_n = 0;
  long long plist_geom = -1;
  dbuf<int> plist_off, plist_src;
};

// walls (walls.cu): t_Wall arrays of all walls back to back + slist_wall + the self-interaction matrices
struct Walls {
  int nwall = 0, NV = 0, NE = 0;
  bool geom_set = false, f_set = false, mat_ok = false;
  long long geom_version = 0, mat_version = -1;
  std::vector<int> h_nvert, h_nele, h_voff, h_eoff;
  dbuf<double> x, f;            // SoA(3,NV)
  dbuf<int> e2v;                // SoA(3,NE), global 0-based vertex numbers
  dbuf<int> ewall, vwall;       // wall index of an element / a vertex
  dbuf<double> area, epsDist;   // [NE]
  dbuf<double> xc, ft;          // SoA(3,NE): centroids (slist_wall%x), PME strengths THRD*sum(fele)*area
  CellList cl, pl;              // centroids by real-space cell / by PME source block
  dbuf<int> src_own, minus1;
  // t_Wall%lhs of every wall as one block-row matrix over the NV vertices (3x3 blocks, row-major)
  int nblk = 0;
  dbuf<int> rowptr, col;
  dbuf<double> val;
};

struct Cells {
  int ncell = 0, nlat = 0, nlon = 0, npc = 0, Np = 0;
  bool mesh_set = false, geom_set = false, f_set = false, g_set = false;
  long long geom_version = 0;  // bumped by every SourceList_UpdateCoord
  std::vector<double> h_th, h_phi, h_w, h_A, h_B, h_area, h_mesh;
  dbuf<double> th, phi, w, A, B, area, meshSize;
  dbuf<double> x, a3, f, g;          // SoA(3,Np), original order (f, g already * detJ*w)
  dbuf<double> spx, spa3, spdetj, spF, spG;  // ABI layout
  // polar patch
  double radius = 0;
  int nrad = 0, nazm = 0;
  dbuf<double> thG, phiG, pw;        // [ilon][ilat][iazm][irad], [nrad]
  dbuf<double> omm;                  // one-minus-mask table [ilat_i][ilat_j][dlon]
  dbuf<int> dlonmax;                 // [ilat_i][ilat_j] largest cyclic |dlon| with mask != 0, -1 if none
  // sorted (by real-space cell) copies used by the pair kernel
  CellList cl;
  dbuf<double> sx, sa3, sf, sgB;     // SoA(3,Np) in sorted order; sgB = g * Bcoef(cell)
  CellList pl;                       // sorted by PME block (spreading)
  dbuf<double> xvint_part;           // partial sums of the linear term
  dbuf<double> sing_xi;              // SoA(3,Np) spline-evaluated target positions (ModRbcSingInt.F90:58)
  // cached double-layer singular path (singular.cu): cell-independent tile tables, per-geometry cache
  bool sg_ok = false, sg_cache_ok = false, spGi_valid = false, spFi_valid = false;
  int sg_ntiles = 0, sg_K = 0, sg_win_max = 0, sg_ntab = 0;
  int sg_npatch_active = 0;          // patch points per target with a non-zero quadrature weight (cached path)
  dbuf<int> sg_tile_tgt, sg_tile_win, sg_idx, sg_cell_active, sg_tile_list, sg_pos;
  int sg_ntl = 0, sg_ntn = 0, sg_ni_max = 0, sg_chunk_stride = 0;
  dbuf<int> sg_rounds, sg_active_list;
  int sg_nactive = 0;
  dbuf<int2> sg_chunk;
  size_t sg_smem = 0;
  dbuf<double> sg_st;                // (s, t) pairs
  dbuf<double> spGi;                 // spline(g detJ) as double2 planes: [cell][6][2 nlat][nlon][2] (phi fastest)
  dbuf<double> spTi;                 // same layout, scratch: spline(x), spline(a3) at geometry time, then spline(f detJ)
  dbuf<double4> sg_cache;            // [slot][row][sorted patch point][half][nlon] x 16 B: (xx.x, xx.y) | (xx.z, w EA (xx.a3))
  // dense same-surface pair kernel (pairself.cu)
  bool ps_ok = false;
  int ps_nwarps = 0;
  dbuf<int> ps_warp_tgt;
  dbuf<unsigned long long> ps_maskbits;
  dbuf<unsigned char> ps_compact;
  dbuf<unsigned char> ps_needmask;   // [patch][patch]: mask table needed for this patch pair
  // geometry cache of the symmetric same-surface double-layer pair sum (pairself.cu): per active cell slot and
  // patch pair the 32-bit set of rotation steps with an in-range pair, per such step 32 coefficients (1 - mask) EA
  bool pc_ok = false;
  bool pc_pending = false;           // geometry changed: the cache is rebuilt by the first double-layer-only pair sum
  int pc_ncached = 0;                // the first pc_ncached cells of sg_active_list are cached
  long long pc_rows = 0;
  dbuf<unsigned> pc_mask;            // [slot][G][G]
  dbuf<int> pc_cnt;                  // [slot * G + I] -> first step of (slot, I) after the scan
  dbuf<double> pc_coef;              // [step][32]
  dbuf<int> src_own;                 // multi-GPU: 1 for the points of the cells this rank spreads
  // density splines built on the device (splinebuild.cu)
  bool sb_ok = false;
  int sb_nlat0 = 0;
  dbuf<double> sb_M0, sb_M1, sb_cs;
  dbuf<double> sb_detj;              // mesh detJ of rbc3d_cells_set_geometry_mesh (scratch)
  dbuf<int> sb_need;                 // several ranks: cells whose density spline this rank reads
};

// device-resident cell velocity solve (solver.cu)
struct Solver {
  bool ok = false;
  int m0 = 0, dofc = 0;
  size_t dof = 0;        // unknowns of all cells
  size_t dof_loc = 0;    // unknowns this rank holds (= dof on one rank)
  int nown = 0, maxown = 0;
  dbuf<int> cells;       // [nown] cells this rank holds the unknowns of
  dbuf<int> own_all;     // several ranks: [rank][maxown] every rank's list (-1 padded)
  dbuf<double> gpack;    // [rank][maxown][3][npc] all-gather buffer of the synthesised density
  long long nmatvec = 0;
  dbuf<double> pb, pbw, cs, dsw, g_raw;
  dbuf<int> ka, kb;
  dbuf<double> V, w, part, h, u, b;
};

// device-resident wall no-slip solve (solver.cu; ModNoSlip.F90:44-149)
struct WallSolver {
  int nindep = 0;
  dbuf<int> indx, last;          // [NV] 0-based independent number of a vertex; [nindep] last vertex with that number
  dbuf<double> V, w, part, h, rhs, x, f0, fw;
  // operator #4 (set traction + walls -> wall vertices) captured once and replayed while nothing it refers to changed
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t gexec = nullptr;
  long long graph_launches = 0;
  long long g_epoch = -1, g_geom = -1, g_mat = -1, g_tl = -1;
  int g_flags[4] = {0, 0, 0, 0};  // NV, nindep, skip_flags, overlap
  Params g_prm{};
};

struct Pme {
  int Nx = 0, Ny = 0, Nz = 0, Nxh = 0;
  size_t G = 0, M = 0;
  dbuf<double> src;                 // [9][Nz][Ny][Nx] spread densities (3 SL + 6 symmetric DL)
  dbuf<cufftDoubleComplex> srcC;    // [9][Nz][Ny][Nxh]
  dbuf<cufftDoubleComplex> vvC;     // [3][Nz][Ny][Nxh]
  dbuf<double> vv;                  // [3][Nz][Ny][Nx]
  dbuf<double> bx, by, bz;          // B-spline modulus factors per axis (ModPME.F90:318-325)
  dbuf<double> str;                 // walk spreading: strengths in sorted order, [point][pass][4]
  cufftHandle planF[3] = {0, 0, 0}; // batch 3, 6, 9
  bool planF_ok[3] = {false, false, false};
  cufftHandle planB = 0;
  bool planB_ok = false;
  dbuf<char> work;
  bool flag_sl = false, flag_dl = false;
  bool distributed = false, transformed = false;
  int nblk[3] = {0, 0, 0};          // PME blocks of PME_BLK^3 mesh cells (interpolation)
  int iblk[3] = {4, 4, 4};          // their edges
  int swalk_mode = -1;              // RBC3D_SPREAD_BLOCKS: 1 = always source blocks, 0 = always pencil walks, unset = by size
  int interp_direct_max = 16384;    // RBC3D_INTERP_DIRECT_MAX: longest target list interpolated one warp per target (k_interp_direct)
  int swalk_min = 1 << 19;          // RBC3D_SPREAD_WALK_MIN: shortest source list that takes the pencil walk (profiles/r02_spread_crossover.jsonl)
  bool walk = false;                // P = 8: register-ring column walks along z (lists keyed z-fastest)
  // slab-decomposed transform (several ranks; ModPFFTW.F90:56-89, 188-316): z-slabs of planes for the 2-D transforms,
  // y-slabs of pencils for the transform in z and the k-space multiplier, all-to-all transposes in between
  bool slab = false;
  int R = 1, rk = 0;
  std::vector<int> zoff, yoff;      // [R + 1] first plane / first y row of every rank
  dbuf<int> d_zoff, d_yoff;
  dbuf<cufftDoubleComplex> sbuf, rbuf, Tz, Vz;
  cufftHandle plan2F = 0, plan2B = 0, plan1[2] = {0, 0};
  bool plan2F_ok = false, plan2B_ok = false, plan1_ok[2] = {false, false};
  int plan1_batch[2] = {0, 0};
  dbuf<int> halo_tmp;               // device scratch of the needed-plane reduction / all-gather
};

}  // namespace rbc3d

struct rbc3d_ctx {
  int device = 0;
  rbc3d::Params prm;
  cudaStream_t stream = nullptr;
  // the PME chain (spread, mesh all-reduce, FFTs, scaling, interpolation) does not depend on the real-space sums:
  // with overlap on it is issued on stream2 while singular / pair kernels run on stream (joined before combine)
  cudaStream_t stream2 = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  int quiet = 0;               // no timing events and no trailing host sync (a stream capture is in progress)
  int resident_collect = 1;    // rbc3d_apply_resident sums the rows over the ranks (0: the sharded solver keeps them local)
  int replicated_density = 0;  // 1: host densities are identical on all ranks -> upload 1/nranks each + all-gather
  int overlap = -1;         // -1: on with several ranks (hides the mesh all-reduce), 0 off, 1 on
  // lookup tables (device): interleaved SL (c1,c2) pairs, DL, mask
  rbc3d::dbuf<double> tab_sl, tab_dl, tab_mask;
  std::vector<double> h_tab_sl1, h_tab_sl2, h_tab_dl, h_tab_mask;
  rbc3d::Cells cells;
  rbc3d::TargetList tl[3];
  rbc3d::Walls walls;
  rbc3d::Pme pme;
  rbc3d::Solver solver;
  rbc3d::WallSolver wsolver;
  int skip_flags = 0;
  int pair_self_mode = 3;   // same-surface pairs: 0 cell list, 1 symmetric patch-pair kernel, 2 dense per-cell kernel,
                            // 3 symmetric kernel streaming a per-geometry coefficient cache (double layer only)
  int sing_cache_mode = 1;  // 0: never cache the singular double-layer integrand, 1: when memory allows
  cudaEvent_t ev[2 * RBC3D_T_COUNT] = {};
  bool ev_used[RBC3D_T_COUNT];
  float ms[RBC3D_T_COUNT];
  long long launches = 0;
  int sm_count = 148;
  void *nccl_comm = nullptr;
};

namespace rbc3d {

// ---- host helpers (host_math.cpp part of capi.cu) ----
void h_bspline_func(double xc, int P, int *imin, double *w);
double h_mask_func_exact(double x);
void h_gauleg(double x1, double x2, int n, double *x, double *w);

// ---- cell list (celllist.cu) ----
int celllist_build_realspace(rbc3d_ctx *c, CellList &cl, int n, const double *x, const int *active);
int celllist_build_pme(rbc3d_ctx *c, CellList &cl, int n, const double *x, const int *active, const int blk[3],
                       bool zfast = false);
int tiles_build(rbc3d_ctx *c, TargetList &t);
int device_exclusive_scan(rbc3d_ctx *c, int *data, int n);  // in place
int celllist_pme_weights(rbc3d_ctx *c, CellList &cl, const double *x);

// ---- real-space operator (pairsum.cu, singular.cu, nearsing.cu) ----
int cells_gather_sorted(rbc3d_ctx *c, bool geom, bool f, bool g);
int pair_sum(rbc3d_ctx *c, TargetList &t, double c1, double c2);
int pairself_mesh_prepare(rbc3d_ctx *c, const std::vector<double> &omm);
int pairself_geometry_prepare(rbc3d_ctx *c);
int pairself_cache_prepare(rbc3d_ctx *c);
bool pairself_available(rbc3d_ctx *c, const TargetList &t);
int pairself_apply(rbc3d_ctx *c, TargetList &t, double c1, double c2);
int cells_active_flags(rbc3d_ctx *c);
int neighbor_signature(rbc3d_ctx *c, TargetList &t, int *count, unsigned long long *sig);
int nearsing_scan(rbc3d_ctx *c, TargetList &t, bool fill);
int nearsing_prepare(rbc3d_ctx *c, TargetList &t);
int nearsing_apply(rbc3d_ctx *c, TargetList &t, double c1, double c2);
int singular_mesh_prepare(rbc3d_ctx *c, const double *thG, const double *phiG, const double *pw);
int singular_prepare(rbc3d_ctx *c);
int singular_density_prepare(rbc3d_ctx *c);
int singular_apply(rbc3d_ctx *c, TargetList &t, double c1, double c2);
int linear_term(rbc3d_ctx *c, TargetList &t, double c2);
int combine(rbc3d_ctx *c, TargetList &t, double *v_dev, bool accumulate, const double *acc2 = nullptr);

// ---- density splines on the device (splinebuild.cu) ----
int spline_builder_prepare(rbc3d_ctx *c, int nlat0);
int spline_build_density(rbc3d_ctx *c, int which);
int spline_build_geometry(rbc3d_ctx *c, const double *detj_dev);

// ---- PME (pme.cu) ----
int pme_block_edge();
int pme_init(rbc3d_ctx *c);
void pme_destroy(rbc3d_ctx *c);
int pme_spread(rbc3d_ctx *c, double c1, double c2, bool use_cells, bool use_walls);
int pme_transform(rbc3d_ctx *c);
int pme_interp(rbc3d_ctx *c, TargetList &t, double *acc = nullptr);  // acc: SoA(3,n) to add into (default t.acc)
int pme_slab_setup(rbc3d_ctx *c);                                    // after the communicator is attached
int pme_source_ownership(rbc3d_ctx *c, int n, const double *x, dbuf<int> &own, const int **flags);

// ---- walls (walls.cu) ----
int walls_set_geometry(rbc3d_ctx *c, int nwall, const int *nvert, const int *nele, const double *x, const int *e2v,
                       const double *area, const double *epsDist);
int walls_set_traction(rbc3d_ctx *c, const double *f, bool from_device = false);
int walls_target_meta(rbc3d_ctx *c, TargetList &t);
int walls_prepare_sing(rbc3d_ctx *c);
int walls_sing_int(rbc3d_ctx *c, double c1, int iwall, double *v_dev);
int walls_add_int(rbc3d_ctx *c, TargetList &t, double c1);
int walls_signature(rbc3d_ctx *c, TargetList &t, int self_skip, int *count, unsigned long long *sig, int *nduffy);
int walls_min_dist_batch(rbc3d_ctx *c, int n, const double *xtar, const double *xtri, double *dist, double *s0,
                         double *t0);
int walls_tri_int_batch(rbc3d_ctx *c, int n, const double *xtri, const double *ftri, const double *xtar,
                        const double *s0, const double *t0, double *rhs, double *lhs);
void walls_release(rbc3d_ctx *c);

// ---- closest-neighbour queries of ModRepulsion on the cell lists (nearsing.cu, walls.cu); device pointers ----
int closest_cells(rbc3d_ctx *c, int n, const double *qx, const int *surf, double epsDist, double *dist, double *x0);
int closest_walls(rbc3d_ctx *c, int n, const double *qx, const int *surf, double *dist, double *x0);

// ---- cell velocity solve on the device (solver.cu) ----
int solver_setup(rbc3d_ctx *c, int nlat0, const double *detj_host);
int solver_matmult(rbc3d_ctx *c, const double *u_dev, double *b_dev);
int solver_gmres(rbc3d_ctx *c, const double *b_dev, double *x_dev, double rtol, int restart, int maxit, int *niter,
                 double *history);
void solver_release(rbc3d_ctx *c);
int wall_noslip_solve(rbc3d_ctx *c, const int *indx_host, int nindep, const double vbkg[3], int use_cells, double rtol, int maxit,
                      double *f_host, int *niter, double *history, double *slip_host);

// ---- multi-GPU (comm.cu) ----
int comm_allreduce_sum(rbc3d_ctx *c, double *buf, size_t n);
int comm_allgather_inplace(rbc3d_ctx *c, double *buf, size_t count);
int comm_allgather_ints(rbc3d_ctx *c, const int *send, int *recv, size_t count);  // device buffers, count per rank
int comm_group_begin();
int comm_group_end();
int comm_send(rbc3d_ctx *c, const void *buf, size_t bytes, int peer);  // bytes: a multiple of 8
int comm_recv(rbc3d_ctx *c, void *buf, size_t bytes, int peer);
void comm_destroy(rbc3d_ctx *c);

// timing helpers
void pme_spread_mode(rbc3d_ctx *c, CellList &L, int n);  // before celllist_build_pme of a source list
void t_begin(rbc3d_ctx *c, int stage);
void t_end(rbc3d_ctx *c, int stage);

}  // namespace rbc3d
