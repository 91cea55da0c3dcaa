// capi.cu -- the C ABI of include/rbc3d.h: context, host-side parameter arithmetic, host<->device mirrors and
// the operator call protocol of the reference (ModPME / ModIntOnRbcs / ModEwaldFunc entry points).
#include <cfloat>
#include <cmath>
#include <cstdarg>

#include "device_math.cuh"
#include "rbc3d_internal.h"

namespace rbc3d {

static thread_local char g_err[1024] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}

static const double H_PI = 3.14159265358979323846;  // ModDataTypes.F90:19
static const double H_TWO_PI = 2 * 3.14159265358979323846;

// ---- host restatements of the init-time helpers (identical arithmetic to the reference, libm on the host) ----
void h_bspline_func(double xc, int P, int *imin, double *w) {  // ModBasicMath.F90:392-417
  double u[64];
  *imin = (int)floor(xc) - (P - 1);
  u[0] = *imin - (xc - P);
  for (int j = 1; j < P; j++) u[j] = u[j - 1] + 1;
  w[0] = 1.;
  for (int j = 1; j < P; j++) w[j] = 0.;
  for (int pp = 2; pp <= P; pp++) {
    for (int j = pp; j >= 2; j--)
      w[j - 1] = u[j - 1] / (pp - 1.) * w[j - 1] + (pp - u[j - 1]) / (pp - 1.) * w[j - 2];
    w[0] = u[0] / (pp - 1.) * w[0];
  }
}

double h_mask_func_exact(double x) {  // ModBasicMath.F90:332-348
  double t = fabs(x);
  if (t < 0.01) return 1.;
  if (t > 0.99) return 0.;
  return exp(2 * (exp(-1. / t)) / (t - 1));
}

static double h_mask_func(const std::vector<double> &tab, double x) {  // ModBasicMath.F90:351-379
  const int N = RBC3D_NTAB;
  double s = fabs(x) * N;
  int i = (int)floor(s);
  if (i >= N) return 0.;
  return tab[i] * (i + 1 - s) + tab[i + 1] * (s - i);
}

void h_gauleg(double x1, double x2, int n, double *x, double *w) {  // ModQuadRule.F90:135-206
  int m = (n + 1) / 2;
  double xm = 0.5 * (x2 + x1), xl = 0.5 * (x2 - x1);
  for (int j = 1; j <= m; j++) {
    double z = cos(H_PI * (j - 0.25) / (n + 0.5)), pp = 0, z1;
    for (int its = 1; its <= 10; its++) {
      double p1 = 1.0, p2 = 0.0, p3;
      for (int k = 1; k <= n; k++) {
        p3 = p2;
        p2 = p1;
        p1 = ((2.0 * k - 1.0) * z * p2 - (k - 1.0) * p3) / k;
      }
      pp = n * (z * p1 - p2) / (z * z - 1.0);
      z1 = z;
      z = z1 - p1 / pp;
      if (!(fabs(z - z1) > 3.e-14)) break;
    }
    x[j - 1] = xm - xl * z;
    x[n - j] = xm + xl * z;
    w[j - 1] = 2.0 * xl / ((1.0 - z * z) * pp * pp);
    w[n - j] = w[j - 1];
  }
}

static void h_polar_patch_build(double th0, double phi0, int nth, const double *thL, int nphi, const double *phiL,
                                double *thG, double *phiG) {  // ModPolarPatch.F90:99-148
  double sin_th0 = sin(th0), cos_th0 = cos(th0);
  for (int i = 0; i < nth; i++) {
    double st = sin(thL[i]), ct = cos(thL[i]);
    for (int j = 0; j < nphi; j++) {
      double sp = sin(phiL[j]), cp = cos(phiL[j]);
      double x0 = st * cp, x1 = st * sp, x2 = ct;
      double y0 = cos_th0 * x0 + 0. * x1 + sin_th0 * x2;
      double y1 = 0. * x0 + 1. * x1 + 0. * x2;
      double y2 = -sin_th0 * x0 + 0. * x1 + cos_th0 * x2;
      y2 = fmax(-1.0, fmin(1.0, y2));
      thG[i + nth * j] = acos(y2);
      double ph = atan2(y1, y0) + phi0;
      phiG[i + nth * j] = ph - floor(ph * (1. / H_TWO_PI)) * H_TWO_PI;
    }
  }
}

static double h_dist_on_sphere(double th0, double phi0, double th1, double phi1) {  // ModPolarPatch.F90:249-257
  double d = cos(th0 - th1) - sin(th0) * sin(th1) * (1. - cos(phi0 - phi1));
  d = fmin(1., fmax(-1., d));
  return acos(d);
}

template <class T>
static int upload(dbuf<T> &d, const T *h, size_t n, cudaStream_t s) {
  RBC_TRY(d.resize(n > 0 ? n : 1));
  if (n) CUDA_TRY(cudaMemcpyAsync(d.p, h, n * sizeof(T), cudaMemcpyHostToDevice, s));
  return RBC3D_OK;
}

__global__ void k_fill_int(int n, int *p, int v) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
__global__ void k_fill_range(int n, int *p, int lo, int hi) {  // p[i] = lo <= i < hi
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = (i >= lo && i < hi) ? 1 : 0;
}
__global__ void k_cell_target_meta(int n, int npc, const double *__restrict__ Acell, int *__restrict__ surf,
                                   double *__restrict__ A) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int c = i / npc;
  surf[i] = c;
  A[i] = Acell[c];
}
__global__ void k_raw_target_meta(int n, int *__restrict__ surf, double *__restrict__ A) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  surf[i] = -1;
  A[i] = 2.0;  // TargetList_CreateFromRaw, ModTargetList.F90:164-165
}

static int target_list_finish(rbc3d_ctx *c, TargetList &t) {
  // cell lists (real-space cells, PME blocks) of the ACTIVE targets, warp tiles, near-singular geometry
  RBC_TRY(celllist_build_realspace(c, t.cl, t.n, t.x.p, t.active.p));
  RBC_TRY(celllist_build_pme(c, t.pl, t.n, t.x.p, t.active.p, c->pme.iblk, c->pme.walk));
  if (c->pme.walk) RBC_TRY(celllist_pme_weights(c, t.pl, t.x.p));
  RBC_TRY(tiles_build(c, t));
  RBC_TRY(t.acc.resize(3 * (size_t)(t.n > 0 ? t.n : 1)));
  RBC_TRY(t.v.resize(3 * (size_t)(t.n > 0 ? t.n : 1)));
  if (c->cells.geom_set) RBC_TRY(nearsing_prepare(c, t));
  t.valid = true;
  t.version++;
  t.wp.valid = false;
  t.plist_valid = false;
  return RBC3D_OK;
}

}  // namespace rbc3d

using namespace rbc3d;

extern "C" {

const char *rbc3d_last_error(void) { return g_err; }
int rbc3d_version(void) { return 100; }

// ModConf.F90:348-408
int rbc3d_set_ewald_prms(const double Lb[3], double alpha, double eps, int P, int nranks, double *rc, int Nb[3]) {
  if (!Lb || !rc || !Nb || nranks < 1 || alpha <= 0 || eps <= 0) return RBC3D_EINVAL;
  double s = 1.;
  for (int iter = 1; iter <= 10; iter++) {
    s = 0.75 * sqrt(H_PI) * eps / (s * s * s + 1.5 * s + 0.75 / s);
    s = sqrt(-log(s));
  }
  double r = sqrt(alpha / H_PI) * s;
  double m = fmin(Lb[0] / 3.001, fmin(Lb[1] / 3.001, Lb[2] / 3.001));
  *rc = fmin(m, r);
  for (int d = 0; d < 3; d++) Nb[d] = 2 * (int)ceil(sqrt(-log(eps) / (H_PI * alpha)) * Lb[d]);
  if (Nb[2] < nranks * P) Nb[2] = nranks * P;
  Nb[2] = (int)ceil((double)Nb[2] / nranks) * nranks;
  return RBC3D_OK;
}

int rbc3d_ewald_coeff_sl_exact(double r, double alpha, double *A, double *B) {  // ModEwaldFunc.F90:25-52
  if (!A || !B) return RBC3D_EINVAL;
  if (alpha <= 0) {
    *A = 1 / (r * r * r);
    *B = 1 / r;
    return RBC3D_OK;
  }
  double r_t = sqrt(H_PI / alpha) * r;
  if (r_t < 1.e-3) {
    *A = 0.;
    *B = 0.;
  } else {
    double c1 = erfc(r_t), c2 = 2. / sqrt(alpha) * exp(-(r_t * r_t));
    double ir = 1. / r, ir2 = ir * ir;
    *A = c1 * ir * ir2 + c2 * ir2;
    *B = c1 * ir - c2;
  }
  return RBC3D_OK;
}

int rbc3d_ewald_coeff_dl_exact(double r, double alpha, double *A) {  // ModEwaldFunc.F90:59-79
  if (!A) return RBC3D_EINVAL;
  if (alpha <= 0) {
    *A = -6 / (r * r * r * r * r);
    return RBC3D_OK;
  }
  double r_t = sqrt(H_PI / alpha) * r;
  if (r_t < 1.e-3) {
    *A = 0.;
  } else {
    double a = exp(-(r_t * r_t)) * (1.5 * r_t + r_t * r_t * r_t) + 0.75 * sqrt(H_PI) * erfc(r_t);
    a = -8 / sqrt(H_PI) * a;
    *A = a / (r * r * r * r * r);
  }
  return RBC3D_OK;
}

int rbc3d_ewald_coeff_sl(const rbc3d_ctx *c, double r, double *A, double *B) {  // ModEwaldFunc.F90:86-131
  if (!c || !A || !B) return RBC3D_EINVAL;
  const int N = RBC3D_NTAB;
  *A = 0.;
  *B = 0.;
  if (r < c->prm.r_eps) return RBC3D_OK;
  double s = N * r / c->prm.rc;
  int i = (int)floor(s);
  if (i >= N) return RBC3D_OK;
  double c1 = c->h_tab_sl1[i] * (i + 1 - s) + c->h_tab_sl1[i + 1] * (s - i);
  double c2 = c->h_tab_sl2[i] * (i + 1 - s) + c->h_tab_sl2[i + 1] * (s - i);
  double ir = 1. / r, ir2 = ir * ir;
  *A = c1 * ir * ir2 + c2 * ir2;
  *B = c1 * ir - c2;
  return RBC3D_OK;
}

int rbc3d_ewald_coeff_dl(const rbc3d_ctx *c, double r, double *A) {  // ModEwaldFunc.F90:141-178
  if (!c || !A) return RBC3D_EINVAL;
  const int N = RBC3D_NTAB;
  *A = 0.;
  if (r < c->prm.r_eps) return RBC3D_OK;
  double s = N * r / c->prm.rc;
  int i = (int)floor(s);
  if (i >= N) return RBC3D_OK;
  double c1 = c->h_tab_dl[i] * (i + 1 - s) + c->h_tab_dl[i + 1] * (s - i);
  double r2 = r * r;
  *A = c1 / (r2 * r2 * r);
  return RBC3D_OK;
}

static int ctx_build(rbc3d_ctx *c, const double Lb[3], double alpha, double eps, int P, double rc, const int Nb[3]);

int rbc3d_ctx_create(rbc3d_ctx **out, const double Lb[3], double alpha, double eps, int P, double rc,
                     const int Nb[3], int device) {
  if (!out || !Lb || !Nb || rc <= 0 || alpha <= 0) return RBC3D_EINVAL;
  *out = nullptr;
  int ndev = 0;
  CUDA_TRY(cudaGetDeviceCount(&ndev));
  if (ndev == 0) {
    set_error("no CUDA device: librbc3d_b200 has no CPU path");
    return RBC3D_ECUDA;
  }
  if (device < 0) CUDA_TRY(cudaGetDevice(&device));
  CUDA_TRY(cudaSetDevice(device));
  rbc3d_ctx *c = new rbc3d_ctx();
  c->device = device;
  const int rc_build = ctx_build(c, Lb, alpha, eps, P, rc, Nb);
  if (rc_build != RBC3D_OK) {  // a half-built context: release what exists (every handle starts out null)
    rbc3d_ctx_destroy(c);
    return rc_build;
  }
  *out = c;
  return RBC3D_OK;
}

static int ctx_build(rbc3d_ctx *c, const double Lb[3], double alpha, double eps, int P, double rc, const int Nb[3]) {
  const int device = c->device;
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  c->sm_count = prop.multiProcessorCount;
  CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CUDA_TRY(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
  CUDA_TRY(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
  if (const char *e = getenv("RBC3D_OVERLAP")) c->overlap = atoi(e);
  for (int i = 0; i < 2 * RBC3D_T_COUNT; i++) CUDA_TRY(cudaEventCreate(&c->ev[i]));
  for (int i = 0; i < RBC3D_T_COUNT; i++) {
    c->ev_used[i] = false;
    c->ms[i] = 0;
  }
  Params &p = c->prm;
  for (int d = 0; d < 3; d++) {
    p.Lb[d] = Lb[d];
    p.iLb[d] = 1. / Lb[d];
    p.Nb[d] = Nb[d];
    p.ih[d] = Nb[d] / Lb[d];  // ih = Nb/Lb, ModPME.F90:418
  }
  p.alpha = alpha;
  p.eps = eps;
  p.rc = rc;
  p.P = P;
  p.nranks = 1;
  p.rank = 0;
  // ModHashTable.F90:99-126 (single node)
  for (int d = 0; d < 3; d++) {
    int nc = (int)floor(Lb[d] / rc);
    p.Nc[d] = nc < 3 ? 3 : nc;
  }
  p.iLbNc[0] = p.iLb[0] * p.Nc[0];
  p.iLbNc[1] = p.iLb[1] * p.Nc[1];
  p.iLbNc[2] = p.Nc[2] / (Lb[2] - 0.);
  // exact threshold replacing "sqrt(r2) > rc"
  double t = rc * rc;
  while (sqrt(t) > rc) t = nextafter(t, 0.0);
  while (sqrt(nextafter(t, INFINITY)) <= rc) t = nextafter(t, INFINITY);
  p.rc2_thr = t;
  p.r_eps = 1.e-3 * sqrt(alpha / H_PI);
  p.tab_scale = RBC3D_NTAB / rc;
  // lookup tables, ModEwaldFunc.F90:96-106,150-160; ModBasicMath.F90:360-366
  const int N = RBC3D_NTAB;
  c->h_tab_sl1.resize(N + 1);
  c->h_tab_sl2.resize(N + 1);
  c->h_tab_dl.resize(N + 1);
  c->h_tab_mask.resize(N + 1);
  std::vector<double> sl(2 * (N + 1) + 2, 0.0);
  for (int i = 0; i <= N; i++) {
    double r_t = sqrt(H_PI / alpha) * (i * rc / N);
    c->h_tab_sl1[i] = erfc(r_t);
    c->h_tab_sl2[i] = 2 / sqrt(alpha) * exp(-(r_t * r_t));
    c->h_tab_dl[i] =
        -8 / sqrt(H_PI) * (exp(-(r_t * r_t)) * (1.5 * r_t + r_t * r_t * r_t) + 0.75 * sqrt(H_PI) * erfc(r_t));
    c->h_tab_mask[i] = h_mask_func_exact((double)i / N);
    sl[2 * i] = c->h_tab_sl1[i];
    sl[2 * i + 1] = c->h_tab_sl2[i];
  }
  int rcode;
  if ((rcode = upload(c->tab_sl, sl.data(), sl.size(), c->stream)) != RBC3D_OK) return rcode;
  if ((rcode = upload(c->tab_dl, c->h_tab_dl.data(), (size_t)N + 1, c->stream)) != RBC3D_OK) return rcode;
  if ((rcode = upload(c->tab_mask, c->h_tab_mask.data(), (size_t)N + 1, c->stream)) != RBC3D_OK) return rcode;
  if ((rcode = pme_init(c)) != RBC3D_OK) return rcode;
  if ((rcode = pme_slab_setup(c)) != RBC3D_OK) return rcode;
  RBC_TRY(c->cells.xvint_part.resize(3 * 296 + 8));
  CUDA_TRY(cudaMemsetAsync(c->cells.xvint_part.p, 0, sizeof(double) * (3 * 296 + 8), c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  for (int k = 0; k < 3; k++) c->tl[k].kind = k;
  return RBC3D_OK;
}

int rbc3d_ctx_destroy(rbc3d_ctx *c) {
  if (!c) return RBC3D_OK;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  pme_destroy(c);
  comm_destroy(c);
  walls_release(c);
  for (int i = 0; i < 2 * RBC3D_T_COUNT; i++)
    if (c->ev[i]) cudaEventDestroy(c->ev[i]);
  if (c->stream) cudaStreamDestroy(c->stream);
  if (c->stream2) cudaStreamDestroy(c->stream2);
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  // device buffers are released with the context's allocations
  auto rel_cl = [](CellList &l) {
    l.cid.release();
    l.order.release();
    l.start.release();
    l.keys_tmp.release();
    l.vals_tmp.release();
    l.cub_tmp.release();
    l.w.release();
  };
  Cells &C = c->cells;
  for (dbuf<double> *b : {&C.th, &C.phi, &C.w, &C.A, &C.B, &C.area, &C.meshSize, &C.x, &C.a3, &C.f, &C.g, &C.spx,
                          &C.spa3, &C.spdetj, &C.spF, &C.spG, &C.thG, &C.phiG, &C.pw, &C.omm, &C.sx, &C.sa3, &C.sf,
                          &C.sgB, &C.xvint_part, &C.sing_xi, &c->tab_sl, &c->tab_dl, &c->tab_mask})
    b->release();
  C.dlonmax.release();
  C.sg_tile_tgt.release();
  C.sg_tile_win.release();
  C.sg_idx.release();
  C.sg_tile_list.release();
  C.sg_rounds.release();
  C.sb_M0.release();
  C.sb_M1.release();
  C.sb_cs.release();
  C.sg_active_list.release();
  C.sg_chunk.release();
  C.sg_pos.release();
  C.sg_cell_active.release();
  C.sg_st.release();
  C.spGi.release();
  C.spTi.release();
  C.sg_cache.release();
  C.ps_warp_tgt.release();
  C.ps_maskbits.release();
  C.ps_compact.release();
  C.ps_needmask.release();
  solver_release(c);
  C.sb_need.release();
  C.sb_detj.release();
  C.pc_mask.release();
  C.pc_cnt.release();
  C.pc_coef.release();
  C.src_own.release();
  rel_cl(C.cl);
  rel_cl(C.pl);
  for (int k = 0; k < 3; k++) {
    TargetList &t = c->tl[k];
    for (dbuf<double> *b : {&t.x, &t.Acoef, &t.acc, &t.acc2, &t.v, &t.host_io, &t.ns.th0, &t.ns.phi0, &t.ns.dist, &t.ns.x0,
                            &t.ns.a30, &t.ns.xi, &t.ns.dv})
      b->release();
    for (dbuf<int> *b : {&t.active, &t.surf, &t.ns.cnt, &t.ns.off, &t.ns.target, &t.ns.cell, &t.ns.pt, &t.ns.flag,
                         &t.ns.overflow})
      b->release();
    t.tiles.release();
    t.plist_off.release();
    t.plist_src.release();
    rel_cl(t.cl);
    rel_cl(t.pl);
  }
  delete c;
  return RBC3D_OK;
}

int rbc3d_host_register(void *ptr, size_t bytes) {
  CUDA_TRY(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
  return RBC3D_OK;
}
int rbc3d_host_unregister(void *ptr) {
  CUDA_TRY(cudaHostUnregister(ptr));
  return RBC3D_OK;
}

int rbc3d_set_sing_cache(rbc3d_ctx *c, int mode) {
  if (!c) return RBC3D_EINVAL;
  c->sing_cache_mode = mode;
  if (mode == 0) {
    c->cells.sg_cache_ok = false;
    c->cells.sg_cache.release();
  }
  return RBC3D_OK;
}

int rbc3d_set_pair_self(rbc3d_ctx *c, int mode) {
  if (!c) return RBC3D_EINVAL;
  c->pair_self_mode = mode;
  return RBC3D_OK;
}

int rbc3d_sing_cache_info(rbc3d_ctx *c, int32_t *cached, int32_t *points_per_target) {
  if (!c) return RBC3D_EINVAL;
  if (cached) *cached = c->cells.sg_cache_ok ? 1 : 0;
  if (points_per_target) *points_per_target = c->cells.sg_npatch_active;
  return RBC3D_OK;
}

int rbc3d_set_replicated_density(rbc3d_ctx *c, int on) {
  if (!c) return RBC3D_EINVAL;
  c->replicated_density = on ? 1 : 0;
  return RBC3D_OK;
}

int rbc3d_pair_cache_info(rbc3d_ctx *c, int32_t *cells_cached, int64_t *rows) {
  if (!c) return RBC3D_EINVAL;
  if (cells_cached) *cells_cached = c->cells.pc_ok ? c->cells.pc_ncached : 0;
  if (rows) *rows = c->cells.pc_ok ? c->cells.pc_rows : 0;
  return RBC3D_OK;
}

int rbc3d_set_overlap(rbc3d_ctx *c, int mode) {
  if (!c || mode < -1 || mode > 1) return RBC3D_EINVAL;
  c->overlap = mode;
  return RBC3D_OK;
}

int rbc3d_set_skip_flags(rbc3d_ctx *c, int flags) {
  if (!c) return RBC3D_EINVAL;
  c->skip_flags = flags;
  return RBC3D_OK;
}

// RBC_Create mesh part (ModRbc.F90:93-95) + RbcPolarPatch_Create (ModPolarPatch.F90:27-76)
int rbc3d_cells_set_mesh(rbc3d_ctx *c, int ncell, int nlat, int nlon, const double *th, const double *phi,
                         const double *w) {
  if (!c || ncell < 0 || nlat < 2 || nlon < 4 || (nlon & 1) || !th || !phi || !w) return RBC3D_EINVAL;
  CUDA_TRY(cudaSetDevice(c->device));
  Cells &C = c->cells;
  C.ncell = ncell;
  C.nlat = nlat;
  C.nlon = nlon;
  C.npc = nlat * nlon;
  C.Np = ncell * C.npc;
  C.h_th.assign(th, th + nlat);
  C.h_phi.assign(phi, phi + nlon);
  C.h_w.assign(w, w + nlat);
  RBC_TRY(upload(C.th, th, nlat, c->stream));
  RBC_TRY(upload(C.phi, phi, nlon, c->stream));
  RBC_TRY(upload(C.w, w, nlat, c->stream));
  // polar patch
  const double radius = H_PI / sqrt((double)nlat);
  const double h = H_PI / nlat;
  const int nrad = 2 * (int)round(radius / h);
  const int nazm = 2 * nrad;
  C.radius = radius;
  C.nrad = nrad;
  C.nazm = nazm;
  std::vector<double> thL(nrad), pw(nrad), phiL(nazm);
  h_gauleg(0., radius, nrad, thL.data(), pw.data());
  for (int ir = 0; ir < nrad; ir++) {
    pw[ir] = pw[ir] * sin(thL[ir]) * (H_TWO_PI / nazm);
    pw[ir] = pw[ir] * h_mask_func(c->h_tab_mask, thL[ir] / radius);
  }
  for (int ia = 0; ia < nazm; ia++) phiL[ia] = ia * H_TWO_PI / nazm;
  const size_t np = (size_t)nrad * nazm;
  std::vector<double> thG(np * C.npc), phiG(np * C.npc);
  for (int ilon = 0; ilon < nlon; ilon++)
    for (int ilat = 0; ilat < nlat; ilat++) {
      size_t off = ((size_t)ilon * nlat + ilat) * np;
      h_polar_patch_build(th[ilat], phi[ilon], nrad, thL.data(), nazm, phiL.data(), &thG[off], &phiG[off]);
    }
  RBC_TRY(upload(C.thG, thG.data(), thG.size(), c->stream));
  RBC_TRY(upload(C.phiG, phiG.data(), phiG.size(), c->stream));
  RBC_TRY(upload(C.pw, pw.data(), pw.size(), c->stream));
  // one-minus-mask table of the pair loop: (1 - MaskFunc(DistOnSphere/radius)), ModIntOnRbcs.F90:82-84,94
  const int nlonh = nlon / 2 + 1;
  std::vector<double> omm((size_t)nlat * nlat * nlonh);
  std::vector<int> dmax((size_t)nlat * nlat);
  for (int i = 0; i < nlat; i++)
    for (int j = 0; j < nlat; j++) {
      int mx = -1;
      for (int dl = 0; dl < nlonh; dl++) {
        double d = h_dist_on_sphere(th[i], phi[dl], th[j], phi[0]);
        double mk = h_mask_func(c->h_tab_mask, d / radius);
        omm[((size_t)i * nlat + j) * nlonh + dl] = 1. - mk;
        if (mk != 0.) mx = dl;
      }
      dmax[(size_t)i * nlat + j] = mx;
    }
  RBC_TRY(upload(C.omm, omm.data(), omm.size(), c->stream));
  RBC_TRY(upload(C.dlonmax, dmax.data(), dmax.size(), c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  RBC_TRY(singular_mesh_prepare(c, thG.data(), phiG.data(), pw.data()));
  RBC_TRY(pairself_mesh_prepare(c, omm));
  C.mesh_set = true;
  C.sb_ok = false;
  C.geom_set = C.f_set = C.g_set = false;
  // everything sized by the old mesh goes: the cell target list, the near-singular entries of the other lists
  // (rebuilt by the next rbc3d_cells_set_geometry), the coefficient caches and the solver tables
  c->tl[RBC3D_TL_CELLS].valid = false;
  c->tl[RBC3D_TL_CELLS].n = 0;
  for (int k = 0; k < 3; k++) c->tl[k].plist_valid = false;
  C.pc_ok = false;
  C.pc_pending = false;
  C.pc_ncached = 0;
  C.sg_cache_ok = false;
  c->solver.ok = false;
  C.geom_version++;
  c->launches = 0;
  return RBC3D_OK;
}

static int set_geometry_common(rbc3d_ctx *c, const double *x, const double *a3, const double *Acoef_cell,
                               const double *Bcoef_cell, const double *area, const double *meshSize,
                               const double *spx, const double *spa3, const double *spdetj, const double *detj,
                               const int32_t *active);

int rbc3d_cells_set_geometry(rbc3d_ctx *c, const double *x, const double *a3, const double *Acoef_cell,
                             const double *Bcoef_cell, const double *area, const double *meshSize,
                             const double *spx, const double *spa3, const double *spdetj, const int32_t *active) {
  if (!c || !c->cells.mesh_set) return RBC3D_ESTATE;
  if (!x || !a3 || !Acoef_cell || !Bcoef_cell || !area || !meshSize || !spx || !spa3 || !spdetj) return RBC3D_EINVAL;
  return set_geometry_common(c, x, a3, Acoef_cell, Bcoef_cell, area, meshSize, spx, spa3, spdetj, nullptr, active);
}

// Same with Rbc_BuildSurfaceSource(xFlag) on the device: the caller passes the mesh field detJ instead of the three
// geometry splines (5.3 GB -> 0.6 GB of host->device traffic per time step at 4096 cells)
int rbc3d_cells_set_geometry_mesh(rbc3d_ctx *c, const double *x, const double *a3, const double *detj,
                                  const double *Acoef_cell, const double *Bcoef_cell, const double *area,
                                  const double *meshSize, const int32_t *active) {
  if (!c || !c->cells.mesh_set) return RBC3D_ESTATE;
  if (!x || !a3 || !detj || !Acoef_cell || !Bcoef_cell || !area || !meshSize) return RBC3D_EINVAL;
  if (!c->cells.sb_ok) {
    set_error("rbc3d_cells_set_geometry_mesh needs rbc3d_cells_enable_device_splines first");
    return RBC3D_ESTATE;
  }
  return set_geometry_common(c, x, a3, Acoef_cell, Bcoef_cell, area, meshSize, nullptr, nullptr, nullptr, detj, active);
}

static int set_geometry_common(rbc3d_ctx *c, const double *x, const double *a3, const double *Acoef_cell,
                               const double *Bcoef_cell, const double *area, const double *meshSize,
                               const double *spx, const double *spa3, const double *spdetj, const double *detj,
                               const int32_t *active) {
  CUDA_TRY(cudaSetDevice(c->device));
  Cells &C = c->cells;
  c->solver.ok = false;  // detJ * w of the device solver belongs to the previous geometry: rbc3d_solver_setup again
  const size_t Np = C.Np, nc = C.ncell;
  const size_t sp1 = (size_t)4 * 2 * C.nlat * C.nlon;
  RBC_TRY(upload(C.x, x, 3 * Np, c->stream));
  RBC_TRY(upload(C.a3, a3, 3 * Np, c->stream));
  RBC_TRY(upload(C.A, Acoef_cell, nc, c->stream));
  RBC_TRY(upload(C.B, Bcoef_cell, nc, c->stream));
  RBC_TRY(upload(C.area, area, nc, c->stream));
  RBC_TRY(upload(C.meshSize, meshSize, nc, c->stream));
  if (detj) {
    RBC_TRY(upload(C.sb_detj, detj, Np, c->stream));
    RBC_TRY(spline_build_geometry(c, C.sb_detj.p));
  } else {
    RBC_TRY(upload(C.spx, spx, 3 * sp1 * nc, c->stream));
    RBC_TRY(upload(C.spa3, spa3, 3 * sp1 * nc, c->stream));
    RBC_TRY(upload(C.spdetj, spdetj, sp1 * nc, c->stream));
  }
  // source cell lists: real-space cells (HashTable_Build) and PME blocks
  RBC_TRY(celllist_build_realspace(c, C.cl, (int)Np, C.x.p, nullptr));
  C.geom_version++;
  // PME sources: with several ranks every rank spreads the points whose B-spline support touches its z-slab of mesh
  // planes (pme.cu; ModPME.F90:428-429) -- or, with the slab decomposition switched off, a contiguous block of cells
  const int *src_own = nullptr;
  if (c->prm.nranks > 1 && Np > 0) {
    RBC_TRY(pme_source_ownership(c, (int)Np, C.x.p, C.src_own, &src_own));
    if (!src_own) {
      const int c_lo = (int)((long long)C.ncell * c->prm.rank / c->prm.nranks);
      const int c_hi = (int)((long long)C.ncell * (c->prm.rank + 1) / c->prm.nranks);
      k_fill_range<<<(int)((Np + 255) / 256), 256, 0, c->stream>>>((int)Np, C.src_own.p, c_lo * C.npc, c_hi * C.npc);
      KERNEL_CHECK();
      src_own = C.src_own.p;
    }
  }
  pme_spread_mode(c, C.pl, (int)Np);
  RBC_TRY(celllist_build_pme(c, C.pl, (int)Np, C.x.p, src_own, C.pl.sblk, C.pl.swalk));
  if (C.pl.swalk) RBC_TRY(celllist_pme_weights(c, C.pl, C.x.p));
  C.geom_set = true;
  RBC_TRY(cells_gather_sorted(c, true, false, false));
  // tlist_rbc: TargetList_Update, ModTargetList.F90:95-135
  TargetList &t = c->tl[RBC3D_TL_CELLS];
  t.kind = RBC3D_TL_CELLS;
  t.n = (int)Np;
  const size_t n1 = Np > 0 ? Np : 1;
  RBC_TRY(t.x.resize(3 * n1));
  RBC_TRY(t.Acoef.resize(n1));
  RBC_TRY(t.surf.resize(n1));
  RBC_TRY(t.active.resize(n1));
  if (Np > 0) {
    CUDA_TRY(cudaMemcpyAsync(t.x.p, C.x.p, sizeof(double) * 3 * Np, cudaMemcpyDeviceToDevice, c->stream));
    k_cell_target_meta<<<(int)((Np + 255) / 256), 256, 0, c->stream>>>((int)Np, C.npc, C.A.p, t.surf.p, t.Acoef.p);
    if (active)
      CUDA_TRY(cudaMemcpyAsync(t.active.p, active, sizeof(int) * Np, cudaMemcpyHostToDevice, c->stream));
    else
      k_fill_int<<<(int)((Np + 255) / 256), 256, 0, c->stream>>>((int)Np, t.active.p, 1);
    KERNEL_CHECK();
  }
  RBC_TRY(target_list_finish(c, t));
  RBC_TRY(cells_active_flags(c));
  RBC_TRY(pairself_geometry_prepare(c));
  RBC_TRY(singular_prepare(c));
  // the coefficient cache of the double-layer pair sum is only read by a c1 = 0 matvec (the GMRES of lambda != 1):
  // it is built by the first such call after this update, so that a step that never solves does not pay for it
  C.pc_ok = false;
  C.pc_ncached = 0;
  C.pc_pending = true;
  // other target lists depend on the cell geometry through their near-singular entries
  for (int k = 1; k < 3; k++)
    if (c->tl[k].valid) RBC_TRY(nearsing_prepare(c, c->tl[k]));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return RBC3D_OK;
}

// Rbc_BuildSurfaceSource(fFlag / gFlag) on the device (ModRbc.F90:760-802): after this call a density passed to
// rbc3d_cells_set_density WITHOUT its spline has the spline built on the GPU (SH filter to degree < nlat0 +
// Spline_Build_on_Sphere), instead of keeping the previous one.
int rbc3d_cells_enable_device_splines(rbc3d_ctx *c, int nlat0) {
  if (!c || !c->cells.mesh_set) return RBC3D_ESTATE;
  CUDA_TRY(cudaSetDevice(c->device));
  return spline_builder_prepare(c, nlat0);
}

int rbc3d_cells_set_density(rbc3d_ctx *c, const double *f, const double *g, const double *spF, const double *spG) {
  if (!c || !c->cells.geom_set) return RBC3D_ESTATE;
  CUDA_TRY(cudaSetDevice(c->device));
  Cells &C = c->cells;
  const size_t Np = C.Np, nc = C.ncell;
  const size_t sp3 = (size_t)12 * 2 * C.nlat * C.nlon;
  t_begin(c, RBC3D_T_H2D);
  const int nr = c->prm.nranks;
  if (nr > 1 && c->replicated_density && nc % nr == 0) {
    // every rank holds the same host arrays (the reference replicates surface state): each rank uploads the rows of
    // its own cell block over PCIe and the blocks are all-gathered over NVLink, component plane by component plane
    const size_t blk = (nc / nr) * (size_t)C.npc;
    for (int which = 0; which < 2; which++) {
      const double *src = which ? g : f;
      dbuf<double> &dst = which ? C.g : C.f;
      if (!src) continue;
      RBC_TRY(dst.resize(3 * Np));
      for (int d = 0; d < 3; d++) {
        const size_t off = (size_t)d * Np + (size_t)c->prm.rank * blk;
        CUDA_TRY(cudaMemcpyAsync(dst.p + off, src + off, sizeof(double) * blk, cudaMemcpyHostToDevice, c->stream));
      }
      for (int d = 0; d < 3; d++) RBC_TRY(comm_allgather_inplace(c, dst.p + (size_t)d * Np, blk));
    }
  } else {
    if (f) RBC_TRY(upload(C.f, f, 3 * Np, c->stream));
    if (g) RBC_TRY(upload(C.g, g, 3 * Np, c->stream));
  }
  if (spF) RBC_TRY(upload(C.spF, spF, sp3 * nc, c->stream));
  if (spG) RBC_TRY(upload(C.spG, spG, sp3 * nc, c->stream));
  t_end(c, RBC3D_T_H2D);
  if (f) C.f_set = true;
  if (g) C.g_set = true;
  if (spG) C.spGi_valid = false;
  if (f || spF) C.spFi_valid = false;
  RBC_TRY(cells_gather_sorted(c, false, f != nullptr, g != nullptr));
  if (spG) RBC_TRY(singular_density_prepare(c));
  if (C.sb_ok) {
    t_begin(c, RBC3D_T_DENSITY);
    if (f && !spF) RBC_TRY(spline_build_density(c, 0));
    if (g && !spG) RBC_TRY(spline_build_density(c, 1));
    t_end(c, RBC3D_T_DENSITY);
  }
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return RBC3D_OK;
}

int rbc3d_targets_set_raw(rbc3d_ctx *c, int n, const double *x, const int32_t *active) {
  if (!c || n < 0 || (n > 0 && !x)) return RBC3D_EINVAL;
  CUDA_TRY(cudaSetDevice(c->device));
  TargetList &t = c->tl[RBC3D_TL_RAW];
  t.kind = RBC3D_TL_RAW;
  t.n = n;
  const size_t n1 = n > 0 ? n : 1;
  RBC_TRY(t.x.resize(3 * n1));
  RBC_TRY(t.Acoef.resize(n1));
  RBC_TRY(t.surf.resize(n1));
  RBC_TRY(t.active.resize(n1));
  if (n > 0) {
    CUDA_TRY(cudaMemcpyAsync(t.x.p, x, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, c->stream));
    k_raw_target_meta<<<(n + 255) / 256, 256, 0, c->stream>>>(n, t.surf.p, t.Acoef.p);
    if (active)
      CUDA_TRY(cudaMemcpyAsync(t.active.p, active, sizeof(int) * n, cudaMemcpyHostToDevice, c->stream));
    else
      k_fill_int<<<(n + 255) / 256, 256, 0, c->stream>>>(n, t.active.p, 1);
    KERNEL_CHECK();
  }
  RBC_TRY(target_list_finish(c, t));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return RBC3D_OK;
}

// ---------------------------------------------------------------------------------------------------------
// walls: SourceList_UpdateCoord(slist_wall) + TargetList_Update(tlist_wall) (ModSourceList.F90:127-146,
// ModTargetList.F90:122-131)
int rbc3d_walls_set(rbc3d_ctx *c, int nwall, const int32_t *nvert, const int32_t *nele, const double *x,
                    const int32_t *e2v, const double *area, const double *epsDist, const int32_t *active) {
  if (!c || nwall < 0 || (nwall > 0 && (!nvert || !nele || !x || !e2v || !area || !epsDist))) return RBC3D_EINVAL;
  CUDA_TRY(cudaSetDevice(c->device));
  RBC_TRY(walls_set_geometry(c, nwall, nvert, nele, x, e2v, area, epsDist));
  Walls &W = c->walls;
  TargetList &t = c->tl[RBC3D_TL_WALLS];
  t.kind = RBC3D_TL_WALLS;
  t.n = W.NV;
  const size_t n1 = W.NV > 0 ? W.NV : 1;
  RBC_TRY(t.x.resize(3 * n1));
  RBC_TRY(t.Acoef.resize(n1));
  RBC_TRY(t.surf.resize(n1));
  RBC_TRY(t.active.resize(n1));
  if (W.NV > 0) {
    CUDA_TRY(cudaMemcpyAsync(t.x.p, W.x.p, sizeof(double) * 3 * W.NV, cudaMemcpyDeviceToDevice, c->stream));
    RBC_TRY(walls_target_meta(c, t));
    if (active)
      CUDA_TRY(cudaMemcpyAsync(t.active.p, active, sizeof(int) * W.NV, cudaMemcpyHostToDevice, c->stream));
    else
      k_fill_int<<<(W.NV + 255) / 256, 256, 0, c->stream>>>(W.NV, t.active.p, 1);
    KERNEL_CHECK();
  }
  RBC_TRY(target_list_finish(c, t));
  for (int k = 0; k < 3; k++) c->tl[k].wp.valid = false;
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return RBC3D_OK;
}

int rbc3d_walls_set_traction(rbc3d_ctx *c, const double *f) {
  if (!c || !f) return RBC3D_EINVAL;
  CUDA_TRY(cudaSetDevice(c->device));
  RBC_TRY(walls_set_traction(c, f));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return RBC3D_OK;
}

int rbc3d_wall_prepare_sing(rbc3d_ctx *c) {  // PrepareSingIntOnWall, ModIntOnWalls.F90:181-308, every wall
  if (!c) return RBC3D_EINVAL;
  CUDA_TRY(cudaSetDevice(c->device));
  return walls_prepare_sing(c);
}

int rbc3d_sing_int_on_wall(rbc3d_ctx *c, double c1, int iwall, double *v) {  // ModIntOnWalls.F90:136-172
  if (!c || !v || iwall < 0 || iwall >= c->walls.nwall) return RBC3D_EINVAL;
  CUDA_TRY(cudaSetDevice(c->device));
  Walls &W = c->walls;
  const int nv = W.h_nvert[iwall];
  TargetList &t = c->tl[RBC3D_TL_WALLS];
  RBC_TRY(t.host_io.resize(3 * (size_t)(t.n > 0 ? t.n : 1)));
  RBC_TRY(walls_sing_int(c, c1, iwall, t.host_io.p));
  if (nv > 0) CUDA_TRY(cudaMemcpyAsync(v, t.host_io.p, sizeof(double) * 3 * nv, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return RBC3D_OK;
}

int rbc3d_wall_matrix_get(rbc3d_ctx *c, int32_t *nblk, int32_t *rowptr, int32_t *col, double *val, int cap) {
  if (!c || !c->walls.mat_ok) return RBC3D_ESTATE;
  CUDA_TRY(cudaSetDevice(c->device));
  Walls &W = c->walls;
  if (nblk) *nblk = W.nblk;
  if (rowptr) CUDA_TRY(cudaMemcpy(rowptr, W.rowptr.p, sizeof(int) * (W.NV + 1), cudaMemcpyDeviceToHost));
  const int m = W.nblk < cap ? W.nblk : cap;
  if (m > 0 && col) CUDA_TRY(cudaMemcpy(col, W.col.p, sizeof(int) * m, cudaMemcpyDeviceToHost));
  if (m > 0 && val) CUDA_TRY(cudaMemcpy(val, W.val.p, sizeof(double) * 9 * m, cudaMemcpyDeviceToHost));
  return RBC3D_OK;
}

int rbc3d_wall_neighbor_signature(rbc3d_ctx *c, int tlist, int self_skip, int32_t *count, uint64_t *sig,
                                  int32_t *nduffy) {
  if (!c || tlist < 0 || tlist > 2 || !c->tl[tlist].valid) return RBC3D_ESTATE;
  CUDA_TRY(cudaSetDevice(c->device));
  return walls_signature(c, c->tl[tlist], self_skip, count, (unsigned long long *)sig, nduffy);
}

int rbc3d_min_dist_to_tri(rbc3d_ctx *c, int n, const double *xtar, const double *xtri, double *dist, double *s0,
                          double *t0) {
  if (!c || n < 0 || !xtar || !xtri || !dist) return RBC3D_EINVAL;
  CUDA_TRY(cudaSetDevice(c->device));
  return walls_min_dist_batch(c, n, xtar, xtri, dist, s0, t0);
}

int rbc3d_tri_int(rbc3d_ctx *c, int n, const double *xtri, const double *ftri, const double *xtar, const double *s0,
                  const double *t0, double *rhs, double *lhs) {
  if (!c || n < 0 || !xtri || !xtar || (!rhs && !lhs) || ((s0 == nullptr) != (t0 == nullptr))) return RBC3D_EINVAL;
  CUDA_TRY(cudaSetDevice(c->device));
  if (rhs) RBC_TRY(walls_tri_int_batch(c, n, xtri, ftri, xtar, s0, t0, rhs, nullptr));
  if (lhs) RBC_TRY(walls_tri_int_batch(c, n, xtri, ftri, xtar, s0, t0, nullptr, lhs));
  return RBC3D_OK;
}

// ---------------------------------------------------------------------------------------------------------
static int check_density(rbc3d_ctx *c, double c1, double c2) {
  if (c1 != 0 && !c->cells.f_set) {
    set_error("c1 != 0 but no single-layer density (f, spF) was set");
    return RBC3D_ESTATE;
  }
  if (c2 != 0 && !c->cells.g_set) {
    set_error("c2 != 0 but no double-layer density (g, spG) was set");
    return RBC3D_ESTATE;
  }
  return RBC3D_OK;
}

static int realspace_cells(rbc3d_ctx *c, TargetList &t, double c1, double c2) {
  // AddIntOnRbcs, ModIntOnRbcs.F90:25-158 (sums are un-normalised; combine() divides by Acoef)
  if (c->cells.ncell == 0) return RBC3D_OK;
  RBC_TRY(check_density(c, c1, c2));
  const int skip = c->skip_flags;
  if (!(skip & 8)) {
    t_begin(c, RBC3D_T_PAIR);
    RBC_TRY(pair_sum(c, t, c1, c2));
    t_end(c, RBC3D_T_PAIR);
  }
  if (!(skip & 1)) {
    t_begin(c, RBC3D_T_SING);
    RBC_TRY(singular_apply(c, t, c1, c2));
    t_end(c, RBC3D_T_SING);
  }
  if (!(skip & 2)) {
    t_begin(c, RBC3D_T_NEARSING);
    RBC_TRY(nearsing_apply(c, t, c1, c2));
    t_end(c, RBC3D_T_NEARSING);
  }
  return RBC3D_OK;
}

static int begin_apply(rbc3d_ctx *c, TargetList &t) {
  for (int i = 0; i < RBC3D_T_COUNT; i++) c->ev_used[i] = false;
  t_begin(c, RBC3D_T_TOTAL);
  CUDA_TRY(cudaMemsetAsync(t.acc.p, 0, sizeof(double) * 3 * (size_t)(t.n > 0 ? t.n : 1), c->stream));
  return RBC3D_OK;
}

static int v_roundtrip_begin(rbc3d_ctx *c, TargetList &t, const double *v) {
  RBC_TRY(t.host_io.resize(3 * (size_t)(t.n > 0 ? t.n : 1)));
  t_begin(c, RBC3D_T_H2D);
  if (t.n) CUDA_TRY(cudaMemcpyAsync(t.host_io.p, v, sizeof(double) * 3 * t.n, cudaMemcpyHostToDevice, c->stream));
  t_end(c, RBC3D_T_H2D);
  return RBC3D_OK;
}
static int v_roundtrip_end(rbc3d_ctx *c, TargetList &t, double *v) {
  t_begin(c, RBC3D_T_D2H);
  if (t.n) CUDA_TRY(cudaMemcpyAsync(v, t.host_io.p, sizeof(double) * 3 * t.n, cudaMemcpyDeviceToHost, c->stream));
  t_end(c, RBC3D_T_D2H);
  t_end(c, RBC3D_T_TOTAL);
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return RBC3D_OK;
}

static int get_tl(rbc3d_ctx *c, int tlist, TargetList **t) {
  if (!c || tlist < 0 || tlist > 2) return RBC3D_EINVAL;
  if (!c->tl[tlist].valid) {
    set_error("target list %d not set", tlist);
    return RBC3D_ESTATE;
  }
  *t = &c->tl[tlist];
  return cudaSetDevice(c->device) == cudaSuccess ? RBC3D_OK : RBC3D_ECUDA;
}

int rbc3d_add_int_on_rbcs(rbc3d_ctx *c, double c1, double c2, int tlist, double *v) {
  TargetList *t;
  RBC_TRY(get_tl(c, tlist, &t));
  RBC_TRY(begin_apply(c, *t));
  RBC_TRY(v_roundtrip_begin(c, *t, v));
  RBC_TRY(realspace_cells(c, *t, c1, c2));
  t_begin(c, RBC3D_T_LINEAR);
  RBC_TRY(linear_term(c, *t, (c->skip_flags & 4) ? 0.0 : c2));
  t_end(c, RBC3D_T_LINEAR);
  RBC_TRY(combine(c, *t, t->host_io.p, true));
  return v_roundtrip_end(c, *t, v);
}

int rbc3d_add_int_on_walls(rbc3d_ctx *c, double c1, int tlist, double *v) {  // ModIntOnWalls.F90:33-130
  TargetList *t;
  RBC_TRY(get_tl(c, tlist, &t));
  RBC_TRY(begin_apply(c, *t));
  RBC_TRY(v_roundtrip_begin(c, *t, v));
  t_begin(c, RBC3D_T_WALL);
  RBC_TRY(walls_add_int(c, *t, c1));
  t_end(c, RBC3D_T_WALL);
  RBC_TRY(linear_term(c, *t, 0.0));
  RBC_TRY(combine(c, *t, t->host_io.p, true));
  return v_roundtrip_end(c, *t, v);
}

int rbc3d_pme_distrib_source(rbc3d_ctx *c, double c1, double c2, int use_cells, int use_walls) {
  if (!c) return RBC3D_EINVAL;
  CUDA_TRY(cudaSetDevice(c->device));
  t_begin(c, RBC3D_T_SPREAD);
  RBC_TRY(pme_spread(c, c1, c2, use_cells != 0, use_walls != 0));
  t_end(c, RBC3D_T_SPREAD);
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return RBC3D_OK;
}

int rbc3d_pme_transform(rbc3d_ctx *c) {
  if (!c) return RBC3D_EINVAL;
  CUDA_TRY(cudaSetDevice(c->device));
  RBC_TRY(pme_transform(c));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return RBC3D_OK;
}

int rbc3d_pme_add_interp_vel(rbc3d_ctx *c, int tlist, double *v) {
  TargetList *t;
  RBC_TRY(get_tl(c, tlist, &t));
  CUDA_TRY(cudaMemsetAsync(t->acc.p, 0, sizeof(double) * 3 * (size_t)(t->n > 0 ? t->n : 1), c->stream));
  RBC_TRY(v_roundtrip_begin(c, *t, v));
  t_begin(c, RBC3D_T_INTERP);
  RBC_TRY(pme_interp(c, *t));
  t_end(c, RBC3D_T_INTERP);
  RBC_TRY(linear_term(c, *t, 0.0));
  RBC_TRY(combine(c, *t, t->host_io.p, true));
  return v_roundtrip_end(c, *t, v);
}

static int pme_chain(rbc3d_ctx *c, TargetList &t, double c1, double c2, int use_cells, int use_walls, double *acc) {
  t_begin(c, RBC3D_T_SPREAD);
  RBC_TRY(pme_spread(c, c1, c2, use_cells != 0, use_walls != 0));
  t_end(c, RBC3D_T_SPREAD);
  RBC_TRY(pme_transform(c));
  t_begin(c, RBC3D_T_INTERP);
  RBC_TRY(pme_interp(c, t, acc));
  t_end(c, RBC3D_T_INTERP);
  return RBC3D_OK;
}

// Returns in *acc2 the second accumulator the caller has to hand to combine() (null when the chain ran in line).
static int apply_common(rbc3d_ctx *c, TargetList &t, double c1, double c2, int use_cells, int use_walls,
                        const double **acc2 = nullptr) {
  // auto: with several ranks always (hides the exchanges); on one rank for large target lists, where the PME chain's
  // low-occupancy kernels fill in beside the real-space kernels (4096 cells: 82.5 -> 78.0 ms); small problems are
  // launch-bound and keep one stream
  const bool overlap = acc2 && (c->overlap == 1 || (c->overlap < 0 && (c->prm.nranks > 1 || t.n >= 100000)));
  if (acc2) *acc2 = nullptr;
  if (overlap) {
    // fork: the PME chain on stream2 (issued first so that its all-reduce and FFTs start early), real space on stream
    RBC_TRY(t.acc2.resize(3 * (size_t)(t.n > 0 ? t.n : 1)));
    CUDA_TRY(cudaEventRecord(c->ev_fork, c->stream));
    CUDA_TRY(cudaStreamWaitEvent(c->stream2, c->ev_fork, 0));
    std::swap(c->stream, c->stream2);
    int rc = cudaMemsetAsync(t.acc2.p, 0, sizeof(double) * 3 * (size_t)(t.n > 0 ? t.n : 1), c->stream) == cudaSuccess
                 ? RBC3D_OK
                 : RBC3D_ECUDA;
    t_begin(c, RBC3D_T_PME_CHAIN);
    if (rc == RBC3D_OK) rc = pme_chain(c, t, c1, c2, use_cells, use_walls, t.acc2.p);
    t_end(c, RBC3D_T_PME_CHAIN);
    if (rc == RBC3D_OK && cudaEventRecord(c->ev_join, c->stream) != cudaSuccess) rc = RBC3D_ECUDA;
    std::swap(c->stream, c->stream2);
    RBC_TRY(rc);
    *acc2 = t.acc2.p;
  }
  t_begin(c, RBC3D_T_REAL_CHAIN);
  if (use_cells) RBC_TRY(realspace_cells(c, t, c1, c2));
  if (use_walls && c1 != 0) {
    t_begin(c, RBC3D_T_WALL);
    RBC_TRY(walls_add_int(c, t, c1));
    t_end(c, RBC3D_T_WALL);
  }
  t_begin(c, RBC3D_T_LINEAR);
  RBC_TRY(linear_term(c, t, (use_cells && !(c->skip_flags & 4)) ? c2 : 0.0));
  t_end(c, RBC3D_T_LINEAR);
  t_end(c, RBC3D_T_REAL_CHAIN);
  if (overlap)
    CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev_join, 0));  // join before combine
  else
    RBC_TRY(pme_chain(c, t, c1, c2, use_cells, use_walls, nullptr));
  return RBC3D_OK;
}

int rbc3d_apply(rbc3d_ctx *c, double c1, double c2, int use_cells, int use_walls, int tlist, double *v) {
  TargetList *t;
  RBC_TRY(get_tl(c, tlist, &t));
  RBC_TRY(begin_apply(c, *t));
  RBC_TRY(v_roundtrip_begin(c, *t, v));
  const double *acc2 = nullptr;
  RBC_TRY(apply_common(c, *t, c1, c2, use_cells, use_walls, &acc2));
  t_begin(c, RBC3D_T_COMBINE);
  RBC_TRY(combine(c, *t, t->host_io.p, true, acc2));
  t_end(c, RBC3D_T_COMBINE);
  return v_roundtrip_end(c, *t, v);
}

// v = operator (not accumulated): the caller's "v = 0" before AddIntOnRbcs (ModVelSolver.F90:473,571) is part of the
// call, so v is never uploaded
int rbc3d_apply_assign(rbc3d_ctx *c, double c1, double c2, int use_cells, int use_walls, int tlist, double *v) {
  TargetList *t;
  RBC_TRY(get_tl(c, tlist, &t));
  RBC_TRY(begin_apply(c, *t));
  RBC_TRY(t->host_io.resize(3 * (size_t)(t->n > 0 ? t->n : 1)));
  const double *acc2 = nullptr;
  RBC_TRY(apply_common(c, *t, c1, c2, use_cells, use_walls, &acc2));
  t_begin(c, RBC3D_T_COMBINE);
  RBC_TRY(combine(c, *t, t->host_io.p, false, acc2));
  t_end(c, RBC3D_T_COMBINE);
  return v_roundtrip_end(c, *t, v);
}

// v = operator summed over the ranks: "v = 0", AddIntOnRbcs/AddIntOnWalls, the PME triple and
// TargetList_CollectArray of ModVelSolver.F90:571-584 as one call.  The sum over ranks happens on the devices
// (ncclAllReduce) before the one device-to-host copy, instead of download + upload + reduce + download.
int rbc3d_apply_collect(rbc3d_ctx *c, double c1, double c2, int use_cells, int use_walls, int tlist, double *v) {
  TargetList *t;
  RBC_TRY(get_tl(c, tlist, &t));
  if (!v) return RBC3D_EINVAL;
  RBC_TRY(begin_apply(c, *t));
  const double *acc2 = nullptr;
  RBC_TRY(apply_common(c, *t, c1, c2, use_cells, use_walls, &acc2));
  t_begin(c, RBC3D_T_COMBINE);
  RBC_TRY(combine(c, *t, t->v.p, false, acc2));
  t_end(c, RBC3D_T_COMBINE);
  if (c->prm.nranks > 1) {
    t_begin(c, RBC3D_T_COMM);
    RBC_TRY(comm_allreduce_sum(c, t->v.p, 3 * (size_t)t->n));
    t_end(c, RBC3D_T_COMM);
  }
  t_begin(c, RBC3D_T_D2H);
  if (t->n) CUDA_TRY(cudaMemcpyAsync(v, t->v.p, sizeof(double) * 3 * t->n, cudaMemcpyDeviceToHost, c->stream));
  t_end(c, RBC3D_T_D2H);
  t_end(c, RBC3D_T_TOTAL);
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return RBC3D_OK;
}

int rbc3d_apply_resident(rbc3d_ctx *c, double c1, double c2, int use_cells, int use_walls, int tlist) {
  TargetList *t;
  RBC_TRY(get_tl(c, tlist, &t));
  RBC_TRY(begin_apply(c, *t));
  // density-dependent part of SourceList_UpdateDensity (ModSourceList.F90:160-187) redone from the resident
  // f / g so that a resident step does the same per-matvec work as ModVelSolver.F90:560-565
  t_begin(c, RBC3D_T_DENSITY);
  if (use_cells) {
    RBC_TRY(cells_gather_sorted(c, false, c1 != 0 && c->cells.f_set, c2 != 0 && c->cells.g_set));
    if (c->cells.sb_ok) {  // the per-matvec Rbc_BuildSurfaceSource of ModVelSolver.F90:563, on the device
      if (c1 != 0 && c->cells.f_set) RBC_TRY(spline_build_density(c, 0));
      if (c2 != 0 && c->cells.g_set) RBC_TRY(spline_build_density(c, 1));
    } else if (c2 != 0 && c1 == 0 && t->kind == RBC3D_TL_CELLS) {
      RBC_TRY(singular_density_prepare(c));
    }
  }
  t_end(c, RBC3D_T_DENSITY);
  const double *acc2 = nullptr;
  RBC_TRY(apply_common(c, *t, c1, c2, use_cells, use_walls, &acc2));
  t_begin(c, RBC3D_T_COMBINE);
  RBC_TRY(combine(c, *t, t->v.p, false, acc2));
  t_end(c, RBC3D_T_COMBINE);
  if (c->prm.nranks > 1 && c->resident_collect) {  // TargetList_CollectArray
    t_begin(c, RBC3D_T_COMM);
    RBC_TRY(comm_allreduce_sum(c, t->v.p, 3 * (size_t)t->n));
    t_end(c, RBC3D_T_COMM);
  }
  t_end(c, RBC3D_T_TOTAL);
  if (!c->quiet) CUDA_TRY(cudaStreamSynchronize(c->stream));
  return RBC3D_OK;
}

// Closest_Neighbor_Cell / Closest_Neighbor_Wall (ModRepulsion.F90:480-613) for n points at once
int rbc3d_closest_neighbors(rbc3d_ctx *c, int n, const double *x, const int32_t *surf_id, double eps_dist,
                            double *dist_cell, double *x0_cell, double *dist_wall, double *x0_wall) {
  if (!c || n < 0 || (n > 0 && (!x || !surf_id))) return RBC3D_EINVAL;
  if (n == 0) return RBC3D_OK;
  CUDA_TRY(cudaSetDevice(c->device));
  const bool cells = c->cells.geom_set && c->cells.Np > 0 && dist_cell && x0_cell;
  const bool walls = c->walls.geom_set && c->walls.NE > 0 && dist_wall && x0_wall;
  dbuf<double> qx, out;
  dbuf<int> sid;
  RBC_TRY(qx.resize(3 * (size_t)n));
  RBC_TRY(out.resize(8 * (size_t)n));
  RBC_TRY(sid.resize(n));
  auto done = [&](int rc) {
    qx.release(), out.release(), sid.release();
    return rc;
  };
  if (cudaMemcpyAsync(qx.p, x, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, c->stream) != cudaSuccess ||
      cudaMemcpyAsync(sid.p, surf_id, sizeof(int) * n, cudaMemcpyHostToDevice, c->stream) != cudaSuccess)
    return done(RBC3D_ECUDA);
  double *dc = out.p, *xc = out.p + n, *dw = out.p + 4 * (size_t)n, *xw = out.p + 5 * (size_t)n;
  int rc = RBC3D_OK;
  if (cells) rc = closest_cells(c, n, qx.p, sid.p, eps_dist, dc, xc);
  if (rc == RBC3D_OK && walls) rc = closest_walls(c, n, qx.p, sid.p, dw, xw);
  if (rc != RBC3D_OK) return done(rc);
  const double huge = HUGE_VAL;  // dist0 = huge(dist0) when there is no such surface (ModRepulsion.F90:494, 571)
  if (dist_cell) {
    if (cells) {
      cudaMemcpyAsync(dist_cell, dc, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream);
      cudaMemcpyAsync(x0_cell, xc, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, c->stream);
    } else {
      for (int i = 0; i < n; i++) dist_cell[i] = huge;
    }
  }
  if (dist_wall) {
    if (walls) {
      cudaMemcpyAsync(dist_wall, dw, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream);
      cudaMemcpyAsync(x0_wall, xw, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, c->stream);
    } else {
      for (int i = 0; i < n; i++) dist_wall[i] = huge;
    }
  }
  if (cudaStreamSynchronize(c->stream) != cudaSuccess) return done(RBC3D_ECUDA);
  return done(RBC3D_OK);
}

int rbc3d_get_velocity(rbc3d_ctx *c, int tlist, double *v) {
  TargetList *t;
  RBC_TRY(get_tl(c, tlist, &t));
  if (t->n) CUDA_TRY(cudaMemcpyAsync(v, t->v.p, sizeof(double) * 3 * t->n, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return RBC3D_OK;
}

// ---------------------------------------------------------------------------------------------------------
int rbc3d_cell_list_get(rbc3d_ctx *c, int32_t Nc[3], int32_t *cid, int32_t *order, int32_t *start) {
  if (!c || !c->cells.geom_set) return RBC3D_ESTATE;
  CUDA_TRY(cudaSetDevice(c->device));
  Cells &C = c->cells;
  const Params &p = c->prm;
  const int ncells = p.Nc[0] * p.Nc[1] * p.Nc[2];
  if (Nc)
    for (int d = 0; d < 3; d++) Nc[d] = p.Nc[d];
  if (cid) CUDA_TRY(cudaMemcpy(cid, C.cl.cid.p, sizeof(int) * C.Np, cudaMemcpyDeviceToHost));
  if (order) CUDA_TRY(cudaMemcpy(order, C.cl.order.p, sizeof(int) * C.Np, cudaMemcpyDeviceToHost));
  if (start) CUDA_TRY(cudaMemcpy(start, C.cl.start.p, sizeof(int) * (ncells + 1), cudaMemcpyDeviceToHost));
  return RBC3D_OK;
}

int rbc3d_neighbor_signature(rbc3d_ctx *c, int tlist, int32_t *count, uint64_t *sig) {
  TargetList *t;
  RBC_TRY(get_tl(c, tlist, &t));
  return neighbor_signature(c, *t, count, (unsigned long long *)sig);
}

int rbc3d_nearsing_get(rbc3d_ctx *c, int tlist, int *n, int32_t *target, int32_t *cell, int32_t *flag, double *th0,
                       double *phi0, double *dist, int cap) {
  TargetList *t;
  RBC_TRY(get_tl(c, tlist, &t));
  NearSing &ns = t->ns;
  if (n) *n = ns.n;
  const int m = ns.n < cap ? ns.n : cap;
  if (m <= 0) return RBC3D_OK;
  if (target) CUDA_TRY(cudaMemcpy(target, ns.target.p, sizeof(int) * m, cudaMemcpyDeviceToHost));
  if (cell) CUDA_TRY(cudaMemcpy(cell, ns.cell.p, sizeof(int) * m, cudaMemcpyDeviceToHost));
  if (flag) CUDA_TRY(cudaMemcpy(flag, ns.flag.p, sizeof(int) * m, cudaMemcpyDeviceToHost));
  if (th0) CUDA_TRY(cudaMemcpy(th0, ns.th0.p, sizeof(double) * m, cudaMemcpyDeviceToHost));
  if (phi0) CUDA_TRY(cudaMemcpy(phi0, ns.phi0.p, sizeof(double) * m, cudaMemcpyDeviceToHost));
  if (dist) CUDA_TRY(cudaMemcpy(dist, ns.dist.p, sizeof(double) * m, cudaMemcpyDeviceToHost));
  return RBC3D_OK;
}

int rbc3d_cells_get_density_spline(rbc3d_ctx *c, int which, double *sp) {
  if (!c || !sp || !c->cells.geom_set) return RBC3D_ESTATE;
  CUDA_TRY(cudaSetDevice(c->device));
  Cells &C = c->cells;
  dbuf<double> &b = which ? C.spG : C.spF;
  const size_t n = (size_t)C.ncell * 12 * 2 * C.nlat * C.nlon;
  if (b.n < n) return RBC3D_ESTATE;
  CUDA_TRY(cudaMemcpy(sp, b.p, sizeof(double) * n, cudaMemcpyDeviceToHost));
  return RBC3D_OK;
}

int rbc3d_cells_get_geometry_spline(rbc3d_ctx *c, int which, double *sp) {
  if (!c || !sp || !c->cells.geom_set || which < 0 || which > 2) return RBC3D_ESTATE;
  CUDA_TRY(cudaSetDevice(c->device));
  Cells &C = c->cells;
  dbuf<double> &b = which == 0 ? C.spx : which == 1 ? C.spa3 : C.spdetj;
  const size_t n = (size_t)C.ncell * (which == 2 ? 4 : 12) * 2 * C.nlat * C.nlon;
  if (b.n < n) return RBC3D_ESTATE;
  CUDA_TRY(cudaMemcpy(sp, b.p, sizeof(double) * n, cudaMemcpyDeviceToHost));
  return RBC3D_OK;
}

int rbc3d_pme_get_grid(rbc3d_ctx *c, double *vv) {
  if (!c || !c->pme.transformed) return RBC3D_ESTATE;
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaMemcpy(vv, c->pme.vv.p, sizeof(double) * 3 * c->pme.G, cudaMemcpyDeviceToHost));
  return RBC3D_OK;
}

int rbc3d_get_timings(rbc3d_ctx *c, float ms[RBC3D_T_COUNT]) {
  if (!c) return RBC3D_EINVAL;
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  for (int i = 0; i < RBC3D_T_COUNT; i++) {
    ms[i] = 0.f;
    if (c->ev_used[i]) {
      float t = 0.f;
      if (cudaEventElapsedTime(&t, c->ev[2 * i], c->ev[2 * i + 1]) == cudaSuccess) ms[i] = t;
    }
  }
  return RBC3D_OK;
}

int rbc3d_get_launch_count(rbc3d_ctx *c, long long *launches) {
  if (!c || !launches) return RBC3D_EINVAL;
  *launches = c->launches;
  return RBC3D_OK;
}

// ---- FP64 FMA peak: 8 independent register chains per thread ----
__global__ void __launch_bounds__(256) k_fp64_peak(double *out, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6,
         x7 = x0 + 7;
  for (int i = 0; i < iters; i++) {
    x0 = fma(x0, a, b);
    x1 = fma(x1, a, b);
    x2 = fma(x2, a, b);
    x3 = fma(x3, a, b);
    x4 = fma(x4, a, b);
    x5 = fma(x5, a, b);
    x6 = fma(x6, a, b);
    x7 = fma(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

int rbc3d_measure_fp64_peak(int device, double *tflops) {
  if (!tflops) return RBC3D_EINVAL;
  if (device >= 0) CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  int dev;
  CUDA_TRY(cudaGetDevice(&dev));
  CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 16;
  double *out;
  CUDA_TRY(cudaMalloc((void **)&out, sizeof(double) * blocks * threads));
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreate(&e0));
  CUDA_TRY(cudaEventCreate(&e1));
  double best = 0;
  for (int rep = 0; rep < 5; rep++) {
    CUDA_TRY(cudaEventRecord(e0));
    k_fp64_peak<<<blocks, threads>>>(out, iters, 0.999999, 1e-7);
    CUDA_TRY(cudaEventRecord(e1));
    CUDA_TRY(cudaEventSynchronize(e1));
    float ms;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    double tf = 2.0 * 8.0 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  *tflops = best;
  return RBC3D_OK;
}

}  // extern "C"
