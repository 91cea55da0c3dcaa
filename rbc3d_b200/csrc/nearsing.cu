// nearsing.cu -- near-singular correction of the real-space sum for targets close to ANOTHER cell's surface:
// the tail of AddIntOnRbcs (ModIntOnRbcs.F90:123-148), Spline_FindProjection (ModSpline.F90:203-269),
// RBC_NearSingInt / _Subtract / _ReAdd (ModRbcSingInt.F90:103-310).
//
// Geometry time (once per SourceList_UpdateCoord): neighbour-cell scan (pairsum.cu) -> entry list
// (target, other cell, closest mesh point) -> projection of the target on the other cell's spline surface,
// signed normal distance and the distance check.  None of that depends on the densities, so GMRES matvecs
// reuse it.  Per application: one warp per active entry evaluates the subtract + re-add quadratures, a second
// kernel adds the entries of each target in list order (deterministic, no atomics).
#include <cfloat>
#include <cub/cub.cuh>

#include "device_math.cuh"
#include "rbc3d_internal.h"

namespace rbc3d {

constexpr int NS_WARPS = 4;

// LAPACK dposv('U', 6, 1) restated: unblocked Cholesky A = U^T U, two triangular solves (ModBasicMath.F90:273)
__device__ __forceinline__ int chol_solve6(double *a, double *b) {
  const int n = 6;
#define A_(i, j) a[(i) + n * (j)]
  for (int j = 0; j < n; j++) {
    double ajj = A_(j, j);
    for (int k = 0; k < j; k++) ajj -= A_(k, j) * A_(k, j);
    if (!(ajj > 0.0)) return j + 1;
    ajj = sqrt(ajj);
    A_(j, j) = ajj;
    for (int i = j + 1; i < n; i++) {
      double s = A_(j, i);
      for (int k = 0; k < j; k++) s -= A_(k, j) * A_(k, i);
      A_(j, i) = s / ajj;
    }
  }
  for (int i = 0; i < n; i++) {
    double s = b[i];
    for (int k = 0; k < i; k++) s -= A_(k, i) * b[k];
    b[i] = s / A_(i, i);
  }
  for (int i = n - 1; i >= 0; i--) {
    double s = b[i];
    for (int k = i + 1; k < n; k++) s -= A_(i, k) * b[k];
    b[i] = s / A_(i, i);
  }
#undef A_
  return 0;
}

// Spline_FindProjection, ModSpline.F90:203-269 (one thread)
__device__ void find_projection(const double *__restrict__ spx, int m, int n, const double xt[3], double &th0,
                                double &phi0, double x0[3]) {
  const int nth = 2, nphi = 8;
  spline_interp<3>(spx, m, n, th0, phi0, x0);
  double h = fmax(RBC_TWO_PI / (double)m, RBC_TWO_PI / (double)n);
  for (int iter = 1; iter <= 3; iter++) {
    double lhs[36], rhs[6];
#pragma unroll
    for (int i = 0; i < 36; i++) lhs[i] = 0.0;
#pragma unroll
    for (int i = 0; i < 6; i++) rhs[i] = 0.0;
    double st0, ct0;
    sincos(th0, &st0, &ct0);
    double xx[3];
    spline_interp<3>(spx, m, n, th0, phi0, xx);
    xx[0] -= xt[0];
    xx[1] -= xt[1];
    xx[2] -= xt[2];
    const double d2_0 = xx[0] * xx[0] + xx[1] * xx[1] + xx[2] * xx[2];
    // accumulate the normal equations of QuadFit_2D (ModBasicMath.F90:241-262), point 0 first
    for (int i = 0; i <= nth * nphi; i++) {
      double px = 0.0, py = 0.0, d2 = d2_0;
      if (i > 0) {
        const int ith = (i - 1) % nth + 1, iphi = (i - 1) / nth + 1;
        const double thL = (double)ith * h / (double)nth;
        const double phL = (double)(iphi - 1) * RBC_TWO_PI / (double)nphi;
        double sp, cp;
        sincos(phL, &sp, &cp);
        px = thL * cp;
        py = thL * sp;
        double thG, phG;
        polar_patch_point(st0, ct0, phi0, thL, phL, thG, phG);
        spline_interp<3>(spx, m, n, thG, phG, xx);
        xx[0] -= xt[0];
        xx[1] -= xt[1];
        xx[2] -= xt[2];
        d2 = xx[0] * xx[0] + xx[1] * xx[1] + xx[2] * xx[2];
      }
      const double u[6] = {1.0, px, py, px * px, px * py, py * py};
#pragma unroll
      for (int ii = 0; ii < 6; ii++) {
#pragma unroll
        for (int jj = ii; jj < 6; jj++) lhs[ii + 6 * jj] += u[ii] * u[jj];
        rhs[ii] += u[ii] * d2;
      }
    }
    chol_solve6(lhs, rhs);
    // Min_Quad_2D, ModBasicMath.F90:300-326
    const double a1 = rhs[1], a2 = rhs[2], a11 = rhs[3], a12 = rhs[4], a22 = rhs[5];
    const double l11 = 2.0 * a11, l22 = 2.0 * a22;
    const double det = l11 * l22 - a12 * a12;
    double xm = 0.0, ym = 0.0;
    if (det > 0) {
      const double idet = 1.0 / det;
      xm = idet * (l22 * (-a1) - a12 * (-a2));
      ym = idet * (-a12 * (-a1) + l11 * (-a2));
    }
    const double thMin_L = sqrt(xm * xm + ym * ym);
    const double phiMin_L = atan2(ym, xm);
    double thMin, phiMin, xMin[3];
    polar_patch_map(th0, phi0, thMin_L, phiMin_L, thMin, phiMin);
    spline_interp<3>(spx, m, n, thMin, phiMin, xMin);
    const double e0 = xMin[0] - xt[0], e1 = xMin[1] - xt[1], e2 = xMin[2] - xt[2];
    const double d2min = e0 * e0 + e1 * e1 + e2 * e2;
    if (d2min > d2_0) break;
    th0 = thMin;
    phi0 = phiMin;
    x0[0] = xMin[0];
    x0[1] = xMin[1];
    x0[2] = xMin[2];
    h = 0.5 * h;
  }
}

struct NsArgs {
  Params prm;
  int n;  // entries
  int Np, npc, nlat, nlon, nt;
  const double *th, *phi, *w;
  const double *x, *a3, *f, *g;  // cell mesh fields SoA(3,Np) (f, g weighted)
  const double *spx, *spa3, *spdetj, *spF, *spG;
  const double *Bcell, *area, *meshSize;
  double radius;
  const double *tx;  // targets SoA(3,nt)
  const int *e_target, *e_cell, *e_pt;
  double *e_th0, *e_phi0, *e_dist, *e_x0, *e_a30, *e_xi;
  int *e_flag;
  const double *tab_sl, *tab_dl, *tab_mask;
  double c1, c2;
  double *e_dv;  // SoA(3,n)
};

// geometry-time: ModIntOnRbcs.F90:131-140 + ModRbcSingInt.F90:118-125
__global__ void k_ns_project(NsArgs a) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= a.n) return;
  const int t = a.e_target[e], cj = a.e_cell[e], pt = a.e_pt[e];
  const int ilon0 = pt / a.nlat, ilat0 = pt - ilon0 * a.nlat;
  double th0 = a.th[ilat0], phi0 = a.phi[ilon0];
  const size_t q = (size_t)cj * a.npc + pt;
  double x0[3], xi[3];
#pragma unroll
  for (int d = 0; d < 3; d++) {
    x0[d] = a.x[(size_t)d * a.Np + q];
    double xx = __dsub_rn(a.tx[(size_t)d * a.nt + t], x0[d]);
    xx = __dsub_rn(xx, __dmul_rn(round(__dmul_rn(xx, a.prm.iLb[d])), a.prm.Lb[d]));
    xi[d] = x0[d] + xx;
  }
  const int m = 2 * a.nlat, n = a.nlon;
  const size_t sp3 = (size_t)12 * m * n;
  find_projection(a.spx + sp3 * cj, m, n, xi, th0, phi0, x0);
  double a30[3];
  spline_interp<3>(a.spa3 + sp3 * cj, m, n, th0, phi0, a30);
  const double dist = a30[0] * (xi[0] - x0[0]) + a30[1] * (xi[1] - x0[1]) + a30[2] * (xi[2] - x0[2]);
  a.e_th0[e] = th0;
  a.e_phi0[e] = phi0;
  a.e_dist[e] = dist;
#pragma unroll
  for (int d = 0; d < 3; d++) {
    a.e_x0[(size_t)d * a.n + e] = x0[d];
    a.e_a30[(size_t)d * a.n + e] = a30[d];
    a.e_xi[(size_t)d * a.n + e] = xi[d];
  }
  a.e_flag[e] = (dist > 2.0 * a.meshSize[cj]) ? 0 : 1;
}

// RBC_NearSingInt_ReAdd (ModRbcSingInt.F90:232-310), one warp; returns the lane-partial sums.
template <bool SL, bool DL>
__device__ void readd(const NsArgs &a, int cj, const double xe[3], const double x0[3], double th0, double phi0,
                      double c2m, double dv[3]) {
  const int lane = threadIdx.x & 31;
  const int NRAD = 16, NAZM = 32;
  const int m = 2 * a.nlat, n = a.nlon;
  const size_t sp3 = (size_t)12 * m * n;
  const double *spx = a.spx + sp3 * cj, *spa3 = a.spa3 + sp3 * cj;
  const double radPat = a.radius;
  const double d0 = xe[0] - x0[0], d1 = xe[1] - x0[1], d2 = xe[2] - x0[2];
  double dist = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
  const double sizePat = radPat * sqrt(a.area[cj] / (4.0 * RBC_PI));
  // radial rule: node irad = lane & 15 (lanes 16..31 duplicate lanes 0..15)
  const int irad = lane & 15;
  const int jroot = irad < 8 ? irad + 1 : 16 - irad;
  double z, wt;
  gauleg_root(NRAD, jroot, z, wt);
  const double sgn = irad < 8 ? -1.0 : 1.0;
  double thP, wP;
  if (dist > DBL_MIN) {
    // GauLeg_Sinh(0, radPat, 0, dist, ...) with dist scaled to the unit sphere, ModQuadRule.F90:220-263
    dist = dist * (radPat / sizePat);
    const double xm = 0.5 * radPat, xl = 0.5 * radPat;
    const double a0 = (0.0 - xm) / xl, b0 = dist * xl;
    const double f1 = (1.0 + a0) / b0, f2 = (1.0 - a0) / b0;
    const double u1 = log(f1 + sqrt(1.0 + f1 * f1)), u2 = log(f2 + sqrt(1.0 + f2 * f2));
    const double mu = 0.5 * (u1 + u2), eta = 0.5 * (u1 - u2);
    const double s = sgn * z;  // GauLeg(-1,1): x = xm -+ xl*z with xm = 0, xl = 1
    const double arg = mu * s - eta;
    double xs = a0 + b0 * sinh(arg);
    double ws = wt * b0 * mu * cosh(arg);
    wP = xl * ws;
    thP = xm + xl * xs;
  } else {
    const double xm = 0.5 * radPat, xl = 0.5 * radPat;
    thP = xm + sgn * xl * z;
    wP = xl * wt;
  }
  wP = wP * sin(thP) * (RBC_TWO_PI / (double)NAZM);
  wP = wP * mask_func(a.tab_mask, thP / radPat);
  double st0, ct0;
  sincos(th0, &st0, &ct0);
  dv[0] = dv[1] = dv[2] = 0.0;
  for (int k = 0; k < (NRAD * NAZM) / 32; k++) {
    const int iazm = (lane >> 4) + 2 * k;
    const double phL = (double)iazm * RBC_TWO_PI / (double)NAZM;
    double thG, phG, xj[3];
    polar_patch_point(st0, ct0, phi0, thP, phL, thG, phG);
    spline_interp<3>(spx, m, n, thG, phG, xj);
    const double xx = xj[0] - xe[0], yy = xj[1] - xe[1], zz = xj[2] - xe[2];
    const double rr = sqrt(xx * xx + yy * yy + zz * zz);
    if (rr > a.prm.rc) continue;
    if (SL) {
      double fj[3], EA, EB;
      spline_interp<3>(a.spF + sp3 * cj, m, n, thG, phG, fj);
      fj[0] *= wP;
      fj[1] *= wP;
      fj[2] *= wP;
      ewald_sl(a.tab_sl, a.prm, rr, EA, EB);
      const double xf = EA * (xx * fj[0] + yy * fj[1] + zz * fj[2]);
      dv[0] += a.c1 * (xf * xx + EB * fj[0]);
      dv[1] += a.c1 * (xf * yy + EB * fj[1]);
      dv[2] += a.c1 * (xf * zz + EB * fj[2]);
    }
    if (DL) {
      double gj[3], nj[3];
      spline_interp<3>(spa3, m, n, thG, phG, nj);
      spline_interp<3>(a.spG + sp3 * cj, m, n, thG, phG, gj);
      const double EA = ewald_dl(a.tab_dl, a.prm, rr);
      const double qd = c2m * EA * wP * (xx * gj[0] + yy * gj[1] + zz * gj[2]) *
                        (xx * nj[0] + yy * nj[1] + zz * nj[2]);
      dv[0] += qd * xx;
      dv[1] += qd * yy;
      dv[2] += qd * zz;
    }
  }
}

// RBC_NearSingInt (ModRbcSingInt.F90:103-167), one warp per entry
template <bool SL, bool DL>
__global__ void __launch_bounds__(NS_WARPS * 32) k_ns_apply(NsArgs a) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int e = blockIdx.x * NS_WARPS + warp;
  if (e >= a.n) return;
  if (!a.e_flag[e]) {
    if (lane < 3) a.e_dv[(size_t)lane * a.n + e] = 0.0;
    return;
  }
  const int cj = a.e_cell[e];
  const double th0 = a.e_th0[e], phi0 = a.e_phi0[e], dist = a.e_dist[e];
  double x0[3], a30[3], xi[3];
#pragma unroll
  for (int d = 0; d < 3; d++) {
    x0[d] = a.e_x0[(size_t)d * a.n + e];
    a30[d] = a.e_a30[(size_t)d * a.n + e];
    xi[d] = a.e_xi[(size_t)d * a.n + e];
  }
  const double c2m = a.c2 * a.Bcell[cj];
  const double radPat = a.radius;
  const double sizePat = radPat * sqrt(a.area[cj] / (4.0 * RBC_PI));
  const double dist1 = copysign(0.01 * sizePat, dist);

  // ---- subtract (ModRbcSingInt.F90:175-226): mesh points inside the patch around the projection,
  //      enumerated like PolarPatch_FindPoints (ModPolarPatch.F90:162-206) ----
  double sx = 0, sy = 0, sz = 0;
  {
    const double eps = 1.e-10;
    const double h_phi = RBC_TWO_PI / (double)a.nlon, ih_phi = 1.0 / h_phi;
    const double cosr0 = cos(radPat);
    const double phi_first = a.phi[0];
    for (int i = 0; i < a.nlat; i++) {
      const double thi = a.th[i];
      if (thi <= th0 - radPat) continue;
      if (thi >= th0 + radPat) break;
      const double dth = thi - th0;
      double dphi = 1.0 - (cos(dth) - cosr0) / ((sin(thi) + eps) * (sin(th0) + eps));
      dphi = fmax(-1.0, fmin(1.0, dphi));
      dphi = acos(dphi);
      int jmin, jmax;
      if (dphi > RBC_PI - eps) {
        jmin = 1;
        jmax = a.nlon;
      } else {
        jmin = (int)ceil((phi0 - dphi - phi_first) * ih_phi) + 1;
        jmax = (int)floor((phi0 + dphi - phi_first) * ih_phi) + 1;
      }
      for (int j = jmin + lane; j <= jmax; j += 32) {
        const int ilon = imodulo(j - 1, a.nlon);
        const size_t q = (size_t)cj * a.npc + (size_t)ilon * a.nlat + i;
        const double xx = a.x[q] - xi[0], yy = a.x[(size_t)a.Np + q] - xi[1], zz = a.x[2 * (size_t)a.Np + q] - xi[2];
        const double rr = sqrt(xx * xx + yy * yy + zz * zz);
        if (rr > a.prm.rc) continue;
        const double dsp = dist_on_sphere(th0, phi0, thi, a.phi[ilon]);
        const double mask = mask_func(a.tab_mask, dsp / radPat);
        if (SL) {
          double EA, EB;
          const double fx = a.f[q], fy = a.f[(size_t)a.Np + q], fz = a.f[2 * (size_t)a.Np + q];
          ewald_sl(a.tab_sl, a.prm, rr, EA, EB);
          const double xf = EA * (xx * fx + yy * fy + zz * fz);
          sx -= a.c1 * mask * (xf * xx + EB * fx);
          sy -= a.c1 * mask * (xf * yy + EB * fy);
          sz -= a.c1 * mask * (xf * zz + EB * fz);
        }
        if (DL) {
          const double gx = a.g[q], gy = a.g[(size_t)a.Np + q], gz = a.g[2 * (size_t)a.Np + q];
          const double nx = a.a3[q], ny = a.a3[(size_t)a.Np + q], nz = a.a3[2 * (size_t)a.Np + q];
          const double EA = ewald_dl(a.tab_dl, a.prm, rr);
          const double qd = c2m * mask * EA * (xx * gx + yy * gy + zz * gz) * (xx * nx + yy * ny + zz * nz);
          sx -= qd * xx;
          sy -= qd * yy;
          sz -= qd * zz;
        }
      }
    }
  }

  // ---- re-add ----
  double r[3];
  if (fabs(dist) >= fabs(dist1)) {
    readd<SL, DL>(a, cj, xi, x0, th0, phi0, c2m, r);
    sx += r[0];
    sy += r[1];
    sz += r[2];
    sx = warp_sum(sx);
    sy = warp_sum(sy);
    sz = warp_sum(sz);
  } else {
    double xi1[3], r0[3];
#pragma unroll
    for (int d = 0; d < 3; d++) xi1[d] = x0[d] + dist1 * a30[d];
    readd<SL, DL>(a, cj, xi1, x0, th0, phi0, c2m, r);
    readd<SL, DL>(a, cj, x0, x0, th0, phi0, c2m, r0);
#pragma unroll
    for (int d = 0; d < 3; d++) {
      r[d] = warp_sum(r[d]);
      r0[d] = warp_sum(r0[d]);
    }
    if (DL) {  // jump condition, ModRbcSingInt.F90:148-160
      const int m = 2 * a.nlat, n = a.nlon;
      double detJ0[1], g0[3];
      spline_interp<1>(a.spdetj + (size_t)4 * m * n * cj, m, n, th0, phi0, detJ0);
      spline_interp<3>(a.spG + (size_t)12 * m * n * cj, m, n, th0, phi0, g0);
      const double sg = dist > 0 ? 1.0 : -1.0;
#pragma unroll
      for (int d = 0; d < 3; d++) r0[d] += sg * c2m * 4.0 * RBC_PI * (g0[d] / detJ0[0]);
    }
    sx = warp_sum(sx) + r0[0] + dist / dist1 * (r[0] - r0[0]);
    sy = warp_sum(sy) + r0[1] + dist / dist1 * (r[1] - r0[1]);
    sz = warp_sum(sz) + r0[2] + dist / dist1 * (r[2] - r0[2]);
  }
  if (lane == 0) {
    a.e_dv[e] = sx;
    a.e_dv[(size_t)a.n + e] = sy;
    a.e_dv[2 * (size_t)a.n + e] = sz;
  }
}

// add the entries of each target in list order
__global__ void k_ns_reduce(int nsorted, const int *__restrict__ torder, const int *__restrict__ cnt,
                            const int *__restrict__ off, const double *__restrict__ e_dv, int n, int nt,
                            double *__restrict__ acc) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nsorted) return;
  const int m = cnt[k];
  if (m == 0) return;
  const int o = off[k], ti = torder[k];
  double s0 = 0, s1 = 0, s2 = 0;
  for (int q = 0; q < m; q++) {
    s0 += e_dv[o + q];
    s1 += e_dv[(size_t)n + o + q];
    s2 += e_dv[2 * (size_t)n + o + q];
  }
  acc[ti] += s0;
  acc[(size_t)nt + ti] += s1;
  acc[2 * (size_t)nt + ti] += s2;
}

static void fill_ns(rbc3d_ctx *c, TargetList &t, NsArgs &a) {
  Cells &C = c->cells;
  NearSing &ns = t.ns;
  a.prm = c->prm;
  a.n = ns.n;
  a.Np = C.Np;
  a.npc = C.npc;
  a.nlat = C.nlat;
  a.nlon = C.nlon;
  a.nt = t.n;
  a.th = C.th.p;
  a.phi = C.phi.p;
  a.w = C.w.p;
  a.x = C.x.p;
  a.a3 = C.a3.p;
  a.f = C.f.p;
  a.g = C.g.p;
  a.spx = C.spx.p;
  a.spa3 = C.spa3.p;
  a.spdetj = C.spdetj.p;
  a.spF = C.spF.p;
  a.spG = C.spG.p;
  a.Bcell = C.B.p;
  a.area = C.area.p;
  a.meshSize = C.meshSize.p;
  a.radius = C.radius;
  a.tx = t.x.p;
  a.e_target = ns.target.p;
  a.e_cell = ns.cell.p;
  a.e_pt = ns.pt.p;
  a.e_th0 = ns.th0.p;
  a.e_phi0 = ns.phi0.p;
  a.e_dist = ns.dist.p;
  a.e_x0 = ns.x0.p;
  a.e_a30 = ns.a30.p;
  a.e_xi = ns.xi.p;
  a.e_flag = ns.flag.p;
  a.tab_sl = c->tab_sl.p;
  a.tab_dl = c->tab_dl.p;
  a.tab_mask = c->tab_mask.p;
  a.c1 = a.c2 = 0;
  a.e_dv = nullptr;
}

int nearsing_prepare(rbc3d_ctx *c, TargetList &t) {
  NearSing &ns = t.ns;
  Cells &C = c->cells;
  ns.n = ns.n_active = 0;
  if (C.Np == 0 || t.ntiles == 0) return RBC3D_OK;
  CellList &tcl = t.cl;
  const int nsorted = tcl.n_sorted;
  RBC_TRY(ns.cnt.resize((size_t)nsorted + 1));
  RBC_TRY(ns.off.resize((size_t)nsorted + 1));
  RBC_TRY(ns.overflow.resize(1));
  CUDA_TRY(cudaMemsetAsync(ns.cnt.p, 0, sizeof(int) * ((size_t)nsorted + 1), c->stream));
  CUDA_TRY(cudaMemsetAsync(ns.overflow.p, 0, sizeof(int), c->stream));
  RBC_TRY(nearsing_scan(c, t, false));
  size_t bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, ns.cnt.p, ns.off.p, nsorted + 1, c->stream);
  RBC_TRY(tcl.cub_tmp.resize(bytes + 256));
  size_t avail = tcl.cub_tmp.n;
  CUDA_TRY(cub::DeviceScan::ExclusiveSum(tcl.cub_tmp.p, avail, ns.cnt.p, ns.off.p, nsorted + 1, c->stream));
  int total = 0, ovf = 0;
  CUDA_TRY(cudaMemcpyAsync(&total, ns.off.p + nsorted, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaMemcpyAsync(&ovf, ns.overflow.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  if (ovf) {
    set_error("neighbour-cell list of a target exceeds %d cells (ModRbcSingInt.F90:319)", RBC3D_NBR_MAX);
    return RBC3D_EOVERFLOW;
  }
  ns.n = total;
  if (total == 0) return RBC3D_OK;
  const size_t n = total;
  RBC_TRY(ns.target.resize(n));
  RBC_TRY(ns.cell.resize(n));
  RBC_TRY(ns.pt.resize(n));
  RBC_TRY(ns.th0.resize(n));
  RBC_TRY(ns.phi0.resize(n));
  RBC_TRY(ns.dist.resize(n));
  RBC_TRY(ns.flag.resize(n));
  RBC_TRY(ns.x0.resize(3 * n));
  RBC_TRY(ns.a30.resize(3 * n));
  RBC_TRY(ns.xi.resize(3 * n));
  RBC_TRY(nearsing_scan(c, t, true));
  NsArgs a;
  fill_ns(c, t, a);
  k_ns_project<<<(total + 63) / 64, 64, 0, c->stream>>>(a);
  KERNEL_CHECK();
  c->launches += 2;
  return RBC3D_OK;
}

int nearsing_apply(rbc3d_ctx *c, TargetList &t, double c1, double c2) {
  NearSing &ns = t.ns;
  if (ns.n == 0) return RBC3D_OK;
  const bool sl = (c1 != 0), dl = (c2 != 0);
  if (!sl && !dl) return RBC3D_OK;
  CellList &tcl = t.cl;
  NsArgs a;
  fill_ns(c, t, a);
  a.c1 = c1;
  a.c2 = c2;
  RBC_TRY(ns.dv.resize(3 * (size_t)ns.n));
  a.e_dv = ns.dv.p;
  const int grid = (ns.n + NS_WARPS - 1) / NS_WARPS;
  if (sl && dl)
    k_ns_apply<true, true><<<grid, NS_WARPS * 32, 0, c->stream>>>(a);
  else if (sl)
    k_ns_apply<true, false><<<grid, NS_WARPS * 32, 0, c->stream>>>(a);
  else
    k_ns_apply<false, true><<<grid, NS_WARPS * 32, 0, c->stream>>>(a);
  KERNEL_CHECK();
  k_ns_reduce<<<(tcl.n_sorted + 255) / 256, 256, 0, c->stream>>>(tcl.n_sorted, tcl.order.p, ns.cnt.p, ns.off.p,
                                                                 a.e_dv, ns.n, t.n, t.acc.p);
  KERNEL_CHECK();
  c->launches += 2;
  return RBC3D_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Closest_Neighbor_Cell (ModRepulsion.F90:480-546) on the GPU cell list.  One warp per query point:
// the lanes scan the points of the 27 neighbouring list cells (sorted copies, coalesced), skip the query's own surface,
// and agree on the closest mesh point (ties: the smaller point index); if it is within 2 epsDist, lane 0 refines it by
// Spline_FindProjection on that cell's spline surface, started from the mesh point's (theta, phi).
struct ClosestArgs {
  Params prm;
  int n, Np, npc, nlat, nlon;
  const double *qx;           // query points SoA(3,n)
  const int *surf;            // surface id of each query point (cells are 1..ncell), ModRepulsion.F90:507
  const int *start, *order;   // cell list of the cell points
  const double *sx;           // sorted coordinates SoA(3,Np)
  const double *x, *spx, *th, *phi;
  double epsDist;
  double *dist, *x0;          // [n], SoA(3,n)
};

__global__ void __launch_bounds__(256) k_closest_cell(ClosestArgs a) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= a.n) return;
  const int i = warp;
  const double xi[3] = {a.qx[i], a.qx[(size_t)a.n + i], a.qx[2 * (size_t)a.n + i]};
  const int *Nc = a.prm.Nc;
  const int i1 = cell_coord(xi[0], a.prm.iLbNc[0], Nc[0]), i2 = cell_coord(xi[1], a.prm.iLbNc[1], Nc[1]),
            i3 = cell_coord(xi[2], a.prm.iLbNc[2], Nc[2]);
  const int sid = a.surf[i];
  double best = INFINITY;
  int bj = 0x7fffffff;
  for (int d3 = -1; d3 <= 1; d3++)
    for (int d2 = -1; d2 <= 1; d2++)
      for (int d1 = -1; d1 <= 1; d1++) {
        const int cj = imodulo(i1 + d1, Nc[0]) + Nc[0] * (imodulo(i2 + d2, Nc[1]) + Nc[1] * imodulo(i3 + d3, Nc[2]));
        for (int s = a.start[cj] + lane; s < a.start[cj + 1]; s += 32) {
          const int j = a.order[s];
          if (j / a.npc + 1 == sid) continue;
          const double xx = min_image(__dsub_rn(xi[0], a.sx[s]), a.prm.iLb[0], a.prm.Lb[0]);
          const double yy = min_image(__dsub_rn(xi[1], a.sx[(size_t)a.Np + s]), a.prm.iLb[1], a.prm.Lb[1]);
          const double zz = min_image(__dsub_rn(xi[2], a.sx[2 * (size_t)a.Np + s]), a.prm.iLb[2], a.prm.Lb[2]);
          const double rr = sqrt(norm2_exact(xx, yy, zz));
          if (rr < best || (rr == best && j < bj)) best = rr, bj = j;
        }
      }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ob = __shfl_xor_sync(FULL_MASK, best, o);
    const int oj = __shfl_xor_sync(FULL_MASK, bj, o);
    if (ob < best || (ob == best && oj < bj)) best = ob, bj = oj;
  }
  if (lane != 0) return;
  double x0[3] = {0, 0, 0}, dist0 = best;
  if (best <= 2.0 * a.epsDist) {  // ModRepulsion.F90:525-542
    const int jc = bj / a.npc, pt = bj - jc * a.npc, ilon0 = pt / a.nlat, ilat0 = pt - ilon0 * a.nlat;
    double th0 = a.th[ilat0], phi0 = a.phi[ilon0], xtar[3];
#pragma unroll
    for (int d = 0; d < 3; d++) {
      x0[d] = a.x[(size_t)d * a.Np + bj];
      const double xx = min_image(__dsub_rn(xi[d], x0[d]), a.prm.iLb[d], a.prm.Lb[d]);
      xtar[d] = x0[d] + xx;
    }
    const int m = 2 * a.nlat, nn = a.nlon;
    find_projection(a.spx + (size_t)12 * m * nn * jc, m, nn, xtar, th0, phi0, x0);
    const double e0 = xtar[0] - x0[0], e1 = xtar[1] - x0[1], e2 = xtar[2] - x0[2];
    dist0 = sqrt(e0 * e0 + e1 * e1 + e2 * e2);
  }
  a.dist[i] = dist0;
#pragma unroll
  for (int d = 0; d < 3; d++) a.x0[(size_t)d * a.n + i] = x0[d];
}

// device buffers in, device buffers out (capi.cu moves the host arrays)
int closest_cells(rbc3d_ctx *c, int n, const double *qx, const int *surf, double epsDist, double *dist, double *x0) {
  Cells &C = c->cells;
  if (n == 0) return RBC3D_OK;
  ClosestArgs a;
  a.prm = c->prm;
  a.n = n, a.Np = C.Np, a.npc = C.npc, a.nlat = C.nlat, a.nlon = C.nlon;
  a.qx = qx, a.surf = surf, a.start = C.cl.start.p, a.order = C.cl.order.p, a.sx = C.sx.p;
  a.x = C.x.p, a.spx = C.spx.p, a.th = C.th.p, a.phi = C.phi.p;
  a.epsDist = epsDist, a.dist = dist, a.x0 = x0;
  k_closest_cell<<<(n + 7) / 8, 256, 0, c->stream>>>(a);
  KERNEL_CHECK();
  c->launches++;
  return RBC3D_OK;
}

}  // namespace rbc3d
