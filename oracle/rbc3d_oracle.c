/*
 * rbc3d_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  See rbc3d_oracle.h.
 *
 * CPU restatement of RBC3D's Ewald boundary-integral operator for cell surfaces
 * (ModConf / ModHashTable / ModEwaldFunc / ModBasicMath / ModPolarPatch /
 * ModQuadRule / ModSpline / ModRbcSingInt / ModIntOnRbcs / ModPME / ModPFFTW).
 * PARITY UNPINNED (no reference tests or runnable reference exist; see header).
 *
 * Build: gcc -O2 -fopenmp -ffp-contract=off -fPIC -shared (see oracle/Makefile).
 * -ffp-contract=off mirrors the reference build (gfortran -O3 on baseline x86-64
 * emits no FMA), which matters for floor()/nint() based indexing.
 */
#include "rbc3d_oracle.h"

#include <complex.h>
#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef double _Complex cplx;

/* ModDataTypes.F90:19-21 */
static const double PI = 3.14159265358979323846;
#define TWO_PI (2 * PI)
#define I_2PI (1. / TWO_PI)
#define THRD (1.0 / 3)

static inline int imodulo(int a, int n) { /* Fortran modulo(): result has the sign of n */
  int r = a % n;
  return (r < 0) ? r + n : r;
}
static inline double fnint(double x) { return round(x); } /* Fortran nint(): half away from zero */

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ====================================================================== */
/* ModConf.F90:348-408  SetEwaldPrms                                        */
void orc_set_ewald_prms(const double Lb[3], double alpha, double eps, int P, int nranks, double *rc,
                        int Nb[3]) {
  double s = 1.;
  for (int iter = 1; iter <= 10; iter++) { /* :369-372 */
    s = 0.75 * sqrt(PI) * eps / (s * s * s + 1.5 * s + 0.75 / s);
    s = sqrt(-log(s));
  }
  double r = sqrt(alpha / PI) * s; /* :373 */
  double m = fmin(Lb[0] / 3.001, fmin(Lb[1] / 3.001, Lb[2] / 3.001));
  *rc = fmin(m, r); /* :376 */
  for (int d = 0; d < 3; d++) Nb[d] = 2 * (int)ceil(sqrt(-log(eps) / (PI * alpha)) * Lb[d]); /* :393 */
  if (Nb[2] < nranks * P) Nb[2] = nranks * P;                                              /* :394 */
  Nb[2] = (int)ceil((double)Nb[2] / nranks) * nranks;                                      /* :395 */
}

/* ModBasicMath.F90:332-348 MaskFunc_Exact */
double orc_mask_func_exact(double x) {
  double t = fabs(x);
  if (t < 0.01) return 1.;
  if (t > 0.99) return 0.;
  return exp(2 * (exp(-1. / t)) / (t - 1));
}

/* ModHashTable.F90:99-126 (single node: nodeZmin = 0, nodeZmax = Lb(3));
 * ModEwaldFunc.F90:96-106,150-160 (tables); ModBasicMath.F90:360-366 (mask table) */
void orc_params_init(orc_params *prm, const double Lb[3], double alpha, double eps, int P, double rc,
                     const int Nb[3]) {
  for (int d = 0; d < 3; d++) {
    prm->Lb[d] = Lb[d];
    prm->iLb[d] = 1. / Lb[d]; /* ModConf: iLb = 1./Lb */
    prm->Nb[d] = Nb[d];
  }
  prm->alpha = alpha;
  prm->eps = eps;
  prm->rc = rc;
  prm->P = P;
  for (int d = 0; d < 3; d++) {
    int nc = (int)floor(Lb[d] / rc);
    if (nc < 3) nc = 3;
    prm->Nc[d] = nc;
  }
  prm->iLbNc[0] = prm->iLb[0] * prm->Nc[0]; /* :122 */
  prm->iLbNc[1] = prm->iLb[1] * prm->Nc[1];
  prm->iLbNc[2] = prm->Nc[2] / (Lb[2] - 0.); /* :123, note the different rounding in z */
  const int N = ORC_NTAB;
  for (int i = 0; i <= N; i++) {
    double r_t = sqrt(PI / alpha) * (i * rc / N);
    prm->sl_c1[i] = erfc(r_t);
    prm->sl_c2[i] = 2 / sqrt(alpha) * exp(-(r_t * r_t));
    prm->dl_c1[i] =
        -8 / sqrt(PI) * (exp(-(r_t * r_t)) * (1.5 * r_t + r_t * r_t * r_t) + 0.75 * sqrt(PI) * erfc(r_t));
    double s = (double)i / N;
    prm->mask_tab[i] = orc_mask_func_exact(s);
  }
  prm->r_eps = 1.e-3 * sqrt(alpha / PI);
}

/* ModEwaldFunc.F90:25-52 */
void orc_ewald_coeff_sl_exact(double r, double alpha, double *A, double *B) {
  if (alpha <= 0) {
    *A = 1 / (r * r * r);
    *B = 1 / r;
    return;
  }
  double r_t = sqrt(PI / alpha) * r;
  if (r_t < 1.e-3) {
    *A = 0.;
    *B = 0.;
  } else {
    double c1 = erfc(r_t);
    double c2 = 2. / sqrt(alpha) * exp(-(r_t * r_t));
    double ir = 1. / r, ir2 = ir * ir;
    *A = c1 * ir * ir2 + c2 * ir2;
    *B = c1 * ir - c2;
  }
}

/* ModEwaldFunc.F90:59-79 */
void orc_ewald_coeff_dl_exact(double r, double alpha, double *A) {
  if (alpha <= 0) {
    *A = -6 / (r * r * r * r * r);
    return;
  }
  double r_t = sqrt(PI / alpha) * r;
  if (r_t < 1.e-3) {
    *A = 0.;
  } else {
    double a = exp(-(r_t * r_t)) * (1.5 * r_t + r_t * r_t * r_t) + 0.75 * sqrt(PI) * erfc(r_t);
    a = -8 / sqrt(PI) * a;
    *A = a / (r * r * r * r * r);
  }
}

/* ModEwaldFunc.F90:86-131 */
void orc_ewald_coeff_sl(const orc_params *prm, double r, double *A, double *B) {
  const int N = ORC_NTAB;
  if (r < prm->r_eps) {
    *A = 0.;
    *B = 0.;
    return;
  }
  double s = N * r / prm->rc;
  int i = (int)floor(s);
  if (i >= N) {
    *A = 0.;
    *B = 0.;
    return;
  }
  double c1 = prm->sl_c1[i] * (i + 1 - s) + prm->sl_c1[i + 1] * (s - i);
  double c2 = prm->sl_c2[i] * (i + 1 - s) + prm->sl_c2[i + 1] * (s - i);
  double ir = 1. / r, ir2 = ir * ir;
  *A = c1 * ir * ir2 + c2 * ir2;
  *B = c1 * ir - c2;
}

/* ModEwaldFunc.F90:141-178 */
void orc_ewald_coeff_dl(const orc_params *prm, double r, double *A) {
  const int N = ORC_NTAB;
  if (r < prm->r_eps) {
    *A = 0.;
    return;
  }
  double s = N * r / prm->rc;
  int i = (int)floor(s);
  if (i >= N) {
    *A = 0.;
    return;
  }
  double c1 = prm->dl_c1[i] * (i + 1 - s) + prm->dl_c1[i + 1] * (s - i);
  double r2 = r * r;
  *A = c1 / (r2 * r2 * r);
}

/* ModBasicMath.F90:351-379 MaskFunc (table lerp) */
double orc_mask_func(const orc_params *prm, double x) {
  const int N = ORC_NTAB;
  double s = fabs(x) * N;
  int i = (int)floor(s);
  if (i >= N) return 0.;
  return prm->mask_tab[i] * (i + 1 - s) + prm->mask_tab[i + 1] * (s - i);
}

/* ModBasicMath.F90:392-417 BsplineFunc */
void orc_bspline_func(double xc, int P, int *imin, double *w) {
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

/* ModPolarPatch.F90:249-257 */
double orc_dist_on_sphere(double th0, double phi0, double th1, double phi1) {
  double d = cos(th0 - th1) - sin(th0) * sin(th1) * (1. - cos(phi0 - phi1));
  d = fmin(1., fmax(-1., d));
  return acos(d);
}

/* ModQuadRule.F90:135-206 GauLeg (Numerical-Recipes Newton iteration, EPS 3e-14, <=10 its) */
void orc_gauleg(double x1, double x2, int n, double *x, double *w) {
  int m = (n + 1) / 2;
  double xm = 0.5 * (x2 + x1), xl = 0.5 * (x2 - x1);
  for (int j = 1; j <= m; j++) {
    double z = cos(PI * (j - 0.25) / (n + 0.5)), pp = 0, z1;
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

/* ModQuadRule.F90:220-263 GauLeg_Sinh (b0 = b*xl is what the reference does) */
void orc_gauleg_sinh(double xmin, double xmax, double a, double b, int n, double *x, double *w) {
  double xm = 0.5 * (xmin + xmax), xl = 0.5 * (xmax - xmin);
  double a0 = (a - xm) / xl, b0 = b * xl;
  double f1 = (1 + a0) / b0, f2 = (1 - a0) / b0;
  double u1 = log(f1 + sqrt(1 + f1 * f1)), u2 = log(f2 + sqrt(1 + f2 * f2)); /* MyAsinh :269-274 */
  double mu = 0.5 * (u1 + u2), eta = 0.5 * (u1 - u2);
  double s[256];
  orc_gauleg(-1., 1., n, s, w);
  for (int i = 0; i < n; i++) {
    x[i] = a0 + b0 * sinh(mu * s[i] - eta);
    w[i] = w[i] * b0 * mu * cosh(mu * s[i] - eta);
  }
  for (int i = 0; i < n; i++) {
    w[i] = xl * w[i];
    x[i] = xm + xl * x[i];
  }
}

/* ModPolarPatch.F90:99-148 PolarPatch_Build; thG/phiG(i,j) stored i (theta) fastest */
void orc_polar_patch_build(double th0, double phi0, int nth, const double *thL, int nphi,
                           const double *phiL, double *thG, double *phiG) {
  double sin_th0 = sin(th0), cos_th0 = cos(th0);
  for (int i = 0; i < nth; i++) {
    double st = sin(thL[i]), ct = cos(thL[i]);
    for (int j = 0; j < nphi; j++) {
      double sp = sin(phiL[j]), cp = cos(phiL[j]);
      double x0 = st * cp, x1 = st * sp, x2 = ct;
      /* x = matmul(A,x), A = [[c,0,s],[0,1,0],[-s,0,c]]; the zero entries are kept as in matmul */
      double y0 = cos_th0 * x0 + 0. * x1 + sin_th0 * x2;
      double y1 = 0. * x0 + 1. * x1 + 0. * x2;
      double y2 = -sin_th0 * x0 + 0. * x1 + cos_th0 * x2;
      y2 = fmax(-1.0, fmin(1.0, y2));
      thG[i + nth * j] = acos(y2);
      double ph = atan2(y1, y0) + phi0;
      phiG[i + nth * j] = ph - floor(ph * I_2PI) * TWO_PI;
    }
  }
}

/* ModPolarPatch.F90:217-242 PolarPatch_Map */
void orc_polar_patch_map(double th0, double phi0, double dth, double dphi, double *th, double *phi) {
  double s1[3] = {cos(th0) * cos(phi0), cos(th0) * sin(phi0), -sin(th0)};
  double s2[3] = {-sin(phi0), cos(phi0), 0.};
  double x0[3] = {sin(th0) * cos(phi0), sin(th0) * sin(phi0), cos(th0)};
  double xl[3] = {sin(dth) * cos(dphi), sin(dth) * sin(dphi), cos(dth)};
  double x[3];
  for (int d = 0; d < 3; d++) x[d] = xl[0] * s1[d] + xl[1] * s2[d] + xl[2] * x0[d];
  double nrm = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
  for (int d = 0; d < 3; d++) x[d] = x[d] / nrm;
  x[2] = fmax(-1., fmin(1., x[2]));
  *th = acos(x[2]);
  double p = atan2(x[1], x[0]);
  if (p < 0) p = p + TWO_PI;
  *phi = p;
}

/* ModPolarPatch.F90:162-206 PolarPatch_FindPoints; ijs = [n][2] (ilat, ilon) 1-based */
int orc_polar_patch_find_points(double th0, double phi0, double r0, int nth, const double *ths,
                                int nphi, const double *phis, int *ijs) {
  const double eps = 1.e-10;
  double h_phi = TWO_PI / nphi, ih_phi = 1.0 / h_phi;
  int n = 0;
  double cosr0 = cos(r0);
  for (int i = 1; i <= nth; i++) {
    if (ths[i - 1] <= th0 - r0) continue;
    if (ths[i - 1] >= th0 + r0) break;
    double dth = ths[i - 1] - th0;
    double dphi = 1. - (cos(dth) - cosr0) / ((sin(ths[i - 1]) + eps) * (sin(th0) + eps));
    dphi = fmax(-1., fmin(1., dphi));
    dphi = acos(dphi);
    int jmin, jmax;
    if (dphi > PI - eps) {
      jmin = 1;
      jmax = nphi;
    } else {
      jmin = (int)ceil((phi0 - dphi - phis[0]) * ih_phi) + 1;
      jmax = (int)floor((phi0 + dphi - phis[0]) * ih_phi) + 1;
    }
    for (int j = jmin; j <= jmax; j++) {
      ijs[2 * n] = i;
      ijs[2 * n + 1] = imodulo(j - 1, nphi) + 1;
      n++;
    }
  }
  return n;
}

/* ModPolarPatch.F90:27-76 RbcPolarPatch_Create; thG/phiG [ilon][ilat][iazm][irad] */
void orc_rbc_polar_patch_create(const orc_params *prm, int nlat, int nlon, const double *th,
                                const double *phi, double *radius_out, int *nrad_out, int *nazm_out,
                                double *thG, double *phiG, double *w) {
  double radius = PI / sqrt((double)nlat);
  double h = PI / nlat;
  int nrad = 2 * (int)fnint(radius / h);
  int nazm = 2 * nrad;
  *radius_out = radius;
  *nrad_out = nrad;
  *nazm_out = nazm;
  if (!thG) return; /* size query */
  double *thL = (double *)malloc(sizeof(double) * nrad), *phiL = (double *)malloc(sizeof(double) * nazm);
  orc_gauleg(0., radius, nrad, thL, w);
  for (int ir = 0; ir < nrad; ir++) {
    w[ir] = w[ir] * sin(thL[ir]) * (TWO_PI / nazm);
    w[ir] = w[ir] * orc_mask_func(prm, thL[ir] / radius);
  }
  for (int ia = 0; ia < nazm; ia++) phiL[ia] = ia * TWO_PI / nazm;
  for (int ilon = 0; ilon < nlon; ilon++)
    for (int ilat = 0; ilat < nlat; ilat++) {
      size_t off = ((size_t)ilon * nlat + ilat) * nazm * nrad;
      orc_polar_patch_build(th[ilat], phi[ilon], nrad, thL, nazm, phiL, thG + off, phiG + off);
    }
  free(thL);
  free(phiL);
}

/* ====================================================================== */
/* ModSpline.F90:150-191 Spline_Interp.  sp = [4][nvar][n][m] (m fastest); hx=2pi/m, hy=2pi/n */
void orc_spline_interp(const double *sp, int m, int n, int nvar, double x, double y, double *f) {
  double hx = TWO_PI / m, hy = TWO_PI / n;
  double ihx = 1. / hx, ihy = 1. / hy;
  int i1 = (int)floor(x * ihx), j1 = (int)floor(y * ihy);
  double s = x * ihx - i1, t = y * ihy - j1;
  i1 = imodulo(i1, m);
  j1 = imodulo(j1, n);
  int i2 = imodulo(i1 + 1, m), j2 = imodulo(j1 + 1, n);
  double cx[4] = {1 + s * s * (-3. + 2. * s), s * s * (3. - 2. * s), hx * s * (1. + s * (-2. + s)),
                  hx * s * s * (-1. + s)};
  double cy[4] = {1 + t * t * (-3. + 2. * t), t * t * (3. - 2. * t), hy * t * (1. + t * (-2. + t)),
                  hy * t * t * (-1. + t)};
  size_t plane = (size_t)m * n, arr = plane * nvar;
  const double *U = sp, *U1 = sp + arr, *U2 = sp + 2 * arr, *U12 = sp + 3 * arr;
  for (int l = 0; l < nvar; l++) {
    size_t o = plane * l;
    size_t a11 = o + i1 + (size_t)m * j1, a12 = o + i1 + (size_t)m * j2;
    size_t a21 = o + i2 + (size_t)m * j1, a22 = o + i2 + (size_t)m * j2;
    double u[4][4] = {{U[a11], U[a12], U2[a11], U2[a12]},
                      {U[a21], U[a22], U2[a21], U2[a22]},
                      {U1[a11], U1[a12], U12[a11], U12[a12]},
                      {U1[a21], U1[a22], U12[a21], U12[a22]}};
    double acc = 0.;
    for (int a = 0; a < 4; a++) {
      double tmp = 0.;
      for (int b = 0; b < 4; b++) tmp += u[a][b] * cy[b];
      acc += cx[a] * tmp;
    }
    f[l] = acc; /* kx = ky = 0 for sphere splines (ModSpline.F90:137-138) */
  }
}

/* LAPACK dposv('U',6,1) restated (unblocked Cholesky A = U^T U, then two triangular solves).
 * a is column-major 6x6 with the upper triangle filled.  Returns info (0 ok). */
static int chol_solve6(double *a, double *b) {
  const int n = 6;
#define A_(i, j) a[(i) + n * (j)]
  for (int j = 0; j < n; j++) {
    double ajj = A_(j, j);
    for (int k = 0; k < j; k++) ajj -= A_(k, j) * A_(k, j);
    if (!(ajj > 0.)) return j + 1;
    ajj = sqrt(ajj);
    A_(j, j) = ajj;
    for (int i = j + 1; i < n; i++) {
      double s = A_(j, i);
      for (int k = 0; k < j; k++) s -= A_(k, j) * A_(k, i);
      A_(j, i) = s / ajj;
    }
  }
  for (int i = 0; i < n; i++) { /* U^T y = b */
    double s = b[i];
    for (int k = 0; k < i; k++) s -= A_(k, i) * b[k];
    b[i] = s / A_(i, i);
  }
  for (int i = n - 1; i >= 0; i--) { /* U x = y */
    double s = b[i];
    for (int k = i + 1; k < n; k++) s -= A_(i, k) * b[k];
    b[i] = s / A_(i, i);
  }
#undef A_
  return 0;
}

/* ModBasicMath.F90:230-287 QuadFit_2D; xy = [npt][2]; c = (a0,a1,a2,a11,a12,a22) */
int orc_quadfit_2d(int npt, const double *xy, const double *f, double c[6]) {
  double lhs[36], rhs[6], u[6];
  memset(lhs, 0, sizeof lhs);
  memset(rhs, 0, sizeof rhs);
  for (int i = 0; i < npt; i++) {
    double xi = xy[2 * i], yi = xy[2 * i + 1];
    u[0] = 1.;
    u[1] = xi;
    u[2] = yi;
    u[3] = xi * xi;
    u[4] = xi * yi;
    u[5] = yi * yi;
    for (int ii = 0; ii < 6; ii++)
      for (int jj = ii; jj < 6; jj++) lhs[ii + 6 * jj] += u[ii] * u[jj];
    for (int ii = 0; ii < 6; ii++) rhs[ii] += u[ii] * f[i];
  }
  for (int ii = 1; ii < 6; ii++)
    for (int jj = 0; jj < ii; jj++) lhs[ii + 6 * jj] = lhs[jj + 6 * ii];
  int info = chol_solve6(lhs, rhs);
  for (int i = 0; i < 6; i++) c[i] = rhs[i];
  return info;
}

/* ModBasicMath.F90:300-326 Min_Quad_2D */
static void min_quad_2d(const double c[6], double xmin[2], double *fmin_) {
  double a0 = c[0], a1 = c[1], a2 = c[2], a11 = c[3], a12 = c[4], a22 = c[5];
  double l11 = 2. * a11, l12 = a12, l21 = a12, l22 = 2. * a22;
  double det = l11 * l22 - l12 * l21;
  double r1 = -a1, r2 = -a2;
  if (det > 0) {
    double idet = 1. / det;
    xmin[0] = idet * (l22 * r1 - l12 * r2);
    xmin[1] = idet * (-l21 * r1 + l11 * r2);
  } else {
    xmin[0] = 0.;
    xmin[1] = 0.;
  }
  if (fmin_)
    *fmin_ = a0 + a1 * xmin[0] + a2 * xmin[1] + a11 * xmin[0] * xmin[0] + a12 * xmin[0] * xmin[1] +
             a22 * xmin[1] * xmin[1];
}

/* ModSpline.F90:203-269 Spline_FindProjection (spln_x: nvar = 3) */
void orc_spline_find_projection(const double *sp, int m, int n, const double xtar[3], double *th0,
                                double *phi0, double x0[3]) {
  enum { nth = 2, nphi = 8, iterMax = 3 };
  double thPat_L[nth], phiPat_L[nphi], thPat[nth * nphi], phiPat[nth * nphi];
  double xyGq_L[(nth * nphi + 1) * 2], dist2Gq[nth * nphi + 1];
  double xx[3], c[6], xyMin_L[2], dist2MinEst, xMin[3], thMin, phiMin;
  orc_spline_interp(sp, m, n, 3, *th0, *phi0, x0);
  double hx = TWO_PI / m, hy = TWO_PI / n;
  double h = fmax(hx, hy);
  for (int iter = 1; iter <= iterMax; iter++) {
    for (int ith = 1; ith <= nth; ith++) thPat_L[ith - 1] = ith * h / nth;
    for (int iphi = 1; iphi <= nphi; iphi++) phiPat_L[iphi - 1] = (iphi - 1) * TWO_PI / nphi;
    orc_polar_patch_build(*th0, *phi0, nth, thPat_L, nphi, phiPat_L, thPat, phiPat);
    xyGq_L[0] = 0.;
    xyGq_L[1] = 0.;
    orc_spline_interp(sp, m, n, 3, *th0, *phi0, xx);
    for (int d = 0; d < 3; d++) xx[d] = xx[d] - xtar[d];
    dist2Gq[0] = xx[0] * xx[0] + xx[1] * xx[1] + xx[2] * xx[2];
    for (int iphi = 1; iphi <= nphi; iphi++)
      for (int ith = 1; ith <= nth; ith++) {
        int i = ith + nth * (iphi - 1);
        xyGq_L[2 * i] = thPat_L[ith - 1] * cos(phiPat_L[iphi - 1]);
        xyGq_L[2 * i + 1] = thPat_L[ith - 1] * sin(phiPat_L[iphi - 1]);
        orc_spline_interp(sp, m, n, 3, thPat[(ith - 1) + nth * (iphi - 1)],
                          phiPat[(ith - 1) + nth * (iphi - 1)], xx);
        for (int d = 0; d < 3; d++) xx[d] = xx[d] - xtar[d];
        dist2Gq[i] = xx[0] * xx[0] + xx[1] * xx[1] + xx[2] * xx[2];
      }
    orc_quadfit_2d(nth * nphi + 1, xyGq_L, dist2Gq, c);
    min_quad_2d(c, xyMin_L, &dist2MinEst);
    double thMin_L = sqrt(xyMin_L[0] * xyMin_L[0] + xyMin_L[1] * xyMin_L[1]);
    double phiMin_L = atan2(xyMin_L[1], xyMin_L[0]);
    orc_polar_patch_map(*th0, *phi0, thMin_L, phiMin_L, &thMin, &phiMin);
    orc_spline_interp(sp, m, n, 3, thMin, phiMin, xMin);
    for (int d = 0; d < 3; d++) xx[d] = xMin[d] - xtar[d];
    double dist2Min = xx[0] * xx[0] + xx[1] * xx[1] + xx[2] * xx[2];
    if (dist2Min > dist2Gq[0]) break; /* fail to converge */
    *th0 = thMin;
    *phi0 = phiMin;
    for (int d = 0; d < 3; d++) x0[d] = xMin[d];
    h = 0.5 * h;
  }
}

/* ====================================================================== */
/* ModHashTable.F90:67-91 HashTable_Index (SINGLE_NODE branch); returns 1-based indices */
void orc_hash_index(const orc_params *prm, const double x[3], int *i1, int *i2, int *i3) {
  *i1 = imodulo((int)floor(x[0] * prm->iLbNc[0]), prm->Nc[0]) + 1;
  *i2 = imodulo((int)floor(x[1] * prm->iLbNc[1]), prm->Nc[1]) + 1;
  *i3 = imodulo((int)floor(x[2] * prm->iLbNc[2]), prm->Nc[2]) + 1;
}

#define HOC(i1, i2, i3) hoc[(i1) + (size_t)n1 * ((i2) + (size_t)n2 * (i3))]

/* ModHashTable.F90:23-58 HashTable_Build.  hoc is (0:Nc1+1,0:Nc2+1,0:Nc3+1), entries are 0-based
 * point indices, -1 = empty (the reference uses 1-based with -1/0 as terminator). */
void orc_hash_build(const orc_params *prm, int n, const double *x, int *hoc, int *next) {
  const int *Nc = prm->Nc;
  int n1 = Nc[0] + 2, n2 = Nc[1] + 2, n3 = Nc[2] + 2;
  size_t tot = (size_t)n1 * n2 * n3;
  for (size_t i = 0; i < tot; i++) hoc[i] = -1;
  for (int i = 0; i < n; i++) next[i] = -1;
  for (int i = 0; i < n; i++) {
    double xi[3] = {x[i], x[n + i], x[2 * (size_t)n + i]};
    int i1, i2, i3;
    orc_hash_index(prm, xi, &i1, &i2, &i3);
    if (i3 >= 0 && i3 <= Nc[2] + 1) {
      next[i] = HOC(i1, i2, i3);
      HOC(i1, i2, i3) = i;
    }
  }
  for (int i3 = 0; i3 < n3; i3++)
    for (int i2 = 0; i2 < n2; i2++) {
      HOC(0, i2, i3) = HOC(Nc[0], i2, i3);
      HOC(Nc[0] + 1, i2, i3) = HOC(1, i2, i3);
    }
  for (int i3 = 0; i3 < n3; i3++)
    for (int i1 = 0; i1 < n1; i1++) {
      HOC(i1, 0, i3) = HOC(i1, Nc[1], i3);
      HOC(i1, Nc[1] + 1, i3) = HOC(i1, 1, i3);
    }
  for (int i2 = 0; i2 < n2; i2++) /* SINGLE_NODE */
    for (int i1 = 0; i1 < n1; i1++) {
      HOC(i1, i2, 0) = HOC(i1, i2, Nc[2]);
      HOC(i1, i2, Nc[2] + 1) = HOC(i1, i2, 1);
    }
}

void orc_cell_ids(const orc_params *prm, int n, const double *x, int *cid) {
#pragma omp parallel for
  for (int i = 0; i < n; i++) {
    double xi[3] = {x[i], x[n + i], x[2 * (size_t)n + i]};
    int i1, i2, i3;
    orc_hash_index(prm, xi, &i1, &i2, &i3);
    cid[i] = (i1 - 1) + prm->Nc[0] * ((i2 - 1) + prm->Nc[1] * (i3 - 1));
  }
}

static inline unsigned long long mix64(unsigned long long z) { /* splitmix64 finaliser */
  z += 0x9e3779b97f4a7c15ULL;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
  return z ^ (z >> 31);
}

/* In-range neighbour set of every target exactly as the pair loop of AddIntOnRbcs finds it
 * (ModIntOnRbcs.F90:62-74): 27-cell scan, minimum image by nint, rr = sqrt(sum(xx*xx)), keep rr <= rc. */
void orc_neighbor_signature(const orc_params *prm, int ns, const double *xs, int nt, const double *xt,
                            int *count, unsigned long long *sig) {
  const int *Nc = prm->Nc;
  int n1 = Nc[0] + 2, n2 = Nc[1] + 2, n3 = Nc[2] + 2;
  int *hoc = (int *)malloc(sizeof(int) * (size_t)n1 * n2 * n3);
  int *next = (int *)malloc(sizeof(int) * (size_t)(ns > 0 ? ns : 1));
  orc_hash_build(prm, ns, xs, hoc, next);
#pragma omp parallel for schedule(dynamic, 64)
  for (int i = 0; i < nt; i++) {
    double xi[3] = {xt[i], xt[nt + i], xt[2 * (size_t)nt + i]};
    int i1, i2, i3, cnt = 0;
    unsigned long long s = 0;
    orc_hash_index(prm, xi, &i1, &i2, &i3);
    for (int j1 = i1 - 1; j1 <= i1 + 1; j1++)
      for (int j2 = i2 - 1; j2 <= i2 + 1; j2++)
        for (int j3 = i3 - 1; j3 <= i3 + 1; j3++) {
          int j = HOC(j1, j2, j3);
          while (j >= 0) {
            double xx[3];
            for (int d = 0; d < 3; d++) {
              xx[d] = xs[(size_t)d * ns + j] - xi[d];
              xx[d] = xx[d] - fnint(xx[d] * prm->iLb[d]) * prm->Lb[d];
            }
            double rr = sqrt(xx[0] * xx[0] + xx[1] * xx[1] + xx[2] * xx[2]);
            if (!(rr > prm->rc)) {
              cnt++;
              s += mix64((unsigned long long)j);
            }
            j = next[j];
          }
        }
    count[i] = cnt;
    sig[i] = s;
  }
  free(hoc);
  free(next);
}

/* ====================================================================== */
/* ModRbcSingInt.F90:29-90 RBC_SingInt.  ilat0/ilon0 1-based; c2 already includes Bcoef. */
void orc_rbc_sing_int(const orc_params *prm, const orc_cells *C, double c1, double c2, int icell,
                      int ilat0, int ilon0, double dv[3]) {
  int m = 2 * C->nlat, n = C->nlon;
  size_t sp3 = (size_t)4 * 3 * m * n;
  const double *spx = C->spx + sp3 * icell, *spa3 = C->spa3 + sp3 * icell;
  const double *spF = C->spF ? C->spF + sp3 * icell : NULL, *spG = C->spG ? C->spG + sp3 * icell : NULL;
  double xi[3], xj[3], fj[3], gj[3], a3j[3], xx[3], EA, EB;
  dv[0] = dv[1] = dv[2] = 0.;
  orc_spline_interp(spx, m, n, 3, C->th[ilat0 - 1], C->phi[ilon0 - 1], xi);
  size_t off = ((size_t)(ilon0 - 1) * C->nlat + (ilat0 - 1)) * C->nazm * C->nrad;
  for (int irad = 0; irad < C->nrad; irad++)
    for (int iazm = 0; iazm < C->nazm; iazm++) {
      double th_j = C->thG[off + irad + (size_t)C->nrad * iazm];
      double phi_j = C->phiG[off + irad + (size_t)C->nrad * iazm];
      orc_spline_interp(spx, m, n, 3, th_j, phi_j, xj);
      for (int d = 0; d < 3; d++) xx[d] = xj[d] - xi[d];
      double rr = sqrt(xx[0] * xx[0] + xx[1] * xx[1] + xx[2] * xx[2]);
      if (rr >= prm->rc) continue;
      if (c1 != 0) {
        orc_spline_interp(spF, m, n, 3, th_j, phi_j, fj);
        for (int d = 0; d < 3; d++) fj[d] = C->patch_w[irad] * fj[d];
        orc_ewald_coeff_sl(prm, rr, &EA, &EB);
        double xf = xx[0] * fj[0] + xx[1] * fj[1] + xx[2] * fj[2];
        for (int d = 0; d < 3; d++) dv[d] = dv[d] + c1 * (EA * xx[d] * xf + EB * fj[d]);
      }
      if (c2 != 0) {
        orc_spline_interp(spG, m, n, 3, th_j, phi_j, gj);
        for (int d = 0; d < 3; d++) gj[d] = C->patch_w[irad] * gj[d];
        orc_spline_interp(spa3, m, n, 3, th_j, phi_j, a3j);
        orc_ewald_coeff_dl(prm, rr, &EA);
        double xg = xx[0] * gj[0] + xx[1] * gj[1] + xx[2] * gj[2];
        double xa = xx[0] * a3j[0] + xx[1] * a3j[1] + xx[2] * a3j[2];
        for (int d = 0; d < 3; d++) dv[d] = dv[d] + c2 * (EA * xx[d] * xg * xa);
      }
    }
}

/* ModRbcSingInt.F90:175-226 RBC_NearSingInt_Subtract */
static void nearsing_subtract(const orc_params *prm, const orc_cells *C, double c1, double c2, int icell,
                              const double xi[3], double th0, double phi0, double radPat, double dv[3]) {
  int nlat = C->nlat, nlon = C->nlon;
  size_t Np = (size_t)C->ncell * nlat * nlon, base = (size_t)icell * nlat * nlon;
  int *ijs = (int *)malloc(sizeof(int) * 2 * (size_t)nlat * nlon * 2);
  dv[0] = dv[1] = dv[2] = 0.;
  int npt = orc_polar_patch_find_points(th0, phi0, radPat, nlat, C->th, nlon, C->phi, ijs);
  for (int p = 0; p < npt; p++) {
    int ilat = ijs[2 * p], ilon = ijs[2 * p + 1];
    size_t q = base + (size_t)(ilon - 1) * nlat + (ilat - 1);
    double xx[3], fj[3], gj[3], a3j[3], EA, EB;
    for (int d = 0; d < 3; d++) xx[d] = C->x[d * Np + q] - xi[d];
    double rr = sqrt(xx[0] * xx[0] + xx[1] * xx[1] + xx[2] * xx[2]);
    if (rr > prm->rc) continue;
    double dth = orc_dist_on_sphere(th0, phi0, C->th[ilat - 1], C->phi[ilon - 1]);
    double mask = orc_mask_func(prm, dth / radPat);
    if (c1 != 0) {
      for (int d = 0; d < 3; d++) fj[d] = C->f[d * Np + q] * C->detj[q] * C->w[ilat - 1];
      orc_ewald_coeff_sl(prm, rr, &EA, &EB);
      double xf = xx[0] * fj[0] + xx[1] * fj[1] + xx[2] * fj[2];
      for (int d = 0; d < 3; d++) dv[d] = dv[d] - c1 * mask * (EA * xx[d] * xf + EB * fj[d]);
    }
    if (c2 != 0) {
      for (int d = 0; d < 3; d++) {
        gj[d] = C->g[d * Np + q] * C->detj[q] * C->w[ilat - 1];
        a3j[d] = C->a3[d * Np + q];
      }
      orc_ewald_coeff_dl(prm, rr, &EA);
      double xg = xx[0] * gj[0] + xx[1] * gj[1] + xx[2] * gj[2];
      double xa = xx[0] * a3j[0] + xx[1] * a3j[1] + xx[2] * a3j[2];
      for (int d = 0; d < 3; d++) dv[d] = dv[d] - c2 * mask * (EA * xx[d] * xg * xa);
    }
  }
  free(ijs);
}

/* ModRbcSingInt.F90:232-310 RBC_NearSingInt_ReAdd */
static void nearsing_readd(const orc_params *prm, const orc_cells *C, double c1, double c2, int icell,
                           const double xi[3], const double x0[3], double th0, double phi0, double radPat,
                           double dv[3]) {
  enum { nrad = 16, nazm = 32 };
  int m = 2 * C->nlat, n = C->nlon;
  size_t sp3 = (size_t)4 * 3 * m * n;
  const double *spx = C->spx + sp3 * icell, *spa3 = C->spa3 + sp3 * icell;
  const double *spF = C->spF ? C->spF + sp3 * icell : NULL, *spG = C->spG ? C->spG + sp3 * icell : NULL;
  double thPat[nrad], phiPat[nazm], wtPat[nrad], thPatG[nrad * nazm], phiPatG[nrad * nazm];
  double d0[3] = {xi[0] - x0[0], xi[1] - x0[1], xi[2] - x0[2]};
  double dist = sqrt(d0[0] * d0[0] + d0[1] * d0[1] + d0[2] * d0[2]);
  double sizePat = radPat * sqrt(C->area[icell] / (4 * PI));
  dv[0] = dv[1] = dv[2] = 0.;
  if (dist > DBL_MIN) { /* tiny(dist) */
    dist = dist * (radPat / sizePat);
    orc_gauleg_sinh(0., radPat, 0., dist, nrad, thPat, wtPat);
  } else {
    orc_gauleg(0., radPat, nrad, thPat, wtPat);
  }
  for (int ir = 0; ir < nrad; ir++) {
    wtPat[ir] = wtPat[ir] * sin(thPat[ir]) * (TWO_PI / nazm);
    wtPat[ir] = wtPat[ir] * orc_mask_func(prm, thPat[ir] / radPat);
  }
  for (int ia = 0; ia < nazm; ia++) phiPat[ia] = ia * TWO_PI / nazm;
  orc_polar_patch_build(th0, phi0, nrad, thPat, nazm, phiPat, thPatG, phiPatG);
  for (int irad = 0; irad < nrad; irad++)
    for (int iazm = 0; iazm < nazm; iazm++) {
      double th_j = thPatG[irad + nrad * iazm], phi_j = phiPatG[irad + nrad * iazm];
      double xj[3], xx[3], fj[3], gj[3], a3j[3], EA, EB;
      orc_spline_interp(spx, m, n, 3, th_j, phi_j, xj);
      for (int d = 0; d < 3; d++) xx[d] = xj[d] - xi[d];
      double rr = sqrt(xx[0] * xx[0] + xx[1] * xx[1] + xx[2] * xx[2]);
      if (rr > prm->rc) continue;
      if (c1 != 0) {
        orc_spline_interp(spF, m, n, 3, th_j, phi_j, fj);
        for (int d = 0; d < 3; d++) fj[d] = wtPat[irad] * fj[d];
        orc_ewald_coeff_sl(prm, rr, &EA, &EB);
        double xf = xx[0] * fj[0] + xx[1] * fj[1] + xx[2] * fj[2];
        for (int d = 0; d < 3; d++) dv[d] = dv[d] + c1 * (EA * xx[d] * xf + EB * fj[d]);
      }
      if (c2 != 0) {
        orc_spline_interp(spa3, m, n, 3, th_j, phi_j, a3j);
        orc_spline_interp(spG, m, n, 3, th_j, phi_j, gj);
        for (int d = 0; d < 3; d++) gj[d] = wtPat[irad] * gj[d];
        orc_ewald_coeff_dl(prm, rr, &EA);
        double xg = xx[0] * gj[0] + xx[1] * gj[1] + xx[2] * gj[2];
        double xa = xx[0] * a3j[0] + xx[1] * a3j[1] + xx[2] * a3j[2];
        for (int d = 0; d < 3; d++) dv[d] = dv[d] + c2 * (EA * xx[d] * xg * xa);
      }
    }
}

/* ModRbcSingInt.F90:103-167 RBC_NearSingInt.  c2 already includes Bcoef of the source cell. */
void orc_rbc_nearsing_int(const orc_params *prm, const orc_cells *C, double c1, double c2, int icell,
                          const double xi[3], const double x0[3], double th0, double phi0, double dv[3]) {
  int m = 2 * C->nlat, n = C->nlon;
  size_t sp3 = (size_t)4 * 3 * m * n, sp1 = (size_t)4 * 1 * m * n;
  double a30[3], dvtmp[3], dv0[3], dv1[3];
  double radPat = C->patch_radius;
  orc_spline_interp(C->spa3 + sp3 * icell, m, n, 3, th0, phi0, a30);
  double dist = a30[0] * (xi[0] - x0[0]) + a30[1] * (xi[1] - x0[1]) + a30[2] * (xi[2] - x0[2]);
  dv[0] = dv[1] = dv[2] = 0.;
  if (dist > 2 * C->meshSize[icell]) return; /* :122-125 (signed distance) */
  double sizePat = radPat * sqrt(C->area[icell] / (4 * PI));
  double dist1 = copysign(0.01 * sizePat, dist); /* sign(0.01*sizePat, dist) */
  nearsing_subtract(prm, C, c1, c2, icell, xi, th0, phi0, radPat, dvtmp);
  for (int d = 0; d < 3; d++) dv[d] = dv[d] + dvtmp[d];
  if (fabs(dist) >= fabs(dist1)) {
    nearsing_readd(prm, C, c1, c2, icell, xi, x0, th0, phi0, radPat, dvtmp);
    for (int d = 0; d < 3; d++) dv[d] = dv[d] + dvtmp[d];
  } else {
    double xi1[3], xi0[3];
    for (int d = 0; d < 3; d++) xi1[d] = x0[d] + dist1 * a30[d];
    nearsing_readd(prm, C, c1, c2, icell, xi1, x0, th0, phi0, radPat, dv1);
    for (int d = 0; d < 3; d++) xi0[d] = x0[d];
    nearsing_readd(prm, C, c1, c2, icell, xi0, x0, th0, phi0, radPat, dv0);
    if (c2 != 0) {
      double detJ0[1], g0[3];
      orc_spline_interp(C->spdetj + sp1 * icell, m, n, 1, th0, phi0, detJ0);
      orc_spline_interp(C->spG + sp3 * icell, m, n, 3, th0, phi0, g0);
      for (int d = 0; d < 3; d++) g0[d] = g0[d] / detJ0[0];
      if (dist > 0)
        for (int d = 0; d < 3; d++) dv0[d] = dv0[d] + c2 * 4 * PI * g0[d];
      else
        for (int d = 0; d < 3; d++) dv0[d] = dv0[d] - c2 * 4 * PI * g0[d];
    }
    for (int d = 0; d < 3; d++) {
      dvtmp[d] = dv0[d] + dist / dist1 * (dv1[d] - dv0[d]);
      dv[d] = dv[d] + dvtmp[d];
    }
  }
}

/* ModIntOnRbcs.F90:162-201 AddLinearInt */
static void add_linear_int(const orc_params *prm, const orc_cells *C, double c2, const orc_targets *tl,
                           double *v) {
  if (c2 == 0) return;
  int nlat = C->nlat, nlon = C->nlon;
  size_t Np = (size_t)C->ncell * nlat * nlon;
  double xvint[3] = {0., 0., 0.};
  for (int ic = 0; ic < C->ncell; ic++)
    for (int ilon = 0; ilon < nlon; ilon++)
      for (int ilat = 0; ilat < nlat; ilat++) {
        size_t q = (size_t)ic * nlat * nlon + (size_t)ilon * nlat + ilat;
        double ds = C->w[ilat] * C->detj[q];
        double vn = C->g[q] * C->a3[q] + C->g[Np + q] * C->a3[Np + q] + C->g[2 * Np + q] * C->a3[2 * Np + q];
        for (int d = 0; d < 3; d++) xvint[d] = xvint[d] + C->Bcoef[ic] * C->x[d * Np + q] * vn * ds;
      }
  double piLb = prm->iLb[0] * prm->iLb[1] * prm->iLb[2];
  for (int d = 0; d < 3; d++) xvint[d] = -8 * PI * piLb * xvint[d];
  size_t n = tl->n;
  for (size_t i = 0; i < n; i++)
    if (tl->active[i])
      for (int d = 0; d < 3; d++) v[d * n + i] = v[d * n + i] + c2 * xvint[d] / tl->Acoef[i];
}

/* ModIntOnRbcs.F90:25-158 AddIntOnRbcs (single-node cell list).  v is SoA(3,n), accumulated into. */
void orc_add_int_on_rbcs(const orc_params *prm, const orc_cells *C, double c1, double c2,
                         const orc_targets *tl, double *v, int flags) {
  if (C->ncell == 0) return;
  const int nlat = C->nlat, nlon = C->nlon, npc = nlat * nlon;
  const size_t Np = (size_t)C->ncell * npc, nt = tl->n;
  const int *Nc = prm->Nc;
  const int n1 = Nc[0] + 2, n2 = Nc[1] + 2, n3 = Nc[2] + 2;
  int *hoc = (int *)malloc(sizeof(int) * (size_t)n1 * n2 * n3);
  int *next = (int *)malloc(sizeof(int) * Np);
  orc_hash_build(prm, (int)Np, C->x, hoc, next); /* SourceList_UpdateCoord, ModSourceList.F90:149 */
  /* SourceList_UpdateDensity (ModSourceList.F90:169-185): f,g <- rbc%f,g * (detj*w) */
  double *sf = NULL, *sg = NULL;
  if (c1 != 0) sf = (double *)malloc(sizeof(double) * 3 * Np);
  if (c2 != 0) sg = (double *)malloc(sizeof(double) * 3 * Np);
#pragma omp parallel for
  for (size_t q = 0; q < Np; q++) {
    double dS = C->detj[q] * C->w[q % nlat];
    for (int d = 0; d < 3; d++) {
      if (sf) sf[d * Np + q] = C->f[d * Np + q] * dS;
      if (sg) sg[d * Np + q] = C->g[d * Np + q] * dS;
    }
  }
  int m = 2 * nlat, n = nlon;
  size_t sp3 = (size_t)4 * 3 * m * n;

#pragma omp parallel for schedule(dynamic, 16)
  for (size_t i = 0; i < nt; i++) {
    if (!tl->active[i]) continue;
    double xi[3] = {tl->x[i], tl->x[nt + i], tl->x[2 * nt + i]};
    double vi[3] = {0., 0., 0.};
    int surfId_i = tl->indx[i];
    int irbc_i = surfId_i - 1 + 1; /* rbcs(1)%Id = 1 */
    double th_i = 0, phi_i = 0;
    int is_cell = (irbc_i >= 1 && irbc_i <= C->ncell);
    if (is_cell) {
      th_i = C->th[tl->indx[nt + i] - 1];
      phi_i = C->phi[tl->indx[2 * nt + i] - 1];
    }
    /* NbrRbcList (ModRbcSingInt.F90:319-364) */
    int nbrN = 0, nbr_indx[ORC_NBR_MAX][3];
    double nbr_dist[ORC_NBR_MAX];
    int i1, i2, i3;
    orc_hash_index(prm, xi, &i1, &i2, &i3);
    if (!(flags & ORC_FLAG_NO_PAIRS) || !(flags & ORC_FLAG_NO_NEARSING))
      for (int j1 = (i1 - 1 > 0 ? i1 - 1 : 0); j1 <= (i1 + 1 < Nc[0] + 1 ? i1 + 1 : Nc[0] + 1); j1++)
        for (int j2 = (i2 - 1 > 0 ? i2 - 1 : 0); j2 <= (i2 + 1 < Nc[1] + 1 ? i2 + 1 : Nc[1] + 1); j2++)
          for (int j3 = (i3 - 1 > 0 ? i3 - 1 : 0); j3 <= (i3 + 1 < Nc[2] + 1 ? i3 + 1 : Nc[2] + 1); j3++) {
            int j = HOC(j1, j2, j3);
            while (j >= 0) {
              double xx[3];
              for (int d = 0; d < 3; d++) {
                xx[d] = C->x[d * Np + j] - xi[d];
                xx[d] = xx[d] - fnint(xx[d] * prm->iLb[d]) * prm->Lb[d];
              }
              double rr = sqrt(xx[0] * xx[0] + xx[1] * xx[1] + xx[2] * xx[2]);
              if (rr > prm->rc) {
                j = next[j];
                continue;
              }
              int jc = j / npc, jr = j % npc;
              int ilat_j = jr % nlat + 1, ilon_j = jr / nlat + 1;
              int surfId_j = jc + 1;
              double mask;
              if (surfId_i == surfId_j) {
                double dth = orc_dist_on_sphere(th_i, phi_i, C->th[ilat_j - 1], C->phi[ilon_j - 1]);
                mask = orc_mask_func(prm, dth / C->patch_radius);
              } else {
                mask = 0.;
                /* NbrRbcList_Insert */
                int found = 0;
                for (int p = 0; p < nbrN; p++)
                  if (nbr_indx[p][0] == surfId_j) {
                    if (rr < nbr_dist[p]) {
                      nbr_indx[p][1] = ilat_j;
                      nbr_indx[p][2] = ilon_j;
                      nbr_dist[p] = rr;
                    }
                    found = 1;
                    break;
                  }
                if (!found) {
                  if (nbrN >= ORC_NBR_MAX) {
                    fprintf(stderr, "orc: NbrRbcList overflow (reference would write out of bounds)\n");
                    abort();
                  }
                  nbr_indx[nbrN][0] = surfId_j;
                  nbr_indx[nbrN][1] = ilat_j;
                  nbr_indx[nbrN][2] = ilon_j;
                  nbr_dist[nbrN] = rr;
                  nbrN++;
                }
              }
              if (!(flags & ORC_FLAG_NO_PAIRS)) {
                if (c1 != 0) {
                  double EA, EB;
                  orc_ewald_coeff_sl(prm, rr, &EA, &EB);
                  double fj[3] = {sf[j], sf[Np + j], sf[2 * Np + j]};
                  double xf = xx[0] * fj[0] + xx[1] * fj[1] + xx[2] * fj[2];
                  for (int d = 0; d < 3; d++)
                    vi[d] = vi[d] + (1. - mask) * c1 / tl->Acoef[i] * (EA * xx[d] * xf + EB * fj[d]);
                }
                if (c2 != 0) {
                  double EA;
                  orc_ewald_coeff_dl(prm, rr, &EA);
                  double xg = xx[0] * sg[j] + xx[1] * sg[Np + j] + xx[2] * sg[2 * Np + j];
                  double xa = xx[0] * C->a3[j] + xx[1] * C->a3[Np + j] + xx[2] * C->a3[2 * Np + j];
                  for (int d = 0; d < 3; d++)
                    vi[d] = vi[d] + (1. - mask) * c2 * C->Bcoef[jc] / tl->Acoef[i] * (EA * xx[d] * xg * xa);
                }
              }
              j = next[j];
            }
          }
    /* Singular integration :114-121 */
    if (is_cell && !(flags & ORC_FLAG_NO_SING)) {
      double dv[3];
      double c2Mod = c2 * C->Bcoef[irbc_i - 1];
      orc_rbc_sing_int(prm, C, c1, c2Mod, irbc_i - 1, tl->indx[nt + i], tl->indx[2 * nt + i], dv);
      for (int d = 0; d < 3; d++) vi[d] = vi[d] + dv[d] / tl->Acoef[i];
    }
    /* Near-singular integration :123-148 */
    if (!(flags & ORC_FLAG_NO_NEARSING))
      for (int p = 0; p < nbrN; p++) {
        int jc = nbr_indx[p][0] - 1, ilat0 = nbr_indx[p][1], ilon0 = nbr_indx[p][2];
        double th0 = C->th[ilat0 - 1], phi0 = C->phi[ilon0 - 1];
        size_t q = (size_t)jc * npc + (size_t)(ilon0 - 1) * nlat + (ilat0 - 1);
        double x0[3] = {C->x[q], C->x[Np + q], C->x[2 * Np + q]};
        double xt[3], xx[3], dv[3];
        for (int d = 0; d < 3; d++) {
          xx[d] = tl->x[d * nt + i] - x0[d];
          xx[d] = xx[d] - fnint(xx[d] * prm->iLb[d]) * prm->Lb[d];
          xt[d] = x0[d] + xx[d];
        }
        orc_spline_find_projection(C->spx + sp3 * jc, m, n, xt, &th0, &phi0, x0);
        double c2Mod = c2 * C->Bcoef[jc];
        orc_rbc_nearsing_int(prm, C, c1, c2Mod, jc, xt, x0, th0, phi0, dv);
        for (int d = 0; d < 3; d++) vi[d] = vi[d] + dv[d] / tl->Acoef[i];
      }
    for (int d = 0; d < 3; d++) v[d * nt + i] += vi[d];
  }
  if (!(flags & ORC_FLAG_NO_LINEAR)) add_linear_int(prm, C, c2, tl, v);
  free(hoc);
  free(next);
  free(sf);
  free(sg);
}

/* ====================================================================== */
/* FFT: plain mixed-radix recursive complex FFT (FFTW is not available; the reference uses FFTW
 * 3.3.10 with FFTW_ESTIMATE, ModPFFTW.F90:110-143).  Unnormalised, sign = exponent sign. */
typedef struct {
  int n;
  cplx *tw; /* tw[j] = exp(+2 pi i j / n) */
} fft_tab;

static void fft_tab_init(fft_tab *t, int n) {
  t->n = n;
  t->tw = (cplx *)malloc(sizeof(cplx) * n);
  for (int j = 0; j < n; j++) {
    double a = TWO_PI * j / n;
    t->tw[j] = cos(a) + I * sin(a);
  }
}

static void fft_rec(const fft_tab *T, int n, const cplx *in, int is, cplx *out, int sign, cplx *scratch) {
  if (n == 1) {
    out[0] = in[0];
    return;
  }
  int p = 2;
  while (n % p) p++;
  int m = n / p;
  for (int r = 0; r < p; r++) fft_rec(T, m, in + (size_t)r * is, is * p, out + (size_t)r * m, sign, scratch);
  int tws = T->n / n;
  if (p == 2) {
    for (int k = 0; k < m; k++) {
      cplx w = T->tw[(size_t)k * tws];
      if (sign < 0) w = conj(w);
      cplx a = out[k], b = w * out[k + m];
      out[k] = a + b;
      out[k + m] = a - b;
    }
    return;
  }
  cplx *tt = scratch + p; /* scratch holds 2p entries: [0,p) outputs, [p,2p) twiddled inputs */
  for (int k = 0; k < m; k++) {
    for (int r = 0; r < p; r++) {
      cplx w = T->tw[((size_t)r * k * tws) % T->n];
      if (sign < 0) w = conj(w);
      tt[r] = w * out[k + (size_t)r * m];
    }
    for (int q = 0; q < p; q++) {
      cplx acc = tt[0];
      for (int r = 1; r < p; r++) {
        cplx w = T->tw[((size_t)((r * q) % p) * (T->n / p)) % T->n];
        if (sign < 0) w = conj(w);
        acc += w * tt[r];
      }
      scratch[q] = acc;
    }
    for (int q = 0; q < p; q++) out[k + (size_t)q * m] = scratch[q];
  }
}

/* out-of-place 1-D transform of a contiguous line */
static void fft_line(const fft_tab *T, const cplx *in, cplx *out, int sign, cplx *scratch) {
  fft_rec(T, T->n, in, 1, out, sign, scratch);
}

struct orc_pme {
  orc_params prm;
  int Nx, Ny, Nz, Nxh;
  fft_tab Tx, Ty, Tz;
  double *ff;  /* [3][Nz][Ny][Nx] */
  double *tt;  /* [9][Nz][Ny][Nx], component c = ii + 3*jj (tt(:,:,:,ii,jj)) */
  double *vv;  /* [3][Nz][Ny][Nx] */
  cplx *ffC;   /* [3][Nz][Ny][Nxh] */
  cplx *ttC;   /* [9][Nz][Ny][Nxh] */
  cplx *vvC;   /* [3][Nz][Ny][Nxh] */
  double *bb;  /* [Nxh][Ny][Nz] */
  int flag_sing_lay, flag_doub_lay;
};

/* ModPFFTW.F90:148-166 Get_That: 2-D r2c over (x,y) per z plane (FFTW forward, sign -1), then 1-D
 * complex transform along z with sign +1 (:110-111).  out[(k*Ny+j)*Nxh+i] = rhoTh(k,j,i). */
static void fft3_forward(const orc_pme *P, const double *in, cplx *out) {
  int Nx = P->Nx, Ny = P->Ny, Nz = P->Nz, Nxh = P->Nxh;
#pragma omp parallel
  {
    int nmax = Nx > Ny ? (Nx > Nz ? Nx : Nz) : (Ny > Nz ? Ny : Nz);
    cplx *a = (cplx *)malloc(sizeof(cplx) * nmax), *b = (cplx *)malloc(sizeof(cplx) * nmax);
    cplx *scr = (cplx *)malloc(sizeof(cplx) * (2 * nmax + 64));
#pragma omp for collapse(2)
    for (int k = 0; k < Nz; k++)
      for (int j = 0; j < Ny; j++) {
        const double *row = in + ((size_t)k * Ny + j) * Nx;
        for (int i = 0; i < Nx; i++) a[i] = row[i];
        fft_line(&P->Tx, a, b, -1, scr);
        cplx *o = out + ((size_t)k * Ny + j) * Nxh;
        for (int i = 0; i < Nxh; i++) o[i] = b[i];
      }
#pragma omp for collapse(2)
    for (int k = 0; k < Nz; k++)
      for (int i = 0; i < Nxh; i++) {
        for (int j = 0; j < Ny; j++) a[j] = out[((size_t)k * Ny + j) * Nxh + i];
        fft_line(&P->Ty, a, b, -1, scr);
        for (int j = 0; j < Ny; j++) out[((size_t)k * Ny + j) * Nxh + i] = b[j];
      }
#pragma omp for collapse(2)
    for (int j = 0; j < Ny; j++)
      for (int i = 0; i < Nxh; i++) {
        for (int k = 0; k < Nz; k++) a[k] = out[((size_t)k * Ny + j) * Nxh + i];
        fft_line(&P->Tz, a, b, +1, scr);
        for (int k = 0; k < Nz; k++) out[((size_t)k * Ny + j) * Nxh + i] = b[k];
      }
    free(a);
    free(b);
    free(scr);
  }
}

/* ModPFFTW.F90:169-185 Get_R: 1-D along z with sign -1 (:113-114), then FFTW 2-D c2r over (x,y):
 * complex backward (+1) along y, then c2r along x, which reads only the real parts of the x = 0 and
 * x = Nx/2 bins (FFTW's halfcomplex convention; DERIVED, see SURVEY.md A.3 caveat).  Unnormalised. */
static void fft3_backward(const orc_pme *P, const cplx *in, double *out) {
  int Nx = P->Nx, Ny = P->Ny, Nz = P->Nz, Nxh = P->Nxh;
  cplx *w = (cplx *)malloc(sizeof(cplx) * (size_t)Nz * Ny * Nxh);
  memcpy(w, in, sizeof(cplx) * (size_t)Nz * Ny * Nxh);
#pragma omp parallel
  {
    int nmax = Nx > Ny ? (Nx > Nz ? Nx : Nz) : (Ny > Nz ? Ny : Nz);
    cplx *a = (cplx *)malloc(sizeof(cplx) * nmax), *b = (cplx *)malloc(sizeof(cplx) * nmax);
    cplx *scr = (cplx *)malloc(sizeof(cplx) * (2 * nmax + 64));
#pragma omp for collapse(2)
    for (int j = 0; j < Ny; j++)
      for (int i = 0; i < Nxh; i++) {
        for (int k = 0; k < Nz; k++) a[k] = w[((size_t)k * Ny + j) * Nxh + i];
        fft_line(&P->Tz, a, b, -1, scr);
        for (int k = 0; k < Nz; k++) w[((size_t)k * Ny + j) * Nxh + i] = b[k];
      }
#pragma omp for collapse(2)
    for (int k = 0; k < Nz; k++)
      for (int i = 0; i < Nxh; i++) {
        for (int j = 0; j < Ny; j++) a[j] = w[((size_t)k * Ny + j) * Nxh + i];
        fft_line(&P->Ty, a, b, +1, scr);
        for (int j = 0; j < Ny; j++) w[((size_t)k * Ny + j) * Nxh + i] = b[j];
      }
#pragma omp for collapse(2)
    for (int k = 0; k < Nz; k++)
      for (int j = 0; j < Ny; j++) {
        const cplx *r = w + ((size_t)k * Ny + j) * Nxh;
        a[0] = creal(r[0]);
        for (int i = 1; i < Nx / 2; i++) {
          a[i] = r[i];
          a[Nx - i] = conj(r[i]);
        }
        a[Nx / 2] = creal(r[Nx / 2]);
        fft_line(&P->Tx, a, b, +1, scr);
        double *o = out + ((size_t)k * Ny + j) * Nx;
        for (int i = 0; i < Nx; i++) o[i] = creal(b[i]);
      }
    free(a);
    free(b);
    free(scr);
  }
  free(w);
}

/* ModPME.F90:252-338 PME_Init */
orc_pme *orc_pme_init(const orc_params *prm) {
  orc_pme *P = (orc_pme *)calloc(1, sizeof(orc_pme));
  P->prm = *prm;
  int Nx = prm->Nb[0], Ny = prm->Nb[1], Nz = prm->Nb[2], Nxh = Nx / 2 + 1, PB = prm->P;
  P->Nx = Nx;
  P->Ny = Ny;
  P->Nz = Nz;
  P->Nxh = Nxh;
  fft_tab_init(&P->Tx, Nx);
  fft_tab_init(&P->Ty, Ny);
  fft_tab_init(&P->Tz, Nz);
  size_t G = (size_t)Nx * Ny * Nz, M = (size_t)Nxh * Ny * Nz;
  P->ff = (double *)calloc(3 * G, sizeof(double));
  P->tt = (double *)calloc(9 * G, sizeof(double));
  P->vv = (double *)calloc(3 * G, sizeof(double));
  P->ffC = (cplx *)calloc(3 * M, sizeof(cplx));
  P->ttC = (cplx *)calloc(9 * M, sizeof(cplx));
  P->vvC = (cplx *)calloc(3 * M, sizeof(cplx));
  P->bb = (double *)malloc(sizeof(double) * M);
  /* B-factor :310-336; bb(k,j,i), k (z) fastest */
  for (size_t q = 0; q < M; q++) P->bb[q] = 1.;
  P->bb[0] = 0.;
  double MP[64];
  int imin;
  orc_bspline_func(PB + DBL_EPSILON, PB, &imin, MP);
  const int Nbd[3] = {Nx, Ny, Nz};
  const int cnt[3] = {Nxh, Ny, Nz};
  double *bd[3];
  for (int ii = 0; ii < 3; ii++) {
    bd[ii] = (double *)malloc(sizeof(double) * cnt[ii]);
    for (int k = 0; k < cnt[ii]; k++) {
      cplx b = 0.;
      for (int mm = 0; mm <= PB - 2; mm++) {
        double ang = TWO_PI * k * mm / (double)Nbd[ii];
        b = b + MP[mm] * (cos(ang) + I * sin(ang));
      }
      double ang = TWO_PI * k * (PB - 1.) / (double)Nbd[ii];
      b = (cos(ang) + I * sin(ang)) / b;
      double b2 = cabs(b);
      bd[ii][k] = b2 * b2;
    }
  }
  /* bb = ((1*b1(i))*b2(j))*b3(k) in the order of the select-case loop :326-334 */
  for (int i = 0; i < Nxh; i++)
    for (int j = 0; j < Ny; j++)
      for (int kz = 0; kz < Nz; kz++) {
        size_t q = ((size_t)i * Ny + j) * Nz + kz;
        P->bb[q] = ((P->bb[q] * bd[0][i]) * bd[1][j]) * bd[2][kz];
      }
  for (int ii = 0; ii < 3; ii++) free(bd[ii]);
  return P;
}

void orc_pme_finalize(orc_pme *P) {
  if (!P) return;
  free(P->ff);
  free(P->tt);
  free(P->vv);
  free(P->ffC);
  free(P->ttC);
  free(P->vvC);
  free(P->bb);
  free(P->Tx.tw);
  free(P->Ty.tw);
  free(P->Tz.tw);
  free(P);
}

const double *orc_pme_vv(orc_pme *P) { return P->vv; }
const double *orc_pme_bb(orc_pme *P) { return P->bb; }

/* ModPME.F90:58-133 PME_Distrib_Source + :405-443 Distrib_Source, for a flat list of point sources.
 * For wall sources the caller passes centroid x and f = mean(f_vert)*area (:119-122).
 * accumulate = 0 zeroes ff, tt first (:75-76); 1 keeps them (second source list of the same call).
 * Threads own z-slabs of the mesh exactly like the reference's MPI ranks (:428-429). */
void orc_pme_distrib_source(orc_pme *P, double c1, double c2, int n, const double *x, const double *f,
                            const double *g, const double *a3, const double *Bcoef, int accumulate) {
  const orc_params *prm = &P->prm;
  int Nx = P->Nx, Ny = P->Ny, Nz = P->Nz, PB = prm->P;
  size_t G = (size_t)Nx * Ny * Nz;
  P->flag_sing_lay = (fabs(c1) > 1.e-10);
  P->flag_doub_lay = (fabs(c2) > 1.e-10);
  if (!accumulate) {
    memset(P->ff, 0, sizeof(double) * 3 * G);
    memset(P->tt, 0, sizeof(double) * 9 * G);
  }
  double ih[3];
  for (int d = 0; d < 3; d++) ih[d] = prm->Nb[d] / prm->Lb[d];
  int fs = P->flag_sing_lay && f, fd = P->flag_doub_lay && g;
#pragma omp parallel
  {
    int nth = 1, tid = 0;
#ifdef _OPENMP
    nth = omp_get_num_threads();
    tid = omp_get_thread_num();
#endif
    int kb = (int)((long long)Nz * tid / nth), ke = (int)((long long)Nz * (tid + 1) / nth) - 1;
    double wx[64], wy[64], wz[64];
    for (int p = 0; p < n; p++) {
      int imin, jmin, kmin;
      orc_bspline_func(x[2 * (size_t)n + p] * ih[2], PB, &kmin, wz);
      int any = 0;
      for (int k0 = 1; k0 <= PB; k0++) {
        int k = imodulo(kmin + k0 - 1, Nz);
        if (k >= kb && k <= ke) any = 1;
      }
      if (!any) continue;
      orc_bspline_func(x[p] * ih[0], PB, &imin, wx);
      orc_bspline_func(x[(size_t)n + p] * ih[1], PB, &jmin, wy);
      double ft[3] = {0, 0, 0}, tm[9];
      if (fs)
        for (int d = 0; d < 3; d++) ft[d] = f[(size_t)d * n + p];
      if (fd) {
        double a3t[3];
        for (int d = 0; d < 3; d++) a3t[d] = a3[(size_t)d * n + p] * Bcoef[p];
        for (int ii = 0; ii < 3; ii++)
          for (int jj = 0; jj < 3; jj++) tm[ii + 3 * jj] = g[(size_t)ii * n + p] * a3t[jj];
      }
      for (int k0 = 1; k0 <= PB; k0++) {
        int k = imodulo(kmin + k0 - 1, Nz);
        if (k < kb || k > ke) continue;
        for (int j0 = 1; j0 <= PB; j0++) {
          int j = imodulo(jmin + j0 - 1, Ny);
          for (int i0 = 1; i0 <= PB; i0++) {
            int i = imodulo(imin + i0 - 1, Nx);
            double wxyz = wx[i0 - 1] * wy[j0 - 1] * wz[k0 - 1];
            size_t q = ((size_t)k * Ny + j) * Nx + i;
            if (fs)
              for (int d = 0; d < 3; d++) P->ff[d * G + q] += (c1 * wxyz) * ft[d];
            if (fd)
              for (int c = 0; c < 9; c++) P->tt[c * G + q] += (c2 * wxyz) * tm[c];
          }
        }
      }
    }
  }
}

/* ModPME.F90:137-222 PME_Transform */
void orc_pme_transform(orc_pme *P) {
  const orc_params *prm = &P->prm;
  int Nx = P->Nx, Ny = P->Ny, Nz = P->Nz, Nxh = P->Nxh;
  size_t G = (size_t)Nx * Ny * Nz, M = (size_t)Nxh * Ny * Nz;
  if (P->flag_sing_lay)
    for (int ii = 0; ii < 3; ii++) fft3_forward(P, P->ff + ii * G, P->ffC + ii * M);
  if (P->flag_doub_lay)
    for (int c = 0; c < 9; c++) fft3_forward(P, P->tt + c * G, P->ttC + c * M);
  const double alpha = prm->alpha;
  const double vol = prm->Lb[0] * prm->Lb[1] * prm->Lb[2];
  const double *iLb = prm->iLb;
#pragma omp parallel for collapse(2)
  for (int i = 0; i < Nxh; i++)
    for (int j = 0; j < Ny; j++)
      for (int k = 0; k < Nz; k++) {
        size_t idx = ((size_t)k * Ny + j) * Nxh + i;
        cplx vC[3] = {0., 0., 0.};
        if (!(i == 0 && j == 0 && k == 0)) {
          double q[3], qt[3];
          q[0] = i * iLb[0]; /* :287-303 */
          q[1] = (j < Ny / 2) ? j * iLb[1] : (j - Ny) * iLb[1];
          double q3 = (k < Nz / 2) ? k * iLb[2] : (k - Nz) * iLb[2];
          q[2] = -q3;
          for (int d = 0; d < 3; d++) qt[d] = sqrt(PI * alpha) * q[d];
          double q2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2];
          double q2t = PI * alpha * q2;
          double expq2t = exp(-q2t);
          double phi0 = expq2t / q2t;
          double phi1 = (expq2t + phi0) / q2t;
          if (P->flag_sing_lay) {
            cplx fC[3], dotp = 0.;
            for (int d = 0; d < 3; d++) {
              fC[d] = P->ffC[d * M + idx];
              dotp += qt[d] * fC[d];
            }
            for (int d = 0; d < 3; d++) vC[d] += 2 * alpha / vol * phi1 * (q2t * fC[d] - qt[d] * dotp);
          }
          if (P->flag_doub_lay) {
            cplx T[3][3]; /* T[ii][jj] = ttC(k,j,i,ii,jj) */
            for (int ii = 0; ii < 3; ii++)
              for (int jj = 0; jj < 3; jj++) T[ii][jj] = P->ttC[(size_t)(ii + 3 * jj) * M + idx];
            cplx tr = T[0][0] + T[1][1] + T[2][2];
            cplx qT[3], Tq[3], qTq = 0.;
            for (int d = 0; d < 3; d++) {
              qT[d] = q[0] * T[0][d] + q[1] * T[1][d] + q[2] * T[2][d]; /* matmul(q,T) */
              Tq[d] = T[d][0] * q[0] + T[d][1] * q[1] + T[d][2] * q[2]; /* matmul(T,q) */
            }
            for (int d = 0; d < 3; d++) qTq += q[d] * Tq[d];
            for (int d = 0; d < 3; d++) {
              cplx t = I * 4 * PI * alpha / vol * phi0 * (q[d] * tr + qT[d] + Tq[d]);
              t = t - I * 8 * PI * PI * alpha * alpha / vol * phi1 * qTq * q[d];
              vC[d] -= t; /* :202 "the sign is tricky" */
            }
          }
        }
        double b = P->bb[((size_t)i * Ny + j) * Nz + k];
        for (int d = 0; d < 3; d++) P->vvC[d * M + idx] = b * vC[d];
      }
  for (int ii = 0; ii < 3; ii++) fft3_backward(P, P->vvC + ii * M, P->vv + ii * G);
}

/* ModPME.F90:228-248 PME_Add_Interp_Vel + :450-489 Interp_Vel (single node: the halo of
 * Update_Buff_Vel makes the z index purely periodic) */
void orc_pme_add_interp_vel(orc_pme *P, const orc_targets *tl, double *v) {
  const orc_params *prm = &P->prm;
  int Nx = P->Nx, Ny = P->Ny, Nz = P->Nz, PB = prm->P;
  size_t G = (size_t)Nx * Ny * Nz, nt = tl->n;
  double ih[3];
  for (int d = 0; d < 3; d++) ih[d] = prm->Nb[d] / prm->Lb[d];
#pragma omp parallel for schedule(static, 256)
  for (size_t t = 0; t < nt; t++) {
    if (!tl->active[t]) continue;
    double wx[64], wy[64], wz[64], dv[3] = {0, 0, 0};
    int imin, jmin, kmin;
    orc_bspline_func(tl->x[2 * nt + t] * ih[2], PB, &kmin, wz);
    orc_bspline_func(tl->x[nt + t] * ih[1], PB, &jmin, wy);
    orc_bspline_func(tl->x[t] * ih[0], PB, &imin, wx);
    for (int k0 = 1; k0 <= PB; k0++) {
      int k = imodulo(kmin + k0 - 1, Nz);
      for (int j0 = 1; j0 <= PB; j0++)
        for (int i0 = 1; i0 <= PB; i0++) {
          int j = imodulo(jmin + j0 - 1, Ny), i = imodulo(imin + i0 - 1, Nx);
          double wxyz = wx[i0 - 1] * wy[j0 - 1] * wz[k0 - 1];
          size_t q = ((size_t)k * Ny + j) * Nx + i;
          for (int d = 0; d < 3; d++) dv[d] = dv[d] + wxyz * P->vv[d * G + q];
        }
    }
    for (int d = 0; d < 3; d++) v[d * nt + t] = v[d * nt + t] + dv[d] / tl->Acoef[t];
  }
}

void orc_fft_forward(const orc_params *prm, const double *real_in, double *cplx_out) {
  orc_pme P;
  memset(&P, 0, sizeof P);
  P.Nx = prm->Nb[0];
  P.Ny = prm->Nb[1];
  P.Nz = prm->Nb[2];
  P.Nxh = P.Nx / 2 + 1;
  fft_tab_init(&P.Tx, P.Nx);
  fft_tab_init(&P.Ty, P.Ny);
  fft_tab_init(&P.Tz, P.Nz);
  fft3_forward(&P, real_in, (cplx *)cplx_out);
  free(P.Tx.tw);
  free(P.Ty.tw);
  free(P.Tz.tw);
}

void orc_fft_backward(const orc_params *prm, const double *cplx_in, double *real_out) {
  orc_pme P;
  memset(&P, 0, sizeof P);
  P.Nx = prm->Nb[0];
  P.Ny = prm->Nb[1];
  P.Nz = prm->Nb[2];
  P.Nxh = P.Nx / 2 + 1;
  fft_tab_init(&P.Tx, P.Nx);
  fft_tab_init(&P.Ty, P.Ny);
  fft_tab_init(&P.Tz, P.Nz);
  fft3_backward(&P, (const cplx *)cplx_in, real_out);
  free(P.Tx.tw);
  free(P.Ty.tw);
  free(P.Tz.tw);
}

/* The cell operator as its callers compose it (ModVelSolver.F90:568-582 for c1=0,c2=-1/4pi;
 * :473-489 for c1=1/4pi,c2=0 without walls): v += AddIntOnRbcs + PME. */
void orc_apply_cells(const orc_params *prm, const orc_cells *C, orc_pme *P, double c1, double c2,
                     const orc_targets *tl, double *v, int flags) {
  orc_add_int_on_rbcs(prm, C, c1, c2, tl, v, flags);
  if (P) {
    size_t Np = (size_t)C->ncell * C->nlat * C->nlon;
    double *sf = NULL, *sg = NULL, *Bp = (double *)malloc(sizeof(double) * Np);
    if (fabs(c1) > 1.e-10) sf = (double *)malloc(sizeof(double) * 3 * Np);
    if (fabs(c2) > 1.e-10) sg = (double *)malloc(sizeof(double) * 3 * Np);
    int npc = C->nlat * C->nlon;
#pragma omp parallel for
    for (size_t q = 0; q < Np; q++) {
      double dS = C->detj[q] * C->w[q % C->nlat];
      Bp[q] = C->Bcoef[q / npc];
      for (int d = 0; d < 3; d++) {
        if (sf) sf[d * Np + q] = C->f[d * Np + q] * dS;
        if (sg) sg[d * Np + q] = C->g[d * Np + q] * dS;
      }
    }
    orc_pme_distrib_source(P, c1, c2, (int)Np, C->x, sf, sg, C->a3, Bp, 0);
    orc_pme_transform(P);
    orc_pme_add_interp_vel(P, tl, v);
    free(sf);
    free(sg);
    free(Bp);
  }
}
