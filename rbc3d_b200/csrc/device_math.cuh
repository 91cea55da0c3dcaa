// device_math.cuh -- FP64 device helpers shared by the kernels.  Each helper names the reference routine
// whose arithmetic it reproduces (paths relative to the reference's common/).
#pragma once
#include "rbc3d_internal.h"

namespace rbc3d {

#define RBC_PI 3.14159265358979323846
#define RBC_TWO_PI (2.0 * RBC_PI)
#define RBC_I_2PI (1.0 / RBC_TWO_PI)
#define FULL_MASK 0xffffffffu

__device__ __forceinline__ int imodulo(int a, int n) {
  int r = a % n;
  return r < 0 ? r + n : r;
}

// ModHashTable.F90:78-82: i = modulo(floor(x*iLbNc), Nc) (+1 in Fortran).  __dmul_rn forbids FMA contraction
// so the product rounds exactly like the reference build (gfortran, baseline x86-64, no FMA).
__device__ __forceinline__ int cell_coord(double x, double iLbNc, int Nc) {
  return imodulo((int)floor(__dmul_rn(x, iLbNc)), Nc);
}

// ModIntOnRbcs.F90:71-72: xx = xj - xi; xx = xx - nint(xx*iLb)*Lb, rounded step by step like the reference.
__device__ __forceinline__ double min_image(double d, double iLb, double Lb) {
  double t = __dmul_rn(d, iLb);
  if (fabs(t) >= 0.5) d = __dsub_rn(d, __dmul_rn(round(t), Lb));
  return d;
}

// sum(xx*xx) evaluated left to right without contraction (ModIntOnRbcs.F90:73)
__device__ __forceinline__ double norm2_exact(double x, double y, double z) {
  return __dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
  return v;
}

// 1/sqrt(x) for normal positive x: the hardware seed (MUFU.RSQ64H) and one third-order correction -- the same
// arithmetic as CUDA's rsqrt() without its special-case branch (callers guarantee r_eps^2 <= x <= rc^2).
__device__ __forceinline__ double rsqrt_pos(double x) {
  double y0;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
  const double h = fma(x, -(y0 * y0), 1.0);
  const double p = fma(h, 0.375, 0.5);
  return fma(p, y0 * h, y0);
}

// ModEwaldFunc.F90:120-121,173: linear interpolation in an 8193-entry table indexed by s = N*r/rc
__device__ __forceinline__ double table_lerp(const double *__restrict__ tab, double s, int i) {
  double t0 = tab[i], t1 = tab[i + 1];
  double fr = s - (double)i;
  return fma(fr, t1 - t0, t0);
}

// ModEwaldFunc.F90:141-178 EwaldCoeff_DL with the table in global memory
__device__ __forceinline__ double ewald_dl(const double *__restrict__ tab, const Params &prm, double r) {
  if (r < prm.r_eps) return 0.0;
  double s = (double)RBC3D_NTAB * r / prm.rc;
  int i = (int)floor(s);
  if (i >= RBC3D_NTAB) return 0.0;
  double c = __ldg(tab + i) * ((double)(i + 1) - s) + __ldg(tab + i + 1) * (s - (double)i);
  double r2 = r * r;
  return c / (r2 * r2 * r);
}

// ModEwaldFunc.F90:86-131 EwaldCoeff_SL, table interleaved (c1,c2)
__device__ __forceinline__ void ewald_sl(const double *__restrict__ tab, const Params &prm, double r, double &A,
                                         double &B) {
  A = 0.0;
  B = 0.0;
  if (r < prm.r_eps) return;
  double s = (double)RBC3D_NTAB * r / prm.rc;
  int i = (int)floor(s);
  if (i >= RBC3D_NTAB) return;
  double w0 = (double)(i + 1) - s, w1 = s - (double)i;
  double c1 = __ldg(tab + 2 * i) * w0 + __ldg(tab + 2 * i + 2) * w1;
  double c2 = __ldg(tab + 2 * i + 1) * w0 + __ldg(tab + 2 * i + 3) * w1;
  double ir = 1.0 / r, ir2 = ir * ir;
  A = c1 * ir * ir2 + c2 * ir2;
  B = c1 * ir - c2;
}

// ModBasicMath.F90:351-379 MaskFunc
__device__ __forceinline__ double mask_func(const double *__restrict__ tab, double x) {
  double s = fabs(x) * (double)RBC3D_NTAB;
  int i = (int)floor(s);
  if (i >= RBC3D_NTAB) return 0.0;
  return tab[i] * ((double)(i + 1) - s) + tab[i + 1] * (s - (double)i);
}

// ModPolarPatch.F90:249-257 DistOnSphere
__device__ __forceinline__ double dist_on_sphere(double th0, double phi0, double th1, double phi1) {
  double d = cos(th0 - th1) - sin(th0) * sin(th1) * (1.0 - cos(phi0 - phi1));
  d = fmin(1.0, fmax(-1.0, d));
  return acos(d);
}

// ModBasicMath.F90:392-417 BsplineFunc, P <= 16.  The recurrence is kept in the reference's order.
template <int PMAX>
__device__ __forceinline__ void bspline_func(double xc, int P, int &imin, double *w) {
  double u[PMAX];
  double fl = floor(xc);
  imin = (int)fl - (P - 1);
  u[0] = (double)imin - (xc - (double)P);
#pragma unroll
  for (int j = 1; j < PMAX; j++)
    if (j < P) u[j] = u[j - 1] + 1.0;
  w[0] = 1.0;
#pragma unroll
  for (int j = 1; j < PMAX; j++)
    if (j < P) w[j] = 0.0;
#pragma unroll
  for (int pp = 2; pp <= PMAX; pp++) {
    if (pp <= P) {
      double inv = 1.0 / ((double)pp - 1.0);
#pragma unroll
      for (int j = PMAX; j >= 2; j--)
        if (j <= pp) w[j - 1] = u[j - 1] * inv * w[j - 1] + ((double)pp - u[j - 1]) * inv * w[j - 2];
      w[0] = u[0] * inv * w[0];
    }
  }
}

// ModSpline.F90:150-191 Spline_Interp on one cell's ABI-layout spline [4][NVAR][n][m] (m fastest).
template <int NVAR>
__device__ __forceinline__ void spline_interp(const double *__restrict__ sp, int m, int n, double x, double y,
                                              double *f) {
  const double hx = RBC_TWO_PI / (double)m, hy = RBC_TWO_PI / (double)n;
  const double ihx = 1.0 / hx, ihy = 1.0 / hy;
  double xs = x * ihx, ys = y * ihy;
  int i1 = (int)floor(xs), j1 = (int)floor(ys);
  double s = xs - (double)i1, t = ys - (double)j1;
  i1 = imodulo(i1, m);
  j1 = imodulo(j1, n);
  int i2 = i1 + 1 == m ? 0 : i1 + 1, j2 = j1 + 1 == n ? 0 : j1 + 1;
  double cx[4] = {1.0 + s * s * (-3.0 + 2.0 * s), s * s * (3.0 - 2.0 * s), hx * s * (1.0 + s * (-2.0 + s)),
                  hx * s * s * (-1.0 + s)};
  double cy[4] = {1.0 + t * t * (-3.0 + 2.0 * t), t * t * (3.0 - 2.0 * t), hy * t * (1.0 + t * (-2.0 + t)),
                  hy * t * t * (-1.0 + t)};
  const size_t plane = (size_t)m * n, arr = plane * NVAR;
  const double *U = sp, *U1 = sp + arr, *U2 = sp + 2 * arr, *U12 = sp + 3 * arr;
  const size_t a11 = i1 + (size_t)m * j1, a12 = i1 + (size_t)m * j2, a21 = i2 + (size_t)m * j1,
               a22 = i2 + (size_t)m * j2;
#pragma unroll
  for (int l = 0; l < NVAR; l++) {
    const size_t o = plane * l;
    double r0 = __ldg(U + o + a11) * cy[0] + __ldg(U + o + a12) * cy[1] + __ldg(U2 + o + a11) * cy[2] +
                __ldg(U2 + o + a12) * cy[3];
    double r1 = __ldg(U + o + a21) * cy[0] + __ldg(U + o + a22) * cy[1] + __ldg(U2 + o + a21) * cy[2] +
                __ldg(U2 + o + a22) * cy[3];
    double r2 = __ldg(U1 + o + a11) * cy[0] + __ldg(U1 + o + a12) * cy[1] + __ldg(U12 + o + a11) * cy[2] +
                __ldg(U12 + o + a12) * cy[3];
    double r3 = __ldg(U1 + o + a21) * cy[0] + __ldg(U1 + o + a22) * cy[1] + __ldg(U12 + o + a21) * cy[2] +
                __ldg(U12 + o + a22) * cy[3];
    f[l] = cx[0] * r0 + cx[1] * r1 + cx[2] * r2 + cx[3] * r3;
  }
}

// ModPolarPatch.F90:99-148 PolarPatch_Build for one (thL, phiL) pair
__device__ __forceinline__ void polar_patch_point(double sin_th0, double cos_th0, double phi0, double thL,
                                                  double phiL, double &thG, double &phiG) {
  double st, ct, sp, cp;
  sincos(thL, &st, &ct);
  sincos(phiL, &sp, &cp);
  double x0 = st * cp, x1 = st * sp, x2 = ct;
  double y0 = cos_th0 * x0 + sin_th0 * x2;
  double y1 = x1;
  double y2 = -sin_th0 * x0 + cos_th0 * x2;
  y2 = fmax(-1.0, fmin(1.0, y2));
  thG = acos(y2);
  double ph = atan2(y1, y0) + phi0;
  phiG = ph - floor(ph * RBC_I_2PI) * RBC_TWO_PI;
}

// ModPolarPatch.F90:217-242 PolarPatch_Map
__device__ __forceinline__ void polar_patch_map(double th0, double phi0, double dth, double dphi, double &th,
                                                double &phi) {
  double st0, ct0, sp0, cp0, sd, cd, sdp, cdp;
  sincos(th0, &st0, &ct0);
  sincos(phi0, &sp0, &cp0);
  sincos(dth, &sd, &cd);
  sincos(dphi, &sdp, &cdp);
  double s1[3] = {ct0 * cp0, ct0 * sp0, -st0};
  double s2[3] = {-sp0, cp0, 0.0};
  double x0[3] = {st0 * cp0, st0 * sp0, ct0};
  double xl[3] = {sd * cdp, sd * sdp, cd};
  double x[3];
#pragma unroll
  for (int d = 0; d < 3; d++) x[d] = xl[0] * s1[d] + xl[1] * s2[d] + xl[2] * x0[d];
  double nrm = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
#pragma unroll
  for (int d = 0; d < 3; d++) x[d] = x[d] / nrm;
  x[2] = fmax(-1.0, fmin(1.0, x[2]));
  th = acos(x[2]);
  double p = atan2(x[1], x[0]);
  if (p < 0) p += RBC_TWO_PI;
  phi = p;
}

// ModQuadRule.F90:135-206 GauLeg: root j (1-based, j <= (n+1)/2) on [-1,1] and its weight
__device__ __forceinline__ void gauleg_root(int n, int j, double &z, double &wt) {
  z = cos(RBC_PI * ((double)j - 0.25) / ((double)n + 0.5));
  double pp = 1.0, z1;
  for (int its = 1; its <= 10; its++) {
    double p1 = 1.0, p2 = 0.0, p3;
    for (int k = 1; k <= n; k++) {
      p3 = p2;
      p2 = p1;
      p1 = ((2.0 * k - 1.0) * z * p2 - ((double)k - 1.0) * p3) / (double)k;
    }
    pp = (double)n * (z * p1 - p2) / (z * z - 1.0);
    z1 = z;
    z = z1 - p1 / pp;
    if (!(fabs(z - z1) > 3.e-14)) break;
  }
  wt = 2.0 / ((1.0 - z * z) * pp * pp);
}

}  // namespace rbc3d
